"""GPU: sliding-window raster inference (gdl_b200/inference.py) on the real kernels against the window-sum of the
oracle model (fp32, CPU restatement of the reference's SegFormer) — same tolerance rule as the model tests."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_sliding_window_segformer_b0(cuda):
    from gdl_b200.inference import SlidingWindowSegmenter, window_origins
    from gdl_b200.models.segformer import SegFormer
    from oracle import segformer as osf
    k, cin, t, stride = 5, 4, 128, 64
    torch.manual_seed(0)
    prod = SegFormer("mit_b0", in_channels=cin, num_classes=k, compute_dtype=torch.bfloat16).cuda().eval()
    with torch.no_grad():
        for n_, p in prod.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
        for m in prod.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.1)
                m.running_var.uniform_(0.5, 1.5)
    sd = {n_: v.detach().clone() for n_, v in prod.state_dict().items()}
    mean, std = [0.45] * cin, [0.22] * cin
    g = torch.Generator().manual_seed(4)
    h, w = 300, 200
    raster = torch.randint(0, 256, (h, w, cin), generator=g, dtype=torch.uint8)
    seg = SlidingWindowSegmenter(prod, tile=t, stride=stride, batch=5, mean=mean, std=std)
    got = seg.logits(raster.cuda())
    cls = seg.predict(raster.pin_memory())  # host raster: copied once
    x = ((raster.float().cuda() / 255.0) - torch.tensor(mean).cuda()) / torch.tensor(std).cuda()
    want = torch.zeros(h, w, k, device="cuda")
    want_ac = torch.zeros(h, w, k, device="cuda")
    nwin = 0
    with torch.no_grad():
        for y in window_origins(h, t, stride):
            for xx in window_origins(w, t, stride):
                tile = x[y:y + t, xx:xx + t].permute(2, 0, 1).unsqueeze(0)
                want[y:y + t, xx:xx + t] += osf.segformer_forward(sd, tile, "mit_b0", training=False)[0].permute(1, 2, 0)
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    o = osf.segformer_forward(sd, tile, "mit_b0", training=False)[0]
                want_ac[y:y + t, xx:xx + t] += o.float().permute(1, 2, 0)
                nwin += 1
    assert seg.windows_done == 2 * nwin  # logits() + predict()
    ep = ((got - want).norm() / want.norm()).item()
    ea = ((want_ac - want).norm() / want.norm()).item()
    print(f"sliding window logits rel err: product {ep:.4f}, autocast oracle {ea:.4f}")
    assert ep < max(2.5 * ea, 5e-3)
    agree = (cls.long().cuda() == want.argmax(2)).float().mean().item()
    agree_ac = (want_ac.argmax(2) == want.argmax(2)).float().mean().item()
    print(f"class agreement with the fp32 oracle: product {agree:.4f}, autocast oracle {agree_ac:.4f}")
    assert agree >= agree_ac - 0.01 and cls.dtype == torch.uint8 and cls.shape == (h, w)


def test_sliding_window_cuda_graph_replay_equals_eager(cuda):
    """cuda_graph=True (window-batch forward captured once, replayed per full batch; ragged last batch eager) returns
    the same logit sums as the eager driver."""
    from gdl_b200.inference import SlidingWindowSegmenter
    from gdl_b200.models.segformer import SegFormer
    torch.manual_seed(0)
    prod = SegFormer("mit_b0", in_channels=3, num_classes=4, compute_dtype=torch.bfloat16).cuda().eval()
    raster = torch.randint(0, 256, (448, 320, 3), generator=torch.Generator().manual_seed(2), dtype=torch.uint8).cuda()
    kw = dict(tile=128, stride=64, batch=4, mean=[0.45] * 3, std=[0.22] * 3)
    eager = SlidingWindowSegmenter(prod, **kw)
    graph = SlidingWindowSegmenter(prod, cuda_graph=True, **kw)
    want = eager.logits(raster)
    got = graph.logits(raster)          # 6 x 4 = 24 windows: 1 eager batch, 1 capture, 4 replays
    assert graph._graph is not None and graph.windows_done == eager.windows_done == 24
    assert torch.allclose(got, want, atol=1e-5, rtol=1e-5)
    again = graph.logits(raster)        # all replays
    assert torch.allclose(again, want, atol=1e-5, rtol=1e-5)
