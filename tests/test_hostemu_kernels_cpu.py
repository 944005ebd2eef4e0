"""The scalar CUDA kernels of libgdlb200, EXECUTED on the CPU: the product's .cu sources compiled for the host
(tests/hostemu: threads of a block as fibers, __syncthreads / warp shuffles as barriers) and driven by the GPU test files
themselves through the same ctypes binding and C ABI.

Two purposes.  (1) The kernel tests that are green on a B200 (test_kernels_gpu.py, test_transformer_kernels_gpu.py) must be
green here too — that validates the emulation.  (2) The kernels written after the round's GPU budget was spent (augmentation,
confusion counts, Dropout2d, ViT training glue, channel pooling: tests/test_zz*_gpu.py) get their CUDA source executed —
indexing, launch geometry, reductions, argument checks — before they meet a B200.  The tensor-core kernels (tcgen05 / TMA) are
NOT covered: where a test needs a convolution it comes from the torch emulation.  This is test infrastructure; the product
path never loads this library.
"""
import pytest

import hostemu

KERNEL_FILES = {
    # B200-verified kernel tests (everything except the tensor-core convolution tests)
    "test_kernels_gpu": dict(exclude=("test_conv_fwd_matches_fp32_conv", "test_conv_reads_channel_slices_and_writes_strided",
                                      "test_conv_wgrad_matches_autograd", "test_conv_dgrad_through_transposed_weights")),
    "test_transformer_kernels_gpu": {},
    "test_upernet_gpu": dict(include=("test_adaptive_pool_and_add_kernels",)),
    # written after the GPU budget was spent
    "test_zz2_augment_metrics_gpu": dict(include=("test_augment_normalize_matches_oracle", "test_augment_full_tile_batch_and_errors",
                                                  "test_argmax_confusion_bit_exact", "test_mean_iou_metric_on_device")),
    "test_zz4_dofa_trainable_gpu": dict(include=("test_vit_training_kernels",)),
    "test_zz5_stochastic_layers_gpu": dict(include=("test_dropout2d_kernel",)),
    "test_zz6_dynamic_encoder_gpu": dict(include=("test_channel_pool_kernels",)),
    "test_upsample_head_gpu": dict(exclude=("test_segformer_fused_head_step_equals_unfused_step", "test_fused_ce_at_baseline_shapes")),  # fused bilinear upsample + loss / argmax head (round 2)
    # ordered reductions (round 2): slot sums + tickets executed on the host, block by block
    "test_determinism_gpu": dict(exclude=("test_wgrad_reproducible", "test_wgrad_reproducible_small", "test_batched_wgrad_reproducible")),
}

CASES = [(f, fn, kw, ident) for f, sel in KERNEL_FILES.items() for fn, kw, ident in hostemu.cases(f, **sel)]


@pytest.mark.parametrize("file,fname,kw", [pytest.param(f, fn, kw, id=ident) for f, fn, kw, ident in CASES])
def test_cuda_source_on_host(monkeypatch, tmp_path, file, fname, kw):
    hostemu.install(monkeypatch)
    hostemu.run_case(file, fname, kw, tmp_path)
