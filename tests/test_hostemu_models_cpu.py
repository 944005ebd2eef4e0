"""Whole-model GPU tests on the CPU with the scalar CUDA kernels EXECUTED (tests/hostemu) and only the tensor-core entry points
(conv forward / wgrad, weight packing) emulated in torch: every normalise / BN / LayerNorm / softmax / depthwise-conv / bilinear /
loss / Adam / augmentation / dropout / channel-pool launch of a training step runs the product's CUDA source through the C ABI on
the operands the host engine really hands it (strides, channel slices, padded leading dimensions, flat-buffer views).

The cheap cases run in the default CPU suite; GDL_HOSTEMU_FULL=1 adds the SegFormer / DOFA train-step parity tests (1-4 minutes
each: one fiber per CUDA thread).  tools/hostemu_asan.sh runs the same files under AddressSanitizer.  Test infrastructure only.
"""
import os

import pytest

import hostemu

FULL = os.environ.get("GDL_HOSTEMU_FULL", "0") == "1"

DEFAULT = {
    "test_zz2_augment_metrics_gpu": ("test_trainer_step_with_augmentation_equals_step_on_augmented_batch",),
    "test_zz3_wds_feeder_gpu": ("test_feeder_to_device_matches_reference_golden", "test_trainer_steps_from_the_feeder"),
    "test_zz6_dynamic_encoder_gpu": ("test_dynamic_segformer_matches_reference_golden",),
    "test_unetpp_gpu": ("test_packed_weight_cache_is_refreshed_in_place",),  # gdl_repack_weights + derived operands
}
SLOW = {
    "test_upsample_head_gpu": ("test_segformer_fused_head_step_equals_unfused_step",),  # fused head inside the trainer step
    "test_unetpp_gpu": None, "test_segformer_gpu": None, "test_upernet_gpu": None, "test_dofa_gpu": None,  # green on a B200 (run 15)
    "test_zz1_inference_gpu": ("test_sliding_window_segformer_b0",),
    "test_zz4_dofa_trainable_gpu": ("test_dofa_unfrozen_train_step_parity", "test_dofa_unfrozen_fused_trainer_reduces_loss"),
    "test_zz5_stochastic_layers_gpu": ("test_segformer_train_step_with_supplied_draws",),
    "test_zz6_dynamic_encoder_gpu": ("test_dynamic_segformer_train_step_parity",),
    "test_zz7_sra_attention_gpu": ("test_segformer_with_fused_attention_equals_three_kernel_model",
                                   "test_dofa_encoder_with_flash_attention_equals_three_kernel_encoder"),
}


# CUDA-graph capture / stream behaviour has no CPU counterpart
EXCLUDE = ("test_sliding_window_cuda_graph_replay_equals_eager", "test_cuda_graph_step_equals_eager_step",
           "test_cuda_graph_step_equals_eager_step_bitwise", "test_set_lr_reaches_a_captured_graph")


def _params(table, slow):
    out = []
    for f, names in table.items():
        for fn, kw, ident in hostemu.cases(f, include=names, exclude=EXCLUDE):
            marks = [pytest.mark.skipif(not FULL, reason="set GDL_HOSTEMU_FULL=1 (minutes per case)")] if slow else []
            out.append(pytest.param(f, fn, kw, id=ident, marks=marks))
    return out


@pytest.mark.parametrize("file,fname,kw", _params(DEFAULT, False) + _params(SLOW, True))
def test_model_step_with_cuda_source_on_host(monkeypatch, tmp_path, file, fname, kw):
    # GDL_HOSTEMU_TC=1: also run the tensor-core kernels on the functional model instead of the torch convolution stand-in
    hostemu.install(monkeypatch, torch_convs=os.environ.get("GDL_HOSTEMU_TC", "0") != "1")
    hostemu.run_case(file, fname, kw, tmp_path)
