"""Pins the CPU oracle (oracle/msda_oracle.c and the torch port of the reference CPU path) to golden vectors that
were produced by the reference's own ms_deform_attn_core_pytorch (tests/golden/make_golden.py).

Mirrors the three checks of the reference's ops/test.py (fwd fp64 allclose default tolerances :40, fwd fp32
rtol 1e-2 / atol 1e-3 :56, gradients :63-86) as asserting tests.
"""
import numpy as np
import pytest
import torch

from oracle import msda_oracle


def _args(c, dtype):
    return (c["value"].astype(dtype), c["spatial_shapes"], c["level_start_index"],
            c["sampling_locations"].astype(dtype), c["attention_weights"].astype(dtype))


def test_forward_fp64_matches_reference(golden):
    out = msda_oracle.forward(*_args(golden, np.float64))
    # reference check_forward_equal_with_pytorch_double: torch.allclose defaults (rtol 1e-5, atol 1e-8)
    np.testing.assert_allclose(out, golden["out"], rtol=1e-9, atol=1e-12)


def test_forward_fp32_matches_reference(golden):
    out = msda_oracle.forward(*_args(golden, np.float32))
    assert out.dtype == np.float32
    # much tighter than the reference's own fp32 bar (rtol 1e-2, atol 1e-3; ops/test.py:56)
    np.testing.assert_allclose(out, golden["out"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(out, golden["out_fp32"], rtol=1e-4, atol=1e-5)


def test_backward_fp64_matches_reference_autograd(golden):
    gv, gl, ga = msda_oracle.backward(golden["grad_output"], *_args(golden, np.float64))
    np.testing.assert_allclose(gv, golden["grad_value"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(ga, golden["grad_attention_weights"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(gl, golden["grad_sampling_locations"], rtol=1e-8, atol=1e-10)


def test_backward_fp32_close_to_fp64_truth(golden):
    gv, gl, ga = msda_oracle.backward(golden["grad_output"].astype(np.float32), *_args(golden, np.float32))
    scale = lambda a: max(1.0, float(np.abs(a).max()))
    assert np.abs(gv - golden["grad_value"]).max() <= 1e-4 * scale(golden["grad_value"])
    assert np.abs(ga - golden["grad_attention_weights"]).max() <= 1e-4 * scale(golden["grad_attention_weights"])
    assert np.abs(gl - golden["grad_sampling_locations"]).max() <= 2e-4 * scale(golden["grad_sampling_locations"])


def test_torch_port_of_cpu_path_matches_reference(golden):
    """The --impl reference arm of bench.py times this port; it must be the reference's function."""
    t = lambda a: torch.from_numpy(np.asarray(a))
    v = t(golden["value"]).requires_grad_(True)
    lo = t(golden["sampling_locations"]).requires_grad_(True)
    at = t(golden["attention_weights"]).requires_grad_(True)
    out = msda_oracle.core_pytorch_port(v, t(golden["spatial_shapes"]), lo, at)
    np.testing.assert_allclose(out.detach().numpy(), golden["out"], rtol=1e-10, atol=1e-13)
    out.backward(t(golden["grad_output"]))
    np.testing.assert_allclose(v.grad.numpy(), golden["grad_value"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(lo.grad.numpy(), golden["grad_sampling_locations"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(at.grad.numpy(), golden["grad_attention_weights"], rtol=1e-9, atol=1e-12)


def test_skipped_sample_window_is_exact():
    """cuh:288 -- a tap is dropped unless h_im > -1, w_im > -1, h_im < H, w_im < W (strict)."""
    shapes = np.array([[4, 4]], dtype=np.int64)
    lsi = np.array([0], dtype=np.int64)
    value = np.ones((1, 16, 1, 2), dtype=np.float64)
    attn = np.ones((1, 1, 1, 1, 1), dtype=np.float64)
    def run(x, y):
        loc = np.array([x, y], dtype=np.float64).reshape(1, 1, 1, 1, 1, 2)
        return msda_oracle.forward(value, shapes, lsi, loc, attn)[0, 0, 0]
    assert run(0.5, 0.5) == 1.0                     # interior: weights sum to 1
    assert run(0.0, 0.5) == 0.5                     # w_im = -0.5: left corners outside -> half weight
    assert run(-0.125, 0.5) == 0.0                  # w_im = -1.0: not > -1 -> dropped
    assert run(1.125, 0.5) == 0.0                   # w_im = 4.0: not < W -> dropped
    assert abs(run(1.124, 0.5) - 0.004) < 1e-12     # just inside: only the left corner (w=3), weight 1-lw
