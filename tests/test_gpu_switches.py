"""The scheduling switches of the tiled kernels change nothing but the order of work: the pixel-patch slot order
(MSDA_PATCHES=1), plain-LDG tap loading (MSDA_NO_TMA=1) and the small-launch tap split over a warp's groups
(MSDA_SPLIT=0/1, normally chosen by launch size) must all give the same results.  Each setting is read once per
process, so every case runs in its own interpreter."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys, torch
sys.path.insert(0, %r)
from uninext_b200.dropin import MultiScaleDeformableAttention as MSDA
from uninext_b200.workloads import CONFIGS, make_inputs
res = {}
for kind in ("enc", "dec"):
    for dt in (torch.float32, torch.bfloat16):
        inp = make_inputs(CONFIGS["cfg1"], kind, "cuda", dtype=dt, seed=3, wild_fraction=0.1)
        a = (inp["value"], inp["spatial_shapes"], inp["level_start_index"], inp["sampling_locations"], inp["attention_weights"])
        out = MSDA.ms_deform_attn_forward(*a, 64)
        gv, gl, ga = MSDA.ms_deform_attn_backward(*a, inp["grad_output"], 64)
        res[(kind, str(dt))] = [t.float().cpu() for t in (out, gv, gl, ga)]
# ragged: odd head count, P=3, pairs not a multiple of the warp's group count, out-of-range taps
g = torch.Generator().manual_seed(5)
ss = torch.tensor([[5, 6], [3, 3]]); lsi = torch.tensor([0, 30])
for dt in (torch.float32, torch.bfloat16):
    v = torch.randn(3, 39, 5, 32, generator=g).to("cuda", dt)
    loc = (torch.rand(3, 13, 5, 2, 3, 2, generator=g) * 1.6 - 0.3).cuda()
    at = torch.rand(3, 13, 5, 2, 3, generator=g).cuda()
    go = torch.randn(3, 13, 160, generator=g).to("cuda", dt)
    a = (v, ss.cuda(), lsi.cuda(), loc, at)
    out = MSDA.ms_deform_attn_forward(*a, 64)
    gv, gl, ga = MSDA.ms_deform_attn_backward(*a, go, 64)
    res[("ragged", str(dt))] = [t.float().cpu() for t in (out, gv, gl, ga)]
torch.save(res, sys.argv[1])
"""


def _run(env, path):
    e = dict(os.environ, **env)
    subprocess.run([sys.executable, "-c", SCRIPT % ROOT, path], check=True, env=e, timeout=300)
    return torch.load(path)


@pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a GPU")
def test_patch_order_and_ldg_taps_match_default(tmp_path):
    base = _run({"MSDA_SPLIT": "0"}, str(tmp_path / "base.pt"))            # linear order, TMA-staged taps, one pair per group
    cases = (("patches", {"MSDA_PATCHES": "1", "MSDA_SPLIT": "0"}), ("no_tma", {"MSDA_NO_TMA": "1", "MSDA_SPLIT": "0"}),
             ("split", {"MSDA_SPLIT": "1"}), ("auto", {}))
    for name, env in cases:
        other = _run(env, str(tmp_path / f"{name}.pt"))
        exact = name in ("patches", "no_tma")          # same per-pair arithmetic; split/auto re-associate the tap sum
        for key, tensors in base.items():
            bf16 = "bfloat16" in key[1]
            for i, (t0, t1) in enumerate(zip(tensors, other[key])):
                scale = t0.abs().max().item()
                if exact and i != 1:
                    assert torch.equal(t0, t1), (name, key, i)
                else:                                  # i == 1: grad_value (atomic order); split: fp32 re-association
                    tol = (1e-2 if bf16 and i in (0, 1) else 2e-5) * scale
                    assert (t0 - t1).abs().max().item() <= tol, (name, key, i)


@pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a GPU")
def test_zero_fill_kernel_and_pdl_give_the_memset_results():
    """MSDA_KNOB_ZERO_FILL: grad_value zero-filled by cudaMemsetAsync (0), by msda_zero_fill (1), or by msda_zero_fill as
    the programmatic-dependent-launch primary of the tiled backward kernel (2).  Same results: grad_loc / grad_attn
    bit-identical, grad_value up to the order of its atomics.  The buffers are pre-filled with garbage and the call is
    repeated back to back, so a backward kernel that did not wait for the fill would show up as lost or stale sums."""
    sys.path.insert(0, ROOT)
    from uninext_b200 import _cabi
    from uninext_b200.dropin import MultiScaleDeformableAttention as MSDA
    from uninext_b200.workloads import CONFIGS, make_inputs
    lib = _cabi.load()
    was = lib.msda_set_knob(_cabi.KNOB_ZERO_FILL, 0)
    try:
        cases = [("cfg1", "enc", torch.float32), ("cfg1", "dec", torch.float32), ("cfg1", "enc", torch.bfloat16),
                 ("cfg1", "dec", torch.float64), ("cfg2", "enc", torch.float32), ("cfg2", "dec", torch.float32)]
        for cfgname, kind, dt in cases:
            inp = make_inputs(CONFIGS[cfgname], kind, "cuda", dtype=dt, seed=11, wild_fraction=0.05)
            a = (inp["value"], inp["spatial_shapes"], inp["level_start_index"], inp["sampling_locations"],
                 inp["attention_weights"])
            lib.msda_set_knob(_cabi.KNOB_ZERO_FILL, 0)
            ref = [t.double() for t in MSDA.ms_deform_attn_backward(*a, inp["grad_output"], 64)]
            scale = ref[0].abs().max().item()
            for mode in (1, 2):
                lib.msda_set_knob(_cabi.KNOB_ZERO_FILL, mode)
                for rep in range(4):
                    junk = torch.full((inp["value"].numel() + 64,), 7.0, device="cuda", dtype=torch.float32)   # dirty the allocator's blocks
                    del junk
                    gv, gl, ga = MSDA.ms_deform_attn_backward(*a, inp["grad_output"], 64)
                    tol = (1e-2 if dt == torch.bfloat16 else 2e-5) * scale
                    assert (gv.double() - ref[0]).abs().max().item() <= tol, (cfgname, kind, dt, mode, rep)
                    assert torch.equal(gl.double(), ref[1]) and torch.equal(ga.double(), ref[2]), (cfgname, kind, dt, mode)
        # inside a CUDA-graph capture the PDL pairing is dropped (plain launches are captured) and replays stay correct
        lib.msda_set_knob(_cabi.KNOB_ZERO_FILL, 2)
        inp = make_inputs(CONFIGS["cfg1"], "enc", "cuda", seed=12)
        a = (inp["value"], inp["spatial_shapes"], inp["level_start_index"], inp["sampling_locations"], inp["attention_weights"])
        want = MSDA.ms_deform_attn_backward(*a, inp["grad_output"], 64)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            got = MSDA.ms_deform_attn_backward(*a, inp["grad_output"], 64)
        for _ in range(3):
            got[0].fill_(3.0)
            g.replay()
        torch.cuda.synchronize()
        assert (got[0] - want[0]).abs().max().item() <= 2e-5 * want[0].abs().max().item()
        assert torch.equal(got[1], want[1]) and torch.equal(got[2], want[2])
    finally:
        lib.msda_set_knob(_cabi.KNOB_ZERO_FILL, was)
