"""The two scheduling switches of the tiled kernels change nothing but the order of work: the pixel-patch slot order
(MSDA_PATCHES=1) and plain-LDG tap loading (MSDA_NO_TMA=1) must give the same results as the default (linear order,
TMA-staged taps).  Each setting is read once per process, so every case runs in its own interpreter."""
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
torch.save(res, sys.argv[1])
"""


def _run(env, path):
    e = dict(os.environ, **env)
    subprocess.run([sys.executable, "-c", SCRIPT % ROOT, path], check=True, env=e, timeout=300)
    return torch.load(path)


@pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a GPU")
def test_patch_order_and_ldg_taps_match_default(tmp_path):
    base = _run({}, str(tmp_path / "base.pt"))
    for name, env in (("patches", {"MSDA_PATCHES": "1"}), ("no_tma", {"MSDA_NO_TMA": "1"})):
        other = _run(env, str(tmp_path / f"{name}.pt"))
        for key, tensors in base.items():
            out0, gv0, gl0, ga0 = tensors
            out1, gv1, gl1, ga1 = other[key]
            assert torch.equal(out0, out1), (name, key)                       # forward: same arithmetic per pair
            assert torch.equal(gl0, gl1) and torch.equal(ga0, ga1), (name, key)
            scale = gv0.abs().max().item()
            assert (gv0 - gv1).abs().max().item() <= 2e-5 * scale + 1e-2 * scale * ("bfloat16" in key[1]), (name, key)
