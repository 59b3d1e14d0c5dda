"""Generate the golden vectors that pin the oracle (and through it the CUDA kernels) to the reference.

Runs ONLY in the build container (needs /root/reference). It imports the reference's own CPU path,
``ms_deform_attn_core_pytorch`` (ops/functions/ms_deform_attn_func.py:43-63), unmodified and from where it lies;
the unbuilt native extension that file imports at module level (``:18``) is satisfied by an empty stub module.
Gradients are torch.autograd through that function in fp64, for a fixed, stored ``grad_output``.

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz

The cases mirror the reference's only test of this path (ops/test.py:21-28: N=1, M=2, D=2, Lq=2, L=2, P=2,
shapes (6,4),(3,2), seed 3, value=rand*0.01, loc=rand, attn=rand+1e-5 normalised over (L,P)) and add the
production geometry (M=8, D=32, L=4, P=4), the D values of the reference gradcheck list that are cheap
(30, 71), ragged / degenerate level shapes and out-of-range sampling locations.
"""
import os
import sys
import types

import numpy as np
import torch

REF_OPS_PARENT = "/root/reference/projects/UNINEXT/uninext/models/deformable_detr"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    sys.modules.setdefault("MultiScaleDeformableAttention", types.ModuleType("MultiScaleDeformableAttention"))
    sys.path.insert(0, REF_OPS_PARENT)
    from ops.functions.ms_deform_attn_func import ms_deform_attn_core_pytorch  # noqa: E402
    return ms_deform_attn_core_pytorch


def make_case(core, name, shapes, N, M, D, Lq, P, seed, loc_mode="unit", value_scale=0.01):
    g = torch.Generator().manual_seed(seed)
    shapes_t = torch.as_tensor(shapes, dtype=torch.long)
    L = len(shapes)
    lsi = torch.cat((shapes_t.new_zeros((1,)), shapes_t.prod(1).cumsum(0)[:-1]))
    S = int(shapes_t.prod(1).sum())
    value = torch.rand(N, S, M, D, generator=g, dtype=torch.float64) * value_scale
    loc = torch.rand(N, Lq, M, L, P, 2, generator=g, dtype=torch.float64)
    if loc_mode == "wide":            # exercise the validity window and the per-corner predicates
        loc = loc * 2.0 - 0.5
    elif loc_mode == "edges":         # hug the borders: most taps have 1-2 corners outside
        loc = torch.where(torch.rand(loc.shape, generator=g) < 0.5, loc * 0.06 - 0.02, 1.0 - (loc * 0.06 - 0.02))
    attn = torch.rand(N, Lq, M, L, P, generator=g, dtype=torch.float64) + 1e-5
    attn = attn / attn.sum(-1, keepdim=True).sum(-2, keepdim=True)
    grad_out = torch.randn(N, Lq, M * D, generator=g, dtype=torch.float64)

    v = value.clone().requires_grad_(True)
    lo = loc.clone().requires_grad_(True)
    at = attn.clone().requires_grad_(True)
    out = core(v, shapes_t, lo, at)
    out.backward(grad_out)
    with torch.no_grad():
        out32 = core(value.float(), shapes_t, loc.float(), attn.float())
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(
        path,
        spatial_shapes=shapes_t.numpy(), level_start_index=lsi.numpy(),
        value=value.numpy(), sampling_locations=loc.numpy(), attention_weights=attn.numpy(),
        grad_output=grad_out.numpy(),
        out=out.detach().numpy(), out_fp32=out32.numpy(),
        grad_value=v.grad.numpy(), grad_sampling_locations=lo.grad.numpy(), grad_attention_weights=at.grad.numpy(),
    )
    print(f"{name}: S={S} out{tuple(out.shape)} -> {os.path.getsize(path) / 1024:.0f} KiB")


def main():
    core = load_reference()
    torch.manual_seed(3)
    # the reference's own test geometry (ops/test.py:21-25)
    make_case(core, "ref_test_geometry", [(6, 4), (3, 2)], N=1, M=2, D=2, Lq=2, P=2, seed=3)
    # production head geometry on a small pyramid, batch 2
    make_case(core, "prod_small", [(12, 20), (6, 10), (3, 5), (2, 3)], N=2, M=8, D=32, Lq=37, P=4, seed=11,
              value_scale=1.0)
    make_case(core, "prod_small_wide", [(12, 20), (6, 10), (3, 5), (2, 3)], N=2, M=8, D=32, Lq=21, P=4, seed=12,
              loc_mode="wide", value_scale=1.0)
    make_case(core, "prod_small_edges", [(9, 7), (5, 4), (3, 2), (1, 1)], N=1, M=8, D=32, Lq=33, P=4, seed=13,
              loc_mode="edges", value_scale=1.0)
    # channel counts from the reference gradcheck list (ops/test.py:85) that are cheap to store
    make_case(core, "d30", [(6, 4), (3, 2)], N=1, M=2, D=30, Lq=2, P=2, seed=14)
    make_case(core, "d71", [(6, 4), (3, 2)], N=1, M=2, D=71, Lq=2, P=2, seed=15)
    make_case(core, "d64", [(7, 5), (4, 3), (2, 2)], N=2, M=4, D=64, Lq=9, P=3, seed=16, loc_mode="wide")
    make_case(core, "d16_l1", [(8, 8)], N=3, M=3, D=16, Lq=5, P=1, seed=17, loc_mode="wide")
    # degenerate maps: single row / single column / single pixel levels
    make_case(core, "degenerate", [(1, 9), (7, 1), (1, 1)], N=1, M=2, D=8, Lq=11, P=2, seed=18, loc_mode="wide")


if __name__ == "__main__":
    main()
