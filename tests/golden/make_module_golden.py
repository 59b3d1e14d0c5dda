"""Golden vectors for the nn.Module around the op.  Build container only (needs /root/reference).

Runs the reference's own ``MSDeformAttn`` module (ops/modules/ms_deform_attn.py) on CPU in fp64, with the native call
``MSDeformAttnFunction.apply`` (which needs the unbuilt CUDA extension) replaced by the reference's own pure-PyTorch
``ms_deform_attn_core_pytorch`` -- both are reference code, nothing of ours is on this path.  Stores the module's
state_dict, inputs, output and the gradients of a fixed ``grad_output`` w.r.t. the query, the features and every
parameter.

    python tests/golden/make_module_golden.py       # rewrites tests/golden/module_*.npz
"""
import os
import sys
import types

import numpy as np
import torch

REF_OPS_PARENT = "/root/reference/projects/UNINEXT/uninext/models/deformable_detr"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    sys.modules.setdefault("MultiScaleDeformableAttention", types.ModuleType("MultiScaleDeformableAttention"))
    sys.path.insert(0, REF_OPS_PARENT)
    import ops.modules.ms_deform_attn as ref_mod
    from ops.functions.ms_deform_attn_func import ms_deform_attn_core_pytorch

    class _CoreAsFunction:            # same call shape as MSDeformAttnFunction.apply (func.py:24)
        @staticmethod
        def apply(value, shapes, lsi, loc, attn, im2col_step):
            return ms_deform_attn_core_pytorch(value, shapes, loc, attn)

    ref_mod.MSDeformAttnFunction = _CoreAsFunction

    shapes = torch.as_tensor([(12, 20), (6, 10), (3, 5), (2, 3)], dtype=torch.long)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    S = int(shapes.prod(1).sum())
    for name, lq, ref_dim, seed in (("module_enc2d", S, 2, 31), ("module_dec4d", 9, 4, 32)):
        torch.manual_seed(seed)
        mod = ref_mod.MSDeformAttn(64, 4, 2, 4).double()          # D = 32: the tiled-kernel head size
        with torch.no_grad():                      # leave the degenerate all-zero init so every path carries signal
            mod.sampling_offsets.weight.normal_(0, 0.02)
            mod.attention_weights.weight.normal_(0, 0.05)
        N = 2
        query = torch.randn(N, lq, 64, dtype=torch.float64, requires_grad=True)
        feats = torch.randn(N, S, 64, dtype=torch.float64, requires_grad=True)
        if ref_dim == 2:
            ref = torch.rand(N, lq, 4, 2, dtype=torch.float64)
        else:
            ref = torch.cat((torch.rand(N, lq, 4, 2, dtype=torch.float64),
                             0.05 + 0.3 * torch.rand(N, lq, 4, 2, dtype=torch.float64)), -1)
        mask = torch.zeros(N, S, dtype=torch.bool)
        mask[1, -7:] = True
        out = mod(query, ref, feats, shapes, lsi, mask)
        gout = torch.randn_like(out)
        out.backward(gout)
        blob = dict(spatial_shapes=shapes.numpy(), level_start_index=lsi.numpy(), query=query.detach().numpy(),
                    input_flatten=feats.detach().numpy(), reference_points=ref.numpy(), padding_mask=mask.numpy(),
                    out=out.detach().numpy(), grad_output=gout.numpy(), grad_query=query.grad.numpy(),
                    grad_input_flatten=feats.grad.numpy())
        for k, v in mod.state_dict().items():
            blob["param." + k] = v.numpy()
        for k, p in mod.named_parameters():
            blob["grad." + k] = p.grad.numpy()
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **{k: (v.astype(np.float32) if v.dtype == np.float64 and k.startswith(("param.", "grad."))
                                         and v.size > 4096 else v) for k, v in blob.items()})
        print(name, f"{os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
