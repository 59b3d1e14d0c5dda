"""GPU tests of the callers of the op: the MSDeformAttn module against golden vectors produced by the REFERENCE module
(tests/golden/make_module_golden.py), the encoder/decoder layers, and the 6+6 layer stack used for frames/s."""
import numpy as np
import pytest
import torch

from tests.conftest import load_golden

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from uninext_b200.modules import MSDeformAttn
    from uninext_b200.modules.deformable_layers import (DeformableStack, DeformableTransformerDecoderLayer,
                                                        DeformableTransformerEncoderLayer)
    from uninext_b200.workloads import CONFIGS, level_tensors

DEV = "cuda"


def _rel(got, want):
    want = torch.as_tensor(want, dtype=torch.float64)
    return ((got.detach().double().cpu() - want).abs().max() / want.abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("name", ["module_enc2d", "module_dec4d"])
@pytest.mark.parametrize("op_dtype,tol", [(None, 2e-4), (torch.bfloat16, 2e-2)])
def test_module_matches_reference_module(name, op_dtype, tol):
    c = load_golden(name)
    mod = MSDeformAttn(64, 4, 2, 4, op_dtype=op_dtype).to(DEV)
    sd = {k[len("param."):]: torch.from_numpy(np.asarray(v)).float() for k, v in c.items() if k.startswith("param.")}
    mod.load_state_dict(sd, strict=True)                      # the reference's keys load unchanged
    t = lambda k, dt=torch.float32: torch.from_numpy(c[k]).to(DEV, dt)
    q = t("query").requires_grad_(True)
    x = t("input_flatten").requires_grad_(True)
    out = mod(q, t("reference_points"), x, t("spatial_shapes", torch.int64), t("level_start_index", torch.int64),
              t("padding_mask", torch.bool))
    out.backward(t("grad_output"))
    assert _rel(out, c["out"]) < tol
    assert _rel(q.grad, c["grad_query"]) < 5 * tol            # passes through softmax + the location arithmetic
    assert _rel(x.grad, c["grad_input_flatten"]) < tol
    for k, p in mod.named_parameters():
        assert _rel(p.grad, c["grad." + k]) < 5 * tol, k


def test_module_under_autocast_is_fp32_like_reference():
    c = load_golden("module_dec4d")
    mod = MSDeformAttn(64, 4, 2, 4).to(DEV)
    t = lambda k, dt=torch.float32: torch.from_numpy(c[k]).to(DEV, dt)
    args = (t("query"), t("reference_points"), t("input_flatten"), t("spatial_shapes", torch.int64),
            t("level_start_index", torch.int64))
    want = mod(*args)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        got = mod(*args)
    assert got.dtype == torch.float32 and torch.allclose(got, want, rtol=1e-5, atol=1e-6)


def test_layers_and_stack_run_forward_backward():
    cfg = CONFIGS["cfg1"]
    shapes = cfg.shapes
    ss, lsi = level_tensors(shapes, DEV)
    torch.manual_seed(0)
    n, s = 2, cfg.S
    src = torch.randn(n, s, 256, device=DEV, requires_grad=True)
    pos = torch.randn(n, s, 256, device=DEV)
    stack = DeformableStack(num_layers=2, num_queries=50).to(DEV)
    out = stack(src, pos, shapes, ss, lsi)
    assert out.shape == (n, 50, 256) and torch.isfinite(out).all()
    out.square().mean().backward()
    assert src.grad is not None and torch.isfinite(src.grad).all()
    missing = [k for k, p in stack.named_parameters() if p.grad is None]
    assert not missing, missing
    enc = DeformableTransformerEncoderLayer(d_ffn=512).to(DEV)
    dec = DeformableTransformerDecoderLayer(d_ffn=512).to(DEV)
    assert set(k.split(".")[0] for k in enc.state_dict()) == {"self_attn", "norm1", "linear1", "linear2", "norm2"}
    assert set(k.split(".")[0] for k in dec.state_dict()) == {"cross_attn", "norm1", "self_attn", "norm2", "linear1",
                                                              "linear2", "norm3"}
    # bf16 op inside the stack stays close to the fp32 op
    torch.manual_seed(1)
    a = DeformableStack(num_layers=1, num_queries=20).to(DEV)
    b = DeformableStack(num_layers=1, num_queries=20, op_dtype=torch.bfloat16).to(DEV)
    b.load_state_dict(a.state_dict())
    ya, yb = a(src, pos, shapes, ss, lsi), b(src, pos, shapes, ss, lsi)
    assert (ya - yb).abs().max().item() < 5e-2 * ya.abs().max().item()


def test_graphed_step_replays_the_eager_step():
    """uninext_b200.graphs.GraphedStep: the whole fwd+bwd of a small stack captured in a CUDA graph reproduces the eager
    gradients (the op and the caller kernels are capture-safe: caller's stream, no syncs, no host reads of the level table)."""
    from uninext_b200.dp import FlatGradBucket
    from uninext_b200.graphs import GraphedStep
    cfg = CONFIGS["cfg1"]
    ss, lsi = level_tensors(cfg.shapes, DEV)
    torch.manual_seed(11)
    stack = DeformableStack(num_layers=2, num_queries=30, d_ffn=256).to(DEV)
    bucket = FlatGradBucket(stack.parameters())
    src = torch.randn(2, cfg.S, 256, device=DEV)
    pos = torch.randn(2, cfg.S, 256, device=DEV)

    def step():
        bucket.zero_()
        stack(src, pos, cfg.shapes, ss, lsi).square().mean().backward()
    step()
    want = bucket.flat.clone()
    g = GraphedStep(step)
    for _ in range(2):
        bucket.flat.fill_(123.0)                    # must be overwritten by the replay
        g.replay()
        torch.cuda.synchronize()
        scale = want.abs().max().item()
        assert (bucket.flat - want).abs().max().item() <= 1e-4 * scale       # grad_value atomics reorder between runs
    src.mul_(0.5)                                   # inputs are read from the same storage on every replay
    g.replay(); step_ref = bucket.flat.clone()
    step()
    assert (bucket.flat - step_ref).abs().max().item() <= 1e-4 * bucket.flat.abs().max().item()
