"""CPU tests: the host-side mirrors of the reference's helpers around the op (reference points, proposals, sine position
embedding, MLP) against the REFERENCE functions themselves, imported from the staged copy ``tests/_ref``
(tests/stage_reference.py; skipped when it has not been staged)."""
import pytest
import torch

from tests import stage_reference

pytestmark = pytest.mark.skipif(not stage_reference.stage(), reason="tests/_ref not staged (needs /root/reference once)")


@pytest.fixture(scope="module")
def ref():
    return stage_reference.import_reference()


SHAPES = [(12, 20), (6, 10), (3, 5), (2, 3)]


def _masks(n, shapes, gen):
    """Padding masks the way the reference pads a batch: valid region top-left, padded right / bottom."""
    out = []
    fr = torch.rand(n, 2, generator=gen) * 0.5 + 0.5
    for h, w in shapes:
        m = torch.ones(n, h, w, dtype=torch.bool)
        for b in range(n):
            vh, vw = max(1, int(round(h * fr[b, 0].item()))), max(1, int(round(w * fr[b, 1].item())))
            m[b, :vh, :vw] = False
        out.append(m)
    return out


def test_sine_pos_embed_and_mlp(ref):
    from uninext_b200.modules.deformable_transformer import MLP, get_sine_pos_embed
    dino = ref[3]
    g = torch.Generator().manual_seed(0)
    pos = torch.rand(2, 7, 4, generator=g)
    for xy in (True, False):
        assert torch.allclose(get_sine_pos_embed(pos, exchange_xy=xy), dino.get_sine_pos_embed(pos, exchange_xy=xy),
                              rtol=0, atol=1e-6)
    assert torch.allclose(get_sine_pos_embed(pos[..., :2], 64, 20), dino.get_sine_pos_embed(pos[..., :2], 64, 20), atol=1e-6)
    torch.manual_seed(1)
    a, b = MLP(512, 256, 256, 2), dino.MLP(512, 256, 256, 2)
    a.load_state_dict(b.state_dict(), strict=True)
    x = torch.randn(3, 5, 512, generator=g)
    assert torch.equal(a(x), b(x))


def test_reference_points_and_valid_ratios(ref):
    from uninext_b200.modules.deformable_transformer import get_reference_points, valid_ratios_from_masks
    dino = ref[3]
    g = torch.Generator().manual_seed(2)
    masks = _masks(3, SHAPES, g)
    holder = dino.DeformableTransformerVLDINO.__new__(dino.DeformableTransformerVLDINO)       # get_valid_ratio uses no state
    want_vr = torch.stack([dino.DeformableTransformerVLDINO.get_valid_ratio(holder, m) for m in masks], 1)
    vr = valid_ratios_from_masks(masks)
    assert torch.equal(vr, want_vr)
    ss = torch.as_tensor(SHAPES)
    want = dino.DeformableTransformerEncoderVL.get_reference_points(ss, want_vr, device="cpu")
    got = get_reference_points(ss, vr)
    assert got.shape == want.shape and torch.allclose(got, want, rtol=1e-6, atol=1e-7)
    assert torch.allclose(get_reference_points(SHAPES, vr), want, rtol=1e-6, atol=1e-7)        # cached grid, list input


def test_encoder_output_proposals(ref):
    from uninext_b200.modules.deformable_transformer import gen_encoder_output_proposals
    dino = ref[3]
    g = torch.Generator().manual_seed(3)
    masks = _masks(2, SHAPES, g)
    flat = torch.cat([m.flatten(1) for m in masks], 1)
    s = flat.shape[1]
    memory = torch.randn(2, s, 16, generator=g)

    class Holder(torch.nn.Module):          # the reference method reads self.enc_output / self.enc_output_norm
        def __init__(self):
            super().__init__()
            self.enc_output = torch.nn.Linear(16, 16)
            self.enc_output_norm = torch.nn.LayerNorm(16)
    h = Holder()
    want_mem, want_prop = dino.DeformableTransformerVLDINO.gen_encoder_output_proposals(h, memory, flat, torch.as_tensor(SHAPES))
    prop, keep = gen_encoder_output_proposals(flat, SHAPES)
    assert torch.equal(torch.isinf(prop), torch.isinf(want_prop))
    fin = ~torch.isinf(want_prop)
    assert torch.allclose(prop[fin], want_prop[fin], rtol=1e-5, atol=1e-6)
    got_mem = h.enc_output_norm(h.enc_output(memory.masked_fill(~keep, 0.0)))
    assert torch.allclose(got_mem, want_mem, rtol=1e-5, atol=1e-6)


def test_layer_signatures_match_reference(ref):
    import inspect
    from uninext_b200.modules.deformable_layers import (DeformableTransformerDecoderLayer,
                                                        DeformableTransformerEncoderLayer)
    from uninext_b200.modules.deformable_transformer import DeformableReidHead
    dino = ref[3]
    # positional signature = the reference's; keyword-only extras (projected_value) are this repo's extensions
    names = lambda f: [n for n, p in inspect.signature(f).parameters.items() if p.kind != p.KEYWORD_ONLY][1:]
    assert names(DeformableTransformerEncoderLayer.forward) == names(dino.DeformableTransformerEncoderLayer.forward)
    assert names(DeformableTransformerDecoderLayer.forward) == names(dino.DeformableTransformerDecoderLayer.forward)
    assert names(DeformableReidHead.forward) == names(dino.DeformableReidHead.forward)
    ours = DeformableReidHead(256, DeformableTransformerDecoderLayer(256, 512, 0.0, "relu", 4, 8, 4), 2)
    theirs = dino.DeformableReidHead(256, dino.DeformableTransformerDecoderLayer(256, 512, 0.0, "relu", 4, 8, 4), 2)
    assert {k: tuple(v.shape) for k, v in ours.state_dict().items()} == \
        {k: tuple(v.shape) for k, v in theirs.state_dict().items()}
    ours.load_state_dict(theirs.state_dict(), strict=True)
