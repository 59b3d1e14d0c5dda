"""Size-independent properties of the CPU oracle (no GPU): the same identities the GPU tests use at full size, plus a
finite-difference check of the oracle's backward -- independent evidence next to the golden vectors."""
import numpy as np
import pytest

from oracle import msda_oracle


def _case(rng, shapes, N, M, D, Lq, P, lo=-0.2, hi=1.2):
    ss = np.array(shapes, dtype=np.int64)
    lsi = np.concatenate(([0], np.cumsum(ss[:, 0] * ss[:, 1])[:-1])).astype(np.int64)
    S = int((ss[:, 0] * ss[:, 1]).sum())
    L = len(shapes)
    value = rng.standard_normal((N, S, M, D))
    loc = rng.uniform(lo, hi, (N, Lq, M, L, P, 2))
    attn = rng.uniform(0.0, 1.0, (N, Lq, M, L, P))
    gout = rng.standard_normal((N, Lq, M * D))
    return value, ss, lsi, loc, attn, gout


def test_sampling_at_pixel_centres_returns_the_pixel():
    rng = np.random.default_rng(0)
    H, W, M, D = 5, 7, 2, 3
    value = rng.standard_normal((1, H * W, M, D))
    ss, lsi = np.array([[H, W]], dtype=np.int64), np.array([0], dtype=np.int64)
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    loc = np.stack(((xs.reshape(-1) + 0.5) / W, (ys.reshape(-1) + 0.5) / H), -1)          # pixel centres, (x, y)
    loc = np.broadcast_to(loc[None, :, None, None, None, :], (1, H * W, M, 1, 1, 2)).copy()
    attn = np.ones((1, H * W, M, 1, 1))
    out = msda_oracle.forward(value, ss, lsi, loc, attn)
    np.testing.assert_allclose(out.reshape(1, H * W, M, D), value, rtol=0, atol=1e-12)


def test_linearity_and_euler_identities():
    rng = np.random.default_rng(1)
    value, ss, lsi, loc, attn, gout = _case(rng, [(6, 5), (3, 4), (1, 2)], N=2, M=3, D=8, Lq=11, P=2)
    out = msda_oracle.forward(value, ss, lsi, loc, attn)
    v2 = rng.standard_normal(value.shape)
    out2 = msda_oracle.forward(v2, ss, lsi, loc, attn)
    np.testing.assert_allclose(msda_oracle.forward(2.5 * value - v2, ss, lsi, loc, attn), 2.5 * out - out2, atol=1e-12)
    np.testing.assert_allclose(msda_oracle.forward(value, ss, lsi, loc, 3 * attn), 3 * out, atol=1e-12)
    gv, gl, ga = msda_oracle.backward(gout, value, ss, lsi, loc, attn)
    inner = float((out * gout).sum())
    assert abs(inner - float((value * gv).sum())) < 1e-9 * abs(inner) + 1e-12       # out is linear in value ...
    assert abs(inner - float((attn * ga).sum())) < 1e-9 * abs(inner) + 1e-12        # ... and in the weights


def test_taps_outside_the_window_contribute_nothing():
    rng = np.random.default_rng(2)
    value, ss, lsi, loc, attn, gout = _case(rng, [(4, 4)], N=1, M=1, D=4, Lq=6, P=3, lo=1.2, hi=3.0)   # x*W-0.5 >= W
    assert np.all(msda_oracle.forward(value, ss, lsi, loc, attn) == 0)
    gv, gl, ga = msda_oracle.backward(gout, value, ss, lsi, loc, attn)
    assert np.all(gv == 0) and np.all(gl == 0) and np.all(ga == 0)


@pytest.mark.parametrize("seed", [3, 4])
def test_backward_matches_finite_differences(seed):
    """Central differences in fp64 on taps kept away from cell boundaries (where the location gradient jumps)."""
    rng = np.random.default_rng(seed)
    shapes = [(5, 6), (3, 3)]
    value, ss, lsi, loc, attn, gout = _case(rng, shapes, N=1, M=2, D=3, Lq=4, P=2, lo=0.05, hi=0.95)
    for l, (h, w) in enumerate(shapes):                 # snap fractional pixel offsets into [0.2, 0.8]
        for k, n in ((0, w), (1, h)):
            pix = loc[:, :, :, l, :, k] * n - 0.5
            frac = pix - np.floor(pix)
            loc[:, :, :, l, :, k] = (np.floor(pix) + 0.2 + 0.6 * frac + 0.5) / n
    gv, gl, ga = msda_oracle.backward(gout, value, ss, lsi, loc, attn)
    f = lambda v, lo_, at: float((msda_oracle.forward(v, ss, lsi, lo_, at) * gout).sum())
    eps = 1e-6
    for arr, grad, name in ((value, gv, "value"), (loc, gl, "loc"), (attn, ga, "attn")):
        flat = arr.reshape(-1)
        for idx in rng.choice(flat.size, size=12, replace=False):
            old = flat[idx]
            flat[idx] = old + eps; fp = f(value, loc, attn)
            flat[idx] = old - eps; fm = f(value, loc, attn)
            flat[idx] = old
            num = (fp - fm) / (2 * eps)
            assert abs(num - grad.reshape(-1)[idx]) < 1e-6 * (1 + abs(num)), (name, idx, num, grad.reshape(-1)[idx])


def test_fp32_and_fp64_oracles_agree():
    rng = np.random.default_rng(5)
    value, ss, lsi, loc, attn, gout = _case(rng, [(8, 8), (4, 4)], N=2, M=4, D=16, Lq=9, P=4)
    o64 = msda_oracle.forward(value, ss, lsi, loc, attn)
    o32 = msda_oracle.forward(value.astype(np.float32), ss, lsi, loc.astype(np.float32), attn.astype(np.float32))
    assert np.abs(o32 - o64).max() < 1e-5 * np.abs(o64).max()
