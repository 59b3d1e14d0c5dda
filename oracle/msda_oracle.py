"""TEST INFRASTRUCTURE ONLY -- Python face of the CPU oracle.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module. The product package ``uninext_b200`` never does (tests/test_no_oracle_in_product.py
enforces it).

Two restatements live here:

* ``forward`` / ``backward``: ctypes calls into ``oracle/libmsda_oracle.so`` (plain C, ``oracle/msda_oracle.c``),
  the bit-level restatement of the reference CUDA kernels' arithmetic
  (``ops/src/cuda/ms_deform_im2col_cuda.cuh:33-159,237-403``). numpy arrays in, numpy arrays out.
* ``core_pytorch_port``: a torch restatement of the reference's *CPU path*, ``ms_deform_attn_core_pytorch``
  (``ops/functions/ms_deform_attn_func.py:43-63``): one ``grid_sample`` per level, weighted sum. The
  ``--impl reference`` arm of bench.py times the reference file itself when ``build()`` staged it (``oracle/refpy.py``
  -> ``oracle/_ref/``) and falls back to this port otherwise.

Parity pin: both are checked against golden vectors generated from the reference's own
``ms_deform_attn_core_pytorch`` (tests/golden/make_golden.py -> tests/golden/*.npz).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmsda_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle/msda_oracle.c with gcc (seconds). Returns the library path."""
    srcs = [os.path.join(_HERE, f) for f in ("msda_oracle.c", "msda_oracle_impl.h")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs if os.path.exists(s))
    if stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "all"])
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        assert _lib.msda_oracle_abi_version() == 1
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _prep(value, shapes, lsi, loc, attn):
    dt = np.float64 if value.dtype == np.float64 else np.float32
    value = np.ascontiguousarray(value, dtype=dt)
    loc = np.ascontiguousarray(loc, dtype=dt)
    attn = np.ascontiguousarray(attn, dtype=dt)
    shapes = np.ascontiguousarray(shapes, dtype=np.int64)
    lsi = np.ascontiguousarray(lsi, dtype=np.int64)
    N, S, M, D = value.shape
    _, Lq, M2, L, P, two = loc.shape
    assert M2 == M and two == 2 and attn.shape == (N, Lq, M, L, P) and shapes.shape == (L, 2)
    dims = [ctypes.c_int(int(v)) for v in (N, S, M, D, L, Lq, P)]
    return dt, value, shapes, lsi, loc, attn, dims, (N, S, M, D, L, Lq, P)


def forward(value, spatial_shapes, level_start_index, sampling_locations, attention_weights):
    """numpy in / numpy out; fp32 or fp64 decided by ``value.dtype``. Returns out[N, Lq, M*D]."""
    dt, value, shapes, lsi, loc, attn, dims, (N, S, M, D, L, Lq, P) = _prep(
        value, spatial_shapes, level_start_index, sampling_locations, attention_weights)
    out = np.empty((N, Lq, M * D), dtype=dt)
    fn = getattr(_load(), "msda_oracle_forward_f64" if dt == np.float64 else "msda_oracle_forward_f32")
    fn.restype = None
    fn(_ptr(value), _ptr(shapes), _ptr(lsi), _ptr(loc), _ptr(attn), *dims, _ptr(out))
    return out


def backward(grad_output, value, spatial_shapes, level_start_index, sampling_locations, attention_weights):
    """Returns (grad_value, grad_sampling_locations, grad_attention_weights)."""
    dt, value, shapes, lsi, loc, attn, dims, (N, S, M, D, L, Lq, P) = _prep(
        value, spatial_shapes, level_start_index, sampling_locations, attention_weights)
    g = np.ascontiguousarray(grad_output, dtype=dt).reshape(N, Lq, M * D)
    gv = np.empty_like(value)
    gl = np.empty_like(loc)
    ga = np.empty_like(attn)
    fn = getattr(_load(), "msda_oracle_backward_f64" if dt == np.float64 else "msda_oracle_backward_f32")
    fn.restype = None
    fn(_ptr(g), _ptr(value), _ptr(shapes), _ptr(lsi), _ptr(loc), _ptr(attn), *dims, _ptr(gv), _ptr(gl), _ptr(ga))
    return gv, gl, ga


def core_pytorch_port(value, spatial_shapes, sampling_locations, attention_weights):
    """torch restatement of the reference CPU path (ms_deform_attn_func.py:43-63).

    Per level: view the level's slab as an image batch [N*M, D, H, W], sample it with
    ``grid_sample(bilinear, zeros, align_corners=False)`` at ``2*loc-1``; then contract the L*P taps with the
    attention weights. Differentiable through torch.autograd (that is how the reference obtains CPU gradients).
    """
    import torch
    import torch.nn.functional as F

    n, s, m, d = value.shape
    lq, nl, npnt = sampling_locations.shape[1], sampling_locations.shape[3], sampling_locations.shape[4]
    hw = [(int(h), int(w)) for h, w in spatial_shapes.tolist()] if hasattr(spatial_shapes, "tolist") \
        else [(int(h), int(w)) for h, w in spatial_shapes]
    grids = sampling_locations * 2 - 1                                   # func.py:50
    # heads become part of the image batch: [N, S, M, D] -> [N, M, D, S]
    planes = value.permute(0, 2, 3, 1)
    taps = []
    start = 0
    for lvl, (h, w) in enumerate(hw):
        img = planes[..., start:start + h * w].reshape(n * m, d, h, w)   # func.py:53-54
        start += h * w
        g = grids[:, :, :, lvl].permute(0, 2, 1, 3, 4).reshape(n * m, lq, npnt, 2)   # func.py:56
        taps.append(F.grid_sample(img, g, mode="bilinear", padding_mode="zeros", align_corners=False))  # :58-59
    sampled = torch.stack(taps, dim=3).reshape(n * m, d, lq, nl * npnt)  # [N*M, D, Lq, L*P]
    wts = attention_weights.permute(0, 2, 1, 3, 4).reshape(n * m, 1, lq, nl * npnt)   # func.py:61
    out = (sampled * wts).sum(-1).reshape(n, m * d, lq)                  # func.py:62
    return out.transpose(1, 2).contiguous()                             # func.py:63
