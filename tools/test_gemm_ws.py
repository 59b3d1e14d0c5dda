#!/usr/bin/env python
"""Bring-up + timing of the W-stationary tcgen05 GEMM (msda_linear_tf32_ex): correctness against fp64 (TF32 tolerance)
incl. the fused tail (bias, row mask, ReLU) and ragged M, then time vs the streaming kernel and cuBLAS TF32.
Run under `timeout` on the GPU box: a pipeline bug in such a kernel is a hang, not a wrong number."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uninext_b200 import _cabi
lib = _cabi.load()
S = lambda: torch.cuda.current_stream().cuda_stream

def ws(a, w, b, mask=None, relu=False):
    m, k = a.shape; n = w.shape[0]
    c = torch.empty((m, n), dtype=torch.float32, device=a.device)
    rc = lib.msda_linear_tf32_ex(a.data_ptr(), w.data_ptr(), b.data_ptr() if b is not None else None,
                                 mask.data_ptr() if mask is not None else None, m, n, k, int(relu), c.data_ptr(), S())
    assert rc == 0, (rc, lib.msda_strerror(rc))
    return c

torch.manual_seed(0)
ok = True
TIMING_ONLY = os.environ.get("MSDA_GEMM_WS_EPI", "")[:1] in ("n", "l", "x")          # experiment modes write no output
for (m, n, k, use_mask, relu, use_bias) in [(128, 256, 256, 0, 0, 1), (64, 64, 32, 0, 0, 1), (1, 128, 64, 0, 0, 1), (300, 256, 256, 1, 0, 1),
                                            (1000, 192, 128, 1, 1, 1), (44646, 256, 256, 1, 0, 1), (44646, 256, 256, 0, 1, 0),
                                            (513, 64, 256, 0, 0, 1), (77, 128, 96, 1, 1, 0), (20000, 256, 32, 0, 0, 1)]:
    assert lib.msda_linear_tf32_ws_ok(n, k) == 1, (n, k)
    a = torch.randn(m, k, device="cuda"); w = torch.randn(n, k, device="cuda") * 0.1
    b = torch.randn(n, device="cuda") if use_bias else None
    mask = (torch.rand(m, device="cuda") < 0.3).to(torch.uint8) if use_mask else None
    c = ws(a, w, b, mask, relu)
    torch.cuda.synchronize()
    ref = a.double() @ w.double().t()
    if b is not None: ref = ref + b.double()
    if relu: ref = ref.clamp_min(0)
    if mask is not None: ref = ref.masked_fill(mask.bool()[:, None], 0.0)
    err = ((c.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()
    exact_mask = mask is None or bool((c[mask.bool()] == 0).all())
    good = (err < 2e-3 and exact_mask) or TIMING_ONLY
    print(f"M={m} N={n} K={k} mask={use_mask} relu={relu} bias={use_bias}: rel err {err:.2e}", "OK" if good else "FAIL", flush=True)
    ok &= good
if not ok:
    sys.exit(1)
torch.backends.cuda.matmul.allow_tf32 = True
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for (m, n, k) in [(44646, 256, 256), (65280, 256, 256), (25500, 256, 256), (44646, 128, 256), (44646, 256, 128)]:
    a = torch.randn(m, k, device="cuda"); w = torch.randn(n, k, device="cuda") * 0.1; b = torch.randn(n, device="cuda")
    c = torch.empty(m, n, device="cuda")
    def stream_kernel():
        os.environ["MSDA_GEMM_WS"] = "0"
    mask = (torch.rand(m, device="cuda") < 0.1)
    mask8 = mask.to(torch.uint8)
    # value_proj as the reference runs it (ops/modules/ms_deform_attn.py:95-97): Linear, then masked_fill of the padded rows
    fns = (("tcgen05-ws", lambda: ws(a, w, b)), ("cublas-tf32", lambda: torch.addmm(b, a, w.t(), out=c)),
           ("tcgen05-ws+mask", lambda: ws(a, w, b, mask8)),
           ("cublas+masked_fill", lambda: torch.addmm(b, a, w.t(), out=c).masked_fill_(mask[:, None], 0.0)))
    for name, fn in fns:
        for _ in range(5): fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(20):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        # back-to-back (L2-warm) figure as well, like round 1's numbers
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        print(f"{name:18s} M={m} N={n} K={k}: cold-L2 median {ts[10]*1e3:.1f} us, back-to-back {e0.elapsed_time(e1)/20*1e3:.1f} us", flush=True)
