#!/usr/bin/env python
"""Regenerate profiles/sass/: one gzip'ed `cuobjdump -sass` extract per production kernel of uninext_b200/lib/libmsda_b200.so
and a table of the mnemonics that show what the code runs on (no GPU needed).

    python tools/sass_extract.py
"""
import gzip
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "uninext_b200", "lib", "libmsda_b200.so")
OUT = os.path.join(ROOT, "profiles", "sass")

# name -> regex on the mangled function name (first match wins)
KERNELS = [
    ("fwd_tiled_f32", r"msda_fwd_tiledIfLi4ELi32ELi16ELi4ELb1ELb0E"),
    ("bwd_tiled_f32", r"msda_bwd_tiledIfLi4ELi32ELi16ELi2ELb1ELb0ELb0E"),
    ("bwd_tiled_f32_split", r"msda_bwd_tiledIfLi4ELi32ELi16ELi2ELb0ELb1ELb0E"),
    ("fwd_tiled_bf16", r"msda_fwd_tiledI13__nv_bfloat16Li8ELi32ELi16ELi4ELb1ELb0E"),
    ("bwd_tiled_bf16", r"msda_bwd_tiledI13__nv_bfloat16Li4ELi32ELi16ELi2ELb1ELb0ELb0E"),
    ("zero_fill", r"msda_zero_fill"),
    ("gemm_tf32_streaming", r"gemm18linear_tf32_kernel"),
    ("gemm_tf32_w_stationary", r"gemm21linear_tf32_ws_kernel"),
    ("gemm_tf32_w_stationary_2cta", r"gemm22linear_tf32_ws2_kernel"),
    ("bwd_slab_f32", r"msda_bwd_slabIfLi16E"),
    ("condinst_fwd", r"msda12condinst_fwd"),
]
COLS = [("LDG.E.128 / .256", r"\bLDG\.E\.(128|256|ENL2\.256)"), ("LDS", r"\bLDS"), ("STS", r"\bSTS"), ("REDG", r"\bREDG"),
        ("FFMA", r"\bFFMA"), ("SHFL", r"\bSHFL"), ("UBLKCP (TMA bulk)", r"\bUBLKCP"), ("UTMALDG (TMA load)", r"\bUTMALDG"),
        ("UTMASTG (TMA store)", r"\bUTMASTG"), ("UTCHMMA (tcgen05.mma)", r"\bUTCHMMA"), ("LDTM (tcgen05.ld)", r"\bLDTM"),
        ("SYNCS (mbarrier)", r"\bSYNCS"), ("ACQBULK (griddepcontrol.wait)", r"\bACQBULK"),
        ("PREEXIT (griddepcontrol.launch_dependents)", r"\bPREEXIT")]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], check=True, capture_output=True, text=True).stdout
    parts = re.split(r"(?m)^\s*Function : ", txt)[1:]
    funcs = {}
    for p in parts:
        name, _, body = p.partition("\n")
        funcs[name.strip()] = body
    os.makedirs(OUT, exist_ok=True)
    rows = []
    for short, pat in KERNELS:
        hit = next((n for n in funcs if re.search(pat, n)), None)
        if hit is None:
            print(f"warning: no function matches {pat}", file=sys.stderr)
            continue
        body = funcs[hit]
        with gzip.open(os.path.join(OUT, short + ".sass.gz"), "wt") as fh:
            fh.write("Function : " + hit + "\n" + body)
        instr = [ln for ln in body.splitlines() if re.search(r"/\*[0-9a-f]{4}\*/\s+\S", ln)]
        counts = [sum(1 for ln in instr if re.search(rx, ln)) for _, rx in COLS]
        rows.append(f"| {short} (`{hit[:64]}`) | {len(instr)} | " + " | ".join(str(c) for c in counts) + " |")
    with open(os.path.join(OUT, "README.md"), "w") as fh:
        fh.write("# SASS extracts of the production kernels (`python tools/sass_extract.py`: cuobjdump -sass uninext_b200/lib/libmsda_b200.so, nvcc 12.9, sm_100a).\n"
                 "# One file per kernel next to this one; the table counts the mnemonics that show what the code runs on.\n\n")
        fh.write("| kernel | instructions | " + " | ".join(c for c, _ in COLS) + " |\n")
        fh.write("|---|---|" + "---|" * len(COLS) + "\n")
        fh.write("\n".join(rows) + "\n")
    print("\n".join(rows))


if __name__ == "__main__":
    main()
