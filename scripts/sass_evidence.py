#!/usr/bin/env python
"""SASS instruction counts of libpmf.so (cuobjdump -sass; runs without a GPU): the proof that the tensor-core kernels
are tcgen05 / TMEM / TMA code (UTCHMMA, LDTM / STTM, UTMALDG / UTMAREDG / UTMASTG) and hold no legacy mma.sync (HMMA).
    python scripts/sass_evidence.py > profiles/r2_sass_evidence.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pathmatfac.jl_b200", "libpmf.so")

WHOLE = [("UTCHMMA", r"UTCHMMA"), ("UTCQMMA", r"UTCQMMA"), ("UTCBAR", r"UTCBAR"), ("LDTM", r"LDTM"), ("STTM", r"STTM"),
         ("UTMALDG", r"UTMALDG"), ("UTMAREDG", r"UTMAREDG"), ("UTMASTG", r"UTMASTG"),
         ("UTMAPF (TMA L2 prefetch, debug instantiation)", r"UTMAPF"), ("USETMAXREG", r"USETMAXREG"),
         ("SYNCS.ARRIVE", r"SYNCS\.ARRIVE"), ("SYNCS.PHASECHK", r"SYNCS\.PHASECHK"), ("REDG.E.ADD.F32x4", r"REDG\.E\.ADD\.F32x4"),
         ("MUFU.EX2", r"MUFU\.EX2"), ("MUFU.LG2", r"MUFU\.LG2"), ("MUFU.RCP", r"MUFU\.RCP"),
         ("HMMA (legacy mma.sync; excludes UTCHMMA)", r"(?<!UTC)HMMA")]
PER = ["UTCHMMA", "UTMALDG", "UTMAREDG", "UTMASTG", "LDTM", "STTM", "USETMAXREG"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    print(f"# SASS evidence (cuobjdump -sass pathmatfac.jl_b200/libpmf.so, sm_100a, built from the tree this file is "
          f"committed with: scripts/sass_evidence.py), instruction counts over the whole library")
    for label, pat in WHOLE:
        print(f"{label:<48} {len(re.findall(pat, sass))}")
    print("\n# per kernel: tcgen05.mma (UTCHMMA) / TMA load / TMA reduce / TMA store / TMEM ld / TMEM st / setmaxnreg")
    for part in re.split(r"\n\s*Function : ", sass)[1:]:
        name = part.split("\n", 1)[0].strip()
        c = {m: len(re.findall(m, part)) for m in PER}
        if c["UTCHMMA"]:
            print("  " + "  ".join(f"{m} {c[m]:>3}" for m in PER) + "   " + name)
    print("\n# data_pass_tc_kernel<DBG, BATCH, THR>: <0,0,0> production, <0,1,*> batch shift/scale layers (two-halves epilogue: "
          "half the TMEM load / store instructions of the pipelined form), <*,*,1> ordinal-threshold gradients, <1,0,0> "
          "ablation / trace build;\n# zlink_kernel / grad_gemm_kernel<A_MN>: K > 64 (wide_tc.cu)")


if __name__ == "__main__":
    sys.exit(main())
