"""Opcode census of the built library's SASS (dev tool): per object file, how many of the instructions that characterise
the design are present.  usage: python tools/sass_census.py > profiles/r2_sass_census.md"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ibgs_b200", "_lib")
OPS = ["LDGSTS", "REDG", "RED.", "ATOMG", "TEX", "SULD", "SUST", "FFMA2", "FMUL2", "FADD2", "MUFU.EX2", "MUFU.RCP", "SHFL", "VOTE",
       "MATCH", "REDUX", "LDS.128", "STS.128", "LDG.E.128", "STG.E.128", "BAR.SYNC", "UTMALDG", "UTCMMA", "LDTM", "HMMA", "LDL", "STL"]
print("# SASS census of libibgs_b200.so (sm_100a), `cuobjdump -sass` per object file\n")
print("Counts of static instructions. `LDL` / `STL` = local-memory traffic (register spills or dynamically indexed arrays);")
print("`UTMALDG` / `UTCMMA` / `LDTM` (TMA, tcgen05) are absent by design: nothing on this path is a dense contraction, and the")
print("record fetch is a per-lane gather (`LDGSTS` = cp.async).  `HMMA`: none either (the colour network's convolutions are cuDNN's).\n")
print("| object | kernels | total | " + " | ".join(OPS) + " |")
print("|---|---|---|" + "---|" * len(OPS))
for f in sorted(os.listdir(LIB)):
    if not f.endswith(".o"):
        continue
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(LIB, f)], capture_output=True, text=True).stdout
    lines = [l for l in sass.splitlines() if re.search(r"/\*[0-9a-f]{4}\*/", l)]
    nk = sass.count("Function :")
    row = [sum(1 for l in lines if re.search(r"\b" + re.escape(op), l)) for op in OPS]
    print(f"| {f} | {nk} | {len(lines)} | " + " | ".join(str(c) for c in row) + " |")
