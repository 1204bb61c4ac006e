"""Summarise an ncu report's source page per CUDA source line (dev tool).
usage: python tools/ncu_src.py report.ncu-rep [--launch N] [--top 25]"""
import argparse, csv, io, subprocess

ap = argparse.ArgumentParser()
ap.add_argument("rep"); ap.add_argument("--launch", type=int, default=0); ap.add_argument("--top", type=int, default=25)
a = ap.parse_args()
cmd = ["ncu", "-i", a.rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(a.launch), "--launch-count", "1"]
txt = subprocess.run(cmd, capture_output=True, text=True).stdout
rows_all, fname, hdr = [], "", None
for r in csv.reader(io.StringIO(txt)):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1]; continue
    if r[0] == "Function Name":
        print("kernel:", r[1][:100]); continue
    if r[0] == "Line No":
        hdr = r
        iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
        stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or r[0] == "":
        continue
    try:
        cur = dict(file=fname, line=r[0], src=r[1].strip(), samples=int(r[iS] or 0), inst=int(r[iI] or 0), stalls={})
    except (ValueError, IndexError):
        continue
    for i in stall_cols:
        if i < len(r) and r[i] not in ("", "0"):
            cur["stalls"][hdr[i][6:]] = int(r[i])
    rows_all.append(cur)
tot = sum(r["samples"] for r in rows_all) or 1
toti = sum(r["inst"] for r in rows_all) or 1
print(f"total samples {tot}, total warp-instructions {toti}")
for r in sorted(rows_all, key=lambda r: -r["samples"])[:a.top]:
    st = ",".join(f"{k}:{v}" for k, v in sorted(r["stalls"].items(), key=lambda kv: -kv[1])[:4])
    print(f"{r['file'].split('/')[-1][:20]:20s}:{r['line']:>4s} smp {100*r['samples']/tot:5.1f}% inst {100*r['inst']/toti:5.1f}% | {r['src'][:64]:64s} | {st}")
