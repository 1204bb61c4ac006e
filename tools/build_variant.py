"""Builds libibgs_b200 with extra nvcc defines into ibgs_b200/_lib/variants/<name>/ (kernel experiments; dev tool).
usage: python tools/build_variant.py <name> -DIBGS_EXP=1 ...   then   IBGS_B200_LIB=<printed path> python tools/quick_ab.py ..."""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ibgs_b200 import build as B
name, defs = sys.argv[1], sys.argv[2:]
out = os.path.join(B.OUT_DIR, "variants", name)
os.makedirs(out, exist_ok=True)
def cc(src):
    obj = os.path.join(out, src.replace(".cu", ".o"))
    r = subprocess.run([B.NVCC] + B.FLAGS + defs + ["-c", os.path.join(B.CSRC, src), "-o", obj], capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError(r.stderr)
    open(obj + ".ptxas.log", "w").write(r.stderr)
    return obj
with ThreadPoolExecutor(8) as ex:
    objs = list(ex.map(cc, B.SOURCES))
lib = os.path.join(out, "libibgs_b200.so")
subprocess.run([B.NVCC, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"], check=True)
print(lib)
