"""Compile side-by-side variants of the CUDA library with different -D knobs
(same C ABI) into casmcode_monte_b200/_variants/ for A/B runs on the GPU box:

  python tools/ab_build.py ctas5:-DCMG_BULK_CTAS=5 pair:-DCMG_BULK_PAIR=1
  CMG_LIB_PATH=casmcode_monte_b200/_variants/lib_ctas5.so python tools/sweep_variants.py
"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from casmcode_monte_b200 import build

out_dir = os.path.join(build.HERE, "_variants")
os.makedirs(out_dir, exist_ok=True)
nvcc = "/usr/local/cuda/bin/nvcc"
src = os.path.join(build.HERE, "csrc", "cmg_capi.cu")
for spec in sys.argv[1:]:
    name, _, flags = spec.partition(":")
    out = os.path.join(out_dir, f"lib_{name}.so")
    cmd = [nvcc] + build.NVCC_FLAGS + ["-Xptxas", "-v"] + [f for f in flags.split(",") if f] + [src, "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        print(r.stderr[-3000:])
        raise SystemExit(1)
    lines = r.stderr.splitlines()
    for i, l in enumerate(lines):
        if "Function properties for" in l and ("bulk2d" in l or "bulk3d" in l):
            print(name, l.split("for ")[1][:40], lines[i + 1].strip(), "|", lines[i + 2].split(":")[1].strip()[:20])
