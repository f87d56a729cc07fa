"""Builds a variant of libdiffphar_b200.so with extra -D flags for ONE translation unit (same-box A/B through
DIFFPHAR_LIB):  python scripts/build_variant.py <tag> <file.cu> -DNAME=VALUE ...  ->  cmd_gen_b200/libdiffphar_b200_<tag>.so"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cmd_gen_b200 import build as B   # noqa: E402

tag, unit, defs = sys.argv[1], sys.argv[2], sys.argv[3:]
B.build()
obj = os.path.join(B.OBJ, unit.replace(".cu", f".{tag}.o"))
subprocess.run([os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")] + [f for f in B.NVCC_FLAGS if f not in ("-Xptxas", "-v")] + defs +
               ["-c", os.path.join(B.CSRC, unit), "-o", obj], check=True)
objs = [obj if s == unit else os.path.join(B.OBJ, s.replace(".cu", ".o")) for s in B.SOURCES]
out = os.path.join(B.HERE, f"libdiffphar_b200_{tag}.so")
subprocess.run([os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc"), "-shared", "-o", out] + objs +
               ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"], check=True)
print(out)
