#!/usr/bin/env python3
"""A/B builds: recompile ONE source of the library with extra nvcc flags and link it with the other (already built)
objects into build_ab/libfhe_<name>.so.  Select it at run time with FHE_B200_LIB=build_ab/libfhe_<name>.so.
usage: tools/build_variant.py NAME SOURCE.cu [-DX=1 ...]"""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from fhe_icp_b200 import _native as N  # noqa: E402

name, src, extra = sys.argv[1], sys.argv[2], sys.argv[3:]
N.build()
out = ROOT / "build_ab"
out.mkdir(exist_ok=True)
obj = out / f"{Path(src).stem}_{name}.o"
flags = [f for f in N.NVCC_FLAGS if f != "-shared"]
subprocess.check_call(["nvcc", *flags, *extra, "-c", "-o", str(obj), str(N._CSRC / src)])
objs = [str(obj) if s == src else str(N._OBJ_DIR / (Path(s).stem + ".o")) for s in N._SOURCES]
so = out / f"libfhe_{name}.so"
subprocess.check_call(["nvcc", "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(so)] + objs)
print(so)
