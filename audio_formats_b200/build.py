"""In-tree build of the CUDA shared library (sm_100a only; nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libl3b200.so"
SOURCES = ["l3_kernels.cu", "l3_entropy.cu", "l12_kernels.cu", "l3_ctx.cu", "l3_raw.cu", "l3_host.cpp", "l3_stream.cpp", "l3_pipeline.cpp"]
HEADERS = ["l3_kernels.cuh", "l3_desc.cuh", "l3_host.hpp", "l3_format.hpp", "l3_device_tables.hpp", "l3_tables_gen.h", "l12_tables.h",
           "../../include/l3b200.h"]

# -fmad=false is part of the numerical contract: the float pipeline must round exactly like the
# un-fused scalar reference (DESIGN.md "Bit-exactness").  Never remove it without re-running parity.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC,-Wall,-O2", "-shared"]


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any((CSRC / f).exists() and (CSRC / f).stat().st_mtime > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False, defines: tuple = (), out: Path | None = None) -> Path:
    """`defines`/`out` build an experimental variant next to the product library (never loaded by default)."""
    target = out or LIB
    if not force and not defines and not needs_build():
        return LIB
    srcs = [str(CSRC / s) for s in SOURCES if (CSRC / s).exists()]
    cmd = ["nvcc", *NVCC_FLAGS, *[f"-D{d}" for d in defines], *(["-Xptxas", "-v"] if verbose else []), "-o", str(target), *srcs]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libl3b200.so")
    if verbose:
        print(res.stderr)
    return target


if __name__ == "__main__":
    build(force=True, verbose=True)
    print(LIB)
