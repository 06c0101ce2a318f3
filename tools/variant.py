"""Build an experimental variant of the library (A/B builds) and print what ptxas made of the stereo bit-exact granule kernel.
usage: python tools/variant.py <name> [DEFINE ...]   ->  audio_formats_b200/_variants/<name>.so   (run with L3B_LIB=<path>, see tools/ab.sh)"""
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from audio_formats_b200 import build  # noqa: E402

name, defines = sys.argv[1], tuple(sys.argv[2:])
out = ROOT / "audio_formats_b200" / "_variants" / f"{name}.so"
out.parent.mkdir(exist_ok=True)
srcs = [str(build.CSRC / s) for s in build.SOURCES]
cmd = ["nvcc", *build.NVCC_FLAGS, *[f"-D{d}" for d in defines], "-Xptxas", "-v", "-o", str(out), *srcs]
res = subprocess.run(cmd, capture_output=True, text=True)
if res.returncode:
    sys.exit(res.stdout + res.stderr)
KERN = "l3_granule_kernelILi2ELi4ELb0ELb0ELb0ELb0"
lines = res.stderr.split("\n")
for i, l in enumerate(lines):
    if "Function properties" in l and KERN in l:
        print(name, "|", lines[i + 1].strip(), "|", lines[i + 2].strip())
sass = subprocess.run(["cuobjdump", "-sass", "-fun", "_ZN3l3b17l3_granule_kernelILi2ELi4ELb0ELb0ELb0ELb0EEEvNS_11BatchParamsEPKNS_4TileEj", str(out)],
                      capture_output=True, text=True).stdout
n = len(re.findall(r"^\s+/\*[0-9a-f]{4,6}\*/\s+\S", sass, re.M))
print(name, "| static SASS instructions", n, "=", n * 16 // 1024, "KB")
