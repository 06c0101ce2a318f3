"""Print registers / spills of every kernel in the product build (ptxas -v), flagging any spill: local memory misses L1 in the
granule kernel (the shared-memory carve-out is at its maximum), so a spill inside its loop costs an L2 round trip."""
import re, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from audio_formats_b200 import build
srcs = [str(build.CSRC / s) for s in build.SOURCES]
defs = [f"-D{d}" for d in sys.argv[1:]]
res = subprocess.run(["nvcc", *build.NVCC_FLAGS, *defs, "-Xptxas", "-v", "-o", "/tmp/_spills.so", *srcs], capture_output=True, text=True)
lines = res.stderr.split("\n")
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function properties for (\S+)", res.stderr)), capture_output=True, text=True).stdout.split("\n")
k = 0
for i, l in enumerate(lines):
    if "Function properties for" in l:
        sp = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", lines[i + 1])
        rg = re.search(r"Used (\d+) registers", lines[i + 2])
        flag = "  <-- SPILLS" if sp and (int(sp.group(1)) or int(sp.group(2))) else ""
        print(f"{rg.group(1) if rg else '?':>4} regs  spill {sp.group(1)}/{sp.group(2)}  {names[k][:110]}{flag}")
        k += 1
