"""Opcode histogram of every kernel in the product library (cuobjdump -sass), written to profiles/<tag>_sass_opcodes.txt.
Evidence for: packed FP32 (FMUL2 / FFMA2) in the granule kernel, no FADD2 fed by a packed multiply in the bit-exact variants
(ptxas would contract that pair into FFMA2), TMA bulk copies (UBLKCP), no tensor-core instructions (UTC*MMA / LDTM / HMMA)."""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
lib = ROOT / "audio_formats_b200" / "libl3b200.so"
txt = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", txt)), capture_output=True, text=True).stdout.split("\n")
out = [f"# cuobjdump -sass {lib.name}: static opcode counts per kernel (sm_100a)"]
watch = ["FMUL", "FMUL2", "FADD", "FADD2", "FFMA", "FFMA2", "MOV", "LDS", "LDS.64", "LDS.128", "STS.64", "STS.128", "STG.E.64", "STG.E", "LDG.E",
         "UBLKCP", "LDGSTS.E", "SYNCS.ARRIVE.TRANS64", "SHFL.UP", "SHFL.DOWN", "BAR.SYNC.DEFER_BLOCKING", "LDL", "STL"]
for k, f in enumerate(re.split(r"\n\s+Function : ", txt)[1:]):
    ops = collections.Counter()
    for line in f.split("\n"):
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(@!?U?P\w+\s+)?(\S+?)[\s;]", line)
        if m:
            ops[m.group(2)] += 1
    tensor = sum(n for o, n in ops.items() if o.startswith("UTC") or o in ("LDTM", "STTM") or "MMA" in o)
    tma = sum(n for o, n in ops.items() if o.startswith("UBLKCP") or o.startswith("UTMA"))
    out.append(f"\n## {names[k]}\ninstructions {sum(ops.values())}; tensor-core instructions {tensor}; TMA bulk copies (UBLKCP*/UTMA*) {tma}")
    out.append("  ".join(f"{o} {ops[o]}" for o in watch if ops[o]))
    out.append("top: " + ", ".join(f"{o} {n}" for o, n in ops.most_common(12)))
(ROOT / "profiles" / f"{tag}_sass_opcodes.txt").write_text("\n".join(out) + "\n")
print("\n".join(out)[:3000])
