#!/usr/bin/env python3
"""Per-stage breakdown of the granule kernel from an `ncu --set full --import-source on` capture:
    python tools/ncu_by_stage.py gpurun_out/r02_kernels.ncu-rep [out.txt]
For every SASS instruction of l3_granule_kernel the source page gives warp-instructions executed and stall samples.  The
arithmetic helpers are inlined, so their instructions carry the helper's line, not the stage's: a stage is therefore taken to
be a contiguous ADDRESS range (ptxas does not move code across the __syncwarp between two stages), found from the instructions
that do carry a line of the kernel body; everything between two such anchors belongs to the earlier one's stage."""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:l3_granule", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
MARKERS = [  # (first code line of the stage in l3_kernels.cu -- only lines that carry instructions appear in the page --, stage name)
    ("const uint32_t phase = k & 1;", "loop head, wait for the staged inputs, descriptor bits"),
    ("const int nsf0 =", "band gains (minimp3.d:714-719)"),
    ("const int nch0 = *reinterpret_cast", "requantisation + MS stereo"),
    ("fence_proxy_async();", "TMA issue for the next granule (one lane) + stereo mode tests"),
    ("const int nlb0 = kind0 == 2", "load + alias reduction + IMDCT-36 + store"),
    ("if (mode >= 1 && lane < NS) {", "DCT-32 (18 of 32 lanes)"),
    ("const int dlo = max(0,", "window, samples 1..15 and 17..31 of every slot"),
    ("for (int k = 0; k < 15; k++) z[k] = col[", "window, samples 0 and 16 (18 of 32 lanes)"),
    ("const T first = D[NS * kDStride];", "history slide"),
]
ins = []      # (address, file, line, text, samples, warp instructions)
marks = {}    # stage index -> first line
cur_file, cur_line, hdr = None, None, None
for row in csv.reader(io.StringIO(raw)):
    if not row:
        continue
    if row[0] == "File Path":
        cur_file = row[1].rsplit("/", 1)[-1]
        continue
    if row[0] == "Line No":
        hdr = row
        i_addr, i_samp, i_exec = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
        i_sass = i_addr + 1
        continue
    if hdr is None or len(row) <= i_exec:
        continue
    if row[0]:
        cur_line = int(row[0])
        if cur_file == "l3_kernels.cu":
            for k, (text, _) in enumerate(MARKERS):
                if row[1].strip().startswith(text) and k not in marks:
                    marks[k] = cur_line
        continue
    if row[i_addr].startswith("0x"):
        ins.append((int(row[i_addr], 16), cur_file, cur_line, row[i_sass].strip(), int(row[i_samp] or 0), int(row[i_exec] or 0)))
assert len(marks) == len(MARKERS), ("marker not found", sorted(set(range(len(MARKERS))) - set(marks)))
first_body = marks[0]
# an inlined instruction is listed under every frame of its inline stack: keep one record per address, the one that carries
# a line of the kernel body if there is one
by_addr = {}
for rec in ins:
    old = by_addr.get(rec[0])
    if old is None or (rec[1] == "l3_kernels.cu" and rec[2] >= first_body and not (old[1] == "l3_kernels.cu" and old[2] >= first_body)):
        by_addr[rec[0]] = rec
ins = sorted(by_addr.values())


def stage_of_line(f, ln):
    if f != "l3_kernels.cu" or ln < first_body:
        return None
    s = None
    for k in sorted(marks, key=lambda k: marks[k]):
        if marks[k] <= ln:
            s = k
    return s


granules = None
stage, per = None, {}
for a, f, ln, text, samp, ex in ins:
    s = stage_of_line(f, ln)
    if s is not None:
        stage = s
    if stage is None:
        key = -1
    else:
        key = stage
    d = per.setdefault(key, {"exec": 0, "fp": 0, "mio": 0, "samp": 0, "static": 0})
    op = re.sub(r"^@!?U?P\d\s+", "", text).split()[0]
    d["exec"] += ex
    d["samp"] += samp
    d["static"] += 1 if ex else 0
    if re.match(r"(FFMA2|FMUL2|FADD2|FADD|FMUL|FFMA)\b", op):
        d["fp"] += ex
    if re.match(r"(LDS|STS|SHFL|LDG|STG|ST|LD|UBLKCP|LDC)\b", op.split(".")[0]):
        d["mio"] += ex
# executions of the loop head's first body instruction = granules (halo included)
granules = max(ex for a, f, ln, text, samp, ex in ins if f == "l3_kernels.cu" and ln == marks[0]) or 1
tot_s = sum(d["samp"] for d in per.values())
print(f"# {rep}: l3_granule_kernel, per granule of one warp (stereo: both channels); {granules} granules incl. halo", file=out)
print(f"# {'stage':62s} {'instr':>7s} {'FP':>6s} {'ld/st':>6s} {'executed-from':>13s} {'stall samples':>13s}", file=out)
names = {-1: "prologue / tile set-up (per tile, amortised)"}
names.update({k: n for k, (_, n) in enumerate(MARKERS)})
tot = {"exec": 0, "fp": 0, "mio": 0, "static": 0}
for k in sorted(per):
    d = per[k]
    print(f"  {names[k]:62s} {d['exec'] / granules:7.1f} {d['fp'] / granules:6.1f} {d['mio'] / granules:6.1f} {d['static']:13d} {100 * d['samp'] / tot_s:12.1f}%", file=out)
    for q in tot:
        tot[q] += d[q]
print(f"  {'total':62s} {tot['exec'] / granules:7.1f} {tot['fp'] / granules:6.1f} {tot['mio'] / granules:6.1f} {tot['static']:13d}", file=out)
