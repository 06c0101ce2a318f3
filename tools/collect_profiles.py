#!/usr/bin/env python3
"""Copy the files tools/final_profile.sh left in gpurun_out/ into profiles/ and derive the summaries:
    python tools/collect_profiles.py r01b"""
import csv, json, shutil, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r01b"
for f in ("bench_ours.json", "bench_reference.json", "launches_bench.csv", "fp32_pipes.txt", "dct32_tc.txt"):
    shutil.copy(f"gpurun_out/{tag}_{f}", f"profiles/{tag}_{f}")
subprocess.run([sys.executable, "tools/summarize_ncu.py", f"gpurun_out/{tag}_kernels.ncu-rep", f"profiles/{tag}_kernels_ncu_summary.txt"],
               stdout=subprocess.DEVNULL, check=True)
raw = subprocess.run(["ncu", "-i", f"gpurun_out/{tag}_kernels.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr = rows[0]
names = {"l3_scf": "scf", "l3_huff_big": "huff_big", "l3_huff_c1": "huff_c1", "l3_granule": "granule"}
out = {}
for r in rows[2:]:
    k = r[hdr.index("Kernel Name")]
    key = [v for n, v in names.items() if n in k][0]
    def g(m):
        i = hdr.index(m)
        return float(r[i]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(rows[1][i], 1)
    rd, wr = g("dram__bytes_read.sum"), g("dram__bytes_write.sum")
    i = hdr.index("gpu__time_duration.sum")
    out[key] = {"kernel": k, "dram_bytes_read": rd, "dram_bytes_write": wr, "traffic": rd + wr,
                "duration_under_ncu": f"{r[i]} {rows[1][i]}",
                "workload": "bench.py default (config 2: 1024 x 60 s), one launch, ncu --set full --clock-control none"}
json.dump(out, open(f"profiles/{tag}_traffic.json", "w"), indent=1)
d = json.load(open(f"profiles/{tag}_bench_ours.json"))
r = d["roofline"]; e = d["e2e"]
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 2), "granule", round(r["ms_per_launch"], 2), "entropy", round(r["entropy_kernels_ms"], 2))
print("e2e", round(e["value"]), round(e["ms_per_step"], 1), "ms", round(e.get("d2h_achieved_gbs", 0), 1), "of", round(e.get("d2h_concurrent_peak_gbs") or 0, 1), "GB/s")
import os
if os.path.exists(f"gpurun_out/{tag}_granule_fused.ncu-rep"):
    subprocess.run([sys.executable, "tools/summarize_ncu.py", f"gpurun_out/{tag}_granule_fused.ncu-rep", f"profiles/{tag}_granule_fused_ncu_summary.txt"],
                   stdout=subprocess.DEVNULL, check=True)
alg = r["hbm"]["algorithmic_bytes_per_launch"]
tot = sum(v["traffic"] for v in out.values())
print("traffic / algorithmic bytes: granule", round(out["granule"]["traffic"] / alg, 3), "whole step", round(tot / alg, 3))
out["_summary"] = {"algorithmic_bytes_per_step": alg, "granule_traffic_over_algorithmic": out["granule"]["traffic"] / alg,
                   "step_traffic_over_algorithmic": tot / alg}
json.dump(out, open(f"profiles/{tag}_traffic.json", "w"), indent=1)
print("cpu", round(d["cpu_baseline"]["value"]), "ref arm", round(json.load(open(f"profiles/{tag}_bench_reference.json"))["value"]))
print({k: (round(v["traffic"] / 1e9, 2), v["duration_under_ncu"]) for k, v in out.items() if not k.startswith("_")})
