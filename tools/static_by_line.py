"""Static SASS instruction count per source line of one kernel (nvdisasm -g on the cubin inside a built library): where the code
bytes of the granule kernel's loop come from.  usage: python tools/static_by_line.py <lib.so> [lo-hi ...]   (line ranges to sum)"""
import collections, re, subprocess, sys, tempfile, os
lib = os.path.abspath(sys.argv[1])
KERN = "_ZN3l3b17l3_granule_kernelILi2ELi4ELb0ELb0ELb0ELb0EEEvNS_11BatchParamsEPKNS_4TileEj"
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, capture_output=True)
    cub = [f for f in os.listdir(d) if f.startswith("l3_kernels.") and f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
sec = txt.split(".text." + KERN + ":")[1].split("\n.text.")[0]
cnt = collections.Counter(); cur = None
for l in sec.split("\n"):
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4}\*/", l) and cur:
        cnt[cur] += 1
tot = sum(cnt.values())
print("total", tot)
ranges = [tuple(map(int, a.split("-"))) for a in sys.argv[2:]]
if ranges:
    for lo, hi in ranges:
        print(f"l3_kernels.cu {lo}-{hi}:", sum(v for (f, n), v in cnt.items() if f == "l3_kernels.cu" and lo <= n <= hi))
    print("other files:", sum(v for (f, n), v in cnt.items() if f != "l3_kernels.cu"))
else:
    for (f, n), v in sorted(cnt.items()):
        if v >= 8: print(f, n, v)
