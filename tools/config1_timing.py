import sys, time
sys.path.insert(0, '.')
import numpy as np
import audio_formats_b200 as af
import oracle
from audio_formats_b200 import synth
ctx = af.Context(0)
st1 = synth.generate(synth.config1_params(1))
def run(read=1024):
    s = af.AudioStream(ctx).openFromMemory(st1.data)
    n = 0
    t_open = time.perf_counter()
    while True:
        c = s.readSamplesFloat(read)
        if len(c) == 0: break
        n += len(c)
    s.close()
    return n
run()
for read in (1024, 8192, 1 << 20):
    ts = []
    for _ in range(8):
        t0 = time.perf_counter(); run(read); ts.append((time.perf_counter() - t0) * 1e3)
    print('read', read, 'ms', ' '.join(f'{t:.2f}' for t in ts))
t0 = time.perf_counter()
for _ in range(5): oracle.transcode_loop(st1.data, 1024, keep=False)
print('cpu ms', (time.perf_counter() - t0) / 5 * 1e3)
# breakdown: open only
ts = []
for _ in range(5):
    t0 = time.perf_counter(); s = af.AudioStream(ctx).openFromMemory(st1.data); t1 = time.perf_counter(); c = s.readSamplesFloat(1024); t2 = time.perf_counter(); s.close()
    ts.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3))
print('open ms, first read ms', ts)
