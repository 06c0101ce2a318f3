#!/usr/bin/env python3
"""A small but varied decode for compute-sanitizer (memcheck / racecheck), checked against the oracle:
    compute-sanitizer --tool racecheck python tools/sanitize_probe.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import audio_formats_b200 as af
import oracle
from audio_formats_b200 import synth

S = synth.SynthParams
cases = [S(seed=1, nframes=20), S(seed=2, nch=1, bitrate_kbps=64, nframes=20, mode_ext_any=1),
         S(seed=3, hz=22050, bitrate_kbps=64, nframes=24, block_mode=1, stereo_mode=2, small_scalefactors=0),
         S(seed=4, hz=16000, nch=1, bitrate_kbps=32, nframes=24, block_mode=2),
         S(seed=5, nframes=24, block_mode=1, stereo_mode=2, istereo_untied=1, scfsi=1, private_bits=1, reservoir=2),
         S(seed=6, bitrate_kbps=200, nframes=20, free_format=1, block_mode=1),
         S(seed=7, bitrate_kbps=160, nframes=30, vbr=1, block_mode=1, stereo_mode=1, table_cycle=1, small_scalefactors=0),
         S(seed=8, hz=8000, bitrate_kbps=24, nframes=24, stereo_mode=1), S(seed=9, hz=48000, bitrate_kbps=320, nframes=16, level=12.0)]
streams = [synth.generate(p) for p in cases]
ctx = af.Context(0)
outs = ctx.decode([s.data for s in streams])
for p, s, o in zip(cases, streams, outs):
    ref, _ = oracle.decode_all(s.data)
    assert o.shape == ref.shape and np.array_equal(o.view(np.uint32), ref.view(np.uint32)), p
st = af.AudioStream(ctx).openFromMemory(streams[4].data)
st.seekPosition(3000)
assert len(st.readSamplesFloat(5000)) == 5000
st.close()
ctx.close()
print("sanitize probe ok:", len(cases), "streams")
