#!/usr/bin/env python3
"""Generate the golden fixtures of tests/golden/ (run from the repo root: python tests/golden/make_golden.py).

The reference ships NO golden vectors for MP3 and cannot be run here (D, no compiler), so these
fixtures are produced by the C oracle (oracle/, a restatement of minimp3.d) on streams from the
deterministic generator.  They pin the oracle, the generator and the host prepass against regressions;
they do not pin the oracle against the D reference ("parity unpinned", see oracle/l3_oracle.h).

Each fixture NAME.npz holds: mp3 (uint8 stream), pcm (float32 [frames, ch]), quantised (int16 [gr, ch, 576],
what the generator encoded), iscf (uint8 [gr, ch, 40]) and the parameter dict as JSON.
"""
import json
import sys
from dataclasses import asdict, replace
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from audio_formats_b200 import synth  # noqa: E402
import oracle  # noqa: E402

HERE = Path(__file__).resolve().parent


def fixtures():
    S = synth.SynthParams
    yield "m1_44k_stereo_long", S(seed=101, nframes=14, bitrate_kbps=128, reservoir=1)
    yield "m1_44k_js_blocks", replace(synth.config3_params(102, 1.0), nframes=14)
    yield "m1_48k_mono_320", S(seed=103, hz=48000, nch=1, bitrate_kbps=192, nframes=14, block_mode=1, scfsi=1,
                               small_scalefactors=0)
    yield "m2_22k_stereo_is", S(seed=104, hz=22050, nch=2, bitrate_kbps=64, nframes=16, block_mode=1, stereo_mode=2,
                                small_scalefactors=0)
    yield "m2_16k_mono", S(seed=105, hz=16000, nch=1, bitrate_kbps=32, nframes=16, block_mode=1, small_scalefactors=0)
    yield "m25_8k_stereo", S(seed=106, hz=8000, nch=2, bitrate_kbps=32, nframes=16, stereo_mode=1, small_scalefactors=0)
    yield "m1_32k_crc_id3", S(seed=107, hz=32000, nch=2, bitrate_kbps=96, nframes=14, block_mode=1, stereo_mode=2,
                              reservoir=2, crc=1, id3v2_bytes=300, id3v1=1, scfsi=1)


def main():
    for name, p in fixtures():
        st = synth.generate(p, want_quantised=True)
        pcm, taps = oracle.decode_all(st.data, taps=st.granules)
        assert len(taps) == st.granules
        assert np.array_equal(taps["is"][:, :p.nch], st.quantised)
        np.savez_compressed(HERE / f"{name}.npz", mp3=np.frombuffer(st.data, np.uint8), pcm=pcm,
                            quantised=st.quantised, iscf=taps["iscf"][:, :p.nch],
                            params=np.frombuffer(json.dumps(asdict(p)).encode(), np.uint8))
        print(f"{name}: {len(st.data)} bytes, pcm {pcm.shape}, rms {np.sqrt((pcm.astype(np.float64) ** 2).mean()):.4f}")


if __name__ == "__main__":
    main()
