"""GPU-side debugging aid: find the first granules whose PCM differs from the oracle and classify them."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import audio_formats_b200 as af
import oracle
from audio_formats_b200 import synth
from dataclasses import replace

def classify(p, label):
    st = synth.generate(p)
    sc = af.Scan(st.data)
    ctx = af.Context(0)
    (pcm,), is_, iscf, ist = af.decode_batch_with_taps(ctx, [sc])
    ref, taps = oracle.decode_all(st.data, taps=sc.granules)
    nch = sc.channels
    d = sc.descs.reshape(-1, nch)
    bad = (pcm.view(np.uint32) != ref.view(np.uint32)).reshape(-1, 576, nch).any(axis=1)  # [granule, ch]
    bt = (d["w1"] >> 29) & 3
    mixed = d["w1"] >> 31
    hb = (d["w3"] >> 27) & 15
    print(label, "granules", sc.granules, "bad granule-channels", int(bad.sum()))
    stats = {}
    for g in range(sc.granules):
        key = (tuple(int(x) for x in bt[g]), tuple(int(x) for x in mixed[g]), int(hb[g, 0]))
        s = stats.setdefault(key, [0, 0])
        s[0] += 1
        s[1] += int(bad[g].any())
    for k, v in sorted(stats.items()):
        print("  bt", k[0], "mixed", k[1], "hdr", bin(k[2]), "count", v[0], "bad", v[1])
    first = np.argwhere(bad.any(axis=1))[:6].ravel()
    for g in first:
        dd = np.abs(pcm.reshape(-1, 576, nch)[g].astype(np.float64) - ref.reshape(-1, 576, nch)[g]).max(axis=0)
        print("  first bad granule", g, "bt", bt[g], "mixed", mixed[g], "hb", bin(int(hb[g, 0])), "maxdelta per ch", dd,
              "prev bt", bt[g - 1] if g else None)
    ctx.close()

if __name__ == "__main__":
    base = synth.config3_params(0, 6.0)
    classify(replace(base, block_mode=0, stereo_mode=0), "long/plain")
    classify(replace(base, block_mode=0, stereo_mode=1), "long/ms")
    classify(replace(base, block_mode=0, stereo_mode=2), "long/ms+is")
    classify(replace(base, block_mode=1, stereo_mode=0), "blocks/plain")
    classify(base, "config3")
