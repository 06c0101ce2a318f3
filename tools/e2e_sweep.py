"""End-to-end pipeline sweep (lanes x wave size x output format) and the concurrent device->host copy ceiling, one rank per GPU.

    python tools/e2e_sweep.py                                              # one GPU
    python -m torch.distributed.run --nproc-per-node N tools/e2e_sweep.py   # N GPUs, all ranks at the same time

Every rank decodes its own 1,024 x 60 s config-2 streams (bench.py's workload) through l3b_pipeline_decode; rank 0 prints one
JSON line per setting: aggregate audio-s/s (slowest rank), device->host GB/s per rank, and where the pipeline's threads spent
their time.  The ceiling is a 1 GiB device->pinned copy issued by every rank between two barriers.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=1024)
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--lanes", default="3,4,6")
    ap.add_argument("--waves", default="16,32")
    ap.add_argument("--formats", default="s16,f32")
    args = ap.parse_args()
    import torch
    import audio_formats_b200 as af
    import bench
    from audio_formats_b200 import api

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    join = bench.Join(world)
    threads = max(1, bench.host_threads() // max(1, world))
    streams = bench.gen_streams(range(rank * args.streams, (rank + 1) * args.streams), args.seconds, threads)
    datas = [s.data for s in streams]
    audio = join.sum(sum(s.seconds for s in streams))
    elems = sum(s.pcm_frames * s.params.nch for s in streams)
    pin = api.PinnedBuffer(4 * (elems + 16 * len(datas) + 4096), near_device=local)
    peak = bench.d2h_rate_gbs(torch, join, pin.view(np.float32))
    peak_1g = bench.d2h_rate_gbs(torch, join)
    peak_mixed = bench.d2h_rate_gbs(torch, join, pin.view(np.float32), h2d_fraction=0.097)   # the 16-bit pipeline's upload : download byte ratio
    if rank == 0:
        print(json.dumps({"n_gpus": world, "host_cpus": bench.host_threads(), "scan_threads_per_rank": threads,
                          "d2h_concurrent_peak_gbs_per_rank": peak, "d2h_concurrent_peak_gbs_per_rank_1GiB_repeated": peak_1g, "d2h_concurrent_peak_gbs_per_rank_with_9.7pct_uploads": peak_mixed, "d2h_concurrent_peak_gbs_aggregate": None if peak is None else peak * world}), flush=True)
    for fmt in args.formats.split(","):
        s16 = fmt == "s16"
        out = pin.view(np.int16) if s16 else pin.view(np.float32)
        for lanes in [int(x) for x in args.lanes.split(",")]:
            for wave in [int(x) for x in args.waves.split(",")]:
                pipe = af.BatchPipeline(device=local, lanes=lanes, wave_streams=wave, prepass_threads=threads, s16=s16)
                pipe.decode_into(datas, out)
                pipe.profile()
                join.barrier()
                t0 = time.perf_counter()
                for _ in range(args.steps):
                    pipe.decode_into(datas, out)
                join.barrier()
                sec = join.max((time.perf_counter() - t0) / args.steps)
                prof = {k: round(v / args.steps, 4) for k, v in pipe.profile().items()}
                pipe.close()
                if rank == 0:
                    gbs = elems * (2 if s16 else 4) / sec / 1e9
                    print(json.dumps({"format": fmt, "lanes": lanes, "wave_streams": wave, "audio_s_per_s": audio / sec, "ms_per_step": sec * 1e3,
                                      "d2h_gbs_per_rank": gbs, "frac_of_concurrent_peak": None if not peak else gbs / peak,
                                      "thread_seconds_per_step_rank0": prof}), flush=True)
    pin.free()
    join.close()


if __name__ == "__main__":
    main()
