"""BASELINE.json configs[4]: file-sharded decode of 65,536 logical 180 s stereo 320 kbps streams (bench.py --workload config5).

472 GB of bitstream in and 4.16 TB of PCM out cannot exist, so: U unique payloads are resident in HBM, the
logical streams reuse them (logical stream j decodes payload j mod U into its own PCM rows), and the job runs as waves of W
logical streams through ONE resident batch whose PCM buffer is recycled from wave to wave.  With U | W every wave has the
same shape, so a wave is one l3b_batch_run.  The logical streams are assigned to the ranks by file, longest first on their
granule-channel count (audio_formats_b200.shard.shard_lpt); there is no collective on the data path: "strong" scaling.

Checked in the run: after the timed passes every wave is run once more and a seeded sample of its logical streams is read
back and compared with the oracle's decode of the payload (crc32 of the float PCM, bit-exact).
"""
from __future__ import annotations

import json
import sys
import time
import zlib
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent


def run(args, rank, world, local, join):
    import torch
    import audio_formats_b200 as af
    import bench
    import oracle
    from audio_formats_b200 import api, shard, synth

    bench.WORKLOAD = "config5"   # `bench` is a second copy of the module that runs as __main__
    U, W, logical = args.unique, args.wave, args.logical
    if W % U or logical % W:
        raise SystemExit("config5: the wave size must be a multiple of the unique payloads, the job a multiple of the wave")
    threads = max(1, bench.host_threads() // max(1, min(world, 8)))
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        payloads = list(ex.map(lambda s: synth.generate(synth.config5_params(s, args.seconds)), range(U)))
        scans = list(ex.map(lambda p: af.Scan(p.data), payloads))
    t_gen = time.perf_counter() - t0
    # file sharding: logical stream j costs the granule-channels of payload j mod U
    cost = [scans[j % U].granules * scans[j % U].channels for j in range(logical)]
    mine = shard.shard_lpt(cost, world)[rank]
    n_waves = -(-len(mine) // W)
    if len(mine) % W:
        raise SystemExit("config5: logical streams per rank must be a multiple of the wave size")
    audio_per_stream = [s.delivered_samples / s.channels / s.samplerate for s in scans]
    audio_rank = sum(audio_per_stream[j % U] for j in mine)

    # one wave: W logical streams over the U payloads (blob holds each payload once; descriptors per logical stream)
    ctx = af.Context(local)
    base = api.HostBatch(scans)                       # U streams: blob + descriptors + stream table
    reps = W // U
    hb = api.HostBatch([], want_taps=False)
    hb.blob = base.blob
    hb.descs = np.tile(base.descs, reps)
    sd = np.tile(base.streams, reps)
    n_desc_u = len(base.descs)
    pcm = 0
    for k in range(W):
        u = k % U
        pcm = (pcm + 3) & ~3
        sd[k]["first_grch"] = (k // U) * n_desc_u + int(base.streams[u]["first_grch"])
        sd[k]["pcm_off"] = pcm
        pcm += int(base.streams[u]["pcm_count"])
    hb.streams, hb.pcm_floats, hb.n_grch = sd, int(pcm), len(hb.descs)
    rb = ctx.upload(hb)
    ext = torch.cuda.ExternalStream(ctx.cuda_stream, device=torch.device("cuda", local))

    for _ in range(max(3, args.warmup)):
        rb.run()
    rb.sync()
    sampler = bench.ClockSampler(local)
    join.barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(ext)
    for _ in range(args.steps * n_waves):
        rb.run()
    ev1.record(ext)
    rb.sync()
    join.barrier()
    clocks = sampler.stop()
    step_ms = join.max(ev0.elapsed_time(ev1) / args.steps)
    nk = min(args.steps * n_waves, 64)
    kern, launches = rb.timing(nk)
    gran_ms, ent_ms = kern[1] / nk, kern[0] / nk
    total_audio = join.sum(audio_rank)
    value = total_audio / (step_ms * 1e-3)

    # ---- per-wave check against the oracle: seeded sample of logical streams, crc32 of their float PCM ----
    with ThreadPoolExecutor(threads) as ex:
        ref_crc = list(ex.map(lambda p: zlib.crc32(oracle.transcode_loop(p.data, 1024, keep=True)[3].tobytes()), payloads))
    rng = np.random.default_rng(5 + rank)
    checked = bad = 0
    for w in range(n_waves):
        rb.run()
        rb.sync()
        for k in rng.choice(W, 2, replace=False):
            got = rb.download(int(sd[k]["pcm_off"]), int(sd[k]["pcm_count"]))
            checked += 1
            bad += int(zlib.crc32(got.tobytes()) != ref_crc[int(k) % U])
    parity = {"waves": n_waves, "logical_streams_checked": checked, "mismatches": bad,
              "compare": "crc32 of the float PCM of 2 seeded logical streams per wave vs the oracle's decode of their payload"}

    # ---- roofline of the granule kernel on this workload ----
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    hbm_peak = float(peaks.get("hbm_gbs", bench.HBM_FALLBACK_GBS))
    n_grch_wave = hb.n_grch
    alg_bytes = hb.pcm_floats * 4 + sum(len(p.data) for p in payloads) * reps
    flops = n_grch_wave * bench.FLOPS_PER_GRCH
    tfl, gbs = flops / (gran_ms * 1e-3) / 1e12, alg_bytes / (gran_ms * 1e-3) / 1e9
    roofline = {"bound": "fp32", "kernel": "l3_granule_kernel<2,4,exact>", "achieved": tfl, "peak": bench.FP32_NOMINAL_TFLOPS,
                "unit": "TFLOP/s", "frac": tfl / bench.FP32_NOMINAL_TFLOPS, "peak_source": "nominal FP32", "traffic": None,
                "ms_per_launch": gran_ms, "entropy_kernels_ms": ent_ms, "launch": "one wave",
                "hbm": {"achieved_gbs": gbs, "peak_gbs": hbm_peak, "frac": gbs / hbm_peak, "algorithmic_bytes_per_launch": alg_bytes}}
    rb.free()

    # ---- end to end on a bounded sample: a quarter wave of logical streams through the library's pipeline, 16-bit delivery ----
    e2e = None
    if not args.no_e2e:
        n_e2e = max(U, W // 4)
        datas = [payloads[k % U].data for k in range(n_e2e)]
        elems = sum(int(scans[k % U].delivered_samples) for k in range(n_e2e))
        pin = api.PinnedBuffer(2 * (elems + 8 * (n_e2e // args.e2e_wave + 2) + 1024), near_device=local)
        out = pin.view(np.int16)
        pipe = af.BatchPipeline(device=local, lanes=args.e2e_lanes, wave_streams=min(args.e2e_wave, 8), prepass_threads=threads, s16=True)
        pipe.decode_into(datas, out)
        join.barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            info = pipe.decode_into(datas, out)
        join.barrier()
        sec = join.max((time.perf_counter() - t0) / args.e2e_steps)
        ref16 = bench.q16(oracle.transcode_loop(payloads[0].data, 1024, keep=True)[3].reshape(-1))
        o, fr, ch, _hz = info[0]
        parity["e2e_s16_first_stream_matches_oracle"] = bool(np.array_equal(out[o:o + fr * ch], ref16))
        audio_e2e = join.sum(sum(audio_per_stream[k % U] for k in range(n_e2e)))
        e2e = {"value": audio_e2e / sec, "unit": bench.UNIT, "h2d_bytes_per_step": sum(len(d) for d in datas),
               "d2h_bytes_per_step": elems * 2, "ms_per_step": sec * 1e3, "steps": args.e2e_steps,
               "sample": f"{n_e2e} logical streams per GPU per step (the full job would move {logical * elems // n_e2e * 2 / 1e12:.2f} TB to the host)",
               "output": "int16 PCM in pinned host memory", "d2h_achieved_gbs": elems * 2 / sec / 1e9}
        pipe.close()
        pin.free()
    ctx.close()

    if rank == 0:
        line = {"metric": bench.METRIC, "value": value, "unit": bench.UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
                "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": bench.workload_config(args), "audio_seconds_per_step": total_audio,
                "logical_streams_per_rank": len(mine), "waves_per_rank_per_step": n_waves, "unique_payloads": U,
                "clocks": clocks, "gpu_launches": (launches // nk) * args.steps * n_waves,
                "roofline": roofline, "e2e": e2e, "cpu_baseline": None, "parity": parity,
                "setup": {"generate_s": t_gen, "host_threads": threads, "sharding": "shard_lpt by granule-channels, no collective"}}
        print(json.dumps(line), flush=True)
