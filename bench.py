#!/usr/bin/env python
"""bench.py -- MP3 (MPEG-1 Layer III) batch decode throughput, audio-seconds decoded per second.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: the CPU decoder on host cores

A "step" is one pass of the hot path over one batch of synthetic streams.  At N=1 the workload is
BASELINE.json configs[1]: 1,024 synthetic 60 s 44.1 kHz stereo 128 kbps MPEG-1 Layer III long-block
streams (seeds 0..1023).  N>1: streams shard by file, 1,024 streams per GPU, no collective (weak scaling).
`--workload config3|config4|config5` runs the other BASELINE configurations at their stated sizes (config5: 65,536
logical 180 s 320 kbps streams over U unique payloads, in device-resident waves, sharded over the ranks: strong scaling).

`value`    device-resident throughput of the bit-exact path: bitstreams + descriptors already in HBM; one step =
           scalefactor kernel + big_values kernel + count1 kernel (integer, "entropy") + fused granule kernel (float).
`e2e`      MP3 bytes in host memory -> 16-bit PCM in pinned host memory through the library's batch pipeline
           (l3b_pipeline_decode): host prepass + H2D + kernels + D2H, all inside the timed region.  `e2e.f32` is the same
           with float delivery.  The reference arm converts to 16 bit too (the transcode example writes 16-bit WAV).
`roofline` of the dominant (granule) kernel against the SLOWER of the FP32 and HBM roofs, bit-exact mode, plus the
           tolerance-mode (FMA-contracted) kernel and its measured deviation; `cpu_baseline`: the C oracle (port of the D
           reference, which cannot be built here: no D compiler) on the host cores, timed in the same run;
`parity`   the measured batch checked in this run: a seeded sample of streams bit-exact against the oracle, and every
           stream's checksum against a second decode (the pipeline's, which tiles and batches them differently).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
import zlib
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

METRIC = "mp3_batch_decode_audio_seconds_per_second"
UNIT = "audio-s/s"
FLOPS_PER_GRCH = 32734          # SURVEY.md 8d / BASELINE.md 3: float add/sub/mul per granule-channel of the reference
FP32_NOMINAL_TFLOPS = 74.45     # 148 SM x 128 lanes x 2 x 1.965 GHz (not measured by the driver)
# tools/microbench/fp32_pipes.cu on this pool's B200: FMUL / FADD / FADD2 / FMUL2 all retire ~125 lane-ops per clock per SM
# (3.8 scalar or 1.95 packed warp-instructions): the FP32 pipe does 148 x 125 x 1.965e9 = 36.4e12 un-fused flop/s
FP32_UNFUSED_MEASURED_TFLOPS = 36.4
HBM_FALLBACK_GBS = 6650.0       # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent
TRAFFIC_FILE = "r02_traffic.json"


def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


WORKLOAD = "config2"   # set from --workload; config2 is the BASELINE.json metric configuration


def gen_streams(seeds, seconds, threads):
    from audio_formats_b200 import synth
    synth.build()
    maker = {"config2": synth.config2_params, "config3": synth.config3_params, "config4": synth.config4_params,
             "config5": synth.config5_params}[WORKLOAD]

    def one(seed):
        return synth.generate(maker(seed, seconds))

    with ThreadPoolExecutor(threads) as ex:
        return list(ex.map(one, seeds))


def q16(x):
    return np.clip(np.rint(x.astype(np.float64) * 32768.0), -32768, 32767).astype(np.int16)


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def cpu_decode_throughput(streams, threads, s16=True):
    """Decode `streams` with the oracle's transcode loop (1,024-frame reads; every chunk converted to 16 bit when s16)
    on `threads` host threads; returns (audio_s, wall_s)."""
    import oracle
    oracle.build()
    oracle.lib()

    def one(st):
        n, nch, hz, _ = oracle.transcode_loop(st.data, 1024, keep=False, s16=s16)
        return n / hz

    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        audio = sum(ex.map(one, streams))
    return audio, time.perf_counter() - t0


def cpu_baseline_note():
    return "C restatement of the reference's D decoder (no D compiler in this image), transcode loop with 1,024-frame reads, " \
           "each chunk converted to 16 bit like the reference's WAV writer (un-dithered)"


def run_reference(args, rank, world):
    """Reference arm: the reference's CPU implementation of the path on the host cores.  The D decoder cannot
    be built in this image (no dmd/ldc2/gdc), so this runs the C port (oracle/), all host threads."""
    if rank != 0:
        return
    threads = host_threads()
    probe = gen_streams(range(min(threads, 8)), args.seconds, threads)
    a, w = cpu_decode_throughput(probe, threads)
    per_stream_s = w * min(threads, len(probe)) / max(1, len(probe))  # thread-seconds per stream
    n_sample = int(max(threads, min(args.streams, args.ref_step_seconds * threads / max(per_stream_s, 1e-6))))
    n_sample = (n_sample // threads) * threads or threads
    streams = gen_streams(range(n_sample), args.seconds, threads)
    for _ in range(args.warmup):
        cpu_decode_throughput(streams, threads)
    audio = wall = 0.0
    for _ in range(args.steps):
        a, w = cpu_decode_throughput(streams, threads)
        audio += a; wall += w
    value = audio / wall
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": workload_scaling(), "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": workload_config(args),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{n_sample} streams x {args.seconds:g} s per step (a bounded sample of the same seeded streams)",
                             "note": cpu_baseline_note()},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_scaling():
    return "strong" if WORKLOAD == "config5" else "weak"


def workload_config(args):
    """Identical in both arms (the driver compares the dicts)."""
    n, sec = args.streams, args.seconds
    names = {
        "config2": f"BASELINE.json configs[1]: {n} synthetic {sec:g} s 44.1 kHz stereo 128 kbps MPEG-1 Layer III long-block "
                   f"streams per GPU (seeds rank*streams+i)",
        "config3": f"BASELINE.json configs[2]: {n} synthetic {sec:g} s 44.1 kHz 128 kbps streams per GPU with mixed long/short/mixed "
                   f"blocks, joint (MS + intensity) stereo, heavy bit-reservoir use",
        "config4": f"BASELINE.json configs[3]: heterogeneous batch of {n} streams x {sec:g} s per GPU mixing 32/44.1/48 kHz MPEG-1 and "
                   f"16/22.05/24 kHz MPEG-2 LSF, 64-320 kbps, mono and stereo",
        "config5": f"BASELINE.json configs[4]: {args.logical} logical 44.1 kHz stereo 320 kbps streams x {sec:g} s over "
                   f"{args.unique} unique payloads, file-sharded over the GPUs, device-resident waves of {args.wave} streams",
    }
    return {"workload": names[WORKLOAD], "streams_per_gpu": n, "seconds_per_stream": sec,
            "sharding": "by file, no collective", "l2": "inputs larger than L2 (no flush needed)"}


# ------------------------------------------------------------------------------------------------
class Join:
    """Barrier + max/sum over ranks.  The data path has no collective (streams are independent), so the only thing the
    ranks exchange is timing scalars: a gloo group on the host does it; NCCL is not used anywhere."""

    def __init__(self, world):
        self.dist = None
        if world > 1:
            import torch.distributed as dist
            dist.init_process_group("gloo")
            self.dist = dist

    def barrier(self):
        import torch
        torch.cuda.synchronize()
        if self.dist:
            self.dist.barrier()

    def _red(self, x, op):
        if not self.dist:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        return self._red(x, self.dist.ReduceOp.MAX) if self.dist else x

    def sum(self, x):
        return self._red(x, self.dist.ReduceOp.SUM) if self.dist else x

    def close(self):
        if self.dist:
            self.dist.destroy_process_group()


def d2h_rate_gbs(torch, join, dst=None, reps=2, h2d_fraction=0.0):
    """Device -> pinned host copy rate of this box, all ranks AT THE SAME TIME between two barriers (per-rank rate of the
    slowest rank): the ceiling of any end-to-end number that delivers PCM to the host.  `dst` (a pinned numpy array, the
    pipeline's own output buffer) makes it a copy of up to 8 GiB into DISTINCT host pages, 1 GiB at a time, like the
    pipeline's deliveries; a 1 GiB copy repeated into one buffer partly lands in the CPU's last-level cache and reads high.
    h2d_fraction > 0: a host -> device copy of that fraction of the bytes runs on a second stream at the same time, like the
    pipeline's uploads of the bitstreams (the two directions share the host's side of the link)."""
    try:
        chunk = 1 << 28   # floats: 1 GiB
        src = torch.empty(chunk, dtype=torch.float32, device="cuda")
        if dst is None:
            host = torch.empty(chunk, dtype=torch.float32).pin_memory()
            n_chunks = 1
        else:
            host = torch.from_numpy(dst.view(np.float32))
            n_chunks = max(1, min(8, host.numel() // chunk))
        up_n = int(chunk * h2d_fraction)
        up_host = torch.empty(max(1, up_n), dtype=torch.float32).pin_memory() if up_n else None
        up_dev = torch.empty(max(1, up_n), dtype=torch.float32, device="cuda") if up_n else None
        side = torch.cuda.Stream() if up_n else None
        best = 0.0
        for _ in range(reps):
            join.barrier()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for k in range(n_chunks):
                if up_n:
                    with torch.cuda.stream(side):
                        up_dev.copy_(up_host, non_blocking=True)
                host[k * chunk:(k + 1) * chunk].copy_(src, non_blocking=True)
            c1.record()
            torch.cuda.synchronize()
            ms = join.max(c0.elapsed_time(c1))     # the slowest rank of a concurrent round
            best = max(best, (n_chunks * chunk * 4) / (ms * 1e-3) / 1e9)
        del src, host
        return best
    except RuntimeError:   # a side measurement must not take the bench down
        return None


def stream_crcs(buf: np.ndarray, spans, threads):
    """crc32 of every stream's PCM bytes in `buf` (spans: (element offset, element count))."""
    def one(sp):
        o, n = sp
        return zlib.crc32(buf[o:o + n].data)

    with ThreadPoolExecutor(threads) as ex:
        return list(ex.map(one, spans))


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=None, help="streams per GPU (default: the configuration's size)")
    ap.add_argument("--seconds", type=float, default=None, help="seconds per stream (default: the configuration's)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-lanes", type=int, default=4)
    ap.add_argument("--e2e-wave", type=int, default=16, help="streams per pipeline wave")
    ap.add_argument("--ref-step-seconds", type=float, default=6.0)
    ap.add_argument("--cpu-baseline-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-fused", action="store_true", help="skip the tolerance-mode kernel measurement")
    ap.add_argument("--parity-sample", type=int, default=16, help="streams compared bit-exactly with the oracle in this run")
    ap.add_argument("--workload", default="config2", choices=["config2", "config3", "config4", "config5"],
                    help="config2 is the metric configuration; config3/4/5 are the other BASELINE configurations at full size")
    ap.add_argument("--logical", type=int, default=65536, help="config5: logical streams of the whole job")
    ap.add_argument("--unique", type=int, default=64, help="config5: unique payloads (device-resident, shared by the logical streams)")
    ap.add_argument("--wave", type=int, default=1024, help="config5: logical streams per device-resident wave")
    args = ap.parse_args()
    global WORKLOAD
    WORKLOAD = args.workload
    defaults = {"config2": (1024, 60.0), "config3": (1024, 60.0), "config4": (4096, 30.0), "config5": (0, 180.0)}[WORKLOAD]
    if args.streams is None:
        args.streams = defaults[0]
    if args.seconds is None:
        args.seconds = defaults[1]
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if WORKLOAD == "config5":
            args.streams = args.unique
        run_reference(args, rank, world)
        return

    import torch
    import audio_formats_b200 as af
    from audio_formats_b200 import api
    from audio_formats_b200 import build as b
    if b.needs_build():
        b.build()
    if not torch.cuda.is_available() or af.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device (this path has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    join = Join(world)
    if WORKLOAD == "config5":
        sys.path.insert(0, str(ROOT / "tools"))
        import bench_config5
        bench_config5.run(args, rank, world, local, join)
        join.close()
        return

    # host threads of this rank: the box's CPUs are shared by the ranks
    threads = max(1, host_threads() // max(1, min(world, 8)))
    t0 = time.perf_counter()
    streams = gen_streams(range(rank * args.streams, (rank + 1) * args.streams), args.seconds, threads)
    t_gen = time.perf_counter() - t0
    datas = [s.data for s in streams]

    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        scans = list(ex.map(af.Scan, datas))
    t_scan = time.perf_counter() - t0
    audio_s = sum(s.delivered_samples / s.channels / s.samplerate for s in scans)
    n_grch = sum(s.granules * s.channels for s in scans)
    n_grch_stereo = sum(s.granules * s.channels for s in scans if s.channels == 2)

    ctx = af.Context(local)
    hb = api.HostBatch(scans)
    # inputs live in pinned host memory
    pin_in = api.PinnedBuffer(hb.blob.size + hb.descs.size * 16 + 128, near_device=local)
    pin_in.u8[:hb.blob.size] = hb.blob
    doff = (hb.blob.size + 63) & ~63
    pin_in.u8[doff:doff + hb.descs.size * 16] = hb.descs.view(np.uint8)
    hb.blob = pin_in.u8[:hb.blob.size]
    hb.descs = pin_in.u8[doff:doff + hb.descs.size * 16].view(api.GRCH_DTYPE)
    spans = [(int(sd["pcm_off"]), int(sd["pcm_count"])) for sd in hb.streams]
    rb = ctx.upload(hb)
    ext = torch.cuda.ExternalStream(ctx.cuda_stream, device=torch.device("cuda", local))

    def timed_steps(rbatch, steps, warmup):
        for _ in range(warmup):
            rbatch.run()
        rbatch.sync()
        join.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(ext)
        for _ in range(steps):
            rbatch.run()
        ev1.record(ext)
        rbatch.sync()
        join.barrier()
        step = join.max(ev0.elapsed_time(ev1) / steps)
        nk = min(steps, 64)
        kern, launches = rbatch.timing(nk)
        return step, [k / nk for k in kern], launches // nk

    # ---- device-resident timing, bit-exact path: W warm-up steps, then EXACTLY K timed steps ------------------------
    sampler = ClockSampler(local)
    sampler.start()
    step_ms, kern_ms, launches_per_step = timed_steps(rb, args.steps, args.warmup)
    clocks = sampler.stop()
    total_audio = join.sum(audio_s)
    value = total_audio / (step_ms * 1e-3)
    ent_ms, gran_ms = kern_ms[0], kern_ms[1]

    # the measured batch, kept on the host for the checks below (pinned: also the e2e destination)
    pin_pcm = api.PinnedBuffer(4 * (hb.pcm_floats + 8 * (len(scans) // args.e2e_wave + 2) + 1024), near_device=local)
    out_f32 = pin_pcm.view(np.float32)
    rb.download_into(out_f32.ctypes.data, 0, hb.pcm_floats)
    crc_resident = stream_crcs(out_f32, spans, threads)

    # ---- parity of the measured batch: a seeded sample of streams against the oracle, bit for bit -------------------
    import oracle
    oracle.build()
    rng = np.random.default_rng(20261017 + rank)
    sample = sorted(rng.choice(len(streams), min(args.parity_sample, len(streams)), replace=False).tolist())

    def oracle_pcm(i):
        return oracle.transcode_loop(streams[i].data, 1024, keep=True)[3]

    with ThreadPoolExecutor(threads) as ex:
        refs = dict(zip(sample, ex.map(oracle_pcm, sample)))
    mism = 0
    for i in sample:
        o, n = spans[i]
        got = out_f32[o:o + n]
        ref = refs[i].reshape(-1)
        if got.size != ref.size or not np.array_equal(got.view(np.uint32), ref.view(np.uint32)):
            mism += 1
    parity = {"oracle_sample_streams": len(sample), "oracle_sample_mismatches": mism, "oracle_compare": "float PCM, bit-identical",
              "sample_seeds": [rank * args.streams + i for i in sample]}

    # ---- tolerance-mode (FMA-contracted) granule kernel: time and deviation on the same batch --------------------------
    fused = None
    if not args.no_fused:
        rb.free()
        hb.flags = api.MATH_FUSED
        rbf = ctx.upload(hb)
        fstep_ms, fkern_ms, _ = timed_steps(rbf, max(3, min(args.steps, 10)), 3)
        mx, ident, cnt = 0.0, 0, 0
        for i in sample:
            o, n = spans[i]
            got = rbf.download(o, n)
            ref = refs[i].reshape(-1)
            mx = max(mx, float(np.abs(got.astype(np.float64) - ref.astype(np.float64)).max()))
            ident += int((q16(got) == q16(ref)).sum()); cnt += n
        fused = {"ms_per_launch": fkern_ms[1], "ms_per_step": fstep_ms, "value": total_audio / (fstep_ms * 1e-3),
                 "max_abs_delta_fs": mx, "identical_after_q16": ident / max(1, cnt), "compared_samples": cnt,
                 "north_star_bar": "max |delta| <= 1e-5 FS and >= 99.99 % identical after 16-bit quantisation",
                 "meets_bar": bool(mx <= 1e-5 and ident / max(1, cnt) >= 0.9999)}
        rbf.free()
        hb.flags = 0
        rb = None

    # ---- roofline of the dominant kernel (granule kernel) ----------------------------------------------
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    hbm_peak = float(peaks.get("hbm_gbs", HBM_FALLBACK_GBS))
    pcm_bytes = hb.pcm_floats * 4
    alg_bytes = pcm_bytes + int(hb.blob.size)          # PCM out + bitstream in (descriptors separate)
    flops = n_grch * FLOPS_PER_GRCH
    hbm_ceiling = hbm_peak * 1e9 / (alg_bytes / audio_s)              # audio-s/s if HBM-bound
    fp32_ceiling = FP32_NOMINAL_TFLOPS * 1e12 / (flops / audio_s)
    slower = "fp32" if fp32_ceiling < hbm_ceiling else "hbm"
    traffic = None
    tf = ROOT / "profiles" / TRAFFIC_FILE
    if tf.exists() and WORKLOAD == "config2" and args.streams == 1024 and args.seconds == 60.0:
        traffic = json.loads(tf.read_text()).get("granule", {}).get("traffic")

    def roof(ms):
        gbs, tfl = alg_bytes / (ms * 1e-3) / 1e9, flops / (ms * 1e-3) / 1e12
        return {"ms_per_launch": ms, "hbm_gbs": gbs, "hbm_frac_of_measured": gbs / hbm_peak, "fp32_tflops": tfl,
                "fp32_frac_of_nominal": tfl / FP32_NOMINAL_TFLOPS, "frac_of_slower_roof": (tfl / FP32_NOMINAL_TFLOPS) if slower == "fp32" else gbs / hbm_peak}

    r_exact = roof(gran_ms)
    roofline = {"bound": slower, "kernel": "l3_granule_kernel<2,4,exact>", "slower_roof": slower,
                "achieved": r_exact["fp32_tflops"] if slower == "fp32" else r_exact["hbm_gbs"],
                "peak": FP32_NOMINAL_TFLOPS if slower == "fp32" else hbm_peak, "unit": "TFLOP/s" if slower == "fp32" else "GB/s",
                "frac": r_exact["frac_of_slower_roof"],
                "peak_source": "nominal FP32 (148 SM x 128 lanes x 2 x 1.965 GHz; the driver measures no FP32 peak)" if slower == "fp32"
                               else ("measured" if peaks else "fallback"),
                "flops_counted": "reference operation count, 32,734 add/sub/mul per granule-channel (SURVEY 8d); an FMA-capable peak counts 2 per lane-slot",
                "hbm": {"achieved_gbs": r_exact["hbm_gbs"], "peak_gbs": hbm_peak, "frac": r_exact["hbm_frac_of_measured"],
                        "peak_source": "measured" if peaks else "fallback", "algorithmic_bytes_per_launch": alg_bytes},
                "traffic": traffic, "traffic_source": f"profiles/{TRAFFIC_FILE} (ncu dram read+write per launch, same workload)" if traffic else None,
                "ms_per_launch": gran_ms,
                "bit_exact": {**r_exact, "unfused_pipe_peak_tflops_measured": FP32_UNFUSED_MEASURED_TFLOPS,
                              "frac_of_unfused_pipe_peak": r_exact["fp32_tflops"] / FP32_UNFUSED_MEASURED_TFLOPS,
                              "note": "every product and sum rounded separately (bit-identical PCM): each flop takes one FP32-pipe lane slot, "
                                      "so the measured un-fused pipe peak (tools/microbench/fp32_pipes.cu) is this mode's ceiling, half the FMA-counted nominal"},
                "fused": ({**roof(fused["ms_per_launch"]), **{k: fused[k] for k in ("max_abs_delta_fs", "identical_after_q16", "meets_bar", "north_star_bar", "compared_samples")},
                           "whole_step_value": fused["value"]} if fused else None),
                "frac_of_slower_roof_whole_step": (total_audio / (step_ms * 1e-3)) / (world * min(fp32_ceiling, hbm_ceiling)),
                "entropy_kernels_ms": ent_ms, "granule_kernel_share_of_step": min(1.0, gran_ms / kern_ms[2]),
                "kernels": ["l3_scf_kernel", "l3_huff_big_kernel<16,8>", "l3_huff_c1_kernel<16>", "l3_granule_kernel<2,4,exact>"]
                           + (["l3_granule_kernel<1,4,exact>"] if n_grch_stereo != n_grch else []),
                "kernels_note": "one launch of each per step, in that order; entropy_kernels_ms spans the first three"}

    # ---- end to end: MP3 bytes (host) -> PCM (pinned host), through the library's pipeline ------------------------------
    e2e = None
    if not args.no_e2e:
        if rb is not None:
            rb.free()
            rb = None
        h2d = int(hb.blob.size) + hb.descs.size * 16 + hb.streams.size * 56

        def run_e2e(s16, steps):
            pipe = af.BatchPipeline(device=local, lanes=args.e2e_lanes, wave_streams=args.e2e_wave, prepass_threads=threads, s16=s16)
            out = pin_pcm.view(np.int16) if s16 else out_f32
            info = pipe.decode_into(datas, out)          # warm-up (allocates the recycled workspaces)
            pipe.profile()
            join.barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                info = pipe.decode_into(datas, out)
            join.barrier()
            sec = join.max((time.perf_counter() - t0) / steps)
            prof = pipe.profile()
            pipe.close()
            return sec, info, {k: v / steps for k, v in prof.items()}, out

        # 16-bit delivery (headline): compare a sample with q16(oracle) and every stream with q16 of the resident run
        sec16, info16, prof16, out16 = run_e2e(True, args.e2e_steps)
        bad16 = 0
        for i in sample:
            o, fr, ch, _hz = info16[i]
            if not np.array_equal(out16[o:o + fr * ch], q16(refs[i].reshape(-1))):
                bad16 += 1
        parity["e2e_s16_sample_mismatches"] = bad16
        # float delivery: every stream's checksum against the device-resident run (a second, differently batched decode)
        sec32, info32, prof32, _ = run_e2e(False, max(2, args.e2e_steps - 1))
        crc_e2e = stream_crcs(out_f32, [(o, fr * ch) for (o, fr, ch, _hz) in info32], threads)
        parity["crc_streams_checked"] = len(crc_e2e)
        parity["crc_mismatches"] = int(sum(a != b2 for a, b2 in zip(crc_e2e, crc_resident)))
        parity["crc_compare"] = "crc32 of every stream's float PCM: device-resident run vs the wave pipeline's decode"
        d2h16 = hb.pcm_floats * 2
        peak = d2h_rate_gbs(torch, join, out_f32)
        peak_mixed = d2h_rate_gbs(torch, join, out_f32, h2d_fraction=h2d / d2h16)
        e2e = {"value": total_audio / sec16, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h16,
               "ms_per_step": sec16 * 1e3, "steps": args.e2e_steps, "output": "int16 PCM in pinned host memory (L3B_OUT_S16)",
               "pipeline": {"lanes": args.e2e_lanes, "wave_streams": args.e2e_wave, "scan_threads": threads,
                            "seconds_per_step_by_phase_summed_over_threads": prof16},
               "includes": "host prepass + H2D (pinned) + kernels + D2H (pinned) of every step, through l3b_pipeline_decode",
               "d2h_achieved_gbs": d2h16 / sec16 / 1e9,
               "d2h_concurrent_peak_gbs": peak, "d2h_concurrent_peak_aggregate_gbs": None if peak is None else peak * world,
               "d2h_frac_of_concurrent_peak": None if not peak else d2h16 / sec16 / 1e9 / peak,
               "d2h_concurrent_peak_with_uploads_gbs": peak_mixed,
               "d2h_frac_of_concurrent_peak_with_uploads": None if not peak_mixed else d2h16 / sec16 / 1e9 / peak_mixed,
               "d2h_peak_note": "up to 8 x 1 GiB device->pinned copies into distinct pages of the output buffer, every rank at the same time "
                                "between two barriers, slowest rank, best of 2; `with_uploads`: the same while a host->device copy of "
                                "h2d_bytes_per_step / d2h_bytes_per_step of the bytes runs on a second stream, like the pipeline's uploads",
               "f32": {"value": total_audio / sec32, "ms_per_step": sec32 * 1e3, "d2h_bytes_per_step": pcm_bytes,
                       "d2h_achieved_gbs": pcm_bytes / sec32 / 1e9, "output": "float PCM in pinned host memory",
                       "seconds_per_step_by_phase_summed_over_threads": prof32}}
    pin_pcm.free()

    # ---- the prepass on the GPU (l3b_raw_open: frame walk, side info, reservoir, main-data gather as kernels) vs on the host ----
    device_prepass = None
    if rank == 0 and rb is None:
        try:
            ctx.raw_prepass(datas[:8])
            device_prepass = ctx.raw_prepass(datas)
            device_prepass["host_prepass_thread_ms"] = t_scan * 1e3 * threads
            device_prepass["note"] = ("open_wall_ms includes the H2D of the raw files from pageable memory and the allocation of the resident batch; "
                                      "prepass_kernels_ms is the device time of the walk / side-info / reservoir kernels")
        except api.L3BError as exc:
            device_prepass = {"error": str(exc)[:200]}

    # ---- BASELINE config 1: the transcode example's loop (open, 1,024-frame reads) on ONE 10 s stream through AudioStream ----
    config1 = None
    if rank == 0:
        from audio_formats_b200 import synth
        st1 = synth.generate(synth.config1_params(1))
        ref1 = oracle.transcode_loop(st1.data, 1024, keep=True)[3]

        def transcode_gpu():
            s = af.AudioStream(ctx).openFromMemory(st1.data)
            chunks = []
            while True:
                c = s.readSamplesFloat(1024)
                if len(c) == 0:
                    break
                chunks.append(c)
            s.close()
            return np.concatenate(chunks)

        got1 = transcode_gpu()                       # warm-up + check
        t0 = time.perf_counter()
        for _ in range(5):
            transcode_gpu()
        t_gpu = (time.perf_counter() - t0) / 5
        t0 = time.perf_counter()
        for _ in range(5):
            oracle.transcode_loop(st1.data, 1024, keep=False)
        t_cpu = (time.perf_counter() - t0) / 5
        config1 = {"workload": "BASELINE.json configs[0]: one synthetic 10 s 44.1 kHz stereo 128 kbps stream, transcode loop (open + 1,024-frame float reads)",
                   "audiostream_gpu_ms": t_gpu * 1e3, "cpu_port_1_thread_ms": t_cpu * 1e3,
                   "bit_identical": bool(got1.shape == ref1.shape and np.array_equal(got1.view(np.uint32), ref1.view(np.uint32))),
                   "note": "a single stream cannot fill a GPU: the AudioStream arm decodes ahead in windows of 2,048 frames (one upload, "
                           "four launches, one download per window); the batch entry point is the throughput path"}

    # ---- CPU baseline (rank 0): the oracle on a bounded sample of the same workload ---------------
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        ht = host_threads()
        a1, w1 = cpu_decode_throughput(streams[:1], 1)
        n_sample = int(max(ht, min(len(streams), args.cpu_baseline_seconds * ht / max(w1, 1e-6))))
        a, w = cpu_decode_throughput(streams[:n_sample], ht)
        cpu = {"value": a / w, "unit": UNIT, "cores": ht, "kind": "port",
               "sample": f"first {n_sample} of the {len(streams)} streams of rank 0, {args.seconds:g} s each",
               "single_thread_value": a1 / w1, "note": cpu_baseline_note()}
    join.barrier()

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": workload_config(args),
                "audio_seconds_per_step": total_audio, "granule_channels_per_gpu": n_grch,
                "clocks": clocks, "gpu_launches": launches_per_step * args.steps,
                "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu, "parity": parity, "config1_transcode": config1, "device_prepass": device_prepass,
                "setup": {"generate_s": t_gen, "prepass_s": t_scan, "host_threads": threads,
                          "join": "gloo (timing scalars only; no NCCL, no collective on the data path)" if world > 1 else "single process"}}
        print(json.dumps(line), flush=True)
    if rb is not None:
        rb.free()
    ctx.close()
    join.close()


if __name__ == "__main__":
    main()
