"""Wave-pipelined batch decode: MP3 bytes in host memory -> float PCM in (pinned) host memory.

The batch is cut into waves; `lanes` worker threads each own a GPU context (its own CUDA stream and a
recycled device workspace) and run  host prepass -> H2D -> entropy kernels -> granule kernel -> D2H  for their
waves.  While one lane copies PCM back over PCIe the other lanes scan and decode, so the copy engine, the
SMs and the host cores work at the same time.  ctypes releases the GIL inside every library call.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Sequence

import numpy as np

from . import api


class BatchPipeline:
    def __init__(self, device: int = 0, lanes: int = 6, wave_streams: int = 16, prepass_threads: int = 8):
        self.device, self.lanes, self.wave_streams = device, lanes, wave_streams
        self._ctxs = [api.Context(device) for _ in range(lanes)]
        self._res = [C.c_void_p() for _ in range(lanes)]
        self._staging: list = [None] * lanes          # pinned H2D staging per lane, grown on demand
        self._pool = ThreadPoolExecutor(max(1, prepass_threads))
        self._lock = threading.Lock()

    def close(self):
        for c, r in zip(self._ctxs, self._res):
            if r:
                c._L.l3b_batch_free(c._h, r)
            c.close()
        for st in self._staging:
            if st is not None:
                st.free()
        self._staging = []
        self._ctxs, self._res = [], []
        self._pool.shutdown(wait=False)

    def decode_into(self, datas: Sequence[bytes], out: np.ndarray):
        """Decode every stream into the float32 buffer `out` (ideally pinned).  Returns a list of
        (offset, frames, channels, samplerate) per stream; streams of one wave are contiguous in `out`."""
        n = len(datas)
        waves = [range(i, min(n, i + self.wave_streams)) for i in range(0, n, self.wave_streams)]
        results: list = [None] * n
        cursor = [0]
        errors: list = []
        next_wave = [0]

        prof = os.environ.get("L3B_PIPELINE_PROFILE") == "1"
        phases = [[0.0] * 5 for _ in range(self.lanes)]   # scan, assemble, upload, run (issue), download (incl. kernels)
        import time as _time

        def lane(k: int):
            ctx = self._ctxs[k]
            L = ctx._L
            try:
                while True:
                    with self._lock:
                        w = next_wave[0]
                        next_wave[0] += 1
                    if w >= len(waves):
                        return
                    idxs = waves[w]
                    t0 = _time.perf_counter()
                    scans = list(self._pool.map(api.Scan, [datas[i] for i in idxs]))   # host prepass
                    t1 = _time.perf_counter()
                    need = sum(int(s._L.l3b_scan_maindata_bytes(s._h)) + 48 + 16 * s.granules * s.channels for s in scans) + 4096
                    if self._staging[k] is None or self._staging[k].nbytes < need:
                        if self._staging[k] is not None:
                            self._staging[k].free()
                        self._staging[k] = api.PinnedBuffer(need + need // 4)
                    hb = api.HostBatch(scans, staging=self._staging[k])
                    with self._lock:                                                   # reserve the output region
                        base = (cursor[0] + 3) & ~3
                        cursor[0] = base + hb.pcm_floats
                    if base + hb.pcm_floats > out.size:
                        raise ValueError("output buffer too small")
                    t2 = _time.perf_counter()
                    ctx._check(L.l3b_batch_upload_reuse(ctx._h, C.byref(hb.c_batch()), C.byref(self._res[k])))
                    t3 = _time.perf_counter()
                    ctx._check(L.l3b_batch_run(ctx._h, self._res[k]))
                    t4 = _time.perf_counter()
                    ctx._check(L.l3b_batch_download(ctx._h, self._res[k], out.ctypes.data + 4 * base, 0, hb.pcm_floats))
                    t5 = _time.perf_counter()
                    if prof:
                        for j, dt in enumerate((t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
                            phases[k][j] += dt
                    for i, s, sd in zip(idxs, scans, hb.streams):
                        results[i] = (base + int(sd["pcm_off"]), int(sd["pcm_count"]) // s.channels, s.channels, s.samplerate)
            except Exception as e:  # noqa: BLE001
                errors.append(e)

        # A lane that comes back from a library call needs the GIL to issue the next one; with the default 5 ms switch
        # interval it can wait that long behind a thread that is running bytecode, and the copy engine idles meanwhile.
        old_interval = sys.getswitchinterval()
        sys.setswitchinterval(float(os.environ.get("L3B_PIPELINE_SWITCH_INTERVAL", "2e-4")))
        try:
            threads = [threading.Thread(target=lane, args=(k,)) for k in range(self.lanes)]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
        finally:
            sys.setswitchinterval(old_interval)
        if prof:
            tot = [sum(p[j] for p in phases) for j in range(5)]
            print("pipeline profile, seconds summed over lanes: scan %.3f assemble %.3f upload %.3f run %.3f download %.3f" % tuple(tot),
                  file=sys.stderr)
        if errors:
            raise errors[0]
        return results
