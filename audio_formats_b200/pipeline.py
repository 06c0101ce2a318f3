"""Wave-pipelined batch decode over one or more GPUs: MP3 bytes in host memory -> PCM in (pinned) host memory.

Thin ctypes wrapper of the library's own pipeline (csrc/l3_pipeline.cpp, `l3b_pipeline_*`): the wave loop, the prepass thread
pool and the per-GPU lanes all run inside the library (no GIL, blocking waits), this class only marshals the arguments.
"""
from __future__ import annotations

import ctypes as C
import sys
from typing import Sequence

import numpy as np

from . import api


class BatchPipeline:
    PHASES = ("scan", "assemble", "upload", "launch", "download", "wait_for_scan")

    def __init__(self, device: int | Sequence[int] = 0, lanes: int = 4, wave_streams: int = 16, prepass_threads: int = 0,
                 s16: bool = False, fused: bool = False):
        """device: one GPU index, or a list of them (streams are then assigned by file, longest first).
        s16: deliver 16-bit PCM (api.OUT_S16); fused: tolerance-mode arithmetic (api.MATH_FUSED)."""
        self._L = api.load_library()
        self.devices = [device] if isinstance(device, int) else list(device)
        self.flags = (api.OUT_S16 if s16 else 0) | (api.MATH_FUSED if fused else 0)
        self.dtype = np.int16 if s16 else np.float32
        opts = api.PipelineOpts(lanes, wave_streams, prepass_threads, self.flags)
        devs = (C.c_int * len(self.devices))(*self.devices)
        h = C.c_void_p()
        rc = self._L.l3b_pipeline_create(devs, len(self.devices), C.byref(opts), C.byref(h))
        if rc:
            raise api.L3BError(rc, (self._L.l3b_last_error(None) or b"").decode())
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.l3b_pipeline_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def decode_into(self, datas: Sequence[bytes], out: np.ndarray):
        """Decode every stream into `out` (float32, or int16 for an s16 pipeline; ideally pinned).  Returns one
        (offset, frames, channels, samplerate) per stream, or None for a stream that could not be decoded (see `status`);
        the streams of one wave are contiguous in `out`."""
        n = len(datas)
        if out.dtype != self.dtype:
            raise TypeError(f"output buffer must be {self.dtype}")
        ptrs = (C.c_char_p * n)(*datas)
        sizes = (C.c_size_t * n)(*[len(d) for d in datas])
        res = (api.StreamResult * n)()
        used = C.c_uint64()
        rc = self._L.l3b_pipeline_decode(self._h, ptrs, sizes, n, out.ctypes.data, out.size, res, C.byref(used))
        if rc:
            raise api.L3BError(rc, (self._L.l3b_pipeline_last_error(self._h) or b"").decode())
        self.status = [r.status for r in res]
        self.device_of = [r.device for r in res]
        self.used = used.value
        return [(int(r.pcm_off), int(r.frames), int(r.channels), int(r.samplerate)) if r.channels and r.status == 0 or r.frames
                else None for r in res]

    def profile(self) -> dict:
        """Seconds per phase summed over the threads since the last call (see l3b_pipeline_profile)."""
        sec = (C.c_double * api.PIPELINE_PHASES)()
        self._L.l3b_pipeline_profile(self._h, C.byref(sec))
        return dict(zip(self.PHASES, list(sec)))
