"""CPU oracle (TEST INFRASTRUCTURE ONLY) -- ctypes binding of oracle/libl3oracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product (audio_formats_b200) never does.  PARITY UNPINNED against the D reference itself;
cross-checked against FFmpeg's mp3float (see oracle/l3_oracle.h, tests/test_oracle_vs_ffmpeg.py).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "libl3oracle.so"


def build(force: bool = False) -> Path:
    srcs = [_HERE / "l3_oracle.c", _HERE / "l3_oracle_ex.c", _HERE / "l3_oracle.h",
            _HERE.parent / "audio_formats_b200" / "csrc" / "l3_tables_gen.h"]
    if force or not _SO.exists() or any(_SO.stat().st_mtime < s.stat().st_mtime for s in srcs):
        subprocess.check_call(["make", "-s", "-C", str(_HERE), "-B"])
    return _SO


class GranuleTap(C.Structure):
    _fields_ = [("is_", (C.c_int16 * 576) * 2), ("iscf", (C.c_uint8 * 40) * 2), ("ist_pos", (C.c_uint8 * 40) * 2),
                ("gain_exp", C.c_int32 * 2), ("scf", (C.c_float * 40) * 2), ("xr", (C.c_float * 576) * 2),
                ("st", (C.c_float * 576) * 2), ("im", (C.c_float * 576) * 2), ("dct", (C.c_float * 576) * 2)]


TAP_DTYPE = np.dtype([("is", np.int16, (2, 576)), ("iscf", np.uint8, (2, 40)), ("ist_pos", np.uint8, (2, 40)),
                      ("gain_exp", np.int32, (2,)), ("scf", np.float32, (2, 40)), ("xr", np.float32, (2, 576)),
                      ("st", np.float32, (2, 576)), ("im", np.float32, (2, 576)), ("dct", np.float32, (2, 576))])
assert TAP_DTYPE.itemsize == C.sizeof(GranuleTap), (TAP_DTYPE.itemsize, C.sizeof(GranuleTap))


class _Tap(C.Structure):
    _fields_ = [("rec", C.c_void_p), ("capacity", C.c_int), ("count", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_SO))
        L.l3o_stream_open_memory.restype = C.c_void_p
        L.l3o_stream_open_memory.argtypes = [C.c_char_p, C.c_size_t]
        L.l3o_stream_close.argtypes = [C.c_void_p]
        for name in ("l3o_stream_channels", "l3o_stream_samplerate", "l3o_stream_tell", "l3o_stream_last_error"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = C.c_int
        L.l3o_stream_length_frames.argtypes = [C.c_void_p]
        L.l3o_stream_length_frames.restype = C.c_longlong
        L.l3o_stream_read_float.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.l3o_stream_read_float.restype = C.c_int
        L.l3o_stream_seek.argtypes = [C.c_void_p, C.c_int]
        L.l3o_stream_seek.restype = C.c_int
        L.l3o_transcode_loop.restype = C.c_longlong
        L.l3o_transcode_loop.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t,
                                         C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.l3o_transcode_loop_s16.restype = C.c_longlong
        L.l3o_transcode_loop_s16.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t,
                                             C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.l3o_set_tap.argtypes = [C.c_void_p]
        L.l3o_enable_timers.argtypes = [C.c_int]
        L.l3o_get_timers.argtypes = [C.POINTER(C.c_double * 7)]
        _lib = L
    return _lib


class OracleStream:
    """AudioStream-shaped view of the oracle (mirrors stream.d's MP3 arms)."""

    def __init__(self, data: bytes):
        self._L = lib()
        self._h = self._L.l3o_stream_open_memory(data, len(data))
        if not self._h:
            raise ValueError("not detected as MP3 by the oracle")
        self.channels = self._L.l3o_stream_channels(self._h)
        self.samplerate = self._L.l3o_stream_samplerate(self._h)
        self.length_frames = self._L.l3o_stream_length_frames(self._h)

    def close(self):
        if self._h:
            self._L.l3o_stream_close(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def read_float(self, frames: int) -> np.ndarray:
        out = np.empty((frames, self.channels), dtype=np.float32)
        n = self._L.l3o_stream_read_float(self._h, out.ctypes.data, frames)
        return out[:n]

    def seek(self, frame: int) -> bool:
        return bool(self._L.l3o_stream_seek(self._h, frame))

    def tell(self) -> int:
        return self._L.l3o_stream_tell(self._h)

    @property
    def last_error(self) -> int:
        return self._L.l3o_stream_last_error(self._h)


def decode_all(data: bytes, chunk_frames: int = 1024, taps: int = 0):
    """Decode a whole stream through the transcode-example loop shape (1024-frame reads,
    examples/transcode/source/main.d:52-78).  Returns (pcm [frames, ch] float32, taps or None)."""
    L = lib()
    s = OracleStream(data)
    tap_arr = None
    tap = None
    if taps:
        tap_arr = np.zeros(taps, dtype=TAP_DTYPE)
        tap = _Tap(tap_arr.ctypes.data, taps, 0)
        L.l3o_set_tap(C.addressof(tap))
    try:
        chunks = []
        while True:
            c = s.read_float(chunk_frames)
            if len(c) == 0:
                break
            chunks.append(c.copy())
    finally:
        L.l3o_set_tap(None)
    pcm = np.concatenate(chunks) if chunks else np.zeros((0, s.channels), np.float32)
    s.close()
    if taps:
        return pcm, tap_arr[:min(tap.count, taps)]
    return pcm, None


def transcode_loop(data: bytes, chunk_frames: int = 1024, keep: bool = False, s16: bool = False):
    """Whole decode inside C (GIL released): returns (frames, channels, hz, pcm or None).
    s16: every chunk is also converted to 16 bit (un-dithered), the result discarded (keep is ignored)."""
    L = lib()
    nch, hz = C.c_int(), C.c_int()
    if s16:
        n = L.l3o_transcode_loop_s16(data, len(data), chunk_frames, None, 0, C.byref(nch), C.byref(hz))
        return n, nch.value, hz.value, None
    if keep:
        probe = OracleStream(data)
        cap = int(probe.length_frames) * probe.channels + 2304 * 4
        probe.close()
        out = np.empty(cap, np.float32)
        n = L.l3o_transcode_loop(data, len(data), chunk_frames, out.ctypes.data, cap, C.byref(nch), C.byref(hz))
        return n, nch.value, hz.value, out[: max(n, 0) * max(nch.value, 1)].reshape(-1, max(nch.value, 1))
    n = L.l3o_transcode_loop(data, len(data), chunk_frames, None, 0, C.byref(nch), C.byref(hz))
    return n, nch.value, hz.value, None
