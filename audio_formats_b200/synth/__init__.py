"""Deterministic synthetic Layer III bitstream generator (ctypes wrapper over l3synth.c).

Bench/test infrastructure: produces legal MPEG-1/2/2.5 Layer III streams at the syntax level and
the signed quantised integers it encoded (the ground truth for bit-exact spectral checks).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, asdict
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "libl3synth.so"


def build(force: bool = False) -> Path:
    srcs = [_HERE / "l3synth.c", _HERE / "l12synth.c"]
    if force or not _SO.exists() or any(_SO.stat().st_mtime < s.stat().st_mtime for s in srcs):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-Wall", "-o", str(_SO), *[str(s) for s in srcs]])
    return _SO


class _Params(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("hz", C.c_int), ("nch", C.c_int), ("bitrate_kbps", C.c_int),
                ("nframes", C.c_int), ("block_mode", C.c_int), ("stereo_mode", C.c_int), ("reservoir", C.c_int),
                ("scfsi", C.c_int), ("crc", C.c_int), ("escapes", C.c_int), ("gain_base", C.c_int),
                ("level", C.c_double), ("small_scalefactors", C.c_int), ("table_cycle", C.c_int),
                ("table_cycle_pos", C.c_int), ("no_padding", C.c_int), ("id3v2_bytes", C.c_int), ("id3v1", C.c_int),
                ("emphasis_bits", C.c_int), ("mixed_only_short", C.c_int), ("free_format", C.c_int), ("vbr", C.c_int), ("mode_ext_any", C.c_int), ("istereo_untied", C.c_int), ("private_bits", C.c_int)]


class _Info(C.Structure):
    _fields_ = [("frames", C.c_int), ("granules", C.c_int), ("samples_per_frame", C.c_int), ("mpeg1", C.c_int),
                ("sr_idx", C.c_int), ("bytes", C.c_longlong)]


class _L12Params(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("layer", C.c_int), ("hz", C.c_int), ("nch", C.c_int), ("bitrate_kbps", C.c_int),
                ("nframes", C.c_int), ("joint", C.c_int), ("crc", C.c_int), ("padding", C.c_int), ("fill", C.c_int),
                ("ref_syntax", C.c_int), ("all_alloc", C.c_int)]


_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_SO))
        _lib.l12s_max_bytes.restype = C.c_size_t
        _lib.l12s_max_bytes.argtypes = [C.POINTER(_L12Params)]
        _lib.l12s_generate.restype = C.c_longlong
        _lib.l12s_generate.argtypes = [C.POINTER(_L12Params), C.c_void_p, C.c_size_t]
        _lib.l3s_max_bytes.restype = C.c_size_t
        _lib.l3s_max_bytes.argtypes = [C.POINTER(_Params)]
        _lib.l3s_generate.restype = C.c_longlong
        _lib.l3s_generate.argtypes = [C.POINTER(_Params), C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(_Info)]
    return _lib


@dataclass
class SynthParams:
    seed: int = 1
    hz: int = 44100
    nch: int = 2
    bitrate_kbps: int = 128
    nframes: int = 383
    block_mode: int = 0
    stereo_mode: int = 0
    reservoir: int = 1
    scfsi: int = 0
    crc: int = 0
    escapes: int = 1
    gain_base: int = 188
    level: float = 3.0
    small_scalefactors: int = 1
    table_cycle: int = 0
    table_cycle_pos: int = 0
    no_padding: int = 0
    id3v2_bytes: int = 0
    id3v1: int = 0
    emphasis_bits: int = 0
    mixed_only_short: int = 0
    free_format: int = 0
    vbr: int = 0
    mode_ext_any: int = 0
    istereo_untied: int = 0
    private_bits: int = 0

    @staticmethod
    def for_seconds(seconds: float, hz: int = 44100, **kw) -> "SynthParams":
        spf = 1152 if hz >= 32000 else 576
        return SynthParams(hz=hz, nframes=max(1, int(round(seconds * hz / spf))), **kw)


@dataclass
class SynthStream:
    data: bytes
    params: SynthParams
    frames: int
    granules: int
    samples_per_frame: int
    quantised: np.ndarray | None  # int16 [granules, nch, 576] ground truth, or None

    @property
    def pcm_frames(self) -> int:
        return self.frames * self.samples_per_frame

    @property
    def seconds(self) -> float:
        return self.pcm_frames / self.params.hz


def generate(params: SynthParams, want_quantised: bool = False) -> SynthStream:
    lib = _load()
    p = _Params(**asdict(params))
    cap = lib.l3s_max_bytes(C.byref(p))
    if cap == 0:
        raise ValueError(f"illegal Layer III format: {params}")
    buf = np.zeros(cap, dtype=np.uint8)
    ngr = 2 if params.hz >= 32000 else 1
    q = np.zeros((params.nframes * ngr, params.nch, 576), dtype=np.int16) if want_quantised else None
    info = _Info()
    n = lib.l3s_generate(C.byref(p), buf.ctypes.data, cap, q.ctypes.data if q is not None else None, C.byref(info))
    if n < 0:
        raise RuntimeError(f"l3s_generate failed with {n} for {params}")
    return SynthStream(data=buf[:n].tobytes(), params=params, frames=info.frames, granules=info.granules,
                       samples_per_frame=info.samples_per_frame, quantised=q)


# The named workload profiles of BASELINE.json `configs`.
def config1_params(seed: int = 1) -> SynthParams:
    """10 s, 44.1 kHz stereo 128 kbps MPEG-1, long blocks, MS off, reservoir on."""
    return SynthParams.for_seconds(10.0, seed=seed, bitrate_kbps=128, reservoir=1)


def config2_params(seed: int, seconds: float = 60.0) -> SynthParams:
    """60 s, 44.1 kHz stereo 128 kbps MPEG-1, long blocks only, moderate reservoir."""
    return SynthParams.for_seconds(seconds, seed=seed, bitrate_kbps=128, reservoir=1)


def config3_params(seed: int, seconds: float = 60.0) -> SynthParams:
    """mixed long/short/mixed blocks, joint (MS + intensity) stereo, heavy reservoir, scfsi, all tables."""
    return SynthParams.for_seconds(seconds, seed=seed, bitrate_kbps=128, block_mode=1, stereo_mode=2, reservoir=2,
                                   scfsi=1, table_cycle=1, table_cycle_pos=seed, small_scalefactors=0)


_CFG4_FORMATS = [(32000, r) for r in (64, 96, 128, 160, 192, 256, 320)] + \
                [(44100, r) for r in (64, 96, 128, 160, 192, 256, 320)] + \
                [(48000, r) for r in (64, 96, 128, 160, 192, 256, 320)] + \
                [(16000, r) for r in (64, 80, 96, 112, 128, 144, 160)] + \
                [(22050, r) for r in (64, 80, 96, 112, 128, 144, 160)] + \
                [(24000, r) for r in (64, 80, 96, 112, 128, 144, 160)]


def config4_params(seed: int, seconds: float = 30.0) -> SynthParams:
    """heterogeneous: 32/44.1/48 kHz MPEG-1 and 16/22.05/24 kHz MPEG-2 LSF, 64-320 kbps, mono and stereo."""
    hz, rate = _CFG4_FORMATS[(seed * 7919) % len(_CFG4_FORMATS)]
    nch = 1 + ((seed // len(_CFG4_FORMATS) + seed) & 1)
    if nch == 1 and hz >= 32000 and rate > 192:
        rate = 192  # keep mono part2_3_length within the 12-bit field comfortably
    return SynthParams.for_seconds(seconds, hz=hz, seed=seed, nch=nch, bitrate_kbps=rate, block_mode=1,
                                   stereo_mode=2 if nch == 2 else 0, reservoir=1, scfsi=1, small_scalefactors=0)


def config5_params(seed: int, seconds: float = 180.0) -> SynthParams:
    """180 s, 44.1 kHz stereo 320 kbps."""
    return SynthParams.for_seconds(seconds, seed=seed, bitrate_kbps=320, reservoir=1, level=12.0, gain_base=186)


@dataclass
class L12Params:
    """Synthetic MPEG-1/2 Layer I / Layer II stream (l12synth.c)."""
    seed: int = 1
    layer: int = 2
    hz: int = 44100
    nch: int = 2
    bitrate_kbps: int = 192
    nframes: int = 60
    joint: int = 0
    crc: int = 0
    padding: int = 1
    fill: int = 90
    ref_syntax: int = 1   # Layer II written the way the D reference reads it (a scfsi field for every band-channel entry)
    all_alloc: int = 0    # every band allocated: with plain stereo the reference's reading and ISO syntax coincide


def generate_l12(params: L12Params) -> bytes:
    lib = _load()
    p = _L12Params(**asdict(params))
    cap = lib.l12s_max_bytes(C.byref(p))
    buf = np.zeros(cap, dtype=np.uint8)
    n = lib.l12s_generate(C.byref(p), buf.ctypes.data, cap)
    if n < 0:
        raise ValueError(f"l12s_generate failed with {n} for {params}")
    return buf[:n].tobytes()
