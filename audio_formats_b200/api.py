"""ctypes mirror of include/l3b200.h.

Names follow the reference's AudioStream (source/audioformats/stream.d): openFromMemory,
openFromFile, readSamplesFloat, readSamplesDouble, seekPosition, tellPosition, getNumChannels,
getLengthInFrames, getSamplerate, isError, errorMessage -- plus the batch entry point
``Context.decode_scans`` the north star adds.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import Sequence

import numpy as np

_HERE = Path(__file__).resolve().parent
import os as _os

# L3B_LIB selects an experimental build of the library (tools/, profiling); the product is libl3b200.so
_LIB_PATH = Path(_os.environ["L3B_LIB"]) if _os.environ.get("L3B_LIB") else _HERE / "libl3b200.so"
_lib = None


class L3BError(RuntimeError):
    def __init__(self, code: int, msg: str = ""):
        super().__init__(f"l3b200 error {code}: {msg}" if msg else f"l3b200 error {code}")
        self.code = code


E_PARAM, E_MEMORY, E_IOERROR, E_USER, E_DECODE, E_NOGPU, E_UNSUPPORTED = -1, -2, -3, -4, -5, -16, -17


class GrchDesc(C.Structure):
    _fields_ = [("bit_start", C.c_uint32), ("w1", C.c_uint32), ("w2", C.c_uint32), ("w3", C.c_uint32)]


class StreamDesc(C.Structure):
    _fields_ = [("maindata_off", C.c_uint64), ("maindata_bytes", C.c_uint32), ("n_granules", C.c_uint32),
                ("first_grch", C.c_uint64), ("pcm_off", C.c_uint64), ("pcm_skip", C.c_uint64),
                ("pcm_count", C.c_uint64), ("nch", C.c_uint8), ("sr_idx", C.c_uint8), ("mpeg1", C.c_uint8),
                ("layer", C.c_uint8), ("reserved2", C.c_uint32)]


class Taps(C.Structure):
    _fields_ = [("is_", C.c_void_p), ("iscf", C.c_void_p), ("ist_pos", C.c_void_p),
                ("xr", C.c_void_p), ("st", C.c_void_p), ("im", C.c_void_p), ("dct", C.c_void_p)]


OUT_S16, MATH_FUSED = 1, 2   # l3b_batch_t.flags


class Batch(C.Structure):
    _fields_ = [("maindata", C.c_void_p), ("maindata_bytes", C.c_uint64), ("grch", C.c_void_p), ("n_grch", C.c_uint64),
                ("streams", C.c_void_p), ("n_streams", C.c_uint32), ("pcm", C.c_void_p), ("pcm_floats", C.c_uint64),
                ("status", C.c_void_p), ("taps", C.c_void_p), ("flags", C.c_uint32), ("reserved", C.c_uint32)]


class PipelineOpts(C.Structure):
    _fields_ = [("lanes", C.c_int32), ("wave_streams", C.c_int32), ("scan_threads", C.c_int32), ("flags", C.c_uint32)]


class StreamResult(C.Structure):
    _fields_ = [("pcm_off", C.c_uint64), ("frames", C.c_uint64), ("channels", C.c_int32), ("samplerate", C.c_int32),
                ("status", C.c_int32), ("device", C.c_int32)]


PIPELINE_PHASES = 6
READ_CB = C.CFUNCTYPE(C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p)   # mp3dec_io_t.read, minimp3_ex.d:61-71
SEEK_CB = C.CFUNCTYPE(C.c_int, C.c_uint64, C.c_void_p)                  # mp3dec_io_t.seek
assert C.sizeof(GrchDesc) == 16 and C.sizeof(StreamDesc) == 56, (C.sizeof(GrchDesc), C.sizeof(StreamDesc))

GRCH_DTYPE = np.dtype([("bit_start", "<u4"), ("w1", "<u4"), ("w2", "<u4"), ("w3", "<u4")])
STREAM_DTYPE = np.dtype([("maindata_off", "<u8"), ("maindata_bytes", "<u4"), ("n_granules", "<u4"),
                         ("first_grch", "<u8"), ("pcm_off", "<u8"), ("pcm_skip", "<u8"), ("pcm_count", "<u8"),
                         ("nch", "u1"), ("sr_idx", "u1"), ("mpeg1", "u1"), ("layer", "u1"), ("reserved2", "<u4")])
assert STREAM_DTYPE.itemsize == 56


def library_path() -> Path:
    return _LIB_PATH


def load_library():
    """Load libl3b200.so.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ["L3B_LIB"]) if os.environ.get("L3B_LIB") else _LIB_PATH   # L3B_LIB: an experimental build (A/B runs)
    if not path.exists():
        raise L3BError(E_NOGPU, f"{path} is missing: run `python -m audio_formats_b200.build` "
                                "(there is no CPU fallback)")
    L = C.CDLL(str(path))
    vp, u8p = C.c_void_p, C.c_char_p
    sig = {
        "l3b_device_count": (C.c_int, []),
        "l3b_ctx_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
        "l3b_ctx_destroy": (None, [vp]),
        "l3b_last_error": (C.c_char_p, [vp]),
        "l3b_host_alloc": (vp, [C.c_size_t]),
        "l3b_host_alloc_near": (vp, [C.c_int, C.c_size_t]),
        "l3b_host_free": (None, [vp]),
        "l3b_decode_batch": (C.c_int, [vp, C.POINTER(Batch)]),
        "l3b_batch_upload": (C.c_int, [vp, C.POINTER(Batch), C.POINTER(vp)]),
        "l3b_batch_upload_reuse": (C.c_int, [vp, C.POINTER(Batch), C.POINTER(vp)]),
        "l3b_batch_reupload": (C.c_int, [vp, vp, C.POINTER(Batch)]),
        "l3b_batch_run": (C.c_int, [vp, vp]),
        "l3b_batch_sync": (C.c_int, [vp]),
        "l3b_batch_download": (C.c_int, [vp, vp, vp, C.c_uint64, C.c_uint64]),
        "l3b_batch_download_taps": (C.c_int, [vp, vp, C.POINTER(Taps)]),
        "l3b_batch_device_pcm": (vp, [vp]),
        "l3b_batch_free": (None, [vp, vp]),
        "l3b_batch_timing": (C.c_int, [vp, C.c_int, C.POINTER(C.c_float * 3), C.POINTER(C.c_int)]),
        "l3b_ctx_cuda_stream": (vp, [vp]),
        "l3b_scan_memory": (C.c_int, [u8p, C.c_size_t, C.POINTER(vp)]),
        "l3b_scan_free": (None, [vp]),
        "l3b_scan_channels": (C.c_int, [vp]),
        "l3b_scan_samplerate": (C.c_int, [vp]),
        "l3b_scan_error": (C.c_int, [vp]),
        "l3b_scan_length_frames": (C.c_uint64, [vp]),
        "l3b_scan_delivered_samples": (C.c_uint64, [vp]),
        "l3b_scan_granules": (C.c_uint32, [vp]),
        "l3b_scan_maindata_bytes": (C.c_uint64, [vp]),
        "l3b_scan_maindata": (vp, [vp]),
        "l3b_scan_descs": (vp, [vp]),
        "l3b_scan_fill_stream_desc": (None, [vp, C.POINTER(StreamDesc)]),
        "l3b_scans_assemble": (C.c_int, [C.POINTER(vp), C.c_uint32, vp, C.c_uint64, vp, C.c_uint64, vp, C.POINTER(Batch)]),
        "l3b_decode_scans": (C.c_int, [vp, C.POINTER(vp), C.c_uint32, C.POINTER(vp), C.POINTER(C.c_int32)]),
        "l3b_raw_open": (C.c_int, [vp, C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_uint32, C.c_uint32, C.POINTER(vp)]),
        "l3b_raw_device_streams": (C.c_uint32, [vp]),
        "l3b_raw_prepass_ms": (C.c_float, [vp]),
        "l3b_raw_channels": (C.c_int, [vp, C.c_uint32]),
        "l3b_raw_samplerate": (C.c_int, [vp, C.c_uint32]),
        "l3b_raw_samples": (C.c_uint64, [vp, C.c_uint32]),
        "l3b_raw_status": (C.c_int, [vp, C.c_uint32]),
        "l3b_raw_decode": (C.c_int, [vp, C.POINTER(vp)]),
        "l3b_raw_free": (None, [vp]),
        "l3b_pipeline_create": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(PipelineOpts), C.POINTER(vp)]),
        "l3b_pipeline_destroy": (None, [vp]),
        "l3b_pipeline_decode": (C.c_int, [vp, C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_uint32, vp, C.c_uint64,
                                          C.POINTER(StreamResult), C.POINTER(C.c_uint64)]),
        "l3b_pipeline_profile": (C.c_int, [vp, C.POINTER(C.c_double * PIPELINE_PHASES)]),
        "l3b_pipeline_last_error": (C.c_char_p, [vp]),
        "l3b_stream_open_memory": (C.c_int, [vp, u8p, C.c_size_t, C.POINTER(vp)]),
        "l3b_stream_open_file": (C.c_int, [vp, C.c_char_p, C.POINTER(vp)]),
        "l3b_stream_open_callbacks": (C.c_int, [vp, READ_CB, SEEK_CB, vp, C.POINTER(vp)]),
        "l3b_stream_close": (None, [vp]),
        "l3b_stream_num_channels": (C.c_int, [vp]),
        "l3b_stream_length_frames": (C.c_int64, [vp]),
        "l3b_stream_samplerate": (C.c_float, [vp]),
        "l3b_stream_read_float": (C.c_int, [vp, vp, C.c_int]),
        "l3b_stream_read_double": (C.c_int, [vp, vp, C.c_int]),
        "l3b_stream_seek": (C.c_int, [vp, C.c_int]),
        "l3b_stream_tell": (C.c_int, [vp]),
        "l3b_stream_is_error": (C.c_int, [vp]),
        "l3b_stream_error_message": (C.c_char_p, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError here means the library does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    L._l3b_signatures = sig
    _lib = L
    return L


def device_count() -> int:
    return load_library().l3b_device_count()


class Scan:
    """Host prepass of one in-memory MP3 stream (frame sync, side info, reservoir slicing)."""

    def __init__(self, data: bytes):
        L = load_library()
        h = C.c_void_p()
        rc = L.l3b_scan_memory(data, len(data), C.byref(h))
        if rc:
            raise L3BError(rc, "scan failed (not an MP3 / unsupported layer)" if rc in (E_USER, E_UNSUPPORTED) else "")
        self._h = h
        self._L = L
        self.channels = L.l3b_scan_channels(h)
        self.samplerate = L.l3b_scan_samplerate(h)
        self.length_frames = L.l3b_scan_length_frames(h)
        self.delivered_samples = L.l3b_scan_delivered_samples(h)
        self.granules = L.l3b_scan_granules(h)
        self.error = L.l3b_scan_error(h)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.l3b_scan_free(self._h)
            self._h = None

    @property
    def descs(self) -> np.ndarray:
        n = self.granules * self.channels
        if not n:
            return np.zeros(0, GRCH_DTYPE)
        buf = (C.c_uint8 * (n * 16)).from_address(self._L.l3b_scan_descs(self._h))
        return np.frombuffer(buf, dtype=GRCH_DTYPE).copy()

    @property
    def maindata(self) -> np.ndarray:
        n = self._L.l3b_scan_maindata_bytes(self._h)
        if not n:
            return np.zeros(0, np.uint8)
        buf = (C.c_uint8 * n).from_address(self._L.l3b_scan_maindata(self._h))
        return np.frombuffer(buf, dtype=np.uint8).copy()

    def stream_desc(self) -> StreamDesc:
        d = StreamDesc()
        self._L.l3b_scan_fill_stream_desc(self._h, C.byref(d))
        return d


class Context:
    """One GPU context (one per GPU, one host thread at a time)."""

    def __init__(self, device: int = 0):
        L = load_library()
        h = C.c_void_p()
        rc = L.l3b_ctx_create(device, C.byref(h))
        if rc:
            raise L3BError(rc, (L.l3b_last_error(None) or b"").decode())
        self._h, self._L, self.device = h, L, device

    def close(self):
        if getattr(self, "_h", None):
            self._L.l3b_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    @property
    def cuda_stream(self) -> int:
        return self._L.l3b_ctx_cuda_stream(self._h)

    def _check(self, rc: int):
        if rc:
            raise L3BError(rc, (self._L.l3b_last_error(self._h) or b"").decode())

    # ---- the batch entry point ---------------------------------------------------------------
    def decode_scans(self, scans: Sequence[Scan]) -> list[np.ndarray]:
        """Decode every stream of the batch on the GPU; returns one [frames, channels] float32 array each."""
        n = len(scans)
        outs = [np.empty(s.delivered_samples, dtype=np.float32) for s in scans]
        hs = (C.c_void_p * n)(*[s._h for s in scans])
        ps = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
        st = (C.c_int32 * n)()
        self._check(self._L.l3b_decode_scans(self._h, hs, n, ps, st))
        return [o.reshape(-1, s.channels) for o, s in zip(outs, scans)]

    def decode(self, datas: Sequence[bytes]) -> list[np.ndarray]:
        return self.decode_scans([Scan(d) for d in datas])

    def decode_raw(self, datas: Sequence[bytes], flags: int = 0):
        """Batch decode with the prepass on the GPU (l3b_raw_*): returns (list of [frames, channels] arrays or None,
        info dict with the number of streams whose prepass ran on the device and its kernel time)."""
        n = len(datas)
        ptrs = (C.c_char_p * n)(*datas)
        sizes = (C.c_size_t * n)(*[len(d) for d in datas])
        h = C.c_void_p()
        self._check(self._L.l3b_raw_open(self._h, ptrs, sizes, n, flags, C.byref(h)))
        try:
            dt = np.int16 if flags & OUT_S16 else np.float32
            status = [self._L.l3b_raw_status(h, i) for i in range(n)]
            outs = [np.empty(self._L.l3b_raw_samples(h, i), dtype=dt) for i in range(n)]
            ps = (C.c_void_p * n)(*[o.ctypes.data if o.size else None for o in outs])
            self._check(self._L.l3b_raw_decode(h, ps))
            res = [o.reshape(-1, self._L.l3b_raw_channels(h, i)) if self._L.l3b_raw_channels(h, i) else None for i, o in enumerate(outs)]
            info = {"device_streams": self._L.l3b_raw_device_streams(h), "prepass_ms": self._L.l3b_raw_prepass_ms(h), "status": status}
        finally:
            self._L.l3b_raw_free(h)
        return res, info

    def raw_prepass(self, datas: Sequence[bytes]) -> dict:
        """Only the prepass of decode_raw (upload of the raw files + the device kernels + layout of the batch), for timing."""
        import time
        n = len(datas)
        ptrs = (C.c_char_p * n)(*datas)
        sizes = (C.c_size_t * n)(*[len(d) for d in datas])
        h = C.c_void_p()
        t0 = time.perf_counter()
        self._check(self._L.l3b_raw_open(self._h, ptrs, sizes, n, 0, C.byref(h)))
        wall = time.perf_counter() - t0
        info = {"device_streams": self._L.l3b_raw_device_streams(h), "streams": n, "prepass_kernels_ms": self._L.l3b_raw_prepass_ms(h),
                "open_wall_ms": wall * 1e3}
        self._L.l3b_raw_free(h)
        return info

    # ---- resident batches (throughput work) ----------------------------------------------------
    def upload(self, batch: "HostBatch") -> "ResidentBatch":
        h = C.c_void_p()
        self._check(self._L.l3b_batch_upload(self._h, C.byref(batch.c_batch(with_taps=batch.want_taps)), C.byref(h)))
        return ResidentBatch(self, h, batch)


class PinnedBuffer:
    """Page-locked host memory (cudaHostAlloc) viewed as a numpy array."""

    def __init__(self, nbytes: int, near_device: int | None = None):
        self._L = load_library()
        self.nbytes = int(nbytes)
        # near_device: place the pages on the NUMA node that GPU hangs off (l3b_host_alloc_near)
        self.ptr = self._L.l3b_host_alloc(self.nbytes) if near_device is None else self._L.l3b_host_alloc_near(near_device, self.nbytes)
        if not self.ptr:
            raise L3BError(E_MEMORY, f"cannot pin {nbytes} bytes of host memory")
        self.u8 = np.frombuffer((C.c_uint8 * self.nbytes).from_address(self.ptr), dtype=np.uint8)

    def view(self, dtype, count=None):
        a = self.u8.view(dtype)
        return a if count is None else a[:count]

    def free(self):
        if getattr(self, "ptr", None):
            self.u8 = None
            self._L.l3b_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class HostBatch:
    """A batch assembled on the host from scans (what the D host hands to the shim).
    `staging` (optional PinnedBuffer) receives the blob and the descriptors so the H2D copies read pinned memory."""

    def __init__(self, scans: Sequence[Scan], want_taps: bool = False, replicate: int = 1, staging: "PinnedBuffer | None" = None,
                 flags: int = 0, float_taps: bool = False):
        self.scans = list(scans)
        self.want_taps = want_taps or float_taps
        self.float_taps = float_taps
        self.flags = flags
        if replicate == 1 and self.scans:
            # assembled inside the library (l3b_scans_assemble): plain memcpys with the GIL released, so that the lanes
            # of a BatchPipeline build their waves in parallel
            L = load_library()
            n = len(self.scans)
            hs = (C.c_void_p * n)(*[sc._h for sc in self.scans])
            b = Batch()
            rc = L.l3b_scans_assemble(hs, n, None, 0, None, 0, None, C.byref(b))
            if rc:
                raise L3BError(rc, "l3b_scans_assemble (size query)")
            nb, nd = int(b.maindata_bytes), int(b.n_grch)
            doff = (nb + 63) & ~63
            if staging is not None and doff + 16 * nd + 64 <= staging.nbytes:
                self.blob = staging.u8[:nb]
                self.descs = staging.u8[doff:doff + 16 * nd].view(GRCH_DTYPE)
            else:
                self.blob = np.empty(nb, np.uint8)
                self.descs = np.empty(nd, GRCH_DTYPE)
            sd = np.zeros(n, dtype=STREAM_DTYPE)
            rc = L.l3b_scans_assemble(hs, n, self.blob.ctypes.data, nb, self.descs.ctypes.data, nd, sd.ctypes.data, C.byref(b))
            if rc:
                raise L3BError(rc, "l3b_scans_assemble")
            self.streams = sd
            self.pcm_floats = int(b.pcm_floats)
            self.n_grch = nd
            self._taps = Taps()
            return
        blobs, descs = [], []
        sd = np.zeros(len(scans) * replicate, dtype=STREAM_DTYPE)
        off = grch = pcm = 0
        k = 0
        cache = [(s.maindata, s.descs, s.stream_desc()) for s in self.scans]
        for _ in range(replicate):
            for (md, ds, d0) in cache:
                pad = (-len(md)) % 16 + 16
                pcm = (pcm + 3) & ~3   # 16-byte aligned PCM rows (stereo stores are 8-byte vectors)
                sd[k] = (off, len(md), d0.n_granules, grch, pcm, d0.pcm_skip, d0.pcm_count, d0.nch, d0.sr_idx, d0.mpeg1, d0.layer, 0)
                blobs.append(md)
                blobs.append(np.zeros(pad, np.uint8))
                descs.append(ds)
                off += len(md) + pad
                grch += len(ds)
                pcm += d0.pcm_count
                k += 1
        n_desc = sum(len(d) for d in descs)
        if staging is not None and off + 16 * n_desc + 64 <= staging.nbytes:
            self.blob = staging.u8[:off]
            np.concatenate(blobs, out=self.blob) if blobs else None
            doff = (off + 63) & ~63
            self.descs = staging.u8[doff:doff + 16 * n_desc].view(GRCH_DTYPE)
            np.concatenate(descs, out=self.descs) if descs else None
        else:
            self.blob = np.concatenate(blobs) if blobs else np.zeros(0, np.uint8)
            self.descs = np.concatenate(descs) if descs else np.zeros(0, GRCH_DTYPE)
        self.streams = sd
        self.pcm_floats = int(pcm)
        self.n_grch = int(grch)
        self._taps = Taps()

    def c_batch(self, pcm: np.ndarray | None = None, with_taps: bool = False) -> Batch:
        b = Batch()
        b.maindata, b.maindata_bytes = self.blob.ctypes.data, self.blob.size
        b.grch, b.n_grch = self.descs.ctypes.data, self.n_grch
        b.streams, b.n_streams = self.streams.ctypes.data, len(self.streams)
        b.pcm = pcm.ctypes.data if pcm is not None else None
        b.pcm_floats = self.pcm_floats
        b.status = None
        if with_taps and self.float_taps:   # non-NULL float tap pointers make the library allocate the float tap buffers
            self._ft = {k: np.zeros((self.n_grch, 576), np.float32) for k in ("xr", "st", "im", "dct")}
            self._taps.xr, self._taps.st = self._ft["xr"].ctypes.data, self._ft["st"].ctypes.data
            self._taps.im, self._taps.dct = self._ft["im"].ctypes.data, self._ft["dct"].ctypes.data
        b.taps = C.cast(C.pointer(self._taps), C.c_void_p) if with_taps else None
        b.flags = self.flags
        return b


class ResidentBatch:
    def __init__(self, ctx: Context, h, host: HostBatch):
        self.ctx, self._h, self.host = ctx, h, host

    def run(self):
        self.ctx._check(self.ctx._L.l3b_batch_run(self.ctx._h, self._h))

    def reupload(self):
        """Per-step H2D of the inputs (blob + descriptors) into the resident device buffers."""
        self.ctx._check(self.ctx._L.l3b_batch_reupload(self.ctx._h, self._h, C.byref(self.host.c_batch())))

    def download_into(self, ptr: int, first: int, count: int):
        self.ctx._check(self.ctx._L.l3b_batch_download(self.ctx._h, self._h, ptr, first, count))

    def sync(self):
        self.ctx._check(self.ctx._L.l3b_batch_sync(self.ctx._h))

    def timing(self, last_runs: int = 1):
        """(ms summed over the last runs [entropy, granule stereo, granule mono], kernels launched)."""
        ms = (C.c_float * 3)()
        n = C.c_int()
        self.ctx._check(self.ctx._L.l3b_batch_timing(self.ctx._h, last_runs, C.byref(ms), C.byref(n)))
        return list(ms), n.value

    def download(self, first: int = 0, count: int | None = None) -> np.ndarray:
        """PCM samples [first, first + count): float32, or int16 when the batch was built with OUT_S16."""
        count = self.host.pcm_floats - first if count is None else count
        out = np.empty(count, np.int16 if self.host.flags & OUT_S16 else np.float32)
        self.ctx._check(self.ctx._L.l3b_batch_download(self.ctx._h, self._h, out.ctypes.data, first, count))
        return out

    def download_taps(self):
        n = self.host.n_grch
        is_ = np.empty((n, 576), np.int16)
        iscf = np.empty((n, 40), np.uint8)
        ist = np.empty((n, 40), np.uint8)
        t = Taps(is_.ctypes.data, iscf.ctypes.data, ist.ctypes.data)
        self.ctx._check(self.ctx._L.l3b_batch_download_taps(self.ctx._h, self._h, C.byref(t)))
        return is_, iscf, ist

    def download_float_taps(self):
        """Float stage snapshots {xr, st, im, dct}: [n_grch, 576] each (the batch must have been built with float_taps)."""
        n = self.host.n_grch
        arrs = {k: np.zeros((n, 576), np.float32) for k in ("xr", "st", "im", "dct")}
        t = Taps(None, None, None, arrs["xr"].ctypes.data, arrs["st"].ctypes.data, arrs["im"].ctypes.data, arrs["dct"].ctypes.data)
        self.ctx._check(self.ctx._L.l3b_batch_download_taps(self.ctx._h, self._h, C.byref(t)))
        return arrs

    @property
    def device_pcm_ptr(self) -> int:
        return self.ctx._L.l3b_batch_device_pcm(self._h)

    def free(self):
        if self._h:
            self.ctx._L.l3b_batch_free(self.ctx._h, self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def decode_batch_with_taps(ctx: Context, scans: Sequence[Scan], float_taps: bool = False):
    """Test helper: decode a resident batch with taps and return
    (list of pcm arrays, is[n_grch,576], iscf[n_grch,40], ist_pos[n_grch,40]) -- plus, with float_taps, a dict of the float
    stage snapshots {xr, st, im, dct} [n_grch, 576] and the stream table (first_grch / pcm_skip locate the delivered granules)."""
    hb = HostBatch(scans, want_taps=True, float_taps=float_taps)
    rb = ctx.upload(hb)
    try:
        rb.run()
        rb.sync()
        pcm = rb.download()
        taps = rb.download_taps()
        ftaps = rb.download_float_taps() if float_taps else None
    finally:
        rb.free()
    outs = []
    for s, sdesc in zip(scans, hb.streams):
        off, n = int(sdesc["pcm_off"]), int(sdesc["pcm_count"])
        outs.append(pcm[off:off + n].reshape(-1, s.channels))
    if float_taps:
        return outs, *taps, ftaps, hb.streams
    return outs, *taps


def decode_mode(ctx: Context, scans: Sequence[Scan], flags: int):
    """Decode a batch in a given output / arithmetic mode (OUT_S16, MATH_FUSED); returns one [frames, channels] array per stream."""
    hb = HostBatch(scans, flags=flags)
    rb = ctx.upload(hb)
    try:
        rb.run()
        rb.sync()
        pcm = rb.download()
    finally:
        rb.free()
    return [pcm[int(sd["pcm_off"]):int(sd["pcm_off"]) + int(sd["pcm_count"])].reshape(-1, s.channels) for s, sd in zip(scans, hb.streams)]


class AudioStream:
    """Mirror of audio-formats' AudioStream for MP3 input (stream.d:102-1925, MP3 arms only)."""

    def __init__(self, ctx: Context | None = None):
        self._ctx = ctx
        self._own_ctx = False
        self._h = None
        self._L = load_library()

    def _ensure_ctx(self):
        if self._ctx is None:
            self._ctx = Context(0)
            self._own_ctx = True

    def openFromMemory(self, data: bytes) -> "AudioStream":
        self._ensure_ctx()
        h = C.c_void_p()
        rc = self._L.l3b_stream_open_memory(self._ctx._h, data, len(data), C.byref(h))
        if rc:
            raise L3BError(rc, "cannot open MP3 stream")
        self._h = h
        return self

    def openFromFile(self, path: str) -> "AudioStream":
        self._ensure_ctx()
        h = C.c_void_p()
        rc = self._L.l3b_stream_open_file(self._ctx._h, str(path).encode(), C.byref(h))
        if rc:
            raise L3BError(rc, "cannot open MP3 file")
        self._h = h
        return self

    def openFromCallbacks(self, read, seek=None) -> "AudioStream":
        """read(n) -> bytes (short = end of input), seek(position) -> None: the IOCallbacks shape of io.d:16-26."""
        self._ensure_ctx()

        def _read(buf, size, _user):
            chunk = read(size)
            C.memmove(buf, chunk, len(chunk))
            return len(chunk)

        def _seek(pos, _user):
            if seek is not None:
                seek(pos)
            return 0

        rcb, scb = READ_CB(_read), SEEK_CB(_seek)
        h = C.c_void_p()
        rc = self._L.l3b_stream_open_callbacks(self._ctx._h, rcb, scb, None, C.byref(h))
        if rc:
            raise L3BError(rc, "cannot open MP3 stream")
        self._h = h
        return self

    def _handle(self):
        if not self._h:
            raise L3BError(E_PARAM, "stream is not open")
        return self._h

    def close(self):
        if self._h:
            self._L.l3b_stream_close(self._h)
            self._h = None
        if self._own_ctx and self._ctx is not None:
            self._ctx.close()
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def getNumChannels(self) -> int:
        return self._L.l3b_stream_num_channels(self._handle())

    def getLengthInFrames(self) -> int:
        return self._L.l3b_stream_length_frames(self._handle())

    def getSamplerate(self) -> float:
        return self._L.l3b_stream_samplerate(self._handle())

    def isError(self) -> bool:
        return bool(self._L.l3b_stream_is_error(self._h))

    def errorMessage(self) -> str:
        return (self._L.l3b_stream_error_message(self._h) or b"").decode()

    def readSamplesFloat(self, frames: int) -> np.ndarray:
        out = np.empty((frames, self.getNumChannels()), np.float32)
        n = self._L.l3b_stream_read_float(self._handle(), out.ctypes.data, frames)
        return out[:n]

    def readSamplesDouble(self, frames: int) -> np.ndarray:
        out = np.empty((frames, self.getNumChannels()), np.float64)
        n = self._L.l3b_stream_read_double(self._handle(), out.ctypes.data, frames)
        return out[:n]

    def seekPosition(self, frame: int) -> bool:
        return bool(self._L.l3b_stream_seek(self._handle(), frame))

    def tellPosition(self) -> int:
        return self._L.l3b_stream_tell(self._handle())
