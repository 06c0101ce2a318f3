"""Independent MP3 decoder for cross-checking the oracle (TEST INFRASTRUCTURE ONLY).

The image ships no D compiler, so the reference decoder itself cannot run here.  It does ship
FFmpeg's libavcodec (bundled inside opencv_python_headless.libs), whose `mp3float` decoder is an implementation of
ISO 11172-3 / 13818-3 Layer III that shares no code with minimp3.  This module drives it through ctypes, one frame per
packet, and returns planar float PCM.  It is used by tests/test_oracle_vs_ffmpeg.py only; nothing in the product
imports it.

Struct offsets used (libavcodec 62 / libavutil 60, x86-64): AVPacket.data @24, AVPacket.size @32;
AVFrame.data[8] @0, AVFrame.nb_samples @112, AVFrame.format @116.  The test skips when the library is absent or its
major version differs.
"""
from __future__ import annotations

import ctypes as C
import glob
import os

import numpy as np

_AV_SAMPLE_FMT_FLTP = 8


def _find(name: str):
    pats = []
    try:
        import cv2  # noqa: F401  (only to locate site-packages; not required)
        pats.append(os.path.join(os.path.dirname(os.path.dirname(cv2.__file__)), "opencv_python_headless.libs", f"{name}-*.so*"))
    except Exception:
        pass
    import sysconfig
    pats.append(os.path.join(sysconfig.get_paths()["purelib"], "opencv_python_headless.libs", f"{name}-*.so*"))
    for p in pats:
        hits = sorted(glob.glob(p))
        if hits:
            return hits[0]
    return None


_libs = None


def available() -> bool:
    try:
        return _load() is not None
    except OSError:
        return False


def _load():
    global _libs
    if _libs is not None:
        return _libs
    pu, pc = _find("libavutil"), _find("libavcodec")
    if not pu or not pc:
        return None
    U = C.CDLL(pu, mode=C.RTLD_GLOBAL)
    A = C.CDLL(pc, mode=C.RTLD_GLOBAL)
    A.avcodec_version.restype = C.c_uint
    if (A.avcodec_version() >> 16) != 62:
        return None
    A.avcodec_find_decoder_by_name.restype = C.c_void_p
    A.avcodec_find_decoder_by_name.argtypes = [C.c_char_p]
    A.avcodec_alloc_context3.restype = C.c_void_p
    A.avcodec_alloc_context3.argtypes = [C.c_void_p]
    A.avcodec_open2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    A.avcodec_free_context.argtypes = [C.c_void_p]
    A.av_packet_alloc.restype = C.c_void_p
    A.av_packet_free.argtypes = [C.c_void_p]
    A.avcodec_send_packet.argtypes = [C.c_void_p, C.c_void_p]
    A.avcodec_receive_frame.argtypes = [C.c_void_p, C.c_void_p]
    U.av_frame_alloc.restype = C.c_void_p
    U.av_frame_free.argtypes = [C.c_void_p]
    U.av_frame_unref.argtypes = [C.c_void_p]
    _libs = (U, A)
    return _libs


def decode_frames(data: bytes, frame_offsets, frame_sizes, decoder: bytes = b"mp3float") -> np.ndarray:
    """Decode the given frames (byte offset + size each, in stream order) with FFmpeg's mp3float (or mp2float / mp1float).
    Returns float32 [total_samples_per_channel, channels]."""
    libs = _load()
    if libs is None:
        raise RuntimeError("libavcodec 62 not found")
    U, A = libs
    codec = A.avcodec_find_decoder_by_name(decoder)
    if not codec:
        raise RuntimeError(f"{decoder.decode()} decoder missing from libavcodec")
    ctx = C.c_void_p(A.avcodec_alloc_context3(codec))
    if A.avcodec_open2(ctx, codec, None) < 0:
        raise RuntimeError("avcodec_open2 failed")
    pkt = C.c_void_p(A.av_packet_alloc())
    frm = C.c_void_p(U.av_frame_alloc())
    pad = 64
    buf = C.create_string_buffer(data + b"\0" * pad, len(data) + pad)
    base = C.addressof(buf)
    out = []
    try:
        for off, size in zip(frame_offsets, frame_sizes):
            C.c_void_p.from_address(pkt.value + 24).value = base + int(off)
            C.c_int.from_address(pkt.value + 32).value = int(size)
            rc = A.avcodec_send_packet(ctx, pkt)
            if rc < 0:
                raise RuntimeError(f"avcodec_send_packet failed: {rc}")
            while A.avcodec_receive_frame(ctx, frm) == 0:
                n = C.c_int.from_address(frm.value + 112).value
                fmt = C.c_int.from_address(frm.value + 116).value
                if fmt != _AV_SAMPLE_FMT_FLTP:
                    raise RuntimeError(f"unexpected sample format {fmt}")
                planes = []
                for ch in range(2):
                    p = C.c_void_p.from_address(frm.value + 8 * ch).value
                    if not p:
                        break
                    planes.append(np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n,)).copy())
                out.append(np.stack(planes, axis=1))
                U.av_frame_unref(frm)
    finally:
        U.av_frame_free(C.byref(frm))
        A.av_packet_free(C.byref(pkt))
        A.avcodec_free_context(C.byref(ctx))
    return np.concatenate(out) if out else np.zeros((0, 2), np.float32)


def split_frames(data: bytes):
    """Frame offsets/sizes of a clean CBR/VBR Layer III stream (ID3v2 skipped), by walking headers (ISO 11172-3 2.4.2.3)."""
    br1 = [0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320]
    br2 = [0, 8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160]
    hz_tab = [44100, 48000, 32000]
    i = 0
    if data[:3] == b"ID3":
        i = 10 + ((data[6] & 0x7F) << 21 | (data[7] & 0x7F) << 14 | (data[8] & 0x7F) << 7 | (data[9] & 0x7F))
    offs, sizes = [], []
    while i + 4 <= len(data):
        h = data[i:i + 4]
        if h[0] != 0xFF or (h[1] & 0xE0) != 0xE0:
            break
        ver = (h[1] >> 3) & 3          # 3 = MPEG-1, 2 = MPEG-2, 0 = MPEG-2.5
        bri, sri, padb = h[2] >> 4, (h[2] >> 2) & 3, (h[2] >> 1) & 1
        if bri in (0, 15) or sri == 3 or ver == 1:
            break
        hz = hz_tab[sri] >> (0 if ver == 3 else (1 if ver == 2 else 2))
        kbps = (br1 if ver == 3 else br2)[bri]
        size = (144 if ver == 3 else 72) * kbps * 1000 // hz + padb
        if i + size > len(data):
            break
        offs.append(i)
        sizes.append(size)
        i += size
    return offs, sizes
