/**
  What a maintainer adds to source/audioformats/stream.d to route the MP3 arms through the B200 path
  (see INTEGRATION.md).  Selected by the DUB version `decodeMP3GPU`; the stock CPU arms stay untouched.

  NOT COMPILED HERE (no D toolchain in the build image).  The same control flow is implemented and tested in
  C++ (audio_formats_b200/csrc/l3_stream.cpp); this file delegates the stream layer to that implementation
  through the C-ABI and adds the batch entry point.
*/
module audioformats.audiostream_mp3gpu;

import core.stdc.stdlib : malloc, free;
import audioformats.l3b200;

nothrow @nogc:

/// One per process and GPU; created lazily by the first MP3 stream.  Not thread-safe, like the library
/// (stream.d:31-33): one host thread per context.
struct Mp3GpuContext
{
nothrow @nogc:
    l3b_ctx_t* ctx;

    bool ensure(int device = 0)
    {
        if (ctx !is null) return true;
        return l3b_ctx_create(device, &ctx) == L3B_OK;   // fails (no CPU fallback) when there is no CUDA device
    }

    void release()
    {
        if (ctx !is null) l3b_ctx_destroy(ctx);
        ctx = null;
    }
}

__gshared Mp3GpuContext g_mp3gpu;

/// State the AudioStream keeps instead of `mp3dec_ex_t* _mp3DecoderNew` (stream.d:1384-1388).
struct Mp3GpuStream
{
nothrow @nogc:
    l3b_stream_t* handle;

    /// stream.d:1706-1749.  `data` is the whole file (openFromFile reads it, openFromMemory already has it).
    bool open(const(ubyte)[] data, out float sampleRate, out int numChannels, out long lengthInFrames)
    {
        if (!g_mp3gpu.ensure()) return false;
        if (l3b_stream_open_memory(g_mp3gpu.ctx, data.ptr, data.length, &handle) != L3B_OK) return false;
        sampleRate = l3b_stream_samplerate(handle);
        numChannels = l3b_stream_num_channels(handle);
        lengthInFrames = l3b_stream_length_frames(handle);
        return true;
    }

    /// The same through the reference's own I/O thunks (mp3_io_read / mp3_io_seek, stream.d:2243-2254, over IOCallbacks,
    /// io.d:16-26): what startDecoding passes to mp3dec_ex_open_cb today (stream.d:1728) goes to the library unchanged;
    /// the library drains the callbacks into memory itself (it decodes ahead in large windows and seeks freely).
    bool openCallbacks(l3b_read_cb readThunk, l3b_seek_cb seekThunk, void* decoderContext,
                       out float sampleRate, out int numChannels, out long lengthInFrames)
    {
        if (!g_mp3gpu.ensure()) return false;
        if (l3b_stream_open_callbacks(g_mp3gpu.ctx, readThunk, seekThunk, decoderContext, &handle) != L3B_OK) return false;
        sampleRate = l3b_stream_samplerate(handle);
        numChannels = l3b_stream_num_channels(handle);
        lengthInFrames = l3b_stream_length_frames(handle);
        return true;
    }

    /// stream.d:537-551
    int readSamplesFloat(float* outData, int frames) { return l3b_stream_read_float(handle, outData, frames); }

    /// stream.d:732-739
    int readSamplesDouble(double* outData, int frames) { return l3b_stream_read_double(handle, outData, frames); }

    /// stream.d:1100-1107
    bool seekPosition(int frame) { return l3b_stream_seek(handle, frame) != 0; }

    /// stream.d:1214-1218
    int tellPosition() { return l3b_stream_tell(handle); }

    bool isError() { return l3b_stream_is_error(handle) != 0; }

    /// stream.d:1443-1456
    void close()
    {
        if (handle !is null) l3b_stream_close(handle);
        handle = null;
    }
}

/// The batch entry point the GPU path adds next to AudioStream: decode many in-memory MP3 files in one go.
/// outPcm[i] must hold lengthInFrames(i) * channels(i) floats (query with mp3BatchLengths first).
/// Returns 0 or a negative L3B_E_* code; status[i] carries per-file results.
int decodeMP3Batch(const(ubyte)[][] files, float*[] outPcm, int[] status, int device = 0)
{
    if (files.length == 0 || outPcm.length != files.length) return L3B_E_PARAM;
    if (!g_mp3gpu.ensure(device)) return L3B_E_NOGPU;
    auto scans = cast(l3b_scan_t**) malloc(files.length * (l3b_scan_t*).sizeof);
    if (scans is null) return L3B_E_MEMORY;
    scope(exit) free(scans);
    size_t made = 0;
    scope(exit) foreach (k; 0 .. made) l3b_scan_free(scans[k]);
    foreach (i, f; files)
    {
        int rc = l3b_scan_memory(f.ptr, f.length, &scans[i]);   // frame sync, side info, reservoir slicing
        if (rc != L3B_OK) return rc;
        made = i + 1;
    }
    return l3b_decode_scans(g_mp3gpu.ctx, scans, cast(uint) files.length, outPcm.ptr, status.length ? status.ptr : null);
}

/// Sizes for decodeMP3Batch: interleaved samples each file will deliver.
int mp3BatchLengths(const(ubyte)[][] files, ulong[] samples, int[] channels, int[] sampleRates)
{
    foreach (i, f; files)
    {
        l3b_scan_t* s;
        int rc = l3b_scan_memory(f.ptr, f.length, &s);
        if (rc != L3B_OK) return rc;
        samples[i] = l3b_scan_delivered_samples(s);
        channels[i] = l3b_scan_channels(s);
        sampleRates[i] = l3b_scan_samplerate(s);
        l3b_scan_free(s);
    }
    return L3B_OK;
}

/// The throughput path: many in-memory files -> PCM in one (ideally page-locked: l3b_host_alloc_near) buffer, pipelined in
/// waves over one or more GPUs inside the library (prepass threads, per-GPU lanes, blocking waits).  `s16` selects 16-bit
/// delivery (the un-dithered conversion of wav.d:475-700), which halves the device -> host traffic that bounds the path.
/// results[i] tells where stream i landed (element offset, frames, channels, rate) or why it could not be decoded.
struct Mp3BatchPipeline
{
nothrow @nogc:
    l3b_pipeline_t* handle;

    bool open(const(int)[] devices, bool s16 = true, int lanes = 4, int waveStreams = 16)
    {
        l3b_pipeline_opts_t o;
        o.lanes = lanes; o.wave_streams = waveStreams; o.scan_threads = 0; o.flags = s16 ? L3B_OUT_S16 : 0;
        return l3b_pipeline_create(devices.ptr, cast(int) devices.length, &o, &handle) == L3B_OK;
    }

    int decode(const(ubyte*)[] data, const(size_t)[] sizes, void* outPcm, ulong capacityElems, l3b_stream_result_t[] results)
    {
        ulong used;
        return l3b_pipeline_decode(handle, data.ptr, sizes.ptr, cast(uint) data.length, outPcm, capacityElems, results.ptr, &used);
    }

    void close() { if (handle !is null) l3b_pipeline_destroy(handle); handle = null; }
}

/// Batch decode with the prepass on the GPU as well (frame walk, side info, reservoir, main-data gather as kernels for
/// well-formed files; anything else takes the host prepass inside the same call).
int decodeMP3BatchRaw(const(ubyte*)[] data, const(size_t)[] sizes, void*[] outPcm, bool s16 = false, int device = 0)
{
    if (!g_mp3gpu.ensure(device)) return L3B_E_NOGPU;
    l3b_raw_t* raw;
    int rc = l3b_raw_open(g_mp3gpu.ctx, data.ptr, sizes.ptr, cast(uint) data.length, s16 ? L3B_OUT_S16 : 0, &raw);
    if (rc != L3B_OK) return rc;
    scope(exit) l3b_raw_free(raw);
    // the caller sizes outPcm[i] from l3b_raw_samples(raw, i) in a real host; here they are assumed large enough
    return l3b_raw_decode(raw, outPcm.ptr);
}
