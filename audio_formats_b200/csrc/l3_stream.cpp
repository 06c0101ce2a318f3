// l3_stream.cpp -- C-ABI layer 2: the AudioStream surface for MP3 (stream.d MP3 arms) on top of the
// GPU shim.  Control flow mirrors mp3dec_ex_read / mp3dec_ex_seek (minimp3_ex.d:662-888); the only
// difference is WHERE samples are computed: frames are walked ahead on the host (l3_host.cpp) and
// their granules are decoded in one GPU batch per decode-ahead window.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <deque>
#include <new>
#include <string>
#include <vector>

#include "l3_host.hpp"

using namespace l3b;

namespace l3b {   // l3_ctx.cu: one recycled stream workspace per context
l3b_resident_t* ctx_take_spare_workspace(l3b_ctx_t* c);
void ctx_give_spare_workspace(l3b_ctx_t* c, l3b_resident_t* r);
}

namespace {
// Frames walked (and decoded in one batch) per refill.  Large on purpose: a refill costs one upload, four launches and one
// download whatever its size, and 64 frames are two warps' worth of work on a 148-SM part.  2,048 frames are 53 s of
// 44.1 kHz audio (19 MB of float stereo PCM in the cache): the transcode loop of a typical file is one or two refills.
constexpr int kDecodeAheadFrames = 2048;
}

struct l3b_stream {
    l3b_ctx_t* ctx = nullptr;
    std::vector<uint8_t> data;
    OpenInfo oi;
    int channels = 0, hz = 0, sr_idx = 0, mpeg1 = 0, layer = 3;
    uint32_t gran = 576;   // PCM frames per granule: 576 (Layer III), 384 (Layer I / II: 12 slots x 32 subbands)
    int64_t length_frames = 0;
    Reader* reader = nullptr;
    // mp3dec_ex_t sample bookkeeping (minimp3_ex.d:79-86)
    uint64_t cur_sample = 0;
    int to_skip = 0, last_error = 0;
    int buffer_samples = 0, buffer_consumed = 0;
    uint32_t head_granule = 0;  // first granule of the frame currently in "dec.buffer"
    // decode run since the last restart
    Program prog;
    std::deque<Reader::Frame> look;  // walked but not yet consumed frames
    bool input_ended = false;
    std::vector<float> cache;        // PCM of granules [cache_g0, cache_g1) of this run
    uint32_t cache_g0 = 0, cache_g1 = 0;
    bool error = false;
    std::string err;
    l3b_resident_t* workspace = nullptr;  // device buffers recycled across decode-ahead windows

    ~l3b_stream() {
        if (workspace) ctx_give_spare_workspace(ctx, workspace);   // the next stream of this context starts with these buffers
        delete reader;
    }
};

static void stream_restart(l3b_stream* s, uint64_t offset) {
    s->reader->restart(offset);
    s->buffer_samples = s->buffer_consumed = 0;
    s->last_error = 0;
    s->prog.clear();
    s->look.clear();
    s->input_ended = false;
    s->cache.clear();
    s->cache_g0 = s->cache_g1 = 0;
    s->head_granule = 0;
}

// Decode granules [cache_g1, prog.granules) of the current run on the GPU and append them to the cache.
static int decode_pending(l3b_stream* s) {
    const uint32_t g_new0 = s->cache_g1, g_new1 = s->prog.granules;
    if (g_new1 <= g_new0) return 0;
    const int nch = s->channels;
    // slice: two granules of halo, widened to a frame boundary so granule 1 can see granule 0's scalefactors
    uint32_t seg0 = g_new0 >= 2 ? g_new0 - 2 : 0;
    if (s->layer == 3 && seg0 > 0 && (s->prog.descs[(size_t)seg0 * nch].w2 >> 31)) seg0--;
    std::vector<l3b_grch_desc_t> descs(s->prog.descs.begin() + (size_t)seg0 * nch, s->prog.descs.begin() + (size_t)g_new1 * nch);
    uint32_t min_bit = 0xFFFFFFFFu;
    for (auto& d : descs) min_bit = std::min(min_bit, d.bit_start);
    // granule 1 may reach back to granule 0's scalefactor bits, which precede it in the blob: already inside the slice
    const size_t byte0 = ((size_t)(min_bit >> 3)) & ~(size_t)15;
    for (auto& d : descs) d.bit_start -= (uint32_t)(byte0 * 8);
    std::vector<uint8_t> blob(s->prog.blob.begin() + byte0, s->prog.blob.end());
    blob.resize(((blob.size() + 15) & ~(size_t)15) + 16, 0);

    l3b_stream_desc_t sd{};
    sd.maindata_off = 0;
    sd.maindata_bytes = (uint32_t)blob.size() - 16;
    sd.n_granules = g_new1 - seg0;
    sd.first_grch = 0;
    sd.pcm_off = 0;
    sd.pcm_skip = (uint64_t)(g_new0 - seg0) * s->gran * nch;
    sd.pcm_count = (uint64_t)(g_new1 - g_new0) * s->gran * nch;
    sd.nch = (uint8_t)nch;
    sd.sr_idx = (uint8_t)s->sr_idx;
    sd.mpeg1 = (uint8_t)s->mpeg1;
    sd.layer = (uint8_t)(s->layer == 3 ? 0 : s->layer);

    // drop cached granules that precede the frame still being consumed
    if (s->head_granule > s->cache_g0) {
        uint32_t drop = std::min(s->head_granule, s->cache_g1) - s->cache_g0;
        s->cache.erase(s->cache.begin(), s->cache.begin() + (size_t)drop * s->gran * nch);
        s->cache_g0 += drop;
    }
    const size_t old = s->cache.size();
    s->cache.resize(old + sd.pcm_count);

    l3b_batch_t b{};
    b.maindata = blob.data();
    b.maindata_bytes = blob.size();
    b.grch = descs.data();
    b.n_grch = descs.size();
    b.streams = &sd;
    b.n_streams = 1;
    b.pcm = s->cache.data() + old;
    b.pcm_floats = sd.pcm_count;
    if (!s->workspace) s->workspace = ctx_take_spare_workspace(s->ctx);
    int rc = l3b_batch_upload_reuse(s->ctx, &b, &s->workspace);
    if (!rc) rc = l3b_batch_run(s->ctx, s->workspace);
    if (!rc) rc = l3b_batch_download(s->ctx, s->workspace, b.pcm, 0, b.pcm_floats);
    if (rc) {
        s->cache.resize(old);
        return rc;
    }
    s->cache_g1 = g_new1;
    return 0;
}

// Walk the next window of frames and decode their granules.
static int refill(l3b_stream* s) {
    for (int k = 0; k < kDecodeAheadFrames && !s->input_ended; k++) {
        Reader::Frame f = s->reader->next(s->oi, &s->prog);
        s->look.push_back(f);
        if (f.end_of_input || f.format_change) s->input_ended = true;
    }
    return decode_pending(s);
}

static const float* frame_pcm(const l3b_stream* s, uint32_t first_granule) {
    return s->cache.data() + (size_t)(first_granule - s->cache_g0) * s->gran * s->channels;
}

// mp3dec_ex_read (minimp3_ex.d:787-888), callback-I/O arm
static long stream_read_samples(l3b_stream* s, float* buf, size_t samples) {
    const size_t requested = samples;
    const uint64_t detected = s->oi.detected_samples;
    if (detected && s->cur_sample >= detected) return 0;
    if (s->last_error) return 0;
    s->reader->begin_call();
    if (s->buffer_consumed < s->buffer_samples) {
        size_t to_copy = std::min<size_t>((size_t)(s->buffer_samples - s->buffer_consumed), samples);
        if (detected && s->cur_sample + to_copy >= detected) to_copy = (size_t)(detected - s->cur_sample);
        s->cur_sample += to_copy;
        memcpy(buf, frame_pcm(s, s->head_granule) + s->buffer_consumed, to_copy * sizeof(float));
        buf += to_copy;
        s->buffer_consumed += (int)to_copy;
        samples -= to_copy;
    }
    while (samples) {
        if (detected && s->cur_sample >= detected) break;
        if (s->look.empty()) {
            if (s->input_ended) break;
            int rc = refill(s);
            if (rc) {
                // Sticky, like a decode error of the reference (minimp3_ex.d:796): the frames just walked have no PCM
                // behind them, so they must not be consumed by a later read.  A seek restarts the run and clears it.
                s->look.clear();
                s->input_ended = true;
                s->last_error = rc;
                s->error = true;
                s->err = std::string("GPU decode failed: ") + l3b_last_error(s->ctx);
                return rc;
            }
            if (s->look.empty()) break;
        }
        Reader::Frame f = s->look.front();
        s->look.pop_front();
        if (f.end_of_input) break;
        s->buffer_consumed = 0;
        if (f.format_change) {
            s->buffer_samples = 0;
            s->last_error = L3B_E_DECODE;
            break;
        }
        s->buffer_samples = f.samples;
        s->head_granule = f.first_granule;
        if (s->buffer_samples) {
            if (s->to_skip) {
                int skip = std::min(s->buffer_samples, s->to_skip);
                s->buffer_consumed += skip;
                s->to_skip -= skip;
            }
            size_t to_copy = std::min<size_t>((size_t)(s->buffer_samples - s->buffer_consumed), samples);
            if (detected && s->cur_sample + to_copy >= detected) to_copy = (size_t)(detected - s->cur_sample);
            s->cur_sample += to_copy;
            memcpy(buf, frame_pcm(s, s->head_granule) + s->buffer_consumed, to_copy * sizeof(float));
            buf += to_copy;
            s->buffer_consumed += (int)to_copy;
            samples -= to_copy;
        } else if (s->to_skip) {
            s->to_skip -= std::min(f.hdr_samples, s->to_skip);
        }
    }
    return (long)(requested - samples);
}

// mp3dec_idx_binary_search (minimp3_ex.d:640-660)
static size_t index_search(const std::vector<IndexEntry>& idx, uint64_t position) {
    size_t end = idx.size(), start = 0, index = 0;
    while (start <= end) {
        size_t mid = (start + end) / 2;
        if (mid >= idx.size()) {  // the reference reads one past the end here; treat it as "move left"
            if (!mid) break;
            end = mid - 1;
            continue;
        }
        if (idx[mid].sample >= position) {
            if (idx[mid].sample == position) return mid;
            if (!mid) break;
            end = mid - 1;
        } else {
            index = mid;
            start = mid + 1;
            if (start == idx.size()) break;
        }
    }
    return index;
}

// mp3dec_ex_seek (minimp3_ex.d:662-785), MP3D_SEEK_TO_SAMPLE
static int stream_seek_samples(l3b_stream* s, uint64_t position) {
    OpenInfo& oi = s->oi;
    const uint8_t* data = s->data.data();
    const size_t size = s->data.size();
    s->cur_sample = position;
    position += (uint64_t)oi.start_delay;
    uint64_t offset;
    if (position == 0) {
        offset = oi.start_offset;
        s->to_skip = 0;
    } else {
        if (!oi.index_started && oi.vbr_tag_found) {  // the length came from the VBR tag: build the index now
            oi.samples = 0;
            OpenInfo tmp = oi;
            tmp.index.clear();
            int rc = open_index(data, size, &tmp, oi.start_offset);
            if (rc) return rc;
            oi.index = std::move(tmp.index);
            oi.index_started = tmp.index_started;
            for (auto& e : oi.index) e.offset += oi.start_offset;
            oi.samples = oi.detected_samples;
        }
        if (!oi.index_started || oi.index.empty()) {
            offset = oi.start_offset;
            s->to_skip = 0;
        } else {
            size_t i = index_search(oi.index, position);
            if (i) {
                int to_fill_bytes = kMaxReservoir;
                i -= std::min<size_t>(i, (size_t)kPredecodeFrames);
                if (oi.info.layer == 3) {
                    while (i && to_fill_bytes) {  // back up until the bit reservoir is covered
                        size_t fo = (size_t)oi.index[i - 1].offset;
                        if (fo + kHdrSize > size) return L3B_E_IOERROR;
                        const uint8_t* hdr = data + fo;
                        int frame_size = Hdr(hdr).frame_bytes(oi.free_format_bytes) + Hdr(hdr).padding();
                        if (fo + (size_t)frame_size > size) return L3B_E_IOERROR;
                        BitReader bs(hdr + kHdrSize, frame_size - kHdrSize);
                        GranuleInfo gr[4];
                        if (Hdr(hdr).has_crc()) bs.get(16);
                        i--;
                        if (parse_side_info(bs, gr, hdr) < 0) break;  // not decodable: start from here
                        int frame_bytes = (bs.limit - bs.pos) / 8;
                        to_fill_bytes -= std::min(to_fill_bytes, frame_bytes);
                    }
                }
            }
            offset = oi.index[i].offset;
            s->to_skip = (int)(position - oi.index[i].sample);
            while ((i + 1) < oi.index.size() && !oi.index[i].sample && !oi.index[i + 1].sample) {
                size_t fo = (size_t)oi.index[i].offset;  // leading frames that decode nothing
                if (fo + kHdrSize > size) return L3B_E_IOERROR;
                s->to_skip += (int)Hdr(data + fo).frame_samples() * oi.info.channels;
                i++;
            }
        }
    }
    stream_restart(s, offset);
    return 0;
}

static int stream_open(l3b_ctx_t* ctx, std::vector<uint8_t>&& bytes, l3b_stream_t** out) {
    *out = nullptr;
    if (!ctx) return L3B_E_NOGPU;  // the MP3 arm needs a GPU context: no CPU fallback
    l3b_stream* s = new (std::nothrow) l3b_stream();
    if (!s) return L3B_E_MEMORY;
    s->ctx = ctx;
    s->data = std::move(bytes);
    const uint8_t* data = s->data.data();
    const size_t size = s->data.size();
    // stream.d:1706-1749: detect with a 32 KiB scratch, then open with MP3D_SEEK_TO_SAMPLE
    int rc = detect_mp3(data, size);
    if (!rc) rc = open_index(data, size, &s->oi);
    if (!rc && !s->oi.info.layer) rc = L3B_E_USER;
    if (rc) { delete s; return rc; }
    s->channels = s->oi.info.channels;
    s->hz = s->oi.info.hz;
    s->layer = s->oi.info.layer;
    s->gran = s->layer == 3 ? 576u : 384u;
    s->length_frames = s->channels ? (int64_t)(s->oi.samples / (uint64_t)s->channels) : 0;
    // sfb row / version from the first frame header the index or the start offset points at
    size_t first = (size_t)(s->oi.index.empty() ? s->oi.start_offset : s->oi.index[0].offset);
    if (first + kHdrSize <= size && Hdr(data + first).valid()) {
        s->sr_idx = Hdr(data + first).sfb_row();
        s->mpeg1 = Hdr(data + first).mpeg1() ? 1 : 0;
    }
    s->reader = new Reader(s->data.data(), s->data.size());
    s->to_skip = s->oi.to_skip;
    stream_restart(s, s->oi.start_offset);
    *out = s;
    return 0;
}

extern "C" {

int l3b_stream_open_memory(l3b_ctx_t* ctx, const uint8_t* data, size_t size, l3b_stream_t** out) {
    if (!data || !out) return L3B_E_PARAM;
    try {
        return stream_open(ctx, std::vector<uint8_t>(data, data + size), out);
    } catch (const std::bad_alloc&) {
        return L3B_E_MEMORY;
    }
}

int l3b_stream_open_file(l3b_ctx_t* ctx, const char* path, l3b_stream_t** out) {
    if (!path || !out) return L3B_E_PARAM;
    FILE* f = fopen(path, "rb");
    if (!f) return L3B_E_IOERROR;
    std::vector<uint8_t> bytes;
    try {
        fseek(f, 0, SEEK_END);
        long n = ftell(f);
        fseek(f, 0, SEEK_SET);
        if (n < 0) { fclose(f); return L3B_E_IOERROR; }
        bytes.resize((size_t)n);
        size_t got = n ? fread(bytes.data(), 1, (size_t)n, f) : 0;
        fclose(f);
        if (got != (size_t)n) return L3B_E_IOERROR;
        return stream_open(ctx, std::move(bytes), out);
    } catch (const std::bad_alloc&) {
        return L3B_E_MEMORY;
    }
}

// mp3dec_io_t-shaped open (minimp3_ex.d:61-71, mp3dec_ex_open_cb :929; thunks stream.d:2243-2254): the GPU path decodes
// ahead in large windows and seeks freely, so the callbacks are drained into memory once, here, instead of by the caller.
int l3b_stream_open_callbacks(l3b_ctx_t* ctx, l3b_read_cb read, l3b_seek_cb seek, void* user, l3b_stream_t** out) {
    if (!read || !out) return L3B_E_PARAM;
    try {
        if (seek && seek(0, user) != 0) return L3B_E_IOERROR;
        std::vector<uint8_t> bytes;
        for (;;) {
            const size_t at = bytes.size();
            bytes.resize(at + kIoSize);
            const size_t got = read(bytes.data() + at, kIoSize, user);
            if (got > kIoSize) return L3B_E_IOERROR;
            bytes.resize(at + got);
            if (got != kIoSize) break;   // short read = end of input (minimp3_ex.d:831-836)
        }
        return stream_open(ctx, std::move(bytes), out);
    } catch (const std::bad_alloc&) {
        return L3B_E_MEMORY;
    }
}

void l3b_stream_close(l3b_stream_t* s) { delete s; }
int l3b_stream_num_channels(const l3b_stream_t* s) { return s ? s->channels : 0; }
int64_t l3b_stream_length_frames(const l3b_stream_t* s) { return s ? s->length_frames : 0; }
float l3b_stream_samplerate(const l3b_stream_t* s) { return s ? (float)s->hz : 0.0f; }
int l3b_stream_is_error(const l3b_stream_t* s) { return (!s || s->error) ? 1 : 0; }   // a stream starts errored (stream.d:1379)
const char* l3b_stream_error_message(const l3b_stream_t* s) { return s ? s->err.c_str() : "stream is not open"; }

// stream.d:537-551
int l3b_stream_read_float(l3b_stream_t* s, float* out, int frames) {
    if (!s || !out || frames < 0) return 0;
    long r;
    try {
        r = stream_read_samples(s, out, (size_t)frames * (size_t)s->channels);
    } catch (const std::bad_alloc&) {
        r = L3B_E_MEMORY;
    }
    if (r < 0) {
        s->error = true;
        if (s->err.empty()) s->err = "Decoding error";  // kErrorDecodingError, internals.d
        return 0;
    }
    return (int)(r / s->channels);
}

// stream.d:732-739: decode to a float buffer, then widen
int l3b_stream_read_double(l3b_stream_t* s, double* out, int frames) {
    if (!s || !out || frames < 0) return 0;
    std::vector<float> tmp((size_t)frames * (size_t)s->channels);
    int n = l3b_stream_read_float(s, tmp.data(), frames);
    for (size_t i = 0; i < (size_t)n * (size_t)s->channels; i++) out[i] = tmp[i];
    return n;
}

// stream.d:1100-1107
int l3b_stream_seek(l3b_stream_t* s, int frame) {
    if (!s || frame < 0 || frame > s->length_frames) return 0;
    return stream_seek_samples(s, (uint64_t)frame * (uint64_t)s->channels) == 0 ? 1 : 0;
}

// stream.d:1214-1218
int l3b_stream_tell(const l3b_stream_t* s) { return (s && s->channels) ? (int)(s->cur_sample / (uint64_t)s->channels) : 0; }

}  // extern "C"
