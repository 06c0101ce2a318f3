// l3_host.cpp -- host prepass (see l3_host.hpp).  No sample arithmetic happens here.
#include "l3_host.hpp"

#include <algorithm>
#include <climits>
#include <cstdlib>
#include <cstring>
#include <new>

namespace l3b {

// ---------------------------------------------------------------------------------------------------
// tag skipping (minimp3_ex.d:93-125)
void skip_id3v1(const uint8_t* buf, size_t* psize) {
    size_t n = *psize;
    if (n >= 128 && !memcmp(buf + n - 128, "TAG", 3)) {
        n -= 128;
        if (n >= 227 && !memcmp(buf + n - 227, "TAG+", 4)) n -= 227;
    }
    if (n > 32 && !memcmp(buf + n - 32, "APETAGEX", 8)) {
        n -= 32;
        const uint8_t* t = buf + n + 8 + 4;
        uint32_t tag_size = ((uint32_t)t[3] << 24) | ((uint32_t)t[2] << 16) | ((uint32_t)t[1] << 8) | t[0];
        if (n >= tag_size) n -= tag_size;
    }
    *psize = n;
}

int skip_id3v2(const uint8_t* buf, size_t size, size_t* out) {
    *out = 0;
    if (size >= (size_t)kId3DetectSize && !memcmp(buf, "ID3", 3) &&
        !((buf[5] & 15) || (buf[6] & 0x80) || (buf[7] & 0x80) || (buf[8] & 0x80) || (buf[9] & 0x80))) {
        size_t n = (size_t)(((buf[6] & 0x7f) << 21) | ((buf[7] & 0x7f) << 14) | ((buf[8] & 0x7f) << 7) | (buf[9] & 0x7f)) + 10;
        if (buf[5] & 16) n += 10;  // footer
        *out = n;
        return 1;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// mp3dec_decode_frame control flow (minimp3.d:1492-1581)
int FrameWalker::step(const uint8_t* mp3, int mp3_bytes, FrameInfo* info, Program* prog) {
    int i = 0, frame_size = 0;
    if (mp3_bytes > 4 && header[0] == 0xff && hdr_compatible(header, mp3)) {
        frame_size = Hdr(mp3).frame_bytes(free_format_bytes) + Hdr(mp3).padding();
        if (frame_size != mp3_bytes && (frame_size + kHdrSize > mp3_bytes || !hdr_compatible(mp3, mp3 + frame_size)))
            frame_size = 0;
    }
    if (!frame_size) {
        // memset(dec, 0, sizeof(mp3dec_t)): overlap, qmf history and reservoir all go to zero
        memset(header, 0, 4);
        free_format_bytes = 0;
        reserv = 0;
        pending_reset = true;
        i = find_frame(mp3, mp3_bytes, &free_format_bytes, &frame_size);
        if (!frame_size || i + frame_size > mp3_bytes) {
            info->frame_bytes = i;
            return 0;
        }
    }
    const uint8_t* hdr = mp3 + i;
    memcpy(header, hdr, kHdrSize);
    Hdr H(hdr);
    info->frame_bytes = i + frame_size;
    info->frame_offset = i;
    info->channels = H.channels();
    info->hz = (int)H.sample_rate_hz();
    info->layer = H.layer();
    info->bitrate_kbps = (int)H.bitrate_kbps();

    BitReader bs(hdr + kHdrSize, frame_size - kHdrSize);
    if (H.has_crc()) bs.get(16);  // skipped, never verified (minimp3.d:1533-1536)

    if (info->layer != 3) {
        // Layer I / II (minimp3.d:1557-1579): no reservoir, one frame = one (Layer I) or three (Layer II) "granules" of 12 slots
        // x 32 subbands.  The frame is dropped, and the decoder state with it, when its bits run past its end.
        if (!l12_frame_fits(hdr, frame_size)) {
            init();
            return 0;
        }
        const int nch = info->channels, parts = H.layer1() ? 1 : 3;
        if (prog) {
            const uint64_t start_bit = (uint64_t)prog->blob.size() * 8u;   // of the frame body (what follows the header)
            for (int part = 0; part < parts; part++) {
                for (int ch = 0; ch < nch; ch++) {
                    l3b_grch_desc_t d;
                    d.bit_start = (uint32_t)std::min<uint64_t>(start_bit, 0xFFFFFFFFull);
                    d.w1 = (uint32_t)hdr[1] | ((uint32_t)hdr[2] << 8) | ((uint32_t)hdr[3] << 16);   // version / layer / crc, rate, mode
                    d.w2 = (uint32_t)part | (part ? 0x80000000u : 0u);
                    d.w3 = pending_reset ? 0x80000000u : 0u;
                    prog->descs.push_back(d);
                }
                pending_reset = false;
                prog->granules++;
            }
            prog->blob.insert(prog->blob.end(), hdr + kHdrSize, hdr + frame_size);
        } else {
            pending_reset = false;
        }
        return (int)H.frame_samples();
    }
    GranuleInfo gr[4];
    int mdb = parse_side_info(bs, gr, hdr);
    if (mdb < 0 || bs.pos > bs.limit) {
        init();  // header[0] = 0: the next call takes the reset path
        return 0;
    }
    // L3_restore_reservoir (minimp3.d:1186-1194)
    const int payload = (bs.limit - bs.pos) / 8;
    const uint8_t* payload_ptr = hdr + kHdrSize + bs.pos / 8;
    const bool success = reserv >= mdb;
    const int nch = info->channels;
    const int ngr = H.mpeg1() ? 2 : 1;
    int consumed_bits = 0;
    if (success) {
        if (prog) {
            uint64_t start_bit = ((uint64_t)prog->blob.size() - (uint64_t)mdb) * 8u;
            for (int g = 0; g < ngr; g++) {
                for (int ch = 0; ch < nch; ch++) {
                    const GranuleInfo& q = gr[g * nch + ch];
                    uint64_t b = start_bit + (uint64_t)consumed_bits;
                    if (b > 0xFFFFFFFFull) b = 0xFFFFFFFFull;  // > 512 MiB of main data in one run: refused upstream
                    prog->descs.push_back(pack_desc(q, (uint32_t)b, hdr[3], g == 1, pending_reset));
                    consumed_bits += q.part_23_length;
                }
                pending_reset = false;
                prog->granules++;
            }
        } else {
            for (int k = 0; k < ngr * nch; k++) consumed_bits += gr[k].part_23_length;
            pending_reset = false;
        }
        // L3_save_reservoir (minimp3.d:1170-1184): keep the unread tail, newest 511 bytes at most
        int remains = (mdb + payload) - (consumed_bits + 7) / 8;
        reserv = std::min(remains, kMaxReservoir);
        if (reserv < 0) reserv = 0;
    } else {
        // nothing decoded, nothing reset; the reservoir keeps sliding (bs.pos == 0 in L3_save_reservoir)
        reserv = std::min(reserv + payload, kMaxReservoir);
    }
    if (prog) prog->blob.insert(prog->blob.end(), payload_ptr, payload_ptr + payload);
    return success ? (int)Hdr(header).frame_samples() : 0;
}

// ---------------------------------------------------------------------------------------------------
// detection (stream.d:1706-1721 -> minimp3_ex.d:197-233 over MemoryContext I/O)
int detect_mp3(const uint8_t* data, size_t size) {
    const size_t buf_size = kBufSize * 2;
    size_t filled = std::min<size_t>(size, (size_t)kId3DetectSize);
    if (filled < (size_t)kId3DetectSize) return L3B_E_USER;
    size_t id3;
    if (skip_id3v2(data, filled, &id3)) return 0;
    filled = std::min(size, buf_size);
    if (filled < kBufSize) skip_id3v1(data, &filled);
    int ffb = 0, frame_size = 0;
    find_frame(data, (int)filled, &ffb, &frame_size);
    return frame_size ? 0 : L3B_E_USER;
}

// ---------------------------------------------------------------------------------------------------
// VBR tag (minimp3_ex.d:144-190)
static int check_vbrtag(const uint8_t* frame, int frame_size, uint32_t* frames, int* delay, int* padding) {
    enum { FRAMES_FLAG = 1, BYTES_FLAG = 2, TOC_FLAG = 4, VBR_SCALE_FLAG = 8 };
    BitReader bs(frame + kHdrSize, frame_size - kHdrSize);
    GranuleInfo gr[4];
    if (Hdr(frame).has_crc()) bs.get(16);
    if (parse_side_info(bs, gr, frame) < 0) return 0;
    const uint8_t* tag = frame + kHdrSize + bs.pos / 8;
    if (memcmp("Xing", tag, 4) && memcmp("Info", tag, 4)) return 0;
    int flags = tag[7];
    if (!(flags & FRAMES_FLAG)) return -1;
    tag += 8;
    *frames = ((uint32_t)tag[0] << 24) | ((uint32_t)tag[1] << 16) | ((uint32_t)tag[2] << 8) | tag[3];
    tag += 4;
    if (flags & BYTES_FLAG) tag += 4;
    if (flags & TOC_FLAG) tag += 100;
    if (flags & VBR_SCALE_FLAG) tag += 4;
    *delay = *padding = 0;
    if (*tag) {  // LAME/Lavc extension
        tag += 21;
        if (tag - frame + 14 >= frame_size) return 0;
        *delay = ((tag[0] << 4) | (tag[1] >> 4)) + (528 + 1);
        *padding = (((tag[1] & 0xF) << 8) | tag[2]) - (528 + 1);
    }
    return 1;
}

// ---------------------------------------------------------------------------------------------------
// mp3dec_iterate_cb + mp3dec_load_index over the in-memory file (minimp3_ex.d:490-621).
// The 128 KiB window is modelled by offsets only: the buffer always holds data[win .. win+filled).
int open_index(const uint8_t* data, size_t size, OpenInfo* oi, uint64_t from_offset) {
    size_t cursor = std::min<uint64_t>(from_offset, size);
    auto io_read = [&](size_t want) { size_t n = std::min(want, size - cursor); cursor += n; return n; };
    size_t win = cursor;
    size_t filled = io_read(kId3DetectSize), consumed = 0;
    uint64_t readed2 = 0;
    bool eof = false;
    FrameWalker walker;
    int buffer_samples = 0;  // dec.buffer_samples while indexing
    if (filled != (size_t)kId3DetectSize) return 0;
    size_t id3v2size;
    if (skip_id3v2(data + win, filled, &id3v2size)) {
        cursor = std::min(id3v2size, size);  // io.seek(id3v2size): absolute (memory_seek clamps, stream.d:2098-2105)
        win = cursor;
        filled = io_read(kIoSize);
        readed2 += id3v2size;
    } else {
        filled += io_read(kIoSize - kId3DetectSize);
    }
    if (filled < kBufSize) skip_id3v1(data + win, &filled);
    // mp3d_find_frame accepts a header when the next ten headers, hopping by their own frame sizes, are compatible with
    // it (minimp3.d:1436-1448).  Walking a clean stream frame by frame repeats nine of those ten checks every time.
    // The memo below remembers, for the header expected next, that the nine behind it are already known to be valid
    // and compatible, so only the tenth is looked at; anything unusual (free format, window end, a mismatch, a frame
    // that does not start where expected) drops the memo and takes the full search, whose result is then the
    // reference's by construction.
    struct { bool valid = false; size_t next_abs = 0, tail_abs = 0; } memo;
    auto full_search = [&](int* ffb, int* frame_size) {
        const uint8_t* base = data + win + consumed;
        const int bytes = (int)(filled - consumed);
        int i = find_frame(base, bytes, ffb, frame_size);
        memo.valid = false;
        if (*frame_size && !Hdr(base + i).free_format()) {   // rebuild the memo: the ten hops from this header
            const uint8_t* h = base + i;
            int pos = 0, matched = 0;
            for (; matched < kMaxSyncMatches; matched++) {
                pos += Hdr(h + pos).frame_bytes(0) + Hdr(h + pos).padding();
                if (i + pos + kHdrSize > bytes || !hdr_compatible(h, h + pos)) break;
            }
            if (matched == kMaxSyncMatches) {
                memo.valid = true;
                memo.next_abs = win + consumed + (size_t)i + (size_t)*frame_size;
                memo.tail_abs = win + consumed + (size_t)i + (size_t)pos;
            }
        }
        return i;
    };
    for (;;) {
        int ffb = 0, frame_size = 0;
        int i;
        const size_t here = win + consumed, win_end = win + filled;
        bool fast = false;
        if (memo.valid && memo.next_abs == here && memo.tail_abs + kHdrSize <= win_end && here + kHdrSize < win_end) {
            const uint8_t* h = data + here;
            const uint8_t* tail = data + memo.tail_abs;
            const int fb = Hdr(h).frame_bytes(0), fb_pad = fb + Hdr(h).padding();
            const size_t q11 = memo.tail_abs + (size_t)(Hdr(tail).frame_bytes(0) + Hdr(tail).padding());
            // the eleventh header has to be inside the window (otherwise the reference's end-of-buffer rule decides: slow path)
            if (fb && here + (size_t)fb_pad <= win_end && q11 + kHdrSize <= win_end && hdr_compatible(h, data + q11)) {
                fast = true;
                i = 0;
                frame_size = fb_pad;
                memo.next_abs = here + (size_t)fb_pad;
                memo.tail_abs = q11;
            }
        }
        if (!fast) i = full_search(&ffb, &frame_size);
        if (i && !frame_size) { consumed += i; continue; }
        if (!frame_size) break;
        const uint8_t* hdr = data + win + consumed + i;
        FrameInfo fi;
        fi.channels = Hdr(hdr).channels();
        fi.hz = (int)Hdr(hdr).sample_rate_hz();
        fi.layer = Hdr(hdr).layer();
        fi.bitrate_kbps = (int)Hdr(hdr).bitrate_kbps();
        fi.frame_bytes = frame_size;
        readed2 += i;
        const size_t cb_buf_size = filled - consumed;  // what the reference passes as buf_size
        const uint64_t offset = readed2;
        // ---- mp3dec_load_index (minimp3_ex.d:566-621) ----
        bool stop = false;
        if (!oi->index_started && !oi->start_offset) {
            oi->info = fi;
            oi->start_offset = offset;
            oi->end_offset = offset + cb_buf_size;
            oi->free_format_bytes = ffb;
            if (fi.layer == 3) {
                uint32_t frames = 0;
                int delay = 0, padding = 0;
                int ret = check_vbrtag(hdr, frame_size, &frames, &delay, &padding);
                if (ret) oi->start_offset = offset + frame_size;
                if (ret > 0) {
                    padding *= fi.channels;
                    oi->start_delay = oi->to_skip = delay * fi.channels;
                    oi->samples = (uint64_t)Hdr(hdr).frame_samples() * fi.channels * (uint64_t)frames;
                    if (oi->samples >= (uint64_t)oi->start_delay) oi->samples -= oi->start_delay;
                    if (padding > 0 && oi->samples >= (uint64_t)padding) oi->samples -= padding;
                    oi->detected_samples = oi->samples;
                    oi->vbr_tag_found = 1;
                    return 0;  // MP3D_E_USER: stop the walk, not an error (minimp3_ex.d:944-945)
                } else if (ret < 0) {
                    stop = true;  // callback returns 0 without indexing this frame
                }
            }
        }
        if (!stop) {
            oi->index_started = true;
            oi->index.push_back({oi->samples, offset});
            if (!buffer_samples && oi->index.size() < 256) {
                // decode (here: walk) up to 255 first frames until one yields samples (minimp3_ex.d:613-619)
                size_t avail = size - (size_t)(hdr - data);
                FrameInfo tmp;
                buffer_samples = walker.step(hdr, (int)std::min<size_t>(std::min(cb_buf_size, avail), INT_MAX), &tmp, nullptr);
                // the reference lets mp3dec_decode_frame overwrite *info and then reads info.channels
                if (tmp.channels) fi.channels = tmp.channels;
                oi->samples += (uint64_t)buffer_samples * fi.channels;
            } else {
                oi->samples += (uint64_t)Hdr(hdr).frame_samples() * fi.channels;
            }
        }
        readed2 += frame_size;
        consumed += i + frame_size;
        if (!eof && filled - consumed < kBufSize) {
            win += consumed;
            filled -= consumed;
            consumed = 0;
            size_t want = kIoSize - filled;
            size_t got = io_read(want);
            if (got != want) eof = true;
            filled += got;
            if (eof) skip_id3v1(data + win, &filled);
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------
void Reader::restart(uint64_t off) {
    offset = off;
    cursor_ = (size_t)std::min<uint64_t>(off, size_);
    win_start_ = cursor_;
    filled_ = consumed_ = 0;
    eof_ = false;
    walker.init();
}

Reader::Frame Reader::next(const OpenInfo& oi, Program* prog) {
    Frame f;
    if (!eof_ && (filled_ - consumed_) < kBufSize) {  // keep >= 16 KiB in the window (minimp3_ex.d:821-837)
        win_start_ += consumed_;
        filled_ -= consumed_;
        consumed_ = 0;
        size_t want = kIoSize - filled_;
        size_t got = std::min(want, size_ - cursor_);
        cursor_ += got;
        if (got != want) eof_ = true;
        filled_ += got;
        if (eof_) skip_id3v1(data_ + win_start_, &filled_);
    }
    if (filled_ == consumed_) {
        f.end_of_input = true;
        return f;
    }
    const uint8_t* dec_buf = data_ + win_start_ + consumed_;
    FrameInfo fi;
    uint32_t g0 = prog ? prog->granules : 0;
    int spc = walker.step(dec_buf, (int)std::min<size_t>(filled_ - consumed_, INT_MAX), &fi, prog);
    consumed_ += fi.frame_bytes;
    f.first_granule = g0;
    if (oi.info.hz != fi.hz || oi.info.layer != fi.layer || oi.info.channels != fi.channels) {
        f.format_change = true;  // MP3D_E_DECODE, sticky (minimp3_ex.d:852-858)
        return f;
    }
    f.samples = spc * fi.channels;
    f.hdr_samples = (int)Hdr(dec_buf).frame_samples() * fi.channels;
    offset += fi.frame_bytes;
    return f;
}

// ---------------------------------------------------------------------------------------------------
int scan_stream(const uint8_t* data, size_t size, ScanResult* out) {
    if (detect_mp3(data, size) != 0) return L3B_E_USER;
    OpenInfo& oi = out->open;
    int rc = open_index(data, size, &oi);
    if (rc) return rc;
    if (!oi.info.layer) return L3B_E_USER;
    out->layer = oi.info.layer;
    out->channels = oi.info.channels;
    out->hz = oi.info.hz;
    out->length_frames = oi.info.channels ? oi.samples / oi.info.channels : 0;

    // one allocation each instead of growth by doubling: the payloads are a little less than the file, and a frame of
    // >= 24 bytes yields at most four descriptors
    out->prog.blob.reserve(size + 64);
    out->prog.descs.reserve(oi.index.size() * 4 + 16);
    Reader rd(data, size);
    rd.restart(oi.start_offset);
    int to_skip = oi.to_skip;
    uint64_t cur = 0, skipped = 0;
    bool first_hdr = true;
    for (;;) {
        if (oi.detected_samples && cur >= oi.detected_samples) break;
        Reader::Frame f = rd.next(oi, &out->prog);
        if (f.end_of_input) break;
        if (f.format_change) { out->last_error = L3B_E_DECODE; break; }
        if (first_hdr && rd.walker.header[0] == 0xff) {
            out->sr_idx = Hdr(rd.walker.header).sfb_row();
            out->mpeg1 = Hdr(rd.walker.header).mpeg1() ? 1 : 0;
            first_hdr = false;
        }
        if (f.samples) {
            int skip = std::min(f.samples, to_skip);
            to_skip -= skip;
            skipped += (uint64_t)skip;
            uint64_t to_copy = (uint64_t)(f.samples - skip);
            if (oi.detected_samples && cur + to_copy >= oi.detected_samples) to_copy = oi.detected_samples - cur;
            cur += to_copy;
        } else if (to_skip) {
            to_skip -= std::min(f.hdr_samples, to_skip);
        }
        if (out->prog.blob.size() > 0x1FFFFFF0ull) { out->last_error = L3B_E_MEMORY; break; }  // bit offsets are 32-bit
    }
    out->pcm_skip = skipped;
    out->pcm_count = cur;
    return 0;
}

}  // namespace l3b

// ---------------------------------------------------------------------------------------------------
// C-ABI, layer 2 (scan part)
extern "C" {

int l3b_scan_memory(const uint8_t* data, size_t size, l3b_scan_t** out) {
    if (!data || !out) return L3B_E_PARAM;
    *out = nullptr;
    l3b_scan* s = new (std::nothrow) l3b_scan();
    if (!s) return L3B_E_MEMORY;
    int rc;
    try {
        rc = l3b::scan_stream(data, size, &s->r);
    } catch (const std::bad_alloc&) {
        rc = L3B_E_MEMORY;
    }
    if (rc) { delete s; return rc; }
    *out = s;
    return 0;
}

void l3b_scan_free(l3b_scan_t* s) { delete s; }
int l3b_scan_channels(const l3b_scan_t* s) { return s ? s->r.channels : 0; }
int l3b_scan_samplerate(const l3b_scan_t* s) { return s ? s->r.hz : 0; }
int l3b_scan_error(const l3b_scan_t* s) { return s ? s->r.last_error : L3B_E_PARAM; }
uint64_t l3b_scan_length_frames(const l3b_scan_t* s) { return s ? s->r.length_frames : 0; }
uint64_t l3b_scan_delivered_samples(const l3b_scan_t* s) { return s ? s->r.pcm_count : 0; }
uint32_t l3b_scan_granules(const l3b_scan_t* s) { return s ? s->r.prog.granules : 0; }
uint64_t l3b_scan_maindata_bytes(const l3b_scan_t* s) { return s ? s->r.prog.blob.size() : 0; }
const uint8_t* l3b_scan_maindata(const l3b_scan_t* s) { return s ? s->r.prog.blob.data() : nullptr; }
const l3b_grch_desc_t* l3b_scan_descs(const l3b_scan_t* s) { return s ? s->r.prog.descs.data() : nullptr; }

void l3b_scan_fill_stream_desc(const l3b_scan_t* s, l3b_stream_desc_t* d) {
    if (!d) return;
    memset(d, 0, sizeof *d);
    if (!s) return;
    d->maindata_bytes = (uint32_t)s->r.prog.blob.size();
    d->n_granules = s->r.prog.granules;
    d->pcm_skip = s->r.pcm_skip;
    d->pcm_count = s->r.pcm_count;
    d->nch = (uint8_t)s->r.channels;
    d->sr_idx = (uint8_t)s->r.sr_idx;
    d->mpeg1 = (uint8_t)s->r.mpeg1;
    d->layer = (uint8_t)(s->r.layer == 3 ? 0 : s->r.layer);
}

}  // extern "C"
