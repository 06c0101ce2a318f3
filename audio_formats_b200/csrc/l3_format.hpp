// l3_format.hpp -- host-side MPEG audio frame syntax: header fields, frame sync, Layer III side info.
//
// Host logic of the drop-in path (what the D host does before calling the CUDA shim).  Behaviour
// follows /root/reference/source/audioformats/minimp3.d; each function cites the lines it mirrors.
#pragma once
#include <cstdint>
#include <cstring>

#include "../../include/l3b200.h"
#include "l12_tables.h"
#include "l3_tables_gen.h"

#ifdef __CUDACC__
#define L3B_HD __host__ __device__
#else
#define L3B_HD
#endif

namespace l3b {

constexpr int kHdrSize = 4;
constexpr int kMaxFreeFormatFrame = 2304;  // minimp3.d:53
constexpr int kMaxSyncMatches = 10;        // minimp3.d:54
constexpr int kMaxReservoir = 511;         // minimp3.d:58

// ---- header field access (minimp3.d:65-148) -----------------------------------------------------
struct Hdr {
    const uint8_t* h;
    L3B_HD explicit Hdr(const uint8_t* p) : h(p) {}
    L3B_HD bool mono() const { return (h[3] & 0xC0) == 0xC0; }
    L3B_HD bool free_format() const { return (h[2] & 0xF0) == 0; }
    L3B_HD bool has_crc() const { return !(h[1] & 1); }
    L3B_HD bool mpeg1() const { return (h[1] & 0x08) != 0; }
    L3B_HD bool not_mpeg25() const { return (h[1] & 0x10) != 0; }
    L3B_HD int layer_bits() const { return (h[1] >> 1) & 3; }
    L3B_HD int bitrate_idx() const { return h[2] >> 4; }
    L3B_HD int sr_bits() const { return (h[2] >> 2) & 3; }
    L3B_HD bool layer1() const { return (h[1] & 6) == 6; }
    L3B_HD bool frame576() const { return (h[1] & 14) == 2; }
    // 0..2 MPEG-2.5, 3..5 MPEG-2, 6..8 MPEG-1 (minimp3.d:135-138)
    L3B_HD int my_sample_rate() const { return sr_bits() + (((h[1] >> 3) & 1) + ((h[1] >> 4) & 1)) * 3; }
    L3B_HD int sfb_row() const { int v = my_sample_rate(); return v - (v != 0); }  // minimp3.d:523
    L3B_HD int channels() const { return mono() ? 1 : 2; }
    L3B_HD int layer() const { return 4 - layer_bits(); }

    // minimp3.d:232-239
    L3B_HD bool valid() const {
        return h[0] == 0xff && ((h[1] & 0xF0) == 0xf0 || (h[1] & 0xFE) == 0xe2) && layer_bits() != 0 &&
               bitrate_idx() != 15 && sr_bits() != 3;
    }
    // minimp3.d:249-257
    L3B_HD unsigned bitrate_kbps_t(const uint8_t* halfrate) const {   // halfrate: L3_HALFRATE, or its copy in device memory
        return 2u * halfrate[((mpeg1() ? 1 : 0) * 3 + (layer_bits() - 1)) * 15 + bitrate_idx()];
    }
    unsigned bitrate_kbps() const { return bitrate_kbps_t(L3_HALFRATE); }
    // minimp3.d:259-263
    L3B_HD unsigned sample_rate_hz() const {
        const unsigned hz = sr_bits() == 0 ? 44100u : (sr_bits() == 1 ? 48000u : 32000u);
        return hz >> (mpeg1() ? 0 : 1) >> (not_mpeg25() ? 0 : 1);
    }
    // minimp3.d:265-268
    L3B_HD unsigned frame_samples() const { return layer1() ? 384u : (1152u >> (frame576() ? 1 : 0)); }
    // minimp3.d:270-278
    L3B_HD int frame_bytes_t(const uint8_t* halfrate, int free_format_size) const {
        int fb = (int)(frame_samples() * bitrate_kbps_t(halfrate) * 125 / sample_rate_hz());
        if (layer1()) fb &= ~3;
        return fb ? fb : free_format_size;
    }
    int frame_bytes(int free_format_size) const { return frame_bytes_t(L3_HALFRATE, free_format_size); }
    // minimp3.d:280-283
    L3B_HD int padding() const { return (h[2] & 2) ? (layer1() ? 4 : 1) : 0; }
};

// minimp3.d:241-247: same stream family (version/layer/sample rate/free-format-ness)
L3B_HD inline bool hdr_compatible(const uint8_t* a, const uint8_t* b) {
    return Hdr(b).valid() && ((a[1] ^ b[1]) & 0xFE) == 0 && ((a[2] ^ b[2]) & 0x0C) == 0 &&
           !(Hdr(a).free_format() ^ Hdr(b).free_format());
}

// minimp3.d:1436-1448
inline bool sync_chain_ok(const uint8_t* hdr, int bytes, int frame_bytes) {
    int pos = 0;
    for (int matched = 0; matched < kMaxSyncMatches; matched++) {
        pos += Hdr(hdr + pos).frame_bytes(frame_bytes) + Hdr(hdr + pos).padding();
        if (pos + kHdrSize > bytes) return matched > 0;
        if (!hdr_compatible(hdr, hdr + pos)) return false;
    }
    return true;
}

// minimp3.d:1450-1485.  Returns the offset of the accepted frame (== bytes when none) and its size.
inline int find_frame(const uint8_t* p, int bytes, int* free_format_bytes, int* frame_size_out) {
    for (int i = 0; i < bytes - kHdrSize; i++, p++) {
        if (!Hdr(p).valid()) continue;
        int fb = Hdr(p).frame_bytes(*free_format_bytes);
        int fb_pad = fb + Hdr(p).padding();
        for (int k = kHdrSize; !fb && k < kMaxFreeFormatFrame && i + 2 * k < bytes - kHdrSize; k++) {
            if (!hdr_compatible(p, p + k)) continue;
            int cand = k - Hdr(p).padding();
            int next = cand + Hdr(p + k).padding();
            if (i + k + next + kHdrSize > bytes || !hdr_compatible(p, p + k + next)) continue;
            fb_pad = k;
            fb = cand;
            *free_format_bytes = cand;
        }
        if ((fb && i + fb_pad <= bytes && sync_chain_ok(p, bytes - i, fb)) || (!i && fb_pad == bytes)) {
            *frame_size_out = fb_pad;
            return i;
        }
        *free_format_bytes = 0;
    }
    *frame_size_out = 0;
    return bytes;
}

// ---- bit reader with the reference's end-of-buffer rule (minimp3.d:216-230) ----------------------
struct BitReader {
    const uint8_t* buf;
    int pos, limit;
    L3B_HD BitReader(const uint8_t* d, int bytes) : buf(d), pos(0), limit(bytes * 8) {}
    L3B_HD uint32_t get(int n) {
        int p = pos;
        pos += n;
        if (pos > limit) return 0;  // still advances
        if (n == 0) return 0;
        const int byte = p >> 3;
        if (byte * 8 + 40 <= limit) {   // five whole bytes in range: one shift instead of a byte loop (n <= 32)
            const uint64_t w = ((uint64_t)buf[byte] << 32) | ((uint64_t)buf[byte + 1] << 24) | ((uint64_t)buf[byte + 2] << 16) |
                               ((uint64_t)buf[byte + 3] << 8) | (uint64_t)buf[byte + 4];
            return (uint32_t)((w >> (40 - (p & 7) - n)) & ((1ull << n) - 1ull));
        }
        uint32_t v = 0;
        while (n > 0) {
            int off = p & 7, take = 8 - off < n ? 8 - off : n;
            v = (v << take) | ((buf[p >> 3] >> (8 - off - take)) & ((1u << take) - 1u));
            p += take;
            n -= take;
        }
        return v;
    }
};

// ---- Layer I / II (minimp3.d:284-470): what the host needs -- does the frame decode, i.e. does the bit position stay inside
// the frame after each of the three L12_dequantize_granule calls (minimp3.d:1571-1575)?  The scale info is parsed exactly like
// the reference parses it, including its reading of a 2-bit scfsi field for EVERY band-channel entry of a Layer II frame
// (minimp3.d:417-421 evaluates get_bits before testing the allocation; ISO 11172-3 and the upstream C read it only for
// allocated entries).  The samples themselves are only counted here; the device parser decodes them.
inline void l12_alloc_table(const uint8_t* hdr, const uint8_t (**rows)[3], int* total_bands, int* stereo_bands) {
    const Hdr H(hdr);
    const int mode = (hdr[3] >> 6) & 3;   // 3 = mono, 1 = joint stereo
    int sb = mode == 3 ? 0 : (mode == 1 ? (((hdr[3] >> 4) & 3) << 2) + 4 : 32), nbands;
    if (H.layer1()) {
        *rows = L12_ALLOC_L1; nbands = 32;
    } else if (!H.mpeg1()) {
        *rows = L12_ALLOC_L2M2; nbands = 30;
    } else {
        unsigned kbps = H.bitrate_kbps() >> (mode != 3 ? 1 : 0);
        if (!kbps) kbps = 192;   // free format
        *rows = L12_ALLOC_L2M1; nbands = 27;
        if (kbps < 56) { *rows = L12_ALLOC_L2M1_LOWRATE; nbands = H.sr_bits() == 2 ? 12 : 8; }
        else if (kbps >= 96 && H.sr_bits() != 1) nbands = 30;
    }
    *total_bands = nbands;
    *stereo_bands = sb < nbands ? sb : nbands;
}

inline bool l12_frame_fits(const uint8_t* hdr, int frame_size) {
    BitReader bs(hdr + kHdrSize, frame_size - kHdrSize);
    if (Hdr(hdr).has_crc()) bs.get(16);
    const uint8_t (*rows)[3];
    int total, stereo;
    l12_alloc_table(hdr, &rows, &total, &stereo);
    const bool l1 = Hdr(hdr).layer1();
    uint8_t bitalloc[64], scfcod[64];
    int k = 0, ba_bits = 0;
    const uint8_t* tab = L12_BITALLOC_CODE_TAB;
    for (int i = 0; i < total; i++) {
        if (i == k) { k += (*rows)[2]; ba_bits = (*rows)[1]; tab = L12_BITALLOC_CODE_TAB + (*rows)[0]; rows++; }
        uint8_t ba = tab[bs.get(ba_bits)];
        bitalloc[2 * i] = ba;
        if (i < stereo) ba = tab[bs.get(ba_bits)];
        bitalloc[2 * i + 1] = stereo ? ba : 0;
    }
    for (int i = 0; i < 2 * total; i++) {
        const uint8_t temp = l1 ? 2 : (uint8_t)bs.get(2);
        scfcod[i] = bitalloc[i] ? temp : 6;
    }
    for (int i = 0; i < 2 * total; i++) {
        const int mask = bitalloc[i] ? 4 + ((19 >> scfcod[i]) & 3) : 0;
        for (int m = 4; m; m >>= 1)
            if (mask & m) bs.get(6);
    }
    for (int i = stereo; i < total; i++) bitalloc[2 * i + 1] = 0;
    const int group = l1 ? 1 : 3;
    int per_call = 0;
    for (int i = 0; i < 2 * total; i++) {
        const int ba = bitalloc[i];
        if (!ba) continue;
        const int mod = (2 << (ba - 17)) + 1;
        per_call += 4 * (ba < 17 ? group * ba : mod + 2 - (mod >> 3));
    }
    for (int igr = 0; igr < 3; igr++) {
        bs.pos += per_call;
        if (bs.pos > bs.limit) return false;
    }
    return true;
}

// ---- Layer III side info (minimp3.d:189-196, 487-611) ---------------------------------------------
struct GranuleInfo {
    uint16_t part_23_length, big_values, scalefac_compress;
    uint8_t global_gain, block_type, mixed_block_flag, n_long_sfb, n_short_sfb;
    uint8_t table_select[3], region_count[3], subblock_gain[3];
    uint8_t preflag, scalefac_scale, count1_table, scfsi;
    const uint8_t* sfbtab;
};

// scalefactor-band width rows: kind 0 long, 1 short, 2 mixed (the host's literal tables; the device prepass passes a functor
// over their copy in device memory, DeviceTables::sfb_width)
struct HostSfbRows {
    const uint8_t* operator()(int row, int kind) const {
        return kind == 0 ? L3_SFB_LONG + row * 23 : (kind == 1 ? L3_SFB_SHORT + row * 40 : L3_SFB_MIXED + row * 40);
    }
};

// Returns main_data_begin, or -1 when the frame must be dropped (minimp3.d:545, 557, 605).
#ifdef __CUDACC__
#pragma nv_exec_check_disable   // instantiated with a host-only functor on the host and a device-side one on the device
#endif
template <class Rows>
L3B_HD inline int parse_side_info_t(BitReader& bs, GranuleInfo* gr, const uint8_t* hdr_bytes, const Rows& rows) {
    Hdr hdr(hdr_bytes);
    const int row = hdr.sfb_row();
    int n = hdr.mono() ? 1 : 2;
    unsigned scfsi = 0;
    int main_data_begin, part_23_sum = 0;
    if (hdr.mpeg1()) {
        n *= 2;
        main_data_begin = (int)bs.get(9);
        scfsi = bs.get(7 + n);  // private bits ride along and leak into scfsi (SURVEY 8c quirk i)
    } else {
        main_data_begin = (int)(bs.get(8 + n) >> n);
    }
    for (int k = 0; k < n; k++, gr++) {
        if (hdr.mono()) scfsi <<= 4;
        gr->part_23_length = (uint16_t)bs.get(12);
        part_23_sum += gr->part_23_length;
        gr->big_values = (uint16_t)bs.get(9);
        if (gr->big_values > 288) return -1;
        gr->global_gain = (uint8_t)bs.get(8);
        gr->scalefac_compress = (uint16_t)bs.get(hdr.mpeg1() ? 4 : 9);
        gr->sfbtab = rows(row, 0);
        gr->n_long_sfb = 22;
        gr->n_short_sfb = 0;
        unsigned tables;
        if (bs.get(1)) {
            gr->block_type = (uint8_t)bs.get(2);
            if (!gr->block_type) return -1;
            gr->mixed_block_flag = (uint8_t)bs.get(1);
            gr->region_count[0] = 7;
            gr->region_count[1] = 255;
            if (gr->block_type == 2) {
                scfsi &= 0x0F0F;
                if (!gr->mixed_block_flag) {
                    gr->region_count[0] = 8;
                    gr->sfbtab = rows(row, 1);
                    gr->n_long_sfb = 0;
                    gr->n_short_sfb = 39;
                } else {
                    gr->sfbtab = rows(row, 2);
                    gr->n_long_sfb = hdr.mpeg1() ? 8 : 6;
                    gr->n_short_sfb = 30;
                }
            }
            tables = bs.get(10) << 5;
            gr->subblock_gain[0] = (uint8_t)bs.get(3);
            gr->subblock_gain[1] = (uint8_t)bs.get(3);
            gr->subblock_gain[2] = (uint8_t)bs.get(3);
            gr->region_count[2] = 255;
        } else {
            gr->block_type = 0;
            gr->mixed_block_flag = 0;
            tables = bs.get(15);
            gr->region_count[0] = (uint8_t)bs.get(4);
            gr->region_count[1] = (uint8_t)bs.get(3);
            gr->region_count[2] = 255;
            gr->subblock_gain[0] = gr->subblock_gain[1] = gr->subblock_gain[2] = 0;
        }
        gr->table_select[0] = (uint8_t)(tables >> 10);
        gr->table_select[1] = (uint8_t)((tables >> 5) & 31);
        gr->table_select[2] = (uint8_t)(tables & 31);
        gr->preflag = hdr.mpeg1() ? (uint8_t)bs.get(1) : (uint8_t)(gr->scalefac_compress >= 500);
        gr->scalefac_scale = (uint8_t)bs.get(1);
        gr->count1_table = (uint8_t)bs.get(1);
        gr->scfsi = (uint8_t)((scfsi >> 12) & 15);
        scfsi <<= 4;
    }
    if (part_23_sum + bs.pos > bs.limit + main_data_begin * 8) return -1;
    return main_data_begin;
}
inline int parse_side_info(BitReader& bs, GranuleInfo* gr, const uint8_t* hdr_bytes) { return parse_side_info_t(bs, gr, hdr_bytes, HostSfbRows()); }

// Pack one granule-channel into the 16-byte device descriptor (layout in l3b200.h).
// The region boundaries are converted from sfb counts to coefficient indices here, so the entropy
// kernel never needs the sfb tables (region loop of minimp3.d:780-853).
L3B_HD inline l3b_grch_desc_t pack_desc(const GranuleInfo& g, uint32_t bit_start, uint8_t hdr3, bool second_granule,
                                 bool reset_before) {
    int acc = 0, i = 0;
    for (; i <= g.region_count[0] && g.sfbtab[i]; i++) acc += g.sfbtab[i];
    int r1 = acc;
    for (int j = 0; j <= g.region_count[1] && g.sfbtab[i]; j++, i++) acc += g.sfbtab[i];
    int r2 = acc;
    l3b_grch_desc_t d;
    d.bit_start = bit_start;
    d.w1 = (uint32_t)g.part_23_length | ((uint32_t)g.big_values << 12) | ((uint32_t)g.global_gain << 21) |
           ((uint32_t)g.block_type << 29) | ((uint32_t)g.mixed_block_flag << 31);
    d.w2 = (uint32_t)g.scalefac_compress | ((uint32_t)g.table_select[0] << 9) | ((uint32_t)g.table_select[1] << 14) |
           ((uint32_t)g.table_select[2] << 19) | ((uint32_t)g.preflag << 24) | ((uint32_t)g.scalefac_scale << 25) |
           ((uint32_t)g.count1_table << 26) | ((uint32_t)g.scfsi << 27) | ((uint32_t)(second_granule ? 1u : 0u) << 31);
    d.w3 = (uint32_t)(r1 / 2) | ((uint32_t)(r2 / 2) << 9) | ((uint32_t)g.subblock_gain[0] << 18) |
           ((uint32_t)g.subblock_gain[1] << 21) | ((uint32_t)g.subblock_gain[2] << 24) | ((uint32_t)(hdr3 >> 4) << 27) |
           ((uint32_t)(reset_before ? 1u : 0u) << 31);
    return d;
}

}  // namespace l3b
