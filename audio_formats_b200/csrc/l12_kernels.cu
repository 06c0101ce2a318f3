// l12_kernels.cu -- Layer I / II front end (integer work): bit allocation, scfsi, scalefactors and sample codes of one
// 12-slot granule -> 2 x 32 x 12 dequantised, scaled subband samples (minimp3.d:284-484).  The synthesis that follows is the
// Layer III one with 12 slots per granule (the L12 instance of l3_granule_kernel, l3_kernels.cu), like the reference, which
// calls mp3d_synth_granule for both (minimp3.d:1549, 1567).
//
// One thread per granule: a Layer I / II frame is a serial bit stream, frames are independent of each other (no reservoir),
// and a batch holds millions of them.  A Layer II frame is parsed by three threads (one per granule of the frame): each reads
// the frame's scale info (a few hundred bits) and then skips to its own third of the samples.
#include "l3_kernels.cuh"

#include <cstring>

#include "l12_tables.h"
#include "l3_tables_gen.h"
#include "l3_desc.cuh"

namespace l3b {

__constant__ uint8_t c_l12_code_tab[92];
__constant__ uint8_t c_l12_alloc[10][3];   // rows: Layer I | Layer II MPEG-2 (3) | Layer II MPEG-1 (4) | Layer II MPEG-1 low rate (2)
__constant__ float c_l12_deq[54];
__constant__ uint8_t c_l12_halfrate[2 * 3 * 15];

void upload_l12_constants() {
    cudaMemcpyToSymbol(c_l12_code_tab, L12_BITALLOC_CODE_TAB, sizeof c_l12_code_tab);
    uint8_t rows[10][3];
    memcpy(rows[0], L12_ALLOC_L1, 3);
    memcpy(rows[1], L12_ALLOC_L2M2, 9);
    memcpy(rows[4], L12_ALLOC_L2M1, 12);
    memcpy(rows[8], L12_ALLOC_L2M1_LOWRATE, 6);
    cudaMemcpyToSymbol(c_l12_alloc, rows, sizeof rows);
    cudaMemcpyToSymbol(c_l12_deq, L12_DEQ, sizeof c_l12_deq);
    cudaMemcpyToSymbol(c_l12_halfrate, L3_HALFRATE, sizeof c_l12_halfrate);
}

namespace {

// MSB-first reader over the 32-bit words of a stream's blob (n <= 16 bits per read)
struct BitRd {
    const uint32_t* words;
    uint32_t nwords, pos;
    __device__ __forceinline__ uint32_t word(uint32_t i) const { return __byte_perm(__ldg(words + min(i, nwords - 1)), 0, 0x0123); }
    __device__ __forceinline__ uint32_t get(int n) {
        const uint32_t wi = pos >> 5;
        const uint32_t v = __funnelshift_l(word(wi + 1), word(wi), pos) >> (32 - n);
        pos += (uint32_t)n;
        return n ? v : 0u;
    }
};

}  // namespace

__global__ void __launch_bounds__(128) l12_parse_kernel(BatchParams p) {
    const uint64_t gi = p.grch_lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= p.grch_hi) return;
    uint32_t si = __ldg(p.group_stream + (gi >> 7));
    while (si + 1 < p.n_streams && p.streams[si + 1].first_grch <= gi) si++;
    const l3b_stream_desc_t* S = p.streams + si;
    if (S->layer != 1 && S->layer != 2) return;              // a Layer III stream: the entropy kernels' business
    const int nch = S->nch;
    if ((gi - S->first_grch) % (uint64_t)nch) return;        // one thread per granule (it handles both channels)
    const Desc d = load_desc(p.grch + gi);
    const uint32_t h1 = d.w1 & 0xFF, h2 = (d.w1 >> 8) & 0xFF, h3 = (d.w1 >> 16) & 0xFF;
    const int part = (int)(d.w2 & 3u);
    const bool l1 = (h1 & 6) == 6, mpeg1 = (h1 & 8) != 0;
    BitRd bs;
    bs.words = reinterpret_cast<const uint32_t*>(p.blob + S->maindata_off);
    bs.nwords = (S->maindata_bytes >> 2) + 4;
    bs.pos = d.bit_start;
    if (!(h1 & 1)) bs.get(16);   // CRC word, skipped (minimp3.d:1533-1536)

    // ---- allocation table (minimp3.d:284-350) ----
    const int mode = (int)(h3 >> 6) & 3;
    int stereo_bands = mode == 3 ? 0 : (mode == 1 ? (int)(((h3 >> 4) & 3) << 2) + 4 : 32), total, row;
    if (l1) { row = 0; total = 32; }
    else if (!mpeg1) { row = 1; total = 30; }
    else {
        const int sr = (int)(h2 >> 2) & 3;
        unsigned kbps = 2u * c_l12_halfrate[(1 * 3 + (int)((h1 >> 1) & 3) - 1) * 15 + (int)(h2 >> 4)];
        kbps >>= (mode != 3 ? 1 : 0);
        if (!kbps) kbps = 192;
        row = 4; total = 27;
        if (kbps < 56) { row = 8; total = sr == 2 ? 12 : 8; }
        else if (kbps >= 96 && sr != 1) total = 30;
    }
    stereo_bands = min(stereo_bands, total);

    // ---- scale info (minimp3.d:387-435) ----
    uint8_t bitalloc[64];
    float scf[64];   // the scalefactor of THIS granule for every band-channel entry
    {
        int k = 0, ba_bits = 0, tab = 0;
        for (int i = 0; i < total; i++) {
            if (i == k) { k += c_l12_alloc[row][2]; ba_bits = c_l12_alloc[row][1]; tab = c_l12_alloc[row][0]; row++; }
            uint8_t ba = c_l12_code_tab[tab + bs.get(ba_bits)];
            bitalloc[2 * i] = ba;
            if (i < stereo_bands) ba = c_l12_code_tab[tab + bs.get(ba_bits)];
            bitalloc[2 * i + 1] = stereo_bands ? ba : 0;
        }
    }
    uint32_t cod_lo = 0, cod_hi = 0, cod_x = 0, cod_y = 0;   // scfcod: 3 bits per entry, 64 entries
    for (int i = 0; i < 2 * total; i++) {
        // the reference evaluates get_bits(2) for EVERY entry of a Layer II frame, allocated or not (minimp3.d:417-421)
        const uint32_t temp = l1 ? 2u : bs.get(2);
        const uint32_t c = bitalloc[i] ? temp : 6u;
        uint32_t& w = i < 16 ? cod_lo : (i < 32 ? cod_hi : (i < 48 ? cod_x : cod_y));
        w |= (c & 3u) << (2 * (i & 15));     // 6 is stored as its two low bits: only read where bitalloc != 0, where c < 4
    }
    for (int i = 0; i < 2 * total; i++) {
        const int ba = bitalloc[i];
        const uint32_t w = i < 16 ? cod_lo : (i < 32 ? cod_hi : (i < 48 ? cod_x : cod_y));
        const int mask = ba ? 4 + ((19 >> ((w >> (2 * (i & 15))) & 3u)) & 3) : 0;
        float s = 0.0f, keep = 0.0f;
        int idx = 0;
        for (int m = 4; m; m >>= 1, idx++) {
            if (mask & m) {
                const int b = (int)bs.get(6);
                s = __fmul_rn(c_l12_deq[ba * 3 - 6 + b % 3], (float)((1 << 21) >> (b / 3)));
            }
            if (idx == part) keep = s;   // Layer I: one granule per frame, part 0, and all three values are equal
        }
        scf[i] = keep;
    }
    for (int i = stereo_bands; i < total; i++) bitalloc[2 * i + 1] = 0;

    // ---- samples of this granule (minimp3.d:437-470), scaled (minimp3.d:472-484) ----
    const int group = l1 ? 1 : 3;
    if (part) {   // Layer II: skip the granules before this one
        uint32_t per_call = 0;
        for (int i = 0; i < 2 * total; i++) {
            const int ba = bitalloc[i];
            if (!ba) continue;
            const int mod = (2 << (ba - 17)) + 1;
            per_call += 4u * (uint32_t)(ba < 17 ? group * ba : mod + 2 - (mod >> 3));
        }
        bs.pos += per_call * (uint32_t)part;
    }
    float* const out0 = p.l12_x + gi * 384;
    float* const out1 = out0 + 384;           // channel 1 (stereo streams only)
    const int n_groups = l1 ? 12 : 4;
    for (int j = 0; j < n_groups; j++) {
        for (int i = 0; i < 2 * total; i++) {
            const int ba = bitalloc[i];
            if (!ba) continue;
            const int band = i >> 1, c = i & 1;
            float* dst = (c ? out1 : out0) + band * 12 + group * j;
            const bool shared = !c && band >= stereo_bands && nch == 2;   // joint stereo: channel 1 repeats channel 0's codes
            if (ba < 17) {
                const int half = (1 << (ba - 1)) - 1;
                for (int k = 0; k < group; k++) {
                    const float v = (float)((int)bs.get(ba) - half);
                    dst[k] = __fmul_rn(v, scf[i]);
                    if (shared) out1[band * 12 + group * j + k] = __fmul_rn(v, scf[i + 1]);
                }
            } else {
                const unsigned mod = (2u << (ba - 17)) + 1u;
                unsigned code = bs.get((int)(mod + 2 - (mod >> 3)));
                for (int k = 0; k < group; k++, code /= mod) {
                    const float v = (float)((int)(code % mod) - (int)(mod / 2));
                    dst[k] = __fmul_rn(v, scf[i]);
                    if (shared) out1[band * 12 + group * j + k] = __fmul_rn(v, scf[i + 1]);
                }
            }
        }
    }
}

int launch_l12_parse(const BatchParams& p, cudaStream_t s) {
    if (p.grch_hi <= p.grch_lo || !p.l12_x) return 0;
    const uint64_t n = p.grch_hi - p.grch_lo;
    l12_parse_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(p);
    return 1;
}

}  // namespace l3b
