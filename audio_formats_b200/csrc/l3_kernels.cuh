// l3_kernels.cuh -- device code of the Layer III granule decode path (sm_100a).
//
// Kernels:
//   l3_scf_kernel, l3_huff_big_kernel, l3_huff_c1_kernel   (l3_entropy.cu) scalefactor decode, band gains and Huffman
//                       decode (minimp3.d:613-644, 659-719, 748-883) -> signed 16-bit quantised spectra + 256-byte
//                       scalefactor/gain records.  Integer work (plus the gain products); the spectra are the
//                       "bit-exact spectral intermediates".
//   l3_granule_kernel   (l3_kernels.cu) one warp per run of consecutive granules of one stream:
//                       requantisation (minimp3.d:727-746, 813-816, 846), MS/intensity stereo (:885-982),
//                       short-block reorder (:984-1000), alias reduction (:1002-1020), IMDCT-36/12 with
//                       overlap-add and frequency inversion (:1022-1168), DCT-32 matrixing (:1232-1298)
//                       and the 512-tap window (:1305-1406), fused through shared memory and registers.
//
// Float arithmetic follows the reference's operation order exactly; this translation unit MUST be
// compiled with -fmad=false (no FMA contraction) and default (IEEE) division/denormal settings so
// that PCM is bit-identical to the un-fused scalar reference on x86-64.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/l3b200.h"

namespace l3b {

constexpr int kSfRecBytes = 96;    // per granule-channel: iscf[40] ist_pos[40] nz_chunks(u16) pad[2] gain(f32) flags(u32) pad[4]
constexpr int kSfGainOff = 84;     // byte offset of the granule-channel gain 2^(gain_exp/4) (minimp3.d:714-716) inside the record;
                                   // the 40 band gains are one table multiplication each in the granule kernel
constexpr int kSfFlagsOff = 88;    // byte offset of the descriptor bits the granule kernel needs (GranFlags, l3_desc.cuh): with them in
                                   // the record the kernel stages one thing less per granule
constexpr int kIsChunks = 72;      // 576 int16 = 72 x 16 bytes
constexpr int kTileGranulesMax = 128;   // longest warp tile of the granule kernel, in granules (l3b_batch_create picks 16 .. 128 per batch)
constexpr int kXrStride = 608;     // spectrum buffer elements: 576 in natural layout / 32x19 in the padded layout

struct Tile {
    uint32_t stream, g0, ng;
};

struct DeviceTables {
    const uint16_t* huff;       // HuffLut::entries
    uint32_t huff_entries;
    uint16_t huff_base[18];     // books 0..14, 15 = all-zero book, 16/17 = count1 A/B
    uint8_t huff_root[18];
    const uint8_t* count1;      // [2][64]
    const uint32_t* huff32;     // HuffLut32::entries
    uint32_t huff32_entries;
    uint32_t huff32_base[16];   // books 0..14, 15 = all-zero book
    uint8_t huff32_root[16];
    const uint16_t* c1code;     // [2][64]
    const uint2* c1val;         // [256]
    const uint8_t* sfb_of_pair; // [8][3][288]
    const uint8_t* sfb_width;   // [8][3][40]
    const uint16_t* sfb_start;  // [8][3][40]
    const uint16_t* perm;       // [8][2][576]
    const float* pow43;         // [129]
    const float* win;           // [240]  L3_WIN layout
};

// One Huffman job per granule-channel (48 bytes = three 16-byte loads, none depending on another): written by the
// scalefactor kernel for the big_values kernel, which in turn leaves the count1 kernel its bit position, its bit
// window and (in place of `par`/`misc`) the partial 16-byte chunk of output that the quads continue in.
struct __align__(16) HuffJob {
    uint32_t word_base;   // first 32-bit word of the stream's main data in the batch blob
    uint32_t nwords;      // words readable from there (the pad words after the stream included)
    uint32_t pos;         // bit position (relative to the stream's main data) where decoding continues
    uint32_t limit;       // bit_start + part2_3_length (minimp3.d:1201)
    uint32_t par[3];      // per region: LUT base | (32 - root_bits) << 16 | linbits << 24
    uint32_t misc;        // big_values (pairs) | region1 start (pairs) << 10 | region2 start (pairs) << 20 | count1_table << 30
    uint32_t w[4];        // the four words from the one holding bit `pos`: two big-endian, two in memory byte order
};

struct BatchParams {
    const uint8_t* blob;
    const l3b_grch_desc_t* grch;
    uint64_t n_grch;
    const l3b_stream_desc_t* streams;
    uint32_t n_streams;
    uint4* is;       // [n_grch][72] packed int16x8
    uint8_t* sf;     // [n_grch][kSfRecBytes]
    uint8_t* nzc;    // [n_grch] nz_chunks again, packed: read a granule ahead by the granule kernel to size its TMA copies
    float* pcm;      // float delivery (NULL when pcm16 is set)
    int16_t* pcm16;  // 16-bit delivery: q = clamp(lrintf(x * 32768)) of the float sample, same element offsets
    float* l12_x;    // [n_grch][384] Layer I / II: dequantised, scaled subband samples [band 32][slot 12] per granule-channel
                     //   (NULL when the batch holds no Layer I / II stream); zeroed before every run
    float *tap_xr, *tap_st, *tap_im, *tap_dct;   // [n_grch][576] float stage snapshots (test taps; TAPS kernels only)
    uint64_t grch_lo, grch_hi;  // granule-channel range the entropy kernels cover in this launch
    int zero_fill;   // the count1 kernel zero-fills the chunks it does not reach (tap mode)
    HuffJob* jobs;        // [n_grch]
    const uint32_t* group_stream;   // [n_grch / 128 + 1] stream index of granule-channel 128 k (search hint)
    uint32_t* counters;   // work-distribution counters of the Huffman kernels: 2 per sub-batch, zeroed before every run
    DeviceTables t;
};

constexpr int kGranuleWarpsStereo = 4;   // warps (= tiles) per CTA of the granule kernel; 16 warps resident per SM.
constexpr int kGranuleWarpsMono = 4;     // 4-warp CTAs measured best (16: 33.7 ms, 8: 28.6 ms, 4: 27.2 ms on config 2)

template <int NCH, int WARPS, bool FUSED, bool TAPS, bool S16, bool L12>
__global__ void l3_granule_kernel(BatchParams p, const Tile* tiles, uint32_t n_tiles);

// lane-decoupled entropy path: scalefactor kernel, big_values kernel, count1 kernel (l3_entropy.cu); returns launches
int launch_entropy_v4(const BatchParams& p, int sub, int sms, cudaStream_t s);
// fused: tolerance-mode arithmetic (hand-contracted FFMA2) instead of the bit-exact one; taps: float stage snapshots
cudaError_t launch_granule(const BatchParams& p, const Tile* tiles_stereo, uint32_t n_stereo, const Tile* tiles_mono,
                           uint32_t n_mono, cudaStream_t s, bool fused, bool taps);
// Layer I / II (l12_kernels.cu): one thread per 12-slot granule parses its share of the frame into p.l12_x; returns launches
int launch_l12_parse(const BatchParams& p, cudaStream_t s);
// the L12 instances of the granule kernel: synthesis only, 12 slots per granule
cudaError_t launch_granule_l12(const BatchParams& p, const Tile* tiles_stereo, uint32_t n_stereo, const Tile* tiles_mono,
                               uint32_t n_mono, cudaStream_t s, bool fused);
void upload_constants();
void upload_entropy_constants();
void upload_l12_constants();

}  // namespace l3b
