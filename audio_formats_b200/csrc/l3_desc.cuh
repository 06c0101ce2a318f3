// l3_desc.cuh -- device-side view of the 16-byte granule-channel descriptor (include/l3b200.h, l3b_grch_desc_t):
// a packed L3_gr_info_t (minimp3.d:189-196) + the bit offset of the granule-channel in the stream's main data.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/l3b200.h"

namespace l3b {

struct Desc {
    uint32_t bit_start, w1, w2, w3;
    __device__ __forceinline__ int part23() const { return w1 & 0xFFF; }
    __device__ __forceinline__ int big_values() const { return (w1 >> 12) & 0x1FF; }
    __device__ __forceinline__ int global_gain() const { return (w1 >> 21) & 0xFF; }
    __device__ __forceinline__ int block_type() const { return (w1 >> 29) & 3; }
    __device__ __forceinline__ int mixed() const { return w1 >> 31; }
    __device__ __forceinline__ int scalefac_compress() const { return w2 & 0x1FF; }
    __device__ __forceinline__ int table_select(int r) const { return (w2 >> (9 + 5 * r)) & 31; }
    __device__ __forceinline__ int preflag() const { return (w2 >> 24) & 1; }
    __device__ __forceinline__ int scalefac_scale() const { return (w2 >> 25) & 1; }
    __device__ __forceinline__ int count1_table() const { return (w2 >> 26) & 1; }
    __device__ __forceinline__ int scfsi() const { return (w2 >> 27) & 15; }
    __device__ __forceinline__ int second_granule() const { return w2 >> 31; }
    __device__ __forceinline__ int region1_start() const { return (w3 & 0x1FF) * 2; }
    __device__ __forceinline__ int region2_start() const { return ((w3 >> 9) & 0x1FF) * 2; }
    __device__ __forceinline__ int subblock_gain(int i) const { return (w3 >> (18 + 3 * i)) & 7; }
    __device__ __forceinline__ int hdr_bits() const { return (w3 >> 27) & 15; }  // header byte 3 >> 4
    __device__ __forceinline__ int reset_before() const { return w3 >> 31; }
    // 0 long, 1 short, 2 mixed
    __device__ __forceinline__ int kind() const { return block_type() == 2 ? (mixed() ? 2 : 1) : 0; }
};

// What the granule kernel needs of a descriptor, packed into one word of the scalefactor record by l3_scf_kernel.
struct GranFlags {
    uint32_t v;
    static __device__ __forceinline__ uint32_t pack(const Desc& d) {
        return (uint32_t)d.block_type() | ((uint32_t)d.mixed() << 2) | ((uint32_t)d.scalefac_scale() << 3) | ((uint32_t)d.second_granule() << 4) |
               ((uint32_t)(d.scalefac_compress() & 1) << 5) | ((uint32_t)d.hdr_bits() << 6) | ((uint32_t)d.reset_before() << 10);
    }
    __device__ __forceinline__ int block_type() const { return v & 3; }
    __device__ __forceinline__ int mixed() const { return (v >> 2) & 1; }
    __device__ __forceinline__ int scalefac_scale() const { return (v >> 3) & 1; }
    __device__ __forceinline__ int second_granule() const { return (v >> 4) & 1; }
    __device__ __forceinline__ int scalefac_compress_lsb() const { return (v >> 5) & 1; }
    __device__ __forceinline__ int hdr_bits() const { return (v >> 6) & 15; }
    __device__ __forceinline__ int reset_before() const { return (v >> 10) & 1; }
    __device__ __forceinline__ int kind() const { return block_type() == 2 ? (mixed() ? 2 : 1) : 0; }
};

__device__ __forceinline__ Desc load_desc(const l3b_grch_desc_t* p) {
    uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    Desc d;
    d.bit_start = v.x; d.w1 = v.y; d.w2 = v.z; d.w3 = v.w;
    return d;
}

}  // namespace l3b
