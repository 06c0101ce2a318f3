// l3_entropy.cu -- lane-decoupled entropy decode (integer only): scalefactors, big_values, count1.
//
// The Huffman data of a granule-channel is a serial bit stream, so the unit of parallelism is the granule-channel and
// its cost (pairs + quads actually coded) varies by 10x from one to the next.  A warp that owns 32 fixed
// granule-channels runs as long as its slowest lane: on the 128 kbps bench streams that wastes half of all lane
// slots.  Here the lanes of a warp are decoupled instead:
//
//   l3_scf_kernel       one thread per granule-channel: scalefactors (minimp3.d:613-644, 659-712) and the granule-channel
//                       gain 2^(gain_exp/4) (minimp3.d:714-716) -> 96-byte record; and a 48-byte HuffJob (bit position
//                       and window, region books, limits) for the two kernels below
//   l3_huff_big_kernel  persistent warps; every LANE pulls granule-channels from a global counter and decodes their
//                       big_values pairs (minimp3.d:789-853), four pairs (one 16-byte chunk) per trip; a lane that
//                       runs out of pairs parks until enough lanes are free, then they write their hand-over to
//                       count1 and refill together; bit-stream words arrive through a cp.async ring
//   l3_huff_c1_kernel   same scheme for the count1 quads (minimp3.d:855-882), continuing at the bit position the
//                       big_values kernel left in the job
//
// Keeping the two Huffman phases in separate kernels keeps every warp on ONE straight-line code path: a warp whose
// lanes sit in different phases would have to run both.  Output format is unchanged: int16 spectra as 72 16-byte
// chunks per granule-channel, `nz_chunks` in the scalefactor record.
#include "l3_kernels.cuh"

#include <algorithm>
#include <cstdlib>

#include "l3_desc.cuh"
#include "l3_tables_gen.h"

namespace l3b {

__constant__ uint8_t e_partitions[84];
__constant__ uint8_t e_scfc_decode[16];
__constant__ uint8_t e_lsf_mod[24];
__constant__ uint8_t e_preamp[10];
__constant__ uint8_t e_linbits[32];
__constant__ int8_t e_sel2book[32];
__constant__ float e_expfrac[4];

void upload_entropy_constants() {
    cudaMemcpyToSymbol(e_partitions, L3_SCF_PARTITIONS, sizeof e_partitions);
    cudaMemcpyToSymbol(e_scfc_decode, L3_SCFC_DECODE, sizeof e_scfc_decode);
    cudaMemcpyToSymbol(e_lsf_mod, L3_LSF_MOD, sizeof e_lsf_mod);
    cudaMemcpyToSymbol(e_preamp, L3_PREAMP, sizeof e_preamp);
    cudaMemcpyToSymbol(e_linbits, L3_LINBITS, sizeof e_linbits);
    cudaMemcpyToSymbol(e_sel2book, L3_SEL2BOOK, sizeof e_sel2book);
    cudaMemcpyToSymbol(e_expfrac, L3_EXPFRAC, sizeof e_expfrac);
}

namespace {

// MSB-first bit reader over the 32-bit words of a stream's main data (scalefactor fields only).  Four words are
// loaded up front (independent loads, one wait) and shifted through registers; further words are fetched two
// crossings before they are needed.
struct ScfReader {
    const uint32_t* words;
    uint32_t nwords, pos, w0, w1, w2, w3, wn;
    // nwords counts the 16 zero pad bytes that follow every stream, so clamping the index makes reads past the end
    // return zero bits without a branch
    __device__ __forceinline__ uint32_t ldw(uint32_t i) const { return __byte_perm(__ldg(words + min(i, nwords - 1)), 0, 0x0123); }
    __device__ __forceinline__ void open(uint32_t bitpos) {
        pos = bitpos;
        const uint32_t wi = bitpos >> 5;
        w0 = ldw(wi); w1 = ldw(wi + 1); w2 = ldw(wi + 2); w3 = ldw(wi + 3);
        wn = wi + 4;
    }
    __device__ __forceinline__ uint32_t peek_at(uint32_t at, int n) const {   // random access, 1 <= n <= 16
        const uint32_t wi = at >> 5;
        return __funnelshift_l(ldw(wi + 1), ldw(wi), at) >> (32 - n);
    }
    __device__ __forceinline__ uint32_t get(int n) {   // 1 <= n <= 16
        const uint32_t v = __funnelshift_l(w1, w0, pos) >> (32 - n);
        const uint32_t np = pos + (uint32_t)n;
        if ((np ^ pos) & ~31u) { w0 = w1; w1 = w2; w2 = w3; w3 = ldw(wn); wn++; }
        pos = np;
        return v;
    }
};

// L3_ldexp_q2 (minimp3.d:646-657): y * 2^(-exp_q2/4) by repeated multiplication, same rounding steps.
// (float)((1 << 30) >> k) is the power of two 2^(30-k), 0 <= k <= 30: built from its exponent bits instead of an
// integer shift + conversion; g_expfrac[e & 3] is picked from registers (a lane-varying constant-bank index would
// be serialised).
__device__ __forceinline__ float ldexp_q2(float y, int exp_q2) {
    const float f0 = e_expfrac[0], f1 = e_expfrac[1], f2 = e_expfrac[2], f3 = e_expfrac[3];
    int e;
    do {
        e = exp_q2 < 120 ? exp_q2 : 120;
        const float frac = (e & 2) ? ((e & 1) ? f3 : f2) : ((e & 1) ? f1 : f0);
        const float pw = __int_as_float((127 + 30 - (e >> 2)) << 23);
        y = __fmul_rn(y, __fmul_rn(frac, pw));
    } while ((exp_q2 -= e) > 0);
    return y;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// Scalefactors + band gains + job setup: one thread per granule-channel
static_assert(kSfRecBytes == 96 && kSfGainOff == 84 && sizeof(HuffJob) == 48, "the staging rows of l3_scf_kernel assume these sizes");
constexpr int kRecRow = kSfRecBytes / 16 + 1;   // staging row of a record in 16-byte units (one spare: conflict-free rows)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) l3_scf_kernel(BatchParams p) {
    // per warp: 32 records of kRecRow x 16 bytes, written out as one contiguous run
    extern __shared__ __align__(16) uint4 s_stage[];
    const uint32_t lane = threadIdx.x & 31;
    uint4* const wstage = s_stage + (threadIdx.x >> 5) * (32 * kRecRow);
    const uint64_t g_first = p.grch_lo + (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u);   // first of this warp
    if (g_first >= p.grch_hi) return;
    const uint32_t n_live = (uint32_t)min((uint64_t)32, p.grch_hi - g_first);
    const uint64_t gi = min(g_first + lane, p.grch_hi - 1);   // tail lanes shadow the last granule-channel
    uint32_t si = __ldg(p.group_stream + (g_first >> 7));   // a stream at or before ours: walk forward (usually 0 steps)
    while (si + 1 < p.n_streams && p.streams[si + 1].first_grch <= gi) si++;
    const l3b_stream_desc_t* S = p.streams + si;
    const int nch = S->nch;
    const int ch = (int)((gi - S->first_grch) % (uint64_t)nch);
    const bool mpeg1 = S->mpeg1 != 0;
    ScfReader br;
    br.words = reinterpret_cast<const uint32_t*>(p.blob + S->maindata_off);
    br.nwords = (S->maindata_bytes >> 2) + 4;  // the batch blob keeps >= 16 zero bytes after each stream
    Desc d = load_desc(p.grch + gi);
    if (S->layer == 1 || S->layer == 2) d.w1 = d.w2 = d.w3 = 0;   // a Layer I / II granule (l12_parse_kernel's): an empty Huffman job
    br.open(d.bit_start);

    // ---------------- scalefactors (minimp3.d:613-644, 659-712) ----------------
    uint32_t recw[kSfRecBytes / 4];
#pragma unroll
    for (int i = 0; i < kSfRecBytes / 4; i++) recw[i] = 0;
    uint8_t* const rec = reinterpret_cast<uint8_t*>(recw);   // thread-local staging of the record (local memory, 96 B)
    const int kind = d.kind();
    const int n_long = kind == 0 ? 22 : (kind == 1 ? 0 : (mpeg1 ? 8 : 6));
    const int n_short = kind == 0 ? 0 : (kind == 1 ? 39 : 30);
    const uint8_t* part = e_partitions + 28 * (kind == 0 ? 0 : (kind == 2 ? 1 : 2));
    const int scf_shift = d.scalefac_scale() + 1;
    uint32_t slen = 0;  // four byte-sized lengths
    int scfsi = d.scfsi();
    const int istereo = d.hdr_bits() & 1;
    if (mpeg1) {
        const int pp = e_scfc_decode[d.scalefac_compress() & 15];
        const uint32_t a = (uint32_t)(pp >> 2), b = (uint32_t)(pp & 3);
        slen = a | (a << 8) | (b << 16) | (b << 24);
    } else {
        const int ist = (istereo && ch) ? 1 : 0;
        int sfc = d.scalefac_compress() >> ist;
        int k = ist * 12;
        for (;; k += 4) {
            int modprod = 1;
            slen = 0;
#pragma unroll
            for (int i = 3; i >= 0; i--) {
                const int m = e_lsf_mod[k + i];
                slen |= (uint32_t)(sfc / modprod % m) << (8 * i);
                modprod *= m;
            }
            sfc -= modprod;
            if (sfc < 0) break;
        }
        part += k + 4;  // the reference's for-loop increments k once more before its exit test (minimp3.d:683-691)
        scfsi = -16;
    }
    // scfsi copies (MPEG-1 only).  The reference copies from `ist_pos`, per-frame scratch that starts zeroed (D default
    // initialisation, minimp3.d:1497) and holds granule 0's values when granule 1 is read (minimp3.d:619-622):
    //  * granule 1 copies what granule 0 of the same channel READ (both are long-type blocks then: a short block
    //    clears the bits, minimp3.d:568); the values are fetched again from granule 0's bit field;
    //  * granule 0 can have scfsi bits too -- the private bits leak into them (minimp3.d:530-540, 600-601) -- and
    //    then "copies" zeros and does not read that partition; granule 1 in turn copies those zeros.
    uint32_t g0_slen = 0, g0_bits = 0;
    int g0_nib = 0;              // granule 0's own scfsi nibble (partitions it did not read)
    bool from_zero = false;      // this IS granule 0: the copy source is the zeroed scratch
    if (scfsi > 0 && d.second_granule() && gi >= S->first_grch + (uint64_t)nch) {
        const Desc d0 = load_desc(p.grch + gi - nch);
        const int pp = e_scfc_decode[d0.scalefac_compress() & 15];
        const uint32_t a = (uint32_t)(pp >> 2), b = (uint32_t)(pp & 3);
        g0_slen = a | (a << 8) | (b << 16) | (b << 24);
        g0_bits = d0.bit_start;
        g0_nib = d0.scfsi();
    } else if (scfsi > 0) {
        from_zero = true;
    }
    {
        const int sbg_sh = 3 - scf_shift;
        int n = 0;
        uint32_t g0_off = g0_bits;
        for (int i = 0; i < 4; i++) {
            const int cnt = part[i];
            if (!cnt) break;
            const int bits = (slen >> (8 * i)) & 0xFF;
            const int bits0 = (g0_slen >> (8 * i)) & 0xFF;
            const bool copy = (scfsi & 8) != 0;
            const bool src_zero = from_zero || (g0_nib & 8) || !bits0;   // the source partition was never read: zeros
            for (int k = 0; k < cnt; k++, n++) {
                int s, ip;
                if (copy) {
                    s = src_zero ? 0 : (int)br.peek_at(g0_off + (uint32_t)(k * bits0), bits0);
                    ip = s;
                } else if (!bits) {
                    s = 0; ip = 0;
                } else {
                    s = (int)br.get(bits);
                    ip = (scfsi < 0 && s == (1 << bits) - 1) ? 255 : s;
                }
                int adj = 0;
                if (n_short) { if (n >= n_long) adj = d.subblock_gain((n - n_long) % 3) << sbg_sh; }
                else if (d.preflag() && n >= 11 && n < 21) adj = e_preamp[n - 11];
                rec[n] = (uint8_t)(s + adj);
                rec[40 + n] = (uint8_t)ip;
            }
            if (!(g0_nib & 8)) g0_off += (uint32_t)(cnt * bits0);   // granule 0 skipped the partitions it "copied"
            scfsi *= 2;
            g0_nib *= 2;
        }
        for (int j = 0; j < 3 && n < 40; j++, n++) {  // scf[0] = scf[1] = scf[2] = 0 after the last partition
            int adj = 0;
            if (n_short) { if (n >= n_long && n < n_long + n_short) adj = d.subblock_gain((n - n_long) % 3) << sbg_sh; }
            rec[n] = (uint8_t)adj;
        }
    }
    {
        // ---------------- gain of the granule-channel (minimp3.d:714-716): 2^(gain_exp/4) as L3_ldexp_q2 computes it.  The 40 band
        // gains scf[i] = gain * 2^(-(iscf[i] << shift)/4) are one more table multiplication each, done by the granule kernel.
        const bool ms_frame = (d.hdr_bits() & 0xE) == 0x6;   // HDR_IS_MS_STEREO: the 1/sqrt(2) of MS stereo is folded in here
        const int gain_exp = d.global_gain() - 4 - 210 - (ms_frame ? 2 : 0);
        recw[kSfGainOff / 4] = __float_as_uint(ldexp_q2(2048.0f, 44 - gain_exp));
        recw[kSfFlagsOff / 4] = GranFlags::pack(d);
        uint4* dst = wstage + lane * kRecRow;
#pragma unroll
        for (int i = 0; i < kSfRecBytes / 16; i++) dst[i] = make_uint4(recw[4 * i], recw[4 * i + 1], recw[4 * i + 2], recw[4 * i + 3]);
        __syncwarp();
        uint4* out = reinterpret_cast<uint4*>(p.sf + g_first * kSfRecBytes);
        for (uint32_t idx = lane; idx < n_live * (kSfRecBytes / 16); idx += 32) out[idx] = wstage[(idx / 6u) * kRecRow + (idx % 6u)];
        __syncwarp();
    }

    // ---------------- Huffman job ----------------
    HuffJob j;
    j.word_base = (uint32_t)(S->maindata_off >> 2);
    j.nwords = br.nwords;
    j.pos = br.pos;
    j.limit = d.bit_start + (uint32_t)d.part23();
#pragma unroll
    for (int r = 0; r < 3; r++) {
        const int sel = d.table_select(r);
        const int book = e_sel2book[sel] < 0 ? L3_NBOOKS : e_sel2book[sel];
        j.par[r] = p.t.huff32_base[book] | ((32u - p.t.huff32_root[book]) << 16) | ((uint32_t)e_linbits[sel] << 24);
    }
    j.misc = (uint32_t)d.big_values() | ((uint32_t)(d.region1_start() >> 1) << 10) | ((uint32_t)(d.region2_start() >> 1) << 20) |
             ((uint32_t)d.count1_table() << 30);
    uint4* jd = wstage + lane * 3;
    jd[0] = make_uint4(j.word_base, j.nwords, j.pos, j.limit);
    jd[1] = make_uint4(j.par[0], j.par[1], j.par[2], j.misc);
    // the first four words of the bit window (the last two in memory byte order): no dependent load at refill
    jd[2] = make_uint4(br.w0, br.w1, __byte_perm(br.w2, 0, 0x0123), __byte_perm(br.w3, 0, 0x0123));
    __syncwarp();
    uint4* jout = reinterpret_cast<uint4*>(p.jobs + g_first);
    for (uint32_t idx = lane; idx < n_live * 3; idx += 32) jout[idx] = wstage[idx];
}

// ---------------------------------------------------------------------------------------------------------------
// Shared pieces of the two Huffman kernels
// ---------------------------------------------------------------------------------------------------------------
namespace {

constexpr uint32_t kRingWords = 8;    // per-lane ring of prefetched bit-stream words (shared memory)
constexpr uint32_t kAhead = 6;        // words requested beyond the one that enters the window next

// shared-window address of p, made opaque so that the compiler keeps it in a register instead of re-deriving it
// (S2R SR_CgaCtaId + LEA) at every use
__device__ __forceinline__ uint32_t smem_addr(const void* p) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p), r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(a));
    return r;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Window over the bit stream.  w0:w1 are the big-endian words holding bit `pos` and the 32 bits after it.  The words
// behind them come from a per-lane ring in shared memory that is topped up with cp.async ONLY at trip boundaries:
// a load issued inside a decode step would be waited for by the next step of every lane of the warp (scoreboards
// are per warp, not per lane), which exposes the full memory latency once per step.  Requests made at one boundary
// are waited for at the next one, a whole trip later.
struct BitWindow {
    const uint32_t* words;
    uint32_t nwords, pos, w0, w1;
    uint32_t wn;       // index of the word that enters w1 at the next crossing
    uint32_t ready;    // words below this index have landed in the ring
    uint32_t filled;   // words below this index have been requested
    uint32_t ring;     // shared address of this lane's slot 0; word i lives at ring + (i % kRingWords) * 128

    __device__ __forceinline__ uint32_t slot(uint32_t i) const { return ring + ((i & (kRingWords - 1)) << 7); }
    __device__ __forceinline__ uint32_t peek32() const { return __funnelshift_l(w1, w0, pos); }
    __device__ __forceinline__ uint32_t raw_word(uint32_t i) const {   // memory byte order
        if (i < ready) return lds32(slot(i));
        return __ldg(words + min(i, nwords - 1));   // starved (escape-heavy data): direct load; nwords counts the zero pad
    }
    __device__ __forceinline__ void advance(uint32_t n) {   // n < 32
        const uint32_t np = pos + n;
        if ((np ^ pos) & ~31u) {
            w0 = w1;
            w1 = __byte_perm(raw_word(wn), 0, 0x0123);
            wn++;
        }
        pos = np;
    }
    // job hand-over: bit position + four words starting at the one holding `pos`
    __device__ __forceinline__ void open(const uint32_t* w, uint32_t nw, uint32_t bitpos, uint4 first) {
        words = w; nwords = nw; pos = bitpos;
        w0 = first.x; w1 = first.y;
        wn = (bitpos >> 5) + 2;
        sts32(slot(wn), first.z);
        sts32(slot(wn + 1), first.w);
        ready = filled = wn + 2;
    }
    __device__ __forceinline__ uint4 close() const { return make_uint4(w0, w1, raw_word(wn), raw_word(wn + 1)); }
    // trip boundary: everything requested a trip ago has landed; request what the coming trips may need
    __device__ __forceinline__ void top_up(bool live) {
        cp_async_wait_all();
        ready = filled;
        if (live) {
            const uint32_t want = wn + kAhead;
            while (filled < want) {
                cp_async4(slot(filled), words + min(filled, nwords - 1));
                filled++;
            }
        }
    }
};

// Work distribution.  A warp takes blocks of 32 consecutive jobs from a global counter (one atomic per block) and
// keeps them in shared memory; its lanes draw from that pool one by one as they run out of work.  The following
// block is already on its way into registers (three coalesced 16-byte loads per lane) while the current one is
// being handed out, so a lane that refills reads its job at shared-memory latency.
template <bool WITH_DESC>
struct JobPool {
    uint4* pool;                 // [96] per warp: job j = pool[3j .. 3j+2]; WITH_DESC: [96..127] = descriptors
    uint4 nx0, nx1, nx2, nxd;    // this lane's pieces (lane, lane+32, lane+64) of the next block (+ its descriptor)
    uint32_t base, used, count;  // current block: first item, items handed out, items it holds
    uint32_t next_base, next_count;

    __device__ __forceinline__ void fetch_next(uint32_t* counter, const uint4* jobs_u4, const uint4* descs, uint32_t n_items, uint32_t lane) {
        uint32_t nb = 0;
        if (lane == 0) nb = atomicAdd(counter, 32u);
        nb = __shfl_sync(0xffffffffu, nb, 0);
        next_base = nb;
        next_count = nb < n_items ? min(32u, n_items - nb) : 0u;
        const uint4* src = jobs_u4 + (size_t)nb * 3u;
        const uint32_t pieces = 3u * next_count;
        if (lane < pieces) nx0 = __ldg(src + lane);
        if (lane + 32u < pieces) nx1 = __ldg(src + lane + 32u);
        if (lane + 64u < pieces) nx2 = __ldg(src + lane + 64u);
        if (WITH_DESC && lane < next_count) nxd = __ldg(descs + nb + lane);
    }
    __device__ __forceinline__ void swap_in(uint32_t* counter, const uint4* jobs_u4, const uint4* descs, uint32_t n_items, uint32_t lane) {
        __syncwarp();
        pool[lane] = nx0; pool[lane + 32] = nx1; pool[lane + 64] = nx2;
        if (WITH_DESC) pool[lane + 96] = nxd;
        __syncwarp();
        base = next_base; count = next_count; used = 0;
        if (count) fetch_next(counter, jobs_u4, descs, n_items, lane);
    }
    __device__ __forceinline__ void init(uint4* smem, uint32_t* counter, const uint4* jobs_u4, const uint4* descs, uint32_t n_items, uint32_t lane) {
        pool = smem;
        nx0 = nx1 = nx2 = nxd = make_uint4(0, 0, 0, 0);
        base = used = count = 0;
        fetch_next(counter, jobs_u4, descs, n_items, lane);
        swap_in(counter, jobs_u4, descs, n_items, lane);
    }
    // Lanes in `need` draw jobs in lane order; returns the pool slot of this lane or -1.  Warp-uniform bookkeeping.
    __device__ __forceinline__ int draw(unsigned need, bool wants, uint32_t lane) {
        const uint32_t idx = used + (uint32_t)__popc(need & ((1u << lane) - 1u));
        const int slot = (wants && idx < count) ? (int)idx : -1;
        used = min(count, used + (uint32_t)__popc(need));
        return slot;
    }
    __device__ __forceinline__ bool empty() const { return used >= count; }
};

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// big_values: one pair per step, four steps (one 16-byte chunk) per trip, lanes refill at trip boundaries.
// When a lane is through with its pairs it leaves the count1 kernel everything that one needs in the job itself:
// the bit position, the bit window and the partial chunk it has to continue in.
// Dynamic shared memory: LUT | WARPS job pools (96 x 16 B) | WARPS word rings (kRingWords x 32 x 4 B)
// ---------------------------------------------------------------------------------------------------------------
template <int K, int WARPS>
__global__ void __launch_bounds__(32 * WARPS) l3_huff_big_kernel(BatchParams p, uint32_t* counter) {
    extern __shared__ __align__(16) uint32_t s_lut[];
    for (uint32_t i = threadIdx.x; i < p.t.huff32_entries; i += blockDim.x) s_lut[i] = p.t.huff32[i];
    __syncthreads();
    const uint32_t lut = smem_addr(s_lut);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n_items = (uint32_t)(p.grch_hi - p.grch_lo);
    const uint32_t* const blob32 = reinterpret_cast<const uint32_t*>(p.blob);
    HuffJob* const jobs = p.jobs + p.grch_lo;
    uint4* const is_base = p.is + p.grch_lo * kIsChunks;
    uint32_t* const s_after_lut = s_lut + ((p.t.huff32_entries + 3u) & ~3u);

    BitWindow bw;
    bw.words = blob32; bw.nwords = 1; bw.pos = 0; bw.w0 = bw.w1 = 0; bw.wn = bw.ready = bw.filled = 0;
    bw.ring = smem_addr(s_after_lut + WARPS * 96 * 4 + warp * (kRingWords * 32) + lane);
    uint32_t item = 0, widx = 0, bvw = 0, nbw = 0, r2w = 0, par1 = 0, par2 = 0;
    uint32_t cur_base = lut, cur_sh = 31, cur_lin = 0;   // current region's book: LUT address, 32 - root bits, linbits
    bool have = false, parked = false;   // parked: finished its pairs, hand-over to count1 not written yet
    uint32_t last[4] = {0, 0, 0, 0};      // the last output chunk of a parked lane
    const uint4* const jobs_u4 = reinterpret_cast<const uint4*>(jobs);
    // What count1 needs, written into the job in place of what big_values no longer needs.  Done when the lane
    // refills (several lanes at once) instead of in the trip the lane finished in (one lane at a time).
    auto hand_over = [&]() {
        uint4* jw = reinterpret_cast<uint4*>(jobs + item);
        reinterpret_cast<uint32_t*>(jw)[2] = bw.pos;
        jw[1] = make_uint4(last[0], last[1], last[2], last[3]);   // the chunk count1 continues in (only used when bvw & 3)
        jw[2] = bw.close();
        parked = false;
    };
    JobPool<false> jp;
    jp.init(reinterpret_cast<uint4*>(s_after_lut) + warp * 96, counter, jobs_u4, nullptr, n_items, lane);

    for (;;) {
        const unsigned need = __ballot_sync(0xffffffffu, !have);
        if (__popc(need) >= K && !jp.empty()) {
            cp_async_wait_all();   // nothing of a finished job may still be landing in a ring that is about to be re-primed
            if (parked) hand_over();
            const int slot = jp.draw(need, !have, lane);
            if (slot >= 0) {
                const uint4 a = jp.pool[3 * slot], b = jp.pool[3 * slot + 1], w = jp.pool[3 * slot + 2];
                item = jp.base + (uint32_t)slot;
                have = true;
                bvw = b.w & 0x3FFu;
                widx = 0;
                cur_base = lut + ((b.x & 0xFFFFu) << 2); cur_sh = (b.x >> 16) & 31u; cur_lin = b.x >> 24;
                par1 = b.y; par2 = b.z;
                nbw = (b.w >> 10) & 0x3FFu;
                r2w = (b.w >> 20) & 0x3FFu;
                bw.open(blob32 + a.x, a.y, a.z, w);
            }
            if (jp.empty()) jp.swap_in(counter, jobs_u4, nullptr, n_items, lane);   // the next block (none left: count = 0)
        }
        if (!__any_sync(0xffffffffu, have)) {
            if (jp.empty()) break;
            continue;
        }
        bw.top_up(have);

        const bool act = have;
        const uint32_t chunk = widx >> 2;
        uint32_t q[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t pk = 0;
            if (act && widx < bvw) {
                if (widx >= nbw) {   // region change (at most twice per granule-channel)
                    const bool one = widx < r2w;
                    const uint32_t par = one ? par1 : par2;
                    cur_base = lut + ((par & 0xFFFFu) << 2); cur_sh = (par >> 16) & 31u; cur_lin = par >> 24;
                    nbw = one ? r2w : 0x7FFFFFFFu;
                }
                const uint32_t bits = bw.peek32();
                uint32_t e = lds32(cur_base + ((bits >> cur_sh) << 2));
                // codes longer than the root table: walk the sub-tables inside the same 32-bit window (the longest
                // code is 19 bits, plus two sign bits)
                uint32_t used = 0;
                while ((int32_t)e < 0) {
                    used += (e >> 22) & 31u;
                    e = lds32(cur_base + (((e & 0xFFFFu) + __funnelshift_r(bits << used, 0u, e >> 16)) << 2));
                }
                if (e & 0x40000000u) {      // linbits escapes (minimp3.d:805-813), out of line
                    int a0 = (int)(e & 15u), a1 = (int)((e >> 16) & 15u);
                    bw.advance(used + ((e >> 8) & 15u));
                    const uint32_t b = bw.peek32();   // linbits, sign, linbits, sign: at most 28 bits
                    uint32_t n = 0;
                    if (a0 == 15) { a0 += (int)(b >> (32u - cur_lin)); n = cur_lin; }
                    if (a0) { if ((b << n) >> 31) a0 = -a0; n++; }
                    if (a1 == 15) { a1 += (int)((b << n) >> (32u - cur_lin)); n += cur_lin; }
                    if (a1) { if ((b << n) >> 31) a1 = -a1; n++; }
                    bw.advance(n);
                    pk = __byte_perm((uint32_t)a0, (uint32_t)a1, 0x5410);
                } else {
                    // the code ends `used + len` bits into the window; the sign bits of the non-zero values follow
                    const uint32_t sb = __funnelshift_l(0u, bits << used, e >> 8);   // bits << (used + len)
                    const uint32_t s0 = (sb >> 31) & (e >> 24);                     // sign of a0 if a0 != 0
                    const uint32_t sb1 = __funnelshift_l(0u, sb, e >> 24);           // skip that bit if it was taken
                    const uint32_t s1 = (sb1 >> 31) & (e >> 15) & 1u;               // sign of a1 if a1 != 0
                    const uint32_t inc = s0 | (s1 << 16);
                    pk = ((e & 0x000F000Fu) ^ (inc * 0xFFFFu)) + inc;                // negate the flagged halves (never zero)
                    bw.advance(used + ((e >> 4) & 15u));                             // code + sign bits in one step
                }
                widx++;
            }
            q[k] = pk;
        }
        if (act) {
            is_base[(uint64_t)item * kIsChunks + chunk] = make_uint4(q[0], q[1], q[2], q[3]);
            if (widx >= bvw) {   // through: park; the hand-over to count1 is written when the lane refills
                have = false;
                parked = true;
                last[0] = q[0]; last[1] = q[1]; last[2] = q[2]; last[3] = q[3];
            }
        }
    }
    if (parked) hand_over();
    cp_async_wait_all();
}

// ---------------------------------------------------------------------------------------------------------------
// count1: one quad per step; output words go through a per-lane ring in shared memory because a quad region starts
// at any pair index (big_values is not a multiple of four pairs)
// ---------------------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(128) l3_huff_c1_kernel(BatchParams p, uint32_t* counter) {
    __shared__ uint16_t s_code[128];
    __shared__ uint2 s_val[256];
    __shared__ uint32_t s_ring[4][8][32];
    __shared__ uint32_t s_words[4][kRingWords][32];
    __shared__ uint4 s_pool[4][128];
    for (uint32_t i = threadIdx.x; i < 128; i += blockDim.x) s_code[i] = p.t.c1code[i];
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) s_val[i] = p.t.c1val[i];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* const ring = &s_ring[warp][0][lane];   // slot k of this lane: ring[k * 32]
    const uint32_t n_items = (uint32_t)(p.grch_hi - p.grch_lo);
    const uint32_t* const blob32 = reinterpret_cast<const uint32_t*>(p.blob);
    const HuffJob* const jobs = p.jobs + p.grch_lo;
    const uint4* const descs = reinterpret_cast<const uint4*>(p.grch + p.grch_lo);
    uint4* const is_base = p.is + p.grch_lo * kIsChunks;
    uint8_t* const sf_base = p.sf + p.grch_lo * kSfRecBytes;
    uint8_t* const nzc_base = p.nzc + p.grch_lo;

    BitWindow bw;
    bw.words = blob32; bw.nwords = 1; bw.pos = 0; bw.w0 = bw.w1 = 0; bw.wn = bw.ready = bw.filled = 0;
    bw.ring = smem_addr(&s_words[warp][0][lane]);
    uint32_t item = 0, widx = 0, limit = 0, cbase = 0;
    bool have = false, fin = false, parked = false;
    const uint4* const jobs_u4 = reinterpret_cast<const uint4*>(jobs);
    JobPool<true> jp;
    jp.init(&s_pool[warp][0], counter, jobs_u4, descs, n_items, lane);

    auto flush = [&](uint32_t chunk) {
        const uint32_t* r = ring + (chunk & 1u) * 128u;
        is_base[(uint64_t)item * kIsChunks + chunk] = make_uint4(r[0], r[32], r[64], r[96]);
    };
    // tail of a finished granule-channel: pad and write the last chunk, record the number of chunks.  Done when the
    // lane refills (several lanes at once) instead of in the trip it finished in.
    auto finish = [&]() {
        if (widx & 3u) {   // pad the last chunk with zeros
            for (uint32_t k = widx & 3u; k < 4; k++) ring[((widx & 4u) + k) * 32u] = 0u;
            flush(widx >> 2);
        }
        const uint32_t chunks = (widx + 3u) >> 2;
        *reinterpret_cast<uint16_t*>(sf_base + (uint64_t)item * kSfRecBytes + 80) = (uint16_t)chunks;
        nzc_base[item] = (uint8_t)chunks;
        if (p.zero_fill)
            for (uint32_t c = chunks; c < (uint32_t)kIsChunks; c++) is_base[(uint64_t)item * kIsChunks + c] = make_uint4(0, 0, 0, 0);
        parked = false;
    };

    for (;;) {
        const unsigned need = __ballot_sync(0xffffffffu, !have);
        if (__popc(need) >= K && !jp.empty()) {
            cp_async_wait_all();   // nothing of a finished job may still be landing in a ring that is about to be re-primed
            if (parked) finish();
            const int slot = jp.draw(need, !have, lane);
            if (slot >= 0) {
                const uint4 a = jp.pool[3 * slot], part = jp.pool[3 * slot + 1], w = jp.pool[3 * slot + 2];
                const uint4 dsc = jp.pool[96 + slot];
                item = jp.base + (uint32_t)slot;
                have = true;
                fin = false;
                widx = (dsc.y >> 12) & 0x1FFu;                 // big_values, in pairs
                limit = a.w;
                cbase = ((dsc.z >> 26) & 1u) * 64u;            // count1_table
                bw.open(blob32 + a.x, a.y, a.z, w);
                if (widx & 3u) {   // the big_values kernel left a partial chunk: continue inside it
                    uint32_t* r = ring + (widx & 4u) * 32u;
                    r[0] = part.x; r[32] = part.y; r[64] = part.z; r[96] = part.w;
                }
            }
            if (jp.empty()) jp.swap_in(counter, jobs_u4, descs, n_items, lane);
        }
        if (!__any_sync(0xffffffffu, have)) {
            if (jp.empty()) break;
            continue;
        }
        bw.top_up(have);
#pragma unroll 2
        for (int k = 0; k < 4; k++) {
            if (have && !fin) {
                const uint32_t bits = bw.peek32();
                const uint32_t e = s_code[cbase + (bits >> 26)];
                const uint32_t len = e & 15u;
                // the limit is tested after the code and before the signs (minimp3.d:866); the band terminator before
                // each half of the quad (minimp3.d:873, 876)
                if (bw.pos + len > limit || widx >= 288u) {
                    fin = true;
                } else {
                    const uint32_t sb = bits << len;
                    const uint2 v = s_val[(e & 0xF0u) | (sb >> 28)];
                    const uint32_t w0i = widx;
                    ring[(w0i & 7u) * 32u] = v.x;
                    widx++;
                    if (widx < 288u) {
                        ring[(widx & 7u) * 32u] = v.y;
                        widx++;
                    }
                    if ((widx ^ w0i) & ~3u) flush(w0i >> 2);
                    bw.advance(e >> 8);
                }
            }
        }
        if (have && fin) { have = false; parked = true; }   // finished: the tail is written when the lane refills
    }
    if (parked) finish();
    cp_async_wait_all();
}

// Launch configuration of the two Huffman kernels (one variant each in the product library):
//   lanes refill in groups of K = 16: measured (big_values / count1 ms) K=4: 4.68/2.46, 8: 4.61/2.47, 16: 4.59/2.42
//   big_values: 8-warp CTAs share one copy of the LUT (more resident warps per SM: the kernel is latency-bound), as many
//               CTAs per SM as fit; count1: fastest at 4 CTAs per SM (2: 3.4 ms, 4: 2.4, 6: 2.7, 10: 3.1)
constexpr int kHuffK = 16, kBigWarps = 8, kC1CtasPerSm = 4;

int launch_entropy_v4(const BatchParams& p, int sub, int sms, cudaStream_t s) {
    if (p.grch_hi <= p.grch_lo) return 0;
    const uint64_t n = p.grch_hi - p.grch_lo;
    l3_scf_kernel<<<(unsigned)((n + 127) / 128), 128, 4 * 32 * kRecRow * sizeof(uint4), s>>>(p);
    uint32_t* cnt = p.counters + 2 * sub;
    {
        // attribute and occupancy are per device and a process may hold contexts on several GPUs: asked every time
        // (two cheap driver calls per launch)
        const size_t smem = (size_t)((p.t.huff32_entries + 3u) & ~3u) * 4 + kBigWarps * (96 * sizeof(uint4) + kRingWords * 32 * 4);
        int occ = 0;
        cudaFuncSetAttribute(l3_huff_big_kernel<kHuffK, kBigWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, l3_huff_big_kernel<kHuffK, kBigWarps>, 32 * kBigWarps, smem);
        occ = std::max(1, occ);
        const unsigned blocks = (unsigned)std::min<uint64_t>((n + 32 * kBigWarps - 1) / (32 * kBigWarps), (uint64_t)sms * occ);
        l3_huff_big_kernel<kHuffK, kBigWarps><<<blocks, 32 * kBigWarps, smem, s>>>(p, cnt);
    }
    const unsigned blocks_c1 = (unsigned)std::min<uint64_t>((n + 127) / 128, (uint64_t)sms * kC1CtasPerSm);
    l3_huff_c1_kernel<kHuffK><<<blocks_c1, 128, 0, s>>>(p, cnt + 1);
    return 3;
}

}  // namespace l3b
