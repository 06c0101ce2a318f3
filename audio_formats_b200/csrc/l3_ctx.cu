// l3_ctx.cu -- C-ABI layer 1 ("shim"): context, batch upload/run/download.  No CPU fallback.
#include <cuda_runtime.h>
#include <sched.h>

#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <new>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/l3b200.h"

#include "l3_device_tables.hpp"
#include "l3_host.hpp"
#include "l3_kernels.cuh"

using namespace l3b;

static thread_local std::string g_create_error;

struct l3b_ctx {
    int device = -1;
    cudaStream_t stream = nullptr;
    std::string err;
    void* d_tables = nullptr;  // one allocation holding every lookup table
    DeviceTables t{};
    cudaStream_t stream2 = nullptr;   // the granule kernels of a run go here, overlapping later entropy launches
    static constexpr int kRing = 64;  // runs whose events are kept
    static constexpr int kMaxSubs = 16;
    // per run: [0] start, [1] last entropy launch done, [2] all done; then 2 per sub-batch around the granule launches
    cudaEvent_t ev[kRing * (3 + 3 * kMaxSubs)] = {};
    int subs_of_run[kRing] = {};
    uint64_t runs = 0;
    int last_launches = 0;
    int sms = 148;                    // multiprocessors of THIS device (grid sizing of the persistent Huffman kernels)
    cudaEvent_t ev_wait = nullptr;    // cudaEventBlockingSync: host waits sleep instead of spinning (one core per waiting lane otherwise)
    // one recycled stream workspace: an AudioStream that closes leaves its device buffers here for the next one that opens
    // (a dozen cudaMalloc / cudaFree calls cost more than decoding a short file)
    std::mutex spare_mutex;
    struct l3b_resident* spare = nullptr;
};

// Wait for everything queued on the context stream without burning a core: a pipeline keeps several lanes per GPU
// waiting on their copies at any time, and a spinning cudaStreamSynchronize takes a host core each.
static cudaError_t ctx_wait(l3b_ctx* c) {
    cudaError_t e = cudaEventRecord(c->ev_wait, c->stream);
    if (e != cudaSuccess) return e;
    return cudaEventSynchronize(c->ev_wait);
}

struct l3b_resident {
    uint8_t* d_blob = nullptr;
    l3b_grch_desc_t* d_grch = nullptr;
    l3b_stream_desc_t* d_streams = nullptr;
    uint4* d_is = nullptr;
    uint8_t* d_sf = nullptr;
    float* d_pcm = nullptr;                 // float PCM, or int16 PCM when flags & L3B_OUT_S16 (cap_pcm counts elements)
    uint8_t* d_nzc = nullptr;
    float* d_ftaps = nullptr;               // 4 x [n_grch][576] float stage snapshots (tap mode with float taps only)
    uint32_t flags = 0;
    float* d_l12x = nullptr;                // Layer I / II: dequantised subband samples, [n_grch][384] (only when such streams exist)
    bool has_l12 = false;
    Tile* d_tiles[4] = {nullptr, nullptr, nullptr, nullptr};  // [0] stereo, [1] mono (Layer III); [2] stereo, [3] mono (Layer I / II)
    HuffJob* d_jobs = nullptr;              // one per granule-channel, written by the scalefactor kernel
    uint32_t* d_group_stream = nullptr;     // stream index of granule-channel 128 k, for every k (search hint)
    uint32_t* d_counters = nullptr;         // 2 per sub-batch: item counters of the two Huffman kernels
    uint32_t n_tiles[4] = {0, 0, 0, 0};
    uint64_t n_grch = 0, pcm_floats = 0;
    uint32_t n_streams = 0;
    // capacities of the device buffers (elements), for l3b_batch_upload_reuse
    uint64_t cap_blob = 0, cap_grch = 0, cap_pcm = 0;
    uint32_t cap_streams = 0, cap_tiles[4] = {0, 0, 0, 0};
    struct Sub { uint64_t grch_lo, grch_hi; uint32_t tile_lo[4], tile_hi[4]; };
    std::vector<Sub> subs;   // the run is issued sub-batch by sub-batch (stream boundaries)
    BatchParams params{};
};

#define CU_TRY(ctx, expr)                                                                     \
    do {                                                                                      \
        cudaError_t e_ = (expr);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(e_);                  \
            return e_ == cudaErrorMemoryAllocation ? L3B_E_MEMORY : L3B_E_NOGPU;              \
        }                                                                                     \
    } while (0)

extern "C" {

int l3b_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* l3b_last_error(const l3b_ctx_t* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int l3b_ctx_create(int device_id, l3b_ctx_t** out) {
    if (!out) return L3B_E_PARAM;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                         " (this library has no CPU fallback)";
        cudaGetLastError();
        return L3B_E_NOGPU;
    }
    if (device_id < 0 || device_id >= n) { g_create_error = "device_id out of range"; return L3B_E_PARAM; }
    l3b_ctx* c = new (std::nothrow) l3b_ctx();
    if (!c) return L3B_E_MEMORY;
    c->device = device_id;
    auto fail = [&](const char* what, cudaError_t err) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(err);
        delete c;
        return L3B_E_NOGPU;
    };
    if ((e = cudaSetDevice(device_id)) != cudaSuccess) return fail("cudaSetDevice", e);
    cudaDeviceGetAttribute(&c->sms, cudaDevAttrMultiProcessorCount, device_id);
    if ((e = cudaEventCreateWithFlags(&c->ev_wait, cudaEventBlockingSync | cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
    {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);   // entropy launches get the higher priority: they fill idle slots
        if ((e = cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, hi)) != cudaSuccess) return fail("cudaStreamCreate", e);
        if ((e = cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, lo)) != cudaSuccess) return fail("cudaStreamCreate", e);
    }

    // lookup tables: one device allocation, sub-allocated at 256-byte granularity
    HuffLut hl = build_huff_lut();
    HuffLut32 hl32 = build_huff_lut32();
    SfbMaps* sm = new SfbMaps();
    build_sfb_maps(sm);
    std::vector<uint8_t> img;
    auto put = [&](const void* src, size_t bytes) {
        size_t at = (img.size() + 255) & ~(size_t)255;
        img.resize(at + bytes);
        memcpy(img.data() + at, src, bytes);
        return at;
    };
    size_t o_huff = put(hl.entries.data(), hl.entries.size() * 2);
    size_t o_c1 = put(hl.count1, sizeof hl.count1);
    size_t o_huff32 = put(hl32.entries.data(), hl32.entries.size() * 4);
    size_t o_c1code = put(hl32.c1code, sizeof hl32.c1code);
    size_t o_c1val = put(hl32.c1val, sizeof hl32.c1val);
    size_t o_pair = put(sm->sfb_of_pair, sizeof sm->sfb_of_pair);
    size_t o_w = put(sm->width, sizeof sm->width);
    size_t o_s = put(sm->start, sizeof sm->start);
    size_t o_perm = put(sm->perm, sizeof sm->perm);
    size_t o_pow = put(L3_POW43, sizeof L3_POW43);
    size_t o_win = put(L3_WIN, sizeof L3_WIN);
    delete sm;
    if ((e = cudaMalloc(&c->d_tables, img.size())) != cudaSuccess) return fail("cudaMalloc(tables)", e);
    if ((e = cudaMemcpy(c->d_tables, img.data(), img.size(), cudaMemcpyHostToDevice)) != cudaSuccess) return fail("cudaMemcpy(tables)", e);
    uint8_t* base = static_cast<uint8_t*>(c->d_tables);
    c->t.huff = reinterpret_cast<const uint16_t*>(base + o_huff);
    c->t.huff_entries = (uint32_t)hl.entries.size();
    for (int i = 0; i < 18; i++) { c->t.huff_base[i] = hl.base[i]; c->t.huff_root[i] = hl.root_bits[i]; }
    c->t.count1 = base + o_c1;
    c->t.huff32 = reinterpret_cast<const uint32_t*>(base + o_huff32);
    c->t.huff32_entries = (uint32_t)hl32.entries.size();
    for (int i = 0; i < 16; i++) { c->t.huff32_base[i] = hl32.base[i]; c->t.huff32_root[i] = hl32.root_bits[i]; }
    c->t.c1code = reinterpret_cast<const uint16_t*>(base + o_c1code);
    c->t.c1val = reinterpret_cast<const uint2*>(base + o_c1val);
    c->t.sfb_of_pair = base + o_pair;
    c->t.sfb_width = base + o_w;
    c->t.sfb_start = reinterpret_cast<const uint16_t*>(base + o_s);
    c->t.perm = reinterpret_cast<const uint16_t*>(base + o_perm);
    c->t.pow43 = reinterpret_cast<const float*>(base + o_pow);
    c->t.win = reinterpret_cast<const float*>(base + o_win);
    upload_constants();
    upload_entropy_constants();
    upload_l12_constants();
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) return fail("constant upload", e);
    *out = c;
    return 0;
}

void l3b_ctx_destroy(l3b_ctx_t* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->spare) { l3b_batch_free(c, c->spare); c->spare = nullptr; }
    if (c->d_tables) cudaFree(c->d_tables);
    for (auto& ev : c->ev)
        if (ev) cudaEventDestroy(ev);
    if (c->ev_wait) cudaEventDestroy(c->ev_wait);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->stream2) cudaStreamDestroy(c->stream2);
    delete c;
}

void* l3b_ctx_cuda_stream(l3b_ctx_t* c) { return c ? (void*)c->stream : nullptr; }

void* l3b_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

// Page placement: cudaHostAlloc pins pages where the allocating thread's policy puts them (first touch, local node), so
// the allocation is made by a thread that runs on the CPUs of the GPU's NUMA node for the duration of the call.
void* l3b_host_alloc_near(int device_id, size_t bytes) {
    cpu_set_t old_set, node_set;
    bool moved = false;
    char bus[32] = {0};
    if (sched_getaffinity(0, sizeof old_set, &old_set) == 0 && cudaDeviceGetPCIBusId(bus, sizeof bus, device_id) == cudaSuccess) {
        for (char* q = bus; *q; q++) *q = (char)tolower(*q);
        char path[128];
        snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
        int node = -1;
        if (FILE* f = fopen(path, "r")) { if (fscanf(f, "%d", &node) != 1) node = -1; fclose(f); }
        if (node >= 0) {
            snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
            if (FILE* f = fopen(path, "r")) {
                CPU_ZERO(&node_set);
                int a, b2, n_set = 0;
                char sep;
                while (fscanf(f, "%d", &a) == 1) {
                    b2 = a;
                    int ch = fgetc(f);
                    if (ch == '-') { if (fscanf(f, "%d", &b2) != 1) b2 = a; ch = fgetc(f); }
                    for (int k = a; k <= b2 && k < CPU_SETSIZE; k++)
                        if (CPU_ISSET(k, &old_set)) { CPU_SET(k, &node_set); n_set++; }
                    sep = (char)ch;
                    if (sep != ',') break;
                }
                fclose(f);
                if (n_set > 0 && sched_setaffinity(0, sizeof node_set, &node_set) == 0) moved = true;
            }
        }
    } else {
        cudaGetLastError();
    }
    void* p = l3b_host_alloc(bytes);
    if (p && moved) {   // touch the pages while still on the node (pinning has usually placed them already)
        volatile uint8_t* q = static_cast<volatile uint8_t*>(p);
        for (size_t i = 0; i < bytes; i += 4096) q[i] = 0;
    }
    if (moved) sched_setaffinity(0, sizeof old_set, &old_set);
    return p;
}

void l3b_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

void l3b_batch_free(l3b_ctx_t* c, l3b_resident_t* r) {
    if (!r) return;
    if (c) cudaSetDevice(c->device);
    cudaFree(r->d_blob);
    cudaFree(r->d_grch);
    cudaFree(r->d_streams);
    cudaFree(r->d_is);
    cudaFree(r->d_sf);
    cudaFree(r->d_pcm);
    cudaFree(r->d_nzc);
    cudaFree(r->d_ftaps);
    cudaFree(r->d_l12x);
    for (int k = 0; k < 4; k++) cudaFree(r->d_tiles[k]);
    cudaFree(r->d_jobs);
    cudaFree(r->d_group_stream);
    cudaFree(r->d_counters);
    delete r;
}

// device_inputs: the blob and the descriptors are produced on the device (l3_raw.cu), only the stream table comes from the host
static int upload_impl(l3b_ctx_t* c, const l3b_batch_t* b, l3b_resident_t** inout, bool reuse, bool device_inputs = false) {
    if (!c || !b || !inout) return L3B_E_PARAM;
    if (!reuse) *inout = nullptr;
    if (!b->n_streams || !b->streams || (!device_inputs && ((b->n_grch && !b->grch) || (b->maindata_bytes && !b->maindata)))) {
        c->err = "empty or inconsistent batch";
        return L3B_E_PARAM;
    }
    if (b->n_grch >= (1ull << 31)) { c->err = "more than 2^31 granule-channels in one batch; split it into waves"; return L3B_E_PARAM; }
    // Granules per warp tile of the granule kernel.  Every tile recomputes a halo of two or three granules, so long tiles
    // waste less (128: 16.8 ms, 64: 16.9 ms on BASELINE config 2), but a batch needs a couple of tiles per resident warp to
    // fill the device and a lone stream is decoded sooner in short ones: the longest of 128 / 64 / 32 / 16 that still gives
    // two tiles per warp slot (16 warps per SM), else 16.
    uint32_t tile_granules = 16;
    uint64_t total_granules = 0;
    {
        uint64_t& total = total_granules;
        for (uint32_t i = 0; i < b->n_streams; i++) {
            const l3b_stream_desc_t& s = b->streams[i];
            const uint64_t per = ((s.layer == 1 || s.layer == 2) ? 384u : 576u) * (uint64_t)(s.nch == 2 ? 2 : 1);
            if (s.pcm_count) total += (s.pcm_skip + s.pcm_count + per - 1) / per - s.pcm_skip / per;
        }
        for (uint32_t t = (uint32_t)kTileGranulesMax; t > 16; t >>= 1)
            if (total / t >= 2ull * 16ull * (uint64_t)c->sms) { tile_granules = t; break; }
    }
    // validate stream table and build the tile lists
    std::vector<Tile> tiles[4];
    uint64_t expect_grch = 0, granules_before = 0;
    bool has_l12 = false;
    for (uint32_t i = 0; i < b->n_streams; i++) {
        const l3b_stream_desc_t& s = b->streams[i];
        const bool l12 = s.layer == 1 || s.layer == 2;
        const uint64_t gran = l12 ? 384u : 576u;   // PCM frames per granule
        if ((s.nch != 1 && s.nch != 2) || s.sr_idx > 7 || (s.maindata_off & 3) || (s.layer != 0 && s.layer != 3 && !l12) ||
            s.maindata_off + s.maindata_bytes > b->maindata_bytes || s.first_grch != expect_grch ||
            s.pcm_off + s.pcm_count > b->pcm_floats || s.pcm_skip + s.pcm_count > (uint64_t)s.n_granules * gran * s.nch ||
            (s.nch == 2 && ((s.pcm_off | s.pcm_skip | s.pcm_count) & 1))) {
            c->err = "stream descriptor " + std::to_string(i) + " is inconsistent";
            return L3B_E_PARAM;
        }
        expect_grch += (uint64_t)s.n_granules * s.nch;
        has_l12 |= l12;
        if (!s.pcm_count) continue;
        const uint64_t per = gran * s.nch;
        uint32_t g0 = (uint32_t)(s.pcm_skip / per), g1 = (uint32_t)((s.pcm_skip + s.pcm_count + per - 1) / per);
        // The tiles are taken in table order, one per warp, and a warp that starts late finishes late: the last eighth of the
        // batch goes into half-length tiles and the last twentieth into quarter-length ones, so that the device drains in a
        // quarter of a tile's time instead of a whole one (the extra halo granules cost less than the idle SMs did).
        uint32_t tg = tile_granules;
        if (granules_before * 20 >= total_granules * 19) tg = std::max(16u, tile_granules / 4);
        else if (granules_before * 8 >= total_granules * 7) tg = std::max(16u, tile_granules / 2);
        granules_before += g1 - g0;
        for (uint32_t g = g0; g < g1; g += tg)
            tiles[(l12 ? 2 : 0) + (s.nch == 2 ? 0 : 1)].push_back({i, g, std::min<uint32_t>(tg, g1 - g)});
    }
    if (expect_grch != b->n_grch) { c->err = "n_grch does not match the stream table"; return L3B_E_PARAM; }
    // sub-batches: cut at stream boundaries into up to kMaxSubs pieces of similar size (>= 64 K granule-channels each)
    std::vector<l3b_resident::Sub> subs;
    {
        int want = 1;  // measured on B200: overlapping the two kernels through sub-batches does not pay (39.2 ms for 1, 39.5 for 4, 41.3 for 16)
        if (getenv("L3B_SUBBATCHES")) want = std::max(1, std::min((int)l3b_ctx::kMaxSubs, atoi(getenv("L3B_SUBBATCHES"))));
        uint64_t per = (b->n_grch + want - 1) / want, lo = 0;
        uint32_t t_at[4] = {0, 0, 0, 0};
        l3b_resident::Sub cur{0, 0, {0, 0, 0, 0}, {0, 0, 0, 0}};
        for (uint32_t i = 0; i < b->n_streams; i++) {
            const l3b_stream_desc_t& s = b->streams[i];
            uint64_t hi = s.first_grch + (uint64_t)s.n_granules * s.nch;
            for (int k = 0; k < 4; k++)
                while (t_at[k] < tiles[k].size() && tiles[k][t_at[k]].stream == i) t_at[k]++;
            if (hi - lo >= per || i + 1 == b->n_streams) {
                cur.grch_lo = lo; cur.grch_hi = hi;
                for (int k = 0; k < 4; k++) cur.tile_hi[k] = t_at[k];
                subs.push_back(cur);
                for (int k = 0; k < 4; k++) cur.tile_lo[k] = t_at[k];
                lo = hi;
            }
        }
    }

    CU_TRY(c, cudaSetDevice(c->device));
    l3b_resident* r = *inout;
    if (!r) {
        r = new (std::nothrow) l3b_resident();
        if (!r) return L3B_E_MEMORY;
    }
    auto bail = [&](int code) { l3b_batch_free(c, r); *inout = nullptr; return code; };
#define CU_TRY_R(expr)                                                                        \
    do {                                                                                      \
        cudaError_t e_ = (expr);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            c->err = std::string(#expr) + ": " + cudaGetErrorString(e_);                      \
            return bail(e_ == cudaErrorMemoryAllocation ? L3B_E_MEMORY : L3B_E_NOGPU);        \
        }                                                                                     \
    } while (0)
    // (re)allocate what is too small; recycled workspaces get 12.5 % headroom so that they settle quickly
    auto grow = [&](uint64_t need) { return reuse ? need + need / 8 + 64 : need; };
    bool stale = false;  // a buffer in use by queued work is about to be freed
    if (b->maindata_bytes + 64 > r->cap_blob || b->n_grch > r->cap_grch || b->pcm_floats > r->cap_pcm || b->n_streams > r->cap_streams)
        stale = true;
    for (int k = 0; k < 4; k++)
        if (tiles[k].size() > r->cap_tiles[k]) stale = true;
    const uint32_t flags = b->flags;
    const size_t pcm_elem = (flags & L3B_OUT_S16) ? sizeof(int16_t) : sizeof(float);
    const bool pcm_kind_changed = ((r->flags ^ flags) & L3B_OUT_S16) != 0;
    if (pcm_kind_changed) stale = true;
    if (stale && r->cap_blob) CU_TRY_R(ctx_wait(c));
    if (b->maindata_bytes + 64 > r->cap_blob) {
        cudaFree(r->d_blob); r->d_blob = nullptr;
        r->cap_blob = grow(b->maindata_bytes + 64);
        CU_TRY_R(cudaMalloc(&r->d_blob, r->cap_blob));
    }
    if (b->n_grch > r->cap_grch || !r->d_grch) {
        cudaFree(r->d_grch); cudaFree(r->d_is); cudaFree(r->d_sf); cudaFree(r->d_jobs); cudaFree(r->d_group_stream); cudaFree(r->d_nzc);
        cudaFree(r->d_ftaps); cudaFree(r->d_l12x);
        r->d_l12x = nullptr;
        r->d_grch = nullptr; r->d_is = nullptr; r->d_sf = nullptr; r->d_jobs = nullptr; r->d_group_stream = nullptr; r->d_nzc = nullptr;
        r->d_ftaps = nullptr;
        r->cap_grch = std::max<uint64_t>(1, grow(b->n_grch));
        CU_TRY_R(cudaMalloc(&r->d_grch, r->cap_grch * sizeof(l3b_grch_desc_t)));
        CU_TRY_R(cudaMalloc(&r->d_is, r->cap_grch * kIsChunks * sizeof(uint4)));
        CU_TRY_R(cudaMalloc(&r->d_sf, r->cap_grch * kSfRecBytes));
        CU_TRY_R(cudaMalloc(&r->d_jobs, r->cap_grch * sizeof(HuffJob)));
        CU_TRY_R(cudaMalloc(&r->d_group_stream, (r->cap_grch / 128 + 2) * sizeof(uint32_t)));
        CU_TRY_R(cudaMalloc(&r->d_nzc, r->cap_grch + 16));
    }
    r->has_l12 = has_l12;
    if (has_l12 && !r->d_l12x) CU_TRY_R(cudaMalloc(&r->d_l12x, r->cap_grch * 384 * sizeof(float)));
    const bool want_ftaps = b->taps && (b->taps->xr || b->taps->st || b->taps->im || b->taps->dct);
    if (want_ftaps && (flags & (L3B_OUT_S16 | L3B_MATH_FUSED))) {
        c->err = "float taps exist in the bit-exact float-delivery mode only";
        return bail(L3B_E_PARAM);
    }
    if (want_ftaps && !r->d_ftaps) {
        CU_TRY_R(cudaMalloc(&r->d_ftaps, r->cap_grch * 4 * 576 * sizeof(float)));
        CU_TRY_R(cudaMemsetAsync(r->d_ftaps, 0, r->cap_grch * 4 * 576 * sizeof(float), c->stream));
    }
    if (!r->d_counters) CU_TRY_R(cudaMalloc(&r->d_counters, 2 * l3b_ctx::kMaxSubs * sizeof(uint32_t)));
    if (b->pcm_floats > r->cap_pcm || !r->d_pcm || pcm_kind_changed) {
        cudaFree(r->d_pcm); r->d_pcm = nullptr;
        r->cap_pcm = std::max<uint64_t>(8, grow(b->pcm_floats));
        CU_TRY_R(cudaMalloc(&r->d_pcm, r->cap_pcm * pcm_elem));
    }
    r->flags = flags;
    if (b->n_streams > r->cap_streams) {
        cudaFree(r->d_streams); r->d_streams = nullptr;
        r->cap_streams = (uint32_t)grow(b->n_streams);
        CU_TRY_R(cudaMalloc(&r->d_streams, r->cap_streams * sizeof(l3b_stream_desc_t)));
    }
    for (int k = 0; k < 4; k++) {
        r->n_tiles[k] = (uint32_t)tiles[k].size();
        if (tiles[k].size() > r->cap_tiles[k]) {
            cudaFree(r->d_tiles[k]); r->d_tiles[k] = nullptr;
            r->cap_tiles[k] = (uint32_t)grow(tiles[k].size());
            CU_TRY_R(cudaMalloc(&r->d_tiles[k], r->cap_tiles[k] * sizeof(Tile)));
        }
    }
    r->n_grch = b->n_grch;
    r->n_streams = b->n_streams;
    r->pcm_floats = b->pcm_floats;
    r->subs = subs;
    CU_TRY_R(cudaMemsetAsync(r->d_blob + b->maindata_bytes, 0, 64, c->stream));
    if (b->maindata_bytes && !device_inputs) CU_TRY_R(cudaMemcpyAsync(r->d_blob, b->maindata, b->maindata_bytes, cudaMemcpyHostToDevice, c->stream));
    if (b->n_grch && !device_inputs) CU_TRY_R(cudaMemcpyAsync(r->d_grch, b->grch, b->n_grch * sizeof(l3b_grch_desc_t), cudaMemcpyHostToDevice, c->stream));
    CU_TRY_R(cudaMemcpyAsync(r->d_streams, b->streams, b->n_streams * sizeof(l3b_stream_desc_t), cudaMemcpyHostToDevice, c->stream));
    for (int k = 0; k < 4; k++)
        if (r->n_tiles[k])
            CU_TRY_R(cudaMemcpyAsync(r->d_tiles[k], tiles[k].data(), tiles[k].size() * sizeof(Tile), cudaMemcpyHostToDevice, c->stream));
    std::vector<uint32_t> group_stream((size_t)(b->n_grch / 128 + 1));
    {
        uint32_t si = 0;
        for (size_t g = 0; g < group_stream.size(); g++) {
            const uint64_t gi = (uint64_t)g * 128;
            while (si + 1 < b->n_streams && b->streams[si + 1].first_grch <= gi) si++;
            group_stream[g] = si;
        }
    }
    CU_TRY_R(cudaMemcpyAsync(r->d_group_stream, group_stream.data(), group_stream.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    CU_TRY_R(ctx_wait(c));  // the host tile vectors go out of scope
#undef CU_TRY_R
    BatchParams& p = r->params;
    p.blob = r->d_blob;
    p.grch = r->d_grch;
    p.n_grch = b->n_grch;
    p.streams = r->d_streams;
    p.n_streams = b->n_streams;
    p.is = r->d_is;
    p.sf = r->d_sf;
    p.pcm = (flags & L3B_OUT_S16) ? nullptr : r->d_pcm;
    p.pcm16 = (flags & L3B_OUT_S16) ? reinterpret_cast<int16_t*>(r->d_pcm) : nullptr;
    p.nzc = r->d_nzc;
    p.l12_x = has_l12 ? r->d_l12x : nullptr;
    p.tap_xr = p.tap_st = p.tap_im = p.tap_dct = nullptr;
    if (want_ftaps) {
        p.tap_xr = r->d_ftaps;
        p.tap_st = r->d_ftaps + r->cap_grch * 576;
        p.tap_im = r->d_ftaps + 2 * r->cap_grch * 576;
        p.tap_dct = r->d_ftaps + 3 * r->cap_grch * 576;
    }
    p.jobs = r->d_jobs;
    p.group_stream = r->d_group_stream;
    p.counters = r->d_counters;
    p.zero_fill = b->taps ? 1 : 0;
    p.t = c->t;
    *inout = r;
    return 0;
}

int l3b_batch_upload(l3b_ctx_t* c, const l3b_batch_t* b, l3b_resident_t** out) { return upload_impl(c, b, out, false); }
int l3b_batch_upload_reuse(l3b_ctx_t* c, const l3b_batch_t* b, l3b_resident_t** inout) { return upload_impl(c, b, inout, true); }

int l3b_batch_reupload(l3b_ctx_t* c, l3b_resident_t* r, const l3b_batch_t* b) {
    if (!c || !r || !b) return L3B_E_PARAM;
    if (b->n_grch != r->n_grch || b->n_streams != r->n_streams || b->pcm_floats != r->pcm_floats) {
        c->err = "reupload: batch shape differs from the resident batch";
        return L3B_E_PARAM;
    }
    CU_TRY(c, cudaSetDevice(c->device));
    if (b->maindata_bytes) CU_TRY(c, cudaMemcpyAsync(r->d_blob, b->maindata, b->maindata_bytes, cudaMemcpyHostToDevice, c->stream));
    if (b->n_grch) CU_TRY(c, cudaMemcpyAsync(r->d_grch, b->grch, b->n_grch * sizeof(l3b_grch_desc_t), cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaMemcpyAsync(r->d_streams, b->streams, b->n_streams * sizeof(l3b_stream_desc_t), cudaMemcpyHostToDevice, c->stream));
    return 0;
}

static cudaEvent_t* run_events(l3b_ctx* c, uint64_t run) {
    cudaEvent_t* ev = c->ev + (run % l3b_ctx::kRing) * (3 + 3 * l3b_ctx::kMaxSubs);
    if (!ev[0])
        for (int i = 0; i < 3 + 3 * l3b_ctx::kMaxSubs; i++) cudaEventCreate(&ev[i]);
    return ev;
}

int l3b_batch_run(l3b_ctx_t* c, l3b_resident_t* r) {
    if (!c || !r) return L3B_E_PARAM;
    CU_TRY(c, cudaSetDevice(c->device));
    cudaEvent_t* ev = run_events(c, c->runs);
    cudaStream_t A = c->stream, B = c->stream2;
    const int ns = (int)r->subs.size();
    CU_TRY(c, cudaEventRecord(ev[0], A));
    CU_TRY(c, cudaStreamWaitEvent(B, ev[0], 0));   // uploads on A are complete before anything on B starts
    int launches = 0;
    CU_TRY(c, cudaMemsetAsync(r->d_counters, 0, 2 * l3b_ctx::kMaxSubs * sizeof(uint32_t), A));
    for (int i = 0; i < ns; i++) {
        const l3b_resident::Sub& sb = r->subs[i];
        BatchParams p = r->params;
        p.grch_lo = sb.grch_lo;
        p.grch_hi = sb.grch_hi;
        launches += launch_entropy_v4(p, i, c->sms, A);
        if (r->has_l12) {   // Layer I / II streams of this sub-batch: unallocated subbands stay +0 (the reference's memset grbuf)
            CU_TRY(c, cudaMemsetAsync(r->d_l12x + sb.grch_lo * 384, 0, (sb.grch_hi - sb.grch_lo) * 384 * sizeof(float), A));
            launches += launch_l12_parse(p, A);
        }
        cudaEvent_t* se = ev + 3 + 3 * i;
        CU_TRY(c, cudaEventRecord(se[0], A));
        CU_TRY(c, cudaStreamWaitEvent(B, se[0], 0));   // granule kernels of sub-batch i wait for its spectra only
        CU_TRY(c, cudaEventRecord(se[1], B));
        const uint32_t n2 = sb.tile_hi[0] - sb.tile_lo[0], n1 = sb.tile_hi[1] - sb.tile_lo[1];
        CU_TRY(c, launch_granule(p, r->d_tiles[0] + sb.tile_lo[0], n2, r->d_tiles[1] + sb.tile_lo[1], n1, B,
                                 (r->flags & L3B_MATH_FUSED) != 0, p.tap_xr != nullptr));
        launches += (n2 > 0) + (n1 > 0);
        if (r->has_l12) {
            const uint32_t m2 = sb.tile_hi[2] - sb.tile_lo[2], m1 = sb.tile_hi[3] - sb.tile_lo[3];
            CU_TRY(c, launch_granule_l12(p, r->d_tiles[2] + sb.tile_lo[2], m2, r->d_tiles[3] + sb.tile_lo[3], m1, B,
                                         (r->flags & L3B_MATH_FUSED) != 0));
            launches += (m2 > 0) + (m1 > 0);
        }
        CU_TRY(c, cudaEventRecord(se[2], B));
    }
    CU_TRY(c, cudaEventRecord(ev[1], A));
    CU_TRY(c, cudaEventRecord(ev[2], B));
    CU_TRY(c, cudaStreamWaitEvent(A, ev[2], 0));       // downloads / the next upload on A see finished PCM
    CU_TRY(c, cudaGetLastError());
    c->subs_of_run[c->runs % l3b_ctx::kRing] = ns;
    c->last_launches = launches;
    c->runs++;
    return 0;
}

int l3b_batch_sync(l3b_ctx_t* c) {
    if (!c) return L3B_E_PARAM;
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, ctx_wait(c));
    CU_TRY(c, cudaGetLastError());
    return 0;
}

int l3b_batch_timing(l3b_ctx_t* c, int last_runs, float ms[3], int* launches) {
    if (!c || !c->runs || last_runs < 1) return L3B_E_PARAM;
    if ((uint64_t)last_runs > c->runs) last_runs = (int)c->runs;
    if (last_runs > l3b_ctx::kRing) last_runs = l3b_ctx::kRing;
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream2));
    float sum[3] = {0, 0, 0};
    for (int k = 0; k < last_runs; k++) {
        const uint64_t run = c->runs - 1 - k;
        cudaEvent_t* ev = run_events(c, run);
        float t = 0;
        CU_TRY(c, cudaEventElapsedTime(&t, ev[0], ev[1]));   // entropy launches, first to last (stream A)
        sum[0] += t;
        for (int i = 0; i < c->subs_of_run[run % l3b_ctx::kRing]; i++) {
            CU_TRY(c, cudaEventElapsedTime(&t, ev[3 + 3 * i + 1], ev[3 + 3 * i + 2]));   // granule launches of sub-batch i
            sum[1] += t;
        }
        CU_TRY(c, cudaEventElapsedTime(&t, ev[0], ev[2]));   // whole run
        sum[2] += t;
    }
    if (ms) memcpy(ms, sum, sizeof sum);
    if (launches) *launches = c->last_launches * last_runs;
    return 0;
}

int l3b_batch_download(l3b_ctx_t* c, l3b_resident_t* r, void* pcm_host, uint64_t first, uint64_t n) {
    if (!c || !r || (!pcm_host && n) || first + n > r->pcm_floats) return L3B_E_PARAM;
    CU_TRY(c, cudaSetDevice(c->device));
    const size_t el = (r->flags & L3B_OUT_S16) ? sizeof(int16_t) : sizeof(float);
    if (n)
        CU_TRY(c, cudaMemcpyAsync(pcm_host, reinterpret_cast<const uint8_t*>(r->d_pcm) + first * el, n * el, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, ctx_wait(c));
    return 0;
}

int l3b_batch_download_taps(l3b_ctx_t* c, l3b_resident_t* r, const l3b_taps_t* taps) {
    if (!c || !r || !taps) return L3B_E_PARAM;
    if (!r->params.zero_fill) { c->err = "batch was uploaded without taps"; return L3B_E_PARAM; }
    CU_TRY(c, cudaSetDevice(c->device));
    const uint64_t n = r->n_grch;
    if (taps->is && n) CU_TRY(c, cudaMemcpyAsync(taps->is, r->d_is, n * 576 * sizeof(int16_t), cudaMemcpyDeviceToHost, c->stream));
    if ((taps->iscf || taps->ist_pos) && n) {
        std::vector<uint8_t> rec(n * kSfRecBytes);
        CU_TRY(c, cudaMemcpyAsync(rec.data(), r->d_sf, rec.size(), cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(c, cudaStreamSynchronize(c->stream));
        for (uint64_t i = 0; i < n; i++) {
            if (taps->iscf) memcpy(taps->iscf + i * 40, rec.data() + i * kSfRecBytes, 40);
            if (taps->ist_pos) memcpy(taps->ist_pos + i * 40, rec.data() + i * kSfRecBytes + 40, 40);
        }
    }
    float* const dst[4] = {taps->xr, taps->st, taps->im, taps->dct};
    for (int k = 0; k < 4; k++)
        if (dst[k] && n) {
            if (!r->d_ftaps) { c->err = "batch was uploaded without float taps"; return L3B_E_PARAM; }
            CU_TRY(c, cudaMemcpyAsync(dst[k], r->d_ftaps + (uint64_t)k * r->cap_grch * 576, n * 576 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        }
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return 0;
}

void* l3b_batch_device_pcm(l3b_resident_t* r) { return r ? r->d_pcm : nullptr; }

int l3b_decode_batch(l3b_ctx_t* c, const l3b_batch_t* b) {
    if (!c || !b) return L3B_E_PARAM;
    if (!b->pcm && b->pcm_floats) { c->err = "pcm destination missing"; return L3B_E_PARAM; }
    l3b_resident_t* r = nullptr;
    int rc = l3b_batch_upload(c, b, &r);
    if (rc) return rc;
    rc = l3b_batch_run(c, r);
    if (!rc) rc = l3b_batch_download(c, r, b->pcm, 0, b->pcm_floats);
    if (!rc && b->taps) rc = l3b_batch_download_taps(c, r, b->taps);
    if (b->status)
        for (uint32_t i = 0; i < b->n_streams; i++) b->status[i] = rc;
    l3b_batch_free(c, r);
    return rc;
}

// ---------------------------------------------------------------------------------------------------
// layer 2: the decode program of a set of scanned streams, assembled into caller-owned memory
int l3b_scans_assemble(l3b_scan_t* const* scans, uint32_t n, uint8_t* blob, uint64_t blob_cap, l3b_grch_desc_t* descs,
                       uint64_t desc_cap, l3b_stream_desc_t* streams, l3b_batch_t* batch) {
    if (!scans || !n || !batch) return L3B_E_PARAM;
    uint64_t n_desc = 0, n_blob = 0, pcm_total = 0;
    for (uint32_t i = 0; i < n; i++) {
        if (!scans[i]) return L3B_E_PARAM;
        n_desc += scans[i]->r.prog.descs.size();
        n_blob += ((scans[i]->r.prog.blob.size() + 15) & ~(size_t)15) + 16;
        pcm_total = ((pcm_total + 3) & ~(uint64_t)3) + scans[i]->r.pcm_count;
    }
    memset(batch, 0, sizeof *batch);
    batch->maindata_bytes = n_blob;
    batch->n_grch = n_desc;
    batch->n_streams = n;
    batch->pcm_floats = pcm_total;
    if (!blob && !descs && !streams) return 0;   // size query
    if (!blob || !descs || !streams || blob_cap < n_blob || desc_cap < n_desc) return L3B_E_PARAM;
    uint64_t boff = 0, doff = 0, poff = 0;
    for (uint32_t i = 0; i < n; i++) {
        const ScanResult& s = scans[i]->r;
        l3b_scan_fill_stream_desc(scans[i], &streams[i]);
        streams[i].maindata_off = boff;
        streams[i].first_grch = doff;
        poff = (poff + 3) & ~(uint64_t)3;   // 16-byte aligned PCM rows (stereo stores are 8-byte vectors)
        streams[i].pcm_off = poff;
        poff += s.pcm_count;
        const size_t nb = s.prog.blob.size(), padded = ((nb + 15) & ~(size_t)15) + 16;
        if (nb) memcpy(blob + boff, s.prog.blob.data(), nb);
        memset(blob + boff + nb, 0, padded - nb);   // >= 16 zero bytes after every stream
        boff += padded;
        if (!s.prog.descs.empty()) memcpy(descs + doff, s.prog.descs.data(), s.prog.descs.size() * sizeof(l3b_grch_desc_t));
        doff += s.prog.descs.size();
    }
    batch->maindata = blob;
    batch->grch = descs;
    batch->streams = streams;
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// layer 2: the batch entry point over scanned streams
int l3b_decode_scans(l3b_ctx_t* c, l3b_scan_t* const* scans, uint32_t n, float* const* pcm, int32_t* status) {
    if (!c || !scans || !n || !pcm) return L3B_E_PARAM;
    std::vector<l3b_stream_desc_t> sd(n);
    std::vector<l3b_grch_desc_t> descs;
    std::vector<uint8_t> blob;
    uint64_t pcm_total = 0;
    size_t n_desc = 0, n_blob = 0;
    for (uint32_t i = 0; i < n; i++) {
        if (!scans[i]) return L3B_E_PARAM;
        n_desc += scans[i]->r.prog.descs.size();
        n_blob += ((scans[i]->r.prog.blob.size() + 15) & ~(size_t)15) + 16;
    }
    descs.reserve(n_desc);
    blob.reserve(n_blob);
    for (uint32_t i = 0; i < n; i++) {
        const ScanResult& s = scans[i]->r;
        l3b_scan_fill_stream_desc(scans[i], &sd[i]);
        sd[i].maindata_off = blob.size();
        sd[i].first_grch = descs.size();
        pcm_total = (pcm_total + 3) & ~(uint64_t)3;  // 16-byte aligned PCM rows (stereo stores are 8-byte vectors)
        sd[i].pcm_off = pcm_total;
        pcm_total += s.pcm_count;
        blob.insert(blob.end(), s.prog.blob.begin(), s.prog.blob.end());
        blob.resize(((blob.size() + 15) & ~(size_t)15) + 16, 0);  // >= 16 zero bytes after every stream
        descs.insert(descs.end(), s.prog.descs.begin(), s.prog.descs.end());
    }
    l3b_batch_t b{};
    b.maindata = blob.data();
    b.maindata_bytes = blob.size();
    b.grch = descs.data();
    b.n_grch = descs.size();
    b.streams = sd.data();
    b.n_streams = n;
    b.pcm = nullptr;
    b.pcm_floats = pcm_total;
    l3b_resident_t* r = nullptr;
    int rc = l3b_batch_upload(c, &b, &r);
    if (!rc) rc = l3b_batch_run(c, r);
    for (uint32_t i = 0; i < n && !rc; i++)
        if (sd[i].pcm_count) {
            cudaError_t e = cudaMemcpyAsync(pcm[i], r->d_pcm + sd[i].pcm_off, sd[i].pcm_count * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
            if (e != cudaSuccess) { c->err = cudaGetErrorString(e); rc = L3B_E_NOGPU; }
        }
    if (!rc) rc = l3b_batch_sync(c);
    if (status)
        for (uint32_t i = 0; i < n; i++) status[i] = rc ? rc : scans[i]->r.last_error;
    if (r) l3b_batch_free(c, r);
    return rc;
}

}  // extern "C"

// internal entry points for l3_raw.cu (the device prepass fills the blob and the descriptor table itself)
namespace l3b {
int resident_for_device_inputs(l3b_ctx_t* c, const l3b_batch_t* shape, l3b_resident_t** inout) { return upload_impl(c, shape, inout, false, true); }
uint8_t* resident_blob(l3b_resident_t* r) { return r->d_blob; }
l3b_resident_t* ctx_take_spare_workspace(l3b_ctx_t* c) {
    std::lock_guard<std::mutex> g(c->spare_mutex);
    l3b_resident_t* r = c->spare;
    c->spare = nullptr;
    return r;
}
void ctx_give_spare_workspace(l3b_ctx_t* c, l3b_resident_t* r) {
    if (!r) return;
    {
        std::lock_guard<std::mutex> g(c->spare_mutex);
        if (!c->spare) { c->spare = r; return; }
    }
    l3b_batch_free(c, r);
}
l3b_grch_desc_t* resident_descs(l3b_resident_t* r) { return r->d_grch; }
const uint8_t* ctx_sfb_width(l3b_ctx_t* c) { return c->t.sfb_width; }
cudaStream_t ctx_stream(l3b_ctx_t* c) { return c->stream; }
std::string& ctx_err(l3b_ctx_t* c) { return c->err; }
int ctx_device(l3b_ctx_t* c) { return c->device; }
}  // namespace l3b

