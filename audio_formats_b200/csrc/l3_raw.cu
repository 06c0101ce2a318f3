// l3_raw.cu -- C-ABI layer 2: batch entry point over RAW Layer III files with the prepass ON THE GPU.
//
// What the host prepass (l3_host.cpp) does per stream -- frame sync, side-info parse, bit-reservoir slicing -- re-expressed for
// the device for the streams it is simple for, which is what a corpus consists of: a clean chain of compatible Layer III
// frames from the end of the ID3v2 tag to the end of the data (ID3v1 / APE trimmed), not free-format, no Xing / Info tag,
// every side info legal.  For such a stream the reference's control flow (minimp3.d:1492-1556, minimp3_ex.d:787-888)
// degenerates to "decode every frame in order; a frame decodes iff the reservoir holds main_data_begin bytes", and
//   k_walk      one THREAD per stream hops from header to header (the only inherently sequential part: 4 bytes per frame)
//   k_sideinfo  one thread per frame parses the side info with the SAME code the host uses (parse_side_info_t, l3_format.hpp)
//   k_reservoir one thread per stream runs the reservoir recurrence (minimp3.d:1170-1194 as counts) and the running sums that
//               place every frame's main data in the linear blob and every granule-channel's descriptor in the table
//   k_gather    one CTA per frame copies the frame's main data into the blob and writes its descriptors (bit_start = bytes of
//               main data before the frame, minus main_data_begin, plus the part2_3 lengths before the granule-channel)
// Anything else (damage, resync, tags, free format, Layer I / II, channel or rate changes, illegal side info) makes the walker
// or the parser raise the stream's `needs_host` flag and the stream takes the host prepass instead: the result is the
// host route's by construction, the device route is the fast path for well-formed input.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/l3b200.h"
#include "l3_format.hpp"
#include "l3_host.hpp"
#include "l3_kernels.cuh"

using namespace l3b;

// internal entry points of l3_ctx.cu
namespace l3b {
int resident_for_device_inputs(l3b_ctx_t* c, const l3b_batch_t* shape, l3b_resident_t** inout);
uint8_t* resident_blob(l3b_resident_t* r);
l3b_grch_desc_t* resident_descs(l3b_resident_t* r);
const uint8_t* ctx_sfb_width(l3b_ctx_t* c);
cudaStream_t ctx_stream(l3b_ctx_t* c);
std::string& ctx_err(l3b_ctx_t* c);
int ctx_device(l3b_ctx_t* c);
}  // namespace l3b

namespace {

struct StreamIn {            // host -> device, per stream
    uint64_t raw_off;        // of the file's first byte in the raw buffer
    uint32_t begin, end;     // [begin, end): after the ID3v2 tag, before ID3v1 / APE
    uint32_t frame_base;     // first slot of this stream in the frame table
    uint32_t frame_cap;
};
struct StreamOut {           // device -> host
    uint32_t n_frames, needs_host, nch, mpeg1, sr_row, hz, n_granules, blob_bytes, n_desc, pad;
};
struct FrameRec {
    uint32_t off;            // of the header, relative to the file
    uint32_t stream;
    uint32_t sum_bits;       // part2_3 bits of the whole frame
    uint16_t mdb, payload;   // main_data_begin; bytes of main data the frame carries
    uint8_t body_off;        // header + CRC + side info
    uint8_t ok, nch, ngr;
    uint32_t blob_off;       // bytes of main data of this stream before this frame
    int32_t desc_off;        // first descriptor of the frame within its stream, -1: not decodable (reservoir underrun)
};
struct DescTmp { uint32_t w1, w2, w3; };

__constant__ uint8_t c_raw_halfrate[2 * 3 * 15];

struct DevSfbRows {
    const uint8_t* w;   // DeviceTables::sfb_width: [8][3][40]
    __host__ __device__ const uint8_t* operator()(int row, int kind) const { return w + (row * 3 + kind) * 40; }
};

__global__ void k_walk(const uint8_t* raw, const StreamIn* in, StreamOut* out, FrameRec* frames, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const StreamIn S = in[i];
    const uint8_t* f = raw + S.raw_off;
    StreamOut o = {};
    uint32_t pos = S.begin, k = 0;
    uint8_t h0[4] = {0, 0, 0, 0};
    bool bad = S.end < S.begin + 4;
    if (!bad) {
        for (int b = 0; b < 4; b++) h0[b] = f[pos + b];
        const Hdr H(h0);
        bad = !H.valid() || H.layer() != 3 || H.free_format();
        if (!bad) {
            o.nch = (uint32_t)H.channels(); o.mpeg1 = H.mpeg1() ? 1u : 0u; o.sr_row = (uint32_t)H.sfb_row(); o.hz = H.sample_rate_hz();
        }
    }
    while (!bad && pos + 4 <= S.end) {
        uint8_t h[4];
        for (int b = 0; b < 4; b++) h[b] = f[pos + b];
        const Hdr H(h);
        if (!hdr_compatible(h0, h) || H.mono() != Hdr(h0).mono() || H.free_format()) { bad = true; break; }
        const uint32_t fs = (uint32_t)(H.frame_bytes_t(c_raw_halfrate, 0) + H.padding());
        if (fs < 4 || pos + fs > S.end || k >= S.frame_cap) { bad = true; break; }
        FrameRec r = {};
        r.off = pos; r.stream = i;
        frames[S.frame_base + k] = r;
        k++;
        pos += fs;
    }
    if (pos != S.end || k < 2) bad = true;   // trailing bytes, a cut frame, or too short for the open's sync rule: the host decides
    o.n_frames = k;
    o.needs_host = bad ? 1u : 0u;
    out[i] = o;
}

__global__ void k_sideinfo(const uint8_t* raw, const StreamIn* in, StreamOut* out, FrameRec* frames, DescTmp* dtmp, uint32_t n_slots,
                           const uint8_t* sfb_width) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slots) return;
    FrameRec r = frames[s];
    // slots past a stream's frame count were never written by the walker: recognise them through the stream's count
    // (every slot of a stream lies in [frame_base, frame_base + frame_cap); the walker wrote `stream` into the used ones only,
    // so find the owner by the slot index instead)
    // -> the launch passes one thread per USED slot through the compacted index below
    const StreamIn S = in[r.stream];
    if (s < S.frame_base || s >= S.frame_base + out[r.stream].n_frames || out[r.stream].needs_host) return;
    const uint8_t* hdr = raw + S.raw_off + r.off;
    uint8_t h[4];
    for (int b = 0; b < 4; b++) h[b] = hdr[b];
    const Hdr H(h);
    const int fs = H.frame_bytes_t(c_raw_halfrate, 0) + H.padding();
    BitReader bs(hdr + kHdrSize, fs - kHdrSize);
    if (H.has_crc()) bs.get(16);
    GranuleInfo gr[4];
    const int mdb = parse_side_info_t(bs, gr, h, DevSfbRows{sfb_width});
    if (mdb < 0 || bs.pos > bs.limit) {   // the reference drops the frame and re-initialises the decoder: host route
        atomicExch(&out[r.stream].needs_host, 1u);
        return;
    }
    // a Xing / Info tag in the first frame changes the stream's length and delay bookkeeping (minimp3_ex.d:144-190): host route
    if (s == S.frame_base) {
        const uint8_t* tag = hdr + kHdrSize + bs.pos / 8;
        if (r.off + kHdrSize + bs.pos / 8 + 4 <= S.end &&
            ((tag[0] == 'X' && tag[1] == 'i' && tag[2] == 'n' && tag[3] == 'g') || (tag[0] == 'I' && tag[1] == 'n' && tag[2] == 'f' && tag[3] == 'o'))) {
            atomicExch(&out[r.stream].needs_host, 1u);
            return;
        }
    }
    const int nch = H.channels(), ngr = H.mpeg1() ? 2 : 1;
    uint32_t sum = 0;
    for (int g = 0; g < ngr; g++)
        for (int ch = 0; ch < nch; ch++) {
            const GranuleInfo& q = gr[g * nch + ch];
            const l3b_grch_desc_t d = pack_desc(q, 0u, h[3], g == 1, false);
            dtmp[(size_t)s * 4 + g * nch + ch] = DescTmp{d.w1, d.w2, d.w3};
            sum += q.part_23_length;
        }
    r.sum_bits = sum;
    r.mdb = (uint16_t)mdb;
    r.payload = (uint16_t)((bs.limit - bs.pos) / 8);
    r.body_off = (uint8_t)(kHdrSize + bs.pos / 8);
    r.ok = 1; r.nch = (uint8_t)nch; r.ngr = (uint8_t)ngr;
    frames[s] = r;
}

// L3_restore_reservoir / L3_save_reservoir as counts (l3_host.cpp, FrameWalker::step), plus the running sums
__global__ void k_reservoir(const StreamIn* in, StreamOut* out, FrameRec* frames, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    StreamOut o = out[i];
    if (o.needs_host) return;
    const StreamIn S = in[i];
    int reserv = 0;
    uint32_t blob = 0, ndesc = 0, ngran = 0;
    for (uint32_t k = 0; k < o.n_frames; k++) {
        FrameRec& r = frames[S.frame_base + k];
        if (!r.ok) { o.needs_host = 1; break; }
        const int mdb = r.mdb, payload = r.payload;
        r.blob_off = blob;
        if (reserv >= mdb) {
            r.desc_off = (int32_t)ndesc;
            ndesc += (uint32_t)r.nch * r.ngr;
            ngran += r.ngr;
            const int remains = (mdb + payload) - (int)((r.sum_bits + 7) / 8);
            reserv = max(0, min(remains, kMaxReservoir));
        } else {
            r.desc_off = -1;
            reserv = min(reserv + payload, kMaxReservoir);
        }
        blob += (uint32_t)payload;
        if (blob > 0x1FFFFFF0u) { o.needs_host = 1; break; }   // 32-bit bit offsets
    }
    o.blob_bytes = blob; o.n_desc = ndesc; o.n_granules = ngran;
    out[i] = o;
}

__global__ void __launch_bounds__(128) k_gather(const uint8_t* raw, const StreamIn* in, const StreamOut* out, const FrameRec* frames,
                                                const DescTmp* dtmp, const uint32_t* used_slot, uint32_t n_used, const uint64_t* blob_base,
                                                const uint64_t* desc_base, uint8_t* blob, l3b_grch_desc_t* descs) {
    const uint32_t u = blockIdx.x;
    if (u >= n_used) return;
    const uint32_t s = used_slot[u];
    const FrameRec r = frames[s];
    const StreamIn S = in[r.stream];
    const uint8_t* src = raw + S.raw_off + r.off + r.body_off;
    uint8_t* dst = blob + blob_base[r.stream] + r.blob_off;
    for (uint32_t b = threadIdx.x; b < r.payload; b += blockDim.x) dst[b] = src[b];
    if (r.desc_off >= 0 && threadIdx.x < (uint32_t)r.nch * r.ngr) {
        const int j = (int)threadIdx.x;
        uint32_t before = 0;                                   // part2_3 bits of the granule-channels before this one
        for (int q = 0; q < j; q++) before += dtmp[(size_t)s * 4 + q].w1 & 0xFFFu;
        const DescTmp t = dtmp[(size_t)s * 4 + j];
        l3b_grch_desc_t d;
        d.bit_start = (r.blob_off - r.mdb) * 8u + before;
        d.w1 = t.w1; d.w2 = t.w2;
        d.w3 = t.w3 | ((r.desc_off == 0 && j < r.nch) ? 0x80000000u : 0u);   // decoder state starts zeroed: the first granule
        descs[desc_base[r.stream] + (uint32_t)r.desc_off + (uint32_t)j] = d;
    }
}

}  // namespace

struct l3b_raw {
    l3b_ctx_t* ctx = nullptr;
    uint32_t n = 0, flags = 0;
    struct Info { int channels = 0, hz = 0, status = 0; uint64_t samples = 0; bool on_device = false; uint32_t dev_index = 0; };
    std::vector<Info> info;
    std::vector<l3b_scan_t*> scans;          // host-route streams (NULL for the others)
    l3b_resident_t* res = nullptr;           // device-route streams, one resident batch
    std::vector<l3b_stream_desc_t> dev_sd;   // its stream table
    l3b_resident_t* res_host = nullptr;      // host-route streams, assembled and uploaded at open
    std::vector<l3b_stream_desc_t> host_sd;
    std::vector<uint32_t> host_ids;
    float prepass_ms = 0;
    uint32_t n_device = 0;
};

extern "C" {

void l3b_raw_free(l3b_raw_t* r) {
    if (!r) return;
    if (r->res) l3b_batch_free(r->ctx, r->res);
    if (r->res_host) l3b_batch_free(r->ctx, r->res_host);
    for (auto* s : r->scans) l3b_scan_free(s);
    delete r;
}

#define RAW_TRY(expr)                                                                 \
    do {                                                                              \
        cudaError_t e_ = (expr);                                                      \
        if (e_ != cudaSuccess) {                                                      \
            ctx_err(c) = std::string(#expr) + ": " + cudaGetErrorString(e_);          \
            rc = e_ == cudaErrorMemoryAllocation ? L3B_E_MEMORY : L3B_E_NOGPU;        \
            goto done;                                                                \
        }                                                                             \
    } while (0)

int l3b_raw_open(l3b_ctx_t* c, const uint8_t* const* data, const size_t* size, uint32_t n, uint32_t flags, l3b_raw_t** out) {
    if (!c || !data || !size || !n || !out) return L3B_E_PARAM;
    *out = nullptr;
    l3b_raw* R = new (std::nothrow) l3b_raw();
    if (!R) return L3B_E_MEMORY;
    R->ctx = c; R->n = n; R->flags = flags;
    R->info.resize(n);
    R->scans.assign(n, nullptr);
    int rc = 0;
    uint8_t* d_raw = nullptr;
    StreamIn* d_in = nullptr;
    StreamOut* d_out = nullptr;
    FrameRec* d_frames = nullptr;
    DescTmp* d_dtmp = nullptr;
    uint32_t* d_used = nullptr;
    uint64_t *d_bbase = nullptr, *d_dbase = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaStream_t st = ctx_stream(c);
    std::vector<StreamIn> in(n);
    std::vector<StreamOut> so(n);
    std::vector<uint32_t> used;
    std::vector<uint64_t> bbase(n, 0), dbase(n, 0);
    uint64_t raw_bytes = 0;
    uint32_t n_slots = 0;
    // ---- host: tag skipping only (O(1) per stream: minimp3_ex.d:93-125) ----
    for (uint32_t i = 0; i < n; i++) {
        size_t id3 = 0, end = size[i];
        if (data[i] && size[i] >= (size_t)kId3DetectSize) skip_id3v2(data[i], size[i], &id3);
        if (id3 > size[i]) id3 = size[i];
        if (data[i]) skip_id3v1(data[i], &end);
        // an APE tag is measured against the reference's 128 KiB read window, not against the file (minimp3_ex.d:101-110): host route
        bool ape = false;
        {
            size_t e1 = size[i];
            if (data[i] && e1 >= 128 && !memcmp(data[i] + e1 - 128, "TAG", 3)) {
                e1 -= 128;
                if (e1 >= 227 && !memcmp(data[i] + e1 - 227, "TAG+", 4)) e1 -= 227;
            }
            ape = data[i] && e1 > 32 && !memcmp(data[i] + e1 - 32, "APETAGEX", 8);
        }
        in[i].raw_off = raw_bytes;
        in[i].begin = (uint32_t)std::min<size_t>(id3, 0xFFFFFFFFu);
        in[i].end = (uint32_t)std::min<size_t>(std::max(end, id3), 0xFFFFFFFFu);
        in[i].frame_base = n_slots;
        in[i].frame_cap = (uint32_t)(size[i] / 48 + 2);     // streams with smaller frames (under 16 kbps) overflow it and take the host route
        if (size[i] > 0x7FFFFFF0u || !data[i] || ape) in[i].end = in[i].begin;   // host route decides
        n_slots += in[i].frame_cap;
        raw_bytes += (size[i] + 63) & ~(size_t)63;
    }
    RAW_TRY(cudaSetDevice(ctx_device(c)));
    RAW_TRY(cudaMalloc(&d_raw, raw_bytes + 64));
    RAW_TRY(cudaMalloc(&d_in, n * sizeof(StreamIn)));
    RAW_TRY(cudaMalloc(&d_out, n * sizeof(StreamOut)));
    RAW_TRY(cudaMalloc(&d_frames, (size_t)n_slots * sizeof(FrameRec)));
    RAW_TRY(cudaMalloc(&d_dtmp, (size_t)n_slots * 4 * sizeof(DescTmp)));
    RAW_TRY(cudaEventCreate(&ev0));
    RAW_TRY(cudaEventCreate(&ev1));
    RAW_TRY(cudaMemcpyToSymbolAsync(c_raw_halfrate, L3_HALFRATE, sizeof c_raw_halfrate, 0, cudaMemcpyHostToDevice, st));
    for (uint32_t i = 0; i < n; i++)
        if (size[i] && data[i]) RAW_TRY(cudaMemcpyAsync(d_raw + in[i].raw_off, data[i], size[i], cudaMemcpyHostToDevice, st));
    RAW_TRY(cudaMemcpyAsync(d_in, in.data(), n * sizeof(StreamIn), cudaMemcpyHostToDevice, st));
    RAW_TRY(cudaMemsetAsync(d_frames, 0, (size_t)n_slots * sizeof(FrameRec), st));
    RAW_TRY(cudaEventRecord(ev0, st));
    k_walk<<<(n + 31) / 32, 32, 0, st>>>(d_raw, d_in, d_out, d_frames, n);
    // the side-info pass needs to know which slots are in use: the walker wrote `stream` only into those; unused slots are
    // zeroed (stream 0), and a zeroed slot inside stream 0's own range is told apart by the stream's frame count
    k_sideinfo<<<(n_slots + 127) / 128, 128, 0, st>>>(d_raw, d_in, d_out, d_frames, d_dtmp, n_slots, ctx_sfb_width(c));
    k_reservoir<<<(n + 31) / 32, 32, 0, st>>>(d_in, d_out, d_frames, n);
    RAW_TRY(cudaEventRecord(ev1, st));
    RAW_TRY(cudaMemcpyAsync(so.data(), d_out, n * sizeof(StreamOut), cudaMemcpyDeviceToHost, st));
    RAW_TRY(cudaStreamSynchronize(st));
    RAW_TRY(cudaGetLastError());
    RAW_TRY(cudaEventElapsedTime(&R->prepass_ms, ev0, ev1));
    {
        // ---- host: lay the device-route streams out in one batch; the others take the host prepass ----
        uint64_t boff = 0, doff = 0, poff = 0;
        for (uint32_t i = 0; i < n; i++) {
            l3b_raw::Info& I = R->info[i];
            if (so[i].needs_host || !so[i].n_granules) {
                l3b_scan_t* sc = nullptr;
                const int src = data[i] ? l3b_scan_memory(data[i], size[i], &sc) : L3B_E_PARAM;
                I.status = src;
                if (!src) {
                    R->scans[i] = sc;
                    I.channels = l3b_scan_channels(sc); I.hz = l3b_scan_samplerate(sc);
                    I.samples = l3b_scan_delivered_samples(sc);
                    I.status = l3b_scan_error(sc);
                    R->host_ids.push_back(i);
                }
                continue;
            }
            I.on_device = true;
            I.dev_index = (uint32_t)R->dev_sd.size();
            I.channels = (int)so[i].nch; I.hz = (int)so[i].hz;
            I.samples = (uint64_t)so[i].n_granules * 576u * so[i].nch;
            l3b_stream_desc_t sd{};
            sd.maindata_off = boff; sd.maindata_bytes = so[i].blob_bytes; sd.n_granules = so[i].n_granules;
            sd.first_grch = doff;
            poff = (poff + 3) & ~(uint64_t)3;
            sd.pcm_off = poff; sd.pcm_skip = 0; sd.pcm_count = I.samples;
            sd.nch = (uint8_t)so[i].nch; sd.sr_idx = (uint8_t)so[i].sr_row; sd.mpeg1 = (uint8_t)so[i].mpeg1;
            bbase[i] = boff; dbase[i] = doff;
            boff += (((uint64_t)so[i].blob_bytes + 15) & ~(uint64_t)15) + 16;
            doff += so[i].n_desc;
            poff += I.samples;
            R->dev_sd.push_back(sd);
            for (uint32_t k = 0; k < so[i].n_frames; k++) used.push_back(in[i].frame_base + k);
        }
        R->n_device = (uint32_t)R->dev_sd.size();
        if (R->n_device) {
            l3b_batch_t shape{};
            shape.maindata_bytes = boff; shape.n_grch = doff;
            shape.streams = R->dev_sd.data(); shape.n_streams = R->n_device;
            shape.pcm_floats = poff; shape.flags = flags;
            rc = resident_for_device_inputs(c, &shape, &R->res);
            if (rc) goto done;
            RAW_TRY(cudaMalloc(&d_used, used.size() * sizeof(uint32_t)));
            RAW_TRY(cudaMalloc(&d_bbase, n * sizeof(uint64_t)));
            RAW_TRY(cudaMalloc(&d_dbase, n * sizeof(uint64_t)));
            RAW_TRY(cudaMemcpyAsync(d_used, used.data(), used.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
            RAW_TRY(cudaMemcpyAsync(d_bbase, bbase.data(), n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
            RAW_TRY(cudaMemcpyAsync(d_dbase, dbase.data(), n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
            RAW_TRY(cudaMemsetAsync(resident_blob(R->res), 0, boff + 64, st));   // the zero bytes after every stream
            k_gather<<<(unsigned)used.size(), 128, 0, st>>>(d_raw, d_in, d_out, d_frames, d_dtmp, d_used, (uint32_t)used.size(), d_bbase, d_dbase,
                                                            resident_blob(R->res), resident_descs(R->res));
            RAW_TRY(cudaStreamSynchronize(st));
            RAW_TRY(cudaGetLastError());
        }
        if (!R->host_ids.empty()) {   // the host route: assemble + upload like l3b_decode_scans
            std::vector<l3b_scan_t*> hs;
            for (uint32_t i : R->host_ids) hs.push_back(R->scans[i]);
            l3b_batch_t b{};
            rc = l3b_scans_assemble(hs.data(), (uint32_t)hs.size(), nullptr, 0, nullptr, 0, nullptr, &b);
            if (rc) goto done;
            std::vector<uint8_t> blob(b.maindata_bytes + 64);
            std::vector<l3b_grch_desc_t> descs(b.n_grch + 1);
            R->host_sd.resize(hs.size());
            rc = l3b_scans_assemble(hs.data(), (uint32_t)hs.size(), blob.data(), blob.size(), descs.data(), descs.size(), R->host_sd.data(), &b);
            if (rc) goto done;
            b.flags = flags;
            rc = l3b_batch_upload(c, &b, &R->res_host);
            if (rc) goto done;
        }
    }
done:
    cudaFree(d_raw); cudaFree(d_in); cudaFree(d_out); cudaFree(d_frames); cudaFree(d_dtmp); cudaFree(d_used); cudaFree(d_bbase); cudaFree(d_dbase);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (rc) { l3b_raw_free(R); return rc; }
    *out = R;
    return 0;
}

uint32_t l3b_raw_device_streams(const l3b_raw_t* r) { return r ? r->n_device : 0; }
float l3b_raw_prepass_ms(const l3b_raw_t* r) { return r ? r->prepass_ms : 0.0f; }
int l3b_raw_channels(const l3b_raw_t* r, uint32_t i) { return (r && i < r->n) ? r->info[i].channels : 0; }
int l3b_raw_samplerate(const l3b_raw_t* r, uint32_t i) { return (r && i < r->n) ? r->info[i].hz : 0; }
uint64_t l3b_raw_samples(const l3b_raw_t* r, uint32_t i) { return (r && i < r->n) ? r->info[i].samples : 0; }
int l3b_raw_status(const l3b_raw_t* r, uint32_t i) { return (r && i < r->n) ? r->info[i].status : L3B_E_PARAM; }

int l3b_raw_decode(l3b_raw_t* R, void* const* pcm) {
    if (!R || !pcm) return L3B_E_PARAM;
    l3b_ctx_t* c = R->ctx;
    int rc = 0;
    if (R->res) rc = l3b_batch_run(c, R->res);
    if (!rc && R->res_host) rc = l3b_batch_run(c, R->res_host);
    for (uint32_t i = 0; i < R->n && !rc; i++) {
        const l3b_raw::Info& I = R->info[i];
        if (!I.samples || !pcm[i]) continue;
        if (I.on_device) {
            rc = l3b_batch_download(c, R->res, pcm[i], R->dev_sd[I.dev_index].pcm_off, I.samples);
        } else {
            const size_t k = (size_t)(std::find(R->host_ids.begin(), R->host_ids.end(), i) - R->host_ids.begin());
            rc = l3b_batch_download(c, R->res_host, pcm[i], R->host_sd[k].pcm_off, R->host_sd[k].pcm_count);
        }
    }
    return rc;
}

}  // extern "C"
