// l3_pipeline.cpp -- C-ABI layer 2: the batch entry point over RAW MP3 streams, pipelined in waves across one or more GPUs.
//
// This is what the D host's batch entry point calls: MP3 bytes in host memory in, PCM in (pinned) host
// memory out.  The batch is cut into waves; per GPU a few LANE threads (context + CUDA stream + recycled device workspace
// + pinned staging each) run  assemble -> H2D -> entropy + granule kernels -> D2H  for their waves, and a pool of SCAN
// threads runs the host prepass (frame sync, side info, reservoir slicing: l3_host.cpp) ahead of them.  While one lane
// waits for its PCM copy the others assemble and launch, so the copy engines, the SMs and the host cores overlap.
// Lanes wait on blocking events (they sleep), so the host cores belong to the scan threads.
//
// Streams are independent (minimp3.d:38-46: all decoder state is per stream), so with several GPUs they are assigned by
// file, longest-processing-time first on their byte size, with no data-path collective.  One stream that
// cannot be decoded does not affect the others: it gets its status and zero frames.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <system_error>
#include <thread>
#include <vector>

#include <sched.h>

#include "../../include/l3b200.h"

namespace {

using Clock = std::chrono::steady_clock;
inline double secs(Clock::time_point a, Clock::time_point b) { return std::chrono::duration<double>(b - a).count(); }

struct Lane {
    l3b_ctx_t* ctx = nullptr;
    l3b_resident_t* ws = nullptr;       // device buffers recycled from wave to wave
    uint8_t* staging = nullptr;         // pinned: blob | descriptors | stream table of the wave being uploaded
    size_t staging_bytes = 0;
};

struct Device {
    int id = 0;
    std::vector<Lane> lanes;
};

struct Wave {
    int dev = 0;
    std::vector<uint32_t> streams;      // indices into the batch
    uint32_t scanned = 0;               // streams whose prepass has finished (guarded by the pipeline mutex)
};

}  // namespace

struct l3b_pipeline {
    std::vector<Device> devs;
    l3b_pipeline_opts_t opts{};
    std::string err;
    double prof[L3B_PIPELINE_PHASES] = {};   // accumulated since the last l3b_pipeline_profile
    std::mutex prof_mu;
};

namespace {

int host_cpus() {
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0) return std::max(1, CPU_COUNT(&set));
    return (int)std::max(1u, std::thread::hardware_concurrency());
}

// Longest-processing-time-first assignment of streams to devices by cost; ties go to the lower device / lower index, so
// the result is deterministic.  (With one device everything goes to it.)
std::vector<std::vector<uint32_t>> assign_lpt(const size_t* cost, uint32_t n, int n_dev) {
    std::vector<std::vector<uint32_t>> out((size_t)n_dev);
    if (n_dev == 1) {
        out[0].resize(n);
        for (uint32_t i = 0; i < n; i++) out[0][i] = i;
        return out;
    }
    std::vector<uint32_t> order(n);
    for (uint32_t i = 0; i < n; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return cost[a] > cost[b]; });
    std::vector<uint64_t> load((size_t)n_dev, 0);
    for (uint32_t i : order) {
        int best = 0;
        for (int d = 1; d < n_dev; d++)
            if (load[(size_t)d] < load[(size_t)best]) best = d;
        out[(size_t)best].push_back(i);
        load[(size_t)best] += cost[i];
    }
    for (auto& v : out) std::sort(v.begin(), v.end());
    return out;
}

struct Run {
    l3b_pipeline* P;
    const uint8_t* const* data;
    const size_t* size;
    uint32_t n;
    uint8_t* out;
    uint64_t out_cap;      // elements
    size_t elem;           // bytes per element
    l3b_stream_result_t* res;

    std::vector<Wave> waves;                  // global scan order: round-robin over the devices
    std::vector<std::vector<uint32_t>> dev_waves;   // per device: indices into `waves`, in order
    std::vector<l3b_scan_t*> scans;           // per stream, owned until its wave is done
    std::vector<uint32_t> wave_of, pos_in_scan_order;

    std::mutex mu;
    std::condition_variable cv;
    uint64_t next_item = 0;                   // next (wave, stream) pair to scan, in global order
    std::vector<uint64_t> item_wave_first;    // first item index of each wave
    std::vector<uint32_t> dev_next_wave;      // per device: next wave (position in dev_waves) a lane takes
    std::vector<uint32_t> dev_done_waves;     // per device: waves fully processed (back-pressure for the scan threads)
    uint64_t cursor = 0;                      // output elements handed out
    int first_error = 0;
    std::string err;
    double prof[L3B_PIPELINE_PHASES] = {};

    void fail(int rc, const std::string& what) {
        std::lock_guard<std::mutex> g(mu);
        if (!first_error) { first_error = rc; err = what; }
        cv.notify_all();   // scan threads waiting for the lanes to catch up give up too
    }
};

void scan_worker(Run* R) {
    const uint64_t total = R->item_wave_first.back();
    for (;;) {
        uint64_t item;
        uint32_t w;
        {
            std::unique_lock<std::mutex> lk(R->mu);
            if (R->next_item >= total || R->first_error) return;
            item = R->next_item;
            // the wave this item belongs to
            w = (uint32_t)(std::upper_bound(R->item_wave_first.begin(), R->item_wave_first.end(), item) - R->item_wave_first.begin() - 1);
            const int d = R->waves[w].dev;
            // position of wave w among its device's waves
            const uint32_t pos = R->pos_in_scan_order[w];
            const uint32_t window = (uint32_t)R->P->devs[(size_t)d].lanes.size() + 2;
            if (pos >= R->dev_done_waves[(size_t)d] + window) {   // far enough ahead of the lanes: wait (bounds the memory held in scans)
                R->cv.wait(lk);
                continue;
            }
            R->next_item++;
        }
        const Wave& W = R->waves[w];
        const uint32_t si = W.streams[(size_t)(item - R->item_wave_first[w])];
        const auto t0 = Clock::now();
        l3b_scan_t* sc = nullptr;
        const int rc = l3b_scan_memory(R->data[si], R->size[si], &sc);
        const auto t1 = Clock::now();
        {
            std::lock_guard<std::mutex> g(R->mu);
            R->scans[si] = rc ? nullptr : sc;
            R->res[si].status = rc;
            R->prof[0] += secs(t0, t1);
            if (++R->waves[w].scanned == W.streams.size()) R->cv.notify_all();
        }
    }
}

void lane_worker(Run* R, int d, int k) {
    Device& D = R->P->devs[(size_t)d];
    Lane& L = D.lanes[(size_t)k];
    const uint32_t flags = R->P->opts.flags;
    for (;;) {
        uint32_t w;
        const auto t_wait0 = Clock::now();
        {
            std::unique_lock<std::mutex> lk(R->mu);
            const uint32_t pos = R->dev_next_wave[(size_t)d];
            if (pos >= R->dev_waves[(size_t)d].size() || R->first_error) return;
            R->dev_next_wave[(size_t)d]++;
            w = R->dev_waves[(size_t)d][pos];
            R->cv.wait(lk, [&] { return R->waves[w].scanned == R->waves[w].streams.size() || R->first_error; });
            if (R->waves[w].scanned != R->waves[w].streams.size()) return;   // abandoned after an error elsewhere
        }
        const auto t0 = Clock::now();
        Wave& W = R->waves[w];
        std::vector<l3b_scan_t*> good;
        std::vector<uint32_t> good_idx;
        for (uint32_t si : W.streams)
            if (R->scans[si]) { good.push_back(R->scans[si]); good_idx.push_back(si); }
        double t_asm = 0, t_up = 0, t_run = 0, t_down = 0;
        if (!good.empty()) {
            // ---- assemble the wave's decode program into this lane's pinned staging ----
            l3b_batch_t b;
            int rc = l3b_scans_assemble(good.data(), (uint32_t)good.size(), nullptr, 0, nullptr, 0, nullptr, &b);
            const size_t blob_bytes = (size_t)b.maindata_bytes, desc_off = (blob_bytes + 63) & ~(size_t)63;
            const size_t sd_off = (desc_off + (size_t)b.n_grch * sizeof(l3b_grch_desc_t) + 63) & ~(size_t)63;
            const size_t need = sd_off + good.size() * sizeof(l3b_stream_desc_t) + 64;
            if (!rc && need > L.staging_bytes) {
                // nothing queued on this lane still reads the old staging: every wave ends with a waited download
                l3b_host_free(L.staging);
                L.staging_bytes = need + need / 4;
                L.staging = static_cast<uint8_t*>(l3b_host_alloc_near(D.id, L.staging_bytes));
                if (!L.staging) { L.staging_bytes = 0; rc = L3B_E_MEMORY; }
            }
            l3b_stream_desc_t* sd = nullptr;
            if (!rc) {
                sd = reinterpret_cast<l3b_stream_desc_t*>(L.staging + sd_off);
                rc = l3b_scans_assemble(good.data(), (uint32_t)good.size(), L.staging, blob_bytes,
                                        reinterpret_cast<l3b_grch_desc_t*>(L.staging + desc_off), b.n_grch, sd, &b);
            }
            uint64_t base = 0;
            if (!rc) {
                std::lock_guard<std::mutex> g(R->mu);   // reserve the wave's output region (16-byte aligned)
                const uint64_t al = 16 / R->elem;
                base = (R->cursor + al - 1) / al * al;
                if (base + b.pcm_floats > R->out_cap) rc = L3B_E_PARAM;
                else R->cursor = base + b.pcm_floats;
            }
            const auto t1 = Clock::now();
            t_asm = secs(t0, t1);
            if (rc == L3B_E_PARAM && sd) {
                R->fail(rc, "output buffer too small for the decoded PCM");
            } else if (!rc) {
                b.flags = flags;
                rc = l3b_batch_upload_reuse(L.ctx, &b, &L.ws);
                const auto t2 = Clock::now();
                if (!rc) rc = l3b_batch_run(L.ctx, L.ws);
                const auto t3 = Clock::now();
                if (!rc) rc = l3b_batch_download(L.ctx, L.ws, R->out + base * R->elem, 0, b.pcm_floats);
                const auto t4 = Clock::now();
                t_up = secs(t1, t2); t_run = secs(t2, t3); t_down = secs(t3, t4);
                if (rc) R->fail(rc, std::string("GPU ") + std::to_string(D.id) + ": " + l3b_last_error(L.ctx));
            } else {
                R->fail(rc, "cannot assemble / stage a wave");
            }
            for (size_t j = 0; j < good.size(); j++) {
                l3b_stream_result_t& r = R->res[good_idx[j]];
                const int nch = l3b_scan_channels(good[j]);
                r.channels = nch;
                r.samplerate = l3b_scan_samplerate(good[j]);
                if (rc) { r.status = rc; continue; }
                r.status = l3b_scan_error(good[j]);   // a sticky decode error met mid-stream (what was decoded before it is delivered)
                r.pcm_off = base + sd[j].pcm_off;
                r.frames = nch ? sd[j].pcm_count / (uint64_t)nch : 0;
                r.device = d;
            }
        }
        for (uint32_t si : W.streams) { l3b_scan_free(R->scans[si]); R->scans[si] = nullptr; }
        {
            std::lock_guard<std::mutex> g(R->mu);
            R->dev_done_waves[(size_t)d]++;
            R->prof[1] += t_asm; R->prof[2] += t_up; R->prof[3] += t_run; R->prof[4] += t_down;
            R->prof[5] += secs(t_wait0, t0);
            R->cv.notify_all();
        }
    }
}

}  // namespace

extern "C" {

int l3b_pipeline_create(const int* device_ids, int n_devices, const l3b_pipeline_opts_t* opts, l3b_pipeline_t** out) {
    if (!out || n_devices < 1 || !device_ids) return L3B_E_PARAM;
    *out = nullptr;
    l3b_pipeline* P = new (std::nothrow) l3b_pipeline();
    if (!P) return L3B_E_MEMORY;
    if (opts) P->opts = *opts;
    if (P->opts.lanes <= 0) P->opts.lanes = 4;
    if (P->opts.wave_streams <= 0) P->opts.wave_streams = 16;
    if (P->opts.scan_threads <= 0) P->opts.scan_threads = std::max(1, host_cpus() - n_devices);
    P->devs.resize((size_t)n_devices);
    for (int d = 0; d < n_devices; d++) {
        P->devs[(size_t)d].id = device_ids[d];
        P->devs[(size_t)d].lanes.resize((size_t)P->opts.lanes);
        for (auto& L : P->devs[(size_t)d].lanes) {
            const int rc = l3b_ctx_create(device_ids[d], &L.ctx);
            if (rc) { l3b_pipeline_destroy(P); return rc; }
        }
    }
    *out = P;
    return 0;
}

void l3b_pipeline_destroy(l3b_pipeline_t* P) {
    if (!P) return;
    for (auto& D : P->devs)
        for (auto& L : D.lanes) {
            if (L.ws) l3b_batch_free(L.ctx, L.ws);
            if (L.staging) l3b_host_free(L.staging);
            if (L.ctx) l3b_ctx_destroy(L.ctx);
        }
    delete P;
}

const char* l3b_pipeline_last_error(const l3b_pipeline_t* P) { return P ? P->err.c_str() : ""; }

int l3b_pipeline_decode(l3b_pipeline_t* P, const uint8_t* const* data, const size_t* size, uint32_t n, void* out,
                        uint64_t out_capacity, l3b_stream_result_t* results, uint64_t* out_used) {
    if (!P || !data || !size || !n || !out || !results) return L3B_E_PARAM;
    Run R;
    R.P = P; R.data = data; R.size = size; R.n = n;
    R.out = static_cast<uint8_t*>(out);
    R.out_cap = out_capacity;
    R.elem = (P->opts.flags & L3B_OUT_S16) ? sizeof(int16_t) : sizeof(float);
    R.res = results;
    memset(results, 0, sizeof(*results) * n);
    const int n_dev = (int)P->devs.size();
    try {
        // ---- streams -> devices (LPT on bytes) -> waves; global scan order interleaves the devices ----
        const auto per_dev = assign_lpt(size, n, n_dev);
        const uint32_t ws = (uint32_t)P->opts.wave_streams;
        std::vector<std::vector<Wave>> tmp((size_t)n_dev);
        size_t max_waves = 0;
        for (int d = 0; d < n_dev; d++) {
            const auto& v = per_dev[(size_t)d];
            for (size_t i = 0; i < v.size(); i += ws) {
                Wave w;
                w.dev = d;
                w.streams.assign(v.begin() + (long)i, v.begin() + (long)std::min(v.size(), i + ws));
                tmp[(size_t)d].push_back(std::move(w));
            }
            max_waves = std::max(max_waves, tmp[(size_t)d].size());
        }
        R.dev_waves.resize((size_t)n_dev);
        for (size_t k = 0; k < max_waves; k++)
            for (int d = 0; d < n_dev; d++)
                if (k < tmp[(size_t)d].size()) {
                    R.dev_waves[(size_t)d].push_back((uint32_t)R.waves.size());
                    R.pos_in_scan_order.push_back((uint32_t)k);
                    R.waves.push_back(std::move(tmp[(size_t)d][k]));
                }
        R.item_wave_first.resize(R.waves.size() + 1, 0);
        for (size_t w = 0; w < R.waves.size(); w++) R.item_wave_first[w + 1] = R.item_wave_first[w] + R.waves[w].streams.size();
        R.scans.assign(n, nullptr);
        R.dev_next_wave.assign((size_t)n_dev, 0);
        R.dev_done_waves.assign((size_t)n_dev, 0);

        std::vector<std::thread> threads;
        const int n_scan = std::max(1, std::min<int>(P->opts.scan_threads, (int)n));
        for (int i = 0; i < n_scan; i++) threads.emplace_back(scan_worker, &R);
        for (int d = 0; d < n_dev; d++)
            for (int k = 0; k < (int)P->devs[(size_t)d].lanes.size(); k++) threads.emplace_back(lane_worker, &R, d, k);
        for (auto& t : threads) t.join();
        for (auto*& s : R.scans)
            if (s) { l3b_scan_free(s); s = nullptr; }   // waves abandoned after an error
    } catch (const std::bad_alloc&) {
        P->err = "out of host memory";
        return L3B_E_MEMORY;
    } catch (const std::system_error& e) {
        P->err = std::string("cannot start worker threads: ") + e.what();
        return L3B_E_MEMORY;
    }
    {
        std::lock_guard<std::mutex> g(P->prof_mu);
        for (int i = 0; i < L3B_PIPELINE_PHASES; i++) P->prof[i] += R.prof[i];
    }
    if (out_used) *out_used = R.cursor;
    if (R.first_error) { P->err = R.err; return R.first_error; }
    return 0;
}

int l3b_pipeline_profile(l3b_pipeline_t* P, double seconds[L3B_PIPELINE_PHASES]) {
    if (!P || !seconds) return L3B_E_PARAM;
    std::lock_guard<std::mutex> g(P->prof_mu);
    for (int i = 0; i < L3B_PIPELINE_PHASES; i++) { seconds[i] = P->prof[i]; P->prof[i] = 0; }
    return 0;
}

}  // extern "C"
