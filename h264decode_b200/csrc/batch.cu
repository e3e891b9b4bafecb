// batch.cu -- h264b_scheduler: many independent streams over the GPUs of one box, in one process.
//
// Replaces the reference's connection-level concurrency (main.go:16-21: one goroutine per accepted connection running
// ByteStreamReader -> handleConnection, h264/server.go:113-166) for a batch of streams that is known up front: one
// worker thread per device, streams dealt to devices longest first (LPT by bytes).  A slice is serial work (~53 ns per
// bin), so a device's share is taken in up to three passes by how long the streams' longest slices are -- the few
// streams that hold the batch's longest slices are staged, copied, split and started first, while the bulk is still
// being staged -- and every pass runs the CABAC engine in up to six launches by slice length, side by side (the slices
// the makespan hangs on one per warp on SMs of their own, the others 32 to a warp), so that short slices are back on the
// host long before the long ones are done.  Results go straight into one pinned arena.  Host code only (one gather
// kernel); the decisions (device, pass, class) are host-only functions shared with h264b_scheduler_plan.
#include <stdlib.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace {

using Clock = std::chrono::steady_clock;

struct TrimmedStream {  // stream i from its first to its last start code
    uint64_t begin, end;
    uint32_t index;
    uint64_t longest;   // ops (or bytes) of its longest slice
};

bool is_sc(const uint8_t *p) { return p[0] == 0 && p[1] == 0 && p[2] == 0 && p[3] == 1; }

// [first start code, end of the last start code) of a stream; false: fewer than two start codes (no NAL unit)
bool trim(const uint8_t *s, uint64_t n, uint64_t *begin, uint64_t *end) {
    if (n < 8) return false;
    uint64_t b = 0;
    while (b + 4 <= n && !is_sc(s + b)) b++;
    if (b + 4 > n) return false;
    uint64_t e = n;
    while (e >= b + 8 && !is_sc(s + e - 4)) e--;
    if (e < b + 8) return false;
    *begin = b;
    *end = e;
    return true;
}

constexpr int kClasses = 6;  // slice-length classes of a pass: their CABAC launches run side by side; class 0: the
                             // slices the whole share waits for, on SMs of their own

struct Grown {  // grow-only raw buffers (pinned host or device)
    void *p = nullptr;
    size_t bytes = 0;
};

constexpr int kPasses = 3;   // of a device's share, by the streams' longest slices (see the worker)

struct Pass {  // buffers and contexts of one pass; grow-only, reused from run to run
    h264b_ctx *ctx[kClasses] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // [1] also runs the split + strip pass
    cudaEvent_t e_scan = nullptr, e_done[kClasses] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    Grown h_stream, h_fin, h_nals, h_small;                              // pinned
    Grown d_stream, d_rbsp, d_nals, d_sum, d_off, d_len, d_snal, d_offp, d_lenp, d_perm, d_nops, d_qp, d_boff, d_bins, d_fin,
        d_ops;                                                           // device
    // the run in progress
    std::vector<uint32_t> order;  // its streams (indices into the trimmed streams), ascending stream index
    std::vector<uint64_t> base;   // their offsets in the staged buffer
    std::vector<uint32_t> rows, perm;
    uint32_t cls_begin[kClasses + 1] = {0, 0, 0, 0, 0, 0, 0};
    const uint64_t *h_boff = nullptr;
    uint64_t arena_base = 0;
    uint64_t n = 0;
    uint32_t n_sl = 0, nal_cap = 0;
    bool launched[kClasses] = {false, false, false, false, false, false};
};

struct Worker {
    int device = 0;
    Pass pass[kPasses];
    std::string err;
    int rc = H264B_OK;
};

}  // namespace

struct h264b_scheduler {
    std::vector<Worker> workers;
    char err[512] = {0};
    // result storage of the last run
    std::vector<int32_t> stream_device;
    std::vector<uint32_t> stream_job;
    std::vector<uint64_t> stream_nal_off;
    std::vector<std::vector<h264b_nal>> stream_nals;
    std::vector<h264b_nal> nals;
    std::vector<h264b_cabac_final> fin;
    std::vector<uint64_t> bins_off;
    uint32_t *bins = nullptr;  // pinned, grow-only: every launch copies its bins straight to where the result says they are
    size_t bins_words = 0;
    std::vector<double> slice_done_ms;
    std::vector<double> device_busy_ms;
    std::vector<uint64_t> device_bytes;
    std::vector<uint32_t> device_jobs;
};

// slice lists in class order: (off, len) of slice perm[k] -> position k
__global__ void __launch_bounds__(256) gather_slices_kernel(const uint64_t *off, const uint32_t *len, const uint32_t *perm,
                                                            uint32_t n, uint64_t *off_p, uint32_t *len_p) {
    for (uint32_t k = blockIdx.x * 256 + threadIdx.x; k < n; k += gridDim.x * 256) {
        off_p[k] = off[perm[k]];
        len_p[k] = len[perm[k]];
    }
}

static int sched_error(h264b_scheduler *s, int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(s->err, sizeof(s->err), fmt, ap);
    va_end(ap);
    return code;
}


// ---------------------------------------------------------------------------------------------- planning (host only)
// The three decisions of a run -- which device a stream goes to, which pass of that device, which launch class a slice
// -- are pure functions of the job; h264b_scheduler_run and h264b_scheduler_plan share them.

// Streams with at least one NAL unit, from their first to their last start code.  false: a stream's slice rows are out
// of range (*bad = its index).
static bool plan_trim(const h264b_batch_job &J, std::vector<TrimmedStream> &ts, uint32_t *bad) {
    ts.clear();
    ts.reserve(J.n_streams);
    for (uint32_t i = 0; i < J.n_streams; i++) {
        const h264b_batch_stream &b = J.streams[i];
        if ((uint64_t)b.first_slice + b.n_slices > J.total_slices) {
            *bad = i;
            return false;
        }
        TrimmedStream t;
        t.index = i;
        if (!b.stream || !trim(b.stream, b.n, &t.begin, &t.end)) continue;  // no NAL unit in it
        t.longest = 0;
        for (uint32_t k = 0; k < b.n_slices; k++)
            t.longest = std::max<uint64_t>(t.longest, J.n_ops ? std::min(J.n_ops[b.first_slice + k], J.n_ops_max) : J.n_ops_max);
        if (!b.n_slices) t.longest = t.end - t.begin;
        ts.push_back(t);
    }
    return true;
}

// streams -> devices: longest first onto the least loaded device (ties: lowest device, lowest stream)
static void plan_devices(const std::vector<TrimmedStream> &ts, uint32_t nd, std::vector<std::vector<uint32_t>> &mine,
                         std::vector<uint64_t> &device_bytes) {
    std::vector<uint32_t> by_size(ts.size());
    for (size_t k = 0; k < ts.size(); k++) by_size[k] = (uint32_t)k;
    std::stable_sort(by_size.begin(), by_size.end(),
                     [&](uint32_t a, uint32_t b) { return ts[a].end - ts[a].begin > ts[b].end - ts[b].begin; });
    mine.assign(nd, {});
    device_bytes.assign(nd, 0);
    for (uint32_t k : by_size) {
        uint32_t best = 0;
        for (uint32_t d = 1; d < nd; d++)
            if (device_bytes[d] < device_bytes[best]) best = d;
        mine[best].push_back(k);
        device_bytes[best] += ts[k].end - ts[k].begin;
    }
}

// A device's streams -> passes: by their longest slice, descending; the first pass takes the streams up to 1/12 of the
// share's bytes (it is staged and copied in a few ms and holds the slices everything waits for), the second up to one
// half, the third the rest.  A share of at most group_bytes, or of fewer than 8 streams, is one pass.
static void plan_passes(const h264b_batch_job &J, const std::vector<TrimmedStream> &ts, const std::vector<uint32_t> &my,
                        uint64_t share, uint64_t group_bytes, std::vector<uint32_t> order[kPasses]) {
    for (int p = 0; p < kPasses; p++) order[p].clear();
    std::vector<uint32_t> by_len(my);
    std::stable_sort(by_len.begin(), by_len.end(), [&](uint32_t x, uint32_t y) { return ts[x].longest > ts[y].longest; });
    const bool cut = J.total_slices != 0 && share > group_bytes && by_len.size() >= 8;
    uint64_t acc = 0;
    for (uint32_t k : by_len) {
        const int p = !cut ? 0 : (acc * 12 < share ? 0 : (acc * 2 < share ? 1 : 2));
        order[p].push_back(k);
        acc += ts[k].end - ts[k].begin;
    }
}

// A pass's slices (ops[k], longest first) -> launch classes [cls_begin[c], cls_begin[c + 1]).
//   class 0: a slice that shares its scheduler runs at ~74 ns per bin instead of ~53; the bulk's launches start once
//   every pass is on the device (~0.1 ms per MB of the share), so a slice longer than (longest x 53 ns - that start) /
//   74 ns would end after the share's longest one does on its own: such slices run one per warp, four to an SM that they
//   have to themselves.  They take SMs away from everything else: *excl_budget (a third of the device over the passes
//   of a share) bounds their number; the longest ones are taken if there are more.
//   classes 1..: more than 1/2, 1/8, 1/32, 1/128 of the pass's longest slice, and the rest.
static void plan_classes(const std::vector<uint64_t> &ops, uint64_t share_top, uint64_t share_bytes, uint32_t *excl_budget,
                         uint32_t cls_begin[kClasses + 1]) {
    const uint32_t n_sl = (uint32_t)ops.size();
    const uint64_t ptop = n_sl ? ops[0] : 0;
    const double t_lone = 53e-6, t_shared = 74e-6;  // ms per bin
    const double start_ms = 0.1 * (double)share_bytes / 1e6;
    double excl = ((double)share_top * t_lone - start_ms) / t_shared;
    if (excl < 0.3 * (double)share_top) excl = 0.3 * (double)share_top;
    uint64_t thr[kClasses - 1] = {(uint64_t)excl, ptop / 2, ptop / 8, ptop / 32, ptop / 128};
    if (thr[1] > thr[0]) thr[1] = thr[0];
    uint32_t n_excl = 0;
    while (n_excl < n_sl && ops[n_excl] > thr[0]) n_excl++;
    if (n_excl > *excl_budget) n_excl = *excl_budget;
    *excl_budget -= n_excl;
    cls_begin[0] = 0;
    uint32_t k = n_excl;
    for (int c = 1; c < kClasses - 1; c++) {
        cls_begin[c] = k;
        while (k < n_sl && ops[k] > thr[c]) k++;
    }
    cls_begin[kClasses - 1] = k;
    cls_begin[kClasses] = n_sl;
}
static uint32_t plan_excl_budget(uint32_t sm_count) { return sm_count * 4u / 3u; }  // slices for class 0 (four to an SM)

extern "C" {

int32_t h264b_scheduler_create(const int32_t *devices, uint32_t n_devices, h264b_scheduler **out) {
    if (!out || !devices || !n_devices) return H264B_E_INVALID;
    *out = nullptr;
    // The classes' launches must not queue behind one another: CUDA maps a process's streams onto 8 hardware queues by
    // default, and two streams on one queue serialise.  (Read when the process initialises CUDA: set it before that.)
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    h264b_scheduler *s = new h264b_scheduler;
    s->workers.resize(n_devices);
    for (uint32_t d = 0; d < n_devices; d++) {
        Worker &w = s->workers[d];
        w.device = devices[d];
        for (int p = 0; p < kPasses; p++) {
            Pass &ps = w.pass[p];
            for (int c = 0; c < kClasses; c++) {
                const int32_t rc = h264b_create(devices[d], &ps.ctx[c]);
                if (rc != H264B_OK) {
                    h264b_scheduler_destroy(s);
                    return rc;
                }
                // several launches share the device: small CTAs, so that the short classes find room beside the long ones
                ps.ctx[c]->cabac_max_warps = 8;
                ps.ctx[c]->cabac_exclusive = c == 0 ? 1 : 0;
                ps.ctx[c]->cabac_pack = 1;  // (a warp decodes 32 slices as fast as one: few warps, each alone on its scheduler)
                if (cudaEventCreateWithFlags(&ps.e_done[c], cudaEventDisableTiming) != cudaSuccess) {
                    h264b_scheduler_destroy(s);
                    return H264B_E_CUDA;
                }
            }
            if (cudaEventCreateWithFlags(&ps.e_scan, cudaEventDisableTiming) != cudaSuccess) {
                h264b_scheduler_destroy(s);
                return H264B_E_CUDA;
            }
        }
    }
    *out = s;
    return H264B_OK;
}

void h264b_scheduler_destroy(h264b_scheduler *s) {
    if (!s) return;
    for (Worker &w : s->workers) {
        if (!w.pass[0].ctx[0]) continue;  // (never created)
        cudaSetDevice(w.device);
        cudaDeviceSynchronize();
        for (Pass &ps : w.pass) {
            for (Grown *g : {&ps.h_stream, &ps.h_fin, &ps.h_nals, &ps.h_small})
                if (g->p) cudaFreeHost(g->p);
            for (Grown *g : {&ps.d_stream, &ps.d_rbsp, &ps.d_nals, &ps.d_sum, &ps.d_off, &ps.d_len, &ps.d_snal, &ps.d_offp,
                             &ps.d_lenp, &ps.d_perm, &ps.d_nops, &ps.d_qp, &ps.d_boff, &ps.d_bins, &ps.d_fin, &ps.d_ops})
                if (g->p) cudaFree(g->p);
            if (ps.e_scan) cudaEventDestroy(ps.e_scan);
            for (int c = 0; c < kClasses; c++) {
                if (ps.e_done[c]) cudaEventDestroy(ps.e_done[c]);
                if (ps.ctx[c]) h264b_destroy(ps.ctx[c]);
            }
        }
    }
    if (s->bins) cudaFreeHost(s->bins);
    delete s;
}

const char *h264b_scheduler_last_error(const h264b_scheduler *s) { return s ? s->err : "no scheduler"; }

int32_t h264b_scheduler_run(h264b_scheduler *s, const h264b_batch_job *job, h264b_batch_result *res) {
    if (!s || !job || !res) return H264B_E_INVALID;
    const h264b_batch_job &J = *job;
    if ((J.n_streams && !J.streams) || (J.total_slices && !J.qp) || (J.n_ops_max && !J.ops))
        return sched_error(s, H264B_E_INVALID, "scheduler_run: null pointer in job");
    const uint32_t nd = (uint32_t)s->workers.size();
    const uint64_t group_bytes = J.group_bytes ? J.group_bytes : (32ull << 20);  // a smaller share is taken in one pass

    // ---- bins: one pinned arena for the whole batch.  A slice's words are fixed by its op count; where they lie is decided
    // by the device, pass and class the slice lands in (a launch's bins are one block), and reported per slice.
    auto words_of = [&](uint32_t r) -> uint64_t {
        uint32_t nb = J.n_ops ? J.n_ops[r] : J.n_ops_max;
        if (nb > J.n_ops_max) nb = J.n_ops_max;
        return ((uint64_t)nb + 1 + 31) / 32;
    };
    uint64_t all_words = 0;
    for (uint32_t r = 0; r < J.total_slices; r++) all_words += words_of(r);
    s->bins_off.assign((size_t)J.total_slices + 1, ~0ull);
    s->bins_off[J.total_slices] = all_words;
    if (all_words + 16 > s->bins_words) {
        if (s->bins) cudaFreeHost(s->bins);
        s->bins = nullptr;
        s->bins_words = 0;
        cudaSetDevice(s->workers[0].device);
        const size_t want = (size_t)(all_words + all_words / 16 + 16);
        if (cudaHostAlloc((void **)&s->bins, want * 4, cudaHostAllocPortable) != cudaSuccess)
            return sched_error(s, H264B_E_NOMEM, "scheduler_run: no pinned memory for %llu MB of bins", (unsigned long long)(want * 4 >> 20));
        s->bins_words = want;
    }
    s->fin.assign(J.total_slices, h264b_cabac_final{});
    s->slice_done_ms.assign(J.total_slices, 0.0);
    s->stream_device.assign(J.n_streams, -1);
    s->stream_job.assign(J.n_streams, 0);
    s->stream_nals.assign(J.n_streams, {});
    s->device_busy_ms.assign(nd, 0.0);
    s->device_bytes.assign(nd, 0);
    s->device_jobs.assign(nd, 0);
    std::vector<TrimmedStream> ts;
    {
        uint32_t bad = 0;
        if (!plan_trim(J, ts, &bad)) return sched_error(s, H264B_E_INVALID, "scheduler_run: stream %u: slice rows out of range", bad);
    }
    std::vector<std::vector<uint32_t>> mine;
    plan_devices(ts, nd, mine, s->device_bytes);
    for (uint32_t d = 0; d < nd; d++)
        for (uint32_t k : mine[d]) s->stream_device[ts[k].index] = (int32_t)d;
    // the devices' regions of the bins arena; slices of streams that go nowhere (no NAL unit in them) get zeroed words
    // behind them
    std::vector<uint64_t> device_words(nd + 1, 0);
    {
        std::vector<char> placed(J.n_streams, 0);
        for (uint32_t d = 0; d < nd; d++) {
            uint64_t wsum = 0;
            for (uint32_t k : mine[d]) {
                const h264b_batch_stream &b = J.streams[ts[k].index];
                placed[ts[k].index] = 1;
                for (uint32_t x = 0; x < b.n_slices; x++) wsum += words_of(b.first_slice + x);
            }
            device_words[d + 1] = device_words[d] + wsum;
        }
        uint64_t at = device_words[nd];
        for (uint32_t i = 0; i < J.n_streams; i++) {
            if (placed[i]) continue;
            const h264b_batch_stream &b = J.streams[i];
            for (uint32_t x = 0; x < b.n_slices; x++) {
                s->bins_off[b.first_slice + x] = at;
                const uint64_t wn = words_of(b.first_slice + x);
                memset(s->bins + at, 0, (size_t)wn * 4);
                at += wn;
            }
        }
        for (uint32_t r = 0; r < J.total_slices; r++)  // rows no stream claims
            if (s->bins_off[r] == ~0ull && at + words_of(r) <= all_words) {
                s->bins_off[r] = at;
                memset(s->bins + at, 0, (size_t)words_of(r) * 4);
                at += words_of(r);
            }
        for (uint32_t r = 0; r < J.total_slices; r++)  // (rows claimed twice leave no room for the others: a caller's error)
            if (s->bins_off[r] == ~0ull) s->bins_off[r] = 0;
    }
    // ---- one worker thread per device.  The device's share is taken in up to kPasses passes: the streams with the longest
    // slices first (a slice is serial work: the pass that holds the batch's longest slices is small, and its CABAC launch
    // is running while the bulk of the share is still being staged).  A pass: stage to pinned memory, copy in, ONE split +
    // strip pass, then its slices in kClasses launches by length (longest class first, each on its own context = CUDA
    // stream, side by side) -- the launch holding the 1 MB slices lasts two orders of magnitude longer than the one
    // holding the 1 KB slices, whose results are on the host long before (per-slice completion times: tail latency).
    const Clock::time_point t_start = Clock::now();
    auto ms_since = [&](Clock::time_point t) { return std::chrono::duration<double, std::milli>(t - t_start).count(); };
    const bool sched_trace = getenv("H264B_SCHED_TRACE") != nullptr;
    std::vector<std::thread> threads;
    for (uint32_t d = 0; d < nd; d++) {
        threads.emplace_back([&, d]() {
            h264b::TraceRange trace_range("h264b:scheduler_worker");
            Worker &w = s->workers[d];
            w.rc = H264B_OK;
            w.err.clear();
            cudaSetDevice(w.device);
            auto fail = [&](int rc, const std::string &what) {
                w.rc = rc;
                w.err = what;
                return false;
            };
            auto grow_pin = [&](Grown &g, size_t bytes) -> bool {
                if (bytes < 256) bytes = 256;
                if (g.bytes >= bytes) return true;
                if (g.p) cudaFreeHost(g.p);
                g.p = nullptr, g.bytes = 0;
                if (cudaHostAlloc(&g.p, bytes + bytes / 8, cudaHostAllocDefault) != cudaSuccess) return false;
                g.bytes = bytes + bytes / 8;
                return true;
            };
            auto grow_dev = [&](Grown &g, size_t bytes) -> bool {
                if (bytes < 256) bytes = 256;
                if (g.bytes >= bytes) return true;
                if (g.p) cudaFree(g.p);
                g.p = nullptr, g.bytes = 0;
                if (cudaMalloc(&g.p, bytes + bytes / 8) != cudaSuccess) return false;
                g.bytes = bytes + bytes / 8;
                return true;
            };
            const std::vector<uint32_t> &my = mine[d];
            if (my.empty()) return;
            const Clock::time_point t_first = Clock::now();
            auto ops_of = [&](uint32_t row) { return J.n_ops ? std::min(J.n_ops[row], J.n_ops_max) : J.n_ops_max; };
            // ---- the share's streams -> passes (plan_passes)
            for (Pass &ps : w.pass) {
                ps.order.clear();
                ps.n = 0;
                ps.n_sl = 0;
                for (bool &l : ps.launched) l = false;
            }
            uint64_t share_top = 0;  // ops of the share's longest slice
            for (uint32_t k : my) share_top = std::max(share_top, ts[k].longest);
            {
                std::vector<uint32_t> order[kPasses];
                plan_passes(J, ts, my, s->device_bytes[d], group_bytes, order);
                for (int p = 0; p < kPasses; p++) w.pass[p].order.swap(order[p]);
            }
            uint32_t n_passes = 0;
            uint64_t arena_at = device_words[d];  // where the next pass's bins go
            uint32_t excl_budget = plan_excl_budget((uint32_t)w.pass[0].ctx[0]->sm_count);

            // ---- the CABAC launches of classes [c_from, c_to) of a pass and their results, each on its own stream
            auto launch_classes = [&](Pass &ps, int c_from, int c_to) -> bool {
                for (int c = c_from; c < c_to; c++) {
                    const uint32_t k0 = ps.cls_begin[c], k1 = ps.cls_begin[c + 1];
                    if (k1 == k0) continue;
                    h264b_ctx *cc = ps.ctx[c];
                    cudaStream_t sc = cc->stream;
                    cudaStreamWaitEvent(sc, ps.e_scan, 0);
                    h264b_cabac_job cj;
                    memset(&cj, 0, sizeof(cj));
                    cj.bytes = (const uint8_t *)ps.d_rbsp.p;
                    cj.total_bytes = ps.n + 16;
                    cj.off = (const uint64_t *)ps.d_offp.p + k0;
                    cj.len = (const uint32_t *)ps.d_lenp.p + k0;
                    cj.n_slices = k1 - k0;
                    cj.n_ctx = J.n_ctx;
                    cj.ops = (const uint16_t *)ps.d_ops.p;
                    cj.n_ops_max = J.n_ops_max;
                    cj.n_ops = (const uint32_t *)ps.d_nops.p + k0;
                    cj.qp = (const h264b_slice_qp *)ps.d_qp.p + k0;
                    cj.bins = (uint32_t *)ps.d_bins.p;
                    cj.bins_off = (const uint64_t *)ps.d_boff.p + k0;
                    cj.final = (h264b_cabac_final *)ps.d_fin.p + k0;
                    cj.flags = J.flags & (H264B_TABLES_SPEC | H264B_BYPASS_SPEC_OR | H264B_CABAC_FINAL_TERMINATE);
                    const int rc = h264b_cabac_decode_dev(cc, &cj);
                    if (rc != H264B_OK) return fail(rc, std::string("cabac: ") + h264b_last_error(cc));
                    const size_t w0 = (size_t)ps.h_boff[k0], w1 = (size_t)ps.h_boff[k1];
                    cudaMemcpyAsync(s->bins + ps.arena_base + w0, (const uint32_t *)ps.d_bins.p + w0, (w1 - w0) * 4, cudaMemcpyDeviceToHost, sc);
                    cudaMemcpyAsync((h264b_cabac_final *)ps.h_fin.p + k0, (const h264b_cabac_final *)ps.d_fin.p + k0,
                                    (size_t)(k1 - k0) * sizeof(h264b_cabac_final), cudaMemcpyDeviceToHost, sc);
                    cudaEventRecord(ps.e_done[c], sc);
                    ps.launched[c] = true;
                }
                return true;
            };
            // ---- one pass: staged, copied in, split + strip, and its class 0 launched; returns with the work in flight
            auto enqueue = [&](Pass &ps, int pass_no) -> bool {
                // streams in stream-index order (so that a pass's slice rows ascend), staged back to back
                std::vector<uint32_t> &order = ps.order;
                std::sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return ts[x].index < ts[y].index; });
                ps.base.assign(order.size() + 1, 0);
                uint32_t n_sl = 0;
                for (size_t k = 0; k < order.size(); k++) {
                    ps.base[k + 1] = ps.base[k] + (ts[order[k]].end - ts[order[k]].begin);
                    n_sl += J.streams[ts[order[k]].index].n_slices;
                    s->stream_job[ts[order[k]].index] = (uint32_t)pass_no;
                }
                const uint64_t n = ps.base.back();
                ps.n = n;
                ps.n_sl = n_sl;
                if (!grow_pin(ps.h_stream, n + 64) || !grow_dev(ps.d_stream, n + 64) || !grow_dev(ps.d_rbsp, n + 64))
                    return fail(H264B_E_NOMEM, "out of memory staging the device's streams");
                {   // staging: a large pass is copied by several threads
                    const unsigned helpers = n > (64ull << 20) ? 4u : 1u;
                    auto stage = [&](size_t k0, size_t k1) {
                        for (size_t k = k0; k < k1; k++) {
                            const TrimmedStream &t = ts[order[k]];
                            memcpy((uint8_t *)ps.h_stream.p + ps.base[k], J.streams[t.index].stream + t.begin, (size_t)(t.end - t.begin));
                        }
                    };
                    if (helpers == 1) {
                        stage(0, order.size());
                    } else {
                        std::vector<std::thread> hs;
                        size_t k0 = 0;
                        for (unsigned h = 0; h < helpers; h++) {  // equal byte shares
                            size_t k1 = k0;
                            const uint64_t until = n * (h + 1) / helpers;
                            while (k1 < order.size() && ps.base[k1 + 1] <= until) k1++;
                            if (h + 1 == helpers) k1 = order.size();
                            hs.emplace_back(stage, k0, k1);
                            k0 = k1;
                        }
                        for (std::thread &t : hs) t.join();
                    }
                }
                const double t_staged = sched_trace ? ms_since(Clock::now()) : 0.0;
                // the pass's slice rows, and their order by length (longest first; ties: the lower row)
                ps.rows.clear();
                ps.rows.reserve(n_sl);
                for (uint32_t k : order) {
                    const h264b_batch_stream &b = J.streams[ts[k].index];
                    for (uint32_t x = 0; x < b.n_slices; x++) ps.rows.push_back(b.first_slice + x);
                }
                const std::vector<uint32_t> &rows = ps.rows;
                std::vector<uint32_t> &perm = ps.perm;
                perm.resize(n_sl);
                for (uint32_t k = 0; k < n_sl; k++) perm[k] = k;
                std::stable_sort(perm.begin(), perm.end(), [&](uint32_t x, uint32_t y) { return ops_of(rows[x]) > ops_of(rows[y]); });
                // launch classes (plan_classes)
                {
                    std::vector<uint64_t> sorted_ops(n_sl);
                    for (uint32_t k = 0; k < n_sl; k++) sorted_ops[k] = ops_of(rows[perm[k]]);
                    plan_classes(sorted_ops, share_top, s->device_bytes[d], &excl_budget, ps.cls_begin);
                }
                const uint32_t nal_cap = (uint32_t)std::min<uint64_t>(n / 64 + 1024 + 2 * (uint64_t)order.size(), 0xFFFFFFF0ull);
                ps.nal_cap = nal_cap;
                const size_t ms = n_sl ? n_sl : 1;
                // per-slice inputs in class order
                if (!grow_pin(ps.h_small, ms * (4 + 4 + sizeof(h264b_slice_qp) + 8) + 8 + 256))
                    return fail(H264B_E_NOMEM, "out of pinned memory");
                uint32_t *h_perm = (uint32_t *)ps.h_small.p;
                uint32_t *h_nops = h_perm + ms;
                h264b_slice_qp *h_qp = (h264b_slice_qp *)(h_nops + ms);
                uint64_t *h_boff = (uint64_t *)(h_qp + ms);
                ps.h_boff = h_boff;
                h_boff[0] = 0;
                for (uint32_t k = 0; k < n_sl; k++) {
                    const uint32_t row = rows[perm[k]];
                    h_perm[k] = perm[k];
                    h_nops[k] = ops_of(row);
                    h_qp[k] = J.qp[row];
                    h_boff[k + 1] = h_boff[k] + ((uint64_t)h_nops[k] + 1 + 31) / 32;
                }
                const size_t total_words = (size_t)h_boff[n_sl];
                ps.arena_base = arena_at;  // the pass's bins in the arena, in class order
                arena_at += total_words;
                for (uint32_t k = 0; k < n_sl; k++) s->bins_off[rows[perm[k]]] = ps.arena_base + h_boff[k];
                if (!grow_dev(ps.d_nals, (size_t)nal_cap * sizeof(h264b_nal)) || !grow_dev(ps.d_sum, 256) ||
                    !grow_dev(ps.d_off, ms * 8) || !grow_dev(ps.d_len, ms * 4) || !grow_dev(ps.d_snal, ms * 4 + 16) ||
                    !grow_dev(ps.d_offp, ms * 8) || !grow_dev(ps.d_lenp, ms * 4) || !grow_dev(ps.d_perm, ms * 4) ||
                    !grow_dev(ps.d_nops, ms * 4) || !grow_dev(ps.d_qp, ms * sizeof(h264b_slice_qp)) ||
                    !grow_dev(ps.d_boff, (ms + 1) * 8) || !grow_dev(ps.d_bins, total_words * 4 + 16) ||
                    !grow_dev(ps.d_fin, ms * sizeof(h264b_cabac_final)) || !grow_dev(ps.d_ops, (size_t)J.n_ops_max * 2 + 16) ||
                    !grow_pin(ps.h_fin, ms * sizeof(h264b_cabac_final)) ||
                    !grow_pin(ps.h_nals, (size_t)nal_cap * sizeof(h264b_nal) + 256))
                    return fail(H264B_E_NOMEM, "out of memory for the device's slice arrays");
                // ---- split + strip, slice list, class order (context 0's stream)
                h264b_ctx *c0 = ps.ctx[1];
                cudaStream_t s0 = c0->stream;
                bool ok = cudaMemcpyAsync(ps.d_stream.p, ps.h_stream.p, n, cudaMemcpyHostToDevice, s0) == cudaSuccess;
                ok = ok && cudaMemcpyAsync(ps.d_perm.p, h_perm, ms * 4, cudaMemcpyHostToDevice, s0) == cudaSuccess;
                ok = ok && cudaMemcpyAsync(ps.d_nops.p, h_nops, ms * 4, cudaMemcpyHostToDevice, s0) == cudaSuccess;
                ok = ok && cudaMemcpyAsync(ps.d_qp.p, h_qp, ms * sizeof(h264b_slice_qp), cudaMemcpyHostToDevice, s0) == cudaSuccess;
                ok = ok && cudaMemcpyAsync(ps.d_boff.p, h_boff, (ms + 1) * 8, cudaMemcpyHostToDevice, s0) == cudaSuccess;
                if (J.n_ops_max)
                    ok = ok && cudaMemcpyAsync(ps.d_ops.p, J.ops, (size_t)J.n_ops_max * 2, cudaMemcpyHostToDevice, s0) == cudaSuccess;
                if (!ok) return fail(H264B_E_CUDA, "copying the device's share in failed");
                double t_copied = 0.0;
                if (sched_trace) {
                    cudaStreamSynchronize(s0);
                    t_copied = ms_since(Clock::now());
                }
                uint32_t *d_ns = (uint32_t *)((uint8_t *)ps.d_sum.p + 64);
                int rc = h264b_annexb_scan_dev(c0, (const uint8_t *)ps.d_stream.p, n, (uint8_t *)ps.d_rbsp.p, (h264b_nal *)ps.d_nals.p,
                                               nullptr, nal_cap, (h264b_scan_summary *)ps.d_sum.p, 0);
                if (rc == H264B_OK)
                    rc = h264b_slice_select_dev(c0, (const h264b_nal *)ps.d_nals.p, (const h264b_scan_summary *)ps.d_sum.p, nal_cap,
                                                J.slice_data_offset, n_sl, (uint64_t *)ps.d_off.p, (uint32_t *)ps.d_len.p,
                                                (uint32_t *)ps.d_snal.p, d_ns);
                if (rc != H264B_OK) return fail(rc, std::string("split + strip: ") + h264b_last_error(c0));
                if (n_sl)
                    gather_slices_kernel<<<(n_sl + 255) / 256, 256, 0, s0>>>((const uint64_t *)ps.d_off.p, (const uint32_t *)ps.d_len.p,
                                                                            (const uint32_t *)ps.d_perm.p, n_sl, (uint64_t *)ps.d_offp.p,
                                                                            (uint32_t *)ps.d_lenp.p);
                cudaEventRecord(ps.e_scan, s0);
                if (sched_trace) {
                    cudaStreamSynchronize(s0);
                    fprintf(stderr, "h264b scheduler: device %d pass %d: %zu streams, %.1f MB staged at %.1f ms, on the device at %.1f ms, "
                            "split + strip done at %.1f ms\n", w.device, pass_no, order.size(), n / 1e6, t_staged, t_copied,
                            ms_since(Clock::now()));
                }
                return launch_classes(ps, 0, 1);  // class 0 at once; the others once every pass is on the device
            };
            // ---- results of the launches that have finished (all passes); false: a launch failed
            auto collect = [&](bool *progressed) -> bool {
                for (int p = 0; p < kPasses; p++) {
                    Pass &ps = w.pass[p];
                    for (int c = 0; c < kClasses; c++) {
                        if (!ps.launched[c]) continue;
                        const cudaError_t q = cudaEventQuery(ps.e_done[c]);
                        if (q == cudaErrorNotReady) continue;
                        if (q != cudaSuccess) return fail(H264B_E_CUDA, std::string("a class launch failed: ") + cudaGetErrorString(q));
                        const double t_done = ms_since(Clock::now());
                        if (sched_trace)
                            fprintf(stderr, "h264b scheduler: device %d pass %d class %d: %u slices (longest %u ops) done at %.1f ms\n",
                                    w.device, p, c, ps.cls_begin[c + 1] - ps.cls_begin[c], ops_of(ps.rows[ps.perm[ps.cls_begin[c]]]), t_done);
                        for (uint32_t k = ps.cls_begin[c]; k < ps.cls_begin[c + 1]; k++) {
                            const uint32_t row = ps.rows[ps.perm[k]];
                            s->fin[row] = ((const h264b_cabac_final *)ps.h_fin.p)[k];
                            s->slice_done_ms[row] = t_done;
                        }
                        ps.launched[c] = false;
                        *progressed = true;
                    }
                }
                return true;
            };
            for (int p = 0; p < kPasses; p++) {
                if (w.pass[p].order.empty()) continue;
                if (!enqueue(w.pass[p], (int)n_passes)) return;
                n_passes++;
            }
            // every pass is on the device and split (their kernels did not have to queue behind CABAC launches that fill the
            // SMs): now the bulk -- the long classes first, they end last
            for (int c = 1; c < kClasses; c++)
                for (int p = 0; p < kPasses; p++) {
                    Pass &ps = w.pass[p];
                    if (ps.order.empty()) continue;
                    if (!launch_classes(ps, c, c + 1)) return;
                }
            for (Pass &ps : w.pass) {  // the NAL index (behind the launch on that context's stream)
                if (ps.order.empty()) continue;
                cudaStream_t s1 = ps.ctx[1]->stream;
                cudaMemcpyAsync(ps.h_nals.p, ps.d_sum.p, 128, cudaMemcpyDeviceToHost, s1);
                cudaMemcpyAsync((uint8_t *)ps.h_nals.p + 256, ps.d_nals.p, (size_t)ps.nal_cap * sizeof(h264b_nal), cudaMemcpyDeviceToHost, s1);
            }
            s->device_jobs[d] = n_passes;
            for (;;) {
                bool pending = false, progressed = false;
                if (!collect(&progressed)) return;
                for (Pass &ps : w.pass)
                    for (bool l : ps.launched) pending = pending || l;
                if (!pending) break;
                if (!progressed) std::this_thread::sleep_for(std::chrono::microseconds(50));
            }
            // ---- the passes' NAL units
            for (Pass &ps : w.pass) {
                if (ps.order.empty()) continue;
                if (cudaStreamSynchronize(ps.ctx[1]->stream) != cudaSuccess) {
                    fail(H264B_E_CUDA, "the NAL index did not arrive");
                    return;
                }
                const h264b_scan_summary *sum = (const h264b_scan_summary *)ps.h_nals.p;
                const uint32_t found = *(const uint32_t *)((const uint8_t *)ps.h_nals.p + 64);
                if (sum->status != H264B_OK) {
                    fail(H264B_E_CAPACITY, "more NAL units than the index holds");
                    return;
                }
                if (found != ps.n_sl) {
                    fail(H264B_E_INVALID, "the device's streams hold " + std::to_string(found) + " slice NAL units, the batch announces " +
                                              std::to_string(ps.n_sl));
                    return;
                }
                // NAL units: those that lie inside one stream's staged extent (the 4-byte unit that the next stream's
                // leading start code forms is nobody's)
                const h264b_nal *un = (const h264b_nal *)((const uint8_t *)ps.h_nals.p + 256);
                size_t k = 0;
                for (uint64_t i = 0; i < sum->n_nals; i++) {
                    const h264b_nal &u = un[i];
                    while (k + 1 < ps.order.size() && u.start >= ps.base[k + 1]) k++;
                    const TrimmedStream &t = ts[ps.order[k]];
                    const uint64_t lo = ps.base[k], hi = ps.base[k + 1];
                    if (u.start < lo + 4 || u.start + u.num_bytes > hi) continue;
                    h264b_nal v = u;
                    v.start = u.start - lo + t.begin;
                    v.rbsp_off = u.rbsp_off - lo + t.begin;
                    s->stream_nals[t.index].push_back(v);
                }
            }
            s->device_busy_ms[d] = std::chrono::duration<double, std::milli>(Clock::now() - t_first).count();
        });
    }
    for (std::thread &t : threads) t.join();
    const double makespan = ms_since(Clock::now());
    for (uint32_t d = 0; d < nd; d++)
        if (s->workers[d].rc != H264B_OK)
            return sched_error(s, s->workers[d].rc, "device %d: %s", s->workers[d].device, s->workers[d].err.c_str());

    // ---- assemble
    s->stream_nal_off.assign((size_t)J.n_streams + 1, 0);
    for (uint32_t i = 0; i < J.n_streams; i++) s->stream_nal_off[i + 1] = s->stream_nal_off[i] + s->stream_nals[i].size();
    s->nals.clear();
    s->nals.reserve((size_t)s->stream_nal_off[J.n_streams]);
    for (uint32_t i = 0; i < J.n_streams; i++) s->nals.insert(s->nals.end(), s->stream_nals[i].begin(), s->stream_nals[i].end());
    memset(res, 0, sizeof(*res));
    res->stream_device = s->stream_device.data();
    res->stream_job = s->stream_job.data();
    res->stream_nal_off = s->stream_nal_off.data();
    res->nals = s->nals.data();
    res->final = s->fin.data();
    res->bins_off = s->bins_off.data();
    res->bins = s->bins;
    res->slice_done_ms = s->slice_done_ms.data();
    res->n_devices = nd;
    res->device_busy_ms = s->device_busy_ms.data();
    res->device_bytes = s->device_bytes.data();
    res->device_jobs = s->device_jobs.data();
    res->makespan_ms = makespan;
    res->total_nals = s->nals.size();
    for (uint32_t r = 0; r < J.total_slices; r++) res->total_bins += s->fin[r].n_bins;
    return H264B_OK;
}

int32_t h264b_scheduler_plan(const h264b_batch_job *job, uint32_t n_devices, uint32_t sm_count, int32_t *stream_device,
                             uint32_t *stream_pass, uint8_t *slice_class) {
    if (!job || !n_devices) return H264B_E_INVALID;
    const h264b_batch_job &J = *job;
    if (J.n_streams && !J.streams) return H264B_E_INVALID;
    const uint64_t group_bytes = J.group_bytes ? J.group_bytes : (32ull << 20);
    std::vector<TrimmedStream> ts;
    uint32_t bad = 0;
    if (!plan_trim(J, ts, &bad)) return H264B_E_INVALID;
    std::vector<std::vector<uint32_t>> mine;
    std::vector<uint64_t> device_bytes;
    plan_devices(ts, n_devices, mine, device_bytes);
    for (uint32_t i = 0; i < J.n_streams; i++) {
        if (stream_device) stream_device[i] = -1;
        if (stream_pass) stream_pass[i] = 0;
    }
    if (slice_class) memset(slice_class, 255, J.total_slices);
    auto ops_of = [&](uint32_t row) { return J.n_ops ? std::min(J.n_ops[row], J.n_ops_max) : J.n_ops_max; };
    for (uint32_t d = 0; d < n_devices; d++) {
        const std::vector<uint32_t> &my = mine[d];
        uint64_t share_top = 0;
        for (uint32_t k : my) share_top = std::max(share_top, ts[k].longest);
        std::vector<uint32_t> order[kPasses];
        plan_passes(J, ts, my, device_bytes[d], group_bytes, order);
        uint32_t excl_budget = plan_excl_budget(sm_count);
        uint32_t pass_no = 0;
        for (int p = 0; p < kPasses; p++) {
            if (order[p].empty()) continue;
            std::sort(order[p].begin(), order[p].end(), [&](uint32_t x, uint32_t y) { return ts[x].index < ts[y].index; });
            std::vector<uint32_t> rows;
            for (uint32_t k : order[p]) {
                const h264b_batch_stream &b = J.streams[ts[k].index];
                if (stream_device) stream_device[ts[k].index] = (int32_t)d;
                if (stream_pass) stream_pass[ts[k].index] = pass_no;
                for (uint32_t x = 0; x < b.n_slices; x++) rows.push_back(b.first_slice + x);
            }
            std::vector<uint32_t> perm(rows.size());
            for (uint32_t k = 0; k < perm.size(); k++) perm[k] = k;
            std::stable_sort(perm.begin(), perm.end(), [&](uint32_t x, uint32_t y) { return ops_of(rows[x]) > ops_of(rows[y]); });
            std::vector<uint64_t> sorted_ops(rows.size());
            for (uint32_t k = 0; k < perm.size(); k++) sorted_ops[k] = ops_of(rows[perm[k]]);
            uint32_t cls_begin[kClasses + 1];
            plan_classes(sorted_ops, share_top, device_bytes[d], &excl_budget, cls_begin);
            if (slice_class)
                for (int c = 0; c < kClasses; c++)
                    for (uint32_t k = cls_begin[c]; k < cls_begin[c + 1]; k++) slice_class[rows[perm[k]]] = (uint8_t)c;
            pass_no++;
        }
    }
    return H264B_OK;
}

}  // extern "C"
