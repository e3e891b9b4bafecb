// batch.cu -- h264b_scheduler: many independent streams over the GPUs of one box, in one process.
//
// Replaces the reference's connection-level concurrency (main.go:16-21: one goroutine per accepted connection running
// ByteStreamReader -> handleConnection, h264/server.go:113-166) for a batch of streams that is known up front: one
// worker thread and one h264b context per device, streams dealt to devices longest first (LPT by bytes), grouped into
// device jobs with the longest slices first, three jobs in flight per device through the same h264b_stream_submit /
// h264b_stream_wait any single-stream caller uses.  Host code only: no kernel lives here.
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

constexpr int kClasses = 5;  // slice-length classes of a device's share: their CABAC launches run side by side

struct Grown {  // grow-only raw buffers (pinned host or device)
    void *p = nullptr;
    size_t bytes = 0;
};

struct Worker {
    int device = 0;
    h264b_ctx *ctx[kClasses] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // [0] also runs the split + strip pass
    cudaEvent_t e_scan = nullptr, e_done[kClasses] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    Grown h_stream, h_bins, h_fin, h_nals, h_small;                      // pinned
    Grown d_stream, d_rbsp, d_nals, d_sum, d_off, d_len, d_snal, d_offp, d_lenp, d_perm, d_nops, d_qp, d_boff, d_bins, d_fin,
        d_ops;                                                           // device
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
    std::vector<uint32_t> bins;
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
        for (int c = 0; c < kClasses; c++) {
            const int32_t rc = h264b_create(devices[d], &w.ctx[c]);
            if (rc != H264B_OK) {
                h264b_scheduler_destroy(s);
                return rc;
            }
            // several launches share the device: small CTAs, so that the short classes find room beside the long ones
            w.ctx[c]->cabac_max_warps = 8;
            if (cudaEventCreateWithFlags(&w.e_done[c], cudaEventDisableTiming) != cudaSuccess) {
                h264b_scheduler_destroy(s);
                return H264B_E_CUDA;
            }
        }
        if (cudaEventCreateWithFlags(&w.e_scan, cudaEventDisableTiming) != cudaSuccess) {
            h264b_scheduler_destroy(s);
            return H264B_E_CUDA;
        }
    }
    *out = s;
    return H264B_OK;
}

void h264b_scheduler_destroy(h264b_scheduler *s) {
    if (!s) return;
    for (Worker &w : s->workers) {
        if (!w.ctx[0]) continue;
        cudaSetDevice(w.device);
        cudaDeviceSynchronize();
        for (Grown *g : {&w.h_stream, &w.h_bins, &w.h_fin, &w.h_nals, &w.h_small})
            if (g->p) cudaFreeHost(g->p);
        for (Grown *g : {&w.d_stream, &w.d_rbsp, &w.d_nals, &w.d_sum, &w.d_off, &w.d_len, &w.d_snal, &w.d_offp, &w.d_lenp,
                         &w.d_perm, &w.d_nops, &w.d_qp, &w.d_boff, &w.d_bins, &w.d_fin, &w.d_ops})
            if (g->p) cudaFree(g->p);
        if (w.e_scan) cudaEventDestroy(w.e_scan);
        for (int c = 0; c < kClasses; c++) {
            if (w.e_done[c]) cudaEventDestroy(w.e_done[c]);
            if (w.ctx[c]) h264b_destroy(w.ctx[c]);
        }
    }
    delete s;
}

const char *h264b_scheduler_last_error(const h264b_scheduler *s) { return s ? s->err : "no scheduler"; }

int32_t h264b_scheduler_run(h264b_scheduler *s, const h264b_batch_job *job, h264b_batch_result *res) {
    if (!s || !job || !res) return H264B_E_INVALID;
    const h264b_batch_job &J = *job;
    if ((J.n_streams && !J.streams) || (J.total_slices && !J.qp) || (J.n_ops_max && !J.ops))
        return sched_error(s, H264B_E_INVALID, "scheduler_run: null pointer in job");
    const uint32_t nd = (uint32_t)s->workers.size();
    const uint64_t group_bytes = J.group_bytes ? J.group_bytes : (16ull << 20);

    // ---- bins layout (fixed by the op counts) and the streams' extents
    s->bins_off.assign((size_t)J.total_slices + 1, 0);
    for (uint32_t r = 0; r < J.total_slices; r++) {
        uint32_t nb = J.n_ops ? J.n_ops[r] : J.n_ops_max;
        if (nb > J.n_ops_max) nb = J.n_ops_max;
        s->bins_off[r + 1] = s->bins_off[r] + ((uint64_t)nb + 1 + 31) / 32;
    }
    s->bins.assign((size_t)s->bins_off[J.total_slices], 0u);
    s->fin.assign(J.total_slices, h264b_cabac_final{});
    s->slice_done_ms.assign(J.total_slices, 0.0);
    s->stream_device.assign(J.n_streams, -1);
    s->stream_job.assign(J.n_streams, 0);
    s->stream_nals.assign(J.n_streams, {});
    s->device_busy_ms.assign(nd, 0.0);
    s->device_bytes.assign(nd, 0);
    s->device_jobs.assign(nd, 0);
    std::vector<TrimmedStream> ts;
    ts.reserve(J.n_streams);
    for (uint32_t i = 0; i < J.n_streams; i++) {
        const h264b_batch_stream &b = J.streams[i];
        if ((uint64_t)b.first_slice + b.n_slices > J.total_slices)
            return sched_error(s, H264B_E_INVALID, "scheduler_run: stream %u: slice rows out of range", i);
        TrimmedStream t;
        t.index = i;
        if (!b.stream || !trim(b.stream, b.n, &t.begin, &t.end)) continue;  // no NAL unit in it
        t.longest = 0;
        for (uint32_t k = 0; k < b.n_slices; k++)
            t.longest = std::max<uint64_t>(t.longest, J.n_ops ? J.n_ops[b.first_slice + k] : J.n_ops_max);
        if (!b.n_slices) t.longest = t.end - t.begin;
        ts.push_back(t);
    }
    // ---- streams -> devices: longest first onto the least loaded device (ties: lowest device, lowest stream)
    std::vector<uint32_t> by_size(ts.size());
    for (size_t k = 0; k < ts.size(); k++) by_size[k] = (uint32_t)k;
    std::stable_sort(by_size.begin(), by_size.end(),
                     [&](uint32_t a, uint32_t b) { return ts[a].end - ts[a].begin > ts[b].end - ts[b].begin; });
    std::vector<std::vector<uint32_t>> mine(nd);
    for (uint32_t k : by_size) {
        uint32_t best = 0;
        for (uint32_t d = 1; d < nd; d++)
            if (s->device_bytes[d] < s->device_bytes[best]) best = d;
        mine[best].push_back(k);
        s->device_bytes[best] += ts[k].end - ts[k].begin;
        s->stream_device[ts[k].index] = (int32_t)best;
    }
    (void)group_bytes;
    // ---- one worker thread per device.  The device's share goes through ONE split + strip pass; its slices then run in
    // kClasses launches by length (longest class first, each on its own context = CUDA stream, side by side): a slice is
    // serial work, so the launch holding the 1 MB slices lasts two orders of magnitude longer than the one holding the
    // 1 KB slices, whose results are on the host long before (per-slice completion times: tail latency).
    const Clock::time_point t_start = Clock::now();
    auto ms_since = [&](Clock::time_point t) { return std::chrono::duration<double, std::milli>(t - t_start).count(); };
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
            s->device_jobs[d] = 1;
            // streams in stream-index order (so that a device's slice rows ascend), staged back to back
            std::vector<uint32_t> order(my);
            std::sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return ts[x].index < ts[y].index; });
            std::vector<uint64_t> base(order.size() + 1, 0);
            uint32_t n_sl = 0;
            for (size_t k = 0; k < order.size(); k++) {
                base[k + 1] = base[k] + (ts[order[k]].end - ts[order[k]].begin);
                n_sl += J.streams[ts[order[k]].index].n_slices;
                s->stream_job[ts[order[k]].index] = 0;
            }
            const uint64_t n = base.back();
            const Clock::time_point t_first = Clock::now();
            if (!grow_pin(w.h_stream, n + 64) || !grow_dev(w.d_stream, n + 64) || !grow_dev(w.d_rbsp, n + 64))
                return fail(H264B_E_NOMEM, "out of memory staging the device's streams");
            for (size_t k = 0; k < order.size(); k++) {
                const TrimmedStream &t = ts[order[k]];
                memcpy((uint8_t *)w.h_stream.p + base[k], J.streams[t.index].stream + t.begin, (size_t)(t.end - t.begin));
            }
            // the device's slice rows, and their order by length (longest first; ties: the lower row)
            std::vector<uint32_t> rows;
            rows.reserve(n_sl);
            for (uint32_t k : order) {
                const h264b_batch_stream &b = J.streams[ts[k].index];
                for (uint32_t x = 0; x < b.n_slices; x++) rows.push_back(b.first_slice + x);
            }
            auto ops_of = [&](uint32_t row) { return J.n_ops ? std::min(J.n_ops[row], J.n_ops_max) : J.n_ops_max; };
            std::vector<uint32_t> perm(n_sl);
            for (uint32_t k = 0; k < n_sl; k++) perm[k] = k;
            std::stable_sort(perm.begin(), perm.end(), [&](uint32_t x, uint32_t y) { return ops_of(rows[x]) > ops_of(rows[y]); });
            // classes: more than 1/2, 1/8, 1/32, 1/128 of the longest slice, and the rest
            uint32_t cls_begin[kClasses + 1];
            {
                const uint64_t top = n_sl ? ops_of(rows[perm[0]]) : 0;
                const uint64_t thr[kClasses - 1] = {top / 2, top / 8, top / 32, top / 128};
                uint32_t k = 0;
                for (int c = 0; c < kClasses - 1; c++) {
                    cls_begin[c] = k;
                    while (k < n_sl && ops_of(rows[perm[k]]) > thr[c]) k++;
                }
                cls_begin[kClasses - 1] = k;
                cls_begin[kClasses] = n_sl;
            }
            const uint32_t nal_cap = (uint32_t)std::min<uint64_t>(n / 64 + 1024 + 2 * (uint64_t)order.size(), 0xFFFFFFF0ull);
            const size_t ms = n_sl ? n_sl : 1;
            // per-slice inputs in class order
            if (!grow_pin(w.h_small, ms * (4 + 4 + sizeof(h264b_slice_qp) + 8) + 8 + 256))
                return fail(H264B_E_NOMEM, "out of pinned memory");
            uint32_t *h_perm = (uint32_t *)w.h_small.p;
            uint32_t *h_nops = h_perm + ms;
            h264b_slice_qp *h_qp = (h264b_slice_qp *)(h_nops + ms);
            uint64_t *h_boff = (uint64_t *)(h_qp + ms);
            h_boff[0] = 0;
            for (uint32_t k = 0; k < n_sl; k++) {
                const uint32_t row = rows[perm[k]];
                h_perm[k] = perm[k];
                h_nops[k] = ops_of(row);
                h_qp[k] = J.qp[row];
                h_boff[k + 1] = h_boff[k] + ((uint64_t)h_nops[k] + 1 + 31) / 32;
            }
            const size_t total_words = (size_t)h_boff[n_sl];
            if (!grow_dev(w.d_nals, (size_t)nal_cap * sizeof(h264b_nal)) || !grow_dev(w.d_sum, 256) ||
                !grow_dev(w.d_off, ms * 8) || !grow_dev(w.d_len, ms * 4) || !grow_dev(w.d_snal, ms * 4 + 16) ||
                !grow_dev(w.d_offp, ms * 8) || !grow_dev(w.d_lenp, ms * 4) || !grow_dev(w.d_perm, ms * 4) ||
                !grow_dev(w.d_nops, ms * 4) || !grow_dev(w.d_qp, ms * sizeof(h264b_slice_qp)) ||
                !grow_dev(w.d_boff, (ms + 1) * 8) || !grow_dev(w.d_bins, total_words * 4 + 16) ||
                !grow_dev(w.d_fin, ms * sizeof(h264b_cabac_final)) || !grow_dev(w.d_ops, (size_t)J.n_ops_max * 2 + 16) ||
                !grow_pin(w.h_bins, total_words * 4 + 16) || !grow_pin(w.h_fin, ms * sizeof(h264b_cabac_final)) ||
                !grow_pin(w.h_nals, (size_t)nal_cap * sizeof(h264b_nal) + 256))
                return fail(H264B_E_NOMEM, "out of memory for the device's slice arrays");
            // ---- split + strip, slice list, class order (context 0's stream)
            h264b_ctx *c0 = w.ctx[0];
            cudaStream_t s0 = c0->stream;
            bool ok = cudaMemcpyAsync(w.d_stream.p, w.h_stream.p, n, cudaMemcpyHostToDevice, s0) == cudaSuccess;
            ok = ok && cudaMemcpyAsync(w.d_perm.p, h_perm, ms * 4, cudaMemcpyHostToDevice, s0) == cudaSuccess;
            ok = ok && cudaMemcpyAsync(w.d_nops.p, h_nops, ms * 4, cudaMemcpyHostToDevice, s0) == cudaSuccess;
            ok = ok && cudaMemcpyAsync(w.d_qp.p, h_qp, ms * sizeof(h264b_slice_qp), cudaMemcpyHostToDevice, s0) == cudaSuccess;
            ok = ok && cudaMemcpyAsync(w.d_boff.p, h_boff, (ms + 1) * 8, cudaMemcpyHostToDevice, s0) == cudaSuccess;
            if (J.n_ops_max)
                ok = ok && cudaMemcpyAsync(w.d_ops.p, J.ops, (size_t)J.n_ops_max * 2, cudaMemcpyHostToDevice, s0) == cudaSuccess;
            if (!ok) return fail(H264B_E_CUDA, "copying the device's share in failed");
            uint32_t *d_ns = (uint32_t *)((uint8_t *)w.d_sum.p + 64);
            int rc = h264b_annexb_scan_dev(c0, (const uint8_t *)w.d_stream.p, n, (uint8_t *)w.d_rbsp.p, (h264b_nal *)w.d_nals.p,
                                           nullptr, nal_cap, (h264b_scan_summary *)w.d_sum.p, 0);
            if (rc == H264B_OK)
                rc = h264b_slice_select_dev(c0, (const h264b_nal *)w.d_nals.p, (const h264b_scan_summary *)w.d_sum.p, nal_cap,
                                            J.slice_data_offset, n_sl, (uint64_t *)w.d_off.p, (uint32_t *)w.d_len.p,
                                            (uint32_t *)w.d_snal.p, d_ns);
            if (rc != H264B_OK) return fail(rc, std::string("split + strip: ") + h264b_last_error(c0));
            if (n_sl)
                gather_slices_kernel<<<(n_sl + 255) / 256, 256, 0, s0>>>((const uint64_t *)w.d_off.p, (const uint32_t *)w.d_len.p,
                                                                        (const uint32_t *)w.d_perm.p, n_sl, (uint64_t *)w.d_offp.p,
                                                                        (uint32_t *)w.d_lenp.p);
            cudaEventRecord(w.e_scan, s0);
            if (getenv("H264B_SCHED_TRACE")) {
                cudaStreamSynchronize(s0);
                fprintf(stderr, "h264b scheduler: device %d: %.1f MB staged, split + strip done at %.1f ms\n", w.device, n / 1e6,
                        ms_since(Clock::now()));
            }
            // ---- the classes' CABAC launches and their results, each on its own stream
            bool launched[kClasses] = {false, false, false, false, false};
            for (int c = 0; c < kClasses; c++) {
                const uint32_t k0 = cls_begin[c], k1 = cls_begin[c + 1];
                if (k1 == k0) continue;
                h264b_ctx *cc = w.ctx[c];
                cudaStream_t sc = cc->stream;
                cudaStreamWaitEvent(sc, w.e_scan, 0);
                h264b_cabac_job cj;
                memset(&cj, 0, sizeof(cj));
                cj.bytes = (const uint8_t *)w.d_rbsp.p;
                cj.total_bytes = n + 16;
                cj.off = (const uint64_t *)w.d_offp.p + k0;
                cj.len = (const uint32_t *)w.d_lenp.p + k0;
                cj.n_slices = k1 - k0;
                cj.n_ctx = J.n_ctx;
                cj.ops = (const uint16_t *)w.d_ops.p;
                cj.n_ops_max = J.n_ops_max;
                cj.n_ops = (const uint32_t *)w.d_nops.p + k0;
                cj.qp = (const h264b_slice_qp *)w.d_qp.p + k0;
                cj.bins = (uint32_t *)w.d_bins.p;
                cj.bins_off = (const uint64_t *)w.d_boff.p + k0;
                cj.final = (h264b_cabac_final *)w.d_fin.p + k0;
                cj.flags = J.flags & (H264B_TABLES_SPEC | H264B_BYPASS_SPEC_OR | H264B_CABAC_FINAL_TERMINATE);
                rc = h264b_cabac_decode_dev(cc, &cj);
                if (rc != H264B_OK) return fail(rc, std::string("cabac: ") + h264b_last_error(cc));
                const size_t w0 = (size_t)h_boff[k0], w1 = (size_t)h_boff[k1];
                cudaMemcpyAsync((uint32_t *)w.h_bins.p + w0, (const uint32_t *)w.d_bins.p + w0, (w1 - w0) * 4, cudaMemcpyDeviceToHost, sc);
                cudaMemcpyAsync((h264b_cabac_final *)w.h_fin.p + k0, (const h264b_cabac_final *)w.d_fin.p + k0,
                                (size_t)(k1 - k0) * sizeof(h264b_cabac_final), cudaMemcpyDeviceToHost, sc);
                cudaEventRecord(w.e_done[c], sc);
                launched[c] = true;
            }
            // the NAL index (context 0's stream, behind its class launch)
            cudaMemcpyAsync(w.h_nals.p, w.d_sum.p, 128, cudaMemcpyDeviceToHost, s0);
            cudaMemcpyAsync((uint8_t *)w.h_nals.p + 256, w.d_nals.p, (size_t)nal_cap * sizeof(h264b_nal), cudaMemcpyDeviceToHost, s0);
            // ---- collect the classes as they finish
            int pending = 0;
            for (int c = 0; c < kClasses; c++) pending += launched[c] ? 1 : 0;
            while (pending) {
                bool progressed = false;
                for (int c = 0; c < kClasses; c++) {
                    if (!launched[c]) continue;
                    const cudaError_t q = cudaEventQuery(w.e_done[c]);
                    if (q == cudaErrorNotReady) continue;
                    if (q != cudaSuccess) return fail(H264B_E_CUDA, std::string("a class launch failed: ") + cudaGetErrorString(q));
                    const double t_done = ms_since(Clock::now());
                    if (getenv("H264B_SCHED_TRACE"))
                        fprintf(stderr, "h264b scheduler: device %d class %d: %u slices (longest %u ops) done at %.1f ms\n", w.device,
                                c, cls_begin[c + 1] - cls_begin[c], ops_of(rows[perm[cls_begin[c]]]), t_done);
                    for (uint32_t k = cls_begin[c]; k < cls_begin[c + 1]; k++) {
                        const uint32_t row = rows[perm[k]];
                        s->fin[row] = ((const h264b_cabac_final *)w.h_fin.p)[k];
                        s->slice_done_ms[row] = t_done;
                        memcpy(s->bins.data() + s->bins_off[row], (const uint32_t *)w.h_bins.p + h_boff[k],
                               (size_t)(h_boff[k + 1] - h_boff[k]) * 4);
                    }
                    launched[c] = false;
                    pending--;
                    progressed = true;
                }
                if (!progressed) std::this_thread::sleep_for(std::chrono::microseconds(50));
            }
            if (cudaStreamSynchronize(s0) != cudaSuccess) return fail(H264B_E_CUDA, "the NAL index did not arrive");
            const h264b_scan_summary *sum = (const h264b_scan_summary *)w.h_nals.p;
            const uint32_t found = *(const uint32_t *)((const uint8_t *)w.h_nals.p + 64);
            if (sum->status != H264B_OK) return fail(H264B_E_CAPACITY, "more NAL units than the index holds");
            if (found != n_sl)
                return fail(H264B_E_INVALID, "the device's streams hold " + std::to_string(found) + " slice NAL units, the batch announces " +
                                                 std::to_string(n_sl));
            // NAL units: those that lie inside one stream's staged extent (the 4-byte unit that the next stream's leading
            // start code forms is nobody's)
            const h264b_nal *un = (const h264b_nal *)((const uint8_t *)w.h_nals.p + 256);
            size_t k = 0;
            for (uint64_t i = 0; i < sum->n_nals; i++) {
                const h264b_nal &u = un[i];
                while (k + 1 < order.size() && u.start >= base[k + 1]) k++;
                const TrimmedStream &t = ts[order[k]];
                const uint64_t lo = base[k], hi = base[k + 1];
                if (u.start < lo + 4 || u.start + u.num_bytes > hi) continue;
                h264b_nal v = u;
                v.start = u.start - lo + t.begin;
                v.rbsp_off = u.rbsp_off - lo + t.begin;
                s->stream_nals[t.index].push_back(v);
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
    res->bins = s->bins.data();
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

}  // extern "C"
