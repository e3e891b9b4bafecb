// batch.cu -- h264b_scheduler: many independent streams over the GPUs of one box, in one process.
//
// Replaces the reference's connection-level concurrency (main.go:16-21: one goroutine per accepted connection running
// ByteStreamReader -> handleConnection, h264/server.go:113-166) for a batch of streams that is known up front: one
// worker thread and one h264b context per device, streams dealt to devices longest first (LPT by bytes), grouped into
// device jobs with the longest slices first, three jobs in flight per device through the same h264b_stream_submit /
// h264b_stream_wait any single-stream caller uses.  Host code only: no kernel lives here.
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

struct DeviceJob {
    std::vector<uint32_t> streams;        // indices into the batch, staging order
    std::vector<uint64_t> base;           // offset of each stream in the staged buffer
    uint64_t bytes = 0;
    uint32_t n_slices = 0;
};

struct Worker {
    int device = 0;
    h264b_ctx *ctx = nullptr;
    uint8_t *stage[kStreamSlots] = {nullptr, nullptr, nullptr};
    size_t stage_bytes[kStreamSlots] = {0, 0, 0};
    std::vector<uint32_t> j_nops[kStreamSlots];
    std::vector<h264b_slice_qp> j_qp[kStreamSlots];
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
    h264b_scheduler *s = new h264b_scheduler;
    s->workers.resize(n_devices);
    for (uint32_t d = 0; d < n_devices; d++) {
        s->workers[d].device = devices[d];
        const int32_t rc = h264b_create(devices[d], &s->workers[d].ctx);
        if (rc != H264B_OK) {
            h264b_scheduler_destroy(s);
            return rc;
        }
    }
    *out = s;
    return H264B_OK;
}

void h264b_scheduler_destroy(h264b_scheduler *s) {
    if (!s) return;
    for (Worker &w : s->workers) {
        if (!w.ctx) continue;
        for (int k = 0; k < kStreamSlots; k++)
            if (w.stage[k]) h264b_host_free(w.ctx, w.stage[k]);
        h264b_destroy(w.ctx);
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
    // ---- per device: streams with the longest slices first, cut into jobs of ~group_bytes
    std::vector<std::vector<DeviceJob>> jobs(nd);
    for (uint32_t d = 0; d < nd; d++) {
        std::stable_sort(mine[d].begin(), mine[d].end(), [&](uint32_t a, uint32_t b) { return ts[a].longest > ts[b].longest; });
        DeviceJob cur;
        for (uint32_t k : mine[d]) {
            const uint64_t len = ts[k].end - ts[k].begin;
            if (!cur.streams.empty() && cur.bytes + len > group_bytes) {
                jobs[d].push_back(std::move(cur));
                cur = DeviceJob();
            }
            cur.base.push_back(cur.bytes);
            cur.streams.push_back(k);
            cur.bytes += len;
            cur.n_slices += J.streams[ts[k].index].n_slices;
            s->stream_job[ts[k].index] = (uint32_t)jobs[d].size();
        }
        if (!cur.streams.empty()) jobs[d].push_back(std::move(cur));
        s->device_jobs[d] = (uint32_t)jobs[d].size();
    }

    // ---- one worker thread per device
    const Clock::time_point t_start = Clock::now();
    auto ms_since = [&](Clock::time_point t) { return std::chrono::duration<double, std::milli>(t - t_start).count(); };
    std::vector<std::thread> threads;
    for (uint32_t d = 0; d < nd; d++) {
        threads.emplace_back([&, d]() {
            Worker &w = s->workers[d];
            w.rc = H264B_OK;
            w.err.clear();
            const std::vector<DeviceJob> &dj = jobs[d];
            auto fail = [&](int rc, const char *what) {
                w.rc = rc;
                w.err = std::string(what) + ": " + h264b_last_error(w.ctx);
            };
            uint64_t ticket[kStreamSlots] = {0, 0, 0};
            Clock::time_point first_submit;
            bool any = false;
            auto collect = [&](size_t q) -> bool {  // wait for device job q and scatter its results
                const DeviceJob &g = dj[q];
                h264b_stream_result r;
                const int32_t rc = h264b_stream_wait(w.ctx, ticket[q % kStreamSlots], &r);
                const double t_done = ms_since(Clock::now());
                if (rc != H264B_OK) {
                    fail(rc, "stream_wait");
                    return false;
                }
                if (r.n_slices != g.n_slices) {
                    w.rc = H264B_E_INVALID;
                    w.err = "a device job found " + std::to_string(r.n_slices) + " slice NAL units, its streams announce " +
                            std::to_string(g.n_slices);
                    return false;
                }
                // slices: job row -> batch row
                uint32_t row = 0;
                for (size_t k = 0; k < g.streams.size(); k++) {
                    const h264b_batch_stream &b = J.streams[ts[g.streams[k]].index];
                    for (uint32_t x = 0; x < b.n_slices; x++, row++) {
                        const uint32_t br = b.first_slice + x;
                        s->fin[br] = r.final[row];
                        s->slice_done_ms[br] = t_done;
                        const uint64_t words = r.bins_off[row + 1] - r.bins_off[row];
                        memcpy(s->bins.data() + s->bins_off[br], r.bins + r.bins_off[row], (size_t)words * 4);
                    }
                }
                // NAL units: those that lie inside one stream's staged extent (the 4-byte unit that the next stream's
                // leading start code forms is nobody's)
                size_t k = 0;
                for (uint64_t i = 0; i < r.scan.n_nals; i++) {
                    const h264b_nal &u = r.nals[i];
                    while (k + 1 < g.streams.size() && u.start >= g.base[k + 1]) k++;
                    const TrimmedStream &t = ts[g.streams[k]];
                    const uint64_t lo = g.base[k], hi = g.base[k] + (t.end - t.begin);
                    if (u.start < lo + 4 || u.start + u.num_bytes > hi) continue;
                    h264b_nal v = u;
                    v.start = u.start - lo + t.begin;
                    v.rbsp_off = u.rbsp_off - lo + t.begin;
                    s->stream_nals[t.index].push_back(v);
                }
                return true;
            };
            for (size_t q = 0; q < dj.size(); q++) {
                if (q >= (size_t)kStreamSlots && !collect(q - kStreamSlots)) return;
                const DeviceJob &g = dj[q];
                const int slot = (int)(q % kStreamSlots);
                if (w.stage_bytes[slot] < g.bytes + 64) {
                    if (w.stage[slot]) h264b_host_free(w.ctx, w.stage[slot]);
                    w.stage[slot] = nullptr;
                    void *p = nullptr;
                    const size_t want = (size_t)(g.bytes + 64) + (size_t)(g.bytes / 8);
                    if (h264b_host_alloc(w.ctx, want, &p) != H264B_OK) return fail(H264B_E_NOMEM, "host_alloc");
                    w.stage[slot] = (uint8_t *)p;
                    w.stage_bytes[slot] = want;
                }
                w.j_nops[slot].clear();
                w.j_qp[slot].clear();
                for (size_t k = 0; k < g.streams.size(); k++) {
                    const TrimmedStream &t = ts[g.streams[k]];
                    const h264b_batch_stream &b = J.streams[t.index];
                    memcpy(w.stage[slot] + g.base[k], b.stream + t.begin, (size_t)(t.end - t.begin));
                    for (uint32_t x = 0; x < b.n_slices; x++) {
                        if (J.n_ops) w.j_nops[slot].push_back(J.n_ops[b.first_slice + x]);
                        w.j_qp[slot].push_back(J.qp[b.first_slice + x]);
                    }
                }
                h264b_stream_job sj;
                memset(&sj, 0, sizeof(sj));
                sj.stream = w.stage[slot];
                sj.n = g.bytes;
                sj.slice_data_offset = J.slice_data_offset;
                sj.n_ctx = J.n_ctx;
                sj.ops = J.ops;
                sj.n_ops_max = J.n_ops_max;
                sj.n_ops = J.n_ops ? w.j_nops[slot].data() : nullptr;
                sj.qp = w.j_qp[slot].data();
                sj.max_slices = g.n_slices;
                sj.flags = J.flags & (H264B_TABLES_SPEC | H264B_BYPASS_SPEC_OR | H264B_CABAC_FINAL_TERMINATE);
                if (!any) {
                    first_submit = Clock::now();
                    any = true;
                }
                if (g.n_slices == 0) sj.qp = nullptr;
                const int32_t rc = h264b_stream_submit(w.ctx, &sj, &ticket[slot]);
                if (rc != H264B_OK) return fail(rc, "stream_submit");
            }
            for (size_t q = dj.size() > (size_t)kStreamSlots ? dj.size() - kStreamSlots : 0; q < dj.size(); q++)
                if (!collect(q)) return;
            if (any) s->device_busy_ms[d] = std::chrono::duration<double, std::milli>(Clock::now() - first_submit).count();
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
