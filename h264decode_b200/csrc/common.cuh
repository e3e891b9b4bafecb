// common.cuh -- shared declarations of libh264b200 (CUDA, sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX 3: ranges show up in Nsight Systems / ncu timelines, no library to link

#include "../../include/h264b200.h"

struct StreamSlot;
// A captured Annex-B pass (annexb_scan.cu): small streams are launch bound (8 stream operations for 1 MB of input), a
// graph of the same operations replays with one launch.  Keyed by every argument of the pass.
struct ScanGraph {
    unsigned char key[96];
    cudaGraphExec_t exec;
    uint64_t used;     // tick of the last replay (least recently used goes first)
    unsigned nodes;    // kernels in the graph (for the launch counter)
};
constexpr int kScanGraphs = 8;
// Jobs in flight per context.  Three: in the steady state one job copies in, one runs its kernels and one copies out, so
// the slowest of the three stages paces the pipeline (with two, H2D + kernels + D2H of one job span two periods).
constexpr int kStreamSlots = 3;
struct h264b_ctx {
    int device;
    int sm_count;
    cudaStream_t own_stream;
    cudaStream_t stream;  // the one "_dev" work goes to (own_stream unless h264b_set_stream was called)
    char err[512];
    uint64_t launches;
    ScanGraph scan_graph[kScanGraphs];
    uint64_t scan_graph_tick;
    int cabac_max_warps;  // 0: as many warps per CTA as fit; else a cap (launches that share the GPU: h264b_scheduler)
    int cabac_pack;       // 1: 32 slices per warp whatever their number (launches that share the GPU: as few warps as the
                          // slices need, so that each has a scheduler to itself); 0: spread over the schedulers first
    int cabac_exclusive;  // 1: one slice per warp, four warps per CTA (one per scheduler), the whole SM's shared memory per
                          // CTA so that nothing else lands beside it: for the few slices a whole batch waits for

    // grow-only device scratch, in banks: [0] the direct ("_dev" and host-pointer) entry points, [1 + s] stream-job slot
    // s, whose kernels run on their own stream and may overlap the other slots'
    int bank;
    void *scan_scratch[1 + kStreamSlots];
    size_t scan_scratch_bytes[1 + kStreamSlots];

    // constant device tables, built once in h264b_create: [0] = REF, [1] = SPEC
    uint64_t *d_cabac_tab[2];   // 128 entries, see cabac_engine.cu
    uint8_t *d_range_lps[2];    // 256 bytes
    uint8_t *d_trans[2];        // 64 LPS + 64 MPS
    int16_t *d_mn[2];           // [5][1024] (m | n << 8) by idc class, see ctx_init.cu
    uint8_t *d_state_lut[2];    // [5][52][1024] state bytes

    // host-level (pinned) staging, grow-only
    void *h_pin[8];
    size_t h_pin_bytes[8];
    void *d_buf[1 + kStreamSlots][20];
    size_t d_buf_bytes[1 + kStreamSlots][20];

    // h264b_stream_submit / h264b_stream_wait: kStreamSlots jobs in flight, each with its own buffers, stream and events
    cudaStream_t s_in, s_out;  // copy streams (host -> device, device -> host)
    struct StreamSlot *slot[kStreamSlots];
    uint64_t next_ticket;
    cudaEvent_t t_ref;  // H264B_TRACE=1: time origin
    int trace;
};

enum { kSlotDev = 20, kSlotPin = 16 };
struct StreamSlot {
    void *d[kSlotDev];
    size_t d_bytes[kSlotDev];
    void *h[kSlotPin];
    size_t h_bytes[kSlotPin];
    cudaStream_t cs;     // the slot's compute stream: the tail of one job's CABAC kernel overlaps the next job's kernels
    cudaEvent_t e_in, e_compute, e_out;
    cudaEvent_t t_in0, t_c0, t_o0;  // H264B_TRACE=1: phase starts (timing events)
    uint64_t ticket;     // job occupying the slot
    bool busy;           // submitted, not yet waited for
    h264b_stream_job job;
    uint32_t nal_cap;
    size_t nal_prefix;   // NAL records copied out with the job (the rest, if any, is fetched by h264b_stream_wait)
    size_t total_words;
};

namespace h264b {

// a named range on the calling host thread for the lifetime of the object (tracing: SURVEY.md section 5)
struct TraceRange {
    explicit TraceRange(const char *name) { nvtxRangePushA(name); }
    ~TraceRange() { nvtxRangePop(); }
    TraceRange(const TraceRange &) = delete;
    TraceRange &operator=(const TraceRange &) = delete;
};

int set_error(h264b_ctx *ctx, int code, const char *fmt, ...);
int ensure_dev(h264b_ctx *ctx, int slot, size_t bytes, void **out);   // grow-only device buffer per slot
int ensure_pin(h264b_ctx *ctx, int slot, size_t bytes, void **out);   // grow-only pinned buffer per slot

#define H264B_CUDA(ctx, call)                                                                        \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return h264b::set_error((ctx), H264B_E_CUDA, "%s failed: %s (%s:%d)", #call,             \
                                    cudaGetErrorString(e_), __FILE__, __LINE__);                     \
    } while (0)

#define H264B_LAUNCH_CHECK(ctx, name)                                                                \
    do {                                                                                             \
        cudaError_t e_ = cudaGetLastError();                                                         \
        if (e_ != cudaSuccess)                                                                       \
            return h264b::set_error((ctx), H264B_E_CUDA, "launch of %s failed: %s", name,            \
                                    cudaGetErrorString(e_));                                         \
        (ctx)->launches++;                                                                           \
    } while (0)

// kernels' host-side launchers (each file owns its kernels)
int build_tables(h264b_ctx *ctx);  // tables.cu
int launch_annexb_scan(h264b_ctx *ctx, const uint8_t *d_stream, uint64_t n, uint8_t *d_rbsp, h264b_nal *d_nals,
                       h264b_nal_ext *d_ext, uint32_t nal_cap, h264b_scan_summary *d_summary, uint32_t flags);
int launch_nal_frames(h264b_ctx *ctx, const uint8_t *d_frames, uint64_t total, const uint64_t *d_off,
                      const uint32_t *d_len, uint32_t n_frames, h264b_nal *d_nals, h264b_nal_ext *d_ext,
                      uint8_t *d_rbsp);
int launch_ctx_init(h264b_ctx *ctx, const h264b_slice_qp *d_params, uint32_t n_slices, uint32_t n_ctx,
                    uint8_t *d_states, uint32_t flags);
int launch_cabac(h264b_ctx *ctx, const h264b_cabac_job *job, const uint32_t *d_n_slices = nullptr);
struct StreamParamSets {  // H264B_STREAM_PARAM_SETS: the stream's own parameter sets, device-resident (param_sets.cu)
    const h264b_sps *sps;
    const h264b_pps *pps;
    const uint32_t *sps_nal, *pps_nal, *counts;
    int32_t *slice_sps, *slice_pps;
    const h264b_sps *initial_sps;  // device copies of job.initial_sps / initial_pps, or NULL
    const h264b_pps *initial_pps;
};
int launch_stream_slice_headers(h264b_ctx *ctx, const h264b_param_sets *params, const uint8_t *d_rbsp, uint64_t total,
                                const h264b_nal *d_nals, const uint32_t *d_slice_nal, const uint32_t *d_n_slices,
                                uint32_t max_slices, h264b_slice_header *d_hdr, uint64_t *d_off, uint32_t *d_len,
                                h264b_slice_qp *d_qp, const StreamParamSets *sp = nullptr);
int launch_pset_select(h264b_ctx *ctx, const h264b_nal *d_nals, const h264b_scan_summary *d_summary, uint32_t nal_cap,
                       uint32_t max_sps, uint32_t max_pps, uint32_t *d_sps_nal, uint32_t *d_pps_nal, uint32_t *d_counts);
int launch_parse_sps(h264b_ctx *ctx, const uint8_t *d_bytes, uint64_t total, const uint64_t *d_off, const uint32_t *d_len,
                     const h264b_nal *d_nals, const uint32_t *d_nal_index, uint32_t n, const uint32_t *d_n,
                     h264b_sps *d_out);
int launch_parse_pps(h264b_ctx *ctx, const uint8_t *d_bytes, uint64_t total, const uint64_t *d_off, const uint32_t *d_len,
                     const h264b_nal *d_nals, const uint32_t *d_nal_index, uint32_t n, const uint32_t *d_n,
                     h264b_pps *d_out);
int launch_slice_select(h264b_ctx *ctx, const h264b_nal *d_nals, const h264b_scan_summary *d_summary,
                        uint32_t nal_cap, uint32_t slice_data_offset, uint32_t max_slices, uint64_t *d_off,
                        uint32_t *d_len, uint32_t *d_slice_nal, uint32_t *d_n_slices, uint32_t *d_n_found = nullptr);

}  // namespace h264b
