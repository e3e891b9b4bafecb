// api.cu -- context management and the host-buffer entry points of libh264b200 (see include/h264b200.h).
// Everything that computes runs in the kernels of annexb_scan.cu / cabac_engine.cu / ctx_init.cu; this file only
// moves bytes between host and device and sequences launches.  There is no CPU implementation of the path here.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace h264b {

int set_error(h264b_ctx *ctx, int code, const char *fmt, ...) {
    if (ctx) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

int ensure_dev(h264b_ctx *ctx, int slot, size_t bytes, void **out) {
    if (bytes < 256) bytes = 256;
    void *&buf = ctx->d_buf[ctx->bank][slot];
    size_t &have = ctx->d_buf_bytes[ctx->bank][slot];
    if (have < bytes) {
        if (buf) {
            H264B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            cudaFree(buf);
            buf = nullptr;
            have = 0;
        }
        const size_t want = bytes + bytes / 8;
        cudaError_t e = cudaMalloc(&buf, want);
        if (e != cudaSuccess) return set_error(ctx, H264B_E_NOMEM, "cudaMalloc(%zu): %s", want, cudaGetErrorString(e));
        have = want;
    }
    *out = buf;
    return H264B_OK;
}

int ensure_pin(h264b_ctx *ctx, int slot, size_t bytes, void **out) {
    if (bytes < 256) bytes = 256;
    if (ctx->h_pin_bytes[slot] < bytes) {
        if (ctx->h_pin[slot]) {
            H264B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            cudaFreeHost(ctx->h_pin[slot]);
            ctx->h_pin[slot] = nullptr;
            ctx->h_pin_bytes[slot] = 0;
        }
        const size_t want = bytes + bytes / 8;
        cudaError_t e = cudaHostAlloc(&ctx->h_pin[slot], want, cudaHostAllocDefault);
        if (e != cudaSuccess)
            return set_error(ctx, H264B_E_NOMEM, "cudaHostAlloc(%zu): %s", want, cudaGetErrorString(e));
        ctx->h_pin_bytes[slot] = want;
    }
    *out = ctx->h_pin[slot];
    return H264B_OK;
}

}  // namespace h264b

using namespace h264b;

#define CHECK_CTX(ctx)                      \
    do {                                    \
        if (!(ctx)) return H264B_E_INVALID; \
        cudaSetDevice((ctx)->device);       \
    } while (0)
#define RC(expr)                  \
    do {                          \
        int rc_ = (expr);         \
        if (rc_) return rc_;      \
    } while (0)

extern "C" {

int32_t h264b_version(void) { return H264B_VERSION; }

int32_t h264b_device_count(int32_t *count) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (count) *count = (e == cudaSuccess) ? n : 0;
    return (e == cudaSuccess && n > 0) ? H264B_OK : H264B_E_NO_DEVICE;
}

int32_t h264b_create(int32_t device, h264b_ctx **out) {
    if (!out) return H264B_E_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return H264B_E_NO_DEVICE;
    if (cudaSetDevice(device) != cudaSuccess) return H264B_E_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return H264B_E_CUDA;
    if (prop.major < 10) return H264B_E_NO_DEVICE;  // sm_100a code only
    h264b_ctx *ctx = (h264b_ctx *)calloc(1, sizeof(h264b_ctx));
    if (!ctx) return H264B_E_NOMEM;
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        free(ctx);
        return H264B_E_CUDA;
    }
    ctx->stream = ctx->own_stream;
    bool ok = cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < kStreamSlots && ok; i++) {
        StreamSlot *sl = (StreamSlot *)calloc(1, sizeof(StreamSlot));
        ok = sl && cudaStreamCreateWithFlags(&sl->cs, cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreate(&sl->e_in) == cudaSuccess && cudaEventCreate(&sl->e_compute) == cudaSuccess &&
             cudaEventCreate(&sl->e_out) == cudaSuccess && cudaEventCreate(&sl->t_in0) == cudaSuccess &&
             cudaEventCreate(&sl->t_c0) == cudaSuccess && cudaEventCreate(&sl->t_o0) == cudaSuccess;
        ctx->slot[i] = sl;
    }
    ctx->trace = getenv("H264B_TRACE") != nullptr;
    ok = ok && cudaEventCreate(&ctx->t_ref) == cudaSuccess;
    if (!ok) {
        h264b_destroy(ctx);
        return H264B_E_CUDA;
    }
    int rc = build_tables(ctx);
    if (rc) {
        fprintf(stderr, "h264b_create: %s\n", ctx->err);
        h264b_destroy(ctx);
        return rc;
    }
    ctx->launches = 0;  // table construction is not counted
    *out = ctx;
    return H264B_OK;
}

void h264b_destroy(h264b_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (int v = 0; v < 2; v++) {
        cudaFree(ctx->d_cabac_tab[v]);
        cudaFree(ctx->d_range_lps[v]);
        cudaFree(ctx->d_trans[v]);
        cudaFree(ctx->d_mn[v]);
        cudaFree(ctx->d_state_lut[v]);
    }
    for (int b = 0; b < 1 + kStreamSlots; b++) {
        for (int i = 0; i < 20; i++) cudaFree(ctx->d_buf[b][i]);
        cudaFree(ctx->scan_scratch[b]);
    }
    for (int i = 0; i < kScanGraphs; i++)
        if (ctx->scan_graph[i].exec) cudaGraphExecDestroy(ctx->scan_graph[i].exec);
    for (int i = 0; i < 8; i++)
        if (ctx->h_pin[i]) cudaFreeHost(ctx->h_pin[i]);
    for (int i = 0; i < kStreamSlots; i++) {
        StreamSlot *sl = ctx->slot[i];
        if (!sl) continue;
        for (int k = 0; k < kSlotDev; k++) cudaFree(sl->d[k]);
        for (int k = 0; k < kSlotPin; k++)
            if (sl->h[k]) cudaFreeHost(sl->h[k]);
        if (sl->cs) cudaStreamDestroy(sl->cs);
        if (sl->e_in) cudaEventDestroy(sl->e_in);
        if (sl->e_compute) cudaEventDestroy(sl->e_compute);
        if (sl->e_out) cudaEventDestroy(sl->e_out);
        if (sl->t_in0) cudaEventDestroy(sl->t_in0);
        if (sl->t_c0) cudaEventDestroy(sl->t_c0);
        if (sl->t_o0) cudaEventDestroy(sl->t_o0);
        free(sl);
    }
    if (ctx->t_ref) cudaEventDestroy(ctx->t_ref);
    if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
    if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    free(ctx);
}

const char *h264b_last_error(const h264b_ctx *ctx) { return ctx ? ctx->err : "null context"; }

int32_t h264b_set_stream(h264b_ctx *ctx, void *cuda_stream) {
    CHECK_CTX(ctx);
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return H264B_OK;
}

int32_t h264b_sync(h264b_ctx *ctx) {
    CHECK_CTX(ctx);
    H264B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return H264B_OK;
}

int32_t h264b_host_alloc(h264b_ctx *ctx, size_t bytes, void **out) {
    CHECK_CTX(ctx);
    if (!out) return H264B_E_INVALID;
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) return set_error(ctx, H264B_E_NOMEM, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e));
    return H264B_OK;
}
int32_t h264b_host_free(h264b_ctx *ctx, void *p) {
    CHECK_CTX(ctx);
    if (p) H264B_CUDA(ctx, cudaFreeHost(p));
    return H264B_OK;
}
int32_t h264b_cut_byte_ranges(const uint8_t *stream, uint64_t n, uint32_t n_ranges, uint64_t *begin, uint64_t *end) {
    if (!n_ranges || !begin || !end || (n && !stream)) return H264B_E_INVALID;
    uint64_t cut = 0;  // begin of the range being closed
    for (uint32_t k = 1; k <= n_ranges; k++) {
        uint64_t next = n;
        if (k < n_ranges) {
            const uint64_t nominal = (uint64_t)(((unsigned __int128)k * n) / n_ranges);
            uint64_t p = nominal >= 3 ? nominal - 3 : 0;  // a start code that straddles the nominal cut counts
            if (p < cut) p = cut;
            for (; p + 4 <= n; p++)
                if (stream[p] == 0 && stream[p + 1] == 0 && stream[p + 2] == 0 && stream[p + 3] == 1) break;
            next = p + 4 <= n ? p : n;
        }
        begin[k - 1] = cut;
        end[k - 1] = k < n_ranges ? (next + 4 < n ? next + 4 : n) : n;
        cut = next;
    }
    return H264B_OK;
}
int32_t h264b_dev_alloc(h264b_ctx *ctx, size_t bytes, void **out) {
    CHECK_CTX(ctx);
    if (!out) return H264B_E_INVALID;
    cudaError_t e = cudaMalloc(out, bytes ? bytes : 1);
    if (e != cudaSuccess) return set_error(ctx, H264B_E_NOMEM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
    return H264B_OK;
}
int32_t h264b_dev_free(h264b_ctx *ctx, void *p) {
    CHECK_CTX(ctx);
    if (p) H264B_CUDA(ctx, cudaFree(p));
    return H264B_OK;
}
int32_t h264b_memcpy_h2d(h264b_ctx *ctx, void *dst, const void *src, size_t bytes) {
    CHECK_CTX(ctx);
    if (bytes) H264B_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return H264B_OK;
}
int32_t h264b_memcpy_d2h(h264b_ctx *ctx, void *dst, const void *src, size_t bytes) {
    CHECK_CTX(ctx);
    if (bytes) H264B_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return H264B_OK;
}
int32_t h264b_launch_count(const h264b_ctx *ctx, uint64_t *count) {
    if (!ctx || !count) return H264B_E_INVALID;
    *count = ctx->launches;
    return H264B_OK;
}

// ------------------------------------------------------------------------------------------------ "_dev" entries
int32_t h264b_annexb_scan_dev(h264b_ctx *ctx, const uint8_t *d_stream, uint64_t n, uint8_t *d_rbsp, h264b_nal *d_nals,
                              h264b_nal_ext *d_ext, uint32_t nal_cap, h264b_scan_summary *d_summary, uint32_t flags) {
    CHECK_CTX(ctx);
    if (!d_rbsp || !d_nals || !d_summary || (!d_stream && n)) return set_error(ctx, H264B_E_INVALID, "null pointer");
    return launch_annexb_scan(ctx, d_stream, n, d_rbsp, d_nals, d_ext, nal_cap, d_summary, flags);
}

int32_t h264b_ctx_init_dev(h264b_ctx *ctx, const h264b_slice_qp *d_params, uint32_t n_slices, uint32_t n_ctx,
                           uint8_t *d_states, uint32_t flags) {
    CHECK_CTX(ctx);
    if (n_slices && (!d_params || !d_states)) return set_error(ctx, H264B_E_INVALID, "null pointer");
    return launch_ctx_init(ctx, d_params, n_slices, n_ctx, d_states, flags);
}

int32_t h264b_slice_select_dev(h264b_ctx *ctx, const h264b_nal *d_nals, const h264b_scan_summary *d_summary,
                               uint32_t nal_cap, uint32_t slice_data_offset, uint32_t max_slices, uint64_t *d_off,
                               uint32_t *d_len, uint32_t *d_slice_nal, uint32_t *d_n_slices) {
    CHECK_CTX(ctx);
    if (!d_nals || !d_summary || !d_off || !d_len || !d_slice_nal || !d_n_slices)
        return set_error(ctx, H264B_E_INVALID, "null pointer");
    return launch_slice_select(ctx, d_nals, d_summary, nal_cap, slice_data_offset, max_slices, d_off, d_len,
                               d_slice_nal, d_n_slices);
}

int32_t h264b_cabac_decode_dev(h264b_ctx *ctx, const h264b_cabac_job *job) {
    CHECK_CTX(ctx);
    if (!job) return H264B_E_INVALID;
    return launch_cabac(ctx, job);
}

// ------------------------------------------------------------------------------------------------ host entries
// device slots: 0 stream/frames/bytes  1 rbsp  2 nals  3 ext  4 summary+counters  5 off  6 len  7 ops  8 n_ops
//               9 qp  10 init/final states  11 bins  12 final  13 slice_nal  14 states(K4)  15 scalar io
//               16 length-bundle sort (cabac_engine.cu)
// pinned slots: 0 nals  1 ext  2 rbsp  3 bins  4 final  5 slice_nal  6 misc
static uint32_t default_nal_cap(uint64_t n) {
    uint64_t c = n / 64 + 1024;
    return (uint32_t)(c > 0x7FFFFFF0ull ? 0x7FFFFFF0ull : c);
}

static int scan_host_common(h264b_ctx *ctx, const uint8_t *stream, uint64_t n, uint32_t flags, uint32_t *cap_io,
                            h264b_scan_summary *summary, uint8_t **d_rbsp_out, h264b_nal **d_nals_out,
                            h264b_nal_ext **d_ext_out, h264b_scan_summary **d_sum_out, bool want_ext) {
    void *d_stream, *d_rbsp, *d_nals, *d_ext = nullptr, *d_sum;
    RC(ensure_dev(ctx, 0, n + 64, &d_stream));
    RC(ensure_dev(ctx, 1, n + 64, &d_rbsp));
    RC(ensure_dev(ctx, 4, 256, &d_sum));
    if (n) H264B_CUDA(ctx, cudaMemcpyAsync(d_stream, stream, n, cudaMemcpyHostToDevice, ctx->stream));
    uint32_t cap = *cap_io;
    for (int attempt = 0; attempt < 2; attempt++) {
        RC(ensure_dev(ctx, 2, (size_t)cap * sizeof(h264b_nal), &d_nals));
        if (want_ext) RC(ensure_dev(ctx, 3, (size_t)cap * sizeof(h264b_nal_ext), &d_ext));
        RC(launch_annexb_scan(ctx, (const uint8_t *)d_stream, n, (uint8_t *)d_rbsp, (h264b_nal *)d_nals,
                              (h264b_nal_ext *)d_ext, cap, (h264b_scan_summary *)d_sum, flags));
        H264B_CUDA(ctx, cudaMemcpyAsync(summary, d_sum, sizeof(*summary), cudaMemcpyDeviceToHost, ctx->stream));
        H264B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (summary->status != H264B_E_CAPACITY) break;
        if (summary->n_start_codes + 1 > 0x7FFFFFF0ull) return set_error(ctx, H264B_E_CAPACITY, "too many NAL units");
        cap = (uint32_t)(summary->n_start_codes + 1);
    }
    if (summary->status != H264B_OK) return set_error(ctx, H264B_E_CAPACITY, "NAL index capacity");
    *cap_io = cap;
    *d_rbsp_out = (uint8_t *)d_rbsp;
    *d_nals_out = (h264b_nal *)d_nals;
    *d_ext_out = (h264b_nal_ext *)d_ext;
    *d_sum_out = (h264b_scan_summary *)d_sum;
    return H264B_OK;
}

int32_t h264b_annexb_scan(h264b_ctx *ctx, const uint8_t *stream, uint64_t n, uint32_t flags, int32_t want_rbsp,
                          const h264b_nal **nals, const h264b_nal_ext **ext, h264b_scan_summary *summary,
                          const uint8_t **rbsp, const uint8_t **d_rbsp_out) {
    CHECK_CTX(ctx);
    if (!summary || (!stream && n)) return set_error(ctx, H264B_E_INVALID, "null pointer");
    uint32_t cap = default_nal_cap(n);
    uint8_t *d_rbsp;
    h264b_nal *d_nals;
    h264b_nal_ext *d_ext;
    h264b_scan_summary *d_sum;
    RC(scan_host_common(ctx, stream, n, flags, &cap, summary, &d_rbsp, &d_nals, &d_ext, &d_sum, ext != nullptr));
    void *h_nals, *h_ext = nullptr, *h_rbsp = nullptr;
    const size_t nn = (size_t)summary->n_nals;
    RC(ensure_pin(ctx, 0, nn * sizeof(h264b_nal), &h_nals));
    if (nn) H264B_CUDA(ctx, cudaMemcpyAsync(h_nals, d_nals, nn * sizeof(h264b_nal), cudaMemcpyDeviceToHost, ctx->stream));
    if (ext) {
        RC(ensure_pin(ctx, 1, nn * sizeof(h264b_nal_ext), &h_ext));
        if (nn)
            H264B_CUDA(ctx, cudaMemcpyAsync(h_ext, d_ext, nn * sizeof(h264b_nal_ext), cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (want_rbsp) {  // position-preserving layout: the buffer is as long as the stream
        RC(ensure_pin(ctx, 2, (size_t)n, &h_rbsp));
        if (n && nn) H264B_CUDA(ctx, cudaMemcpyAsync(h_rbsp, d_rbsp, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    }
    H264B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (nals) *nals = (const h264b_nal *)h_nals;
    if (ext) *ext = (const h264b_nal_ext *)h_ext;
    if (rbsp) *rbsp = (const uint8_t *)h_rbsp;
    if (d_rbsp_out) *d_rbsp_out = d_rbsp;
    return H264B_OK;
}

int32_t h264b_nal_units(h264b_ctx *ctx, const uint8_t *frames, uint64_t total_bytes, const uint64_t *frame_off,
                        const uint32_t *frame_len, uint32_t n_frames, uint32_t flags, h264b_nal *nals,
                        h264b_nal_ext *ext, uint8_t *rbsp) {
    (void)flags;
    CHECK_CTX(ctx);
    if (!n_frames) return H264B_OK;
    if (!frame_off || !frame_len || !nals || !rbsp || (!frames && total_bytes))
        return set_error(ctx, H264B_E_INVALID, "null pointer");
    for (uint32_t i = 0; i < n_frames; i++)
        if (frame_off[i] + frame_len[i] > total_bytes) return set_error(ctx, H264B_E_INVALID, "frame %u out of range", i);
    void *d_in, *d_out, *d_nals, *d_ext = nullptr, *d_off, *d_len;
    RC(ensure_dev(ctx, 0, total_bytes + 64, &d_in));
    RC(ensure_dev(ctx, 1, total_bytes + 64, &d_out));
    RC(ensure_dev(ctx, 2, (size_t)n_frames * sizeof(h264b_nal), &d_nals));
    if (ext) RC(ensure_dev(ctx, 3, (size_t)n_frames * sizeof(h264b_nal_ext), &d_ext));
    RC(ensure_dev(ctx, 5, (size_t)n_frames * 8, &d_off));
    RC(ensure_dev(ctx, 6, (size_t)n_frames * 4, &d_len));
    if (total_bytes) H264B_CUDA(ctx, cudaMemcpyAsync(d_in, frames, total_bytes, cudaMemcpyHostToDevice, ctx->stream));
    H264B_CUDA(ctx, cudaMemcpyAsync(d_off, frame_off, (size_t)n_frames * 8, cudaMemcpyHostToDevice, ctx->stream));
    H264B_CUDA(ctx, cudaMemcpyAsync(d_len, frame_len, (size_t)n_frames * 4, cudaMemcpyHostToDevice, ctx->stream));
    RC(launch_nal_frames(ctx, (const uint8_t *)d_in, total_bytes, (const uint64_t *)d_off, (const uint32_t *)d_len,
                         n_frames, (h264b_nal *)d_nals, (h264b_nal_ext *)d_ext, (uint8_t *)d_out));
    H264B_CUDA(ctx, cudaMemcpyAsync(nals, d_nals, (size_t)n_frames * sizeof(h264b_nal), cudaMemcpyDeviceToHost, ctx->stream));
    if (ext)
        H264B_CUDA(ctx, cudaMemcpyAsync(ext, d_ext, (size_t)n_frames * sizeof(h264b_nal_ext), cudaMemcpyDeviceToHost, ctx->stream));
    if (total_bytes) H264B_CUDA(ctx, cudaMemcpyAsync(rbsp, d_out, total_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    H264B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return H264B_OK;
}

int32_t h264b_ctx_init(h264b_ctx *ctx, const h264b_slice_qp *params, uint32_t n_slices, uint32_t n_ctx,
                       uint8_t *states, uint32_t flags) {
    CHECK_CTX(ctx);
    if (!n_slices) return H264B_OK;
    if (!params || !states) return set_error(ctx, H264B_E_INVALID, "null pointer");
    void *d_p, *d_s;
    RC(ensure_dev(ctx, 9, (size_t)n_slices * sizeof(h264b_slice_qp), &d_p));
    RC(ensure_dev(ctx, 14, (size_t)n_slices * n_ctx, &d_s));
    H264B_CUDA(ctx, cudaMemcpyAsync(d_p, params, (size_t)n_slices * sizeof(h264b_slice_qp), cudaMemcpyHostToDevice, ctx->stream));
    RC(launch_ctx_init(ctx, (const h264b_slice_qp *)d_p, n_slices, n_ctx, (uint8_t *)d_s, flags));
    H264B_CUDA(ctx, cudaMemcpyAsync(states, d_s, (size_t)n_slices * n_ctx, cudaMemcpyDeviceToHost, ctx->stream));
    H264B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return H264B_OK;
}

int32_t h264b_cabac_decode(h264b_ctx *ctx, const h264b_cabac_job *job) {
    CHECK_CTX(ctx);
    if (!job) return H264B_E_INVALID;
    const h264b_cabac_job &j = *job;
    if (!j.n_slices) return H264B_OK;
    if (!j.bytes || !j.off || !j.len || !j.bins || !j.final || (!j.ops && j.n_ops_max) || (!j.qp && !j.init_states))
        return set_error(ctx, H264B_E_INVALID, "cabac: null pointer in job");
    for (uint32_t s = 0; s < j.n_slices; s++)
        if (j.off[s] + j.len[s] > j.total_bytes) return set_error(ctx, H264B_E_INVALID, "slice %u out of range", s);
    const size_t ns = j.n_slices;
    void *d_bytes, *d_off, *d_len, *d_ops, *d_nops = nullptr, *d_qp = nullptr, *d_init = nullptr, *d_bins, *d_fin,
         *d_fst = nullptr;
    RC(ensure_dev(ctx, 0, j.total_bytes + 64, &d_bytes));
    RC(ensure_dev(ctx, 5, ns * 8, &d_off));
    RC(ensure_dev(ctx, 6, ns * 4, &d_len));
    RC(ensure_dev(ctx, 7, (size_t)j.n_ops_max * 2 + 16, &d_ops));
    if (j.bins_off) return set_error(ctx, H264B_E_INVALID, "cabac: bins_off is for the _dev entry point");
    RC(ensure_dev(ctx, 11, ns * j.bins_stride_words * 4, &d_bins));
    RC(ensure_dev(ctx, 12, ns * sizeof(h264b_cabac_final), &d_fin));
    H264B_CUDA(ctx, cudaMemcpyAsync(d_bytes, j.bytes, j.total_bytes, cudaMemcpyHostToDevice, ctx->stream));
    H264B_CUDA(ctx, cudaMemcpyAsync(d_off, j.off, ns * 8, cudaMemcpyHostToDevice, ctx->stream));
    H264B_CUDA(ctx, cudaMemcpyAsync(d_len, j.len, ns * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (j.n_ops_max)
        H264B_CUDA(ctx, cudaMemcpyAsync(d_ops, j.ops, (size_t)j.n_ops_max * 2, cudaMemcpyHostToDevice, ctx->stream));
    if (j.n_ops) {
        RC(ensure_dev(ctx, 8, ns * 4, &d_nops));
        H264B_CUDA(ctx, cudaMemcpyAsync(d_nops, j.n_ops, ns * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (j.qp) {
        RC(ensure_dev(ctx, 9, ns * sizeof(h264b_slice_qp), &d_qp));
        H264B_CUDA(ctx, cudaMemcpyAsync(d_qp, j.qp, ns * sizeof(h264b_slice_qp), cudaMemcpyHostToDevice, ctx->stream));
    }
    const size_t st_bytes = ns * j.n_ctx;
    if (j.init_states || j.final_states) {
        void *d_st;
        RC(ensure_dev(ctx, 10, 2 * st_bytes, &d_st));
        if (j.init_states) {
            d_init = d_st;
            H264B_CUDA(ctx, cudaMemcpyAsync(d_init, j.init_states, st_bytes, cudaMemcpyHostToDevice, ctx->stream));
        }
        if (j.final_states) d_fst = (uint8_t *)d_st + st_bytes;
    }
    h264b_cabac_job dj = j;
    dj.bytes = (const uint8_t *)d_bytes;
    dj.off = (const uint64_t *)d_off;
    dj.len = (const uint32_t *)d_len;
    dj.ops = (const uint16_t *)d_ops;
    dj.n_ops = (const uint32_t *)d_nops;
    dj.qp = (const h264b_slice_qp *)d_qp;
    dj.init_states = (const uint8_t *)d_init;
    dj.bins = (uint32_t *)d_bins;
    dj.final = (h264b_cabac_final *)d_fin;
    dj.final_states = (uint8_t *)d_fst;
    RC(launch_cabac(ctx, &dj));
    H264B_CUDA(ctx, cudaMemcpyAsync(j.bins, d_bins, ns * j.bins_stride_words * 4, cudaMemcpyDeviceToHost, ctx->stream));
    H264B_CUDA(ctx, cudaMemcpyAsync(j.final, d_fin, ns * sizeof(h264b_cabac_final), cudaMemcpyDeviceToHost, ctx->stream));
    if (j.final_states)
        H264B_CUDA(ctx, cudaMemcpyAsync(j.final_states, d_fst, st_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    H264B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return H264B_OK;
}

// ---- the whole front end of one stream, asynchronously: kStreamSlots jobs in flight per context -------------------------
// device buffers of a slot: 0 stream  1 rbsp  2 nals  3 summary + slice count  4 off  5 len  6 slice_nal  7 ops
//                           8 n_ops  9 qp  10 bins_off  11 bins  12 final  13 ext  14 slice headers
// pinned buffers of a slot: 0 nals  1 bins_off  2 bins  3 final  4 slice_nal  5 summary + slice count
//                           9 rbsp  10 ext (H264B_STREAM_WANT_RBSP)  11 slice headers
//                           6 ops  7 n_ops  8 qp (staging of the caller's small arrays: they may be pageable, and a
//                           pageable source would make the copies -- and with them the whole submit -- synchronous)
static int slot_dev(h264b_ctx *ctx, StreamSlot *sl, int i, size_t bytes, void **out) {
    if (bytes < 256) bytes = 256;
    if (sl->d_bytes[i] < bytes) {
        if (sl->d[i]) {
            H264B_CUDA(ctx, cudaDeviceSynchronize());
            cudaFree(sl->d[i]);
            sl->d[i] = nullptr;
            sl->d_bytes[i] = 0;
        }
        const size_t want = bytes + bytes / 16;
        cudaError_t e = cudaMalloc(&sl->d[i], want);
        if (e != cudaSuccess) return set_error(ctx, H264B_E_NOMEM, "cudaMalloc(%zu): %s", want, cudaGetErrorString(e));
        sl->d_bytes[i] = want;
    }
    *out = sl->d[i];
    return H264B_OK;
}
static int slot_pin(h264b_ctx *ctx, StreamSlot *sl, int i, size_t bytes, void **out) {
    if (bytes < 256) bytes = 256;
    if (sl->h_bytes[i] < bytes) {
        if (sl->h[i]) {
            H264B_CUDA(ctx, cudaDeviceSynchronize());
            cudaFreeHost(sl->h[i]);
            sl->h[i] = nullptr;
            sl->h_bytes[i] = 0;
        }
        const size_t want = bytes + bytes / 16;
        cudaError_t e = cudaHostAlloc(&sl->h[i], want, cudaHostAllocDefault);
        if (e != cudaSuccess) return set_error(ctx, H264B_E_NOMEM, "cudaHostAlloc(%zu): %s", want, cudaGetErrorString(e));
        sl->h_bytes[i] = want;
    }
    *out = sl->h[i];
    return H264B_OK;
}

static int32_t stream_submit_on_slot(h264b_ctx *ctx, const h264b_stream_job *job, int slot_idx, uint32_t nal_cap);

int32_t h264b_stream_submit(h264b_ctx *ctx, const h264b_stream_job *job, uint64_t *ticket) {
    CHECK_CTX(ctx);
    if (!job || !ticket) return H264B_E_INVALID;
    // the job's kernels go to its slot's own stream and scratch bank (restored whatever happens)
    const int slot_idx = (int)(ctx->next_ticket % kStreamSlots);
    StreamSlot *sl = ctx->slot[slot_idx];
    if (sl->busy)
        return set_error(ctx, H264B_E_INVALID, "stream_submit: %d jobs are in flight, wait for one first", kStreamSlots);
    const cudaStream_t saved_stream = ctx->stream;
    const int saved_bank = ctx->bank;
    ctx->stream = sl->cs;
    ctx->bank = 1 + slot_idx;
    const int32_t rc = stream_submit_on_slot(ctx, job, slot_idx, default_nal_cap(job->n));
    ctx->stream = saved_stream;
    ctx->bank = saved_bank;
    if (rc == H264B_OK) {
        sl->busy = true;
        sl->ticket = ctx->next_ticket;
        *ticket = ctx->next_ticket++;
    }
    return rc;
}

// Enqueues the job on slot slot_idx (its copies and kernels; ctx->stream / ctx->bank are the slot's).  nal_cap: records of
// the NAL index.  h264b_stream_wait calls this again, with the bounds the first run reported, when one of them was too
// small (the job's host buffers are still valid then: they have to be until the wait returns).
static int32_t stream_submit_on_slot(h264b_ctx *ctx, const h264b_stream_job *job, int slot_idx, uint32_t nal_cap) {
    TraceRange trace_range("h264b:stream_submit");
    const h264b_stream_job &j = *job;
    // (H264B_STREAM_PARAM_SETS implies H264B_STREAM_SLICE_HEADERS)
    const bool from_headers = (j.flags & (H264B_STREAM_SLICE_HEADERS | H264B_STREAM_PARAM_SETS)) != 0 && j.max_slices != 0;
    const bool own_psets = from_headers && (j.flags & H264B_STREAM_PARAM_SETS) != 0;
    const uint32_t max_sps = j.max_sps ? j.max_sps : 64u, max_pps = j.max_pps ? j.max_pps : 64u;
    if ((!j.stream && j.n) || (j.max_slices && ((!j.qp && !from_headers) || (!j.ops && j.n_ops_max))) ||
        (from_headers && !own_psets && !j.param_sets))
        return set_error(ctx, H264B_E_INVALID, "stream_submit: null pointer in job");
    StreamSlot *sl = ctx->slot[slot_idx];
    // the slot's previous results may still be on their way out: reuse its buffers only after that
    H264B_CUDA(ctx, cudaEventSynchronize(sl->e_out));
    const size_t ms = j.max_slices ? j.max_slices : 1;
    const uint32_t cap = nal_cap;
    void *d_stream, *d_rbsp, *d_nals, *d_sum, *d_off, *d_len, *d_snal, *d_ops, *d_nops = nullptr, *d_qp, *d_boff, *d_bins,
        *d_fin, *h_boff;
    // bins layout: fixed by the caller's op counts, so it is known before the device knows how many slices there are
    RC(slot_pin(ctx, sl, 1, (ms + 1) * 8, &h_boff));
    uint64_t *boff = (uint64_t *)h_boff;
    boff[0] = 0;
    for (size_t s = 0; s < j.max_slices; s++) {
        uint32_t nb = j.n_ops ? j.n_ops[s] : j.n_ops_max;
        if (nb > j.n_ops_max) nb = j.n_ops_max;
        boff[s + 1] = boff[s] + ((uint64_t)nb + 1 + 31) / 32;
    }
    const size_t total_words = (size_t)boff[j.max_slices];
    RC(slot_dev(ctx, sl, 0, j.n + 64, &d_stream));
    RC(slot_dev(ctx, sl, 1, j.n + 64, &d_rbsp));
    RC(slot_dev(ctx, sl, 2, (size_t)cap * sizeof(h264b_nal), &d_nals));
    RC(slot_dev(ctx, sl, 3, 256, &d_sum));
    RC(slot_dev(ctx, sl, 4, ms * 8, &d_off));
    RC(slot_dev(ctx, sl, 5, ms * 4, &d_len));
    RC(slot_dev(ctx, sl, 6, ms * 4 + 16, &d_snal));
    RC(slot_dev(ctx, sl, 7, (size_t)j.n_ops_max * 2 + 16, &d_ops));
    if (j.n_ops) RC(slot_dev(ctx, sl, 8, ms * 4, &d_nops));
    RC(slot_dev(ctx, sl, 9, ms * sizeof(h264b_slice_qp), &d_qp));
    RC(slot_dev(ctx, sl, 10, (ms + 1) * 8, &d_boff));
    RC(slot_dev(ctx, sl, 11, total_words * 4, &d_bins));
    RC(slot_dev(ctx, sl, 12, ms * sizeof(h264b_cabac_final), &d_fin));
    // The NAL index is copied out with the job up to a length the caller's slice bound makes likely (its true
    // length is only known on the device); h264b_stream_wait fetches the rest in the rare case there is more.
    size_t nal_prefix = 2 * ms + 65536;
    if (nal_prefix > cap) nal_prefix = cap;
    void *h_nals, *h_bins, *h_fin, *h_snal, *h_sum;
    RC(slot_pin(ctx, sl, 0, nal_prefix * sizeof(h264b_nal), &h_nals));
    RC(slot_pin(ctx, sl, 2, total_words * 4, &h_bins));
    RC(slot_pin(ctx, sl, 3, ms * sizeof(h264b_cabac_final), &h_fin));
    RC(slot_pin(ctx, sl, 4, ms * 4, &h_snal));
    RC(slot_pin(ctx, sl, 5, 256, &h_sum));

    // 1. inputs: host -> device on the copy-in stream (overlaps the kernels of the job before)
    cudaStream_t in = ctx->s_in, out = ctx->s_out, cs = ctx->stream;
    if (ctx->trace && ctx->next_ticket == 0 && !sl->busy) cudaEventRecord(ctx->t_ref, in);
    H264B_CUDA(ctx, cudaStreamWaitEvent(in, sl->e_compute, 0));  // the slot's previous kernels have read its inputs
    if (ctx->trace) cudaEventRecord(sl->t_in0, in);
    if (j.n) H264B_CUDA(ctx, cudaMemcpyAsync(d_stream, j.stream, j.n, cudaMemcpyHostToDevice, in));
    void *h_ops, *h_nops, *h_qp;
    RC(slot_pin(ctx, sl, 6, (size_t)j.n_ops_max * 2, &h_ops));
    RC(slot_pin(ctx, sl, 7, ms * 4, &h_nops));
    RC(slot_pin(ctx, sl, 8, ms * sizeof(h264b_slice_qp) + sizeof(h264b_sps) + sizeof(h264b_pps), &h_qp));  // (+ initial sets)
    if (j.n_ops_max && j.max_slices) {
        memcpy(h_ops, j.ops, (size_t)j.n_ops_max * 2);
        H264B_CUDA(ctx, cudaMemcpyAsync(d_ops, h_ops, (size_t)j.n_ops_max * 2, cudaMemcpyHostToDevice, in));
    }
    if (j.max_slices) {
        if (!from_headers) {
            memcpy(h_qp, j.qp, (size_t)j.max_slices * sizeof(h264b_slice_qp));
            H264B_CUDA(ctx, cudaMemcpyAsync(d_qp, h_qp, ms * sizeof(h264b_slice_qp), cudaMemcpyHostToDevice, in));
        }
        if (j.n_ops) {
            memcpy(h_nops, j.n_ops, (size_t)j.max_slices * 4);
            H264B_CUDA(ctx, cudaMemcpyAsync(d_nops, h_nops, ms * 4, cudaMemcpyHostToDevice, in));
        }
        H264B_CUDA(ctx, cudaMemcpyAsync(d_boff, boff, (ms + 1) * 8, cudaMemcpyHostToDevice, in));
    }
    H264B_CUDA(ctx, cudaEventRecord(sl->e_in, in));

    // 2. kernels on the compute stream: split + strip, slice list, CABAC over the slices the device found
    H264B_CUDA(ctx, cudaStreamWaitEvent(cs, sl->e_in, 0));
    H264B_CUDA(ctx, cudaStreamWaitEvent(cs, sl->e_out, 0));  // (results of the slot's previous job have left)
    if (ctx->trace) cudaEventRecord(sl->t_c0, cs);
    uint32_t *d_ns = (uint32_t *)((uint8_t *)d_sum + 64);
    void *d_ext = nullptr;
    if (j.flags & H264B_STREAM_WANT_RBSP) RC(slot_dev(ctx, sl, 13, (size_t)cap * sizeof(h264b_nal_ext), &d_ext));
    RC(launch_annexb_scan(ctx, (const uint8_t *)d_stream, j.n, (uint8_t *)d_rbsp, (h264b_nal *)d_nals,
                          (h264b_nal_ext *)d_ext, cap, (h264b_scan_summary *)d_sum, j.flags));
    RC(launch_slice_select(ctx, (const h264b_nal *)d_nals, (const h264b_scan_summary *)d_sum, cap,
                           from_headers ? 0u : j.slice_data_offset, j.max_slices, (uint64_t *)d_off, (uint32_t *)d_len,
                           (uint32_t *)d_snal, d_ns, d_ns + 1));
    void *d_hdr = nullptr, *d_psl = nullptr, *d_sps = nullptr, *d_pps = nullptr, *d_sps_of = nullptr;
    if (from_headers) {  // SliceQPY, cabac_init_idc and the start of the CABAC data come from the slices' own headers
        RC(slot_dev(ctx, sl, 14, ms * sizeof(h264b_slice_header), &d_hdr));
        StreamParamSets sp;
        if (own_psets) {  // ... and the parameter sets from the stream's SPS / PPS NAL units
            RC(slot_dev(ctx, sl, 15, ((size_t)max_sps + max_pps + 4) * 4, &d_psl));
            RC(slot_dev(ctx, sl, 16, ((size_t)max_sps + 1) * sizeof(h264b_sps), &d_sps));  // (+ 1: job.initial_sps)
            RC(slot_dev(ctx, sl, 17, ((size_t)max_pps + 1) * sizeof(h264b_pps), &d_pps));
            RC(slot_dev(ctx, sl, 18, ms * 8, &d_sps_of));
            uint32_t *counts = (uint32_t *)d_psl, *sps_nal = counts + 4, *pps_nal = sps_nal + max_sps;
            RC(launch_pset_select(ctx, (const h264b_nal *)d_nals, (const h264b_scan_summary *)d_sum, cap, max_sps, max_pps,
                                  sps_nal, pps_nal, counts));
            RC(launch_parse_sps(ctx, (const uint8_t *)d_rbsp, j.n + 16, nullptr, nullptr, (const h264b_nal *)d_nals, sps_nal,
                                max_sps, counts, (h264b_sps *)d_sps));
            RC(launch_parse_pps(ctx, (const uint8_t *)d_rbsp, j.n + 16, nullptr, nullptr, (const h264b_nal *)d_nals, pps_nal,
                                max_pps, counts + 1, (h264b_pps *)d_pps));
            sp.sps = (const h264b_sps *)d_sps;
            sp.pps = (const h264b_pps *)d_pps;
            sp.sps_nal = sps_nal;
            sp.pps_nal = pps_nal;
            sp.counts = counts;
            sp.slice_sps = (int32_t *)d_sps_of;
            sp.slice_pps = sp.slice_sps + ms;
            sp.initial_sps = nullptr;
            sp.initial_pps = nullptr;
            if (j.initial_sps) {  // the sets the batch inherits: staged through the slot's pinned memory
                uint8_t *hp = (uint8_t *)h_qp + ms * sizeof(h264b_slice_qp);  // (behind the slot's qp staging area)
                memcpy(hp, j.initial_sps, sizeof(h264b_sps));
                H264B_CUDA(ctx, cudaMemcpyAsync((h264b_sps *)d_sps + max_sps, hp, sizeof(h264b_sps), cudaMemcpyHostToDevice, cs));
                sp.initial_sps = (const h264b_sps *)d_sps + max_sps;
                if (j.initial_pps) {
                    memcpy(hp + sizeof(h264b_sps), j.initial_pps, sizeof(h264b_pps));
                    H264B_CUDA(ctx, cudaMemcpyAsync((h264b_pps *)d_pps + max_pps, hp + sizeof(h264b_sps), sizeof(h264b_pps),
                                                    cudaMemcpyHostToDevice, cs));
                    sp.initial_pps = (const h264b_pps *)d_pps + max_pps;
                }
            }
        }
        RC(launch_stream_slice_headers(ctx, j.param_sets, (const uint8_t *)d_rbsp, j.n + 16, (const h264b_nal *)d_nals,
                                       (const uint32_t *)d_snal, d_ns, j.max_slices, (h264b_slice_header *)d_hdr,
                                       (uint64_t *)d_off, (uint32_t *)d_len, (h264b_slice_qp *)d_qp,
                                       own_psets ? &sp : nullptr));
    }
    if (j.max_slices) {
        h264b_cabac_job cj;
        memset(&cj, 0, sizeof(cj));
        cj.bytes = (const uint8_t *)d_rbsp;
        cj.total_bytes = j.n + 16;
        cj.off = (const uint64_t *)d_off;
        cj.len = (const uint32_t *)d_len;
        cj.n_slices = j.max_slices;  // the bound; the kernels read the actual count from d_ns
        cj.n_ctx = j.n_ctx;
        cj.ops = (const uint16_t *)d_ops;
        cj.n_ops_max = j.n_ops_max;
        cj.n_ops = (const uint32_t *)d_nops;
        cj.qp = (const h264b_slice_qp *)d_qp;
        cj.bins = (uint32_t *)d_bins;
        cj.bins_off = (const uint64_t *)d_boff;
        cj.final = (h264b_cabac_final *)d_fin;
        cj.flags = j.flags;
        // the schedule's context working set: only those rows have to live in shared memory
        uint32_t top = 0;
        for (uint32_t k = 0; k < j.n_ops_max; k++)
            if ((j.ops[k] >> 14) == H264B_OP_DECISION && (j.ops[k] & 0x3FFu) > top) top = j.ops[k] & 0x3FFu;
        cj.n_ctx_used = top + 1 < j.n_ctx ? top + 1 : 0;
        RC(launch_cabac(ctx, &cj, d_ns));
    }
    H264B_CUDA(ctx, cudaEventRecord(sl->e_compute, cs));

    // 3. results: device -> host on the copy-out stream (overlaps the kernels of the job after)
    H264B_CUDA(ctx, cudaStreamWaitEvent(out, sl->e_compute, 0));
    if (ctx->trace) cudaEventRecord(sl->t_o0, out);
    H264B_CUDA(ctx, cudaMemcpyAsync(h_sum, d_sum, 128, cudaMemcpyDeviceToHost, out));
    H264B_CUDA(ctx, cudaMemcpyAsync(h_nals, d_nals, nal_prefix * sizeof(h264b_nal), cudaMemcpyDeviceToHost, out));
    if (j.max_slices) {
        if (total_words) H264B_CUDA(ctx, cudaMemcpyAsync(h_bins, d_bins, total_words * 4, cudaMemcpyDeviceToHost, out));
        H264B_CUDA(ctx, cudaMemcpyAsync(h_fin, d_fin, ms * sizeof(h264b_cabac_final), cudaMemcpyDeviceToHost, out));
        H264B_CUDA(ctx, cudaMemcpyAsync(h_snal, d_snal, ms * 4, cudaMemcpyDeviceToHost, out));
    }
    if (from_headers) {
        void *h_hdr;
        RC(slot_pin(ctx, sl, 11, ms * sizeof(h264b_slice_header), &h_hdr));
        H264B_CUDA(ctx, cudaMemcpyAsync(h_hdr, d_hdr, ms * sizeof(h264b_slice_header), cudaMemcpyDeviceToHost, out));
    }
    if (own_psets) {
        void *h_psl, *h_sps, *h_pps, *h_sps_of;
        RC(slot_pin(ctx, sl, 12, ((size_t)max_sps + max_pps + 4) * 4, &h_psl));
        RC(slot_pin(ctx, sl, 13, (size_t)max_sps * sizeof(h264b_sps), &h_sps));
        RC(slot_pin(ctx, sl, 14, (size_t)max_pps * sizeof(h264b_pps), &h_pps));
        RC(slot_pin(ctx, sl, 15, ms * 8, &h_sps_of));
        H264B_CUDA(ctx, cudaMemcpyAsync(h_psl, d_psl, ((size_t)max_sps + max_pps + 4) * 4, cudaMemcpyDeviceToHost, out));
        H264B_CUDA(ctx, cudaMemcpyAsync(h_sps, d_sps, (size_t)max_sps * sizeof(h264b_sps), cudaMemcpyDeviceToHost, out));
        H264B_CUDA(ctx, cudaMemcpyAsync(h_pps, d_pps, (size_t)max_pps * sizeof(h264b_pps), cudaMemcpyDeviceToHost, out));
        H264B_CUDA(ctx, cudaMemcpyAsync(h_sps_of, d_sps_of, ms * 8, cudaMemcpyDeviceToHost, out));
    }
    if (j.flags & H264B_STREAM_WANT_RBSP) {  // position-preserving layout: the buffer is as long as the stream
        void *h_rbsp, *h_ext;
        RC(slot_pin(ctx, sl, 9, j.n, &h_rbsp));
        RC(slot_pin(ctx, sl, 10, nal_prefix * sizeof(h264b_nal_ext), &h_ext));
        if (j.n) H264B_CUDA(ctx, cudaMemcpyAsync(h_rbsp, d_rbsp, j.n, cudaMemcpyDeviceToHost, out));
        H264B_CUDA(ctx, cudaMemcpyAsync(h_ext, d_ext, nal_prefix * sizeof(h264b_nal_ext), cudaMemcpyDeviceToHost, out));
    }
    H264B_CUDA(ctx, cudaEventRecord(sl->e_out, out));
    sl->job = j;
    sl->nal_cap = cap;
    sl->nal_prefix = nal_prefix;
    sl->total_words = total_words;
    return H264B_OK;
}

int32_t h264b_stream_wait(h264b_ctx *ctx, uint64_t ticket, h264b_stream_result *res) {
    CHECK_CTX(ctx);
    if (!res) return H264B_E_INVALID;
    TraceRange trace_range("h264b:stream_wait");
    StreamSlot *sl = ctx->slot[ticket % kStreamSlots];
    if (!sl->busy || sl->ticket != ticket) return set_error(ctx, H264B_E_INVALID, "stream_wait: unknown ticket");
    H264B_CUDA(ctx, cudaEventSynchronize(sl->e_out));
    sl->busy = false;
    if (ctx->trace) {
        float a = 0, b = 0, c = 0, d = 0, e = 0, f = 0;
        cudaEventElapsedTime(&a, ctx->t_ref, sl->t_in0);
        cudaEventElapsedTime(&b, ctx->t_ref, sl->e_in);
        cudaEventElapsedTime(&c, ctx->t_ref, sl->t_c0);
        cudaEventElapsedTime(&d, ctx->t_ref, sl->e_compute);
        cudaEventElapsedTime(&e, ctx->t_ref, sl->t_o0);
        cudaEventElapsedTime(&f, ctx->t_ref, sl->e_out);
        fprintf(stderr, "h264b trace: job %llu  H2D %.1f..%.1f  kernels %.1f..%.1f  D2H %.1f..%.1f ms\n",
                (unsigned long long)ticket, a, b, c, d, e, f);
    }
    // A bound of the job that turned out too small -- the NAL index (streams of very short NAL units), the SPS / PPS
    // lists, the slice list of a job whose per-slice inputs all come from the stream -- is raised to what the run
    // reported and the job runs once more on its slot (its host buffers are still valid: the wait has not returned).
    for (int attempt = 0;; attempt++) {
        memset(res, 0, sizeof(*res));
        memcpy(&res->scan, sl->h[5], sizeof(res->scan));
        h264b_stream_job jr = sl->job;
        uint32_t cap = sl->nal_cap;
        bool again = false;
        if (res->scan.status != H264B_OK) {
            if (res->scan.n_start_codes + 1 > 0xFFFFFFFFull)
                return set_error(ctx, H264B_E_CAPACITY, "stream: %llu NAL units", (unsigned long long)res->scan.n_nals);
            cap = (uint32_t)(res->scan.n_start_codes + 1);
            again = true;
        } else {
            const bool headers = (jr.flags & (H264B_STREAM_SLICE_HEADERS | H264B_STREAM_PARAM_SETS)) != 0 && jr.max_slices;
            const uint32_t slices_found = *(const uint32_t *)((const uint8_t *)sl->h[5] + 68);
            if (headers && (jr.flags & H264B_STREAM_PARAM_SETS) && !jr.n_ops && slices_found > jr.max_slices) {
                jr.max_slices = slices_found;
                again = true;
            }
            if (headers && (jr.flags & H264B_STREAM_PARAM_SETS)) {
                const uint32_t max_sps = jr.max_sps ? jr.max_sps : 64u, max_pps = jr.max_pps ? jr.max_pps : 64u;
                const uint32_t *counts = (const uint32_t *)sl->h[12];
                if (counts[2] > max_sps || counts[3] > max_pps) {
                    jr.max_sps = counts[2] > max_sps ? counts[2] : max_sps;
                    jr.max_pps = counts[3] > max_pps ? counts[3] : max_pps;
                    again = true;
                }
            }
        }
        if (!again) break;
        if (attempt >= 2) return set_error(ctx, H264B_E_CAPACITY, "stream: the job's bounds keep falling short");
        const cudaStream_t saved_stream = ctx->stream;
        const int saved_bank = ctx->bank;
        const int slot_idx = (int)(ticket % kStreamSlots);
        ctx->stream = sl->cs;
        ctx->bank = 1 + slot_idx;
        const int32_t rc = stream_submit_on_slot(ctx, &jr, slot_idx, cap);
        ctx->stream = saved_stream;
        ctx->bank = saved_bank;
        if (rc != H264B_OK) return rc;
        H264B_CUDA(ctx, cudaEventSynchronize(sl->e_out));
    }
    uint32_t n_slices = *(const uint32_t *)((const uint8_t *)sl->h[5] + 64);
    if (n_slices > sl->job.max_slices) n_slices = sl->job.max_slices;
    void *h_nals = sl->h[0];
    const size_t nn = (size_t)res->scan.n_nals;
    if (nn > sl->nal_prefix) {  // more NAL units than the prefix copied with the job: fetch the whole index now (this
        // queues behind whatever the copy-out stream is doing for the next job)
        RC(slot_pin(ctx, sl, 0, nn * sizeof(h264b_nal), &h_nals));
        H264B_CUDA(ctx, cudaMemcpyAsync(h_nals, sl->d[2], nn * sizeof(h264b_nal), cudaMemcpyDeviceToHost, ctx->s_out));
        if (sl->job.flags & H264B_STREAM_WANT_RBSP) {
            void *h_ext;
            RC(slot_pin(ctx, sl, 10, nn * sizeof(h264b_nal_ext), &h_ext));
            H264B_CUDA(ctx, cudaMemcpyAsync(h_ext, sl->d[13], nn * sizeof(h264b_nal_ext), cudaMemcpyDeviceToHost, ctx->s_out));
        }
        H264B_CUDA(ctx, cudaStreamSynchronize(ctx->s_out));
        sl->nal_prefix = nn;
    }
    res->nals = (const h264b_nal *)h_nals;
    res->n_slices = n_slices;
    res->bins_off = (const uint64_t *)sl->h[1];
    res->bins = (const uint32_t *)sl->h[2];
    res->final = (const h264b_cabac_final *)sl->h[3];
    res->slice_nal = (const uint32_t *)sl->h[4];
    for (uint32_t s = 0; s < n_slices; s++) res->total_bins += res->final[s].n_bins;
    res->rbsp = (sl->job.flags & H264B_STREAM_WANT_RBSP) ? (const uint8_t *)sl->h[9] : nullptr;
    res->d_rbsp = (const uint8_t *)sl->d[1];
    res->ext = (sl->job.flags & H264B_STREAM_WANT_RBSP) ? (const h264b_nal_ext *)sl->h[10] : nullptr;
    res->headers = ((sl->job.flags & (H264B_STREAM_SLICE_HEADERS | H264B_STREAM_PARAM_SETS)) && sl->job.max_slices)
                       ? (const h264b_slice_header *)sl->h[11]
                       : nullptr;
    if (res->headers && (sl->job.flags & H264B_STREAM_PARAM_SETS)) {
        const uint32_t max_sps = sl->job.max_sps ? sl->job.max_sps : 64u;
        const uint32_t *counts = (const uint32_t *)sl->h[12];
        res->n_sps = counts[0];
        res->n_pps = counts[1];
        res->sps_nal = counts + 4;
        res->pps_nal = counts + 4 + max_sps;
        res->sps = (const h264b_sps *)sl->h[13];
        res->pps = (const h264b_pps *)sl->h[14];
        res->slice_sps = (const int32_t *)sl->h[15];
        res->slice_pps = res->slice_sps + (sl->job.max_slices ? sl->job.max_slices : 1);
    }
    return H264B_OK;
}

int32_t h264b_stream_decode(h264b_ctx *ctx, const h264b_stream_job *job, h264b_stream_result *res) {
    uint64_t ticket;
    RC(h264b_stream_submit(ctx, job, &ticket));
    return h264b_stream_wait(ctx, ticket, res);
}

}  // extern "C"
