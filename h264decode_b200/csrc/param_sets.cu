// param_sets.cu -- rows S1 / f4: NewSPS (h264/sps.go:192-437) and NewPPS (h264/pps.go:40-133) for all parameter-set
// NAL units at once, one thread per NAL (a serial Exp-Golomb walk of a few hundred bits; parameter sets are independent
// of each other -- NewPPS reads its *SPS only after it has already panicked).  The walk itself is param_sets.cuh, shared
// with the CPU emulation of tests/.  Also here: the ordered lists of the type-7 / type-8 NAL units of a scanned stream
// (handleConnection's dispatch, h264/server.go:147-158), which keep the parameter sets device-resident for the
// slice-header kernel.
#include "common.cuh"
#include "param_sets.cuh"

namespace h264b {

struct PsetArgs {
    const uint8_t *bytes;
    uint64_t total_bytes;
    const uint64_t *off;
    const uint32_t *len;
    const h264b_nal *nals;
    const uint32_t *nal_index;
    uint32_t n;
    const uint32_t *n_dev;
};

__device__ __forceinline__ bool pset_locate(const PsetArgs &a, uint32_t i, uint64_t *off, uint64_t *len) {
    if (i >= a.n || (a.n_dev && i >= *a.n_dev)) return false;
    if (a.nals) {
        const h264b_nal u = a.nals[a.nal_index[i]];
        *off = u.rbsp_off;
        *len = u.rbsp_len;
    } else {
        *off = a.off[i];
        *len = a.len[i];
    }
    if (*off > a.total_bytes) *off = a.total_bytes;
    if (*len > a.total_bytes - *off) *len = a.total_bytes - *off;
    return true;
}

// The records are zeroed and filled in place (4.3 KB each: too large for registers; global stores of a handful of
// threads are not a bottleneck of anything).
__global__ void __launch_bounds__(64) parse_sps_kernel(PsetArgs a, h264b_sps *out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t off, len;
    if (!pset_locate(a, i, &off, &len)) return;
    h264b_sps *o = out + i;
    uint64_t *w = reinterpret_cast<uint64_t *>(o);
    for (uint32_t k = 0; k < sizeof(h264b_sps) / 8; k++) w[k] = 0;
    o->status = parse_sps(a.bytes + off, len, o);
}

__global__ void __launch_bounds__(64) parse_pps_kernel(PsetArgs a, h264b_pps *out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t off, len;
    if (!pset_locate(a, i, &off, &len)) return;
    h264b_pps o = {};
    o.status = parse_pps(a.bytes + off, len, &o);
    out[i] = o;
}

// Ordered lists of the NAL units of type 7 and of type 8.  Single CTA, ballot-based stable compaction of both kinds in
// one walk over the index; counts[0..1] = entries kept, counts[2..3] = entries found.
__global__ void __launch_bounds__(1024) pset_select_kernel(const h264b_nal *nals, const h264b_scan_summary *summary,
                                                           uint32_t nal_cap, uint32_t max_sps, uint32_t max_pps,
                                                           uint32_t *sps_nal, uint32_t *pps_nal, uint32_t *counts) {
    __shared__ uint32_t warp_cnt[2][32];
    __shared__ uint32_t base[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint64_t n = summary->n_nals;
    if (n > nal_cap) n = nal_cap;
    if (summary->status != H264B_OK) n = 0;  // an incomplete index holds no records at all
    if (tid < 2) base[tid] = 0;
    __syncthreads();
    for (uint64_t k0 = 0; k0 < n; k0 += 1024) {
        const uint64_t k = k0 + tid;
        const uint32_t type = k < n ? nals[k].type : 0u;
        const uint32_t m7 = __ballot_sync(0xFFFFFFFFu, type == 7u), m8 = __ballot_sync(0xFFFFFFFFu, type == 8u);
        if (lane == 0) {
            warp_cnt[0][warp] = __popc(m7);
            warp_cnt[1][warp] = __popc(m8);
        }
        __syncthreads();
        if (type == 7u || type == 8u) {
            const int q = type == 8u;
            uint32_t wbase = 0;
            for (int w2 = 0; w2 < warp; w2++) wbase += warp_cnt[q][w2];
            const uint32_t idx = base[q] + wbase + __popc((q ? m8 : m7) & ((1u << lane) - 1u));
            if (idx < (q ? max_pps : max_sps)) (q ? pps_nal : sps_nal)[idx] = (uint32_t)k;
        }
        __syncthreads();
        if (tid < 2) {
            uint32_t tot = 0;
            for (int w2 = 0; w2 < 32; w2++) tot += warp_cnt[tid][w2];
            base[tid] += tot;
        }
        __syncthreads();
    }
    if (tid < 2) {
        const uint32_t cap = tid ? max_pps : max_sps;
        counts[tid] = base[tid] < cap ? base[tid] : cap;
        counts[2 + tid] = base[tid];
    }
}

int launch_pset_select(h264b_ctx *ctx, const h264b_nal *d_nals, const h264b_scan_summary *d_summary, uint32_t nal_cap,
                       uint32_t max_sps, uint32_t max_pps, uint32_t *d_sps_nal, uint32_t *d_pps_nal, uint32_t *d_counts) {
    pset_select_kernel<<<1, 1024, 0, ctx->stream>>>(d_nals, d_summary, nal_cap, max_sps, max_pps, d_sps_nal, d_pps_nal,
                                                    d_counts);
    H264B_LAUNCH_CHECK(ctx, "pset_select_kernel");
    return H264B_OK;
}

static PsetArgs pset_args(const uint8_t *d_bytes, uint64_t total, const uint64_t *d_off, const uint32_t *d_len,
                          const h264b_nal *d_nals, const uint32_t *d_nal_index, uint32_t n, const uint32_t *d_n) {
    PsetArgs a;
    a.bytes = d_bytes;
    a.total_bytes = total;
    a.off = d_off;
    a.len = d_len;
    a.nals = d_nals;
    a.nal_index = d_nal_index;
    a.n = n;
    a.n_dev = d_n;
    return a;
}

int launch_parse_sps(h264b_ctx *ctx, const uint8_t *d_bytes, uint64_t total, const uint64_t *d_off, const uint32_t *d_len,
                     const h264b_nal *d_nals, const uint32_t *d_nal_index, uint32_t n, const uint32_t *d_n,
                     h264b_sps *d_out) {
    if (!n) return H264B_OK;
    parse_sps_kernel<<<(n + 63) / 64, 64, 0, ctx->stream>>>(pset_args(d_bytes, total, d_off, d_len, d_nals, d_nal_index, n, d_n),
                                                            d_out);
    H264B_LAUNCH_CHECK(ctx, "parse_sps_kernel");
    return H264B_OK;
}

int launch_parse_pps(h264b_ctx *ctx, const uint8_t *d_bytes, uint64_t total, const uint64_t *d_off, const uint32_t *d_len,
                     const h264b_nal *d_nals, const uint32_t *d_nal_index, uint32_t n, const uint32_t *d_n,
                     h264b_pps *d_out) {
    if (!n) return H264B_OK;
    parse_pps_kernel<<<(n + 63) / 64, 64, 0, ctx->stream>>>(pset_args(d_bytes, total, d_off, d_len, d_nals, d_nal_index, n, d_n),
                                                            d_out);
    H264B_LAUNCH_CHECK(ctx, "parse_pps_kernel");
    return H264B_OK;
}

// host buffers -> device scratch, parse, records back
template <typename Rec, typename Launch>
static int parse_host(h264b_ctx *ctx, const uint8_t *bytes, uint64_t total_bytes, const uint64_t *off, const uint32_t *len,
                      uint32_t n, Rec *out, Launch launch) {
    if (!n) return H264B_OK;
    if (!bytes || !off || !len || !out) return set_error(ctx, H264B_E_INVALID, "parse_sps/pps: null pointer");
    void *d_bytes, *d_off, *d_len, *d_out;
    int rc;
    if ((rc = ensure_dev(ctx, 0, total_bytes + 64, &d_bytes))) return rc;
    if ((rc = ensure_dev(ctx, 5, (size_t)n * 8, &d_off))) return rc;
    if ((rc = ensure_dev(ctx, 6, (size_t)n * 4, &d_len))) return rc;
    if ((rc = ensure_dev(ctx, 19, (size_t)n * sizeof(Rec), &d_out))) return rc;
    H264B_CUDA(ctx, cudaMemcpyAsync(d_bytes, bytes, total_bytes, cudaMemcpyHostToDevice, ctx->stream));
    H264B_CUDA(ctx, cudaMemcpyAsync(d_off, off, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    H264B_CUDA(ctx, cudaMemcpyAsync(d_len, len, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = launch(ctx, (const uint8_t *)d_bytes, total_bytes, (const uint64_t *)d_off, (const uint32_t *)d_len, nullptr,
                     nullptr, n, nullptr, (Rec *)d_out)))
        return rc;
    H264B_CUDA(ctx, cudaMemcpyAsync(out, d_out, (size_t)n * sizeof(Rec), cudaMemcpyDeviceToHost, ctx->stream));
    H264B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return H264B_OK;
}

}  // namespace h264b

using namespace h264b;

extern "C" {

int32_t h264b_parse_sps_dev(h264b_ctx *ctx, const uint8_t *d_bytes, uint64_t total_bytes, const uint64_t *d_off,
                            const uint32_t *d_len, const h264b_nal *d_nals, const uint32_t *d_nal_index, uint32_t n,
                            const uint32_t *d_n, h264b_sps *d_out) {
    if (!ctx) return H264B_E_INVALID;
    cudaSetDevice(ctx->device);
    if (n && (!d_bytes || !d_out || (d_nals ? !d_nal_index : (!d_off || !d_len))))
        return set_error(ctx, H264B_E_INVALID, "parse_sps: null pointer");
    return launch_parse_sps(ctx, d_bytes, total_bytes, d_off, d_len, d_nals, d_nal_index, n, d_n, d_out);
}

int32_t h264b_parse_pps_dev(h264b_ctx *ctx, const uint8_t *d_bytes, uint64_t total_bytes, const uint64_t *d_off,
                            const uint32_t *d_len, const h264b_nal *d_nals, const uint32_t *d_nal_index, uint32_t n,
                            const uint32_t *d_n, h264b_pps *d_out) {
    if (!ctx) return H264B_E_INVALID;
    cudaSetDevice(ctx->device);
    if (n && (!d_bytes || !d_out || (d_nals ? !d_nal_index : (!d_off || !d_len))))
        return set_error(ctx, H264B_E_INVALID, "parse_pps: null pointer");
    return launch_parse_pps(ctx, d_bytes, total_bytes, d_off, d_len, d_nals, d_nal_index, n, d_n, d_out);
}

int32_t h264b_parse_sps(h264b_ctx *ctx, const uint8_t *bytes, uint64_t total_bytes, const uint64_t *off,
                        const uint32_t *len, uint32_t n, h264b_sps *out) {
    if (!ctx) return H264B_E_INVALID;
    cudaSetDevice(ctx->device);
    return parse_host(ctx, bytes, total_bytes, off, len, n, out, launch_parse_sps);
}

int32_t h264b_parse_pps(h264b_ctx *ctx, const uint8_t *bytes, uint64_t total_bytes, const uint64_t *off,
                        const uint32_t *len, uint32_t n, h264b_pps *out) {
    if (!ctx) return H264B_E_INVALID;
    cudaSetDevice(ctx->device);
    return parse_host(ctx, bytes, total_bytes, off, len, n, out, launch_parse_pps);
}

int32_t h264b_make_param_sets(const h264b_sps *sps, const h264b_pps *pps, h264b_param_sets *out) {
    if (!sps || !pps || !out) return H264B_E_INVALID;
    *out = make_param_sets(*sps, *pps);
    return H264B_OK;
}

int32_t h264b_param_set_select_dev(h264b_ctx *ctx, const h264b_nal *d_nals, const h264b_scan_summary *d_summary,
                                   uint32_t nal_cap, uint32_t max_sps, uint32_t max_pps, uint32_t *d_sps_nal,
                                   uint32_t *d_pps_nal, uint32_t *d_counts) {
    if (!ctx) return H264B_E_INVALID;
    cudaSetDevice(ctx->device);
    if (!d_nals || !d_summary || !d_counts || (max_sps && !d_sps_nal) || (max_pps && !d_pps_nal))
        return set_error(ctx, H264B_E_INVALID, "param_set_select: null pointer");
    return launch_pset_select(ctx, d_nals, d_summary, nal_cap, max_sps, max_pps, d_sps_nal, d_pps_nal, d_counts);
}

}  // extern "C"
