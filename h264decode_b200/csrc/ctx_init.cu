// ctx_init.cu -- K4: CABAC context-variable initialisation, and the constant device tables.
//
// Reference: PreCtxState (h264/cabac.go:118-121) = Clip3(1,126, ((m*Clip3(0,51,qp))>>4)+n), the state split of
// initCabac (cabac.go:158-164) and the (m,n) lookups MNVars (h264/mn_vars.go:15-175) / CodedblockPatternMN
// (mn_vars.go:184-440).  The result for a slice depends only on (clipped qp, idc class), 52 x 5 combinations, so the
// formula is evaluated once per (class, qp, ctxIdx) by state_lut_kernel (one thread each, on the GPU) into a
// 266 KB table that stays in L2, and the per-slice kernel is a pure gather-copy at HBM write bandwidth:
// algorithmic bytes per slice = n_ctx (states written) + 8 (params read).
#include "common.cuh"
#include "tables.inc"

namespace h264b {

constexpr int kCtxMax = 1024;
constexpr int kClasses = 5;  // idc -1, 0, 1, 2, other

// idc class of a cabac_init_idc value: Go map / switch semantics for keys the tables do not have
__host__ __device__ __forceinline__ int idc_class(int idc) { return (idc >= -1 && idc <= 2) ? idc + 1 : 4; }

// (m | n << 8) for (class, ctxIdx): MNVars has keys -1 (ctx 0..10) or 0,1,2 (ctx 11..39); CodedblockPatternMN
// returns the I/SI column for every idc outside 0..2.
__global__ void mn_table_kernel(const int8_t *m_tab, const int8_t *n_tab, int16_t *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kClasses * kCtxMax) return;
    const int cls = i / kCtxMax, c = i % kCtxMax;
    int col;
    if (c >= 70 && c <= 104)
        col = (cls >= 1 && cls <= 3) ? cls : 0;
    else
        col = cls <= 3 ? cls : -1;
    int m = 0, n = 0;
    if (col >= 0) {
        m = m_tab[col * kCtxMax + c];
        n = n_tab[col * kCtxMax + c];
    }
    out[i] = (int16_t)((m & 0xFF) | (n << 8));
}

__device__ __forceinline__ int clip3(int x, int y, int z) { return z < x ? x : (z > y ? y : z); }

// one thread per (class, qp, ctxIdx): PreCtxState + state split; >> on a negative int is arithmetic in Go and here
__global__ void state_lut_kernel(const int16_t *mn, uint8_t *lut) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kClasses * 52 * kCtxMax) return;
    const int c = i % kCtxMax, qp = (i / kCtxMax) % 52, cls = i / (kCtxMax * 52);
    const int16_t e = mn[cls * kCtxMax + c];
    const int m = (int8_t)(e & 0xFF), n = e >> 8;
    const int pre = clip3(1, 126, ((m * clip3(0, 51, qp)) >> 4) + n);
    lut[i] = (uint8_t)(pre <= 63 ? (63 - pre) : ((pre - 64) | 64));
}

// CABAC engine table, indexed by the state byte s = pStateIdx | valMPS << 6:
//   bits  0..31  rangeTabLPS[pStateIdx][0..3]
//   bits 32..39  next state after an MPS        bits 40..47  bin value of the MPS
//   bits 48..55  next state after an LPS        bits 56..63  bin value of the LPS
// (StateTransitionProcess, h264/cabac.go:544-553: valMPS flips on an LPS at pStateIdx 0)
__global__ void cabac_table_kernel(const uint8_t *range_lps, const uint8_t *trans, uint64_t *out) {
    const int s = threadIdx.x;
    if (s >= 128) return;
    const uint32_t p = s & 63, v = s >> 6;
    const uint32_t lo = range_lps[p * 4] | (range_lps[p * 4 + 1] << 8) | (range_lps[p * 4 + 2] << 16) |
                        ((uint32_t)range_lps[p * 4 + 3] << 24);
    const uint32_t next_mps = trans[64 + p] | (v << 6);
    const uint32_t v_lps = (p == 0) ? 1 - v : v;
    const uint32_t next_lps = trans[p] | (v_lps << 6);
    const uint32_t hi = next_mps | (v << 8) | (next_lps << 16) | ((1 - v) << 24);
    out[s] = ((uint64_t)hi << 32) | lo;
}

int build_tables(h264b_ctx *ctx) {
    const uint8_t *h_range[2] = {h264b_range_tab_lps_ref, h264b_range_tab_lps_spec};
    const uint8_t *h_lps[2] = {h264b_trans_idx_lps_ref, h264b_trans_idx_lps_spec};
    const uint8_t *h_mps[2] = {h264b_trans_idx_mps_ref, h264b_trans_idx_mps_spec};
    const int8_t *h_m[2] = {h264b_mn_m_ref, h264b_mn_m_spec};
    const int8_t *h_n[2] = {h264b_mn_n_ref, h264b_mn_n_spec};
    for (int v = 0; v < 2; v++) {
        int8_t *d_m = nullptr, *d_n = nullptr;
        H264B_CUDA(ctx, cudaMalloc(&ctx->d_range_lps[v], 256));
        H264B_CUDA(ctx, cudaMalloc(&ctx->d_trans[v], 128));
        H264B_CUDA(ctx, cudaMalloc(&ctx->d_cabac_tab[v], 128 * 8));
        H264B_CUDA(ctx, cudaMalloc(&ctx->d_mn[v], kClasses * kCtxMax * 2));
        H264B_CUDA(ctx, cudaMalloc(&ctx->d_state_lut[v], kClasses * 52 * kCtxMax));
        H264B_CUDA(ctx, cudaMalloc(&d_m, 4 * kCtxMax));
        H264B_CUDA(ctx, cudaMalloc(&d_n, 4 * kCtxMax));
        H264B_CUDA(ctx, cudaMemcpyAsync(ctx->d_range_lps[v], h_range[v], 256, cudaMemcpyHostToDevice, ctx->stream));
        H264B_CUDA(ctx, cudaMemcpyAsync(ctx->d_trans[v], h_lps[v], 64, cudaMemcpyHostToDevice, ctx->stream));
        H264B_CUDA(ctx, cudaMemcpyAsync(ctx->d_trans[v] + 64, h_mps[v], 64, cudaMemcpyHostToDevice, ctx->stream));
        H264B_CUDA(ctx, cudaMemcpyAsync(d_m, h_m[v], 4 * kCtxMax, cudaMemcpyHostToDevice, ctx->stream));
        H264B_CUDA(ctx, cudaMemcpyAsync(d_n, h_n[v], 4 * kCtxMax, cudaMemcpyHostToDevice, ctx->stream));
        mn_table_kernel<<<(kClasses * kCtxMax + 255) / 256, 256, 0, ctx->stream>>>(d_m, d_n, ctx->d_mn[v]);
        H264B_LAUNCH_CHECK(ctx, "mn_table_kernel");
        state_lut_kernel<<<(kClasses * 52 * kCtxMax + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_mn[v],
                                                                                          ctx->d_state_lut[v]);
        H264B_LAUNCH_CHECK(ctx, "state_lut_kernel");
        cabac_table_kernel<<<1, 128, 0, ctx->stream>>>(ctx->d_range_lps[v], ctx->d_trans[v], ctx->d_cabac_tab[v]);
        H264B_LAUNCH_CHECK(ctx, "cabac_table_kernel");
        H264B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(d_m);
        cudaFree(d_n);
    }
    return H264B_OK;
}

// ------------------------------------------------------------------------------------------------ K4
// states[s][c] = lut[class(idc_s)][clip(qp_s)][c].  Each thread moves 16 bytes per step (one LUT granule from L2 ->
// one streaming 16-byte store); consecutive threads write consecutive granules, so every warp store is 512
// contiguous bytes.
__global__ void __launch_bounds__(256) ctx_init_kernel(const h264b_slice_qp *__restrict__ params, uint32_t n_slices,
                                                       uint32_t gran_per_slice, const uint8_t *__restrict__ lut,
                                                       uint8_t *__restrict__ states) {
    const uint64_t total = (uint64_t)n_slices * gran_per_slice;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += stride) {
        const uint32_t s = (uint32_t)(g / gran_per_slice), j = (uint32_t)(g % gran_per_slice);
        const h264b_slice_qp p = params[s];
        const int row = idc_class(p.cabac_init_idc) * 52 + clip3(0, 51, p.slice_qp_y);
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(lut + (size_t)row * kCtxMax) + j);
        __stcs(reinterpret_cast<uint4 *>(states) + g, v);
    }
}
// n_ctx not a multiple of 16 (rows are not 16-byte aligned): byte-granular variant
__global__ void __launch_bounds__(256) ctx_init_bytes_kernel(const h264b_slice_qp *__restrict__ params,
                                                             uint32_t n_slices, uint32_t n_ctx,
                                                             const uint8_t *__restrict__ lut,
                                                             uint8_t *__restrict__ states) {
    const uint64_t total = (uint64_t)n_slices * n_ctx;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const uint32_t s = (uint32_t)(i / n_ctx), c = (uint32_t)(i % n_ctx);
        const h264b_slice_qp p = params[s];
        const int row = idc_class(p.cabac_init_idc) * 52 + clip3(0, 51, p.slice_qp_y);
        states[i] = lut[(size_t)row * kCtxMax + c];
    }
}

int launch_ctx_init(h264b_ctx *ctx, const h264b_slice_qp *d_params, uint32_t n_slices, uint32_t n_ctx,
                    uint8_t *d_states, uint32_t flags) {
    TraceRange trace_range("h264b:ctx_init");
    if (n_ctx < 1 || n_ctx > kCtxMax) return set_error(ctx, H264B_E_INVALID, "ctx_init: n_ctx must be 1..1024");
    if (!n_slices) return H264B_OK;
    const uint8_t *lut = ctx->d_state_lut[(flags & H264B_TABLES_SPEC) ? 1 : 0];
    if ((n_ctx & 15) == 0 && ((uintptr_t)d_states & 15) == 0) {
        const uint64_t total = (uint64_t)n_slices * (n_ctx / 16);
        uint64_t blocks = (total + 255) / 256;
        const uint64_t cap = (uint64_t)ctx->sm_count * 8 * 4;  // 8 resident CTAs per SM, ~4 granules per thread
        if (blocks > cap) blocks = cap;
        ctx_init_kernel<<<(int)blocks, 256, 0, ctx->stream>>>(d_params, n_slices, n_ctx / 16, lut, d_states);
        H264B_LAUNCH_CHECK(ctx, "ctx_init_kernel");
    } else {
        const uint64_t total = (uint64_t)n_slices * n_ctx;
        uint64_t blocks = (total + 255) / 256;
        const uint64_t cap = (uint64_t)ctx->sm_count * 32;
        if (blocks > cap) blocks = cap;
        ctx_init_bytes_kernel<<<(int)blocks, 256, 0, ctx->stream>>>(d_params, n_slices, n_ctx, lut, d_states);
        H264B_LAUNCH_CHECK(ctx, "ctx_init_bytes_kernel");
    }
    return H264B_OK;
}

// scalar drop-ins -----------------------------------------------------------------------------------------------
__global__ void pre_ctx_state_kernel(int m, int n, int qp, int *out) {
    *out = clip3(1, 126, ((m * clip3(0, 51, qp)) >> 4) + n);
}
__global__ void mn_lookup_kernel(const int16_t *mn, int ctx_idx, int idc, int *out) {
    int m = 0, n = 0;
    if (ctx_idx >= 0 && ctx_idx < kCtxMax) {
        const int16_t e = mn[idc_class(idc) * kCtxMax + ctx_idx];
        m = (int8_t)(e & 0xFF);
        n = e >> 8;
    }
    out[0] = m;
    out[1] = n;
}

}  // namespace h264b

extern "C" int32_t h264b_pre_ctx_state(h264b_ctx *ctx, int32_t m, int32_t n, int32_t qp, int32_t *out) {
    if (!ctx || !out) return H264B_E_INVALID;
    void *d;
    int rc = h264b::ensure_dev(ctx, 15, 64, &d);
    if (rc) return rc;
    h264b::pre_ctx_state_kernel<<<1, 1, 0, ctx->stream>>>(m, n, qp, (int *)d);
    H264B_LAUNCH_CHECK(ctx, "pre_ctx_state_kernel");
    H264B_CUDA(ctx, cudaMemcpyAsync(out, d, 4, cudaMemcpyDeviceToHost, ctx->stream));
    H264B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return H264B_OK;
}

extern "C" int32_t h264b_mn(h264b_ctx *ctx, int32_t ctx_idx, int32_t idc, uint32_t flags, int32_t *m, int32_t *n) {
    if (!ctx || !m || !n) return H264B_E_INVALID;
    void *d;
    int rc = h264b::ensure_dev(ctx, 15, 64, &d);
    if (rc) return rc;
    h264b::mn_lookup_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_mn[(flags & H264B_TABLES_SPEC) ? 1 : 0], ctx_idx, idc,
                                                      (int *)d);
    H264B_LAUNCH_CHECK(ctx, "mn_lookup_kernel");
    int tmp[2];
    H264B_CUDA(ctx, cudaMemcpyAsync(tmp, d, 8, cudaMemcpyDeviceToHost, ctx->stream));
    H264B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *m = tmp[0];
    *n = tmp[1];
    return H264B_OK;
}
