// harness_gpu.cu -- the synthetic-data harness of harness_core.h as CUDA kernels, for bench inputs too large to
// encode on the host (BASELINE configs[3]: ~4 GB of CABAC slice data).  INPUT GENERATION ONLY: never inside a timed
// region, not part of the product (libh264b200.so) and not part of the oracle.  Same encoder source as the CPU
// build (harness.c), so tests can cross-check the two byte for byte.
#include <cuda_runtime.h>
#include <stdint.h>

#include "harness_core.h"
#include "harness_tables.h"

#define HZ_TABLES_SPEC 1u
#define HZ_ESCAPE 2u
#define HZG_MAX_CTX 1024

__device__ __forceinline__ void hzg_mn(const int8_t *tm, const int8_t *tn, int ctx, int idc, int *m, int *n) {
    *m = *n = 0;
    int col;
    if (ctx >= 70 && ctx <= 104)
        col = (idc >= 0 && idc <= 2) ? idc + 1 : 0;
    else if (idc >= -1 && idc <= 2)
        col = idc + 1;
    else
        return;
    *m = tm[col * HZ_N_CTX_MAX + ctx];
    *n = tn[col * HZ_N_CTX_MAX + ctx];
}

// one thread per slice; context states live in a per-thread slab of global scratch (d_states, n_ctx bytes each)
__global__ void __launch_bounds__(64) hzg_encode_kernel(uint32_t flags, uint32_t config, uint64_t id_base,
                                                        int64_t n_slices, const uint16_t *ops, const uint32_t *n_ops,
                                                        uint32_t n_active, uint32_t n_ctx, const int32_t *qp,
                                                        const int32_t *idc, uint8_t *data, int64_t stride,
                                                        int64_t *lens, uint32_t *bins, int64_t bins_stride,
                                                        uint8_t *states_all, const uint8_t *range_lps,
                                                        const uint8_t *trans_lps, const uint8_t *trans_mps,
                                                        const int8_t *mn_m, const int8_t *mn_n, int *overflow) {
    __shared__ uint8_t s_range[256], s_lps[64], s_mps[64];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_range[i] = range_lps[i];
    for (int i = threadIdx.x; i < 64; i += blockDim.x) {
        s_lps[i] = trans_lps[i];
        s_mps[i] = trans_mps[i];
    }
    __syncthreads();
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slices) return;
    hz_tables t;
    t.range_lps = s_range;
    t.trans_lps = s_lps;
    t.trans_mps = s_mps;
    uint8_t *states = states_all + s * n_ctx;
    for (uint32_t c = 0; c < n_ctx; c++) {
        int m, n;
        hzg_mn(mn_m, mn_n, (int)c, idc[s], &m, &n);
        states[c] = hz_ctx_state(m, n, qp[s]);
    }
    hz_rng r = hz_seed(config, id_base + (uint64_t)s);
    hz_sink sink;
    hz_sink_init(&sink, data + s * stride, stride, (flags & HZ_ESCAPE) ? 1 : 0);
    hz_encode_slice(&t, &r, ops, n_ops[s], n_active, states, &sink, bins ? bins + s * bins_stride : nullptr);
    lens[s] = sink.n;
    if (sink.overflow) *overflow = 1;
}

// Annex-B assembly: NAL i = 00 00 00 01 | hdr[i] | data[i*stride .. +lens[i]) written at stream + pos[i];
// when pre_len > 0 and i % params_every == 0 the `pre` bytes (start codes + SPS + PPS) go right before it.
// One CTA per slice; the terminating start code is written by CTA 0 at `end_pos`.
__global__ void __launch_bounds__(256) hzg_assemble_kernel(const uint8_t *data, int64_t stride, const int64_t *lens,
                                                           const int64_t *pos, const uint8_t *hdr, int64_t n_slices,
                                                           uint8_t *stream, const uint8_t *pre, int pre_len,
                                                           int64_t params_every, int64_t end_pos) {
    const int64_t i = blockIdx.x;
    if (i >= n_slices) return;
    uint8_t *dst = stream + pos[i];
    if (pre_len > 0 && params_every > 0 && i % params_every == 0) {
        for (int k = threadIdx.x; k < pre_len; k += blockDim.x) dst[k - pre_len] = pre[k];
    }
    if (threadIdx.x < 5) dst[threadIdx.x] = threadIdx.x < 3 ? 0 : (threadIdx.x == 3 ? 1 : hdr[i]);
    const uint8_t *src = data + i * stride;
    for (int64_t k = threadIdx.x; k < lens[i]; k += blockDim.x) dst[5 + k] = src[k];
    if (i == 0 && threadIdx.x < 4) stream[end_pos + threadIdx.x] = threadIdx.x == 3 ? 1 : 0;
}

// Random payload of SURVEY.md §8(d) C1 written with emulation-prevention escaping, one thread per slice.
__global__ void __launch_bounds__(64) hzg_random_payload_kernel(uint32_t config, uint64_t id_base, int64_t n_slices,
                                                                const uint32_t *raw_len, uint8_t *data, int64_t stride,
                                                                int64_t *lens, int *overflow) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slices) return;
    hz_rng r = hz_seed(config, id_base + (uint64_t)s);
    hz_sink sink;
    hz_sink_init(&sink, data + s * stride, stride, 1);
    for (uint32_t i = 0; i < raw_len[s]; i++) {
        uint64_t x = hz_next(&r);
        uint32_t sel = (uint32_t)(x & 15u);
        uint8_t b = sel < 2 ? 0 : (sel < 5 ? (uint8_t)(sel - 1) : (uint8_t)(4 + ((x >> 8) % 252u)));
        hz_put_byte(&sink, b);
    }
    hz_sink_finish(&sink);
    lens[s] = sink.n;
    if (sink.overflow) *overflow = 1;
}

struct HzgTables {
    uint8_t *range_lps, *trans_lps, *trans_mps;
    int8_t *mn_m, *mn_n;
    int *overflow;
    int device;
    int ready;
};
static HzgTables g_tab[2][16];

static int ensure_tables(int v) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return 1;
    HzgTables &t = g_tab[v][dev];
    if (t.ready) return 0;
    const uint8_t *r = v ? hz_range_tab_lps_spec : hz_range_tab_lps_ref;
    const uint8_t *l = v ? hz_trans_idx_lps_spec : hz_trans_idx_lps_ref;
    const uint8_t *m = v ? hz_trans_idx_mps_spec : hz_trans_idx_mps_ref;
    const int8_t *mm = v ? hz_mn_m_spec : hz_mn_m_ref;
    const int8_t *mn = v ? hz_mn_n_spec : hz_mn_n_ref;
    if (cudaMalloc(&t.range_lps, 256) || cudaMalloc(&t.trans_lps, 64) || cudaMalloc(&t.trans_mps, 64) ||
        cudaMalloc(&t.mn_m, 4 * HZ_N_CTX_MAX) || cudaMalloc(&t.mn_n, 4 * HZ_N_CTX_MAX) || cudaMalloc(&t.overflow, 4))
        return 1;
    cudaMemcpy(t.range_lps, r, 256, cudaMemcpyHostToDevice);
    cudaMemcpy(t.trans_lps, l, 64, cudaMemcpyHostToDevice);
    cudaMemcpy(t.trans_mps, m, 64, cudaMemcpyHostToDevice);
    cudaMemcpy(t.mn_m, mm, 4 * HZ_N_CTX_MAX, cudaMemcpyHostToDevice);
    cudaMemcpy(t.mn_n, mn, 4 * HZ_N_CTX_MAX, cudaMemcpyHostToDevice);
    t.ready = 1;
    return 0;
}

extern "C" {
#pragma GCC visibility push(default)

// all pointers are device pointers on the current device; synchronous.  Returns 0 ok, 1 overflow, <0 CUDA error.
int hzg_gen_cabac_slices(int device, uint32_t flags, uint32_t config, uint64_t id_base, int64_t n_slices, const uint16_t *d_ops,
                         const uint32_t *d_n_ops, uint32_t n_active, uint32_t n_ctx, const int32_t *d_qp,
                         const int32_t *d_idc, uint8_t *d_data, int64_t stride, int64_t *d_lens, uint32_t *d_bins,
                         int64_t bins_stride, uint8_t *d_states /* n_slices * n_ctx scratch, ends as final states */) {
    const int v = (flags & HZ_TABLES_SPEC) ? 1 : 0;
    if (cudaSetDevice(device) != cudaSuccess) return -1;
    if (n_ctx > HZG_MAX_CTX || ensure_tables(v)) return -1;
    int dev = 0;
    cudaGetDevice(&dev);
    HzgTables &t = g_tab[v][dev];
    cudaMemset(t.overflow, 0, 4);
    const int blocks = (int)((n_slices + 63) / 64);
    if (blocks)
        hzg_encode_kernel<<<blocks, 64>>>(flags, config, id_base, n_slices, d_ops, d_n_ops, n_active, n_ctx, d_qp,
                                          d_idc, d_data, stride, d_lens, d_bins, bins_stride, d_states, t.range_lps,
                                          t.trans_lps, t.trans_mps, t.mn_m, t.mn_n, t.overflow);
    int ov = 0;
    cudaError_t e = cudaMemcpy(&ov, t.overflow, 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return -(int)e;
    return ov;
}

int hzg_random_payloads(int device, uint32_t config, uint64_t id_base, int64_t n_slices, const uint32_t *d_raw_len,
                        uint8_t *d_data, int64_t stride, int64_t *d_lens) {
    if (cudaSetDevice(device) != cudaSuccess) return -1;
    if (ensure_tables(0)) return -1;
    int dev = 0;
    cudaGetDevice(&dev);
    HzgTables &t = g_tab[0][dev];
    cudaMemset(t.overflow, 0, 4);
    const int blocks = (int)((n_slices + 63) / 64);
    if (blocks) hzg_random_payload_kernel<<<blocks, 64>>>(config, id_base, n_slices, d_raw_len, d_data, stride, d_lens, t.overflow);
    int ov = 0;
    cudaError_t e = cudaMemcpy(&ov, t.overflow, 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return -(int)e;
    return ov;
}

int hzg_assemble(int device, const uint8_t *d_data, int64_t stride, const int64_t *d_lens, const int64_t *d_pos,
                 const uint8_t *d_hdr, int64_t n_slices, uint8_t *d_stream, const uint8_t *d_pre, int pre_len,
                 int64_t params_every, int64_t end_pos) {
    if (cudaSetDevice(device) != cudaSuccess) return -1;
    if (n_slices)
        hzg_assemble_kernel<<<(int)n_slices, 256>>>(d_data, stride, d_lens, d_pos, d_hdr, n_slices, d_stream, d_pre,
                                                    pre_len, params_every, end_pos);
    cudaError_t e = cudaDeviceSynchronize();
    return e == cudaSuccess ? 0 : -(int)e;
}

#pragma GCC visibility pop
}
