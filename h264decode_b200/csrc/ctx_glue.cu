// ctx_glue.cu -- rows I5 / f3: CtxIdx, NewBinarization, initCabac and the mb_type bin-string tables for batches of
// queries, one thread per query (ctx_glue.cuh has the logic and the reference lines).  Scalar host-side logic in the
// reference; kept on the device here like every other entry point (no CPU path), which also keeps the results next to
// the engine's inputs for a future syntax-element-level decode.
#include "common.cuh"
#include "ctx_glue.cuh"

namespace h264b {

struct GlueArgs {
    uint32_t n, op;
    const int64_t *a, *b, *c, *d, *e;  // op-specific int64 inputs
    const int32_t *i0, *i1;            // op-specific int32 inputs
    const uint32_t *u0, *u1;
    const uint8_t *b0;
    int64_t *o64;
    int32_t *o0, *o1;
    uint32_t *ou;
    h264b_binarization *obin;
    const int16_t *mn;  // [5][1024] (m | n << 8) by idc class (ctx_init.cu)
};
enum { kGlueCtxIdx, kGlueBinarization, kGlueInitCabac, kGlueBinString, kGlueMatch };

__global__ void __launch_bounds__(128) ctx_glue_kernel(GlueArgs g) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    switch (g.op) {
        case kGlueCtxIdx:
            g.o64[i] = ctx_idx_ref(g.a[i], g.b[i], g.c[i]);
            break;
        case kGlueBinarization:
            g.obin[i] = new_binarization_ref(g.i0[i], g.i1[i]);
            break;
        case kGlueInitCabac: {
            const int64_t ctx_idx = ctx_idx_ref(g.a[i], g.b[i], g.c[i]);
            int64_t m = 0, n = 0;
            if (ctx_idx >= 0 && ctx_idx <= 39) {  // MNVars has keys 0..39; [ctxIdx][0] = the cabac_init_idc 0 column
                const int16_t v = g.mn[1 * 1024 + ctx_idx];
                m = (int8_t)(v & 0xFF);
                n = (int8_t)((v >> 8) & 0xFF);
            }
            const int64_t qp = (int64_t)((uint64_t)26 + (uint64_t)g.d[i] + (uint64_t)g.e[i]);  // SliceQPy, cabac.go:113
            const int64_t q = qp < 0 ? 0 : (qp > 51 ? 51 : qp);
            int64_t pre = ((m * q) >> 4) + n;
            pre = pre < 1 ? 1 : (pre > 126 ? 126 : pre);
            g.o0[i] = (int32_t)(pre <= 63 ? 63 - pre : pre - 64);
            g.o1[i] = pre <= 63 ? 0 : 1;
            if (g.o64) g.o64[i] = ctx_idx;
            break;
        }
        case kGlueBinString: {
            int32_t len;
            uint32_t bits;
            mb_bin_string_ref(g.i0[i], g.a[i], g.b0[i] != 0, &len, &bits);
            g.o0[i] = len;
            g.ou[i] = bits;
            break;
        }
        case kGlueMatch:
            g.o0[i] = bin_string_match_ref(g.i0[i], g.u0[i], g.i1[i], g.u1[i]);
            break;
    }
}

// stage the host arrays of one call in one device buffer, run the kernel, bring the outputs back
struct Stage {
    h264b_ctx *ctx;
    uint8_t *base;
    size_t used, cap;
    int rc;
    template <typename T>
    const T *in(const T *h, size_t n) {
        T *d = reinterpret_cast<T *>(base + used);
        used += (n * sizeof(T) + 15) / 16 * 16;
        if (!rc && cudaMemcpyAsync(d, h, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) rc = H264B_E_CUDA;
        return d;
    }
    template <typename T>
    T *out(size_t n) {
        T *d = reinterpret_cast<T *>(base + used);
        used += (n * sizeof(T) + 15) / 16 * 16;
        return d;
    }
};
static int stage_begin(h264b_ctx *ctx, size_t bytes, Stage *s) {
    void *d;
    int rc = ensure_dev(ctx, 19, bytes + 512, &d);
    if (rc) return rc;
    s->ctx = ctx;
    s->base = (uint8_t *)d;
    s->used = 0;
    s->cap = bytes + 512;
    s->rc = 0;
    return H264B_OK;
}
static int glue_launch(h264b_ctx *ctx, const GlueArgs &g) {
    ctx_glue_kernel<<<(g.n + 127) / 128, 128, 0, ctx->stream>>>(g);
    H264B_LAUNCH_CHECK(ctx, "ctx_glue_kernel");
    return H264B_OK;
}
template <typename T>
static int fetch(h264b_ctx *ctx, T *h, const T *d, size_t n) {
    H264B_CUDA(ctx, cudaMemcpyAsync(h, d, n * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    return H264B_OK;
}

}  // namespace h264b

using namespace h264b;

#define GLUE_ENTER(ctx, n, cond)                                                            \
    if (!(ctx)) return H264B_E_INVALID;                                                     \
    cudaSetDevice((ctx)->device);                                                           \
    if (!(n)) return H264B_OK;                                                              \
    if (!(cond)) return set_error((ctx), H264B_E_INVALID, "%s: null pointer", __func__);    \
    Stage st;                                                                               \
    GlueArgs g;                                                                             \
    memset(&g, 0, sizeof(g));                                                               \
    g.n = (n);                                                                              \
    int rc
#define GLUE_RC(x)           \
    do {                     \
        if ((rc = (x))) return rc; \
    } while (0)

extern "C" {

int32_t h264b_ctx_idx(h264b_ctx *ctx, uint32_t n, const int64_t *bin_idx, const int64_t *max_bin_idx_ctx,
                      const int64_t *ctx_idx_offset, int64_t *out) {
    GLUE_ENTER(ctx, n, bin_idx && max_bin_idx_ctx && ctx_idx_offset && out);
    GLUE_RC(stage_begin(ctx, (size_t)n * 32 + 64, &st));
    g.op = kGlueCtxIdx;
    g.a = st.in(bin_idx, n);
    g.b = st.in(max_bin_idx_ctx, n);
    g.c = st.in(ctx_idx_offset, n);
    g.o64 = st.out<int64_t>(n);
    if (st.rc) return set_error(ctx, st.rc, "ctx_idx: copy failed");
    GLUE_RC(glue_launch(ctx, g));
    GLUE_RC(fetch(ctx, out, g.o64, n));
    H264B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return H264B_OK;
}

int32_t h264b_new_binarization(h264b_ctx *ctx, uint32_t n, const int32_t *syntax_element, const int32_t *slice_type_name,
                               h264b_binarization *out) {
    GLUE_ENTER(ctx, n, syntax_element && slice_type_name && out);
    GLUE_RC(stage_begin(ctx, (size_t)n * (8 + sizeof(h264b_binarization)) + 64, &st));
    g.op = kGlueBinarization;
    g.i0 = st.in(syntax_element, n);
    g.i1 = st.in(slice_type_name, n);
    g.obin = st.out<h264b_binarization>(n);
    if (st.rc) return set_error(ctx, st.rc, "new_binarization: copy failed");
    GLUE_RC(glue_launch(ctx, g));
    GLUE_RC(fetch(ctx, out, g.obin, n));
    H264B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return H264B_OK;
}

int32_t h264b_init_cabac(h264b_ctx *ctx, uint32_t flags, uint32_t n, const int64_t *bin_idx, const int64_t *max_prefix,
                         const int64_t *off_prefix, const int64_t *pic_init_qp_minus26, const int64_t *slice_qp_delta,
                         int32_t *p_state_idx, int32_t *val_mps, int64_t *ctx_idx_out) {
    GLUE_ENTER(ctx, n, bin_idx && max_prefix && off_prefix && pic_init_qp_minus26 && slice_qp_delta && p_state_idx && val_mps);
    GLUE_RC(stage_begin(ctx, (size_t)n * 56 + 128, &st));
    g.op = kGlueInitCabac;
    g.a = st.in(bin_idx, n);
    g.b = st.in(max_prefix, n);
    g.c = st.in(off_prefix, n);
    g.d = st.in(pic_init_qp_minus26, n);
    g.e = st.in(slice_qp_delta, n);
    g.o0 = st.out<int32_t>(n);
    g.o1 = st.out<int32_t>(n);
    g.o64 = st.out<int64_t>(n);
    g.mn = ctx->d_mn[(flags & H264B_TABLES_SPEC) ? 1 : 0];
    if (st.rc) return set_error(ctx, st.rc, "init_cabac: copy failed");
    GLUE_RC(glue_launch(ctx, g));
    GLUE_RC(fetch(ctx, p_state_idx, g.o0, n));
    GLUE_RC(fetch(ctx, val_mps, g.o1, n));
    if (ctx_idx_out) GLUE_RC(fetch(ctx, ctx_idx_out, g.o64, n));
    H264B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return H264B_OK;
}

int32_t h264b_mb_bin_string(h264b_ctx *ctx, uint32_t n, const int32_t *slice_type_name, const int64_t *mb_type,
                            const uint8_t *sub_mb, int32_t *len, uint32_t *bits) {
    GLUE_ENTER(ctx, n, slice_type_name && mb_type && sub_mb && len && bits);
    GLUE_RC(stage_begin(ctx, (size_t)n * 21 + 128, &st));
    g.op = kGlueBinString;
    g.i0 = st.in(slice_type_name, n);
    g.a = st.in(mb_type, n);
    g.b0 = st.in(sub_mb, n);
    g.o0 = st.out<int32_t>(n);
    g.ou = st.out<uint32_t>(n);
    if (st.rc) return set_error(ctx, st.rc, "mb_bin_string: copy failed");
    GLUE_RC(glue_launch(ctx, g));
    GLUE_RC(fetch(ctx, len, g.o0, n));
    GLUE_RC(fetch(ctx, bits, g.ou, n));
    H264B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return H264B_OK;
}

int32_t h264b_bin_string_match(h264b_ctx *ctx, uint32_t n, const int32_t *bin_len, const uint32_t *bin_bits,
                               const int32_t *n_bits, const uint32_t *bits, int32_t *out) {
    GLUE_ENTER(ctx, n, bin_len && bin_bits && n_bits && bits && out);
    GLUE_RC(stage_begin(ctx, (size_t)n * 20 + 128, &st));
    g.op = kGlueMatch;
    g.i0 = st.in(bin_len, n);
    g.u0 = st.in(bin_bits, n);
    g.i1 = st.in(n_bits, n);
    g.u1 = st.in(bits, n);
    g.o0 = st.out<int32_t>(n);
    if (st.rc) return set_error(ctx, st.rc, "bin_string_match: copy failed");
    GLUE_RC(glue_launch(ctx, g));
    GLUE_RC(fetch(ctx, out, g.o0, n));
    H264B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return H264B_OK;
}

}  // extern "C"
