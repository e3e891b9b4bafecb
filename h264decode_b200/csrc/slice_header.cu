// slice_header.cu -- "next" row f1: the slice-header walk of NewSliceContext (h264/slice.go:835-1048) for all slice
// NAL units at once, one thread per slice (the walk is a serial Exp-Golomb parse of a few dozen bits; slices are
// independent).  The parse itself is slice_header.cuh (shared with the CPU emulation of tests/).
#include "common.cuh"
#include "param_sets.cuh"

namespace h264b {

struct SliceHeaderArgs {
    h264b_param_sets ps;
    const uint8_t *bytes;
    uint64_t total_bytes;
    const uint64_t *off;
    const uint32_t *len;
    const uint8_t *nal_type, *nal_ref_idc;
    const h264b_nal *nals;
    const uint32_t *slice_nal;
    uint32_t n_slices;
    const uint32_t *n_slices_dev;  // actual count on the device (n_slices is then the bound), or NULL
    h264b_slice_header *out;
    // H264B_STREAM_PARAM_SETS: the stream's own parameter sets (param_sets.cu) instead of `ps`
    const h264b_sps *sps;
    const h264b_pps *pps;
    const uint32_t *sps_nal, *pps_nal;  // their NAL ordinals, ascending
    const uint32_t *ps_counts;          // [0] SPS kept, [1] PPS kept
    int32_t *slice_sps, *slice_pps;     // out: the sets slice s used (-1: none, -2: the initial set)
    const h264b_sps *initial_sps;       // the sets in force when the batch begins (batched ingest), or NULL
    const h264b_pps *initial_pps;
};

// index of the last entry below `key` in an ascending list, -1 if there is none
__device__ __forceinline__ int32_t last_below(const uint32_t *list, uint32_t n, uint32_t key) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (list[mid] < key) lo = mid + 1;
        else hi = mid;
    }
    return (int32_t)lo - 1;
}

__global__ void __launch_bounds__(128) slice_header_kernel(SliceHeaderArgs a) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.n_slices || (a.n_slices_dev && s >= *a.n_slices_dev)) return;
    uint64_t off, len;
    uint32_t type, ref_idc;
    if (a.nals) {
        const h264b_nal u = a.nals[a.slice_nal[s]];
        off = u.rbsp_off;
        len = u.rbsp_len;
        type = u.type;
        ref_idc = u.ref_idc;
    } else {
        off = a.off[s];
        len = a.len[s];
        type = a.nal_type[s];
        ref_idc = a.nal_ref_idc[s];
    }
    if (off > a.total_bytes) off = a.total_bytes;
    if (len > a.total_bytes - off) len = a.total_bytes - off;
    h264b_slice_header h;
    if (a.sps) {
        // handleConnection (server.go:147-162): a slice belongs to the last VideoStream, i.e. the last SPS before it, and
        // uses the PPS stored there last, i.e. the last PPS behind that SPS.  No SPS: VideoStreams[-1] panics; no PPS:
        // NewSliceContext dereferences nil; a parameter set NewSPS / NewPPS panicked on never got that far.
        const uint32_t k = a.slice_nal[s];
        int32_t si = last_below(a.sps_nal, a.ps_counts[0], k);
        int32_t pi = last_below(a.pps_nal, a.ps_counts[1], k);
        const h264b_sps *sps = nullptr;
        const h264b_pps *pps = nullptr;
        if (si >= 0) {  // a VideoStream of this batch: its PPS must come behind its SPS
            sps = a.sps + si;
            if (pi >= 0 && a.pps_nal[pi] > a.sps_nal[si]) pps = a.pps + pi;
            else pi = -1;
        } else if (a.initial_sps) {  // the VideoStream the batch inherited: a PPS of this batch replaces the inherited one
            si = -2;
            sps = a.initial_sps;
            if (pi >= 0) pps = a.pps + pi;
            else if (a.initial_pps) pps = a.initial_pps, pi = -2;
        } else {
            pi = -1;
        }
        a.slice_sps[s] = si;
        a.slice_pps[s] = pi;
        if (!sps || !pps || sps->status != H264B_SH_OK || pps->status != H264B_SH_OK) {
            h = h264b_slice_header{};
            h.status = H264B_SH_PANIC;
        } else {
            parse_slice_header_record(make_param_sets(*sps, *pps), type, ref_idc, a.bytes + off, len, &h);
        }
    } else {
        parse_slice_header_record(a.ps, type, ref_idc, a.bytes + off, len, &h);
    }
    a.out[s] = h;
}

// The CABAC stage's per-slice inputs from the parsed headers: (SliceQPY, cabac_init_idc) for the K4 rule and the start
// of the CABAC data = the first byte boundary behind the header (cabac_alignment_one_bit, slice.go:583-587).
__global__ void __launch_bounds__(256) slice_params_kernel(const h264b_slice_header *hdr, const uint32_t *n_slices_dev,
                                                           uint32_t max_slices, uint64_t *off, uint32_t *len,
                                                           h264b_slice_qp *qp) {
    const uint32_t n = *n_slices_dev < max_slices ? *n_slices_dev : max_slices;
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        const h264b_slice_header &h = hdr[s];
        h264b_slice_qp p;
        p.slice_qp_y = 0;
        p.cabac_init_idc = -1;
        if (h.status != H264B_SH_OK) {
            len[s] = 0;  // nothing to decode: the engine flags the slice (H264B_F_OVERRUN)
        } else {
            const int64_t q = h.slice_qp_y < -1000000 ? -1000000 : (h.slice_qp_y > 1000000 ? 1000000 : h.slice_qp_y);
            p.slice_qp_y = (int32_t)q;  // (PreCtxState clips to 0..51 anyway)
            const int64_t t5 = (h.slice_type >= 0 && h.slice_type <= 9) ? h.slice_type % 5 : -1;
            const bool intra = t5 == 2 || t5 == 4;  // I, SI: no cabac_init_idc in the header -> the I / SI column
            const int64_t idc = h.cabac_init_idc;
            p.cabac_init_idc = intra ? -1 : (int32_t)(idc < -2 ? -2 : (idc > 1000 ? 1000 : idc));
            const uint64_t skip = (h.header_bits + 7u) / 8u;
            const uint64_t l = len[s];
            off[s] += skip < l ? skip : l;
            len[s] = (uint32_t)(skip < l ? l - skip : 0u);
        }
        qp[s] = p;
    }
}

int launch_slice_headers(h264b_ctx *ctx, const SliceHeaderArgs &a) {
    if (!a.n_slices) return H264B_OK;
    slice_header_kernel<<<(a.n_slices + 127) / 128, 128, 0, ctx->stream>>>(a);
    H264B_LAUNCH_CHECK(ctx, "slice_header_kernel");
    return H264B_OK;
}

int launch_stream_slice_headers(h264b_ctx *ctx, const h264b_param_sets *params, const uint8_t *d_rbsp, uint64_t total,
                                const h264b_nal *d_nals, const uint32_t *d_slice_nal, const uint32_t *d_n_slices,
                                uint32_t max_slices, h264b_slice_header *d_hdr, uint64_t *d_off, uint32_t *d_len,
                                h264b_slice_qp *d_qp, const StreamParamSets *sp) {
    if (!max_slices) return H264B_OK;
    SliceHeaderArgs a;
    a.ps = params ? *params : h264b_param_sets{};
    a.sps = sp ? sp->sps : nullptr;
    a.pps = sp ? sp->pps : nullptr;
    a.sps_nal = sp ? sp->sps_nal : nullptr;
    a.pps_nal = sp ? sp->pps_nal : nullptr;
    a.ps_counts = sp ? sp->counts : nullptr;
    a.slice_sps = sp ? sp->slice_sps : nullptr;
    a.slice_pps = sp ? sp->slice_pps : nullptr;
    a.initial_sps = sp ? sp->initial_sps : nullptr;
    a.initial_pps = sp ? sp->initial_pps : nullptr;
    a.bytes = d_rbsp;
    a.total_bytes = total;
    a.off = nullptr;
    a.len = nullptr;
    a.nal_type = a.nal_ref_idc = nullptr;
    a.nals = d_nals;
    a.slice_nal = d_slice_nal;
    a.n_slices = max_slices;
    a.n_slices_dev = d_n_slices;
    a.out = d_hdr;
    int rc = launch_slice_headers(ctx, a);
    if (rc) return rc;
    int blocks = (int)((max_slices + 255) / 256);
    if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
    slice_params_kernel<<<blocks, 256, 0, ctx->stream>>>(d_hdr, d_n_slices, max_slices, d_off, d_len, d_qp);
    H264B_LAUNCH_CHECK(ctx, "slice_params_kernel");
    return H264B_OK;
}

}  // namespace h264b

using namespace h264b;

extern "C" int32_t h264b_slice_headers_dev(h264b_ctx *ctx, const h264b_param_sets *params, const uint8_t *d_bytes,
                                           uint64_t total_bytes, const uint64_t *d_off, const uint32_t *d_len,
                                           const uint8_t *d_nal_type, const uint8_t *d_nal_ref_idc,
                                           const h264b_nal *d_nals, const uint32_t *d_slice_nal, uint32_t n_slices,
                                           h264b_slice_header *d_out) {
    if (!ctx) return H264B_E_INVALID;
    cudaSetDevice(ctx->device);
    if (!params || (n_slices && (!d_bytes || !d_out)))
        return set_error(ctx, H264B_E_INVALID, "slice_headers: null pointer");
    if (n_slices && !d_nals && (!d_off || !d_len || !d_nal_type || !d_nal_ref_idc))
        return set_error(ctx, H264B_E_INVALID, "slice_headers: need either the NAL index or off/len/type/ref_idc");
    if (n_slices && d_nals && !d_slice_nal) return set_error(ctx, H264B_E_INVALID, "slice_headers: slice_nal is null");
    SliceHeaderArgs a;
    a.ps = *params;
    a.bytes = d_bytes;
    a.total_bytes = total_bytes;
    a.off = d_off;
    a.len = d_len;
    a.nal_type = d_nal_type;
    a.nal_ref_idc = d_nal_ref_idc;
    a.nals = d_nals;
    a.slice_nal = d_slice_nal;
    a.n_slices = n_slices;
    a.n_slices_dev = nullptr;
    a.out = d_out;
    a.sps = nullptr;
    a.pps = nullptr;
    a.sps_nal = a.pps_nal = a.ps_counts = nullptr;
    a.slice_sps = a.slice_pps = nullptr;
    a.initial_sps = nullptr;
    a.initial_pps = nullptr;
    return launch_slice_headers(ctx, a);
}

extern "C" int32_t h264b_slice_headers(h264b_ctx *ctx, const h264b_param_sets *params, const uint8_t *bytes,
                                       uint64_t total_bytes, const uint64_t *off, const uint32_t *len,
                                       const uint8_t *nal_type, const uint8_t *nal_ref_idc, uint32_t n_slices,
                                       h264b_slice_header *out) {
    if (!ctx) return H264B_E_INVALID;
    cudaSetDevice(ctx->device);
    if (!n_slices) return H264B_OK;
    if (!params || !bytes || !off || !len || !nal_type || !nal_ref_idc || !out)
        return set_error(ctx, H264B_E_INVALID, "slice_headers: null pointer");
    void *d_bytes, *d_off, *d_len, *d_type, *d_ref, *d_out;
    int rc;
    if ((rc = ensure_dev(ctx, 0, total_bytes + 64, &d_bytes))) return rc;
    if ((rc = ensure_dev(ctx, 5, (size_t)n_slices * 8, &d_off))) return rc;
    if ((rc = ensure_dev(ctx, 6, (size_t)n_slices * 4, &d_len))) return rc;
    if ((rc = ensure_dev(ctx, 17, (size_t)n_slices * 2, &d_type))) return rc;
    d_ref = (uint8_t *)d_type + n_slices;
    if ((rc = ensure_dev(ctx, 18, (size_t)n_slices * sizeof(h264b_slice_header), &d_out))) return rc;
    H264B_CUDA(ctx, cudaMemcpyAsync(d_bytes, bytes, total_bytes, cudaMemcpyHostToDevice, ctx->stream));
    H264B_CUDA(ctx, cudaMemcpyAsync(d_off, off, (size_t)n_slices * 8, cudaMemcpyHostToDevice, ctx->stream));
    H264B_CUDA(ctx, cudaMemcpyAsync(d_len, len, (size_t)n_slices * 4, cudaMemcpyHostToDevice, ctx->stream));
    H264B_CUDA(ctx, cudaMemcpyAsync(d_type, nal_type, n_slices, cudaMemcpyHostToDevice, ctx->stream));
    H264B_CUDA(ctx, cudaMemcpyAsync(d_ref, nal_ref_idc, n_slices, cudaMemcpyHostToDevice, ctx->stream));
    rc = h264b_slice_headers_dev(ctx, params, (const uint8_t *)d_bytes, total_bytes, (const uint64_t *)d_off,
                                 (const uint32_t *)d_len, (const uint8_t *)d_type, (const uint8_t *)d_ref, nullptr, nullptr,
                                 n_slices, (h264b_slice_header *)d_out);
    if (rc) return rc;
    H264B_CUDA(ctx, cudaMemcpyAsync(out, d_out, (size_t)n_slices * sizeof(h264b_slice_header), cudaMemcpyDeviceToHost,
                                    ctx->stream));
    H264B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return H264B_OK;
}
