// mb_type.cu -- row f3 of SURVEY.md section 8: mb_type decoded as a syntax element, one slice per lane, every lane on
// its own, data-dependent sequence of ops.
//
// Reference: the walk sketched at h264/slice.go:639-672 -- NewBinarization("MbType") h264/cabac.go:340-427, per bin
// CtxIdx h264/cabac.go:557-758, the bin strings binIdxMbMap h264/cabac.go:180-303 and IsBinStringMatch :429-436 --
// composed with DecodeDecision / DecodeTerminate (cabac.go:486-553) the way include/h264b200.h states.
//
// The binarisation is a trie: a node names the context of the next bin (or DecodeTerminate) and its two children, a leaf
// the mb_type.  The host builds it once per context from the same table functions the scalar glue uses (ctx_glue.cuh:
// mb_bin_string_ref, ctx_idx_ref), the kernel keeps it in shared memory.  Lanes of a warp sit at different nodes, on
// different contexts, and finish their elements after different numbers of bins: none of cabac_decode_kernel's
// warp-uniform structure applies (two ballots per 32 ops, rows of 32 equal contexts, the op kinds as masks).  What is
// left is the per-lane engine (LaneDecoder) stepped warp-synchronously, one bin per lane and step.
#include <vector>

#include "cabac_lane.cuh"
#include "common.cuh"
#include "ctx_glue.cuh"

namespace h264b {

// node: leaf  = 0x80000000 | I_PCM << 8 | mb_type   (I_PCM: the walk of the slice ends with this element)
//       inner = terminate << 30 | add_prev << 29 | ctxIdx << 16 | child(bin 1) << 8 | child(bin 0); child 0xFF: no mb_type
//       has this bin string (the reference's loop would never end)
constexpr uint32_t kLeaf = 0x80000000u, kNoChild = 0xFFu;
constexpr int kTrieMax = 128;

struct TrieBuilder {
    std::vector<uint32_t> nodes;
    // the I-slice table (26 strings) below `offset`; leaves carry base + type
    int build_i_table(int64_t offset, int base, bool root_adds_prev) {
        struct Str {
            int32_t len;
            uint32_t bits;
        } str[26];
        for (int t = 0; t < 26; t++) mb_bin_string_ref(2, t, false, &str[t].len, &str[t].bits);
        return build(offset, base, root_adds_prev, [&](int n, uint32_t bits, int *type) {
            int state = 0;  // 0: no string has this prefix, 1: proper prefix, 2: exact
            for (int t = 0; t < 26; t++) {
                if (str[t].len < n) continue;
                const uint32_t m = (1u << n) - 1u;
                if ((str[t].bits & m) != (bits & m)) continue;
                if (str[t].len == n) {
                    *type = t;
                    return 2;
                }
                state = 1;
            }
            return state;
        });
    }
    template <class Classify>
    int build(int64_t offset, int base, bool root_adds_prev, const Classify &classify) {
        return node(offset, base, root_adds_prev, 0, 0u, classify);
    }
    template <class Classify>
    int node(int64_t offset, int base, bool root_adds_prev, int n, uint32_t bits, const Classify &classify) {
        const int me = (int)nodes.size();
        nodes.push_back(0);
        // context of bin n given the bins so far: CtxIdx where it answers, the clause it points to where it does not
        int64_t inc = ctx_idx_ref(n, 0, offset);
        bool add_prev = false, term = false;
        if (inc == kNaCtxId) {
            if (offset == 3 && n == 0) inc = 0, add_prev = root_adds_prev;
            else if (offset == 3 && n == 4) inc = ((bits >> 3) & 1u) ? 5 : 6;
            else if (offset == 3 && n == 5) inc = ((bits >> 3) & 1u) ? 6 : 7;
            else if (offset == 17 && n == 4) inc = ((bits >> 3) & 1u) ? 2 : 3;
        }
        if (inc == 276) term = true;
        int64_t ctx = inc == kNaCtxId ? 0 : offset + inc;
        uint32_t child[2];
        for (uint32_t b = 0; b < 2; b++) {
            int type = 0;
            const uint32_t nb = bits | (b << n);
            const int st = classify(n + 1, nb, &type);
            if (st == 2) {
                child[b] = (uint32_t)nodes.size();
                nodes.push_back(kLeaf | (type == 25 ? 0x100u : 0u) | (uint32_t)(base + type));  // (I table only: 25 = I_PCM)
            } else if (st == 1) {
                child[b] = (uint32_t)node(offset, base, root_adds_prev, n + 1, nb, classify);
            } else {
                child[b] = kNoChild;
            }
        }
        nodes[me] = (term ? 1u << 30 : 0u) | (add_prev ? 1u << 29 : 0u) | ((uint32_t)ctx << 16) | (child[1] << 8) | child[0];
        return me;
    }
};

// [0]: I slices; [1]: P / SP slices (prefix on offset 14, the intra suffix on offset 17)
static void build_tries(uint32_t out[2][kTrieMax]) {
    memset(out, 0, sizeof(uint32_t) * 2 * kTrieMax);
    {
        TrieBuilder b;
        b.build_i_table(3, 0, true);
        for (size_t k = 0; k < b.nodes.size() && k < (size_t)kTrieMax; k++) out[0][k] = b.nodes[k];
    }
    {
        TrieBuilder b;
        // prefix: bin 0 on ctx 14; 0 -> bins 1, 2 (ctx 15; 16 or 17) -> P types 0..3; 1 -> the I table on offset 17, + 5
        b.nodes.assign(4, 0);  // 0: root, 1: after {0}, 2: after {0,0}, 3: after {0,1}
        const int suffix = b.build_i_table(17, 5, false);
        auto leaf = [&](int type) {
            b.nodes.push_back(kLeaf | (uint32_t)type);
            return (uint32_t)b.nodes.size() - 1u;
        };
        // binIdxMbMap["P"]: 0 {0,0,0}  1 {0,1,1}  2 {0,1,0}  3 {0,0,1}
        const uint32_t l000 = leaf(0), l001 = leaf(3), l010 = leaf(2), l011 = leaf(1);
        b.nodes[0] = (14u << 16) | ((uint32_t)suffix << 8) | 1u;
        b.nodes[1] = (15u << 16) | (3u << 8) | 2u;
        b.nodes[2] = (16u << 16) | (l001 << 8) | l000;  // b1 != 1: ctxIdxInc 2
        b.nodes[3] = (17u << 16) | (l011 << 8) | l010;  // b1 == 1: ctxIdxInc 3
        for (size_t k = 0; k < b.nodes.size() && k < (size_t)kTrieMax; k++) out[1][k] = b.nodes[k];
    }
}

struct MbArgs {
    h264b_mb_type_job j;
    const uint64_t *tab;
    const uint8_t *lut;
    const uint32_t *trie;  // [2][kTrieMax]
};

__device__ __forceinline__ int mb_idc_class(int idc) { return (idc >= -1 && idc <= 2) ? idc + 1 : 4; }

constexpr int kMbWarps = 4;

__global__ void __launch_bounds__(kMbWarps * 32) mb_type_kernel(MbArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint64_t *s_tab = reinterpret_cast<uint64_t *>(smem);            // 128 x 8 B
    uint32_t *s_trie = reinterpret_cast<uint32_t *>(smem + 1024);    // 2 x kTrieMax
    uint8_t *s_state_all = smem + 1024 + 2 * kTrieMax * 4;           // [warp][n_ctx][32]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 128; i += blockDim.x) s_tab[i] = a.tab[i];
    for (int i = tid; i < 2 * kTrieMax; i += blockDim.x) s_trie[i] = a.trie[i];
    __syncthreads();
    const h264b_mb_type_job &j = a.j;
    const uint32_t n_ctx = j.n_ctx;
    uint8_t *s_state = s_state_all + (size_t)warp * n_ctx * 32;
    const uint32_t slice = (blockIdx.x * kMbWarps + warp) * 32 + lane;
    const bool own = slice < j.n_slices;
    const uint32_t src = own ? slice : j.n_slices - 1;  // (lanes past the end shadow the last slice and store nothing)
    uint64_t off = j.off[src];
    uint32_t len = j.len[src];
    bool bad_range = false;
    if (off > j.total_bytes) off = j.total_bytes, bad_range = true;
    if ((uint64_t)len > j.total_bytes - off) len = (uint32_t)(j.total_bytes - off), bad_range = true;
    {
        const uint8_t *st;
        if (j.init_states) {
            st = j.init_states + (size_t)src * n_ctx;
        } else {
            const h264b_slice_qp p = j.qp[src];
            const int q = p.slice_qp_y < 0 ? 0 : (p.slice_qp_y > 51 ? 51 : p.slice_qp_y);
            st = a.lut + ((size_t)mb_idc_class(p.cabac_init_idc) * 52 + q) * 1024;
        }
        for (uint32_t c = 0; c < n_ctx; c++) s_state[c * 32 + lane] = st[c];
    }
    __syncwarp();
    LaneDecoder eng = {};
    eng.init(j.bytes, j.total_bytes, off, true);
    const uint32_t *trie = s_trie + (j.slice_kind[src] ? kTrieMax : 0);
    uint32_t want = j.n_mb[src];
    if (want > j.n_mb_max) want = j.n_mb_max;
    uint8_t *out = j.mb_type + (size_t)src * j.n_mb_max;
    uint32_t done = 0, n_bins = 0, idx = 0, prev = 0;
    bool active = want > 0;
    // the reference's reader panics when a bin needs bits past the slice's last byte: the slice stops before that bin
    uint64_t bits_ok = 0;       // bits_read after the last bin that fitted
    int64_t r_ok = 0, o_ok = 0;
    bool overrun = false;
    auto snapshot = [&]() { bits_ok = eng.bits_read(), r_ok = eng.cod_i_range(), o_ok = eng.cod_i_offset(); };
    snapshot();
    if (bits_ok > 8ull * len) overrun = true, active = false;
    while (__any_sync(0xFFFFFFFFu, active)) {
        if (active) {
            if (eng.must_refill()) eng.refill_if_room();
            const uint32_t node = trie[idx];
            uint32_t ctx = ((node >> 16) & 0x1FFFu) + (((node >> 29) & 1u) ? prev : 0u);
            if (ctx >= n_ctx) ctx = 0;
            uint8_t *sp = s_state + ctx * 32 + lane;
            const uint8_t s0 = *sp;
            uint32_t bin;
            if ((node >> 30) & 1u) {
                bin = eng.terminate();
            } else {
                uint8_t ns;
                bin = eng.decision(s_tab[s0 & 127u], &ns);
                *sp = ns;
            }
            if (eng.bits_read() > 8ull * len) {  // this bin ran off the data: everything as it was before it
                if (!((node >> 30) & 1u)) *sp = s0;
                overrun = true;
                active = false;
            } else {
                snapshot();
                n_bins++;
                const uint32_t child = bin ? (node >> 8) & 0xFFu : node & 0xFFu;
                if (child == kNoChild) {
                    active = false;  // no mb_type has this bin string
                } else {
                    const uint32_t nn = trie[child];
                    if (nn & kLeaf) {
                        const uint32_t t = nn & 0xFFu;
                        if (own) out[done] = (uint8_t)t;
                        done++;
                        prev = t != 0u ? 1u : 0u;
                        idx = 0;
                        if (done >= want || (nn & 0x100u)) active = false;  // (I_PCM)
                    } else {
                        idx = child;
                    }
                }
            }
        }
    }
    if (own) {
        h264b_mb_final f;
        f.cod_i_range = r_ok;
        f.cod_i_offset = o_ok;
        f.bits_read = bits_ok;
        f.flags = (overrun || bad_range) ? H264B_F_OVERRUN : 0u;
        f.n_bins = n_bins;
        f.n_mb = done;
        f.reserved = 0;
        j.final[slice] = f;
        if (j.final_states) {
            uint8_t *dst = j.final_states + (size_t)slice * n_ctx;
            for (uint32_t c = 0; c < n_ctx; c++) dst[c] = s_state[c * 32 + lane];
        }
    }
}

static int ensure_trie(h264b_ctx *ctx, const uint32_t **d_trie) {
    void *d;
    const bool fresh = ctx->d_buf_bytes[0][19] == 0;
    const int saved_bank = ctx->bank;
    ctx->bank = 0;
    int rc = ensure_dev(ctx, 19, sizeof(uint32_t) * 2 * kTrieMax, &d);
    ctx->bank = saved_bank;
    if (rc) return rc;
    if (fresh) {
        uint32_t h[2][kTrieMax];
        build_tries(h);
        H264B_CUDA(ctx, cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice));
    }
    *d_trie = (const uint32_t *)d;
    return H264B_OK;
}

static int launch_mb_type(h264b_ctx *ctx, const h264b_mb_type_job *job) {
    TraceRange trace_range("h264b:mb_type_decode");
    const h264b_mb_type_job &j = *job;
    if (j.n_ctx < 21 || j.n_ctx > 1024) return set_error(ctx, H264B_E_INVALID, "mb_type: n_ctx must be 21..1024");
    if (!j.n_slices) return H264B_OK;
    if (!j.bytes || !j.off || !j.len || !j.slice_kind || !j.n_mb || !j.mb_type || !j.final || (!j.qp && !j.init_states))
        return set_error(ctx, H264B_E_INVALID, "mb_type: null pointer in job");
    if ((uintptr_t)j.bytes & 3) return set_error(ctx, H264B_E_INVALID, "mb_type: bytes must be 4-byte aligned");
    MbArgs a;
    a.j = j;
    const int v = (j.flags & H264B_TABLES_SPEC) ? 1 : 0;
    a.tab = ctx->d_cabac_tab[v];
    a.lut = ctx->d_state_lut[v];
    int rc = ensure_trie(ctx, &a.trie);
    if (rc) return rc;
    const size_t smem = 1024 + 2 * kTrieMax * 4 + (size_t)kMbWarps * j.n_ctx * 32;
    H264B_CUDA(ctx, cudaFuncSetAttribute(mb_type_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t blocks = (j.n_slices + kMbWarps * 32 - 1) / (kMbWarps * 32);
    mb_type_kernel<<<blocks, kMbWarps * 32, smem, ctx->stream>>>(a);
    H264B_LAUNCH_CHECK(ctx, "mb_type_kernel");
    return H264B_OK;
}

}  // namespace h264b

using namespace h264b;

extern "C" int32_t h264b_mb_type_decode_dev(h264b_ctx *ctx, const h264b_mb_type_job *job) {
    if (!ctx || !job) return H264B_E_INVALID;
    cudaSetDevice(ctx->device);
    return launch_mb_type(ctx, job);
}

extern "C" int32_t h264b_mb_type_decode(h264b_ctx *ctx, const h264b_mb_type_job *job) {
    if (!ctx || !job) return H264B_E_INVALID;
    cudaSetDevice(ctx->device);
    const h264b_mb_type_job &j = *job;
    if (!j.n_slices) return H264B_OK;
    if (!j.bytes || !j.off || !j.len || !j.slice_kind || !j.n_mb || !j.mb_type || !j.final || (!j.qp && !j.init_states))
        return set_error(ctx, H264B_E_INVALID, "mb_type: null pointer in job");
    for (uint32_t s = 0; s < j.n_slices; s++)
        if (j.off[s] > j.total_bytes || j.len[s] > j.total_bytes - j.off[s])
            return set_error(ctx, H264B_E_INVALID, "mb_type: slice %u lies outside the buffer", s);
    const size_t ns = j.n_slices;
    struct Buf {
        void *d = nullptr;
    };
    void *d_bytes, *d_off, *d_len, *d_kind, *d_nmb, *d_qp = nullptr, *d_init = nullptr, *d_out, *d_fin, *d_fst = nullptr;
    int rc;
#define DEV(slot, bytes, out)                         \
    if ((rc = ensure_dev(ctx, slot, bytes, out))) return rc;
    DEV(0, j.total_bytes + 64, &d_bytes);
    DEV(5, ns * 8, &d_off);
    DEV(6, ns * 4, &d_len);
    DEV(7, ns, &d_kind);
    DEV(8, ns * 4, &d_nmb);
    if (j.init_states) {
        DEV(10, ns * j.n_ctx, &d_init);
    } else {
        DEV(9, ns * sizeof(h264b_slice_qp), &d_qp);
    }
    DEV(11, ns * (size_t)j.n_mb_max + 16, &d_out);
    DEV(12, ns * sizeof(h264b_mb_final), &d_fin);
    if (j.final_states) DEV(14, ns * j.n_ctx, &d_fst);
#undef DEV
    cudaStream_t st = ctx->stream;
    H264B_CUDA(ctx, cudaMemsetAsync((uint8_t *)d_bytes + (j.total_bytes & ~(uint64_t)3), 0, 64 - (j.total_bytes & 3), st));
    H264B_CUDA(ctx, cudaMemcpyAsync(d_bytes, j.bytes, j.total_bytes, cudaMemcpyHostToDevice, st));
    H264B_CUDA(ctx, cudaMemcpyAsync(d_off, j.off, ns * 8, cudaMemcpyHostToDevice, st));
    H264B_CUDA(ctx, cudaMemcpyAsync(d_len, j.len, ns * 4, cudaMemcpyHostToDevice, st));
    H264B_CUDA(ctx, cudaMemcpyAsync(d_kind, j.slice_kind, ns, cudaMemcpyHostToDevice, st));
    H264B_CUDA(ctx, cudaMemcpyAsync(d_nmb, j.n_mb, ns * 4, cudaMemcpyHostToDevice, st));
    if (j.init_states) H264B_CUDA(ctx, cudaMemcpyAsync(d_init, j.init_states, ns * j.n_ctx, cudaMemcpyHostToDevice, st));
    else H264B_CUDA(ctx, cudaMemcpyAsync(d_qp, j.qp, ns * sizeof(h264b_slice_qp), cudaMemcpyHostToDevice, st));
    H264B_CUDA(ctx, cudaMemsetAsync(d_out, 0, ns * (size_t)j.n_mb_max, st));
    h264b_mb_type_job dj = j;
    dj.bytes = (const uint8_t *)d_bytes;
    dj.off = (const uint64_t *)d_off;
    dj.len = (const uint32_t *)d_len;
    dj.slice_kind = (const uint8_t *)d_kind;
    dj.n_mb = (const uint32_t *)d_nmb;
    dj.qp = (const h264b_slice_qp *)d_qp;
    dj.init_states = (const uint8_t *)d_init;
    dj.mb_type = (uint8_t *)d_out;
    dj.final = (h264b_mb_final *)d_fin;
    dj.final_states = (uint8_t *)d_fst;
    rc = launch_mb_type(ctx, &dj);
    if (rc) return rc;
    H264B_CUDA(ctx, cudaMemcpyAsync(j.mb_type, d_out, ns * (size_t)j.n_mb_max, cudaMemcpyDeviceToHost, st));
    H264B_CUDA(ctx, cudaMemcpyAsync(j.final, d_fin, ns * sizeof(h264b_mb_final), cudaMemcpyDeviceToHost, st));
    if (j.final_states) H264B_CUDA(ctx, cudaMemcpyAsync(j.final_states, d_fst, ns * j.n_ctx, cudaMemcpyDeviceToHost, st));
    H264B_CUDA(ctx, cudaStreamSynchronize(st));
    return H264B_OK;
}
