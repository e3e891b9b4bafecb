// cabac_engine.cu -- K3: the CABAC arithmetic-decoding engine, one slice per warp lane.
//
// Reference functions replaced (h264/cabac.go): initDecodingEngine :439-446, BinaryDecision core :525-536 +
// StateTransitionProcess :544-553 + RenormD :503-511 composed as DecodeDecision, DecodeBypass :468-481,
// DecodeTerminate :486-499; tables h264/rangeTabLPS.go and h264/stateTransxTab.go (folded into one 128 x 64-bit
// table indexed by the context's state byte, see ctx_init.cu).
//
// Mapping.  CABAC is serial inside a slice and slices share nothing, so a lane owns a slice.  All slices follow one
// shared op schedule, hence every lane of a warp executes the same op kind on the same ctxIdx at the same time: no
// divergence on the op, and the context-state access  state[ctxIdx][lane]  is one conflict-free 32-byte row of
// shared memory.  Per-lane bit windows drift apart (bits per bin are data dependent); refills are warp-synchronised:
// when any lane runs low, every lane that has room takes 32 more bits, so the refill code runs once per ~15-30 bins
// instead of (divergently) on almost every bin.  The next 32-bit word of each lane's bitstream is prefetched one
// refill ahead, which hides the (uncoalesced, L2-resident) load completely; bitstream traffic is ~0.11 B per bin.
#include "cabac_lane.cuh"
#include "common.cuh"
#include <stdlib.h>

namespace h264b {

#ifndef H264B_CABAC_PIPELINE
#define H264B_CABAC_PIPELINE 1  // fast loop: context state / table entry of a decision requested ahead of time
#endif
constexpr int kMaxWarpsPerCta = 20;            // one CTA per SM, five warps per scheduler
constexpr int kTab16Rows = 130;                // 128 states + the bypass pseudo state 128 (+ 1 spare)
constexpr int kTab16Bytes = kTab16Rows * 8 * 16; // the fast loops' table: rows x 8 copies x 16 bytes
constexpr size_t kMaxSmemPerCta = 227 * 1024;  // opt-in maximum of dynamic shared memory per CTA on sm_100

// explicit shared-memory accesses (32-bit shared addresses: no generic-address arithmetic in the inner loop)
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint2 lds_u32x2(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ uint4 lds_u32x4(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {  // (no masking of the selector)
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
__device__ __forceinline__ uint32_t opaque(uint32_t x) {  // keeps an address in a register instead of re-deriving it
    asm volatile("mov.b32 %0, %0;" : "+r"(x));
    return x;
}
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

struct CabacArgs {
    h264b_cabac_job j;
    const uint64_t *tab;   // 128-entry engine table
    const uint8_t *lut;    // K4 state LUT [5][52][1024]
    const uint32_t *order; // slices sorted by length (longest first) or NULL: lane -> slice = order[index]
    const uint32_t *d_n;   // actual slice count on the device (<= j.n_slices, which then is only the bound) or NULL
    uint32_t lanes_per_warp;
    uint32_t n_warps;      // bundles of lanes_per_warp slices
    uint32_t map_mode;     // 0: warp 0 of a CTA takes its longest bundle, 1: the highest warp does, 2: launch order,
                           // 3: slot_bundle[CTA x warps + warp] (bundle_assign_kernel), -1: none
    const int32_t *slot_bundle;
    uint32_t rounds;       // map_mode 3: bundles per slot (else 1)
    uint32_t n_rows;       // context rows per slice in shared memory
};

// ---------------------------------------------------------------------------------------------- length bundles
// A warp runs until its longest slice is done, so the 32 slices of a warp should be equally long: slices are bucketed
// by op count (2048 linear buckets between the shortest and the longest: a counting sort, order inside a bucket is
// arbitrary) and dealt to warps longest first, which also puts the long warps at the front of the launch.
constexpr int kLenBuckets = 2048;

struct SortScratch {
    uint32_t min_ops, max_ops;
    uint32_t hist[kLenBuckets];    // then: running cursor of each bucket
};

__device__ __forceinline__ uint32_t len_bucket(uint32_t ops, uint32_t lo, uint32_t hi) {  // longest -> bucket 0
    const uint64_t span = (uint64_t)(hi - lo) + 1;
    return (uint32_t)(((uint64_t)(hi - ops) * kLenBuckets) / span);
}

__global__ void __launch_bounds__(256) sort_minmax_kernel(const uint32_t *n_ops, uint32_t n, const uint32_t *d_n, uint32_t cap,
                                                          SortScratch *s) {
    if (d_n && *d_n < n) n = *d_n;
    uint32_t lo = 0xFFFFFFFFu, hi = 0;
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const uint32_t v = n_ops[i] < cap ? n_ops[i] : cap;
        lo = min(lo, v);
        hi = max(hi, v);
    }
    lo = __reduce_min_sync(0xFFFFFFFFu, lo);
    hi = __reduce_max_sync(0xFFFFFFFFu, hi);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&s->min_ops, lo);
        atomicMax(&s->max_ops, hi);
    }
}

__global__ void __launch_bounds__(256) sort_hist_kernel(const uint32_t *n_ops, uint32_t n, const uint32_t *d_n, uint32_t cap,
                                                          SortScratch *s) {
    if (d_n && *d_n < n) n = *d_n;
    const uint32_t lo = s->min_ops, hi = s->max_ops;
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256)
        atomicAdd(&s->hist[len_bucket(n_ops[i] < cap ? n_ops[i] : cap, lo, hi)], 1u);
}

__global__ void __launch_bounds__(1024) sort_scan_kernel(SortScratch *s) {  // exclusive scan of the 2048 counts
    __shared__ uint32_t warp_sum[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t a = s->hist[2 * tid], b = s->hist[2 * tid + 1];
    uint32_t x = a + b;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) warp_sum[warp] = x;
    __syncthreads();
    uint32_t base = 0;
    for (int w = 0; w < warp; w++) base += warp_sum[w];
    const uint32_t excl = base + x - (a + b);
    s->hist[2 * tid] = excl;
    s->hist[2 * tid + 1] = excl + a;
}

__global__ void __launch_bounds__(256) sort_scatter_kernel(const uint32_t *n_ops, uint32_t n, const uint32_t *d_n, uint32_t cap,
                                                           SortScratch *s, uint32_t *order) {
    if (d_n && *d_n < n) n = *d_n;
    const uint32_t lo = s->min_ops, hi = s->max_ops;
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256)
        order[atomicAdd(&s->hist[len_bucket(n_ops[i] < cap ? n_ops[i] : cap, lo, hi)], 1u)] = i;
}

// ---------------------------------------------------------------------------------------------- bundles -> schedulers
// One wave of one CTA per SM: warp w of CTA b runs on scheduler (b, w & 3) from start to end, so which bundles share a
// scheduler decides when the launch ends.  A scheduler is saturated by ~2.2 warps of the fast loop (a lone warp needs
// ~115 cycles per op, the issue port ~52), so its time is  max(longest bundle x 115, all its ops x 52)  and a few bundles
// -- the 32 longest slices of 80 000 are 1.9 x the mean -- want a scheduler (almost) to themselves.  Greedy "longest
// first onto the least loaded scheduler" with such bundles weighted up does that (tools/cabac_assign_model.py has the
// model next to the measurements: 80 ms -> 62 ms for BASELINE configs[3]).  One warp, ~0.2 ms for 2 500 bundles.
constexpr int kSlotsMax = 5;  // warps per scheduler = kMaxWarpsPerCta / 4

__global__ void __launch_bounds__(256) bundle_assign_kernel(const uint32_t *n_ops, const uint32_t *order, uint32_t n_bound,
                                                             const uint32_t *d_n, uint32_t cap_ops, uint32_t lpw,
                                                             uint32_t n_sched, uint32_t slots, uint32_t rounds, uint32_t W,
                                                             int32_t *slot_bundle) {
    extern __shared__ uint32_t sm_u32[];
    uint32_t *len = sm_u32;                       // [n_bundles] ops of each bundle (its longest slice)
    uint32_t *sum = len + n_sched * slots * rounds;  // [n_sched] weighted ops placed so far
    uint32_t *cnt = sum + n_sched;                // [n_sched] bundles placed so far
    __shared__ unsigned long long total_s;
    const uint32_t n_slices = d_n && *d_n < n_bound ? *d_n : n_bound;
    const uint32_t n_bundles = (n_slices + lpw - 1) / lpw;
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) total_s = 0;
    for (uint32_t k = tid; k < n_sched * slots * rounds; k += blockDim.x) slot_bundle[k] = -1;
    for (uint32_t k = tid; k < n_sched; k += blockDim.x) sum[k] = 0, cnt[k] = 0;
    __syncthreads();
    unsigned long long part = 0;
    for (uint32_t g = tid; g < n_bundles; g += blockDim.x) {
        uint32_t m = 0;
        for (uint32_t l = 0; l < lpw && g * lpw + l < n_slices; l++) {
            const uint32_t v = n_ops[order[g * lpw + l]];
            m = max(m, v < cap_ops ? v : cap_ops);
        }
        len[g] = m;
        part += m;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, d);
    if (lane == 0 && part) atomicAdd(&total_s, part);
    __syncthreads();
    if (tid >= 32) return;
    // a bundle whose own chain (x 2.2) comes near the balanced load of a scheduler counts 1.5 times
    const unsigned long long critical = total_s * 4 / (5ull * n_sched);  // 0.8 x total / schedulers
    // key of a scheduler: weighted ops / 8 (22 bits: 2^25 ops; more saturate) << 10 | its index; the least loaded one
    // (ties: the lowest index) is the minimum key -- one warp reduction per bundle.  Lane l keeps the keys of schedulers
    // l, l + 32, ... in registers (kKeys of them: 640 schedulers = 160 SMs).
    constexpr int kKeys = 20;
    uint32_t key[kKeys], load[kKeys], filled[kKeys];
#pragma unroll
    for (int q = 0; q < kKeys; q++) {
        const uint32_t k = (uint32_t)lane + 32u * q;
        load[q] = 0, filled[q] = 0;
        key[q] = k < n_sched ? k : 0xFFFFFFFFu;
    }
    for (uint32_t g = 0; g < n_bundles; g++) {  // bundles come longest first
        uint32_t my_key = key[0];
#pragma unroll
        for (int q = 1; q < kKeys; q++) my_key = min(my_key, key[q]);
        const uint32_t sidx = __reduce_min_sync(0xFFFFFFFFu, my_key) & 1023u;
        if ((sidx & 31u) == (uint32_t)lane) {  // its owner places the bundle
            const uint32_t L = len[g];
            const uint32_t wgt = ((unsigned long long)L * 11 / 5 > critical) ? L + L / 2 : L;
            const uint32_t mine = sidx >> 5;
#pragma unroll
            for (int q = 0; q < kKeys; q++) {
                if ((uint32_t)q == mine) {
                    // the scheduler's bundles fill its warps round by round: the longest ones run side by side first
                    const uint32_t kpos = filled[q], round = kpos / slots, slot = kpos % slots;
                    slot_bundle[(round * (n_sched >> 2) + (sidx >> 2)) * W + slot * 4 + (sidx & 3u)] = (int32_t)g;
                    filled[q] = kpos + 1;
                    load[q] += wgt;
                    key[q] = kpos + 1 >= slots * rounds ? 0xFFFFFFFFu
                                                        : ((load[q] >= (1u << 25) ? 0x3FFFFEu : load[q] >> 3) << 10) | sidx;
                }
            }
        }
    }
}

__device__ __forceinline__ int idc_class_dev(int idc) { return (idc >= -1 && idc <= 2) ? idc + 1 : 4; }
__device__ __forceinline__ int clip3_dev(int x, int y, int z) { return z < x ? x : (z > y ? y : z); }

// 16-byte form of an engine-table entry for the pipelined fast loop (kLoop 1):
//   x  rangeTabLPS[state][0..3]                      (as in the 8-byte entry)
//   y  renormalisation shift after an LPS, per q:    clz(rangeLPS) - 23  (RenormD's step count for codIRange = rangeLPS)
//   z  next state after an MPS | bin << 15 | next state after an LPS << 16 | bin << 31   (the r1 "fast" form)
// Kept 8 times in shared memory, copy c in the 16-byte column c of every 128-byte row: lane l reads copy l & 7, so the
// 8 lanes of a quarter warp (one LDS.128 wavefront) always hit 8 different bank groups -- no conflicts for any states.
__device__ __forceinline__ uint4 entry16(uint64_t e) {
    uint4 r;
    r.x = (uint32_t)e;
    uint32_t y = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const uint32_t lps = (r.x >> (8 * q)) & 0xFFu;
        y |= ((lps ? (uint32_t)__clz((int)lps) - 23u : 9u) & 0xFFu) << (8 * q);
    }
    r.y = y;
    const uint64_t f = (e & 0x00FF00FFFFFFFFFFull) | ((e & 0x0100010000000000ull) << 7);
    r.z = (uint32_t)(f >> 32);
    r.w = 0u;
    return r;
}

// kRounds: a warp may run several bundles in turn (a launch whose context rows leave room for few warps).  The usual
// launch has one bundle per warp and takes the instantiation without that loop: under it the compiler keeps fewer
// warp-uniform values on the uniform datapath, which a warp on its own pays with a third more cycles per bin in the
// older loops (configs[1]: 193 against 138 cycles per bin with kLoop 0, 244 against 177 on the literal engine).
template <int kLoop, bool kRounds>
__global__ void __launch_bounds__(kMaxWarpsPerCta * 32, 1) cabac_decode_kernel(CabacArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint64_t *s_tab = reinterpret_cast<uint64_t *>(smem);                  // 128 x 8 B (generic loop)
    uint64_t *s_tab_fast = reinterpret_cast<uint64_t *>(smem + 1024);      // kLoop 0: 256 x 8 B, the r1 fast loop's form
    uint4 *s_tab16 = reinterpret_cast<uint4 *>(smem + 1024);               // kLoop 1: 128 x 8 x 16 B
    uint8_t *s_state_all = smem + 1024 + (kLoop ? kTab16Bytes : 2048);     // [warp][n_ctx][32]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t W = blockDim.x >> 5;
    for (int i = tid; i < 128; i += blockDim.x) s_tab[i] = a.tab[i];
    if (kLoop == 0) {
        for (int i = tid; i < 256; i += blockDim.x) {
            const uint64_t e = a.tab[i & 127];
            // bins from bit 0 to bit 7 of their bytes (bits 40 -> 47, 56 -> 63); entries 128..255 repeat 0..127
            s_tab_fast[i] = (e & 0x00FF00FFFFFFFFFFull) | ((e & 0x0100010000000000ull) << 7);
        }
    } else {
        for (int i = tid; i < kTab16Rows * 8; i += blockDim.x) {
            // rows 128..: the bypass pseudo state (kLoop 2): rangeLPS 0, shift 0, next state 128 either way, bins (0, 1)
            s_tab16[i] = (i >> 3) < 128 ? entry16(a.tab[i >> 3]) : make_uint4(0u, 0u, 0x80800080u, 0xFFFFFFFFu);
        }
    }
    __syncthreads();
    const h264b_cabac_job &j = a.j;
    const uint32_t n_ctx = a.n_rows;  // context rows kept in shared memory (j.n_ctx, or the caller's n_ctx_used)
    uint8_t *s_state = s_state_all + (size_t)warp * (n_ctx + 2) * 32;  // (+ 2: kLoop 2's bypass pseudo contexts)
    const uint32_t n_slices = a.d_n && *a.d_n < j.n_slices ? *a.d_n : j.n_slices;
    const uint32_t n_bundles = (n_slices + a.lanes_per_warp - 1) / a.lanes_per_warp;

    // Bundle (lanes_per_warp slices of similar length, longest bundles first) -> warp.  One wave: the CTAs are spread over
    // the SMs and the position of a warp is  rank of the warp inside its CTA x CTAs + CTA,  so the longest bundles are
    // spread over all SMs and sit on the warps the issue arbiter prefers (map_mode 1: the highest warp of a CTA takes
    // its longest bundle).  More bundles than one wave holds: small CTAs in launch order.
    const uint32_t rank = a.map_mode == 1 ? (W - 1u - (uint32_t)warp) : (uint32_t)warp;
    uint32_t gw = a.map_mode == 2 ? blockIdx.x * W + (uint32_t)warp : rank * gridDim.x + blockIdx.x;
    // map_mode 3: the warp runs the bundles of its slot one after the other (rounds: as many as the launch needs for all
    // bundles to have a slot: one, unless the context rows leave room for few warps); -1: no bundle
    const uint32_t n_rounds = kRounds ? a.rounds : 1u;
    for (uint32_t round = 0; round < n_rounds; round++) {
    if (a.map_mode == 3) gw = (uint32_t)a.slot_bundle[(round * gridDim.x + blockIdx.x) * W + (uint32_t)warp];
    if (kRounds) {
        if (__all_sync(0xFFFFFFFFu, gw >= n_bundles)) continue;  // (a vote: gw is the same in every lane)
    } else {
        if (gw >= n_bundles) return;
    }
    const uint32_t index = gw * a.lanes_per_warp + lane;
    // Lanes without a slice of their own (a partly filled warp) shadow the warp's first lane: they decode the same
    // slice and store nothing, so the loops below never have to predicate on "is there a slice in this lane".
    const bool own = lane < (int)a.lanes_per_warp && index < n_slices;
    const uint32_t src_index = own ? index : gw * a.lanes_per_warp;
    const uint32_t slice = a.order ? a.order[src_index] : src_index;

    // ---- per-lane setup
    uint32_t my_ops = j.n_ops ? j.n_ops[slice] : j.n_ops_max;
    if (my_ops > j.n_ops_max) my_ops = j.n_ops_max;
    uint64_t off = j.off[slice];
    uint32_t len = j.len[slice];
    // a slice that does not lie inside the buffer (a caller's bad offset) decodes the buffer's last bytes and is flagged
    bool bad_range = false;
    if (off > j.total_bytes) off = j.total_bytes, bad_range = true;
    if ((uint64_t)len > j.total_bytes - off) len = (uint32_t)(j.total_bytes - off), bad_range = true;
    // initial context states: given, or the K4 rule (state LUT row of this slice's (idc class, clipped qp))
    uint32_t st_or = 0;
    bool st_63 = false;  // a caller-given pStateIdx 63 (rangeLPS 2: seven renormalisation shifts; unreachable otherwise)
    {
        const uint8_t *src;
        if (j.init_states) {
            src = j.init_states + (size_t)slice * j.n_ctx;
        } else {
            const h264b_slice_qp p = j.qp[slice];
            src = a.lut + ((size_t)idc_class_dev(p.cabac_init_idc) * 52 + clip3_dev(0, 51, p.slice_qp_y)) * 1024;
        }
        for (uint32_t c = 0; c < n_ctx; c++) {
            const uint8_t v = src[c];
            st_or |= v;
            st_63 |= (v & 63u) == 63u;
            s_state[c * 32 + lane] = v;
        }
    }
    __syncwarp();

    LaneDecoder eng = {};
    eng.init(j.bytes, j.total_bytes, off, (j.flags & H264B_BYPASS_SPEC_OR) != 0);

    uint32_t warp_ops = my_ops;
#pragma unroll
    for (int d = 16; d; d >>= 1) warp_ops = max(warp_ops, __shfl_xor_sync(0xFFFFFFFFu, warp_ops, d));

    uint32_t *bins = j.bins + (j.bins_off ? (size_t)j.bins_off[slice] : (size_t)slice * j.bins_stride_words);
    uint32_t word = 0;
    uint32_t i = 0;
    bool live = true;  // this lane's slice is not finished (kLoop 2 finishes lanes one by one, see below)
    // one op of the generic form (any engine, any lane state) for the lanes that ask for it
    const auto generic_op = [&](uint32_t at, uint32_t op, bool active) {
        if (__any_sync(0xFFFFFFFFu, active && eng.must_refill())) {
            if (active) eng.refill_if_room();
        }
        const uint32_t kind = op >> 14;
        uint32_t bin = 0;
        if (kind == H264B_OP_DECISION) {
            uint32_t c = op & 0x3FFu;
            if (c >= n_ctx) c = 0;
            uint8_t *sp = s_state + c * 32 + lane;
            if (active) {
                const uint64_t e = s_tab[*sp & 127u];
                uint8_t ns;
                bin = eng.decision(e, &ns);
                *sp = ns;
            }
        } else if (kind == H264B_OP_BYPASS) {
            if (active) bin = eng.bypass();
        } else {
            if (active) bin = eng.terminate();
        }
        if (active) {  // a lane that has finished keeps its last partial word for finish_lane()
            word |= bin << (at & 31u);
            if ((at & 31u) == 31u) {
                if (own) bins[at >> 5] = word;
                word = 0;
            }
        }
    };
    // the end of a lane's slice: optional final DecodeTerminate, flush, final record
    const auto finish_lane = [&]() {
        if (!own) return;
        uint32_t n_bins = my_ops;
        if (j.flags & H264B_CABAC_FINAL_TERMINATE) {
            if (eng.must_refill()) eng.refill_if_room();
            const uint32_t bin = eng.terminate();
            word |= bin << (my_ops & 31u);  // `word` holds bins (my_ops & ~31) .. my_ops-1 (empty after a flush)
            n_bins++;
        }
        if (n_bins & 31u) bins[n_bins >> 5] = word;
        else if ((j.flags & H264B_CABAC_FINAL_TERMINATE) && (n_bins & 31u) == 0) bins[(n_bins - 1) >> 5] = word;
        h264b_cabac_final f;
        const uint64_t bits_read = eng.bits_read();
        f.cod_i_range = eng.cod_i_range();
        f.cod_i_offset = eng.cod_i_offset();
        f.bits_read = bits_read;
        f.flags = (bits_read > 8ull * len || bad_range) ? H264B_F_OVERRUN : 0u;
        f.n_bins = n_bins;
        j.final[slice] = f;
        if (j.final_states) {
            uint8_t *dst = j.final_states + (size_t)slice * j.n_ctx;
            for (uint32_t c = 0; c < n_ctx; c++) dst[c] = s_state[c * 32 + lane];
            if (n_ctx < j.n_ctx) {  // contexts the schedule never touches: as they were initialised
                const uint8_t *src;
                if (j.init_states) {
                    src = j.init_states + (size_t)slice * j.n_ctx;
                } else {
                    const h264b_slice_qp p = j.qp[slice];
                    src = a.lut + ((size_t)idc_class_dev(p.cabac_init_idc) * 52 + clip3_dev(0, 51, p.slice_qp_y)) * 1024;
                }
                for (uint32_t c = n_ctx; c < j.n_ctx; c++) dst[c] = src[c];
            }
        }
    };
    // ---- fast loop: blocks of 32 ops (one word of bins) while every lane of the warp is active and on the window
    // engine -- the usual case for all but the tail of a length bundle.  Same arithmetic as CabacLane::decision /
    // bypass / terminate.  A warp's time is its serial chain codIRange / codIOffset -> next bin plus every taken branch
    // (tens of cycles for a warp that has its scheduler to itself, as the long slices at the end of a launch do), so:
    //   * the op kinds of a block are two ballots, so every branch on them is warp-uniform (no convergence barriers);
    //   * the context state and the table entry of a decision are requested ahead of the arithmetic;
    //   * shared memory is addressed with explicit 32-bit shared addresses (ld/st.shared), not generic pointers;
    //   * codIRange is kept as R << 22, aligned with codIOffset in the window: (R22 >> 28) is 4 + qCodIRangeIdx, which as
    //     a byte-permute selector picks rangeTabLPS[state][q] out of the table word directly;
    //   * the table carries the bin in bit 7 of its byte, so one byte permute yields (next state -> byte 0, bin ->
    //     bit 31) and one funnel shift appends the bin; the word is bit-reversed once per 32 bins;
    //   * refills are checked once per two ops (16 bits cover two decisions).
    // (every condition that steers the loop is a vote result: the compiler then knows the warp stays converged and
    //  emits the shuffles and votes inside without divergence checks)
    if (kLoop == 2) {
      // kLoop 2: no branch inside a chunk of 8 ops, and as few instructions as the arithmetic allows.  A lone warp -- the
      // long slices at the end of a launch, or a launch of a few slices -- pays ~28 cycles for every predicate -> taken
      // branch and the static waits of predicated-off code all the same (ncu: profiles/r2_lone_*_sass.txt), and the
      // op-kind dispatch plus the refill vote of the branchy loops take 2 - 3 branches per op.  Here
      //  * a bypass is a decision on a pseudo context: its row holds the state byte 128 for ever, table entry 128 has
      //    rangeLPS 0, shift 0, bins (0, 1) and an all-ones word w that keeps codIRange on the "LPS" side; the window's
      //    bypass shift rides on the renormalisation shift of the op BEFORE it (bit 7 of the next op's state byte);
      //    nothing in the straight line asks for the op kind;
      //  * the window is 96 bits (codIOffset + up to 86 stream bits), so one warp vote per chunk covers its refills:
      //    8 ops take at most 8 x 6 + 1 bits (pStateIdx 63, rangeLPS 2, is unreachable: such caller-given states stay
      //    with the other loops), a lane holding fewer than 49 refills, and with it every lane that has room;
      //  * row addresses are broadcast three ops ahead (for the state byte fetched two ops ahead) and reused for the
      //    store; a chunk in which a context comes back within two ops (known from the schedule) runs the variant that
      //    re-requests what the store has overtaken;
      //  * the selects on the decision's outcome are written as selp (a predicate used as a guard costs ~10 cycles more
      //    than one used as data).
      // Chunks with a terminate op (1 in 48) run op by op.
      if (__all_sync(0xFFFFFFFFu, !eng.lit && !(st_or & 0x80u) && !st_63) && j.total_bytes < (1ull << 33)) {
        CabacLane &w = eng.w;
        const uint32_t st_lane = opaque(smem_addr(s_state) + (uint32_t)lane);  // &state[0][lane]
        const uint32_t tabl = opaque(smem_addr(s_tab16) + (uint32_t)(lane & 7) * 16u);
        const uint32_t sel_mps = opaque(0x1440u), sel_lps = opaque(0x3442u);  // byte-permute selectors, kept in registers
        // two bypass rows, for ops at even / odd positions: neighbouring bypass ops never look like one context
        sts_u8(st_lane + n_ctx * 32u, 128u);
        sts_u8(st_lane + n_ctx * 32u + 32u, 128u);
        // window: hi = codIOffset (10 bits) | 22 stream bits, mid, lo = the next 64; fbits of them are valid
        uint32_t R22, hi, mid, lo;
        int32_t fbits;
        // the bit feed as a word index, two words ahead: pf0 = words[widx], pf1 = words[widx + 1] (indices clamped)
        const uint32_t *words = reinterpret_cast<const uint32_t *>(j.bytes);
        const uint32_t last_idx = (uint32_t)(w.feed.last_word - words);
        const uint32_t mis8 = w.feed.mis8;
        uint32_t widx, widx0, refills0, cur, pf0, pf1;
        const auto enter = [&]() {  // CabacLane -> registers of this loop
            R22 = w.R << 22, hi = w.hi, mid = w.lo, lo = 0u, fbits = w.fbits;
            widx = (uint32_t)(w.feed.next - words);
            widx0 = widx, refills0 = w.refills;
            cur = w.feed.cur, pf0 = w.feed.pf, pf1 = words[widx + 1u < last_idx ? widx + 1u : last_idx];
        };
        const auto leave = [&]() {  // ... and back: a lane holding more than 54 stream bits gives the last 32 back
            if (fbits > 54) {
                const uint32_t keep = (uint32_t)(fbits - 32 - 22);  // valid bits that stay in mid (1 .. 32)
                mid = keep >= 32u ? mid : (mid & ~(0xFFFFFFFFu >> keep));
                fbits -= 32;
                widx--;
            }
            w.R = R22 >> 22, w.hi = hi, w.lo = mid, w.fbits = fbits;
            w.refills = refills0 + (widx - widx0);
            w.feed.next = words + widx;
            const uint32_t prev = widx ? widx - 1u : 0u;
            w.feed.cur = bswap32(words[prev < last_idx ? prev : last_idx]);
            w.feed.pf = words[widx < last_idx ? widx : last_idx];
        };
        const auto next32 = [&]() -> uint32_t {  // the next 32 stream bits; the feed moves on one word
            const uint32_t nxt = bswap32(pf0);
            const uint32_t v = funnel_l(nxt, cur, mis8);
            cur = nxt;
            pf0 = pf1;
            widx++;
            pf1 = words[widx + 1u < last_idx ? widx + 1u : last_idx];
            return v;
        };
        const auto refill_chunk = [&]() {  // every lane that has room; a lane below 23 twice (rare)
            if (fbits <= 22) {
                const uint32_t v = next32();
                const uint32_t s = (uint32_t)(22 - fbits);
                hi |= funnel_l(v, 0u, s);  // v >> (32 - s), 0 for s == 0
                mid |= v << s;
                fbits += 32;
            }
            if (fbits <= 54) {  // (> 22 here)
                const uint32_t v = next32();
                const uint32_t t = (uint32_t)(fbits - 22);  // 1 .. 32
                mid |= __funnelshift_rc(v, 0u, t);          // v >> t, 0 for t == 32
                lo |= __funnelshift_rc(0u, v, t);           // v << (32 - t)
                fbits += 32;
            }
        };
        const auto shift3 = [&](uint32_t k) {
            hi = __funnelshift_l(mid, hi, k);
            mid = __funnelshift_l(lo, mid, k);
            lo <<= k;
            fbits -= (int32_t)k;
        };
        bool left = false;
        uint4 e_cur = make_uint4(0u, 0u, 0u, 0u);  // table entry of the op at hand
        uint32_t s1 = 0;                           // state byte of the op after it
        uint32_t a0 = 0, a1 = 0, a2 = 0;           // state addresses of the op at hand and of the two after it
        // A lane whose slice ends before the warp's longest one finishes on its own (the < 32 ops that do not fill a block
        // op by op, then its final record) and rides along from there, decoding on without storing anything: the block
        // loop goes on for the other lanes.  (The 32 longest slices of a launch differ by tens of per cent.)
        for (;;) {
        enter();
        uint32_t next_op = i + (uint32_t)lane < j.n_ops_max ? (uint32_t)j.ops[i + (uint32_t)lane] : 0u;
        while (!left && __all_sync(0xFFFFFFFFu, !live || i + 32u <= my_ops) && __any_sync(0xFFFFFFFFu, live)) {
            const uint32_t my_op = next_op;
            next_op = i + 32u + (uint32_t)lane < j.n_ops_max ? (uint32_t)j.ops[i + 32u + (uint32_t)lane] : 0u;
            const uint32_t my_kind = my_op >> 14;
            const uint32_t dec_mask = __ballot_sync(0xFFFFFFFFu, my_kind == H264B_OP_DECISION);
            const uint32_t byp_mask = __ballot_sync(0xFFFFFFFFu, my_kind == H264B_OP_BYPASS);
            const uint32_t trm_mask = ~(dec_mask | byp_mask);
            uint32_t my_row = (my_kind == H264B_OP_DECISION ? ((my_op & 0x3FFu) < n_ctx ? (my_op & 0x3FFu) : 0u)
                                                            : n_ctx + ((uint32_t)lane & 1u)) * 32u;
            // op p stores the state that the requests already made for op p + 1 / p + 2 have read
            const uint32_t r1 = __shfl_down_sync(0xFFFFFFFFu, my_row, 1), r2 = __shfl_down_sync(0xFFFFFFFFu, my_row, 2);
            const uint32_t haz_mask = __ballot_sync(0xFFFFFFFFu, my_kind == H264B_OP_DECISION &&
                                                                     ((lane < 31 && r1 == my_row) || (lane < 30 && r2 == my_row)));
            // per chunk c: bit c = holds a terminate op (runs op by op), bit 4 + c = a context comes back within two ops
            uint32_t chunk_flags = 0;
#pragma unroll
            for (int c = 0; c < 4; c++)
                chunk_flags |= (((trm_mask >> (8 * c)) & 0xFFu) ? 1u << c : 0u) | (((haz_mask >> (8 * c)) & 0xFFu) ? 16u << c : 0u);
            bool primed = false;  // e_cur / s1 / a0 / a1 / a2 are those of the op at hand (straight-line chunks hand them on)
            uint32_t k = 0;
#pragma unroll 1
            for (uint32_t k8 = 0; k8 < 32u && !left; k8 += 8u) {
                const uint32_t row8 = my_row;  // lanes 0..7: this chunk's ops, lanes 8..10: the next chunk's first three
                my_row = __shfl_sync(0xFFFFFFFFu, my_row, (lane + 8) & 31);
                const uint32_t cf = chunk_flags >> (k8 >> 3);
                if (__any_sync(0xFFFFFFFFu, fbits < 49)) {
                    __syncwarp();  // (keeps this block a branch: predicated off it would still pay its static waits)
                    refill_chunk();
                }
                if (cf & 1u) {  // ---- a chunk with a terminate op: op by op, nothing in flight
                    const uint32_t dm = dec_mask >> k8, bm = byp_mask >> k8;
#pragma unroll 1
                    for (uint32_t u = 0; u < 8u; u++) {
                        const uint32_t row = __shfl_sync(0xFFFFFFFFu, row8, (int)u);
                        if ((dm >> u) & 1u) {
                            const uint32_t addr = row + st_lane;
                            const uint4 e = lds_u32x4(tabl + lds_u8(addr) * 128u);
                            const uint32_t q4 = R22 >> 28;
                            const uint32_t lps22 = prmt(0u, e.x, q4) << 22;
                            const uint32_t sh_lps = prmt(0u, e.y, q4);
                            const uint32_t rm22 = R22 - lps22;
                            const bool is_lps = hi >= rm22;
                            const uint32_t sh = is_lps ? sh_lps : ((rm22 >> 30) ^ 1u);
                            const uint32_t r22 = is_lps ? lps22 : rm22;
                            hi = is_lps ? hi - rm22 : hi;
                            const uint32_t sel = prmt(e.z, 0u, is_lps ? sel_lps : sel_mps);
                            sts_u8(addr, sel);
                            R22 = r22 << sh;
                            shift3(sh);
                            word = __funnelshift_l(sel, word, 1);
                        } else if ((bm >> u) & 1u) {
                            shift3(1u);
                            const bool one = hi >= R22;
                            if (one) hi -= R22;
                            word = (word << 1) | (one ? 1u : 0u);
                        } else {  // DecodeTerminate (CabacLane::terminate on this window)
                            R22 -= 2u << 22;
                            const uint32_t bin = (hi >= R22 && live) ? 1u : 0u;  // (a lane that rides along: whatever)
                            if (!bin) {
                                const uint32_t sh = (uint32_t)__clz((int)R22) - 1u;
                                R22 <<= sh;
                                shift3(sh);
                            }
                            word = (word << 1) | bin;
                            if (__any_sync(0xFFFFFFFFu, bin)) {  // a slice that goes on after its end: the generic loop takes over
                                left = true;
                                k = k8 + u + 1u;
                                break;
                            }
                        }
                    }
                    primed = false;
                    continue;
                }
                // ---- the straight line
                const bool hand_on = k8 < 24u && !((cf >> 1) & 1u);  // the next chunk of this block runs here too
                const uint32_t pre_mask = hand_on ? 1u : 0u;
                if (!primed) {
                    // the first op's own bypass shift (the op before it ran elsewhere), its entry, and the state byte of
                    // the second op
                    shift3((byp_mask >> k8) & 1u);
                    a0 = __shfl_sync(0xFFFFFFFFu, row8, 0) + st_lane;
                    a1 = __shfl_sync(0xFFFFFFFFu, row8, 1) + st_lane;
                    a2 = __shfl_sync(0xFFFFFFFFu, row8, 2) + st_lane;
                    e_cur = lds_u32x4(tabl + lds_u8(a0) * 128u);
                    s1 = lds_u8(a1);
                    primed = true;
                }
#define H264B_UNIFIED_OP(u, kHaz)                                                                                        \
    {                                                                                                                    \
        const uint32_t a3 = __shfl_sync(0xFFFFFFFFu, row8, (u) + 3) + st_lane; /* address of op u + 3's state */         \
        uint32_t s2 = lds_u8(a2);                                              /* state byte of op u + 2 */              \
        const uint4 e = e_cur;                                                                                           \
        e_cur = lds_u32x4(tabl + s1 * 128u); /* entry of op u + 1, ahead of this op's arithmetic */                      \
        const uint32_t q4 = R22 >> 28;                                                                                   \
        const uint32_t lps22 = prmt(0u, e.x, q4) << 22; /* rangeTabLPS[state][q] << 22; bypass: 0 */                     \
        const uint32_t sh_lps = prmt(0u, e.y, q4);      /* bypass: 0 */                                                  \
        const uint32_t rm22 = R22 - lps22;                                                                               \
        const uint32_t r_lps = lps22 | (R22 & e.w);     /* a bypass keeps codIRange either way */                        \
        const uint32_t sh_mps = (rm22 >> 30) ^ 1u;      /* codIRange - rangeLPS >= 128: at most one doubling */           \
        const uint32_t hi_lps = hi - rm22;                                                                               \
        uint32_t r22, sh, selr;                                                                                          \
        asm("{\n\t.reg .pred p;\n\t"                                                                                     \
            "setp.ge.u32 p, %3, %4;\n\t"                                                                                 \
            "selp.u32 %0, %5, %4, p;\n\t"                                                                                \
            "selp.u32 %1, %6, %7, p;\n\t"                                                                                \
            "selp.u32 %2, %8, %9, p;\n\t"                                                                                \
            "selp.u32 %3, %10, %3, p;\n\t}"                                                                              \
            : "=r"(r22), "=r"(sh), "=r"(selr), "+r"(hi)                                                                  \
            : "r"(rm22), "r"(r_lps), "r"(sh_lps), "r"(sh_mps), "r"(sel_lps), "r"(sel_mps), "r"(hi_lps));                 \
        const uint32_t sel = prmt(e.z, 0u, selr); /* next state | bin << 31 */                                           \
        sts_u8(a0, sel);                                                                                                 \
        if (kHaz) { /* this context again within two ops: what the store has overtaken is asked for again */             \
            if (a1 == a0) e_cur = lds_u32x4(tabl + (sel & 0xFFu) * 128u);                                                \
            if (a2 == a0) s2 = sel & 0xFFu;                                                                              \
        }                                                                                                                \
        R22 = r22 << sh;                                                                                                 \
        /* + the bypass shift of the op that follows (state byte 128) */                                                 \
        const uint32_t sht = sh + ((u) == 7 ? ((s1 >> 7) & pre_mask) : (s1 >> 7));                                       \
        shift3(sht);                                                                                                     \
        word = __funnelshift_l(sel, word, 1);                                                                            \
        s1 = s2;                                                                                                         \
        a0 = a1;                                                                                                         \
        a1 = a2;                                                                                                         \
        a2 = a3;                                                                                                         \
    }
                if (!(cf & 16u)) {
                    H264B_UNIFIED_OP(0, false) H264B_UNIFIED_OP(1, false) H264B_UNIFIED_OP(2, false)
                    H264B_UNIFIED_OP(3, false) H264B_UNIFIED_OP(4, false) H264B_UNIFIED_OP(5, false)
                    H264B_UNIFIED_OP(6, false) H264B_UNIFIED_OP(7, false)
                } else {
                    H264B_UNIFIED_OP(0, true) H264B_UNIFIED_OP(1, true) H264B_UNIFIED_OP(2, true)
                    H264B_UNIFIED_OP(3, true) H264B_UNIFIED_OP(4, true) H264B_UNIFIED_OP(5, true)
                    H264B_UNIFIED_OP(6, true) H264B_UNIFIED_OP(7, true)
                }
#undef H264B_UNIFIED_OP
                primed = hand_on;
            }
            if (!left) {
                i += 32u;
                if (own && live) bins[(i >> 5) - 1u] = __brev(word);
                word = 0;
            } else {
                i += k;
                word = k < 32u ? __brev(word) >> (32u - k) : __brev(word);  // bins 0..k-1 in bits 0..k-1
                if (k == 32u) {
                    if (own && live) bins[(i >> 5) - 1u] = word;
                    word = 0;
                }
            }
        }
        leave();
        if (left) {
            if (live && hi >= R22) eng.to_literal();  // the terminate bin of 1: codIOffset >= codIRange from here on
            break;
        }
        // lanes whose slice ends inside the next block: their last ops one by one, then their final record
        const bool ending = live && my_ops < i + 32u;
        if (!__any_sync(0xFFFFFFFFu, ending)) break;  // (no lane is live any more)
        uint32_t end_at = ending ? my_ops : 0u;
#pragma unroll
        for (int d = 16; d; d >>= 1) end_at = max(end_at, __shfl_xor_sync(0xFFFFFFFFu, end_at, d));
        for (uint32_t at = i; at < end_at; at++) generic_op(at, j.ops[at], ending && at < my_ops);
        if (ending) {
            finish_lane();
            live = false;
            word = 0;
        }
        if (!__any_sync(0xFFFFFFFFu, live) || __any_sync(0xFFFFFFFFu, live && eng.lit)) break;
        }
      }
    } else
    if (kLoop == 1) {
      // kLoop 1.  The serial chain of a decision is  codIRange -> rangeLPS -> compare -> new codIRange: nothing else may
      // sit on it.  (1) The entry of the next decision is requested before this decision's arithmetic, from the state
      // byte fetched one decision earlier; only when the next decision (or the one after it) uses this decision's own
      // context -- known from the schedule: two ballots per block -- a rare, warp-uniform branch re-requests it from the
      // state just written.  (2) The renormalisation shift comes from the table (LPS: static per state and q) or from
      // bit 30 of the MPS range, not from a count-leading-zeros (XU pipe, ~25 cycles).  (3) The table is replicated so
      // that its reads are bank-conflict free for any combination of states (entry16()).
      if (__all_sync(0xFFFFFFFFu, my_ops >= 32u && !eng.lit && !(st_or & 0x80u))) {
        CabacLane &w = eng.w;
        const uint32_t st_lane = opaque(smem_addr(s_state) + (uint32_t)lane);  // &state[0][lane]
        const uint32_t tabl = opaque(smem_addr(s_tab16) + (uint32_t)(lane & 7) * 16u);
        const uint32_t sel_mps = opaque(0x1440u), sel_lps = opaque(0x3442u);  // byte-permute selectors, kept in registers
        uint32_t R22 = w.R << 22, hi = w.hi, lo = w.lo;
        int32_t fbits = w.fbits;
        bool left = false;
        uint4 e_cur = make_uint4(0u, 0u, 0u, 0u);  // table entry of the next decision
        uint32_t s1 = 0;                           // state byte of the decision after that
        uint32_t next_op = i + (uint32_t)lane < j.n_ops_max ? (uint32_t)j.ops[i + (uint32_t)lane] : 0u;
        while (!left && __all_sync(0xFFFFFFFFu, i + 32u <= my_ops)) {
            const uint32_t my_op = next_op;
            next_op = i + 32u + (uint32_t)lane < j.n_ops_max ? (uint32_t)j.ops[i + 32u + (uint32_t)lane] : 0u;
            const uint32_t my_kind = my_op >> 14;
            const uint32_t dec_mask = __ballot_sync(0xFFFFFFFFu, my_kind == H264B_OP_DECISION);
            const uint32_t byp_mask = __ballot_sync(0xFFFFFFFFu, my_kind == H264B_OP_BYPASS);
            uint32_t my_row = ((my_op & 0x3FFu) < n_ctx ? (my_op & 0x3FFu) : 0u) * 32u;  // as below: ctx 0
            const uint32_t above = dec_mask & ~((2u << lane) - 1u);
            const uint32_t above2 = above & (above - 1u);
            const uint32_t nrow1 = __shfl_sync(0xFFFFFFFFu, my_row, above ? __ffs((int)above) - 1 : lane);
            uint32_t my_nn = __shfl_sync(0xFFFFFFFFu, my_row, above2 ? __ffs((int)above2) - 1 : lane);
            const uint32_t fwd1_mask = __ballot_sync(0xFFFFFFFFu, above != 0u && nrow1 == my_row);
            const uint32_t fwd2_mask = __ballot_sync(0xFFFFFFFFu, above2 != 0u && my_nn == my_row);
            if (!above2) my_nn = 0u;  // (no second decision behind this op in the block: a harmless load of row 0)
            {   // the block's first two decisions: entry of the first, state of the second (once per block, in place)
                const uint32_t rest = dec_mask & (dec_mask - 1u);
                const uint32_t row_d0 = __shfl_sync(0xFFFFFFFFu, my_row, dec_mask ? __ffs((int)dec_mask) - 1 : 0);
                const uint32_t row_d1 = __shfl_sync(0xFFFFFFFFu, my_row, rest ? __ffs((int)rest) - 1 : 0);
                e_cur = lds_u32x4(tabl + lds_u8(row_d0 + st_lane) * 128u);
                s1 = lds_u8(row_d1 + st_lane);
            }
            uint32_t k = 0;
#pragma unroll 1
            for (uint32_t k8 = 0; k8 < 32u && !left; k8 += 8u) {
                // lanes 0..7 hold the rows of this chunk's ops (the rows rotate by 8 lanes per chunk), so the shuffles
                // below have constant source lanes
                const uint32_t dm = dec_mask >> k8, bm = byp_mask >> k8;
                const uint32_t row8 = my_row;
                my_row = __shfl_sync(0xFFFFFFFFu, my_row, (lane + 8) & 31);
                const uint32_t f1m = fwd1_mask >> k8, f2m = fwd2_mask >> k8, nn8 = my_nn;
                const uint32_t hzm = f1m | f2m;
                my_nn = __shfl_sync(0xFFFFFFFFu, my_nn, (lane + 8) & 31);
#pragma unroll
                for (uint32_t u = 0; u < 8u; u++) {
                    if ((u & 1u) == 0u) {
                        if (__builtin_expect(__any_sync(0xFFFFFFFFu, fbits < 16), 0)) {
                            __syncwarp();  // (also keeps this rare block a branch instead of 20 predicated instructions)
                            if (fbits <= 22) {
                                w.hi = hi, w.lo = lo, w.fbits = fbits;
                                w.refill();
                                hi = w.hi, lo = w.lo, fbits = w.fbits;
                            }
                        }
                    }
                    if (dm & (1u << u)) {
                        const uint32_t s2 = lds_u8(__shfl_sync(0xFFFFFFFFu, nn8, (int)u) + st_lane);  // state of the one after next
                        const uint32_t addr = __shfl_sync(0xFFFFFFFFu, row8, (int)u) + st_lane;
                        const uint4 e = e_cur;
                        e_cur = lds_u32x4(tabl + s1 * 128u);  // the next decision's entry, ahead of this one's arithmetic
                        const uint32_t q4 = R22 >> 28;
                        const uint32_t lps22 = prmt(0u, e.x, q4) << 22;  // rangeTabLPS[state][q] << 22
                        const uint32_t sh_lps = prmt(0u, e.y, q4);
                        const uint32_t rm22 = R22 - lps22;
                        const bool is_lps = hi >= rm22;
                        const uint32_t sh_mps = (rm22 >> 30) ^ 1u;  // codIRange - rangeLPS >= 128: at most one doubling
                        const uint32_t r22 = is_lps ? lps22 : rm22;
                        const uint32_t sh = is_lps ? sh_lps : sh_mps;
                        const uint32_t hi_lps = hi - rm22;
                        hi = is_lps ? hi_lps : hi;
                        const uint32_t sel = prmt(e.z, 0u, is_lps ? sel_lps : sel_mps);  // next state | bin << 31
                        sts_u8(addr, sel);
                        R22 = r22 << sh;
                        hi = __funnelshift_l(lo, hi, sh);
                        lo <<= sh;
                        fbits -= (int32_t)sh;
                        word = __funnelshift_l(sel, word, 1);
                        s1 = s2;
                        if (__builtin_expect((hzm >> u) & 1u, 0)) {  // this context again within two decisions
                            __syncwarp();
                            if ((f1m >> u) & 1u) e_cur = lds_u32x4(tabl + (sel & 0xFFu) * 128u);
                            if ((f2m >> u) & 1u) s1 = sel & 0xFFu;
                        }
                    } else if (bm & (1u << u)) {
                        hi = __funnelshift_l(lo, hi, 1);
                        lo <<= 1;
                        fbits -= 1;
                        const bool one = hi >= R22;
                        if (one) hi -= R22;
                        word = (word << 1) | (one ? 1u : 0u);
                    } else {
                        w.R = R22 >> 22, w.hi = hi, w.lo = lo, w.fbits = fbits;
                        const uint32_t bin = w.terminate();
                        R22 = w.R << 22, hi = w.hi, lo = w.lo, fbits = w.fbits;
                        word = (word << 1) | bin;
                        if (__any_sync(0xFFFFFFFFu, bin)) {  // a slice that goes on after its end: the generic loop takes over
                            if (bin) eng.to_literal();
                            left = true;
                            k = k8 + u + 1u;
                            break;
                        }
                    }
                }
            }
            if (!left) {
                i += 32u;
                if (own) bins[(i >> 5) - 1u] = __brev(word);
                word = 0;
            } else {
                i += k;
                word = k < 32u ? __brev(word) >> (32u - k) : __brev(word);  // bins 0..k-1 in bits 0..k-1
                if (k == 32u) {
                    if (own) bins[(i >> 5) - 1u] = word;
                    word = 0;
                }
            }
        }
        if (!eng.lit) w.R = R22 >> 22, w.hi = hi, w.lo = lo, w.fbits = fbits;
      }
    } else
    if (__all_sync(0xFFFFFFFFu, my_ops >= 32u && !eng.lit)) {
        CabacLane &w = eng.w;
        const uint32_t st_lane = opaque(smem_addr(s_state) + (uint32_t)lane);  // &state[0][lane]
        const uint32_t tab_fast = opaque(smem_addr(s_tab_fast));
        const uint32_t sel_mps = opaque(0x1440u), sel_lps = opaque(0x3442u);  // byte-permute selectors, kept in registers
        uint32_t R22 = w.R << 22, hi = w.hi, lo = w.lo;
        int32_t fbits = w.fbits;
        bool left = false;
#if H264B_CABAC_PIPELINE
        uint2 e_cur = make_uint2(0u, 0u);  // table entry of the next decision
        uint32_t s1 = 0;                   // state byte of the decision after that
#endif
        // the ops of a block are loaded one block ahead (a global load per 32 ops that is never waited for)
        uint32_t next_op = i + (uint32_t)lane < j.n_ops_max ? (uint32_t)j.ops[i + (uint32_t)lane] : 0u;
        while (!left && __all_sync(0xFFFFFFFFu, i + 32u <= my_ops)) {
            const uint32_t my_op = next_op;
            next_op = i + 32u + (uint32_t)lane < j.n_ops_max ? (uint32_t)j.ops[i + 32u + (uint32_t)lane] : 0u;
            const uint32_t my_kind = my_op >> 14;
            const uint32_t dec_mask = __ballot_sync(0xFFFFFFFFu, my_kind == H264B_OP_DECISION);
            const uint32_t byp_mask = __ballot_sync(0xFFFFFFFFu, my_kind == H264B_OP_BYPASS);
            uint32_t my_row = ((my_op & 0x3FFu) < n_ctx ? (my_op & 0x3FFu) : 0u) * 32u;  // as below: ctx 0
#if H264B_CABAC_PIPELINE
            // Loads ahead of the arithmetic (the schedule is known, only the states are data): a decision requests the
            // state byte of the second decision after it and, once its own new state is known, the table entry of the
            // next one.  The new state is forwarded into both when they use the same context (two ballots per block), so
            // there is no hazard path and no branch; the pipeline restarts at every block.
            const uint32_t above = dec_mask & ~((2u << lane) - 1u);
            const uint32_t above2 = above & (above - 1u);
            const uint32_t nrow1 = __shfl_sync(0xFFFFFFFFu, my_row, above ? __ffs((int)above) - 1 : lane);
            uint32_t my_nn = __shfl_sync(0xFFFFFFFFu, my_row, above2 ? __ffs((int)above2) - 1 : lane);
            // the state a decision writes is forwarded to the loads already made for the next two decisions when they
            // use the same context
            const uint32_t fwd1_mask = __ballot_sync(0xFFFFFFFFu, above != 0u && nrow1 == my_row);
            const uint32_t fwd2_mask = __ballot_sync(0xFFFFFFFFu, above2 != 0u && my_nn == my_row);
            if (!above2) my_nn = 0u;  // (no second decision behind this op in the block: a harmless load of row 0)
            {   // the block's first two decisions: entry of the first, state of the second (once per block, in place)
                const uint32_t rest = dec_mask & (dec_mask - 1u);
                const uint32_t row_d0 = __shfl_sync(0xFFFFFFFFu, my_row, dec_mask ? __ffs((int)dec_mask) - 1 : 0);
                const uint32_t row_d1 = __shfl_sync(0xFFFFFFFFu, my_row, rest ? __ffs((int)rest) - 1 : 0);
                e_cur = lds_u32x2(tab_fast + lds_u8(row_d0 + st_lane) * 8u);
                s1 = lds_u8(row_d1 + st_lane);
            }
#endif
            uint32_t k = 0;
#pragma unroll 1
            for (uint32_t k8 = 0; k8 < 32u && !left; k8 += 8u) {
                // lanes 0..7 hold the rows of this chunk's ops (the rows rotate by 8 lanes per chunk), so the shuffles
                // below have constant source lanes
                const uint32_t dm = dec_mask >> k8, bm = byp_mask >> k8;
                const uint32_t row8 = my_row;
                my_row = __shfl_sync(0xFFFFFFFFu, my_row, (lane + 8) & 31);
#if H264B_CABAC_PIPELINE
                const uint32_t f1m = fwd1_mask >> k8, f2m = fwd2_mask >> k8, nn8 = my_nn;
                my_nn = __shfl_sync(0xFFFFFFFFu, my_nn, (lane + 8) & 31);
#endif
#pragma unroll
                for (uint32_t u = 0; u < 8u; u++) {
                    if ((u & 1u) == 0u) {
                        if (__builtin_expect(__any_sync(0xFFFFFFFFu, fbits < 16), 0)) {
                            __syncwarp();  // (also keeps this rare block a branch instead of 20 predicated instructions)
                            if (fbits <= 22) {
                                w.hi = hi, w.lo = lo, w.fbits = fbits;
                                w.refill();
                                hi = w.hi, lo = w.lo, fbits = w.fbits;
                            }
                        }
                    }
                    if (dm & (1u << u)) {
#if H264B_CABAC_PIPELINE
                        const uint32_t s2 = lds_u8(__shfl_sync(0xFFFFFFFFu, nn8, (int)u) + st_lane);  // state of the one after next
                        const uint32_t addr = __shfl_sync(0xFFFFFFFFu, row8, (int)u) + st_lane;
                        const uint2 e = e_cur;
#else
                        const uint32_t addr = __shfl_sync(0xFFFFFFFFu, row8, (int)u) + st_lane;
                        const uint32_t st = lds_u8(addr);
                        const uint2 e = lds_u32x2(tab_fast + st * 8u);
#endif
                        const uint32_t lps22 = prmt(0u, e.x, R22 >> 28) << 22;  // rangeTabLPS[state][q] << 22
                        const uint32_t rm22 = R22 - lps22;
                        const bool is_lps = hi >= rm22;
                        const uint32_t r22 = is_lps ? lps22 : rm22;
                        const uint32_t hi_lps = hi - rm22;
                        hi = is_lps ? hi_lps : hi;
                        const uint32_t sel = prmt(e.y, 0u, is_lps ? sel_lps : sel_mps);  // next state | bin << 31
                        sts_u8(addr, sel);
                        const uint32_t sh = (uint32_t)__clz((int)r22) - 1u;
                        R22 = r22 << sh;
                        hi = __funnelshift_l(lo, hi, sh);
                        lo <<= sh;
                        fbits -= (int32_t)sh;
                        word = __funnelshift_l(sel, word, 1);
#if H264B_CABAC_PIPELINE
                        // forwarding as one bitwise select each: the masks are 0xFF or 0 (warp-uniform), states are bytes
                        const uint32_t m1 = (0u - ((f1m >> u) & 1u)) & 0xFFu, m2 = (0u - ((f2m >> u) & 1u)) & 0xFFu;
                        e_cur = lds_u32x2(tab_fast + ((sel & m1) | (s1 & ~m1)) * 8u);  // next decision's entry
                        s1 = (sel & m2) | (s2 & ~m2);
#endif
                    } else if (bm & (1u << u)) {
                        hi = __funnelshift_l(lo, hi, 1);
                        lo <<= 1;
                        fbits -= 1;
                        const bool one = hi >= R22;
                        if (one) hi -= R22;
                        word = (word << 1) | (one ? 1u : 0u);
                    } else {
                        w.R = R22 >> 22, w.hi = hi, w.lo = lo, w.fbits = fbits;
                        const uint32_t bin = w.terminate();
                        R22 = w.R << 22, hi = w.hi, lo = w.lo, fbits = w.fbits;
                        word = (word << 1) | bin;
                        if (__any_sync(0xFFFFFFFFu, bin)) {  // a slice that goes on after its end: the generic loop takes over
                            if (bin) eng.to_literal();
                            left = true;
                            k = k8 + u + 1u;
                            break;
                        }
                    }
                }
            }
            if (!left) {
                i += 32u;
                if (own) bins[(i >> 5) - 1u] = __brev(word);
                word = 0;
            } else {
                i += k;
                word = k < 32u ? __brev(word) >> (32u - k) : __brev(word);  // bins 0..k-1 in bits 0..k-1
                if (k == 32u) {
                    if (own) bins[(i >> 5) - 1u] = word;
                    word = 0;
                }
            }
        }
        if (!eng.lit) w.R = R22 >> 22, w.hi = hi, w.lo = lo, w.fbits = fbits;
    }
    // ---- the same block structure for warps whose lanes are all on the literal engine (the reference's own bypass form,
    // H264B_BYPASS_SPEC_OR clear): uniform op kinds by ballot, ops one block ahead, explicit shared addresses; the
    // arithmetic is LiteralLane's (64-bit codIOffset, stream bits from its window).  A terminate bin of 1 does not end
    // anything here: the literal engine simply goes on, like the reference.
    if (__all_sync(0xFFFFFFFFu, (i & 31u) == 0u && eng.lit && i + 32u <= my_ops)) {  // (i is warp-uniform; a vote says so)
        LiteralLane &l = eng.l;
        const uint32_t st_lane = opaque(smem_addr(s_state) + (uint32_t)lane);
        // (entry of a state byte: rangeLPS word, next-state word; kLoop 1 keeps them in the replicated 16-byte table, whose
        //  128 rows need the state byte without a stray bit 7 of caller-given initial states)
        const uint32_t tab_fast = opaque(kLoop ? smem_addr(s_tab16) + (uint32_t)(lane & 7) * 16u : smem_addr(s_tab_fast));
        const auto entry_of = [&](uint32_t st) -> uint2 {
            if (kLoop) {
                const uint4 v = lds_u32x4(tab_fast + (st & 127u) * 128u);
                return make_uint2(v.x, v.z);
            }
            return lds_u32x2(tab_fast + st * 8u);
        };
        uint32_t next_op = i + (uint32_t)lane < j.n_ops_max ? (uint32_t)j.ops[i + (uint32_t)lane] : 0u;
        while (__all_sync(0xFFFFFFFFu, i + 32u <= my_ops)) {
            const uint32_t my_op = next_op;
            next_op = i + 32u + (uint32_t)lane < j.n_ops_max ? (uint32_t)j.ops[i + 32u + (uint32_t)lane] : 0u;
            const uint32_t my_kind = my_op >> 14;
            const uint32_t dec_mask = __ballot_sync(0xFFFFFFFFu, my_kind == H264B_OP_DECISION);
            const uint32_t byp_mask = __ballot_sync(0xFFFFFFFFu, my_kind == H264B_OP_BYPASS);
            uint32_t my_row = ((my_op & 0x3FFu) < n_ctx ? (my_op & 0x3FFu) : 0u) * 32u;
            // the same load pipeline as the window loop above: entry of the next decision and state of the one after it
            // in flight, the state a decision writes forwarded into them where the contexts coincide
            const uint32_t above = dec_mask & ~((2u << lane) - 1u);
            const uint32_t above2 = above & (above - 1u);
            const uint32_t nrow1 = __shfl_sync(0xFFFFFFFFu, my_row, above ? __ffs((int)above) - 1 : lane);
            uint32_t my_nn = __shfl_sync(0xFFFFFFFFu, my_row, above2 ? __ffs((int)above2) - 1 : lane);
            const uint32_t fwd1_mask = __ballot_sync(0xFFFFFFFFu, above != 0u && nrow1 == my_row);
            const uint32_t fwd2_mask = __ballot_sync(0xFFFFFFFFu, above2 != 0u && my_nn == my_row);
            if (!above2) my_nn = 0u;
            uint2 e_cur;
            uint32_t s1;
            {
                const uint32_t rest = dec_mask & (dec_mask - 1u);
                const uint32_t row_d0 = __shfl_sync(0xFFFFFFFFu, my_row, dec_mask ? __ffs((int)dec_mask) - 1 : 0);
                const uint32_t row_d1 = __shfl_sync(0xFFFFFFFFu, my_row, rest ? __ffs((int)rest) - 1 : 0);
                e_cur = entry_of(lds_u8(row_d0 + st_lane));
                s1 = lds_u8(row_d1 + st_lane);
            }
#pragma unroll 1
            for (uint32_t k8 = 0; k8 < 32u; k8 += 8u) {
                const uint32_t dm = dec_mask >> k8, bm = byp_mask >> k8;
                const uint32_t f1m = fwd1_mask >> k8, f2m = fwd2_mask >> k8;
                const uint32_t row8 = my_row, nn8 = my_nn;
                my_row = __shfl_sync(0xFFFFFFFFu, my_row, (lane + 8) & 31);
                my_nn = __shfl_sync(0xFFFFFFFFu, my_nn, (lane + 8) & 31);
#pragma unroll
                for (uint32_t u = 0; u < 8u; u++) {
                    if ((u & 1u) == 0u) {  // two ops take at most 2 x 9 stream bits
                        if (__any_sync(0xFFFFFFFFu, l.avail < 18u)) {
                            __syncwarp();
                            if (l.avail <= 32u) l.top_up();
                        }
                    }
                    if (dm & (1u << u)) {
                        const uint32_t s2 = lds_u8(__shfl_sync(0xFFFFFFFFu, nn8, (int)u) + st_lane);
                        const uint32_t addr = __shfl_sync(0xFFFFFFFFu, row8, (int)u) + st_lane;
                        const uint2 e = e_cur;
                        // LiteralLane::decision on the fast table's form (bin in bit 7 of the state bytes)
                        const int64_t lps = (int64_t)prmt(e.x, 0u, 0x4440u | ((uint32_t)(l.R >> 6) & 3u));
                        l.R -= lps;
                        const bool is_lps = l.O >= l.R;
                        l.O = is_lps ? (int64_t)((uint64_t)l.O - (uint64_t)l.R) : l.O;
                        l.R = is_lps ? lps : l.R;
                        const uint32_t sel = prmt(e.y, 0u, is_lps ? 0x3442u : 0x1440u);  // next state | bin << 31
                        sts_u8(addr, sel);
                        const uint32_t m1 = (0u - ((f1m >> u) & 1u)) & 0xFFu, m2 = (0u - ((f2m >> u) & 1u)) & 0xFFu;
                        e_cur = entry_of((sel & m1) | (s1 & ~m1));
                        s1 = (sel & m2) | (s2 & ~m2);
                        l.renorm<false>();
                        word = __funnelshift_l(sel, word, 1);
                    } else if (bm & (1u << u)) {
                        word = (word << 1) | l.bypass<false>();
                    } else {
                        word = (word << 1) | l.terminate<false>();
                    }
                }
            }
            i += 32u;
            if (own) bins[(i >> 5) - 1u] = __brev(word);
            word = 0;
        }
    }
    // ---- generic loop: lanes that have finished, lanes on the literal engine
    warp_ops = live ? my_ops : 0u;
#pragma unroll
    for (int d = 16; d; d >>= 1) warp_ops = max(warp_ops, __shfl_xor_sync(0xFFFFFFFFu, warp_ops, d));
    uint32_t next_op = i < warp_ops ? j.ops[i] : 0;
    for (; i < warp_ops; i++) {
        const uint32_t op = next_op;
        if (i + 1 < warp_ops) next_op = j.ops[i + 1];
        generic_op(i, op, live && i < my_ops);
    }
    if (live) finish_lane();
    __syncwarp();  // every lane is done with the context rows before the next bundle's states go there
    }
}

static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

int launch_cabac(h264b_ctx *ctx, const h264b_cabac_job *job, const uint32_t *d_n_slices) {
    TraceRange trace_range("h264b:cabac_decode");
    const h264b_cabac_job &j = *job;
    if (j.n_ctx < 1 || j.n_ctx > 1024) return set_error(ctx, H264B_E_INVALID, "cabac: n_ctx must be 1..1024");
    if (!j.n_slices) return H264B_OK;
    if (!j.bytes || !j.off || !j.len || !j.bins || !j.final || (!j.ops && j.n_ops_max))
        return set_error(ctx, H264B_E_INVALID, "cabac: null pointer in job");
    if (!j.qp && !j.init_states) return set_error(ctx, H264B_E_INVALID, "cabac: need qp or init_states");
    if (!j.bins_off && j.bins_stride_words < (j.n_ops_max + 1 + 31) / 32)
        return set_error(ctx, H264B_E_INVALID, "cabac: bins_stride_words too small");
    if ((uintptr_t)j.bytes & 3) return set_error(ctx, H264B_E_INVALID, "cabac: bytes must be 4-byte aligned");
    // measurement knobs (tools/cabac_balance_exp.py); the defaults are the shipped configuration
    // the branch-free loop (2) is the window engine's; jobs in the reference's own bypass form run on the literal engine,
    // whose block loop is fastest in the instantiation with the 8-byte table (0)
    const int k_loop = env_int("H264B_CABAC_LOOP", (j.flags & H264B_BYPASS_SPEC_OR) ? 2 : 0), k_w = env_int("H264B_CABAC_W", 0),
              k_map = env_int("H264B_CABAC_MAP", 3);
    const int v = (j.flags & H264B_TABLES_SPEC) ? 1 : 0;
    CabacArgs a;
    a.j = j;
    a.tab = ctx->d_cabac_tab[v];
    a.lut = ctx->d_state_lut[v];
    // Few slices: spread them one (or a few) per warp so every slice gets its own scheduler slot and no lane waits on
    // a neighbour's bank conflict; many slices: 32 per warp.
    const uint32_t sms = (uint32_t)ctx->sm_count;
    const uint32_t target_warps = sms * 4;  // one warp per scheduler before lanes are doubled up
    uint32_t lpw = (j.n_slices + target_warps - 1) / target_warps;
    if (lpw < 1) lpw = 1;
    if (lpw > 32) lpw = 32;
    const bool exclusive = ctx->cabac_exclusive != 0;
    if (exclusive) lpw = 1;
    else if (ctx->cabac_pack) lpw = 32;
    a.lanes_per_warp = lpw;
    a.n_warps = (j.n_slices + lpw - 1) / lpw;
    a.order = nullptr;
    a.d_n = d_n_slices;
    a.map_mode = (uint32_t)k_map;
    // One wave of one CTA per SM with W warps, each warp with its own n_ctx x 32 bytes of context rows; what does not fit
    // one wave runs as small CTAs in launch order (the hardware hands them out as earlier ones finish).
    const size_t tab_bytes = 1024 + (k_loop ? kTab16Bytes : 2048);
    // the schedule's context working set, when the caller names it: only those rows live in shared memory
    const uint32_t n_rows = j.n_ctx_used && j.n_ctx_used < j.n_ctx ? j.n_ctx_used : j.n_ctx;
    a.n_rows = n_rows;
    uint32_t w_fit = (uint32_t)((kMaxSmemPerCta - tab_bytes) / ((size_t)(n_rows + 2) * 32));
    if (w_fit > (uint32_t)kMaxWarpsPerCta) w_fit = kMaxWarpsPerCta;
    if (ctx->cabac_max_warps > 0 && w_fit > (uint32_t)ctx->cabac_max_warps) w_fit = (uint32_t)ctx->cabac_max_warps;
    // bundles -> schedulers by bundle_assign_kernel: one wave of full CTAs with a slot to spare on every scheduler; when
    // the context rows leave room for fewer warps than there are bundles, every warp runs several bundles in turn
    uint32_t slots = (a.n_warps + sms * 4 - 1) / (sms * 4) + 1;
    if (slots > (uint32_t)kSlotsMax) slots = kSlotsMax;
    if (slots * 4 > w_fit) slots = w_fit / 4;
    uint32_t rounds = slots ? (uint32_t)((a.n_warps + (uint64_t)sms * 4 * slots - 1) / ((uint64_t)sms * 4 * slots)) : 0;
    const bool assign = !exclusive && a.map_mode == 3 && k_w <= 0 && lpw > 1 && j.n_ops && slots >= 1 && rounds >= 1 && rounds <= 64 &&
                        a.n_warps > sms * 4 && sms * 4 <= 640 && ((uint64_t)sms * 4 * slots * rounds + 2ull * sms * 4) * 4 <= 200 * 1024;
    if (a.map_mode == 3 && !assign) a.map_mode = 1;
    uint32_t W, grid;
    if (exclusive) {  // slices in the caller's order, four to an SM, each warp on a scheduler of its own
        W = w_fit < 4 ? w_fit : 4;
        if (W < 1) W = 1;
        grid = (a.n_warps + W - 1) / W;
        a.map_mode = 2;
    } else if (assign) {
        W = slots * 4;
        grid = sms;
    } else {
        W = (a.n_warps + sms - 1) / sms;
        if (k_w > 0) W = (uint32_t)k_w;
        if (W > w_fit || (k_w > 0 && k_w < 4)) {
            W = w_fit < 4 ? w_fit : 4;
            if (k_w > 0 && k_w < 4) W = (uint32_t)k_w;
            a.map_mode = 2;
        }
        if (W < 1) W = 1;
        grid = (a.n_warps + W - 1) / W;
    }
    a.slot_bundle = nullptr;
    a.rounds = 1;
    if (lpw > 1 && j.n_ops) {  // bundles of equally long slices
        void *d_sort;
        const size_t order_bytes = ((size_t)j.n_slices * 4 + 15) & ~(size_t)15;
        int rc = ensure_dev(ctx, 16, sizeof(SortScratch) + order_bytes + (size_t)sms * 4 * (assign ? slots * rounds : 1) * 4, &d_sort);
        if (rc) return rc;
        SortScratch *ss = (SortScratch *)d_sort;
        uint32_t *order = (uint32_t *)(ss + 1);
        H264B_CUDA(ctx, cudaMemsetAsync(ss, 0, sizeof(SortScratch), ctx->stream));
        H264B_CUDA(ctx, cudaMemsetAsync(&ss->min_ops, 0xFF, 4, ctx->stream));
        const int sb = (int)((j.n_slices + 255) / 256 < sms * 4 ? (j.n_slices + 255) / 256 : sms * 4);
        sort_minmax_kernel<<<sb, 256, 0, ctx->stream>>>(j.n_ops, j.n_slices, d_n_slices, j.n_ops_max, ss);
        H264B_LAUNCH_CHECK(ctx, "sort_minmax_kernel");
        sort_hist_kernel<<<sb, 256, 0, ctx->stream>>>(j.n_ops, j.n_slices, d_n_slices, j.n_ops_max, ss);
        H264B_LAUNCH_CHECK(ctx, "sort_hist_kernel");
        sort_scan_kernel<<<1, 1024, 0, ctx->stream>>>(ss);
        H264B_LAUNCH_CHECK(ctx, "sort_scan_kernel");
        sort_scatter_kernel<<<sb, 256, 0, ctx->stream>>>(j.n_ops, j.n_slices, d_n_slices, j.n_ops_max, ss, order);
        H264B_LAUNCH_CHECK(ctx, "sort_scatter_kernel");
        a.order = order;
        if (assign) {
            int32_t *slot_bundle = (int32_t *)((uint8_t *)order + order_bytes);
            const size_t sm_assign = ((size_t)sms * 4 * slots * rounds + 2 * (size_t)sms * 4) * 4;
            if (sm_assign > 48 * 1024)
                H264B_CUDA(ctx, cudaFuncSetAttribute(bundle_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_assign));
            bundle_assign_kernel<<<1, 256, sm_assign, ctx->stream>>>(j.n_ops, order, j.n_slices, d_n_slices, j.n_ops_max, lpw,
                                                                     sms * 4, slots, rounds, W, slot_bundle);
            H264B_LAUNCH_CHECK(ctx, "bundle_assign_kernel");
            a.slot_bundle = slot_bundle;
            a.rounds = rounds;
        }
    }
    const size_t smem = exclusive ? (size_t)kMaxSmemPerCta : tab_bytes + (size_t)W * (n_rows + 2) * 32;
    // (one carveout for every launch of these kernels: launches that differ in their shared-memory split cannot share an
    //  SM, and h264b_scheduler runs several side by side; half of the SM's 228 KB: room for every launch shape at 64
    //  contexts, and an L1 for the op schedule and the bitstream words; launches that need more get more)
    const auto launch = [&](auto kernel, int which) -> int {
        static bool attr_set[6][64] = {{false}};
        if (ctx->device >= 64 || !attr_set[which][ctx->device]) {
            H264B_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                                 env_int("H264B_CABAC_CARVEOUT", 50)));
            H264B_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmemPerCta));
            if (ctx->device < 64) attr_set[which][ctx->device] = true;
        }
        kernel<<<grid, W * 32, smem, ctx->stream>>>(a);
        return H264B_OK;
    };
    int lrc;
    const bool multi = a.rounds > 1;
    if (k_loop == 2) {
        lrc = multi ? launch(cabac_decode_kernel<2, true>, 0) : launch(cabac_decode_kernel<2, false>, 1);
    } else if (k_loop) {
        lrc = multi ? launch(cabac_decode_kernel<1, true>, 2) : launch(cabac_decode_kernel<1, false>, 3);
    } else {
        lrc = multi ? launch(cabac_decode_kernel<0, true>, 4) : launch(cabac_decode_kernel<0, false>, 5);
    }
    if (lrc) return lrc;
    H264B_LAUNCH_CHECK(ctx, "cabac_decode_kernel");
    return H264B_OK;
}

// ---------------------------------------------------------------------------------------------- scalar drop-ins
struct StepIo {
    int64_t R, O;
    int32_t p_state, val_mps, bin, pad;
    uint32_t bits_used, n_bits;
    uint8_t bits[32];
};

// One engine primitive with explicit state, bit-at-a-time like the reference (used only for per-call drop-in of
// the Go functions; the batch kernel above is the throughput path).
__global__ void engine_step_kernel(StepIo *io, uint32_t kind, uint32_t flags, const uint64_t *tab) {
    int64_t R = io->R, O = io->O;
    uint32_t used = 0;
    auto read_bit = [&]() -> uint32_t {
        uint32_t b = 0;
        if (used < io->n_bits && used < 256) b = (io->bits[used >> 3] >> (7 - (used & 7))) & 1u;
        used++;
        return b;
    };
    auto renorm = [&]() {
        while (R < 256) {
            R <<= 1;
            O = (int64_t)(((uint64_t)O << 1) | read_bit());
        }
    };
    int32_t bin = 0;
    if (kind == H264B_OP_DECISION) {
        const uint32_t s = (uint32_t)(io->p_state & 63) | ((uint32_t)(io->val_mps & 1) << 6);
        const uint64_t e = tab[s];
        const uint32_t tlo = (uint32_t)e, thi = (uint32_t)(e >> 32);
        const int64_t lps = (tlo >> ((uint32_t)((R >> 6) & 3) * 8)) & 0xFF;
        R -= lps;
        uint32_t sel = thi;
        if (O >= R) {
            O = (int64_t)((uint64_t)O - (uint64_t)R);
            R = lps;
            sel = thi >> 16;
        }
        bin = (sel >> 8) & 1;
        if (!(flags & 0x80000000u)) {  // composed DecodeDecision; the bare BinaryDecision core stops here (A6)
            io->p_state = sel & 63;
            io->val_mps = (sel >> 6) & 1;
            renorm();
        }
    } else if (kind == H264B_OP_BYPASS) {
        uint64_t o = (uint64_t)O << 1;
        const uint32_t b = read_bit();
        o = (flags & H264B_BYPASS_SPEC_OR) ? (o | b) : (o << b);
        O = (int64_t)o;
        if (O >= R) {
            O = (int64_t)((uint64_t)O - (uint64_t)R);
            bin = 1;
        }
    } else if (kind == H264B_OP_TERMINATE) {
        R -= 2;
        if (O >= R) {
            bin = 1;
        } else {
            renorm();
        }
    } else if (kind == 3u) {  // StateTransitionProcess alone: bin given in io->bin
        const uint32_t s = (uint32_t)(io->p_state & 63) | ((uint32_t)(io->val_mps & 1) << 6);
        const uint32_t thi = (uint32_t)(tab[s] >> 32);
        const uint32_t sel = (io->bin == io->val_mps) ? thi : (thi >> 16);
        io->p_state = sel & 63;
        io->val_mps = (sel >> 6) & 1;
        bin = io->bin;
    } else if (kind == 4u) {  // RenormD alone
        renorm();
    } else if (kind == 5u) {  // initDecodingEngine
        R = 510;
        O = 0;
        for (int i = 0; i < 9; i++) O = (O << 1) | read_bit();
    }
    io->R = R;
    io->O = O;
    io->bin = bin;
    io->bits_used = used;
}

}  // namespace h264b

extern "C" int32_t h264b_engine_step(h264b_ctx *ctx, uint32_t kind, uint32_t flags, const uint8_t *bits,
                                     uint32_t n_bits, int64_t *R, int64_t *O, int32_t *p_state, int32_t *val_mps,
                                     int32_t *bin_val, uint32_t *bits_used) {
    using namespace h264b;
    if (!ctx || !R || !O) return H264B_E_INVALID;
    if (n_bits > 256) n_bits = 256;
    StepIo h;
    memset(&h, 0, sizeof(h));
    h.R = *R;
    h.O = *O;
    h.p_state = p_state ? *p_state : 0;
    h.val_mps = val_mps ? *val_mps : 0;
    h.bin = bin_val ? *bin_val : 0;
    h.n_bits = n_bits;
    if (bits && n_bits) memcpy(h.bits, bits, (n_bits + 7) / 8);
    void *d;
    int rc = ensure_dev(ctx, 15, sizeof(StepIo), &d);
    if (rc) return rc;
    H264B_CUDA(ctx, cudaMemcpyAsync(d, &h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
    engine_step_kernel<<<1, 1, 0, ctx->stream>>>((StepIo *)d, kind, flags,
                                                 ctx->d_cabac_tab[(flags & H264B_TABLES_SPEC) ? 1 : 0]);
    H264B_LAUNCH_CHECK(ctx, "engine_step_kernel");
    H264B_CUDA(ctx, cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    H264B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *R = h.R;
    *O = h.O;
    if (p_state) *p_state = h.p_state;
    if (val_mps) *val_mps = h.val_mps;
    if (bin_val) *bin_val = h.bin;
    if (bits_used) *bits_used = h.bits_used;
    return H264B_OK;
}

extern "C" int32_t h264b_binary_decision(h264b_ctx *ctx, uint32_t flags, int32_t p_state, int32_t val_mps, int64_t *R,
                                         int64_t *O, int32_t *bin_val) {
    return h264b_engine_step(ctx, H264B_OP_DECISION, flags | 0x80000000u, nullptr, 0, R, O, &p_state, &val_mps,
                             bin_val, nullptr);
}

extern "C" int32_t h264b_state_transition(h264b_ctx *ctx, uint32_t flags, int32_t *p_state, int32_t *val_mps,
                                          int32_t bin_val) {
    int64_t R = 0, O = 0;
    return h264b_engine_step(ctx, 3u, flags, nullptr, 0, &R, &O, p_state, val_mps, &bin_val, nullptr);
}
