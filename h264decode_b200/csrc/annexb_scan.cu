// annexb_scan.cu -- K1/K2: single-pass Annex-B split + RBSP emulation-prevention strip for sm_100a.
//
// Replaces the byte-at-a-time loops of readNalUnit (h264/server.go:64-111) and NewNalUnit (h264/nalUnit.go:75-131)
// of the reference with one streaming pass: every input byte is read from HBM once and every kept byte written once
// (algorithmic traffic N_in + N_rbsp + 20 B per NAL).
//
// Structure (one CTA = 256 threads, tiles of 16 KiB handed out by an atomic ticket so that tile i is always started
// before tile i+1):
//   load     one elected thread issues a TMA 1-D bulk copy (cp.async.bulk + mbarrier complete_tx) of the tile plus a
//            16-byte halo on each side into shared memory
//   detect   each thread reads 16-byte granules (LDS.128, conflict-free interleaved mapping), finds zero bytes with
//            word-parallel arithmetic and only where two zeros precede a byte looks for 03 (emulation prevention)
//            and 01 (start-code end); start-code bits go to a per-tile bitmap
//   adjust   granules with a start code within reach (header bytes, the 2-byte tail rule, zeros that belong to a
//            header) recompute their 16 keep bits from the start-code bitmap (keep_mask_near_sc, annexb_local.cuh)
//   scan     packed (kept bytes | start codes << 16) block scan: shuffle scan per warp, 32 warp totals by warp 0
//   chain    decoupled look-back over per-tile descriptors gives the tile's exclusive (kept bytes, NAL index) prefix
//   compact  the few 512-byte rows that lose bytes are compacted in place in the tile buffer (overlaps the look-back)
//   store    every row is now `len` contiguous bytes: lanes exchange neighbour words by shuffle so that each lane
//            writes one aligned 16-byte granule of the destination; only a row's two ragged ends use byte stores
//   index    threads owning a start code write (start, rbsp offset, first 4 NAL bytes) for NAL k = prefix + rank
// A tiny finalize kernel turns those into h264b_nal records (lengths are differences of neighbours).
#include "annexb_local.cuh"
#include "common.cuh"

namespace h264b {

constexpr int kThreads = 256;
constexpr int kRows = 4;                          // granules per thread per tile
constexpr int kGranules = kThreads * kRows;       // 1024
constexpr int kTile = kGranules * 16;             // 16384 bytes
constexpr int kHalo = 16;
constexpr int kInBytes = kHalo + kTile + kHalo;   // 16416
constexpr uint64_t kStatusAgg = 1ull << 62, kStatusPrefix = 2ull << 62, kValueMask = (1ull << 62) - 1;

struct ScanScratchHeader {   // device scratch, zeroed / initialised by scan_init_kernel
    unsigned int ticket;
    unsigned int pad;
    unsigned long long first_start;
    unsigned long long total_sc;   // start codes found = slots handed out of the NAL record buffer
    unsigned long long total_kept;
    unsigned long long n_epb;
    unsigned int n_fix;            // entries of fix_list
    unsigned int pad2;
    unsigned long long reserved[2];
};  // 64 bytes

struct ScanArgs {
    const uint8_t *in;
    uint64_t n;
    uint8_t *out;
    ScanScratchHeader *hdr;
    uint32_t *tile_tot;        // per tile: packed segmented total (seg_combine format): NAL start? | start codes | EPBs
    uint32_t *tile_slot;       // per tile: first slot of its records in the unordered record buffer
    uint32_t *tile_ord;        // per tile: ordinal of its first start code (exclusive scan of the start-code counts)
    uint32_t *fix_list;        // NAL ordinals whose later pieces must slide left (written by scan_finalize_kernel)
    // per start code, in slot order (tiles reserve slots with one atomicAdd): written by the main pass
    unsigned long long *rec_start;
    unsigned long long *rec_epb;
    uint32_t *rec_hdr;
    // the same in stream order (written by nal_permute_kernel, read by scan_finalize_kernel)
    unsigned long long *nal_start;
    unsigned long long *nal_epb;  // [k]: EPBs removed from the NAL that ends at start code k
    uint32_t *nal_hdr;
    uint32_t nal_cap;
    uint32_t n_tiles;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// 16-byte tile descriptors: one 128-bit transaction each way (L2 is the point of coherence: .cg / volatile)
__device__ __forceinline__ ulonglong2 ld_desc(const ulonglong2 *p) {
    ulonglong2 v;
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_desc(ulonglong2 *p, unsigned long long x, unsigned long long y) {
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(x), "l"(y) : "memory");
}

// ------------------------------------------------------------------------------------------------ init
__global__ void scan_init_kernel(ScanScratchHeader *hdr, uint64_t n) {
    hdr->ticket = 0;
    hdr->first_start = n;
    hdr->total_sc = 0;
    hdr->total_kept = 0;
    hdr->n_epb = 0;
    hdr->n_fix = 0;
}

// ------------------------------------------------------------------------------------------------ first start code
// first_start = 1 + position of the first 00 00 00 01 (bytes before it are not part of any NAL, server.go:68-72).
// Chunks are visited in order by each CTA and a CTA stops as soon as a smaller position is already known.
__global__ void __launch_bounds__(256) first_start_kernel(const uint8_t *in, uint64_t n, ScanScratchHeader *hdr) {
    const uint64_t n_gran = (n + 15) / 16;
    for (uint64_t chunk = blockIdx.x;; chunk += gridDim.x) {
        uint64_t g = chunk * 256 + threadIdx.x;
        if (chunk * 256 >= n_gran) return;
        if (ld_relaxed(&hdr->first_start) <= chunk * 256 * 16) return;
        if (g < n_gran) {
            uint64_t pos = g * 16;
            uint4 v = *reinterpret_cast<const uint4 *>(in + pos);
            uint32_t w[4] = {v.x, v.y, v.z, v.w};
            uint32_t prev = pos ? *reinterpret_cast<const uint32_t *>(in + pos - 4) : 0xFFFFFFFFu;
            GranuleMasks m = granule_masks(w, prev);
            if (m.sc) {
                uint64_t q = pos + (uint64_t)(__ffs(m.sc) - 1);
                if (q < n) atomicMin(&hdr->first_start, (unsigned long long)(q + 1));
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ main pass
struct __align__(16) ScanSmem {
    uint8_t in[kInBytes];                 // [0,16): low halo, [16,16+kTile): tile, then high halo
    uint16_t scbits[kGranules + 2];       // start-code-end bits per granule, [0] = halo granule before the tile
    uint32_t row_tot[kRows * 8];          // packed segmented elements per (row, warp); then exclusive prefixes
    uint8_t row_class[kRows * 8];         // 0 untouched, 1 only EPBs removed, 2 contains NAL boundaries / stream ends
    unsigned long long nal_slot0;         // first record slot reserved for this tile's start codes
    uint32_t tile;
    unsigned long long mbar;
};

__device__ __forceinline__ void store_row(uint8_t *out, uint64_t o, uint32_t len, const uint32_t w[4], int lane,
                                          const uint8_t *prev_tail, bool next_joins) {
    uint32_t wp[4];
    if ((uint32_t)o & 15u) {  // warp-uniform: only shifted rows need the neighbour's bytes
#pragma unroll
        for (int k = 0; k < 4; k++) wp[k] = __shfl_up_sync(0xFFFFFFFFu, w[k], 1);
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) wp[k] = 0;
    }
    store_row_lane(out, o, len, wp, w, lane, prev_tail, next_joins);
}

__device__ __forceinline__ uint32_t warp_seg_scan(uint32_t x, int lane) {  // inclusive segmented scan over the warp
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
        if (lane >= d) x = seg_combine(y, x);
    }
    return x;
}


__global__ void __launch_bounds__(kThreads, 5) annexb_scan_kernel(ScanArgs a) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    ScanSmem &sm = *reinterpret_cast<ScanSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t mbar = smem_u32(&sm.mbar);

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint64_t e0 = ld_relaxed(&a.hdr->first_start);  // written by first_start_kernel (previous launch)
    const uint64_t n16 = (a.n + 15) & ~15ull;
    uint32_t parity = 0;

    for (;;) {
        // ---------------------------------------------------------------- ticket + TMA load
        if (tid == 0) sm.tile = atomicAdd(&a.hdr->ticket, 1u);
        __syncthreads();  // also: everybody is done with the previous tile's shared memory
        const uint32_t tile = sm.tile;
        if (tile >= a.n_tiles) break;
        const uint64_t base = (uint64_t)tile * kTile;
        const uint64_t lo = tile ? base - kHalo : base;
        uint64_t hi = base + kTile + kHalo;
        if (hi > n16) hi = n16;
        if (tid == 0) {
            uint32_t bytes = (uint32_t)(hi - lo);
            uint32_t dst = smem_u32(sm.in + (lo - (base - kHalo)));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                "l"(a.in + lo), "r"(bytes), "r"(mbar)
                : "memory");
        }
        // bytes outside the stream read as 0xFF (they match no predicate): low halo of tile 0 ...
        const bool edge_tile = tile == 0 || hi < base + kTile + kHalo || hi > a.n;  // CTA-uniform
        if (tile == 0 && tid < 4) reinterpret_cast<uint32_t *>(sm.in)[tid] = 0xFFFFFFFFu;
        // ... and everything the copy does not write at the end of the stream
        const uint32_t loaded_end = (uint32_t)(hi - (base - kHalo));  // offset in sm.in
        for (uint32_t o = loaded_end + tid * 4; o < (uint32_t)kInBytes; o += kThreads * 4)
            *reinterpret_cast<uint32_t *>(sm.in + o) = 0xFFFFFFFFu;
        {
            uint32_t done = 0;
            while (!done) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                    "selp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(done)
                    : "r"(mbar), "r"(parity)
                    : "memory");
            }
            parity ^= 1;
        }
        if (edge_tile) {
            if (hi > a.n && hi - a.n < 16) {  // the last 16-byte granule holds bytes past n: blank them
                uint32_t first_bad = (uint32_t)(a.n - (base - kHalo));
                if (tid < 16 && first_bad + tid < loaded_end) sm.in[first_bad + tid] = 0xFF;
            }
            __syncthreads();  // the 0xFF fills above are plain stores other threads read
        }
        uint8_t *tile_in = sm.in + kHalo;  // tile_in[i] = s[base + i], valid for i in [-16, kTile+16)
#ifdef H264B_EXP_LOADONLY
        continue;
#endif

        // ---------------------------------------------------------------- detect
        uint32_t em[kRows];  // raw EPB mask | start-code mask << 16
#pragma unroll
        for (int r = 0; r < kRows; r++) {
            const int gi = r * kThreads + tid;
            const uint4 v = *reinterpret_cast<const uint4 *>(tile_in + gi * 16);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
            const uint32_t prev = *reinterpret_cast<const uint32_t *>(tile_in + gi * 16 - 4);
#ifdef H264B_EXP_NODETECT
            GranuleMasks m;
            m.e = (w[0] == 0x12345678u && prev == 0x9ABCDEFu) ? 1u : 0u;
            m.sc = 0;
#else
            GranuleMasks m = granule_masks(w, prev);
#endif
            em[r] = m.e | (m.sc << 16);
            sm.scbits[gi + 1] = (uint16_t)m.sc;
        }
        if (tid < 2) {  // halo granules: only start codes ending in [base-6, base-1] and at base+kTile matter
            const int gi = tid ? kGranules : -1;
            const uint4 v = *reinterpret_cast<const uint4 *>(tile_in + gi * 16);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
            const uint32_t prev = tid ? *reinterpret_cast<const uint32_t *>(tile_in + gi * 16 - 4) : 0xFFFFFFFFu;
            sm.scbits[gi + 1] = (uint16_t)granule_masks(w, prev).sc;
        }
        __syncthreads();

        // ---------------------------------------------------------------- adjust + classify + per-row segmented scan
        const bool tile_has_head = base < e0;            // some bytes precede the first NAL
        const bool tile_has_end = base + kTile > a.n;    // some granules reach past the stream
        auto get = [&](int64_t p) -> uint32_t { return tile_in[p - (int64_t)base]; };
        uint32_t ks[kRows];    // keep mask | start-code mask << 16
        uint32_t ee[kRows];    // emulation-prevention bytes really removed
        uint32_t incl[kRows];  // inclusive segmented scan inside the (row, warp) group
        uint32_t cls = 0;      // 2 bits per row, warp-uniform
#pragma unroll
        for (int r = 0; r < kRows; r++) {
            const int gi = r * kThreads + tid;
            const uint64_t gpos = base + (uint64_t)gi * 16;
            uint32_t e16 = em[r] & 0xFFFFu;
            uint32_t k16 = ~e16 & 0xFFFFu;
            // start-code ends q in [g-6, g+16] change what this granule keeps
            const uint32_t near = ((uint32_t)sm.scbits[gi] >> 10) | sm.scbits[gi + 1] | (sm.scbits[gi + 2] & 1u);
            if (near)  // rare: header bytes, the 2-byte tail rule and the EPB guard, all in the bit domain
                k16 = keep_mask_near_sc(get, (int64_t)gpos, e16, sm.scbits[gi], sm.scbits[gi + 1], sm.scbits[gi + 2],
                                        &e16);
            uint32_t sc = em[r] >> 16;
            if (tile_has_head) {
                if (gpos + 16 <= e0) {
                    k16 = 0;
                    e16 = 0;
                } else if (gpos < e0) {
                    const uint32_t m = ~((1u << (uint32_t)(e0 - gpos)) - 1u);
                    k16 &= m;
                    e16 &= m;
                }
            }
            if (tile_has_end) {
                if (gpos >= a.n) {
                    k16 = 0;
                    sc = 0;
                    e16 = 0;
                } else if (gpos + 16 > a.n) {
                    const uint32_t valid = (1u << (uint32_t)(a.n - gpos)) - 1u;
                    k16 &= valid;
                    sc &= valid;
                    e16 &= valid;
                }
            }
            ks[r] = k16 | (sc << 16);
            ee[r] = e16;
            uint32_t x;
            if (__all_sync(0xFFFFFFFFu, k16 == 0xFFFFu)) {  // the usual row: nothing removed
                x = 0;
            } else if (__all_sync(0xFFFFFFFFu, (k16 | e16) == 0xFFFFu && sc == 0)) {  // only EPBs removed
                cls |= 1u << (2 * r);
                x = warp_seg_scan(bits_popc(e16), lane);
            } else {
                cls |= 2u << (2 * r);
                x = warp_seg_scan(seg_element(e16, sc), lane);
            }
            incl[r] = x;
            if (lane == 31) {
                sm.row_tot[r * 8 + warp] = x;
                sm.row_class[r * 8 + warp] = (uint8_t)((cls >> (2 * r)) & 3u);
            }
        }
        __syncthreads();

        // ---------------------------------------------------------------- tile carry (warp 0) | in-place compaction
        if (warp == 0) {
            uint32_t x = sm.row_tot[lane];
            const uint32_t own = x;
            x = warp_seg_scan(x, lane);
            uint32_t ex = __shfl_up_sync(0xFFFFFFFFu, x, 1);  // exclusive prefix of row `lane` inside the tile
            if (lane == 0) ex = 0;
            (void)own;
            sm.row_tot[lane] = ex;
            const uint32_t total = __shfl_sync(0xFFFFFFFFu, x, 31);
            const bool has_start = (total >> 31) != 0;
            const unsigned long long t_val = total & 0x7FFFu, t_nsc = (total >> 16) & 0x1FFFu;
            // No inter-tile communication: a tile counts the EPBs of its open NAL from zero (see nal_pieces in
            // annexb_local.cuh for how NALs that span tiles are finished).  It leaves its packed total for the
            // post-pass and, when it holds start codes, reserves that many record slots with one atomicAdd.
            if (lane == 0) {
                a.tile_tot[tile] = total;
                if (t_nsc) {
                    sm.nal_slot0 = atomicAdd(&a.hdr->total_sc, t_nsc);
                    a.tile_slot[tile] = (uint32_t)sm.nal_slot0;
                }
            }
            (void)has_start;
            (void)t_val;
        }
        // Rows that only lose emulation-prevention bytes are compacted in place inside their own 512-byte span of the
        // tile buffer (nobody else reads it any more): afterwards they are `len` contiguous bytes like untouched rows.
#pragma unroll
        for (int r = 0; r < kRows; r++) {
            if (((cls >> (2 * r)) & 3u) != 1u) continue;  // warp-uniform
            const int gi = r * kThreads + tid;
            const uint32_t k16 = ks[r] & 0xFFFFu;
            const uint4 v = *reinterpret_cast<const uint4 *>(tile_in + gi * 16);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
            __syncwarp();  // every lane has its bytes in registers before anyone overwrites the span
            uint32_t loff = 16u * (uint32_t)lane - (incl[r] - bits_popc(ee[r]));  // kept bytes of the lanes before
            uint8_t *row = tile_in + (r * kThreads + warp * 32) * 16;
            if (k16 == 0xFFFFu && (loff & 3u) == 0) {
                uint32_t *d = reinterpret_cast<uint32_t *>(row + loff);
                d[0] = w[0];
                d[1] = w[1];
                d[2] = w[2];
                d[3] = w[3];
            } else {
#pragma unroll
                for (int j = 0; j < 16; j++)
                    if (k16 & (1u << j)) row[loff++] = (uint8_t)(w[j >> 2] >> ((j & 3) * 8));
            }
        }
        __syncthreads();
        const uint64_t carry_in = 0;  // per-tile counting (see above)
        const uint64_t slot0 = sm.nal_slot0;

        // ---------------------------------------------------------------- store rows + NAL index
#ifdef H264B_EXP_NOSTORE
        if (slot0 == 0x123456789ull)
#endif
#pragma unroll
        for (int r = 0; r < kRows; r++) {
            const int gi = r * kThreads + tid;
            const int t = r * 8 + warp;
            const uint32_t rowpre = sm.row_tot[t];  // warp-uniform
            const uint32_t c2 = (cls >> (2 * r)) & 3u;
            const uint4 v = *reinterpret_cast<const uint4 *>(tile_in + gi * 16);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
            if (c2 != 2u) {
                // one contiguous run of len bytes, shifted left by the EPBs removed from its NAL so far
                const uint64_t c_row = (rowpre >> 31) ? (uint64_t)(rowpre & 0x7FFFu) : carry_in + (rowpre & 0x7FFFu);
                const uint32_t removed = __shfl_sync(0xFFFFFFFFu, incl[r], 31) & 0x7FFFu;
                const uint64_t o = base + 512u * (uint32_t)t - c_row;
                const uint8_t *prev_tail = nullptr;
                if (t > 0 && sm.row_class[t - 1] != 2) {
                    const uint32_t prev_removed = (rowpre - sm.row_tot[t - 1]) & 0x7FFFu;
                    prev_tail = tile_in + 512 * t - prev_removed;
                }
                const bool next_joins = t < kRows * 8 - 1 && sm.row_class[t + 1] != 2;
                store_row(a.out, o, 512u - removed, w, lane, prev_tail, next_joins);
            } else {
                uint32_t ex = __shfl_up_sync(0xFFFFFFFFu, incl[r], 1);
                if (lane == 0) ex = 0;
                const uint32_t pre = seg_combine(rowpre, ex);
                const uint64_t c = (pre >> 31) ? (uint64_t)(pre & 0x7FFFu) : carry_in + (pre & 0x7FFFu);
                const uint64_t gpos = base + (uint64_t)gi * 16;
                uint64_t k = slot0 + ((pre >> 16) & 0x1FFFu);  // record slot of the first NAL that starts in this granule
                store_granule_bytes(a.out, gpos, w, ks[r] & 0xFFFFu, ee[r], ks[r] >> 16, c, [&](int j, uint64_t c_end) {
                    if (k < a.nal_cap) {
                        const uint64_t st = gpos + (uint64_t)j + 1;  // the new NAL's first byte
                        a.rec_start[k] = st;
                        a.rec_epb[k] = c_end;  // EPBs removed from the NAL that ends with this start code
                        uint32_t h = 0;  // its first 4 bytes, from the (L2-resident) input
#pragma unroll
                        for (int q = 0; q < 4; q++)
                            h |= (uint32_t)(st + q < a.n ? a.in[st + q] : (uint8_t)0xFF) << (8 * q);
                        a.rec_hdr[k] = h;
                    }
                    k++;
                });
            }
        }
        // the loop-top __syncthreads orders these shared-memory reads before the next tile's writes
    }
}

// ------------------------------------------------------------------------------------------------ finalize
__device__ __forceinline__ void decode_nal_header(uint32_t hdr4, h264b_nal &o, h264b_nal_ext *ext) {
    const uint32_t b0 = hdr4 & 0xFF, b1 = (hdr4 >> 8) & 0xFF, b2 = (hdr4 >> 16) & 0xFF, b3 = hdr4 >> 24;
    o.forbidden_zero_bit = (uint8_t)(b0 >> 7);
    o.ref_idc = (uint8_t)((b0 >> 5) & 3);
    o.type = (uint8_t)(b0 & 31);
    o.header_bytes = (uint8_t)nal_header_bytes(b0, b1);
    if (!ext) return;
    h264b_nal_ext e;
    memset(&e, 0, sizeof(e));
    const uint32_t t = b0 & 31;
    if (t == 14 || t == 20 || t == 21) {
        const uint32_t bits = (b1 << 16) | (b2 << 8) | b3;  // 24 extension bits, MSB first
        const uint32_t flag = bits >> 23;
        if (t != 21) e.svc_extension_flag = (uint8_t)flag; else e.avc_3d_extension_flag = (uint8_t)flag;
        if (t != 21 && flag) {  // nalUnit.go:39-51
            e.idr_flag = (bits >> 22) & 1;
            e.priority_id = (bits >> 16) & 63;
            e.no_inter_layer_pred_flag = (bits >> 15) & 1;
            e.dependency_id = (bits >> 12) & 7;
            e.quality_id = (bits >> 8) & 15;
            e.temporal_id = (bits >> 5) & 7;
            e.use_ref_base_pic_flag = (bits >> 4) & 1;
            e.discardable_flag = (bits >> 3) & 1;
            e.output_flag = (bits >> 2) & 1;
            e.reserved_three_2bits = bits & 3;
        } else if (t == 21 && flag) {  // nalUnit.go:53-61 (16 bits)
            e.view_idx = (bits >> 15) & 255;
            e.depth_flag = (bits >> 14) & 1;
            e.non_idr_flag = (bits >> 13) & 1;
            e.temporal_id = (bits >> 10) & 7;
            e.anchor_pic_flag = (bits >> 9) & 1;
            e.inter_view_flag = (bits >> 8) & 1;
        } else {  // nalUnit.go:62-71
            e.non_idr_flag = (bits >> 22) & 1;
            e.priority_id = (bits >> 16) & 63;
            e.view_id = (bits >> 6) & 1023;
            e.temporal_id = (bits >> 3) & 7;
            e.anchor_pic_flag = (bits >> 2) & 1;
            e.inter_view_flag = (bits >> 1) & 1;
            e.reserved_one_bit = bits & 1;
        }
    }
    *ext = e;
}

// Post-pass 1: ordinal of every tile's first start code = exclusive scan of the per-tile start-code counts (one CTA;
// the array has one entry per 16 KiB of stream; warps read it coalesced, 1024 entries per step).
__global__ void __launch_bounds__(1024) nal_order_kernel(const uint32_t *tile_tot, uint32_t *tile_ord, uint32_t n_tiles) {
    __shared__ uint32_t warp_sum[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t running = 0;  // identical in every thread
    for (uint32_t base = 0; base < n_tiles; base += 1024) {
        const uint32_t i = base + (uint32_t)tid;
        const uint32_t c = i < n_tiles ? ((tile_tot[i] >> 16) & 0x1FFFu) : 0u;
        uint32_t x = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) warp_sum[warp] = x;
        __syncthreads();
        uint32_t w = warp_sum[lane], wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t y = __shfl_up_sync(0xFFFFFFFFu, wi, d);
            if (lane >= d) wi += y;
        }
        const uint32_t wbase = __shfl_sync(0xFFFFFFFFu, wi - w, warp);
        const uint32_t tot = __shfl_sync(0xFFFFFFFFu, wi, 31);
        if (i < n_tiles) tile_ord[i] = running + wbase + x - c;
        running += tot;
        __syncthreads();
    }
}

// Post-pass 2: move every tile's records from its reserved slots to their ordinals (stream order).
__global__ void __launch_bounds__(256) nal_permute_kernel(ScanArgs a) {
    const uint64_t cap = a.nal_cap;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < a.n_tiles; t += gridDim.x * blockDim.x) {
        const uint32_t c = (a.tile_tot[t] >> 16) & 0x1FFFu;
        if (!c) continue;
        const uint64_t slot = a.tile_slot[t], ord = a.tile_ord[t];
        for (uint32_t i = 0; i < c; i++) {
            if (slot + i < cap && ord + i < cap) {
                a.nal_start[ord + i] = a.rec_start[slot + i];
                a.nal_epb[ord + i] = a.rec_epb[slot + i];
                a.nal_hdr[ord + i] = a.rec_hdr[slot + i];
            }
        }
    }
}

// Post-pass 3: the h264b_nal records; NALs whose later tile pieces have to slide left are queued for post-pass 4.
__global__ void __launch_bounds__(256) scan_finalize_kernel(ScanArgs a, h264b_nal *nals, h264b_nal_ext *ext,
                                                             h264b_scan_summary *summary) {
    const uint64_t K = a.hdr->total_sc;
    const uint64_t n_nals = K ? K - 1 : 0;
    // more start codes than record slots: the index is incomplete (status H264B_E_CAPACITY, the caller retries
    // with n_start_codes + 1 slots), so no record is produced at all
    const uint64_t lim = K > a.nal_cap ? 0 : n_nals;
    unsigned long long epb = 0, rbsp = 0;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < lim;
         k += (uint64_t)gridDim.x * blockDim.x) {
        h264b_nal o;
        o.start = a.nal_start[k];
        const uint64_t next = a.nal_start[k + 1];
        o.num_bytes = (uint32_t)(next - o.start);
        decode_nal_header(a.nal_hdr[k], o, ext ? &ext[k] : nullptr);
        bool fix = false;
        const uint64_t removed = nal_pieces(o.start, next, o.header_bytes, a.nal_epb[k + 1], a.tile_tot, (uint64_t)kTile,
                                            [&](uint64_t, uint64_t, uint64_t) { fix = true; });
        if (fix) a.fix_list[atomicAdd(&a.hdr->n_fix, 1u)] = (uint32_t)k;
        // body = NumBytes - HeaderBytes - 2 bytes (a NAL shorter than that has no body); its RBSP sits at the body's
        // own position in the output buffer
        const int64_t body = (int64_t)o.num_bytes - (int64_t)o.header_bytes - 2;
        o.rbsp_off = o.start + o.header_bytes;
        o.rbsp_len = (uint32_t)((body > 0 ? body : 0) - (int64_t)removed);
        o.flags = (removed ? H264B_F_HAS_EPB : 0u) | (o.num_bytes < 8 ? H264B_F_SHORT_NAL : 0u);
        epb += removed;
        rbsp += o.rbsp_len;
        nals[k] = o;
    }
    if (epb) atomicAdd(&a.hdr->n_epb, epb);
    if (rbsp) atomicAdd(&a.hdr->total_kept, rbsp);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        summary->n_start_codes = K;
        summary->n_nals = n_nals;
        summary->first_start = a.hdr->first_start;
        summary->status = (K > a.nal_cap) ? H264B_E_CAPACITY : H264B_OK;
        summary->reserved = 0;
    }
}

// Post-pass 4: slide the later pieces of the queued NALs left (one CTA per NAL, pieces in stream order, 4 KiB at a
// time: everything is read into registers before anything is written, so the overlapping move is safe).
__global__ void __launch_bounds__(256) nal_fixup_kernel(ScanArgs a) {
    const uint32_t n_fix = a.hdr->n_fix;
    for (uint32_t f = blockIdx.x; f < n_fix; f += gridDim.x) {
        const uint64_t k = a.fix_list[f];
        const uint64_t st = a.nal_start[k], next = a.nal_start[k + 1];
        h264b_nal o;
        decode_nal_header(a.nal_hdr[k], o, nullptr);
        nal_pieces(st, next, o.header_bytes, a.nal_epb[k + 1], a.tile_tot, (uint64_t)kTile,
                   [&](uint64_t ps, uint64_t len, uint64_t G) {
                       for (uint64_t off = 0; off < len; off += 256 * 16) {
                           const uint64_t p = ps + off + (uint64_t)threadIdx.x * 16;
                           uint8_t v[16];
                           const uint64_t end = ps + len;
#pragma unroll
                           for (int j = 0; j < 16; j++) v[j] = p + j < end ? a.out[p + j] : (uint8_t)0;
                           __syncthreads();
#pragma unroll
                           for (int j = 0; j < 16; j++)
                               if (p + j < end) a.out[p + j - G] = v[j];
                           __syncthreads();
                       }
                   });
    }
}

__global__ void scan_summary_epb_kernel(const ScanScratchHeader *hdr, h264b_scan_summary *summary) {
    summary->n_epb = hdr->n_epb;
    summary->rbsp_bytes = hdr->total_kept;  // RBSP bytes of all emitted NAL units
}

// ------------------------------------------------------------------------------------------------ frames (NewNalUnit)
// One CTA per frame; RBSP of frame i is written at rbsp + off[i] (never longer than the frame).
__global__ void __launch_bounds__(256) nal_frames_kernel(const uint8_t *in, uint64_t total, const uint64_t *off,
                                                         const uint32_t *len, uint32_t n_frames, h264b_nal *nals,
                                                         h264b_nal_ext *ext, uint8_t *rbsp) {
    __shared__ uint32_t warp_tot[8];
    __shared__ uint32_t running;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t f = blockIdx.x; f < n_frames; f += gridDim.x) {
        const int64_t a0 = (int64_t)off[f], N = (int64_t)len[f];
        auto get = [&](int64_t p) -> uint32_t {
            return (p >= a0 && p < a0 + N && (uint64_t)p < total) ? (uint32_t)in[p] : 0xFFu;
        };
        const uint32_t b0 = get(a0), b1 = get(a0 + 1);
        const uint32_t H = nal_header_bytes(b0, b1);
        if (tid == 0) running = 0;
        __syncthreads();
        int any_epb = 0;
        for (int64_t chunk = 0; chunk < N; chunk += 256 * 16) {
            const int64_t p0 = a0 + chunk + (int64_t)tid * 16;
            uint32_t k16 = 0;
            for (int j = 0; j < 16; j++) {
                if (p0 + j >= a0 + N) break;
                if (keep_byte_frame(get, a0, N, H, p0 + j)) k16 |= 1u << j;
                if (is_epb_frame(get, a0, N, H, p0 + j)) any_epb = 1;
            }
            uint32_t x = __popc(k16);
            const uint32_t own = x;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
                if (lane >= d) x += y;
            }
            if (lane == 31) warp_tot[warp] = x;
            __syncthreads();
            uint32_t wbase = 0, tot = 0;
            for (int w2 = 0; w2 < 8; w2++) {
                if (w2 < warp) wbase += warp_tot[w2];
                tot += warp_tot[w2];
            }
            uint32_t o = running + wbase + x - own;
            for (int j = 0; j < 16; j++)
                if (k16 & (1u << j)) rbsp[a0 + o++] = (uint8_t)get(p0 + j);
            __syncthreads();
            if (tid == 0) running += tot;
            __syncthreads();
        }
        any_epb = __syncthreads_or(any_epb);
        if (tid == 0) {
            h264b_nal o;
            o.start = (uint64_t)a0;
            o.rbsp_off = (uint64_t)a0;
            o.num_bytes = (uint32_t)N;
            o.rbsp_len = running;
            const uint32_t hdr4 = b0 | (b1 << 8) | (get(a0 + 2) << 16) | (get(a0 + 3) << 24);
            decode_nal_header(hdr4, o, ext ? &ext[f] : nullptr);
            // a frame shorter than its header makes the reference panic (bit_reader.go:298): flag it
            o.flags = ((int64_t)o.header_bytes > N ? H264B_F_OVERRUN : 0u) | (N < 8 ? H264B_F_SHORT_NAL : 0u) |
                      (any_epb ? H264B_F_HAS_EPB : 0u);
            nals[f] = o;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ slice selection
// Ordered list of the NAL units of type 1 / 5 (server.go:147-162 dispatches exactly those to the slice parser).
// Single CTA, ballot-based stable compaction; the list is small (one entry per slice).
__global__ void __launch_bounds__(1024) slice_select_kernel(const h264b_nal *nals, const h264b_scan_summary *summary,
                                                            uint32_t nal_cap, uint32_t data_off, uint32_t max_slices,
                                                            uint64_t *s_off, uint32_t *s_len, uint32_t *s_nal,
                                                            uint32_t *n_out) {
    __shared__ uint32_t warp_cnt[32];
    __shared__ uint32_t base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint64_t n = summary->n_nals;
    if (n > nal_cap) n = nal_cap;
    if (tid == 0) base = 0;
    __syncthreads();
    for (uint64_t k0 = 0; k0 < n; k0 += 1024) {
        const uint64_t k = k0 + tid;
        bool is_slice = false;
        h264b_nal u;
        if (k < n) {
            u = nals[k];
            is_slice = (u.type == 1 || u.type == 5);
        }
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, is_slice);
        if (lane == 0) warp_cnt[warp] = __popc(m);
        __syncthreads();
        uint32_t wbase = 0, tot = 0;
        for (int w2 = 0; w2 < 32; w2++) {
            if (w2 < warp) wbase += warp_cnt[w2];
            tot += warp_cnt[w2];
        }
        if (is_slice) {
            const uint32_t idx = base + wbase + __popc(m & ((1u << lane) - 1u));
            if (idx < max_slices) {
                const uint32_t skip = data_off < u.rbsp_len ? data_off : u.rbsp_len;
                s_off[idx] = u.rbsp_off + skip;
                s_len[idx] = u.rbsp_len - skip;
                s_nal[idx] = (uint32_t)k;
            }
        }
        __syncthreads();
        if (tid == 0) base += tot;
        __syncthreads();
    }
    if (tid == 0) *n_out = base < max_slices ? base : max_slices;
}

// ------------------------------------------------------------------------------------------------ launchers
struct ScratchOffsets {
    uint64_t tile_tot, tile_slot, tile_ord, fix_list, rec_start, rec_epb, rec_hdr, nal_start, nal_epb, nal_hdr, total;
};
static ScratchOffsets scratch_layout(uint64_t n, uint32_t nal_cap) {
    const uint64_t n_tiles = (n + kTile - 1) / kTile;
    ScratchOffsets o;
    uint64_t p = sizeof(ScanScratchHeader);
    auto take = [&](uint64_t bytes) {
        const uint64_t at = p;
        p = (p + bytes + 15) & ~15ull;
        return at;
    };
    o.tile_tot = take(n_tiles * 4);
    o.tile_slot = take(n_tiles * 4);
    o.tile_ord = take(n_tiles * 4);
    o.fix_list = take((uint64_t)nal_cap * 4);
    o.rec_start = take((uint64_t)nal_cap * 8);
    o.rec_epb = take((uint64_t)nal_cap * 8);
    o.rec_hdr = take((uint64_t)nal_cap * 4);
    o.nal_start = take((uint64_t)nal_cap * 8);
    o.nal_epb = take((uint64_t)nal_cap * 8);
    o.nal_hdr = take((uint64_t)nal_cap * 4);
    o.total = (p + 255) & ~255ull;
    return o;
}

int launch_annexb_scan(h264b_ctx *ctx, const uint8_t *d_stream, uint64_t n, uint8_t *d_rbsp, h264b_nal *d_nals,
                       h264b_nal_ext *d_ext, uint32_t nal_cap, h264b_scan_summary *d_summary, uint32_t flags) {
    (void)flags;
    if (((uintptr_t)d_stream & 15) || ((uintptr_t)d_rbsp & 15))
        return set_error(ctx, H264B_E_INVALID, "annexb_scan: d_stream and d_rbsp must be 16-byte aligned");
    if (n >= (1ull << 45)) return set_error(ctx, H264B_E_INVALID, "annexb_scan: stream too long");
    const ScratchOffsets so = scratch_layout(n, nal_cap);
    if (so.total > ctx->scan_scratch_bytes) {
        if (ctx->scan_scratch) {
            H264B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            cudaFree(ctx->scan_scratch);
        }
        ctx->scan_scratch = nullptr;
        ctx->scan_scratch_bytes = 0;
        H264B_CUDA(ctx, cudaMalloc(&ctx->scan_scratch, so.total));
        ctx->scan_scratch_bytes = so.total;
    }
    uint8_t *s = (uint8_t *)ctx->scan_scratch;
    const uint64_t n_tiles = (n + kTile - 1) / kTile;
    ScanArgs a;
    a.in = d_stream;
    a.n = n;
    a.out = d_rbsp;
    a.hdr = (ScanScratchHeader *)s;
    a.tile_tot = (uint32_t *)(s + so.tile_tot);
    a.tile_slot = (uint32_t *)(s + so.tile_slot);
    a.tile_ord = (uint32_t *)(s + so.tile_ord);
    a.fix_list = (uint32_t *)(s + so.fix_list);
    a.rec_start = (unsigned long long *)(s + so.rec_start);
    a.rec_epb = (unsigned long long *)(s + so.rec_epb);
    a.rec_hdr = (uint32_t *)(s + so.rec_hdr);
    a.nal_start = (unsigned long long *)(s + so.nal_start);
    a.nal_epb = (unsigned long long *)(s + so.nal_epb);
    a.nal_hdr = (uint32_t *)(s + so.nal_hdr);
    a.nal_cap = nal_cap;
    a.n_tiles = (uint32_t)n_tiles;

    scan_init_kernel<<<1, 1, 0, ctx->stream>>>(a.hdr, n);
    H264B_LAUNCH_CHECK(ctx, "scan_init_kernel");
    if (n_tiles) {
        const uint64_t chunks = (n + 4095) / 4096;
        const int fs_blocks = (int)(chunks < (uint64_t)ctx->sm_count * 4 ? chunks : (uint64_t)ctx->sm_count * 4);
        first_start_kernel<<<fs_blocks, 256, 0, ctx->stream>>>(d_stream, n, a.hdr);
        H264B_LAUNCH_CHECK(ctx, "first_start_kernel");

        static bool attr_set = false;
        const size_t smem = sizeof(ScanSmem);
        if (!attr_set) {
            H264B_CUDA(ctx, cudaFuncSetAttribute(annexb_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)smem));
            attr_set = true;
        }
        int occ = 0;
        H264B_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, annexb_scan_kernel, kThreads, smem));
        if (occ < 1) occ = 1;
        uint64_t grid = (uint64_t)ctx->sm_count * occ;
        if (grid > n_tiles) grid = n_tiles;
        annexb_scan_kernel<<<(int)grid, kThreads, smem, ctx->stream>>>(a);
        H264B_LAUNCH_CHECK(ctx, "annexb_scan_kernel");
        nal_order_kernel<<<1, 1024, 0, ctx->stream>>>(a.tile_tot, a.tile_ord, a.n_tiles);
        H264B_LAUNCH_CHECK(ctx, "nal_order_kernel");
        uint64_t pb = (n_tiles + 255) / 256;
        if (pb > (uint64_t)ctx->sm_count * 8) pb = (uint64_t)ctx->sm_count * 8;
        nal_permute_kernel<<<(int)pb, 256, 0, ctx->stream>>>(a);
        H264B_LAUNCH_CHECK(ctx, "nal_permute_kernel");
    }
    int fin_blocks = ctx->sm_count * 2;
    scan_finalize_kernel<<<fin_blocks, 256, 0, ctx->stream>>>(a, d_nals, d_ext, d_summary);
    H264B_LAUNCH_CHECK(ctx, "scan_finalize_kernel");
    if (n_tiles > 1) {
        nal_fixup_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(a);
        H264B_LAUNCH_CHECK(ctx, "nal_fixup_kernel");
    }
    scan_summary_epb_kernel<<<1, 1, 0, ctx->stream>>>(a.hdr, d_summary);
    H264B_LAUNCH_CHECK(ctx, "scan_summary_epb_kernel");
    return H264B_OK;
}

int launch_nal_frames(h264b_ctx *ctx, const uint8_t *d_frames, uint64_t total, const uint64_t *d_off,
                      const uint32_t *d_len, uint32_t n_frames, h264b_nal *d_nals, h264b_nal_ext *d_ext,
                      uint8_t *d_rbsp) {
    if (!n_frames) return H264B_OK;
    int blocks = (int)(n_frames < (uint32_t)ctx->sm_count * 8 ? n_frames : (uint32_t)ctx->sm_count * 8);
    nal_frames_kernel<<<blocks, 256, 0, ctx->stream>>>(d_frames, total, d_off, d_len, n_frames, d_nals, d_ext, d_rbsp);
    H264B_LAUNCH_CHECK(ctx, "nal_frames_kernel");
    return H264B_OK;
}

int launch_slice_select(h264b_ctx *ctx, const h264b_nal *d_nals, const h264b_scan_summary *d_summary,
                        uint32_t nal_cap, uint32_t slice_data_offset, uint32_t max_slices, uint64_t *d_off,
                        uint32_t *d_len, uint32_t *d_slice_nal, uint32_t *d_n_slices) {
    slice_select_kernel<<<1, 1024, 0, ctx->stream>>>(d_nals, d_summary, nal_cap, slice_data_offset, max_slices, d_off,
                                                     d_len, d_slice_nal, d_n_slices);
    H264B_LAUNCH_CHECK(ctx, "slice_select_kernel");
    return H264B_OK;
}

}  // namespace h264b

extern "C" uint64_t h264b_annexb_scratch_bytes(uint64_t n) { return h264b::scratch_layout(n, 0).total; }
