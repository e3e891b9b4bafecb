// annexb_scan.cu -- K1/K2: single-pass Annex-B split + RBSP emulation-prevention strip for sm_100a.
//
// Replaces the byte-at-a-time loops of readNalUnit (h264/server.go:64-111) and NewNalUnit (h264/nalUnit.go:75-131)
// of the reference with one streaming pass: every input byte is read from HBM once and every kept byte written once
// (algorithmic traffic N_in + N_rbsp + 20 B per NAL).
//
// The RBSP of a NAL is written at the position of the NAL's own body (h264b_nal.rbsp_off = start + header_bytes), so
// a span of the stream that holds neither a start code nor an emulation-prevention byte -- all but a few KiB per MB
// of entropy-coded data -- is an aligned copy.  The kernel is built around that:
//
//   pieces   the stream is cut into spans of chunks ("pieces", 128 KiB by default); a warp takes a piece by atomic
//            ticket and walks its 2 KiB chunks front to back.  Warps never talk to each other: no block barrier, no
//            look-back.  What a NAL lost in an EARLIER piece is made up for by the post-pass (nal_pieces).
//   ring     every warp owns a ring of kStages shared-memory slots; lane 0 keeps it full with TMA 1-D bulk loads
//            (cp.async.bulk + mbarrier complete_tx) of chunk + 16-byte halos, kStages-1 chunks ahead of the consumer
//   detect   LDS.128 per lane and granule, word-parallel search for two adjacent zero bytes (zero_pair_bits): every
//            00 00 03 and 00 00 00 01 needs one
//   clean    no pair anywhere and nothing removed from the open NAL so far: lane 0 hands the slot to the TMA again,
//            one 2 KiB bulk store straight from shared memory to out + pos (no registers, no STG)
//   dirty    otherwise the chunk takes the general path: exact masks (granule_masks), start-code bitmap, bit-domain
//            fix-up near start codes (keep_mask_near_sc), per-row segmented scan of (EPBs | start codes), in-place
//            compaction of rows that only lose EPBs, shuffle-aligned 16-byte row stores, byte stores and NAL records
//            (start, EPB count, first 4 bytes, rank in piece) for rows with boundaries
// Post-passes (tiny): exclusive scan of the per-piece start-code counts, permutation of the records into stream
// order, h264b_nal records (lengths are differences of neighbours), and the slide of NAL parts described above.
#include "annexb_local.cuh"
#include "common.cuh"

namespace h264b {

#ifndef H264B_SCAN_ROWS
#define H264B_SCAN_ROWS 4
#endif
#ifndef H264B_SCAN_STAGES
#define H264B_SCAN_STAGES 4
#endif
#ifndef H264B_SCAN_WARPS
#define H264B_SCAN_WARPS 4
#endif
constexpr int kRows = H264B_SCAN_ROWS;             // 512-byte rows per chunk (one granule per lane and row), <= 16
constexpr int kChunkGran = 32 * kRows;             // 128 granules
constexpr int kChunk = kChunkGran * 16;            // 2048 bytes
constexpr int kHalo = 16;
constexpr int kSlotBytes = kHalo + kChunk + kHalo; // 2080
constexpr int kStages = H264B_SCAN_STAGES;         // ring slots per warp
constexpr int kWarps = H264B_SCAN_WARPS;           // warps per CTA (independent of each other)
constexpr uint32_t kMaxSpanBytes = 128 * 1024;     // piece size for large streams

struct ScanScratchHeader {   // device scratch, initialised by scan_init_kernel
    unsigned int ticket;           // next piece
    unsigned int n_fix;            // entries of fix_list
    unsigned long long first_start;
    unsigned long long total_sc;   // start codes found = slots handed out of the NAL record buffer
    unsigned long long total_kept;
    unsigned long long n_epb;
    unsigned long long reserved[3];
};  // 64 bytes

struct ScanArgs {
    const uint8_t *in;
    uint64_t n;
    uint8_t *out;
    ScanScratchHeader *hdr;
    uint32_t *piece_epb;       // per piece: EPBs after its last NAL start (or in the whole piece when it has none)
    uint32_t *piece_nsc;       // per piece: start codes
    uint32_t *piece_ord;       // per piece: ordinal of its first start code (exclusive scan of piece_nsc)
    uint32_t *fix_list;        // NAL ordinals whose later parts must slide left (written by scan_finalize_kernel)
    // per start code, in slot order (chunks reserve slots with one atomicAdd): written by the main pass
    unsigned long long *rec_start;
    uint32_t *rec_epb;
    uint32_t *rec_hdr;
    uint32_t *rec_rank;        // rank of the start code inside its piece
    // the same in stream order (written by nal_permute_kernel, read by scan_finalize_kernel)
    unsigned long long *nal_start;
    uint32_t *nal_epb;         // [k]: EPBs removed (within the piece of start code k) from the NAL that ends there
    uint32_t *nal_hdr;
    uint32_t nal_cap;
    uint32_t n_chunks;
    uint32_t n_pieces;
    uint32_t span_chunks;      // chunks per piece
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------------------------------------ init
__global__ void scan_init_kernel(ScanScratchHeader *hdr, uint64_t n) {
    hdr->ticket = 0;
    hdr->n_fix = 0;
    hdr->first_start = n;
    hdr->total_sc = 0;
    hdr->total_kept = 0;
    hdr->n_epb = 0;
}

// ------------------------------------------------------------------------------------------------ main pass
struct __align__(16) WarpRing {
    uint8_t slot[kStages][kSlotBytes];    // each: [0,16) low halo, [16,16+kChunk) chunk, high halo
    ulonglong2 meta[kStages];             // .x = stream offset of the chunk in the slot (~0: no more work), .y = its piece
    unsigned long long mbar[kStages];     // "bytes have landed"
    uint16_t scbits[kChunkGran + 2];      // start-code-end bits per granule, [0] = halo granule before the chunk
    uint16_t pad[(8 - (kChunkGran + 2 + 4 * kStages) % 8) % 8];
};
static_assert(sizeof(WarpRing) % 16 == 0 && kSlotBytes % 16 == 0, "ring slots must stay 16-byte aligned");

// TMA bulk store shared -> global (16-byte aligned on both sides, size a multiple of 16)
__device__ __forceinline__ void bulk_store(uint8_t *dst_global, const uint8_t *src_shared, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_global),
                 "r"(smem_u32(src_shared)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    }
}

__device__ __forceinline__ uint32_t warp_seg_scan(uint32_t x, int lane) {  // inclusive segmented scan over the warp
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
        if (lane >= d) x = seg_combine(y, x);
    }
    return x;
}

// The general path is rare for real streams; its pieces are kept out of line so that the hot loop stays small.
__device__ __noinline__ uint32_t keep_near_sc(const uint8_t *tile_in, uint64_t base, uint64_t gpos, uint32_t e16,
                                              uint32_t sc_prev, uint32_t sc_own, uint32_t sc_next, uint32_t *epb_eff) {
    auto get = [&](int64_t p) -> uint32_t { return tile_in[p - (int64_t)base]; };
    return keep_mask_near_sc(get, (int64_t)gpos, e16, sc_prev, sc_own, sc_next, epb_eff);
}

__device__ __noinline__ void compact_row_in_place(uint8_t *row, const uint8_t *src, uint32_t k16, uint32_t loff) {
    const uint4 v = *reinterpret_cast<const uint4 *>(src);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    __syncwarp();  // every lane has its bytes in registers before anyone overwrites the span
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

// one lane's granule of a row with NAL boundaries: kept bytes one by one, one record per start-code end
//   c     EPBs removed so far (in this piece) from the NAL open at the granule's first byte
//   k     record slot of the first start code of the granule;  rank: its rank inside the piece
__device__ __noinline__ void store_boundary_granule(const ScanArgs &a, uint64_t gpos, const uint8_t *src, uint32_t k16,
                                                    uint32_t ee, uint32_t sc, uint64_t c, uint64_t k, uint32_t rank,
                                                    bool first_of_chunk) {
    const uint4 v = *reinterpret_cast<const uint4 *>(src);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    store_granule_bytes(a.out, gpos, w, k16, ee, sc, c, [&](int j, uint64_t c_end) {
        const uint64_t st = gpos + (uint64_t)j + 1;  // the new NAL's first byte
        if (first_of_chunk) {
            atomicMin(&a.hdr->first_start, (unsigned long long)st);
            first_of_chunk = false;
        }
        if (k < a.nal_cap) {
            a.rec_start[k] = st;
            a.rec_epb[k] = (uint32_t)c_end;  // EPBs removed (in this piece) from the NAL that ends with this start code
            uint32_t h = 0;                  // its first 4 bytes, from the (L2-resident) input
#pragma unroll
            for (int q = 0; q < 4; q++) h |= (uint32_t)(st + q < a.n ? a.in[st + q] : (uint8_t)0xFF) << (8 * q);
            a.rec_hdr[k] = h;
            a.rec_rank[k] = rank;
        }
        k++;
        rank++;
    });
}

__device__ __noinline__ void store_row(uint8_t *out, uint64_t o, uint32_t len, const uint8_t *src, int lane,
                                       const uint8_t *prev_tail, bool next_joins) {
    const uint4 v = *reinterpret_cast<const uint4 *>(src);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
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

struct ChunkResult {
    uint32_t carry_epb;  // EPBs removed from the open NAL since its start / the start of the piece
    uint32_t piece_nsc;  // start codes of the piece so far
    uint32_t clean;      // nothing to remove after all and nothing shifted: the caller bulk-stores the chunk
};

// General path of one chunk (whole warp).  tile_in[i] = s[pos + i] for i in [-16, kChunk + 16), bytes outside the
// stream read as 0xFF.
__device__ __noinline__ ChunkResult general_chunk(const ScanArgs &a, uint16_t *scbits, uint8_t *buf, uint64_t pos,
                                                  uint32_t carry_epb, uint32_t piece_nsc, int lane) {
    uint8_t *tile_in = buf + kHalo;
    if (pos == 0 || pos + kChunk + kHalo > a.n) {  // bytes outside the stream read as 0xFF (they match no predicate)
        const uint64_t n16 = (a.n + 15) & ~15ull;
        if (pos == 0 && lane < 4) reinterpret_cast<uint32_t *>(buf)[lane] = 0xFFFFFFFFu;
        uint64_t hi = pos + kChunk + kHalo;
        if (hi > n16) hi = n16;
        const uint32_t loaded_end = (uint32_t)(hi + kHalo - pos);  // offset in the slot
        for (uint32_t o = loaded_end + lane * 4; o < (uint32_t)kSlotBytes; o += 128)
            *reinterpret_cast<uint32_t *>(buf + o) = 0xFFFFFFFFu;
        if (hi > a.n) {  // the last 16-byte granule holds bytes past n: blank them
            const uint32_t first_bad = (uint32_t)(a.n + kHalo - pos);
            if (lane < 16 && first_bad + lane < loaded_end) buf[first_bad + lane] = 0xFF;
        }
        __syncwarp();
    }
    // ---------------------------------------------------------------- exact masks + start-code bitmap
    uint32_t em[kRows];  // raw EPB mask | start-code mask << 16
#pragma unroll
    for (int r = 0; r < kRows; r++) {
        const int gi = r * 32 + lane;
        const uint4 v = *reinterpret_cast<const uint4 *>(tile_in + gi * 16);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        const uint32_t prev = *reinterpret_cast<const uint32_t *>(tile_in + gi * 16 - 4);
        const GranuleMasks m = granule_masks(w, prev);
        em[r] = m.e | (m.sc << 16);
        scbits[gi + 1] = (uint16_t)m.sc;
    }
    if (lane < 2) {  // halo granules: only start codes ending in [pos-6, pos-1] and at pos+kChunk matter
        const int gi = lane ? kChunkGran : -1;
        const uint4 v = *reinterpret_cast<const uint4 *>(tile_in + gi * 16);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        const uint32_t prev = lane ? *reinterpret_cast<const uint32_t *>(tile_in + gi * 16 - 4) : 0xFFFFFFFFu;
        scbits[gi + 1] = (uint16_t)granule_masks(w, prev).sc;
    }
    __syncwarp();

    // ---------------------------------------------------------------- adjust + classify + per-row segmented scan
    const bool has_end = pos + kChunk > a.n;  // some granules reach past the stream
    uint32_t ks[kRows];    // keep mask | start-code mask << 16
    uint32_t ee[kRows];    // emulation-prevention bytes really removed
    uint32_t incl[kRows];  // inclusive segmented scan inside the row
    uint32_t rp[kRows + 1];  // exclusive prefix of each row inside the chunk (warp-uniform); [kRows] = chunk total
    uint32_t cls = 0;        // 2 bits per row, warp-uniform: 0 untouched, 1 only EPBs removed, 2 boundaries / stream end
    rp[0] = 0;
#pragma unroll
    for (int r = 0; r < kRows; r++) {
        const int gi = r * 32 + lane;
        const uint64_t gpos = pos + (uint64_t)gi * 16;
        uint32_t e16 = em[r] & 0xFFFFu;
        uint32_t k16 = ~e16 & 0xFFFFu;
        // start-code ends q in [g-6, g+16] change what this granule keeps
        const uint32_t near = ((uint32_t)scbits[gi] >> 10) | scbits[gi + 1] | (scbits[gi + 2] & 1u);
        if (near)  // header bytes, the 2-byte tail rule and the EPB guard, all in the bit domain
            k16 = keep_near_sc(tile_in, pos, gpos, e16, scbits[gi], scbits[gi + 1], scbits[gi + 2], &e16);
        uint32_t sc = em[r] >> 16;
        if (has_end) {
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
        if (__all_sync(0xFFFFFFFFu, k16 == 0xFFFFu)) {  // nothing removed
            x = 0;
        } else if (__all_sync(0xFFFFFFFFu, (k16 | e16) == 0xFFFFu && sc == 0)) {  // only EPBs removed
            cls |= 1u << (2 * r);
            x = warp_seg_scan(bits_popc(e16), lane);
        } else {
            cls |= 2u << (2 * r);
            x = warp_seg_scan(seg_element(e16, sc), lane);
        }
        incl[r] = x;
        rp[r + 1] = seg_combine(rp[r], __shfl_sync(0xFFFFFFFFu, x, 31));
    }
    ChunkResult res;
    res.carry_epb = carry_epb;
    res.piece_nsc = piece_nsc;
    res.clean = 0;
    if (cls == 0 && carry_epb == 0) {  // false alarm (00 00 xx with xx > 3): a verbatim copy after all
        res.clean = 1;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // (the fills above)
        return res;
    }
    const uint32_t total = rp[kRows];
    const uint32_t n_sc = (total >> 16) & 0x1FFFu;
    unsigned long long slot0 = 0;
    if (n_sc) {
        if (lane == 0) slot0 = atomicAdd(&a.hdr->total_sc, (unsigned long long)n_sc);
        slot0 = __shfl_sync(0xFFFFFFFFu, slot0, 0);
    }

    // ---------------------------------------------------------------- in-place compaction of EPB-only rows
    // inside their own 512-byte span of the slot: afterwards they are `len` contiguous bytes like untouched rows
#pragma unroll
    for (int r = 0; r < kRows; r++) {
        if (((cls >> (2 * r)) & 3u) != 1u) continue;  // warp-uniform
        const int gi = r * 32 + lane;
        const uint32_t loff = 16u * (uint32_t)lane - (incl[r] - bits_popc(ee[r]));  // kept bytes of the lanes before
        compact_row_in_place(tile_in + r * 512, tile_in + gi * 16, ks[r] & 0xFFFFu, loff);
    }
    __syncwarp();

    // ---------------------------------------------------------------- store rows + NAL records
#pragma unroll
    for (int r = 0; r < kRows; r++) {
        const int gi = r * 32 + lane;
        const uint32_t c2 = (cls >> (2 * r)) & 3u;
        if (c2 != 2u) {
            // one contiguous run of len bytes, shifted left by the EPBs removed from its NAL so far in this piece
            const uint64_t c_row = seg_apply(rp[r], carry_epb);
            const uint32_t removed = c2 ? ((rp[r + 1] - rp[r]) & 0x7FFFu) : 0u;
            const uint64_t o = pos + 512u * (uint32_t)r - c_row;
            const uint8_t *prev_tail = nullptr;
            if (r > 0 && ((cls >> (2 * r - 2)) & 3u) != 2u) {
                const uint32_t prev_removed = (rp[r] - rp[r - 1]) & 0x7FFFu;
                prev_tail = tile_in + 512 * r - prev_removed;
            }
            const bool next_joins = r < kRows - 1 && ((cls >> (2 * r + 2)) & 3u) != 2u;
            store_row(a.out, o, 512u - removed, tile_in + gi * 16, lane, prev_tail, next_joins);
        } else {
            uint32_t ex = __shfl_up_sync(0xFFFFFFFFu, incl[r], 1);
            if (lane == 0) ex = 0;
            const uint32_t pre = seg_combine(rp[r], ex);
            const uint64_t c = seg_apply(pre, carry_epb);
            const uint32_t before = (pre >> 16) & 0x1FFFu;  // start codes of the chunk before this granule
            store_boundary_granule(a, pos + (uint64_t)gi * 16, tile_in + gi * 16, ks[r] & 0xFFFFu, ee[r], ks[r] >> 16, c,
                                   slot0 + before, piece_nsc + before, before == 0);
        }
    }
    res.carry_epb = seg_apply(total, carry_epb);
    res.piece_nsc = piece_nsc + n_sc;
    // generic-proxy writes to the slot (fills, compaction) are ordered before the TMA writes that will reuse it
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    return res;
}

// Producer side of a warp's ring (lane 0 only): pieces by ticket, their chunks in order.
struct Producer {
    uint64_t pos;    // stream offset of the next chunk to load
    uint32_t left;   // chunks of the current piece still to load
    uint32_t piece;
    uint32_t done;   // the tickets have run out
    uint32_t next_static;  // (H264B_SCAN_STATIC: pieces dealt round-robin instead of by ticket)
};

__device__ __forceinline__ void tma_load(uint32_t dst, const uint8_t *src, uint32_t bytes, uint32_t bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// everything but an interior chunk of the current piece: next piece, first / last chunks of the stream, the end
__device__ __noinline__ void produce_slow(const ScanArgs &a, WarpRing &ring, int s, Producer *p) {
    if (!p->done && p->left == 0) {
#ifdef H264B_SCAN_STATIC
        p->piece = p->next_static;
        p->next_static += gridDim.x * kWarps;
#else
        p->piece = atomicAdd(&a.hdr->ticket, 1u);
#endif
        if (p->piece < a.n_pieces) {
            const uint32_t first = p->piece * a.span_chunks;
            p->left = first + a.span_chunks < a.n_chunks ? a.span_chunks : a.n_chunks - first;
            p->pos = (uint64_t)first * kChunk;
        } else {
            p->done = 1;
        }
    }
    const uint32_t bar = smem_u32(&ring.mbar[s]);
    if (p->done) {  // an arrival without bytes: the consumer sees "no more work"
        ring.meta[s] = make_ulonglong2(~0ull, 0ull);
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
        return;
    }
    const uint64_t pos = p->pos, n16 = (a.n + 15) & ~15ull;
    ring.meta[s] = make_ulonglong2(pos, (unsigned long long)p->piece);
    const uint64_t lo = pos ? pos - kHalo : 0;
    uint64_t hi = pos + kChunk + kHalo;
    if (hi > n16) hi = n16;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tma_load(smem_u32(ring.slot[s] + (lo + kHalo - pos)), a.in + lo, (uint32_t)(hi - lo), bar);
    p->pos = pos + kChunk;
    p->left--;
}

__global__ void __launch_bounds__(kWarps * 32) annexb_scan_kernel(ScanArgs a) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpRing &ring = reinterpret_cast<WarpRing *>(smem_raw)[warp];
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kStages; s++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&ring.mbar[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const uint64_t n16 = (a.n + 15) & ~15ull;

    Producer prod = {0, 0, 0, 0, blockIdx.x * kWarps + (uint32_t)warp};
    if (lane == 0) {
        Producer t = prod;
#pragma unroll 1
#ifdef H264B_SCAN_STG
        for (int s = 0; s < kStages; s++) produce_slow(a, ring, s, &t);
#else
        for (int s = 0; s < kStages - 1; s++) produce_slow(a, ring, s, &t);
#endif
        prod = t;
    }

    uint32_t cur_piece = 0xFFFFFFFFu;
    uint32_t carry_epb = 0, piece_nsc = 0;
    int s = 0;            // slot of the chunk being consumed
    uint32_t parity = 0;  // phase parity of its barrier
#pragma unroll 1
    for (;;) {
        mbar_wait(smem_u32(&ring.mbar[s]), parity);
        const ulonglong2 meta = ring.meta[s];
        const uint64_t pos = meta.x;
        if (pos == ~0ull) break;
        if ((uint32_t)meta.y != cur_piece) {
            if (cur_piece != 0xFFFFFFFFu && lane == 0) {
                a.piece_epb[cur_piece] = carry_epb;
                a.piece_nsc[cur_piece] = piece_nsc;
            }
            cur_piece = (uint32_t)meta.y;
            carry_epb = 0;
            piece_nsc = 0;
        }
        uint8_t *buf = ring.slot[s];
        uint8_t *tile_in = buf + kHalo;  // tile_in[i] = s[pos + i], valid for i in [-16, kChunk+16)

        // ---------------------------------------------------------------- detect: two adjacent zero bytes anywhere?
        uint32_t acc = 0xFFFFFFFFu;
#ifdef H264B_SCAN_STG
        uint4 vv[kRows];
#pragma unroll
        for (int r = 0; r < kRows; r++) vv[r] = *reinterpret_cast<const uint4 *>(tile_in + (r * 32 + lane) * 16);
#endif
#if !defined(H264B_EXP_NODETECT) && !defined(H264B_EXP_LOADONLY)
#pragma unroll
        for (int r = 0; r < kRows; r++) {
            const int gi = r * 32 + lane;
#ifdef H264B_SCAN_STG
            const uint4 v = vv[r];
#else
            const uint4 v = *reinterpret_cast<const uint4 *>(tile_in + gi * 16);
#endif
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
            acc = zero_pair_acc(acc, w, *reinterpret_cast<const uint32_t *>(tile_in + gi * 16 - 4));
        }
        {   // a start code ending in the last bytes before the chunk still reaches into it (same words for all lanes)
            const uint2 t8 = *reinterpret_cast<const uint2 *>(buf + 8);
            acc = zero_pair_acc_tail8(acc, t8.x, t8.y);
        }
#endif
        // chunks that touch the ends of the stream always take the general path (it blanks the bytes outside)
        const bool edge = pos == 0 || pos + kChunk + kHalo > a.n;
        bool clean = !(__any_sync(0xFFFFFFFFu, acc_has_pair(acc)) || carry_epb != 0 || edge);
#ifdef H264B_EXP_LOADONLY
        clean = false;
#else
        if (!clean) {
            const ChunkResult res = general_chunk(a, ring.scbits, buf, pos, carry_epb, piece_nsc, lane);
            carry_epb = res.carry_epb;
            piece_nsc = res.piece_nsc;
            clean = res.clean != 0;
        }
#endif
        __syncwarp();  // every lane is done reading (and writing) the slot
#ifdef H264B_SCAN_STG
        const int s_prev = s;  // the registers hold the chunk: its slot is free at once
#ifndef H264B_EXP_NOSTORE
        if (clean) {
#pragma unroll
            for (int r = 0; r < kRows; r++)
                *reinterpret_cast<uint4 *>(a.out + pos + (uint32_t)(r * 32 + lane) * 16u) = vv[r];
        }
#endif
#else
        const int s_prev = s ? s - 1 : kStages - 1;
#endif
        if (lane == 0) {
#ifndef H264B_SCAN_STG
#ifndef H264B_EXP_NOSTORE
            // the bulk of a real stream leaves through the TMA: one bulk store straight from the slot
            if (clean) bulk_store(a.out + pos, tile_in, kChunk);
#endif
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            // all but the newest store have finished reading shared memory: the slot of the previous chunk is free
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
#endif
            if (prod.left != 0 && prod.pos + kChunk + kHalo <= n16) {  // interior chunk of the current piece
                ring.meta[s_prev] = make_ulonglong2(prod.pos, (unsigned long long)prod.piece);
                tma_load(smem_u32(ring.slot[s_prev]), a.in + prod.pos - kHalo, kSlotBytes, smem_u32(&ring.mbar[s_prev]));
                prod.pos += kChunk;
                prod.left--;
            } else {
                Producer t = prod;
                produce_slow(a, ring, s_prev, &t);
                prod = t;
            }
        }
        if (++s == kStages) {
            s = 0;
            parity ^= 1u;
        }
    }
    if (lane == 0) {
        if (cur_piece != 0xFFFFFFFFu) {
            a.piece_epb[cur_piece] = carry_epb;
            a.piece_nsc[cur_piece] = piece_nsc;
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

// ------------------------------------------------------------------------------------------------ main pass, LDG form
// The same pass with the chunk held in registers: coalesced LDG.128 (one granule per lane and row, the next chunk of
// the piece already in flight while this one is examined), the previous word by shuffle, STG.128 for clean chunks.
// Only chunks that take the general path are staged in shared memory.
struct __align__(16) WarpStage {
    uint8_t buf[kSlotBytes];
    uint16_t scbits[kChunkGran + 2];
    uint16_t pad[(8 - (kChunkGran + 2) % 8) % 8];
};

struct ChunkRegs {
    uint4 v[kRows];
    uint2 t8;  // the 8 bytes before the chunk
};

__device__ __forceinline__ void load_chunk(ChunkRegs &c, const uint8_t *in, uint64_t pos, uint64_t n16, int lane) {
#pragma unroll
    for (int r = 0; r < kRows; r++) {
        const uint64_t g = pos + (uint32_t)(r * 32 + lane) * 16u;
        c.v[r] = g < n16 ? __ldcs(reinterpret_cast<const uint4 *>(in + g)) : make_uint4(~0u, ~0u, ~0u, ~0u);
    }
    c.t8 = pos ? *reinterpret_cast<const uint2 *>(in + pos - 8) : make_uint2(~0u, ~0u);
}

__global__ void __launch_bounds__(kWarps * 32) annexb_scan_ldg_kernel(ScanArgs a) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpStage &st = reinterpret_cast<WarpStage *>(smem_raw)[warp];
    const uint64_t n16 = (a.n + 15) & ~15ull;
#ifdef H264B_SCAN_STATIC
    uint32_t next_static = blockIdx.x * kWarps + (uint32_t)warp;
#endif
#pragma unroll 1
    for (;;) {
#ifdef H264B_SCAN_STATIC
        const uint32_t piece = next_static;
        next_static += gridDim.x * kWarps;
#else
        uint32_t piece = 0;
        if (lane == 0) piece = atomicAdd(&a.hdr->ticket, 1u);
        piece = __shfl_sync(0xFFFFFFFFu, piece, 0);
#endif
        if (piece >= a.n_pieces) break;
        const uint32_t c0 = piece * a.span_chunks;
        const uint32_t c1 = c0 + a.span_chunks < a.n_chunks ? c0 + a.span_chunks : a.n_chunks;
        uint32_t carry_epb = 0, piece_nsc = 0;
        ChunkRegs cur;
        load_chunk(cur, a.in, (uint64_t)c0 * kChunk, n16, lane);
#pragma unroll 1
        for (uint32_t c = c0; c < c1; c++) {
            const uint64_t pos = (uint64_t)c * kChunk;
            ChunkRegs nxt;
            if (c + 1 < c1) load_chunk(nxt, a.in, pos + kChunk, n16, lane);
            // ---------------------------------------------------------------- detect
            uint32_t acc = 0xFFFFFFFFu;
#ifndef H264B_EXP_NODETECT
#pragma unroll
            for (int r = 0; r < kRows; r++) {
                uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, cur.v[r].w, 1);
                const uint32_t wrap = r ? __shfl_sync(0xFFFFFFFFu, cur.v[r ? r - 1 : 0].w, 31) : cur.t8.y;
                if (lane == 0) prev = wrap;
                const uint32_t w[4] = {cur.v[r].x, cur.v[r].y, cur.v[r].z, cur.v[r].w};
                acc = zero_pair_acc(acc, w, prev);
            }
            acc = zero_pair_acc_tail8(acc, cur.t8.x, cur.t8.y);
#endif
            const bool edge = pos == 0 || pos + kChunk + kHalo > a.n;
            bool clean = !(__any_sync(0xFFFFFFFFu, acc_has_pair(acc)) || carry_epb != 0 || edge);
            if (!clean) {
                // stage the chunk and its halos exactly as a bulk load of [lo, hi) would have left them
                uint8_t *tile_in = st.buf + kHalo;
#pragma unroll
                for (int r = 0; r < kRows; r++)
                    if (pos + (uint32_t)(r * 32 + lane) * 16u < n16)
                        *reinterpret_cast<uint4 *>(tile_in + (r * 32 + lane) * 16) = cur.v[r];
                if (lane == 0 && pos) *reinterpret_cast<uint4 *>(st.buf) = *reinterpret_cast<const uint4 *>(a.in + pos - kHalo);
                if (lane == 1 && pos + kChunk < n16)
                    *reinterpret_cast<uint4 *>(tile_in + kChunk) = *reinterpret_cast<const uint4 *>(a.in + pos + kChunk);
                __syncwarp();
                const ChunkResult res = general_chunk(a, st.scbits, st.buf, pos, carry_epb, piece_nsc, lane);
                carry_epb = res.carry_epb;
                piece_nsc = res.piece_nsc;
                clean = res.clean != 0;
                __syncwarp();
            }
#ifndef H264B_EXP_NOSTORE
            if (clean) {
#pragma unroll
                for (int r = 0; r < kRows; r++)
                    __stcs(reinterpret_cast<uint4 *>(a.out + pos + (uint32_t)(r * 32 + lane) * 16u), cur.v[r]);
            }
#endif
            cur = nxt;
        }
        if (lane == 0) {
            a.piece_epb[piece] = carry_epb;
            a.piece_nsc[piece] = piece_nsc;
        }
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

// Post-pass 1: ordinal of every piece's first start code = exclusive scan of the per-piece start-code counts.  One CTA;
// the array has one entry per piece (128 KiB of stream): every thread sums a contiguous run, one block scan, done.
__global__ void __launch_bounds__(1024) piece_order_kernel(const uint32_t *piece_nsc, uint32_t *piece_ord,
                                                           uint32_t n_pieces) {
    __shared__ uint32_t warp_sum[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t per = (n_pieces + 1023u) / 1024u;
    const uint32_t lo = (uint32_t)tid * per < n_pieces ? (uint32_t)tid * per : n_pieces;
    const uint32_t hi = lo + per < n_pieces ? lo + per : n_pieces;
    uint32_t own = 0;
    for (uint32_t i = lo; i < hi; i++) own += piece_nsc[i];
    uint32_t x = own;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) warp_sum[warp] = x;
    __syncthreads();
    uint32_t w = warp_sum[lane], wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, wi, d);
        if (lane >= d) wi += y;
    }
    uint32_t run = __shfl_sync(0xFFFFFFFFu, wi - w, warp) + x - own;  // start codes before this thread's run
    for (uint32_t i = lo; i < hi; i++) {
        piece_ord[i] = run;
        run += piece_nsc[i];
    }
}

// Post-pass 2: move every record from its slot to its ordinal (stream order): ordinal = first ordinal of the piece
// that holds the start code + the record's rank inside that piece.
__global__ void __launch_bounds__(256) nal_permute_kernel(ScanArgs a) {
    const uint64_t cap = a.nal_cap;
    uint64_t K = a.hdr->total_sc;
    if (K > cap) K = cap;
    const uint64_t piece_bytes = (uint64_t)a.span_chunks * kChunk;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < K; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t st = a.rec_start[i];
        const uint64_t ord = (uint64_t)a.piece_ord[(st - 1) / piece_bytes] + a.rec_rank[i];
        if (ord < cap) {
            a.nal_start[ord] = st;
            a.nal_epb[ord] = a.rec_epb[i];
            a.nal_hdr[ord] = a.rec_hdr[i];
        }
    }
}

// Post-pass 3: the h264b_nal records; NALs whose later parts have to slide left are queued for post-pass 4.
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
        const uint64_t removed = nal_pieces(o.start, next, o.header_bytes, a.nal_epb[k + 1], a.piece_epb,
                                            (uint64_t)a.span_chunks * kChunk,
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

// Post-pass 4: slide the later parts of the queued NALs left (one CTA per NAL, parts in stream order, 4 KiB at a
// time: everything is read into registers before anything is written, so the overlapping move is safe).
__global__ void __launch_bounds__(256) nal_fixup_kernel(ScanArgs a, h264b_scan_summary *summary) {
    const uint32_t n_fix = a.hdr->n_fix;
    if (blockIdx.x == 0 && threadIdx.x == 0) {  // totals of scan_finalize_kernel (complete: previous launch)
        summary->n_epb = a.hdr->n_epb;
        summary->rbsp_bytes = a.hdr->total_kept;  // RBSP bytes of all emitted NAL units
    }
    for (uint32_t f = blockIdx.x; f < n_fix; f += gridDim.x) {
        const uint64_t k = a.fix_list[f];
        const uint64_t st = a.nal_start[k], next = a.nal_start[k + 1];
        h264b_nal o;
        decode_nal_header(a.nal_hdr[k], o, nullptr);
        nal_pieces(st, next, o.header_bytes, a.nal_epb[k + 1], a.piece_epb, (uint64_t)a.span_chunks * kChunk,
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
    uint64_t piece_epb, piece_nsc, piece_ord, fix_list, rec_start, rec_epb, rec_hdr, rec_rank, nal_start, nal_epb,
        nal_hdr, total;
};
static ScratchOffsets scratch_layout(uint64_t n_pieces, uint32_t nal_cap) {
    ScratchOffsets o;
    uint64_t p = sizeof(ScanScratchHeader);
    auto take = [&](uint64_t bytes) {
        const uint64_t at = p;
        p = (p + bytes + 15) & ~15ull;
        return at;
    };
    o.piece_epb = take(n_pieces * 4);
    o.piece_nsc = take(n_pieces * 4);
    o.piece_ord = take(n_pieces * 4);
    o.fix_list = take((uint64_t)nal_cap * 4);
    o.rec_start = take((uint64_t)nal_cap * 8);
    o.rec_epb = take((uint64_t)nal_cap * 4);
    o.rec_hdr = take((uint64_t)nal_cap * 4);
    o.rec_rank = take((uint64_t)nal_cap * 4);
    o.nal_start = take((uint64_t)nal_cap * 8);
    o.nal_epb = take((uint64_t)nal_cap * 4);
    o.nal_hdr = take((uint64_t)nal_cap * 4);
    o.total = (p + 255) & ~255ull;
    return o;
}

int launch_annexb_scan(h264b_ctx *ctx, const uint8_t *d_stream, uint64_t n, uint8_t *d_rbsp, h264b_nal *d_nals,
                       h264b_nal_ext *d_ext, uint32_t nal_cap, h264b_scan_summary *d_summary, uint32_t flags) {
    (void)flags;
    if (((uintptr_t)d_stream & 15) || ((uintptr_t)d_rbsp & 15))
        return set_error(ctx, H264B_E_INVALID, "annexb_scan: d_stream and d_rbsp must be 16-byte aligned");
    if (n >= (1ull << 42)) return set_error(ctx, H264B_E_INVALID, "annexb_scan: stream too long");

    // launch shape: as many warps as fit, each with its own ring of slots
    static bool attr_set = false;
#ifdef H264B_SCAN_LDG
    const auto main_kernel = annexb_scan_ldg_kernel;
    const size_t smem = sizeof(WarpStage) * kWarps;
#else
    const auto main_kernel = annexb_scan_kernel;
    const size_t smem = sizeof(WarpRing) * kWarps;
#endif
    if (!attr_set) {
        H264B_CUDA(ctx, cudaFuncSetAttribute(main_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    int occ = 0;
    H264B_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, main_kernel, kWarps * 32, smem));
    if (occ < 1) occ = 1;
    const uint64_t max_ctas = (uint64_t)ctx->sm_count * occ;

    // pieces: 128 KiB for large streams; smaller when the stream would otherwise leave most warps without work
    const uint64_t n_chunks = (n + kChunk - 1) / kChunk;
    uint64_t span = ctx->scan_span_chunks;
    if (!span) {
        span = n_chunks / (max_ctas * kWarps * 4);
        if (span > kMaxSpanBytes / kChunk) span = kMaxSpanBytes / kChunk;
        if (span < 1) span = 1;
    }
    const uint64_t n_pieces = (n_chunks + span - 1) / span;

    const ScratchOffsets so = scratch_layout(n_pieces, nal_cap);
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
    ScanArgs a;
    a.in = d_stream;
    a.n = n;
    a.out = d_rbsp;
    a.hdr = (ScanScratchHeader *)s;
    a.piece_epb = (uint32_t *)(s + so.piece_epb);
    a.piece_nsc = (uint32_t *)(s + so.piece_nsc);
    a.piece_ord = (uint32_t *)(s + so.piece_ord);
    a.fix_list = (uint32_t *)(s + so.fix_list);
    a.rec_start = (unsigned long long *)(s + so.rec_start);
    a.rec_epb = (uint32_t *)(s + so.rec_epb);
    a.rec_hdr = (uint32_t *)(s + so.rec_hdr);
    a.rec_rank = (uint32_t *)(s + so.rec_rank);
    a.nal_start = (unsigned long long *)(s + so.nal_start);
    a.nal_epb = (uint32_t *)(s + so.nal_epb);
    a.nal_hdr = (uint32_t *)(s + so.nal_hdr);
    a.nal_cap = nal_cap;
    a.n_chunks = (uint32_t)n_chunks;
    a.n_pieces = (uint32_t)n_pieces;
    a.span_chunks = (uint32_t)span;

    scan_init_kernel<<<1, 1, 0, ctx->stream>>>(a.hdr, n);
    H264B_LAUNCH_CHECK(ctx, "scan_init_kernel");
    if (n_pieces) {
        uint64_t grid = (n_pieces + kWarps - 1) / kWarps;
        if (grid > max_ctas) grid = max_ctas;
        main_kernel<<<(int)grid, kWarps * 32, smem, ctx->stream>>>(a);
        H264B_LAUNCH_CHECK(ctx, "annexb_scan_kernel");
        piece_order_kernel<<<1, 1024, 0, ctx->stream>>>(a.piece_nsc, a.piece_ord, a.n_pieces);
        H264B_LAUNCH_CHECK(ctx, "piece_order_kernel");
        nal_permute_kernel<<<ctx->sm_count * 2, 256, 0, ctx->stream>>>(a);
        H264B_LAUNCH_CHECK(ctx, "nal_permute_kernel");
    }
    scan_finalize_kernel<<<ctx->sm_count * 2, 256, 0, ctx->stream>>>(a, d_nals, d_ext, d_summary);
    H264B_LAUNCH_CHECK(ctx, "scan_finalize_kernel");
    nal_fixup_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(a, d_summary);
    H264B_LAUNCH_CHECK(ctx, "nal_fixup_kernel");
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

// upper bound: the smallest pieces (one chunk each); the record arrays are sized by nal_cap on top of this
extern "C" uint64_t h264b_annexb_scratch_bytes(uint64_t n) {
    return h264b::scratch_layout((n + h264b::kChunk - 1) / h264b::kChunk, 0).total;
}
