// annexb_scan.cu -- K1/K2: single-pass Annex-B split + RBSP emulation-prevention strip for sm_100a.
//
// Replaces the byte-at-a-time loops of readNalUnit (h264/server.go:64-111) and NewNalUnit (h264/nalUnit.go:75-131)
// of the reference with one streaming pass: every input byte is read from HBM once and every kept byte written once
// (algorithmic traffic N_in + N_rbsp + 20 B per NAL).
//
// The RBSP of a NAL is written at the position of the NAL's own body (h264b_nal.rbsp_off = start + header_bytes), so
// a span of the stream that holds neither a start code nor an emulation-prevention byte -- all but a few KiB per MB
// of entropy-coded data -- is an aligned copy.  The pass is built around that, as two kernels over 2 KiB chunks:
//
//   K1a  annexb_copy_kernel   one warp per chunk, no shared memory, 36 registers, full occupancy: coalesced LDG.128
//                             (one 16-byte granule per lane and row), word-parallel search for two adjacent zero
//                             bytes (zero_pair_acc: every 00 00 03 and 00 00 00 01 needs such a pair), and -- for a
//                             chunk without an emulation-prevention candidate -- STG.128 of the same registers: a copy
//                             at copy speed (start codes in such a chunk only add records).  A chunk with a candidate
//                             (or at an end of the stream) is only flagged.
//   K1b  annexb_dirty_kernel  the flagged chunks (a few per cent of random data, every chunk of an EPB-dense stream):
//                             the chunk + 16-byte halos are staged in shared memory by a TMA 1-D bulk load
//                             (cp.async.bulk + mbarrier complete_tx, three buffers per warp).  First walk
//                             (chunk_masks): exact masks (granule_masks), start-code bitmap, bit-domain fix-up near
//                             start codes (keep_mask_near_sc), segmented scan of (EPBs | start codes); the chunk's
//                             counts are published.  Look-back (lookback_carry) over the published counts of the
//                             chunks before it: how many bytes the NAL open at the chunk's first byte has lost so
//                             far.  Second walk (chunk_store): the output image of the chunk is built in place
//                             (emulation-prevention bytes spliced out in registers, 32-bit stores, neighbouring
//                             lanes hand over the bytes that straddle a word) and stored with aligned 16-byte
//                             stores at its final position; NAL records for the start codes.
// A kept byte at stream position p goes to out[p - (EPBs removed from p's NAL before p)].  Dirty chunks are handed out in
// ascending order by ticket (dirty_list_kernel compacts the flags first), so a chunk that is waited for is always in the
// hands of a running warp, and a warp takes the counts of its NEXT chunk before it looks back for the current one, so a
// neighbour rarely waits.  What is left for the post-passes: exclusive scans of the per-chunk
// counts (NAL ordinals; S[] for the per-NAL totals; the segmented carry that finds the chunks the copy kernel stored
// verbatim although their NAL had already lost bytes -- those are copied again, shifted, from the input), permutation
// of the records into stream order, and the h264b_nal records (lengths are differences of neighbours).
//
// (Measured alternatives -- 16 KiB CTA tiles with decoupled look-back, warp-private TMA rings with bulk stores,
// register pipelines over 128 KiB pieces, chunks compacted on their own plus a per-NAL slide -- are in the git history
// and profiles/r1_scan_*, profiles/r2_dense_*; DESIGN.md has the numbers.)
#include <stdlib.h>

#include "annexb_local.cuh"
#include "common.cuh"

namespace h264b {

#ifndef H264B_SCAN_ROWS
#define H264B_SCAN_ROWS 4
#endif
#ifndef H264B_SCAN_WARPS
#define H264B_SCAN_WARPS 8
#endif
constexpr int kRows = H264B_SCAN_ROWS;             // 512-byte rows per chunk (one granule per lane and row), <= 16
constexpr int kChunkGran = 32 * kRows;             // 128 granules
constexpr int kChunk = kChunkGran * 16;            // 2048 bytes: the piece of nal_pieces()
constexpr int kHalo = 16;
constexpr int kSlotBytes = kHalo + kChunk + kHalo; // 2080
constexpr int kWarpsA = H264B_SCAN_WARPS;          // chunks per CTA of the copy kernel
constexpr int kWarpsB = 4;                         // warps per CTA of the dirty-chunk kernel (32 chunk flags each)
constexpr int kOrderTile = 4096;                   // chunks per CTA of the ordinal scan
// piece[] word of a chunk
constexpr uint32_t kPieceEpb = 0x7FFFu;            // bits 14:0   EPBs (<= 683 per chunk)
constexpr uint32_t kPieceDirty = 0x8000u;          // bit 15      the copy kernel left the chunk to the dirty-chunk kernel
constexpr uint32_t kPieceNscShift = 16;            // bits 28:16  start codes (<= 512 per chunk)
constexpr uint32_t kPieceNsc = 0x1FFFu;
constexpr uint32_t kPieceReady = 0x80000000u;      // bit 31      (dirty chunks) the fields above are final

struct ScanScratchHeader {   // device scratch; zeroed (together with the two per-chunk arrays behind it) before every pass
    unsigned int n_shift;          // entries of shift_list
    unsigned int n_dirty;          // entries of dirty_list (dirty_list_kernel)
    unsigned long long first_inv;  // max over NAL starts of ~start (0: no start code): first_start = ~first_inv
    unsigned long long total_sc;   // start codes found = slots handed out of the NAL record buffer
    unsigned long long total_kept;
    unsigned long long n_epb;
    unsigned long long ticket;     // next entry of dirty_list to hand out (annexb_dirty_kernel)
    unsigned long long reserved[2];
};  // 64 bytes

struct ScanArgs {
    const uint8_t *in;
    uint64_t n;
    uint8_t *out;
    ScanScratchHeader *hdr;
    uint32_t *piece;           // per chunk, see kPiece*: start codes << 16 | EPBs after its last NAL start (or in the
                               // whole chunk when it has none) | flags; stays 0 for a chunk the copy kernel stored
                               // verbatim and found no start code in
    uint32_t *piece_carry;     // per dirty chunk: kPieceReady | EPBs removed from the NAL open at the END of the chunk
                               // since that NAL's start (the look-back's short cut)
    uint32_t *tile_dirty;      // per kOrderTile chunks: dirty chunks among them (counted by the copy kernel)
    uint32_t *dirty_list;      // the dirty chunks in ascending order (dirty_list_kernel)
    uint32_t *piece_ord;       // per chunk: ordinal of its first start code (exclusive scan of the counts)
    uint32_t *piece_S;         // per chunk: exclusive scan of the EPB fields
    uint2 *shift_list;         // (chunk, G): chunks the copy kernel stored verbatim although the NAL open at their first
                               // byte had lost G emulation-prevention bytes in earlier chunks (order_apply_kernel)
    uint4 *tile_sum;           // per kOrderTile chunks: (start codes, EPB fields, segmented EPB count: flag, count)
    // One 16-byte record per start code: .x/.y = offset of the byte after it (the next NAL's first byte), .z = that
    // NAL's first 4 bytes, .w = EPBs removed (within the start code's chunk) from the NAL that ENDS at this start code
    // | rank of the start code inside its chunk << 16.
    uint4 *rec;                // in slot order (chunks reserve slots with one atomicAdd): written by the dirty-chunk kernel
    uint4 *nal_rec;            // in stream order (written by nal_permute_kernel, read by scan_finalize_kernel)
    uint32_t nal_cap;
    uint32_t n_chunks;
};

__device__ __forceinline__ uint64_t rec_start(const uint4 &r) { return (uint64_t)r.x | ((uint64_t)r.y << 32); }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------------------------------------ K1a: copy + detect
// NAL records of a chunk that holds start codes but no emulation-prevention candidate (rare path of the copy kernel,
// kept out of line).  sc[r]: start-code-end mask of this lane's granule in row r.  Such a chunk loses no byte, so every
// count of removed bytes in its records is zero.
struct ScMasks {
    uint32_t m[kRows];
};
__device__ __noinline__ void emit_records(const uint8_t *in, ScanScratchHeader *hdr, uint32_t *piece, uint4 *rec,
                                          uint32_t nal_cap, uint32_t chunk, int lane, ScMasks scm) {
    const uint32_t *sc = scm.m;
    const uint64_t pos = (uint64_t)chunk * kChunk;
    uint32_t before[kRows], total = 0;  // start codes of the chunk before this lane's granule of row r
#pragma unroll
    for (int r = 0; r < kRows; r++) {
        const uint32_t c = bits_popc(sc[r]);
        uint32_t x = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
            if (lane >= d) x += y;
        }
        before[r] = total + x - c;
        total += __shfl_sync(0xFFFFFFFFu, x, 31);
    }
    unsigned long long slot0 = 0;
    if (lane == 0) {
        slot0 = atomicAdd(&hdr->total_sc, (unsigned long long)total);
        piece[chunk] = total << 16;
    }
    slot0 = __shfl_sync(0xFFFFFFFFu, slot0, 0);
#pragma unroll
    for (int r = 0; r < kRows; r++) {
        uint32_t m = sc[r], rank = before[r];
        while (m) {
            const int j = __ffs((int)m) - 1;
            m &= m - 1;
            const uint64_t st = pos + (uint32_t)(r * 32 + lane) * 16u + (uint32_t)j + 1u;  // the new NAL's first byte
            if (rank == 0) atomicMax(&hdr->first_inv, ~(unsigned long long)st);
            if (slot0 + rank < nal_cap) {
                uint32_t h = 0;  // its first 4 bytes (inside the stream: the chunk is not at its end)
#pragma unroll
                for (int q = 0; q < 4; q++) h |= (uint32_t)in[st + q] << (8 * q);
                rec[slot0 + rank] = make_uint4((uint32_t)st, (uint32_t)(st >> 32), h, rank << 16);
            }
            rank++;
        }
    }
}

// One warp per chunk.  Everything a thread needs is its own granules (4 x 16 bytes, all loads in flight at once), the
// last word of the lane before it (shuffle) and the 4 bytes in front of the chunk (one broadcast load).
//   * no emulation-prevention candidate in the chunk (all but one chunk in thousands): nothing moves, the chunk is
//     stored as it is -- the bytes the reference drops around a start code (its last two bytes, the NAL header) lie
//     between two RBSPs of the position-preserving layout, where the buffer is unspecified; start codes only add records
//   * otherwise, and at the two ends of the stream: the chunk is listed for the dirty-chunk kernel
__global__ void __launch_bounds__(kWarpsA * 32) annexb_copy_kernel(ScanArgs a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t chunk = blockIdx.x * kWarpsA + (uint32_t)warp;
    if (chunk >= a.n_chunks) return;
    const uint64_t pos = (uint64_t)chunk * kChunk;
    // chunks that touch the ends of the stream always take the general path (it blanks the bytes outside)
    if (pos == 0 || pos + kChunk + kHalo > a.n) {
        if (lane == 0) {
            a.piece[chunk] = kPieceDirty;
            atomicAdd(&a.tile_dirty[chunk / kOrderTile], 1u);
        }
        return;
    }
    const uint8_t *src = a.in + pos + lane * 16;
    uint4 v[kRows];
#pragma unroll
    for (int r = 0; r < kRows; r++) v[r] = __ldcs(reinterpret_cast<const uint4 *>(src + r * 512));
    const uint32_t before = *reinterpret_cast<const uint32_t *>(a.in + pos - 4);  // the 4 bytes before the chunk

    uint32_t any_e = 0, any_sc = 0;
    ScMasks sc;
#pragma unroll
    for (int r = 0; r < kRows; r++) {
        uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, v[r].w, 1);
        const uint32_t wrap = r ? __shfl_sync(0xFFFFFFFFu, v[r ? r - 1 : 0].w, 31) : before;
        if (lane == 0) prev = wrap;
        const uint32_t w[4] = {v[r].x, v[r].y, v[r].z, v[r].w};
#ifdef H264B_EXP_NODETECT
        sc.m[r] = 0;
#else
        const GranuleMasks m = granule_masks_filtered(w, prev);
        any_e |= m.e;
        sc.m[r] = m.sc;
        any_sc |= m.sc;
#endif
    }
    if (__any_sync(0xFFFFFFFFu, any_e != 0)) {
        if (lane == 0) {
            a.piece[chunk] = kPieceDirty;
            atomicAdd(&a.tile_dirty[chunk / kOrderTile], 1u);  // (one address per 8 MiB of stream: no hot spot)
        }
        return;
    }
#ifndef H264B_EXP_NOSTORE
    uint8_t *dst = a.out + pos + lane * 16;
#pragma unroll
    for (int r = 0; r < kRows; r++) __stcs(reinterpret_cast<uint4 *>(dst + r * 512), v[r]);
#endif
    if (__any_sync(0xFFFFFFFFu, any_sc != 0)) emit_records(a.in, a.hdr, a.piece, a.rec, a.nal_cap, chunk, lane, sc);
}

// ------------------------------------------------------------------------------------------------ K1b: dirty chunks
constexpr int kStageBufs = 3;  // chunk being written, chunk whose counts are being taken, chunk on its way
struct __align__(16) WarpStage {
    uint8_t buf[kStageBufs][kSlotBytes];  // each: [0,16) low halo, [16,16+kChunk) chunk, high halo
    unsigned long long mbar[kStageBufs];  // "bytes have landed"
    unsigned long long pad0;
    uint16_t scbits[kChunkGran + 2];      // start-code-end bits per granule, [0] = halo granule before the chunk
    uint16_t pad[(8 - (kChunkGran + 2) % 8) % 8];
};
static_assert(sizeof(WarpStage) % 16 == 0 && kSlotBytes % 16 == 0, "staging buffers must stay 16-byte aligned");

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

// Pieces that few lanes need are kept out of line so that the walk of a chunk stays small.
__device__ __noinline__ uint32_t keep_near_sc(const uint8_t *tile_in, uint64_t base, uint64_t gpos, uint32_t e16,
                                              uint32_t sc_prev, uint32_t sc_own, uint32_t sc_next, uint32_t *epb_eff) {
    auto get = [&](int64_t p) -> uint32_t { return tile_in[p - (int64_t)base]; };
    return keep_mask_near_sc(get, (int64_t)gpos, e16, sc_prev, sc_own, sc_next, epb_eff);
}

// EPBs removed so far from the NAL that is open at the first byte of `chunk`: the chunks before it are walked
// backwards, 32 at a time, up to the nearest one that holds a NAL start or has published its own carry.  A dirty chunk
// that has not published its counts yet is waited for; it never waits itself before publishing them, and the warp that
// owns it is running (chunks are dealt to the warps of one wave in ascending order), so the wait ends.
__device__ __noinline__ uint32_t lookback_carry(const ScanArgs &a, uint32_t chunk, int lane) {
    uint32_t carry = 0;
    const volatile uint32_t *piece = a.piece, *pc = a.piece_carry;
    for (int64_t base = (int64_t)chunk; base > 0; base -= 32) {
        const int64_t idx = base - 1 - lane;  // lane 0: the nearest chunk
        uint32_t w, c, term, pending;
        for (;;) {
            if (idx < 0) {  // in front of the stream: nothing is open there
                w = 0;
                c = 0;
                term = 1;
                pending = 0;
            } else {
                w = piece[idx];  // (both loads in flight together)
                c = pc[idx];
                pending = (w & kPieceDirty) && !(w & kPieceReady);
                term = !pending && ((((w >> kPieceNscShift) & kPieceNsc) != 0) || (c & kPieceReady));
            }
            const uint32_t tmask = __ballot_sync(0xFFFFFFFFu, term), pmask = __ballot_sync(0xFFFFFFFFu, pending);
            const uint32_t need = tmask ? ((tmask & (0u - tmask)) - 1u) : 0xFFFFFFFFu;  // the lanes nearer than the first stop
            if ((pmask & need) == 0) {
                uint32_t x = 0;
                if (tmask) {
                    const int first = __ffs((int)tmask) - 1;
                    if (lane < first) x = w & kPieceEpb;
                    // a chunk with a NAL start: the EPBs after its last start; else its published carry
                    if (lane == first) x = idx < 0 ? 0u : (((w >> kPieceNscShift) & kPieceNsc) ? (w & kPieceEpb) : (c & ~kPieceReady));
                } else {
                    x = w & kPieceEpb;
                }
                carry += __reduce_add_sync(0xFFFFFFFFu, x);
                if (tmask) return carry;
                break;
            }
            __nanosleep(20);
        }
    }
    return carry;
}

// What the first walk over a staged chunk leaves for the second one (per lane: its granule of every row).
struct ChunkMasks {
    uint32_t es[kRows];   // emulation-prevention bytes really removed | start-code-end mask << 16
    uint32_t pre[kRows];  // segmented count in front of the granule (inside the chunk): seg_combine format
    uint32_t total;       // ... of the whole chunk (warp-uniform)
};

// First walk (whole warp).  tile_in[i] = s[pos + i] for i in [-16, kChunk + 16); bytes outside the stream are made 0xFF
// here.  Exact masks, start-code bitmap, the bit-domain fix-up near start codes and the segmented counts.
__device__ __forceinline__ ChunkMasks chunk_masks(const ScanArgs &a, uint16_t *scbits, uint8_t *buf, uint64_t pos, int lane) {
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
    const bool has_end = pos + kChunk > a.n;  // some granules reach past the stream
    ChunkMasks cm;
    uint32_t rp = 0;  // segmented count in front of the row (warp-uniform)
#pragma unroll
    for (int r = 0; r < kRows; r++) {
        const int gi = r * 32 + lane;
        const uint64_t gpos = pos + (uint64_t)gi * 16;
        uint32_t e16 = em[r] & 0xFFFFu;
        // start-code ends q in [g-6, g+16] change what this granule keeps
        const uint32_t near = ((uint32_t)scbits[gi] >> 10) | scbits[gi + 1] | (scbits[gi + 2] & 1u);
        if (near)  // header bytes, the 2-byte tail rule and the EPB guard, all in the bit domain
            (void)keep_near_sc(tile_in, pos, gpos, e16, scbits[gi], scbits[gi + 1], scbits[gi + 2], &e16);
        uint32_t sc = em[r] >> 16;
        if (has_end) {
            if (gpos >= a.n) {
                sc = 0;
                e16 = 0;
            } else if (gpos + 16 > a.n) {
                const uint32_t valid = (1u << (uint32_t)(a.n - gpos)) - 1u;
                sc &= valid;
                e16 &= valid;
            }
        }
        cm.es[r] = e16 | (sc << 16);
        uint32_t x;  // inclusive segmented scan inside the row
        if (__all_sync(0xFFFFFFFFu, (e16 | sc) == 0)) {
            x = 0;
        } else if (__all_sync(0xFFFFFFFFu, sc == 0)) {
            x = bits_popc(e16);
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
                if (lane >= d) x += y;
            }
        } else {
            x = warp_seg_scan(seg_element(e16, sc), lane);
        }
        uint32_t ex = __shfl_up_sync(0xFFFFFFFFu, x, 1);
        if (lane == 0) ex = 0;
        cm.pre[r] = seg_combine(rp, ex);
        rp = seg_combine(rp, __shfl_sync(0xFFFFFFFFu, x, 31));
    }
    cm.total = rp;
    return cm;
}

// NAL records of the start-code ends in one lane's granule (few lanes have any).  hdr4 bytes are read from the staged
// chunk, so this runs before the chunk is rewritten in place.
__device__ __noinline__ void emit_dirty_records(const ScanArgs &a, const uint8_t *tile_in, uint64_t pos, int gi, uint32_t ee,
                                                uint32_t sc, uint32_t pre, uint64_t slot0) {
    uint32_t rank = (pre >> 16) & 0x1FFFu;  // start codes of the chunk in front of this granule
    uint32_t c = pre & 0x7FFFu;             // EPBs since the last NAL start / the start of the chunk
    int prev = -1;
    while (sc) {
        const int j = __ffs((int)sc) - 1;
        sc &= sc - 1;
        const uint32_t between = (ee >> (prev + 1)) & ((1u << (j - prev - 1)) - 1u);  // (no EPB at a start-code end)
        c += bits_popc(between);
        const uint64_t st = pos + (uint64_t)gi * 16 + (uint64_t)j + 1;  // the new NAL's first byte
        if (rank == 0) atomicMax(&a.hdr->first_inv, ~(unsigned long long)st);
        if (slot0 + rank < a.nal_cap) {
            const uint8_t *hb = tile_in + gi * 16 + j + 1;  // its first 4 bytes (0xFF past the end of the stream)
            const uint32_t h = (uint32_t)hb[0] | ((uint32_t)hb[1] << 8) | ((uint32_t)hb[2] << 16) | ((uint32_t)hb[3] << 24);
            // .w: EPBs the NAL that ends here lost inside this chunk (nal_removed adds the chunks before from S[])
            a.rec[slot0 + rank] = make_uint4((uint32_t)st, (uint32_t)(st >> 32), h, c | (rank << 16));
        }
        rank++;
        c = 0;
        prev = j;
    }
}

// A granule with start-code ends, byte by byte into the image (few lanes).  x0 = image offset of the granule's first
// byte if nothing had been removed so far; c = EPBs removed so far from the open NAL (plus the sub-granule part of
// the carry while that NAL is the one open at the chunk's first byte); after a start-code end the count restarts.
__device__ __noinline__ void image_boundary_granule(uint8_t *tile_in, int gi, uint4 v4, uint32_t ee, uint32_t sc,
                                                    uint32_t c) {
    const uint32_t v[4] = {v4.x, v4.y, v4.z, v4.w};
    for (int j = 0; j < 16; j++) {
        const uint32_t bit = 1u << j;
        if (ee & bit) {
            c++;
        } else {
            tile_in[gi * 16 + j - (int)c] = (uint8_t)granule_byte(v, j);  // (dropped bytes land between two RBSPs)
        }
        if (sc & bit) c = 0;
    }
}

// 16-byte granules [x0, x1) of the image (offsets relative to tile_in, which is 16-byte aligned like `base`) to
// out[base + x]: aligned 16-byte stores inside, single bytes at the two ragged ends (neighbouring chunks own the
// bytes beyond them).
__device__ __forceinline__ void store_image(uint8_t *out, uint64_t base, const uint8_t *tile_in, int x0, int x1, int lane) {
    if (x1 <= x0) return;
    const int g0 = x0 >> 4, g1 = (x1 + 15) >> 4;  // (arithmetic shift: x0 may be negative)
    const bool head = (x0 & 15) != 0, tail = (x1 & 15) != 0 && (g1 - 1 > g0 || !head);
    for (int g = g0 + (head ? 1 : 0) + lane; g < g1 - ((x1 & 15) ? 1 : 0); g += 32)
        *reinterpret_cast<uint4 *>(out + base + (int64_t)g * 16) = *reinterpret_cast<const uint4 *>(tile_in + g * 16);
    // lanes 0..15: the head granule's bytes; lanes 16..31: the tail granule's
    if (lane < 16) {
        const int x = g0 * 16 + lane;
        if (head && x >= x0 && x < x1) out[base + (int64_t)x] = tile_in[x];
    } else {
        const int x = (g1 - 1) * 16 + (lane - 16);
        if (tail && x >= x0 && x < x1) out[base + (int64_t)x] = tile_in[x];
    }
}

// Second walk (whole warp): the chunk's output image is built in place in the staging buffer -- every byte that is
// not an emulation-prevention byte moves left by the number of those removed from its NAL so far -- and stored with
// aligned 16-byte stores.  Bytes the reference drops around a start code (its last two bytes, the NAL header) are
// carried along: they land between two RBSPs of the position-preserving layout, where the buffer is unspecified.
//   image offset x <-> out[pos - (carry & ~15) + x] for the NAL open at the chunk's first byte (its bytes start at
//   x = -(carry & 15), inside the low halo), out[pos + x] for everything behind the chunk's first start code.
// Lanes exchange the 0..3 bytes that straddle a 32-bit word so that the image is written with word stores.
__device__ __forceinline__ void chunk_store(const ScanArgs &a, uint8_t *buf, uint64_t pos, const ChunkMasks &cm,
                                            uint32_t carry, int lane) {
    uint8_t *tile_in = buf + kHalo;
    uint32_t *tile_w = reinterpret_cast<uint32_t *>(tile_in);
    const uint32_t total = cm.total;
    const uint32_t n_sc = (total >> 16) & 0x1FFFu;
    const uint32_t cb = carry & 15u;
    const uint64_t base_a = pos - (uint64_t)(carry & ~15u);
    const int64_t limit = (int64_t)a.n - (int64_t)pos;  // image offsets (second mapping) stay below this
    if ((total & 0x7FFFu) == 0 && n_sc == 0 && cb == 0) {  // nothing removed, nothing to restart: an aligned copy
        store_image(a.out, base_a, tile_in, 0, (int)(limit < kChunk ? limit : kChunk), lane);
        return;
    }
    // start codes: records first (they read the staged bytes)
    int first_end = 0;    // (n_sc != 0) image offset of the byte behind the chunk's first start code, second mapping,
    int first_a_end = 0;  // and the end of the bytes of the NAL it ends, first mapping
    if (n_sc) {
        unsigned long long slot0 = 0;
        if (lane == 0) slot0 = atomicAdd(&a.hdr->total_sc, (unsigned long long)n_sc);
        slot0 = __shfl_sync(0xFFFFFFFFu, slot0, 0);
        int fe = -1, fa = 0;  // one granule of the chunk holds its first start code: that lane sets these
#pragma unroll
        for (int r = 0; r < kRows; r++) {
            const uint32_t sc = cm.es[r] >> 16;
            if (!sc) continue;
            const int gi = r * 32 + lane;
            emit_dirty_records(a, tile_in, pos, gi, cm.es[r] & 0xFFFFu, sc, cm.pre[r], slot0);
            if (((cm.pre[r] >> 16) & 0x1FFFu) == 0 && fe < 0) {  // no start code of the chunk in front of this granule
                const int j = __ffs((int)sc) - 1;
                fe = gi * 16 + j + 1;
                // the NAL's bytes end two bytes earlier; they have moved by cb + the EPBs in front of them
                fa = fe - 2 - (int)cb - (int)((cm.pre[r] & 0x7FFFu) + bits_popc(cm.es[r] & ((1u << j) - 1u)));
            }
        }
        const int owner = __ffs((int)__ballot_sync(0xFFFFFFFFu, fe >= 0)) - 1;
        first_end = __shfl_sync(0xFFFFFFFFu, fe, owner);
        first_a_end = __shfl_sync(0xFFFFFFFFu, fa, owner);
    }
    // every lane's granules into registers before anything is rewritten
    uint32_t v[kRows][4];
#pragma unroll
    for (int r = 0; r < kRows; r++) {
        const uint4 t = *reinterpret_cast<const uint4 *>(tile_in + (r * 32 + lane) * 16);
        v[r][0] = t.x;
        v[r][1] = t.y;
        v[r][2] = t.z;
        v[r][3] = t.w;
    }
    __syncwarp();
    uint32_t spill_prev_row = 0;  // lane 31's spill of the previous row
    bool reg_prev_row = false;    // ... and whether it feeds this row's lane 0
    uint32_t late_word[kRows];    // bytes this lane has to write itself once the words are in place
    int late_x[kRows], late_n[kRows];
#pragma unroll
    for (int r = 0; r < kRows; r++) {
        const int gi = r * 32 + lane;
        uint32_t ee = cm.es[r] & 0xFFFFu;
        const uint32_t sc = cm.es[r] >> 16;
        const bool regular = sc == 0;
        const uint32_t pre = cm.pre[r];
        const int shift = (int)(pre & 0x7FFFu) + ((pre >> 31) ? 0 : (int)cb);
        const int q = gi * 16 - shift;  // image offset of the granule's first byte
        const int n = 16 - (int)bits_popc(ee);
        // drop the emulation-prevention bytes (highest first: the lower positions stay put)
        uint32_t w0 = v[r][0], w1 = v[r][1], w2 = v[r][2], w3 = v[r][3];
        if (regular) {
            while (ee) {
                const int j = bits_msb(ee);
                ee &= ~(1u << j);
                const uint32_t s0 = __funnelshift_r(w0, w1, 8), s1 = __funnelshift_r(w1, w2, 8),
                               s2 = __funnelshift_r(w2, w3, 8), s3 = w3 >> 8;
                const uint32_t m = (1u << ((j & 3) * 8)) - 1u;  // bytes of word j>>2 that stay
                const int wj = j >> 2;
                w0 = wj > 0 ? w0 : (w0 & m) | (s0 & ~m);
                w1 = wj > 1 ? w1 : (wj == 1 ? (w1 & m) | (s1 & ~m) : s1);
                w2 = wj > 2 ? w2 : (wj == 2 ? (w2 & m) | (s2 & ~m) : s2);
                w3 = wj == 3 ? (w3 & m) | (s3 & ~m) : s3;
            }
        }
        // shifted by the byte offset inside the first word: x0 .. x4 are the image words q>>2 ..
        const uint32_t a8 = (uint32_t)(q & 3) * 8u;
        const uint32_t x0 = w0 << a8, x1 = __funnelshift_l(w0, w1, a8), x2 = __funnelshift_l(w1, w2, a8),
                       x3 = __funnelshift_l(w2, w3, a8), x4 = __funnelshift_l(w3, 0u, a8);
        const int wq = q >> 2;                       // first word (arithmetic shift)
        const int cnt = ((q + n) >> 2) - wq;         // words whose last byte is this lane's: 2 .. 4
        const int rem = (q + n) & 3;                 // bytes of the word behind them: the next lane completes it
        const uint32_t spill = cnt == 2 ? x2 : (cnt == 3 ? x3 : x4);
        // the lane in front (the previous row's last lane for lane 0) hands over its spill if it is a regular granule
        uint32_t sp = __shfl_up_sync(0xFFFFFFFFu, spill, 1);
        uint32_t rg = __shfl_up_sync(0xFFFFFFFFu, (uint32_t)regular, 1);
        if (lane == 0) {
            sp = spill_prev_row;
            rg = reg_prev_row ? 1u : 0u;
        }
        spill_prev_row = __shfl_sync(0xFFFFFFFFu, spill, 31);
        reg_prev_row = __shfl_sync(0xFFFFFFFFu, (uint32_t)regular, 31) != 0;
        uint32_t rn = __shfl_down_sync(0xFFFFFFFFu, (uint32_t)regular, 1);  // does the next lane take this lane's spill?
        if (lane == 31) rn = r + 1 < kRows ? 2u : 0u;                       // (2: decided below, from the next row)
        if (regular) {
            tile_w[wq] = x0 | (rg ? sp : 0u);
            tile_w[wq + 1] = x1;
            if (cnt > 2) tile_w[wq + 2] = x2;
            if (cnt > 3) tile_w[wq + 3] = x3;
        }
        late_word[r] = spill;
        late_x[r] = (wq + cnt) * 4;
        late_n[r] = regular ? (rn == 1u ? 0 : (rn == 2u ? -rem : rem)) : -100;  // -100: a boundary granule; < 0: see below
    }
    // lane 31's spill is taken by the next row's lane 0 if that one is regular
#pragma unroll
    for (int r = 0; r + 1 < kRows; r++) {
        const uint32_t next_reg = __shfl_sync(0xFFFFFFFFu, (uint32_t)((cm.es[r + 1] >> 16) == 0), 0);
        if (lane == 31 && late_n[r] > -100 && late_n[r] <= 0) late_n[r] = next_reg ? 0 : -late_n[r];
    }
    __syncwarp();
    // single bytes: spills nobody took, and the granules with start codes
#pragma unroll
    for (int r = 0; r < kRows; r++) {
        if (late_n[r] == -100) {
            const uint32_t pre = cm.pre[r];
            const uint32_t c = (pre & 0x7FFFu) + ((pre >> 31) ? 0u : cb);
            image_boundary_granule(tile_in, r * 32 + lane, make_uint4(v[r][0], v[r][1], v[r][2], v[r][3]), cm.es[r] & 0xFFFFu,
                                   cm.es[r] >> 16, c);
        } else {
            for (int k = 0; k < late_n[r]; k++) tile_in[late_x[r] + k] = (uint8_t)(late_word[r] >> (8 * k));
        }
    }
    __syncwarp();
    // the image to the output buffer
    const int removed_end = (int)(total & 0x7FFFu);  // EPBs of the NAL open at the chunk's end, since its start / the chunk's
    if (n_sc == 0) {
        int x1 = kChunk - (int)cb - removed_end;
        const int64_t lim = limit - (int64_t)cb;  // (first mapping: out[base_a + x] with base_a + x < n - carry)
        if ((int64_t)x1 > lim) x1 = (int)lim;
        store_image(a.out, base_a, tile_in, -(int)cb, x1, lane);
    } else {
        int xb1 = kChunk - removed_end;
        if ((int64_t)xb1 > limit) xb1 = (int)limit;
        if (base_a == pos) {  // one mapping: one run, the gap at the start code included
            store_image(a.out, pos, tile_in, -(int)cb, xb1, lane);
        } else {
            store_image(a.out, base_a, tile_in, -(int)cb, first_a_end, lane);
            store_image(a.out, pos, tile_in, first_end, xb1, lane);
        }
    }
}

// The dirty chunks in ascending order: one CTA per kOrderTile chunks adds up the counts of the tiles before its own (the
// copy kernel counted while it flagged) and compacts its tile's flags in order.
__global__ void __launch_bounds__(256) dirty_list_kernel(ScanArgs a) {
    __shared__ uint32_t warp_sum[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t before = 0;
    for (uint32_t t = (uint32_t)tid; t < blockIdx.x; t += 256) before += a.tile_dirty[t];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) before += __shfl_xor_sync(0xFFFFFFFFu, before, d);
    if (lane == 0) warp_sum[warp] = before;
    __syncthreads();
    before = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) before += warp_sum[w];
    __syncthreads();
    constexpr int kPer = kOrderTile / 256;  // consecutive chunks per thread
    const uint32_t first = blockIdx.x * kOrderTile + (uint32_t)tid * kPer;
    uint32_t flags = 0, own = 0;
#pragma unroll
    for (int k = 0; k < kPer; k++)
        if (first + k < a.n_chunks && (a.piece[first + k] & kPieceDirty)) {
            flags |= 1u << k;
            own++;
        }
    uint32_t incl = own;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += y;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    uint32_t at = before + incl - own, total = before;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        if (w < warp) at += warp_sum[w];
        total += warp_sum[w];
    }
#pragma unroll
    for (int k = 0; k < kPer; k++)
        if (flags & (1u << k)) a.dirty_list[at++] = first + (uint32_t)k;
    if (blockIdx.x == gridDim.x - 1 && tid == 0) a.hdr->n_dirty = total;
}

// The dirty chunks, handed out in ascending order by ticket (one atomic per chunk; a warp holds the tickets of its next
// two chunks ahead of time).  Per chunk: the TMA stages it; first walk (chunk_masks), after which its counts are
// published; look-back for the carry; second walk (chunk_store).  The first walk of the warp's NEXT chunk runs before
// the look-back of the current one, so the counts a neighbour waits for are out a whole chunk early.  Ascending tickets
// are what the look-back relies on: a chunk a warp can wait for has a lower ticket, so its warp is running -- whatever
// part of the grid is resident beside other kernels -- and the warp with the lowest ticket never waits.
#ifndef H264B_DIRTY_MIN_CTAS
#define H264B_DIRTY_MIN_CTAS 6
#endif
__global__ void __launch_bounds__(kWarpsB * 32, H264B_DIRTY_MIN_CTAS) annexb_dirty_kernel(ScanArgs a) {
    __shared__ WarpStage stage[kWarpsB];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpStage &st = stage[warp];
    if (lane == 0) {
#pragma unroll
        for (int b = 0; b < kStageBufs; b++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&st.mbar[b])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const uint64_t n16 = (a.n + 15) & ~15ull;
    auto stage_chunk = [&](uint32_t chunk, int b) {  // lane 0: TMA bulk load of the chunk and its halos, clipped to the stream
        const uint64_t pos = (uint64_t)chunk * kChunk;
        const uint64_t lo = pos ? pos - kHalo : 0;
        uint64_t hi = pos + kChunk + kHalo;
        if (hi > n16) hi = n16;
        const uint32_t bytes = (uint32_t)(hi - lo), bar = smem_u32(&st.mbar[b]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                smem_u32(st.buf[b] + (lo + kHalo - pos))),
            "l"(a.in + lo), "r"(bytes), "r"(bar)
            : "memory");
    };
    const uint32_t n_dirty = a.hdr->n_dirty;
    // (more warps than dirty chunks: the surplus leaves before it adds to the queue at the ticket counter)
    if ((uint64_t)blockIdx.x * kWarpsB + (uint64_t)warp >= (uint64_t)n_dirty) return;
    const auto next_dirty = [&]() -> int64_t {  // the next dirty chunk nobody has taken, -1: none left (warp-uniform)
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(&a.hdr->ticket, 1ull);
        t = __shfl_sync(0xFFFFFFFFu, t, 0);
        return t < (unsigned long long)n_dirty ? (int64_t)a.dirty_list[t] : -1;
    };
    // c_cur: counted and published, waits for its second walk (buffer b_cur); c_nxt: staged (buffer b_cur + 1);
    // c_far: staged (buffer b_cur + 2)
    int64_t c_cur = -1, c_nxt = next_dirty(), c_far = -1;
    if (c_nxt < 0) return;
    int b_cur = kStageBufs - 1;  // so that c_nxt sits in buffer 0
    if (lane == 0) stage_chunk((uint32_t)c_nxt, 0);
    c_far = next_dirty();
    if (lane == 0 && c_far >= 0) stage_chunk((uint32_t)c_far, 1);
    uint32_t parity = 0;  // bit b: phase parity of buffer b's barrier
    ChunkMasks cm_cur, cm_nxt;
    while (c_cur >= 0 || c_nxt >= 0) {
        const int b_nxt = b_cur + 1 == kStageBufs ? 0 : b_cur + 1;
        if (c_nxt >= 0) {
            mbar_wait(smem_u32(&st.mbar[b_nxt]), (parity >> b_nxt) & 1u);
            parity ^= 1u << b_nxt;
            cm_nxt = chunk_masks(a, st.scbits, st.buf[b_nxt], (uint64_t)c_nxt * kChunk, lane);
            // the chunk's own counts are final: out they go, nobody is kept waiting while this warp waits
            if (lane == 0)
                *(volatile uint32_t *)&a.piece[c_nxt] = kPieceReady | kPieceDirty | (cm_nxt.total & 0x1FFF0000u) | (cm_nxt.total & kPieceEpb);
        }
        if (c_cur >= 0) {
            const uint32_t carry = lookback_carry(a, (uint32_t)c_cur, lane);
            if (lane == 0) *(volatile uint32_t *)&a.piece_carry[c_cur] = kPieceReady | seg_apply(cm_cur.total, carry);
            chunk_store(a, st.buf[b_cur], (uint64_t)c_cur * kChunk, cm_cur, carry, lane);
            // generic-proxy writes to the buffer are ordered before the TMA write that reuses it
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            // this buffer is free: the chunk after c_far goes there (c_far stays in its own)
            const int64_t c_new = next_dirty();
            if (c_nxt >= 0 && c_far >= 0) {
                // (buffers: b_nxt = c_nxt, b_nxt + 1 = c_far, b_cur = free)
            }
            if (lane == 0 && c_new >= 0) stage_chunk((uint32_t)c_new, b_cur);
            // rotate
            c_cur = c_nxt;
            cm_cur = cm_nxt;
            c_nxt = c_far;
            c_far = c_new;
            b_cur = b_nxt;
        } else {  // first round: nothing to store yet
            c_cur = c_nxt;
            cm_cur = cm_nxt;
            c_nxt = c_far;
            c_far = next_dirty();
            b_cur = b_nxt;
            if (lane == 0 && c_far >= 0) stage_chunk((uint32_t)c_far, 2);
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

// Post-pass 1: exclusive scans over the chunks of (a) the start-code counts -> ordinal of every chunk's first start
// code, (b) the EPB fields -> S[] and (c) the SEGMENTED EPB count (it restarts in every chunk with a NAL start) -> G, the
// bytes the NAL open at a chunk's first byte has lost so far; a chunk the copy kernel stored verbatim with G > 0 goes
// on the list of post-pass 4.  Two phases over tiles of kOrderTile chunks: tile totals, then every CTA combines the
// totals of the tiles before its own (a few hundred values for a 4 GB stream) and scans its tile.  Every thread owns
// kOrderTile / 256 consecutive chunks, so that (c), which does not commute, can be folded in order.
struct SegCount {
    uint32_t f, c;  // a NAL start inside the span; EPBs since the last one (since the start of the span without one)
};
__device__ __forceinline__ SegCount segc_combine(SegCount a, SegCount b) {  // a: earlier span, b: later span
    return SegCount{a.f | b.f, b.f ? b.c : a.c + b.c};
}
struct Ord3 {
    uint32_t nsc, epb;
    SegCount seg;
};
__device__ __forceinline__ Ord3 ord3_of(uint32_t p) {
    const uint32_t nsc = (p >> kPieceNscShift) & kPieceNsc, epb = p & kPieceEpb;
    return Ord3{nsc, epb, SegCount{nsc ? 1u : 0u, epb}};
}
__device__ __forceinline__ Ord3 ord3_combine(Ord3 a, Ord3 b) {
    return Ord3{a.nsc + b.nsc, a.epb + b.epb, segc_combine(a.seg, b.seg)};
}
__device__ __forceinline__ Ord3 ord3_shfl_up(Ord3 x, int d) {
    Ord3 y;
    y.nsc = __shfl_up_sync(0xFFFFFFFFu, x.nsc, d);
    y.epb = __shfl_up_sync(0xFFFFFFFFu, x.epb, d);
    y.seg.f = __shfl_up_sync(0xFFFFFFFFu, x.seg.f, d);
    y.seg.c = __shfl_up_sync(0xFFFFFFFFu, x.seg.c, d);
    return y;
}
constexpr Ord3 kOrd3Zero = {0u, 0u, {0u, 0u}};

// In-order scan over the 256 threads of a CTA (thread t's value covers the span right after thread t-1's):
// returns the exclusive prefix of this thread; *total receives the CTA's total.
__device__ __forceinline__ Ord3 block_scan3_256(Ord3 own, Ord3 *warp_sum, int tid, Ord3 *total) {
    const int lane = tid & 31, warp = tid >> 5;
    Ord3 x = own;  // inclusive over the warp
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const Ord3 y = ord3_shfl_up(x, d);
        if (lane >= d) x = ord3_combine(y, x);
    }
    Ord3 ex = ord3_shfl_up(x, 1);
    if (lane == 0) ex = kOrd3Zero;
    if (lane == 31) warp_sum[warp] = x;
    __syncthreads();
    Ord3 before = kOrd3Zero, all = kOrd3Zero;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        if (w < warp) before = ord3_combine(before, warp_sum[w]);
        all = ord3_combine(all, warp_sum[w]);
    }
    __syncthreads();
    *total = all;
    return ord3_combine(before, ex);
}

__global__ void __launch_bounds__(256) order_reduce_kernel(const uint32_t *piece, uint4 *tile_sum, uint32_t n_chunks) {
    __shared__ Ord3 warp_sum[8];
    const int tid = threadIdx.x;
    constexpr int kPer = kOrderTile / 256;
    const uint32_t first = blockIdx.x * kOrderTile + (uint32_t)tid * kPer;
    Ord3 own = kOrd3Zero;
#pragma unroll
    for (int k = 0; k < kPer; k++)
        if (first + k < n_chunks) own = ord3_combine(own, ord3_of(piece[first + k]));
    Ord3 total;
    (void)block_scan3_256(own, warp_sum, tid, &total);
    if (tid == 0) tile_sum[blockIdx.x] = make_uint4(total.nsc, total.epb, total.seg.f, total.seg.c);
}

__global__ void __launch_bounds__(256) order_apply_kernel(ScanArgs a) {
    __shared__ Ord3 warp_sum[8];
    const int tid = threadIdx.x, lane = tid & 31;
    // the tiles before this one, in order: thread t folds tiles [t m, (t + 1) m)
    const uint32_t nb = blockIdx.x, m = (nb + 255u) / 256u;
    Ord3 mine = kOrd3Zero;
    for (uint32_t t = (uint32_t)tid * m; t < (uint32_t)(tid + 1) * m && t < nb; t++) {
        const uint4 s = a.tile_sum[t];
        mine = ord3_combine(mine, Ord3{s.x, s.y, SegCount{s.z, s.w}});
    }
    Ord3 before;
    (void)block_scan3_256(mine, warp_sum, tid, &before);
    constexpr int kPer = kOrderTile / 256;  // consecutive chunks per thread
    const uint32_t first = blockIdx.x * kOrderTile + (uint32_t)tid * kPer;
    uint32_t p[kPer];
    Ord3 own = kOrd3Zero;
#pragma unroll
    for (int k = 0; k < kPer; k++) {
        p[k] = first + k < a.n_chunks ? a.piece[first + k] : 0u;
        own = ord3_combine(own, ord3_of(p[k]));
    }
    Ord3 total;
    Ord3 run = ord3_combine(before, block_scan3_256(own, warp_sum, tid, &total));
    uint32_t G[kPer];
    uint32_t n_mine = 0;  // chunks of this thread for the list
#pragma unroll
    for (int k = 0; k < kPer; k++) {
        G[k] = 0;
        if (first + k < a.n_chunks) {
            a.piece_ord[first + k] = run.nsc;
            a.piece_S[first + k] = run.epb;
            // stored verbatim, some NAL is open at its first byte, and that NAL has lost bytes already
            if (!(p[k] & kPieceDirty) && run.nsc && run.seg.c) {
                G[k] = run.seg.c;
                n_mine++;
            }
        }
        run = ord3_combine(run, ord3_of(p[k]));
    }
    // one slot reservation per warp
    uint32_t incl = n_mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += y;
    }
    const uint32_t warp_total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    if (warp_total) {
        uint32_t slot = 0;
        if (lane == 31) slot = atomicAdd(&a.hdr->n_shift, warp_total);
        slot = __shfl_sync(0xFFFFFFFFu, slot, 31) + incl - n_mine;
#pragma unroll
        for (int k = 0; k < kPer; k++)
            if (G[k]) a.shift_list[slot++] = make_uint2(first + (uint32_t)k, G[k]);
    }
}

// Post-pass 2: move every record from its slot to its ordinal (stream order): ordinal = first ordinal of the chunk
// that holds the start code + the record's rank inside that chunk.
__global__ void __launch_bounds__(256) nal_permute_kernel(ScanArgs a) {
    const uint64_t cap = a.nal_cap;
    uint64_t K = a.hdr->total_sc;
    if (K > cap) K = cap;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < K; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 r = a.rec[i];
        const uint64_t ord = (uint64_t)a.piece_ord[(rec_start(r) - 1) / kChunk] + (r.w >> 16);
        if (ord < cap) a.nal_rec[ord] = r;
    }
}

// Post-pass 3: the h264b_nal records.
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
        const uint4 r0 = a.nal_rec[k], r1 = a.nal_rec[k + 1];
        o.start = rec_start(r0);
        const uint64_t next = rec_start(r1);
        o.num_bytes = (uint32_t)(next - o.start);
        decode_nal_header(r0.z, o, ext ? &ext[k] : nullptr);
        uint32_t later_shift;
        const uint64_t removed = nal_removed(o.start, next, r1.w & 0xFFFFu, a.piece_S, (uint64_t)kChunk, &later_shift);
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
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {  // one pair of atomics per warp
        epb += __shfl_xor_sync(0xFFFFFFFFu, epb, d);
        rbsp += __shfl_xor_sync(0xFFFFFFFFu, rbsp, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (epb) atomicAdd(&a.hdr->n_epb, epb);
        if (rbsp) atomicAdd(&a.hdr->total_kept, rbsp);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        summary->n_start_codes = K;
        summary->n_nals = n_nals;
        summary->first_start = K ? ~a.hdr->first_inv : a.n;
        summary->status = (K > a.nal_cap) ? H264B_E_CAPACITY : H264B_OK;
        summary->reserved = 0;
    }
}

// Post-pass 4: chunks the copy kernel stored verbatim although the NAL open at their first byte had lost emulation-
// prevention bytes in an earlier chunk (G > 0 of them): that NAL's bytes of the chunk are copied again from the input
// stream, G bytes to the left.  Every such chunk is independent of every other (the dirty-chunk kernel has written its
// own chunks at their final place already), and real streams have few: a warp looks at 32 chunks, then copies the
// flagged ones, one destination-aligned 16-byte granule per lane and step, assembled from aligned source words.
__device__ __forceinline__ void shifted_copy_warp(const uint8_t *in, uint8_t *out, uint64_t ps, uint64_t pe, uint64_t G,
                                                  int lane) {
    const uint64_t d0 = ps - G, d1 = pe - G;  // destination range: at most kChunk bytes, kChunk / 512 + 1 rounds of the warp
    constexpr int kSteps = kChunk / 512 + 1;
    const uint64_t Dl = (d0 & ~15ull) + (uint64_t)lane * 16u;
    uint32_t y[kSteps][4];
#pragma unroll
    for (int k = 0; k < kSteps; k++) {  // all loads first
        const uint64_t D = Dl + (uint64_t)k * 512u;
        if (D < d1) {
            const uint64_t src = D + G;
            const uint32_t *wp = reinterpret_cast<const uint32_t *>(in + (src & ~3ull));
            const uint32_t sh = (uint32_t)(src & 3u) * 8u;
            const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3], w4 = sh ? wp[4] : 0u;
            y[k][0] = __funnelshift_r(w0, w1, sh);
            y[k][1] = __funnelshift_r(w1, w2, sh);
            y[k][2] = __funnelshift_r(w2, w3, sh);
            y[k][3] = __funnelshift_r(w3, w4, sh);
        }
    }
#pragma unroll
    for (int k = 0; k < kSteps; k++) {
        const uint64_t D = Dl + (uint64_t)k * 512u;
        if (D < d1) {
            if (D >= d0 && D + 16 <= d1) {
                *reinterpret_cast<uint4 *>(out + D) = make_uint4(y[k][0], y[k][1], y[k][2], y[k][3]);
            } else {
#pragma unroll
                for (int j = 0; j < 16; j++)
                    if (D + j >= d0 && D + j < d1) out[D + j] = (uint8_t)granule_byte(y[k], j);
            }
        }
    }
}

__global__ void __launch_bounds__(256) chunk_shift_kernel(ScanArgs a, h264b_scan_summary *summary) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {  // totals of scan_finalize_kernel (complete: previous launch)
        summary->n_epb = a.hdr->n_epb;
        summary->rbsp_bytes = a.hdr->total_kept;  // RBSP bytes of all emitted NAL units
    }
    const int lane = threadIdx.x & 31;
    const uint32_t n = a.hdr->n_shift;
    const uint32_t warps = gridDim.x * 8u;
    for (uint32_t i = blockIdx.x * 8u + (threadIdx.x >> 5); i < n; i += warps) {  // one warp per listed chunk
        const uint2 e = a.shift_list[i];
        const uint64_t t = e.x, G = e.y;
        const uint32_t p = a.piece[t];
        const uint64_t k = (uint64_t)a.piece_ord[t] - 1u;  // the NAL open at the chunk's first byte
        if (k + 1 >= a.nal_cap) continue;  // (an index too small for the stream is reported as such: nothing to do)
        const uint64_t lo = t * (uint64_t)kChunk;
        uint64_t ps = lo, pe = lo + kChunk;
        const uint4 r0 = a.nal_rec[k];
        const uint64_t body = rec_start(r0) + nal_header_bytes(r0.z & 0xFFu, (r0.z >> 8) & 0xFFu);
        if (body > ps) ps = body;
        if ((p >> kPieceNscShift) & kPieceNsc)  // the NAL ends in this chunk: its last two bytes stay out
            pe = rec_start(a.nal_rec[k + 1]) - 2;
        if (pe > ps) shifted_copy_warp(a.in, a.out, ps, pe, G, lane);
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
                                                            uint32_t *n_out, uint32_t *n_found) {
    __shared__ uint32_t warp_cnt[32];
    __shared__ uint32_t base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint64_t n = summary->n_nals;
    if (n > nal_cap) n = nal_cap;
    if (summary->status != H264B_OK) n = 0;  // an incomplete index holds no records at all (scan_finalize_kernel)
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
    if (tid == 0) {
        *n_out = base < max_slices ? base : max_slices;
        if (n_found) *n_found = base;  // all slice NAL units, whatever the list holds
    }
}

// ------------------------------------------------------------------------------------------------ launchers
struct ScratchOffsets {
    uint64_t piece, piece_carry, tile_dirty, zero_end, dirty_list, piece_ord, piece_S, shift_list, tile_sum, rec, nal_rec, total;
};
static ScratchOffsets scratch_layout(uint64_t n, uint32_t nal_cap) {
    const uint64_t n_chunks = (n + kChunk - 1) / kChunk;
    ScratchOffsets o;
    uint64_t p = sizeof(ScanScratchHeader);
    auto take = [&](uint64_t bytes) {
        const uint64_t at = p;
        p = (p + bytes + 15) & ~15ull;
        return at;
    };
    o.piece = take(n_chunks * 4);  // directly behind the header, then piece_carry and tile_dirty: one memset clears all
    o.piece_carry = take(n_chunks * 4);
    o.tile_dirty = take((n_chunks + kOrderTile - 1) / kOrderTile * 4);
    o.zero_end = p;  // everything up to here is cleared before a pass
    o.dirty_list = take(n_chunks * 4);
    o.piece_ord = take(n_chunks * 4);
    o.piece_S = take(n_chunks * 4);
    o.shift_list = take(n_chunks * 8);
    o.tile_sum = take((n_chunks + kOrderTile - 1) / kOrderTile * 16);
    o.rec = take((uint64_t)nal_cap * 16);
    o.nal_rec = take((uint64_t)nal_cap * 16);
    o.total = (p + 255) & ~255ull;
    return o;
}

int launch_annexb_scan(h264b_ctx *ctx, const uint8_t *d_stream, uint64_t n, uint8_t *d_rbsp, h264b_nal *d_nals,
                       h264b_nal_ext *d_ext, uint32_t nal_cap, h264b_scan_summary *d_summary, uint32_t flags) {
    TraceRange trace_range("h264b:annexb_scan");
    (void)flags;
    if (((uintptr_t)d_stream & 15) || ((uintptr_t)d_rbsp & 15))
        return set_error(ctx, H264B_E_INVALID, "annexb_scan: d_stream and d_rbsp must be 16-byte aligned");
    if (n >= (1ull << 42)) return set_error(ctx, H264B_E_INVALID, "annexb_scan: stream too long");
    const ScratchOffsets so = scratch_layout(n, nal_cap);
    void *&scratch = ctx->scan_scratch[ctx->bank];
    size_t &scratch_bytes = ctx->scan_scratch_bytes[ctx->bank];
    if (so.total > scratch_bytes) {
        if (scratch) {
            H264B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            cudaFree(scratch);
        }
        scratch = nullptr;
        scratch_bytes = 0;
        H264B_CUDA(ctx, cudaMalloc(&scratch, so.total));
        scratch_bytes = so.total;
    }
    uint8_t *s = (uint8_t *)scratch;
    const uint64_t n_chunks = (n + kChunk - 1) / kChunk;
    ScanArgs a;
    a.in = d_stream;
    a.n = n;
    a.out = d_rbsp;
    a.hdr = (ScanScratchHeader *)s;
    a.piece = (uint32_t *)(s + so.piece);
    a.piece_carry = (uint32_t *)(s + so.piece_carry);
    a.tile_dirty = (uint32_t *)(s + so.tile_dirty);
    a.dirty_list = (uint32_t *)(s + so.dirty_list);
    a.piece_ord = (uint32_t *)(s + so.piece_ord);
    a.piece_S = (uint32_t *)(s + so.piece_S);
    a.shift_list = (uint2 *)(s + so.shift_list);
    a.tile_sum = (uint4 *)(s + so.tile_sum);
    a.rec = (uint4 *)(s + so.rec);
    a.nal_rec = (uint4 *)(s + so.nal_rec);
    a.nal_cap = nal_cap;
    a.n_chunks = (uint32_t)n_chunks;

    // Small streams are launch bound: the pass is captured once per set of arguments and replayed as one graph launch.
    static const bool graphs_on = !(getenv("H264B_SCAN_GRAPH") && atoi(getenv("H264B_SCAN_GRAPH")) == 0);
    static int occ_b = 0;  // resident CTAs per SM of the dirty-chunk kernel: it walks its list grid-stride, one wave
    if (!occ_b) {
        H264B_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_b, annexb_dirty_kernel, kWarpsB * 32, 0));
        if (occ_b < 1) occ_b = 1;
    }
    const auto enqueue = [&]() -> int {
        unsigned launched = 0;
        // header + the per-chunk counts and carries: all zero
        H264B_CUDA(ctx, cudaMemsetAsync(s, 0, so.zero_end, ctx->stream));
        if (n_chunks) {
            annexb_copy_kernel<<<(unsigned)((n_chunks + kWarpsA - 1) / kWarpsA), kWarpsA * 32, 0, ctx->stream>>>(a);
            H264B_LAUNCH_CHECK(ctx, "annexb_copy_kernel");
            const unsigned tiles_d = (unsigned)((n_chunks + kOrderTile - 1) / kOrderTile);
            dirty_list_kernel<<<tiles_d, 256, 0, ctx->stream>>>(a);
            H264B_LAUNCH_CHECK(ctx, "dirty_list_kernel");
            uint64_t grid_b = (n_chunks + kWarpsB - 1) / kWarpsB;  // (the list is at most that long)
            if (grid_b > (uint64_t)ctx->sm_count * occ_b) grid_b = (uint64_t)ctx->sm_count * occ_b;
            annexb_dirty_kernel<<<(unsigned)grid_b, kWarpsB * 32, 0, ctx->stream>>>(a);
            H264B_LAUNCH_CHECK(ctx, "annexb_dirty_kernel");
            const unsigned tiles = (unsigned)((n_chunks + kOrderTile - 1) / kOrderTile);
            order_reduce_kernel<<<tiles, 256, 0, ctx->stream>>>(a.piece, a.tile_sum, a.n_chunks);
            H264B_LAUNCH_CHECK(ctx, "order_reduce_kernel");
            order_apply_kernel<<<tiles, 256, 0, ctx->stream>>>(a);
            H264B_LAUNCH_CHECK(ctx, "order_apply_kernel");
            nal_permute_kernel<<<ctx->sm_count * 2, 256, 0, ctx->stream>>>(a);
            H264B_LAUNCH_CHECK(ctx, "nal_permute_kernel");
            launched += 6;
        }
        scan_finalize_kernel<<<ctx->sm_count * 2, 256, 0, ctx->stream>>>(a, d_nals, d_ext, d_summary);
        H264B_LAUNCH_CHECK(ctx, "scan_finalize_kernel");
        {
            uint64_t grid_s = (n_chunks + 7) / 8;  // one warp per listed chunk, grid-stride
            if (grid_s > (uint64_t)ctx->sm_count * 8) grid_s = (uint64_t)ctx->sm_count * 8;
            if (grid_s < 1) grid_s = 1;
            chunk_shift_kernel<<<(unsigned)grid_s, 256, 0, ctx->stream>>>(a, d_summary);
            H264B_LAUNCH_CHECK(ctx, "chunk_shift_kernel");
        }
        (void)launched;
        return H264B_OK;
    };
    if (!graphs_on || n > (8ull << 20)) return enqueue();
    ScanGraph key;
    memset(&key, 0, sizeof(key));
    {
        unsigned char *k = key.key;
        const void *ptrs[6] = {d_stream, d_rbsp, d_nals, d_ext, d_summary, s};
        memcpy(k, ptrs, sizeof(ptrs));
        memcpy(k + 48, &n, 8);
        memcpy(k + 56, &nal_cap, 4);
        const void *st = (const void *)ctx->stream;
        memcpy(k + 64, &st, 8);
    }
    ScanGraph *hit = nullptr, *victim = nullptr;  // victim: an empty entry, else the least recently used one
    for (int i = 0; i < kScanGraphs; i++) {
        ScanGraph &g = ctx->scan_graph[i];
        if (g.exec && memcmp(g.key, key.key, sizeof(key.key)) == 0) hit = &g;
        if (!g.exec) {
            if (!victim || victim->exec) victim = &g;
        } else if (!victim || (victim->exec && g.used < victim->used)) {
            victim = &g;
        }
    }
    if (!hit) {
        const uint64_t launches_before = ctx->launches;
        if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError();
            return enqueue();  // (a stream that cannot be captured: the legacy default stream, or one already capturing)
        }
        const int rc = enqueue();
        cudaGraph_t graph = nullptr;
        const cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
        if (rc != H264B_OK || e != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            ctx->launches = launches_before;
            return rc != H264B_OK ? rc : enqueue();
        }
        cudaGraphExec_t exec = nullptr;
        if (cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
            cudaGraphDestroy(graph);
            cudaGetLastError();
            ctx->launches = launches_before;
            return enqueue();
        }
        cudaGraphDestroy(graph);
        if (victim->exec) cudaGraphExecDestroy(victim->exec);
        memcpy(victim->key, key.key, sizeof(key.key));
        victim->exec = exec;
        victim->nodes = (unsigned)(ctx->launches - launches_before);
        ctx->launches = launches_before;
        hit = victim;
    }
    hit->used = ++ctx->scan_graph_tick;
    H264B_CUDA(ctx, cudaGraphLaunch(hit->exec, ctx->stream));
    ctx->launches += hit->nodes;
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
                        uint32_t *d_len, uint32_t *d_slice_nal, uint32_t *d_n_slices, uint32_t *d_n_found) {
    slice_select_kernel<<<1, 1024, 0, ctx->stream>>>(d_nals, d_summary, nal_cap, slice_data_offset, max_slices, d_off,
                                                     d_len, d_slice_nal, d_n_slices, d_n_found);
    H264B_LAUNCH_CHECK(ctx, "slice_select_kernel");
    return H264B_OK;
}

}  // namespace h264b

// (the record arrays are sized by nal_cap on top of this)
extern "C" uint64_t h264b_annexb_scratch_bytes(uint64_t n) { return h264b::scratch_layout(n, 0).total; }
