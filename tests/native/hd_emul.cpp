// hd_emul.cpp -- CPU harness around the product's __host__ __device__ headers (annexb_local.cuh, cabac_lane.cuh).
// Built by tests/test_hd_logic.py with g++.  It drives the very same per-byte predicates / per-lane arithmetic the
// CUDA kernels use, in the same order of decisions (fast path vs exact path), so that their logic can be checked
// against the oracle without a GPU.  Test infrastructure: not part of the product.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../h264decode_b200/csrc/annexb_local.cuh"
#include "../../h264decode_b200/csrc/cabac_lane.cuh"
#include "../../h264decode_b200/csrc/ctx_glue.cuh"
#include "../../h264decode_b200/csrc/param_sets.cuh"
#include "../../h264decode_b200/csrc/tables.inc"

using namespace h264b;

extern "C" {

// Whole-stream split + strip with the position-local rules, granule by granule like annexb_scan_kernel:
// granules with a start-code end in [g-6, g+16] use keep_mask_near_sc, the others use ~raw-EPB mask; both are
// compared with the exact per-byte predicate keep_byte_stream.
// Outputs: nal_start / nal_rbsp_off / nal_hdr per start code (K entries), rbsp bytes; returns K.
// *fast_slow_mismatch counts granules where the fast mask differs from the exact predicate although no start code
// is near (must be 0).
int64_t emul_stream(const uint8_t *s, int64_t n, uint64_t *nal_start, uint64_t *nal_rbsp_off, uint32_t *nal_hdr,
                    int64_t cap, uint8_t *rbsp, int64_t *rbsp_total, int64_t *first_start, int64_t *fast_slow_mismatch) {
    auto get = [&](int64_t p) -> uint32_t { return (p >= 0 && p < n) ? s[p] : 0xFFu; };
    const int64_t n_gran = (n + 15) / 16;
    std::vector<uint16_t> sc(n_gran + 2, 0), em(n_gran + 2, 0);
    int64_t e0 = n;
    for (int64_t g = 0; g < n_gran; g++) {
        uint32_t w[4];
        for (int i = 0; i < 4; i++)
            w[i] = get(g * 16 + i * 4) | (get(g * 16 + i * 4 + 1) << 8) | (get(g * 16 + i * 4 + 2) << 16) |
                   (get(g * 16 + i * 4 + 3) << 24);
        uint32_t prev = get(g * 16 - 4) | (get(g * 16 - 3) << 8) | (get(g * 16 - 2) << 16) | (get(g * 16 - 1) << 24);
        GranuleMasks m = granule_masks(w, prev);
        sc[g + 1] = (uint16_t)m.sc;
        em[g + 1] = (uint16_t)m.e;
        if (m.sc && e0 == n) {
            int64_t q = g * 16 + __builtin_ctz(m.sc);
            if (q < n) e0 = q + 1;
        }
    }
    *first_start = e0;
    int64_t K = 0, kept = 0, mism = 0;
    for (int64_t g = 0; g < n_gran; g++) {
        const int64_t gpos = g * 16;
        uint32_t k16 = ~(uint32_t)em[g + 1] & 0xFFFFu;
        const uint32_t near = ((uint32_t)sc[g] >> 10) | sc[g + 1] | (sc[g + 2] & 1u);
        uint32_t exact = 0;
        for (int j = 0; j < 16; j++)
            if (keep_byte_stream(get, gpos + j)) exact |= 1u << j;
        if (near) {
            uint32_t ee_unused;
            k16 = keep_mask_near_sc(get, gpos, em[g + 1], sc[g], sc[g + 1], sc[g + 2], &ee_unused);
            if (k16 != exact) mism++;
        } else if (k16 != exact) {
            mism++;
        }
        if (gpos + 16 <= e0)
            k16 = 0;
        else if (gpos < e0)
            k16 &= ~((1u << (uint32_t)(e0 - gpos)) - 1u);
        uint32_t scm = sc[g + 1];
        if (gpos + 16 > n) {
            uint32_t valid = (1u << (uint32_t)(n - gpos)) - 1u;
            k16 &= valid;
            scm &= valid;
        }
        for (int j = 0; j < 16; j++) {
            if (scm & (1u << j)) {
                if (K < cap) {
                    const int64_t a = gpos + j + 1;
                    nal_start[K] = (uint64_t)a;
                    nal_rbsp_off[K] = (uint64_t)kept;
                    nal_hdr[K] = get(a) | (get(a + 1) << 8) | (get(a + 2) << 16) | (get(a + 3) << 24);
                }
                K++;
            }
            if (k16 & (1u << j)) rbsp[kept++] = s[gpos + j];
        }
    }
    *rbsp_total = kept;
    *fast_slow_mismatch = mism;
    return K;
}

uint32_t emul_header_bytes(uint32_t b0, uint32_t b1) { return nal_header_bytes(b0, b1); }

// Chunk-by-chunk emulation of annexb_copy_kernel + annexb_dirty_kernel and the post-passes (same phases, same helper
// functions; the lanes of a warp as loops).  The product treats every 2 KiB chunk as a piece of its own
// (span_chunks == 1); larger pieces (a carry across the chunks of a piece) are exercised too because the helpers
// support them.  Pieces are visited in a pseudo-random order (order_seed) because the GPU runs them in any order;
// records go to "slots" in visiting order and are permuted into stream order exactly as order_*_kernel /
// nal_permute_kernel do.  RBSP bytes of a NAL land at the NAL body's own position in `out`
// (position-preserving layout); `out` holds out_shift + n + 64 bytes, pre-filled by the caller so that stray writes
// are detectable; out_shift (a multiple of 16 in the product) shifts the whole destination to exercise alignment.
// nal_start / nal_epb / nal_hdr (cap entries) receive the per-start-code index (nal_epb: the NAL's EPB total, as
// scan_finalize_kernel computes it).  Returns the number of start codes.
int64_t emul_stream_tiles(const uint8_t *s, int64_t n, uint8_t *out_base, int64_t out_shift, uint64_t *nal_start,
                          uint64_t *nal_epb, uint32_t *nal_hdr, int64_t cap, int64_t span_chunks, uint32_t order_seed,
                          int64_t *stats) {
    const int kRows = 4, kGran = 32 * kRows, kChunk = kGran * 16, kHalo = 16;
    uint8_t *out = out_base + out_shift;
    auto gets = [&](int64_t p) -> uint32_t { return (p >= 0 && p < n) ? s[p] : 0xFFu; };
    const int64_t n_chunks = (n + kChunk - 1) / kChunk;
    if (span_chunks < 1) span_chunks = 1;
    const int64_t n_pieces = (n_chunks + span_chunks - 1) / span_chunks;
    const uint64_t piece_bytes = (uint64_t)span_chunks * kChunk;
    std::vector<uint32_t> piece_epb((size_t)n_pieces + 1, 0), piece_nsc((size_t)n_pieces + 1, 0),
        piece_ord((size_t)n_pieces + 1, 0);
    std::vector<uint64_t> rec_start;
    std::vector<uint32_t> rec_epb, rec_hdr, rec_rank;
    std::vector<int64_t> order((size_t)n_pieces);
    for (int64_t i = 0; i < n_pieces; i++) order[i] = i;
    uint32_t lcg = order_seed;
    if (order_seed)
        for (int64_t i = n_pieces - 1; i > 0; i--) {
            lcg = lcg * 1664525u + 1013904223u;
            std::swap(order[i], order[(lcg >> 8) % (uint32_t)(i + 1)]);
        }
    int64_t n_fast = 0, n_false_alarm = 0, n_general = 0, n_filter_mismatch = 0;
    std::vector<uint8_t> buf(kHalo + kChunk + kHalo);
    std::vector<uint16_t> scb(kGran + 2);
    for (int64_t oi = 0; oi < n_pieces; oi++) {
        const int64_t piece = order[oi];
        uint32_t carry_epb = 0, pnsc = 0;
        const int64_t c_end = std::min((piece + 1) * span_chunks, n_chunks);
        for (int64_t chunk = piece * span_chunks; chunk < c_end; chunk++) {
            const int64_t pos = chunk * kChunk;
            for (int i = 0; i < kHalo + kChunk + kHalo; i++) buf[i] = (uint8_t)gets(pos - kHalo + i);
            uint8_t *tile_in = buf.data() + kHalo;
            // ---- detect (fast path)
            // the copy kernel's filter (annexb_copy_kernel): per granule, exact masks only where the cheap test fires
            uint32_t any_e = 0, any_sc = 0;
            std::vector<uint32_t> sc_f(kGran);
            for (int gi = 0; gi < kGran; gi++) {
                uint32_t w[4], prev;
                memcpy(w, tile_in + gi * 16, 16);
                memcpy(&prev, tile_in + gi * 16 - 4, 4);
                const GranuleMasks mf = granule_masks_filtered(w, prev), mx = granule_masks(w, prev);
                if (mf.e != mx.e || mf.sc != mx.sc) n_filter_mismatch++;  // the filter may only skip empty granules
                any_e |= mf.e;
                any_sc |= mf.sc;
                sc_f[gi] = mf.sc;
                // the cheap part of the filter against its definition
                bool brute = false;
                for (int i = -1; i < 16; i++)
                    brute = brute || (tile_in[gi * 16 + i] == 0 && tile_in[gi * 16 + i - 1] == 0);
                if (brute != (acc_has_pair(zero_pair_acc(0xFFFFFFFFu, w, prev)) || (prev >> 16) == 0u)) n_filter_mismatch++;
            }
            const bool edge = pos == 0 || pos + kChunk + kHalo > n;
            // no emulation-prevention candidate: the copy kernel stores the chunk as it is and writes the records itself
            const bool filter_clean = any_e == 0 && !edge;
            // pieces longer than one chunk (not what the product runs) also carry the open NAL's EPB count along
            bool clean = filter_clean && carry_epb == 0;
            {   // ---- general path: exact masks (computed for every chunk here, to hold the filter against them)
                std::vector<uint32_t> em(kGran), ks(kGran), ee(kGran), incl(kGran);
                auto masks_at = [&](int gi, bool have_prev) {
                    uint32_t w[4];
                    memcpy(w, tile_in + gi * 16, 16);
                    uint32_t prev = 0xFFFFFFFFu;
                    if (have_prev) memcpy(&prev, tile_in + gi * 16 - 4, 4);
                    return granule_masks(w, prev);
                };
                for (int gi = 0; gi < kGran; gi++) {
                    GranuleMasks m = masks_at(gi, true);
                    em[gi] = m.e | (m.sc << 16);
                    scb[gi + 1] = (uint16_t)m.sc;
                }
                scb[0] = (uint16_t)masks_at(-1, false).sc;
                scb[kGran + 1] = (uint16_t)masks_at(kGran, true).sc;
                auto get = [&](int64_t p) -> uint32_t { return tile_in[p - pos]; };
                uint32_t rp[kRows + 1], cls[kRows];
                rp[0] = 0;
                for (int r = 0; r < kRows; r++) {
                    bool all_full = true, epb_only = true;
                    for (int lane = 0; lane < 32; lane++) {
                        const int gi = r * 32 + lane;
                        const int64_t gpos = pos + (int64_t)gi * 16;
                        uint32_t e16 = em[gi] & 0xFFFFu;
                        uint32_t k16 = ~e16 & 0xFFFFu;
                        const uint32_t near = ((uint32_t)scb[gi] >> 10) | scb[gi + 1] | (scb[gi + 2] & 1u);
                        if (near) k16 = keep_mask_near_sc(get, gpos, e16, scb[gi], scb[gi + 1], scb[gi + 2], &e16);
                        uint32_t sc = em[gi] >> 16;
                        if (pos + kChunk > n) {
                            if (gpos >= n) { k16 = 0; sc = 0; e16 = 0; }
                            else if (gpos + 16 > n) { uint32_t v = (1u << (uint32_t)(n - gpos)) - 1u; k16 &= v; sc &= v; e16 &= v; }
                        }
                        ks[gi] = k16 | (sc << 16);
                        ee[gi] = e16;
                        if (k16 != 0xFFFFu) all_full = false;
                        if (!((k16 | e16) == 0xFFFFu && sc == 0)) epb_only = false;
                    }
                    cls[r] = all_full ? 0 : (epb_only ? 1 : 2);
                    uint32_t run = 0;
                    for (int lane = 0; lane < 32; lane++) {
                        const int gi = r * 32 + lane;
                        const uint32_t el = cls[r] == 0 ? 0u : (cls[r] == 1 ? bits_popc(ee[gi]) : seg_element(ee[gi], ks[gi] >> 16));
                        run = lane ? seg_combine(run, el) : el;
                        incl[gi] = run;
                    }
                    rp[r + 1] = seg_combine(rp[r], run);
                }
                bool all0 = true;
                for (int r = 0; r < kRows; r++) all0 = all0 && cls[r] == 0;
                if (filter_clean) {  // the copy kernel's claim: nothing is removed in this chunk
                    for (int gi = 0; gi < kGran; gi++)
                        if (ee[gi]) n_filter_mismatch++;
                }
                if (clean) {
                    n_fast++;  // left to the copy kernel: verbatim copy (below) + records, all counts zero
                    uint32_t n_sc = 0;
                    for (int gi = 0; gi < kGran; gi++) n_sc += bits_popc(sc_f[gi]);
                    const uint64_t slot0 = rec_start.size();
                    rec_start.resize(slot0 + n_sc, 0);
                    rec_epb.resize(slot0 + n_sc, 0);
                    rec_hdr.resize(slot0 + n_sc, 0);
                    rec_rank.resize(slot0 + n_sc, 0);
                    uint32_t rank = 0;
                    for (int gi = 0; gi < kGran; gi++)
                        for (int j = 0; j < 16; j++)
                            if (sc_f[gi] & (1u << j)) {
                                const int64_t st = pos + gi * 16 + j + 1;
                                rec_start[slot0 + rank] = (uint64_t)st;
                                rec_hdr[slot0 + rank] = gets(st) | (gets(st + 1) << 8) | (gets(st + 2) << 16) | (gets(st + 3) << 24);
                                rec_rank[slot0 + rank] = pnsc + rank;
                                rank++;
                            }
                    pnsc += n_sc;
                } else if (all0 && carry_epb == 0) {
                    clean = true;
                    n_false_alarm++;
                } else {
                    n_general++;
                    const uint32_t total = rp[kRows];
                    const uint32_t n_sc = (total >> 16) & 0x1FFFu;
                    const uint64_t slot0 = rec_start.size();
                    rec_start.resize(slot0 + n_sc, 0);
                    rec_epb.resize(slot0 + n_sc, 0);
                    rec_hdr.resize(slot0 + n_sc, 0);
                    rec_rank.resize(slot0 + n_sc, 0);
                    // rows with boundaries first (bytes + records; header bytes come from the not yet compacted tile)
                    for (int r = 0; r < kRows; r++) {
                        if (cls[r] != 2) continue;
                        for (int lane = 0; lane < 32; lane++) {
                            const int gi = r * 32 + lane;
                            uint32_t w[4];
                            memcpy(w, tile_in + gi * 16, 16);
                            const uint32_t ex = lane ? incl[gi - 1] : 0u;
                            const uint32_t pre = seg_combine(rp[r], ex);
                            const uint64_t c = seg_apply(pre, carry_epb);
                            const uint32_t before = (pre >> 16) & 0x1FFFu;
                            const uint64_t gpos = (uint64_t)pos + (uint64_t)gi * 16;
                            const uint32_t k16 = ks[gi] & 0xFFFFu, sc = ks[gi] >> 16;
                            if (k16 == 0xFFFFu && sc == 0 && ((gpos - c) & 15u) == 0) {  // ordinary granule of such a row
                                memcpy(out + gpos - c, w, 16);
                                continue;
                            }
                            uint64_t k = slot0 + before;
                            uint32_t rank = pnsc + before;
                            store_granule_bytes(out, gpos, w, k16, ee[gi], sc, c, [&](int j, uint64_t c_end2) {
                                const uint64_t st = gpos + j + 1;
                                const uint8_t *hb = tile_in + gi * 16 + j + 1;
                                rec_start[k] = st;
                                rec_epb[k] = (uint32_t)c_end2;
                                rec_hdr[k] = (uint32_t)hb[0] | ((uint32_t)hb[1] << 8) | ((uint32_t)hb[2] << 16) |
                                             ((uint32_t)hb[3] << 24);
                                rec_rank[k] = rank;
                                k++;
                                rank++;
                            });
                        }
                    }
                    // in-place compaction of EPB-only rows
                    for (int r = 0; r < kRows; r++) {
                        if (cls[r] != 1) continue;
                        uint32_t w[32][4];
                        for (int lane = 0; lane < 32; lane++) memcpy(w[lane], tile_in + (r * 32 + lane) * 16, 16);
                        uint8_t *row = tile_in + 512 * r;
                        for (int lane = 0; lane < 32; lane++) {
                            const int gi = r * 32 + lane;
                            const uint32_t k16 = ks[gi] & 0xFFFFu;
                            uint32_t loff = 16u * lane - (incl[gi] - bits_popc(ee[gi]));
                            for (int j = 0; j < 16; j++)
                                if (k16 & (1u << j)) row[loff++] = (uint8_t)(w[lane][j >> 2] >> ((j & 3) * 8));
                        }
                    }
                    // the other rows
                    for (int r = 0; r < kRows; r++) {
                        if (cls[r] == 2) continue;
                        uint32_t w[32][4];
                        for (int lane = 0; lane < 32; lane++) memcpy(w[lane], tile_in + (512 * r + lane * 16), 16);
                        const uint64_t c_row = seg_apply(rp[r], carry_epb);
                        const uint32_t removed = cls[r] ? ((rp[r + 1] - rp[r]) & 0x7FFFu) : 0u;
                        const uint64_t o = (uint64_t)pos + 512u * r - c_row;
                        const uint8_t *prev_tail = nullptr;
                        if (r > 0 && cls[r - 1] != 2) prev_tail = tile_in + 512 * r - ((rp[r] - rp[r - 1]) & 0x7FFFu);
                        const bool next_joins = r < kRows - 1 && cls[r + 1] != 2;
                        for (int lane = 0; lane < 32; lane++)
                            store_row_lane(out, o, 512u - removed, lane ? w[lane - 1] : w[0], w[lane], lane, prev_tail, next_joins);
                    }
                    carry_epb = seg_apply(total, carry_epb);
                    pnsc += n_sc;
                }
            }
            if (clean) memcpy(out + pos, tile_in, kChunk);  // the TMA bulk store
        }
        piece_epb[piece] = carry_epb;
        piece_nsc[piece] = pnsc;
    }
    // post-pass 1 + 2: ordinals, permutation into stream order
    uint32_t run = 0;
    for (int64_t t = 0; t < n_pieces; t++) {
        piece_ord[t] = run;
        run += piece_nsc[t];
    }
    const int64_t Kall = (int64_t)rec_start.size();
    std::vector<uint32_t> epb_local((size_t)std::min(Kall, cap) + 1, 0);
    for (int64_t i = 0; i < Kall; i++) {
        const uint64_t ord = (uint64_t)piece_ord[(rec_start[i] - 1) / piece_bytes] + rec_rank[i];
        if ((int64_t)ord < cap) {
            nal_start[ord] = rec_start[i];
            epb_local[ord] = rec_epb[i];
            nal_hdr[ord] = rec_hdr[i];
        }
    }
    // post-pass 3 + 4: NALs that span pieces -- total their EPB counts, slide their later parts left
    std::vector<uint32_t> packed((size_t)n_pieces + 1, 0), S((size_t)n_pieces + 1, 0);
    uint32_t srun = 0;
    for (int64_t t = 0; t < n_pieces; t++) {
        packed[t] = (piece_nsc[t] << 16) | (piece_epb[t] & 0xFFFFu);
        S[t] = srun;
        srun += piece_epb[t];
    }
    const int64_t K = Kall < cap ? Kall : cap;
    for (int64_t k = 0; k < K; k++) nal_epb[k] = 0;
    for (int64_t k = 0; k + 1 < K; k++) {
        const uint32_t H = nal_header_bytes(nal_hdr[k] & 0xFF, (nal_hdr[k] >> 8) & 0xFF);
        uint32_t later;
        nal_epb[k + 1] = nal_removed(nal_start[k], nal_start[k + 1], epb_local[k + 1], S.data(), piece_bytes, &later);
        if (later)
            nal_pieces(nal_start[k], nal_start[k + 1], H, epb_local[k + 1], packed.data(), S.data(), piece_bytes,
                       [&](uint64_t ps, uint64_t len, uint64_t G) { memmove(out + ps - G, out + ps, len); });
    }
    if (stats) {
        stats[0] = n_fast;
        stats[1] = n_false_alarm;
        stats[2] = n_general;
        stats[3] = n_filter_mismatch;
    }
    return Kall;
}

// ---------------------------------------------------------------------------------------------------------------
// The round-2 pipeline of annexb_scan.cu, lane by lane: copy kernel (filter, verbatim chunks, dirty flags), ordered
// dirty list, per dirty chunk chunk_masks -> published counts -> lookback_carry -> chunk_store (output image built in
// place with word stores and spill hand-over, boundary granules byte by byte, store_image in two mappings), the
// segmented carry scan with its re-copy list, chunk_shift_kernel, permutation and nal_removed.  Dirty chunks are taken
// in ticket (= ascending) order, one at a time: the device's waits are synchronisation only, what is waited for is what
// is read here.  Same outputs as emul_stream_tiles; `out` must be 16-byte aligned like the device buffer when out_shift
// is a multiple of 16 (other shifts exercise nothing new here: stores are emulated byte-wise).
int64_t emul_stream_carry(const uint8_t *s, int64_t n, uint8_t *out_base, int64_t out_shift, uint64_t *nal_start,
                          uint64_t *nal_epb, uint32_t *nal_hdr, int64_t cap, int64_t *stats) {
    const int kRows = 4, kGran = 32 * kRows, kChunk = kGran * 16, kHalo = 16;
    const uint32_t kEpb = 0x7FFFu, kDirty = 0x8000u, kNscShift = 16, kNsc = 0x1FFFu, kReady = 0x80000000u;
    uint8_t *out = out_base + out_shift;
    auto gets = [&](int64_t p) -> uint32_t { return (p >= 0 && p < n) ? s[p] : 0xFFu; };
    const int64_t n_chunks = (n + kChunk - 1) / kChunk;
    std::vector<uint32_t> piece((size_t)n_chunks + 1, 0), piece_carry((size_t)n_chunks + 1, 0);
    std::vector<uint64_t> rec_start;
    std::vector<uint32_t> rec_epb, rec_hdr, rec_rank;
    int64_t n_verbatim = 0, n_dirty = 0, n_spin_would_wait = 0, n_recopied = 0;
    std::vector<uint8_t> buf(kHalo + kChunk + kHalo);
    std::vector<uint16_t> scb(kGran + 2);
    std::vector<int64_t> dirty_list;
    auto stage = [&](int64_t pos) {
        for (int i = 0; i < kHalo + kChunk + kHalo; i++) buf[i] = (uint8_t)gets(pos - kHalo + i);
    };
    // ---- annexb_copy_kernel
    for (int64_t chunk = 0; chunk < n_chunks; chunk++) {
        const int64_t pos = chunk * kChunk;
        if (pos == 0 || pos + kChunk + kHalo > n) {
            piece[chunk] = kDirty;
            dirty_list.push_back(chunk);
            continue;
        }
        stage(pos);
        const uint8_t *tile_in = buf.data() + kHalo;
        uint32_t any_e = 0;
        std::vector<uint32_t> sc_f(kGran);
        for (int gi = 0; gi < kGran; gi++) {
            uint32_t w[4], prev;
            memcpy(w, tile_in + gi * 16, 16);
            memcpy(&prev, tile_in + gi * 16 - 4, 4);
            const GranuleMasks mf = granule_masks_filtered(w, prev);
            any_e |= mf.e;
            sc_f[gi] = mf.sc;
        }
        if (any_e) {
            piece[chunk] = kDirty;
            dirty_list.push_back(chunk);
            continue;
        }
        n_verbatim++;
        memcpy(out + pos, tile_in, kChunk);
        uint32_t rank = 0;
        for (int gi = 0; gi < kGran; gi++)
            for (int j = 0; j < 16; j++)
                if (sc_f[gi] & (1u << j)) {
                    const uint64_t st = (uint64_t)pos + (uint64_t)gi * 16 + j + 1;
                    uint32_t h = 0;
                    for (int q = 0; q < 4; q++) h |= gets((int64_t)st + q) << (8 * q);
                    rec_start.push_back(st);
                    rec_hdr.push_back(h);
                    rec_epb.push_back(0);
                    rec_rank.push_back(rank++);
                }
        piece[chunk] = rank << kNscShift;
    }
    // ---- annexb_dirty_kernel, tickets in ascending order
    for (int64_t chunk : dirty_list) {
        n_dirty++;
        const int64_t pos = chunk * kChunk;
        stage(pos);
        uint8_t *tile_in = buf.data() + kHalo;
        // chunk_masks
        std::vector<uint32_t> es(kGran), pre(kGran);
        {
            std::vector<uint32_t> em(kGran);
            auto masks_at = [&](int gi, bool have_prev) {
                uint32_t w[4];
                memcpy(w, tile_in + gi * 16, 16);
                uint32_t prev = 0xFFFFFFFFu;
                if (have_prev) memcpy(&prev, tile_in + gi * 16 - 4, 4);
                return granule_masks(w, prev);
            };
            for (int gi = 0; gi < kGran; gi++) {
                const GranuleMasks m = masks_at(gi, true);
                em[gi] = m.e | (m.sc << 16);
                scb[gi + 1] = (uint16_t)m.sc;
            }
            scb[0] = (uint16_t)masks_at(-1, false).sc;
            scb[kGran + 1] = (uint16_t)masks_at(kGran, true).sc;
            auto get = [&](int64_t p) -> uint32_t { return tile_in[p - pos]; };
            uint32_t rp = 0;
            for (int r = 0; r < kRows; r++) {
                uint32_t run = 0;
                for (int lane = 0; lane < 32; lane++) {
                    const int gi = r * 32 + lane;
                    const int64_t gpos = pos + (int64_t)gi * 16;
                    uint32_t e16 = em[gi] & 0xFFFFu;
                    const uint32_t near = ((uint32_t)scb[gi] >> 10) | scb[gi + 1] | (scb[gi + 2] & 1u);
                    if (near) (void)keep_mask_near_sc(get, gpos, e16, scb[gi], scb[gi + 1], scb[gi + 2], &e16);
                    uint32_t sc = em[gi] >> 16;
                    if (pos + kChunk > n) {
                        if (gpos >= n) {
                            sc = 0;
                            e16 = 0;
                        } else if (gpos + 16 > n) {
                            const uint32_t v = (1u << (uint32_t)(n - gpos)) - 1u;
                            sc &= v;
                            e16 &= v;
                        }
                    }
                    es[gi] = e16 | (sc << 16);
                    pre[gi] = seg_combine(rp, lane ? run : 0u);
                    const uint32_t el = seg_element(e16, sc);
                    run = lane ? seg_combine(run, el) : el;
                }
                rp = seg_combine(rp, run);
            }
            // publish
            const uint32_t total = rp;
            piece[chunk] = kReady | kDirty | (total & 0x1FFF0000u) | (total & kEpb);
            // lookback_carry
            uint32_t carry = 0;
            bool done = false;
            for (int64_t base = chunk; base > 0 && !done; base -= 32) {
                uint32_t w[32], c[32];
                uint32_t tmask = 0;
                for (int lane = 0; lane < 32; lane++) {
                    const int64_t idx = base - 1 - lane;
                    bool term;
                    if (idx < 0) {
                        w[lane] = 0;
                        c[lane] = 0;
                        term = true;
                    } else {
                        w[lane] = piece[idx];
                        c[lane] = piece_carry[idx];
                        if ((w[lane] & kDirty) && !(w[lane] & kReady)) n_spin_would_wait++;  // (cannot happen in ticket order)
                        term = (((w[lane] >> kNscShift) & kNsc) != 0) || (c[lane] & kReady);
                    }
                    if (term) tmask |= 1u << lane;
                }
                const int first = tmask ? __builtin_ctz(tmask) : 32;
                for (int lane = 0; lane < 32 && lane <= first; lane++) {
                    const int64_t idx = base - 1 - lane;
                    if (lane < first) carry += w[lane] & kEpb;
                    else carry += idx < 0 ? 0u : (((w[lane] >> kNscShift) & kNsc) ? (w[lane] & kEpb) : (c[lane] & ~kReady));
                }
                if (tmask) done = true;
            }
            piece_carry[chunk] = kReady | seg_apply(total, carry);

            // ---- chunk_store
            const uint32_t n_sc = (total >> 16) & 0x1FFFu;
            const uint32_t cb = carry & 15u;
            const int64_t base_a = pos - (int64_t)(carry & ~15u);
            const int64_t limit = n - pos;
            auto store_image = [&](int64_t base, int x0, int x1) {  // (aligned stores and ragged ends alike: bytes [x0, x1))
                for (int x = x0; x < x1; x++) out[base + x] = tile_in[x];
            };
            if ((total & 0x7FFFu) == 0 && n_sc == 0 && cb == 0) {
                store_image(base_a, 0, (int)(limit < kChunk ? limit : kChunk));
                continue;
            }
            int first_end = 0, first_a_end = 0;
            if (n_sc) {
                const uint64_t slot0 = rec_start.size();
                rec_start.resize(slot0 + n_sc);
                rec_hdr.resize(slot0 + n_sc);
                rec_epb.resize(slot0 + n_sc);
                rec_rank.resize(slot0 + n_sc);
                bool have_first = false;
                for (int gi = 0; gi < kGran; gi++) {
                    uint32_t sc = es[gi] >> 16;
                    if (!sc) continue;
                    const uint32_t ee = es[gi] & 0xFFFFu;
                    if (((pre[gi] >> 16) & 0x1FFFu) == 0 && !have_first) {
                        const int j = __builtin_ctz(sc);
                        first_end = gi * 16 + j + 1;
                        first_a_end = first_end - 2 - (int)cb - (int)((pre[gi] & 0x7FFFu) + bits_popc(ee & ((1u << j) - 1u)));
                        have_first = true;
                    }
                    // emit_dirty_records
                    uint32_t rank = (pre[gi] >> 16) & 0x1FFFu, c = pre[gi] & 0x7FFFu;
                    int prev = -1;
                    while (sc) {
                        const int j = __builtin_ctz(sc);
                        sc &= sc - 1;
                        const uint32_t between = (ee >> (prev + 1)) & ((1u << (j - prev - 1)) - 1u);
                        c += bits_popc(between);
                        const uint64_t st = (uint64_t)pos + (uint64_t)gi * 16 + (uint64_t)j + 1;
                        const uint8_t *hb = tile_in + gi * 16 + j + 1;
                        rec_start[slot0 + rank] = st;
                        rec_hdr[slot0 + rank] = (uint32_t)hb[0] | ((uint32_t)hb[1] << 8) | ((uint32_t)hb[2] << 16) | ((uint32_t)hb[3] << 24);
                        rec_epb[slot0 + rank] = c;
                        rec_rank[slot0 + rank] = rank;
                        rank++;
                        c = 0;
                        prev = j;
                    }
                }
            }
            // every lane's granules into registers before anything is rewritten
            std::vector<uint32_t> v((size_t)kGran * 4);
            memcpy(v.data(), tile_in, kChunk);
            auto put_word = [&](int wq, uint32_t x) { memcpy(tile_in + 4 * wq, &x, 4); };
            std::vector<uint32_t> late_word(kGran);
            std::vector<int> late_x(kGran), late_n(kGran);
            uint32_t spill_prev = 0;
            bool reg_prev = false;
            for (int gi = 0; gi < kGran; gi++) {
                const int lane = gi & 31, r = gi >> 5;
                uint32_t ee = es[gi] & 0xFFFFu;
                const uint32_t sc = es[gi] >> 16;
                const bool regular = sc == 0;
                const int shift = (int)(pre[gi] & 0x7FFFu) + ((pre[gi] >> 31) ? 0 : (int)cb);
                const int q = gi * 16 - shift;
                const int nb = 16 - (int)bits_popc(ee);
                uint32_t w0 = v[gi * 4], w1 = v[gi * 4 + 1], w2 = v[gi * 4 + 2], w3 = v[gi * 4 + 3];
                if (regular) {
                    while (ee) {
                        const int j = bits_msb(ee);
                        ee &= ~(1u << j);
                        const uint32_t s0 = funnel_r(w0, w1, 8), s1 = funnel_r(w1, w2, 8), s2 = funnel_r(w2, w3, 8), s3 = w3 >> 8;
                        const uint32_t m = (1u << ((j & 3) * 8)) - 1u;
                        const int wj = j >> 2;
                        w0 = wj > 0 ? w0 : (w0 & m) | (s0 & ~m);
                        w1 = wj > 1 ? w1 : (wj == 1 ? (w1 & m) | (s1 & ~m) : s1);
                        w2 = wj > 2 ? w2 : (wj == 2 ? (w2 & m) | (s2 & ~m) : s2);
                        w3 = wj == 3 ? (w3 & m) | (s3 & ~m) : s3;
                    }
                }
                const uint32_t a8 = (uint32_t)(q & 3) * 8u;
                auto fl = [&](uint32_t lo, uint32_t hi) { return a8 ? ((hi << a8) | (lo >> (32 - a8))) : hi; };  // __funnelshift_l
                const uint32_t x0 = w0 << a8, x1 = fl(w0, w1), x2 = fl(w1, w2), x3 = fl(w2, w3), x4 = fl(w3, 0u);
                const int wq = q >> 2;
                const int cnt = ((q + nb) >> 2) - wq;
                const int rem = (q + nb) & 3;
                const uint32_t spill = cnt == 2 ? x2 : (cnt == 3 ? x3 : x4);
                const uint32_t sp = spill_prev;
                const bool rg = gi == 0 ? false : reg_prev;
                spill_prev = spill;
                reg_prev = regular;
                const bool next_regular = gi + 1 < kGran && (es[gi + 1] >> 16) == 0;
                (void)lane;
                (void)r;
                if (regular) {
                    put_word(wq, x0 | (rg ? sp : 0u));
                    put_word(wq + 1, x1);
                    if (cnt > 2) put_word(wq + 2, x2);
                    if (cnt > 3) put_word(wq + 3, x3);
                }
                late_word[gi] = spill;
                late_x[gi] = (wq + cnt) * 4;
                late_n[gi] = regular ? (next_regular ? 0 : rem) : -100;
            }
            for (int gi = 0; gi < kGran; gi++) {
                if (late_n[gi] == -100) {  // image_boundary_granule
                    uint32_t c = (pre[gi] & 0x7FFFu) + ((pre[gi] >> 31) ? 0u : cb);
                    const uint32_t ee = es[gi] & 0xFFFFu, sc = es[gi] >> 16;
                    const uint32_t vv[4] = {v[gi * 4], v[gi * 4 + 1], v[gi * 4 + 2], v[gi * 4 + 3]};
                    for (int j = 0; j < 16; j++) {
                        const uint32_t bit = 1u << j;
                        if (ee & bit) c++;
                        else tile_in[gi * 16 + j - (int)c] = (uint8_t)granule_byte(vv, j);
                        if (sc & bit) c = 0;
                    }
                } else {
                    for (int k = 0; k < late_n[gi]; k++) tile_in[late_x[gi] + k] = (uint8_t)(late_word[gi] >> (8 * k));
                }
            }
            const int removed_end = (int)(total & 0x7FFFu);
            if (n_sc == 0) {
                int x1 = kChunk - (int)cb - removed_end;
                const int64_t lim = limit - (int64_t)cb;
                if ((int64_t)x1 > lim) x1 = (int)lim;
                store_image(base_a, -(int)cb, x1);
            } else {
                int xb1 = kChunk - removed_end;
                if ((int64_t)xb1 > limit) xb1 = (int)limit;
                if (base_a == pos) {
                    store_image(pos, -(int)cb, xb1);
                } else {
                    store_image(base_a, -(int)cb, first_a_end);
                    store_image(pos, first_end, xb1);
                }
            }
        }
    }
    // ---- order_reduce / order_apply: ordinals, S[], segmented carry -> re-copy list
    std::vector<uint32_t> piece_ord((size_t)n_chunks + 1, 0), S((size_t)n_chunks + 1, 0);
    struct Shift { int64_t t; uint32_t G; };
    std::vector<Shift> shift_list;
    {
        uint32_t nsc_run = 0, epb_run = 0, seg_f = 0, seg_c = 0;
        for (int64_t t = 0; t < n_chunks; t++) {
            piece_ord[t] = nsc_run;
            S[t] = epb_run;
            if (!(piece[t] & kDirty) && nsc_run && seg_c) shift_list.push_back({t, seg_c});
            const uint32_t nsc = (piece[t] >> kNscShift) & kNsc, epb = piece[t] & kEpb;
            nsc_run += nsc;
            epb_run += epb;
            if (nsc) seg_f = 1, seg_c = epb; else seg_c += epb;
        }
        (void)seg_f;
        S[n_chunks] = epb_run;
    }
    // ---- nal_permute_kernel
    const int64_t Kall = (int64_t)rec_start.size();
    std::vector<uint32_t> epb_local((size_t)std::min(Kall, cap) + 1, 0);
    for (int64_t i = 0; i < Kall; i++) {
        const uint64_t ord = (uint64_t)piece_ord[(rec_start[i] - 1) / kChunk] + rec_rank[i];
        if ((int64_t)ord < cap) {
            nal_start[ord] = rec_start[i];
            epb_local[ord] = rec_epb[i];
            nal_hdr[ord] = rec_hdr[i];
        }
    }
    const int64_t K = Kall < cap ? Kall : cap;
    // ---- chunk_shift_kernel (reads the INPUT)
    for (const Shift &e : shift_list) {
        const int64_t t = e.t, k = (int64_t)piece_ord[t] - 1;
        if (k + 1 >= cap) continue;
        const int64_t lo = t * kChunk;
        int64_t ps = lo, pe = lo + kChunk;
        const int64_t body = (int64_t)nal_start[k] + nal_header_bytes(nal_hdr[k] & 0xFFu, (nal_hdr[k] >> 8) & 0xFFu);
        if (body > ps) ps = body;
        if ((piece[t] >> kNscShift) & kNsc) pe = (int64_t)nal_start[k + 1] - 2;
        if (pe > ps) {
            n_recopied++;
            for (int64_t p = ps; p < pe; p++) out[p - e.G] = s[p];
        }
    }
    // ---- scan_finalize_kernel
    for (int64_t k = 0; k < K; k++) nal_epb[k] = 0;
    for (int64_t k = 0; k + 1 < K; k++) {
        uint32_t later;
        nal_epb[k + 1] = nal_removed(nal_start[k], nal_start[k + 1], epb_local[k + 1], S.data(), (uint64_t)kChunk, &later);
    }
    if (stats) {
        stats[0] = n_verbatim;
        stats[1] = n_dirty;
        stats[2] = n_recopied;
        stats[3] = n_spin_would_wait;
    }
    return Kall;
}

// the slice-header walk of slice_header_kernel, one slice
void emul_slice_header(const h264b_param_sets *ps, uint32_t nal_type, uint32_t nal_ref_idc, const uint8_t *rbsp,
                       uint64_t len, h264b_slice_header *out) {
    parse_slice_header_record(*ps, nal_type, nal_ref_idc, rbsp, len, out);
}

// NewSPS / NewPPS as parse_sps_kernel / parse_pps_kernel run them, one parameter set
void emul_parse_sps(const uint8_t *rbsp, uint64_t len, h264b_sps *out) {
    memset(out, 0, sizeof(*out));
    out->status = parse_sps(rbsp, len, out);
}
void emul_parse_pps(const uint8_t *rbsp, uint64_t len, h264b_pps *out) {
    memset(out, 0, sizeof(*out));
    out->status = parse_pps(rbsp, len, out);
}
void emul_make_param_sets(const h264b_sps *s, const h264b_pps *p, h264b_param_sets *out) { *out = make_param_sets(*s, *p); }

// the syntax-element glue as ctx_glue_kernel evaluates it
int64_t emul_ctx_idx(int64_t bin_idx, int64_t max_bin_idx_ctx, int64_t off) { return ctx_idx_ref(bin_idx, max_bin_idx_ctx, off); }
void emul_new_binarization(int32_t se, int32_t st, h264b_binarization *out) { *out = new_binarization_ref(se, st); }
void emul_mb_bin_string(int32_t st, int64_t mb_type, int32_t sub, int32_t *len, uint32_t *bits) {
    mb_bin_string_ref(st, mb_type, sub != 0, len, bits);
}
int32_t emul_bin_string_match(int32_t len, uint32_t bin, int32_t n, uint32_t bits) { return bin_string_match_ref(len, bin, n, bits); }

// NewNalUnit on one frame with keep_byte_frame; returns rbsp length
int64_t emul_frame(const uint8_t *f, int64_t N, uint8_t *rbsp) {
    auto get = [&](int64_t p) -> uint32_t { return (p >= 0 && p < N) ? f[p] : 0xFFu; };
    const uint32_t H = nal_header_bytes(get(0), get(1));
    int64_t k = 0;
    for (int64_t p = 0; p < N; p++)
        if (keep_byte_frame(get, 0, N, H, p)) rbsp[k++] = f[p];
    return k;
}

static void build_tab(int spec, uint64_t *tab) {
    const uint8_t *range_lps = spec ? h264b_range_tab_lps_spec : h264b_range_tab_lps_ref;
    const uint8_t *tl = spec ? h264b_trans_idx_lps_spec : h264b_trans_idx_lps_ref;
    const uint8_t *tm = spec ? h264b_trans_idx_mps_spec : h264b_trans_idx_mps_ref;
    for (uint32_t s = 0; s < 128; s++) {  // same packing as cabac_table_kernel (ctx_init.cu)
        const uint32_t p = s & 63, v = s >> 6;
        const uint32_t lo = range_lps[p * 4] | (range_lps[p * 4 + 1] << 8) | (range_lps[p * 4 + 2] << 16) |
                            ((uint32_t)range_lps[p * 4 + 3] << 24);
        const uint32_t next_mps = tm[p] | (v << 6);
        const uint32_t v_lps = (p == 0) ? 1 - v : v;
        const uint32_t next_lps = tl[p] | (v_lps << 6);
        const uint32_t hi = next_mps | (v << 8) | (next_lps << 16) | ((1 - v) << 24);
        tab[s] = ((uint64_t)hi << 32) | lo;
    }
}

struct EmulFinal {
    int64_t R, O;
    uint64_t bits_read;
    uint32_t overrun, n_bins;
};

// One lane of cabac_decode_kernel: flags bit0 = SPEC tables, bit1 = SPEC_OR bypass, bit2 = final terminate,
// bit3 = also refill at random eligible moments (what happens when another lane of the warp triggers the refill)
void emul_cabac(const uint8_t *buf, uint64_t total, uint64_t off, uint32_t len, const uint16_t *ops, uint32_t n_ops,
                uint8_t *states, uint32_t n_ctx, uint32_t flags, uint32_t *bins, EmulFinal *fin) {
    uint64_t tab[128];
    build_tab(flags & 1, tab);
    uint32_t word = 0, n_bins = n_ops;
    LaneDecoder eng = {};
    eng.init(buf, total, off, (flags & 2u) != 0);
    uint32_t lcg = 12345u + n_ops;
    for (uint32_t i = 0; i < n_ops; i++) {
        if (eng.must_refill()) eng.refill_if_room();
        if (flags & 8u) {  // a neighbouring lane triggered the warp-wide refill: take bits early whenever there is room
            lcg = lcg * 1664525u + 1013904223u;
            if ((lcg >> 28) < 5u) eng.refill_if_room();
        }
        const uint32_t kind = ops[i] >> 14;
        uint32_t bin;
        if (kind == 0) {
            uint32_t c = ops[i] & 0x3FFu;
            if (c >= n_ctx) c = 0;
            uint8_t ns;
            bin = eng.decision(tab[states[c] & 127u], &ns);
            states[c] = ns;
        } else if (kind == 1) {
            bin = eng.bypass();
        } else {
            bin = eng.terminate();
        }
        word |= bin << (i & 31u);
        if ((i & 31u) == 31u) {
            bins[i >> 5] = word;
            word = 0;
        }
    }
    if (flags & 4u) {
        if (eng.must_refill()) eng.refill_if_room();
        word |= eng.terminate() << (n_ops & 31u);
        n_bins++;
    }
    if (n_bins & 31u)
        bins[n_bins >> 5] = word;
    else if ((flags & 4u))
        bins[(n_bins - 1) >> 5] = word;
    fin->R = eng.cod_i_range();
    fin->O = eng.cod_i_offset();
    fin->bits_read = eng.bits_read();
    fin->overrun = fin->bits_read > 8ull * len;
    fin->n_bins = n_bins;
}

}  // extern "C"
