// annexb_local.cuh -- position-local form of the reference's sequential Annex-B split + RBSP strip.
//
// The reference walks the stream one byte at a time (readNalUnit, h264/server.go:64-111) and then walks each NAL
// again (NewNalUnit body loop, h264/nalUnit.go:106-126).  Both walks are equivalent to per-byte predicates that
// look at most 9 bytes back and 1 byte ahead, which is what makes the path data-parallel:
//
//   SC(p)    <=> s[p-3..p] == 00 00 00 01                       (isStartSequence, server.go:28-39; 4-byte only, A8)
//   a        =  p+1 for SC(p): first byte of a NAL; the NAL runs up to and including the NEXT start code (A8)
//   H(a)     =  1, or 4 for types 14/20, or for type 21: 3 if s[a+1]&0x80 else 4        (nalUnit.go:79,86-103)
//   EPB(p)   <=> s[p]==3 && s[p-1]==0 && s[p-2]==0 && p-2 >= a+H                         (nalUnit.go:32-37,113; A7)
//   keep(p)  <=> a <= p exists && p >= a+H && !SC(p) && !SC(p+1) && !EPB(p)
//               (!SC(p) && !SC(p+1): the Peek failure at nalUnit.go:107-111 drops the NAL's last two bytes, which in
//                stream framing are the 00 01 that end the next start code)
//
// Everything here is __host__ __device__ so that tests/ can run the exact same predicates on the CPU against the
// oracle's sequential restatement.
#pragma once
#include <stdint.h>

#ifndef H264B_HD
#if defined(__CUDACC__)
#define H264B_HD __host__ __device__ __forceinline__
#else
#define H264B_HD static inline
#endif
#endif

namespace h264b {

// NAL header length from the first two NAL bytes (nalUnit.go:79,86-103)
H264B_HD uint32_t nal_header_bytes(uint32_t b0, uint32_t b1) {
    uint32_t t = b0 & 31u;
    if (t == 14u || t == 20u) return 4u;  // SVC (flag 1) and MVC (flag 0) extensions are both 3 bytes
    if (t == 21u) return (b1 & 0x80u) ? 3u : 4u;  // 3D-AVC ext is 2 bytes, MVC 3
    return 1u;
}

// Byte accessor over a window with out-of-range positions reading as 0xFF (matches nothing).
template <class Get>
H264B_HD bool is_sc_end(const Get& get, int64_t p) {  // SC(p): p is the 01 of 00 00 00 01
    return get(p) == 1u && get(p - 1) == 0u && get(p - 2) == 0u && get(p - 3) == 0u;
}

// keep(p) in stream mode for a byte known to lie at or after the first NAL start.
// get(q) must return s[q] for every q in [p-9, p+2] that lies inside the stream and 0xFF otherwise.
template <class Get>
H264B_HD bool keep_byte_stream(const Get& get, int64_t p) {
    if (is_sc_end(get, p) || is_sc_end(get, p + 1)) return false;  // positions b-1 and b-2
    bool epb_ok = true;
#pragma unroll
    for (int d = 0; d < 6; d++) {  // most recent NAL start a = p-d, if within reach of its header / EPB guard
        if (is_sc_end(get, p - d - 1)) {
            uint32_t H = nal_header_bytes(get(p - d), get(p - d + 1));
            if ((uint32_t)d < H) return false;  // header byte
            epb_ok = (uint32_t)d >= H + 2u;     // both zeros of a 00 00 03 must be body bytes (A7)
            break;
        }
    }
    if (epb_ok && get(p) == 3u && get(p - 1) == 0u && get(p - 2) == 0u) return false;  // emulation prevention
    return true;
}

// keep(p) for NewNalUnit called directly on one frame [a, a+N): no start codes involved; the body is
// [a+H, a+N-3], plus the byte a+N-2 when the frame ends in an emulation-prevention triple (the match at cursor
// N-3 copies both zeros, nalUnit.go:113-117).
// is p the 03 of an emulation-prevention triple that NewNalUnit removes from frame [a, a+N)?
template <class Get>
H264B_HD bool is_epb_frame(const Get& get, int64_t a, int64_t N, uint32_t H, int64_t p) {
    return p - a - 2 >= (int64_t)H && p - a < N && get(p) == 3u && get(p - 1) == 0u && get(p - 2) == 0u;
}
template <class Get>
H264B_HD bool keep_byte_frame(const Get& get, int64_t a, int64_t N, uint32_t H, int64_t p) {
    int64_t rel = p - a;
    if (rel < (int64_t)H) return false;
    if (rel <= N - 3) return !is_epb_frame(get, a, N, H, p);
    if (rel == N - 2) return is_epb_frame(get, a, N, H, p + 1);
    return false;
}

// ---- word-parallel detection used by the kernel's fast path ------------------------------------------------
// bit 7 of each byte of the result is set iff that byte of w is zero (exact, no cross-byte carries)
H264B_HD uint32_t zero_bytes(uint32_t w) { return ~(((w & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | w | 0x7F7F7F7Fu); }
// gather bit 7 of each byte into bits 0..3 (byte 0 -> bit 0)
H264B_HD uint32_t pack_msb4(uint32_t m) { return ((m & 0x80808080u) * 0x00204081u) >> 28; }

// For a 16-byte granule (little-endian words w[0..3], byte j of the granule = byte j&3 of w[j>>2]) and the 4 bytes
// before it (prev, byte 3 = the byte just before the granule), return per-byte bit masks (bit j = granule byte j):
//   z  : byte == 0          e : raw emulation-prevention candidate  s[p]==3 && s[p-1]==0 && s[p-2]==0
//   sc : start-code end     s[p]==1 && s[p-1]==s[p-2]==s[p-3]==0
struct GranuleMasks {
    uint32_t z, e, sc;
};
H264B_HD GranuleMasks granule_masks(const uint32_t w[4], uint32_t prev) {
    uint32_t z = pack_msb4(zero_bytes(w[0])) | (pack_msb4(zero_bytes(w[1])) << 4) |
                 (pack_msb4(zero_bytes(w[2])) << 8) | (pack_msb4(zero_bytes(w[3])) << 12);
    uint32_t zp = pack_msb4(zero_bytes(prev));  // bits 0..3 = bytes g-4..g-1
    uint32_t zz = (z << 4) | zp;                // bit k <-> position g-4+k, k in [0,20)
    // two / three zeros immediately before position g+j  <=> zz bits (j+3, j+2) / (j+3, j+2, j+1)
    uint32_t two = (zz >> 3) & (zz >> 2);
    uint32_t three = two & (zz >> 1);
    GranuleMasks m;
    m.z = z;
    m.e = 0;
    m.sc = 0;
    if ((two & 0xFFFFu) != 0) {  // rare for entropy-coded payloads: only now look for 03 / 01 bytes
        uint32_t t = pack_msb4(zero_bytes(w[0] ^ 0x03030303u)) | (pack_msb4(zero_bytes(w[1] ^ 0x03030303u)) << 4) |
                     (pack_msb4(zero_bytes(w[2] ^ 0x03030303u)) << 8) |
                     (pack_msb4(zero_bytes(w[3] ^ 0x03030303u)) << 12);
        uint32_t o = pack_msb4(zero_bytes(w[0] ^ 0x01010101u)) | (pack_msb4(zero_bytes(w[1] ^ 0x01010101u)) << 4) |
                     (pack_msb4(zero_bytes(w[2] ^ 0x01010101u)) << 8) |
                     (pack_msb4(zero_bytes(w[3] ^ 0x01010101u)) << 12);
        m.e = t & two & 0xFFFFu;
        m.sc = o & three & 0xFFFFu;
    }
    return m;
}

// keep mask of a granule at stream position gpos when start codes end within [gpos-6, gpos+16], in the bit domain
// (same result as 16 x keep_byte_stream, checked exhaustively on the CPU by tests/test_hd_logic.py):
//   e16      raw emulation-prevention mask of the granule (granule_masks().e)
//   sc_prev / sc_own / sc_next   start-code-end masks of the previous, this and the next granule
//   get(p)   stream byte accessor, used only for the (at most two) header bytes of each NAL that starts in reach
template <class Get>
H264B_HD uint32_t keep_mask_near_sc(const Get& get, int64_t gpos, uint32_t e16, uint32_t sc_prev, uint32_t sc_own,
                                    uint32_t sc_next) {
    // bit (16 + j) of these 64-bit masks <-> stream position gpos + j, j in [-16, 32)
    const uint64_t S = (uint64_t)(sc_prev & 0xFFFFu) | ((uint64_t)(sc_own & 0xFFFFu) << 16) |
                       ((uint64_t)(sc_next & 0xFFFFu) << 32);
    uint64_t drop = S | (S >> 1);  // q and q-1: the NAL's last two bytes are never copied (nalUnit.go:107-111)
    uint64_t no_epb = 0;
    uint64_t cand = S & 0x7FFFFC00ull;  // start-code ends q in [gpos-6, gpos+14]: NAL starts that reach this granule
    while (cand) {
#if defined(__CUDA_ARCH__)
        const int b = __ffsll((long long)cand) - 1;
#else
        const int b = __builtin_ctzll(cand);
#endif
        cand &= cand - 1;
        const int64_t a = gpos + (b - 16) + 1;  // first byte of the NAL
        const uint32_t H = nal_header_bytes(get(a), get(a + 1));
        drop |= ((1ull << H) - 1ull) << (b + 1);           // header bytes a .. a+H-1
        no_epb |= ((1ull << (H + 2u)) - 1ull) << (b + 1);  // a 03 at a .. a+H+1 has a header byte among its zeros (A7)
    }
    const uint32_t d16 = (uint32_t)(drop >> 16) & 0xFFFFu, n16 = (uint32_t)(no_epb >> 16) & 0xFFFFu;
    return ~(d16 | (e16 & ~n16)) & 0xFFFFu;
}

#if defined(__CUDA_ARCH__)
H264B_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s) { return __funnelshift_r(lo, hi, s); }
H264B_HD void store16(uint8_t *dst, const uint32_t v[4]) {  // dst is 16-byte aligned
    *reinterpret_cast<uint4 *>(dst) = make_uint4(v[0], v[1], v[2], v[3]);
}
#else
H264B_HD void store16(uint8_t *dst, const uint32_t v[4]) {
    for (int b = 0; b < 16; b++) dst[b] = (uint8_t)(v[b >> 2] >> ((b & 3) * 8));
}
H264B_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s) {
    s &= 31u;
    return s ? ((lo >> s) | (hi << (32u - s))) : lo;
}
#endif

// byte b (0..15) of a 16-byte granule held in four words, without dynamic register indexing
H264B_HD uint32_t granule_byte(const uint32_t y[4], int b) {
    const uint32_t lo = (b & 4) ? y[1] : y[0], hi = (b & 4) ? y[3] : y[2];
    return (((b & 8) ? hi : lo) >> ((b & 3) * 8)) & 0xFFu;
}

// Store row t of a tile (one lane's part; the kernel calls this for all 32 lanes of the row's warp) -- `len` (<= 512) contiguous bytes, lane l holding row bytes [16l, 16l+16) in w -- to
// out[o .. o+len).  After the in-place compaction every row of the tile is a contiguous run in shared memory (row t at
// tile_in + 512 t), and rows follow each other without gaps in the output, so:
//   * lanes exchange neighbours' words by shuffle and each writes one ALIGNED 16-byte granule of the destination;
//   * the granule that straddles the seam with the previous row is written once, by this row's lane 0, which fetches
//     the previous row's last bytes from shared memory;
//   * only the first ragged granule of a tile (its other bytes belong to the previous tile) and the last one (next
//     tile) are written byte by byte.
// rowoff[0..31] = exclusive kept-byte offsets of the tile's rows (low 16 bits), x_row = rowoff[t], K = tile total.
H264B_HD void store_row_lane(uint8_t *out, uint64_t o, uint32_t len, const uint32_t wp[4], const uint32_t w[4],
                             int lane, int t, uint32_t x_row, uint32_t K, const uint8_t *tile_in,
                             const uint32_t *rowoff) {
    const uint32_t sb = (uint32_t)o & 15u;  // warp-uniform
    uint32_t x[8];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        x[k] = wp[k];  // previous lane's granule (row bytes 16l-16 .. 16l-1); by shuffle on the device
        x[4 + k] = w[k];
    }
    // this lane's output granule = row bytes [16l - sb, 16l - sb + 16) = bytes [16 - sb, 32 - sb) of x
    const uint32_t off = 16u - sb, br8 = (off & 3u) * 8u;
    uint32_t y[4];
    switch (off >> 2) {  // warp-uniform
        case 0:
#pragma unroll
            for (int k = 0; k < 4; k++) y[k] = funnel_r(x[k], x[k + 1], br8);
            break;
        case 1:
#pragma unroll
            for (int k = 0; k < 4; k++) y[k] = funnel_r(x[1 + k], x[2 + k], br8);
            break;
        case 2:
#pragma unroll
            for (int k = 0; k < 4; k++) y[k] = funnel_r(x[2 + k], x[3 + k], br8);
            break;
        case 3:
#pragma unroll
            for (int k = 0; k < 4; k++) y[k] = funnel_r(x[3 + k], x[4 + k], br8);
            break;
        default:  // sb == 0: already aligned
#pragma unroll
            for (int k = 0; k < 4; k++) y[k] = w[k];
            break;
    }
    const int lo_b = 16 * lane - (int)sb;  // first row byte of this lane's output granule
    uint8_t *dst = out + (o - sb) + 16u * (uint32_t)lane;
    if (lo_b >= 0 && lo_b + 16 <= (int)len) {
        store16(dst, y);
    } else if (lane == 0 && sb != 0 && len >= 16u - sb) {
        // seam granule: its first sb bytes are the sb output bytes before this row.  The ones that belong to this tile
        // (tile-output coordinates x_row - have .. x_row - 1) are fetched from shared memory; any others are the
        // previous tile's and are written by it.
        const uint32_t have = x_row < sb ? x_row : sb;  // how many of the sb bytes this tile holds
        const uint32_t plen = t > 0 ? ((x_row - rowoff[t - 1]) & 0xFFFFu) : 0u;
        uint32_t z[4] = {0u, 0u, 0u, 0u};
        if (have == sb && plen >= sb) {  // usual case: all of them are the tail of row t-1
            const uint32_t addr = 512u * (uint32_t)(t - 1) + plen - sb;  // byte offset in the tile buffer
            const uint32_t *p = reinterpret_cast<const uint32_t *>(tile_in + (addr & ~3u));
            const uint32_t s8 = (addr & 3u) * 8u;
            const uint32_t q0 = p[0], q1 = p[1], q2 = p[2], q3 = p[3], q4 = p[4];
            z[0] = funnel_r(q0, q1, s8);
            z[1] = funnel_r(q1, q2, s8);
            z[2] = funnel_r(q2, q3, s8);
            z[3] = funnel_r(q3, q4, s8);
        } else {  // rare: tiny rows in between and / or the tile's first bytes; walk back byte by byte
            int tt = t - 1;
            for (int b = (int)sb - 1; b >= (int)(sb - have); b--) {
                const uint32_t xo = x_row - sb + (uint32_t)b;  // tile-output coordinate of this byte
                while (tt > 0 && (rowoff[tt] & 0xFFFFu) > xo) tt--;
                const uint32_t v = tile_in[512 * tt + (int)(xo - (rowoff[tt] & 0xFFFFu))];
                const uint32_t sh = (uint32_t)(b & 3) * 8u;
                if (b < 4) z[0] |= v << sh; else if (b < 8) z[1] |= v << sh;
                else if (b < 12) z[2] |= v << sh; else z[3] |= v << sh;
            }
        }
        uint32_t r4[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {  // bytes < sb from z, the rest from y
            const int nb = (int)sb - 4 * k;
            const uint32_t m = nb >= 4 ? 0xFFFFFFFFu : (nb <= 0 ? 0u : ((1u << (8 * nb)) - 1u));
            r4[k] = (z[k] & m) | (y[k] & ~m);
        }
        if (have == sb) {
            store16(dst, r4);
        } else {  // first ragged granule of the tile
            for (int b = (int)(sb - have); b < 16; b++) dst[b] = (uint8_t)granule_byte(r4, b);
        }
    }
    // the last ragged granule of the TILE (nothing after this row in the tile): bytes only
    if (x_row + len == K && ((sb + len) & 15u) != 0) {
        const uint32_t jt = (sb + len) >> 4, nb = (sb + len) & 15u;  // granule index in the row, valid bytes in it
        if (jt < 32u) {
            if ((uint32_t)lane == jt) {
                const int b0 = lo_b < 0 ? -lo_b : 0;
                for (int b = b0; b < (int)nb; b++) dst[b] = (uint8_t)granule_byte(y, b);
            }
        } else if (lane == 31) {  // 33rd granule: the last sb bytes of lane 31's data
            uint8_t *d2 = out + o + 496;
            for (uint32_t b = 16u - sb; b < 16u && 496u + b < len; b++) d2[b] = (uint8_t)granule_byte(w, (int)b);
        }
    }
}

}  // namespace h264b
