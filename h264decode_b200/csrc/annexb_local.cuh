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
#include <stddef.h>
#include <stdint.h>

#ifndef H264B_HD
#if defined(__CUDACC__)
#define H264B_HD __host__ __device__ __forceinline__
#else
#define H264B_HD static inline
#endif
#endif

namespace h264b {

#if defined(__CUDA_ARCH__)
H264B_HD uint32_t bits_popc(uint32_t x) { return (uint32_t)__popc(x); }
H264B_HD int bits_msb(uint32_t x) { return 31 - __clz((int)x); }  // x != 0
#else
H264B_HD uint32_t bits_popc(uint32_t x) { return (uint32_t)__builtin_popcount(x); }
H264B_HD int bits_msb(uint32_t x) { return 31 - __builtin_clz(x); }
#endif

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

// byte b (0..15) of a 16-byte granule held in four words, without dynamic register indexing
H264B_HD uint32_t granule_byte(const uint32_t y[4], int b) {
    const uint32_t lo = (b & 4) ? y[1] : y[0], hi = (b & 4) ? y[3] : y[2];
    return (((b & 8) ? hi : lo) >> ((b & 3) * 8)) & 0xFFu;
}

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
    uint32_t cand = two & 0xFFFFu;  // positions with two zero bytes right before them: rare for entropy-coded payloads
    if (cand && (cand & (cand - 1u)) == 0u) {  // a single candidate (the usual case): look at that byte alone
        const int j = bits_msb(cand);
        const uint32_t b = granule_byte(w, j);
        if (b == 3u) m.e = cand;
        if (b == 1u) m.sc = cand & three;
    } else if (cand) {  // several (runs of zeros): byte-parallel search for 03 / 01
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

// Cheap filter in front of granule_masks: does some byte p of a span have s[p] == 0 && s[p-1] == 0?  Every
// emulation-prevention byte and every start-code end at p, p+1 or p+2 needs such a pair, so a span of the stream
// without one is copied verbatim (the fast path of annexb_scan_kernel).  Two adjacent zero bytes are a zero 16-bit
// half of either the word itself or the word shifted by one byte, so a running per-halfword minimum over both views
// (one VIMNMX3.U16x2 per word on sm_100a) ends with a zero half iff the span holds a pair.
H264B_HD uint32_t funnel_l8(uint32_t lo, uint32_t hi) { return (hi << 8) | (lo >> 24); }
H264B_HD uint32_t min3_u16x2(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    return __vimin3_u16x2(a, b, c);
#else
    uint32_t lo = a & 0xFFFFu, hi = a >> 16;
    if ((b & 0xFFFFu) < lo) lo = b & 0xFFFFu;
    if ((c & 0xFFFFu) < lo) lo = c & 0xFFFFu;
    if ((b >> 16) < hi) hi = b >> 16;
    if ((c >> 16) < hi) hi = c >> 16;
    return lo | (hi << 16);
#endif
}
// acc: running minimum (start with 0xFFFFFFFF); w: a 16-byte granule; prev: the 4 bytes before it
H264B_HD uint32_t zero_pair_acc(uint32_t acc, const uint32_t w[4], uint32_t prev) {
    acc = min3_u16x2(acc, w[0], funnel_l8(prev, w[0]));
    acc = min3_u16x2(acc, w[1], funnel_l8(w[0], w[1]));
    acc = min3_u16x2(acc, w[2], funnel_l8(w[1], w[2]));
    acc = min3_u16x2(acc, w[3], funnel_l8(w[2], w[3]));
    return acc;
}
// the same for the last 7 positions of the 8 bytes (lo, hi) that precede a chunk: a start code ending there still
// reaches into the chunk with its NAL header and its emulation-prevention guard
H264B_HD uint32_t zero_pair_acc_tail8(uint32_t acc, uint32_t lo, uint32_t hi) {
    // pairs (b-8,b-7) .. (b-2,b-1): halves of lo, of hi, and of the 4 bytes b-7 .. b-4 (the pair (b-9,b-8) is out of reach:
    // the byte shifted in below is made non-zero)
    acc = min3_u16x2(acc, lo, (lo << 8) | 0xFFu);
    return min3_u16x2(acc, hi, funnel_l8(lo, hi));
}
H264B_HD bool acc_has_pair(uint32_t acc) { return (acc & 0xFFFFu) == 0u || (acc >> 16) == 0u; }

// Two-level filter of the copy kernel, per granule: the exact masks, but computed only when the cheap test fires.
// An emulation-prevention candidate or a start-code end at p needs two zero bytes right before it, i.e. a pair ending
// at g-1 .. g+14: the cheap test looks for pairs ending at g-1 (the last two bytes of prev) .. g+15.
H264B_HD GranuleMasks granule_masks_filtered(const uint32_t w[4], uint32_t prev) {
    if (!acc_has_pair(zero_pair_acc(0xFFFFFFFFu, w, prev)) && (prev >> 16) != 0u) {
        GranuleMasks m;
        m.z = m.e = m.sc = 0;
        return m;
    }
    return granule_masks(w, prev);
}

// keep mask of a granule at stream position gpos when start codes end within [gpos-6, gpos+16], in the bit domain
// (same result as 16 x keep_byte_stream, checked exhaustively on the CPU by tests/test_hd_logic.py):
//   e16      raw emulation-prevention mask of the granule (granule_masks().e)
//   sc_prev / sc_own / sc_next   start-code-end masks of the previous, this and the next granule
//   get(p)   stream byte accessor, used only for the (at most two) header bytes of each NAL that starts in reach
//   *epb_eff receives the emulation-prevention bytes that are really removed (raw candidates minus those whose zeros
//   belong to a NAL header)
template <class Get>
H264B_HD uint32_t keep_mask_near_sc(const Get& get, int64_t gpos, uint32_t e16, uint32_t sc_prev, uint32_t sc_own,
                                    uint32_t sc_next, uint32_t *epb_eff) {
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
    *epb_eff = e16 & ~n16 & ~d16;
    return ~(d16 | (e16 & ~n16)) & 0xFFFFu;
}

// ---- segmented EPB count ("how far has this NAL's RBSP shifted left so far") ----------------------------------
// The RBSP of a NAL is written at the NAL's own position in the output buffer: a kept byte at stream position p goes
// to out[p - c(p)], c(p) = emulation-prevention bytes removed in p's NAL before p.  c restarts at every NAL start, so
// it is a SEGMENTED running count; an element of the scan is packed in 32 bits:
//   bit 31      the span contains a NAL start (the count below is then "since the last start in the span")
//   bits 28:16  start codes in the span (for NAL numbering; <= 512 per 2 KiB chunk)
//   bits 14:0   EPBs removed (<= 683 per chunk)
H264B_HD uint32_t seg_combine(uint32_t a, uint32_t b) {  // a = earlier span, b = later span
    const uint32_t val = ((b >> 31) ? 0u : (a & 0x7FFFu)) + (b & 0x7FFFu);
    return ((a | b) & 0x80000000u) | ((a + b) & 0x1FFF0000u) | val;
}
// element of one granule: ee = effective EPB mask, sc = start-code-end mask
H264B_HD uint32_t seg_element(uint32_t ee, uint32_t sc) {
    if (!sc) return bits_popc(ee);
    const int last = bits_msb(sc);  // bit of the last start-code end; the new NAL starts just above it
    const uint32_t above = last >= 15 ? 0u : (ee >> (last + 1));
    return 0x80000000u | (bits_popc(sc) << 16) | bits_popc(above);
}
// count carried into a span whose exclusive prefix (inside the tile) is `pre`, given the tile's carry-in
H264B_HD uint32_t seg_apply(uint32_t pre, uint32_t carry_in) { return (pre >> 31) ? (pre & 0x7FFFu) : carry_in + (pre & 0x7FFFu); }

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

// Store one row of a tile (one lane's part; the kernel calls this for all 32 lanes of the row's warp): `len`
// (<= 512) contiguous bytes, lane l holding row bytes [16l, 16l+16) in w, to out[o .. o+len).
//   * lanes exchange neighbours' words (wp = previous lane's granule) and each writes one ALIGNED 16-byte granule;
//     when the row has not shifted (o is 16-byte aligned, the usual case) that is a plain aligned copy;
//   * prev_tail != nullptr: the previous row is contiguous with this one in the output and its bytes end at
//     prev_tail (in the tile buffer): lane 0 completes the granule straddling the seam from there.  Otherwise this
//     row's part of that granule is written byte by byte;
//   * next_joins: the next row will complete the last, ragged granule; otherwise it is written byte by byte here.
H264B_HD void store_row_lane(uint8_t *out, uint64_t o, uint32_t len, const uint32_t wp[4], const uint32_t w[4],
                             int lane, const uint8_t *prev_tail, bool next_joins) {
    const uint32_t sb = (uint32_t)o & 15u;  // warp-uniform
    uint32_t y[4];
    if (sb == 0) {
#pragma unroll
        for (int k = 0; k < 4; k++) y[k] = w[k];
    } else {
        uint32_t x[8];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            x[k] = wp[k];  // previous lane's granule (row bytes 16l-16 .. 16l-1); by shuffle on the device
            x[4 + k] = w[k];
        }
        // this lane's output granule = row bytes [16l - sb, 16l - sb + 16) = bytes [16 - sb, 32 - sb) of x
        const uint32_t off = 16u - sb, br8 = (off & 3u) * 8u;
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
            default:
#pragma unroll
                for (int k = 0; k < 4; k++) y[k] = funnel_r(x[3 + k], x[4 + k], br8);
                break;
        }
    }
    const int lo_b = 16 * lane - (int)sb;  // first row byte of this lane's output granule
    uint8_t *dst = out + (o - sb) + 16u * (uint32_t)lane;
    if (lo_b >= 0 && lo_b + 16 <= (int)len) {
        store16(dst, y);
    } else if (lane == 0 && sb != 0) {  // seam granule (len >= 16 always holds for these rows)
        if (prev_tail) {
            const uint32_t addr = (uint32_t)(uintptr_t)(prev_tail - sb);  // only its low 2 bits matter below
            const uint32_t *p = reinterpret_cast<const uint32_t *>((uintptr_t)(prev_tail - sb) & ~(uintptr_t)3);
            const uint32_t s8 = (addr & 3u) * 8u;
            const uint32_t q0 = p[0], q1 = p[1], q2 = p[2], q3 = p[3], q4 = p[4];
            const uint32_t z[4] = {funnel_r(q0, q1, s8), funnel_r(q1, q2, s8), funnel_r(q2, q3, s8),
                                   funnel_r(q3, q4, s8)};
            uint32_t r4[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {  // bytes < sb from the previous row, the rest from this one
                const int nb = (int)sb - 4 * k;
                const uint32_t m = nb >= 4 ? 0xFFFFFFFFu : (nb <= 0 ? 0u : ((1u << (8 * nb)) - 1u));
                r4[k] = (z[k] & m) | (y[k] & ~m);
            }
            store16(dst, r4);
        } else {
            for (int b = (int)sb; b < 16; b++) dst[b] = (uint8_t)granule_byte(y, b);
        }
    }
    if (!next_joins && ((sb + len) & 15u) != 0) {  // last ragged granule, nobody else will complete it
        const uint32_t jt = (sb + len) >> 4, nb = (sb + len) & 15u;  // granule index in the row, valid bytes in it
        if (jt < 32u) {
            if ((uint32_t)lane == jt) {
                const int b0 = lo_b < 0 ? -lo_b : 0;
                for (int b = b0; b < (int)nb; b++) dst[b] = (uint8_t)granule_byte(y, b);
            }
        } else if (lane == 31) {  // 33rd granule: the last bytes of lane 31's data
            uint8_t *d2 = out + o + 496;
            for (uint32_t b = 16u - sb; b < 16u && 496u + b < len; b++) d2[b] = (uint8_t)granule_byte(w, (int)b);
        }
    }
}

// A row that contains NAL boundaries (or stream ends): one lane writes its granule's kept bytes one by one.
//   c      EPB count of the open NAL at the granule's first byte
//   k16 / ee / sc   keep, effective-EPB and start-code-end masks of the granule
// on_start(j, c_end) is called for every start-code end at granule byte j with the EPB count of the NAL it ends.
template <class OnStart>
H264B_HD void store_granule_bytes(uint8_t *out, uint64_t gpos, const uint32_t w[4], uint32_t k16, uint32_t ee,
                                  uint32_t sc, uint64_t c, const OnStart &on_start) {
    for (int j = 0; j < 16; j++) {
        const uint32_t bit = 1u << j;
        if (ee & bit) {
            c++;
        } else if (k16 & bit) {
            out[gpos + (uint64_t)j - c] = (uint8_t)granule_byte(w, j);
        }
        if (sc & bit) {
            on_start(j, c);
            c = 0;
        }
    }
}

// ---- NALs whose body spans several chunks ---------------------------------------------------------------------
// Inside a chunk ("piece": kChunk bytes of the stream) a kept byte at stream position p goes to
// out[p - (EPBs removed from p's NAL earlier in this chunk) - G], G = what the NAL lost in the chunks before.  The
// device kernels carry G into the chunk (look-back over the per-chunk counts, annexb_scan.cu).  The CPU emulation of
// the kernels' logic (tests/native/hd_emul.cpp) keeps the round-1 formulation as a second, independent one: every
// chunk compacted on its own (G = 0), then the later parts of a NAL slide left by their G (nal_pieces below).  Both
// must produce the oracle's bytes.  nal_removed is shared: the device's scan_finalize_kernel totals a NAL's count
// with it.
//   tail[t]   per chunk, low 16 bits: EPBs after the chunk's last NAL start, or in the whole chunk when it holds none
//   S[t]      exclusive prefix sum of tail[] (mod 2^32): for chunks Tq < t <= Tb of one NAL, G(t) = S[t] - S[Tq]
//   a, b      first byte of this NAL / of the next one (b-1 is the 01 of the start code that ends it)
//   end_local EPB count the main pass recorded at that start code: EPBs since the NAL's start if it began in the
//             same chunk, else since the start of the chunk
// EPBs removed from the whole NAL [a, b); *later_shift = G of its last part (non-zero: some part may have to move)
H264B_HD uint64_t nal_removed(uint64_t a, uint64_t b, uint64_t end_local, const uint32_t *S, uint64_t piece_bytes,
                              uint32_t *later_shift) {
    const uint64_t Tq = (a - 1) / piece_bytes, Tb = (b - 1) / piece_bytes;
    const uint32_t G = Tq == Tb ? 0u : S[Tb] - S[Tq];
    *later_shift = G;
    return (uint64_t)G + end_local;
}
// move(start, len, G) is called, in stream order, for every run of later parts that has to slide left by G > 0
// (parts that lost nothing themselves are contiguous with their successor and share its G: they move as one run).
// The walk may be cut into windows of parts [t_begin, t_end) (the GPU stages tail[] / S[] of a window in shared memory:
// both are indexed [t - t0]); `run` carries the open run from one window to the next, nal_pieces_flush ends the walk.
struct MoveRun {
    uint64_t ps, len, G;
};
template <class Move>
H264B_HD void nal_pieces_window(uint64_t a, uint64_t b, uint32_t H, uint64_t end_local, const uint32_t *tail,
                                const uint32_t *S, uint64_t t0, uint32_t S_Tq, uint64_t piece_bytes, uint64_t t_begin,
                                uint64_t t_end, MoveRun &run, const Move &move) {
    const uint64_t Tb = (b - 1) / piece_bytes;
    for (uint64_t t = t_begin; t < t_end; t++) {
        const uint64_t G = (uint32_t)(S[t - t0] - S_Tq);
        if (!G) continue;
        const uint64_t lo = t * piece_bytes, body = a + H;
        const uint64_t ps = lo > body ? lo : body;                    // first kept byte of the part
        const uint64_t pe = t < Tb ? (t + 1) * piece_bytes : b - 2;   // kept bytes are < pe (b-2, b-1: the tail rule)
        const uint64_t e = t < Tb ? (uint64_t)(tail[t - t0] & 0xFFFFu) : end_local;
        if (!(pe > ps && pe - ps > e)) continue;
        const uint64_t len = pe - ps - e;
        if (run.len && G == run.G && ps == run.ps + run.len) {
            run.len += len;
        } else {
            if (run.len) move(run.ps, run.len, run.G);
            run.ps = ps;
            run.len = len;
            run.G = G;
        }
    }
}
template <class Move>
H264B_HD void nal_pieces_flush(MoveRun &run, const Move &move) {
    if (run.len) move(run.ps, run.len, run.G);
    run.len = 0;
}
// the whole walk in one go (tail[] / S[] indexed by chunk)
template <class Move>
H264B_HD void nal_pieces(uint64_t a, uint64_t b, uint32_t H, uint64_t end_local, const uint32_t *tail, const uint32_t *S,
                         uint64_t piece_bytes, const Move &move) {
    const uint64_t Tq = (a - 1) / piece_bytes, Tb = (b - 1) / piece_bytes;
    MoveRun run = {0, 0, 0};
    nal_pieces_window(a, b, H, end_local, tail, S, 0, S[Tq], piece_bytes, Tq + 1, Tb + 1, run, move);
    nal_pieces_flush(run, move);
}

}  // namespace h264b
