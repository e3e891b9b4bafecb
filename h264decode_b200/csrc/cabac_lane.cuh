// cabac_lane.cuh -- per-lane CABAC arithmetic decoding (one slice per lane), __host__ __device__ so that tests/
// can run the exact same arithmetic on the CPU against the oracle.
//
// Reference functions (h264/cabac.go): initDecodingEngine :439-446, BinaryDecision core :525-536,
// StateTransitionProcess :544-553, RenormD :503-511, DecodeBypass :468-481, DecodeTerminate :486-499.
//
// Representation.  The reference keeps codIOffset (O, 9 bits once O < codIRange holds) and pulls stream bits in
// one at a time.  A lane instead keeps a 64-bit window  hi:lo  whose top 10 bits are O and whose next `fbits` bits
// are the not-yet-consumed stream bits that follow:
//        bit 63 ....... 54 | 53 .................... 54-fbits | rest zero
//             O (10 bits)  |  next fbits bits of the stream   |
//   * O >= R          <=>  hi >= (R << 22)          (the fraction bits cannot change the outcome)
//   * O -= R          <=>  hi -= (R << 22)
//   * RenormD by k    <=>  window <<= k, fbits -= k  (k = clz(R) - 23: R<256 doubles until >= 256; k <= 7 since R >= 2)
//   * bypass           =   window <<= 1 first ((O<<1)|bit is exactly that), O may then need the 10th bit
// 32 fresh bits are OR-ed in below the fraction whenever fbits <= 22; a lane MUST refill before an op when fbits < 8.
// bitsRead of the reference (9 + sum of renorm shifts + bypass bins) = 32 * (1 + refills) - fbits.
//
// The window is valid while O < R at op boundaries, which conformant slices guarantee under the SPEC_OR bypass form
// (H264B_BYPASS_SPEC_OR).  Everything else -- the reference's literal bypass (O <<= 1; O <<= bit, A5), a first
// 9-bit value >= 510, ops after a DecodeTerminate that returned 1 -- runs on LiteralLane (see LaneDecoder).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#ifndef H264B_HD
#define H264B_HD __host__ __device__ __forceinline__
#endif
#define H264B_HDM __host__ __device__ __forceinline__
#else
#ifndef H264B_HD
#define H264B_HD static inline
#endif
#define H264B_HDM inline
#endif

namespace h264b {

#if defined(__CUDA_ARCH__)
H264B_HD uint32_t funnel_l(uint32_t lo, uint32_t hi, uint32_t s) { return __funnelshift_l(lo, hi, s); }
H264B_HD uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }
H264B_HD uint32_t clz32(uint32_t x) { return (uint32_t)__clz((int)x); }
#else
H264B_HD uint32_t funnel_l(uint32_t lo, uint32_t hi, uint32_t s) {
    s &= 31u;
    return s ? ((hi << s) | (lo >> (32u - s))) : hi;
}
H264B_HD uint32_t bswap32(uint32_t x) { return (x >> 24) | ((x >> 8) & 0xFF00u) | ((x << 8) & 0xFF0000u) | (x << 24); }
H264B_HD uint32_t clz32(uint32_t x) { return x ? (uint32_t)__builtin_clz(x) : 32u; }
#endif

// Sequential reader of big-endian 32-bit groups from an arbitrarily aligned byte position, using aligned word loads
// clamped to [words, last_word] (bytes past the slice are don't-care: they can never influence a comparison).
struct BitFeed {
    const uint32_t *next;       // aligned word to load next
    const uint32_t *last_word;  // last readable aligned word of the buffer
    uint32_t cur;               // big-endian value of the previous aligned word
    uint32_t pf;                // prefetched raw word at `next`
    uint32_t mis8;              // 8 * (byte offset & 3)

    H264B_HDM uint32_t load(const uint32_t *p) const { return *(p <= last_word ? p : last_word); }
    H264B_HDM void init(const uint8_t *buf, uint64_t total_bytes, uint64_t off) {
        const uint32_t *words = reinterpret_cast<const uint32_t *>(buf);  // buf is at least 4-byte aligned
        last_word = words + ((total_bytes ? total_bytes - 1 : 0) >> 2);
        next = words + (off >> 2);
        mis8 = (uint32_t)(off & 3u) * 8u;
        cur = bswap32(load(next));
        next++;
        pf = load(next);
    }
    H264B_HDM uint32_t get32() {
        const uint32_t nxt = bswap32(pf);
        const uint32_t v = funnel_l(nxt, cur, mis8);
        cur = nxt;
        next++;
        pf = load(next);
        return v;
    }
};

struct CabacLane {
    uint32_t R;      // codIRange
    uint32_t hi, lo; // window
    int32_t fbits;
    uint32_t refills;
    BitFeed feed;

    // initDecodingEngine, cabac.go:439-446: codIRange = 510, codIOffset = next 9 bits
    H264B_HDM void init(const uint8_t *buf, uint64_t total_bytes, uint64_t off) {
        feed.init(buf, total_bytes, off);
        const uint32_t v0 = feed.get32();
        R = 510u;
        hi = v0 >> 1;
        lo = v0 << 31;
        fbits = 23;
        refills = 0;
    }
    H264B_HDM bool can_refill() const { return fbits <= 22; }
    H264B_HDM bool must_refill() const { return fbits < 8; }
    H264B_HDM void refill() {
        const uint32_t v = feed.get32();
        const uint32_t s = (uint32_t)(22 - fbits);
        hi |= funnel_l(v, 0u, s);  // v >> (32 - s), 0 for s == 0
        lo |= v << s;
        fbits += 32;
        refills++;
    }
    H264B_HDM void shift(uint32_t k) {
        hi = funnel_l(lo, hi, k);
        lo <<= k;
        fbits -= (int32_t)k;
    }
    // "DecodeDecision": cabac.go:525-536 -> :544-553 -> :503-511.  tab_entry is the 64-bit engine table entry for
    // the current state byte (see ctx_init.cu: rangeLPS x4 | nextMPS | binMPS | nextLPS | binLPS).
    H264B_HDM uint32_t decision(uint64_t tab_entry, uint8_t *state_out) {
        const uint32_t tlo = (uint32_t)tab_entry, thi = (uint32_t)(tab_entry >> 32);
        const uint32_t q = (R >> 6) & 3u;
        const uint32_t lps = (tlo >> (q * 8u)) & 0xFFu;
        const uint32_t rm = R - lps;
        const uint32_t x = rm << 22;
        const bool is_lps = hi >= x;
        if (is_lps) hi -= x;
        R = is_lps ? lps : rm;
        const uint32_t sel = is_lps ? (thi >> 16) : thi;
        *state_out = (uint8_t)(sel & 0xFFu);
        const uint32_t bin = (sel >> 8) & 1u;
        const uint32_t k = clz32(R) - 23u;  // R in [2, 510]
        R <<= k;
        shift(k);
        return bin;
    }
    // DecodeBypass, SPEC_OR form of cabac.go:468-481
    H264B_HDM uint32_t bypass() {
        shift(1);
        const uint32_t x = R << 22;
        const bool one = hi >= x;
        if (one) hi -= x;
        return one ? 1u : 0u;
    }
    // DecodeTerminate, cabac.go:486-499 (no renormalisation when the bin is 1)
    H264B_HDM uint32_t terminate() {
        R -= 2u;
        const uint32_t x = R << 22;
        if (hi >= x) return 1u;
        const uint32_t k = clz32(R) - 23u;
        R <<= k;
        shift(k);
        return 0u;
    }
    H264B_HDM int64_t cod_i_offset() const { return (int64_t)(hi >> 22); }
    H264B_HDM uint64_t bits_read() const { return 32ull * (1ull + refills) - (uint64_t)fbits; }
};

// Literal 64-bit engine: codIOffset is a Go int (int64) that may exceed codIRange and wraps silently; compares are
// signed; one bit at a time, like the reference.  Used
//   * for the reference's own bypass form (H264B_BYPASS_SPEC_OR clear: O <<= 1; O <<= bit, cabac.go:470-473, A5),
//     under which O grows without bound, and
//   * as the continuation of a window lane whose invariant O < R broke: the first 9 bits were >= 510, or a
//     DecodeTerminate returned 1 (cabac.go:488-493 leaves O >= R) and the schedule keeps going.  Conformant slices
//     never do either, but the primitives are defined there and parity is bit-exact on every input.
struct LiteralLane {
    int64_t R, O;
    uint64_t bitpos;  // bits consumed from the slice start
    // the stream bits that follow, left-aligned in `win` (`avail` of them, <= 64), behind them the slice's bit feed
    uint64_t win;
    uint32_t avail;
    BitFeed feed;
    bool spec_or;

    // next k bits of the slice (k <= 32), MSB first; bytes past the buffer repeat its last word (don't-care: the slice
    // is flagged H264B_F_OVERRUN long before)
    // kRefill = false: the caller has made sure that avail >= k (cabac_decode_kernel's block loop tops the window up by
    // a warp vote every two ops)
    H264B_HDM void top_up() {  // avail <= 32: room for 32 more
        win |= (uint64_t)feed.get32() << (32u - avail);
        avail += 32u;
    }
    template <bool kRefill = true>
    H264B_HDM uint32_t read_bits(uint32_t k) {
        if (kRefill && avail < k) top_up();
        const uint32_t v = k ? (uint32_t)(win >> (64u - k)) : 0u;
        win = k ? win << k : win;
        avail -= k;
        bitpos += k;
        return v;
    }
    H264B_HDM void attach(const uint8_t *, uint64_t, uint64_t, bool spec_or_bypass) { spec_or = spec_or_bypass; }
    H264B_HDM void init(const uint8_t *b, uint64_t total, uint64_t o) {  // initDecodingEngine, cabac.go:439-446
        feed.init(b, total, o);
        win = 0;
        avail = 0;
        bitpos = 0;
        R = 510;
        O = (int64_t)read_bits(9);
    }
    template <bool kRefill = true>
    H264B_HDM void renorm() {  // RenormD, cabac.go:503-511: R doubles until >= 256, O takes one stream bit per step
        // k = the number of doublings: 0 for R >= 256, clz32(R) - 23 for R in [1, 255] (9 - bit length), and the cap of 9
        // steps for R <= 0 (a state only stream garbage reaches; the reference would spin there).  No loop: a chain of
        // compare-and-branch steps costs a lone warp tens of cycles each.
        const uint32_t rc = R >= 256 ? 256u : (R <= 0 ? 0u : (uint32_t)R);  // clz32(256) = 23, clz32(0) = 32
        const uint32_t k = clz32(rc) - 23u;
        R = (int64_t)((uint64_t)R << k);
        O = (int64_t)(((uint64_t)O << k) | (uint64_t)read_bits<kRefill>(k));
    }
    H264B_HDM uint32_t decision(uint64_t tab_entry, uint8_t *state_out) {
        const uint32_t tlo = (uint32_t)tab_entry, thi = (uint32_t)(tab_entry >> 32);
        const uint32_t q = (uint32_t)(R >> 6) & 3u;
        const int64_t lps = (tlo >> (q * 8u)) & 0xFFu;
        R -= lps;
        uint32_t sel;
        if (O >= R) {
            O = (int64_t)((uint64_t)O - (uint64_t)R);
            R = lps;
            sel = thi >> 16;
        } else {
            sel = thi;
        }
        *state_out = (uint8_t)(sel & 0xFFu);
        renorm();
        return (sel >> 8) & 1u;
    }
    template <bool kRefill = true>
    H264B_HDM uint32_t bypass() {  // cabac.go:468-481
        uint64_t o = (uint64_t)O << 1;
        const uint32_t b = read_bits<kRefill>(1);
        o = spec_or ? (o | b) : (o << b);
        O = (int64_t)o;
        if (O >= R) {
            O = (int64_t)((uint64_t)O - (uint64_t)R);
            return 1u;
        }
        return 0u;
    }
    template <bool kRefill = true>
    H264B_HDM uint32_t terminate() {  // cabac.go:486-499
        R -= 2;
        if (O >= R) return 1u;
        renorm<kRefill>();
        return 0u;
    }
};

// What a lane of cabac_decode_kernel runs: the window engine while O < R is guaranteed, the literal engine otherwise.
struct LaneDecoder {
    CabacLane w;
    LiteralLane l;
    bool lit;

    H264B_HDM void to_literal() {  // the window engine's look-ahead bits and bit feed go on where they are
        l.R = (int64_t)w.R;
        l.O = w.cod_i_offset();
        l.bitpos = w.bits_read();
        l.win = ((((uint64_t)w.hi << 32) | w.lo) << 10);
        l.avail = w.fbits > 0 ? (uint32_t)w.fbits : 0u;
        l.feed = w.feed;
        lit = true;
    }
    H264B_HDM void init(const uint8_t *buf, uint64_t total_bytes, uint64_t off, bool spec_or_bypass) {
        l.attach(buf, total_bytes, off, spec_or_bypass);
        if (!spec_or_bypass) {
            lit = true;
            l.init(buf, total_bytes, off);
            return;
        }
        lit = false;
        w.init(buf, total_bytes, off);
        if (w.cod_i_offset() >= 510) to_literal();
    }
    H264B_HDM bool must_refill() const { return !lit && w.must_refill(); }
    H264B_HDM void refill_if_room() {
        if (!lit && w.can_refill()) w.refill();
    }
    H264B_HDM uint32_t decision(uint64_t tab_entry, uint8_t *state_out) {
        return lit ? l.decision(tab_entry, state_out) : w.decision(tab_entry, state_out);
    }
    H264B_HDM uint32_t bypass() { return lit ? l.bypass() : w.bypass(); }
    H264B_HDM uint32_t terminate() {
        if (lit) return l.terminate();
        const uint32_t bin = w.terminate();
        if (bin) to_literal();  // O >= R from here on
        return bin;
    }
    H264B_HDM int64_t cod_i_range() const { return lit ? l.R : (int64_t)w.R; }
    H264B_HDM int64_t cod_i_offset() const { return lit ? l.O : w.cod_i_offset(); }
    H264B_HDM uint64_t bits_read() const { return lit ? l.bitpos : w.bits_read(); }
};

}  // namespace h264b
