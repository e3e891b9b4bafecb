/* harness_core.h -- synthetic-data harness shared by the CPU build (harness.c, gcc) and the GPU generator
 * (harness_gpu.cu, nvcc).  TEST / BENCH INPUT GENERATION ONLY: it is not part of the product and not part of
 * the oracle.  Nothing here comes from the reference: the reference has no encoder.  The arithmetic encoder is
 * the informative encoder of ITU-T H.264 9.3.4, parameterised on the same tables as the
 * decoder under test so that encode -> decode round-trips (with the SPEC_OR bypass form).
 *
 * RNG and input shapes follow SURVEY.md §8(d): splitmix64, seed = 0x4832363400000000 + config*0x1000 + id.
 */
#ifndef HARNESS_CORE_H
#define HARNESS_CORE_H
#include <stdint.h>

#if defined(__CUDACC__)
#define HZ_HD __host__ __device__ __forceinline__
#else
#define HZ_HD static inline
#endif

#define HZ_SEED_BASE 0x4832363400000000ull

/* op encoding shared with the product ABI (include/h264b200.h): kind in bits 14..15, ctxIdx in bits 0..9 */
#define HZ_OP_DECISION 0u
#define HZ_OP_BYPASS 1u
#define HZ_OP_TERMINATE 2u

typedef struct {
    uint64_t s;
} hz_rng;

HZ_HD uint64_t hz_next(hz_rng *r) {
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
HZ_HD hz_rng hz_seed(uint32_t config, uint64_t id) {
    hz_rng r;
    r.s = HZ_SEED_BASE + (uint64_t)config * 0x1000ull + id;
    return r;
}

/* ---- tables the encoder needs (pointers so the GPU build can point them at shared memory) */
typedef struct {
    const uint8_t *range_lps; /* [64*4] */
    const uint8_t *trans_lps; /* [64]   */
    const uint8_t *trans_mps; /* [64]   */
} hz_tables;

/* ---- byte sink with optional emulation-prevention escaping (insert 03 before a byte <= 3 that follows 00 00) */
typedef struct {
    uint8_t *out;
    int64_t cap, n;
    uint32_t acc;   /* pending bits, MSB-first */
    int nacc;
    int zeros;      /* run of emitted zero bytes (escaped stream) */
    int escape;
    int overflow;
} hz_sink;

HZ_HD void hz_sink_init(hz_sink *s, uint8_t *out, int64_t cap, int escape) {
    s->out = out;
    s->cap = cap;
    s->n = 0;
    s->acc = 0;
    s->nacc = 0;
    s->zeros = 0;
    s->escape = escape;
    s->overflow = 0;
}
HZ_HD void hz_put_raw(hz_sink *s, uint8_t b) {
    if (s->n < s->cap)
        s->out[s->n] = b;
    else
        s->overflow = 1;
    s->n++;
}
HZ_HD void hz_put_byte(hz_sink *s, uint8_t b) {
    if (s->escape) {
        if (s->zeros >= 2 && b <= 3) {
            hz_put_raw(s, 3);
            s->zeros = 0;
        }
        s->zeros = (b == 0) ? s->zeros + 1 : 0;
    }
    hz_put_raw(s, b);
}
HZ_HD void hz_put_bit_raw(hz_sink *s, uint32_t bit) {
    s->acc = (s->acc << 1) | (bit & 1u);
    if (++s->nacc == 8) {
        hz_put_byte(s, (uint8_t)s->acc);
        s->acc = 0;
        s->nacc = 0;
    }
}
/* pad the last partial byte with zero bits; with escaping on, a payload ending in 00 gets a trailing 03 */
HZ_HD void hz_sink_finish(hz_sink *s) {
    while (s->nacc != 0) hz_put_bit_raw(s, 0);
    if (s->escape && s->zeros > 0) {
        hz_put_raw(s, 3);
        s->zeros = 0;
    }
}

/* ---- arithmetic encoder, H.264 9.3.4.x */
typedef struct {
    uint32_t low, range;
    uint32_t first_bit_flag, bits_outstanding;
    hz_sink *sink;
} hz_enc;

HZ_HD void hz_enc_init(hz_enc *e, hz_sink *sink) {
    e->low = 0;
    e->range = 510;
    e->first_bit_flag = 1;
    e->bits_outstanding = 0;
    e->sink = sink;
}
HZ_HD void hz_enc_put_bit(hz_enc *e, uint32_t b) { /* 9.3.4.2 PutBit */
    if (e->first_bit_flag)
        e->first_bit_flag = 0;
    else
        hz_put_bit_raw(e->sink, b);
    while (e->bits_outstanding > 0) {
        hz_put_bit_raw(e->sink, 1u - b);
        e->bits_outstanding--;
    }
}
HZ_HD void hz_enc_renorm(hz_enc *e) { /* RenormE */
    while (e->range < 256) {
        if (e->low < 256) {
            hz_enc_put_bit(e, 0);
        } else if (e->low >= 512) {
            e->low -= 512;
            hz_enc_put_bit(e, 1);
        } else {
            e->low -= 256;
            e->bits_outstanding++;
        }
        e->range <<= 1;
        e->low <<= 1;
    }
}
/* state byte = pStateIdx | valMPS << 6, the same packing the decoder side uses */
HZ_HD void hz_enc_decision(hz_enc *e, const hz_tables *t, uint8_t *state, uint32_t bin) {
    uint32_t p = *state & 63u, v = (*state >> 6) & 1u;
    uint32_t q = (e->range >> 6) & 3u;
    uint32_t lps = t->range_lps[p * 4 + q];
    e->range -= lps;
    if (bin != v) {
        e->low += e->range;
        e->range = lps;
        if (p == 0) v = 1u - v;
        p = t->trans_lps[p];
    } else {
        p = t->trans_mps[p];
    }
    *state = (uint8_t)(p | (v << 6));
    hz_enc_renorm(e);
}
HZ_HD void hz_enc_bypass(hz_enc *e, uint32_t bin) {
    e->low <<= 1;
    if (bin) e->low += e->range;
    if (e->low >= 1024) {
        hz_enc_put_bit(e, 1);
        e->low -= 1024;
    } else if (e->low < 512) {
        hz_enc_put_bit(e, 0);
    } else {
        e->low -= 512;
        e->bits_outstanding++;
    }
}
HZ_HD void hz_enc_terminate(hz_enc *e, uint32_t bin) {
    e->range -= 2;
    if (bin) {
        e->low += e->range;
        e->range = 2;
        hz_enc_renorm(e);
        hz_enc_put_bit(e, (e->low >> 9) & 1u);
        hz_put_bit_raw(e->sink, (e->low >> 8) & 1u); /* WriteBits(((low >> 7) & 3) | 1, 2) */
        hz_put_bit_raw(e->sink, 1u);
    } else {
        hz_enc_renorm(e);
    }
}

/* ---- context initialisation used to seed encoder states (formula of H.264 9.3.1.1; independent of oracle/) */
HZ_HD uint8_t hz_ctx_state(int m, int n, int qp) {
    int q = qp < 0 ? 0 : (qp > 51 ? 51 : qp);
    int prod = m * q;
    int sh = prod >= 0 ? (prod >> 4) : -((-prod + 15) >> 4);
    int pre = sh + n;
    pre = pre < 1 ? 1 : (pre > 126 ? 126 : pre);
    return (uint8_t)(pre <= 63 ? (63 - pre) : ((pre - 64) | 64));
}

/* ---- the shared op schedule of SURVEY.md §8(d) C2: bin i is terminate(0) if i % 384 == 383, else a decision
 * on a uniformly drawn ctx in [0, n_active) with probability 0.70, else bypass. */
HZ_HD uint16_t hz_sched_op(hz_rng *r, uint64_t i, uint32_t n_active) {
    if (i % 384 == 383) return (uint16_t)(HZ_OP_TERMINATE << 14);
    uint64_t x = hz_next(r);
    if ((x & 0xFFFFu) < 45875u) /* 0.70 * 65536 */
        return (uint16_t)((HZ_OP_DECISION << 14) | (uint32_t)((x >> 16) % n_active));
    return (uint16_t)(HZ_OP_BYPASS << 14);
}

/* Encode one slice: n_ops scheduled bins + the final terminate(1).  Bin values: a decision equals the current MPS
 * with probability 0.5 + 0.45*ctx/n_active (16-bit fixed point), a bypass bin is uniform, terminate(0) inside the
 * schedule.  states[n_ctx] are the initial context states (updated in place).  bins_out (may be NULL) receives the
 * encoded bins, 1 bit per bin, LSB-first in 32-bit words, n_ops+1 bins.  Returns bytes written to the sink. */
HZ_HD int64_t hz_encode_slice(const hz_tables *t, hz_rng *r, const uint16_t *ops, uint32_t n_ops, uint32_t n_active,
                              uint8_t *states, hz_sink *sink, uint32_t *bins_out) {
    hz_enc e;
    hz_enc_init(&e, sink);
    uint32_t word = 0;
    for (uint32_t i = 0; i < n_ops; i++) {
        uint32_t kind = ops[i] >> 14, ctx = ops[i] & 0x3FFu, bin;
        if (kind == HZ_OP_DECISION) {
            uint32_t u = (uint32_t)(hz_next(r) & 0xFFFFu);
            uint32_t thr = 32768u + (29491u * ctx) / n_active;
            uint32_t mps = (states[ctx] >> 6) & 1u;
            bin = (u < thr) ? mps : 1u - mps;
            hz_enc_decision(&e, t, &states[ctx], bin);
        } else if (kind == HZ_OP_BYPASS) {
            bin = (uint32_t)(hz_next(r) & 1u);
            hz_enc_bypass(&e, bin);
        } else {
            bin = 0;
            hz_enc_terminate(&e, 0);
        }
        word |= bin << (i & 31);
        if ((i & 31) == 31) {
            if (bins_out) bins_out[i >> 5] = word;
            word = 0;
        }
    }
    hz_enc_terminate(&e, 1);
    word |= 1u << (n_ops & 31);
    if (bins_out) bins_out[n_ops >> 5] = word;
    hz_sink_finish(sink);
    return sink->n;
}

#endif
