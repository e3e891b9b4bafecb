/* harness.c -- CPU build of the synthetic-data harness (see harness_core.h).  Test/bench input generation only. */
#include "harness_core.h"
#include "harness_tables.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define HZ_TABLES_SPEC 1u
#define HZ_ESCAPE 2u

static hz_tables tables_for(uint32_t flags) {
    hz_tables t;
    if (flags & HZ_TABLES_SPEC) {
        t.range_lps = hz_range_tab_lps_spec;
        t.trans_lps = hz_trans_idx_lps_spec;
        t.trans_mps = hz_trans_idx_mps_spec;
    } else {
        t.range_lps = hz_range_tab_lps_ref;
        t.trans_lps = hz_trans_idx_lps_ref;
        t.trans_mps = hz_trans_idx_mps_ref;
    }
    return t;
}

/* (m,n) lookup with the reference's column rules (see tools/extract_tables.py): idc -1 -> column 0 */
static void hz_mn(uint32_t flags, int ctx, int idc, int *m, int *n) {
    *m = *n = 0;
    if (ctx < 0 || ctx >= HZ_N_CTX_MAX) return;
    int col;
    if (ctx >= 70 && ctx <= 104)
        col = (idc >= 0 && idc <= 2) ? idc + 1 : 0;
    else if (idc >= -1 && idc <= 2)
        col = idc + 1;
    else
        return;
    *m = ((flags & HZ_TABLES_SPEC) ? hz_mn_m_spec : hz_mn_m_ref)[col * HZ_N_CTX_MAX + ctx];
    *n = ((flags & HZ_TABLES_SPEC) ? hz_mn_n_spec : hz_mn_n_ref)[col * HZ_N_CTX_MAX + ctx];
}

void hz_init_states(uint32_t flags, int qp, int idc, int n_ctx, uint8_t *states) {
    for (int c = 0; c < n_ctx; c++) {
        int m, n;
        hz_mn(flags, c, idc, &m, &n);
        states[c] = hz_ctx_state(m, n, qp);
    }
}

void hz_gen_schedule(uint32_t config, uint64_t id, uint64_t n_ops, uint32_t n_active, uint16_t *ops) {
    hz_rng r = hz_seed(config, id);
    for (uint64_t i = 0; i < n_ops; i++) ops[i] = hz_sched_op(&r, i, n_active);
}

typedef struct {
    uint32_t flags, config;
    int64_t first, last; /* slice range */
    uint64_t id_base;
    const uint16_t *ops;
    const uint32_t *n_ops;
    uint32_t n_active, n_ctx;
    const int32_t *qp, *idc;
    uint8_t *data;
    int64_t stride;
    int64_t *lens;
    uint32_t *bins;
    int64_t bins_stride;
    uint8_t *final_states;
    int overflow;
} gen_job;

static void *gen_worker(void *arg) {
    gen_job *j = (gen_job *)arg;
    hz_tables t = tables_for(j->flags);
    uint8_t *states = (uint8_t *)malloc(j->n_ctx);
    for (int64_t s = j->first; s < j->last; s++) {
        hz_init_states(j->flags, j->qp[s], j->idc[s], (int)j->n_ctx, states);
        hz_rng r = hz_seed(j->config, j->id_base + (uint64_t)s);
        hz_sink sink;
        hz_sink_init(&sink, j->data + s * j->stride, j->stride, (j->flags & HZ_ESCAPE) ? 1 : 0);
        hz_encode_slice(&t, &r, j->ops, j->n_ops[s], j->n_active, states, &sink,
                        j->bins ? j->bins + s * j->bins_stride : NULL);
        j->lens[s] = sink.n;
        if (sink.overflow) j->overflow = 1;
        if (j->final_states) memcpy(j->final_states + s * j->n_ctx, states, j->n_ctx);
    }
    free(states);
    return NULL;
}

/* Encode n_slices slices with the shared schedule `ops`; slice s codes n_ops[s] scheduled bins plus the final
 * terminate(1).  Slice s is seeded with (config, id_base + s) and starts from the states given by (qp[s], idc[s]).
 * data[s*stride ...] receives the bytes (emulation-prevention escaped when HZ_ESCAPE), lens[s] their count.
 * Returns 0, or 1 if some slice did not fit its stride. */
int hz_gen_cabac_slices(uint32_t flags, uint32_t config, uint64_t id_base, int64_t n_slices, const uint16_t *ops,
                        const uint32_t *n_ops, uint32_t n_active, uint32_t n_ctx, const int32_t *qp,
                        const int32_t *idc, uint8_t *data, int64_t stride, int64_t *lens, uint32_t *bins,
                        int64_t bins_stride, uint8_t *final_states, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    if ((int64_t)n_threads > n_slices) n_threads = (int)(n_slices > 0 ? n_slices : 1);
    pthread_t th[256];
    gen_job jobs[256];
    int64_t per = (n_slices + n_threads - 1) / n_threads;
    for (int i = 0; i < n_threads; i++) {
        gen_job *j = &jobs[i];
        j->flags = flags;
        j->config = config;
        j->first = i * per;
        j->last = (i + 1) * per < n_slices ? (i + 1) * per : n_slices;
        if (j->first > n_slices) j->first = n_slices;
        j->id_base = id_base;
        j->ops = ops;
        j->n_ops = n_ops;
        j->n_active = n_active;
        j->n_ctx = n_ctx;
        j->qp = qp;
        j->idc = idc;
        j->data = data;
        j->stride = stride;
        j->lens = lens;
        j->bins = bins;
        j->bins_stride = bins_stride;
        j->final_states = final_states;
        j->overflow = 0;
        pthread_create(&th[i], NULL, gen_worker, j);
    }
    int overflow = 0;
    for (int i = 0; i < n_threads; i++) {
        pthread_join(th[i], NULL);
        overflow |= jobs[i].overflow;
    }
    return overflow;
}

/* Random payload of SURVEY.md §8(d) C1: P(00)=1/8, P(01)=P(02)=P(03)=1/16, else uniform in 0x04..0xFF. */
void hz_random_payload(uint32_t config, uint64_t id, int64_t n, uint8_t *out) {
    hz_rng r = hz_seed(config, id);
    for (int64_t i = 0; i < n; i++) {
        uint64_t x = hz_next(&r);
        uint32_t sel = (uint32_t)(x & 15u);
        uint8_t b;
        if (sel < 2)
            b = 0;
        else if (sel < 5)
            b = (uint8_t)(sel - 1); /* 1, 2, 3 */
        else
            b = (uint8_t)(4 + ((x >> 8) % 252u));
        out[i] = b;
    }
}

/* Standard emulation-prevention insertion.  out must hold n + n/2 + 2 bytes.  Returns bytes written. */
int64_t hz_escape(const uint8_t *in, int64_t n, uint8_t *out) {
    hz_sink s;
    hz_sink_init(&s, out, n + n / 2 + 2, 1);
    for (int64_t i = 0; i < n; i++) hz_put_byte(&s, in[i]);
    hz_sink_finish(&s);
    return s.n;
}

/* Encode an explicit list of (op, bin) pairs -- decision bins on the ctxIdx of the op word, bypass bins, terminate bins --
 * then the final terminate(1).  For data-dependent op sequences (syntax elements), where the schedule is the caller's.
 * states[n_ctx]: initial context states, updated in place.  Returns the bytes written (0: out too small). */
int64_t hz_encode_explicit(uint32_t flags, const uint16_t *ops, const uint8_t *bins, int64_t n, uint8_t *states, int64_t n_ctx,
                           uint8_t *out, int64_t cap) {
    const hz_tables t = tables_for(flags);
    hz_sink sink;
    hz_sink_init(&sink, out, cap, (flags & HZ_ESCAPE) ? 1 : 0);
    hz_enc e;
    hz_enc_init(&e, &sink);
    for (int64_t i = 0; i < n; i++) {
        uint32_t kind = ops[i] >> 14, ctx = ops[i] & 0x3FFu;
        if (kind == HZ_OP_DECISION) {
            if ((int64_t)ctx >= n_ctx) ctx = 0;
            hz_enc_decision(&e, &t, &states[ctx], bins[i] & 1u);
        } else if (kind == HZ_OP_BYPASS) {
            hz_enc_bypass(&e, bins[i] & 1u);
        } else {
            hz_enc_terminate(&e, bins[i] & 1u);
            if (bins[i] & 1u) { /* the slice's data ends here (I_PCM / end of slice): nothing may follow */
                hz_sink_finish(&sink);
                return sink.overflow ? 0 : sink.n;
            }
        }
    }
    hz_enc_terminate(&e, 1);
    hz_sink_finish(&sink);
    return sink.overflow ? 0 : sink.n;
}
