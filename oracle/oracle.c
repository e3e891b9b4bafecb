/* oracle.c -- see oracle.h.  TEST INFRASTRUCTURE ONLY; parity unpinned by the reference (no vectors, no Go).
 * A literal, sequential, CPU restatement of the reference's hot path.  Paths cited are under /root/reference. */
#include "oracle.h"
#include "ref_tables.h"
#include <stdlib.h>
#include <string.h>

/* ======================================================================= bit reader (h264/bit_reader.go) */

void orc_br_init(orc_bit_reader *b, const uint8_t *bytes, int64_t len) {
    b->bytes = bytes;
    b->len = len;
    b->byteOffset = b->bitOffset = b->bitsRead = 0;
    b->panicked = 0;
}

/* setOffset, bit_reader.go:169-172 */
static void set_offset(orc_bit_reader *b) {
    b->byteOffset = b->bitsRead / 8;
    b->bitOffset = b->bitsRead % 8;
}

/* degolomb.BitArray(x)[i] -- external, unvendored package; MSB-first is the only reading under which
 * nalUnit.go:82-84 parses forbidden_zero(1)/ref_idc(2)/type(5). */
static int64_t bit_array(uint8_t x, int64_t i) { return (x >> (7 - i)) & 1; }

/* Read, bit_reader.go:292-314.  The EOF tests at :293 and :307 can never fire before b.bytes[b.byteOffset]
 * (:298) indexes out of range, which is a Go panic (A10). */
int64_t orc_br_read(orc_bit_reader *b, int64_t *buf, int64_t n) {
    int64_t i = 0;
    for (;;) {
        if (b->byteOffset < 0 || b->byteOffset >= b->len) { /* :298 index out of range */
            b->panicked = 1;
            return i;
        }
        uint8_t cur = b->bytes[b->byteOffset];
        int64_t from = b->bitOffset; /* the range expression [bitOffset:8] is evaluated once */
        for (int64_t k = from; k < 8; k++) {
            if (i >= n) { /* buf[i] with i == len(buf): only reachable for n == 0 */
                b->panicked = 1;
                return i;
            }
            buf[i] = bit_array(cur, k);
            i++;
            b->bitsRead += 1;
            set_offset(b);
            if (i >= n) return n;
        }
    }
}

/* bitVal, bit_reader.go:50-59 (1 << k wraps to 0 for k >= 64 in Go) */
int64_t orc_bit_val(const int64_t *bits, int64_t n) {
    uint64_t t = 0;
    for (int64_t i = 0; i < n; i++) {
        if (bits[i] == 1) {
            int64_t sh = (n - 1) - i;
            if (sh < 64) t += (uint64_t)1 << sh;
        }
    }
    return (int64_t)t;
}

/* NextField, bit_reader.go:315-325 */
int64_t orc_br_next_field(orc_bit_reader *b, int64_t bits) {
    int64_t buf[64];
    if (bits > 64) bits = 64;
    orc_br_read(b, buf, bits);
    if (b->panicked) return -1;
    return orc_bit_val(buf, bits);
}

/* ReadOneBit, bit_reader.go:232-236 */
int64_t orc_br_read_one_bit(orc_bit_reader *b) {
    int64_t buf[1] = {0};
    orc_br_read(b, buf, 1);
    return buf[0];
}

/* golomb, bit_reader.go:174-196: leading zeros up to and including the first 1, then `zeros` more bits */
int64_t orc_br_golomb(orc_bit_reader *b, int64_t *bits, int64_t cap) {
    int64_t zeros = -1, bit = 0, nb = 0;
    while (bit != 1) {
        zeros += 1;
        if (b->byteOffset >= b->len) { /* :180 index out of range */
            b->panicked = 1;
            return nb;
        }
        bit = bit_array(b->bytes[b->byteOffset], b->bitOffset);
        b->bitsRead += 1;
        set_offset(b);
        if (nb < cap) bits[nb] = bit;
        nb++;
    }
    if (zeros == 0) return nb;
    for (int64_t i = 0; i < zeros; i++) {
        if (b->byteOffset >= b->len) { /* :189 */
            b->panicked = 1;
            return nb;
        }
        bit = bit_array(b->bytes[b->byteOffset], b->bitOffset);
        b->bitsRead += 1;
        set_offset(b);
        if (nb < cap) bits[nb] = bit;
        nb++;
    }
    return nb;
}

/* ue, bit_reader.go:62-64 */
int64_t orc_ue(const int64_t *bits, int64_t n) { return orc_bit_val(bits, n) - 1; }

/* se, bit_reader.go:158-161.  codeNum/2 is Go integer division BEFORE math.Ceil, so the magnitude is
 * floor(codeNum/2) (A9): codeNum 1 -> 0, 3 -> +1, 5 -> +2; even codeNums give the correct negatives. */
int64_t orc_se(const int64_t *bits, int64_t n) {
    int64_t codeNum = orc_bit_val(bits, n) - 1;
    int64_t sign = ((codeNum + 1) % 2 == 0) ? 1 : -1; /* (-1)^(codeNum+1) */
    return sign * (codeNum / 2);
}

/* MoreRBSPData, bit_reader.go:199-219 */
int orc_br_more_rbsp_data(orc_bit_reader *b) {
    if (b->len - b->byteOffset == 0) return 0;
    int64_t buf[1] = {0};
    int64_t cnt = 0;
    while (buf[0] != 1) {
        orc_br_read(b, buf, 1);
        if (b->panicked) return 0;
        cnt++;
    }
    return cnt > 0;
}

/* HasMoreData, bit_reader.go:220-226 */
int orc_br_has_more_data(const orc_bit_reader *b) { return b->len - b->byteOffset > 0; }

/* PeekBytes :263-269 / ReadByte :272-279 / ReadBytes :280-290 */
static int peek_bytes(const orc_bit_reader *b, int64_t n, const uint8_t **out) {
    if (b->len >= b->byteOffset + n) {
        *out = b->bytes + b->byteOffset;
        return 0;
    }
    return -1;
}
static int read_byte(orc_bit_reader *b, uint8_t *out) {
    if (b->len > b->byteOffset) {
        *out = b->bytes[b->byteOffset];
        b->byteOffset += 1;
        return 0;
    }
    *out = 0;
    return -1;
}
static int64_t read_bytes(orc_bit_reader *b, int64_t n, uint8_t *buf) {
    int64_t got = 0;
    for (int64_t i = 0; i < n; i++) {
        uint8_t x;
        if (read_byte(b, &x) == 0)
            buf[got++] = x;
        else
            return got;
    }
    return got;
}

/* ======================================================================= NAL units */

/* isStartSequence, server.go:28-39 (InitialNALU = 00 00 00 01, server.go:19) */
int orc_is_start_sequence(const uint8_t *packet, int64_t len) {
    static const uint8_t InitialNALU[4] = {0, 0, 0, 1};
    if (len < 4) return 0;
    const uint8_t *seg = packet + len - 4;
    for (int i = 0; i < 4; i++)
        if (seg[i] != InitialNALU[i]) return 0;
    return 1;
}

/* isEmulationPreventionThreeByte, nalUnit.go:32-37 */
static int is_epb3(const uint8_t *b, int64_t len) {
    if (len != 3) return 0;
    return b[0] == 0 && b[1] == 0 && b[2] == 3;
}

/* NewNalUnit, nalUnit.go:75-131 (extension headers :39-71) */
int orc_new_nal_unit(const uint8_t *frame, int64_t frame_len, int64_t num_bytes_in_nal, orc_nal_unit *u,
                     uint8_t *rbsp_out) {
    memset(u, 0, sizeof(*u));
    u->NumBytes = num_bytes_in_nal;
    u->HeaderBytes = 1;
    orc_bit_reader br, *b = &br;
    orc_br_init(b, frame, frame_len);
#define NF(field, nbits)                       \
    do {                                       \
        u->field = orc_br_next_field(b, nbits); \
        if (b->panicked) return ORC_PANIC;     \
    } while (0)
    NF(ForbiddenZeroBit, 1);
    NF(RefIdc, 2);
    NF(Type, 5);
    if (u->Type == 14 || u->Type == 20 || u->Type == 21) {
        if (u->Type != 21)
            NF(SvcExtensionFlag, 1);
        else
            NF(Avc3dExtensionFlag, 1);
        if (u->SvcExtensionFlag == 1) { /* NalUnitHeaderSvcExtension, :39-51 */
            NF(IdrFlag, 1);
            NF(PriorityId, 6);
            NF(NoInterLayerPredFlag, 1);
            NF(DependencyId, 3);
            NF(QualityId, 4);
            NF(TemporalId, 3);
            NF(UseRefBasePicFlag, 1);
            NF(DiscardableFlag, 1);
            NF(OutputFlag, 1);
            NF(ReservedThree2Bits, 2);
            u->HeaderBytes += 3;
        } else if (u->Avc3dExtensionFlag == 1) { /* NalUnitHeader3davcExtension, :53-61 */
            NF(ViewIdx, 8);
            NF(DepthFlag, 1);
            NF(NonIdrFlag, 1);
            NF(TemporalId, 3);
            NF(AnchorPicFlag, 1);
            NF(InterViewFlag, 1);
            u->HeaderBytes += 2;
        } else { /* NalUnitHeaderMvcExtension, :62-71 */
            NF(NonIdrFlag, 1);
            NF(PriorityId, 6);
            NF(ViewId, 10);
            NF(TemporalId, 3);
            NF(AnchorPicFlag, 1);
            NF(InterViewFlag, 1);
            NF(ReservedOneBit, 1);
            u->HeaderBytes += 3;
        }
    }
#undef NF
    /* body loop, :106-126.  The byte cursor (b.byteOffset) equals i throughout. */
    int64_t nr = 0;
    for (int64_t i = u->HeaderBytes; i < u->NumBytes; i++) {
        const uint8_t *next3;
        if (peek_bytes(b, 3, &next3) != 0) break; /* :107-111 -- the last two bytes are never copied (A7) */
        if (i + 2 < u->NumBytes && is_epb3(next3, 3)) {
            uint8_t tmp[3];
            read_bytes(b, 3, tmp);
            rbsp_out[nr++] = tmp[0];
            rbsp_out[nr++] = tmp[1];
            i += 2;
            u->EmulationPreventionThreeByte = tmp[2];
        } else {
            uint8_t x;
            if (read_byte(b, &x) == 0)
                rbsp_out[nr++] = x;
            else
                break;
        }
    }
    u->rbsp_len = nr;
    return ORC_OK;
}

/* readNalUnit (server.go:64-111) driven by the for-loop of handleConnection (server.go:144-146). */
int64_t orc_read_nal_units(const uint8_t *stream, int64_t n, orc_stream_nal *out, int64_t cap, uint8_t *rbsp_buf,
                           int64_t rbsp_cap, int64_t *rbsp_total, int literal) {
    /* H264Reader: bytes accumulate forever (bit_reader.go:33); byteOffset == len(bytes) (bit_reader.go:37) */
    uint8_t *acc = NULL;
    int64_t acc_cap = 0, len = 0, pos = 0, count = 0, rtot = 0;
    int64_t rc = 0;
    const uint8_t *bytes = literal ? NULL : stream;
/* BufferToReader(1), bit_reader.go:27-39: one byte from the io.Reader appended to h.bytes */
#define BUFFER_ONE()                                             \
    do {                                                         \
        if (pos >= n) goto end_of_stream; /* Read error -> nil */ \
        if (literal) {                                           \
            if (len == acc_cap) {                                \
                acc_cap = acc_cap ? acc_cap * 2 : 8;             \
                acc = (uint8_t *)realloc(acc, (size_t)acc_cap);  \
            }                                                    \
            acc[len] = stream[pos];                              \
            bytes = acc;                                         \
        }                                                        \
        pos++;                                                   \
        len++;                                                   \
    } while (0)
    for (;;) {
        /* :68-72 read to start of NAL */
        while (!orc_is_start_sequence(bytes, len)) BUFFER_ONE();
        int64_t startOffset = len; /* :88 */
        int64_t so = len;          /* :92 */
        while (so == startOffset || !orc_is_start_sequence(bytes, len)) { /* :93-98 */
            so = len;
            BUFFER_ONE();
        }
        int64_t endOffset = len; /* :103; the rewind at :99-102 is commented out, so the NAL keeps the start code (A8) */
        if (count >= cap) {
            rc = -ORC_CAPACITY;
            goto done;
        }
        int64_t nb = endOffset - startOffset;
        if (rtot + nb > rbsp_cap) {
            rc = -ORC_CAPACITY;
            goto done;
        }
        orc_stream_nal *o = &out[count];
        o->start_offset = startOffset;
        o->end_offset = endOffset;
        o->rbsp_off = rtot;
        /* :105,109 NewNalUnit(r.Bytes()[startOffset:], len) */
        int st = orc_new_nal_unit(bytes + startOffset, nb, nb, &o->nal, rbsp_buf + rtot);
        if (st != ORC_OK) { /* panic -> recover -> os.Exit(1), server.go:136-143 */
            rc = -ORC_PANIC;
            goto done;
        }
        rtot += o->nal.rbsp_len;
        count++;
    }
end_of_stream: /* readNalUnit returns nil (server.go:69-71,95-97); nalUnit.Type then panics (:147): the NAL after
                  the last start code is never emitted */
    rc = count;
done:
#undef BUFFER_ONE
    if (rbsp_total) *rbsp_total = rtot;
    free(acc);
    return rc;
}

/* ======================================================================= CABAC engine (h264/cabac.go) */

static const uint8_t *range_tab(uint32_t flags) {
    return (flags & ORC_TABLES_SPEC) ? orc_range_tab_lps_spec : orc_range_tab_lps_ref;
}
static const uint8_t *trans_lps(uint32_t flags) {
    return (flags & ORC_TABLES_SPEC) ? orc_trans_idx_lps_spec : orc_trans_idx_lps_ref;
}
static const uint8_t *trans_mps(uint32_t flags) {
    return (flags & ORC_TABLES_SPEC) ? orc_trans_idx_mps_spec : orc_trans_idx_mps_ref;
}

/* initDecodingEngine, cabac.go:439-446 */
void orc_init_decoding_engine(orc_bit_reader *b, int64_t *codIRange, int64_t *codIOffset) {
    *codIRange = 510;
    *codIOffset = orc_br_next_field(b, 9);
}

/* BinaryDecision arithmetic core, cabac.go:525-536.  The reference re-derives (pStateIdx, valMPS) through
 * initCabac on every call (:523) and ignores ctxIdx; here the context is an explicit argument (A6). */
void orc_binary_decision(uint32_t flags, int64_t pStateIdx, int64_t valMPS, int64_t *codIRange, int64_t *codIOffset,
                         int64_t *binVal) {
    int64_t qCodIRangeIdx = (*codIRange >> 6) & 3;
    int64_t codIRangeLPS = range_tab(flags)[pStateIdx * 4 + qCodIRangeIdx];
    *codIRange = *codIRange - codIRangeLPS;
    if (*codIOffset >= *codIRange) {
        *binVal = 1 - valMPS;
        *codIOffset -= *codIRange;
        *codIRange = codIRangeLPS;
    } else {
        *binVal = valMPS;
    }
}

/* StateTransitionProcess, cabac.go:544-553 */
void orc_state_transition(uint32_t flags, int64_t *pStateIdx, int64_t *valMPS, int64_t binVal) {
    if (binVal == *valMPS) {
        *pStateIdx = trans_mps(flags)[*pStateIdx];
    } else {
        if (*pStateIdx == 0) *valMPS = 1 - *valMPS;
        *pStateIdx = trans_lps(flags)[*pStateIdx];
    }
}

/* RenormD, cabac.go:503-511 (tail recursion written as the equivalent loop) */
void orc_renorm_d(orc_bit_reader *b, int64_t *codIRange, int64_t *codIOffset) {
    while (*codIRange < 256) {
        *codIRange = *codIRange << 1;
        *codIOffset = (int64_t)((uint64_t)*codIOffset << 1);
        *codIOffset = *codIOffset | orc_br_read_one_bit(b);
        if (b->panicked) return;
    }
}

/* DecodeBypass, cabac.go:468-481.  REF: codIOffset <<= 1 then codIOffset <<= bit (:470,473); Go int wraps
 * silently at 64 bits and the compare at :474 is signed (A5).  SPEC_OR: (codIOffset << 1) | bit. */
void orc_decode_bypass(uint32_t flags, orc_bit_reader *b, int64_t codIRange, int64_t *codIOffset, int64_t *binVal) {
    uint64_t o = (uint64_t)*codIOffset << 1;
    int64_t bit = orc_br_read_one_bit(b);
    if (b->panicked) return;
    if (flags & ORC_BYPASS_SPEC_OR)
        o = o | (uint64_t)bit;
    else
        o = o << (unsigned)bit;
    int64_t so = (int64_t)o;
    if (so >= codIRange) {
        *binVal = 1;
        so = (int64_t)((uint64_t)so - (uint64_t)codIRange);
    } else {
        *binVal = 0;
    }
    *codIOffset = so;
}

/* DecodeTerminate, cabac.go:486-499 */
void orc_decode_terminate(orc_bit_reader *b, int64_t *codIRange, int64_t *codIOffset, int64_t *binVal) {
    *codIRange -= 2;
    if (*codIOffset >= *codIRange) {
        *binVal = 1; /* no renormalisation on this branch (:488-493) */
        return;
    }
    *binVal = 0;
    orc_renorm_d(b, codIRange, codIOffset);
}

/* Composition 9.3.3.2.1 the reference's section comments point to: :525-536 -> :544-553 -> :503-511 */
void orc_decode_decision(uint32_t flags, orc_bit_reader *b, uint8_t *ctx_state, int64_t *codIRange,
                         int64_t *codIOffset, int64_t *binVal) {
    int64_t p = *ctx_state & 63, v = (*ctx_state >> 6) & 1;
    orc_binary_decision(flags, p, v, codIRange, codIOffset, binVal);
    orc_state_transition(flags, &p, &v, *binVal);
    *ctx_state = (uint8_t)(p | (v << 6));
    orc_renorm_d(b, codIRange, codIOffset);
}

int orc_cabac_decode_slice(uint32_t flags, const uint8_t *bytes, int64_t len, const uint16_t *ops, int64_t n_ops,
                           uint8_t *ctx_state, int64_t n_ctx, uint32_t *bins_out, orc_cabac_final *fin) {
    orc_bit_reader br;
    orc_br_init(&br, bytes, len);
    int64_t R, O;
    int64_t done = 0;
    orc_init_decoding_engine(&br, &R, &O);
    if (!br.panicked) {
        for (int64_t i = 0; i < n_ops; i++) {
            uint32_t kind = ops[i] >> 14, ctx = ops[i] & 0x3ffu;
            int64_t bin = 0;
            int64_t R0 = R, O0 = O, bits0 = br.bitsRead;
            uint8_t st0 = 0;
            if (kind == ORC_OP_DECISION) {
                if ((int64_t)ctx >= n_ctx) ctx = 0;
                st0 = ctx_state[ctx];
                orc_decode_decision(flags, &br, &ctx_state[ctx], &R, &O, &bin);
            } else if (kind == ORC_OP_BYPASS) {
                orc_decode_bypass(flags, &br, R, &O, &bin);
            } else {
                orc_decode_terminate(&br, &R, &O, &bin);
            }
            if (br.panicked) { /* the op that ran off the end produces nothing; state as before it */
                R = R0;
                O = O0;
                br.bitsRead = bits0;
                if (kind == ORC_OP_DECISION) ctx_state[ctx] = st0;
                break;
            }
            if ((i & 31) == 0) bins_out[i >> 5] = 0;
            bins_out[i >> 5] |= (uint32_t)bin << (i & 31);
            done++;
        }
    }
    fin->codIRange = R;
    fin->codIOffset = O;
    fin->bitsRead = br.bitsRead;
    fin->flags = br.panicked ? 1u : 0u;
    fin->n_bins = (uint32_t)done;
    return br.panicked ? ORC_PANIC : ORC_OK;
}

/* ======================================================================= context initialisation */

/* Clip3, cabac.go:131-139 */
int64_t orc_clip3(int64_t x, int64_t y, int64_t z) {
    if (z < x) return x;
    if (z > y) return y;
    return z;
}

/* PreCtxState, cabac.go:118-121.  Go's >> on a negative int is an arithmetic shift (floor). */
int64_t orc_pre_ctx_state(int64_t m, int64_t n, int64_t sliceQPy) {
    int64_t prod = m * orc_clip3(0, 51, sliceQPy);
    int64_t sh = prod >= 0 ? (prod >> 4) : -((-prod + 15) >> 4); /* floor(prod / 16) without relying on signed >> */
    return orc_clip3(1, 126, sh + n);
}

/* SliceQPy, cabac.go:113-115 */
int64_t orc_slice_qpy(int64_t pic_init_qp_minus26, int64_t slice_qp_delta) {
    return 26 + pic_init_qp_minus26 + slice_qp_delta;
}

/* MNVars[ctxIdx][idc] (mn_vars.go:15-175) / CodedblockPatternMN(ctxIdx, idc) (mn_vars.go:184-440).
 * A missing Go map key yields MN{0,0}; CodedblockPatternMN returns the I/SI column for any idc outside 0..2. */
void orc_mn(uint32_t flags, int64_t ctxIdx, int64_t cabacInitIdc, int64_t *m, int64_t *n) {
    *m = 0;
    *n = 0;
    if (ctxIdx < 0 || ctxIdx >= ORC_N_CTX_MAX) return;
    int64_t col;
    if (ctxIdx >= 70 && ctxIdx <= 104)
        col = (cabacInitIdc >= 0 && cabacInitIdc <= 2) ? cabacInitIdc + 1 : 0;
    else if (cabacInitIdc >= -1 && cabacInitIdc <= 2)
        col = cabacInitIdc + 1;
    else
        return; /* no such key in MNVars[ctxIdx] */
    const int8_t *tm = (flags & ORC_TABLES_SPEC) ? orc_mn_m_spec : orc_mn_m_ref;
    const int8_t *tn = (flags & ORC_TABLES_SPEC) ? orc_mn_n_spec : orc_mn_n_ref;
    *m = tm[col * ORC_N_CTX_MAX + ctxIdx];
    *n = tn[col * ORC_N_CTX_MAX + ctxIdx];
}

/* initCabac state split, cabac.go:158-164 */
uint8_t orc_ctx_state(int64_t preCtxState) {
    int64_t pStateIdx, valMPS;
    if (preCtxState <= 63) {
        pStateIdx = 63 - preCtxState;
        valMPS = 0;
    } else {
        pStateIdx = preCtxState - 64;
        valMPS = 1;
    }
    return (uint8_t)(pStateIdx | (valMPS << 6));
}

void orc_ctx_init(uint32_t flags, const int32_t *qp, const int32_t *idc, int64_t n_slices, int64_t n_ctx,
                  uint8_t *states) {
    for (int64_t s = 0; s < n_slices; s++)
        for (int64_t c = 0; c < n_ctx; c++) {
            int64_t m, n;
            orc_mn(flags, c, idc[s], &m, &n);
            states[s * n_ctx + c] = orc_ctx_state(orc_pre_ctx_state(m, n, qp[s]));
        }
}

/* ======================================================================= SPS / PPS */

#define GOLOMB_CAP 130
#define PANIC_CHECK()                          \
    do {                                       \
        if (b->panicked) {                     \
            out->bits_read = b->bitsRead;      \
            return ORC_PANIC;                  \
        }                                      \
    } while (0)
#define FIELD(dst, nbits)                    \
    do {                                     \
        (dst) = orc_br_next_field(b, nbits); \
        PANIC_CHECK();                       \
    } while (0)
#define FLAG(dst)                                \
    do {                                         \
        int64_t v_ = orc_br_next_field(b, 1);    \
        PANIC_CHECK();                           \
        (dst) = (v_ == 1);                       \
    } while (0)
/* ue(b.golomb()) / se(b.golomb()) over ALL bits of the code word, however long (sh_golomb / sh_se below) */
static int sh_golomb(orc_bit_reader *b, int64_t *val);
static int64_t sh_se(int64_t bitval);
#define UE(dst)                                        \
    do {                                               \
        int64_t bv_;                                   \
        (void)gb;                                      \
        sh_golomb(b, &bv_);                            \
        PANIC_CHECK();                                 \
        (dst) = (int64_t)((uint64_t)bv_ - 1u);         \
    } while (0)
#define SE(dst)                                        \
    do {                                               \
        int64_t bv_;                                   \
        (void)gb;                                      \
        sh_golomb(b, &bv_);                            \
        PANIC_CHECK();                                 \
        (dst) = sh_se(bv_);                            \
    } while (0)

/* scalingList, sps.go:172-191.  The decoded values land in package-global default lists (sps.go:141-155), which
 * nothing on this path reads back; only the bits consumed matter here. */
static int scaling_list(orc_bit_reader *b, int64_t size) {
    int64_t gb[GOLOMB_CAP];
    int64_t lastScale = 8, nextScale = 8;
    for (int64_t i = 0; i < size; i++) {
        if (nextScale != 0) {
            int64_t bv;
            (void)gb;
            sh_golomb(b, &bv);
            if (b->panicked) return ORC_PANIC;
            int64_t deltaScale = sh_se(bv);
            nextScale = (int64_t)((uint64_t)lastScale + (uint64_t)deltaScale + 256u) % 256;
        }
        lastScale = (nextScale == 0) ? lastScale : nextScale;
    }
    return ORC_OK;
}

/* hrdParameters closure, sps.go:197-216: the four *_length fields are read INSIDE the SchedSelIdx loop (A12) */
static int hrd_parameters(orc_bit_reader *b, orc_sps *out) {
    int64_t gb[GOLOMB_CAP];
    UE(out->CpbCntMinus1);
    FIELD(out->BitRateScale, 4);
    FIELD(out->CpbSizeScale, 4);
    for (int64_t i = 0; i <= out->CpbCntMinus1; i++) {
        int64_t a, c, cbr;
        UE(a);
        UE(c);
        FLAG(cbr);
        if (out->n_hrd < ORC_MAX_LIST) { /* n_hrd = len() of the three lists; the first ORC_MAX_LIST are kept */
            out->BitRateValueMinus1[out->n_hrd] = a;
            out->CpbSizeValueMinus1[out->n_hrd] = c;
            out->Cbr[out->n_hrd] = cbr;
        }
        out->n_hrd++;
        FIELD(out->InitialCpbRemovalDelayLengthMinus1, 5);
        FIELD(out->CpbRemovalDelayLengthMinus1, 5);
        FIELD(out->DpbOutputDelayLengthMinus1, 5);
        FIELD(out->TimeOffsetLength, 5);
    }
    return ORC_OK;
}

/* NewSPS, sps.go:192-437 */
int orc_new_sps(const uint8_t *rbsp, int64_t len, orc_sps *out) {
    static const int64_t isProfileIDC[] = {100, 110, 122, 244, 44, 83, 86, 118, 128, 138, 139, 134, 135};
    int64_t gb[GOLOMB_CAP];
    orc_bit_reader br, *b = &br;
    memset(out, 0, sizeof(*out));
    orc_br_init(b, rbsp, len);
    int64_t tmp;
    FIELD(out->Profile, 8);
    FIELD(out->Constraint0, 1);
    FIELD(out->Constraint1, 1);
    FIELD(out->Constraint2, 1);
    FIELD(out->Constraint3, 1);
    FIELD(out->Constraint4, 1);
    FIELD(out->Constraint5, 1);
    FIELD(tmp, 2); /* ReservedZeroBits */
    (void)tmp;
    FIELD(out->Level, 8);
    UE(out->ID);
    UE(out->ChromaFormat);
    int special = 0;
    for (unsigned i = 0; i < sizeof(isProfileIDC) / sizeof(isProfileIDC[0]); i++)
        if (isProfileIDC[i] == out->Profile) special = 1;
    if (special) {
        if (out->ChromaFormat == 3) FLAG(out->UseSeparateColorPlane);
        UE(out->BitDepthLumaMinus8);
        UE(out->BitDepthChromaMinus8);
        FLAG(out->QPrimeYZeroTransformBypass);
        FLAG(out->SeqScalingMatrixPresent);
        if (out->SeqScalingMatrixPresent) {
            int64_t max = (out->ChromaFormat != 3) ? 8 : 12;
            for (int64_t i = 0; i < max; i++) {
                int64_t present;
                FLAG(present);
                out->SeqScalingList[out->n_SeqScalingList++] = present;
                if (present) {
                    if (i < 6) {
                        /* DefaultScalingMatrix4x4[i] has 2 rows (sps.go:106-109): index >= 2 panics (A12) */
                        if (i >= 2) {
                            out->bits_read = b->bitsRead;
                            return ORC_PANIC;
                        }
                        if (scaling_list(b, 16) != ORC_OK) PANIC_CHECK();
                    } else {
                        if (i - 6 >= 2) { /* DefaultScalingMatrix8x8[i-6], sps.go:111-128 */
                            out->bits_read = b->bitsRead;
                            return ORC_PANIC;
                        }
                        if (scaling_list(b, 64) != ORC_OK) PANIC_CHECK();
                    }
                }
            }
        }
    }
    UE(out->Log2MaxFrameNumMinus4);
    UE(out->PicOrderCountType);
    if (out->PicOrderCountType == 0) {
        UE(out->Log2MaxPicOrderCntLSBMin4);
    } else if (out->PicOrderCountType == 1) {
        FLAG(out->DeltaPicOrderAlwaysZero);
        SE(out->OffsetForNonRefPic);
        SE(out->OffsetForTopToBottomField);
        UE(out->NumRefFramesInPicOrderCntCycle);
        for (int64_t i = 0; i < out->NumRefFramesInPicOrderCntCycle; i++) {
            int64_t v;
            SE(v);
            if (out->n_OffsetForRefFrameList < ORC_MAX_LIST)
                out->OffsetForRefFrameList[out->n_OffsetForRefFrameList] = v;
            out->n_OffsetForRefFrameList++;
        }
    }
    UE(out->MaxNumRefFrames);
    FLAG(out->GapsInFrameNumValueAllowed);
    UE(out->PicWidthInMbsMinus1);
    UE(out->PicHeightInMapUnitsMinus1);
    FLAG(out->FrameMbsOnly);
    if (!out->FrameMbsOnly) FLAG(out->MBAdaptiveFrameField);
    FLAG(out->Direct8x8Inference);
    FLAG(out->FrameCropping);
    if (out->FrameCropping) {
        UE(out->FrameCropLeftOffset);
        UE(out->FrameCropRightOffset);
        UE(out->FrameCropTopOffset);
        UE(out->FrameCropBottomOffset);
    }
    FLAG(out->VuiParametersPresent);
    if (out->VuiParametersPresent) {
        FLAG(out->AspectRatioInfoPresent);
        if (out->AspectRatioInfoPresent) {
            FIELD(out->AspectRatio, 8);
            if (out->AspectRatio == 999) { /* EXTENDED_SAR := 999, sps.go:347 -- never true for an 8-bit field (A12) */
                FIELD(out->SarWidth, 16);
                FIELD(out->SarHeight, 16);
            }
        }
        FLAG(out->OverscanInfoPresent);
        if (out->OverscanInfoPresent) FLAG(out->OverscanAppropriate);
        FLAG(out->VideoSignalTypePresent);
        if (out->VideoSignalTypePresent) FIELD(out->VideoFormat, 3);
        if (out->VideoSignalTypePresent) {
            FLAG(out->VideoFullRange);
            FLAG(out->ColorDescriptionPresent);
            if (out->ColorDescriptionPresent) {
                FIELD(out->ColorPrimaries, 8);
                FIELD(out->TransferCharacteristics, 8);
                FIELD(out->MatrixCoefficients, 8);
            }
        }
        FLAG(out->ChromaLocInfoPresent);
        if (out->ChromaLocInfoPresent) {
            UE(out->ChromaSampleLocTypeTopField);
            UE(out->ChromaSampleLocTypeBottomField);
        }
        FLAG(out->TimingInfoPresent);
        if (out->TimingInfoPresent) {
            FIELD(out->NumUnitsInTick, 32);
            FIELD(out->TimeScale, 32);
            FLAG(out->FixedFrameRate);
        }
        FLAG(out->NalHrdParametersPresent);
        if (out->NalHrdParametersPresent)
            if (hrd_parameters(b, out) != ORC_OK) return ORC_PANIC;
        FLAG(out->VclHrdParametersPresent);
        if (out->VclHrdParametersPresent)
            if (hrd_parameters(b, out) != ORC_OK) return ORC_PANIC;
        if (out->NalHrdParametersPresent || out->VclHrdParametersPresent) FLAG(out->LowHrdDelay);
        FLAG(out->PicStructPresent);
        FLAG(out->BitstreamRestriction);
        if (out->BitstreamRestriction) {
            FLAG(out->MotionVectorsOverPicBoundaries);
            UE(out->MaxBytesPerPicDenom);
            UE(out->MaxBitsPerMbDenom);
            UE(out->Log2MaxMvLengthHorizontal);
            UE(out->Log2MaxMvLengthVertical);
            UE(out->MaxNumReorderFrames); /* order as in the reference: reorder before dec-buffering, sps.go:427-428 */
            UE(out->MaxDecFrameBuffering);
        }
    }
    out->bits_read = b->bitsRead;
    return ORC_OK;
}

/* NewPPS, pps.go:40-133 */
int orc_new_pps(int64_t sps_chroma_format, const uint8_t *rbsp, int64_t len, orc_pps *out) {
    int64_t gb[GOLOMB_CAP];
    orc_bit_reader br, *b = &br;
    (void)sps_chroma_format;
    memset(out, 0, sizeof(*out));
    orc_br_init(b, rbsp, len);
    UE(out->ID);
    UE(out->SPSID);
    FIELD(out->EntropyCodingMode, 1);
    FLAG(out->BottomFieldPicOrderInFramePresent);
    UE(out->NumSliceGroupsMinus1);
    if (out->NumSliceGroupsMinus1 > 0) {
        UE(out->SliceGroupMapType);
        if (out->SliceGroupMapType == 0 || out->SliceGroupMapType == 2 || out->SliceGroupMapType == 6) {
            /* pps.go:61,65-66,74 assign into nil slices: the first iteration panics (A11).  For type 6 the size
             * field is read first (:72). */
            if (out->SliceGroupMapType == 6) UE(out->PicSizeInMapUnitsMinus1);
            /* type 6: the loop `i <= PicSizeInMapUnitsMinus1` (pps.go:73) does not run for a negative size (a code word
             * of 64 leading zeros or more wraps to -1) */
            if (out->SliceGroupMapType != 6 || out->PicSizeInMapUnitsMinus1 >= 0) {
                out->bits_read = b->bitsRead;
                return ORC_PANIC;
            }
        } else if (out->SliceGroupMapType > 2 && out->SliceGroupMapType < 6) {
            FLAG(out->SliceGroupChangeDirection);
            UE(out->SliceGroupChangeRateMinus1);
        }
    }
    UE(out->NumRefIdxL0DefaultActiveMinus1);
    UE(out->NumRefIdxL1DefaultActiveMinus1);
    FLAG(out->WeightedPred);
    FIELD(out->WeightedBipred, 2);
    SE(out->PicInitQpMinus26);
    SE(out->PicInitQsMinus26);
    SE(out->ChromaQpIndexOffset);
    FLAG(out->DeblockingFilterControlPresent);
    FLAG(out->ConstrainedIntraPred);
    FLAG(out->RedundantPicCntPresent);
    if (orc_br_has_more_data(b)) { /* byte-granular (bit_reader.go:225): with A8's trailing 00 00 always true */
        FIELD(out->Transform8x8Mode, 1);
        FLAG(out->PicScalingMatrixPresent);
        if (out->PicScalingMatrixPresent) {
            /* pps.go:103 writes PicScalingListPresent[i] of a nil slice after reading the flag: panic (A11) */
            int64_t f;
            FLAG(f);
            (void)f;
            out->bits_read = b->bitsRead;
            return ORC_PANIC;
        }
        orc_br_more_rbsp_data(b); /* consumes bits up to and including the next 1; panics if there is none */
        PANIC_CHECK();
    }
    out->bits_read = b->bitsRead;
    return ORC_OK;
}

/* ====================================================================================================== slice header
 * NewSliceContext, h264/slice.go:835-1048, up to the call of NewSliceData (:1046).  Test infrastructure. */
#include <math.h>

/* golomb() + bitVal() over ALL bits of the code (bit_reader.go:174-196, :50-59): bitVal adds 1 << (len-1-i) per set
 * bit and Go shifts of 64 or more give 0, so the value is the code's bit string modulo 2^64 */
static int sh_golomb(orc_bit_reader *b, int64_t *val) {
    uint64_t t = 0;
    int64_t zeros = -1, bit = 0;
    while (bit != 1) {
        zeros += 1;
        if (b->byteOffset >= b->len) { /* :180 index out of range */
            b->panicked = 1;
            return ORC_PANIC;
        }
        bit = bit_array(b->bytes[b->byteOffset], b->bitOffset);
        b->bitsRead += 1;
        set_offset(b);
        t = (t << 1) | (uint64_t)bit;
    }
    for (int64_t i = 0; i < zeros; i++) {
        if (b->byteOffset >= b->len) { /* :189 */
            b->panicked = 1;
            return ORC_PANIC;
        }
        bit = bit_array(b->bytes[b->byteOffset], b->bitOffset);
        b->bitsRead += 1;
        set_offset(b);
        t = (t << 1) | (uint64_t)bit;
    }
    *val = (int64_t)t;
    return ORC_OK;
}
/* se(), bit_reader.go:158-161, through float64 exactly as written: int(math.Pow(-1, float64(codeNum+1)) *
 * math.Ceil(float64(codeNum/2))) */
static int64_t sh_se(int64_t bitval) {
    const int64_t codeNum = (int64_t)((uint64_t)bitval - 1u);
    const double sign = pow(-1.0, (double)(int64_t)((uint64_t)codeNum + 1u));
    return (int64_t)(sign * ceil((double)(codeNum / 2)));
}
/* NextField(name, n), bit_reader.go:315-325 -> Read (:292-314): make([]int, n) panics for n < 0 */
static int sh_field(orc_bit_reader *b, int64_t n, int64_t *val) {
    uint64_t t = 0;
    if (n < 0) {
        b->panicked = 1;
        return ORC_PANIC;
    }
    for (int64_t i = 0; i < n; i++) {
        if (b->byteOffset >= b->len) { /* :298 */
            b->panicked = 1;
            return ORC_PANIC;
        }
        t = (t << 1) | (uint64_t)bit_array(b->bytes[b->byteOffset], b->bitOffset);
        b->bitsRead += 1;
        set_offset(b);
    }
    *val = (int64_t)t;
    return ORC_OK;
}
/* sliceTypeMap, slice.go:105-116: 0 "P", 1 "B", 2 "I", 3 "SP", 4 "SI", 5..9 the same again, anything else "" */
enum { ST_P, ST_B, ST_I, ST_SP, ST_SI, ST_NONE };
static int slice_type_name(int64_t t) { return (t >= 0 && t <= 9) ? (int)(t % 5) : ST_NONE; }

#define SH_UE(dst)                               \
    do {                                         \
        int64_t v_;                              \
        if (sh_golomb(b, &v_)) goto panic;       \
        (dst) = (int64_t)((uint64_t)v_ - 1u);    \
    } while (0)
#define SH_SE(dst)                         \
    do {                                   \
        int64_t v_;                        \
        if (sh_golomb(b, &v_)) goto panic; \
        (dst) = sh_se(v_);                 \
    } while (0)
#define SH_FLAG(dst)                          \
    do {                                      \
        int64_t v_;                           \
        if (sh_field(b, 1, &v_)) goto panic;  \
        (dst) = (v_ == 1);                    \
    } while (0)

int orc_new_slice_header(const orc_sps *sps, const orc_pps *pps, int64_t nal_type, int64_t nal_ref_idc,
                         const uint8_t *rbsp, int64_t len, orc_slice_header *h) {
    orc_bit_reader br, *b = &br;
    memset(h, 0, sizeof(*h));
    orc_br_init(b, rbsp, len);
    const int idrPic = nal_type == 5;                                       /* :841-844 */
    h->ChromaArrayType = sps->UseSeparateColorPlane ? 0 : sps->ChromaFormat; /* :846-850 */
    SH_UE(h->FirstMbInSlice);                                                /* :858 */
    SH_UE(h->SliceType);                                                     /* :859 */
    const int st = slice_type_name(h->SliceType);                            /* :860 */
    SH_UE(h->PPSID);                                                         /* :862 */
    if (sps->UseSeparateColorPlane) {                                        /* :863-865 */
        if (sh_field(b, 2, &h->ColorPlaneID)) goto panic;
    }
    /* frame_num is not read (:866-867) */
    if (!sps->FrameMbsOnly) { /* :868-873 */
        SH_FLAG(h->FieldPic);
        if (h->FieldPic) SH_FLAG(h->BottomField);
    }
    if (idrPic) SH_UE(h->IDRPicID); /* :874-876 */
    if (sps->PicOrderCountType == 0) { /* :877-882 */
        if (sh_field(b, sps->Log2MaxPicOrderCntLSBMin4 + 4, &h->PicOrderCntLsb)) goto panic;
        if (pps->BottomFieldPicOrderInFramePresent && !h->FieldPic) SH_SE(h->DeltaPicOrderCntBottom);
    }
    if (sps->PicOrderCountType == 1 && !sps->DeltaPicOrderAlwaysZero) { /* :883-888 */
        SH_SE(h->DeltaPicOrderCnt0);
        if (pps->BottomFieldPicOrderInFramePresent && !h->FieldPic) SH_SE(h->DeltaPicOrderCnt1);
    }
    if (pps->RedundantPicCntPresent) SH_UE(h->RedundantPicCnt); /* :889-891 */
    if (st == ST_B) SH_FLAG(h->DirectSpatialMvPred);            /* :892-894 */
    if (st == ST_B || st == ST_SP || st == ST_B) {              /* :895-903 ("B" twice, no "P") */
        SH_FLAG(h->NumRefIdxActiveOverride);
        if (h->NumRefIdxActiveOverride) {
            SH_UE(h->NumRefIdxL0ActiveMinus1);
            if (st == ST_B) SH_UE(h->NumRefIdxL1ActiveMinus1);
        }
    }
    if (nal_type == 20 || nal_type == 21) { /* :905-908: nothing */
    } else {
        if (h->SliceType % 5 != 2 && h->SliceType % 5 != 4) { /* :911-924 */
            SH_FLAG(h->RefPicListModificationFlagL0);
            if (h->RefPicListModificationFlagL0) {
                while (h->ModificationOfPicNums != 3) {
                    SH_UE(h->ModificationOfPicNums);
                    if (h->ModificationOfPicNums == 0 || h->ModificationOfPicNums == 1)
                        SH_UE(h->AbsDiffPicNumMinus1);
                    else if (h->ModificationOfPicNums == 2)
                        SH_UE(h->LongTermPicNum);
                }
            }
        }
        if (h->SliceType % 5 == 1) { /* :925-937; ModificationOfPicNums keeps its value from list 0 */
            SH_FLAG(h->RefPicListModificationFlagL1);
            if (h->RefPicListModificationFlagL1) {
                while (h->ModificationOfPicNums != 3) {
                    SH_UE(h->ModificationOfPicNums);
                    if (h->ModificationOfPicNums == 0 || h->ModificationOfPicNums == 1)
                        SH_UE(h->AbsDiffPicNumMinus1);
                    else if (h->ModificationOfPicNums == 2)
                        SH_UE(h->LongTermPicNum);
                }
            }
        }
    }
    if ((pps->WeightedPred && (st == ST_P || st == ST_SP)) || (pps->WeightedBipred == 1 && st == ST_B)) { /* :942-990 */
        SH_UE(h->LumaLog2WeightDenom);
        if (h->ChromaArrayType != 0) SH_UE(h->ChromaLog2WeightDenom);
        for (int64_t i = 0; i <= h->NumRefIdxL0ActiveMinus1; i++) {
            int64_t f, v;
            SH_FLAG(f);
            if (f) {
                SH_SE(v); /* LumaWeightL0 = append(...) */
                SH_SE(v); /* LumaOffsetL0 */
                h->NLumaWeightL0++;
            }
            if (h->ChromaArrayType != 0) {
                SH_FLAG(f);
                if (f) {
                    h->NChromaWeightL0++;                 /* ChromaWeightL0 = append(ChromaWeightL0, []int{}) */
                    if (i >= h->NChromaWeightL0) goto panic; /* ChromaWeightL0[i]: index out of range (:964) */
                    for (int j = 0; j < 2; j++) {
                        SH_SE(v);
                        SH_SE(v);
                    }
                }
            }
        }
        if (h->SliceType % 5 == 1) {
            for (int64_t i = 0; i <= h->NumRefIdxL1ActiveMinus1; i++) {
                int64_t f, v;
                SH_FLAG(f);
                if (f) {
                    SH_SE(v);
                    SH_SE(v);
                    h->NLumaWeightL1++;
                }
                if (h->ChromaArrayType != 0) {
                    SH_FLAG(f);
                    if (f) {
                        h->NChromaWeightL1++;
                        if (i >= h->NChromaWeightL1) goto panic; /* :983 */
                        for (int j = 0; j < 2; j++) {
                            SH_SE(v);
                            SH_SE(v);
                        }
                    }
                }
            }
        }
    }
    if (nal_ref_idc != 0) { /* :991-1018 */
        if (idrPic) {
            SH_FLAG(h->NoOutputOfPriorPicsFlag);
            SH_FLAG(h->LongTermReferenceFlag);
        } else {
            SH_FLAG(h->AdaptiveRefPicMarkingModeFlag);
            if (h->AdaptiveRefPicMarkingModeFlag) {
                SH_UE(h->MemoryManagementControlOperation);
                const int64_t op = h->MemoryManagementControlOperation;
                if (op != 0 && !(op == 1 || op == 2 || op == 3 || op == 4 || op == 6)) {
                    h->bits_read = b->bitsRead; /* the loop body reads nothing and the operation never changes */
                    return ORC_HANG;
                }
                while (h->MemoryManagementControlOperation != 0) { /* never re-read: ends in the panic below */
                    if (op == 1 || op == 3) SH_UE(h->DifferenceOfPicNumsMinus1);
                    if (op == 2) SH_UE(h->LongTermPicNum);
                    if (op == 3 || op == 6) SH_UE(h->LongTermFrameIdx);
                    if (op == 4) SH_UE(h->MaxLongTermFrameIdxPlus1);
                }
            }
        }
    }
    if (pps->EntropyCodingMode == 1 && st != ST_I && st != ST_SI) SH_UE(h->CabacInit); /* :1019-1021 */
    SH_SE(h->SliceQpDelta);                                                             /* :1022 */
    if (st == ST_SP || st == ST_SI) {                                                   /* :1023-1028 */
        if (st == ST_SP) SH_FLAG(h->SpForSwitch);
        SH_SE(h->SliceQsDelta);
    }
    if (pps->DeblockingFilterControlPresent) { /* :1029-1035 */
        SH_UE(h->DisableDeblockingFilter);
        if (h->DisableDeblockingFilter != 1) {
            SH_SE(h->SliceAlphaC0OffsetDiv2);
            SH_SE(h->SliceBetaOffsetDiv2);
        }
    }
    if (pps->NumSliceGroupsMinus1 > 0 && pps->SliceGroupMapType >= 3 && pps->SliceGroupMapType <= 5) { /* :1036-1040 */
        if (pps->SliceGroupChangeRateMinus1 == 0) goto panic; /* integer divide by zero */
        const int64_t q = (int64_t)((uint64_t)(pps->PicSizeInMapUnitsMinus1 / pps->SliceGroupChangeRateMinus1) + 1u);
        /* int(math.Ceil(math.Log2(float64(q)))): NaN / -Inf for q <= 0 convert to the most negative int -> make panics */
        if (q <= 0) goto panic;
        const int64_t n = (int64_t)ceil(log2((double)q));
        if (sh_field(b, n, &h->SliceGroupChangeCycle)) goto panic;
    }
    h->SliceQPy = (int64_t)((uint64_t)26 + (uint64_t)pps->PicInitQpMinus26 + (uint64_t)h->SliceQpDelta); /* cabac.go:113-115 */
    h->bits_read = b->bitsRead;
    return ORC_OK;
panic:
    h->bits_read = b->bitsRead;
    return ORC_PANIC;
}

/* ================================================================================ syntax-element glue (rows I5 / f3)
 * Literal restatements, TEST INFRASTRUCTURE: CtxIdx (cabac.go:557-758), NewBinarization (:340-427), initCabac
 * (:148-174), binIdxMbMap / binIdxSubMbMap (:180-303), IsBinStringMatch (:429-436). */
#define NA_CTX_ID 10000

int64_t orc_ctx_idx(int64_t binIdx, int64_t maxBinIdxCtx, int64_t ctxIdxOffset) {
    int64_t ctxIdx = NA_CTX_ID;
    (void)maxBinIdxCtx;
    switch (ctxIdxOffset) {
        case 0:
            if (binIdx != 0) return NA_CTX_ID;
            break;
        case 3:
            switch (binIdx) {
                case 0: break;
                case 1: ctxIdx = 276; break;
                case 2: ctxIdx = 3; break;
                case 3: ctxIdx = 4; break;
                case 4: break;
                case 5: break;
                default: ctxIdx = 7;
            }
            break;
        case 11:
            if (binIdx != 0) return NA_CTX_ID;
            break;
        case 14:
            if (binIdx == 0) ctxIdx = 0;
            if (binIdx == 1) ctxIdx = 1;
            if (binIdx == 2) { /* 9.3.3.1.2 */ }
            if (binIdx > 2) return NA_CTX_ID;
            break;
        case 17:
            switch (binIdx) {
                case 0: ctxIdx = 0; break;
                case 1: ctxIdx = 276; break;
                case 2: ctxIdx = 1; break;
                case 3: ctxIdx = 2; break;
                case 4: break;
                default: ctxIdx = 3;
            }
            break;
        case 21:
            if (binIdx < 3) ctxIdx = binIdx;
            else return NA_CTX_ID;
            break;
        case 24:
            if (binIdx != 0) return NA_CTX_ID;
            break;
        case 27:
            switch (binIdx) {
                case 0: break;
                case 1: ctxIdx = 3; break;
                case 2: break;
                default: ctxIdx = 5;
            }
            break;
        case 32:
            switch (binIdx) {
                case 0: ctxIdx = 0; break;
                case 1: ctxIdx = 276; break;
                case 2: ctxIdx = 1; break;
                case 3: ctxIdx = 2; break;
                case 4: break;
                default: ctxIdx = 3;
            }
            break;
        case 36:
            if (binIdx == 0 || binIdx == 1) ctxIdx = binIdx;
            if (binIdx == 2) { /* 9.3.3.1.2 */ }
            if (binIdx > 2 && binIdx < 6) ctxIdx = 3;
            break;
        case 40: /* fallthrough */
        case 47:
            switch (binIdx) {
                case 0: break;
                case 1: ctxIdx = 3; break;
                case 2: ctxIdx = 4; break;
                case 3: ctxIdx = 5; break;
                default: ctxIdx = 6;
            }
            break;
        case 54:
            if (binIdx == 1) ctxIdx = 4;
            if (binIdx > 1) ctxIdx = 5;
            break;
        case 60:
            if (binIdx == 1) ctxIdx = 2;
            if (binIdx > 1) ctxIdx = 3;
            break;
        case 64:
            if (binIdx == 0) { /* 9.3.3.1.1.8 */
            } else if (binIdx == 1 || binIdx == 2) {
                ctxIdx = 3;
            } else {
                return NA_CTX_ID;
            }
            break;
        case 68:
            if (binIdx != 0) return NA_CTX_ID;
            ctxIdx = 0;
            break;
        case 69:
            if (binIdx >= 0 && binIdx < 3) ctxIdx = 0;
            return NA_CTX_ID; /* :714-718: unconditional */
        case 70:
            if (binIdx != 0) return NA_CTX_ID;
            break;
        case 73:
            switch (binIdx) {
                case 0: case 1: case 2: case 3: break;
                default: return NA_CTX_ID;
            }
            break;
        case 77:
            if (binIdx == 0) {
            } else if (binIdx == 1) {
            } else {
                return NA_CTX_ID;
            }
            break;
        case 276:
            if (binIdx != 0) return NA_CTX_ID;
            ctxIdx = 0;
            break;
        case 399:
            if (binIdx != 0) return NA_CTX_ID;
            break;
    }
    return ctxIdx;
}

/* NewBinarization(syntaxElement, data): se = index into the names below (anything else: no case), st = sliceTypeMap
 * name index (0 P, 1 B, 2 I, 3 SP, 4 SI).  out[16]: SyntaxElement, PrefixSuffix, FixedLength, Unary, TruncatedUnary, CMax,
 * UEGk, CMaxValue, MaxBinIdxCtx{IsPrefixSuffix, Prefix, Suffix}, CtxIdxOffset{IsPrefixSuffix, Prefix, Suffix},
 * UseDecodeBypass, 0 */
static const char *const se_names[] = {"CodedBlockPattern", "IntraChromaPredMode", "MbQpDelta", "MvdLnEnd0", "MvdLnEnd1",
                                       "MbType", "MbFieldDecodingFlag", "PrevIntra4x4PredModeFlag",
                                       "PrevIntra8x8PredModeFlag", "RefIdxL0", "RefIdxL1", "RemIntra4x4PredMode",
                                       "RemIntra8x8PredMode", "TransformSize8x8Flag"};
void orc_new_binarization(int32_t se, int32_t st, int32_t *out) {
    enum { SE, PS, FL, UN, TU, CM, UEGK, CMV, MAX_PS, MAX_P, MAX_S, OFF_PS, OFF_P, OFF_S, BYP };
    const char *name = (se >= 0 && se < 14) ? se_names[se] : "";
    memset(out, 0, 16 * sizeof(int32_t));
    out[SE] = se;
#define IS(s) (strcmp(name, s) == 0)
    if (IS("CodedBlockPattern")) {
        out[PS] = 1;
        out[MAX_PS] = 1, out[MAX_P] = 3, out[MAX_S] = 1;
        out[OFF_PS] = 1, out[OFF_P] = 73, out[OFF_S] = 77;
    } else if (IS("IntraChromaPredMode")) {
        out[TU] = 1, out[CM] = 1, out[CMV] = 3;
        out[MAX_P] = 1;
        out[OFF_P] = 64;
    } else if (IS("MbQpDelta")) {
        out[MAX_P] = 2;
        out[OFF_P] = 60;
    } else if (IS("MvdLnEnd0") || IS("MvdLnEnd1")) {
        out[BYP] = 1;
        out[UEGK] = 1;
        out[MAX_PS] = 1, out[MAX_P] = 4, out[MAX_S] = -1;
        out[OFF_PS] = 1, out[OFF_P] = IS("MvdLnEnd0") ? 40 : 47, out[OFF_S] = -1;
    } else if (IS("MbType")) {
        if (st == 4) { /* SI */
            out[PS] = 1;
            out[MAX_PS] = 1, out[MAX_P] = 0, out[MAX_S] = 6;
            out[OFF_PS] = 1, out[OFF_P] = 0, out[OFF_S] = 3;
        } else if (st == 2) { /* I */
            out[MAX_P] = 6;
            out[OFF_P] = 3;
        } else if (st == 3 || st == 0) { /* SP falls through to P */
            out[PS] = 1;
            out[MAX_PS] = 1, out[MAX_P] = 2, out[MAX_S] = 5;
            out[OFF_PS] = 1, out[OFF_P] = 14, out[OFF_S] = 17;
        }
    } else if (IS("MbFieldDecodingFlag")) {
        out[FL] = 1, out[CM] = 1, out[CMV] = 1;
        out[OFF_P] = 70;
    } else if (IS("PrevIntra4x4PredModeFlag") || IS("PrevIntra8x8PredModeFlag")) {
        out[FL] = 1, out[CM] = 1, out[CMV] = 1;
        out[OFF_P] = 68;
    } else if (IS("RefIdxL0") || IS("RefIdxL1")) {
        out[UN] = 1;
        out[MAX_P] = 2;
        out[OFF_P] = 54;
    } else if (IS("RemIntra4x4PredMode") || IS("RemIntra8x8PredMode")) {
        out[FL] = 1, out[CM] = 1, out[CMV] = 7;
        out[OFF_P] = 69;
    } else if (IS("TransformSize8x8Flag")) {
        out[FL] = 1, out[CM] = 1, out[CMV] = 1;
        out[OFF_P] = 399;
    }
#undef IS
}

/* initCabac: mn := MNVars[ctxIdx]; mn[0] -- the MNVars map only (keys 0..39), column cabac_init_idc 0 */
void orc_init_cabac(uint32_t flags, int64_t binIdx, int64_t maxPrefix, int64_t offPrefix, int64_t picInitQpMinus26,
                    int64_t sliceQpDelta, int64_t *pStateIdx, int64_t *valMPS, int64_t *ctxIdxOut) {
    int64_t ctxIdx = orc_ctx_idx(binIdx, maxPrefix, offPrefix);
    int64_t m = 0, n = 0;
    if (ctxIdx >= 0 && ctxIdx <= 39) orc_mn(flags, ctxIdx, 0, &m, &n);
    int64_t sliceQPy = (int64_t)((uint64_t)26 + (uint64_t)picInitQpMinus26 + (uint64_t)sliceQpDelta);
    int64_t pre = orc_pre_ctx_state(m, n, sliceQPy);
    if (pre <= 63) {
        *pStateIdx = 63 - pre;
        *valMPS = 0;
    } else {
        *pStateIdx = pre - 64;
        *valMPS = 1;
    }
    if (ctxIdxOut) *ctxIdxOut = ctxIdx;
}

/* binIdxMbMap / binIdxSubMbMap, written out as in the reference; returns the length, bits[k] = element k */
static const int8_t mb_I[26][8] = {
    {1, 0}, {6, 1, 0, 0, 0, 0, 0}, {6, 1, 0, 0, 0, 0, 1}, {6, 1, 0, 0, 0, 1, 0}, {6, 1, 0, 0, 0, 1, 1},
    {7, 1, 0, 0, 1, 0, 0, 0}, {7, 1, 0, 0, 1, 0, 0, 1}, {7, 1, 0, 0, 1, 0, 1, 0}, {7, 1, 0, 0, 1, 0, 1, 1},
    {7, 1, 0, 0, 1, 1, 0, 0}, {7, 1, 0, 0, 1, 1, 0, 1}, {7, 1, 0, 0, 1, 1, 1, 0}, {7, 1, 0, 0, 1, 1, 1, 1},
    {6, 1, 0, 1, 0, 0, 0}, {6, 1, 0, 1, 0, 0, 1}, {6, 1, 0, 1, 0, 1, 0}, {6, 1, 0, 1, 0, 1, 1},
    {7, 1, 0, 1, 1, 0, 0, 0}, {7, 1, 0, 1, 1, 0, 0, 1}, {7, 1, 0, 1, 1, 0, 1, 0}, {7, 1, 0, 1, 1, 0, 1, 1},
    {7, 1, 0, 1, 1, 1, 0, 0}, {7, 1, 0, 1, 1, 1, 0, 1}, {7, 1, 0, 1, 1, 1, 1, 0}, {7, 1, 0, 1, 1, 1, 1, 1},
    {2, 1, 1}};
int32_t orc_mb_bin_string(int32_t st, int64_t mbType, int32_t sub, int32_t *bits /* [8] */) {
    memset(bits, 0, 8 * sizeof(int32_t));
    if (sub) {
        if (!(st == 0 || st == 3)) return 0;
        switch (mbType) {
            case 0: bits[0] = 1; return 1;
            case 1: return 2;
            case 2: bits[1] = 1, bits[2] = 1; return 3;
            case 3: bits[1] = 1; return 3;
        }
        return 0;
    }
    if (st == 2) {
        if (mbType < 0 || mbType > 25) return 0;
        for (int k = 0; k < mb_I[mbType][0]; k++) bits[k] = mb_I[mbType][1 + k];
        return mb_I[mbType][0];
    }
    if (st == 0 || st == 3) {
        if (mbType < 0 || mbType > 30) return 0;
        switch (mbType) {
            case 0: return 3;
            case 1: bits[1] = 1, bits[2] = 1; return 3;
            case 2: bits[1] = 1; return 3;
            case 3: bits[2] = 1; return 3;
            case 4: return 0;
        }
        bits[0] = 1;
        return 1;
    }
    return 0;
}

/* IsBinStringMatch: 1 / 0, ORC_HANG + 0 = 2 for the index-out-of-range panic */
int32_t orc_bin_string_match(const int32_t *binString, int32_t len, const int32_t *bits, int32_t n) {
    for (int32_t i = 0; i < n; i++) {
        if (i >= len) return 2;
        if (binString[i] != bits[i]) return 0;
    }
    return len == n;
}

/* ======================================================================= mb_type as a syntax element (row f3)
 * The walk the reference sketches at h264/slice.go:639-672 -- NewBinarization("MbType") (cabac.go:340-427), then bin
 * after bin: CtxIdx(binIdx, MaxBinIdxCtx.Prefix, CtxIdxOffset.Prefix) (cabac.go:557-758), decode, IsBinStringMatch
 * against binIdxMbMap[sliceTypeName] (cabac.go:180-303, :429-436) -- composed with the engine the way the reference's
 * section comments point to (the reference itself reads raw bits there and never decodes: A6, A16).  Where CtxIdx
 * leaves a binIdx to a "9.3.3.1.x" comment and answers NaCtxId, the rule of that clause is used:
 *   I slices, ctxIdxOffset 3:  binIdx 0 -> ctxIdxInc = condTermFlagA + condTermFlagB (9.3.3.1.1.3) with A the macroblock
 *     decoded before this one in the slice (1 unless there is none or it was I_NxN) and B not available;
 *     binIdx 1 -> 276: DecodeTerminate; binIdx 4 -> b3 != 0 ? 5 : 6; binIdx 5 -> b3 != 0 ? 6 : 7 (9.3.3.1.2); the other
 *     increments are the values CtxIdx returns (Table 9-39 as the reference has it);
 *   P / SP slices, prefix ctxIdxOffset 14: binIdx 2 -> b1 != 1 ? 2 : 3; the prefix bin 1 (binIdxMbMap gives {1} for every
 *     mb_type 5..30) is followed by the I-slice bin string decoded with the suffix offset 17 (binIdx 4 -> b3 != 0 ? 2 : 3)
 *     and mb_type = 5 + that I-slice value (Table 9-37 of the standard).
 * A bin string no mb_type matches ends the walk (the reference's loop would spin): ORC_PANIC.  I_PCM (the terminate
 * bin of 1) ends the slice's walk after that element: pcm samples follow, not CABAC data. */
static int64_t mbt_ctx_inc(int64_t offset, int64_t binIdx, const int32_t *bits, int64_t prev_not_nxn) {
    int64_t v = orc_ctx_idx(binIdx, 0, offset);
    if (v != NA_CTX_ID) return v;
    if (offset == 3) {
        if (binIdx == 0) return prev_not_nxn;
        if (binIdx == 4) return bits[3] != 0 ? 5 : 6;
        if (binIdx == 5) return bits[3] != 0 ? 6 : 7;
    } else if (offset == 14) {
        if (binIdx == 2) return bits[1] != 1 ? 2 : 3;
    } else if (offset == 17) {
        if (binIdx == 4) return bits[3] != 0 ? 2 : 3;
    }
    return NA_CTX_ID;
}

/* one bin string against binIdxMbMap[st]: the mb_type whose string it is, -1: a proper prefix of some string,
 * -2: of none */
static int64_t mbt_match(int32_t st, const int32_t *bits, int32_t n, int64_t n_types) {
    int64_t prefix_of_some = 0;
    for (int64_t t = 0; t < n_types; t++) {
        int32_t bs[8];
        int32_t len = orc_mb_bin_string(st, t, 0, bs);
        if (len == 0) continue; /* nil / empty string: never matches a decoded bin */
        if (len == n && orc_bin_string_match(bs, len, bits, n) == 1) return t;
        if (len > n) {
            int same = 1;
            for (int32_t k = 0; k < n; k++) same = same && bs[k] == bits[k];
            prefix_of_some |= same;
        }
    }
    return prefix_of_some ? -1 : -2;
}

/* kind 0: I slice, 1: P / SP slice.  Decodes up to n_mb mb_type elements; out_types[k] receives element k.
 * fin->n_bins = bins decoded, *n_done = elements decoded.  Status as orc_cabac_decode_slice (a bin that runs off the
 * data leaves everything as it was before that bin). */
int orc_decode_mb_types(uint32_t flags, int32_t kind, const uint8_t *bytes, int64_t len, int64_t n_mb, uint8_t *ctx_state,
                        int64_t n_ctx, uint8_t *out_types, int64_t *n_done, orc_cabac_final *fin) {
    orc_bit_reader br;
    orc_br_init(&br, bytes, len);
    int64_t R, O, done = 0, bins = 0;
    int status = ORC_OK;
    orc_init_decoding_engine(&br, &R, &O);
    int64_t prev_not_nxn = 0;
    while (!br.panicked && done < n_mb) {
        int32_t bits[16];
        int32_t n = 0;
        int64_t offset = kind == 0 ? 3 : 14;
        int32_t st = kind == 0 ? 2 : 0; /* sliceTypeName I / P */
        int64_t base_type = 0, found = -1;
        int stop = 0;
        for (;;) {
            int64_t inc = mbt_ctx_inc(offset, n, bits, prev_not_nxn);
            int64_t bin = 0, R0 = R, O0 = O, bits0 = br.bitsRead;
            if (inc == 276) {
                orc_decode_terminate(&br, &R, &O, &bin);
            } else {
                int64_t ctx = offset + inc;
                if (inc == NA_CTX_ID || ctx >= n_ctx) ctx = 0;
                uint8_t s0 = ctx_state[ctx];
                orc_decode_decision(flags, &br, &ctx_state[ctx], &R, &O, &bin);
                if (br.panicked) ctx_state[ctx] = s0;
            }
            if (br.panicked) {
                R = R0, O = O0, br.bitsRead = bits0;
                break;
            }
            bins++;
            bits[n++] = (int32_t)bin;
            if (kind == 1 && offset == 14 && n == 1 && bin == 1) { /* the prefix {1}: an intra macroblock in a P slice */
                offset = 17, st = 2, base_type = 5, n = 0;
                continue;
            }
            int64_t t = mbt_match(st, bits, n, st == 2 ? 26 : 5);
            if (t >= 0) {
                found = base_type + t;
                break;
            }
            if (t == -2 || n >= 15) {
                status = ORC_PANIC;
                stop = 1;
                break;
            }
        }
        if (br.panicked || stop) break;
        out_types[done++] = (uint8_t)found;
        if (kind == 0) prev_not_nxn = found != 0;
        if (found == 25 + base_type && st == 2) break; /* I_PCM */
    }
    fin->codIRange = R;
    fin->codIOffset = O;
    fin->bitsRead = br.bitsRead;
    fin->flags = br.panicked ? 1u : 0u;
    fin->n_bins = (uint32_t)bins;
    *n_done = done;
    return br.panicked ? ORC_PANIC : status;
}
