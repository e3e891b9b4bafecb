/* oracle.h -- CPU restatement of the hot path of mrmod/h264decode (pure Go), in plain C.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under h264decode_b200/ (the product) may include, link or call
 * this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it,
 * and only as the checker / the timed CPU baseline.
 *
 * PARITY STATUS: "parity unpinned" by the reference itself.  The reference has no runnable test, no
 * golden vectors, does not build as shipped and no Go toolchain exists in this image (SURVEY.md §0, §4,
 * §8c), so this oracle cannot be checked against outputs of the reference.  It is pinned instead by the
 * hand-derived known-answer vectors of SURVEY.md Appendix B (tests/test_oracle_kat.py), which were
 * produced by an independent reading of the same source lines.
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference).  Go `int` is
 * restated as int64_t.  A Go runtime panic (index out of range on a slice) is restated as the status
 * ORC_PANIC -- a distinct, well-defined outcome (h264/server.go:136-143 turns it into exit(1)).
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_OK 0
#define ORC_PANIC 1   /* the Go code would have panicked (index out of range / nil slice write) */
#define ORC_CAPACITY 2 /* caller's output buffer too small (oracle-side condition, not a reference outcome) */

/* behaviour switches.  0 everywhere = the reference, literally. */
#define ORC_TABLES_SPEC 1u      /* A1..A3 corrected tables instead of the reference's */
#define ORC_BYPASS_SPEC_OR 2u   /* A5: codIOffset = (codIOffset<<1)|bit instead of the reference's <<1 then <<bit */

/* ---------------------------------------------------------------- bit reader (h264/bit_reader.go:11-17) */
typedef struct {
    const uint8_t *bytes;
    int64_t len;
    int64_t byteOffset, bitOffset, bitsRead;
    int panicked;
} orc_bit_reader;

void orc_br_init(orc_bit_reader *b, const uint8_t *bytes, int64_t len);
/* Read (bit_reader.go:292-314): fills buf[0..n) with single bits, MSB first. Sets panicked on overrun. */
int64_t orc_br_read(orc_bit_reader *b, int64_t *buf, int64_t n);
int64_t orc_br_next_field(orc_bit_reader *b, int64_t bits);          /* bit_reader.go:315-325 */
int64_t orc_br_read_one_bit(orc_bit_reader *b);                      /* bit_reader.go:232-236 */
int64_t orc_bit_val(const int64_t *bits, int64_t n);                 /* bit_reader.go:50-59 */
int64_t orc_br_golomb(orc_bit_reader *b, int64_t *bits, int64_t cap); /* bit_reader.go:174-196; returns nbits */
int64_t orc_ue(const int64_t *bits, int64_t n);                      /* bit_reader.go:62-64 */
int64_t orc_se(const int64_t *bits, int64_t n);                      /* bit_reader.go:158-161 (floor quirk A9) */
int orc_br_more_rbsp_data(orc_bit_reader *b);                        /* bit_reader.go:199-219 */
int orc_br_has_more_data(const orc_bit_reader *b);                   /* bit_reader.go:220-226 */

/* ---------------------------------------------------------------- NAL unit (h264/nalUnit.go:3-30) */
typedef struct {
    int64_t NumBytes, ForbiddenZeroBit, RefIdc, Type, SvcExtensionFlag, Avc3dExtensionFlag, IdrFlag, PriorityId,
        NoInterLayerPredFlag, DependencyId, QualityId, TemporalId, UseRefBasePicFlag, DiscardableFlag, OutputFlag,
        ReservedThree2Bits, HeaderBytes, NonIdrFlag, ViewId, AnchorPicFlag, InterViewFlag, ReservedOneBit, ViewIdx,
        DepthFlag;
    int64_t EmulationPreventionThreeByte;
    int64_t rbsp_len; /* len(nalUnit.rbsp) */
} orc_nal_unit;

/* isStartSequence (server.go:28-39) */
int orc_is_start_sequence(const uint8_t *packet, int64_t len);
/* NewNalUnit (nalUnit.go:75-131).  rbsp_out must hold frame_len bytes. */
int orc_new_nal_unit(const uint8_t *frame, int64_t frame_len, int64_t num_bytes_in_nal, orc_nal_unit *out,
                     uint8_t *rbsp_out);

typedef struct {
    int64_t start_offset, end_offset; /* server.go:88,103 */
    int64_t rbsp_off;                 /* offset of this NAL's rbsp in the caller's rbsp buffer */
    orc_nal_unit nal;
} orc_stream_nal;

/* The readNalUnit loop of handleConnection (server.go:64-111,144-146) over an in-memory stream, one byte per
 * BufferToReader call (bit_reader.go:27-39; the debug-file tee and the logging are not restated).  Stops at end of
 * stream exactly where the reference dies (server.go:69-71,95-97,147).  Returns the number of NAL units emitted, or
 * -ORC_CAPACITY.  `literal` != 0 keeps the per-byte append-to-growing-buffer access pattern of the Go code (the
 * timed "reference" behaviour); 0 scans the caller's buffer in place (same results). */
int64_t orc_read_nal_units(const uint8_t *stream, int64_t n, orc_stream_nal *out, int64_t cap, uint8_t *rbsp_buf,
                           int64_t rbsp_cap, int64_t *rbsp_total, int literal);

/* ---------------------------------------------------------------- CABAC engine (h264/cabac.go:439-553) */
void orc_init_decoding_engine(orc_bit_reader *b, int64_t *codIRange, int64_t *codIOffset); /* cabac.go:439-446 */
/* arithmetic core of BinaryDecision (cabac.go:525-536) with the context passed explicitly */
void orc_binary_decision(uint32_t flags, int64_t pStateIdx, int64_t valMPS, int64_t *codIRange, int64_t *codIOffset,
                         int64_t *binVal);
void orc_state_transition(uint32_t flags, int64_t *pStateIdx, int64_t *valMPS, int64_t binVal); /* cabac.go:544-553 */
void orc_renorm_d(orc_bit_reader *b, int64_t *codIRange, int64_t *codIOffset);                  /* cabac.go:503-511 */
void orc_decode_bypass(uint32_t flags, orc_bit_reader *b, int64_t codIRange, int64_t *codIOffset,
                       int64_t *binVal);                                                        /* cabac.go:468-481 */
void orc_decode_terminate(orc_bit_reader *b, int64_t *codIRange, int64_t *codIOffset, int64_t *binVal); /* :486-499 */
/* "DecodeDecision" = cabac.go:525-536 -> :544-553 -> :503-511 on a persistent context */
void orc_decode_decision(uint32_t flags, orc_bit_reader *b, uint8_t *ctx_state, int64_t *codIRange,
                         int64_t *codIOffset, int64_t *binVal);

/* op schedule entry: kind in bits 14..15 (0 decision, 1 bypass, 2 terminate), ctxIdx in bits 0..9 */
#define ORC_OP_DECISION 0u
#define ORC_OP_BYPASS 1u
#define ORC_OP_TERMINATE 2u
#define ORC_OP(kind, ctx) ((uint16_t)(((kind) << 14) | ((ctx)&0x3ffu)))

typedef struct {
    int64_t codIRange, codIOffset, bitsRead;
    uint32_t flags; /* bit 0: ORC_PANIC (read past the end of the slice bytes, A10) */
    uint32_t n_bins;
} orc_cabac_final;

/* Decode n_ops bins of one slice: initDecodingEngine then one primitive per op.  ctx_state[n_ctx] holds
 * pStateIdx | valMPS<<6 per context, updated in place.  bins_out gets 1 bit per bin, LSB-first in 32-bit words.
 * Decoding stops at the op that panics (the bin of that op is not produced). */
int orc_cabac_decode_slice(uint32_t flags, const uint8_t *bytes, int64_t len, const uint16_t *ops, int64_t n_ops,
                           uint8_t *ctx_state, int64_t n_ctx, uint32_t *bins_out, orc_cabac_final *fin);

/* ---------------------------------------------------------------- context init (cabac.go:113-174, mn_vars.go) */
int64_t orc_clip3(int64_t x, int64_t y, int64_t z);              /* cabac.go:131-139 */
int64_t orc_pre_ctx_state(int64_t m, int64_t n, int64_t sliceQPy); /* cabac.go:118-121 */
int64_t orc_slice_qpy(int64_t pic_init_qp_minus26, int64_t slice_qp_delta); /* cabac.go:113-115 */
/* (m,n) for ctxIdx and cabac_init_idc (-1 = NoCabacInitIdc / I,SI column): MNVars for 0..39 (mn_vars.go:15-175),
 * CodedblockPatternMN for 70..104 (:184-440), MN{0,0} otherwise. */
void orc_mn(uint32_t flags, int64_t ctxIdx, int64_t cabacInitIdc, int64_t *m, int64_t *n);
/* state split of initCabac (cabac.go:158-164): returns pStateIdx | valMPS<<6 */
uint8_t orc_ctx_state(int64_t preCtxState);
/* states[s*n_ctx + c] for every slice s with (qp[s], idc[s]) */
void orc_ctx_init(uint32_t flags, const int32_t *qp, const int32_t *idc, int64_t n_slices, int64_t n_ctx,
                  uint8_t *states);

/* ---------------------------------------------------------------- SPS / PPS (h264/sps.go, h264/pps.go) */
#define ORC_MAX_LIST 256
typedef struct {
    int64_t Profile, Constraint0, Constraint1, Constraint2, Constraint3, Constraint4, Constraint5, Level, ID,
        ChromaFormat, UseSeparateColorPlane, BitDepthLumaMinus8, BitDepthChromaMinus8, QPrimeYZeroTransformBypass,
        SeqScalingMatrixPresent, Log2MaxFrameNumMinus4, PicOrderCountType, Log2MaxPicOrderCntLSBMin4,
        DeltaPicOrderAlwaysZero, OffsetForNonRefPic, OffsetForTopToBottomField, NumRefFramesInPicOrderCntCycle,
        MaxNumRefFrames, GapsInFrameNumValueAllowed, PicWidthInMbsMinus1, PicHeightInMapUnitsMinus1, FrameMbsOnly,
        MBAdaptiveFrameField, Direct8x8Inference, FrameCropping, FrameCropLeftOffset, FrameCropRightOffset,
        FrameCropTopOffset, FrameCropBottomOffset, VuiParametersPresent, AspectRatioInfoPresent, AspectRatio, SarWidth,
        SarHeight, OverscanInfoPresent, OverscanAppropriate, VideoSignalTypePresent, VideoFormat, VideoFullRange,
        ColorDescriptionPresent, ColorPrimaries, TransferCharacteristics, MatrixCoefficients, ChromaLocInfoPresent,
        ChromaSampleLocTypeTopField, ChromaSampleLocTypeBottomField, CpbCntMinus1, BitRateScale, CpbSizeScale,
        InitialCpbRemovalDelayLengthMinus1, CpbRemovalDelayLengthMinus1, DpbOutputDelayLengthMinus1, TimeOffsetLength,
        TimingInfoPresent, NumUnitsInTick, TimeScale, NalHrdParametersPresent, FixedFrameRate, VclHrdParametersPresent,
        LowHrdDelay, PicStructPresent, BitstreamRestriction, MotionVectorsOverPicBoundaries, MaxBytesPerPicDenom,
        MaxBitsPerMbDenom, Log2MaxMvLengthHorizontal, Log2MaxMvLengthVertical, MaxDecFrameBuffering,
        MaxNumReorderFrames;
    int64_t n_SeqScalingList, SeqScalingList[12];
    int64_t n_OffsetForRefFrameList, OffsetForRefFrameList[ORC_MAX_LIST];
    int64_t n_hrd, BitRateValueMinus1[ORC_MAX_LIST], CpbSizeValueMinus1[ORC_MAX_LIST], Cbr[ORC_MAX_LIST];
    int64_t bits_read; /* BitReader.bitsRead when NewSPS returned (or panicked) */
} orc_sps;

typedef struct {
    int64_t ID, SPSID, EntropyCodingMode, NumSliceGroupsMinus1, BottomFieldPicOrderInFramePresent, SliceGroupMapType,
        SliceGroupChangeDirection, SliceGroupChangeRateMinus1, PicSizeInMapUnitsMinus1, NumRefIdxL0DefaultActiveMinus1,
        NumRefIdxL1DefaultActiveMinus1, WeightedPred, WeightedBipred, PicInitQpMinus26, PicInitQsMinus26,
        ChromaQpIndexOffset, DeblockingFilterControlPresent, ConstrainedIntraPred, RedundantPicCntPresent,
        Transform8x8Mode, PicScalingMatrixPresent, SecondChromaQpIndexOffset;
    int64_t bits_read;
} orc_pps;

/* ---- slice header: NewSliceContext up to (not including) NewSliceData, h264/slice.go:835-1048 ("next" row f1) ----
 * Literal restatement, quirks kept: frame_num is never read (:864-865); num_ref_idx_active_override is read for B and
 * SP slices only (:897 tests "B" twice, never "P"); ModificationOfPicNums is not reset between list 0 and list 1
 * (:917,:930); the memory_management_control_operation loop never reads the next operation (:1003-1016), so it either
 * reads until the Go code panics at the end of the data or -- for operations that read nothing -- never ends
 * (ORC_HANG); chroma weights index a slice that only grows when the flag is set (:960-966, panic when an earlier flag
 * was clear); se() has the floor quirk A9 and goes through float64 (math.Pow / math.Ceil, bit_reader.go:158-161);
 * SliceGroupChangeCycle divides by SliceGroupChangeRateMinus1 (panic when 0, :1033-1036). */
#define ORC_HANG 2    /* the Go code would loop forever */
typedef struct {
    int64_t FirstMbInSlice, SliceType, PPSID, ColorPlaneID, FieldPic, BottomField, IDRPicID, PicOrderCntLsb,
        DeltaPicOrderCntBottom, DeltaPicOrderCnt0, DeltaPicOrderCnt1, RedundantPicCnt, DirectSpatialMvPred,
        NumRefIdxActiveOverride, NumRefIdxL0ActiveMinus1, NumRefIdxL1ActiveMinus1, RefPicListModificationFlagL0,
        RefPicListModificationFlagL1, ModificationOfPicNums, AbsDiffPicNumMinus1, LongTermPicNum, LumaLog2WeightDenom,
        ChromaLog2WeightDenom, NLumaWeightL0, NChromaWeightL0, NLumaWeightL1, NChromaWeightL1, NoOutputOfPriorPicsFlag,
        LongTermReferenceFlag, AdaptiveRefPicMarkingModeFlag, MemoryManagementControlOperation,
        DifferenceOfPicNumsMinus1, LongTermFrameIdx, MaxLongTermFrameIdxPlus1, CabacInit, SliceQpDelta, SpForSwitch,
        SliceQsDelta, DisableDeblockingFilter, SliceAlphaC0OffsetDiv2, SliceBetaOffsetDiv2, SliceGroupChangeCycle,
        ChromaArrayType;
    int64_t SliceQPy;   /* cabac.go:113-115: 26 + PicInitQpMinus26 + SliceQpDelta */
    int64_t bits_read;  /* BitReader.bitsRead when the header ends: where slice_data() starts */
} orc_slice_header;
int orc_new_slice_header(const orc_sps *sps, const orc_pps *pps, int64_t nal_type, int64_t nal_ref_idc,
                         const uint8_t *rbsp, int64_t len, orc_slice_header *out);

int orc_new_sps(const uint8_t *rbsp, int64_t len, orc_sps *out);                         /* sps.go:192-437 */
int orc_new_pps(int64_t sps_chroma_format, const uint8_t *rbsp, int64_t len, orc_pps *out); /* pps.go:40-133 */

/* ---- syntax-element glue (rows I5 / f3): CtxIdx cabac.go:557-758, NewBinarization :340-427, initCabac :148-174,
 * binIdxMbMap / binIdxSubMbMap :180-303, IsBinStringMatch :429-436 */
int64_t orc_ctx_idx(int64_t binIdx, int64_t maxBinIdxCtx, int64_t ctxIdxOffset);
void orc_new_binarization(int32_t se, int32_t st, int32_t *out /* [16] */);
void orc_init_cabac(uint32_t flags, int64_t binIdx, int64_t maxPrefix, int64_t offPrefix, int64_t picInitQpMinus26,
                    int64_t sliceQpDelta, int64_t *pStateIdx, int64_t *valMPS, int64_t *ctxIdxOut);
int32_t orc_mb_bin_string(int32_t st, int64_t mbType, int32_t sub, int32_t *bits /* [8] */);
int32_t orc_bin_string_match(const int32_t *binString, int32_t len, const int32_t *bits, int32_t n);
/* mb_type as a syntax element: h264/slice.go:639-672, h264/cabac.go:180-303, :340-436, :557-758 */
int orc_decode_mb_types(uint32_t flags, int32_t kind, const uint8_t *bytes, int64_t len, int64_t n_mb, uint8_t *ctx_state,
                        int64_t n_ctx, uint8_t *out_types, int64_t *n_done, orc_cabac_final *fin);

#ifdef __cplusplus
}
#endif
#endif
