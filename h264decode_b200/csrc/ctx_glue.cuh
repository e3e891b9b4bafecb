// ctx_glue.cuh -- rows I5 / f3 of SURVEY.md section 8: the scalar glue between syntax elements and the engine, as
// __host__ __device__ functions (one thread per query on the GPU, ctx_glue.cu; the same code compiled for the CPU by
// tests/native/hd_emul.cpp):
//   CtxIdx            h264/cabac.go:557-758  Table 9-39 as the reference has it: ctxIdx INCREMENTS / special values, not
//                                            offset + increment; its maxBinIdxCtx argument is never read; offset 69 always
//                                            answers NaCtxId; offset 21 hands a negative binIdx straight back
//   NewBinarization   h264/cabac.go:340-427  Table 9-34 rows (type flags, maxBinIdxCtx, ctxIdxOffset, bypass flag); MbType
//                                            by slice type, nothing for B slices; unknown names give the zero value
//   initCabac         h264/cabac.go:148-174  CtxIdx(binIdx, MaxBinIdxCtx.Prefix, CtxIdxOffset.Prefix) -> MNVars[ctxIdx][0]
//                                            (cabac_init_idc 0 whatever the slice says; MNVars only, i.e. ctxIdx 0..39) ->
//                                            PreCtxState with SliceQPy -> (pStateIdx, valMPS)
//   binIdxMbMap / binIdxSubMbMap / IsBinStringMatch   h264/cabac.go:180-303, :429-436   mb_type / sub_mb_type bin strings
// Unlike the reference's switch statements these are table driven: one row per ctxIdxOffset / syntax element.
#pragma once
#include <stdint.h>

#include "../../include/h264b200.h"

#ifndef H264B_HD
#if defined(__CUDACC__)
#define H264B_HD __host__ __device__ __forceinline__
#else
#define H264B_HD static inline
#endif
#endif

namespace h264b {

constexpr int64_t kNaCtxId = 10000;  // NaCtxId, cabac.go:4

// CtxIdx: per ctxIdxOffset the answers for binIdx 0..5, for binIdx >= 6 and for binIdx < 0 (N = NaCtxId; kSelf: binIdx)
H264B_HD int64_t ctx_idx_ref(int64_t bin_idx, int64_t /*max_bin_idx_ctx: never read, cabac.go:557*/, int64_t ctx_idx_offset) {
    constexpr int16_t N = 10000, kSelf = -1;
    struct Row {
        int16_t offset, v[6], ge6, neg;
    };
    const Row rows[] = {
        {3, {N, 276, 3, 4, N, N}, 7, 7},      {14, {0, 1, N, N, N, N}, N, N},   {17, {0, 276, 1, 2, N, 3}, 3, 3},
        {21, {0, 1, 2, N, N, N}, N, kSelf},   {27, {N, 3, N, 5, 5, 5}, 5, 5},   {32, {0, 276, 1, 2, N, 3}, 3, 3},
        {36, {0, 1, N, 3, 3, 3}, N, N},       {40, {N, 3, 4, 5, 6, 6}, 6, 6},   {47, {N, 3, 4, 5, 6, 6}, 6, 6},
        {54, {N, 4, 5, 5, 5, 5}, 5, N},       {60, {N, 2, 3, 3, 3, 3}, 3, N},   {64, {N, 3, 3, N, N, N}, N, N},
        {68, {0, N, N, N, N, N}, N, N},       {276, {0, N, N, N, N, N}, N, N},
    };  // offsets 0, 11, 24, 69, 70, 73, 77, 399 and every other value: NaCtxId whatever binIdx is
    for (const Row &r : rows) {
        if (r.offset != ctx_idx_offset) continue;
        const int16_t v = bin_idx < 0 ? r.neg : (bin_idx < 6 ? r.v[bin_idx] : r.ge6);
        return v == kSelf ? bin_idx : (int64_t)v;
    }
    return kNaCtxId;
}

enum {  // syntax element names of NewBinarization, in the order of its switch
    kSeCodedBlockPattern, kSeIntraChromaPredMode, kSeMbQpDelta, kSeMvdLnEnd0, kSeMvdLnEnd1, kSeMbType, kSeMbFieldDecodingFlag,
    kSePrevIntra4x4PredModeFlag, kSePrevIntra8x8PredModeFlag, kSeRefIdxL0, kSeRefIdxL1, kSeRemIntra4x4PredMode,
    kSeRemIntra8x8PredMode, kSeTransformSize8x8Flag, kSeCount
};

// slice_type_name: 0 P, 1 B, 2 I, 3 SP, 4 SI, anything else "" (sliceTypeMap, slice.go:105-116)
H264B_HD h264b_binarization new_binarization_ref(int32_t se, int32_t slice_type_name) {
    h264b_binarization b = {};
    b.syntax_element = se;
    // type flags: bit 0 PrefixSuffix, 1 FixedLength, 2 Unary, 3 TruncatedUnary, 4 CMax, 5 UEGk
    struct Row {
        uint8_t type;
        int16_t cmax_value;
        int8_t max_ps;
        int16_t max_prefix, max_suffix;
        int8_t off_ps;
        int16_t off_prefix, off_suffix;
        int8_t bypass;
    };
    const Row rows[kSeCount] = {
        {1, 0, 1, 3, 1, 1, 73, 77, 0},     // CodedBlockPattern
        {8 | 16, 3, 0, 1, 0, 0, 64, 0, 0}, // IntraChromaPredMode
        {0, 0, 0, 2, 0, 0, 60, 0, 0},      // MbQpDelta
        {32, 0, 1, 4, -1, 1, 40, -1, 1},   // MvdLnEnd0 (NA_SUFFIX = -1)
        {32, 0, 1, 4, -1, 1, 47, -1, 1},   // MvdLnEnd1
        {0, 0, 0, 0, 0, 0, 0, 0, 0},       // MbType: by slice type, below
        {2 | 16, 1, 0, 0, 0, 0, 70, 0, 0}, // MbFieldDecodingFlag
        {2 | 16, 1, 0, 0, 0, 0, 68, 0, 0}, // PrevIntra4x4PredModeFlag
        {2 | 16, 1, 0, 0, 0, 0, 68, 0, 0}, // PrevIntra8x8PredModeFlag
        {4, 0, 0, 2, 0, 0, 54, 0, 0},      // RefIdxL0
        {4, 0, 0, 2, 0, 0, 54, 0, 0},      // RefIdxL1
        {2 | 16, 7, 0, 0, 0, 0, 69, 0, 0}, // RemIntra4x4PredMode
        {2 | 16, 7, 0, 0, 0, 0, 69, 0, 0}, // RemIntra8x8PredMode
        {2 | 16, 1, 0, 0, 0, 0, 399, 0, 0} // TransformSize8x8Flag
    };
    if (se < 0 || se >= kSeCount) return b;  // e.g. "Transform8x8Flag" (slice.go:780): no case, the zero value
    Row r = rows[se];
    if (se == kSeMbType) {
        if (slice_type_name == 4) r = Row{1, 0, 1, 0, 6, 1, 0, 3, 0};                               // SI
        else if (slice_type_name == 2) r = Row{0, 0, 0, 6, 0, 0, 3, 0, 0};                          // I
        else if (slice_type_name == 0 || slice_type_name == 3) r = Row{1, 0, 1, 2, 5, 1, 14, 17, 0}; // P, SP
    }
    b.prefix_suffix = r.type & 1;
    b.fixed_length = (r.type >> 1) & 1;
    b.unary = (r.type >> 2) & 1;
    b.truncated_unary = (r.type >> 3) & 1;
    b.cmax = (r.type >> 4) & 1;
    b.uegk = (r.type >> 5) & 1;
    b.cmax_value = r.cmax_value;
    b.max_is_prefix_suffix = r.max_ps;
    b.max_prefix = r.max_prefix;
    b.max_suffix = r.max_suffix;
    b.off_is_prefix_suffix = r.off_ps;
    b.off_prefix = r.off_prefix;
    b.off_suffix = r.off_suffix;
    b.use_decode_bypass = r.bypass;
    return b;
}

// binIdxMbMap[sliceTypeName][mbType] / binIdxSubMbMap[sliceTypeName][subMbType]: length, bits (element k in bit k).
// A missing key is a nil slice: length 0 (so is P / SP mb_type 4, an empty literal).
H264B_HD void mb_bin_string_ref(int32_t slice_type_name, int64_t mb_type, bool sub, int32_t *len, uint32_t *bits) {
    *len = 0;
    *bits = 0;
    const bool p_like = slice_type_name == 0 || slice_type_name == 3;
    if (sub) {
        if (!p_like || mb_type < 0 || mb_type > 3) return;
        const uint8_t l[4] = {1, 2, 3, 3}, v[4] = {1, 0, 6, 2};  // {1} {0,0} {0,1,1} {0,1,0}
        *len = l[mb_type];
        *bits = v[mb_type];
        return;
    }
    if (slice_type_name == 2) {  // I: Table 9-36
        if (mb_type < 0 || mb_type > 25) return;
        if (mb_type == 0) {
            *len = 1;  // {0}
        } else if (mb_type == 25) {
            *len = 2, *bits = 3;  // {1,1}
        } else {
            // 1 0 <b2> <...>: types 1..12 start 1,0,0, types 13..24 start 1,0,1; inside a dozen: four of
            // (0, x, y), then eight of (1, x, y, z)
            const int64_t t = (mb_type - 1) % 12;
            uint32_t b = 1u | ((mb_type > 12 ? 1u : 0u) << 2);
            if (t < 4) {
                b |= 0u << 3 | (uint32_t)((t >> 1) & 1) << 4 | (uint32_t)(t & 1) << 5;
                *len = 6;
            } else {
                const int64_t s = t - 4;
                b |= 1u << 3 | (uint32_t)((s >> 2) & 1) << 4 | (uint32_t)((s >> 1) & 1) << 5 | (uint32_t)(s & 1) << 6;
                *len = 7;
            }
            *bits = b;
        }
        return;
    }
    if (p_like) {  // Table 9-37 as the reference has it
        if (mb_type < 0 || mb_type > 30) return;
        if (mb_type < 4) {
            const uint8_t v[4] = {0, 6, 2, 4};  // {0,0,0} {0,1,1} {0,1,0} {0,0,1}
            *len = 3;
            *bits = v[mb_type];
        } else if (mb_type > 4) {
            *len = 1, *bits = 1;  // {1}
        }
    }
}

// IsBinStringMatch (cabac.go:429-436): 1 match, 0 no match, 2 the reference indexes past the end of binString (panic)
H264B_HD int32_t bin_string_match_ref(int32_t len, uint32_t bin, int32_t n, uint32_t bits) {
    for (int32_t k = 0; k < n; k++) {
        if (k >= len) return 2;
        if (((bin >> k) & 1u) != ((bits >> k) & 1u)) return 0;
    }
    return len == n ? 1 : 0;
}

}  // namespace h264b
