// slice_header.cuh -- the slice-header walk of NewSliceContext (h264/slice.go:835-1048, up to NewSliceData) as one
// __host__ __device__ function: one thread per slice NAL on the GPU (slice_header_kernel), and the same code compiled
// for the CPU by tests/native/hd_emul.cpp so that it can be fuzzed against the oracle without a GPU.
//
// What the CABAC stage needs from it: SliceQPY (cabac.go:113-115), cabac_init_idc, the slice type and the bit offset
// at which slice_data() starts.  The walk itself reproduces the reference, deviations from ITU-T H.264 included:
// frame_num is not read; num_ref_idx_active_override_flag is read for B and SP slices but not for P; the
// modification_of_pic_nums_idc loop of list 1 starts from list 0's last value; the memory_management_control_operation
// loop never fetches the next operation (it runs into the end of the data, or never ends: H264B_SH_HANG); chroma weight
// tables index a list that only grows when the flag is set; se() uses floor(codeNum/2) and float64 rounding for
// |codeNum| >= 2^53; SliceGroupChangeCycle divides by slice_group_change_rate_minus1.
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

struct ShBits {  // MSB-first cursor over rbsp[0 .. len); reading past the end is the reference's panic
    const uint8_t *p;
    uint64_t len_bits, pos;
    bool panicked;
};
H264B_HD bool sh_bit(ShBits &b, uint32_t *bit) {
    if (b.pos >= b.len_bits) {
        b.panicked = true;
        return false;
    }
    *bit = (b.p[b.pos >> 3] >> (7u - (uint32_t)(b.pos & 7u))) & 1u;
    b.pos++;
    return true;
}
// golomb() + bitVal(): the whole code word modulo 2^64 (bit_reader.go:174-196, :50-59)
H264B_HD bool sh_golomb(ShBits &b, uint64_t *val) {
    uint64_t t = 0, zeros = 0;
    uint32_t bit = 0;
    for (;;) {
        if (!sh_bit(b, &bit)) return false;
        t = (t << 1) | bit;
        if (bit) break;
        zeros++;
    }
    for (uint64_t i = 0; i < zeros; i++) {
        if (!sh_bit(b, &bit)) return false;
        t = (t << 1) | bit;
    }
    *val = t;
    return true;
}
H264B_HD bool sh_ue(ShBits &b, int64_t *v) {  // ue(), bit_reader.go:62-64
    uint64_t t;
    if (!sh_golomb(b, &t)) return false;
    *v = (int64_t)(t - 1u);
    return true;
}
// se(), bit_reader.go:158-161: int(math.Pow(-1, float64(codeNum+1)) * math.Ceil(float64(codeNum/2))).  The two
// int64 -> float64 conversions round to nearest even above 2^53; (-1)^y is +1 for every y that large.
H264B_HD bool sh_se(ShBits &b, int64_t *v) {
    uint64_t t;
    if (!sh_golomb(b, &t)) return false;
    const int64_t code = (int64_t)(t - 1u);
    const int64_t y = (int64_t)((uint64_t)code + 1u);
    const double yd = (double)y;
    const bool odd = (yd > -9007199254740992.0 && yd < 9007199254740992.0) ? ((y & 1) != 0) : false;
    const double mag = (double)(code / 2);  // integral, so math.Ceil changes nothing
    *v = (int64_t)(odd ? -mag : mag);
    return true;
}
H264B_HD bool sh_field(ShBits &b, int64_t n, int64_t *v) {  // NextField, bit_reader.go:315-325
    if (n < 0) {
        b.panicked = true;  // make([]int, n) with n < 0
        return false;
    }
    uint64_t t = 0;
    uint32_t bit;
    for (int64_t i = 0; i < n; i++) {
        if (!sh_bit(b, &bit)) return false;
        t = (t << 1) | bit;
    }
    *v = (int64_t)t;
    return true;
}
H264B_HD bool sh_flag(ShBits &b, int64_t *v) {
    uint32_t bit;
    if (!sh_bit(b, &bit)) return false;
    *v = bit;
    return true;
}

enum { kStP = 0, kStB = 1, kStI = 2, kStSP = 3, kStSI = 4, kStNone = 5 };  // sliceTypeMap, slice.go:105-116

#define SH_TRY(x)            \
    do {                     \
        if (!(x)) return st; \
    } while (0)

// Returns H264B_SH_OK / H264B_SH_PANIC / H264B_SH_HANG; *h holds every field read up to that point.
H264B_HD uint32_t parse_slice_header(const h264b_param_sets &ps, uint32_t nal_type, uint32_t nal_ref_idc,
                                     const uint8_t *rbsp, uint64_t len, h264b_slice_header *h) {
    ShBits b = {rbsp, len * 8u, 0, false};
    uint32_t st = H264B_SH_PANIC;  // what an early return means
    const bool idr = nal_type == 5;
    h->chroma_array_type = ps.use_separate_color_plane ? 0 : ps.chroma_format;
    SH_TRY(sh_ue(b, &h->first_mb_in_slice));
    SH_TRY(sh_ue(b, &h->slice_type));
    const int name = (h->slice_type >= 0 && h->slice_type <= 9) ? (int)(h->slice_type % 5) : kStNone;
    SH_TRY(sh_ue(b, &h->pps_id));
    if (ps.use_separate_color_plane) SH_TRY(sh_field(b, 2, &h->color_plane_id));
    if (!ps.frame_mbs_only) {
        SH_TRY(sh_flag(b, &h->field_pic));
        if (h->field_pic) SH_TRY(sh_flag(b, &h->bottom_field));
    }
    if (idr) SH_TRY(sh_ue(b, &h->idr_pic_id));
    if (ps.pic_order_count_type == 0) {
        SH_TRY(sh_field(b, (int64_t)ps.log2_max_pic_order_cnt_lsb_min4 + 4, &h->pic_order_cnt_lsb));
        if (ps.bottom_field_pic_order_in_frame_present && !h->field_pic) SH_TRY(sh_se(b, &h->delta_pic_order_cnt_bottom));
    }
    if (ps.pic_order_count_type == 1 && !ps.delta_pic_order_always_zero) {
        SH_TRY(sh_se(b, &h->delta_pic_order_cnt[0]));
        if (ps.bottom_field_pic_order_in_frame_present && !h->field_pic) SH_TRY(sh_se(b, &h->delta_pic_order_cnt[1]));
    }
    if (ps.redundant_pic_cnt_present) SH_TRY(sh_ue(b, &h->redundant_pic_cnt));
    if (name == kStB) SH_TRY(sh_flag(b, &h->direct_spatial_mv_pred));
    if (name == kStB || name == kStSP) {
        SH_TRY(sh_flag(b, &h->num_ref_idx_active_override));
        if (h->num_ref_idx_active_override) {
            SH_TRY(sh_ue(b, &h->num_ref_idx_l0_active_minus1));
            if (name == kStB) SH_TRY(sh_ue(b, &h->num_ref_idx_l1_active_minus1));
        }
    }
    if (!(nal_type == 20 || nal_type == 21)) {
        for (int list = 0; list < 2; list++) {
            const int64_t m5 = h->slice_type % 5;
            if (list == 0 ? !(m5 != 2 && m5 != 4) : m5 != 1) continue;
            int64_t *flag = list == 0 ? &h->ref_pic_list_modification_flag_l0 : &h->ref_pic_list_modification_flag_l1;
            SH_TRY(sh_flag(b, flag));
            if (*flag) {
                while (h->modification_of_pic_nums != 3) {
                    SH_TRY(sh_ue(b, &h->modification_of_pic_nums));
                    if (h->modification_of_pic_nums == 0 || h->modification_of_pic_nums == 1)
                        SH_TRY(sh_ue(b, &h->abs_diff_pic_num_minus1));
                    else if (h->modification_of_pic_nums == 2)
                        SH_TRY(sh_ue(b, &h->long_term_pic_num));
                }
            }
        }
    }
    if ((ps.weighted_pred && (name == kStP || name == kStSP)) || (ps.weighted_bipred == 1 && name == kStB)) {
        SH_TRY(sh_ue(b, &h->luma_log2_weight_denom));
        if (h->chroma_array_type != 0) SH_TRY(sh_ue(b, &h->chroma_log2_weight_denom));
        for (int list = 0; list < 2; list++) {
            if (list == 1 && h->slice_type % 5 != 1) break;
            const int64_t last = list == 0 ? h->num_ref_idx_l0_active_minus1 : h->num_ref_idx_l1_active_minus1;
            int64_t *n_luma = list == 0 ? &h->n_luma_weight_l0 : &h->n_luma_weight_l1;
            int64_t *n_chroma = list == 0 ? &h->n_chroma_weight_l0 : &h->n_chroma_weight_l1;
            for (int64_t i = 0; i <= last; i++) {
                int64_t f, v;
                SH_TRY(sh_flag(b, &f));
                if (f) {
                    SH_TRY(sh_se(b, &v));
                    SH_TRY(sh_se(b, &v));
                    (*n_luma)++;
                }
                if (h->chroma_array_type != 0) {
                    SH_TRY(sh_flag(b, &f));
                    if (f) {
                        (*n_chroma)++;
                        if (i >= *n_chroma) return st;  // the reference indexes past the end of its list: panic
                        for (int j = 0; j < 4; j++) SH_TRY(sh_se(b, &v));
                    }
                }
            }
        }
    }
    if (nal_ref_idc != 0) {
        if (idr) {
            SH_TRY(sh_flag(b, &h->no_output_of_prior_pics_flag));
            SH_TRY(sh_flag(b, &h->long_term_reference_flag));
        } else {
            SH_TRY(sh_flag(b, &h->adaptive_ref_pic_marking_mode_flag));
            if (h->adaptive_ref_pic_marking_mode_flag) {
                SH_TRY(sh_ue(b, &h->memory_management_control_operation));
                const int64_t op = h->memory_management_control_operation;
                if (op != 0) {
                    if (!(op == 1 || op == 2 || op == 3 || op == 4 || op == 6)) {
                        h->header_bits = b.pos;
                        return H264B_SH_HANG;
                    }
                    for (;;) {  // the operation is never read again: this ends at the end of the data
                        if (op == 1 || op == 3) SH_TRY(sh_ue(b, &h->difference_of_pic_nums_minus1));
                        if (op == 2) SH_TRY(sh_ue(b, &h->long_term_pic_num));
                        if (op == 3 || op == 6) SH_TRY(sh_ue(b, &h->long_term_frame_idx));
                        if (op == 4) SH_TRY(sh_ue(b, &h->max_long_term_frame_idx_plus1));
                    }
                }
            }
        }
    }
    if (ps.entropy_coding_mode == 1 && name != kStI && name != kStSI) SH_TRY(sh_ue(b, &h->cabac_init_idc));
    SH_TRY(sh_se(b, &h->slice_qp_delta));
    if (name == kStSP || name == kStSI) {
        if (name == kStSP) SH_TRY(sh_flag(b, &h->sp_for_switch));
        SH_TRY(sh_se(b, &h->slice_qs_delta));
    }
    if (ps.deblocking_filter_control_present) {
        SH_TRY(sh_ue(b, &h->disable_deblocking_filter));
        if (h->disable_deblocking_filter != 1) {
            SH_TRY(sh_se(b, &h->slice_alpha_c0_offset_div2));
            SH_TRY(sh_se(b, &h->slice_beta_offset_div2));
        }
    }
    if (ps.num_slice_groups_minus1 > 0 && ps.slice_group_map_type >= 3 && ps.slice_group_map_type <= 5) {
        if (ps.slice_group_change_rate_minus1 == 0) return st;  // integer divide by zero
        const int64_t q = (int64_t)ps.pic_size_in_map_units_minus1 / (int64_t)ps.slice_group_change_rate_minus1 + 1;
        if (q <= 0) return st;  // Log2 of a non-positive number -> negative length -> make() panics
        int64_t n = 0;          // ceil(log2(q)), q < 2^33
        while (((int64_t)1 << n) < q) n++;
        SH_TRY(sh_field(b, n, &h->slice_group_change_cycle));
    }
    h->slice_qp_y = (int64_t)((uint64_t)26 + (uint64_t)(int64_t)ps.pic_init_qp_minus26 + (uint64_t)h->slice_qp_delta);
    h->header_bits = b.pos;
    return H264B_SH_OK;
}

// early returns above leave the position in `b`, which is local: the wrapper records it
H264B_HD void parse_slice_header_record(const h264b_param_sets &ps, uint32_t nal_type, uint32_t nal_ref_idc,
                                        const uint8_t *rbsp, uint64_t len, h264b_slice_header *h) {
    *h = h264b_slice_header{};
    h->status = parse_slice_header(ps, nal_type, nal_ref_idc, rbsp, len, h);
}

}  // namespace h264b
