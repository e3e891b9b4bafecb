// param_sets.cuh -- NewSPS (h264/sps.go:192-437, scalingList :172-191) and NewPPS (h264/pps.go:40-133) as
// __host__ __device__ functions: one thread per parameter-set NAL on the GPU (param_sets.cu), the same code compiled
// for the CPU by tests/native/hd_emul.cpp and fuzzed against the oracle without a GPU.  Rows S1 / f4 of SURVEY.md §8.
//
// The walk reproduces the reference, deviations from ITU-T H.264 included:
//   * seq_parameter_set_id is read as ue(v) and chroma_format_idc is read for EVERY profile (sps.go:231-232);
//   * a present sequence scaling list i indexes DefaultScalingMatrix4x4[i] (2 rows) resp. DefaultScalingMatrix8x8[i-6]
//     (2 rows) before anything of the list is read: i in 2..5 and i >= 8 panic; the decoded scales go to package-level
//     lists nothing reads back, so only the bits consumed matter;
//   * aspect_ratio_idc is compared with 999, so the SAR fields are never read;
//   * the four *_length_minus1 / time_offset_length fields of hrd_parameters() are read inside the SchedSelIdx loop;
//   * max_num_reorder_frames is read before max_dec_frame_buffering;
//   * NewPPS writes slice-group and scaling-list tables through nil slices: slice_group_map_type 0, 2 and (with a
//     non-negative size) 6, and pic_scaling_matrix_present_flag = 1 panic;
//   * the 8x8 / scaling part of the PPS is entered when any BYTE is left (HasMoreData), and MoreRBSPData then consumes
//     bits up to and including the next 1 bit;
//   * se() is floor(codeNum / 2) with the sign of (-1)^(codeNum+1), evaluated through float64.
#pragma once
#include "slice_header.cuh"

namespace h264b {

#define PS_TRY(x)                       \
    do {                                \
        if (!(x)) {                     \
            o->bits_read = b.pos;       \
            return H264B_SH_PANIC;      \
        }                               \
    } while (0)

// scalingList, sps.go:172-191: only the bits consumed are observable
H264B_HD bool ps_scaling_list(ShBits &b, int size) {
    int64_t last = 8, next = 8;
    for (int i = 0; i < size; i++) {
        if (next != 0) {
            int64_t delta;
            if (!sh_se(b, &delta)) return false;
            next = (int64_t)((uint64_t)last + (uint64_t)delta + 256u) % 256;
        }
        last = next == 0 ? last : next;
    }
    return true;
}

// hrdParameters closure, sps.go:197-216
H264B_HD bool ps_hrd(ShBits &b, h264b_sps *o) {
    if (!sh_ue(b, &o->cpb_cnt_minus1)) return false;
    if (!sh_field(b, 4, &o->bit_rate_scale)) return false;
    if (!sh_field(b, 4, &o->cpb_size_scale)) return false;
    for (int64_t i = 0; i <= o->cpb_cnt_minus1; i++) {
        int64_t br, cs, cbr;
        if (!sh_ue(b, &br)) return false;
        if (o->n_hrd < H264B_SPS_MAX_HRD) o->bit_rate_value_minus1[o->n_hrd] = br;  // appended one by one
        if (!sh_ue(b, &cs)) return false;
        if (o->n_hrd < H264B_SPS_MAX_HRD) o->cpb_size_value_minus1[o->n_hrd] = cs;
        if (!sh_flag(b, &cbr)) return false;
        if (o->n_hrd < H264B_SPS_MAX_HRD) o->cbr[o->n_hrd] = cbr;
        o->n_hrd++;
        if (!sh_field(b, 5, &o->initial_cpb_removal_delay_length_minus1)) return false;
        if (!sh_field(b, 5, &o->cpb_removal_delay_length_minus1)) return false;
        if (!sh_field(b, 5, &o->dpb_output_delay_length_minus1)) return false;
        if (!sh_field(b, 5, &o->time_offset_length)) return false;
    }
    return true;
}

// NewSPS.  Returns H264B_SH_OK / H264B_SH_PANIC; *o holds every field read up to that point, o->bits_read the bits
// consumed.
H264B_HD uint32_t parse_sps(const uint8_t *rbsp, uint64_t len, h264b_sps *o) {
    ShBits b = {rbsp, len * 8u, 0, false};
    int64_t tmp;
    PS_TRY(sh_field(b, 8, &o->profile));
    PS_TRY(sh_field(b, 1, &o->constraint0));
    PS_TRY(sh_field(b, 1, &o->constraint1));
    PS_TRY(sh_field(b, 1, &o->constraint2));
    PS_TRY(sh_field(b, 1, &o->constraint3));
    PS_TRY(sh_field(b, 1, &o->constraint4));
    PS_TRY(sh_field(b, 1, &o->constraint5));
    PS_TRY(sh_field(b, 2, &tmp));  // reserved_zero_2bits
    PS_TRY(sh_field(b, 8, &o->level));
    PS_TRY(sh_ue(b, &o->id));
    PS_TRY(sh_ue(b, &o->chroma_format));
    const int64_t p = o->profile;
    if (p == 100 || p == 110 || p == 122 || p == 244 || p == 44 || p == 83 || p == 86 || p == 118 || p == 128 ||
        p == 138 || p == 139 || p == 134 || p == 135) {
        if (o->chroma_format == 3) PS_TRY(sh_flag(b, &o->use_separate_color_plane));
        PS_TRY(sh_ue(b, &o->bit_depth_luma_minus8));
        PS_TRY(sh_ue(b, &o->bit_depth_chroma_minus8));
        PS_TRY(sh_flag(b, &o->qprime_y_zero_transform_bypass));
        PS_TRY(sh_flag(b, &o->seq_scaling_matrix_present));
        if (o->seq_scaling_matrix_present) {
            const int max = o->chroma_format != 3 ? 8 : 12;
            for (int i = 0; i < max; i++) {
                int64_t present;
                PS_TRY(sh_flag(b, &present));
                o->seq_scaling_list[o->n_seq_scaling_list++] = present;
                if (present) {
                    if (i < 6 ? i >= 2 : i - 6 >= 2) {  // DefaultScalingMatrix4x4[i] / 8x8[i-6]: index out of range
                        o->bits_read = b.pos;
                        return H264B_SH_PANIC;
                    }
                    PS_TRY(ps_scaling_list(b, i < 6 ? 16 : 64));
                }
            }
        }
    }
    PS_TRY(sh_ue(b, &o->log2_max_frame_num_minus4));
    PS_TRY(sh_ue(b, &o->pic_order_count_type));
    if (o->pic_order_count_type == 0) {
        PS_TRY(sh_ue(b, &o->log2_max_pic_order_cnt_lsb_min4));
    } else if (o->pic_order_count_type == 1) {
        PS_TRY(sh_flag(b, &o->delta_pic_order_always_zero));
        PS_TRY(sh_se(b, &o->offset_for_non_ref_pic));
        PS_TRY(sh_se(b, &o->offset_for_top_to_bottom_field));
        PS_TRY(sh_ue(b, &o->num_ref_frames_in_pic_order_cnt_cycle));
        for (int64_t i = 0; i < o->num_ref_frames_in_pic_order_cnt_cycle; i++) {
            int64_t v;
            PS_TRY(sh_se(b, &v));
            if (o->n_offset_for_ref_frame < H264B_SPS_MAX_REF_FRAMES) o->offset_for_ref_frame[o->n_offset_for_ref_frame] = v;
            o->n_offset_for_ref_frame++;
        }
    }
    PS_TRY(sh_ue(b, &o->max_num_ref_frames));
    PS_TRY(sh_flag(b, &o->gaps_in_frame_num_value_allowed));
    PS_TRY(sh_ue(b, &o->pic_width_in_mbs_minus1));
    PS_TRY(sh_ue(b, &o->pic_height_in_map_units_minus1));
    PS_TRY(sh_flag(b, &o->frame_mbs_only));
    if (!o->frame_mbs_only) PS_TRY(sh_flag(b, &o->mb_adaptive_frame_field));
    PS_TRY(sh_flag(b, &o->direct_8x8_inference));
    PS_TRY(sh_flag(b, &o->frame_cropping));
    if (o->frame_cropping) {
        PS_TRY(sh_ue(b, &o->frame_crop_left_offset));
        PS_TRY(sh_ue(b, &o->frame_crop_right_offset));
        PS_TRY(sh_ue(b, &o->frame_crop_top_offset));
        PS_TRY(sh_ue(b, &o->frame_crop_bottom_offset));
    }
    PS_TRY(sh_flag(b, &o->vui_parameters_present));
    if (o->vui_parameters_present) {
        PS_TRY(sh_flag(b, &o->aspect_ratio_info_present));
        if (o->aspect_ratio_info_present) {
            PS_TRY(sh_field(b, 8, &o->aspect_ratio));
            if (o->aspect_ratio == 999) {  // EXTENDED_SAR := 999 (sps.go:347): never equal to an 8-bit field
                PS_TRY(sh_field(b, 16, &o->sar_width));
                PS_TRY(sh_field(b, 16, &o->sar_height));
            }
        }
        PS_TRY(sh_flag(b, &o->overscan_info_present));
        if (o->overscan_info_present) PS_TRY(sh_flag(b, &o->overscan_appropriate));
        PS_TRY(sh_flag(b, &o->video_signal_type_present));
        if (o->video_signal_type_present) {
            PS_TRY(sh_field(b, 3, &o->video_format));
            PS_TRY(sh_flag(b, &o->video_full_range));
            PS_TRY(sh_flag(b, &o->color_description_present));
            if (o->color_description_present) {
                PS_TRY(sh_field(b, 8, &o->color_primaries));
                PS_TRY(sh_field(b, 8, &o->transfer_characteristics));
                PS_TRY(sh_field(b, 8, &o->matrix_coefficients));
            }
        }
        PS_TRY(sh_flag(b, &o->chroma_loc_info_present));
        if (o->chroma_loc_info_present) {
            PS_TRY(sh_ue(b, &o->chroma_sample_loc_type_top_field));
            PS_TRY(sh_ue(b, &o->chroma_sample_loc_type_bottom_field));
        }
        PS_TRY(sh_flag(b, &o->timing_info_present));
        if (o->timing_info_present) {
            PS_TRY(sh_field(b, 32, &o->num_units_in_tick));
            PS_TRY(sh_field(b, 32, &o->time_scale));
            PS_TRY(sh_flag(b, &o->fixed_frame_rate));
        }
        PS_TRY(sh_flag(b, &o->nal_hrd_parameters_present));
        if (o->nal_hrd_parameters_present) PS_TRY(ps_hrd(b, o));
        PS_TRY(sh_flag(b, &o->vcl_hrd_parameters_present));
        if (o->vcl_hrd_parameters_present) PS_TRY(ps_hrd(b, o));
        if (o->nal_hrd_parameters_present || o->vcl_hrd_parameters_present) PS_TRY(sh_flag(b, &o->low_hrd_delay));
        PS_TRY(sh_flag(b, &o->pic_struct_present));
        PS_TRY(sh_flag(b, &o->bitstream_restriction));
        if (o->bitstream_restriction) {
            PS_TRY(sh_flag(b, &o->motion_vectors_over_pic_boundaries));
            PS_TRY(sh_ue(b, &o->max_bytes_per_pic_denom));
            PS_TRY(sh_ue(b, &o->max_bits_per_mb_denom));
            PS_TRY(sh_ue(b, &o->log2_max_mv_length_horizontal));
            PS_TRY(sh_ue(b, &o->log2_max_mv_length_vertical));
            PS_TRY(sh_ue(b, &o->max_num_reorder_frames));
            PS_TRY(sh_ue(b, &o->max_dec_frame_buffering));
        }
    }
    o->bits_read = b.pos;
    return H264B_SH_OK;
}

// NewPPS.  (Its *SPS argument is only read on a path that has panicked before, pps.go:99-103.)
H264B_HD uint32_t parse_pps(const uint8_t *rbsp, uint64_t len, h264b_pps *o) {
    ShBits b = {rbsp, len * 8u, 0, false};
    PS_TRY(sh_ue(b, &o->id));
    PS_TRY(sh_ue(b, &o->sps_id));
    PS_TRY(sh_field(b, 1, &o->entropy_coding_mode));
    PS_TRY(sh_flag(b, &o->bottom_field_pic_order_in_frame_present));
    PS_TRY(sh_ue(b, &o->num_slice_groups_minus1));
    if (o->num_slice_groups_minus1 > 0) {
        PS_TRY(sh_ue(b, &o->slice_group_map_type));
        const int64_t t = o->slice_group_map_type;
        if (t == 0 || t == 2) {  // RunLengthMinus1[0] / TopLeft[0] of a nil slice (pps.go:61,65)
            o->bits_read = b.pos;
            return H264B_SH_PANIC;
        } else if (t > 2 && t < 6) {
            PS_TRY(sh_flag(b, &o->slice_group_change_direction));
            PS_TRY(sh_ue(b, &o->slice_group_change_rate_minus1));
        } else if (t == 6) {
            PS_TRY(sh_ue(b, &o->pic_size_in_map_units_minus1));
            if (o->pic_size_in_map_units_minus1 >= 0) {  // SliceGroupId[0] of a nil slice (pps.go:74)
                o->bits_read = b.pos;
                return H264B_SH_PANIC;
            }
        }
    }
    PS_TRY(sh_ue(b, &o->num_ref_idx_l0_default_active_minus1));
    PS_TRY(sh_ue(b, &o->num_ref_idx_l1_default_active_minus1));
    PS_TRY(sh_flag(b, &o->weighted_pred));
    PS_TRY(sh_field(b, 2, &o->weighted_bipred));
    PS_TRY(sh_se(b, &o->pic_init_qp_minus26));
    PS_TRY(sh_se(b, &o->pic_init_qs_minus26));
    PS_TRY(sh_se(b, &o->chroma_qp_index_offset));
    PS_TRY(sh_flag(b, &o->deblocking_filter_control_present));
    PS_TRY(sh_flag(b, &o->constrained_intra_pred));
    PS_TRY(sh_flag(b, &o->redundant_pic_cnt_present));
    if (len > (b.pos >> 3)) {  // HasMoreData, bit_reader.go:220-226: byte granular
        PS_TRY(sh_field(b, 1, &o->transform_8x8_mode));
        PS_TRY(sh_flag(b, &o->pic_scaling_matrix_present));
        if (o->pic_scaling_matrix_present) {  // PicScalingListPresent[0] of a nil slice (pps.go:103), flag read first
            int64_t f;
            PS_TRY(sh_flag(b, &f));
            o->bits_read = b.pos;
            return H264B_SH_PANIC;
        }
        if (len > (b.pos >> 3)) {  // MoreRBSPData, bit_reader.go:199-219: up to and including the next 1 bit
            uint32_t bit = 0;
            while (bit != 1) PS_TRY(sh_bit(b, &bit));
        }
    }
    o->bits_read = b.pos;
    return H264B_SH_OK;
}

// What the slice-header walk reads from the active parameter sets (slice.go:835-1048).
H264B_HD h264b_param_sets make_param_sets(const h264b_sps &s, const h264b_pps &p) {
    h264b_param_sets r;
    r.use_separate_color_plane = s.use_separate_color_plane;
    r.chroma_format = s.chroma_format;
    r.frame_mbs_only = s.frame_mbs_only;
    r.pic_order_count_type = s.pic_order_count_type;
    r.log2_max_pic_order_cnt_lsb_min4 = s.log2_max_pic_order_cnt_lsb_min4;
    r.delta_pic_order_always_zero = s.delta_pic_order_always_zero;
    r.bottom_field_pic_order_in_frame_present = p.bottom_field_pic_order_in_frame_present;
    r.redundant_pic_cnt_present = p.redundant_pic_cnt_present;
    r.weighted_pred = p.weighted_pred;
    r.weighted_bipred = p.weighted_bipred;
    r.entropy_coding_mode = p.entropy_coding_mode;
    r.deblocking_filter_control_present = p.deblocking_filter_control_present;
    r.num_slice_groups_minus1 = p.num_slice_groups_minus1;
    r.slice_group_map_type = p.slice_group_map_type;
    r.pic_size_in_map_units_minus1 = p.pic_size_in_map_units_minus1;
    r.slice_group_change_rate_minus1 = p.slice_group_change_rate_minus1;
    r.pic_init_qp_minus26 = p.pic_init_qp_minus26;
    r.reserved = 0;
    return r;
}

}  // namespace h264b
