"""ctypes binding of include/h264b200.h -- one Python method per exported symbol, numpy in / numpy out.

No compute happens here.  If libh264b200.so is missing the import fails (build it with
`python -m h264decode_b200.build`); if no CUDA device is present Context() raises H264BError(NO_DEVICE).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# h264b_scheduler runs several launches side by side on one device: more hardware queues than CUDA's default of 8, so that
# no two of its streams share one (read when the process initialises CUDA, which importing this module does not do)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
LIB_PATH = os.environ.get("H264B_LIB") or os.path.join(_HERE, "libh264b200.so")  # (H264B_LIB: experiment builds)

OK, E_INVALID, E_CUDA, E_NOMEM, E_CAPACITY, E_NO_DEVICE = range(6)
TABLES_SPEC, BYPASS_SPEC_OR, CABAC_FINAL_TERMINATE, STREAM_WANT_RBSP, STREAM_SLICE_HEADERS = 1, 2, 4, 8, 16
STREAM_PARAM_SETS = 32
F_OVERRUN, F_HAS_EPB, F_SHORT_NAL = 1, 2, 4
OP_DECISION, OP_BYPASS, OP_TERMINATE = 0, 1, 2

# every symbol include/h264b200.h declares (tests check that the library exports exactly these)
SYMBOLS = [
    "h264b_version", "h264b_device_count", "h264b_create", "h264b_destroy", "h264b_last_error", "h264b_set_stream",
    "h264b_sync", "h264b_host_alloc", "h264b_host_free", "h264b_cut_byte_ranges", "h264b_dev_alloc", "h264b_dev_free", "h264b_memcpy_h2d",
    "h264b_memcpy_d2h", "h264b_launch_count", "h264b_annexb_scratch_bytes", "h264b_annexb_scan_dev",
    "h264b_annexb_scan", "h264b_nal_units", "h264b_ctx_init_dev", "h264b_ctx_init", "h264b_pre_ctx_state", "h264b_mn",
    "h264b_cabac_decode_dev", "h264b_cabac_decode", "h264b_engine_step", "h264b_binary_decision",
    "h264b_state_transition", "h264b_stream_decode", "h264b_stream_submit", "h264b_stream_wait",
    "h264b_slice_headers_dev", "h264b_slice_headers",
    "h264b_slice_select_dev",
    "h264b_parse_sps", "h264b_parse_pps", "h264b_parse_sps_dev", "h264b_parse_pps_dev", "h264b_make_param_sets",
    "h264b_param_set_select_dev",
    "h264b_ctx_idx", "h264b_new_binarization", "h264b_init_cabac", "h264b_mb_bin_string", "h264b_bin_string_match",
    "h264b_mb_type_decode_dev", "h264b_mb_type_decode",
    "h264b_scheduler_create", "h264b_scheduler_destroy", "h264b_scheduler_last_error", "h264b_scheduler_run", "h264b_scheduler_plan",
]

NAL_DTYPE = np.dtype([("start", "<u8"), ("rbsp_off", "<u8"), ("num_bytes", "<u4"), ("rbsp_len", "<u4"),
                      ("forbidden_zero_bit", "u1"), ("ref_idc", "u1"), ("type", "u1"), ("header_bytes", "u1"),
                      ("flags", "<u4")])
NAL_EXT_DTYPE = np.dtype([(n, "u1") for n in (
    "svc_extension_flag", "avc_3d_extension_flag", "idr_flag", "priority_id", "no_inter_layer_pred_flag",
    "dependency_id", "quality_id", "temporal_id", "use_ref_base_pic_flag", "discardable_flag", "output_flag",
    "reserved_three_2bits", "non_idr_flag", "anchor_pic_flag", "inter_view_flag", "reserved_one_bit")] +
    [("view_id", "<u2"), ("view_idx", "u1"), ("depth_flag", "u1"), ("pad", "<u4")])
FINAL_DTYPE = np.dtype([("cod_i_range", "<i8"), ("cod_i_offset", "<i8"), ("bits_read", "<u8"), ("flags", "<u4"),
                        ("n_bins", "<u4")])
SLICE_QP_DTYPE = np.dtype([("slice_qp_y", "<i4"), ("cabac_init_idc", "<i4")])
assert NAL_DTYPE.itemsize == 32 and NAL_EXT_DTYPE.itemsize == 24 and FINAL_DTYPE.itemsize == 32


class ScanSummary(C.Structure):
    _fields_ = [("n_start_codes", C.c_uint64), ("n_nals", C.c_uint64), ("rbsp_bytes", C.c_uint64),
                ("first_start", C.c_uint64), ("n_epb", C.c_uint64), ("status", C.c_uint32), ("reserved", C.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


class CabacJob(C.Structure):
    _fields_ = [("bytes", C.c_void_p), ("total_bytes", C.c_uint64), ("off", C.c_void_p), ("len", C.c_void_p),
                ("n_slices", C.c_uint32), ("n_ctx", C.c_uint32), ("ops", C.c_void_p), ("n_ops_max", C.c_uint32),
                ("n_ops", C.c_void_p), ("qp", C.c_void_p), ("init_states", C.c_void_p), ("bins", C.c_void_p),
                ("bins_off", C.c_void_p), ("bins_stride_words", C.c_uint32), ("final", C.c_void_p), ("final_states", C.c_void_p),
                ("flags", C.c_uint32), ("n_ctx_used", C.c_uint32)]


class StreamJob(C.Structure):
    _fields_ = [("stream", C.c_void_p), ("n", C.c_uint64), ("slice_data_offset", C.c_uint32), ("n_ctx", C.c_uint32),
                ("ops", C.c_void_p), ("n_ops_max", C.c_uint32), ("n_ops", C.c_void_p), ("qp", C.c_void_p),
                ("max_slices", C.c_uint32), ("flags", C.c_uint32), ("param_sets", C.c_void_p),
                ("max_sps", C.c_uint32), ("max_pps", C.c_uint32), ("initial_sps", C.c_void_p), ("initial_pps", C.c_void_p)]


class MbTypeJob(C.Structure):
    _fields_ = [("bytes", C.c_void_p), ("total_bytes", C.c_uint64), ("off", C.c_void_p), ("len", C.c_void_p),
                ("n_slices", C.c_uint32), ("n_ctx", C.c_uint32), ("slice_kind", C.c_void_p), ("n_mb", C.c_void_p),
                ("n_mb_max", C.c_uint32), ("flags", C.c_uint32), ("qp", C.c_void_p), ("init_states", C.c_void_p),
                ("mb_type", C.c_void_p), ("final", C.c_void_p), ("final_states", C.c_void_p)]


MB_FINAL_DTYPE = np.dtype([("cod_i_range", "<i8"), ("cod_i_offset", "<i8"), ("bits_read", "<u8"), ("flags", "<u4"),
                           ("n_bins", "<u4"), ("n_mb", "<u4"), ("reserved", "<u4")])


class BatchStream(C.Structure):
    _fields_ = [("stream", C.c_void_p), ("n", C.c_uint64), ("first_slice", C.c_uint32), ("n_slices", C.c_uint32)]


class BatchJob(C.Structure):
    _fields_ = [("streams", C.c_void_p), ("n_streams", C.c_uint32), ("total_slices", C.c_uint32), ("n_ctx", C.c_uint32),
                ("n_ops_max", C.c_uint32), ("ops", C.c_void_p), ("n_ops", C.c_void_p), ("qp", C.c_void_p),
                ("slice_data_offset", C.c_uint32), ("flags", C.c_uint32), ("group_bytes", C.c_uint64)]


class BatchResult(C.Structure):
    _fields_ = [("stream_device", C.c_void_p), ("stream_job", C.c_void_p), ("stream_nal_off", C.c_void_p),
                ("nals", C.c_void_p), ("final", C.c_void_p), ("bins_off", C.c_void_p), ("bins", C.c_void_p),
                ("slice_done_ms", C.c_void_p), ("n_devices", C.c_uint32), ("reserved", C.c_uint32),
                ("device_busy_ms", C.c_void_p), ("device_bytes", C.c_void_p), ("device_jobs", C.c_void_p),
                ("makespan_ms", C.c_double), ("total_bins", C.c_uint64), ("total_nals", C.c_uint64)]


class StreamResult(C.Structure):
    _fields_ = [("scan", ScanSummary), ("nals", C.c_void_p), ("n_slices", C.c_uint32), ("reserved", C.c_uint32),
                ("slice_nal", C.c_void_p), ("bins_off", C.c_void_p), ("bins", C.c_void_p), ("final", C.c_void_p),
                ("total_bins", C.c_uint64), ("rbsp", C.c_void_p), ("d_rbsp", C.c_void_p), ("ext", C.c_void_p), ("headers", C.c_void_p),
                ("n_sps", C.c_uint32), ("n_pps", C.c_uint32), ("sps", C.c_void_p), ("pps", C.c_void_p),
                ("sps_nal", C.c_void_p), ("pps_nal", C.c_void_p), ("slice_sps", C.c_void_p), ("slice_pps", C.c_void_p)]


PARAM_SET_FIELDS = ["use_separate_color_plane", "chroma_format", "frame_mbs_only", "pic_order_count_type",
                    "log2_max_pic_order_cnt_lsb_min4", "delta_pic_order_always_zero",
                    "bottom_field_pic_order_in_frame_present", "redundant_pic_cnt_present", "weighted_pred",
                    "weighted_bipred", "entropy_coding_mode", "deblocking_filter_control_present",
                    "num_slice_groups_minus1", "slice_group_map_type", "pic_size_in_map_units_minus1",
                    "slice_group_change_rate_minus1", "pic_init_qp_minus26", "reserved"]


class ParamSets(C.Structure):
    _fields_ = [(n, C.c_int64) for n in PARAM_SET_FIELDS]


SPS_MAX_REF_FRAMES, SPS_MAX_HRD = 256, 64
SPS_SCALARS = [
    "profile", "constraint0", "constraint1", "constraint2", "constraint3", "constraint4", "constraint5", "level", "id",
    "chroma_format", "use_separate_color_plane", "bit_depth_luma_minus8", "bit_depth_chroma_minus8",
    "qprime_y_zero_transform_bypass", "seq_scaling_matrix_present", "log2_max_frame_num_minus4", "pic_order_count_type",
    "log2_max_pic_order_cnt_lsb_min4", "delta_pic_order_always_zero", "offset_for_non_ref_pic",
    "offset_for_top_to_bottom_field", "num_ref_frames_in_pic_order_cnt_cycle", "max_num_ref_frames",
    "gaps_in_frame_num_value_allowed", "pic_width_in_mbs_minus1", "pic_height_in_map_units_minus1", "frame_mbs_only",
    "mb_adaptive_frame_field", "direct_8x8_inference", "frame_cropping", "frame_crop_left_offset",
    "frame_crop_right_offset", "frame_crop_top_offset", "frame_crop_bottom_offset", "vui_parameters_present",
    "aspect_ratio_info_present", "aspect_ratio", "sar_width", "sar_height", "overscan_info_present",
    "overscan_appropriate", "video_signal_type_present", "video_format", "video_full_range", "color_description_present",
    "color_primaries", "transfer_characteristics", "matrix_coefficients", "chroma_loc_info_present",
    "chroma_sample_loc_type_top_field", "chroma_sample_loc_type_bottom_field", "cpb_cnt_minus1", "bit_rate_scale",
    "cpb_size_scale", "initial_cpb_removal_delay_length_minus1", "cpb_removal_delay_length_minus1",
    "dpb_output_delay_length_minus1", "time_offset_length", "timing_info_present", "num_units_in_tick", "time_scale",
    "nal_hrd_parameters_present", "fixed_frame_rate", "vcl_hrd_parameters_present", "low_hrd_delay", "pic_struct_present",
    "bitstream_restriction", "motion_vectors_over_pic_boundaries", "max_bytes_per_pic_denom", "max_bits_per_mb_denom",
    "log2_max_mv_length_horizontal", "log2_max_mv_length_vertical", "max_dec_frame_buffering", "max_num_reorder_frames"]
SPS_DTYPE = np.dtype([(n, "<i8") for n in SPS_SCALARS] +
                     [("n_seq_scaling_list", "<i8"), ("seq_scaling_list", "<i8", (12,)),
                      ("n_offset_for_ref_frame", "<i8"), ("offset_for_ref_frame", "<i8", (SPS_MAX_REF_FRAMES,)),
                      ("n_hrd", "<i8"), ("bit_rate_value_minus1", "<i8", (SPS_MAX_HRD,)),
                      ("cpb_size_value_minus1", "<i8", (SPS_MAX_HRD,)), ("cbr", "<i8", (SPS_MAX_HRD,)),
                      ("bits_read", "<u8"), ("status", "<u4"), ("reserved", "<u4")])
PPS_SCALARS = [
    "id", "sps_id", "entropy_coding_mode", "num_slice_groups_minus1", "bottom_field_pic_order_in_frame_present",
    "slice_group_map_type", "slice_group_change_direction", "slice_group_change_rate_minus1",
    "pic_size_in_map_units_minus1", "num_ref_idx_l0_default_active_minus1", "num_ref_idx_l1_default_active_minus1",
    "weighted_pred", "weighted_bipred", "pic_init_qp_minus26", "pic_init_qs_minus26", "chroma_qp_index_offset",
    "deblocking_filter_control_present", "constrained_intra_pred", "redundant_pic_cnt_present", "transform_8x8_mode",
    "pic_scaling_matrix_present", "second_chroma_qp_index_offset"]
PPS_DTYPE = np.dtype([(n, "<i8") for n in PPS_SCALARS] + [("bits_read", "<u8"), ("status", "<u4"), ("reserved", "<u4")])


BINARIZATION_FIELDS = ["syntax_element", "prefix_suffix", "fixed_length", "unary", "truncated_unary", "cmax", "uegk",
                       "cmax_value", "max_is_prefix_suffix", "max_prefix", "max_suffix", "off_is_prefix_suffix",
                       "off_prefix", "off_suffix", "use_decode_bypass", "reserved"]
BINARIZATION_DTYPE = np.dtype([(n, "<i4") for n in BINARIZATION_FIELDS])
NA_CTX_ID = 10000

SLICE_HEADER_FIELDS = [
    "first_mb_in_slice", "slice_type", "pps_id", "color_plane_id", "field_pic", "bottom_field", "idr_pic_id",
    "pic_order_cnt_lsb", "delta_pic_order_cnt_bottom", "delta_pic_order_cnt0", "delta_pic_order_cnt1",
    "redundant_pic_cnt", "direct_spatial_mv_pred", "num_ref_idx_active_override", "num_ref_idx_l0_active_minus1",
    "num_ref_idx_l1_active_minus1", "ref_pic_list_modification_flag_l0", "ref_pic_list_modification_flag_l1",
    "modification_of_pic_nums", "abs_diff_pic_num_minus1", "long_term_pic_num", "luma_log2_weight_denom",
    "chroma_log2_weight_denom", "n_luma_weight_l0", "n_chroma_weight_l0", "n_luma_weight_l1", "n_chroma_weight_l1",
    "no_output_of_prior_pics_flag", "long_term_reference_flag", "adaptive_ref_pic_marking_mode_flag",
    "memory_management_control_operation", "difference_of_pic_nums_minus1", "long_term_frame_idx",
    "max_long_term_frame_idx_plus1", "cabac_init_idc", "slice_qp_delta", "sp_for_switch", "slice_qs_delta",
    "disable_deblocking_filter", "slice_alpha_c0_offset_div2", "slice_beta_offset_div2", "slice_group_change_cycle",
    "chroma_array_type", "slice_qp_y"]
SLICE_HEADER_DTYPE = np.dtype([(n, "<i8") for n in SLICE_HEADER_FIELDS] +
                              [("header_bits", "<u8"), ("status", "<u4"), ("reserved", "<u4")])
SH_OK, SH_PANIC, SH_HANG = 0, 1, 2


class H264BError(RuntimeError):
    def __init__(self, code, msg=""):
        super().__init__("h264b status %d: %s" % (code, msg))
        self.code = code


def load():
    if not os.path.exists(LIB_PATH):
        raise ImportError("libh264b200.so is not built (%s); run `python -m h264decode_b200.build` -- there is no "
                          "fallback implementation" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32
    P = C.POINTER
    sig = {
        "h264b_version": (i32, []),
        "h264b_device_count": (i32, [P(i32)]),
        "h264b_create": (i32, [i32, P(vp)]),
        "h264b_destroy": (None, [vp]),
        "h264b_last_error": (C.c_char_p, [vp]),
        "h264b_set_stream": (i32, [vp, vp]),
        "h264b_sync": (i32, [vp]),
        "h264b_host_alloc": (i32, [vp, C.c_size_t, P(vp)]),
        "h264b_host_free": (i32, [vp, vp]),
        "h264b_cut_byte_ranges": (i32, [vp, C.c_uint64, C.c_uint32, vp, vp]),
        "h264b_dev_alloc": (i32, [vp, C.c_size_t, P(vp)]),
        "h264b_dev_free": (i32, [vp, vp]),
        "h264b_memcpy_h2d": (i32, [vp, vp, vp, C.c_size_t]),
        "h264b_memcpy_d2h": (i32, [vp, vp, vp, C.c_size_t]),
        "h264b_launch_count": (i32, [vp, P(u64)]),
        "h264b_annexb_scratch_bytes": (u64, [u64]),
        "h264b_annexb_scan_dev": (i32, [vp, vp, u64, vp, vp, vp, u32, vp, u32]),
        "h264b_annexb_scan": (i32, [vp, vp, u64, u32, i32, P(vp), P(vp), P(ScanSummary), P(vp), P(vp)]),
        "h264b_nal_units": (i32, [vp, vp, u64, vp, vp, u32, u32, vp, vp, vp]),
        "h264b_ctx_init_dev": (i32, [vp, vp, u32, u32, vp, u32]),
        "h264b_ctx_init": (i32, [vp, vp, u32, u32, vp, u32]),
        "h264b_pre_ctx_state": (i32, [vp, i32, i32, i32, P(i32)]),
        "h264b_mn": (i32, [vp, i32, i32, u32, P(i32), P(i32)]),
        "h264b_cabac_decode_dev": (i32, [vp, P(CabacJob)]),
        "h264b_cabac_decode": (i32, [vp, P(CabacJob)]),
        "h264b_engine_step": (i32, [vp, u32, u32, vp, u32, P(C.c_int64), P(C.c_int64), P(i32), P(i32), P(i32), P(u32)]),
        "h264b_binary_decision": (i32, [vp, u32, i32, i32, P(C.c_int64), P(C.c_int64), P(i32)]),
        "h264b_state_transition": (i32, [vp, u32, P(i32), P(i32), i32]),
        "h264b_stream_decode": (i32, [vp, P(StreamJob), P(StreamResult)]),
        "h264b_stream_submit": (i32, [vp, P(StreamJob), P(u64)]),
        "h264b_slice_headers_dev": (i32, [vp, P(ParamSets), vp, u64, vp, vp, vp, vp, vp, vp, u32, vp]),
        "h264b_slice_headers": (i32, [vp, P(ParamSets), vp, u64, vp, vp, vp, vp, u32, vp]),
        "h264b_stream_wait": (i32, [vp, u64, P(StreamResult)]),
        "h264b_slice_select_dev": (i32, [vp, vp, vp, u32, u32, u32, vp, vp, vp, vp]),
        "h264b_parse_sps": (i32, [vp, vp, u64, vp, vp, u32, vp]),
        "h264b_parse_pps": (i32, [vp, vp, u64, vp, vp, u32, vp]),
        "h264b_parse_sps_dev": (i32, [vp, vp, u64, vp, vp, vp, vp, u32, vp, vp]),
        "h264b_parse_pps_dev": (i32, [vp, vp, u64, vp, vp, vp, vp, u32, vp, vp]),
        "h264b_make_param_sets": (i32, [vp, vp, P(ParamSets)]),
        "h264b_param_set_select_dev": (i32, [vp, vp, vp, u32, u32, u32, vp, vp, vp]),
        "h264b_ctx_idx": (i32, [vp, u32, vp, vp, vp, vp]),
        "h264b_new_binarization": (i32, [vp, u32, vp, vp, vp]),
        "h264b_init_cabac": (i32, [vp, u32, u32, vp, vp, vp, vp, vp, vp, vp, vp]),
        "h264b_mb_bin_string": (i32, [vp, u32, vp, vp, vp, vp, vp]),
        "h264b_bin_string_match": (i32, [vp, u32, vp, vp, vp, vp, vp]),
        "h264b_mb_type_decode_dev": (i32, [vp, P(MbTypeJob)]),
        "h264b_mb_type_decode": (i32, [vp, P(MbTypeJob)]),
        "h264b_scheduler_create": (i32, [vp, u32, P(vp)]),
        "h264b_scheduler_destroy": (None, [vp]),
        "h264b_scheduler_last_error": (C.c_char_p, [vp]),
        "h264b_scheduler_run": (i32, [vp, P(BatchJob), P(BatchResult)]),
        "h264b_scheduler_plan": (i32, [P(BatchJob), u32, u32, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        if os.environ.get("H264B_LIB") and not hasattr(L, name):
            continue  # (an experiment build of an older revision)
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    return L


_lib = load()


def lib():
    return _lib


def make_op(kind, ctx=0):
    return (kind << 14) | (ctx & 0x3FF)


def cut_byte_ranges(stream, n_ranges):
    """h264b_cut_byte_ranges (host-only planner, needs no context / GPU) -> [(begin, end)] * n_ranges"""
    s = np.ascontiguousarray(stream, dtype=np.uint8)
    b, e = np.zeros(n_ranges, np.uint64), np.zeros(n_ranges, np.uint64)
    rc = _lib.h264b_cut_byte_ranges(C.c_void_p(s.ctypes.data), len(s), n_ranges, C.c_void_p(b.ctypes.data),
                                    C.c_void_p(e.ctypes.data))
    if rc:
        raise H264BError(rc, "h264b_cut_byte_ranges")
    return [(int(x), int(y)) for x, y in zip(b, e)]


def _from_ptr(ptr, dtype, count):
    """copy `count` records of `dtype` from a host pointer the library owns"""
    if not count:
        return np.zeros(0, dtype=dtype)
    nbytes = int(count) * np.dtype(dtype).itemsize
    buf = (C.c_uint8 * nbytes).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=int(count)).copy()


class Context:
    """One h264b_ctx: one GPU, one stream.  Methods mirror the C entry points one to one."""

    def __init__(self, device=0):
        h = C.c_void_p()
        rc = _lib.h264b_create(device, C.byref(h))
        if rc != OK:
            raise H264BError(rc, "h264b_create(device=%d) failed%s" % (
                device, " -- no CUDA device; this library has no CPU path" if rc == E_NO_DEVICE else ""))
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            _lib.h264b_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != OK:
            raise H264BError(rc, (_lib.h264b_last_error(self.h) or b"").decode(errors="replace"))

    # ---- plumbing
    def set_stream(self, cuda_stream):
        self._check(_lib.h264b_set_stream(self.h, C.c_void_p(cuda_stream)))

    def sync(self):
        self._check(_lib.h264b_sync(self.h))

    def launch_count(self):
        n = C.c_uint64()
        self._check(_lib.h264b_launch_count(self.h, C.byref(n)))
        return n.value

    def host_alloc(self, nbytes):
        """pinned host memory as a uint8 numpy view (free with host_free(arr))"""
        p = C.c_void_p()
        self._check(_lib.h264b_host_alloc(self.h, nbytes, C.byref(p)))
        arr = np.frombuffer((C.c_uint8 * max(nbytes, 1)).from_address(p.value), dtype=np.uint8, count=nbytes)
        arr.flags.writeable = True
        self._pinned = getattr(self, "_pinned", {})
        self._pinned[arr.ctypes.data] = p.value
        return arr

    def host_free(self, arr):
        p = self._pinned.pop(arr.ctypes.data)
        self._check(_lib.h264b_host_free(self.h, C.c_void_p(p)))

    def dev_alloc(self, nbytes):
        p = C.c_void_p()
        self._check(_lib.h264b_dev_alloc(self.h, nbytes, C.byref(p)))
        return p.value

    def dev_free(self, p):
        self._check(_lib.h264b_dev_free(self.h, C.c_void_p(p)))

    def h2d(self, dptr, arr):
        arr = np.ascontiguousarray(arr)
        self._check(_lib.h264b_memcpy_h2d(self.h, C.c_void_p(dptr), C.c_void_p(arr.ctypes.data), arr.nbytes))
        return arr  # keep alive until sync

    def d2h(self, arr, dptr):
        self._check(_lib.h264b_memcpy_d2h(self.h, C.c_void_p(arr.ctypes.data), C.c_void_p(dptr), arr.nbytes))

    # ---- Annex-B
    def annexb_scan(self, stream, flags=0, want_rbsp=True, want_ext=True):
        """-> (summary dict, nals[NAL_DTYPE], ext[NAL_EXT_DTYPE] | None, rbsp uint8[len(stream)] | None);
        the RBSP of NAL k is rbsp[nals[k].rbsp_off : nals[k].rbsp_off + nals[k].rbsp_len]"""
        s = np.ascontiguousarray(stream, dtype=np.uint8)
        nals, ext, rbsp, drbsp = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        summ = ScanSummary()
        self._check(_lib.h264b_annexb_scan(self.h, C.c_void_p(s.ctypes.data), len(s), flags, 1 if want_rbsp else 0,
                                           C.byref(nals), C.byref(ext) if want_ext else None, C.byref(summ),
                                           C.byref(rbsp), C.byref(drbsp)))
        n = summ.n_nals
        out_n = _from_ptr(nals.value, NAL_DTYPE, n)
        out_e = _from_ptr(ext.value, NAL_EXT_DTYPE, n) if want_ext else None
        out_r = _from_ptr(rbsp.value, np.uint8, len(s) if n else 0) if want_rbsp else None
        d = summ.as_dict()
        d["d_rbsp"] = drbsp.value
        return d, out_n, out_e, out_r

    def annexb_scan_dev(self, d_stream, n, d_rbsp, d_nals, d_ext, nal_cap, d_summary, flags=0):
        self._check(_lib.h264b_annexb_scan_dev(self.h, C.c_void_p(d_stream), n, C.c_void_p(d_rbsp), C.c_void_p(d_nals),
                                               C.c_void_p(d_ext) if d_ext else None, nal_cap, C.c_void_p(d_summary),
                                               flags))

    def nal_units(self, frames, flags=0):
        """NewNalUnit for a list of byte strings -> (nals, ext, [rbsp bytes per frame])"""
        lens = np.array([len(f) for f in frames], dtype=np.uint32)
        offs = np.zeros(len(frames), dtype=np.uint64)
        if len(frames):
            offs[1:] = np.cumsum(lens[:-1].astype(np.uint64))
        cat = np.frombuffer(b"".join(bytes(f) for f in frames), dtype=np.uint8).copy() if len(frames) else \
            np.zeros(0, np.uint8)
        nals = np.zeros(len(frames), dtype=NAL_DTYPE)
        ext = np.zeros(len(frames), dtype=NAL_EXT_DTYPE)
        rbsp = np.zeros(len(cat) + 16, dtype=np.uint8)
        self._check(_lib.h264b_nal_units(self.h, C.c_void_p(cat.ctypes.data), len(cat), C.c_void_p(offs.ctypes.data),
                                         C.c_void_p(lens.ctypes.data), len(frames), flags,
                                         C.c_void_p(nals.ctypes.data), C.c_void_p(ext.ctypes.data),
                                         C.c_void_p(rbsp.ctypes.data)))
        outs = [bytes(rbsp[int(o):int(o) + int(nl["rbsp_len"])]) for o, nl in zip(offs, nals)]
        return nals, ext, outs

    # ---- context init
    @staticmethod
    def slice_qp(qp, idc):
        p = np.zeros(len(qp), dtype=SLICE_QP_DTYPE)
        p["slice_qp_y"] = qp
        p["cabac_init_idc"] = idc
        return p

    def ctx_init(self, qp, idc, n_ctx, flags=0):
        p = self.slice_qp(qp, idc)
        out = np.zeros((len(p), n_ctx), dtype=np.uint8)
        self._check(_lib.h264b_ctx_init(self.h, C.c_void_p(p.ctypes.data), len(p), n_ctx,
                                        C.c_void_p(out.ctypes.data), flags))
        return out

    def ctx_init_dev(self, d_params, n_slices, n_ctx, d_states, flags=0):
        self._check(_lib.h264b_ctx_init_dev(self.h, C.c_void_p(d_params), n_slices, n_ctx, C.c_void_p(d_states), flags))

    def pre_ctx_state(self, m, n, qp):
        out = C.c_int32()
        self._check(_lib.h264b_pre_ctx_state(self.h, m, n, qp, C.byref(out)))
        return out.value

    def mn(self, ctx_idx, idc, flags=0):
        m, n = C.c_int32(), C.c_int32()
        self._check(_lib.h264b_mn(self.h, ctx_idx, idc, flags, C.byref(m), C.byref(n)))
        return m.value, n.value

    # ---- CABAC
    def cabac_decode(self, data, off, length, ops, n_ops, n_ctx, qp=None, idc=None, init_states=None, flags=0,
                     want_states=True):
        """host buffers -> (bins uint32[n, stride], final[FINAL_DTYPE], final_states | None)"""
        data = np.ascontiguousarray(data, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        length = np.ascontiguousarray(length, dtype=np.uint32)
        ops = np.ascontiguousarray(ops, dtype=np.uint16)
        n = len(off)
        nops = None if n_ops is None else np.ascontiguousarray(n_ops, dtype=np.uint32)
        stride = (len(ops) + 1 + 31) // 32
        bins = np.zeros((n, stride), dtype=np.uint32)
        fin = np.zeros(n, dtype=FINAL_DTYPE)
        fst = np.zeros((n, n_ctx), dtype=np.uint8) if want_states else None
        p = self.slice_qp(qp, idc) if qp is not None else None
        init = None if init_states is None else np.ascontiguousarray(init_states, dtype=np.uint8)
        j = CabacJob()
        j.bytes = data.ctypes.data
        j.total_bytes = len(data)
        j.off = off.ctypes.data
        j.len = length.ctypes.data
        j.n_slices = n
        j.n_ctx = n_ctx
        j.ops = ops.ctypes.data
        j.n_ops_max = len(ops)
        j.n_ops = nops.ctypes.data if nops is not None else None
        j.qp = p.ctypes.data if p is not None else None
        j.init_states = init.ctypes.data if init is not None else None
        j.bins = bins.ctypes.data
        j.bins_stride_words = stride
        j.final = fin.ctypes.data
        j.final_states = fst.ctypes.data if fst is not None else None
        j.flags = flags
        self._check(_lib.h264b_cabac_decode(self.h, C.byref(j)))
        return bins, fin, fst

    def slice_select_dev(self, d_nals, d_summary, nal_cap, slice_data_offset, max_slices, d_off, d_len, d_slice_nal,
                         d_n_slices):
        self._check(_lib.h264b_slice_select_dev(self.h, C.c_void_p(d_nals), C.c_void_p(d_summary), nal_cap,
                                                slice_data_offset, max_slices, C.c_void_p(d_off), C.c_void_p(d_len),
                                                C.c_void_p(d_slice_nal), C.c_void_p(d_n_slices)))

    def cabac_decode_dev(self, **kw):
        j = CabacJob()
        for k, v in kw.items():
            setattr(j, k, v)
        self._check(_lib.h264b_cabac_decode_dev(self.h, C.byref(j)))

    def engine_step(self, kind, R, O, bits=b"", n_bits=None, p_state=0, val_mps=0, bin_val=0, flags=0):
        """-> dict(R, O, p_state, val_mps, bin, bits_used)"""
        r, o = C.c_int64(R), C.c_int64(O)
        p, v, b, used = C.c_int32(p_state), C.c_int32(val_mps), C.c_int32(bin_val), C.c_uint32()
        buf = np.frombuffer(bytes(bits) + b"\x00", dtype=np.uint8).copy()
        nb = len(bits) * 8 if n_bits is None else n_bits
        self._check(_lib.h264b_engine_step(self.h, kind, flags, C.c_void_p(buf.ctypes.data), nb, C.byref(r),
                                           C.byref(o), C.byref(p), C.byref(v), C.byref(b), C.byref(used)))
        return dict(R=r.value, O=o.value, p_state=p.value, val_mps=v.value, bin=b.value, bits_used=used.value)

    def binary_decision(self, p_state, val_mps, R, O, flags=0):
        r, o, b = C.c_int64(R), C.c_int64(O), C.c_int32()
        self._check(_lib.h264b_binary_decision(self.h, flags, p_state, val_mps, C.byref(r), C.byref(o), C.byref(b)))
        return b.value, r.value, o.value

    def state_transition(self, p_state, val_mps, bin_val, flags=0):
        p, v = C.c_int32(p_state), C.c_int32(val_mps)
        self._check(_lib.h264b_state_transition(self.h, flags, C.byref(p), C.byref(v), bin_val))
        return p.value, v.value

    # ---- whole front end
    def stream_decode(self, stream, ops, n_ops, qp, idc, n_ctx, slice_data_offset=0, flags=0):
        """-> dict(scan, nals, slice_nal, bins, final, total_bins)"""
        s = np.ascontiguousarray(stream, dtype=np.uint8)
        ops = np.ascontiguousarray(ops, dtype=np.uint16)
        p = self.slice_qp(qp, idc)
        nops = None if n_ops is None else np.ascontiguousarray(n_ops, dtype=np.uint32)
        j = StreamJob()
        j.stream = s.ctypes.data
        j.n = len(s)
        j.slice_data_offset = slice_data_offset
        j.n_ctx = n_ctx
        j.ops = ops.ctypes.data
        j.n_ops_max = len(ops)
        j.n_ops = nops.ctypes.data if nops is not None else None
        j.qp = p.ctypes.data
        j.max_slices = len(p)
        j.flags = flags
        r = StreamResult()
        self._check(_lib.h264b_stream_decode(self.h, C.byref(j), C.byref(r)))
        return self._stream_result(r)

    @staticmethod
    def param_sets(**kw):
        p = ParamSets()
        for k, v in kw.items():
            setattr(p, k, int(v))
        return p

    def _parse_psets(self, fn, dtype, data, off, length):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        length = np.ascontiguousarray(length, dtype=np.uint32)
        out = np.zeros(len(off), dtype=dtype)
        self._check(fn(self.h, data.ctypes.data if len(data) else None, len(data), off.ctypes.data, length.ctypes.data,
                       len(off), out.ctypes.data))
        return out

    def parse_sps(self, data, off, length):
        """NewSPS for every RBSP data[off[i] : off[i] + length[i]] -> SPS_DTYPE[n]"""
        return self._parse_psets(_lib.h264b_parse_sps, SPS_DTYPE, data, off, length)

    def parse_pps(self, data, off, length):
        """NewPPS -> PPS_DTYPE[n]"""
        return self._parse_psets(_lib.h264b_parse_pps, PPS_DTYPE, data, off, length)

    @staticmethod
    def make_param_sets(sps_rec, pps_rec):
        """ParamSets (what the slice-header walk reads) from one SPS_DTYPE and one PPS_DTYPE record"""
        a = np.ascontiguousarray(np.asarray(sps_rec, dtype=SPS_DTYPE).reshape(1))
        b = np.ascontiguousarray(np.asarray(pps_rec, dtype=PPS_DTYPE).reshape(1))
        p = ParamSets()
        rc = _lib.h264b_make_param_sets(a.ctypes.data, b.ctypes.data, C.byref(p))
        if rc:
            raise H264BError(rc)
        return p

    # ---- syntax-element glue (rows I5 / f3): batches of queries, one device thread each
    def ctx_idx(self, bin_idx, max_bin_idx_ctx, ctx_idx_offset):
        a, b, c = (np.ascontiguousarray(x, dtype=np.int64) for x in (bin_idx, max_bin_idx_ctx, ctx_idx_offset))
        out = np.zeros(len(a), np.int64)
        self._check(_lib.h264b_ctx_idx(self.h, len(a), a.ctypes.data, b.ctypes.data, c.ctypes.data, out.ctypes.data))
        return out

    def new_binarization(self, syntax_element, slice_type_name):
        a, b = (np.ascontiguousarray(x, dtype=np.int32) for x in (syntax_element, slice_type_name))
        out = np.zeros(len(a), BINARIZATION_DTYPE)
        self._check(_lib.h264b_new_binarization(self.h, len(a), a.ctypes.data, b.ctypes.data, out.ctypes.data))
        return out

    def init_cabac(self, bin_idx, max_prefix, off_prefix, pic_init_qp_minus26, slice_qp_delta, flags=0):
        arrs = [np.ascontiguousarray(x, dtype=np.int64) for x in (bin_idx, max_prefix, off_prefix, pic_init_qp_minus26,
                                                                   slice_qp_delta)]
        n = len(arrs[0])
        p, v, c = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int64)
        self._check(_lib.h264b_init_cabac(self.h, flags, n, *[a.ctypes.data for a in arrs], p.ctypes.data, v.ctypes.data,
                                          c.ctypes.data))
        return p, v, c

    def mb_bin_string(self, slice_type_name, mb_type, sub_mb):
        a = np.ascontiguousarray(slice_type_name, dtype=np.int32)
        b = np.ascontiguousarray(mb_type, dtype=np.int64)
        c = np.ascontiguousarray(sub_mb, dtype=np.uint8)
        ln, bits = np.zeros(len(a), np.int32), np.zeros(len(a), np.uint32)
        self._check(_lib.h264b_mb_bin_string(self.h, len(a), a.ctypes.data, b.ctypes.data, c.ctypes.data, ln.ctypes.data,
                                             bits.ctypes.data))
        return ln, bits

    def bin_string_match(self, bin_len, bin_bits, n_bits, bits):
        a, c = (np.ascontiguousarray(x, dtype=np.int32) for x in (bin_len, n_bits))
        b, d = (np.ascontiguousarray(x, dtype=np.uint32) for x in (bin_bits, bits))
        out = np.zeros(len(a), np.int32)
        self._check(_lib.h264b_bin_string_match(self.h, len(a), a.ctypes.data, b.ctypes.data, c.ctypes.data, d.ctypes.data,
                                                out.ctypes.data))
        return out

    def slice_headers(self, params, data, off, length, nal_type, nal_ref_idc):
        """host buffers -> SLICE_HEADER_DTYPE[n]"""
        data = np.ascontiguousarray(data, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        length = np.ascontiguousarray(length, dtype=np.uint32)
        t = np.ascontiguousarray(nal_type, dtype=np.uint8)
        r = np.ascontiguousarray(nal_ref_idc, dtype=np.uint8)
        out = np.zeros(len(off), dtype=SLICE_HEADER_DTYPE)
        self._check(_lib.h264b_slice_headers(self.h, C.byref(params), data.ctypes.data, len(data), off.ctypes.data,
                                             length.ctypes.data, t.ctypes.data, r.ctypes.data, len(off), out.ctypes.data))
        return out

    def slice_headers_dev(self, params, d_bytes, total_bytes, d_nals, d_slice_nal, n_slices, d_out):
        self._check(_lib.h264b_slice_headers_dev(self.h, C.byref(params), d_bytes, total_bytes, None, None, None, None,
                                                 d_nals, d_slice_nal, n_slices, d_out))

    def mb_type_decode(self, data, off, length, slice_kind, n_mb, n_ctx, qp=None, idc=None, init_states=None, flags=0,
                       want_states=True):
        """mb_type as a syntax element for every slice (slice_kind 0: I, 1: P / SP) -> (mb_type uint8[n_slices][n_mb_max],
        final MB_FINAL_DTYPE[n_slices], final states | None)"""
        d = np.ascontiguousarray(data, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        length = np.ascontiguousarray(length, dtype=np.uint32)
        kind = np.ascontiguousarray(slice_kind, dtype=np.uint8)
        n_mb = np.ascontiguousarray(n_mb, dtype=np.uint32)
        ns = len(off)
        n_max = int(n_mb.max()) if ns else 0
        p = self.slice_qp(qp, idc) if qp is not None else None
        init = None if init_states is None else np.ascontiguousarray(init_states, dtype=np.uint8)
        out = np.zeros((ns, max(n_max, 1)), dtype=np.uint8)
        fin = np.zeros(ns, dtype=MB_FINAL_DTYPE)
        fst = np.zeros((ns, n_ctx), dtype=np.uint8) if want_states else None
        j = MbTypeJob()
        j.bytes, j.total_bytes = d.ctypes.data, len(d)
        j.off, j.len, j.n_slices, j.n_ctx = off.ctypes.data, length.ctypes.data, ns, n_ctx
        j.slice_kind, j.n_mb, j.n_mb_max, j.flags = kind.ctypes.data, n_mb.ctypes.data, max(n_max, 1), flags
        j.qp = p.ctypes.data if p is not None else None
        j.init_states = init.ctypes.data if init is not None else None
        j.mb_type, j.final = out.ctypes.data, fin.ctypes.data
        j.final_states = fst.ctypes.data if fst is not None else None
        self._check(_lib.h264b_mb_type_decode(self.h, C.byref(j)))
        return out, fin, fst

    def mb_type_decode_dev(self, **kw):
        """device pointers (ints): bytes, total_bytes, off, len, n_slices, n_ctx, slice_kind, n_mb, n_mb_max, flags, qp,
        init_states, mb_type, final, final_states"""
        j = MbTypeJob()
        for k, v in kw.items():
            setattr(j, k, v)
        self._check(_lib.h264b_mb_type_decode_dev(self.h, C.byref(j)))

    def stream_submit(self, stream, ops, n_ops, qp, idc, n_ctx, slice_data_offset=0, flags=0, param_sets=None,
                      max_slices=None, max_sps=0, max_pps=0, initial_sps=None, initial_pps=None):
        """asynchronous form: returns (ticket, keepalive); pass both to stream_wait.  Up to three jobs in flight (H264B_STREAM_JOBS_IN_FLIGHT).
        param_sets (ParamSets): take SliceQPY / cabac_init_idc / the CABAC data offset from the slice headers
        (qp, idc may then be None; max_slices bounds the slice count)."""
        s = np.ascontiguousarray(stream, dtype=np.uint8)
        ops = np.ascontiguousarray(ops, dtype=np.uint16)
        p = self.slice_qp(qp, idc) if qp is not None else None
        nops = None if n_ops is None else np.ascontiguousarray(n_ops, dtype=np.uint32)
        j = StreamJob()
        j.stream = s.ctypes.data
        j.n = len(s)
        j.slice_data_offset = slice_data_offset
        j.n_ctx = n_ctx
        j.ops = ops.ctypes.data
        j.n_ops_max = len(ops)
        j.n_ops = nops.ctypes.data if nops is not None else None
        j.qp = p.ctypes.data if p is not None else None
        j.max_slices = len(p) if p is not None else int(max_slices)
        j.flags = flags | (STREAM_SLICE_HEADERS if param_sets is not None else 0)
        if flags & STREAM_PARAM_SETS:   # the stream's own SPS / PPS NAL units
            j.flags |= STREAM_SLICE_HEADERS
        j.param_sets = C.addressof(param_sets) if param_sets is not None else None
        j.max_sps, j.max_pps = max_sps, max_pps
        keep = []
        for name, rec, dt in (("initial_sps", initial_sps, SPS_DTYPE), ("initial_pps", initial_pps, PPS_DTYPE)):
            if rec is not None:   # the sets in force when the batch begins (one SPS_DTYPE / PPS_DTYPE record each)
                a = np.ascontiguousarray(np.asarray(rec, dtype=dt).reshape(1))
                keep.append(a)
                setattr(j, name, a.ctypes.data)
        t = C.c_uint64()
        self._check(_lib.h264b_stream_submit(self.h, C.byref(j), C.byref(t)))
        return t.value, (s, ops, p, nops, j, param_sets, keep)

    def stream_wait(self, ticket, keepalive=None):
        r = StreamResult()
        self._check(_lib.h264b_stream_wait(self.h, ticket, C.byref(r)))
        return self._stream_result(r)

    @staticmethod
    def _stream_result(r):
        ns = r.n_slices
        boff = _from_ptr(r.bins_off, np.uint64, ns + 1)
        flat = _from_ptr(r.bins, np.uint32, int(boff[-1]) if ns else 0)
        return dict(scan=r.scan.as_dict(), nals=_from_ptr(r.nals, NAL_DTYPE, r.scan.n_nals),
                    slice_nal=_from_ptr(r.slice_nal, np.uint32, ns), bins_off=boff, bins_flat=flat,
                    bins=[flat[int(boff[i]):int(boff[i + 1])] for i in range(ns)],
                    final=_from_ptr(r.final, FINAL_DTYPE, ns), total_bins=r.total_bins,
                    headers=_from_ptr(r.headers, SLICE_HEADER_DTYPE, ns) if r.headers else None,
                    sps=_from_ptr(r.sps, SPS_DTYPE, r.n_sps) if r.sps else None,
                    pps=_from_ptr(r.pps, PPS_DTYPE, r.n_pps) if r.pps else None,
                    sps_nal=_from_ptr(r.sps_nal, np.uint32, r.n_sps) if r.sps else None,
                    pps_nal=_from_ptr(r.pps_nal, np.uint32, r.n_pps) if r.pps else None,
                    slice_sps=_from_ptr(r.slice_sps, np.int32, ns) if r.sps else None,
                    slice_pps=_from_ptr(r.slice_pps, np.int32, ns) if r.sps else None)


def scheduler_plan(streams, slices_per_stream, n_ops, n_ops_max, n_devices, sm_count=148, group_bytes=0):
    """h264b_scheduler_plan: host only (works without a GPU).  -> (stream_device, stream_pass, slice_class)"""
    arrs = [np.ascontiguousarray(s, dtype=np.uint8) for s in streams]
    per = np.asarray(slices_per_stream, dtype=np.int64)
    first = np.concatenate([[0], np.cumsum(per)]).astype(np.int64)
    total = int(first[-1])
    bs = (BatchStream * max(len(arrs), 1))()
    for i, a in enumerate(arrs):
        bs[i].stream = a.ctypes.data if len(a) else None
        bs[i].n = len(a)
        bs[i].first_slice = int(first[i])
        bs[i].n_slices = int(per[i])
    nops = None if n_ops is None else np.ascontiguousarray(n_ops, dtype=np.uint32)
    j = BatchJob()
    j.streams = C.addressof(bs)
    j.n_streams = len(arrs)
    j.total_slices = total
    j.n_ctx = 64
    j.n_ops_max = int(n_ops_max)
    j.n_ops = nops.ctypes.data if nops is not None else None
    j.group_bytes = group_bytes
    dev = np.zeros(max(len(arrs), 1), np.int32)
    pas = np.zeros(max(len(arrs), 1), np.uint32)
    cls = np.zeros(max(total, 1), np.uint8)
    rc = _lib.h264b_scheduler_plan(C.byref(j), n_devices, sm_count, dev.ctypes.data, pas.ctypes.data, cls.ctypes.data)
    if rc != OK:
        raise H264BError(rc, "h264b_scheduler_plan")
    return dev[:len(arrs)], pas[:len(arrs)], cls[:total]


class Scheduler:
    """h264b_scheduler: a batch of independent streams over several GPUs of this process (one worker thread and one
    context per device, LPT by bytes, device jobs with the longest slices first, three in flight per device)."""

    def __init__(self, devices):
        devs = np.ascontiguousarray(devices, dtype=np.int32)
        h = C.c_void_p()
        rc = _lib.h264b_scheduler_create(C.c_void_p(devs.ctypes.data), len(devs), C.byref(h))
        if rc != OK:
            raise H264BError(rc, "h264b_scheduler_create failed" + (": no usable CUDA device" if rc == E_NO_DEVICE else ""))
        self.h = h
        self.n_devices = len(devs)

    def close(self):
        if self.h:
            _lib.h264b_scheduler_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, streams, slices_per_stream, ops, n_ops, qp, idc, n_ctx, flags=0, slice_data_offset=0, group_bytes=0):
        """streams: list of uint8 arrays; slices_per_stream[i]: slice NAL units of stream i (rows of n_ops / qp / idc in
        stream order).  -> dict(stream_device, stream_job, nals (list per stream), final, bins (list per slice),
        slice_done_ms, device_busy_ms, device_bytes, device_jobs, makespan_ms, total_bins, total_nals)"""
        arrs = [np.ascontiguousarray(s, dtype=np.uint8) for s in streams]
        per = np.asarray(slices_per_stream, dtype=np.int64)
        first = np.concatenate([[0], np.cumsum(per)]).astype(np.int64)
        total = int(first[-1])
        bs = (BatchStream * max(len(arrs), 1))()
        for i, a in enumerate(arrs):
            bs[i].stream = a.ctypes.data if len(a) else None
            bs[i].n = len(a)
            bs[i].first_slice = int(first[i])
            bs[i].n_slices = int(per[i])
        ops = np.ascontiguousarray(ops, dtype=np.uint16)
        nops = None if n_ops is None else np.ascontiguousarray(n_ops, dtype=np.uint32)
        p = Context.slice_qp(qp, idc)
        j = BatchJob()
        j.streams = C.addressof(bs)
        j.n_streams = len(arrs)
        j.total_slices = total
        j.n_ctx = n_ctx
        j.n_ops_max = len(ops)
        j.ops = ops.ctypes.data if len(ops) else None
        j.n_ops = nops.ctypes.data if nops is not None else None
        j.qp = p.ctypes.data
        j.slice_data_offset = slice_data_offset
        j.flags = flags
        j.group_bytes = group_bytes
        r = BatchResult()
        rc = _lib.h264b_scheduler_run(self.h, C.byref(j), C.byref(r))
        if rc != OK:
            raise H264BError(rc, (_lib.h264b_scheduler_last_error(self.h) or b"").decode(errors="replace"))
        ns = len(arrs)
        noff = _from_ptr(r.stream_nal_off, np.uint64, ns + 1).copy()
        nals = _from_ptr(r.nals, NAL_DTYPE, int(noff[-1]) if ns else 0).copy()
        boff = _from_ptr(r.bins_off, np.uint64, total + 1).copy()
        flat = _from_ptr(r.bins, np.uint32, int(boff[-1]) if total else 0).copy()   # (the library's arena lives until the next run)
        nb = np.minimum(nops, len(ops)) if nops is not None else np.full(total, len(ops), np.uint32)
        words = (nb.astype(np.int64) + 1 + 31) // 32   # (a slice's words lie where its launch copied them)
        return dict(stream_device=_from_ptr(r.stream_device, np.int32, ns).copy(),
                    stream_job=_from_ptr(r.stream_job, np.uint32, ns).copy(),
                    nals=[nals[int(noff[i]):int(noff[i + 1])] for i in range(ns)],
                    final=_from_ptr(r.final, FINAL_DTYPE, total).copy(), bins_off=boff, bins_flat=flat,
                    bins=[flat[int(boff[i]):int(boff[i]) + int(words[i])] for i in range(total)],
                    slice_done_ms=_from_ptr(r.slice_done_ms, np.float64, total).copy(),
                    device_busy_ms=_from_ptr(r.device_busy_ms, np.float64, r.n_devices).copy(),
                    device_bytes=_from_ptr(r.device_bytes, np.uint64, r.n_devices).copy(),
                    device_jobs=_from_ptr(r.device_jobs, np.uint32, r.n_devices).copy(),
                    makespan_ms=r.makespan_ms, total_bins=r.total_bins, total_nals=r.total_nals)
