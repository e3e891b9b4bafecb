/* h264b200.h -- C ABI of libh264b200.so: the B200-native (sm_100a) data-parallel front end of an H.264
 * decoder, a drop-in for the corresponding hot path of mrmod/h264decode (pure Go, package h264).
 *
 * The reference has no FFI layer; its boundary for this path is the exported Go API of package h264.  Each
 * entry point below names the reference function(s) it replaces (paths relative to the reference root).  A Go
 * shim binds these symbols through cgo (INTEGRATION.md shows the stub); tests bind them through ctypes.
 *
 * Conventions
 *   - every function returns an int32 status (H264B_OK == 0); h264b_last_error() gives text for the last failure
 *   - plain pointers and sizes only; no C++ or torch types
 *   - "_dev" entry points take DEVICE pointers, enqueue work on the context's CUDA stream and return without
 *     synchronising (except where stated); all other entry points take HOST pointers and are synchronous
 *   - there is no CPU fallback: without a CUDA device h264b_create fails with H264B_E_NO_DEVICE
 *   - one h264b_ctx drives one GPU; contexts are independent (one per process-rank or several per process);
 *     a context must not be used from two threads at once
 *   - behaviour flags select the reference-exact (REF, default, bit 0) or ITU-T-corrected (SPEC) variant of each
 *     documented deviation of the reference
 */
#ifndef H264B200_H
#define H264B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden; these are its exports */
#endif

#define H264B_VERSION 100 /* 0.1.0 */

/* status codes */
#define H264B_OK 0
#define H264B_E_INVALID 1   /* bad argument */
#define H264B_E_CUDA 2      /* a CUDA call failed; see h264b_last_error */
#define H264B_E_NOMEM 3
#define H264B_E_CAPACITY 4  /* an output buffer was too small; the summary says how much is needed */
#define H264B_E_NO_DEVICE 5 /* no usable CUDA device: this library has no CPU path */

/* behaviour flags */
#define H264B_TABLES_SPEC 0x1u           /* corrected rangeTabLPS / transIdx / (m,n) tables (A1..A3); default REF */
#define H264B_BYPASS_SPEC_OR 0x2u        /* DecodeBypass as (O<<1)|bit (A5); default REF: O<<=1 then O<<=bit */
#define H264B_CABAC_FINAL_TERMINATE 0x4u /* after a slice's n_ops bins decode one more DecodeTerminate bin */
#define H264B_STREAM_WANT_RBSP 0x8u      /* h264b_stream_*: also copy the RBSP buffer and the extension headers back */
#define H264B_STREAM_SLICE_HEADERS 0x10u /* h264b_stream_*: every slice NAL starts with a slice header: parse it on the
                                            device (job.param_sets) and take SliceQPY, cabac_init_idc (I / SI slices:
                                            NoCabacInitIdc) and the start of the CABAC data (the byte boundary after
                                            the header, slice.go:583-587) from it; job.qp and slice_data_offset are
                                            ignored.  A slice whose header the reference cannot walk is not decoded
                                            (its final record carries H264B_F_OVERRUN). */

#define H264B_STREAM_PARAM_SETS 0x20u   /* h264b_stream_* (implies H264B_STREAM_SLICE_HEADERS): the parameter sets come from
                                            the stream's own SPS / PPS NAL units, parsed on the device; a slice uses
                                            the last SPS before it and the last PPS after that SPS (handleConnection's
                                            VideoStreams bookkeeping, h264/server.go:147-162).  job.param_sets is
                                            ignored; a slice without both, or whose parameter sets the reference
                                            cannot parse, gets header status H264B_SH_PANIC and is not decoded. */

/* per-unit flag bits written by kernels (never abort a batch; SURVEY.md §5 failure handling) */
#define H264B_F_OVERRUN 0x1u    /* the reference would have panicked reading past the slice's last byte (A10) */
#define H264B_F_HAS_EPB 0x2u    /* NAL: at least one emulation-prevention byte was removed */
#define H264B_F_SHORT_NAL 0x4u  /* NAL shorter than 8 bytes: the reference's log line server.go:108 may panic (A13) */

typedef struct h264b_ctx h264b_ctx;

/* ------------------------------------------------------------------ context / memory / stream plumbing */
int32_t h264b_version(void);
int32_t h264b_device_count(int32_t *count);
int32_t h264b_create(int32_t device, h264b_ctx **out);
void h264b_destroy(h264b_ctx *ctx);
const char *h264b_last_error(const h264b_ctx *ctx);
/* Run all "_dev" work on the caller's cudaStream_t (e.g. torch's current stream); NULL = the context's own. */
int32_t h264b_set_stream(h264b_ctx *ctx, void *cuda_stream);
int32_t h264b_sync(h264b_ctx *ctx);
/* Pinned host memory: ingest code reads sockets/files straight into it (replaces the growing []byte of
 * H264Reader.BufferToReader, h264/bit_reader.go:27-39). */
int32_t h264b_host_alloc(h264b_ctx *ctx, size_t bytes, void **out);
int32_t h264b_host_free(h264b_ctx *ctx, void *p);
/* Host-side planner, no device work: cut one long Annex-B stream held in host memory into n_ranges byte ranges that
 * can be scanned independently (one per GPU / context).  Each nominal cut k * n / n_ranges moves forward to
 * the next start code 00 00 00 01 (h264/server.go:19, :28-39: the only boundary the reference knows) and a range keeps
 * the 4 bytes that open the next one, because the reference's NAL unit is payload plus the following start code
 * (h264/server.go:64-111).  The NAL units of range r are exactly those of the whole stream whose start code lies in
 * [begin[r], begin[r+1]); offsets in a range's results are relative to begin[r].  Ranges may be empty.  Only the bytes
 * between a nominal cut and the next start code are read (O(n_ranges x NAL size)). */
int32_t h264b_cut_byte_ranges(const uint8_t *stream, uint64_t n, uint32_t n_ranges, uint64_t *begin, uint64_t *end);
int32_t h264b_dev_alloc(h264b_ctx *ctx, size_t bytes, void **out);
int32_t h264b_dev_free(h264b_ctx *ctx, void *p);
int32_t h264b_memcpy_h2d(h264b_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes); /* async on the stream */
int32_t h264b_memcpy_d2h(h264b_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes); /* async on the stream */
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int32_t h264b_launch_count(const h264b_ctx *ctx, uint64_t *count);

/* ------------------------------------------------------------------ Annex-B split + RBSP strip (K1/K2)
 * Replaces, for a whole byte stream at once:
 *   isStartSequence               h264/server.go:28-39      (00 00 00 01 only, A8)
 *   readNalUnit                   h264/server.go:64-111     (NAL k = bytes after start code k up to AND INCLUDING
 *                                                            start code k+1; the NAL after the last start code and the
 *                                                            bytes before the first one are not emitted)
 *   NewNalUnit                    h264/nalUnit.go:75-131    (header parse, extension headers :39-71, and the
 *                                                            emulation-prevention strip loop :106-126 incl. A7)
 *   NalUnit / (*NalUnit).RBSP()   h264/nalUnit.go:3-30,72
 */
typedef struct {
    uint64_t start;       /* stream offset of the NAL's first byte (startOffset, server.go:88) */
    uint64_t rbsp_off;    /* offset of its RBSP in the rbsp output buffer (= start + header_bytes: the RBSP of a NAL
                             is written at the position of the NAL's own body, so RBSPs never overlap and the buffer
                             is as long as the stream; the bytes between two RBSPs are unspecified) */
    uint32_t num_bytes;   /* NalUnit.NumBytes (includes the following start code) */
    uint32_t rbsp_len;    /* len(NalUnit.rbsp) */
    uint8_t forbidden_zero_bit, ref_idc, type, header_bytes; /* nalUnit.go:82-84, :79,94,97,100 */
    uint32_t flags;       /* H264B_F_HAS_EPB | H264B_F_SHORT_NAL */
} h264b_nal; /* 32 bytes */

/* extension-header fields of NAL types 14/20/21 (nalUnit.go:39-71); all zero for other types */
typedef struct {
    uint8_t svc_extension_flag, avc_3d_extension_flag, idr_flag, priority_id;
    uint8_t no_inter_layer_pred_flag, dependency_id, quality_id, temporal_id;
    uint8_t use_ref_base_pic_flag, discardable_flag, output_flag, reserved_three_2bits;
    uint8_t non_idr_flag, anchor_pic_flag, inter_view_flag, reserved_one_bit;
    uint16_t view_id;
    uint8_t view_idx, depth_flag;
    uint32_t pad;
} h264b_nal_ext; /* 24 bytes */

typedef struct {
    uint64_t n_start_codes; /* occurrences of 00 00 00 01 */
    uint64_t n_nals;        /* = max(n_start_codes, 1) - 1 : NAL units the reference would emit */
    uint64_t rbsp_bytes;    /* RBSP bytes of those NAL units */
    uint64_t first_start;   /* offset just after the first start code (== n when there is none) */
    uint64_t n_epb;         /* emulation-prevention bytes removed inside emitted NAL units */
    uint32_t status;        /* H264B_OK or H264B_E_CAPACITY (nal_cap too small; n_nals says what is needed) */
    uint32_t reserved;
} h264b_scan_summary; /* 48 bytes */

/* Bytes of device scratch h264b_annexb_scan_dev needs for an n-byte stream (the context owns and grows it). */
uint64_t h264b_annexb_scratch_bytes(uint64_t n);

/* Device-resident scan.  d_stream: n bytes, 16-byte aligned, readable up to n + 32 (loads are whole 16-byte granules
 * and whole words; what lies past n is never used).  d_rbsp: 16-byte aligned, capacity >= n + 32 (the pass itself
 * writes below n; h264b_cabac_decode_dev on this buffer reads whole words past the end of the last slice).
 * d_nals: nal_cap records.  d_ext: nal_cap records or NULL.  d_summary: one record.  Asynchronous. */
int32_t h264b_annexb_scan_dev(h264b_ctx *ctx, const uint8_t *d_stream, uint64_t n, uint8_t *d_rbsp,
                              h264b_nal *d_nals, h264b_nal_ext *d_ext, uint32_t nal_cap,
                              h264b_scan_summary *d_summary, uint32_t flags);

/* Host-buffer scan (the call a Go caller makes): copies the stream in, runs the kernels, copies the NAL index
 * and the RBSP buffer (n bytes, indexed by h264b_nal.rbsp_off) back into context-owned pinned buffers that stay valid
 * until the next call on ctx.
 * want_rbsp == 0 leaves the RBSP on the device (d_rbsp_out, if not NULL, receives its device address for a
 * following h264b_cabac_decode_dev). */
int32_t h264b_annexb_scan(h264b_ctx *ctx, const uint8_t *stream, uint64_t n, uint32_t flags, int32_t want_rbsp,
                          const h264b_nal **nals, const h264b_nal_ext **ext, h264b_scan_summary *summary,
                          const uint8_t **rbsp, const uint8_t **d_rbsp_out);

/* NewNalUnit(frame, numBytesInNal) for a batch of independent frames (direct-call semantics, including the
 * 00 00 03-at-the-end edge of nalUnit.go:113-117): frames are concatenated in `frames`, frame i occupying
 * [frame_off[i], frame_off[i] + frame_len[i]).  RBSP of frame i is written at rbsp + frame_off[i].
 * Host buffers; synchronous.  nals[i].start = frame_off[i]. */
int32_t h264b_nal_units(h264b_ctx *ctx, const uint8_t *frames, uint64_t total_bytes, const uint64_t *frame_off,
                        const uint32_t *frame_len, uint32_t n_frames, uint32_t flags, h264b_nal *nals,
                        h264b_nal_ext *ext, uint8_t *rbsp);

/* ------------------------------------------------------------------ context-variable initialisation (K4)
 * Replaces PreCtxState (h264/cabac.go:118-121), Clip3 (:131-139), the state split of initCabac (:158-164) and
 * the MNVars / CodedblockPatternMN lookups (h264/mn_vars.go:15-175,184-440), for every (slice, ctxIdx) at once.
 * State byte = pStateIdx | valMPS << 6.  cabac_init_idc -1 (= NoCabacInitIdc, mn_vars.go:7) selects the
 * single-column entries of ctxIdx 0..10 and the I/SI column of ctxIdx 70..104. */
typedef struct {
    int32_t slice_qp_y;     /* SliceQPy (cabac.go:113-115); clipped to 0..51 inside PreCtxState */
    int32_t cabac_init_idc; /* -1, 0, 1, 2; anything else: Go missing-key semantics (MN{0,0}, I column for 70..104) */
} h264b_slice_qp;

int32_t h264b_ctx_init_dev(h264b_ctx *ctx, const h264b_slice_qp *d_params, uint32_t n_slices, uint32_t n_ctx,
                           uint8_t *d_states /* [n_slices][n_ctx] */, uint32_t flags);
int32_t h264b_ctx_init(h264b_ctx *ctx, const h264b_slice_qp *params, uint32_t n_slices, uint32_t n_ctx,
                       uint8_t *states, uint32_t flags);
/* scalar drop-ins (each runs a one-thread kernel: there is no CPU implementation in this library) */
int32_t h264b_pre_ctx_state(h264b_ctx *ctx, int32_t m, int32_t n, int32_t slice_qp_y, int32_t *pre_ctx_state);
int32_t h264b_mn(h264b_ctx *ctx, int32_t ctx_idx, int32_t cabac_init_idc, uint32_t flags, int32_t *m, int32_t *n);

/* ------------------------------------------------------------------ CABAC arithmetic decoding engine (K3)
 * Replaces initDecodingEngine (h264/cabac.go:439-446), the arithmetic core of BinaryDecision (:525-536) composed
 * with StateTransitionProcess (:544-553) and RenormD (:503-511) as "DecodeDecision", DecodeBypass (:468-481),
 * DecodeTerminate (:486-499), with rangeTabLPS (h264/rangeTabLPS.go) and stateTransxTab (h264/stateTransxTab.go).
 * One slice per warp lane; all slices follow one shared op schedule, slice s using its first n_ops[s] entries. */
#define H264B_OP_DECISION 0u
#define H264B_OP_BYPASS 1u
#define H264B_OP_TERMINATE 2u
#define H264B_OP(kind, ctx_idx) ((uint16_t)(((kind) << 14) | ((ctx_idx)&0x3FFu)))

typedef struct {
    int64_t cod_i_range;
    int64_t cod_i_offset;
    uint64_t bits_read; /* BitReader.bitsRead relative to the slice's first byte: 9 + renorm shifts + bypass bins */
    uint32_t flags;     /* H264B_F_OVERRUN: everything else in this record and the slice's bins are unspecified */
    uint32_t n_bins;    /* bins produced (n_ops[s], +1 with H264B_CABAC_FINAL_TERMINATE) */
} h264b_cabac_final;    /* 32 bytes */

typedef struct {
    const uint8_t *bytes;          /* slice data; slice s = bytes[off[s] .. off[s]+len[s]) */
    uint64_t total_bytes;          /* readable bytes at `bytes` (loads are clamped to it) */
    const uint64_t *off;           /* [n_slices] */
    const uint32_t *len;           /* [n_slices] */
    uint32_t n_slices;
    uint32_t n_ctx;                /* context variables per slice, 1..1024 */
    const uint16_t *ops;           /* [n_ops_max] shared schedule, H264B_OP(kind, ctxIdx) */
    uint32_t n_ops_max;
    const uint32_t *n_ops;         /* [n_slices] or NULL (= n_ops_max for every slice) */
    const h264b_slice_qp *qp;      /* [n_slices]: initial states by the K4 rule; used when init_states == NULL */
    const uint8_t *init_states;    /* [n_slices][n_ctx] or NULL */
    uint32_t *bins;                /* bin i of slice s = bit (i & 31) of word (i >> 5) of the slice's row */
    const uint64_t *bins_off;      /* [n_slices] word offset of each slice's row in `bins` (compact layout), or NULL:
                                      row s starts at s * bins_stride_words */
    uint32_t bins_stride_words;    /* >= (n_ops_max + 1 + 31) / 32 when bins_off == NULL */
    h264b_cabac_final *final;      /* [n_slices] */
    uint8_t *final_states;         /* [n_slices][n_ctx] or NULL */
    uint32_t flags;
    uint32_t n_ctx_used;           /* 0, or a promise: the schedule's decisions only use ctxIdx < n_ctx_used (<= n_ctx).  Only
                                      those context rows are then kept in shared memory (32 bytes per context and warp: what
                                      decides how many slices an SM decodes at once); the others pass from init to final
                                      untouched.  A ctxIdx >= n_ctx_used in the schedule is treated like one >= n_ctx: as 0. */
} h264b_cabac_job;

/* all pointers in job are DEVICE pointers; asynchronous */
int32_t h264b_cabac_decode_dev(h264b_ctx *ctx, const h264b_cabac_job *job);
/* all pointers in job are HOST pointers; synchronous */
int32_t h264b_cabac_decode(h264b_ctx *ctx, const h264b_cabac_job *job);

/* Single engine steps with explicit state, for per-call drop-in use of the Go functions (each runs a one-thread
 * kernel).  bits/n_bits: the bits the step may consume, MSB-first in `bits`; *bits_used reports how many it took.
 * kind: H264B_OP_*; for a decision *p_state_idx / *val_mps are the context and are updated. */
int32_t h264b_engine_step(h264b_ctx *ctx, uint32_t kind, uint32_t flags, const uint8_t *bits, uint32_t n_bits,
                          int64_t *cod_i_range, int64_t *cod_i_offset, int32_t *p_state_idx, int32_t *val_mps,
                          int32_t *bin_val, uint32_t *bits_used);
/* the un-composed reference primitives (cabac.go:525-536 and :544-553 on their own, A6) */
int32_t h264b_binary_decision(h264b_ctx *ctx, uint32_t flags, int32_t p_state_idx, int32_t val_mps,
                              int64_t *cod_i_range, int64_t *cod_i_offset, int32_t *bin_val);
int32_t h264b_state_transition(h264b_ctx *ctx, uint32_t flags, int32_t *p_state_idx, int32_t *val_mps,
                               int32_t bin_val);

/* ------------------------------------------------------------------ slice headers ("next" row f1)
 * Replaces the header walk of NewSliceContext (h264/slice.go:835-1048, everything before NewSliceData) for all slice
 * NAL units at once, one thread per slice: what the CABAC stage needs from the stream instead of from the caller --
 * SliceQPY (SliceQPy, h264/cabac.go:113-115), cabac_init_idc, the slice type and the bit at which slice_data()
 * begins.  The reference's deviations from ITU-T H.264 are reproduced (slice_header.cuh lists them). */
typedef struct { /* the SPS / PPS fields the walk reads (h264/sps.go, h264/pps.go); Go ints -> int64 */
    int64_t use_separate_color_plane, chroma_format, frame_mbs_only, pic_order_count_type;
    int64_t log2_max_pic_order_cnt_lsb_min4, delta_pic_order_always_zero;
    int64_t bottom_field_pic_order_in_frame_present, redundant_pic_cnt_present, weighted_pred, weighted_bipred;
    int64_t entropy_coding_mode, deblocking_filter_control_present, num_slice_groups_minus1, slice_group_map_type;
    int64_t pic_size_in_map_units_minus1, slice_group_change_rate_minus1, pic_init_qp_minus26, reserved;
} h264b_param_sets; /* 144 bytes; h264b_make_param_sets fills it from a parsed SPS + PPS */

#define H264B_SH_OK 0u
#define H264B_SH_PANIC 1u /* the reference would have panicked (read past the end of the RBSP, index out of range,
                             divide by zero); the other fields are unspecified */
#define H264B_SH_HANG 2u  /* the reference would loop forever (memory_management_control_operation 5 or >= 7) */

typedef struct { /* SliceHeader, h264/slice.go:35-103; Go ints -> int64 */
    int64_t first_mb_in_slice, slice_type, pps_id, color_plane_id, field_pic, bottom_field, idr_pic_id;
    int64_t pic_order_cnt_lsb, delta_pic_order_cnt_bottom, delta_pic_order_cnt[2], redundant_pic_cnt;
    int64_t direct_spatial_mv_pred, num_ref_idx_active_override, num_ref_idx_l0_active_minus1;
    int64_t num_ref_idx_l1_active_minus1, ref_pic_list_modification_flag_l0, ref_pic_list_modification_flag_l1;
    int64_t modification_of_pic_nums, abs_diff_pic_num_minus1, long_term_pic_num;
    int64_t luma_log2_weight_denom, chroma_log2_weight_denom;
    int64_t n_luma_weight_l0, n_chroma_weight_l0, n_luma_weight_l1, n_chroma_weight_l1; /* list lengths only */
    int64_t no_output_of_prior_pics_flag, long_term_reference_flag, adaptive_ref_pic_marking_mode_flag;
    int64_t memory_management_control_operation, difference_of_pic_nums_minus1, long_term_frame_idx;
    int64_t max_long_term_frame_idx_plus1, cabac_init_idc, slice_qp_delta, sp_for_switch, slice_qs_delta;
    int64_t disable_deblocking_filter, slice_alpha_c0_offset_div2, slice_beta_offset_div2, slice_group_change_cycle;
    int64_t chroma_array_type;
    int64_t slice_qp_y;   /* 26 + pic_init_qp_minus26 + slice_qp_delta */
    uint64_t header_bits; /* BitReader.bitsRead at the end of the header */
    uint32_t status;      /* H264B_SH_* */
    uint32_t reserved;
} h264b_slice_header; /* 46 x 8 + 8 bytes */

/* Slice s = RBSP bytes [off[s], off[s] + len[s]) of `bytes`, from the NAL with (nal_type[s], nal_ref_idc[s]).  Device
 * pointers; asynchronous.  With d_nals != NULL, off/len/type/ref_idc are taken from nals[slice_nal[s]] instead (the
 * output of h264b_annexb_scan_dev + h264b_slice_select_dev with slice_data_offset 0) and d_off .. d_ref_idc may be
 * NULL. */
int32_t h264b_slice_headers_dev(h264b_ctx *ctx, const h264b_param_sets *params /* host */, const uint8_t *d_bytes,
                                uint64_t total_bytes, const uint64_t *d_off, const uint32_t *d_len,
                                const uint8_t *d_nal_type, const uint8_t *d_nal_ref_idc, const h264b_nal *d_nals,
                                const uint32_t *d_slice_nal, uint32_t n_slices, h264b_slice_header *d_out);
/* host pointers; synchronous */
int32_t h264b_slice_headers(h264b_ctx *ctx, const h264b_param_sets *params, const uint8_t *bytes, uint64_t total_bytes,
                            const uint64_t *off, const uint32_t *len, const uint8_t *nal_type,
                            const uint8_t *nal_ref_idc, uint32_t n_slices, h264b_slice_header *out);

/* ------------------------------------------------------------------ parameter sets (rows S1 / f4)
 * Replaces NewSPS (h264/sps.go:192-437, with scalingList :172-191) and NewPPS (h264/pps.go:40-133): the field
 * extraction from the RBSP of NAL units of type 7 / 8, one device thread per parameter set, so that the active
 * parameter sets of a stream stay on the device for the slice-header walk.  The reference's deviations from ITU-T
 * H.264 are reproduced (param_sets.cuh lists them).  Go ints and bools -> int64; `status` is H264B_SH_OK or
 * H264B_SH_PANIC (the reference would have panicked: the fields read before that point are set, the rest are 0). */
#define H264B_SPS_MAX_REF_FRAMES 256 /* entries kept of OffsetForRefFrameList (n_offset_for_ref_frame counts all) */
#define H264B_SPS_MAX_HRD 64         /* entries kept of BitRateValueMinus1 / CpbSizeValueMinus1 / Cbr (n_hrd counts all) */
typedef struct { /* SPS, h264/sps.go:13-103 */
    int64_t profile, constraint0, constraint1, constraint2, constraint3, constraint4, constraint5, level, id;
    int64_t chroma_format, use_separate_color_plane, bit_depth_luma_minus8, bit_depth_chroma_minus8;
    int64_t qprime_y_zero_transform_bypass, seq_scaling_matrix_present, log2_max_frame_num_minus4;
    int64_t pic_order_count_type, log2_max_pic_order_cnt_lsb_min4, delta_pic_order_always_zero;
    int64_t offset_for_non_ref_pic, offset_for_top_to_bottom_field, num_ref_frames_in_pic_order_cnt_cycle;
    int64_t max_num_ref_frames, gaps_in_frame_num_value_allowed, pic_width_in_mbs_minus1;
    int64_t pic_height_in_map_units_minus1, frame_mbs_only, mb_adaptive_frame_field, direct_8x8_inference;
    int64_t frame_cropping, frame_crop_left_offset, frame_crop_right_offset, frame_crop_top_offset;
    int64_t frame_crop_bottom_offset, vui_parameters_present, aspect_ratio_info_present, aspect_ratio, sar_width;
    int64_t sar_height, overscan_info_present, overscan_appropriate, video_signal_type_present, video_format;
    int64_t video_full_range, color_description_present, color_primaries, transfer_characteristics;
    int64_t matrix_coefficients, chroma_loc_info_present, chroma_sample_loc_type_top_field;
    int64_t chroma_sample_loc_type_bottom_field, cpb_cnt_minus1, bit_rate_scale, cpb_size_scale;
    int64_t initial_cpb_removal_delay_length_minus1, cpb_removal_delay_length_minus1;
    int64_t dpb_output_delay_length_minus1, time_offset_length, timing_info_present, num_units_in_tick, time_scale;
    int64_t nal_hrd_parameters_present, fixed_frame_rate, vcl_hrd_parameters_present, low_hrd_delay;
    int64_t pic_struct_present, bitstream_restriction, motion_vectors_over_pic_boundaries, max_bytes_per_pic_denom;
    int64_t max_bits_per_mb_denom, log2_max_mv_length_horizontal, log2_max_mv_length_vertical;
    int64_t max_dec_frame_buffering, max_num_reorder_frames; /* 74 scalars */
    int64_t n_seq_scaling_list, seq_scaling_list[12];         /* SeqScalingList (the present flags) */
    int64_t n_offset_for_ref_frame, offset_for_ref_frame[H264B_SPS_MAX_REF_FRAMES];
    int64_t n_hrd, bit_rate_value_minus1[H264B_SPS_MAX_HRD], cpb_size_value_minus1[H264B_SPS_MAX_HRD],
        cbr[H264B_SPS_MAX_HRD];
    uint64_t bits_read; /* BitReader.bitsRead when NewSPS returned (or panicked) */
    uint32_t status, reserved;
} h264b_sps;

typedef struct { /* PPS, h264/pps.go:10-38 */
    int64_t id, sps_id, entropy_coding_mode, num_slice_groups_minus1, bottom_field_pic_order_in_frame_present;
    int64_t slice_group_map_type, slice_group_change_direction, slice_group_change_rate_minus1;
    int64_t pic_size_in_map_units_minus1, num_ref_idx_l0_default_active_minus1;
    int64_t num_ref_idx_l1_default_active_minus1, weighted_pred, weighted_bipred, pic_init_qp_minus26;
    int64_t pic_init_qs_minus26, chroma_qp_index_offset, deblocking_filter_control_present, constrained_intra_pred;
    int64_t redundant_pic_cnt_present, transform_8x8_mode, pic_scaling_matrix_present;
    int64_t second_chroma_qp_index_offset; /* 22 scalars */
    uint64_t bits_read;
    uint32_t status, reserved;
} h264b_pps;

/* Parameter set i = RBSP bytes [off[i], off[i] + len[i]) of `bytes`.  Host pointers; synchronous. */
int32_t h264b_parse_sps(h264b_ctx *ctx, const uint8_t *bytes, uint64_t total_bytes, const uint64_t *off,
                        const uint32_t *len, uint32_t n, h264b_sps *out);
int32_t h264b_parse_pps(h264b_ctx *ctx, const uint8_t *bytes, uint64_t total_bytes, const uint64_t *off,
                        const uint32_t *len, uint32_t n, h264b_pps *out);
/* Device pointers; asynchronous.  With d_nals != NULL parameter set i is the RBSP of nals[nal_index[i]] and d_off /
 * d_len may be NULL; d_n (device, may be NULL) holds the actual count, n is then the bound. */
int32_t h264b_parse_sps_dev(h264b_ctx *ctx, const uint8_t *d_bytes, uint64_t total_bytes, const uint64_t *d_off,
                            const uint32_t *d_len, const h264b_nal *d_nals, const uint32_t *d_nal_index, uint32_t n,
                            const uint32_t *d_n, h264b_sps *d_out);
int32_t h264b_parse_pps_dev(h264b_ctx *ctx, const uint8_t *d_bytes, uint64_t total_bytes, const uint64_t *d_off,
                            const uint32_t *d_len, const h264b_nal *d_nals, const uint32_t *d_nal_index, uint32_t n,
                            const uint32_t *d_n, h264b_pps *d_out);
/* The fields the slice-header walk reads, from a parsed SPS + PPS (host-side convenience; no device work). */
int32_t h264b_make_param_sets(const h264b_sps *sps, const h264b_pps *pps, h264b_param_sets *out);
/* Ordered lists of the NAL units of type 7 and of type 8 of a finished h264b_annexb_scan_dev (handleConnection's
 * dispatch, h264/server.go:147-158).  Device pointers; d_counts[0] / [1] receive min(count, max_sps / max_pps). */
int32_t h264b_param_set_select_dev(h264b_ctx *ctx, const h264b_nal *d_nals, const h264b_scan_summary *d_summary,
                                   uint32_t nal_cap, uint32_t max_sps, uint32_t max_pps, uint32_t *d_sps_nal,
                                   uint32_t *d_pps_nal, uint32_t *d_counts);

/* ------------------------------------------------------------------ syntax-element glue (rows I5 / f3)
 * Replaces, for batches of queries (one device thread each; host pointers, synchronous):
 *   CtxIdx            h264/cabac.go:557-758   (Table 9-39 as the reference has it; maxBinIdxCtx is never read)
 *   NewBinarization   h264/cabac.go:340-427   (Table 9-34 rows)
 *   initCabac         h264/cabac.go:148-174   (CtxIdx -> MNVars[ctxIdx][0] -> PreCtxState(SliceQPy) -> state split)
 *   binIdxMbMap / binIdxSubMbMap, (*Binarization).IsBinStringMatch   h264/cabac.go:180-303, :429-436 */
#define H264B_NA_CTX_ID 10000 /* NaCtxId, cabac.go:4 */
#define H264B_NA_SUFFIX (-1)  /* NA_SUFFIX, cabac.go:5 */
enum { /* syntax element names NewBinarization knows (anything else gives the zero value) */
    H264B_SE_CODED_BLOCK_PATTERN, H264B_SE_INTRA_CHROMA_PRED_MODE, H264B_SE_MB_QP_DELTA, H264B_SE_MVD_LN_END0,
    H264B_SE_MVD_LN_END1, H264B_SE_MB_TYPE, H264B_SE_MB_FIELD_DECODING_FLAG, H264B_SE_PREV_INTRA4X4_PRED_MODE_FLAG,
    H264B_SE_PREV_INTRA8X8_PRED_MODE_FLAG, H264B_SE_REF_IDX_L0, H264B_SE_REF_IDX_L1, H264B_SE_REM_INTRA4X4_PRED_MODE,
    H264B_SE_REM_INTRA8X8_PRED_MODE, H264B_SE_TRANSFORM_SIZE_8X8_FLAG, H264B_SE_OTHER
};
enum { H264B_ST_P, H264B_ST_B, H264B_ST_I, H264B_ST_SP, H264B_ST_SI, H264B_ST_NONE }; /* sliceTypeMap names, slice.go:105 */
typedef struct { /* Binarization, cabac.go:318-338 (its private binIdx / binString travel separately) */
    int32_t syntax_element;
    int32_t prefix_suffix, fixed_length, unary, truncated_unary, cmax, uegk, cmax_value; /* BinarizationType */
    int32_t max_is_prefix_suffix, max_prefix, max_suffix;                                /* MaxBinIdxCtx */
    int32_t off_is_prefix_suffix, off_prefix, off_suffix;                                /* CtxIdxOffset */
    int32_t use_decode_bypass, reserved;
} h264b_binarization; /* 64 bytes */
int32_t h264b_ctx_idx(h264b_ctx *ctx, uint32_t n, const int64_t *bin_idx, const int64_t *max_bin_idx_ctx,
                      const int64_t *ctx_idx_offset, int64_t *out);
int32_t h264b_new_binarization(h264b_ctx *ctx, uint32_t n, const int32_t *syntax_element, const int32_t *slice_type_name,
                               h264b_binarization *out);
/* initCabac for query i: binarization fields (bin_idx, max_prefix, off_prefix) and the slice's qp terms ->
 * (PStateIdx, ValMPS); ctx_idx_out (may be NULL) receives the CtxIdx result it used */
int32_t h264b_init_cabac(h264b_ctx *ctx, uint32_t flags, uint32_t n, const int64_t *bin_idx, const int64_t *max_prefix,
                         const int64_t *off_prefix, const int64_t *pic_init_qp_minus26, const int64_t *slice_qp_delta,
                         int32_t *p_state_idx, int32_t *val_mps, int64_t *ctx_idx_out);
/* bin string of mb_type (sub_mb == 0) / sub_mb_type (sub_mb != 0): len[i] elements, element k in bit k of bits[i] */
int32_t h264b_mb_bin_string(h264b_ctx *ctx, uint32_t n, const int32_t *slice_type_name, const int64_t *mb_type,
                            const uint8_t *sub_mb, int32_t *len, uint32_t *bits);
/* IsBinStringMatch(bits): 1 match, 0 no match, 2 the reference panics (more bits than the bin string holds after a
 * matching prefix) */
int32_t h264b_bin_string_match(h264b_ctx *ctx, uint32_t n, const int32_t *bin_len, const uint32_t *bin_bits,
                               const int32_t *n_bits, const uint32_t *bits, int32_t *out);

/* ------------------------------------------------------------------ whole front end of one stream
 * split + strip + (slice NALs of type 1 / 5) context init + CABAC bins, the slice data staying on the device
 * between the stages.  The CABAC data of a slice NAL starts at RBSP byte `slice_data_offset` (the reference's
 * slice-header parser, h264/slice.go:835-1048, is not on this path: SURVEY.md §8 f1).  Slice j (j-th NAL of type
 * 1 or 5 in stream order) uses qp[j] and n_ops[j].  Host pointers; synchronous; outputs in context-owned pinned
 * memory valid until the next call on ctx. */
typedef struct {
    const uint8_t *stream;
    uint64_t n;
    uint32_t slice_data_offset;
    uint32_t n_ctx;
    const uint16_t *ops;
    uint32_t n_ops_max;
    const uint32_t *n_ops;        /* [max_slices] or NULL */
    const h264b_slice_qp *qp;     /* [max_slices] */
    uint32_t max_slices;          /* 0: split + strip only (no CABAC stage; ops / n_ops / qp are ignored) */
    uint32_t flags;
    const h264b_param_sets *param_sets; /* H264B_STREAM_SLICE_HEADERS: the active SPS / PPS fields (host pointer) */
    uint32_t max_sps, max_pps;    /* H264B_STREAM_PARAM_SETS: bounds on the SPS / PPS NAL units kept (0: 64 each) */
    /* H264B_STREAM_PARAM_SETS, batched ingest: the parameter sets in force when this batch begins (host pointers or
     * NULL), i.e. the last SPS of the batches before and the last PPS behind it.  Slices in front of the batch's first
     * SPS use them (slice_sps / slice_pps then read -2); a PPS of this batch in front of its first SPS replaces
     * initial_pps, as handleConnection stores it into the current VideoStream (h264/server.go:153-155). */
    const h264b_sps *initial_sps;
    const h264b_pps *initial_pps;
} h264b_stream_job;

typedef struct {
    h264b_scan_summary scan;
    const h264b_nal *nals;          /* [scan.n_nals] */
    uint32_t n_slices;
    uint32_t reserved;
    const uint32_t *slice_nal;      /* [n_slices] index into nals */
    const uint64_t *bins_off;       /* [n_slices + 1] word offsets: slice s owns bins[bins_off[s] .. bins_off[s+1]) */
    const uint32_t *bins;           /* compact: (n_ops[s] + 1 + 31) / 32 words per slice */
    const h264b_cabac_final *final; /* [n_slices] */
    uint64_t total_bins;
    const uint8_t *rbsp;            /* job.n bytes, indexed by h264b_nal.rbsp_off (H264B_STREAM_WANT_RBSP), else NULL */
    const uint8_t *d_rbsp;          /* the same buffer on the device (for follow-up h264b_*_dev calls) */
    const h264b_nal_ext *ext;       /* [scan.n_nals] (H264B_STREAM_WANT_RBSP), else NULL */
    const h264b_slice_header *headers; /* [n_slices] (H264B_STREAM_SLICE_HEADERS), else NULL */
    /* H264B_STREAM_PARAM_SETS (else 0 / NULL): the stream's parameter sets in stream order */
    uint32_t n_sps, n_pps;
    const h264b_sps *sps;           /* [n_sps] */
    const h264b_pps *pps;           /* [n_pps] */
    const uint32_t *sps_nal;        /* [n_sps] index into nals */
    const uint32_t *pps_nal;        /* [n_pps] */
    const int32_t *slice_sps;       /* [n_slices] index into sps of the slice's active SPS, -1: none, -2: job.initial_sps */
    const int32_t *slice_pps;       /* [n_slices] index into pps, -1: none after that SPS, -2: job.initial_pps */
} h264b_stream_result;

int32_t h264b_stream_decode(h264b_ctx *ctx, const h264b_stream_job *job, h264b_stream_result *result);

/* The same, asynchronously ("batch submit + wait"): up to H264B_STREAM_JOBS_IN_FLIGHT (3) jobs per context, each with
 * its own buffers and compute stream, so that the host -> device
 * copy of job k+1, the kernels of job k and the device -> host copy of job k-1 overlap -- what an ingest loop that
 * batches NAL units into pinned buffers (the modified handleConnection, h264/server.go:113-166) needs.  The job's
 * host buffers must stay valid and unchanged until h264b_stream_wait returns for its ticket (they should be pinned:
 * h264b_host_alloc).  A result stays valid until H264B_STREAM_JOBS_IN_FLIGHT further submits on ctx.  h264b_stream_decode is
 * submit + wait. */
#define H264B_STREAM_JOBS_IN_FLIGHT 3
int32_t h264b_stream_submit(h264b_ctx *ctx, const h264b_stream_job *job, uint64_t *ticket);
int32_t h264b_stream_wait(h264b_ctx *ctx, uint64_t ticket, h264b_stream_result *result);

/* Device-resident building block of the above: the ordered list of slice NAL units (type 1 / 5, the ones
 * handleConnection hands to the slice parser, h264/server.go:147-162) of a finished h264b_annexb_scan_dev, as
 * (rbsp offset + slice_data_offset, remaining length) pairs ready for h264b_cabac_decode_dev.  All pointers are
 * device pointers; d_n_slices receives min(count, max_slices).  Asynchronous. */
int32_t h264b_slice_select_dev(h264b_ctx *ctx, const h264b_nal *d_nals, const h264b_scan_summary *d_summary,
                               uint32_t nal_cap, uint32_t slice_data_offset, uint32_t max_slices, uint64_t *d_off,
                               uint32_t *d_len, uint32_t *d_slice_nal, uint32_t *d_n_slices);

/* ------------------------------------------------------------------ mb_type as a syntax element (row f3)
 * The walk the reference sketches at h264/slice.go:639-672: NewBinarization("MbType") (h264/cabac.go:340-427), then bin
 * after bin  CtxIdx(binIdx, MaxBinIdxCtx.Prefix, CtxIdxOffset.Prefix) (:557-758) -> decode -> IsBinStringMatch against
 * binIdxMbMap[sliceTypeName] (:180-303, :429-436), composed with the engine (the reference reads raw bits there and
 * never decodes).  Unlike h264b_cabac_decode, where every slice follows one shared op schedule, each lane's sequence of
 * (context, decision | terminate) ops depends on the bins it has decoded: the binarisation is a trie in shared memory
 * (built once per context from the reference's tables), a lane's position in it names the next context.
 * Where CtxIdx leaves a binIdx to a "9.3.3.1.x" comment and answers NaCtxId the rule of that clause is used:
 *   I slices (ctxIdxOffset 3): binIdx 0 -> ctxIdxInc = 1 if a macroblock was decoded before this one in the slice and it
 *   was not I_NxN, else 0 (neighbour A = the previous macroblock, B not available); binIdx 1 -> 276 = DecodeTerminate;
 *   binIdx 4 -> b3 != 0 ? 5 : 6; binIdx 5 -> b3 != 0 ? 6 : 7.   P / SP slices (prefix offset 14): binIdx 2 -> b1 != 1 ?
 *   2 : 3; the prefix bin 1 is followed by the I-slice bin string on the suffix offset 17 (binIdx 4 -> b3 != 0 ? 2 : 3)
 *   and mb_type = 5 + that value.
 * I_PCM (the terminate bin of 1) ends a slice's walk after that element (pcm samples follow, not CABAC data). */
typedef struct {
    int64_t cod_i_range;
    int64_t cod_i_offset;
    uint64_t bits_read;
    uint32_t flags;     /* H264B_F_OVERRUN */
    uint32_t n_bins;    /* bins decoded */
    uint32_t n_mb;      /* mb_type elements decoded */
    uint32_t reserved;
} h264b_mb_final;       /* 40 bytes */

typedef struct {
    const uint8_t *bytes;          /* slice data; slice s = bytes[off[s] .. off[s]+len[s]), 4-byte aligned base */
    uint64_t total_bytes;
    const uint64_t *off;           /* [n_slices] */
    const uint32_t *len;           /* [n_slices] */
    uint32_t n_slices;
    uint32_t n_ctx;                /* context variables per slice, 21..1024 (mb_type uses ctxIdx 3..10 and 14..20) */
    const uint8_t *slice_kind;     /* [n_slices] 0: I slice (Table 9-36), 1: P / SP slice (Table 9-37 + the I suffix) */
    const uint32_t *n_mb;          /* [n_slices] mb_type elements to decode, each <= n_mb_max */
    uint32_t n_mb_max;
    uint32_t flags;                /* H264B_TABLES_SPEC */
    const h264b_slice_qp *qp;      /* [n_slices]: initial states by the K4 rule; used when init_states == NULL */
    const uint8_t *init_states;    /* [n_slices][n_ctx] or NULL */
    uint8_t *mb_type;              /* [n_slices][n_mb_max] */
    h264b_mb_final *final;         /* [n_slices] */
    uint8_t *final_states;         /* [n_slices][n_ctx] or NULL */
} h264b_mb_type_job;

/* all pointers in job are DEVICE pointers; asynchronous */
int32_t h264b_mb_type_decode_dev(h264b_ctx *ctx, const h264b_mb_type_job *job);
/* all pointers in job are HOST pointers; synchronous */
int32_t h264b_mb_type_decode(h264b_ctx *ctx, const h264b_mb_type_job *job);

/* ------------------------------------------------------------------ many streams over the GPUs of one box
 * The reference's unit of concurrency is a connection: main.go:16-21 starts one goroutine per accepted connection,
 * each running ByteStreamReader -> handleConnection (h264/server.go:113-166) on its own stream.  Here a scheduler owns
 * one context and one worker thread per device and takes a whole batch of independent streams ("multi-camera batch",
 * BASELINE configs[4]):
 *   - streams are dealt to the devices longest first onto the least loaded one (LPT by bytes);
 *   - a device's share is taken in up to three passes, the streams with the longest slices first (a twelfth of the
 *     share's bytes, then up to one half, then the rest): a slice is serial work, so the slices everything waits for are
 *     staged, copied and started within milliseconds, while the bulk of the share is still being staged;
 *   - a pass goes through one split + strip pass; its slices then run in up to six CABAC launches by length, the longest
 *     class first, side by side on streams of their own: the slices longer than 0.7 of the share's longest -- the chains
 *     the makespan hangs on -- one per warp, four to an SM that they have to themselves; then more than 1/2, 1/8, 1/32,
 *     1/128 of the pass's longest slice, and the rest: the launch with the 1 MB slices lasts two orders of magnitude
 *     longer than the one with the 1 KB slices, whose results reach the host long before;
 *   - every slice's result carries the time at which it reached host memory (tail latency), every device the time it
 *     was busy.
 * No collective, no exchange between devices: slices share nothing.
 * A stream is taken from its first to its last start code 00 00 00 01 (what comes before and after yields no NAL unit
 * in the reference either: h264/server.go:64-111); its results are those of h264b_stream_decode on it alone. */
typedef struct h264b_scheduler h264b_scheduler;
int32_t h264b_scheduler_create(const int32_t *devices, uint32_t n_devices, h264b_scheduler **out);
void h264b_scheduler_destroy(h264b_scheduler *s);
const char *h264b_scheduler_last_error(const h264b_scheduler *s);

typedef struct {
    const uint8_t *stream;  /* host memory */
    uint64_t n;
    uint32_t first_slice;   /* its slice NAL units (type 1 / 5, in stream order) are rows first_slice .. + n_slices - 1 */
    uint32_t n_slices;      /* of the batch's per-slice arrays */
} h264b_batch_stream;

typedef struct {
    const h264b_batch_stream *streams;
    uint32_t n_streams;
    uint32_t total_slices;
    uint32_t n_ctx;
    uint32_t n_ops_max;
    const uint16_t *ops;       /* shared op schedule */
    const uint32_t *n_ops;     /* [total_slices] or NULL */
    const h264b_slice_qp *qp;  /* [total_slices] */
    uint32_t slice_data_offset;
    uint32_t flags;            /* H264B_TABLES_SPEC | H264B_BYPASS_SPEC_OR | H264B_CABAC_FINAL_TERMINATE */
    uint64_t group_bytes;      /* a device's share of at most this many stream bytes is taken in one pass (0: 32 MiB) */
} h264b_batch_job;

typedef struct {
    const int32_t *stream_device;    /* [n_streams] index into the scheduler's devices */
    const uint32_t *stream_job;      /* [n_streams] pass of that device the stream ran in (0: the first) */
    const uint64_t *stream_nal_off;  /* [n_streams + 1] the stream's NAL units are nals[stream_nal_off[i] .. [i + 1]) */
    const h264b_nal *nals;           /* start / rbsp_off relative to the stream's own first byte */
    const h264b_cabac_final *final;  /* [total_slices] */
    const uint64_t *bins_off;        /* [total_slices + 1]: slice r owns bins[bins_off[r] .. + (final[r].n_bins + 31) / 32)
                                        (its launch copied them there; the offsets are in no particular order);
                                        bins_off[total_slices] = words in all */
    const uint32_t *bins;            /* pinned host memory */
    const double *slice_done_ms;     /* [total_slices] from the start of the run to the slice's result in host memory */
    uint32_t n_devices;
    uint32_t reserved;
    const double *device_busy_ms;    /* [n_devices] first submit to last wait */
    const uint64_t *device_bytes;    /* [n_devices] stream bytes dealt to the device */
    const uint32_t *device_jobs;     /* [n_devices] passes (1..3, or 0 for a device without a stream) */
    double makespan_ms;
    uint64_t total_bins, total_nals;
} h264b_batch_result;

/* Synchronous; the result stays valid until the next run or the scheduler's destruction. */
int32_t h264b_scheduler_run(h264b_scheduler *s, const h264b_batch_job *job, h264b_batch_result *res);

/* Host only (no device is touched; the streams' bytes are searched for their first and last start code): how
 * h264b_scheduler_run deals this batch over n_devices devices of sm_count SMs each -- the same code decides there.
 * Outputs, each may be NULL:
 *   stream_device[n_streams]   index of the device, -1: the stream holds no NAL unit
 *   stream_pass[n_streams]     pass of that device the stream is taken in (0: the first)
 *   slice_class[total_slices]  launch class inside its pass: 0 = one slice per warp on an SM of their own, 1..5 by
 *                              length (longest first); 255: the slice's stream goes nowhere */
int32_t h264b_scheduler_plan(const h264b_batch_job *job, uint32_t n_devices, uint32_t sm_count, int32_t *stream_device,
                             uint32_t *stream_pass, uint8_t *slice_class);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif
