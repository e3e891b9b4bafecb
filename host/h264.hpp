// host/h264.hpp -- C++ host side above the C ABI of libh264b200.so, mirroring the exported API of the reference's
// Go package `h264` for the hot path (same names, argument meaning and error behaviour).  The reference is compiled
// Go; no Go toolchain exists in this image, so the host layer a Go maintainer would write with cgo
//  is written here in C++ instead and exercised by tests/native/host_test.cpp.
//
// Nothing in this file computes anything about the path itself: every function forwards to the library (which runs
// CUDA kernels and has no CPU fallback) and only moves results into the reference's own types.
//
//   reference (h264/...)                                   here
//   NalUnit, (*NalUnit).RBSP            nalUnit.go:3-30,72   h264::NalUnit, NalUnit::RBSP()
//   NewNalUnit                          nalUnit.go:75-131    h264::NewNalUnit, NewNalUnits (batch)
//   readNalUnit loop over a stream      server.go:64-111     h264::ReadNalUnits
//   handleConnection / ByteStreamReader server.go:113-172    h264::ByteStreamReader (batches into pinned buffers)
//   PreCtxState, SliceQPy, Clip3        cabac.go:113-139     h264::PreCtxState, SliceQPy, Clip3
//   MNVars / CodedblockPatternMN, MN    mn_vars.go           h264::MNVars, h264::MN
//   ArithmeticDecoding.{DecodeBypass, DecodeTerminate, RenormD, BinaryDecision}  cabac.go:468-540
//   (*CABAC).StateTransitionProcess     cabac.go:544-553     h264::CABAC::StateTransitionProcess
//   NewSliceContext (header part)       slice.go:835-1048    h264::SliceHeaders
//   CtxIdx, NewBinarization, initCabac  cabac.go:557, :340, :148  h264::CtxIdx, NewBinarization, InitCabac
//   NewSPS, NewPPS                      sps.go:192, pps.go:40 h264::NewSPS, h264::NewPPS (+ h264::ParamSets for the walk)
//   (new, batch)                                             h264::InitContexts, h264::DecodeBins
#pragma once
#include <stdint.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <functional>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "../include/h264b200.h"

namespace h264 {

// Go panics (index out of range in BitReader.Read, bit_reader.go:298) and library failures both surface as exceptions.
struct Panic : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// One context per GPU (one per process-rank in the sharded setup).
class Device {
   public:
    explicit Device(int device = 0) {
        const int32_t rc = h264b_create(device, &ctx_);
        if (rc != H264B_OK) throw std::runtime_error("h264b_create: status " + std::to_string(rc) + " (no CPU fallback)");
    }
    ~Device() { h264b_destroy(ctx_); }
    Device(const Device &) = delete;
    Device &operator=(const Device &) = delete;
    static Device &Default() {
        static Device d(0);
        return d;
    }
    h264b_ctx *ctx() const { return ctx_; }
    void check(int32_t rc) const {
        if (rc != H264B_OK) throw std::runtime_error("h264b status " + std::to_string(rc) + ": " + h264b_last_error(ctx_));
    }

   private:
    h264b_ctx *ctx_ = nullptr;
};

// ------------------------------------------------------------------------------------------------ NAL units
struct NalUnit {  // nalUnit.go:3-30 (Go ints)
    int NumBytes = 0, ForbiddenZeroBit = 0, RefIdc = 0, Type = 0;
    int SvcExtensionFlag = 0, Avc3dExtensionFlag = 0, IdrFlag = 0, PriorityId = 0, NoInterLayerPredFlag = 0;
    int DependencyId = 0, QualityId = 0, TemporalId = 0, UseRefBasePicFlag = 0, DiscardableFlag = 0, OutputFlag = 0;
    int ReservedThree2Bits = 0, HeaderBytes = 0, NonIdrFlag = 0, ViewId = 0, AnchorPicFlag = 0, InterViewFlag = 0;
    int ReservedOneBit = 0, ViewIdx = 0, DepthFlag = 0;
    uint8_t EmulationPreventionThreeByte = 0;
    std::vector<uint8_t> rbsp;
    uint64_t startOffset = 0;  // server.go:88 (absolute stream offset of the NAL's first byte)
    const std::vector<uint8_t> &RBSP() const { return rbsp; }  // nalUnit.go:72
};

inline NalUnit make_nal_unit(const h264b_nal &c, const h264b_nal_ext *e, const uint8_t *rbsp_buf, uint64_t base) {
    NalUnit u;
    u.NumBytes = (int)c.num_bytes;
    u.ForbiddenZeroBit = c.forbidden_zero_bit;
    u.RefIdc = c.ref_idc;
    u.Type = c.type;
    u.HeaderBytes = c.header_bytes;
    if (e) {
        u.SvcExtensionFlag = e->svc_extension_flag;
        u.Avc3dExtensionFlag = e->avc_3d_extension_flag;
        u.IdrFlag = e->idr_flag;
        u.PriorityId = e->priority_id;
        u.NoInterLayerPredFlag = e->no_inter_layer_pred_flag;
        u.DependencyId = e->dependency_id;
        u.QualityId = e->quality_id;
        u.TemporalId = e->temporal_id;
        u.UseRefBasePicFlag = e->use_ref_base_pic_flag;
        u.DiscardableFlag = e->discardable_flag;
        u.OutputFlag = e->output_flag;
        u.ReservedThree2Bits = e->reserved_three_2bits;
        u.NonIdrFlag = e->non_idr_flag;
        u.ViewId = e->view_id;
        u.AnchorPicFlag = e->anchor_pic_flag;
        u.InterViewFlag = e->inter_view_flag;
        u.ReservedOneBit = e->reserved_one_bit;
        u.ViewIdx = e->view_idx;
        u.DepthFlag = e->depth_flag;
    }
    if (c.flags & H264B_F_HAS_EPB) u.EmulationPreventionThreeByte = 3;  // nalUnit.go:117
    u.rbsp.assign(rbsp_buf + c.rbsp_off, rbsp_buf + c.rbsp_off + c.rbsp_len);
    u.startOffset = base + c.start;
    return u;
}

// NewNalUnit(frame, numBytesInNal) for a batch of independent frames (one kernel launch for all of them)
inline std::vector<NalUnit> NewNalUnits(const std::vector<std::vector<uint8_t>> &frames, Device &dev = Device::Default()) {
    std::vector<uint8_t> cat;
    std::vector<uint64_t> off;
    std::vector<uint32_t> len;
    for (const auto &f : frames) {
        off.push_back(cat.size());
        len.push_back((uint32_t)f.size());
        cat.insert(cat.end(), f.begin(), f.end());
        while (cat.size() & 15) cat.push_back(0xFF);  // frames start on 16-byte boundaries
    }
    const size_t n = frames.size();
    std::vector<h264b_nal> nals(n);
    std::vector<h264b_nal_ext> ext(n);
    std::vector<uint8_t> rbsp(cat.size() + 16);
    dev.check(h264b_nal_units(dev.ctx(), cat.data(), cat.size(), off.data(), len.data(), (uint32_t)n, 0, nals.data(),
                              ext.data(), rbsp.data()));
    std::vector<NalUnit> out;
    for (size_t i = 0; i < n; i++) {
        if (nals[i].flags & H264B_F_OVERRUN) throw Panic("NewNalUnit: index out of range (bit_reader.go:298)");
        out.push_back(make_nal_unit(nals[i], &ext[i], rbsp.data(), 0));
    }
    return out;
}

// nalUnit.go:75
inline NalUnit NewNalUnit(const uint8_t *frame, int numBytesInNal, Device &dev = Device::Default()) {
    return NewNalUnits({std::vector<uint8_t>(frame, frame + numBytesInNal)}, dev)[0];
}

// One long stream as byte ranges that can be scanned independently, e.g. one per GPU (h264b_cut_byte_ranges; the same
// rule as h264decode_b200/sharding.py cut_byte_ranges): every nominal cut k * n / n_ranges moves forward to the next start code
// 00 00 00 01 (server.go:19, :28-39) and a range keeps the start code that opens the next one, because the reference's
// NAL unit is payload plus the following start code (server.go:64-111).  The NAL units of range r are exactly those of
// the whole stream whose start code lies in [begin_r, begin_{r+1}); ReadNalUnits(stream + begin, end - begin, dev,
// begin) yields them with whole-stream offsets.  Host work is O(n_ranges x NAL size): no pass over the stream.
inline std::vector<std::pair<size_t, size_t>> CutByteRanges(const uint8_t *stream, size_t n, unsigned n_ranges) {
    std::vector<uint64_t> b(n_ranges), e(n_ranges);
    if (h264b_cut_byte_ranges(stream, n, n_ranges, b.data(), e.data()) != H264B_OK)
        throw std::invalid_argument("CutByteRanges");
    std::vector<std::pair<size_t, size_t>> out;
    for (unsigned r = 0; r < n_ranges; r++) out.emplace_back((size_t)b[r], (size_t)e[r]);
    return out;
}

// Every NalUnit the readNalUnit loop (server.go:64-111) produces from a byte stream held in memory (`base`: offset of
// `stream` in a longer stream, added to startOffset).
inline std::vector<NalUnit> ReadNalUnits(const uint8_t *stream, size_t n, Device &dev = Device::Default(), uint64_t base = 0) {
    const h264b_nal *nals = nullptr;
    const h264b_nal_ext *ext = nullptr;
    const uint8_t *rbsp = nullptr;
    h264b_scan_summary sum;
    dev.check(h264b_annexb_scan(dev.ctx(), stream, n, 0, 1, &nals, &ext, &sum, &rbsp, nullptr));
    std::vector<NalUnit> out;
    out.reserve(sum.n_nals);
    for (uint64_t i = 0; i < sum.n_nals; i++) out.push_back(make_nal_unit(nals[i], &ext[i], rbsp, base));
    return out;
}

// ------------------------------------------------------------------------------------------------ ingest
using SPS = h264b_sps;  // (NewSPS / NewPPS further down)
using PPS = h264b_pps;
// What handleConnection does with each NAL unit (server.go:145-162), as callbacks: every unit, then by type NewSPS, NewPPS,
// and the header part of NewSliceContext with the parameter sets in force.  A record whose status is H264B_SH_PANIC is
// one the reference would have panicked on (its handleConnection recovers and exits, server.go:136-143).
struct IngestHandlers {
    std::function<void(const NalUnit &)> on_nal;
    std::function<void(const SPS &)> on_sps;
    std::function<void(const PPS &)> on_pps;
    std::function<void(const NalUnit &, const h264b_slice_header &)> on_slice;
};
// The reference's handleConnection (server.go:113-166) appends one byte at a time to a growing []byte
// (bit_reader.go:27-39) and re-tests isStartSequence after every byte.  Here the connection is read straight into
// pinned, device-bound buffers; whole batches go through h264b_stream_submit / h264b_stream_wait with two batches in
// flight (the socket read of batch k+1 overlaps the GPU pass over batch k).  The bytes after the last start code of
// a batch belong to a NAL that is not complete yet: they are carried over in front of the next batch, so the NAL
// units (and their absolute startOffset) are exactly those of one pass over the whole stream.
class ByteStreamReader {
   public:
    explicit ByteStreamReader(Device &dev = Device::Default(), size_t batch_bytes = 64u << 20, size_t carry_room = 8u << 20)
        : dev_(dev), batch_(batch_bytes), room_(carry_room) {
        for (int i = 0; i < 2; i++) alloc(i, room_ + batch_);
    }
    ~ByteStreamReader() {
        for (int i = 0; i < 2; i++) h264b_host_free(dev_.ctx(), buf_[i]);
    }
    // Reads fd until end of file; on_nal is called for every NAL unit in stream order.  Returns the number of units.
    uint64_t Run(int fd, const std::function<void(const NalUnit &)> &on_nal) {
        IngestHandlers h;
        h.on_nal = on_nal;
        return Run(fd, h);
    }
    // The same with handleConnection's dispatch: parameter sets and slice headers are parsed on the device in the same
    // job as the split (H264B_STREAM_PARAM_SETS); the sets in force carry over from batch to batch.  max_slices bounds
    // the slice NAL units of one batch.
    uint64_t Run(int fd, const IngestHandlers &hd, uint32_t max_slices = 1u << 16) {
        const bool dispatch = (bool)hd.on_sps || (bool)hd.on_pps || (bool)hd.on_slice;
        bool have_sps = false, have_pps = false;
        SPS cur_sps;
        PPS cur_pps;
        uint64_t n_units = 0, consumed = 0;  // consumed: stream offset of the first byte of the pending batch's new data
        int cur = 0;
        bool eof = false, in_flight = false;
        uint64_t ticket = 0, flight_base = 0;
        size_t flight_len = 0, flight_begin = 0;  // the in-flight batch is buf_[cur ^ 1][flight_begin .. flight_begin+flight_len)
        size_t carry = 0;                         // bytes in front of buf_[cur]'s new data (at [room_ - carry, room_))
        std::vector<uint8_t> tail;                // (carry bytes, kept outside the pinned buffers between batches)
        while (!eof || in_flight) {
            // 1. fill the free buffer from the connection while the other batch is on the GPU
            size_t fill = 0;
            if (!eof) {
                uint8_t *dst = buf_[cur] + room_;
                while (fill < batch_) {
                    const ssize_t r = ::read(fd, dst + fill, batch_ - fill);  // straight into pinned memory
                    if (r < 0) throw std::runtime_error("read failed");
                    if (r == 0) {
                        eof = true;
                        break;
                    }
                    fill += (size_t)r;
                }
            }
            // 2. collect the batch in flight: its NAL units, and what it leaves for the next one
            if (in_flight) {
                h264b_stream_result res;
                dev_.check(h264b_stream_wait(dev_.ctx(), ticket, &res));
                uint32_t i_sps = 0, i_pps = 0, i_slice = 0;
                for (uint64_t i = 0; i < res.scan.n_nals; i++) {
                    const NalUnit u = make_nal_unit(res.nals[i], res.ext ? &res.ext[i] : nullptr, res.rbsp, flight_base);
                    if (hd.on_nal) hd.on_nal(u);
                    n_units++;
                    if (!dispatch) continue;
                    if (u.Type == 7 && i_sps < res.n_sps) {
                        if (hd.on_sps) hd.on_sps(res.sps[i_sps]);
                        i_sps++;
                    } else if (u.Type == 8 && i_pps < res.n_pps) {
                        if (hd.on_pps) hd.on_pps(res.pps[i_pps]);
                        i_pps++;
                    } else if (u.Type == 1 || u.Type == 5) {
                        if (i_slice >= res.n_slices)
                            throw std::runtime_error("ByteStreamReader: more slice NAL units in a batch than max_slices");
                        if (hd.on_slice) hd.on_slice(u, res.headers[i_slice]);
                        i_slice++;
                    }
                }
                if (dispatch && res.n_sps) {  // the sets in force behind this batch: its last SPS and the last PPS behind it
                    cur_sps = res.sps[res.n_sps - 1];
                    have_sps = true;
                    have_pps = false;
                    for (uint32_t q = res.n_pps; q-- > 0;)
                        if (res.pps_nal[q] > res.sps_nal[res.n_sps - 1]) {
                            cur_pps = res.pps[q];
                            have_pps = true;
                            break;
                        }
                } else if (dispatch && res.n_pps && have_sps) {  // a PPS for the VideoStream the batch inherited
                    cur_pps = res.pps[res.n_pps - 1];
                    have_pps = true;
                }
                // everything from the last start code on is carried over (none found: a start code may still straddle
                // the batch boundary, keep the last 3 bytes)
                size_t keep_from;
                if (res.scan.n_start_codes == 0)
                    keep_from = flight_len > 3 ? flight_len - 3 : 0;
                else if (res.scan.n_nals == 0)
                    keep_from = (size_t)res.scan.first_start - 4;
                else
                    keep_from = (size_t)(res.nals[res.scan.n_nals - 1].start + res.nals[res.scan.n_nals - 1].num_bytes) - 4;
                const uint8_t *src = buf_[cur ^ 1] + flight_begin;
                tail.assign(src + keep_from, src + flight_len);
                carry_base_ = flight_base + keep_from;
                in_flight = false;
            }
            carry = tail.size();
            if (carry > room_) {  // a NAL larger than the reserved room: grow both buffers
                grow(carry);
            }
            if (carry) memcpy(buf_[cur] + room_ - carry, tail.data(), carry);
            // 3. submit carry + new bytes
            if (fill == 0 && eof) break;  // nothing new: what is left is the (dropped) tail after the last start code
            h264b_stream_job job;
            memset(&job, 0, sizeof(job));
            job.stream = buf_[cur] + room_ - carry;
            job.n = carry + fill;
            job.flags = H264B_STREAM_WANT_RBSP;
            if (dispatch) {  // (no CABAC ops: headers only)
                job.flags |= H264B_STREAM_PARAM_SETS;
                job.max_slices = max_slices;
                job.n_ctx = 1;
                job.initial_sps = have_sps ? &cur_sps : nullptr;
                job.initial_pps = have_sps && have_pps ? &cur_pps : nullptr;
            }
            dev_.check(h264b_stream_submit(dev_.ctx(), &job, &ticket));
            in_flight = true;
            flight_begin = room_ - carry;
            flight_len = carry + fill;
            flight_base = carry ? carry_base_ : consumed;
            consumed += fill;
            tail.clear();
            cur ^= 1;
        }
        return n_units;
    }

   private:
    void alloc(int i, size_t bytes) {
        void *p = nullptr;
        dev_.check(h264b_host_alloc(dev_.ctx(), bytes, &p));
        buf_[i] = (uint8_t *)p;
    }
    void grow(size_t need) {
        const size_t new_room = need * 2;
        for (int i = 0; i < 2; i++) {
            void *p = nullptr;
            dev_.check(h264b_host_alloc(dev_.ctx(), new_room + batch_, &p));
            memcpy((uint8_t *)p + new_room, buf_[i] + room_, batch_);
            h264b_host_free(dev_.ctx(), buf_[i]);
            buf_[i] = (uint8_t *)p;
        }
        room_ = new_room;
    }
    Device &dev_;
    size_t batch_, room_;
    uint8_t *buf_[2] = {nullptr, nullptr};
    uint64_t carry_base_ = 0;
};

// ------------------------------------------------------------------------------------------------ context init
inline int Clip3(int x, int y, int z) { return z < x ? x : (z > y ? y : z); }                            // cabac.go:131-139
inline int SliceQPy(int picInitQpMinus26, int sliceQpDelta) { return 26 + picInitQpMinus26 + sliceQpDelta; }  // :113-115
inline int PreCtxState(int m, int n, int sliceQPY, Device &dev = Device::Default()) {                    // :118-121
    int32_t v = 0;
    dev.check(h264b_pre_ctx_state(dev.ctx(), m, n, sliceQPY, &v));
    return v;
}
struct MN {  // mn_vars.go:3-5
    int M, N;
};
const int NoCabacInitIdc = -1;  // mn_vars.go:7
inline MN MNVars(int ctxIdx, int cabacInitIdc, Device &dev = Device::Default()) {  // mn_vars.go:15-175, 184-440
    int32_t m = 0, n = 0;
    dev.check(h264b_mn(dev.ctx(), ctxIdx, cabacInitIdc, 0, &m, &n));
    return MN{m, n};
}
// new, batch: pStateIdx | valMPS << 6 for every (slice, ctxIdx)
inline std::vector<uint8_t> InitContexts(const std::vector<int> &sliceQPY, const std::vector<int> &cabacInitIdc, int nCtx,
                                         Device &dev = Device::Default()) {
    std::vector<h264b_slice_qp> p(sliceQPY.size());
    for (size_t i = 0; i < p.size(); i++) p[i] = h264b_slice_qp{sliceQPY[i], cabacInitIdc[i]};
    std::vector<uint8_t> out(p.size() * (size_t)nCtx);
    dev.check(h264b_ctx_init(dev.ctx(), p.data(), (uint32_t)p.size(), (uint32_t)nCtx, out.data(), 0));
    return out;
}

// ------------------------------------------------------------------------------------------------ CABAC engine
// MSB-first bit cursor over a byte slice (bit_reader.go:11, 232-236, 292-325); reading past the end panics.
struct BitReader {
    const uint8_t *bytes = nullptr;
    size_t n = 0;
    uint64_t bitsRead = 0;
    // the next `count` bits, MSB first, packed for h264b_engine_step; bits past the end are not there (n_avail)
    void peek(uint8_t out[32], uint32_t *n_avail, uint32_t count = 256) const {
        memset(out, 0, 32);
        uint32_t k = 0;
        for (; k < count && bitsRead + k < 8 * (uint64_t)n; k++) {
            const uint64_t p = bitsRead + k;
            const uint32_t b = (bytes[p >> 3] >> (7 - (p & 7))) & 1u;
            out[k >> 3] |= (uint8_t)(b << (7 - (k & 7)));
        }
        *n_avail = k;
    }
    void advance(uint32_t used, uint32_t avail) {
        if (used > avail) throw Panic("BitReader: index out of range (bit_reader.go:298)");
        bitsRead += used;
    }
};

struct CABAC {  // cabac.go:141-145
    int PStateIdx = 0, ValMPS = 0;
    void StateTransitionProcess(int binVal, Device &dev = Device::Default()) {  // cabac.go:544-553
        int32_t p = PStateIdx, v = ValMPS;
        dev.check(h264b_state_transition(dev.ctx(), 0, &p, &v, binVal));
        PStateIdx = p;
        ValMPS = v;
    }
};

struct ArithmeticDecoding {  // cabac.go:513-517; the methods keep the reference's argument and result order
    BitReader *bits = nullptr;
    Device *dev = &Device::Default();
    uint32_t flags = 0;  // REF behaviour (H264B_BYPASS_SPEC_OR / H264B_TABLES_SPEC select the corrected variants)

    std::pair<int64_t, int> DecodeBypass(int64_t codIRange, int64_t codIOffset) {  // :468-481 -> (codIOffset, binVal)
        int32_t bin = 0;
        step(H264B_OP_BYPASS, &codIRange, &codIOffset, nullptr, nullptr, &bin);
        return {codIOffset, bin};
    }
    std::tuple<int64_t, int64_t, int> DecodeTerminate(int64_t codIRange, int64_t codIOffset) {  // :486-499
        int32_t bin = 0;
        step(H264B_OP_TERMINATE, &codIRange, &codIOffset, nullptr, nullptr, &bin);
        return {codIRange, codIOffset, bin};
    }
    std::pair<int64_t, int64_t> RenormD(int64_t codIRange, int64_t codIOffset) {  // :503-511
        int32_t bin = 0;
        step(4u, &codIRange, &codIOffset, nullptr, nullptr, &bin);
        return {codIRange, codIOffset};
    }
    // the arithmetic core of BinaryDecision (:525-536) for an explicit context state: no transition, no renorm (A6)
    std::tuple<int, int64_t, int64_t> BinaryDecision(const CABAC &c, int64_t codIRange, int64_t codIOffset) {
        int32_t bin = 0;
        dev->check(h264b_binary_decision(dev->ctx(), flags, c.PStateIdx, c.ValMPS, &codIRange, &codIOffset, &bin));
        return {bin, codIRange, codIOffset};
    }
    // composed DecodeDecision: core + StateTransitionProcess + RenormD
    int DecodeDecision(CABAC &c, int64_t &codIRange, int64_t &codIOffset) {
        int32_t bin = 0, p = c.PStateIdx, v = c.ValMPS;
        step(H264B_OP_DECISION, &codIRange, &codIOffset, &p, &v, &bin);
        c.PStateIdx = p;
        c.ValMPS = v;
        return bin;
    }
    std::pair<int64_t, int64_t> InitDecodingEngine() {  // initDecodingEngine, :439-446 -> (codIRange, codIOffset)
        int64_t r = 0, o = 0;
        int32_t bin = 0;
        step(5u, &r, &o, nullptr, nullptr, &bin);
        return {r, o};
    }

   private:
    void step(uint32_t kind, int64_t *r, int64_t *o, int32_t *p, int32_t *v, int32_t *bin) {
        uint8_t buf[32];
        uint32_t avail = 0, used = 0;
        bits->peek(buf, &avail);
        dev->check(h264b_engine_step(dev->ctx(), kind, flags, buf, avail, r, o, p, v, bin, &used));
        bits->advance(used, avail);
    }
};

// ------------------------------------------------------------------------------------------------ slice headers
// The header part of NewSliceContext (slice.go:835-1048) for a batch of slice NAL units: header fields, SliceQPy,
// and where slice_data() starts.  A header on which the reference would panic throws h264::Panic.
inline std::vector<h264b_slice_header> SliceHeaders(const h264b_param_sets &ps, const std::vector<NalUnit> &slices,
                                                    Device &dev = Device::Default()) {
    std::vector<uint8_t> cat, type, ref;
    std::vector<uint64_t> off;
    std::vector<uint32_t> len;
    for (const auto &u : slices) {
        off.push_back(cat.size());
        len.push_back((uint32_t)u.rbsp.size());
        type.push_back((uint8_t)u.Type);
        ref.push_back((uint8_t)u.RefIdc);
        cat.insert(cat.end(), u.rbsp.begin(), u.rbsp.end());
    }
    cat.resize(cat.size() + 8);
    std::vector<h264b_slice_header> out(slices.size());
    dev.check(h264b_slice_headers(dev.ctx(), &ps, cat.data(), cat.size(), off.data(), len.data(), type.data(), ref.data(),
                                  (uint32_t)slices.size(), out.data()));
    for (const auto &h : out)
        if (h.status == H264B_SH_PANIC) throw Panic("NewSliceContext: the reference panics on this slice header");
    return out;
}

// ------------------------------------------------------------------------------------------------ parameter sets
// NewSPS(rbsp, showPacket) (sps.go:192) / NewPPS(sps, rbsp, showPacket) (pps.go:40): field extraction on the device, one
// thread per parameter set.  A parameter set on which the reference would panic throws h264::Panic (the reference's
// handleConnection recovers it and exits, server.go:136-143).  showPacket only controlled debug logging.
inline SPS NewSPS(const std::vector<uint8_t> &rbsp, bool showPacket = false, Device &dev = Device::Default()) {
    (void)showPacket;
    std::vector<uint8_t> buf(rbsp);
    buf.resize(buf.size() + 8);
    const uint64_t off = 0;
    const uint32_t len = (uint32_t)rbsp.size();
    SPS out;
    dev.check(h264b_parse_sps(dev.ctx(), buf.data(), buf.size(), &off, &len, 1, &out));
    if (out.status != H264B_SH_OK) throw Panic("NewSPS: the reference panics on this parameter set");
    return out;
}
inline PPS NewPPS(const SPS *sps, const std::vector<uint8_t> &rbsp, bool showPacket = false, Device &dev = Device::Default()) {
    (void)sps;  // only read by the reference after it has panicked (pps.go:99-103)
    (void)showPacket;
    std::vector<uint8_t> buf(rbsp);
    buf.resize(buf.size() + 8);
    const uint64_t off = 0;
    const uint32_t len = (uint32_t)rbsp.size();
    PPS out;
    dev.check(h264b_parse_pps(dev.ctx(), buf.data(), buf.size(), &off, &len, 1, &out));
    if (out.status != H264B_SH_OK) throw Panic("NewPPS: the reference panics on this parameter set");
    return out;
}
// VideoStream{SPS, PPS} (server.go:149-158) as the slice-header walk needs it
inline h264b_param_sets ParamSets(const SPS &sps, const PPS &pps) {
    h264b_param_sets ps;
    if (h264b_make_param_sets(&sps, &pps, &ps) != H264B_OK) throw std::runtime_error("h264b_make_param_sets");
    return ps;
}

// ------------------------------------------------------------------------------------------------ syntax-element glue
constexpr int NaCtxId = H264B_NA_CTX_ID;  // cabac.go:4
// CtxIdx(binIdx, maxBinIdxCtx, ctxIdxOffset) (cabac.go:557): Table 9-39 as the reference has it
inline int64_t CtxIdx(int64_t binIdx, int64_t maxBinIdxCtx, int64_t ctxIdxOffset, Device &dev = Device::Default()) {
    int64_t out = 0;
    dev.check(h264b_ctx_idx(dev.ctx(), 1, &binIdx, &maxBinIdxCtx, &ctxIdxOffset, &out));
    return out;
}
// NewBinarization(syntaxElement, data) (cabac.go:340): the name as an H264B_SE_* value, data.SliceTypeName as H264B_ST_*
using Binarization = h264b_binarization;
inline Binarization NewBinarization(int32_t syntaxElement, int32_t sliceTypeName, Device &dev = Device::Default()) {
    Binarization b;
    dev.check(h264b_new_binarization(dev.ctx(), 1, &syntaxElement, &sliceTypeName, &b));
    return b;
}
// initCabac(binarization, context) (cabac.go:148): the binarization's private binIdx is always 0 in the reference
inline CABAC InitCabac(const Binarization &b, int64_t picInitQpMinus26, int64_t sliceQpDelta, int64_t binIdx = 0,
                       uint32_t flags = 0, Device &dev = Device::Default()) {
    const int64_t maxp = b.max_prefix, offp = b.off_prefix;
    int32_t p = 0, v = 0;
    dev.check(h264b_init_cabac(dev.ctx(), flags, 1, &binIdx, &maxp, &offp, &picInitQpMinus26, &sliceQpDelta, &p, &v, nullptr));
    CABAC c;
    c.PStateIdx = p;
    c.ValMPS = v;
    return c;
}

// new, batch: the whole engine for many slices at once (see h264b_cabac_job)
inline void DecodeBins(const h264b_cabac_job &job, Device &dev = Device::Default()) {
    dev.check(h264b_cabac_decode(dev.ctx(), &job));
}

// new, batch: mb_type as a syntax element for many slices at once -- the walk of slice.go:639-672 (NewBinarization,
// CtxIdx per bin, IsBinStringMatch) with every slice on its own, data-dependent sequence of contexts (see
// h264b_mb_type_job; host pointers)
inline void DecodeMbTypes(const h264b_mb_type_job &job, Device &dev = Device::Default()) {
    dev.check(h264b_mb_type_decode(dev.ctx(), &job));
}

// The reference serves every TCP connection from a goroutine of its own (main.go:16-21 -> ByteStreamReader ->
// handleConnection, h264/server.go:113-166).  Scheduler is that concurrency for whole streams held in host memory: a
// batch of independent streams goes over the GPUs of this process (longest processing time first by bytes; a GPU takes its
// share in up to three passes, longest slices first, each one split + strip pass and the CABAC engine in up to six
// launches by slice length, side by side), and comes back as NAL units,
// bins and final engine states per stream, with the time every slice's result reached host memory.
// What Scheduler::Run would decide for a batch over n_devices devices of sm_count SMs each: per stream its device (-1:
// no NAL unit in it) and its pass on that device, per slice its launch class (0: on a scheduler of its own, 1..5 by
// length, 255: not decoded).  Host only: no device is touched.
struct BatchPlan {
    std::vector<int32_t> stream_device;
    std::vector<uint32_t> stream_pass;
    std::vector<uint8_t> slice_class;
};
inline BatchPlan PlanBatch(const h264b_batch_job &job, uint32_t n_devices, uint32_t sm_count = 148) {
    BatchPlan p;
    p.stream_device.assign(job.n_streams, -1);
    p.stream_pass.assign(job.n_streams, 0);
    p.slice_class.assign(job.total_slices, 255);
    const int32_t rc = h264b_scheduler_plan(&job, n_devices, sm_count, p.stream_device.data(), p.stream_pass.data(),
                                            p.slice_class.data());
    if (rc != H264B_OK) throw std::runtime_error("h264b_scheduler_plan: status " + std::to_string(rc));
    return p;
}

class Scheduler {
   public:
    // devices: CUDA ordinals; empty = every device of the process
    explicit Scheduler(std::vector<int32_t> devices = {}) {
        if (devices.empty()) {
            int32_t n = 0;
            h264b_device_count(&n);
            for (int32_t d = 0; d < n; d++) devices.push_back(d);
        }
        const int32_t rc = h264b_scheduler_create(devices.data(), (uint32_t)devices.size(), &s_);
        if (rc != H264B_OK) throw std::runtime_error("h264b_scheduler_create: status " + std::to_string(rc) + " (no CPU fallback)");
    }
    ~Scheduler() { h264b_scheduler_destroy(s_); }
    Scheduler(const Scheduler &) = delete;
    Scheduler &operator=(const Scheduler &) = delete;
    // synchronous; the result's arrays stay valid until the next Run or the scheduler's destruction
    h264b_batch_result Run(const h264b_batch_job &job) {
        h264b_batch_result res;
        const int32_t rc = h264b_scheduler_run(s_, &job, &res);
        if (rc != H264B_OK)
            throw std::runtime_error("h264b_scheduler_run: status " + std::to_string(rc) + ": " + h264b_scheduler_last_error(s_));
        return res;
    }

   private:
    h264b_scheduler *s_ = nullptr;
};

}  // namespace h264
