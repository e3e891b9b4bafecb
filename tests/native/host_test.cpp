// host_test.cpp -- drives host/h264.hpp (the C++ mirror of the reference's Go API over the C ABI) and prints what it
// gets as text; tests/test_host_cpp.py (GPU) builds it with g++, runs it and compares with the oracle.
// Test infrastructure: not part of the product.
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>

#include <fstream>
#include <thread>

#include "../../host/h264.hpp"

static std::vector<uint8_t> slurp(const char *path) {
    std::ifstream f(path, std::ios::binary);
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
static uint64_t fnv(const std::vector<uint8_t> &v) {
    uint64_t h = 1469598103934665603ull;
    for (uint8_t b : v) h = (h ^ b) * 1099511628211ull;
    return h;
}
static void print_nal(const h264::NalUnit &u) {
    printf("nal %llu %d %d %d %d %d %zu %016llx %d %d %d %d %d %d %d\n", (unsigned long long)u.startOffset, u.NumBytes,
           u.ForbiddenZeroBit, u.RefIdc, u.Type, u.HeaderBytes, u.RBSP().size(), (unsigned long long)fnv(u.RBSP()),
           (int)u.EmulationPreventionThreeByte, u.SvcExtensionFlag, u.Avc3dExtensionFlag, u.PriorityId, u.ViewId,
           u.TemporalId, u.ViewIdx);
}

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    const std::string mode = argv[1];
    try {
        if (mode == "nals") {  // readNalUnit loop over a whole buffer
            const auto s = slurp(argv[2]);
            for (const auto &u : h264::ReadNalUnits(s.data(), s.size())) print_nal(u);
        } else if (mode == "ranges") {  // CutByteRanges alone (no device call)
            const auto s = slurp(argv[2]);
            for (const auto &r : h264::CutByteRanges(s.data(), s.size(), (unsigned)atoi(argv[3])))
                printf("range %zu %zu\n", r.first, r.second);
        } else if (mode == "nals_ranges") {  // the stream as n byte ranges, each scanned on its own
            const auto s = slurp(argv[2]);
            for (const auto &r : h264::CutByteRanges(s.data(), s.size(), (unsigned)atoi(argv[3])))
                for (const auto &u : h264::ReadNalUnits(s.data() + r.first, r.second - r.first, h264::Device::Default(), r.first))
                    print_nal(u);
        } else if (mode == "ingest") {  // the same stream through a pipe in odd-sized writes, batched ingest
            const auto s = slurp(argv[2]);
            const size_t batch = strtoull(argv[3], nullptr, 10), chunk = strtoull(argv[4], nullptr, 10),
                         room = strtoull(argv[5], nullptr, 10);
            int fds[2];
            if (pipe(fds)) return 3;
            std::thread writer([&] {
                size_t off = 0, k = 0;
                while (off < s.size()) {
                    size_t n = chunk + (k++ % 7) * 13;  // ragged writes
                    if (n > s.size() - off) n = s.size() - off;
                    const ssize_t w = write(fds[1], s.data() + off, n);
                    if (w <= 0) break;
                    off += (size_t)w;
                }
                close(fds[1]);
            });
            h264::ByteStreamReader reader(h264::Device::Default(), batch, room);
            const uint64_t n = reader.Run(fds[0], [](const h264::NalUnit &u) { print_nal(u); });
            writer.join();
            close(fds[0]);
            printf("units %llu\n", (unsigned long long)n);
        } else if (mode == "frame") {  // NewNalUnit(frame, len(frame)); a Go panic prints "panic"
            const auto f = slurp(argv[2]);
            try {
                print_nal(h264::NewNalUnit(f.data(), (int)f.size()));
            } catch (const h264::Panic &) {
                printf("panic\n");
            }
        } else if (mode == "psets") {  // handleConnection's dispatch (server.go:145-162): NewSPS / NewPPS / slice headers
            const auto s = slurp(argv[2]);
            bool have_sps = false, have_pps = false;
            h264::SPS sps;
            h264::PPS pps;
            for (const auto &u : h264::ReadNalUnits(s.data(), s.size())) {
                try {
                    if (u.Type == 7) {
                        sps = h264::NewSPS(u.RBSP());
                        have_sps = true;
                        have_pps = false;
                        printf("sps %lld %lld %lld %lld %lld %lld %llu\n", (long long)sps.profile, (long long)sps.level,
                               (long long)sps.pic_width_in_mbs_minus1, (long long)sps.pic_height_in_map_units_minus1,
                               (long long)sps.pic_order_count_type, (long long)sps.n_hrd, (unsigned long long)sps.bits_read);
                    } else if (u.Type == 8) {
                        pps = h264::NewPPS(have_sps ? &sps : nullptr, u.RBSP());
                        have_pps = have_sps;
                        printf("pps %lld %lld %lld %lld %lld %llu\n", (long long)pps.id, (long long)pps.entropy_coding_mode,
                               (long long)pps.pic_init_qp_minus26, (long long)pps.chroma_qp_index_offset,
                               (long long)pps.transform_8x8_mode, (unsigned long long)pps.bits_read);
                    } else if ((u.Type == 1 || u.Type == 5) && have_sps && have_pps) {
                        const auto h = h264::SliceHeaders(h264::ParamSets(sps, pps), {u});
                        printf("slice %lld %lld %lld %llu\n", (long long)h[0].slice_type, (long long)h[0].slice_qp_y,
                               (long long)h[0].cabac_init_idc, (unsigned long long)h[0].header_bits);
                    }
                } catch (const h264::Panic &) {  // (the reference exits here; going on, the failed set is not usable)
                    printf("panic %d\n", u.Type);
                    if (u.Type == 7) have_sps = have_pps = false;
                    if (u.Type == 8) have_pps = false;
                }
            }
        } else if (mode == "ingest_psets") {  // handleConnection's dispatch inside the batched ingest (sets carry over)
            const auto s = slurp(argv[2]);
            const size_t batch = strtoull(argv[3], nullptr, 10), chunk = strtoull(argv[4], nullptr, 10);
            int fds[2];
            if (pipe(fds)) return 3;
            std::thread writer([&] {
                size_t off = 0;
                while (off < s.size()) {
                    size_t n = chunk < s.size() - off ? chunk : s.size() - off;
                    const ssize_t w = write(fds[1], s.data() + off, n);
                    if (w <= 0) break;
                    off += (size_t)w;
                }
                close(fds[1]);
            });
            h264::IngestHandlers hd;
            hd.on_sps = [](const h264::SPS &sps) {
                if (sps.status != H264B_SH_OK) {
                    printf("panic 7\n");
                    return;
                }
                printf("sps %lld %lld %lld %lld %lld %lld %llu\n", (long long)sps.profile, (long long)sps.level,
                       (long long)sps.pic_width_in_mbs_minus1, (long long)sps.pic_height_in_map_units_minus1,
                       (long long)sps.pic_order_count_type, (long long)sps.n_hrd, (unsigned long long)sps.bits_read);
            };
            hd.on_pps = [](const h264::PPS &pps) {
                if (pps.status != H264B_SH_OK) {
                    printf("panic 8\n");
                    return;
                }
                printf("pps %lld %lld %lld %lld %lld %llu\n", (long long)pps.id, (long long)pps.entropy_coding_mode,
                       (long long)pps.pic_init_qp_minus26, (long long)pps.chroma_qp_index_offset,
                       (long long)pps.transform_8x8_mode, (unsigned long long)pps.bits_read);
            };
            hd.on_slice = [](const h264::NalUnit &u, const h264b_slice_header &h) {
                if (h.status != H264B_SH_OK) {
                    printf("panic %d\n", u.Type);
                    return;
                }
                printf("slice %lld %lld %lld %llu\n", (long long)h.slice_type, (long long)h.slice_qp_y,
                       (long long)h.cabac_init_idc, (unsigned long long)h.header_bits);
            };
            h264::ByteStreamReader reader(h264::Device::Default(), batch, 4096);
            const uint64_t n = reader.Run(fds[0], hd, 64);
            writer.join();
            close(fds[0]);
            printf("units %llu\n", (unsigned long long)n);
        } else if (mode == "glue") {  // CtxIdx / NewBinarization / InitCabac, one call each like the Go functions
            for (int off : {3, 17, 21, 69, 276})
                for (int b : {-1, 0, 1, 2, 5, 9}) printf("ctxidx %d %d %lld\n", b, off, (long long)h264::CtxIdx(b, 6, off));
            for (int se = 0; se < 15; se++)
                for (int st : {H264B_ST_P, H264B_ST_I, H264B_ST_SI, H264B_ST_B}) {
                    const auto bz = h264::NewBinarization(se, st);
                    const auto c = h264::InitCabac(bz, -3, 5);
                    printf("bin %d %d %d %d %d %d %d %d\n", se, st, bz.prefix_suffix, bz.max_prefix, bz.off_prefix,
                           bz.use_decode_bypass, c.PStateIdx, c.ValMPS);
                }
            const auto c = h264::InitCabac(h264::NewBinarization(H264B_SE_MB_TYPE, H264B_ST_P), 4, -9, 2);
            printf("init %d %d\n", c.PStateIdx, c.ValMPS);
        } else if (mode == "ctx") {  // PreCtxState / MNVars / InitContexts
            for (int i = 2; i + 2 < argc; i += 3)
                printf("pre %d\n", h264::PreCtxState(atoi(argv[i]), atoi(argv[i + 1]), atoi(argv[i + 2])));
            for (int c : {0, 5, 10, 11, 39, 40, 70, 104, 105, 1023})
                for (int idc : {-1, 0, 1, 2, 3}) {
                    const h264::MN mn = h264::MNVars(c, idc);
                    printf("mn %d %d %d %d\n", c, idc, mn.M, mn.N);
                }
            const auto st = h264::InitContexts({26, 0, 51, 30}, {0, -1, 2, 1}, 128);
            printf("init %016llx %d %d\n", (unsigned long long)fnv(st), h264::SliceQPy(-3, 5), h264::Clip3(0, 51, 77));
        } else if (mode == "engine") {  // the per-call engine methods on a bit string given as a byte file
            const auto bytes = slurp(argv[2]);
            h264::BitReader br;
            br.bytes = bytes.data();
            br.n = bytes.size();
            h264::ArithmeticDecoding ad;
            ad.bits = &br;
            auto ro = ad.InitDecodingEngine();
            int64_t R = ro.first, O = ro.second;
            printf("init %lld %lld %llu\n", (long long)R, (long long)O, (unsigned long long)br.bitsRead);
            h264::CABAC c;
            c.PStateIdx = 20;
            c.ValMPS = 1;
            for (int i = 0; i < 24; i++) {
                int bin;
                if (i % 5 == 3) {
                    auto r = ad.DecodeBypass(R, O);  // REF form (A5)
                    O = r.first;
                    bin = r.second;
                } else if (i % 11 == 10) {
                    auto r = ad.DecodeTerminate(R, O);
                    R = std::get<0>(r);
                    O = std::get<1>(r);
                    bin = std::get<2>(r);
                } else {
                    bin = ad.DecodeDecision(c, R, O);
                }
                printf("step %d %d %lld %lld %d %d %llu\n", i, bin, (long long)R, (long long)O, c.PStateIdx, c.ValMPS,
                       (unsigned long long)br.bitsRead);
            }
            auto bd = ad.BinaryDecision(c, R, O);  // bare core: no transition, no renorm
            printf("core %d %lld %lld\n", std::get<0>(bd), (long long)std::get<1>(bd), (long long)std::get<2>(bd));
            c.StateTransitionProcess(1 - c.ValMPS);
            printf("trans %d %d\n", c.PStateIdx, c.ValMPS);
        } else if (mode == "plan") {  // h264::PlanBatch on a batch file (see "sched"), argv[3] devices (no device call)
            const auto f = slurp(argv[2]);
            const uint8_t *p = f.data();
            auto take = [&](void *dst, size_t n) {
                memcpy(dst, p, n);
                p += n;
            };
            uint32_t hdr[5];
            take(hdr, sizeof(hdr));
            std::vector<std::vector<uint8_t>> bytes(hdr[0]);
            std::vector<h264b_batch_stream> streams(hdr[0]);
            uint32_t total = 0;
            for (uint32_t i = 0; i < hdr[0]; i++) {
                uint64_t n;
                uint32_t ns;
                take(&n, 8);
                take(&ns, 4);
                bytes[i].resize(n);
                take(bytes[i].data(), n);
                streams[i].stream = bytes[i].data();
                streams[i].n = n;
                streams[i].first_slice = total;
                streams[i].n_slices = ns;
                total += ns;
            }
            p += (size_t)hdr[2] * 2;  // (the op schedule: not needed to plan)
            std::vector<uint32_t> n_ops(total);
            take(n_ops.data(), n_ops.size() * 4);
            h264b_batch_job job;
            memset(&job, 0, sizeof(job));
            job.streams = streams.data();
            job.n_streams = hdr[0];
            job.total_slices = total;
            job.n_ctx = hdr[1];
            job.n_ops_max = hdr[2];
            job.n_ops = n_ops.data();
            job.group_bytes = strtoull(argv[4], nullptr, 10);
            const h264::BatchPlan plan = h264::PlanBatch(job, (uint32_t)atoi(argv[3]));
            for (uint32_t i = 0; i < hdr[0]; i++) printf("stream %u device %d pass %u\n", i, plan.stream_device[i], plan.stream_pass[i]);
            for (uint32_t k = 0; k < total; k++) printf("slice %u class %u\n", k, (unsigned)plan.slice_class[k]);
        } else if (mode == "sched") {  // a batch of streams through h264::Scheduler; argv[2]: the batch as one file (see test)
            const auto f = slurp(argv[2]);
            const uint8_t *p = f.data();
            auto take = [&](void *dst, size_t n) {
                memcpy(dst, p, n);
                p += n;
            };
            uint32_t hdr[5];  // n_streams, n_ctx, n_ops_max, flags, workers
            take(hdr, sizeof(hdr));
            std::vector<std::vector<uint8_t>> bytes(hdr[0]);
            std::vector<h264b_batch_stream> streams(hdr[0]);
            uint32_t total = 0;
            for (uint32_t i = 0; i < hdr[0]; i++) {
                uint64_t n;
                uint32_t ns;
                take(&n, 8);
                take(&ns, 4);
                bytes[i].resize(n);
                take(bytes[i].data(), n);
                streams[i].stream = bytes[i].data();
                streams[i].n = n;
                streams[i].first_slice = total;
                streams[i].n_slices = ns;
                total += ns;
            }
            std::vector<uint16_t> ops(hdr[2]);
            take(ops.data(), ops.size() * 2);
            std::vector<uint32_t> n_ops(total);
            take(n_ops.data(), n_ops.size() * 4);
            std::vector<h264b_slice_qp> qp(total);
            take(qp.data(), qp.size() * sizeof(h264b_slice_qp));
            h264b_batch_job job;
            memset(&job, 0, sizeof(job));
            job.streams = streams.data();
            job.n_streams = hdr[0];
            job.total_slices = total;
            job.n_ctx = hdr[1];
            job.n_ops_max = hdr[2];
            job.ops = ops.data();
            job.n_ops = n_ops.data();
            job.qp = qp.data();
            job.flags = hdr[3];
            h264::Scheduler sched(std::vector<int32_t>(hdr[4], 0));  // (workers on device 0: the test box has one GPU)
            const h264b_batch_result r = sched.Run(job);
            for (uint32_t i = 0; i < hdr[0]; i++)
                printf("stream %u worker %d nals %llu\n", i, r.stream_device[i],
                       (unsigned long long)(r.stream_nal_off[i + 1] - r.stream_nal_off[i]));
            for (uint32_t k = 0; k < total; k++) {
                uint64_t sum = 0;  // the slice's whole bin words, position-weighted
                const uint32_t words = r.final[k].n_bins / 32;
                for (uint32_t w = 0; w < words; w++) sum = sum * 1000003ull + r.bins[r.bins_off[k] + w];
                printf("slice %u bins %u range %lld offset %lld flags %u sum %llu done %d\n", k, r.final[k].n_bins,
                       (long long)r.final[k].cod_i_range, (long long)r.final[k].cod_i_offset, r.final[k].flags,
                       (unsigned long long)sum, r.slice_done_ms[k] > 0 && r.slice_done_ms[k] <= r.makespan_ms);
            }
            printf("total bins %llu nals %llu\n", (unsigned long long)r.total_bins, (unsigned long long)r.total_nals);
        } else {
            return 2;
        }
    } catch (const std::exception &e) {
        printf("error %s\n", e.what());
        return 1;
    }
    return 0;
}
