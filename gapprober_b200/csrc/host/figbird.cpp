// figbird.cpp — drop-in replacement for the reference's figbird.cpp (the binary GapProber.py:41-44 builds
// with a bare `g++ ../figbird.cpp -o figbird.out` and runs as `./figbird.out <scaffolds.fa> ' 0' 250`).
//
// Same argv, same input files in the working directory (stat.txt, myout.sam, gapInfo.txt, Gaps/gaps_<n>.sam),
// same output files (gapout.txt, filledContigs.fa, mappedInserts.txt, out.txt) and the same stdout lines
// (figbird.cpp:2744, 2503, 2382, 2852, 2863, 2877, 2947).  The preamble runs on the host (fb_preamble.hpp);
// GapFiller::fillGap for ALL gaps runs on the GPU through the C ABI of libfigbird_cuda.so, loaded with
// dlopen so this file still compiles with no flags (glibc >= 2.34 needs no -ldl / -lpthread).
//
// Library lookup: $FIGBIRD_CUDA_LIB, then next to the executable, then its parent directory, then the
// dynamic loader's search path.  Optional environment:
//   FIGBIRD_DEVICE=<n>       CUDA device (default 0)
//   FIGBIRD_GPUS=<n>         fill the gaps on n GPUs (devices FIGBIRD_DEVICE .. +n-1): one context and one host thread per
//                            GPU, gaps dealt by fb_set_partition, no exchange between them
//   FIGBIRD_MODEL_OUT=<path> also write the learned model ("FBM1" file)
//   FIGBIRD_PREAMBLE_ONLY=1  stop after the preamble (used by the benchmark to build its model)
//   FIGBIRD_FROM_SAM=<path>  in-memory hand-off: split <path> (bowtie2's result.sam) in this process the way
//                            main.cpp does (fb_split.hpp) and feed the proper pairs to the preamble and the gap records to
//                            the GPU without writing myout.sam or Gaps/*.sam (gapInfo.txt and stat.txt are still written)
// There is no CPU fallback: if the library or a GPU is missing the program prints why and exits 1.

// GapProber.py builds this file without any flags, i.e. at -O0; the host stages (splitter, preamble) are text parsing
// loops that run twice as fast optimised, so ask for it here.  No fast-math: the preamble stays bit-exact.
#if defined(__GNUC__) && !defined(__clang__) && !defined(__OPTIMIZE__)
#pragma GCC optimize("O2")
#endif

#include <dlfcn.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/figbird_cuda.h"
#include "fb_io.hpp"
#include "fb_preamble.hpp"
#include "fb_split.hpp"

namespace {

struct Api {
    void* h = nullptr;
    decltype(&fb_create) create = nullptr;
    decltype(&fb_warmup) warmup = nullptr;
    decltype(&fb_destroy) destroy = nullptr;
    decltype(&fb_add_gap) add_gap = nullptr;
    decltype(&fb_run) run = nullptr;
    decltype(&fb_fill_gap_files) fill_gap_files = nullptr;   // optional: absent in older builds of the library
    decltype(&fb_set_partition) set_partition = nullptr;
    decltype(&fb_owns_gap) owns_gap = nullptr;
    decltype(&fb_result) result = nullptr;
    decltype(&fb_last_error) last_error = nullptr;
    decltype(&fb_status_str) status_str = nullptr;
    decltype(&fb_abi_version) abi_version = nullptr;
};

std::string exe_dir() {
    char buf[4096];
    ssize_t n = readlink("/proc/self/exe", buf, sizeof buf - 1);
    if (n <= 0) return ".";
    buf[n] = 0;
    std::string p(buf);
    size_t s = p.rfind('/');
    return s == std::string::npos ? "." : p.substr(0, s);
}

bool load_api(Api& a, std::string& why) {
    std::vector<std::string> cand;
    if (const char* e = getenv("FIGBIRD_CUDA_LIB")) cand.push_back(e);
    const std::string d = exe_dir();
    cand.push_back(d + "/libfigbird_cuda.so");
    cand.push_back(d + "/../libfigbird_cuda.so");
    cand.push_back("libfigbird_cuda.so");
    for (auto& c : cand) {
        a.h = dlopen(c.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (a.h) break;
        why = dlerror();
    }
    if (!a.h) return false;
#define FB_SYM(field, name)                                      \
    a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.h, name)); \
    if (!a.field) { why = std::string("missing symbol ") + name; return false; }
    FB_SYM(create, "fb_create")
    FB_SYM(warmup, "fb_warmup")
    FB_SYM(destroy, "fb_destroy")
    FB_SYM(add_gap, "fb_add_gap")
    FB_SYM(run, "fb_run")
    FB_SYM(set_partition, "fb_set_partition")
    FB_SYM(owns_gap, "fb_owns_gap")
    FB_SYM(result, "fb_result")
    FB_SYM(last_error, "fb_last_error")
    FB_SYM(status_str, "fb_status_str")
    FB_SYM(abi_version, "fb_abi_version")
#undef FB_SYM
    a.fill_gap_files = reinterpret_cast<decltype(a.fill_gap_files)>(dlsym(a.h, "fb_fill_gap_files"));
    if (a.abi_version() != FB_ABI_VERSION) { why = "ABI version mismatch"; return false; }
    return true;
}

}  // namespace

int main(int argc, char* argv[]) {
    if (argc < 3) {
        printf("usage: %s <scaffolds.fa> <f> [ignored]\n", argv[0]);
        return 1;
    }
    const char* contig_path = argv[1];
    const int f_mode = atoi(argv[2]);
    const bool trace = getenv("FB_TRACE") != nullptr;      // phase wall times on stderr
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!trace) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[fb_trace] figbird: %-24s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    FILE* out_txt = fopen("out.txt", "w");   // the reference creates it and never writes to it
    fbhost::Scaffolds sc;
    if (!fbhost::read_scaffolds(contig_path, sc)) {
        printf("Can't open contig file\n");
        return 1;
    }
    lap("scaffolds");
    // The CUDA context and the kernels come up on a second thread while this one splits and learns the model.
    Api api;
    std::string why;
    bool api_ok = false;
    const int device = getenv("FIGBIRD_DEVICE") ? atoi(getenv("FIGBIRD_DEVICE")) : 0;
    const int ngpu = getenv("FIGBIRD_GPUS") ? std::max(1, atoi(getenv("FIGBIRD_GPUS"))) : 1;
    std::thread warm;
    if (!getenv("FIGBIRD_PREAMBLE_ONLY"))
        warm = std::thread([&]() {
            api_ok = load_api(api, why);
            if (api_ok)
                for (int k = 0; k < ngpu; k++) api.warmup(device + k);     // failures are reported by fb_create below
        });
    struct Joiner {
        std::thread& t;
        ~Joiner() { if (t.joinable()) t.join(); }
    } joiner{warm};
    const char* from_sam = getenv("FIGBIRD_FROM_SAM");
    fbhost::Splitter* split = nullptr;
    std::string myout_text;
    long total_count = 0, un_count = 0;
    int max_read_len = 0, max_insert = 0;
    if (from_sam) {
        split = new fbhost::Splitter(sc);
        split->myout_text = &myout_text;
        if (!split->run(from_sam, nullptr, nullptr, nullptr, nullptr, false)) {
            printf("Can't open map file\n");
            return 1;
        }
        total_count = split->total_count;
        un_count = split->un_count;
        max_read_len = (int)split->max_read_len;
        max_insert = 5000;
        if (FILE* gi = fopen("gapInfo.txt", "w")) { split->write_gap_info(gi); fclose(gi); }
        if (FILE* st = fopen("stat.txt", "w")) { split->write_stat(st); fclose(st); }
    } else {
        FILE* probe = fopen("myout.sam", "r");
        if (!probe) {
            printf("Can't open map file\n");
            return 1;
        }
        fclose(probe);
        if (FILE* st = fopen("stat.txt", "r")) {
            if (fscanf(st, "%ld %ld %d %d", &total_count, &un_count, &max_read_len, &max_insert) != 4) max_read_len = 0;
            fclose(st);
        }
    }
    if (max_read_len < 1) {
        printf("Can't read stat.txt\n");
        return 1;
    }

    lap("split / stat.txt");
    // ---- preamble (host): model tables + gapProbCutOff
    fbhost::Preamble pre(max_read_len, max_insert, sc.lengths());
    fbhost::Model hm;
    if (!pre.learn("myout.sam", hm, from_sam ? &myout_text : nullptr)) {
        printf("Can't open map file\n");
        return 1;
    }
    if (FILE* mi = fopen("mappedInserts.txt", "w")) {
        for (long i = 0; i < (long)hm.inserts_mapped.size(); i++)
            if (hm.inserts_mapped[i] > 0) fprintf(mi, "%ld,%ld\n", i, hm.inserts_mapped[i]);
        fclose(mi);
    }
    if (out_txt) fclose(out_txt);
    lap("preamble");
    std::cout << "insertThresholdMin = " << hm.thr_min << "...insertThresholdMax = " << hm.thr_max << std::endl;
    if (const char* mo = getenv("FIGBIRD_MODEL_OUT")) hm.save(mo);
    if (getenv("FIGBIRD_PREAMBLE_ONLY")) return 0;

    // ---- the hot path, on the GPU
    if (warm.joinable()) warm.join();
    if (!api_ok) {
        printf("figbird: cannot load libfigbird_cuda.so (%s); there is no CPU fallback\n", why.c_str());
        return 1;
    }
    fb_model m;
    memset(&m, 0, sizeof m);
    m.max_read_len = hm.max_read_len;
    m.error_pos = hm.error_pos.data();
    m.in_pos = hm.in_pos.data();
    m.del_pos = hm.del_pos.data();
    for (int i = 0; i < 5; i++)
        for (int j = 0; j < 5; j++) m.error_type[i][j] = hm.error_type[i][j];
    m.ild = hm.ild_smoothed.data();
    m.ild_len = hm.max_insert;
    m.thr_min = hm.thr_min;
    m.thr_max = hm.thr_max;
    m.gap_prob_cutoff = hm.cutoff;
    m.max_distance = 200;   // maxDistance, figbird.cpp:103
    m.num_iter = 5;         // num_itr, figbird.cpp:2466
    m.f_mode = f_mode;
    std::vector<fb_ctx*> ctxs(ngpu, nullptr);
    for (int k = 0; k < ngpu; k++) {
        int rc = api.create(&m, device + k, &ctxs[k]);
        if (rc != FB_OK) {
            printf("figbird: fb_create failed on device %d: %s%s\n", device + k, api.status_str(rc),
                   rc == FB_EUNSUPPORTED ? " (f must be 0 or 1)" : "");
            return 1;
        }
    }

    lap("library + context");
    time_t t0, t1;
    time(&t0);
    std::vector<fbhost::GapInfo> gaps;
    if (split) {
        for (const fbhost::SplitGap& g : split->gaps) gaps.push_back({g.contig_no, g.start, g.length});
    } else if (!fbhost::read_gap_info("gapInfo.txt", gaps)) {
        printf("Can't open gapInfo.txt\n");
        return 1;
    }
    std::vector<int> handle(gaps.size(), -1);
    std::vector<int> run_rc(ngpu, FB_OK);
    // File route on one GPU: the whole gap loop in one library call — the gap files are read, parsed and packed by the
    // library's host threads, and all but the first batch of that work overlaps the EM scan (fb_fill_gap_files).
    const bool one_call = !split && ngpu == 1 && api.fill_gap_files && !getenv("FIGBIRD_NO_PIPE");
    if (one_call) {
        std::vector<std::string> names;
        std::vector<size_t> which;
        std::vector<fb_gap_desc> descs;
        names.reserve(gaps.size());
        for (size_t g = 0; g < gaps.size(); g++) {
            char name[64];
            snprintf(name, sizeof name, "Gaps/gaps_%zu.sam", g);
            if (access(name, R_OK) != 0) {
                printf("Can't open %s\n", name);
                return 1;
            }
            const fbhost::GapInfo& gi = gaps[g];
            if (gi.contig_no < 0 || gi.contig_no >= (int)sc.seqs.size()) continue;
            const std::string& ctg = sc.seqs[gi.contig_no];
            names.push_back(name);
            which.push_back(g);
            fb_gap_desc d;
            d.contig_no = gi.contig_no;
            d.gap_start = gi.start;
            d.orig_gap = gi.length;
            d.contig = ctg.c_str();
            d.contig_len = (long)ctg.size();
            d.gap_file = nullptr;
            descs.push_back(d);
        }
        for (size_t k = 0; k < descs.size(); k++) descs[k].gap_file = names[k].c_str();
        std::vector<int> h(descs.size(), -1);
        const int rc1 = descs.empty() ? FB_OK : api.fill_gap_files(ctxs[0], descs.data(), (int)descs.size(), h.data(), 0);
        run_rc[0] = rc1 == FB_EDOMAIN ? FB_OK : rc1;     // refused gaps stay N-runs (counted below), the others are filled
        for (size_t k = 0; k < descs.size(); k++) {
            handle[which[k]] = h[k];
            if (h[k] < 0) fprintf(stderr, "figbird: gap %zu left unfilled: %s\n", which[k], api.status_str(h[k]));
        }
        lap("gap files + EM (one call)");
    } else {
        for (size_t g = 0; g < gaps.size(); g++) {
            char name[64];
            snprintf(name, sizeof name, "Gaps/gaps_%zu.sam", g);
            std::vector<fbhost::GapRecord> recs;
            if (split) {
                recs.swap(split->gaps[g].records);
            } else if (!fbhost::read_gap_file(name, recs)) {
                printf("Can't open %s\n", name);
                return 1;
            }
            std::vector<fb_record> rv(recs.size());
            for (size_t r = 0; r < recs.size(); r++) {
                rv[r].pos1 = recs[r].pos1;
                rv[r].len1 = recs[r].len1;
                rv[r].strand1 = recs[r].strand1;
                rv[r].seq2 = recs[r].seq2.c_str();
                rv[r].len2 = (int)recs[r].seq2.size();
            }
            const fbhost::GapInfo& gi = gaps[g];
            if (gi.contig_no < 0 || gi.contig_no >= (int)sc.seqs.size()) continue;
            const std::string& ctg = sc.seqs[gi.contig_no];
            for (int k = 0; k < ngpu; k++) {     // every context sees every gap; the partition decides who fills it
                int h = api.add_gap(ctxs[k], gi.contig_no, gi.start, gi.length, ctg.c_str(), (long)ctg.size(), rv.data(), (int)rv.size());
                if (h < 0 && k == 0) fprintf(stderr, "figbird: gap %zu left unfilled: %s\n", g, api.last_error(ctxs[k]));
                handle[g] = h;
            }
        }
        lap("gap files + pack");
        if (ngpu == 1) {
            run_rc[0] = api.run(ctxs[0]);
        } else {
            std::vector<std::thread> pool;
            for (int k = 0; k < ngpu; k++)
                pool.emplace_back([&, k]() {
                    int rc = api.set_partition(ctxs[k], k, ngpu);
                    run_rc[k] = rc == FB_OK ? api.run(ctxs[k]) : rc;
                });
            for (auto& t : pool) t.join();
        }
    }
    for (int k = 0; k < ngpu; k++)
        if (run_rc[k] != FB_OK) {
            printf("figbird: fb_run failed on device %d: %s: %s\n", device + k, api.status_str(run_rc[k]), api.last_error(ctxs[k]));
            return 1;
        }
    lap("fb_run");
    int rc = FB_OK;
    FILE* gap_out = fopen("gapout.txt", "w");
    std::vector<std::string> filled(gaps.size());
    for (size_t g = 0; g < gaps.size(); g++) {
        const fbhost::GapInfo& gi = gaps[g];
        if (handle[g] >= 0) {
            fb_gap_result res;
            std::vector<char> buf((size_t)2 * gi.length + 16);
            fb_ctx* owner = ctxs[0];
            for (int k = 0; k < ngpu; k++)
                if (api.owns_gap(ctxs[k], handle[g])) owner = ctxs[k];
            rc = api.result(owner, handle[g], &res, buf.data(), (int)buf.size());
            if (rc != FB_OK) {
                printf("figbird: fb_result failed: %s\n", api.last_error(owner));
                return 1;
            }
            filled[g] = buf.data();
            std::cout << "----##############MaxGapEstimate = " << res.best_estimate << "\t" << "MaxLikelihood = " << res.best_lik << std::endl;
            std::cout << "Total_Reads = " << res.total_reads << "\t" << "Discard = " << res.discarded << std::endl;
        } else {
            filled[g].assign((size_t)gi.length, 'N');
        }
        fprintf(gap_out, "%d\t%d\t%ld\t%d\t%d\t%s\n", (int)g, gi.contig_no, gi.start, gi.length, (int)filled[g].size(), filled[g].c_str());
        std::cout << "Gap " << g << " is done." << std::endl;
    }
    fclose(gap_out);
    {
        // Gaps the library refused (closer than maxDistance to a scaffold end, where the reference's own arithmetic leaves
        // its arrays — figbird.cpp:1623 vs 2149 — or mates longer than maxDistance) stay N-runs; say how many, on both streams.
        size_t refused = 0;
        for (size_t g = 0; g < gaps.size(); g++) refused += handle[g] < 0;
        if (refused) {
            std::cout << "figbird: " << refused << " of " << gaps.size() << " gaps left unfilled (outside the reference's defined territory)" << std::endl;
            fprintf(stderr, "figbird: %zu of %zu gaps left unfilled\n", refused, gaps.size());
        }
    }
    for (fb_ctx* cx : ctxs) api.destroy(cx);
    time(&t1);
    std::cout << "Total time to fill gaps in seconds = " << difftime(t1, t0) << std::endl;

    lap("results + gapout.txt");
    // ---- stitch
    std::cout << "# of Contig = " << sc.seqs.size() << std::endl;
    fbhost::stitch(sc, filled, "filledContigs.fa");
    lap("stitch");
    std::cout << "Figbird.cpp ends successfully" << std::endl;
    return 0;
}
