// fb_preamble.hpp — host re-implementation of the reference's preamble: one pass over myout.sam that
// learns the error / insert-size model (figbird.cpp:154-823), and a second pass that scores every pair in
// long double to derive gapProbCutOff (figbird.cpp:931-1351, 2778-2794).  Its outputs are the hot path's
// constant inputs (fb_model in include/figbird_cuda.h), so it is kept bit-exact: same accumulation order,
// same integer/double/long double types, same parsing quirks.  Out of scope for the GPU (≈4 % of the
// reference's runtime, text-bound) — SURVEY §8 row f-2.
#pragma once

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace fbhost {

constexpr int kMaxRecLen = 1024;   // MAX_REC_LEN, figbird.cpp:22

inline int md_base_code(char ch) {   // the switch at figbird.cpp:413-447
    switch (ch) {
        case 'A': return 0;
        case 'C': return 1;
        case 'G': return 2;
        case 'T': return 3;
        default: return 4;
    }
}

inline bool is_md_base(char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'N'; }

struct Model {
    int max_read_len = 0;
    int max_insert = 0;            // maxInsertSize
    int thr_min = 0, thr_max = 0;  // insertThresholdMin/Max
    int cutoff = 0;                // gapProbCutOff
    int max_distance = 200;
    double error_type[5][5];
    std::vector<double> error_pos, in_pos, del_pos, ild_smoothed;
    std::vector<long> inserts_mapped;   // insertCountsMapped, written to mappedInserts.txt by main()

    bool save(const char* path) const {   // "FBM1" layout, see oracle/patch_ref.py
        FILE* f = fopen(path, "wb");
        if (!f) return false;
        int hdr[6] = {max_read_len, max_insert, thr_min, thr_max, cutoff, max_distance};
        fwrite("FBM1", 1, 4, f);
        fwrite(hdr, sizeof(int), 6, f);
        fwrite(error_type, sizeof(double), 25, f);
        fwrite(error_pos.data(), sizeof(double), (size_t)max_read_len, f);
        fwrite(in_pos.data(), sizeof(double), (size_t)max_read_len, f);
        fwrite(del_pos.data(), sizeof(double), (size_t)max_read_len, f);
        fwrite(ild_smoothed.data(), sizeof(double), (size_t)max_insert, f);
        fclose(f);
        return true;
    }
};

// One alignment line of the splitter's dialect (writeSam, main.cpp:180): qname flag contigNo pos cigar tlen seq md IH:i:n
struct SamFields {
    char* qname = nullptr;
    int flag = 0;
    char* rname = nullptr;
    int pos = 0;
    char* cigar = nullptr;
    int tlen = 0;
    char* seq = nullptr;
};

// Tokenise the 7 leading columns in place with strtok semantics on '\t' (empty fields collapse).
inline bool split_leading(char* line, SamFields& f, char** save) {
    char* t;
    if (!(f.qname = strtok_r(line, "\t", save))) return false;
    if (!(t = strtok_r(nullptr, "\t", save))) return false;
    f.flag = atoi(t);
    if (!(f.rname = strtok_r(nullptr, "\t", save))) return false;
    if (!(t = strtok_r(nullptr, "\t", save))) return false;
    f.pos = atoi(t);
    if (!(f.cigar = strtok_r(nullptr, "\t", save))) return false;
    if (!(t = strtok_r(nullptr, "\t", save))) return false;
    f.tlen = atoi(t);
    if (!(f.seq = strtok_r(nullptr, "\t", save))) return false;
    return true;
}

class Preamble {
public:
    // stat.txt's maxReadLength / MAX_INSERT_SIZE (figbird.cpp:2680-2690) and the scaffold lengths
    Preamble(int max_read_len, int max_insert_stat, const std::vector<long>& contig_lengths)
        : mrl_(max_read_len), contig_len_(contig_lengths) {
        hard_max_insert_ = max_insert_stat > 20000 ? max_insert_stat : 20000;
        insert_counts_.assign((size_t)hard_max_insert_, 1);      // pseudo-count 1, figbird.cpp:154-162
        for (auto& row : err_types_)
            for (long& v : row) v = 1;
        for (long& v : base_counts_) v = 1;
        err_pos_.assign(mrl_, 1);
        in_pos_.assign(mrl_, 1);
        in_len_.assign(mrl_, 1);
        del_pos_.assign(mrl_, 1);
        del_len_.assign(mrl_, 1);
        read_len_hist_.assign(mrl_, 0);
        no_error_cigar_ = std::to_string(mrl_) + "M";
    }

    // processMapping, figbird.cpp:825-900.  `line` is modified.
    void add_mapping(char* line) {
        char* save = nullptr;
        SamFields f;
        if (!split_leading(line, f, &save)) return;
        const int strand = (f.flag & 16) >> 4;
        char* tok;
        while ((tok = strtok_r(nullptr, "\t\n", &save)) != nullptr) {
            if (tok[0] == 'M' && tok[1] == 'D') md_ = tok;                 // sticky across lines, like the reference's locals
            else if (tok[0] == 'I' && tok[1] == 'H') nh_ = atoi(tok + 5);
        }
        if (nh_ == 1 && md_.size() > 5 && md_[5] != '^') {
            long ctg = atol(f.rname);
            if (ctg >= 0 && ctg < (long)contig_len_.size() && contig_len_[ctg] > 0) count_insert(f.tlen);   // inputMean is 0
            count_errors(f.cigar, md_.c_str(), f.seq, strand);
            unique_reads_++;
        }
    }

    // Pass 1 over the file, then computeProbabilites, then pass 2 (computeLikelihood) and the cut-off.
    // With `mem` the same text is read from memory (the splitter's in-memory hand-off) instead of from myout_path.
    // Both passes only accumulate integer counts, so large inputs are cut into chunks scanned by several host threads
    // (FIGBIRD_THREADS, default one per hardware thread up to 16) and the counts are added up: same result, bit for bit.
    bool learn(const char* myout_path, Model& out, const std::string* mem = nullptr) {
        std::string owned;
        const std::string* text = mem;
        if (!text) {
            FILE* fp = fopen(myout_path, "rb");
            if (!fp) return false;
            fseek(fp, 0, SEEK_END);
            long sz = ftell(fp);
            fseek(fp, 0, SEEK_SET);
            owned.resize(sz > 0 ? (size_t)sz : 0);
            if (sz > 0 && fread(&owned[0], 1, (size_t)sz, fp) != (size_t)sz) { fclose(fp); return false; }
            fclose(fp);
            text = &owned;
        }
        int threads = 1;
        if (const char* e = getenv("FIGBIRD_THREADS")) threads = atoi(e);
        else threads = (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
        std::vector<size_t> cut1, cut2;
        if (threads > 1 && text->size() >= (size_t)(4 << 20)) plan_chunks(*text, threads, cut1, cut2);
        if (cut1.size() <= 2) {
            cut1 = {0, text->size()};
            cut2 = cut1;
        }
        const int T = (int)cut1.size() - 1;
        // ---- pass 1
        {
            std::vector<Preamble> workers;
            for (int t = 1; t < T; t++) workers.emplace_back(mrl_, hard_max_insert_, contig_len_);
            std::vector<std::thread> pool;
            for (int t = 1; t < T; t++)
                pool.emplace_back([&, t]() { workers[t - 1].scan_range(text->data() + cut1[t], cut1[t + 1] - cut1[t]); });
            scan_range(text->data() + cut1[0], cut1[1] - cut1[0]);
            for (auto& th : pool) th.join();
            for (Preamble& w : workers) absorb(w);
        }
        derive_distributions();
        // ---- pass 2
        {
            gap_hist_.assign(1000, 0);
            std::vector<std::vector<long>> hist(T, std::vector<long>(1000, 0)), mapped(T, std::vector<long>((size_t)max_insert_, 0));
            std::vector<char> ok(T, 1);
            std::vector<std::thread> pool;
            auto work = [&](int t) {
                std::vector<long> eff(eff_len_);
                ok[t] = score_range(text->data() + cut2[t], cut2[t + 1] - cut2[t], hist[t], mapped[t], eff, /*flush_last_group=*/t != T - 1);
            };
            for (int t = 1; t < T; t++) pool.emplace_back(work, t);
            work(0);
            for (auto& th : pool) th.join();
            for (int t = 0; t < T; t++) {
                if (!ok[t]) return false;
                for (int i = 0; i < 1000; i++) gap_hist_[i] += hist[t][i];
                for (int i = 0; i < max_insert_; i++) inserts_mapped_[i] += mapped[t][i];
            }
        }
        // gapProbCutOff: first bin where the cumulative count reaches 99 % (figbird.cpp:2778-2794)
        long total = 0;
        for (long v : gap_hist_) total += v;
        long run = 0;
        cutoff_ = 0;
        for (int i = 0; i < 1000; i++) {
            run += gap_hist_[i];
            if (run >= 0.99 * total) { cutoff_ = i; break; }
        }
        out.max_read_len = mrl_;
        out.max_insert = max_insert_;
        out.thr_min = thr_min_;
        out.thr_max = thr_max_;
        out.cutoff = cutoff_;
        for (int i = 0; i < 5; i++)
            for (int j = 0; j < 5; j++) out.error_type[i][j] = err_type_p_[i][j];
        out.error_pos = err_pos_d_;
        out.in_pos = in_pos_d_;
        out.del_pos = del_pos_d_;
        out.ild_smoothed = ild_smooth_;
        out.inserts_mapped = inserts_mapped_;
        return true;
    }

    int thr_min() const { return thr_min_; }
    int thr_max() const { return thr_max_; }

private:
    static FILE* open_range(const char* p, size_t n) { return n ? fmemopen((void*)p, n, "r") : fopen("/dev/null", "r"); }

    // pass 1 over a range of whole lines
    void scan_range(const char* p, size_t n) {
        FILE* fp = open_range(p, n);
        if (!fp) return;
        std::vector<char> line(kMaxRecLen);
        while (fgets(line.data(), kMaxRecLen, fp)) {
            if (line[0] == '@') continue;
            add_mapping(line.data());
        }
        fclose(fp);
    }

    // add another worker's pass-1 counts to this one (every counter starts from the same pseudo-count)
    void absorb(const Preamble& w) {
        if (w.insert_counts_.size() > insert_counts_.size()) insert_counts_.resize(w.insert_counts_.size(), 1);
        for (size_t i = 0; i < w.insert_counts_.size(); i++) insert_counts_[i] += w.insert_counts_[i] - 1;
        discarded_ += w.discarded_;
        unique_reads_ += w.unique_reads_;
        for (int i = 0; i < 5; i++) {
            base_counts_[i] += w.base_counts_[i] - 1;
            for (int j = 0; j < 5; j++) err_types_[i][j] += w.err_types_[i][j] - 1;
        }
        for (int i = 0; i < mrl_; i++) {
            err_pos_[i] += w.err_pos_[i] - 1;
            in_pos_[i] += w.in_pos_[i] - 1;
            in_len_[i] += w.in_len_[i] - 1;
            del_pos_[i] += w.del_pos_[i] - 1;
            del_len_[i] += w.del_len_[i] - 1;
            read_len_hist_[i] += w.read_len_hist_[i];
        }
    }

    // Cut the text into about `parts` ranges.  cut1: at line starts (pass 1).  cut2: at the start of a line PAIR whose
    // names differ from the pair before it, so that no multi-hit group of pass 2 straddles a cut.  Refuses (leaves the
    // cuts empty -> single range) when the text has '@' lines, lines the reference's 1023-byte fgets would split, or
    // lines without an MD or IH tag (their values carry over from the line before, which a chunk start would lose).
    static void plan_chunks(const std::string& text, int parts, std::vector<size_t>& cut1, std::vector<size_t>& cut2) {
        const size_t n = text.size();
        const char* d = text.data();
        if (n == 0 || d[n - 1] != '\n') return;
        std::vector<size_t> tentative;
        for (int t = 1; t < parts; t++) {
            size_t pos = n / parts * t;
            const char* nl = (const char*)memchr(d + pos, '\n', n - pos);
            if (!nl) break;
            pos = (size_t)(nl - d) + 1;
            if (pos < n && (tentative.empty() || pos > tentative.back())) tentative.push_back(pos);
        }
        // line parity of every tentative cut + sanity of the whole text, chunk by chunk in parallel
        std::vector<size_t> bounds = {0};
        bounds.insert(bounds.end(), tentative.begin(), tentative.end());
        bounds.push_back(n);
        const int C = (int)bounds.size() - 1;
        std::vector<size_t> lines(C, 0);
        std::vector<char> bad(C, 0);
        std::vector<std::thread> pool;
        for (int c = 0; c < C; c++)
            pool.emplace_back([&, c]() {
                const char* p = d + bounds[c];
                const char* e = d + bounds[c + 1];
                size_t cnt = 0;
                while (p < e) {
                    if (*p == '@') bad[c] = 1;
                    const char* nl = (const char*)memchr(p, '\n', (size_t)(e - p));
                    if (!nl) { bad[c] = 1; break; }
                    if (nl - p >= kMaxRecLen - 1) bad[c] = 1;
                    // the MD and IH values are sticky across lines in the reference (locals that a line without the tag
                    // does not reset): a chunk may only start afresh if every line carries both tags itself
                    if (!memmem(p, (size_t)(nl - p), "\tMD", 3) || !memmem(p, (size_t)(nl - p), "\tIH", 3)) bad[c] = 1;
                    cnt++;
                    p = nl + 1;
                }
                lines[c] = cnt;
            });
        for (auto& th : pool) th.join();
        for (int c = 0; c < C; c++)
            if (bad[c]) return;
        cut1 = bounds;
        // pass-2 cuts: move each cut forward to an even line, then further while the pair repeats the names of the pair before
        auto next_line = [&](size_t pos) { const char* nl = (const char*)memchr(d + pos, '\n', n - pos); return nl ? (size_t)(nl - d) + 1 : n; };
        auto name_at = [&](size_t pos) { size_t e = pos; while (e < n && d[e] != '\t' && d[e] != '\n') e++; return std::string(d + pos, e - pos); };
        auto prev_line = [&](size_t pos) { size_t p = pos - 1; while (p > 0 && d[p - 1] != '\n') p--; return p; };   // pos = start of a line, > 0
        cut2 = {0};
        size_t seen = 0;
        for (int c = 1; c < C; c++) {
            seen += lines[c - 1];
            size_t pos = bounds[c];
            if (seen & 1) pos = next_line(pos);                  // an odd number of lines before: skip the pair's second line
            while (pos < n) {
                const size_t l2 = next_line(pos);
                if (l2 >= n) { pos = n; break; }
                const size_t p2 = prev_line(pos), p1 = prev_line(p2);
                if (name_at(pos) != name_at(p1) || name_at(l2) != name_at(p2)) break;   // a new group starts here
                pos = next_line(l2);
            }
            if (pos < n && pos > cut2.back()) cut2.push_back(pos);
        }
        cut2.push_back(n);
        if (cut2.size() != cut1.size()) {       // keep the two cut lists the same length (one range per thread)
            while (cut2.size() < cut1.size()) cut2.insert(cut2.end() - 1, cut2[cut2.size() - 2]);
        }
    }

    // updateInsertCounts, figbird.cpp:164-203
    void count_insert(int v) {
        if (v <= 0) return;
        if (v < (int)insert_counts_.size()) { insert_counts_[v]++; return; }
        if (v > hard_max_insert_) { discarded_++; return; }
        size_t grown = std::max(insert_counts_.size() * 2, (size_t)v + 1);
        insert_counts_.resize(grown, 1);
        insert_counts_[v]++;
    }

    // processErrorTypes, figbird.cpp:269-465
    void count_errors(const char* cigar, const char* md, const char* read, int strand) {
        int rl = 0;
        for (; read[rl]; rl++) base_counts_[md_base_code(read[rl])]++;      // getLength, figbird.cpp:233-253
        if (rl < 1 || rl > mrl_) return;
        read_len_hist_[rl - 1]++;
        if (strcmp(md, no_error_cigar_.c_str()) == 0) return;   // compares an MD tag with a CIGAR: never equal
        const long md_len = (long)strlen(md) - 5;
        std::vector<int> ins(rl + 1, 0);
        int index = 0, cur = 0;
        // CIGAR: <n><op>...
        for (const char* p = cigar; *p;) {
            while (*p && strchr("IDMS^\t\n ", *p)) p++;
            if (!*p) break;
            const char* q = p;
            while (*q && !strchr("IDMS^\t\n ", *q)) q++;
            const int n = atoi(p);
            const char op = *q;
            if (op == 'M') { index += n; cur += n; }
            else if (op == 'I' || op == 'S') {
                bump(in_pos_, strand == 0 ? index : rl - index - 1);
                bump(in_len_, n - 1);
                if (cur >= 0 && cur <= rl) ins[cur] = n;
                index += n;
            } else if (op == 'D') {
                bump(del_pos_, strand == 0 ? index : rl - index - 1);
                bump(del_len_, n - 1);
            }
            p = *q ? q + 1 : q;
        }
        // MD: numbers separated by mismatched bases or ^deletions
        walk_md(md, md_len, rl, ins, read, strand, nullptr);
    }

    void bump(std::vector<long>& v, long i) {
        if (i >= 0 && i < (long)v.size()) v[i]++;
    }

    // Shared MD walk (figbird.cpp:366-462 and 1030-1128).  With prob == nullptr it counts (pass 1);
    // otherwise it multiplies the mismatch factors into *prob (pass 2).
    void walk_md(const char* md, long md_len, int rl, const std::vector<int>& ins, const char* read, int strand, long double* prob) {
        const char* body = md + 5;
        long total = 0;
        int index = 0;
        const char* p = body;
        while (*p) {
            while (*p && strchr("ACGTN^\t\n ", *p)) p++;   // strtok skips delimiter runs
            if (!*p) break;
            const char* q = p;
            while (*q && !strchr("ACGTN^\t\n ", *q)) q++;
            const long tok_len = q - p;
            const int num = atoi(p);
            total += tok_len;
            if (total < md_len) {
                char from = body[total];
                if (from == '^') {
                    total++;
                    index += num;
                    for (long i = total; i < md_len; i++) {
                        if (is_md_base(body[total])) total++;
                        else break;
                    }
                } else if (is_md_base(from)) {
                    total++;
                    index += num + 1;
                    int cur = 0;
                    for (int i = 0; i < index && i <= rl; i++) cur += ins[i];
                    const int rp = index - 1 + cur;
                    const char to = (rp >= 0 && rp < rl) ? read[rp] : 'N';
                    const long slot = strand == 0 ? rp : rl - index - cur;
                    const int f = md_base_code(from), t = md_base_code(to);
                    if (!prob) {
                        bump(err_pos_, slot);
                        if (f != t) err_types_[f][t]++;
                    } else if (slot >= 0 && slot < mrl_) {
                        *prob = *prob * err_pos_d_[slot] / (1 - err_pos_d_[slot] - in_pos_d_[slot] - del_pos_d_[slot]);   // :1073
                        if (f != t) *prob *= base_err_[f] * err_type_p_[f][t];                                                // :1120
                    }
                } else {
                    break;
                }
            }
            p = q;
        }
    }

    // computeProbabilites, figbird.cpp:475-823
    void derive_distributions() {
        for (int i = 0; i < 5; i++) {
            int cnt = 0;
            for (int j = 0; j < 5; j++) cnt += (int)err_types_[i][j];
            for (int j = 0; j < 5; j++) err_type_p_[i][j] = (double)err_types_[i][j] / cnt;
            base_err_[i] = cnt / (double)base_counts_[i];
        }
        double sum = 0;
        for (int i = 0; i < 4; i++) sum += base_err_[i];
        for (int i = 0; i < 4; i++) base_err_[i] = 4 * base_err_[i] / sum;
        base_err_[4] = 1;
        for (int i = mrl_ - 1; i > 0; i--) read_len_hist_[i - 1] = read_len_hist_[i] + read_len_hist_[i - 1];   // reads covering position i
        err_pos_d_.resize(mrl_);
        in_pos_d_.resize(mrl_);
        del_pos_d_.resize(mrl_);
        in_len_d_.resize(mrl_);
        del_len_d_.resize(mrl_);
        for (int i = 0; i < mrl_; i++) err_pos_d_[i] = (double)err_pos_[i] / read_len_hist_[i];
        for (int i = 0; i < mrl_; i++) in_pos_d_[i] = (double)in_pos_[i] / read_len_hist_[i];
        int in_total = 0;
        for (int i = 0; i < mrl_; i++) in_total += (int)in_len_[i];
        for (int i = 0; i < mrl_; i++) in_len_d_[i] = (double)in_len_[i] / in_total;
        for (int i = 0; i < mrl_; i++) del_pos_d_[i] = (double)del_pos_[i] / read_len_hist_[i];
        int del_total = 0;
        for (int i = 0; i < mrl_; i++) del_total += (int)del_len_[i];
        for (int i = 0; i < mrl_; i++) del_len_d_[i] = (double)del_len_[i] / del_total;

        max_insert_ = (int)insert_counts_.size();
        const int M = max_insert_;
        ild_.resize(M);
        long ins_count = discarded_;
        sum = 0;
        for (int i = 0; i < M; i++) {
            ins_count += (insert_counts_[i] - 1);
            sum += i * (insert_counts_[i] - 1);
        }
        ins_mean_ = sum / ins_count;
        sum = 0;
        for (int i = 0; i < M; i++) {
            ild_[i] = (double)insert_counts_[i] / ins_count;
            sum += (insert_counts_[i] - 1) * (ins_mean_ - i) * (ins_mean_ - i);
        }
        no_err_.resize(mrl_);
        double ne = 1.0;
        for (int i = 0; i < mrl_; i++) {
            ne *= (1 - err_pos_d_[i] - in_pos_d_[i] - del_pos_d_[i]);
            no_err_[i] = ne;
        }
        eff_len_.assign(M, -1);
        long total_len = 0;
        for (long l : contig_len_) total_len += l;
        eff_len_[0] = total_len;
        inserts_mapped_.assign(M, 0);

        // moving average of half-width 12, ends copied (figbird.cpp:624-649), then the pseudo-count shift (:651-655)
        const int W = 12;
        ild_smooth_.assign(M, 0.0);
        for (int i = 0; i < W; i++) ild_smooth_[i] = ild_[i];
        double win = 0;
        for (int i = 0; i < 2 * W + 1; i++) win += ild_[i];
        ild_smooth_[W] = win / (2 * W + 1);
        for (int i = W + 1; i < M - W; i++) {
            win -= ild_[i - W - 1];
            win += ild_[i + W];
            ild_smooth_[i] = win / (2 * W + 1);
        }
        for (int i = M - W; i < M; i++) ild_smooth_[i] = ild_[i];
        for (int i = 0; i < M; i++) ild_smooth_[i] = ild_smooth_[i] - 1 / (double)(ins_count) + (1 / (double)M) / (double)(ins_count + 1);

        // one-sided standard deviations around the mean (figbird.cpp:761-789)
        double isum = 0, icount = 0;
        for (int i = (int)(ins_mean_ + 1); i < M; i++) {
            isum = isum + (insert_counts_[i] - 1) * (i - ins_mean_) * (i - ins_mean_);
            icount += (insert_counts_[i] - 1);
        }
        const double right_sd = sqrt(isum / icount);
        isum = 0;
        icount = 0;
        for (int i = std::max((int)(ins_mean_ - 10 * right_sd), 0); i < ins_mean_; i++) {
            isum = isum + (insert_counts_[i] - 1) * (ins_mean_ - i) * (ins_mean_ - i);
            icount += (insert_counts_[i] - 1);
        }
        const double left_sd = sqrt(isum / icount);
        if (right_sd > 1000 || left_sd > 1000) fprintf(stderr, "Switching to conservative mode\n");
        thr_min_ = std::max((int)(ins_mean_ - 2.5 * left_sd), 1);
        thr_max_ = std::min((int)(ins_mean_ + 2.5 * right_sd), M);
    }

    // getEffectiveLength, figbird.cpp:902-929
    long effective_length(int ins, std::vector<long>& cache) const {
        auto compute = [&](int v) {
            long e = 0;
            for (long l : contig_len_)
                if (l >= v) e += (l - v + 1);
            return e;
        };
        if (ins < 0) return cache[0];
        if (ins >= max_insert_) return compute(ins);
        if (cache[ins] == -1) cache[ins] = compute(ins);
        return cache[ins];
    }

    // computeErrorProb, figbird.cpp:931-1132 (long double)
    long double read_error_prob(const char* cigar, const char* md, const char* read, int strand) {
        const int rl = (int)strlen(read);
        if (rl < 1 || rl > mrl_) return 0;
        long double prob = no_err_[rl - 1];
        if (md[5] == '^') return prob;
        const long md_len = (long)strlen(md) - 5;
        std::vector<int> ins(rl + 1, 0);
        int index = 0, cur = 0;
        for (const char* p = cigar; *p;) {
            while (*p && strchr("IDM^\t\n ", *p)) p++;
            if (!*p) break;
            const char* q = p;
            while (*q && !strchr("IDM^\t\n ", *q)) q++;
            const int n = atoi(p);
            const char op = *q;
            if (op == 'M') { index += n; cur += n; }
            else if (op == 'I') {
                long i = strand == 0 ? index : rl - index - 1;
                if (i >= 0 && i < mrl_ && n >= 1 && n <= mrl_)
                    prob = prob * in_pos_d_[i] * in_len_d_[n - 1] / (1 - err_pos_d_[i] - in_pos_d_[i] - del_pos_d_[i]);   // :995
                if (cur >= 0 && cur <= rl) ins[cur] = n;
                index += n;
            } else if (op == 'D') {
                long i = strand == 0 ? index : rl - index - 1;
                if (i >= 0 && i < mrl_ && n >= 1 && n <= mrl_)
                    prob = prob * del_pos_d_[i] * del_len_d_[n - 1] / (1 - err_pos_d_[i] - in_pos_d_[i] - del_pos_d_[i]); // :1014
            }
            p = *q ? q + 1 : q;
        }
        walk_md(md, md_len, rl, ins, read, strand, &prob);
        return prob;
    }

    // computeLikelihood, figbird.cpp:1135-1351: only its side effects (gapProbs histogram, insertCountsMapped) are used.
    // One range of line pairs.  The reference never flushes its last group (figbird.cpp:1339-1346 only adds the log);
    // a range that is not the last one of the file does flush it, because the group is complete there.
    bool score_range(const char* p, size_t n, std::vector<long>& hist, std::vector<long>& mapped, std::vector<long>& eff_cache,
                     bool flush_last_group) {
        FILE* fp = open_range(p, n);
        if (!fp) return false;
        std::vector<char> l1(kMaxRecLen), l2(kMaxRecLen);
        std::string pre1 = "*", pre2 = "*", md1, md2;
        long double sum = 0.0, logsum = 0.0, best = 0, gap_p = 0;
        int best_ins = 0;
        bool any = false;
        auto flush_hist = [&]() {
            long double v = -logl(gap_p);
            int idx = (std::isnan(v) || v >= 2147483648.0L || v <= -2147483649.0L) ? INT_MIN : (int)v;
            if (idx < 1000 && idx >= 0) hist[idx]++;
            else hist[999]++;
            if (best_ins >= max_insert_) mapped[max_insert_ - 1]++;
            else mapped[best_ins]++;
        };
        while (fgets(l1.data(), kMaxRecLen, fp)) {
            if (l1[0] == '@') continue;
            if (!fgets(l2.data(), kMaxRecLen, fp)) break;
            char *s1 = nullptr, *s2 = nullptr, *tok;
            SamFields a, b;
            if (!split_leading(l1.data(), a, &s1)) continue;
            while ((tok = strtok_r(nullptr, "\t\n", &s1)) != nullptr)
                if (tok[0] == 'M' && tok[1] == 'D') md1 = tok;
            if (!split_leading(l2.data(), b, &s2)) continue;
            while ((tok = strtok_r(nullptr, "\t\n", &s2)) != nullptr)
                if (tok[0] == 'M' && tok[1] == 'D') md2 = tok;
            const int st1 = (a.flag & 16) >> 4, st2 = (b.flag & 16) >> 4;
            const int ins = std::max(a.tlen, b.tlen);
            long double ins_p = 0;
            if (ins >= 0 && ins < max_insert_) ins_p = ild_[ins];
            if (ins_p == 0) ins_p = 1 / (double)unique_reads_;
            const long double e1 = read_error_prob(a.cigar, md1.c_str(), a.seq, st1);
            const long double e2 = read_error_prob(b.cigar, md2.c_str(), b.seq, st2);
            const long eff = effective_length(ins, eff_cache);
            any = true;
            const long double prob = (1 / (long double)(eff)) * ins_p * e1 * e2;
            if (pre1 == a.qname && pre2 == b.qname) {
                if (best < prob) {
                    best = prob;
                    best_ins = ins < 0 ? 0 : ins;
                    gap_p = ins_p * e2;
                }
                sum += prob;
            } else if (pre1 != "*" && pre2 != "*") {
                if (sum < 1e-320 || std::isnan(sum)) sum = 1e-320;
                logsum += logl(sum);
                flush_hist();
                sum = prob;
                best = prob;
                best_ins = ins < 0 ? 0 : ins;
                gap_p = ins_p * e2;
            } else {
                sum = prob;
                best = prob;
                best_ins = ins < 0 ? 0 : ins;
                gap_p = ins_p * e2;
            }
            pre1 = a.qname;
            pre2 = b.qname;
            if (std::isinf(logsum)) { fclose(fp); return false; }
        }
        fclose(fp);
        if (flush_last_group && any) flush_hist();
        return true;
    }

    int mrl_;
    std::vector<long> contig_len_;
    int hard_max_insert_ = 20000;
    std::vector<long> insert_counts_;
    long discarded_ = 0, unique_reads_ = 0;
    long err_types_[5][5];
    long base_counts_[5];
    std::vector<long> err_pos_, in_pos_, in_len_, del_pos_, del_len_, read_len_hist_;
    std::string no_error_cigar_, md_;
    int nh_ = 0;
    // derived
    double err_type_p_[5][5];
    double base_err_[5];
    std::vector<double> err_pos_d_, in_pos_d_, del_pos_d_, in_len_d_, del_len_d_, ild_, ild_smooth_, no_err_;
    std::vector<long> eff_len_, inserts_mapped_, gap_hist_;
    int max_insert_ = 0, thr_min_ = 0, thr_max_ = 0, cutoff_ = 0;
    double ins_mean_ = 0;
};

}  // namespace fbhost
