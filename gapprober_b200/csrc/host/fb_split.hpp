// fb_split.hpp — the stage in front of the hot path: GapProber's SAM splitter (reference main.cpp), restated so
// that its products can be handed to figbird in memory instead of through Gaps/*.sam.
//
// What the reference does (main.cpp:820-1205): find the N-runs of the scaffolds (gaps), then stream result.sam:
//   * proper pairs (FLAG & 2)  -> myout.sam (the preamble's input), two writeSam lines per pair (main.cpp:221-429)
//   * other groups: consecutive lines of one read, then of its mate (main.cpp:1070-1115); when exactly one mate is
//     mapped and its position falls in a gap's window (checkPos, main.cpp:197-219) the (anchor, mate) pair goes to that
//     gap's file (main.cpp:640-726) — these pairs are the records of the hot path.
// Here the same decisions are taken once and kept as GapRecord vectors; the files can still be written for
// compatibility (byte-identical where the reference's output is defined — it prints uninitialised md/ih fields for
// unpaired records, main.cpp:180, which are written as "*" / 0 here and which figbird never reads).
//
// The reference streams the file with one state machine.  Here the (memory-mapped) file is cut into stretches that
// run the SAME machine on several host threads.  A stretch starts speculatively at a line where the read name
// changes and runs until the machine is back at its top level at or behind the stretch's end.  The stretches are then
// taken in file order: a stretch counts only if it started exactly where (and how) its predecessor stopped and does
// not continue the predecessor's last group of proper pairs — otherwise it is run again from the predecessor's state —
// and its outputs are appended.  The one thing the sequential machine carries across a cut, the group of proper pairs
// it has not flushed yet (it is flushed at the next name change), is carried by the merge and flushed at that same
// point, so every output byte, every counter and every floating-point addition happens in the reference's order.
#pragma once

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "fb_io.hpp"

namespace fbhost {

struct SplitGap {
    int contig_no;
    long start;
    int length;
    std::vector<GapRecord> records;   // (anchor pos/len/strand, mate sequence) in file order
    std::string text;                 // the gap file's lines, when file output is wanted
};

struct SamRec {   // getSAM, main.cpp:755-810 (tokens split on tab, newline AND blank).  One copy of the line, fields by offset.
    char line[1024];
    unsigned short o_qname = 0, o_rname = 0, o_cigar = 0, o_seq = 0, o_qual = 0, o_md = 0;
    unsigned short seq_len = 0;
    int flag = 0, pos = 0, tlen = 0, nm = -1;
    long contig_no = -1;
    long ih = 0;
    bool has_md = false;
    SamRec() { line[0] = 0; line[1] = '*'; line[2] = 0; }
    const char* qname() const { return line + o_qname; }
    const char* rname() const { return line + o_rname; }
    const char* cigar() const { return line + o_cigar; }
    const char* seq() const { return line + o_seq; }
    const char* qual() const { return line + o_qual; }
    const char* md() const { return has_md ? line + o_md : "*"; }
    bool same_name(const SamRec& o) const { return strcmp(qname(), o.qname()) == 0; }
};

// What the machine needs to know and never changes while it runs.
struct SplitShared {
    int max_distance = 200;                              // main.cpp:80
    const std::vector<SplitGap>* gaps = nullptr;         // the workers read the positions only
    const std::unordered_map<std::string, long>* name_to_no = nullptr;
    const std::vector<std::pair<int, int>>* gaps_of_contig = nullptr;   // [contig] -> index range of its gaps (they are consecutive)
    bool want_myout = false, want_single = false, want_linking = false, want_unmapped = false, keep_text = false;
};

// The state the sequential machine has when it stands at its top level: the names of the last proper pair it saw
// and the group of proper pairs it is still collecting.
struct ProperState {
    std::string pre1 = "*", pre2 = "*";
    std::vector<SamRec> reads1, reads2;
};

// Everything one stretch of the SAM produces, in file order.
struct SplitOut {
    std::string myout, single, linking, unmapped;
    struct Hit {
        int gap;
        GapRecord rec;
    };
    std::vector<Hit> hits;                               // records for the gap files
    std::vector<std::pair<int, std::string>> hit_text;   // (gap, two lines) when file output is wanted
    long total_count = 0, un_count = 0, warnings = 0;
    unsigned long max_read_len = 0;
    // NM:i:0 counting (main.cpp:451-491) adds 1/n0 to a double per hit: the additions are logged and replayed in file
    // order.  In a speculative stretch {-1, 0} marks its first name change among proper pairs — the point where the
    // sequential machine flushes the group it carried in.
    std::vector<std::pair<int, int>> perfect;
    std::unordered_map<unsigned long long, int> pair_counts;   // readCounts[c1][c2] increments
    bool saw_proper = false;
    std::string first_pre1, first_pre2;                  // names of the first proper pair
    ProperState tail;                                    // the machine's state at the end of the stretch
    size_t stop = 0;            // where the next stretch has to start
    bool stop_pending = false;  // ... with a line the machine had already taken as a record (no '@' header test)
};

class SplitWorker {
public:
    SplitWorker(const SplitShared& sh, SplitOut& out) : sh_(sh), out_(out) {}

    // Run the reference's loop (main.cpp:1055-1130) over base[begin, file_end), handing over once the machine is at its
    // top level at or behind `limit`.  first_is_record: the first line was already accepted as a record by the
    // previous stretch (it is not tested for '@').  init: the state the machine arrives with; a speculative stretch has
    // none (nullptr) and only marks where its first flush of proper pairs would have happened.
    void run(const char* base, size_t begin, size_t limit, size_t file_end, bool first_is_record, ProperState* init) {
        base_ = base;
        pos_ = begin;
        end_ = file_end;
        const bool speculative = init == nullptr;
        std::string pre1 = "*", pre2 = "*", q;
        if (init) {
            pre1 = init->pre1;
            pre2 = init->pre2;
            reads1_.swap(init->reads1);
            reads2_.swap(init->reads2);
        }
        SamRec r1, r2;
        size_t r1_start = begin;
        bool end = false, handed_over = false, pending = first_is_record;
        for (;;) {
            if (!pending) {
                if (pos_ >= limit || pos_ >= end_) break;
                if (base_[pos_] == '@') {        // header lines are only skipped at the top level (main.cpp:1057)
                    pos_ += line_len();
                    continue;
                }
            }
            if (!next(r1, r1_start)) break;
            pending = false;
            while ((r1.flag & 2) == 0) {
                // lines of one read, then lines of its mate (main.cpp:1073-1113)
                q = r1.qname();
                int seg = r1.flag & 192;
                bool stuck = false;
                while (q == r1.qname() && (r1.flag & 192) == seg) {
                    mixed1_.push_back(r1);
                    if (!next(r1, r1_start)) { stuck = true; break; }   // the reference spins forever here; stop instead
                }
                if (stuck) { end = true; flush_mixed(); break; }
                q = r1.qname();
                seg = r1.flag & 192;
                while (q == r1.qname() && (r1.flag & 192) == seg) {
                    mixed2_.push_back(r1);
                    if (!next(r1, r1_start)) { end = true; break; }
                }
                flush_mixed();
                if (end) break;
                if (r1_start >= limit) {          // r1 is the first record of the next stretch
                    handed_over = true;
                    break;
                }
            }
            if (end || handed_over) break;
            size_t r2_start;
            if (!next(r2, r2_start)) r2 = r1;
            if (!out_.saw_proper) {
                out_.saw_proper = true;
                out_.first_pre1 = r1.qname();
                out_.first_pre2 = r2.qname();
            }
            if (pre1 != r1.qname() || pre2 != r2.qname()) {
                const bool first_change = speculative && pre1 == "*" && pre2 == "*" && !marked_;
                pre1 = r1.qname();
                pre2 = r2.qname();
                if (first_change) {
                    marked_ = true;
                    out_.perfect.emplace_back(-1, 0);     // the carried group is flushed here, by the merge
                } else {
                    flush_group(reads1_, reads2_);
                }
            }
            reads1_.push_back(r1);
            reads2_.push_back(r2);
        }
        out_.tail.pre1 = pre1;
        out_.tail.pre2 = pre2;
        out_.tail.reads1.swap(reads1_);
        out_.tail.reads2.swap(reads2_);
        out_.stop = handed_over ? r1_start : pos_;
        out_.stop_pending = handed_over;
    }

    // printVectors, main.cpp:221-429
    void flush_group(std::vector<SamRec>& g1, std::vector<SamRec>& g2) {
        const size_t ih = g1.size();
        count_perfect(g1);
        count_perfect(g2);
        for (size_t i = 0; i < ih; i++) {
            SamRec& r1 = g1[i];
            SamRec& r2 = g2[i];
            if (strcmp(r1.rname(), "*") == 0 || strcmp(r2.rname(), "*") == 0) continue;   // "never here" in the reference
            if (strcmp(r1.rname(), r2.rname()) != 0) {
                if (r1.contig_no >= 0 && r2.contig_no >= 0) count_pair(r1.contig_no, r2.contig_no);
                out_.warnings++;           // "Warning: contig names different", printed by the merge
                continue;
            }
            if (r1.contig_no >= 0) count_pair(r1.contig_no, r2.contig_no);
            if (r1.seq_len > out_.max_read_len) out_.max_read_len = r1.seq_len;
            if (r2.seq_len > out_.max_read_len) out_.max_read_len = r2.seq_len;
            r1.ih = r2.ih = (long)ih;
            if (sh_.want_myout) {
                sam_line(r1, out_.myout);
                sam_line(r2, out_.myout);
            }
        }
        out_.total_count++;
        g1.clear();
        g2.clear();
    }

private:
    // fgets(line, 1024, fp): at most 1023 bytes, up to and including the newline
    size_t line_len() const {
        const size_t room = std::min<size_t>(1023, end_ - pos_);
        const void* nl = memchr(base_ + pos_, '\n', room);
        return nl ? (size_t)((const char*)nl - (base_ + pos_)) + 1 : room;
    }
    bool next(SamRec& r, size_t& start) {
        if (pos_ >= end_) return false;
        const size_t n = line_len();
        start = pos_;
        parse(base_ + pos_, n, r);
        pos_ += n;
        return true;
    }

    void parse(const char* src, size_t len, SamRec& r) {
        static const char* D = "\t\n ";
        r.flag = r.pos = r.tlen = 0;
        r.nm = -1;
        r.ih = 0;
        r.has_md = false;
        r.seq_len = 0;
        size_t n = strnlen(src, len);
        if (n > sizeof r.line - 4) n = sizeof r.line - 4;
        memcpy(r.line + 3, src, n);             // bytes 0..2 hold "\0*\0": offset 0 = empty string, offset 1 = "*"
        r.line[0] = 0; r.line[1] = '*'; r.line[2] = 0;
        r.line[3 + n] = 0;
        r.o_qname = r.o_cigar = r.o_seq = r.o_qual = r.o_md = 0;
        r.o_rname = 0;
        char* save = nullptr;
        auto off = [&](char* t) { return (unsigned short)(t - r.line); };
        char* t = strtok_r(r.line + 3, D, &save);
        if (!t) { r.contig_no = -1; return; }
        r.o_qname = off(t);
        if ((t = strtok_r(nullptr, D, &save))) r.flag = atoi(t);
        if ((t = strtok_r(nullptr, D, &save))) r.o_rname = off(t);
        if ((t = strtok_r(nullptr, D, &save))) r.pos = atoi(t);
        strtok_r(nullptr, D, &save);                                  // mapq
        if ((t = strtok_r(nullptr, D, &save))) r.o_cigar = off(t);
        strtok_r(nullptr, D, &save);                                  // rnext
        strtok_r(nullptr, D, &save);                                  // pnext
        if ((t = strtok_r(nullptr, D, &save))) r.tlen = atoi(t);
        if ((t = strtok_r(nullptr, D, &save))) { r.o_seq = off(t); r.seq_len = (unsigned short)strlen(t); }
        if ((t = strtok_r(nullptr, D, &save))) r.o_qual = off(t);
        while ((t = strtok_r(nullptr, D, &save)) != nullptr) {
            if (t[0] == 'M' && t[1] == 'D') { r.o_md = off(t); r.has_md = true; }
            if (t[0] == 'N' && t[1] == 'M') r.nm = atoi(t + 5);
        }
        if (r.o_rname == last_rname_off_valid_ && last_rname_ == r.rname()) {
            r.contig_no = last_contig_;
        } else {
            auto it = sh_.name_to_no->find(r.rname());
            r.contig_no = it == sh_.name_to_no->end() ? -1 : it->second;
            last_rname_ = r.rname();
            last_contig_ = r.contig_no;
            last_rname_off_valid_ = r.o_rname;
        }
    }

    static void sam_line(const SamRec& r, std::string& out) {   // writeSam, main.cpp:175-183
        char buf[1400];
        int n = snprintf(buf, sizeof buf, "%s\t%d\t%ld\t%d\t%s\t%d\t%s\t%s\tIH:i:%ld\n", r.qname(), r.flag, r.contig_no, r.pos,
                         r.cigar(), r.tlen, r.seq(), r.md(), r.ih);
        out.append(buf, (size_t)(n < (int)sizeof buf ? n : (int)sizeof buf - 1));
    }

    static std::string revcomp_upper_only(const char* s, size_t len) {   // main.cpp's reverse(): upper case only, rest -> N
        std::string o(len, 'N');
        for (size_t i = 0; i < len; i++) {
            char ch = s[i], c = 'N';
            if (ch == 'A') c = 'T';
            else if (ch == 'C') c = 'G';
            else if (ch == 'G') c = 'C';
            else if (ch == 'T') c = 'A';
            o[len - 1 - i] = c;
        }
        return o;
    }

    void count_perfect(const std::vector<SamRec>& v) {   // contigReadCounts from NM:i:0 hits (main.cpp:451-491)
        int n0 = 0;
        for (const SamRec& r : v)
            if (strcmp(r.rname(), "*") != 0 && r.nm == 0) n0++;
        if (n0 > 0)
            for (const SamRec& r : v)
                if (strcmp(r.rname(), "*") != 0 && r.nm == 0 && r.contig_no >= 0) out_.perfect.emplace_back((int)r.contig_no, n0);
    }

    int check_pos(long contig_no, long pos, int strand) const {   // main.cpp:197-219: first gap whose window holds the anchor
        const std::vector<SplitGap>& gaps = *sh_.gaps;
        // the reference walks all gaps and skips those of other scaffolds; a scaffold's gaps are consecutive, so walk only them
        if (contig_no < 0 || contig_no >= (long)sh_.gaps_of_contig->size()) return -1;
        const std::pair<int, int> range = (*sh_.gaps_of_contig)[(size_t)contig_no];
        for (int i = range.first; i < range.second; i++) {
            const SplitGap& g = gaps[(size_t)i];
            if (contig_no != g.contig_no) continue;
            const long gs = g.start, ge = g.start + g.length;
            if ((strand == 0 && pos > gs - sh_.max_distance && pos < ge) || (strand == 1 && pos > ge && pos < ge + sh_.max_distance)) return (int)i;
        }
        return -1;
    }

    void fastq_entry(const SamRec& r) {
        if (!sh_.want_unmapped) return;
        const int strand = (r.flag & 16) >> 4;
        const std::string s = strand == 1 ? revcomp_upper_only(r.seq(), r.seq_len) : std::string(r.seq());
        std::string& o = out_.unmapped;
        o += '@'; o += r.qname(); o += '\n'; o += s; o += "\n+"; o += r.qname(); o += '\n'; o += r.qual(); o += '\n';
    }

    void to_gap(int gi, const SamRec& anchor, const SamRec& mate) {
        SplitOut::Hit h;
        h.gap = gi;
        h.rec.pos1 = anchor.pos;
        h.rec.len1 = (int)anchor.seq_len;
        h.rec.strand1 = (anchor.flag & 16) >> 4;
        h.rec.seq2 = mate.seq();
        out_.hits.push_back(std::move(h));
        if (sh_.keep_text) {
            out_.hit_text.emplace_back(gi, std::string());
            sam_line(anchor, out_.hit_text.back().second);
            sam_line(mate, out_.hit_text.back().second);
        }
    }

    void count_pair(long c1, long c2) { out_.pair_counts[((unsigned long long)c1 << 32) | (unsigned long long)(unsigned)c2]++; }

    // printMixedVectors, main.cpp:432-753
    void flush_mixed() {
        std::vector<SamRec>& a = mixed1_;
        std::vector<SamRec>& b = mixed2_;
        count_perfect(a);
        count_perfect(b);
        if (!a.empty() && !b.empty()) {
            bool done = false;
            for (size_t i = 0; i < a.size() && !done; i++) {
                for (size_t j = 0; j < b.size() && !done; j++) {
                    SamRec& r1 = a[i];
                    SamRec& r2 = b[j];
                    if (i == 0 && j == 0) {
                        if (r1.seq_len > out_.max_read_len) out_.max_read_len = r1.seq_len;
                        if (r2.seq_len > out_.max_read_len) out_.max_read_len = r2.seq_len;
                        size_t n1 = 0, n2 = 0;
                        for (const char* q = r1.seq(); *q; q++) n1 += (*q == 'N' || *q == 'n');
                        for (const char* q = r2.seq(); *q; q++) n2 += (*q == 'N' || *q == 'n');
                        if (n1 / (double)r1.seq_len < 0.8 && n2 / (double)r2.seq_len < 0.8) {
                            out_.un_count++;
                            out_.total_count++;
                            fastq_entry(r1);
                            fastq_entry(r2);
                        } else {
                            done = true;
                            break;
                        }
                    }
                    const bool u1 = (r1.flag & 4) != 0, u2 = (r2.flag & 4) != 0;
                    if (u1 && u2) {
                        done = true;
                    } else if (!u1 && u2) {
                        if (sh_.want_single) {
                            for (const SamRec& r : a) sam_line(r, out_.single);
                            for (const SamRec& r : b) sam_line(r, out_.single);
                        }
                        for (const SamRec& r : a) {
                            int gi = check_pos(r.contig_no, r.pos, (r.flag & 16) >> 4);
                            if (gi >= 0) to_gap(gi, r, b[0]);
                        }
                        done = true;
                    } else if (u1 && !u2) {
                        if (sh_.want_single) {
                            for (const SamRec& r : a) sam_line(r, out_.single);
                            for (const SamRec& r : b) sam_line(r, out_.single);
                        }
                        for (const SamRec& r : b) {
                            int gi = check_pos(r.contig_no, r.pos, (r.flag & 16) >> 4);
                            if (gi >= 0) to_gap(gi, r, a[0]);
                        }
                        done = true;
                    } else if (strcmp(r1.rname(), r2.rname()) != 0) {
                        r1.ih = (long)a.size();
                        r2.ih = (long)b.size();
                        if (r1.contig_no >= 0 && r2.contig_no >= 0) count_pair(r1.contig_no, r2.contig_no);
                        if (sh_.want_linking) {
                            sam_line(r1, out_.linking);
                            sam_line(r2, out_.linking);
                        }
                    }
                }
            }
        }
        a.clear();
        b.clear();
    }

    const SplitShared& sh_;
    SplitOut& out_;
    const char* base_ = nullptr;
    size_t pos_ = 0, end_ = 0;
    bool marked_ = false;
    std::string last_rname_;          // one-entry cache in front of the name map
    long last_contig_ = -1;
    unsigned short last_rname_off_valid_ = 0xffff;
    std::vector<SamRec> mixed1_, mixed2_, reads1_, reads2_;
};

class Splitter {
public:
    int max_distance = 200;          // main.cpp:80
    long total_count = 0, un_count = 0;
    unsigned long max_read_len = 0;
    std::vector<SplitGap> gaps;
    std::vector<double> contig_read_counts;
    std::vector<std::vector<int>> read_counts;
    int threads = 0;                 // 0: FIGBIRD_THREADS, else the hardware's (at most 16)
    size_t stretch_bytes = 0;        // 0: FIGBIRD_SPLIT_STRETCH, else 8 MiB
    long stretches = 0, reruns = 0;  // reruns: stretches whose speculative start was wrong (0 on well-formed input)
    std::string* myout_text = nullptr;   // in-memory use: myout.sam's text is appended here instead of written to a file

    explicit Splitter(const Scaffolds& sc) : sc_(sc) {
        // gaps = N-runs followed by a non-N; the run state is NOT reset between scaffolds (main.cpp:938-972)
        bool in_run = false;
        int count = 0;
        long start = 0;
        for (size_t c = 0; c < sc.seqs.size(); c++) {
            const std::string& s = sc.seqs[c];
            for (size_t j = 0; j < s.size(); j++) {
                if (s[j] == 'N' || s[j] == 'n') {
                    if (!in_run) { in_run = true; count = 1; start = (long)j; }
                    else count++;
                } else if (in_run) {
                    if (count >= 1) gaps.push_back(SplitGap{(int)c, start, count, {}, {}});
                    in_run = false;
                }
            }
        }
        for (size_t c = 0; c < sc.names.size() && c < sc.seqs.size(); c++) name_to_no_.emplace(sc.names[c], (long)c);   // first wins, like the hash chain
        gaps_of_contig_.assign(sc.seqs.size(), std::make_pair(0, 0));
        for (size_t i = 0; i < gaps.size(); i++) {            // contig numbers never decrease along the vector
            std::pair<int, int>& r = gaps_of_contig_[(size_t)gaps[i].contig_no];
            if (r.second == 0) r.first = (int)i;
            r.second = (int)i + 1;
        }
        contig_read_counts.assign(sc.seqs.size(), 0.0);
        read_counts.assign(sc.seqs.size(), std::vector<int>(sc.seqs.size(), 0));
    }

    // Split a SAM file.  myout/single/linking/unmapped may be null (in-memory use); keep_text fills SplitGap::text.
    bool run(const char* sam_path, FILE* myout, FILE* single, FILE* linking, FILE* unmapped, bool keep_text) {
        const int fd = open(sam_path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0) { close(fd); return false; }
        const size_t size = (size_t)st.st_size;
        const char* base = nullptr;
        if (size) {
            void* m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m == MAP_FAILED) { close(fd); return false; }
            base = (const char*)m;
            madvise(m, size, MADV_SEQUENTIAL);
        }
        SplitShared sh;
        sh.max_distance = max_distance;
        sh.gaps = &gaps;
        sh.name_to_no = &name_to_no_;
        sh.gaps_of_contig = &gaps_of_contig_;
        sh.want_myout = myout != nullptr || myout_text != nullptr;
        sh.want_single = single != nullptr;
        sh.want_linking = linking != nullptr;
        sh.want_unmapped = unmapped != nullptr;
        sh.keep_text = keep_text;

        int nth = threads;
        if (nth <= 0 && getenv("FIGBIRD_THREADS")) nth = atoi(getenv("FIGBIRD_THREADS"));
        if (nth <= 0) nth = (int)std::min<unsigned>(16, std::max<unsigned>(1, std::thread::hardware_concurrency()));
        size_t sb = stretch_bytes;
        if (!sb && getenv("FIGBIRD_SPLIT_STRETCH")) sb = (size_t)atol(getenv("FIGBIRD_SPLIT_STRETCH"));
        if (!sb) sb = (size_t)8 << 20;

        // speculative stretch starts: the first line start at or behind k * sb whose read name differs from the line before
        std::vector<size_t> starts{0};
        for (size_t target = sb; target < size; target += sb) {
            const size_t s = name_change_at_or_after(base, std::max(target, starts.back() + 1), size);
            if (s >= size) break;
            if (s > starts.back()) starts.push_back(s);
        }
        const size_t nstretch = starts.size();
        stretches = (long)nstretch;

        ProperState carried;                          // what the sequential machine holds at the current cut
        size_t expect = 0;                            // where the next stretch has to start, and how
        bool expect_pending = false;
        // waves of nth stretches; the workers of the next wave run while this thread checks and appends the previous one
        struct Wave {
            size_t w0 = 0, w1 = 0;
            std::vector<SplitOut> outs;
            std::vector<std::thread> pool;
        };
        auto launch = [&](Wave& w, size_t w0) {
            w.w0 = w0;
            w.w1 = std::min(nstretch, w0 + (size_t)nth);
            w.outs.clear();
            w.outs.resize(w.w1 - w.w0);
            w.pool.clear();
            for (size_t k = w.w0; k < w.w1; k++)
                w.pool.emplace_back([&sh, &w, &starts, base, size, nstretch, k]() {
                    SplitWorker wk(sh, w.outs[k - w.w0]);
                    wk.run(base, starts[k], k + 1 < nstretch ? starts[k + 1] : size, size, false, nullptr);
                });
        };
        Wave waves[2];
        int cur = 0;
        launch(waves[cur], 0);
        for (;;) {
            Wave& w = waves[cur];
            for (auto& t : w.pool) t.join();
            w.pool.clear();
            const bool more = w.w1 < nstretch;
            if (more) launch(waves[cur ^ 1], w.w1);
            for (size_t k = w.w0; k < w.w1; k++) {
                SplitOut& o = w.outs[k - w.w0];
                const size_t limit = k + 1 < nstretch ? starts[k + 1] : size;
                // the stretch is the sequential machine's continuation only if it began where (and how) the previous one
                // stopped and its first proper pair starts a new group
                const bool same_start = starts[k] == expect && (!expect_pending || expect >= size || base[expect] != '@');
                const bool new_group = !o.saw_proper || o.first_pre1 != carried.pre1 || o.first_pre2 != carried.pre2;
                if (same_start && new_group) {
                    merge(sh, o, myout, single, linking, unmapped, &carried);
                    if (o.saw_proper) carried = std::move(o.tail);
                } else {
                    SplitOut redo;
                    SplitWorker wk(sh, redo);
                    wk.run(base, expect, std::max(limit, expect), size, expect_pending, &carried);   // continues from the carried state
                    merge(sh, redo, myout, single, linking, unmapped, nullptr);
                    carried = std::move(redo.tail);
                    o.stop = redo.stop;
                    o.stop_pending = redo.stop_pending;
                    reruns++;
                }
                expect = o.stop;
                expect_pending = o.stop_pending;
            }
            w.outs.clear();
            if (!more) break;
            cur ^= 1;
        }
        {   // the final printVectors call (main.cpp:1130)
            SplitOut last;
            SplitWorker wk(sh, last);
            wk.flush_group(carried.reads1, carried.reads2);
            merge(sh, last, myout, single, linking, unmapped, nullptr);
        }
        if (base) munmap((void*)base, size);
        close(fd);
        return true;
    }

    void write_gap_info(FILE* f) const {
        for (const SplitGap& g : gaps) fprintf(f, "%d\t%ld\t%d\n", g.contig_no, g.start, g.length);   // main.cpp:965
    }
    void write_stat(FILE* f) const { fprintf(f, "%ld %ld %ld %ld", total_count, un_count, max_read_len, 5000L); }   // main.cpp:1192
    void write_counts(FILE* f) const {
        for (size_t i = 0; i < sc_.seqs.size(); i++) fprintf(f, "%d %ld %lf\n", (int)i, (long)sc_.seqs[i].size(), contig_read_counts[i]);
    }
    void write_scaffold_matrix(FILE* f) const {   // main.cpp:1162-1182
        for (size_t i = 0; i < sc_.seqs.size(); i++) fprintf(f, ";%s", name(i));
        fprintf(f, ";\n");
        for (size_t i = 0; i < sc_.seqs.size(); i++) {
            fprintf(f, "%s;", name(i));
            for (size_t j = 0; j < sc_.seqs.size(); j++) fprintf(f, "%d;", (i != j && read_counts[i][j] > 10) ? read_counts[i][j] : 0);
            fprintf(f, "\n");
        }
    }

private:
    const char* name(size_t i) const { return i < sc_.names.size() ? sc_.names[i].c_str() : ""; }

    static const char* field0_end(const char* p, const char* e) {
        while (p < e && *p != '\t' && *p != '\n' && *p != ' ') p++;
        return p;
    }
    // first line start >= from whose first field differs from the previous line's (the lines of a read are consecutive)
    static size_t name_change_at_or_after(const char* base, size_t from, size_t size) {
        const char* e = base + size;
        if (from >= size) return size;
        const void* nl = memchr(base + from - 1, '\n', size - (from - 1));
        size_t cur = nl ? (size_t)((const char*)nl - base) + 1 : size;      // a line start at or behind `from`
        if (cur >= size) return size;
        size_t prev = cur - 1;                                               // the line before it
        while (prev > 0 && base[prev - 1] != '\n') prev--;
        while (cur < size) {
            const char* a0 = base + prev;
            const char* a1 = field0_end(a0, e);
            const char* b0 = base + cur;
            const char* b1 = field0_end(b0, e);
            if ((a1 - a0) != (b1 - b0) || memcmp(a0, b0, (size_t)(a1 - a0)) != 0) return cur;
            const void* n2 = memchr(base + cur, '\n', size - cur);
            if (!n2) return size;
            prev = cur;
            cur = (size_t)((const char*)n2 - base) + 1;
        }
        return size;
    }

    // Append one stretch's products.  carried: the group of proper pairs the sequential machine still holds when it
    // enters a speculative stretch; it is flushed where the stretch marked its first name change.
    void merge(const SplitShared& sh, SplitOut& o, FILE* myout, FILE* single, FILE* linking, FILE* unmapped, ProperState* carried) {
        bool flushed = false;
        for (const auto& e : o.perfect) {
            if (e.first < 0) {
                if (carried) {
                    SplitOut fl;
                    SplitWorker wk(sh, fl);
                    wk.flush_group(carried->reads1, carried->reads2);
                    merge(sh, fl, myout, single, linking, unmapped, nullptr);     // its lines precede this stretch's in myout.sam
                    flushed = true;
                }
            } else {
                contig_read_counts[e.first] += 1 / (double)e.second;
            }
        }
        (void)flushed;
        if (myout && !o.myout.empty()) fwrite(o.myout.data(), 1, o.myout.size(), myout);
        if (myout_text) myout_text->append(o.myout);
        if (single && !o.single.empty()) fwrite(o.single.data(), 1, o.single.size(), single);
        if (linking && !o.linking.empty()) fwrite(o.linking.data(), 1, o.linking.size(), linking);
        if (unmapped && !o.unmapped.empty()) fwrite(o.unmapped.data(), 1, o.unmapped.size(), unmapped);
        for (SplitOut::Hit& h : o.hits) gaps[h.gap].records.push_back(std::move(h.rec));
        for (auto& t : o.hit_text) gaps[t.first].text += t.second;
        total_count += o.total_count;
        un_count += o.un_count;
        if (o.max_read_len > max_read_len) max_read_len = o.max_read_len;
        for (long i = 0; i < o.warnings; i++) printf("Warning: contig names different\n");
        for (const auto& pc : o.pair_counts) read_counts[(size_t)(pc.first >> 32)][(size_t)(pc.first & 0xffffffffu)] += pc.second;
    }

    const Scaffolds& sc_;
    std::unordered_map<std::string, long> name_to_no_;
    std::vector<std::pair<int, int>> gaps_of_contig_;
};

}  // namespace fbhost
