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
#pragma once

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
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

class Splitter {
public:
    int max_distance = 200;          // main.cpp:80
    long total_count = 0, un_count = 0;
    unsigned long max_read_len = 0;
    std::vector<SplitGap> gaps;
    std::vector<double> contig_read_counts;
    std::vector<std::vector<int>> read_counts;

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
        contig_read_counts.assign(sc.seqs.size(), 0.0);
        read_counts.assign(sc.seqs.size(), std::vector<int>(sc.seqs.size(), 0));
    }

    // Stream a SAM file.  myout/single/linking/unmapped may be null (in-memory use); keep_text fills SplitGap::text.
    bool run(const char* sam_path, FILE* myout, FILE* single, FILE* linking, FILE* unmapped, bool keep_text) {
        FILE* fp = fopen(sam_path, "r");
        if (!fp) return false;
        myout_ = myout; single_ = single; linking_ = linking; unmapped_ = unmapped; keep_text_ = keep_text;
        char line[1024];
        std::string pre1 = "*", pre2 = "*", q;
        SamRec r1, r2;
        bool end = false;
        while (fgets(line, sizeof line, fp)) {
            if (line[0] == '@') continue;
            parse(line, r1);
            while ((r1.flag & 2) == 0) {
                // lines of one read, then lines of its mate (main.cpp:1073-1113)
                q = r1.qname();
                int seg = r1.flag & 192;
                bool stuck = false;
                while (q == r1.qname() && (r1.flag & 192) == seg) {
                    mixed1_.push_back(r1);
                    if (fgets(line, sizeof line, fp)) parse(line, r1);
                    else { stuck = true; break; }   // the reference spins forever here; stop instead
                }
                if (stuck) { end = true; flush_mixed(); break; }
                q = r1.qname();
                seg = r1.flag & 192;
                while (q == r1.qname() && (r1.flag & 192) == seg) {
                    mixed2_.push_back(r1);
                    if (fgets(line, sizeof line, fp)) parse(line, r1);
                    else { end = true; break; }
                }
                flush_mixed();
                if (end) break;
            }
            if (end) break;
            if (fgets(line, sizeof line, fp)) parse(line, r2);
            else r2 = r1;
            if (pre1 != r1.qname() || pre2 != r2.qname()) {
                pre1 = r1.qname();
                pre2 = r2.qname();
                flush_proper();
            }
            reads1_.push_back(r1);
            reads2_.push_back(r2);
        }
        flush_proper();
        fclose(fp);
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

    void parse(const char* src, SamRec& r) {
        static const char* D = "\t\n ";
        r.flag = r.pos = r.tlen = 0;
        r.nm = -1;
        r.ih = 0;
        r.has_md = false;
        r.seq_len = 0;
        size_t n = strlen(src);
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
            auto it = name_to_no_.find(r.rname());
            r.contig_no = it == name_to_no_.end() ? -1 : it->second;
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
    static void write_sam(const SamRec& r, FILE* f) {
        if (!f) return;
        std::string s;
        sam_line(r, s);
        fwrite(s.data(), 1, s.size(), f);
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
                if (strcmp(r.rname(), "*") != 0 && r.nm == 0 && r.contig_no >= 0) contig_read_counts[r.contig_no] += 1 / (double)n0;
    }

    int check_pos(long contig_no, long pos, int strand) const {   // main.cpp:197-219: first gap whose window holds the anchor
        for (size_t i = 0; i < gaps.size(); i++) {
            const SplitGap& g = gaps[i];
            if (contig_no != g.contig_no) continue;
            const long gs = g.start, ge = g.start + g.length;
            if ((strand == 0 && pos > gs - max_distance && pos < ge) || (strand == 1 && pos > ge && pos < ge + max_distance)) return (int)i;
        }
        return -1;
    }

    void fastq_entry(const SamRec& r) {
        if (!unmapped_) return;
        const int strand = (r.flag & 16) >> 4;
        const std::string s = strand == 1 ? revcomp_upper_only(r.seq(), r.seq_len) : std::string(r.seq());
        fprintf(unmapped_, "@%s\n%s\n+%s\n%s\n", r.qname(), s.c_str(), r.qname(), r.qual());
    }

    void to_gap(int gi, const SamRec& anchor, const SamRec& mate) {
        SplitGap& g = gaps[gi];
        GapRecord rec;
        rec.pos1 = anchor.pos;
        rec.len1 = (int)anchor.seq_len;
        rec.strand1 = (anchor.flag & 16) >> 4;
        rec.seq2 = mate.seq();
        g.records.push_back(rec);
        if (keep_text_) {
            sam_line(anchor, g.text);
            sam_line(mate, g.text);
        }
    }

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
                        if (r1.seq_len > max_read_len) max_read_len = r1.seq_len;
                        if (r2.seq_len > max_read_len) max_read_len = r2.seq_len;
                        size_t n1 = 0, n2 = 0;
                        for (const char* q = r1.seq(); *q; q++) n1 += (*q == 'N' || *q == 'n');
                        for (const char* q = r2.seq(); *q; q++) n2 += (*q == 'N' || *q == 'n');
                        if (n1 / (double)r1.seq_len < 0.8 && n2 / (double)r2.seq_len < 0.8) {
                            un_count++;
                            total_count++;
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
                        for (const SamRec& r : a) write_sam(r, single_);
                        for (const SamRec& r : b) write_sam(r, single_);
                        for (const SamRec& r : a) {
                            int gi = check_pos(r.contig_no, r.pos, (r.flag & 16) >> 4);
                            if (gi >= 0) to_gap(gi, r, b[0]);
                        }
                        done = true;
                    } else if (u1 && !u2) {
                        for (const SamRec& r : a) write_sam(r, single_);
                        for (const SamRec& r : b) write_sam(r, single_);
                        for (const SamRec& r : b) {
                            int gi = check_pos(r.contig_no, r.pos, (r.flag & 16) >> 4);
                            if (gi >= 0) to_gap(gi, r, a[0]);
                        }
                        done = true;
                    } else if (strcmp(r1.rname(), r2.rname()) != 0) {
                        r1.ih = (long)a.size();
                        r2.ih = (long)b.size();
                        if (r1.contig_no >= 0 && r2.contig_no >= 0) read_counts[r1.contig_no][r2.contig_no]++;
                        write_sam(r1, linking_);
                        write_sam(r2, linking_);
                    }
                }
            }
        }
        a.clear();
        b.clear();
    }

    // printVectors, main.cpp:221-429
    void flush_proper() {
        const size_t ih = reads1_.size();
        count_perfect(reads1_);
        count_perfect(reads2_);
        for (size_t i = 0; i < ih; i++) {
            SamRec& r1 = reads1_[i];
            SamRec& r2 = reads2_[i];
            if (strcmp(r1.rname(), "*") == 0 || strcmp(r2.rname(), "*") == 0) continue;   // "never here" in the reference
            if (strcmp(r1.rname(), r2.rname()) != 0) {
                if (r1.contig_no >= 0 && r2.contig_no >= 0) read_counts[r1.contig_no][r2.contig_no]++;
                printf("Warning: contig names different\n");
                continue;
            }
            if (r1.contig_no >= 0) read_counts[r1.contig_no][r2.contig_no]++;
            if (r1.seq_len > max_read_len) max_read_len = r1.seq_len;
            if (r2.seq_len > max_read_len) max_read_len = r2.seq_len;
            r1.ih = r2.ih = (long)ih;
            write_sam(r1, myout_);
            write_sam(r2, myout_);
        }
        total_count++;
        reads1_.clear();
        reads2_.clear();
    }

    const Scaffolds& sc_;
    std::unordered_map<std::string, long> name_to_no_;
    std::string last_rname_;          // one-entry cache in front of the name map
    long last_contig_ = -1;
    unsigned short last_rname_off_valid_ = 0xffff;
    std::vector<SamRec> mixed1_, mixed2_, reads1_, reads2_;
    FILE *myout_ = nullptr, *single_ = nullptr, *linking_ = nullptr, *unmapped_ = nullptr;
    bool keep_text_ = false;
};

}  // namespace fbhost
