// fb_io.hpp — the file contract either side of the hot path, as figbird's main() implements it:
// scaffold FASTA (figbird.cpp:2586-2658), stat.txt (2680), gapInfo.txt (2819-2837), Gaps/gaps_<n>.sam
// (1718-1801), gapout.txt (2848) and the filledContigs.fa stitcher (2867-2946).
#pragma once

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace fbhost {

struct Scaffolds {
    std::vector<std::string> names;
    std::vector<std::string> seqs;   // upper-cased
    std::vector<long> lengths() const {
        std::vector<long> l;
        for (auto& s : seqs) l.push_back((long)s.size());
        return l;
    }
};

// Reads with a 1024-byte fgets buffer like the reference, so over-long lines and a missing final
// newline behave identically (a last line without '\n' loses its final base there too).
inline bool read_scaffolds(const char* path, Scaffolds& out) {
    FILE* fp = fopen(path, "r");
    if (!fp) return false;
    constexpr int kBuf = 1024;
    char line[kBuf];
    std::string cur;
    bool any = false;
    while (fgets(line, kBuf, fp)) {
        if (line[0] == ';') continue;
        if (line[0] == '>') {
            std::string name(line + 1);
            if (!name.empty()) name.pop_back();
            size_t e = name.find_first_of(" \t\n");
            size_t b = name.find_first_not_of(" \t\n");
            if (b == std::string::npos) name.clear();
            else name = name.substr(b, e == std::string::npos || e < b ? std::string::npos : e - b);
            out.names.push_back(name);
            if (!cur.empty()) {
                out.seqs.push_back(cur);
                cur.clear();
            }
            any = true;
        } else {
            size_t n = strlen(line);
            if (n < (size_t)kBuf - 1) cur.append(line, n ? n - 1 : 0);
            else cur.append(line, n);
        }
    }
    fclose(fp);
    (void)any;
    out.seqs.push_back(cur);
    for (auto& s : out.seqs)
        for (auto& ch : s) ch = (char)toupper((unsigned char)ch);
    return true;
}

struct GapInfo {
    int contig_no;
    long start;
    int length;
};

inline bool read_gap_info(const char* path, std::vector<GapInfo>& out) {
    FILE* fp = fopen(path, "r");
    if (!fp) return false;
    char line[1024];
    while (fgets(line, sizeof line, fp)) {
        char* save = nullptr;
        char* a = strtok_r(line, "\t", &save);
        char* b = a ? strtok_r(nullptr, "\t", &save) : nullptr;
        char* c = b ? strtok_r(nullptr, "\t\n", &save) : nullptr;
        if (!c) continue;
        out.push_back({atoi(a), atol(b), atoi(c)});
    }
    fclose(fp);
    return true;
}

struct GapRecord {
    int pos1, len1, strand1;
    std::string seq2;
};

// Line pairs (anchor, unmapped mate): columns qname flag rname pos cigar tlen seq ... split on tabs.
inline bool read_gap_file(const char* path, std::vector<GapRecord>& out) {
    FILE* fp = fopen(path, "r");
    if (!fp) return false;
    char l1[1024], l2[1024];
    while (fgets(l1, sizeof l1, fp)) {
        if (l1[0] == '@') continue;
        if (!fgets(l2, sizeof l2, fp)) break;
        char *s1 = nullptr, *s2 = nullptr;
        char* col[7];
        bool ok = true;
        col[0] = strtok_r(l1, "\t", &s1);
        for (int k = 1; k < 7 && ok; k++) ok = (col[k] = strtok_r(nullptr, "\t", &s1)) != nullptr;
        if (!ok || !col[0]) continue;
        GapRecord r;
        r.strand1 = (atoi(col[1]) & 16) >> 4;
        r.pos1 = atoi(col[3]);
        r.len1 = (int)strlen(col[6]);
        char* m = strtok_r(l2, "\t", &s2);
        for (int k = 1; k < 7 && m; k++) m = strtok_r(nullptr, "\t", &s2);
        if (!m) continue;
        r.seq2 = m;
        out.push_back(r);
    }
    fclose(fp);
    return true;
}

// filledContigs.fa: every scaffold split at its N-runs, each run replaced by the next gapout.txt string
// (figbird.cpp:2867-2946).  A run is spliced when a non-N follows it; the run state carries across
// scaffolds exactly like the reference's nStart flag.
inline bool stitch(const Scaffolds& sc, const std::vector<std::string>& filled, const char* path) {
    FILE* out = fopen(path, "w");
    if (!out) return false;
    bool in_run = false;
    size_t next = 0;
    for (size_t c = 0; c < sc.seqs.size(); c++) {
        const std::string& s = sc.seqs[c];
        fprintf(out, ">%s\n", c < sc.names.size() ? sc.names[c].c_str() : "");
        std::string seg;
        for (size_t j = 0; j < s.size(); j++) {
            if (s[j] == 'N' || s[j] == 'n') {
                in_run = true;
            } else {
                if (in_run) {
                    fprintf(out, "%s\n", seg.c_str());
                    fprintf(out, "%s\n", next < filled.size() ? filled[next].c_str() : "");
                    next++;
                    seg.clear();
                    in_run = false;
                }
                seg.push_back(s[j]);
            }
        }
        fprintf(out, "%s\n", seg.c_str());
    }
    fclose(out);
    return true;
}

}  // namespace fbhost
