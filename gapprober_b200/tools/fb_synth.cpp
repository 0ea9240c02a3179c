// fb_synth — seeded synthetic workload generator for the figbird hot path.
//
// Emits what GapProber's pipeline would see for a wgsim-style paired-end library:
//   scaffolds.fa   scaffolds with N-runs (gaps)
//   truth.txt      gapNo ctg start scafLen trueLen trueSeq
//   result.sam     bowtie2-dialect SAM (the input of the reference splitter, main.cpp:1055-1153)
// and, with --direct, the files the splitter writes for figbird (main.cpp:175-183, 965, 1192):
//   gapInfo.txt  stat.txt  myout.sam  Gaps/gaps_<n>.sam
// bowtie2 / wgsim / samtools are not available in the build image, so this stands in for them.
// SAM dialect requirements come from the reference's parser (main.cpp:755-810): >=11 fields, no
// blanks inside fields, MD:Z/NM:i tags on mapped reads, proper pairs carry flag&2, pairs with one
// unmapped mate are two consecutive lines with equal qname and different flag&192, the unmapped mate
// carries its FASTQ-orientation sequence and the anchor's RNAME/POS.
//
// A read "maps" iff it does not overlap a gap of the truth genome.  Everything is driven by one
// splitmix64 stream so a (seed, options) pair always yields the same bytes.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <sys/stat.h>
#include <vector>

namespace {

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0x1234567ull) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double uni() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
    uint64_t below(uint64_t n) { return n ? next() % n : 0; }
    double normal() {
        double u1 = uni(), u2 = uni();
        if (u1 < 1e-300) u1 = 1e-300;
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
};

struct GapSpec {
    int ctg;
    long tstart;   // start in truth coordinates
    int tlen;      // true length
    long sstart;   // start in scaffold coordinates
    int slen;      // N-run length in the scaffold
};

struct Opt {
    std::string out = "synth_out";
    uint64_t seed = 1;
    long genome = 100000;
    int scaffolds = 1;
    int gaps = 1;
    int gapMin = 500, gapMax = 500;
    std::vector<int> gapList;
    double cov = 30;
    int L = 100;
    double insMean = 500, insSd = 50;
    double err = 0.0, nrate = 0.0;
    int multihit = 1;
    long modelPairs = -1;   // cap on proper pairs written (-1 = all)
    int lenNoise = 0;       // scaffold N-run length = true length + U[-lenNoise, lenNoise]
    bool direct = false, sam = true;
    int margin = 1000;      // keep gaps this far from scaffold ends
    int spacing = 800;      // minimum distance between gaps
};

const char BASES[5] = "ACGT";

inline char comp(char c) {
    switch (c) {
        case 'A': return 'T';
        case 'C': return 'G';
        case 'G': return 'C';
        case 'T': return 'A';
        default: return 'N';
    }
}

void revcomp(const char* in, int n, char* out) {
    for (int i = 0; i < n; i++) out[n - 1 - i] = comp(in[i]);
    out[n] = 0;
}

// MD:Z / NM:i for a substitution-only alignment of `read` against `ref`.
int makeMd(const char* ref, const char* read, int n, char* md) {
    int nm = 0, run = 0;
    char* p = md;
    for (int i = 0; i < n; i++) {
        if (ref[i] == read[i]) {
            run++;
        } else {
            p += sprintf(p, "%d%c", run, ref[i]);
            run = 0;
            nm++;
        }
    }
    sprintf(p, "%d", run);
    return nm;
}

void usage() {
    fprintf(stderr,
            "fb_synth --out DIR [--seed S] [--genome N] [--scaffolds K] [--gaps G] [--gap-min a] [--gap-max b]\n"
            "         [--gap-list l1,l2,...] [--cov C] [--read-len L] [--ins-mean m] [--ins-sd s] [--err e]\n"
            "         [--nrate r] [--multihit k] [--model-pairs M] [--len-noise d] [--direct] [--no-sam]\n");
    exit(2);
}

}  // namespace

int main(int argc, char** argv) {
    Opt o;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto need = [&](void) -> const char* {
            if (i + 1 >= argc) usage();
            return argv[++i];
        };
        if (a == "--out") o.out = need();
        else if (a == "--seed") o.seed = strtoull(need(), nullptr, 10);
        else if (a == "--genome") o.genome = atol(need());
        else if (a == "--scaffolds") o.scaffolds = atoi(need());
        else if (a == "--gaps") o.gaps = atoi(need());
        else if (a == "--gap-min") o.gapMin = atoi(need());
        else if (a == "--gap-max") o.gapMax = atoi(need());
        else if (a == "--gap-list") {
            const char* s = need();
            while (*s) {
                o.gapList.push_back(atoi(s));
                while (*s && *s != ',') s++;
                if (*s == ',') s++;
            }
        } else if (a == "--cov") o.cov = atof(need());
        else if (a == "--read-len") o.L = atoi(need());
        else if (a == "--ins-mean") o.insMean = atof(need());
        else if (a == "--ins-sd") o.insSd = atof(need());
        else if (a == "--err") o.err = atof(need());
        else if (a == "--nrate") o.nrate = atof(need());
        else if (a == "--multihit") o.multihit = atoi(need());
        else if (a == "--model-pairs") o.modelPairs = atol(need());
        else if (a == "--len-noise") o.lenNoise = atoi(need());
        else if (a == "--margin") o.margin = atoi(need());
        else if (a == "--spacing") o.spacing = atoi(need());
        else if (a == "--direct") o.direct = true;
        else if (a == "--no-sam") o.sam = false;
        else usage();
    }
    if (o.L < 10 || o.L > 150) { fprintf(stderr, "read length out of range\n"); return 2; }
    if (!o.gapList.empty()) o.gaps = (int)o.gapList.size();

    Rng rng(o.seed);
    mkdir(o.out.c_str(), 0755);
    auto path = [&](const char* f) { return o.out + "/" + f; };

    // ---- truth genome, split into scaffolds -------------------------------------------------
    const int K = std::max(1, o.scaffolds);
    std::vector<std::string> truth(K);
    for (int c = 0; c < K; c++) {
        long len = o.genome / K;
        truth[c].resize(len);
        for (long i = 0; i < len; i++) truth[c][i] = BASES[rng.next() & 3];
    }

    // ---- gap placement: round-robin over scaffolds, rejection-sampled for spacing ------------
    std::vector<std::vector<GapSpec>> perCtg(K);
    for (int g = 0; g < o.gaps; g++) {
        int len = o.gapList.empty() ? o.gapMin + (int)rng.below((uint64_t)(o.gapMax - o.gapMin + 1)) : o.gapList[g];
        int c = g % K;
        long tl = (long)truth[c].size();
        bool placed = false;
        for (int attempt = 0; attempt < 10000 && !placed; attempt++) {
            if (tl < 2L * o.margin + len + 1) break;
            long st = o.margin + (long)rng.below((uint64_t)(tl - 2L * o.margin - len));
            bool ok = true;
            for (auto& q : perCtg[c])
                if (st < q.tstart + q.tlen + o.spacing && q.tstart < st + len + o.spacing) { ok = false; break; }
            if (!ok) continue;
            GapSpec s;
            s.ctg = c; s.tstart = st; s.tlen = len; s.sstart = 0;
            s.slen = len;
            if (o.lenNoise > 0) {
                s.slen = len - o.lenNoise + (int)rng.below((uint64_t)(2 * o.lenNoise + 1));
                if (s.slen < 1) s.slen = 1;
            }
            perCtg[c].push_back(s);
            placed = true;
        }
        if (!placed) { fprintf(stderr, "could not place gap %d (len %d)\n", g, len); return 1; }
    }

    // ---- scaffolds (truth with gaps replaced by N-runs) + coordinate shifts ------------------
    std::vector<std::string> scaf(K);
    std::vector<GapSpec> gaps;   // in gapInfo order (scaffold-major, position-minor), as main.cpp:938-972 finds them
    for (int c = 0; c < K; c++) {
        auto& v = perCtg[c];
        std::sort(v.begin(), v.end(), [](const GapSpec& a, const GapSpec& b) { return a.tstart < b.tstart; });
        long prev = 0, shift = 0;
        for (auto& g : v) {
            scaf[c].append(truth[c], prev, g.tstart - prev);
            g.sstart = g.tstart + shift;
            scaf[c].append((size_t)g.slen, 'N');
            shift += g.slen - g.tlen;
            prev = g.tstart + g.tlen;
            gaps.push_back(g);
        }
        scaf[c].append(truth[c], prev, std::string::npos);
    }
    {
        FILE* f = fopen(path("scaffolds.fa").c_str(), "w");
        if (!f) { perror("scaffolds.fa"); return 1; }
        for (int c = 0; c < K; c++) {
            fprintf(f, ">scaf%d\n", c);
            for (size_t i = 0; i < scaf[c].size(); i += 60) {
                size_t n = std::min<size_t>(60, scaf[c].size() - i);
                fwrite(scaf[c].data() + i, 1, n, f);
                fputc('\n', f);
            }
        }
        fclose(f);
        f = fopen(path("truth.txt").c_str(), "w");
        for (size_t g = 0; g < gaps.size(); g++)
            fprintf(f, "%zu\t%d\t%ld\t%d\t%d\t%s\n", g, gaps[g].ctg, gaps[g].sstart, gaps[g].slen, gaps[g].tlen,
                    truth[gaps[g].ctg].substr(gaps[g].tstart, gaps[g].tlen).c_str());
        fclose(f);
    }

    // ---- output streams ---------------------------------------------------------------------
    FILE* fsam = nullptr;
    if (o.sam) {
        fsam = fopen(path("result.sam").c_str(), "w");
        fprintf(fsam, "@HD\tVN:1.0\tSO:unsorted\n");
        for (int c = 0; c < K; c++) fprintf(fsam, "@SQ\tSN:scaf%d\tLN:%zu\n", c, scaf[c].size());
        fprintf(fsam, "@PG\tID:fb_synth\tPN:fb_synth\n");
    }
    FILE* fmy = nullptr;
    std::vector<FILE*> fgap;
    if (o.direct) {
        fmy = fopen(path("myout.sam").c_str(), "w");
        mkdir(path("Gaps").c_str(), 0755);
        FILE* gi = fopen(path("gapInfo.txt").c_str(), "w");
        for (auto& g : gaps) fprintf(gi, "%d\t%ld\t%d\n", g.ctg, g.sstart, g.slen);
        fclose(gi);
        fgap.resize(gaps.size(), nullptr);
    }
    // Gap files are buffered in memory and written at the end so that thousands of gaps do not
    // need thousands of open descriptors.
    std::vector<std::string> gapText(o.direct ? gaps.size() : 0);

    // first gap index per scaffold, for the splitter's linear search order (main.cpp:197-219)
    std::vector<int> firstGap(K + 1, 0);
    {
        int idx = 0;
        for (int c = 0; c < K; c++) { firstGap[c] = idx; idx += (int)perCtg[c].size(); }
        firstGap[K] = idx;
    }
    const int D = 200;  // maxDistance, main.cpp:80

    const int L = o.L;
    std::vector<char> fwd1(L + 1), fwd2(L + 1), tmp(L + 1), ref(L + 1);
    char md1[1024], md2[1024];
    long properWritten = 0, totalGroups = 0, unGroups = 0;
    long pairId = 0;
    const std::string qual(L, 'I');

    auto mutate = [&](const char* src, char* dst) {
        for (int i = 0; i < L; i++) {
            char b = src[i];
            if (o.err > 0 && rng.uni() < o.err) {
                int k = (int)(rng.next() % 3);
                for (int t = 0, seen = 0; t < 4; t++) {
                    if (BASES[t] == b) continue;
                    if (seen++ == k) { b = BASES[t]; break; }
                }
            }
            if (o.nrate > 0 && rng.uni() < o.nrate) b = 'N';
            dst[i] = b;
        }
        dst[L] = 0;
    };

    for (int c = 0; c < K; c++) {
        const std::string& T = truth[c];
        const long tl = (long)T.size();
        const auto& cg = perCtg[c];
        long npairs = (long)(o.cov * tl / (2.0 * L));
        // truth -> scaffold coordinate, or -1 when [p, p+L) touches a gap (binary search over sorted gaps)
        std::vector<long> gEnd(cg.size()), shiftAfter(cg.size() + 1, 0);
        for (size_t k = 0; k < cg.size(); k++) {
            gEnd[k] = cg[k].tstart + cg[k].tlen;
            shiftAfter[k + 1] = shiftAfter[k] + cg[k].slen - cg[k].tlen;
        }
        auto mapPos = [&](long p) -> long {
            size_t k = std::upper_bound(gEnd.begin(), gEnd.end(), p) - gEnd.begin();   // first gap ending after p
            if (k < cg.size() && p + L > cg[k].tstart) return -1;
            return p + shiftAfter[k];
        };
        for (long n = 0; n < npairs; n++, pairId++) {
            long ins = (long)std::llround(o.insMean + o.insSd * rng.normal());
            if (ins < L) ins = L;
            if (ins > tl) ins = tl;
            long s = (long)rng.below((uint64_t)(tl - ins + 1));
            bool flip = rng.next() & 1;   // flip: mate1 is the right-hand (reverse-strand) read
            long pl = s, pr = s + ins - L;
            long ml = mapPos(pl), mr = mapPos(pr);
            bool proper = ml >= 0 && mr >= 0;
            if (proper && o.modelPairs >= 0 && properWritten >= o.modelPairs) continue;
            // reads in forward-strand orientation with sequencing errors
            mutate(T.data() + pl, fwd1.data());   // left read
            mutate(T.data() + pr, fwd2.data());   // right read
            char q[32];
            snprintf(q, sizeof q, "r%ld", pairId);
            if (proper) {
                int nm1 = makeMd(T.data() + pl, fwd1.data(), L, md1);
                int nm2 = makeMd(T.data() + pr, fwd2.data(), L, md2);
                long tlen = mr + L - ml;
                int fl = flip ? 163 : 99, fr = flip ? 83 : 147;
                // mate1 first in the file
                struct Line { int flag; long pos, pnext, tl; const char* seq; const char* md; int nm; };
                Line left{fl, ml + 1, mr + 1, tlen, fwd1.data(), md1, nm1};
                Line right{fr, mr + 1, ml + 1, -tlen, fwd2.data(), md2, nm2};
                Line first = flip ? right : left, second = flip ? left : right;
                if (fsam) {
                    for (const Line* ln : {&first, &second})
                        fprintf(fsam, "%s\t%d\tscaf%d\t%ld\t42\t%dM\t=\t%ld\t%ld\t%s\t%s\tAS:i:0\tXN:i:0\tNM:i:%d\tMD:Z:%s\tYS:i:0\tYT:Z:CP\n",
                                q, ln->flag, c, ln->pos, L, ln->pnext, ln->tl, ln->seq, qual.c_str(), ln->nm, ln->md);
                }
                if (fmy) {
                    for (const Line* ln : {&first, &second})
                        fprintf(fmy, "%s\t%d\t%d\t%ld\t%dM\t%ld\t%s\tMD:Z:%s\tIH:i:1\n", q, ln->flag, c, ln->pos, L, ln->tl,
                                ln->seq, ln->md);
                }
                properWritten++;
                totalGroups++;
                continue;
            }
            if (ml < 0 && mr < 0) {
                // both unmapped (77/141): the splitter drops these (main.cpp:627-639) but still counts them
                if (fsam) {
                    revcomp(fwd2.data(), L, tmp.data());
                    const char* s1 = flip ? tmp.data() : fwd1.data();
                    std::string s2;
                    if (flip) s2 = fwd1.data();
                    else s2 = tmp.data();
                    fprintf(fsam, "%s\t77\t*\t0\t0\t*\t*\t0\t0\t%s\t%s\tYT:Z:UP\n", q, s1, qual.c_str());
                    fprintf(fsam, "%s\t141\t*\t0\t0\t*\t*\t0\t0\t%s\t%s\tYT:Z:UP\n", q, s2.c_str(), qual.c_str());
                }
                totalGroups++;
                unGroups++;
                continue;
            }
            // exactly one mate maps: the anchor
            bool anchorLeft = ml >= 0;
            long apos = (anchorLeft ? ml : mr) + 1;
            const char* aseq = anchorLeft ? fwd1.data() : fwd2.data();
            int anm = makeMd(T.data() + (anchorLeft ? pl : pr), aseq, L, md1);
            int astrand = anchorLeft ? 0 : 1;
            bool anchorIsMate1 = anchorLeft ? !flip : flip;
            int aflag = 1 + 8 + (astrand ? 16 : 0) + (anchorIsMate1 ? 64 : 128);
            int uflag = 1 + 4 + (astrand ? 32 : 0) + (anchorIsMate1 ? 128 : 64);
            // the unmapped mate in FASTQ orientation: right-hand reads are sequenced from the reverse strand
            std::string useq;
            if (anchorLeft) { revcomp(fwd2.data(), L, tmp.data()); useq = tmp.data(); }
            else useq = fwd1.data();
            if (fsam) {
                auto emitAnchor = [&]() {
                    for (int k = 0; k < o.multihit; k++)
                        fprintf(fsam, "%s\t%d\tscaf%d\t%ld\t42\t%dM\t=\t%ld\t0\t%s\t%s\tAS:i:0\tXN:i:0\tNM:i:%d\tMD:Z:%s\tYT:Z:UP\n", q,
                                aflag, c, apos, L, apos, aseq, qual.c_str(), anm, md1);
                };
                auto emitMate = [&]() {
                    fprintf(fsam, "%s\t%d\tscaf%d\t%ld\t0\t*\t=\t%ld\t0\t%s\t%s\tYT:Z:UP\n", q, uflag, c, apos, apos, useq.c_str(),
                            qual.c_str());
                };
                if (anchorIsMate1) { emitAnchor(); emitMate(); }
                else { emitMate(); emitAnchor(); }
            }
            totalGroups++;
            unGroups++;
            if (o.direct) {
                // splitter's gap lookup: first gap (in file order) whose window holds the anchor (main.cpp:197-219)
                int hit = -1;
                for (int gi = firstGap[c]; gi < firstGap[c + 1]; gi++) {
                    const GapSpec& g = gaps[gi];
                    long gs = g.sstart, ge = g.sstart + g.slen;
                    if ((astrand == 0 && apos > gs - D && apos < ge) || (astrand == 1 && apos > ge && apos < ge + D)) { hit = gi; break; }
                }
                if (hit >= 0) {
                    char buf[1024];
                    for (int k = 0; k < o.multihit; k++) {
                        int n1 = snprintf(buf, sizeof buf, "%s\t%d\t%d\t%ld\t%dM\t0\t%s\tMD:Z:%s\tIH:i:0\n", q, aflag, c, apos, L, aseq, md1);
                        gapText[hit].append(buf, n1);
                        int n2 = snprintf(buf, sizeof buf, "%s\t%d\t%d\t%ld\t*\t0\t%s\t*\tIH:i:0\n", q, uflag, c, apos, useq.c_str());
                        gapText[hit].append(buf, n2);
                    }
                }
            }
        }
    }
    if (fsam) fclose(fsam);
    if (o.direct) {
        fclose(fmy);
        for (size_t g = 0; g < gaps.size(); g++) {
            char name[64];
            snprintf(name, sizeof name, "Gaps/gaps_%zu.sam", g);
            FILE* f = fopen(path(name).c_str(), "w");
            fwrite(gapText[g].data(), 1, gapText[g].size(), f);
            fclose(f);
        }
        // totalCount counts one extra flush of the (empty) proper-pair buffer, main.cpp:1137-1144, 415
        FILE* st = fopen(path("stat.txt").c_str(), "w");
        fprintf(st, "%ld %ld %d %d", properWritten + 1 + unGroups, unGroups, L, 5000);
        fclose(st);
    }
    fprintf(stderr, "fb_synth: %d scaffolds, %zu gaps, %ld pairs (%ld proper written, %ld with an unmapped mate)\n", K, gaps.size(),
            pairId, properWritten, unGroups);
    return 0;
}
