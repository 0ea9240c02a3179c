/* figbird_oracle.c — CPU restatement of the figbird EM hot path.  TEST INFRASTRUCTURE ONLY
 * (see figbird_oracle.h).  Plain C, scalar, one thread; arithmetic kept in the reference's
 * evaluation order so results are bit-identical to /root/reference/figbird.cpp built without FMA
 * contraction (compile with -ffp-contract=off).
 *
 * Layout differs from the reference (flat row-major arrays instead of pointer-per-row), the
 * arithmetic does not.
 */
#include "figbird_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAX_REC_LEN 1024 /* figbird.cpp:22 */

/* charCodes table, figbird.cpp:2753-2776: only upper-case ACGT are bases. */
static inline int base_code(char ch) {
    switch (ch) {
        case 'A': return 0;
        case 'C': return 1;
        case 'G': return 2;
        case 'T': return 3;
        default: return 4;
    }
}

/* reverse(), figbird.cpp:1382-1403: reverse complement, lower case accepted, anything else -> 'N'. */
static void revcomp_into(char* dst, const char* src, int n) {
    for (int i = 0; i < n; i++) {
        char ch = src[i], o;
        if (ch == 'A' || ch == 'a') o = 'T';
        else if (ch == 'C' || ch == 'c') o = 'G';
        else if (ch == 'G' || ch == 'g') o = 'C';
        else if (ch == 'T' || ch == 't') o = 'A';
        else o = 'N';
        dst[n - 1 - i] = o;
    }
    dst[n] = '\0';
}

void orc_gap_range(int orig_gap, int* gap_min, int* gap_max) {
    /* figbird.cpp:2438-2453 — the fractions are float and int*float is evaluated in float. */
    float lo_frac, hi_frac;
    if (orig_gap > 300) lo_frac = .95;
    else if (orig_gap < 100) lo_frac = 0.1;
    else lo_frac = 0.85;
    if (orig_gap > 200) hi_frac = 1.05;
    else hi_frac = 1.5;
    *gap_min = orig_gap * lo_frac < 10 ? 1 : orig_gap * lo_frac;
    *gap_max = orig_gap * hi_frac;
}

typedef struct {
    const orc_model* m;
    const char* contig;
    long gap_start;
    int orig_gap;
    int D;
    long start_pos, end_pos; /* figbird.cpp:1532 */
    int gap_len;             /* this->gapLength */
    double* counts;          /* [max_rows][5] */
    double* probs;
    double* eprobs;
    int nrec;
    const orc_record* recs;
    char** seq;    /* mate sequence as scored (reverse-complemented when the anchor is forward) */
    int* len2;
    char* consensus;
    int shards;       /* record-sharded restatement (see orc_fill_gap_sharded); 0 or 1: the reference's single sum */
    double* part;     /* [max_rows][5] one shard's partial counts */
    double* total;    /* [max_rows][5] rank-ordered sum of the partial counts */
} filler;

/* computeProbsGap, figbird.cpp:1568-1587.  0/0 -> NaN is kept. */
static void update_probs(filler* g) {
    long rows = g->end_pos - g->start_pos;
    for (long i = 0; i < rows; i++) {
        double* c = g->counts + 5 * i;
        double total = 0;
        for (int j = 0; j < 5; j++) total = total + c[j];
        double ncount = c[4];
        for (int j = 0; j < 4; j++) g->probs[5 * i + j] = ((c[j] + ncount / 4) / total);
        g->probs[5 * i + 4] = 0;
    }
}

/* computeErrorProbsGap, figbird.cpp:1589-1608. */
static void update_error_probs(filler* g) {
    long rows = g->end_pos - g->start_pos;
    for (long i = 0; i < rows; i++) {
        for (int j = 0; j < 5; j++) {
            double sum = 0;
            for (int k = 0; k < 4; k++) {
                if (j == k) continue;
                sum += g->probs[5 * i + k] * (g->m->error_type[k][j]);
            }
            g->eprobs[5 * i + j] = sum;
        }
    }
}

/* initialize, figbird.cpp:1620-1675. */
static void init_profile(filler* g, int gp) {
    const int D = g->D;
    g->gap_len = gp;
    g->start_pos = (g->gap_start - D) < 0 ? 0 : (g->gap_start - D);
    g->end_pos = g->gap_start + gp + D;
    long rows = g->end_pos - g->start_pos;
    memset(g->counts, 0, sizeof(double) * 5 * rows);
    memset(g->probs, 0, sizeof(double) * 5 * rows);
    memset(g->eprobs, 0, sizeof(double) * 5 * rows);
    for (long i = 0; i < D; i++) g->counts[5 * i + base_code(g->contig[i + g->start_pos])]++;
    for (long i = D; i < D + gp; i++) g->counts[5 * i + 4]++;
    for (long i = D + gp; i < D + gp + D; i++) g->counts[5 * i + base_code(g->contig[i + g->start_pos - gp + g->orig_gap])]++;
    update_probs(g);
    update_error_probs(g);
}

/* placeReads, figbird.cpp:1677-1965 (parsing hoisted out: records are already in memory). */
static double place_reads(filler* g) {
    const orc_model* m = g->m;
    const int D = g->D, gp = g->gap_len, f = m->f_mode;
    const long gs = g->gap_start, sp = g->start_pos;
    long rows = g->end_pos - g->start_pos;
    for (long i = D; i < rows - D; i++)
        for (int j = 0; j < 5; j++) g->counts[5 * i + j] = 0;

    double lik = 0;
    const int shards = g->shards > 1 ? g->shards : 1;
    double* const counts_home = g->counts;
    for (int sh = 0; sh < shards; sh++) {
    /* shard sh takes the records [sh*R/S, (sh+1)*R/S) and accumulates them, in the reference's order, into its own
     * zeroed array; the shards' arrays are then added in shard order (what the ranks of the GPU path do) */
    const int r_lo = (int)((long)sh * g->nrec / shards), r_hi = (int)((long)(sh + 1) * g->nrec / shards);
    if (shards > 1) {
        for (long i = D; i < rows - D; i++)
            for (int j = 0; j < 5; j++) g->part[5 * i + j] = 0;
        g->counts = g->part;
    }
    for (int r = r_lo; r < r_hi; r++) {
        const orc_record* rc = &g->recs[r];
        const char* s2 = g->seq[r];
        const int L2 = g->len2[r];
        const int rev = rc->strand1 == 0; /* isReverse, figbird.cpp:1803-1812 */
        double best = (f == 1) ? -DBL_MAX : 0;
        const int left = rc->pos1 < gs;
        for (long i = gs - L2 + 1; i < gs + gp; i++) {
            int t;
            if (left) t = (int)(gs - rc->pos1 + L2) + (int)(i - gs);                      /* figbird.cpp:1824-1828 */
            else t = (int)(rc->pos1 - (gs + gp) + rc->len1) + (int)(gs + gp - i);          /* figbird.cpp:1887-1891 */
            if (t < m->thr_min || t > m->thr_max) continue;
            double s = (f == 1) ? log(m->ild[t]) : m->ild[t];
            for (int j = 0; j < L2; j++) {
                int c = base_code(s2[j]);
                long x = i + j - sp;
                int k = rev ? L2 - 1 - j : j;
                if (c < 4) {
                    double aa = (g->probs[5 * x + c] * (1 - m->error_pos[k]) + m->error_pos[k] * g->eprobs[5 * x + c]); /* :1855 */
                    if (f == 1) s += log(aa);
                    else s *= aa;
                } else {
                    if (f == 1) s += log(m->error_pos[k] * g->eprobs[5 * x + 4]);
                    else s *= (m->error_pos[k] * g->eprobs[5 * x + 4]);                                                 /* :1862 */
                }
            }
            if (s > best) best = s;
            for (int j = 0; j < L2; j++) {
                long x = i - sp + j;
                if (x >= D && x < D + gp) {
                    if (f == 1) g->counts[5 * x + base_code(s2[j])] += pow(10, s); /* :1878 */
                    else g->counts[5 * x + base_code(s2[j])] += s;                 /* :1879 */
                }
            }
        }
        if (f == 1) lik += best;
        else if (best > 0) lik += log(best); /* :1951-1955 */
    }
    if (shards > 1) {
        g->counts = counts_home;
        for (long i = D; i < rows - D; i++)
            for (int j = 0; j < 5; j++) {
                if (sh == 0) g->total[5 * i + j] = g->part[5 * i + j];
                else g->total[5 * i + j] = g->total[5 * i + j] + g->part[5 * i + j];
            }
    }
    }
    if (shards > 1)
        for (long i = D; i < rows - D; i++)
            for (int j = 0; j < 5; j++) g->counts[5 * i + j] = g->total[5 * i + j];
    return lik;
}

/* computeSequence, figbird.cpp:1984-2047.  checkDuplicate (1967-1982) only ever inspects arr[0]
 * once (sizeof(pointer)/sizeof(double) == 1) and so returns 0 whenever 0 <= (int)arr[0] < 10000;
 * outside that range the reference indexes off its stack array — reported as an error here. */
static int consensus_of(filler* g, int check) {
    const int D = g->D;
    long rows = g->end_pos - g->start_pos;
    for (long i = D; i < rows - D; i++) {
        double mx = 0;
        int idx = 0, seen = 0;
        for (int j = 0; j <= 4; j++) {
            if (g->counts[5 * i + j] > mx) {
                mx = g->counts[5 * i + j];
                idx = j;
                seen = 1;
            }
        }
        if (check == 1) {
            double a0 = g->counts[5 * i];
            if (!(a0 >= 0 && a0 < 10000)) return -3;
        }
        g->consensus[i - D] = (seen || !check) ? "ACGTN"[idx] : 'N';
    }
    g->consensus[rows - 2 * D] = '\0';
    return 0;
}

/* finalize, figbird.cpp:2087-2393. */
static int hard_assign(filler* g, int gb, const char* best, int* total, int* discarded) {
    const orc_model* m = g->m;
    const int D = g->D;
    const long gs = g->gap_start;
    int max_size = gb + 2 * D;
    char* gstr = (char*)malloc((size_t)max_size + 1);
    for (long i = 0; i < D; i++) gstr[i] = g->contig[i + g->start_pos];
    for (long i = D; i < D + gb; i++) gstr[i] = best[i - D];
    for (long i = D + gb; i < D + gb + D; i++) gstr[i] = g->contig[i + g->start_pos - gb + g->orig_gap];
    gstr[max_size] = '\0';

    long rows = g->end_pos - g->start_pos; /* still the last scanned estimate's extent (:2140) */
    for (long i = D; i < rows - D; i++)
        for (int j = 0; j < 5; j++) g->counts[5 * i + j] = 0;

    g->start_pos = gs - D;                 /* :2149, no clamp */
    g->end_pos = gs + g->gap_len + D;      /* :2150, gapLength is still the last scanned estimate */
    g->gap_len = gb;                       /* :2152 */
    const long sp = g->start_pos;

    int tot = 0, disc = 0;
    for (int r = 0; r < g->nrec; r++) {
        const orc_record* rc = &g->recs[r];
        const char* s2 = g->seq[r];
        const int L2 = g->len2[r];
        const int rev = rc->strand1 == 0; /* strandNo2 = 1, :2222-2231 */
        double best_p = 0;
        long best_pos = 0;
        tot++;
        const int left = rc->pos1 < gs;
        for (long i = gs - L2 + 1; i < gs + gb; i++) {
            int t;
            if (left) t = (int)(gs - rc->pos1 + L2) + (int)(i - gs);
            else t = (int)(rc->pos1 - (gs + gb) + rc->len1) + (int)(gs + gb - i);
            if (t < m->thr_min || t > m->thr_max) continue;
            double s = m->ild[t];
            for (int j = 0; j < L2; j++) {
                int to = base_code(s2[j]);
                int from = base_code(gstr[i + j - sp]);
                int k = rev ? L2 - j - 1 : j;
                if (from == to) s *= (1 - m->error_pos[k] - m->in_pos[k] - m->del_pos[k]); /* :2272 */
                else s *= m->error_pos[k] * m->error_type[from][to];                       /* :2276 */
            }
            if (s > best_p) {
                best_p = s;
                best_pos = i - sp;
            }
        }
        if (-log(best_p) <= m->gap_prob_cutoff) { /* :2293 */
            for (int j = 0; j < L2; j++)
                if (best_pos + j >= D && best_pos + j < D + gb) g->counts[5 * (best_pos + j) + base_code(s2[j])] += 1;
        } else {
            disc++;
        }
    }
    free(gstr);
    *total = tot;
    *discarded = disc;
    return consensus_of(g, 1);
}

int orc_fill_gap(const orc_model* m, const char* contig, long contig_len, long gap_start, int orig_gap,
                 const orc_record* recs, int nrec, orc_result* out) {
    return orc_fill_gap_sharded(m, contig, contig_len, gap_start, orig_gap, recs, nrec, 1, out);
}

int orc_fill_gap_sharded(const orc_model* m, const char* contig, long contig_len, long gap_start, int orig_gap,
                         const orc_record* recs, int nrec, int shards, orc_result* out) {
    return orc_fill_gap_subset(m, contig, contig_len, gap_start, orig_gap, recs, nrec, shards, NULL, out);
}

int orc_fill_gap_subset(const orc_model* m, const char* contig, long contig_len, long gap_start, int orig_gap,
                        const orc_record* recs, int nrec, int shards, const unsigned char* scan, orc_result* out) {
    const int D = m->max_distance;
    /* Outside these bounds the reference reads past its arrays; refuse instead. */
    if (orig_gap < 1 || gap_start < D || gap_start + orig_gap + D > contig_len) return -2;

    filler g;
    memset(&g, 0, sizeof g);
    g.m = m;
    g.contig = contig;
    g.gap_start = gap_start;
    g.orig_gap = orig_gap;
    g.D = D;
    g.nrec = nrec;
    g.recs = recs;

    int gmin, gmax;
    orc_gap_range(orig_gap, &gmin, &gmax);
    int cap = (gmax > 2 * orig_gap ? gmax : 2 * orig_gap);
    long max_rows = (long)cap + 2 * D + 1;
    g.counts = (double*)calloc((size_t)max_rows * 5, sizeof(double));
    g.probs = (double*)calloc((size_t)max_rows * 5, sizeof(double));
    g.eprobs = (double*)calloc((size_t)max_rows * 5, sizeof(double));
    g.shards = shards;
    g.part = (double*)calloc((size_t)max_rows * 5, sizeof(double));
    g.total = (double*)calloc((size_t)max_rows * 5, sizeof(double));
    g.consensus = (char*)calloc((size_t)cap + 2, 1);
    char* best = (char*)calloc((size_t)cap + 2, 1);
    g.seq = (char**)calloc((size_t)(nrec > 0 ? nrec : 1), sizeof(char*));
    g.len2 = (int*)calloc((size_t)(nrec > 0 ? nrec : 1), sizeof(int));
    int rc = 0;
    for (int r = 0; r < nrec; r++) {
        int n = (int)strlen(recs[r].seq2);
        if (n > m->max_read_len || n > D || n < 1) rc = -2;
        g.len2[r] = n;
        g.seq[r] = (char*)malloc((size_t)n + 1);
        if (recs[r].strand1 == 0) revcomp_into(g.seq[r], recs[r].seq2, n);
        else memcpy(g.seq[r], recs[r].seq2, (size_t)n + 1);
    }

    if (rc == 0) {
        /* fillGap, figbird.cpp:2454-2510 */
        int best_est = gmin;
        double best_lik = -DBL_MAX, lik = 0;
        for (int gp = gmin; gp <= gmax; gp++) {
            if (scan && !scan[gp - gmin]) {      /* test aid: estimate left out of the scan */
                if (out->lik_per_estimate) out->lik_per_estimate[gp - gmin] = NAN;
                continue;
            }
            init_profile(&g, gp);
            for (int it = 0; it < m->num_iter; it++) {
                lik = place_reads(&g);
                update_probs(&g);
                update_error_probs(&g);
            }
            if (out->lik_per_estimate) out->lik_per_estimate[gp - gmin] = lik;
            if (out->dump_estimate == gp) {
                size_t n = (size_t)(gp + 2 * D) * 5;
                if (out->probs_dump) memcpy(out->probs_dump, g.probs, n * sizeof(double));
                if (out->counts_dump) memcpy(out->counts_dump, g.counts, n * sizeof(double));
            }
            if (lik >= best_lik) { /* last estimate wins ties, :2479 */
                best_lik = lik;
                best_est = gp;
                consensus_of(&g, 0);
                strcpy(best, g.consensus);
            }
        }
        if (gmax < gmin) rc = -2;
        if (rc == 0 && scan) {
            /* finalize sees the extent of the LAST estimate of the full scan (gapMax), whatever subset was scanned */
            g.gap_len = gmax;
            g.start_pos = (g.gap_start - D) < 0 ? 0 : (g.gap_start - D);
            g.end_pos = g.gap_start + gmax + D;
        }
        if (rc == 0) {
            if (out->best_string) strcpy(out->best_string, best);
            rc = hard_assign(&g, best_est, best, &out->total_reads, &out->discarded);
            out->gap_min = gmin;
            out->gap_max = gmax;
            out->best_estimate = best_est;
            out->best_lik = best_lik;
            out->filled_len = best_est; /* getConcensus, :2050-2063 */
            memcpy(out->filled, g.consensus, (size_t)best_est);
            out->filled[best_est] = '\0';
        }
    }
    for (int r = 0; r < nrec; r++) free(g.seq[r]);
    free(g.seq);
    free(g.len2);
    free(g.counts);
    free(g.probs);
    free(g.eprobs);
    free(g.part);
    free(g.total);
    free(g.consensus);
    free(best);
    return rc;
}

long orc_count_likelihoods(const orc_model* m, long gap_start, int gp, const orc_record* recs, int nrec) {
    long n = 0;
    for (int r = 0; r < nrec; r++) {
        int L2 = (int)strlen(recs[r].seq2);
        for (long i = gap_start - L2 + 1; i < gap_start + gp; i++) {
            long t = recs[r].pos1 < gap_start ? i - recs[r].pos1 + L2 : (long)recs[r].pos1 + recs[r].len1 - i;
            if (t >= m->thr_min && t <= m->thr_max) n++;
        }
    }
    return n;
}

int orc_read_gap_file(const char* path, orc_record** recs_out) {
    FILE* fp = fopen(path, "r");
    *recs_out = NULL;
    if (!fp) return -1;
    int n = 0, cap = 64;
    orc_record* v = (orc_record*)malloc(sizeof(orc_record) * cap);
    char l1[ORC_MAX_REC_LEN], l2[ORC_MAX_REC_LEN];
    while (fgets(l1, ORC_MAX_REC_LEN, fp)) {
        if (l1[0] == '@') continue;
        if (!fgets(l2, ORC_MAX_REC_LEN, fp)) break;
        /* columns: qname flag rname pos cigar tlen seq ... (writeSam, main.cpp:180) */
        char* tok = strtok(l1, "\t"); (void)tok;
        int flag1 = atoi(strtok(NULL, "\t"));
        strtok(NULL, "\t");
        int pos1 = atoi(strtok(NULL, "\t"));
        strtok(NULL, "\t");
        strtok(NULL, "\t");
        char* seq1 = strtok(NULL, "\t");
        int len1 = (int)strlen(seq1);
        strtok(l2, "\t");
        for (int k = 0; k < 5; k++) strtok(NULL, "\t");
        char* seq2 = strtok(NULL, "\t");
        if (n == cap) {
            cap *= 2;
            v = (orc_record*)realloc(v, sizeof(orc_record) * cap);
        }
        v[n].pos1 = pos1;
        v[n].len1 = len1;
        v[n].strand1 = (flag1 & 16) >> 4;
        v[n].seq2 = strdup(seq2);
        n++;
    }
    fclose(fp);
    *recs_out = v;
    return n;
}

void orc_free_records(orc_record* recs, int n) {
    if (!recs) return;
    for (int i = 0; i < n; i++) free((void*)recs[i].seq2);
    free(recs);
}
