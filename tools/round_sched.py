"""List-scheduling model of the EM kernel's rounds (host-only arithmetic; nothing runs on a GPU, nothing under oracle/
is used): how long a round's task queue keeps 16 warps busy before the barrier, for the task shapes that were measured
in profiles/r2_tail_pack_experiment.txt.

    python tools/round_sched.py <workdir with fb_model.bin, gapInfo.txt, Gaps/> [gaps to sample, default 40]

Model: a unit (gap, estimate) runs rounds r = -1 .. nchunks; round r holds the E tasks of chunk r (16 records) and the M
tasks of chunk r-1; warps take tasks from one queue in the given order; the round ends when the last warp is done.
Task costs are warp instructions, from the SASS of the kernels (per trip of 8 bases: two chains 134, one chain 92, tail
task 87, pair task 162; M task: 4.25 per list entry + 12 per record).  The model knows nothing about shared-memory
bandwidth or the second CTA of the SM: it only says what the queue's shape does to the length of a round.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gapprober_b200.figbird import FigbirdModel, gap_range, read_gap_file, read_gap_info  # noqa: E402

C2, C1, CT, CP = 134 / 8, 92 / 8, 87 / 8, 162 / 8      # instructions per base
NW, CH = 16, 16


def makespan(costs):
    free = [0.0] * NW
    for c in costs:
        i = min(range(NW), key=free.__getitem__)
        free[i] += c
    return max(free), sum(costs)


def unit_tasks(win, D, gp, has_pre, mode):
    """Per chunk: (E task costs in queue order, extra tasks).  mode: 'base', 'tails', 'tails+pairs'."""
    out = []
    for c0 in range(0, len(win), CH):
        recs = win[c0:c0 + CH]
        geo = []
        for lo0, hi0, L2, rev in recs:
            lo, hi = max(lo0, D - L2 + 1), min(hi0, D + gp - 1)
            geo.append((lo, hi - lo + 1, L2, rev))
        packed = [False] * len(geo)
        tails = []
        if mode != "base":
            tasks = []
            for i, (lo, W, L2, rev) in enumerate(geo):
                nt = W & 31 if W > 0 else 0
                if not nt or (has_pre and D - (lo + (W & ~31) + 31) >= 4):
                    continue
                for t in tasks:
                    if t[0] == (L2, rev) and t[1] + nt <= 32:
                        t[1] += nt; t[2].append(i); break
                else:
                    tasks.append([(L2, rev), nt, [i]])
            for t in tasks:
                if len(t[2]) >= 2:
                    tails.append(CT * t[0][0])
                    for i in t[2]:
                        packed[i] = True
        nch = [((W >> 5) if packed[i] else ((W + 31) >> 5)) if W > 0 else 0 for i, (lo, W, L2, rev) in enumerate(geo)]
        paired = [False] * len(geo)
        pairs = []
        if mode == "tails+pairs":
            open_ = {}
            for i, (lo, W, L2, rev) in enumerate(geo):
                if nch[i] & 1 and not (has_pre and D - (lo + 32 * (nch[i] - 1) + 31) >= 4):
                    k = (L2, rev)
                    if k in open_:
                        j = open_.pop(k); paired[i] = paired[j] = True; pairs.append(CP * L2)
                    else:
                        open_[k] = i
        nblk = max(1, max((n + 1) // 2 for n in nch) if nch else 1)
        e = []
        for blk in range(nblk):                 # block-major, as the kernel hands them out
            for i, (lo, W, L2, rev) in enumerate(geo):
                n = nch[i] - (1 if paired[i] else 0)
                left = n - 2 * blk
                if left <= 0:
                    e.append(0.3)                                   # zero fill only
                    continue
                steps = L2
                if has_pre:                                          # left-flank prefix of chain A of this block
                    js = D - (lo + 64 * blk + 31)
                    if js >= 4:
                        steps -= min(L2, js & ~3)
                e.append((C2 if left >= 2 else C1) * max(steps, 0) + 6)
        out.append((e, pairs, tails))
    return out


def m_tasks(win, seqs, D, gp):
    """Per chunk: costs of the (32-row block, class) tasks."""
    out = []
    nrb = (gp + 31) // 32
    for c0 in range(0, len(win), CH):
        costs = []
        for rb in range(nrb):
            x0 = D + rb * 32
            per_class = np.zeros(5)
            nrec = 0
            for (lo0, hi0, L2, rev), cls in zip(win[c0:c0 + CH], seqs[c0:c0 + CH]):
                lo, hi = max(lo0, D - L2 + 1), min(hi0, D + gp - 1)
                jhi, jlo = min(L2 - 1, x0 + 31 - lo), max(0, x0 - hi)
                nrec += 1
                if hi >= lo and jhi >= jlo:
                    per_class += np.bincount(cls[jlo:jhi + 1], minlength=5)[:5]
            costs += [4.25 * n + 12 * nrec / 5 + 20 for n in per_class]
        out.append(costs)
    return out


def main(wd, max_gaps=40):
    m = FigbirdModel.load(os.path.join(wd, "fb_model.bin"))
    D = m.max_distance
    gi = read_gap_info(os.path.join(wd, "gapInfo.txt"))
    idx = np.random.default_rng(1).permutation(len(gi))[:max_gaps]
    code = np.full(256, 4, dtype=np.int64)
    for k, ch in enumerate(b"ACGT"):
        code[ch] = k
    variants = [("base, M tasks first", "base", "ME"), ("base, M tasks last", "base", "EM"),
                ("tails, M first, tails first", "tails", "MTE"), ("tails, M last, tails first", "tails", "TEM"),
                ("tails, M first, tails after E", "tails", "MET"), ("tails, M last, tails after E", "tails", "ETM"),
                ("tails + pairs, E, pairs, tails, M", "tails+pairs", "ETM"), ("tails + pairs, pairs and tails first", "tails+pairs", "TEM")]
    span = {v[0]: 0.0 for v in variants}
    work = {v[0]: 0.0 for v in variants}
    for n in idx:
        _, gs, orig = gi[n]
        recs = read_gap_file(os.path.join(wd, "Gaps", f"gaps_{n}.sam"))
        a, b = gap_range(orig)
        sp = gs - D
        win, seqs = [], []
        for pos1, len1, strand1, seq2 in recs:
            L2 = len(seq2)
            rev = strand1 == 0
            if pos1 < gs:
                tb = sp - pos1 + L2
                win.append((m.thr_min - tb, m.thr_max - tb, L2, rev))
            else:
                tb = pos1 + len1 - sp
                win.append((tb - m.thr_max, tb - m.thr_min, L2, rev))
            seqs.append(code[np.frombuffer(seq2, dtype=np.uint8)])     # class counts per position range only: orientation does not matter
        for gp in range(a, b + 1, 5):
            mt = m_tasks(win, seqs, D, gp)
            cache = {}
            for name, mode, order in variants:
                if mode not in cache:
                    cache[mode] = unit_tasks(win, D, gp, True, mode)
                et = cache[mode]
                nchunks = len(et)
                for r in range(0, nchunks + 1):
                    e, pairs, tails = et[r] if r < nchunks else ([], [], [])
                    mm = mt[r - 1] if r >= 1 else []
                    q = []
                    for ch in order:
                        q += {"M": mm, "E": e, "T": pairs + tails}[ch]
                    if q:
                        ms, w = makespan(q)
                        span[name] += ms
                        work[name] += w
    base = span[variants[0][0]]
    print(f"{wd}: {len(idx)} gaps, every fifth estimate; round length = list schedule of the round's queue on {NW} warps")
    for name, _, _ in variants:
        print(f"  {name:42s} sum of round lengths {span[name] / base:6.3f}   work {work[name] / work[variants[0][0]]:6.3f}   warps busy {work[name] / span[name] / NW:5.3f}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
