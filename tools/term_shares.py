"""Count, on a generated working directory, how much of the EM scan's work the bit-exact shortcuts can remove
(host-only arithmetic over the placement windows; nothing is run on a GPU and nothing under oracle/ is used).

    python tools/term_shares.py <workdir with fb_model.bin, gapInfo.txt, Gaps/> [gaps to sample, default all]

Geometry (gapprober_b200/csrc/fb_context.cpp: pack_gap): profile row 0 = scaffold position gapStart - D; placement u =
row of a mate's first base; a record's window is [max(a_lo, D - L2 + 1), min(a_hi, D + gp - 1)] and does not depend on
the estimate gp except through the last clamp; lanes of a warp take 32 consecutive placements (a "chain"), a task is
two chains.  Reported:
  * left-flank terms: a placement u < D multiplies its first D - u terms against left-flank rows (fb_prefix_kernel);
    per lane, and what a chain can skip as a whole (every lane of it has that many);
  * first iteration: chain steps on which all placements of a task are on gap rows (FbEmParams::t0_tab);
  * first iteration shared between the estimates of a gap: placements that do not reach the right flank, at placement
    and at chain granularity; M-step entries of rows only such placements touch.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gapprober_b200.figbird import FigbirdModel, gap_range, read_gap_file, read_gap_info  # noqa: E402


def main(wd, max_gaps=None):
    m = FigbirdModel.load(os.path.join(wd, "fb_model.bin"))
    D = m.max_distance
    gi = read_gap_info(os.path.join(wd, "gapInfo.txt"))
    idx = np.random.default_rng(1).permutation(len(gi))[:max_gaps or len(gi)]
    terms = flank_lane = 0            # E-step terms per iteration; of them against left-flank rows
    chain_terms = chain_skip = 0      # the same at chain granularity (a chain's 32 lanes move together)
    steps = uni = 0                   # chain steps of a task after the prefix skip; of them all-lanes-on-gap-rows
    share_terms = share_chain = 0     # first iteration: placements / whole chains that do not reach the right flank
    m_entries = m_shared = 0
    for n in idx:
        _, gs, orig = gi[n]
        recs = read_gap_file(os.path.join(wd, "Gaps", f"gaps_{n}.sam"))
        a, b = gap_range(orig)
        sp = gs - D
        win = []
        for pos1, len1, _, seq2 in recs:
            L2 = len(seq2)
            if pos1 < gs:
                tb = sp - pos1 + L2
                win.append((m.thr_min - tb, m.thr_max - tb, L2))
            else:
                tb = pos1 + len1 - sp
                win.append((tb - m.thr_max, tb - m.thr_min, L2))
        for gp in range(a, b + 1):
            for lo0, hi0, L2 in win:
                lo, hi = max(lo0, D - L2 + 1), min(hi0, D + gp - 1)
                if hi < lo:
                    continue
                us = np.arange(lo, hi + 1)
                npl = hi - lo + 1
                terms += npl * L2
                flank_lane += int(np.clip(D - us, 0, L2).sum())
                nch = (npl + 31) // 32
                chain_terms += nch * L2
                ns = max(0, min(hi, D + gp - L2) - lo + 1)
                share_terms += ns * L2
                share_chain += (ns // 32 if ns < npl else nch) * L2
                for c in range(nch):
                    js = D - (lo + 32 * c + 31)
                    chain_skip += js & ~3 if js >= 4 else 0
                for blk in range((npl + 63) // 64):
                    u0 = lo + 64 * blk
                    two = u0 + 32 <= hi
                    js_a = D - (u0 + 31)
                    sk_a = js_a & ~3 if js_a >= 4 else 0
                    sk_b = sk_a - 32 if two and js_a >= 36 else 0
                    steps += (L2 - sk_a) + ((L2 - sk_b) if two else 0)
                    g0 = max(sk_a, (D - u0 + 3) & ~3)
                    g1 = min(L2, D + gp - (63 if two else 31) - u0) & ~3
                    if g1 > g0:
                        uni += (2 if two else 1) * (g1 - g0)
                lo_r, hi_r = np.maximum(us, D), np.minimum(us + L2 - 1, D + gp - 1)
                m_entries += int(np.clip(hi_r - lo_r + 1, 0, None).sum())
                m_shared += int(np.clip(np.minimum(us + L2 - 1, D + gp - L2) - lo_r + 1, 0, None).sum())
    print(f"workdir {wd}: {len(idx)} of {len(gi)} gaps, insert window {m.thr_min}..{m.thr_max}, D {D}")
    print(f"E-step terms per iteration            {terms:.4e}   lane use of the chains {terms / chain_terms / 32:.3f}")
    print(f"  against left-flank rows, per lane   {flank_lane / terms:.4f}")
    print(f"  a chain can skip as a whole         {chain_skip / chain_terms:.4f}")
    print(f"first iteration, chain steps left     {steps:.4e}   all lanes on gap rows {uni / steps:.4f}")
    print(f"first iteration shared per gap        placements {share_terms / terms:.4f}   whole chains {share_chain / chain_terms:.4f}")
    print(f"M-step entries per iteration          {m_entries:.4e}   on rows only shared placements touch {m_shared / m_entries:.4f}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None)
