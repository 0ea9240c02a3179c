"""GPU parity against the committed golden fixtures (answers of the reference itself, tests/golden/make_golden.py).
Through the C ABI; bit-exact on likelihoods, counts, probabilities and sequences."""
import numpy as np
import pytest

import golden_lib as G

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", G.ALL_CASES)      # logmode_f1 included: the reference's log-domain mode (argv[2] = 1)
def test_cuda_matches_reference_golden(name, built):
    from gapprober_b200 import FigbirdContext
    case = G.Case(name)
    with FigbirdContext(case.figbird_model()) as fb:
        gfs = [fb.GapFiller().initGapFiller(0, case.D, g["orig"], g["recs"], g["window"]) for g in case.gaps]
        fb.fillGaps()
        for gi, (gf, g) in enumerate(zip(gfs, case.gaps)):
            r = gf.result()
            assert np.array_equal(gf.estimateLikelihoods(), g["liks"], equal_nan=True), f"{name} gap {gi}: per-estimate likelihoods"
            assert r["best_estimate"] == g["best_est"] and r["best_lik"] == g["best_lik"]
            assert gf.bestString() == g["best_string"]
            assert r["filled"] == g["filled"] and r["filled_len"] == len(g["filled"])
            assert (r["total_reads"], r["discarded"]) == g["counts"]
            probs, counts = gf.dumpProfile(g["best_est"])
            # stated bar: 1e-5 relative with identical NaN pattern; the kernels are built to be bit-exact
            ref_p = g["rows"][:, :5]
            assert np.array_equal(np.isnan(probs), np.isnan(ref_p))
            ok = ~np.isnan(ref_p)
            assert np.all(np.abs(probs[ok] - ref_p[ok]) <= 1e-5 * np.maximum(np.abs(ref_p[ok]), 1e-300))
            assert np.array_equal(probs, ref_p, equal_nan=True)
            assert np.array_equal(counts, g["rows"][:, 5:], equal_nan=True)
