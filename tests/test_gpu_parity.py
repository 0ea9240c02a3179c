"""GPU parity: the CUDA path, through the C ABI, against the reference itself (oracle/_ref binaries run on the
same freshly generated inputs) and against the C restatement (oracle/liboracle.so).

Bars: filled sequences byte-identical; per-position probabilities within 1e-5 relative with
identical NaN pattern.  The kernels are built to be bit-exact, so the tests additionally insist that the
per-estimate likelihoods are bit-identical (that is what keeps likelihood plateaus tied)."""
import os

import numpy as np
import pytest

import oracle_lib as O
import refrun

pytestmark = pytest.mark.gpu

CASES = {
    "cfg1_500bp_30x": (1, ["--genome", "100000", "--gap-list", "500", "--cov", "30"]),
    "lib180_100bp_60x": (4, ["--genome", "60000", "--gap-list", "100", "--cov", "60", "--ins-mean", "180", "--ins-sd", "20"]),
    "multi_gap_err1pct": (8, ["--genome", "120000", "--scaffolds", "2", "--gap-list", "50,150,250,400,99", "--cov", "50",
                              "--ins-mean", "180", "--ins-sd", "20", "--err", "0.01"]),
    "tiny_gaps_N_multihit": (9, ["--genome", "80000", "--gap-list", "300,20,7", "--cov", "40", "--ins-mean", "300", "--ins-sd", "30",
                                 "--err", "0.02", "--nrate", "0.01", "--multihit", "3", "--len-noise", "10"]),
    "plateau_1000bp": (6, ["--genome", "100000", "--gap-list", "1000", "--cov", "40"]),
}


def _rel_close(a, b, tol=1e-5):
    nan_same = np.array_equal(np.isnan(a), np.isnan(b))
    m = ~np.isnan(a)
    denom = np.maximum(np.abs(b[m]), 1e-300)
    return nan_same and bool(np.all(np.abs(a[m] - b[m]) <= tol * denom))


@pytest.mark.parametrize("name", list(CASES))
def test_cuda_vs_reference(name, built, tmp_path):
    if not refrun.have_ref():
        pytest.skip("oracle/_ref binaries not built")
    from gapprober_b200 import FigbirdContext, FigbirdModel, synth
    seed, args = CASES[name]
    wd = str(tmp_path / name)
    synth.generate(wd, args, seed=seed)
    ref = refrun.run_reference(wd)
    model = FigbirdModel.load(os.path.join(wd, "fb_model.bin"))
    with FigbirdContext(model) as fb:
        gfs = fb.add_workdir(wd)
        fb.fillGaps()
        st = fb.stats()
        assert st["em_launches"] >= 1 and st["likelihoods"] > 0
        _, seqs = O.read_fasta(os.path.join(wd, "scaffolds.fa"))
        for gi, gf in enumerate(gfs):
            d = ref["dump"][f"Gaps/gaps_{gi}.sam"]
            r = gf.result()
            ref_liks = np.array([d["est"][e] for e in range(r["gap_min"], r["gap_max"] + 1)])
            liks = gf.estimateLikelihoods()
            assert np.array_equal(liks, ref_liks), f"gap {gi}: per-estimate likelihoods differ from the reference"
            assert r["best_estimate"] == d["best"][0]
            assert r["best_lik"] == d["best"][1]
            assert gf.bestString().decode() == d["best"][2]
            assert r["filled"].decode() == ref["gapout"][gi]["filled"]
            assert r["filled_len"] == ref["gapout"][gi]["filled_len"]
            # profile of the winning estimate and of both ends of the scan
            for est in sorted({r["best_estimate"], r["gap_min"], r["gap_max"]}):
                rows = d["rows"][est]
                probs, counts = gf.dumpProfile(est)
                assert _rel_close(probs, rows[:, :5]), f"gap {gi} estimate {est}: probabilities off"
                assert np.array_equal(probs, rows[:, :5], equal_nan=True), f"gap {gi} estimate {est}: probabilities not bit-exact"
                assert np.array_equal(counts, rows[:, 5:], equal_nan=True), f"gap {gi} estimate {est}: counts not bit-exact"
            # the C restatement agrees too (it is what the CPU baseline times)
            recs = O.read_gap_file(os.path.join(wd, "Gaps", f"gaps_{gi}.sam"))
            c, s, l = O.read_gap_info(os.path.join(wd, "gapInfo.txt"))[gi]
            orc = O.fill_gap(ref["model"], seqs[c], s, l, recs)
            assert orc["filled"] == r["filled"] and np.array_equal(orc["liks"], liks)
