"""BASELINE.json's bench configuration (configs[1], R. sphaeroides scale) at full size through the C ABI, checked with
size-independent properties plus oracle spot checks on seeded gaps."""
import hashlib
import os

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


def test_config2_full_size_properties(built, tmp_path):
    from gapprober_b200 import FigbirdContext, FigbirdModel, build, synth
    wd = str(tmp_path / "cfg2")
    synth.generate(wd, list(synth.CONFIGS["cfg2_rhodobacter"]) + ["--model-pairs", "100000"], seed=2)
    model = FigbirdModel.load(build.learn_model(wd))
    omodel = O.Model.load(os.path.join(wd, "fb_model.bin"))
    _, seqs = O.read_fasta(os.path.join(wd, "scaffolds.fa"))
    gaps = O.read_gap_info(os.path.join(wd, "gapInfo.txt"))
    assert len(gaps) == 200

    def digest(fb, gfs, only=None):
        h = hashlib.sha256()
        for i, gf in enumerate(gfs):
            if only is None or only(i):
                h.update(gf.result()["filled"] + b"|")
        return h.hexdigest()

    with FigbirdContext(model) as fb:
        gfs = fb.add_workdir(wd)
        fb.fillGaps()
        st = fb.stats()
        full = [gf.result() for gf in gfs]
        # every estimate in range, every filled string over the ACGTN alphabet and of the winning length
        for r in full:
            assert r["gap_min"] <= r["best_estimate"] <= r["gap_max"] and r["filled_len"] == r["best_estimate"]
            assert set(r["filled"]) <= set(b"ACGTN")
        # the analytic likelihood counter (the metric's numerator) agrees with the oracle's counter
        rng = np.random.default_rng(5)
        picks = sorted(rng.choice(len(gaps), size=5, replace=False).tolist())
        for gi in picks:
            c, s, l = gaps[gi]
            recs = O.read_gap_file(os.path.join(wd, "Gaps", f"gaps_{gi}.sam"))
            a, b = O.gap_range(l)
            n = sum(O.count_likelihoods(omodel, s, gp, recs) for gp in range(a, b + 1)) * omodel.num_iter
            assert full[gi]["likelihoods"] == n
            orc = O.fill_gap(omodel, seqs[c], s, l, recs)         # oracle spot check, bit-exact
            assert orc["filled"] == full[gi]["filled"] and np.array_equal(orc["liks"], gfs[gi].estimateLikelihoods())
        assert st["likelihoods"] == sum(r["likelihoods"] for r in full)
        d_full = digest(fb, gfs)
        fb.runDevice()
        assert digest(fb, gfs) == d_full                          # idempotent
        # partition invariance: two shares reproduce the single-GPU answers
        parts = {}
        for rank in range(2):
            fb.setPartition(rank, 2)
            fb.fillGaps()
            for i, gf in enumerate(gfs):
                if fb.ownsGap(i):
                    parts[i] = gf.result()["filled"]
        assert [parts[i] for i in range(len(gfs))] == [r["filled"] for r in full]
