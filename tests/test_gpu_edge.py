"""Edge cases of the hot path on the GPU, each checked against the oracle (C restatement) on the same inputs:
empty and ragged inputs, the longest gaps (global-memory table variant), heavy multi-hit duplication, partition
invariance, idempotence, and the error contract."""
import os

import numpy as np
import pytest

import golden_lib as G
import oracle_lib as O

pytestmark = pytest.mark.gpu


def _check_against_oracle(fb, gf, omodel, window, D, orig, recs):
    r = gf.result()
    orc = O.fill_gap(omodel, window, D, orig, recs)
    assert np.array_equal(gf.estimateLikelihoods(), orc["liks"])
    assert r["best_estimate"] == orc["best_estimate"]
    assert r["filled"] == orc["filled"]
    assert (r["total_reads"], r["discarded"]) == (orc["total_reads"], orc["discarded"])
    return r, orc


def test_gap_without_records(built):
    from gapprober_b200 import FigbirdContext
    case = G.Case("lib180_100bp_60x")
    g = case.gaps[0]
    with FigbirdContext(case.figbird_model()) as fb:
        gf = fb.GapFiller().initGapFiller(0, case.D, g["orig"], [], g["window"])
        fb.fillGaps()
        r, orc = _check_against_oracle(fb, gf, case.oracle_model(), g["window"], case.D, g["orig"], [])
        assert r["best_estimate"] == r["gap_max"] and set(r["filled"]) == {ord("N")}   # all-zero likelihood ties: last estimate wins


def test_ragged_read_lengths_and_lower_case(built):
    from gapprober_b200 import FigbirdContext
    case = G.Case("multi_gap_err1pct")
    rng = np.random.default_rng(11)
    with FigbirdContext(case.figbird_model()) as fb:
        todo = []
        for g in case.gaps[:3]:
            recs = []
            for p, l1, s1, seq in g["recs"]:
                n = int(rng.integers(25, len(seq) + 1))
                q = seq[:n] if rng.random() < 0.5 else seq[len(seq) - n:]
                if rng.random() < 0.2:
                    q = q.lower()          # reverse() accepts lower case, charCodes does not (figbird.cpp:1390, 2753)
                recs.append((p, l1, s1, q))
            todo.append((fb.GapFiller().initGapFiller(0, case.D, g["orig"], recs, g["window"]), g, recs))
        fb.fillGaps()
        for gf, g, recs in todo:
            _check_against_oracle(fb, gf, case.oracle_model(), g["window"], case.D, g["orig"], recs)


@pytest.mark.parametrize("compact", [True, False])
def test_long_gap(built, tmp_path, compact):
    """A 3000 bp gap scans estimates 2850..3149.  By default the middle rows that no placement can reach are left out
    of the tables and the unit runs from shared memory; with FB_NO_COMPACT=1 the full profile no longer fits and the
    kernel variant that keeps it in global memory runs.  Same answers as the oracle either way, NaN rows included."""
    from gapprober_b200 import FigbirdContext, FigbirdModel, build, synth
    wd = str(tmp_path / "long")
    synth.generate(wd, ["--genome", "60000", "--gap-list", "3000", "--cov", "10", "--ins-mean", "180", "--ins-sd", "20", "--err", "0.01"], seed=13)
    model = FigbirdModel.load(build.learn_model(wd))
    omodel = O.Model.load(os.path.join(wd, "fb_model.bin"))
    _, seqs = O.read_fasta(os.path.join(wd, "scaffolds.fa"))
    c, s, l = O.read_gap_info(os.path.join(wd, "gapInfo.txt"))[0]
    recs = O.read_gap_file(os.path.join(wd, "Gaps", "gaps_0.sam"))
    old = os.environ.pop("FB_NO_COMPACT", None)
    if not compact:
        os.environ["FB_NO_COMPACT"] = "1"
    try:
        with FigbirdContext(model) as fb:
            gf = fb.GapFiller().initGapFiller(c, s, l, os.path.join(wd, "Gaps", "gaps_0.sam"), seqs[c])
            fb.fillGaps()
            r, orc = _check_against_oracle(fb, gf, omodel, seqs[c], s, l, recs)
            assert r["gap_max"] == O.gap_range(3000)[1] == 3149      # 3000*1.05f truncates (float arithmetic, figbird.cpp:2453)
            probs, counts = gf.dumpProfile(r["best_estimate"])
            o2 = O.fill_gap(omodel, seqs[c], s, l, recs, dump_estimate=r["best_estimate"])
            assert np.array_equal(probs, o2["probs"], equal_nan=True) and np.array_equal(counts, o2["counts"], equal_nan=True)
            assert np.isnan(probs[200 + 1500, :4]).all()                # the middle of the gap: 0/0 in the reference too
    finally:
        os.environ.pop("FB_NO_COMPACT", None)
        if old is not None:
            os.environ["FB_NO_COMPACT"] = old


def test_multihit_duplicates_many_chunks(built, tmp_path):
    """bowtie2 -k style duplication: 10 copies of every record, several hundred records in one gap."""
    from gapprober_b200 import FigbirdContext, FigbirdModel, build, synth
    wd = str(tmp_path / "k10")
    synth.generate(wd, ["--genome", "60000", "--gap-list", "100", "--cov", "30", "--ins-mean", "180", "--ins-sd", "20", "--err", "0.01", "--multihit", "10"], seed=17)
    model = FigbirdModel.load(build.learn_model(wd))
    omodel = O.Model.load(os.path.join(wd, "fb_model.bin"))
    _, seqs = O.read_fasta(os.path.join(wd, "scaffolds.fa"))
    c, s, l = O.read_gap_info(os.path.join(wd, "gapInfo.txt"))[0]
    recs = O.read_gap_file(os.path.join(wd, "Gaps", "gaps_0.sam"))
    assert len(recs) > 150
    with FigbirdContext(model) as fb:
        gf = fb.GapFiller().initGapFiller(c, s, l, recs, seqs[c])
        fb.fillGaps()
        _check_against_oracle(fb, gf, omodel, seqs[c], s, l, recs)


def test_long_reads_125bp(built, tmp_path):
    """Mates of 125 bp (beyond the reference's 100-byte scratch buffer, figbird.cpp:1698; within the packed layout's 128)."""
    from gapprober_b200 import FigbirdContext, FigbirdModel, build, synth
    wd = str(tmp_path / "l125")
    synth.generate(wd, ["--genome", "60000", "--gap-list", "150,40", "--cov", "40", "--read-len", "125", "--ins-mean", "300", "--ins-sd", "25",
                        "--err", "0.01"], seed=19)
    model = FigbirdModel.load(build.learn_model(wd))
    assert model.max_read_len == 125
    omodel = O.Model.load(os.path.join(wd, "fb_model.bin"))
    _, seqs = O.read_fasta(os.path.join(wd, "scaffolds.fa"))
    with FigbirdContext(model) as fb:
        gfs = fb.add_workdir(wd)
        fb.fillGaps()
        for gi, (c, s, l) in enumerate(O.read_gap_info(os.path.join(wd, "gapInfo.txt"))):
            recs = O.read_gap_file(os.path.join(wd, "Gaps", f"gaps_{gi}.sam"))
            _check_against_oracle(fb, gfs[gi], omodel, seqs[c], s, l, recs)


def test_partition_invariance_and_idempotence(built):
    from gapprober_b200 import FigbirdContext
    case = G.Case("multi_gap_err1pct")
    with FigbirdContext(case.figbird_model()) as fb:
        gfs = [fb.GapFiller().initGapFiller(0, case.D, g["orig"], g["recs"], g["window"]) for g in case.gaps]
        fb.fillGaps()
        full = [(gf.result()["filled"], gf.estimateLikelihoods().tobytes()) for gf in gfs]
        fb.runDevice()                                  # same inputs, second pass: identical
        assert full == [(gf.result()["filled"], gf.estimateLikelihoods().tobytes()) for gf in gfs]
        seen = {}
        for rank in range(3):
            fb.setPartition(rank, 3)
            fb.fillGaps()
            for i, gf in enumerate(gfs):
                if fb.ownsGap(i):
                    assert i not in seen
                    seen[i] = (gf.result()["filled"], gf.estimateLikelihoods().tobytes())
        assert [seen[i] for i in range(len(gfs))] == full


def test_estimate_level_partition(built):
    """(gap, estimate) units dealt to three emulated ranks: each scans its share, the per-estimate likelihoods are merged
    (what multigpu.fill_partitioned all-gathers), and the rank holding the winning estimate finalizes the gap."""
    from gapprober_b200 import FigbirdContext, multigpu
    for name in ("plateau_1000bp", "multi_gap_err1pct"):
        case = G.Case(name)
        world = 3
        ctxs = [FigbirdContext(case.figbird_model()) for _ in range(world)]
        try:
            gfs = []
            for rank, fb in enumerate(ctxs):
                gfs.append([fb.GapFiller().initGapFiller(0, case.D, g["orig"], g["recs"], g["window"]) for g in case.gaps])
                fb.setPartitionUnits(rank, world)
                fb.upload()
                fb.runScan()
            total_units = sum(fb.stats()["n_units"] for fb in ctxs)
            assert total_units == sum(len(g["liks"]) for g in case.gaps)          # every unit scanned exactly once
            assert all(fb.stats()["n_units"] > 0 for fb in ctxs)
            for gi, g in enumerate(case.gaps):
                parts = [fb.scanLiks(gi) for fb in ctxs]
                assert np.array_equal(np.sum([~np.isnan(p) for p in parts], axis=0), np.ones(len(g["liks"])))
                merged = multigpu.merge_liks(parts)
                assert np.array_equal(merged, g["liks"])
                for fb in ctxs:
                    fb.mergeLiks(gi, merged)
            for fb in ctxs:
                fb.runFinalize()
            for gi, g in enumerate(case.gaps):
                owners = [r for r, fb in enumerate(ctxs) if fb.hasResult(gi)]
                assert len(owners) == 1
                r = gfs[owners[0]][gi].result()
                assert r["filled"] == g["filled"] and r["best_estimate"] == g["best_est"] and r["best_lik"] == g["best_lik"]
        finally:
            for fb in ctxs:
                fb.close()


def test_finalize_needs_merged_likelihoods(built):
    from gapprober_b200 import FigbirdContext, FigbirdError
    case = G.Case("lib180_100bp_60x")
    g = case.gaps[0]
    with FigbirdContext(case.figbird_model()) as fb:
        fb.GapFiller().initGapFiller(0, case.D, g["orig"], g["recs"], g["window"])
        fb.setPartitionUnits(0, 2)
        fb.upload()
        fb.runScan()
        with pytest.raises(FigbirdError) as e:
            fb.runFinalize()
        assert e.value.status == -5


def test_error_contract(built):
    from gapprober_b200 import FigbirdContext, FigbirdError
    case = G.Case("lib180_100bp_60x")
    g = case.gaps[0]
    with FigbirdContext(case.figbird_model()) as fb:
        with pytest.raises(FigbirdError) as e:      # gap closer than max_distance to the scaffold start: reference reads out of bounds
            fb.GapFiller().initGapFiller(0, case.D - 1, g["orig"], g["recs"], g["window"][1:])
        assert e.value.status == -2
        gf = fb.GapFiller().initGapFiller(0, case.D, g["orig"], g["recs"], g["window"])
        with pytest.raises(FigbirdError) as e:      # result before run
            gf.result()
        assert e.value.status == -5
        fb.fillGaps()
        assert gf.result()["filled"] == g["filled"]


def _fill_with_env(case, gaps, env):
    """Fill `gaps` in a fresh context with tuning knobs set (they are read when the gap set is uploaded)."""
    from gapprober_b200 import FigbirdContext
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        with FigbirdContext(case.figbird_model()) as fb:
            gfs = [fb.GapFiller().initGapFiller(0, case.D, g["orig"], g["recs"], g["window"]) for g in gaps]
            fb.fillGaps()
            out = []
            for gf in gfs:
                r = gf.result()
                probs, counts = gf.dumpProfile(r["best_estimate"])
                out.append((r["filled"], r["best_estimate"], gf.estimateLikelihoods().tobytes(), probs.tobytes(), counts.tobytes()))
            return out
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_kernel_variants_agree_bit_for_bit(built):
    """One CTA per unit, clusters of 2 and 4 CTAs per unit, 512- and 1024-thread CTAs, other chunk sizes: the same bits.
    plateau_1000bp scans estimates up to 1 050 rows (one CTA per SM), the other cases fit two CTAs per SM."""
    for name, n in (("multi_gap_err1pct", 3), ("plateau_1000bp", 1), ("lib180_100bp_60x", 1)):
        case = G.Case(name)
        gaps = case.gaps[:n]
        ref = _fill_with_env(case, gaps, {"FB_CLUSTER": "1", "FB_THREADS1": "512"})
        for g, r in zip(gaps, ref):
            assert r[0] == g["filled"] and r[1] == g["best_est"] and r[2] == np.asarray(g["liks"]).tobytes()
        # FB_HOST_LOG=1: per-record maxima to the host, log() and the record-order sum by the host libm (the round-1 route)
        # instead of the device's restatement of glibc's log — the likelihood bits must not depend on who takes the log
        for env in ({"FB_CLUSTER": "2"}, {"FB_CLUSTER": "4"}, {"FB_CLUSTER": "1"}, {"FB_CLUSTER": "4", "FB_CHUNK": "8"},
                    {"FB_NO_COMPACT": "1"}, {"FB_NO_COMPACT": "1", "FB_CLUSTER": "2"},
                    {"FB_CLUSTER": "1", "FB_CHUNK": "32"}, {"FB_CLUSTER": "2", "FB_CHUNK": "5"},
                    {"FB_HOST_LOG": "1"}, {"FB_HOST_LOG": "1", "FB_CLUSTER": "4"},
                    {"FB_TWO_CALL": "1"},                      # scan and finalize as two calls with the selection on the host
                    {"FB_ONE_CTA": "1", "FB_CLUSTER": "2"}, {"FB_ONE_CTA": "1", "FB_CLUSTER": "1"},
                    {"FB_NO_PREFIX": "1"}, {"FB_NO_PREFIX": "1", "FB_CLUSTER": "2"},    # every chain from ILD[t], no left-flank prefix table
                    {"FB_NO_T0": "1"}, {"FB_NO_T0": "1", "FB_NO_PREFIX": "1"}):         # first iteration from the tables like the others
            assert _fill_with_env(case, gaps, env) == ref, (name, env)


def test_log_mode_variants_agree_bit_for_bit(built):
    """f=1 (log-domain scores): cluster sizes, chunk sizes and the host-libm likelihood route give the golden bits."""
    case = G.Case("logmode_f1")
    gaps = case.gaps
    ref = _fill_with_env(case, gaps, {"FB_CLUSTER": "1"})
    for g, r in zip(gaps, ref):
        assert r[0] == g["filled"] and r[1] == g["best_est"] and r[2] == np.asarray(g["liks"]).tobytes()
    for env in ({"FB_CLUSTER": "2"}, {"FB_CLUSTER": "4", "FB_CHUNK": "8"}, {"FB_HOST_LOG": "1"}, {"FB_NO_COMPACT": "1", "FB_CLUSTER": "2"},
                {"FB_TWO_CALL": "1"}, {"FB_NO_PREFIX": "1"}, {"FB_NO_T0": "1"}):
        assert _fill_with_env(case, gaps, env) == ref, env
