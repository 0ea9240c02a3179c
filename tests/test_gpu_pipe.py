"""fb_fill_gap_files — the whole gap loop of figbird's main() (figbird.cpp:2819-2855) in one call, the gap set handed to
the device in two batches so that the host's parsing/packing/planning of the second overlaps the scan of the first.
Gaps are independent, so every result must be the one fb_add_gap_files + fb_run gives, bit for bit, and a seeded subset
is also checked against the oracle."""
import os

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


def _make(tmp_path, args, seed):
    from gapprober_b200 import FigbirdModel, build, synth
    wd = str(tmp_path / "wd")
    synth.generate(wd, args, seed=seed)
    model = FigbirdModel.load(build.learn_model(wd, out="mine.bin"))
    _, seqs = O.read_fasta(os.path.join(wd, "scaffolds.fa"))
    gaps = O.read_gap_info(os.path.join(wd, "gapInfo.txt"))
    batch = [(c, s, l, os.path.join(wd, "Gaps", f"gaps_{i}.sam"), seqs[c]) for i, (c, s, l) in enumerate(gaps)]
    return wd, model, seqs, gaps, batch


def _all(gfs):
    return [(gf.result(), gf.estimateLikelihoods().tobytes(), gf.bestString()) for gf in gfs]


def test_two_batches_equal_one(built, tmp_path):
    from gapprober_b200 import FigbirdContext, FigbirdError
    wd, model, seqs, gaps, batch = _make(
        tmp_path, ["--genome", "1500000", "--scaffolds", "6", "--gaps", "160", "--gap-min", "50", "--gap-max", "500", "--cov", "40",
                   "--err", "0.01", "--ins-mean", "180", "--ins-sd", "20", "--model-pairs", "100000"], 5)
    assert len(gaps) >= 128                                   # above FB_PIPE_MIN: the two-batch route is taken
    with FigbirdContext(model) as fb:
        gfs = fb.add_gap_files(batch)
        fb.fillGaps()
        plain, st_plain = _all(gfs), fb.stats()
    with FigbirdContext(model) as fb:
        gfs = fb.fill_gap_files(batch)
        st = fb.stats()
        handles = [gf.index for gf in gfs]
        assert sorted(handles) == list(range(len(gaps))) and handles != list(range(len(gaps)))   # two batches: second batch's gaps first
        piped = _all(gfs)
        assert piped == plain
        for k in ("n_gaps", "n_units", "n_records", "likelihoods", "base_terms", "algorithmic_bytes", "finalize_likelihoods"):
            assert st[k] == st_plain[k], k
        assert st["em_launches"] > st_plain["em_launches"] and st["step_ms"] > 0
        # a gap of either batch: the profile tap goes to the context that ran it
        omodel = O.Model.load(os.path.join(wd, "mine.bin"))
        for gi in (3, len(gaps) - 2):
            c, s, l = gaps[gi]
            orc = O.fill_gap(omodel, seqs[c], s, l, O.read_gap_file(os.path.join(wd, "Gaps", f"gaps_{gi}.sam")), dump_estimate=piped[gi][0]["best_estimate"])
            probs, counts = gfs[gi].dumpProfile(piped[gi][0]["best_estimate"])
            assert np.array_equal(counts, orc["counts"]) and np.array_equal(probs, orc["probs"], equal_nan=True)
            assert piped[gi][0]["filled"] == orc["filled"] and piped[gi][1] == np.asarray(orc["liks"]).tobytes()
        # the context can be run again as it is, read, or cleared; it refuses to grow or to be re-partitioned
        fb.runDevice()
        assert _all(gfs) == plain
        with pytest.raises(FigbirdError):
            fb.add_gap_files(batch[:2])
        with pytest.raises(FigbirdError):
            fb.setPartition(0, 2)
        with pytest.raises(FigbirdError):
            fb.fill_gap_files(batch)                          # needs a context without gaps
        fb.clearGaps()
        gfs = fb.fill_gap_files(batch)
        assert _all(gfs) == plain
        # below the threshold (and with FB_NO_PIPE) the plain route runs: handles in the order given
        fb.clearGaps()
        few = fb.fill_gap_files(batch[:40])
        assert [gf.index for gf in few] == list(range(40)) and _all(few) == plain[:40]


def test_small_batches_and_log_mode(built, tmp_path):
    """FB_PIPE_MIN lowers the threshold: 24 gaps in batches of 3 + 21, in f=1 mode as well."""
    from gapprober_b200 import FigbirdContext, FigbirdModel
    wd, model, seqs, gaps, batch = _make(
        tmp_path, ["--genome", "300000", "--scaffolds", "3", "--gaps", "24", "--gap-min", "40", "--gap-max", "300", "--cov", "30",
                   "--err", "0.01", "--ins-mean", "180", "--ins-sd", "20", "--model-pairs", "50000"], 9)
    old = os.environ.get("FB_PIPE_MIN")
    try:
        for f_mode in (0, 1):
            m = FigbirdModel.load(os.path.join(wd, "mine.bin"))
            m.f_mode = f_mode
            os.environ.pop("FB_PIPE_MIN", None)
            with FigbirdContext(m) as fb:
                plain = _all(fb.fill_gap_files(batch))            # 24 < 128: plain route
            os.environ["FB_PIPE_MIN"] = "8"
            with FigbirdContext(m) as fb:
                gfs = fb.fill_gap_files(batch)
                assert [gf.index for gf in gfs][:3] == [21, 22, 23]
                assert _all(gfs) == plain
    finally:
        if old is None:
            os.environ.pop("FB_PIPE_MIN", None)
        else:
            os.environ["FB_PIPE_MIN"] = old


def test_refused_gaps_in_either_batch(built, tmp_path):
    """A gap closer than max_distance to a scaffold end is outside the reference's defined territory (FB_EDOMAIN): its
    handle is negative, the call says so, and every other gap of both batches is filled as usual."""
    import ctypes as C

    from gapprober_b200 import FigbirdContext, GapBatch
    wd, model, seqs, gaps, batch = _make(
        tmp_path, ["--genome", "300000", "--scaffolds", "3", "--gaps", "24", "--gap-min", "40", "--gap-max", "300", "--cov", "30",
                   "--err", "0.01", "--ins-mean", "180", "--ins-sd", "20", "--model-pairs", "50000"], 9)
    with FigbirdContext(model) as fb:
        plain = _all(fb.fill_gap_files(batch))
    bad = {1, 17}                                             # one in the first batch (3 gaps), one in the second
    cut = list(batch)
    for i in bad:
        c, s, l, path, contig = cut[i]
        cut[i] = (c, s, l, path, contig[:s + l + 50])         # the scaffold now ends 50 bp after the gap
    old = os.environ.get("FB_PIPE_MIN")
    os.environ["FB_PIPE_MIN"] = "8"
    try:
        with FigbirdContext(model) as fb:
            gb = GapBatch(cut)
            rc = fb._L.fb_fill_gap_files(fb._h, gb.descs, gb.n, gb.handles, 0)
            assert rc == -2                                   # FB_EDOMAIN: some gaps were refused, the rest ran
            handles = list(gb.handles)[:gb.n]
            assert [h < 0 for h in handles] == [i in bad for i in range(gb.n)]
            assert sorted(h for h in handles if h >= 0) == list(range(gb.n - len(bad)))
            from gapprober_b200.figbird import _FbGapResult
            for i, h in enumerate(handles):
                if h < 0:
                    assert h == -2
                    continue
                r = _FbGapResult()
                buf = C.create_string_buffer(4096)
                assert fb._L.fb_result(fb._h, h, C.byref(r), buf, 4096) == 0
                assert buf.value == plain[i][0]["filled"] and r.best_estimate == plain[i][0]["best_estimate"]
    finally:
        if old is None:
            os.environ.pop("FB_PIPE_MIN", None)
        else:
            os.environ["FB_PIPE_MIN"] = old
