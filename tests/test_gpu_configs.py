"""BASELINE.json configs[2..4] at full size through the C ABI, against the oracle (bit-exact, seeded gaps) and against
the reference binary itself (oracle/_ref/figbird_dump.out run on the box's host cores):

  cfg4_chr14   the target workload of north_star: 5 000 gaps; 50 seeded gaps vs the oracle, a 300-gap slice vs the
               reference binary, 8-way partition invariance of all 5 000 fills
  cfg3_staph   all 100 gaps vs the oracle, dealt by gap, and record-sharded over 2 and 3 local shards
  cfg5 corners 5 000 bp x 300x x k=10 and 100 bp x 300x x k=50: every estimate on the GPU, seeded estimates (and the
               winner) vs the oracle, which needs minutes per full scan of these gaps on one core
"""
import hashlib
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import oracle_lib as O
import refrun

pytestmark = pytest.mark.gpu

THREADS = max(2, min(16, os.cpu_count() or 2))      # the oracle is called through ctypes, which drops the GIL


def _make(tmp_path, cfg, seed):
    from gapprober_b200 import FigbirdModel, build, synth
    wd = str(tmp_path / cfg)
    synth.generate(wd, synth.config_args(cfg), seed=seed)
    model = FigbirdModel.load(build.learn_model(wd, out="mine.bin"))
    omodel = O.Model.load(os.path.join(wd, "mine.bin"))
    _, seqs = O.read_fasta(os.path.join(wd, "scaffolds.fa"))
    gaps = O.read_gap_info(os.path.join(wd, "gapInfo.txt"))
    return wd, model, omodel, seqs, gaps


def _oracle_many(omodel, wd, seqs, gaps, ids, **kw):
    def one(gi):
        c, s, l = gaps[gi]
        recs = O.read_gap_file(os.path.join(wd, "Gaps", f"gaps_{gi}.sam"))
        return gi, O.fill_gap(omodel, seqs[c], s, l, recs, **kw)
    with ThreadPoolExecutor(THREADS) as ex:
        return dict(ex.map(one, ids))


def _sha(fills):
    h = hashlib.sha256()
    for f in fills:
        h.update(f + b"|")
    return h.hexdigest()


def test_cfg4_chr14_target_workload(built, tmp_path):
    from gapprober_b200 import FigbirdContext
    wd, model, omodel, seqs, gaps = _make(tmp_path, "cfg4_chr14", 2)
    assert len(gaps) == 5000
    with FigbirdContext(model) as fb:
        gfs = fb.add_workdir(wd)
        fb.fillGaps()
        st = fb.stats()
        full = [gf.result() for gf in gfs]
        liks = {gi: gfs[gi].estimateLikelihoods() for gi in range(len(gaps))}
        assert st["n_gaps"] == 5000 and st["likelihoods"] == sum(r["likelihoods"] for r in full) > 3e9
        # ---- 50 seeded gaps against the oracle, bit for bit
        rng = np.random.default_rng(14)
        picks = sorted(rng.choice(len(gaps), size=50, replace=False).tolist())
        for gi, orc in _oracle_many(omodel, wd, seqs, gaps, picks).items():
            r = full[gi]
            assert np.array_equal(liks[gi], orc["liks"]), gi
            assert (r["best_estimate"], r["filled"], r["total_reads"], r["discarded"]) == \
                   (orc["best_estimate"], orc["filled"], orc["total_reads"], orc["discarded"]), gi
        # ---- a 300-gap slice against the reference binary on the host cores: every estimate's likelihood, the winner, the fill
        if refrun.have_ref():
            ids = list(range(1700, 2000))
            ref = refrun.run_reference_on_gaps(wd, ids, THREADS, str(tmp_path))
            assert open(os.path.join(str(tmp_path), "refpart0", "fb_model.bin"), "rb").read() == open(os.path.join(wd, "mine.bin"), "rb").read()
            for gi in ids:
                r, d = full[gi], ref[gi]
                want = np.array([d["est"][e] for e in range(r["gap_min"], r["gap_max"] + 1)])
                assert np.array_equal(liks[gi], want), gi
                assert r["best_estimate"] == d["best"][0] and r["filled"].decode() == d["filled"], gi
        # ---- gap-partitioned over 8 ranks (north_star's split): the union of the shares is the single-GPU answer
        parts = {}
        for rank in range(8):
            fb.setPartition(rank, 8)
            fb.fillGaps()
            for i, gf in enumerate(gfs):
                if fb.ownsGap(i):
                    assert i not in parts
                    parts[i] = gf.result()["filled"]
        assert len(parts) == 5000
        assert _sha(parts[i] for i in range(5000)) == _sha(r["filled"] for r in full)


def test_cfg3_staph_by_gap_and_record_sharded(built, tmp_path):
    from gapprober_b200 import FigbirdContext, multigpu
    wd, model, omodel, seqs, gaps = _make(tmp_path, "cfg3_staph", 3)
    assert len(gaps) == 100
    ids = list(range(100))
    orc = _oracle_many(omodel, wd, seqs, gaps, ids)
    with FigbirdContext(model) as fb:
        gfs = fb.add_workdir(wd)
        fb.fillGaps()
        for gi in ids:
            r = gfs[gi].result()
            assert np.array_equal(gfs[gi].estimateLikelihoods(), orc[gi]["liks"]), gi
            assert (r["best_estimate"], r["filled"], r["discarded"]) == (orc[gi]["best_estimate"], orc[gi]["filled"], orc[gi]["discarded"]), gi
        parts = {}
        for rank in range(2):                     # dealt by gap
            fb.setPartition(rank, 2)
            fb.fillGaps()
            parts.update({i: gf.result()["filled"] for i, gf in enumerate(gfs) if fb.ownsGap(i)})
        assert [parts[i] for i in ids] == [orc[i]["filled"] for i in ids]
    # ---- read-sharded (BASELINE config 3): 2 and 3 shards, all in this process on cuda:0
    for world in (2, 3):
        sh = _oracle_many(omodel, wd, seqs, gaps, ids, shards=world)
        ctxs = [FigbirdContext(model) for _ in range(world)]
        try:
            gfs = [fb.add_workdir(wd) for fb in ctxs]
            multigpu.scan_record_sharded_local(ctxs)
            for gi in ids:
                l0 = ctxs[0].scanLiks(gi)
                assert np.array_equal(l0, sh[gi]["liks"]), (world, gi)                 # the sharded grouping, bit for bit
                assert np.allclose(l0, orc[gi]["liks"], rtol=1e-9, atol=0), (world, gi)  # the reference's, to rounding
            for fb in ctxs:
                fb.runFinalize()
            for gi in ids:
                r = gfs[gi % world][gi].result()
                assert r["filled"] == orc[gi]["filled"] == sh[gi]["filled"], (world, gi)   # identical sequences
                assert r["best_estimate"] == orc[gi]["best_estimate"], (world, gi)
        finally:
            for fb in ctxs:
                fb.close()


@pytest.mark.parametrize("cfg", ["cfg5_stress_g5000_c300_k10", "cfg5_stress_g100_c300_k50"])
def test_cfg5_stress_corner(built, tmp_path, cfg):
    from gapprober_b200 import FigbirdContext
    wd, model, omodel, seqs, gaps = _make(tmp_path, cfg, 5)
    assert len(gaps) == 1
    c, s, l = gaps[0]
    recs = O.read_gap_file(os.path.join(wd, "Gaps", "gaps_0.sam"))
    assert len(recs) > 4000
    with FigbirdContext(model) as fb:
        gf = fb.add_workdir(wd)[0]
        fb.fillGaps()
        r, liks = gf.result(), gf.estimateLikelihoods()
        assert r["likelihoods"] == sum(O.count_likelihoods(omodel, s, gp, recs) for gp in range(r["gap_min"], r["gap_max"] + 1)) * omodel.num_iter
        # the oracle scans seeded estimates + both ends of the range + the GPU's winner; selection and finalize then run
        # over that subset, so the winner and the fill must come out the same when the likelihood bits agree
        rng = np.random.default_rng(55)
        sub = sorted(set(rng.choice(np.arange(r["gap_min"], r["gap_max"] + 1), size=4, replace=False).tolist())
                     | {r["gap_min"], r["gap_max"], r["best_estimate"]})
        best_in_sub = max(sub, key=lambda e: (liks[e - r["gap_min"]], e))
        assert best_in_sub == r["best_estimate"]
        halves = [sub[0::2], sub[1::2]]           # two oracle calls side by side
        with ThreadPoolExecutor(2) as ex:
            outs = list(ex.map(lambda es: O.fill_gap(omodel, seqs[c], s, l, recs, estimates=es), halves))
        for es, o in zip(halves, outs):
            for e in es:
                assert o["liks"][e - r["gap_min"]] == liks[e - r["gap_min"]], e
        o = O.fill_gap(omodel, seqs[c], s, l, recs, estimates=[r["best_estimate"]])
        assert o["best_string"] == gf.bestString()
        if not o.get("undefined_finalize"):     # >= 10000 reads on one base: the reference's finalize leaves its arrays there
            assert (o["filled"], o["total_reads"], o["discarded"]) == (r["filled"], r["total_reads"], r["discarded"])
        # idempotent, and the estimate-level split (what a multi-GPU run of a single gap uses) changes nothing
        d0 = (r["best_estimate"], r["filled"])
        fb.runDevice()
        assert (gf.result()["best_estimate"], gf.result()["filled"]) == d0
        merged = None
        per_rank = []
        for rank in range(2):
            fb.setPartitionUnits(rank, 2)
            fb.upload()
            fb.runScan()
            per_rank.append(fb.scanLiks(0))
        merged = np.where(np.isnan(per_rank[0]), per_rank[1], per_rank[0])
        assert np.array_equal(merged, liks)
