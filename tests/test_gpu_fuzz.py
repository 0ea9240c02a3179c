"""Seeded randomised parity: libraries, error rates, coverage, gap sizes, duplication and N rates drawn at random;
every case is run through the REFERENCE binary (oracle/_ref/figbird_dump.out) and through the CUDA path on the GPU,
and the per-estimate likelihoods, the winning estimate and the filled sequences must be identical."""
import os

import numpy as np
import pytest

import refrun

pytestmark = pytest.mark.gpu


def _draw(seed):
    r = np.random.default_rng(1000 + seed)
    mean = int(r.integers(150, 520))
    sd = int(r.integers(8, max(9, mean // 9)))
    ngaps = int(r.integers(2, 6))
    kinds = r.integers(0, 4, size=ngaps)
    gaps = [int(r.integers(1, 30)) if k == 0 else int(r.integers(30, 200)) if k == 1 else int(r.integers(200, 420)) if k == 2
            else int(r.integers(420, 760)) for k in kinds]
    args = ["--genome", str(int(r.integers(60, 140)) * 1000), "--scaffolds", str(int(r.integers(1, 3))),
            "--gap-list", ",".join(map(str, gaps)), "--cov", str(int(r.integers(8, 90))), "--ins-mean", str(mean), "--ins-sd", str(sd),
            "--read-len", str(int(r.choice([36, 50, 76, 99, 100]))), "--err", f"{float(r.choice([0, 0.002, 0.01, 0.03])):g}",
            "--nrate", f"{float(r.choice([0, 0, 0.005])):g}", "--multihit", str(int(r.choice([1, 1, 1, 2, 4]))),
            "--len-noise", str(int(r.choice([0, 0, 5, 25])))]
    return args


# FB_FUZZ_CASES=<n> widens the sweep (the default 8 keep the suite short); every third case runs the reference's
# log-domain mode (argv[2] = 1: log() per base term, pow(10, score) into the counts)
@pytest.mark.parametrize("seed", range(int(os.environ.get("FB_FUZZ_CASES", "8"))))
def test_random_case_matches_reference(seed, built, tmp_path):
    if not refrun.have_ref():
        pytest.skip("oracle/_ref binaries not built")
    from gapprober_b200 import FigbirdContext, FigbirdModel, build, synth
    wd = str(tmp_path / f"fuzz{seed}")
    args = _draw(seed)
    f_mode = 1 if seed % 3 == 2 else 0
    synth.generate(wd, args, seed=500 + seed)
    ref = refrun.run_reference(wd, dump_probs=False, f_mode=f_mode)
    model = FigbirdModel.load(build.learn_model(wd, out="mine.bin"))           # host preamble ...
    assert open(os.path.join(wd, "mine.bin"), "rb").read() == open(os.path.join(wd, "fb_model.bin"), "rb").read(), args   # ... equals the reference's model
    model.f_mode = f_mode
    with FigbirdContext(model) as fb:
        gfs = fb.add_workdir(wd)
        fb.fillGaps()
        for gi, gf in enumerate(gfs):
            d = ref["dump"][f"Gaps/gaps_{gi}.sam"]
            r = gf.result()
            want = np.array([d["est"][e] for e in range(r["gap_min"], r["gap_max"] + 1)])
            assert np.array_equal(gf.estimateLikelihoods(), want, equal_nan=True), (args, f_mode, gi)
            assert r["best_estimate"] == d["best"][0] and gf.bestString().decode() == d["best"][2], (args, f_mode, gi)
            assert r["filled"].decode() == ref["gapout"][gi]["filled"], (args, f_mode, gi)


@pytest.mark.parametrize("seed", range(100, 100 + int(os.environ.get("FB_FUZZ_F1_CASES", "3"))))
def test_random_case_log_mode_matches_reference(seed, built, tmp_path):
    """f=1 only (FB_FUZZ_F1_CASES widens it): per-estimate likelihoods, winner, bestString and fill against the reference binary."""
    if not refrun.have_ref():
        pytest.skip("oracle/_ref binaries not built")
    from gapprober_b200 import FigbirdContext, FigbirdModel, build, synth
    wd = str(tmp_path / f"f1_{seed}")
    args = _draw(seed)
    synth.generate(wd, args, seed=500 + seed)
    # the reference takes a log() per base term in this mode: its gaps run side by side on the host cores
    n_gaps = len(open(os.path.join(wd, "gapInfo.txt")).read().splitlines())
    ref = refrun.run_reference_on_gaps(wd, list(range(n_gaps)), min(os.cpu_count() or 1, n_gaps), str(tmp_path), f_mode=1)
    model = FigbirdModel.load(build.learn_model(wd, out="mine.bin"))
    model.f_mode = 1
    with FigbirdContext(model) as fb:
        gfs = fb.add_workdir(wd)
        fb.fillGaps()
        assert len(gfs) == n_gaps
        for gi, gf in enumerate(gfs):
            d = ref[gi]
            r = gf.result()
            want = np.array([d["est"][e] for e in range(r["gap_min"], r["gap_max"] + 1)])
            assert np.array_equal(gf.estimateLikelihoods(), want, equal_nan=True), (args, gi)
            assert r["best_estimate"] == d["best"][0] and gf.bestString().decode() == d["best"][2], (args, gi)
            assert r["filled"].decode() == d["filled"], (args, gi)


def _draw_long(seed):
    r = np.random.default_rng(7000 + seed)
    mean, sd = [(180, 20), (300, 30), (500, 50), (420, 12)][int(r.integers(0, 4))]
    gaps = [int(r.integers(600, 4000)) for _ in range(int(r.integers(1, 3)))]
    return ["--genome", str(int(r.integers(40, 70)) * 1000), "--gap-list", ",".join(map(str, gaps)), "--cov", str(int(r.integers(5, 22))),
            "--ins-mean", str(mean), "--ins-sd", str(sd), "--read-len", str(int(r.choice([76, 100]))),
            "--err", f"{float(r.choice([0, 0.01])):g}", "--multihit", str(int(r.choice([1, 1, 3])))]


@pytest.mark.parametrize("seed", range(int(os.environ.get("FB_FUZZ_LONG_CASES", "4"))))
def test_random_long_gap_matches_oracle(seed, built, tmp_path):
    """Long gaps (600 - 4 000 bp): the kernel variants that only they reach — 1024-thread CTAs, tables without the rows
    no placement touches, global-memory tables — against the oracle (pinned to the reference by the golden fixtures;
    the reference binary itself needs minutes per case here), profile dump of the winning estimate included."""
    import oracle_lib as O
    from gapprober_b200 import FigbirdContext, FigbirdModel, build, synth
    wd = str(tmp_path / f"long{seed}")
    args = _draw_long(seed)
    synth.generate(wd, args, seed=900 + seed)
    model = FigbirdModel.load(build.learn_model(wd))
    omodel = O.Model.load(os.path.join(wd, "fb_model.bin"))
    _, seqs = O.read_fasta(os.path.join(wd, "scaffolds.fa"))
    with FigbirdContext(model) as fb:
        gfs = fb.add_workdir(wd)
        fb.fillGaps()
        for gi, (c, s, l) in enumerate(O.read_gap_info(os.path.join(wd, "gapInfo.txt"))):
            recs = O.read_gap_file(os.path.join(wd, "Gaps", f"gaps_{gi}.sam"))
            r = gfs[gi].result()
            orc = O.fill_gap(omodel, seqs[c], s, l, recs, dump_estimate=0)
            assert np.array_equal(gfs[gi].estimateLikelihoods(), orc["liks"]), (args, gi)
            assert (r["best_estimate"], r["filled"]) == (orc["best_estimate"], orc["filled"]), (args, gi)
            assert (r["total_reads"], r["discarded"]) == (orc["total_reads"], orc["discarded"]), (args, gi)
            o2 = O.fill_gap(omodel, seqs[c], s, l, recs, dump_estimate=r["best_estimate"])
            probs, counts = gfs[gi].dumpProfile(r["best_estimate"])
            assert np.array_equal(probs, o2["probs"], equal_nan=True) and np.array_equal(counts, o2["counts"], equal_nan=True), (args, gi)
