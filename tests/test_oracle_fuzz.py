"""The oracle against the REFERENCE binary on freshly generated inputs (CPU only, both values of argv[2]).

tests/test_oracle_golden.py pins oracle/figbird_oracle.c on the six committed fixtures; this widens the pin to seeded
random workloads that are generated, run through oracle/_ref/figbird_dump.out (the unmodified reference plus output
taps, built by oracle/Makefile) and through the oracle at test time: per-estimate likelihoods bit for bit, winner,
bestString and the filled sequence.  Small on purpose (the scalar oracle and the reference take seconds per gap):
FB_ORACLE_FUZZ_CASES widens it.  TEST INFRASTRUCTURE ONLY - nothing of the product path is exercised here.
"""
import os

import numpy as np
import pytest

import oracle_lib as O
import refrun


def _draw(seed):
    r = np.random.default_rng(4000 + seed)
    mean = int(r.integers(150, 300))       # longer inserts put the anchors of short mates out of the splitter's reach: no records
    sd = int(r.integers(8, max(9, mean // 9)))
    gaps = [int(r.integers(1, 25)), int(r.integers(25, 90))] + ([int(r.integers(90, 160))] if seed % 2 else [])
    return ["--genome", str(int(r.integers(40, 70)) * 1000), "--scaffolds", str(int(r.integers(1, 3))),
            "--gap-list", ",".join(map(str, gaps)), "--cov", str(int(r.integers(6, 22))), "--ins-mean", str(mean), "--ins-sd", str(sd),
            "--read-len", str(int(r.choice([36, 50, 76, 100]))), "--err", f"{float(r.choice([0, 0.01, 0.03])):g}",
            "--nrate", f"{float(r.choice([0, 0.005])):g}", "--multihit", str(int(r.choice([1, 1, 3]))),
            "--len-noise", str(int(r.choice([0, 5, 25])))]


@pytest.mark.parametrize("seed", range(int(os.environ.get("FB_ORACLE_FUZZ_CASES", "8"))))
def test_oracle_matches_reference_binary_on_fresh_input(seed, built, tmp_path):
    if not refrun.have_ref():
        pytest.skip("oracle/_ref binaries not built")
    from gapprober_b200 import synth
    wd = str(tmp_path / f"of{seed}")
    args = _draw(seed)
    f_mode = seed % 2                      # odd seeds: the log-domain mode
    synth.generate(wd, args, seed=900 + seed)
    gaps = O.read_gap_info(os.path.join(wd, "gapInfo.txt"))
    ref = refrun.run_reference_on_gaps(wd, list(range(len(gaps))), min(os.cpu_count() or 1, len(gaps)), str(tmp_path), f_mode=f_mode)
    model = O.Model.load(os.path.join(str(tmp_path), "refpart0", "fb_model.bin"))   # the model the reference itself learned (written by the tap)
    model.f_mode = f_mode
    _, seqs = O.read_fasta(os.path.join(wd, "scaffolds.fa"))
    scored = 0
    for gi, (c, s, l) in enumerate(gaps):
        recs = O.read_gap_file(os.path.join(wd, "Gaps", f"gaps_{gi}.sam"))
        o = O.fill_gap(model, seqs[c], s, l, recs)
        scored += int(len(recs) > 0)
        d = ref[gi]
        want = np.array([d["est"][e] for e in range(o["gap_min"], o["gap_max"] + 1)])
        assert np.array_equal(o["liks"], want, equal_nan=True), (args, f_mode, gi)
        assert o["best_estimate"] == d["best"][0] and o["best_string"].decode() == d["best"][2], (args, f_mode, gi)
        assert o["filled"].decode() == d["filled"], (args, f_mode, gi)
    assert scored, (args, "no gap of this case had any record: the comparison would be vacuous")
