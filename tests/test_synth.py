"""The workload generator: deterministic, and its --direct output equals what the reference's splitter (main.cpp)
writes from the generator's own result.sam (checked against oracle/_ref/main.out when it has been built)."""
import filecmp
import os
import subprocess

import pytest

import refrun

ARGS = ["--genome", "50000", "--scaffolds", "2", "--gap-list", "120,60,300", "--cov", "20", "--ins-mean", "180", "--ins-sd", "20",
        "--err", "0.01", "--multihit", "2"]


def test_generator_is_deterministic(built, tmp_path):
    from gapprober_b200 import synth
    a, b = str(tmp_path / "a"), str(tmp_path / "b")
    synth.generate(a, ARGS, seed=5)
    synth.generate(b, ARGS, seed=5)
    for f in ("scaffolds.fa", "myout.sam", "gapInfo.txt", "stat.txt", "Gaps/gaps_0.sam", "Gaps/gaps_2.sam"):
        assert filecmp.cmp(os.path.join(a, f), os.path.join(b, f), shallow=False), f


def test_direct_output_equals_reference_splitter(built, tmp_path):
    if not refrun.have_ref():
        pytest.skip("oracle/_ref binaries not built")
    from gapprober_b200 import synth
    d, r = str(tmp_path / "direct"), str(tmp_path / "viaref")
    synth.generate(d, ARGS, seed=7, direct=True, sam=True)
    synth.generate(r, ARGS, seed=7, direct=False, sam=True)
    os.makedirs(os.path.join(r, "Gaps"), exist_ok=True)
    subprocess.run([refrun.REF_MAIN, "scaffolds.fa", "250"], cwd=r, check=True, stdout=subprocess.DEVNULL)
    for f in ("myout.sam", "gapInfo.txt", "stat.txt"):
        assert filecmp.cmp(os.path.join(d, f), os.path.join(r, f), shallow=False), f
    for k in range(3):
        # the splitter leaves md/IH of unpaired records uninitialised (main.cpp:180): compare the 7 columns figbird reads
        cols = lambda p: [ln.split("\t")[:7] for ln in open(p)]
        assert cols(os.path.join(d, "Gaps", f"gaps_{k}.sam")) == cols(os.path.join(r, "Gaps", f"gaps_{k}.sam"))
