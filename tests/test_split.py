"""The splitter restatement (csrc/host/fb_split.hpp) against the reference's main.out on the same
result.sam: every file byte-identical, except the md/IH columns the reference prints from uninitialised memory for
unpaired records (main.cpp:180) — there the seven columns figbird reads are compared."""
import filecmp
import os
import shutil
import subprocess
import time

import pytest

import refrun

ARGS = ["--genome", "90000", "--scaffolds", "3", "--gap-list", "120,60,300,45,700", "--cov", "25", "--ins-mean", "200", "--ins-sd", "25",
        "--err", "0.01", "--nrate", "0.002", "--multihit", "2"]
SAME = ["myout.sam", "gapInfo.txt", "stat.txt", "linking.sam", "unmapped.txt", "scaffold.txt", "counts.txt"]


def _cols7(path):
    return [ln.split(b"\t")[:7] for ln in open(path, "rb")]     # binary: the reference's md column can be arbitrary bytes


def test_split_equals_reference_main_out(built, tmp_path):
    if not refrun.have_ref():
        pytest.skip("oracle/_ref binaries not built")
    from gapprober_b200 import build, synth
    a, b = str(tmp_path / "ours"), str(tmp_path / "ref")
    synth.generate(a, ARGS, seed=41, direct=False, sam=True)
    shutil.copytree(a, b)
    t0 = time.perf_counter()
    subprocess.run([os.path.join(build.BIN, "fb_split"), "scaffolds.fa", "250"], cwd=a, check=True)
    t1 = time.perf_counter()
    subprocess.run([refrun.REF_MAIN, "scaffolds.fa", "250"], cwd=b, check=True, stdout=subprocess.DEVNULL)
    t2 = time.perf_counter()
    for f in SAME:
        assert filecmp.cmp(os.path.join(a, f), os.path.join(b, f), shallow=False), f
    assert _cols7(os.path.join(a, "singletons.sam")) == _cols7(os.path.join(b, "singletons.sam"))
    for k in range(5):
        assert _cols7(os.path.join(a, "Gaps", f"gaps_{k}.sam")) == _cols7(os.path.join(b, "Gaps", f"gaps_{k}.sam"))
    mb = os.path.getsize(os.path.join(a, "result.sam")) / 1e6
    print(f"split {mb:.1f} MB SAM: ours {mb / (t1 - t0):.0f} MB/s, reference {mb / (t2 - t1):.0f} MB/s")


def test_in_memory_handoff_learns_the_same_model(built, tmp_path):
    """FIGBIRD_FROM_SAM: split + preamble in one process, nothing but gapInfo.txt/stat.txt written; same model bytes as
    the file route (fb_split -> myout.sam -> preamble)."""
    from gapprober_b200 import build, synth
    wd = str(tmp_path / "w")
    synth.generate(wd, ARGS, seed=42, direct=True, sam=True)
    via_files = open(build.learn_model(wd, out="files.bin"), "rb").read()
    os.remove(os.path.join(wd, "myout.sam"))
    shutil.rmtree(os.path.join(wd, "Gaps"))
    env = dict(os.environ, FIGBIRD_FROM_SAM="result.sam", FIGBIRD_PREAMBLE_ONLY="1", FIGBIRD_MODEL_OUT="mem.bin")
    subprocess.run([build.build_host(), "scaffolds.fa", " 0", "250"], cwd=wd, env=env, check=True, stdout=subprocess.DEVNULL)
    assert open(os.path.join(wd, "mem.bin"), "rb").read() == via_files
    assert not os.path.exists(os.path.join(wd, "myout.sam"))
