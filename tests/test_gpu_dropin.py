"""The drop-in binary (gapprober_b200/bin/figbird.out: host preamble + GPU hot path) against the reference binary
(oracle/_ref/figbird.out) in the same generated working directory: same files, same stdout."""
import filecmp
import os
import shutil
import subprocess

import pytest

import refrun

pytestmark = pytest.mark.gpu

ARGS = ["--genome", "150000", "--scaffolds", "3", "--gap-list", "60,120,250,90,400,33", "--cov", "40", "--ins-mean", "180", "--ins-sd", "20", "--err", "0.01"]


@pytest.mark.parametrize("f", [" 0", " 1"])     # argv[2]: 0 is what GapProber.py passes, 1 what run_EM_eval.sh passes (log-domain scores)
def test_dropin_outputs_equal_reference(built, tmp_path, f):
    if not refrun.have_ref():
        pytest.skip("oracle/_ref binaries not built")
    from gapprober_b200 import build, synth
    a, b = str(tmp_path / "ours"), str(tmp_path / "ref")
    synth.generate(a, ARGS, seed=31)
    shutil.copytree(a, b)
    ours = subprocess.run([build.build_host(), "scaffolds.fa", f, "250"], cwd=a, stdout=subprocess.PIPE, text=True)
    assert ours.returncode == 0, ours.stdout
    ref = subprocess.run([refrun.REF_FIGBIRD, "scaffolds.fa", f, "250"], cwd=b, stdout=subprocess.PIPE, text=True, check=True)
    for f in ("gapout.txt", "filledContigs.fa", "mappedInserts.txt"):
        assert filecmp.cmp(os.path.join(a, f), os.path.join(b, f), shallow=False), f
    strip = lambda s: [ln for ln in s.splitlines() if not ln.startswith("Total time to fill gaps")]
    assert strip(ours.stdout) == strip(ref.stdout)


def test_in_memory_handoff_from_result_sam(built, tmp_path):
    """FIGBIRD_FROM_SAM: result.sam -> split -> preamble -> GPU in one process, no myout.sam / Gaps/ on disk; same gapout.txt
    and filledContigs.fa as the reference pipeline main.out -> figbird.out."""
    if not refrun.have_ref():
        pytest.skip("oracle/_ref binaries not built")
    from gapprober_b200 import build, synth
    a, b = str(tmp_path / "ours"), str(tmp_path / "ref")
    synth.generate(a, ARGS, seed=33, direct=False, sam=True)
    shutil.copytree(a, b)
    os.makedirs(os.path.join(b, "Gaps"), exist_ok=True)
    shutil.rmtree(os.path.join(a, "Gaps"), ignore_errors=True)
    env = dict(os.environ, FIGBIRD_FROM_SAM="result.sam")
    ours = subprocess.run([build.build_host(), "scaffolds.fa", " 0", "250"], cwd=a, env=env, stdout=subprocess.PIPE, text=True)
    assert ours.returncode == 0, ours.stdout
    subprocess.run([refrun.REF_MAIN, "scaffolds.fa", "250"], cwd=b, check=True, stdout=subprocess.DEVNULL)
    ref = subprocess.run([refrun.REF_FIGBIRD, "scaffolds.fa", " 0", "250"], cwd=b, stdout=subprocess.PIPE, text=True, check=True)
    for f in ("gapout.txt", "filledContigs.fa", "mappedInserts.txt", "gapInfo.txt", "stat.txt"):
        assert filecmp.cmp(os.path.join(a, f), os.path.join(b, f), shallow=False), f
    assert not os.path.exists(os.path.join(a, "myout.sam")) and not os.path.exists(os.path.join(a, "Gaps"))
    strip = lambda s: [ln for ln in s.splitlines() if not ln.startswith("Total time to fill gaps")]
    assert strip(ours.stdout) == strip(ref.stdout)


def test_dropin_two_gpus(built, tmp_path):
    """FIGBIRD_GPUS=2: one context and host thread per GPU inside the drop-in binary, gaps dealt by fb_set_partition."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    if not refrun.have_ref():
        pytest.skip("oracle/_ref binaries not built")
    from gapprober_b200 import build, synth
    a, b = str(tmp_path / "ours"), str(tmp_path / "ref")
    synth.generate(a, ARGS, seed=35)
    shutil.copytree(a, b)
    ours = subprocess.run([build.build_host(), "scaffolds.fa", " 0", "250"], cwd=a, env=dict(os.environ, FIGBIRD_GPUS="2"),
                          stdout=subprocess.PIPE, text=True)
    assert ours.returncode == 0, ours.stdout
    subprocess.run([refrun.REF_FIGBIRD, "scaffolds.fa", " 0", "250"], cwd=b, stdout=subprocess.DEVNULL, check=True)
    for f in ("gapout.txt", "filledContigs.fa"):
        assert filecmp.cmp(os.path.join(a, f), os.path.join(b, f), shallow=False), f
