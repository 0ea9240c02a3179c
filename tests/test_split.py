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


def _outputs(d):
    names = SAME + ["singletons.sam"] + sorted(os.path.join("Gaps", f) for f in os.listdir(os.path.join(d, "Gaps")))
    return {n: open(os.path.join(d, n), "rb").read() for n in names}


def _mangle(path, seed):
    """Make a SAM the parallel splitter cannot cut blindly: /1 /2 suffixes on some proper pairs (the name changes inside
    a pair), mate lines dropped from some unpaired groups (the machine then pairs lines of different reads), an
    over-long line (split by the reference's 1023-byte fgets) and a record whose name repeats after a gap."""
    import random
    rng = random.Random(seed)
    lines = open(path, "rb").read().split(b"\n")
    out = []
    i = 0
    while i < len(lines):
        ln = lines[i]
        f = ln.split(b"\t")
        if len(f) > 10 and not ln.startswith(b"@"):
            flag = int(f[1])
            r = rng.random()
            if flag & 2 and r < 0.15 and i + 1 < len(lines):
                g = lines[i + 1].split(b"\t")
                f[0] += b"/1"
                g[0] += b"/2"
                out += [b"\t".join(f), b"\t".join(g)]
                i += 2
                continue
            if not flag & 2 and r < 0.05:
                i += 1                      # drop this line: the groups behind it shift
                continue
            if r > 0.999:
                f[9] = f[9] * 12            # > 1023 characters
                f[10] = f[10] * 12
                ln = b"\t".join(f)
        out.append(ln)
        i += 1
    open(path, "wb").write(b"\n".join(out))


@pytest.mark.parametrize("mangled", [False, True])
def test_split_is_the_same_for_every_stretch_size(built, tmp_path, mangled):
    """The file is cut into stretches that run on several threads; whatever the cut points, the outputs are those of
    the one-stretch (= sequential) run, which test_split_equals_reference_main_out pins to main.out."""
    from gapprober_b200 import build, synth
    base = str(tmp_path / "base")
    synth.generate(base, ARGS, seed=43, direct=False, sam=True)
    if mangled:
        _mangle(os.path.join(base, "result.sam"), 7)
    want = None
    for stretch, threads in (("4000000000", "1"), ("700", "3"), ("5000", "8"), ("65536", "2"), ("1000000", "4")):
        d = str(tmp_path / f"s{stretch}")
        shutil.copytree(base, d)
        env = dict(os.environ, FIGBIRD_SPLIT_STRETCH=stretch, FIGBIRD_THREADS=threads, FB_TRACE="1")
        r = subprocess.run([os.path.join(build.BIN, "fb_split"), "scaffolds.fa", "250"], cwd=d, env=env, check=True,
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
        got = dict(_outputs(d), stdout=r.stdout)
        if want is None:
            want = got
            assert b"1 stretches, 0 run again" in r.stderr
        else:
            assert got.keys() == want.keys()
            for k in want:
                assert got[k] == want[k], (stretch, k)
            if not mangled:
                assert b" 0 run again" in r.stderr, r.stderr      # well-formed input: every speculative start was right
