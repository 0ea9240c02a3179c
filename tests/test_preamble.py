"""Host preamble (gapprober_b200/csrc/host/fb_preamble.hpp, the model learning of figbird.cpp:154-1351) against
the model the reference itself learned from the same myout.sam (golden fixture made by running the reference)."""
import gzip
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "golden", "preamble")


def _workdir(tmp_path):
    wd = str(tmp_path)
    for f in ("myout.sam", "scaffolds.fa"):
        open(os.path.join(wd, f), "wb").write(gzip.open(os.path.join(SRC, f + ".gz")).read())
    for f in ("stat.txt", "gapInfo.txt"):
        shutil.copy(os.path.join(SRC, f), wd)
    return wd


def test_preamble_model_is_bit_identical_to_the_reference(built, tmp_path):
    from gapprober_b200 import build
    wd = _workdir(tmp_path)
    path = build.learn_model(wd)
    want = gzip.open(os.path.join(SRC, "fb_model.bin.gz")).read()
    assert open(path, "rb").read() == want
    assert open(os.path.join(wd, "mappedInserts.txt")).read() == open(os.path.join(SRC, "mappedInserts.txt")).read()


def test_preamble_prints_the_threshold_line(built, tmp_path):
    from gapprober_b200 import FigbirdModel, build
    wd = _workdir(tmp_path)
    env = dict(os.environ, FIGBIRD_PREAMBLE_ONLY="1", FIGBIRD_MODEL_OUT="m.bin")
    out = subprocess.run([build.build_host(), "scaffolds.fa", " 0", "250"], cwd=wd, env=env, stdout=subprocess.PIPE, text=True, check=True).stdout
    m = FigbirdModel.load(os.path.join(wd, "m.bin"))
    assert out.strip() == f"insertThresholdMin = {m.thr_min}...insertThresholdMax = {m.thr_max}"   # figbird.cpp:2744


def test_threaded_preamble_equals_serial_and_reference(built, tmp_path):
    """Large myout.sam cut into chunks scanned by several host threads: the model must be byte-identical to the
    single-thread scan and to the reference's (both passes only add integer counts).  Every pair is written three
    times in a row (multi-hit groups, figbird.cpp:1268-1280) so that the pass-2 cuts have groups to respect."""
    import refrun
    from gapprober_b200 import build, synth
    wd = str(tmp_path / "big")
    synth.generate(wd, ["--genome", "400000", "--scaffolds", "2", "--gap-list", "120,300", "--cov", "45", "--ins-mean", "200", "--ins-sd", "25",
                        "--err", "0.01", "--nrate", "0.002"], seed=51)
    src = open(os.path.join(wd, "myout.sam")).read().splitlines(keepends=True)
    with open(os.path.join(wd, "myout.sam"), "w") as f:
        for k in range(0, len(src), 2):
            f.write((src[k] + src[k + 1]) * 3)
    assert os.path.getsize(os.path.join(wd, "myout.sam")) > (8 << 20)
    models = {}
    for threads in ("1", "3", "8"):
        env = dict(os.environ, FIGBIRD_PREAMBLE_ONLY="1", FIGBIRD_MODEL_OUT=f"m{threads}.bin", FIGBIRD_THREADS=threads)
        subprocess.run([build.build_host(), "scaffolds.fa", " 0", "250"], cwd=wd, env=env, check=True, stdout=subprocess.DEVNULL)
        models[threads] = (open(os.path.join(wd, f"m{threads}.bin"), "rb").read(), open(os.path.join(wd, "mappedInserts.txt")).read())
    assert models["1"] == models["3"] == models["8"]
    if refrun.have_ref():
        for k in range(2):                                           # preamble is what is compared: give the gaps no records
            open(os.path.join(wd, "Gaps", f"gaps_{k}.sam"), "w").close()
        subprocess.run([refrun.REF_DUMP, "scaffolds.fa", " 0", "250"], cwd=wd, check=True, stdout=subprocess.DEVNULL)
        assert open(os.path.join(wd, "fb_model.bin"), "rb").read() == models["8"][0]
        assert open(os.path.join(wd, "mappedInserts.txt")).read() == models["8"][1]


def test_lines_without_md_or_ih_tags_are_not_chunked(built, tmp_path):
    """A line without an MD or IH tag reuses the previous line's values (sticky state, as in the reference's second pass,
    figbird.cpp:1151, 1204).  A chunk that started at such a line would lose that state, so the threaded scan must fall
    back to one range and give the serial model.  (No comparison with the reference here: its first pass reads
    uninitialised stack locals for such lines, figbird.cpp:835-838, 883 — the splitter never writes them.)"""
    from gapprober_b200 import build, synth
    wd = str(tmp_path / "tagless")
    synth.generate(wd, ["--genome", "400000", "--scaffolds", "2", "--gap-list", "120,300", "--cov", "45", "--ins-mean", "200", "--ins-sd", "25",
                        "--err", "0.01"], seed=52)
    src = open(os.path.join(wd, "myout.sam")).read().splitlines(keepends=True)
    with open(os.path.join(wd, "myout.sam"), "w") as f:
        for rep in range(3):
            for k, ln in enumerate(src):
                if k % 997 == 5:                              # replace the two trailing tag columns (MD:Z:..., IH:i:n)
                    ln = "\t".join(ln.rstrip("\n").split("\t")[:7] + ["XX:i:0"]) + "\n"
                f.write(ln)
    assert os.path.getsize(os.path.join(wd, "myout.sam")) > (8 << 20)
    models = {}
    for threads in ("1", "8"):
        env = dict(os.environ, FIGBIRD_PREAMBLE_ONLY="1", FIGBIRD_MODEL_OUT=f"m{threads}.bin", FIGBIRD_THREADS=threads)
        subprocess.run([build.build_host(), "scaffolds.fa", " 0", "250"], cwd=wd, env=env, check=True, stdout=subprocess.DEVNULL)
        models[threads] = open(os.path.join(wd, f"m{threads}.bin"), "rb").read()
    assert models["1"] == models["8"]
