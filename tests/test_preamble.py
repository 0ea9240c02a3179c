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
