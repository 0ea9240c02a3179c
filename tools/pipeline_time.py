"""Wall time of the whole figbird stage of GapProber on one generated data set, reference binaries vs this repo:
    result.sam --(main.out | fb_split)--> myout.sam, Gaps/ --(figbird.out)--> gapout.txt, filledContigs.fa
and the in-memory route (FIGBIRD_FROM_SAM: one process, nothing but gapInfo.txt / stat.txt in between).
The outputs are compared byte for byte.  TEST/measurement tool: runs oracle/_ref (the unmodified reference, built by
oracle/Makefile); the product never does.
    python tools/pipeline_time.py [--config cfg2_rhodobacter] [--seed 2] [--out gpurun_out/pipeline.json]
"""
import argparse
import filecmp
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref")


def timed(cmd, cwd, env=None):
    t = time.perf_counter()
    subprocess.run(cmd, cwd=cwd, env=env, check=True, stdout=subprocess.DEVNULL)
    return time.perf_counter() - t


def main():
    from gapprober_b200 import build, synth
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg2_rhodobacter")
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--out", default="")
    ap.add_argument("--skip-reference-em", action="store_true", help="time the reference's splitter and preamble only")
    a = ap.parse_args()
    top = tempfile.mkdtemp(prefix="fb_pipeline_")
    base = os.path.join(top, "base")
    synth.generate(base, synth.config_args(a.config), seed=a.seed, direct=False, sam=True)
    shutil.rmtree(os.path.join(base, "Gaps"), ignore_errors=True)
    for f in ("myout.sam", "gapInfo.txt", "stat.txt"):
        if os.path.exists(os.path.join(base, f)):
            os.remove(os.path.join(base, f))
    res = {"config": a.config, "seed": a.seed, "sam_mb": os.path.getsize(os.path.join(base, "result.sam")) / 1e6}
    host = build.build_host()
    split = os.path.join(build.BIN, "fb_split")
    dirs = {}
    for name in ("ref", "ours", "mem"):
        dirs[name] = os.path.join(top, name)
        shutil.copytree(base, dirs[name])
        if name != "mem":
            os.makedirs(os.path.join(dirs[name], "Gaps"), exist_ok=True)
    # reference: main.out, then figbird.out (preamble + EM + stitch in one process; the preamble alone via empty gap files)
    res["ref_split_s"] = timed([os.path.join(REF, "main.out"), "scaffolds.fa", "250"], dirs["ref"])
    pre = os.path.join(top, "ref_pre")
    shutil.copytree(dirs["ref"], pre)
    for f in os.listdir(os.path.join(pre, "Gaps")):
        open(os.path.join(pre, "Gaps", f), "w").close()
    res["ref_preamble_s"] = timed([os.path.join(REF, "figbird.out"), "scaffolds.fa", " 0", "250"], pre)
    if not a.skip_reference_em:
        res["ref_figbird_s"] = timed([os.path.join(REF, "figbird.out"), "scaffolds.fa", " 0", "250"], dirs["ref"])
    # this repo, file route
    res["ours_split_s"] = timed([split, "scaffolds.fa", "250"], dirs["ours"])
    res["ours_preamble_s"] = timed([host, "scaffolds.fa", " 0", "250"], dirs["ours"], dict(os.environ, FIGBIRD_PREAMBLE_ONLY="1"))
    res["ours_figbird_s"] = timed([host, "scaffolds.fa", " 0", "250"], dirs["ours"])
    # this repo, in-memory route
    res["ours_from_sam_s"] = timed([host, "scaffolds.fa", " 0", "250"], dirs["mem"], dict(os.environ, FIGBIRD_FROM_SAM="result.sam"))
    same = {}
    for f in ("gapout.txt", "filledContigs.fa", "mappedInserts.txt"):
        same[f] = filecmp.cmp(os.path.join(dirs["ours"], f), os.path.join(dirs["mem"], f), shallow=False)
        if not a.skip_reference_em:
            same[f] = same[f] and filecmp.cmp(os.path.join(dirs["ours"], f), os.path.join(dirs["ref"], f), shallow=False)
    res["outputs_identical"] = same
    if not a.skip_reference_em:
        res["ref_total_s"] = res["ref_split_s"] + res["ref_figbird_s"]
        res["speedup_file_route"] = res["ref_total_s"] / (res["ours_split_s"] + res["ours_figbird_s"])
        res["speedup_in_memory"] = res["ref_total_s"] / res["ours_from_sam_s"]
    line = json.dumps(res)
    print(line)
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        open(a.out, "w").write(line + "\n")
    shutil.rmtree(top, ignore_errors=True)


if __name__ == "__main__":
    main()
