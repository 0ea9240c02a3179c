#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE.

Runs only where /root/reference exists (the build container): oracle/Makefile builds the unmodified reference
and an instrumented copy into oracle/_ref/, fb_synth makes the inputs, the reference's own splitter (main.out)
and figbird (figbird_dump.out) produce the expected outputs.  What is committed is small: per case one .npz with
the scaffold windows around the gaps, the gap records, the model tables and the reference's answers.

  python tests/golden/make_golden.py            # regenerate everything

Coordinates in a fixture are shifted so that every gap sits at offset max_distance of its own window
(the hot path only ever looks max_distance bases either side of the gap; all its arithmetic is relative
to gapStart), which keeps the fixtures a few hundred KB in total.
"""
import gzip
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib as O  # noqa: E402
import refrun  # noqa: E402
from gapprober_b200 import synth  # noqa: E402

CASES = {
    # name: (seed, f_mode, fb_synth args)
    "cfg1_500bp_30x": (1, 0, ["--genome", "100000", "--gap-list", "500", "--cov", "30"]),
    "lib180_100bp_60x": (4, 0, ["--genome", "60000", "--gap-list", "100", "--cov", "60", "--ins-mean", "180", "--ins-sd", "20"]),
    "multi_gap_err1pct": (8, 0, ["--genome", "120000", "--scaffolds", "2", "--gap-list", "50,150,250,400,99", "--cov", "50",
                                 "--ins-mean", "180", "--ins-sd", "20", "--err", "0.01"]),
    "tiny_gaps_N_multihit": (9, 0, ["--genome", "80000", "--gap-list", "300,20,7", "--cov", "40", "--ins-mean", "300", "--ins-sd", "30",
                                    "--err", "0.02", "--nrate", "0.01", "--multihit", "3", "--len-noise", "10"]),
    "plateau_1000bp": (6, 0, ["--genome", "100000", "--gap-list", "1000", "--cov", "40"]),
    "logmode_f1": (4, 1, ["--genome", "60000", "--gap-list", "100,260", "--cov", "60", "--ins-mean", "180", "--ins-sd", "20", "--err", "0.01"]),
}


def parse_counts(stdout):
    out = []
    for ln in stdout.splitlines():
        if ln.startswith("Total_Reads"):
            a, b = ln.split("\t")
            out.append((int(a.split("=")[1]), int(b.split("=")[1])))
    return out


def make_case(name, seed, f_mode, args, keep_sam=False):
    wd = tempfile.mkdtemp(prefix="fbgold_")
    try:
        synth.generate(wd, args, seed=seed, direct=False, sam=True)
        os.makedirs(os.path.join(wd, "Gaps"), exist_ok=True)
        subprocess.run([refrun.REF_MAIN, "scaffolds.fa", "250"], cwd=wd, check=True, stdout=subprocess.DEVNULL)   # the reference's own splitter
        ref = refrun.run_reference(wd, dump_probs=True, f_mode=f_mode)
        m = ref["model"]
        D = m.max_distance
        _, seqs = O.read_fasta(os.path.join(wd, "scaffolds.fa"))
        gaps = O.read_gap_info(os.path.join(wd, "gapInfo.txt"))
        counts = parse_counts(ref["stdout"])
        z = dict(
            f_mode=np.int32(f_mode), max_read_len=np.int32(m.max_read_len), thr_min=np.int32(m.thr_min), thr_max=np.int32(m.thr_max),
            cutoff=np.int32(m.cutoff), max_distance=np.int32(D), ild_len=np.int32(m.ild.size),
            ild_head=m.ild[:m.thr_max + 1].copy(),   # the hot path reads only [thr_min, thr_max]
            error_type=m.error_type, error_pos=m.error_pos, in_pos=m.in_pos, del_pos=m.del_pos, n_gaps=np.int32(len(gaps)))
        for gi, (c, s, l) in enumerate(gaps):
            recs = O.read_gap_file(os.path.join(wd, "Gaps", f"gaps_{gi}.sam"))
            d = ref["dump"][f"Gaps/gaps_{gi}.sam"]
            shift = s - D
            z[f"g{gi}_window"] = np.frombuffer(seqs[c][s - D:s + l + D], dtype=np.uint8).copy()
            z[f"g{gi}_orig"] = np.int32(l)
            z[f"g{gi}_meta"] = np.array([[p - shift, l1, s1] for p, l1, s1, _ in recs], dtype=np.int32).reshape(-1, 3)
            z[f"g{gi}_seqs"] = np.frombuffer(b"\n".join(q for _, _, _, q in recs), dtype=np.uint8).copy()
            a, b = O.gap_range(l)
            z[f"g{gi}_liks"] = np.array([d["est"][e] for e in range(a, b + 1)])
            z[f"g{gi}_best_est"] = np.int32(d["best"][0])
            z[f"g{gi}_best_lik"] = np.float64(d["best"][1])
            z[f"g{gi}_best_string"] = np.frombuffer(d["best"][2].encode(), dtype=np.uint8).copy()
            z[f"g{gi}_filled"] = np.frombuffer(ref["gapout"][gi]["filled"].encode(), dtype=np.uint8).copy()
            z[f"g{gi}_counts"] = np.array(counts[gi], dtype=np.int32)
            z[f"g{gi}_rows"] = d["rows"][d["best"][0]]     # probsGap | countsGap of the winning estimate, after the last iteration
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **z)
        print(f"{name}: {len(gaps)} gaps -> {os.path.getsize(os.path.join(HERE, name + '.npz')) // 1024} KB")
    finally:
        shutil.rmtree(wd, ignore_errors=True)


def make_preamble_case():
    """A small myout.sam (splitter dialect) + stat.txt + scaffold lengths, with the model the reference learns from it."""
    wd = tempfile.mkdtemp(prefix="fbgold_")
    try:
        synth.generate(wd, ["--genome", "40000", "--scaffolds", "2", "--gap-list", "120,60", "--cov", "12", "--ins-mean", "180", "--ins-sd", "20",
                            "--err", "0.02", "--nrate", "0.003"], seed=21, direct=False, sam=True)
        os.makedirs(os.path.join(wd, "Gaps"), exist_ok=True)
        subprocess.run([refrun.REF_MAIN, "scaffolds.fa", "250"], cwd=wd, check=True, stdout=subprocess.DEVNULL)
        refrun.run_reference(wd, dump_probs=False)
        dst = os.path.join(HERE, "preamble")
        os.makedirs(dst, exist_ok=True)
        for f in ("myout.sam", "scaffolds.fa", "fb_model.bin"):
            with open(os.path.join(wd, f), "rb") as src, gzip.GzipFile(os.path.join(dst, f + ".gz"), "wb", mtime=0) as out:
                out.write(src.read())
        for f in ("stat.txt", "gapInfo.txt", "mappedInserts.txt"):
            shutil.copy(os.path.join(wd, f), os.path.join(dst, f))
        print("preamble:", sum(os.path.getsize(os.path.join(dst, f)) for f in os.listdir(dst)) // 1024, "KB")
    finally:
        shutil.rmtree(wd, ignore_errors=True)


def main():
    if not os.path.isdir("/root/reference"):
        raise SystemExit("make_golden.py needs /root/reference (it runs the reference to obtain the answers)")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all", "ref"])
    for name, (seed, f_mode, args) in CASES.items():
        make_case(name, seed, f_mode, args)
    make_preamble_case()


if __name__ == "__main__":
    main()
