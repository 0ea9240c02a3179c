"""Seeded synthetic GapProber workloads (driver for tools/fb_synth.cpp; bowtie2/wgsim are not in the image).

The named configurations are the ones BASELINE.json lists.
"""
import os
import subprocess

from . import build as _build

# name -> fb_synth arguments
CONFIGS = {
    # 100 kb scaffold, one 500 bp gap, error-free 100 bp pairs at 30x
    "cfg1_single_gap": ["--genome", "100000", "--gap-list", "500", "--cov", "30", "--ins-mean", "500", "--ins-sd", "50"],
    # R. sphaeroides scale: 4.6 Mb in 20 scaffolds, ~200 gaps of 50-500 bp, 1 % errors, 50x
    "cfg2_rhodobacter": ["--genome", "4600000", "--scaffolds", "20", "--gaps", "200", "--gap-min", "50", "--gap-max", "500",
                         "--cov", "50", "--err", "0.01", "--ins-mean", "180", "--ins-sd", "20"],
    # S. aureus scale: 2.9 Mb, ~100 gaps
    "cfg3_staph": ["--genome", "2900000", "--scaffolds", "10", "--gaps", "100", "--gap-min", "50", "--gap-max", "500",
                   "--cov", "50", "--err", "0.01", "--ins-mean", "180", "--ins-sd", "20"],
    # human chr14 scale: 88 Mb, ~5000 gaps, 40x; the model is learned from a 400k-pair sample of the proper pairs
    "cfg4_chr14": ["--genome", "88000000", "--scaffolds", "40", "--gaps", "5000", "--gap-min", "50", "--gap-max", "500",
                   "--cov", "40", "--err", "0.01", "--ins-mean", "180", "--ins-sd", "20", "--model-pairs", "400000"],
    # single-gap stress sweep, two corners: long gap x high coverage x duplicated anchors (bowtie2 -k style)
    "cfg5_stress_g5000_c300_k10": ["--genome", "200000", "--gap-list", "5000", "--cov", "300", "--err", "0.01", "--ins-mean", "500",
                                   "--ins-sd", "50", "--multihit", "10", "--model-pairs", "200000"],
    "cfg5_stress_g100_c300_k50": ["--genome", "200000", "--gap-list", "100", "--cov", "300", "--err", "0.01", "--ins-mean", "180",
                                  "--ins-sd", "20", "--multihit", "50", "--model-pairs", "200000"],
}


def config_args(name):
    """fb_synth arguments of a named configuration.  Besides the entries of CONFIGS, any point of the single-gap
    stress sweep (BASELINE config 5) can be named cfg5_stress_g<gap>_c<coverage>_k<copies>."""
    if name in CONFIGS:
        return list(CONFIGS[name])
    import re
    m = re.fullmatch(r"cfg5_stress_g(\d+)_c(\d+)_k(\d+)", name)
    if not m:
        raise KeyError(f"unknown workload {name!r}")
    g, cov, k = (int(x) for x in m.groups())
    lib = ["--ins-mean", "180", "--ins-sd", "20"] if g < 500 else ["--ins-mean", "500", "--ins-sd", "50"]
    args = ["--genome", "200000", "--gap-list", str(g), "--cov", str(cov), "--err", "0.01"] + lib
    if k > 1:
        args += ["--multihit", str(k)]
    return args + ["--model-pairs", "200000"]


def fb_synth_path():
    exe = os.path.join(_build.BIN, "fb_synth")
    if not os.path.exists(exe):
        _build.build_tools()
    return exe


def generate(out_dir, args, seed=1, direct=True, sam=False):
    """Run fb_synth into out_dir.  direct: also write the splitter's outputs (myout.sam, Gaps/, gapInfo.txt, stat.txt)."""
    os.makedirs(os.path.join(out_dir, "Gaps"), exist_ok=True)
    cmd = [fb_synth_path(), "--out", out_dir, "--seed", str(seed)] + [str(a) for a in args]
    if direct:
        cmd.append("--direct")
    if not sam:
        cmd.append("--no-sam")
    subprocess.run(cmd, check=True, stderr=subprocess.PIPE)
    return out_dir
