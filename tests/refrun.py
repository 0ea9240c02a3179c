"""Run the REFERENCE binaries (oracle/_ref, built from /root/reference by oracle/Makefile) on a generated
working directory.  TEST INFRASTRUCTURE: the prebuilt binaries travel to the GPU box; /root/reference does not
and is never read here."""
import os
import subprocess

import oracle_lib as O

REF_DUMP = os.path.join(O.REF_DIR, "figbird_dump.out")
REF_FIGBIRD = os.path.join(O.REF_DIR, "figbird.out")
REF_MAIN = os.path.join(O.REF_DIR, "main.out")


def have_ref():
    return all(os.path.exists(p) for p in (REF_DUMP, REF_FIGBIRD, REF_MAIN))


def run_reference(wd, dump_probs=True, f_mode=0):
    """Run the instrumented reference figbird in wd (needs scaffolds.fa, stat.txt, myout.sam, gapInfo.txt, Gaps/).
    Returns dict(model=..., dump=..., gapout=[rows], stdout=str)."""
    env = dict(os.environ)
    if dump_probs:
        env["FB_DUMP_PROBS"] = "1"
    else:
        env.pop("FB_DUMP_PROBS", None)
    out = subprocess.run([REF_DUMP, "scaffolds.fa", f" {f_mode}", "250"], cwd=wd, env=env, check=True,
                         stdout=subprocess.PIPE).stdout.decode()
    model = O.Model.load(os.path.join(wd, "fb_model.bin"))
    model.f_mode = f_mode
    dump = O.parse_dump(os.path.join(wd, "fb_dump.txt"))
    gapout = []
    with open(os.path.join(wd, "gapout.txt")) as f:
        for ln in f:
            p = ln.rstrip("\n").split("\t")
            gapout.append(dict(gap=int(p[0]), ctg=int(p[1]), start=int(p[2]), orig=int(p[3]), filled_len=int(p[4]),
                               filled=p[5] if len(p) > 5 else ""))
    return dict(model=model, dump=dump, gapout=gapout, stdout=out)


def run_reference_on_gaps(wd, gap_ids, procs, tmp, f_mode=0):
    """The instrumented reference on a SUBSET of the gaps of wd, spread over `procs` concurrent processes (each gets a
    working directory with its share of the gaps renumbered from 0 and the big inputs symlinked).  Returns
    {gap_id: dict(est={estimate: lik}, best=(estimate, lik, bestString), filled=str)}."""
    import shutil
    lines = open(os.path.join(wd, "gapInfo.txt")).read().splitlines()
    shares = [list(gap_ids[k::procs]) for k in range(procs)]
    shares = [s for s in shares if s]
    dirs = []
    for k, share in enumerate(shares):
        d = os.path.join(tmp, f"refpart{k}")
        if os.path.isdir(d):
            shutil.rmtree(d)
        os.makedirs(os.path.join(d, "Gaps"))
        for f in ("scaffolds.fa", "myout.sam", "stat.txt"):
            os.symlink(os.path.join(wd, f), os.path.join(d, f))
        with open(os.path.join(d, "gapInfo.txt"), "w") as f:
            for j, gi in enumerate(share):
                f.write(lines[gi] + "\n")
                os.symlink(os.path.join(wd, "Gaps", f"gaps_{gi}.sam"), os.path.join(d, "Gaps", f"gaps_{j}.sam"))
        dirs.append(d)
    env = dict(os.environ)
    env.pop("FB_DUMP_PROBS", None)
    ps = [subprocess.Popen([REF_DUMP, "scaffolds.fa", f" {f_mode}", "250"], cwd=d, env=env, stdout=subprocess.DEVNULL) for d in dirs]
    rcs = [p.wait() for p in ps]
    out = {}
    for d, share, rc in zip(dirs, shares, rcs):
        dump = O.parse_dump(os.path.join(d, "fb_dump.txt"))
        rows = open(os.path.join(d, "gapout.txt")).read().splitlines()
        # The gap loop closes gapout.txt (figbird.cpp:2858) before the stitcher re-reads it; on a gap list that is only a
        # share of the scaffolds' gaps the reference's stitcher can die in malloc afterwards.  Its output is not used
        # here: a non-zero exit only counts when the gap loop's own outputs are incomplete.
        if rc != 0 and (len(rows) != len(share) or any(f"Gaps/gaps_{j}.sam" not in dump for j in range(len(share)))):
            raise RuntimeError(f"reference figbird_dump.out failed in {d} (exit {rc})")
        for j, gi in enumerate(share):
            e = dump[f"Gaps/gaps_{j}.sam"]
            p = rows[j].split("\t")
            out[gi] = dict(est=e["est"], best=e["best"], filled=p[5] if len(p) > 5 else "")
    return out
