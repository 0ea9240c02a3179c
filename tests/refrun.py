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
