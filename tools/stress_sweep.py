"""Single-gap stress sweep (BASELINE config 5): gap length x coverage x duplicated anchor records, one GPU.

Runs bench.py once per point (device-resident and end-to-end arms, no CPU leg) and writes one JSON line per point.
    python tools/stress_sweep.py [--out gpurun_out/stress_sweep.jsonl] [--points g100_c30_k1,g5000_c300_k10,...]
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT = ["g100_c30_k1", "g100_c300_k1", "g100_c300_k50", "g500_c30_k1", "g500_c100_k10", "g500_c300_k1", "g1000_c30_k1",
           "g1000_c100_k10", "g1000_c300_k1", "g2500_c30_k1", "g2500_c100_k1", "g2500_c300_k10", "g5000_c30_k1", "g5000_c100_k1",
           "g5000_c300_k10"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "stress_sweep.jsonl"))
    ap.add_argument("--points", default=",".join(DEFAULT))
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as out:
        for pt in a.points.split(","):
            cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--config", "cfg5_stress_" + pt, "--steps", str(a.steps),
                   "--warmup", str(a.warmup), "--no-cpu", "--no-check"]
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
            line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ""
            try:
                d = json.loads(line)
                row = {"point": pt, "units": d["work"]["units_this_rank"], "records": d["work"]["records_this_rank"],
                       "likelihoods_per_step": d["work"]["likelihoods_per_step"], "ms_per_step": d["ms_per_step"],
                       "em_kernel_ms": d["roofline"]["kernel_ms_per_step"], "value": d["value"], "e2e": d["e2e"]["value"],
                       "smem_frac": d["smem"]["frac"]}
            except Exception:
                row = {"point": pt, "error": (r.stderr or line)[-400:]}
            out.write(json.dumps(row) + "\n")
            out.flush()
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
