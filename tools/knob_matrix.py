"""Run bench.py on a few single-gap workloads under different planning knobs; one line per (workload, knobs)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORK = sys.argv[1].split(",") if len(sys.argv) > 1 else ["cfg1_single_gap", "cfg5_stress_g100_c30_k1", "cfg5_stress_g500_c300_k1", "cfg5_stress_g500_c30_k1", "cfg5_stress_g1000_c30_k1"]
KNOBS = [{}, {"FB_NO_GRAPH": "1"}, {"FB_TWO_CALL": "1"}, {"FB_NO_FEW": "1"}]
for w in WORK:
    for k in KNOBS:
        env = dict(os.environ, **k)
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--config", w, "--no-cpu", "--no-check", "--steps", "5"], env=env,
                           capture_output=True, text=True)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
            print(w, k, "step_ms %.4f em_ms %.4f value %.3e" % (d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["value"]), flush=True)
        except Exception:
            print(w, k, "FAILED", r.stderr[-300:], flush=True)
