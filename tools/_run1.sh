mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2s_pytest.log
run() { # name, env, args...
  name=$1; envs=$2; shift 2
  env $envs timeout 600 python bench.py "$@" --no-cpu --no-check --steps 5 > gpurun_out/r2s_$name.json 2> gpurun_out/r2s_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2s_$name.json").read().strip().splitlines()[-1])
    print("$name step_ms %.4f em_ms %.4f value %.4e e2e %.4e (%.3f) smem %.3f units %d" % (d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["value"]/d["value"], d["smem"]["frac"], d["work"]["units_this_rank"]))
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/r2s_$name.err").read()[-600:])
PY
}
run cfg2_f1 "A=1" --config cfg2_rhodobacter --f-mode 1
run cfg4_f1 "A=1" --f-mode 1
run g5000_gt "FB_NO_COMPACT=1 FB_TRACE=1" --config cfg5_stress_g5000_c100_k1
run g5000_smem "FB_TRACE=1" --config cfg5_stress_g5000_c100_k1
grep "bin:" gpurun_out/r2s_g5000_gt.err | sort | uniq -c | head -5
timeout 900 python tools/pipeline_time.py --out gpurun_out/r2s_pipeline.json > gpurun_out/r2s_pipeline.log 2>&1; echo "pipeline rc=$?"; tail -c 1500 gpurun_out/r2s_pipeline.json
