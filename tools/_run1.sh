set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2h_pytest.log
for k in "" "FB_NO_PREFIX=1"; do
  env $k python bench.py --config cfg2_rhodobacter --scaling weak --no-cpu --no-check --steps 10 > gpurun_out/r2h_cfg2_$k.json 2> gpurun_out/r2h_cfg2_$k.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r2h_cfg2_$k.json").read().strip().splitlines()[-1])
print("cfg2 [$k] step_ms %.4f em_ms %.4f value %.4e e2e %.4e sha %s" % (d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["value"], d["e2e"]["value"], d["result"]["filled_sha256"][:12]))
PY
done
for k in "" "FB_NO_PREFIX=1"; do
  env $k python bench.py --no-cpu --no-check --steps 5 > gpurun_out/r2h_cfg4_$k.json 2> gpurun_out/r2h_cfg4_$k.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r2h_cfg4_$k.json").read().strip().splitlines()[-1])
print("cfg4 [$k] step_ms %.4f em_ms %.4f value %.4e e2e %.4e sha %s" % (d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["value"], d["e2e"]["value"], d["result"]["filled_sha256"][:12]))
PY
done
