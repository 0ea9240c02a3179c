mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_parity.py tests/test_gpu_edge.py -m gpu -x -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2o_pytest.log
for cfg in cfg2_rhodobacter cfg4_chr14; do
for k in "" "FB_E_REC_MAJOR=1"; do
  sc=strong; [ $cfg = cfg2_rhodobacter ] && sc=weak
  env $k timeout 600 python bench.py --config $cfg --scaling $sc --no-cpu --no-check --steps 5 > gpurun_out/r2o_${cfg}_$k.json 2> gpurun_out/r2o_${cfg}_$k.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2o_${cfg}_$k.json").read().strip().splitlines()[-1])
    print("$cfg [$k] step_ms %.4f em_ms %.4f value %.4e e2e %.4e (%.3f) sha %s" % (d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["value"]/d["value"], d["result"]["filled_sha256"][:12]))
except Exception as e:
    print("$cfg [$k] FAILED", e); print(open("gpurun_out/r2o_${cfg}_$k.err").read()[-600:])
PY
done
done
