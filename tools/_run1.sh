set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_pipe.py tests/test_gpu_golden.py tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_gpu_fuzz.py tests/test_gpu_dropin.py -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2i_pytest.log
for cfg in cfg2_rhodobacter cfg4_chr14; do
for k in "" "FB_NO_PIPE=1" "FB_NO_T0=1"; do
  sc=strong; [ $cfg = cfg2_rhodobacter ] && sc=weak
  env $k python bench.py --config $cfg --scaling $sc --no-cpu --no-check --steps 5 > gpurun_out/r2i_${cfg}_$k.json 2> gpurun_out/r2i_${cfg}_$k.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r2i_${cfg}_$k.json").read().strip().splitlines()[-1])
print("$cfg [$k] step_ms %.4f em_ms %.4f value %.4e e2e %.4e (%.3f) sha %s" % (d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["value"]/d["value"], d["result"]["filled_sha256"][:12]))
PY
done
done
