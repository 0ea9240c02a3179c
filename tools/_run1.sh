mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2j_pytest.log
for k in "FB_CHUNK=12" "FB_CHUNK=24"; do
  env $k python bench.py --config cfg2_rhodobacter --scaling weak --no-cpu --no-check --steps 5 > gpurun_out/r2j_cfg2_$k.json 2> gpurun_out/r2j_cfg2_$k.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r2j_cfg2_$k.json").read().strip().splitlines()[-1])
print("cfg2 [$k] step_ms %.4f em_ms %.4f value %.4e e2e %.4e (%.3f)" % (d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["value"]/d["value"]))
PY
done
export FB_PROFILE_STEP=1
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-check"
$CMD > gpurun_out/r2j_plain.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2j_launches.csv $CMD > gpurun_out/r2j_ncu1.log 2>&1
echo "launch list rc=$?"
ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/r2j_full $CMD > gpurun_out/r2j_ncu2.log 2>&1
echo "full rc=$?"; ls -la gpurun_out/r2j_full.ncu-rep
