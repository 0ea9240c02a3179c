mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2p_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2p_bench_ref.json 2> gpurun_out/r2p_bench_ref.err; echo "ref rc=$?"
for cfg in cfg1_single_gap cfg2_rhodobacter cfg3_staph; do
  sc=strong
  timeout 600 python bench.py --config $cfg --scaling $sc --no-cpu --no-check --steps 10 > gpurun_out/r2p_${cfg}.json 2> gpurun_out/r2p_${cfg}.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2p_${cfg}.json").read().strip().splitlines()[-1])
    print("$cfg step_ms %.4f em_ms %.4f value %.4e e2e %.4e (%.3f) smem %.3f" % (d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["value"]/d["value"], d["smem"]["frac"]))
except Exception as e:
    print("$cfg FAILED", e); print(open("gpurun_out/r2p_${cfg}.err").read()[-600:])
PY
done
timeout 1200 python tools/stress_sweep.py --out gpurun_out/r2p_stress_sweep.jsonl > gpurun_out/r2p_stress.log 2>&1; echo "sweep rc=$?"
export FB_PROFILE_STEP=1 FB_SERIAL_BINS=1
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-check"
$CMD > gpurun_out/r2p_plain.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2p_launches.csv $CMD > gpurun_out/r2p_ncu1.log 2>&1
echo "launch list rc=$?"
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:fb_ -o gpurun_out/r2p_full $CMD > gpurun_out/r2p_ncu2.log 2>&1
echo "full rc=$?"; ls -la gpurun_out/r2p_full.ncu-rep
