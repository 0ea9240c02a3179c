mkdir -p gpurun_out
python - <<'PY'
from gapprober_b200 import synth, build
build.build_all()
synth.generate('/tmp/wd2', synth.config_args('cfg2_rhodobacter'), seed=2)
PY
cd /tmp/wd2
for k in "A=1" "FIGBIRD_NO_PIPE=1" "A=2" "FIGBIRD_NO_PIPE=1"; do
  echo "=== $k"
  t0=$(date +%s.%N)
  env $k FB_TRACE=1 /root/repo/gapprober_b200/bin/figbird.out scaffolds.fa " 0" 250 2>/tmp/err.txt >/dev/null
  t1=$(date +%s.%N)
  echo "wall $(echo "$t1 - $t0" | bc) s"
  grep -v "bin:\|threads\|upload:  " /tmp/err.txt | tail -40
done
