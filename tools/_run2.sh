mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multiproc.py tests/test_gpu_dropin.py tests/test_gpu_record_shard.py -m gpu -x -q > gpurun_out/r2q_pytest_n2.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2q_pytest_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2q_bench_n2.json 2> gpurun_out/r2q_bench_n2.err; echo "bench n2 rc=$?"
tail -c 1500 gpurun_out/r2q_bench_n2.json
