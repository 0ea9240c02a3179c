mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2r_bench_n8.json 2> gpurun_out/r2r_bench_n8.err; echo "bench n8 rc=$?"
tail -c 300 gpurun_out/r2r_bench_n8.err
