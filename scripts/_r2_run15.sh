set -x
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "marginalize or golden or pass2_restart or both_ode" 2>&1 | grep -v "^$" | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2_c.json 2> gpurun_out/r2_bench_n2_c.err; echo rc=$?; tail -5 gpurun_out/r2_bench_n2_c.err
