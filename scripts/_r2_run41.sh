N=${N:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/r2_bench_n${N}_k.json 2> gpurun_out/r2_bench_n${N}_k.err; echo rc=$?; tail -3 gpurun_out/r2_bench_n${N}_k.err
