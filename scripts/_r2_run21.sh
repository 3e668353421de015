set -x
python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_e.json 2> gpurun_out/r2_bench_n1_e.err; echo rc=$?; tail -3 gpurun_out/r2_bench_n1_e.err
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_ref_e.json 2>/dev/null; echo rc=$?
