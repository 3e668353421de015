python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "I290 or lane" 2>&1 | grep -v "^$" | tail -3
B="--steps 1 --warmup 1 --no-sampler --no-cpu-baseline --no-secondary"
ncu --set full --clock-control none --import-source on -k regex:"k_score|k_ode|k_mid" -c 12 -o gpurun_out/r2_ncu_I290_post -f python bench.py $B --workload I290 --chains 65536 --substeps 4 --theta posterior > gpurun_out/ncu_I290_post.log 2>&1
python profiles/extract_ncu_constants.py gpurun_out/r2_ncu_I290_post.ncu-rep I290 65536 4 posterior > /dev/null
python profiles/summarize_ncu.py gpurun_out/r2_ncu_I290_post.ncu-rep > gpurun_out/r2_ncu_I290_post.txt 2>/dev/null
rm -f gpurun_out/r2_ncu_I290_post.ncu-rep
cp profiles/ncu_constants.json gpurun_out/ncu_constants.json
