B="--steps 1 --warmup 1 --no-sampler --no-cpu-baseline --no-secondary"
ncu --set full --clock-control none --import-source on -k regex:"k_score|k_ode|k_mid" -c 12 -o gpurun_out/r2_ncu_X300 -f python bench.py $B --workload X300 --chains 65536 --substeps 4 --theta narrow > gpurun_out/ncu_X300.log 2>&1
python profiles/extract_ncu_constants.py gpurun_out/r2_ncu_X300.ncu-rep X300 65536 4 narrow > /dev/null
python profiles/summarize_ncu.py gpurun_out/r2_ncu_X300.ncu-rep > gpurun_out/r2_ncu_X300.txt 2>/dev/null
rm -f gpurun_out/r2_ncu_X300.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:"k_score|k_ode|k_mid" -c 8 -o gpurun_out/r2_ncu_A300_8192 -f python bench.py $B --workload A300 --chains 8192 --substeps 4 --theta posterior > gpurun_out/ncu_A300_8192.log 2>&1
python profiles/summarize_ncu.py gpurun_out/r2_ncu_A300_8192.ncu-rep > gpurun_out/r2_ncu_A300_8192.txt 2>/dev/null
python profiles/source_hotspots.py gpurun_out/r2_ncu_A300_8192.ncu-rep "k_ode_age2" "k_ode_age2ILi2ELb1E" 40 > gpurun_out/r2_c_k_ode_age2_pass2_lines.txt 2>gpurun_out/hs2.err
rm -f gpurun_out/r2_ncu_A300_8192.ncu-rep
cp profiles/ncu_constants.json gpurun_out/ncu_constants.json
