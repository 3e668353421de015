set -x
ncu --set full --clock-control none --import-source on -k regex:"k_ode_age4" -c 2 -o gpurun_out/r2_ncu_age4 -f python bench.py --chains 8192 --steps 1 --warmup 1 --no-sampler --no-cpu-baseline --no-secondary > gpurun_out/ncu_age4.log 2>&1
python profiles/summarize_ncu.py gpurun_out/r2_ncu_age4.ncu-rep > gpurun_out/r2_ncu_age4.txt 2>/dev/null
python profiles/source_hotspots.py gpurun_out/r2_ncu_age4.ncu-rep "k_ode_age4<\(int\)2, \(bool\)1" "k_ode_age4ILi2ELb1ELi4E" 60 > gpurun_out/r2_age4_pass2_lines.txt 2>gpurun_out/hs.err
rm -f gpurun_out/r2_ncu_age4.ncu-rep
