B="--steps 1 --warmup 1 --no-sampler --no-cpu-baseline --no-secondary"
ncu --set full --clock-control none --import-source on -k regex:"k_ode_age2" -c 2 -o gpurun_out/r2_ncu_p2 -f \
    python bench.py $B --workload A300 --chains 65536 --substeps 4 --theta posterior > gpurun_out/ncu_p2.log 2>&1
python profiles/source_hotspots.py gpurun_out/r2_ncu_p2.ncu-rep "k_ode_age2" "k_ode_age2ILi2ELb1ELi4E" 70 > gpurun_out/r2_e_k_ode_age2_pass2_lines.txt 2>gpurun_out/hs.err
python profiles/source_hotspots.py gpurun_out/r2_ncu_p2.ncu-rep "k_ode_age2" "k_ode_age2ILi2ELb0ELi4E" 40 > gpurun_out/r2_e_k_ode_age2_pass1_lines.txt 2>>gpurun_out/hs.err
rm -f gpurun_out/r2_ncu_p2.ncu-rep
