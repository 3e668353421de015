for l in 1 4; do
EPI_ODE_LANES=$l ncu --set full --clock-control none --import-source on -k regex:k_ode -c 6 -o gpurun_out/r2_ode_small_l$l -f python bench.py --chains 8192 --steps 1 --warmup 1 --no-sampler --no-cpu-baseline > gpurun_out/ncu_small_l$l.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -3
