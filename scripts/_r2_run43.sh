cp scripts/_libepimcmc_cyc.so epi_mcmc_b200/csrc/libepimcmc.so
python bench.py --chains 65536 --steps 2 --warmup 1 --no-sampler --no-cpu-baseline --no-secondary 2>&1 | grep SCORE | tail -12
