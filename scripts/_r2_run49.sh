P='import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["kernels"].items()})'
for rep in 1 2; do
for v in old new; do
cp scripts/_libepimcmc_$v.so epi_mcmc_b200/csrc/libepimcmc.so
echo "== NE120 4096 $v"; python bench.py --workload NE120 --chains 4096 --theta narrow --steps 400 --warmup 10 --no-sampler --no-cpu-baseline --no-secondary 2>/dev/null | python -c "$P"
done
done
for v in old new; do
cp scripts/_libepimcmc_$v.so epi_mcmc_b200/csrc/libepimcmc.so
echo "== S120 65536 $v"; python bench.py --workload S120 --chains 65536 --theta narrow --steps 100 --warmup 5 --no-sampler --no-cpu-baseline --no-secondary 2>/dev/null | python -c "$P"
done
