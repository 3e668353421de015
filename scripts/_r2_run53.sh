python -m pytest tests/test_gpu_sampler.py -m gpu -x -q 2>&1 | grep -v "^$" | tail -3
python bench.py --chains 8192 --steps 50 --warmup 5 --no-cpu-baseline --no-secondary 2>/dev/null | python -c '
import json,sys; d=json.loads(sys.stdin.read()); s=d["sampler"]; print("8192 chains: step", round(d["ms_per_step"],4), "burnin evals/s", round(s["burnin_evals_per_s"]/1e6,2), "M  recorded", round(s["evals_per_s"]/1e6,2), "M  accept", round(s["accept_rate"],3), "ess", round(s["ess_median"]))'
