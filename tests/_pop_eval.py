"""diagnostic (not collected): evaluate a dumped sampler population (tests/_pop_dump.py) three times and print the
kernel times; used under ncu to profile pass 2 on a posterior-spread theta cloud"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from epi_mcmc_b200.model import Model, load_dataset
theta = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "_pop_A300.npy"))
M = Model(load_dataset("A300"), substeps=4)
M.set_timing(True)
for _ in range(3):
    l, p, s = M.calclogpost_batch(theta)
print(np.round(M.last_kernel_ms(), 3), int((s != 0).sum()))
