"""diagnostic (not collected): 30 DE-MC iterations at 65,536 chains with recording on, for a launch list under ncu"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from epi_mcmc_b200.model import Model, load_dataset
from epi_mcmc_b200.sampler import Sampler
M = Model(load_dataset("A300"), substeps=4)
S = Sampler(M, 65536, kind="demc", seed=1)
S.record(1, 64)
S.run(30)
print(S.stats())
