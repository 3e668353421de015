import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
from epi_mcmc_b200.model import Model, load_dataset
import bench
ds=load_dataset('A300'); M=Model(ds, substeps=4)
n=int(sys.argv[1]) if len(sys.argv)>1 else 65536
mn,mx,init=M.df_params['min'],M.df_params['max'],M.df_params['init']
theta=bench.make_theta('A300', n, 1234, mn, mx, init)
th=torch.from_numpy(np.ascontiguousarray(theta.T)).cuda(); ll=torch.empty(n,dtype=torch.float64,device='cuda'); lp=torch.empty_like(ll); st=torch.empty(n,dtype=torch.int32,device='cuda')
s=torch.cuda.Stream()
M.set_timing(True)
k=np.zeros(4)
for i in range(8):
    M.eval_device(th.data_ptr(), n, ll.data_ptr(), lp.data_ptr(), st.data_ptr(), s.cuda_stream)
    if i>=3: k+=np.array(M.last_kernel_ms())
k/=5
print('n',n,'kernel ms ode1 %.3f mid %.3f ode2 %.3f score %.3f total %.3f  evals/s %.3e'%(k[0],k[1],k[2],k[3],k.sum(), n/k.sum()*1e3), 'finite', int(torch.isfinite(ll).sum()))
