import sys; sys.path.insert(0,'/root/repo')
import mcpower_b200 as M
from mcpower_b200 import spec as S
s=S.sweep_spec(); grid=list(range(20,1001,20)); seed=M.splitmix64(2137)
M.run_batch(s,grid,4096,seed)
r=M.run_batch(s,grid,1<<18,seed)
print(r.kernel_ms, (1<<18)*50/(r.kernel_ms*1e-3))
