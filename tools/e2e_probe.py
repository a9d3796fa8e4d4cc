import sys, time, os; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch
from helpers import build_product
from bench import make_batch, workload
seq=build_product("teleop/allegro_hand_right",device=0); opt=seq.optimizer
B=65536
kp,x0=make_batch(opt.robot.kin, workload(seq), B, 1)
kp_p=torch.from_numpy(kp).pin_memory(); x0_p=torch.from_numpy(x0).pin_memory(); out=torch.empty((B,16),dtype=torch.float32).pin_memory()
for _ in range(3): opt.retarget_batch_host(keypoints=kp_p,last_qpos=x0_p,out=out)
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(20): opt.retarget_batch_host(keypoints=kp_p,last_qpos=x0_p,out=out)
torch.cuda.synchronize(); ms=(time.perf_counter()-t)/20*1e3
# raw copy times
d=torch.empty((B,21,3),dtype=torch.float32,device='cuda'); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record(); 
for _ in range(10): d.copy_(kp_p,non_blocking=True)
e1.record(); torch.cuda.synchronize()
print(os.environ.get("DEXR_HOST_CHUNKS"),"e2e ms %.3f -> %.3e f/s ; H2D 16.5MB %.3f ms"%(ms,B/ms*1e3,e0.elapsed_time(e1)/10))
