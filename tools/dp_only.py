import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import acrobotics_b200 as ab
from acrobotics_b200 import _native
from acrobotics_b200.synthetic import random_configs_device, scene16
lib=_native.load(); robot, scene = ab.Kuka(), scene16()
hk, hf, hs = robot._handle(kinematics_only=True), robot._handle(), scene.native_handle()
P,S=int(sys.argv[1]) if len(sys.argv)>1 else 1000,400; m=P*S
q=random_configs_device(m,6,seed=7); st=torch.cuda.current_stream().cuda_stream
T=torch.empty((m,4,4),dtype=torch.float64,device='cuda'); _native.check(lib.acro_fk_f64(hk,q.data_ptr(),m,0,T.data_ptr(),0,st))
q_all=torch.empty((m,8,6),dtype=torch.float64,device='cuda'); valid=torch.empty(m,dtype=torch.uint8,device='cuda'); free=torch.empty(m,dtype=torch.uint8,device='cuda')
q_free=torch.empty((8*m,6),dtype=torch.float64,device='cuda'); offsets=torch.zeros(P+1,dtype=torch.int64,device='cuda')
_native.check(lib.acro_joint_solutions_csr_f64(hf,hs,T.data_ptr(),P,S,q_all.data_ptr(),valid.data_ptr(),free.data_ptr(),q_free.data_ptr(),8*m,offsets.data_ptr(),st))
off=offsets.cpu().numpy().astype(np.int64); M=int(off[-1])
value=torch.empty(M,dtype=torch.float64,device='cuda'); pred=torch.empty(M,dtype=torch.int32,device='cuda'); offd=torch.empty(P+1,dtype=torch.int64,device='cuda'); rows=torch.empty(P,dtype=torch.int64,device='cuda'); length=torch.empty(1,dtype=torch.float64,device='cuda')
for _ in range(2):
    _native.check(lib.acro_shortest_path_f64(q_free.data_ptr(),6,off.ctypes.data,P,2,None,None,value.data_ptr(),pred.data_ptr(),offd.data_ptr(),rows.data_ptr(),length.data_ptr(),st))
    torch.cuda.synchronize()
print(float(length.item()))
