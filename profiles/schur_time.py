import sys, os
sys.path.insert(0, os.getcwd())
import ctypes as C, numpy as np, torch
import cameracalibrations_b200 as cc
from cameracalibrations_b200 import _lib
lib=_lib.lib
rng=np.random.default_rng(7)
nv,nc=10000,280
views=np.concatenate([rng.normal(0,0.3,(nv,3)), np.array([-10.0,-7.0,40.0])+rng.normal(0,2.0,(nv,3))],1)
obj=np.array([[a,b,0.0] for b in range(14) for a in range(20)],dtype=np.float64)
intr=(2800.0,2800.0,1080.0,1920.0,-0.12,1.0)
tv=torch.from_numpy(views).cuda(); to=torch.from_numpy(obj).cuda()
c=cc.Calibration(intr[:4],[(views[0,:3],views[0,3:])],1.0,intr[4],["extrinsic.png"])
img=torch.empty((nv,nc,2),dtype=torch.float64,device="cuda")
for i in range(0,nv,1000):
    for j in range(i,i+1000):
        pass
# synth image points: project with first view only (values irrelevant for timing)
img.normal_(1000.0,300.0)
pv,sh=cc.reproj_jtj(intr,1.0,tv,to,img)
yz=torch.empty((nv,30),dtype=torch.float64,device="cuda"); schur=torch.empty(21,dtype=torch.float64,device="cuda")
h=_lib.context(0).handle
st=C.c_void_p(torch.cuda.current_stream().cuda_stream)
def f(): _lib.check(lib.cc_lm_schur_f64(h,C.c_void_p(pv.data_ptr()),nv,1e-3,C.c_void_p(yz.data_ptr()),C.c_void_p(schur.data_ptr()),st))
for _ in range(5): f()
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): f()
e1.record(); torch.cuda.synchronize()
print("lm_schur (schur + reduce) us:", e0.elapsed_time(e1)/50*1e3, "checksum", float(schur.sum()))
