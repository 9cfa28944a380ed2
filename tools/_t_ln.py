import sys, math, ctypes as Ct
sys.path.insert(0, ".")
import torch
from diff_foley_b200 import _lib as L
DEV="cuda"; lib=L.lib()
M,C,N=32,1280,3840
sp_p=int(sys.argv[1]); sp_c=int(sys.argv[2]); bn=int(sys.argv[3]); deep=int(sys.argv[4])
a=torch.randn(M,C,device=DEV).half(); wp=(torch.randn(C,C,device=DEV)/math.sqrt(C)).half()
x32=torch.zeros(M,C,device=DEV); x16=torch.zeros(M,C,device=DEV,dtype=torch.float16)
stats=torch.zeros(M,64,2,device=DEV); tiles=Ct.c_int(0)
L.check(lib.dfb_gemm_stats(L.ptr(a),L.ptr(wp),M,C,C,None,None,L.ptr(x32),L.ptr(x16),sp_p,L.ptr(stats),Ct.byref(tiles),L.cur_stream()),"p")
torch.cuda.synchronize(); print("producer ok tiles",tiles.value, flush=True)
w=(torch.randn(N,C,device=DEV)/math.sqrt(C)).half(); s_n=w.float().sum(1); t_n=torch.zeros(N,device=DEV)
out=torch.zeros(M,N,device=DEV,dtype=torch.float16)
st=stats.view(-1)[:M*tiles.value*2].clone()
if bn: lib.dfb_debug_igemm_force(bn, deep)
L.check(lib.dfb_gemm_ln(L.ptr(x16),L.ptr(w),M,N,C,L.ptr(t_n),L.ptr(s_n),L.ptr(st),tiles.value,1e-5,0,None,L.ptr(out),sp_c,L.cur_stream()),"c")
torch.cuda.synchronize(); print("consumer ok", flush=True)
