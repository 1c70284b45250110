# quick device timing of the construction kernels (no parity): prints per-kernel ms
python - <<'PY'
import ctypes as C, sys, os
sys.path.insert(0, os.getcwd())
import torch, ossg_b200
from ossg_b200 import _lib
from ossg_b200.grids import make_params
L=_lib.lib(); dev=torch.device('cuda',0)
for typ,dt,name in ((_lib.OSSG_F64,torch.float64,'f64'),(_lib.OSSG_F32,torch.float32,'f32')):
    Nx,Ny,H=4320,2160,4
    p=make_params((Nx,Ny,100),(H,H,H),-80.0,55.0,70.0,6371.0e3)
    out=[torch.empty((Ny+2*H,Nx+2*H),dtype=dt,device=dev) for _ in range(20)]
    optr=_lib.ptr_array([t.data_ptr() for t in out])
    nb=L.ossg_tripolar_workspace_bytes(C.byref(p),typ); ws=torch.empty(max(nb,16),dtype=torch.uint8,device=dev)
    s=torch.cuda.current_stream(dev)
    prof=(C.c_float*4)(); acc=[0.0]*4; n=20
    for it in range(n+5):
        _lib.check(L.ossg_tripolar_build_profiled(C.byref(p),typ,optr,ws.data_ptr(),ws.numel(),s.cuda_stream,prof))
        if it>=5:
            for k in range(4): acc[k]+=prof[k]/n
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(20): _lib.check(L.ossg_tripolar_build(C.byref(p),typ,optr,ws.data_ptr(),ws.numel(),s.cuda_stream))
    e1.record(s); torch.cuda.synchronize()
    print(name,'tables %.4f coords %.4f tile %.4f frame %.4f | step %.4f ms'%(*acc,e0.elapsed_time(e1)/20))
PY
