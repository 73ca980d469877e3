"""DCT timing on one GPU: Stockham r2r kernels (r2r_engine=0) vs the chirp-z kernels (r2r_engine=1), 512^3 float64."""
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import mpi4py_fft_b200 as B
from mpi4py_fft_b200 import _lib
torch.cuda.set_device(0)
shape=(512,512,512)
a=B.fftw.aligned(shape,dtype='d'); b=B.fftw.aligned(shape,dtype='d')
a.tensor.copy_(torch.rand(shape,dtype=torch.float64,device='cuda'))
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
st=torch.cuda.current_stream()
for eng in (0,1):
    _lib.set_option('r2r_engine',eng)
    for axis, typ in ((2,2),(1,2),(0,2),(2,3),(2,4)):
        p=B.fftw.dctn(a,axes=(axis,),type=typ,output_array=b)
        for _ in range(3): p()
        torch.cuda.synchronize(); e0.record(st)
        for _ in range(10): p()
        e1.record(st); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)/10
        print("r2r_engine=%d DCT type %d 512^3 f64 axis %d: %.3f ms  %.0f GB/s  [%s]"%(eng,typ,axis,ms,2*a.nbytes/ms/1e6,p.plan().describe().strip()[:40]))
_lib.set_option('r2r_engine',0)
