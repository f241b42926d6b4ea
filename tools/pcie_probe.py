import torch, time
n=4_000_000_000//8
d=torch.empty(n,dtype=torch.float64,device='cuda'); h=torch.empty(n,dtype=torch.float64).pin_memory()
for name,fn in (("D2H",lambda: h.copy_(d,non_blocking=True)),("H2D",lambda: d.copy_(h,non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t=time.perf_counter(); fn(); torch.cuda.synchronize(); dt=time.perf_counter()-t
    print(name, "%.1f GB/s"%(n*8/dt/1e9))
# 2D strided like the e2e path: 500 columns, chunk of 134144 rows out of 1e6
import ctypes
rows,cols,ld=134144,500,1_000_000
hd=torch.empty(cols*ld,dtype=torch.float64).pin_memory(); dd=torch.empty(rows*cols,dtype=torch.float64,device='cuda')
cudart=ctypes.CDLL("libcudart.so")
def c2d():
    cudart.cudaMemcpy2DAsync(ctypes.c_void_p(hd.data_ptr()), ctypes.c_size_t(ld*8), ctypes.c_void_p(dd.data_ptr()), ctypes.c_size_t(rows*8), ctypes.c_size_t(rows*8), ctypes.c_size_t(cols), 2, ctypes.c_void_p(0))
c2d(); torch.cuda.synchronize()
t=time.perf_counter(); c2d(); torch.cuda.synchronize(); dt=time.perf_counter()-t
print("2D D2H", "%.1f GB/s"%(rows*cols*8/dt/1e9))
