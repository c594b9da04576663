"""Scratch timing harness (not the contract bench): times device-resident configs with CUDA events."""
import sys, os, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genfft_b200 as g

def time_it(fn, iters=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts=[]
    for _ in range(iters):
        a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[0], ts[len(ts)//2]

def c2c(n, batch, dt=np.float32):
    cd = torch.complex64 if dt==np.float32 else torch.complex128
    x=torch.randn(batch, n, dtype=cd, device='cuda'); y=torch.empty_like(x)
    p=g.FFT(n, dt, batch=batch)
    best,med=time_it(lambda: p.transform(y,x))
    bytes_=2*x.numel()*x.element_size()
    print(f"c2c {dt.__name__} n={n} batch={batch}: best {best*1e3:.1f} us med {med*1e3:.1f} us  {bytes_/best/1e6:.0f} GB/s  {5*n*np.log2(n)*batch/best/1e9:.1f} TFLOP/s   {p.describe()}", flush=True)

def r2c(n,batch):
    x=torch.randn(batch,n,device='cuda'); y=torch.empty(batch,n//2+1,dtype=torch.complex64,device='cuda')
    p=g.RealFFT(n,np.float32,half=True,batch=batch)
    best,med=time_it(lambda: p.forward(y,x), iters=10, warm=3)
    bytes_=x.numel()*4+y.numel()*8
    print(f"r2c n={n} batch={batch}: best {best:.3f} ms med {med:.3f}  {bytes_/best/1e6:.0f} GB/s {p.describe()}", flush=True)

def c2r(n,batch):
    x=torch.randn(batch,n//2+1,dtype=torch.complex64,device='cuda'); y=torch.empty(batch,n,device='cuda')
    p=g.InverseRealFFT(n,np.float32,batch=batch)
    best,med=time_it(lambda: p.inverse(y,x), iters=10, warm=3)
    bytes_=x.numel()*8+y.numel()*4
    print(f"c2r n={n} batch={batch}: best {best:.3f} ms med {med:.3f}  {bytes_/best/1e6:.0f} GB/s {p.describe()}", flush=True)

def fft2d(w,h):
    x=torch.randn(h,w,dtype=torch.complex64,device='cuda'); y=torch.empty_like(x)
    p=g.FFT2D(w,h,np.float32)
    best,med=time_it(lambda: p.transform(y,x), iters=10, warm=3)
    bytes_=2*x.numel()*8
    print(f"2d {w}x{h}: best {best:.3f} ms med {med:.3f} {bytes_/best/1e6:.0f} GB/s {p.describe()}", flush=True)

if __name__=="__main__":
    which = sys.argv[1:] or ["c2","small","c3","c4","2d"]
    if "c2" in which:
        c2c(4096, 1<<16)
        c2c(4096, 1<<14)
    if "small" in which:
        for n in (256,512,1024,2048,8192,16384):
            c2c(n, (1<<28)//n)
        c2c(1024,1)
        c2c(4096, 1<<15, np.float64)
    if "c3" in which:
        c2c(1<<24, 1, np.float64)
        c2c(1<<24, 1, np.float32)
        c2c(1<<21, 256, np.float32)
        c2c(1<<20, 16, np.float64)
    if "c4" in which:
        r2c(1<<22, 256)
    if "2d" in which:
        fft2d(4096,4096); fft2d(8192,8192); fft2d(32768,32768)
    if "c2r" in which:
        c2r(4096, 1<<16); c2r(1<<22, 64)
    if "r2csmall" in which:
        for n,b in ((4096, 1<<16), (32768, 1<<13)):
            r2c(n,b)
    if "2dmid" in which:
        fft2d(4096, 512); fft2d(4096, 1024); fft2d(4096, 2048); fft2d(1024, 1024); fft2d(2048, 2048)
