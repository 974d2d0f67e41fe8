import torch
def t(f, n=10):
    f(); torch.cuda.synchronize()
    best=1e9
    for _ in range(n):
        a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); best=min(best,a.elapsed_time(b))
    return best
N=1<<31
x=torch.empty(N,dtype=torch.uint8,device='cuda'); y=torch.empty(N,dtype=torch.uint8,device='cuda')
print("memset 2GiB: %.0f GB/s"%(N/t(lambda:x.zero_())/1e6))
xf=x.view(torch.float16); yf=y.view(torch.float16)
print("fill fp16 2GiB: %.0f GB/s"%(N/t(lambda:xf.fill_(1.0))/1e6))
print("copy 2GiB: %.0f GB/s (r+w)"%(2*N/t(lambda:y.copy_(x))/1e6))
xi=x.view(torch.int32)
print("read (sum) 2GiB: %.0f GB/s"%(N/t(lambda:xi.sum())/1e6))
