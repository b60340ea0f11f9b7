import torch, time
x = torch.empty(1<<30, dtype=torch.uint8).pin_memory()
d = torch.empty(1<<30, dtype=torch.uint8, device="cuda")
for chunk in (1<<30, 32*102400, 102400):
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    n = (1<<30)//chunk
    for i in range(n):
        d[i*chunk:(i+1)*chunk].copy_(x[i*chunk:(i+1)*chunk], non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print("H2D chunk", chunk, "GB/s", n*chunk/e0.elapsed_time(e1)/1e6)
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record(); x.copy_(d, non_blocking=True); e1.record(); torch.cuda.synchronize()
print("D2H GB/s", (1<<30)/e0.elapsed_time(e1)/1e6)
