import torch, time
for mb in (0.5, 4, 22, 64, 256):
    n=int(mb*1e6)
    h=torch.empty(n,dtype=torch.uint8).pin_memory(); d=torch.empty(n,dtype=torch.uint8,device='cuda')
    for fresh in (False, True):
        ts=[]
        for r in range(6):
            if fresh: h.add_(1)
            torch.cuda.synchronize()
            e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
            e0.record(); d.copy_(h,non_blocking=True); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(mb,'MB fresh' if fresh else 'MB cold ', 'ms',round(min(ts),3),'GB/s',round(n/min(ts)/1e6,1))
