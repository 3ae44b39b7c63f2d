import torch, time
dev = torch.device('cuda', 0)
for mb in [8, 32, 64, 256]:
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(mb << 20, dtype=torch.uint8, device=dev)
    for _ in range(2): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): d.copy_(h, non_blocking=True)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    a.record()
    for _ in range(5): h.copy_(d, non_blocking=True)
    b.record(); torch.cuda.synchronize()
    ms2 = a.elapsed_time(b) / 5
    print(f"{mb} MB  H2D {mb / 1024 / (ms / 1e3):.1f} GB/s   D2H {mb / 1024 / (ms2 / 1e3):.1f} GB/s")
