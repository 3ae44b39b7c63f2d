"""Host->device bandwidth probe behind DESIGN.md §5: torch pinned memory vs write-combined pinned memory for one 67 MB step of
fp32 latents (B200 box: 53.6 vs 50.4 GB/s).  Usage: python devtools/h2d_probe.py"""
import ctypes, torch, time
rt = ctypes.CDLL("libcudart.so.12") if False else None
import torch.cuda
torch.cuda.init()
cudart = torch.cuda.cudart()
N = 67115008
dev = torch.empty(N, dtype=torch.uint8, device="cuda")
def bw(host_ptr_tensor, label, reps=20):
    s = torch.cuda.Stream()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        for _ in range(3):
            dev.copy_(host_ptr_tensor, non_blocking=True)
        a.record(s)
        for _ in range(reps):
            dev.copy_(host_ptr_tensor, non_blocking=True)
        b.record(s)
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print(f"{label}: {ms:.3f} ms per 67 MB = {N / ms / 1e6:.1f} GB/s")
h = torch.empty(N, dtype=torch.uint8).pin_memory()
h.fill_(3)
bw(h, "torch pinned")
# write-combined pinned
lib = ctypes.CDLL("libcudart.so")
p = ctypes.c_void_p()
rc = lib.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(N), ctypes.c_uint(0x04))  # cudaHostAllocWriteCombined
print("cudaHostAlloc WC rc", rc)
if rc == 0:
    buf = (ctypes.c_uint8 * N).from_address(p.value)
    t = torch.frombuffer(buf, dtype=torch.uint8)
    t.fill_(3)
    print("is_pinned", t.is_pinned())
    bw(t, "write-combined pinned")
    # two half copies on two streams
