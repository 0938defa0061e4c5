"""Which kernels (tile / cluster shape in the name) cuBLAS picks for the four DiT GEMM shapes, via torch.profiler."""
import torch
from torch.profiler import profile, ProfilerActivity
M = 24272
for name, N, K in (("qkv", 3072, 1024), ("out", 1024, 1024), ("ff1", 2048, 1024), ("ff2", 1024, 2048)):
    a = torch.randn(M, K, device="cuda").bfloat16(); b = torch.randn(N, K, device="cuda").bfloat16()
    for _ in range(3): c = a @ b.t()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5): c = a @ b.t()
        torch.cuda.synchronize()
    for ev in prof.key_averages():
        if ev.device_time_total > 0:
            print(f"{name} N={N} K={K}: {ev.key}  x{ev.count}  {ev.device_time_total/ev.count:.1f} us")
