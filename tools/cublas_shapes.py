import torch, time
torch.backends.cuda.matmul.allow_bf16_reduced_precision_reduction = True
M=24272
for name,N,K in (("qkv",3072,1024),("out",1024,1024),("ff1",2048,1024),("ff2",1024,2048),("sq8192",8192,8192)):
    m = 8192 if name=="sq8192" else M
    a=torch.randn(m,K,device="cuda").bfloat16(); b=torch.randn(N,K,device="cuda").bfloat16()
    for _ in range(3): c=a@b.t()
    torch.cuda.synchronize()
    reps=200 if name!="sq8192" else 40
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): c=a@b.t()
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/reps
    print(f"cuBLAS {name:7s} M={m} N={N} K={K}: {ms*1e3:8.1f} us  {2.0*m*N*K/ms/1e9:8.1f} TFLOP/s (plain GEMM, bf16 out, {reps} back to back)")
