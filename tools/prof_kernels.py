"""Micro-driver for ncu captures and event timing of the hot kernels at the bench shape (B=8, T=1501 -> M=24272).
usage: python tools/prof_kernels.py [attn|gemm|all] [reps]"""
import ctypes as C
import math
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vietvoice_tts_b200 import _lib
from vietvoice_tts_b200.arch import FULL

what = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
lib = _lib.load()
h = C.c_void_p()
carch = FULL.to_c()
s = torch.cuda.Stream()
torch.cuda.set_stream(s)
_lib.check(lib.vv_engine_create(C.byref(carch), 0, C.c_void_p(s.cuda_stream), C.byref(h)))
P = lambda t: C.c_void_p(t.data_ptr())
T, nseq, heads, dim = 1501, 16, 16, 1024
gap = 16
offs = [i * (T + gap) for i in range(nseq)]
M = (nseq * (T + gap) + 7) // 8 * 8
g = torch.Generator(device="cuda").manual_seed(0)


def timeit(fn, flops, name):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:28s} {ms*1e3:9.1f} us  {flops/ms/1e9:8.1f} TFLOP/s")


if what in ("attn", "all"):
    qkv = torch.randn(M, 3 * dim, device="cuda", generator=g).bfloat16()
    out = torch.zeros(M, dim, device="cuda", dtype=torch.bfloat16)
    so = (C.c_int32 * nseq)(*offs)
    sl = (C.c_int32 * nseq)(*([T] * nseq))
    fn = lambda: _lib.check(lib.vv_attention_bf16(h, P(qkv), P(out), M, so, sl, nseq, heads))
    timeit(fn, nseq * 4.0 * T * T * dim, "attention 16x1501 h16")

if what in ("gemm", "all"):
    for name, N, K, resid, act in (("qkv", 3072, 1024, False, 0), ("out", 1024, 1024, True, 0),
                                    ("ff1", 2048, 1024, False, 1), ("ff2", 1024, 2048, True, 0)):
        A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
        B = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
        bias = torch.randn(N, device="cuda", generator=g)
        gate = torch.randn(N, device="cuda", generator=g)
        ep = _lib.VVGemmEpilogue()
        ep.bias = bias.data_ptr()
        if resid:
            x = torch.randn(M, N, device="cuda", generator=g)
            ep.gate = gate.data_ptr(); ep.resid = x.data_ptr(); ep.ld_resid = N; ep.out_f32 = x.data_ptr(); ep.ld_f32 = N
        else:
            o = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
            ep.out_bf16 = o.data_ptr(); ep.ld_bf16 = N; ep.act = act
        for bn in (256, 512):
            fn = lambda: _lib.check(lib.vv_gemm_bf16(h, P(A), K, P(B), K, M, N, K, C.byref(ep), bn))
            timeit(fn, 2.0 * M * N * K, f"gemm {name} N={N} K={K} bn={bn}")
            if bn == 512 and hasattr(lib, "vv_gemm_timing_dump"):
                sys.stdout.flush()
                lib.vv_gemm_timing_dump()
torch.cuda.synchronize()
lib.vv_engine_destroy(h)
if what in ("attn", "all") and hasattr(lib, "vv_attn_trace_dump"):
    lib.vv_attn_trace_dump(b"gpurun_out/attn_trace.csv")
if what in ("attn", "all"):
    for name in ("vv_attn_timing_dump",):
        if hasattr(lib, name):
            sys.stdout.flush()
            getattr(lib, name)()
