"""LayerNorm-modulate alone at the bench shape (M 24272 x 1024): us per launch and GB/s; median of 20 launches with the L2 flushed before each."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vietvoice_tts_b200 import _lib
from vietvoice_tts_b200.arch import FULL
lib = _lib.load()
h = C.c_void_p(); carch = FULL.to_c()
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
_lib.check(lib.vv_engine_create(C.byref(carch), 0, C.c_void_p(s.cuda_stream), C.byref(h)))
P = lambda t: C.c_void_p(t.data_ptr())
M, K = 24272, 1024
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(M, K, device="cuda", generator=g) * 2 + 0.5
sh = torch.randn(K, device="cuda", generator=g); sc = torch.randn(K, device="cuda", generator=g) * 0.3
out = torch.empty(M, K, device="cuda", dtype=torch.bfloat16)
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
ref = torch.nn.functional.layer_norm(x, (K,), eps=1e-6) * (1 + sc) + sh
fn = lambda: _lib.check(lib.vv_ln_modulate(h, P(x), M, K, P(sh), P(sc), 1e-6, P(out)))
fn(); torch.cuda.synchronize()
err = ((out.float() - ref).norm() / ref.norm()).item()
ts = []
for _ in range(20):
    flush.zero_()                      # cold L2, as behind a GEMM that streamed 200 MB through it
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
ms = ts[len(ts) // 2]
print(f"ln_modulate reverse {os.environ.get('VVB200_LN_REVERSE', '1')}: {ms*1e3:7.1f} us  {M*K*6/ms/1e6:7.1f} GB/s  rel err {err:.2e}")
lib.vv_engine_destroy(h)
