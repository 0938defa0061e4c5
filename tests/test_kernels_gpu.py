"""Kernel-level parity on the B200, through the C ABI (device pointers): tcgen05 GEMM + fused epilogues, the
grouped conv as implicit GEMM, flash attention, LayerNorm-modulate.  Reference = plain PyTorch fp32 of the same op
on the same bf16-rounded operands; tolerance = bf16 output rounding (rel 2^-8) + fp32 accumulation noise."""
import ctypes as C
import math

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from vietvoice_tts_b200 import _lib
from vietvoice_tts_b200.arch import TINY


@pytest.fixture(scope="module")
def eng():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    lib = _lib.load()
    h = C.c_void_p()
    carch = TINY.to_c()
    torch.cuda.init()
    # The engine launches on the stream it is given; a NULL handle (torch's default stream) would make it create its
    # own non-blocking stream, unordered with the torch ops that fill the operands.  Run torch and the engine on ONE
    # explicit stream for the whole module.
    stream = torch.cuda.Stream()
    prev = torch.cuda.current_stream()
    torch.cuda.set_stream(stream)
    _lib.check(lib.vv_engine_create(C.byref(carch), 0, C.c_void_p(stream.cuda_stream), C.byref(h)))
    yield lib, h
    torch.cuda.synchronize()
    lib.vv_engine_destroy(h)
    torch.cuda.set_stream(prev)


def P(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def rel_err(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))


def run_gemm(lib, h, A, B, M, N, K, bn, **kw):
    ep = _lib.VVGemmEpilogue()
    for k, v in kw.items():
        if isinstance(v, torch.Tensor):
            setattr(ep, k, v.data_ptr())
        else:
            setattr(ep, k, v)
    _lib.check(lib.vv_gemm_bf16(h, P(A), A.stride(0), P(B), B.stride(0), M, N, K, C.byref(ep), bn))
    torch.cuda.synchronize()


@pytest.mark.parametrize("M,N,K,bn", [
    (128, 128, 64, 128), (256, 256, 128, 256), (300, 192, 256, 64), (1000, 1024, 1024, 256),
    (3002, 3072, 1024, 256), (3002, 1024, 2048, 128), (777, 100, 1024, 128), (129, 1026, 512, 256),
])
def test_gemm_plain(eng, M, N, K, bn):
    lib, h = eng
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    B = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    ldo = (N + 3) // 4 * 4
    out = torch.full((M, ldo), float("nan"), device="cuda")
    run_gemm(lib, h, A, B, M, N, K, bn, bias=bias, out_f32=out, ld_f32=ldo)
    ref = A.float() @ B.float().t() + bias
    assert torch.isfinite(out[:, :N]).all()
    assert rel_err(out[:, :N], ref) < 2e-5
    if ldo > N:
        assert torch.isnan(out[:, N:]).all()          # nothing written past N


def test_gemm_persistent_many_tiles(eng):
    """more tiles than SMs: the ring of smem stages and both TMEM accumulators wrap several times"""
    lib, h = eng
    M, N, K = 128 * 40, 256 * 9, 320
    g = torch.Generator(device="cuda").manual_seed(3)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    B = torch.randn(N, K, device="cuda", generator=g).bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    run_gemm(lib, h, A, B, M, N, K, 256, out_bf16=out, ld_bf16=N)
    ref = A.float() @ B.float().t()
    assert rel_err(out, ref) < 4e-3


# ---------------------------------------------------------------------------------------------- CTA-pair kernel
@pytest.mark.parametrize("M,N,K", [
    (256, 256, 64), (128, 256, 128), (300, 512, 256), (1000, 1024, 1024), (3034, 3072, 1024), (3034, 1024, 2048),
    (257, 256, 64 * 7), (6656, 1024, 128), (24272, 2048, 64),
])
def test_gemm_pair_plain(eng, M, N, K):
    """bn=512: 256x256 tile on a 2-CTA cluster (tcgen05 cta_group::2); ragged M, rings wrapping, odd pair count.
    On 74 clusters the last partial wave of tiles is cut into column slices: 4 x 64 columns for (1000, 1024) [16 tiles],
    (300, 512), (257, 256); 2 x 128 columns for (6656, 1024) [104 tiles] and (24272, 2048) [760 tiles]; none for the
    3072-column cases [144 tiles]"""
    lib, h = eng
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N + K)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    B = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    out = torch.full((M + 3, N), float("nan"), device="cuda")
    run_gemm(lib, h, A, B, M, N, K, 512, bias=bias, out_f32=out, ld_f32=N)
    ref = A.float() @ B.float().t() + bias
    assert torch.isfinite(out[:M]).all()
    assert rel_err(out[:M], ref) < 2e-5
    assert torch.isnan(out[M:]).all()                 # nothing written past M


def test_gemm_pair_many_tiles_bf16_gelu(eng):
    """more pair tiles than clusters: smem ring and both TMEM accumulators wrap; bf16 + GELU-tanh epilogue (ffn-up)"""
    lib, h = eng
    M, N, K = 128 * 41, 256 * 8, 320
    g = torch.Generator(device="cuda").manual_seed(3)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    B = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    run_gemm(lib, h, A, B, M, N, K, 512, bias=bias, act=1, out_bf16=out, ld_bf16=N)
    ref = torch.nn.functional.gelu(A.float() @ B.float().t() + bias, approximate="tanh")
    assert rel_err(out, ref) < 5e-3


@pytest.mark.parametrize("M,N,K,masked", [(700, 256, 512, True), (5000, 1024, 1024, False), (24272, 1024, 256, False),
                                          (6656, 1024, 192, False)])
def test_gemm_pair_gate_residual_inplace(eng, M, N, K, masked):
    """x += gate * (A B^T + bias) in place through the TMA-staged residual epilogue (out-proj / ffn-down form)"""
    lib, h = eng
    g = torch.Generator(device="cuda").manual_seed(11 + M)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    B = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    gate = torch.randn(N, device="cuda", generator=g)
    x = torch.randn(M + 2, N, device="cuda", generator=g)
    x0 = x.clone()
    kw = {}
    mask = None
    if masked:
        mask = (torch.rand(M, device="cuda", generator=g) > 0.2).to(torch.uint8)
        kw["row_mask"] = mask
    run_gemm(lib, h, A, B, M, N, K, 512, bias=bias, gate=gate, resid=x, ld_resid=N, out_f32=x, ld_f32=N, **kw)
    ref = x0[:M] + gate * (A.float() @ B.float().t() + bias)
    if masked:
        ref = ref * mask[:, None].float()
    assert rel_err(x[:M], ref) < 2e-5
    assert torch.equal(x[M:], x0[M:])                 # rows past M untouched


def test_gemm_pair_residual_separate_out_and_bf16(eng):
    """residual + fp32 + bf16 outputs (input-embedding form): the direct (non-TMA) pair epilogue"""
    lib, h = eng
    M, N, K = 1500, 512, 128
    g = torch.Generator(device="cuda").manual_seed(21)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    B = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    r = torch.randn(M, N, device="cuda", generator=g)
    o = torch.empty(M, N, device="cuda")
    ob = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    run_gemm(lib, h, A, B, M, N, K, 512, resid=r, ld_resid=N, out_f32=o, ld_f32=N, out_bf16=ob, ld_bf16=N)
    ref = A.float() @ B.float().t() + r
    assert rel_err(o, ref) < 2e-5
    assert rel_err(ob, ref) < 4e-3


def test_gemm_pair_rope_epilogue(eng):
    lib, h = eng
    dim, M = 256, 1400
    N, K = 3 * dim, 256
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    B = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    pos = torch.randint(0, 2000, (M,), device="cuda", generator=g, dtype=torch.int32)
    out = torch.empty(M, N, device="cuda")
    run_gemm(lib, h, A, B, M, N, K, 512, bias=bias, out_f32=out, ld_f32=N, row_pos=pos, rope_dim=64, rope_off2=dim)
    pre = A.float() @ B.float().t() + bias
    inv = 1.0 / (10000.0 ** (torch.arange(0, 64, 2, device="cuda", dtype=torch.float64) / 64))
    ang = pos.double()[:, None] * inv[None]
    cos = torch.repeat_interleave(torch.cos(ang), 2, -1).float()
    sin = torch.repeat_interleave(torch.sin(ang), 2, -1).float()

    def rot(x):
        x1, x2 = x[..., 0::2], x[..., 1::2]
        return torch.stack((-x2, x1), -1).reshape(x.shape)

    ref = pre.clone()
    for o in (0, dim):
        seg = pre[:, o:o + 64]
        ref[:, o:o + 64] = seg * cos + rot(seg) * sin
    assert rel_err(out, ref) < 2e-5


def test_gemm_pair_rejects_bad_n(eng):
    lib, h = eng
    A = torch.zeros(256, 64, device="cuda", dtype=torch.bfloat16)
    B = torch.zeros(192, 64, device="cuda", dtype=torch.bfloat16)
    out = torch.zeros(256, 192, device="cuda")
    ep = _lib.VVGemmEpilogue()
    ep.out_f32 = out.data_ptr(); ep.ld_f32 = 192
    assert lib.vv_gemm_bf16(h, P(A), 64, P(B), 64, 256, 192, 64, C.byref(ep), 512) < 0


@pytest.mark.parametrize("act", [1, 2, 3])
def test_gemm_activation_bf16(eng, act):
    lib, h = eng
    M, N, K = 515, 512, 256
    g = torch.Generator(device="cuda").manual_seed(act)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    B = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    run_gemm(lib, h, A, B, M, N, K, 128, bias=bias, act=act, out_bf16=out, ld_bf16=N)
    pre = A.float() @ B.float().t() + bias
    F = torch.nn.functional
    ref = {1: F.gelu(pre, approximate="tanh"), 2: F.gelu(pre), 3: F.mish(pre)}[act]
    assert (out.float() - ref).abs().max() < 2e-2
    assert rel_err(out, ref) < 5e-3


def test_gemm_gate_residual_inplace_and_mask(eng):
    lib, h = eng
    M, N, K = 700, 256, 512
    g = torch.Generator(device="cuda").manual_seed(11)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    B = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    gate = torch.randn(N, device="cuda", generator=g)
    x = torch.randn(M, N, device="cuda", generator=g)
    x0 = x.clone()
    mask = (torch.rand(M, device="cuda", generator=g) > 0.2).to(torch.uint8)
    outb = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    run_gemm(lib, h, A, B, M, N, K, 128, bias=bias, gate=gate, resid=x, ld_resid=N, out_f32=x, ld_f32=N,
             out_bf16=outb, ld_bf16=N, row_mask=mask)
    ref = (x0 + gate * (A.float() @ B.float().t() + bias)) * mask[:, None].float()
    assert rel_err(x, ref) < 2e-5
    assert rel_err(outb, ref) < 4e-3


def test_gemm_rope_epilogue(eng):
    lib, h = eng
    dim, M = 256, 400
    N, K = 3 * dim, 256
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    B = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    pos = torch.randint(0, 2000, (M,), device="cuda", generator=g, dtype=torch.int32)
    out = torch.empty(M, N, device="cuda")
    rope_heads = 2
    run_gemm(lib, h, A, B, M, N, K, 256, bias=bias, out_f32=out, ld_f32=N, row_pos=pos, rope_dim=rope_heads * 64,
             rope_off2=dim)
    pre = A.float() @ B.float().t() + bias
    inv = 1.0 / (10000.0 ** (torch.arange(0, 64, 2, device="cuda", dtype=torch.float64) / 64))
    ang = pos.double()[:, None] * inv[None]
    cos = torch.repeat_interleave(torch.cos(ang), 2, -1).float().repeat(1, rope_heads)
    sin = torch.repeat_interleave(torch.sin(ang), 2, -1).float().repeat(1, rope_heads)

    def rot(x):
        x1, x2 = x[..., 0::2], x[..., 1::2]
        return torch.stack((-x2, x1), -1).reshape(x.shape)

    ref = pre.clone()
    rd = rope_heads * 64
    for o in (0, dim):
        seg = pre[:, o:o + rd]
        ref[:, o:o + rd] = seg * cos + rot(seg) * sin
    assert rel_err(out, ref) < 2e-5


@pytest.mark.parametrize("groups,taps,T,extras", [(4, 31, 333, False), (16, 31, 2100, True), (2, 7, 700, False),
                                                 (3, 33, 257, True), (1, 1, 40, False)])
def test_conv_rows_grouped(eng, groups, taps, T, extras):
    """conv_pos_embed as implicit GEMM: grouped Conv1d over time with zero padding + Mish.  The resident-halo kernel
    (conv.cu) reads tap t through a descriptor advanced by t rows: every tap count / tile count / ragged end here
    exercises another phase of the 8-row swizzle pattern; `extras` adds the residual, row mask and bf16 output of the
    second conv of the DiT"""
    lib, h = eng
    dim = groups * 64
    g = torch.Generator(device="cuda").manual_seed(9 + T)
    x = torch.randn(T, dim, device="cuda", generator=g).bfloat16()
    w = (torch.randn(dim, 64, taps, device="cuda", generator=g) / math.sqrt(64 * taps)).bfloat16()
    bias = torch.randn(dim, device="cuda", generator=g) * 0.1
    # [g][tap][co][ci]
    wt = w.view(groups, 64, 64, taps).permute(0, 3, 1, 2).contiguous().view(groups * taps * 64, 64)
    out = torch.full((T + 1, dim), float("nan"), device="cuda")
    ep = _lib.VVGemmEpilogue()
    ep.bias = bias.data_ptr(); ep.act = 3; ep.out_f32 = out.data_ptr(); ep.ld_f32 = dim
    ref = torch.nn.functional.conv1d(x.float().t()[None], w.float(), bias, padding=taps // 2, groups=groups)[0].t()
    ref = torch.nn.functional.mish(ref)
    if extras:
        res = torch.randn(T, dim, device="cuda", generator=g)
        mask = (torch.rand(T, device="cuda", generator=g) > 0.1).to(torch.uint8)
        outb = torch.empty(T, dim, device="cuda", dtype=torch.bfloat16)
        ep.resid = res.data_ptr(); ep.ld_resid = dim; ep.row_mask = mask.data_ptr()
        ep.out_bf16 = outb.data_ptr(); ep.ld_bf16 = dim
        ref = (ref + res) * mask[:, None].float()
    _lib.check(lib.vv_conv_rows_bf16(h, P(x), dim, P(wt), T, groups, taps, C.byref(ep)))
    torch.cuda.synchronize()
    assert rel_err(out[:T], ref) < 1e-4
    assert torch.isnan(out[T:]).all()                 # nothing written past the last row
    if extras:
        assert rel_err(outb, ref) < 4e-3


@pytest.mark.parametrize("lens,heads", [([128], 1), ([256], 2), ([100], 4), ([300, 77, 513], 4), ([1501], 16),
                                         ([1877, 845], 2)])
def test_attention(eng, lens, heads):
    lib, h = eng
    dim = heads * 64
    gap = 16
    offs, r = [], 0
    for L in lens:
        offs.append(r)
        r += L + gap
    rows = r
    g = torch.Generator(device="cuda").manual_seed(sum(lens) + heads)
    qkv = torch.randn(rows, 3 * dim, device="cuda", generator=g).bfloat16()
    out = torch.zeros(rows, dim, device="cuda", dtype=torch.bfloat16)
    so = (C.c_int32 * len(lens))(*offs)
    sl = (C.c_int32 * len(lens))(*lens)
    _lib.check(lib.vv_attention_bf16(h, P(qkv), P(out), rows, so, sl, len(lens), heads))
    torch.cuda.synchronize()
    for o, L in zip(offs, lens):
        q, k, v = qkv[o:o + L].float().split(dim, dim=-1)
        qh = q.view(L, heads, 64).transpose(0, 1)
        kh = k.view(L, heads, 64).transpose(0, 1)
        vh = v.view(L, heads, 64).transpose(0, 1)
        ref = torch.nn.functional.scaled_dot_product_attention(qh, kh, vh).transpose(0, 1).reshape(L, dim)
        got = out[o:o + L].float()
        assert torch.isfinite(got).all()
        assert rel_err(got, ref) < 1e-2, (L, rel_err(got, ref))
        assert (got - ref).abs().max() < 3e-2
    # gap rows untouched
    for o, L in zip(offs, lens):
        assert (out[o + L:o + L + gap] == 0).all()


def test_attention_large_logits(eng):
    """peaky softmax: exercises the lazy-rescale path (row max grows by > 2^8 between kv tiles)"""
    lib, h = eng
    L, heads = 640, 2
    dim = heads * 64
    g = torch.Generator(device="cuda").manual_seed(1)
    qkv = torch.randn(L, 3 * dim, device="cuda", generator=g)
    qkv[:, :2 * dim] *= 4.0
    ramp = torch.linspace(0.2, 3.0, L, device="cuda")[:, None]
    qkv[:, dim:2 * dim] *= ramp                       # later keys have larger norm -> max keeps growing
    qkv = qkv.bfloat16()
    out = torch.zeros(L, dim, device="cuda", dtype=torch.bfloat16)
    so = (C.c_int32 * 1)(0)
    sl = (C.c_int32 * 1)(L)
    _lib.check(lib.vv_attention_bf16(h, P(qkv), P(out), L, so, sl, 1, heads))
    torch.cuda.synchronize()
    q, k, v = qkv.float().split(dim, dim=-1)
    ref = torch.nn.functional.scaled_dot_product_attention(
        q.view(L, heads, 64).transpose(0, 1), k.view(L, heads, 64).transpose(0, 1),
        v.view(L, heads, 64).transpose(0, 1)).transpose(0, 1).reshape(L, dim)
    assert torch.isfinite(out.float()).all()
    assert rel_err(out, ref) < 1.5e-2


def test_attention_reference_outgrown_by_overflow(eng):
    """the kernel keeps the first kv tile's row maximum as exp2 reference and only re-references when a tile's row
    sum exceeds 2^40: here keys of later tiles score hundreds of bits above the first tile (exp2 overflows to inf),
    then fall back again — the re-reference path, the O / l rescale and the masked tail all have to be exact"""
    lib, h = eng
    L, heads = 128 * 4 + 57, 2
    dim = heads * 64
    g = torch.Generator(device="cuda").manual_seed(2)
    qkv = torch.randn(L, 3 * dim, device="cuda", generator=g)
    qkv[:, dim:2 * dim] *= 0.05                       # tile 0: tiny scores
    qkv[128:256, dim:2 * dim] *= 400.0                # tile 1: scores ~ +-200 -> far beyond 2^40, partly > 2^127
    qkv[256:384, dim:2 * dim] *= 40.0                 # tile 2: moderate
    qkv = qkv.bfloat16()
    out = torch.zeros(L, dim, device="cuda", dtype=torch.bfloat16)
    so = (C.c_int32 * 1)(0)
    sl = (C.c_int32 * 1)(L)
    _lib.check(lib.vv_attention_bf16(h, P(qkv), P(out), L, so, sl, 1, heads))
    torch.cuda.synchronize()
    q, k, v = qkv.float().split(dim, dim=-1)
    ref = torch.nn.functional.scaled_dot_product_attention(
        q.view(L, heads, 64).transpose(0, 1), k.view(L, heads, 64).transpose(0, 1),
        v.view(L, heads, 64).transpose(0, 1)).transpose(0, 1).reshape(L, dim)
    assert torch.isfinite(out.float()).all()
    assert rel_err(out, ref) < 1.5e-2


@pytest.mark.parametrize("dim", [128, 256, 512, 1024])
def test_ln_modulate(eng, dim):
    lib, h = eng
    rows = 1003
    g = torch.Generator(device="cuda").manual_seed(dim)
    x = torch.randn(rows, dim, device="cuda", generator=g) * 3 + 1
    shift = torch.randn(dim, device="cuda", generator=g)
    scale = torch.randn(dim, device="cuda", generator=g) * 0.3
    out = torch.empty(rows, dim, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.vv_ln_modulate(h, P(x), rows, dim, P(shift), P(scale), 1e-6, P(out)))
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(x, (dim,), eps=1e-6) * (1 + scale) + shift
    assert (out.float() - ref).abs().max() < 4e-2
    assert rel_err(out, ref) < 4e-3
