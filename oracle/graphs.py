"""ORACLE — test infrastructure only.  PARITY UNPINNED.

CPU fp32 (PyTorch) restatement of the three graphs the reference executes through ONNX
Runtime: `preprocess`, `transformer`, `decode`
(/root/reference/vietvoicetts/core/tts_engine.py:133-146, 148-174, 176-187; sessions built at
/root/reference/vietvoicetts/core/model.py:98-106).

Why "unpinned": the arithmetic lives in artefacts that are NOT under /root/reference —
  * executor: `onnxruntime-gpu>=1.20.2` (pyproject.toml:41), not installed, not in the wheelhouse;
  * graphs + weights: https://huggingface.co/nguyenvulebinh/VietVoice-TTS/resolve/main/model-bin.pt
    (core/model_config.py:26), branch `main`, no revision pin, unreachable offline;
and the reference's tests mock every session (tests/test_tts_engine_full.py:55-75), so there is no
golden vector for this path.  The body below is the published F5-TTS-Base / Vocos-mel-24k
architecture named by BASELINE.json north_star, as specified in SURVEY.md Appendix A; what it is
anchored on from the reference itself is the positional I/O contract (3->8, 8->2, 2->1), dtypes,
the `nfe_step-1` loop and the host constants (24 kHz, hop 256, NFE 32).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.  The product (vietvoice-tts_b200/) never does.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from vietvoice_tts_b200.arch import ArchConfig


def _t(a) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a)).float()


# =========================================================================================
# shared helpers
# =========================================================================================
def time_grid(arch: ArchConfig, nfe: Optional[int] = None) -> torch.Tensor:
    """t_i = linspace(0,1,nfe) with sway sampling; float64 then cast."""
    nfe = nfe or arch.nfe
    t = torch.linspace(0, 1, nfe, dtype=torch.float64)
    t = t + arch.sway * (torch.cos(math.pi / 2 * t) - 1 + t)
    return t


def rope_tables(arch: ArchConfig, T: int):
    """x-transformers interleaved RoPE over head_dim. -> cos, sin [T, head_dim]"""
    hd = arch.head_dim
    inv = 1.0 / (arch.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.float64) / hd))
    ang = torch.arange(T, dtype=torch.float64)[:, None] * inv[None, :]
    ang = torch.repeat_interleave(ang, 2, dim=-1)
    return torch.cos(ang).float(), torch.sin(ang).float()


def _rotate_half_interleaved(x):
    x1 = x[..., 0::2]
    x2 = x[..., 1::2]
    return torch.stack((-x2, x1), dim=-1).reshape(x.shape)


def _ln(x, eps, g=None, b=None):
    return F.layer_norm(x, (x.shape[-1],), g, b, eps)


# =========================================================================================
# preprocess graph  (core/tts_engine.py:133-146; outputs unpacked at :229-230)
# =========================================================================================
class OraclePreprocess:
    input_names = ["audio", "text_ids", "max_duration"]
    output_names = ["noise", "rope_cos_q", "rope_sin_q", "rope_cos_k", "rope_sin_k",
                    "cat_mel_text", "cat_mel_text_drop", "ref_signal_len"]

    def __init__(self, arch: ArchConfig, W: Dict[str, np.ndarray], seed: int = 9527):
        self.arch = arch
        self.W = {k: _t(v) for k, v in W.items() if k.startswith("pre.")}
        self.gen = torch.Generator().manual_seed(seed)      # stands for ort.set_seed (core/model.py:133)
        hd = arch.text_dim
        freqs = 1.0 / (10000.0 ** (torch.arange(0, hd, 2, dtype=torch.float64)[: hd // 2] / hd))
        ang = torch.arange(arch.pos_table_len, dtype=torch.float64)[:, None] * freqs[None, :]
        self.pos_table = torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1).float()

    # -- mel front-end
    def mel(self, audio_i16: np.ndarray) -> torch.Tensor:
        a = self.arch
        x = torch.from_numpy(np.asarray(audio_i16).reshape(-1).astype(np.float32)) / 32768.0
        rms = torch.sqrt(torch.mean(x * x))
        if rms < a.target_rms and rms > 0:
            x = x * (a.target_rms / rms)
        win = torch.hann_window(a.n_fft, periodic=True)
        spec = torch.stft(x, a.n_fft, hop_length=a.hop, win_length=a.n_fft, window=win, center=True,
                          pad_mode="reflect", return_complex=True)
        mag = spec.abs().transpose(0, 1)                      # [T_ref, n_bins]
        mel = mag @ self.W["pre.mel_fb"]                      # [T_ref, n_mel]
        return torch.log(torch.clamp(mel, min=a.mel_clamp))

    # -- text embedding
    def text_embed(self, ids_plus1: torch.Tensor, T: int) -> torch.Tensor:
        a, W = self.arch, self.W
        ids = ids_plus1[:T]
        ids = F.pad(ids, (0, T - ids.shape[0]), value=0)
        x = W["pre.text_embed"][ids]                          # [T, text_dim]
        pos = torch.clamp(torch.arange(T), max=a.pos_table_len - 1)
        x = x + self.pos_table[pos]
        for i in range(a.text_layers):
            p = f"pre.text_blocks.{i}"
            r = x
            h = F.conv1d(x.t()[None], W[p + ".dw.w"][:, None, :], W[p + ".dw.b"], padding=3,
                         groups=a.text_dim)[0].t()
            h = _ln(h, a.ln_eps, W[p + ".ln.g"], W[p + ".ln.b"])
            h = F.gelu(h @ W[p + ".pw1.w"].t() + W[p + ".pw1.b"])
            gx = torch.sqrt(torch.sum(h * h, dim=0, keepdim=True))          # L2 over TIME, per channel
            nx = gx / (gx.mean(dim=-1, keepdim=True) + 1e-6)
            h = W[p + ".grn.g"] * (h * nx) + W[p + ".grn.b"] + h
            h = h @ W[p + ".pw2.w"].t() + W[p + ".pw2.b"]
            x = r + h
        return x

    def run(self, audio: np.ndarray, text_ids: np.ndarray, max_duration: np.ndarray,
            noise: Optional[np.ndarray] = None):
        a = self.arch
        T = int(np.asarray(max_duration).reshape(-1)[0])
        mel = self.mel(audio)
        T_ref = mel.shape[0]
        assert T_ref == np.asarray(audio).size // a.hop + 1   # == host ref_audio_len (core/tts_engine.py:55)
        cond = torch.zeros(T, a.n_mel)
        cond[: min(T_ref, T)] = mel[:T]
        ids = torch.from_numpy(np.asarray(text_ids).reshape(-1).astype(np.int64)) + 1
        text_c = self.text_embed(ids, T)
        text_u = self.text_embed(torch.zeros(0, dtype=torch.int64), T)
        cat_c = torch.cat([cond, text_c], dim=-1)[None]
        cat_u = torch.cat([torch.zeros(T, a.n_mel), text_u], dim=-1)[None]
        if noise is None:
            noise_t = torch.randn(1, T, a.n_mel, generator=self.gen)
        else:
            noise_t = _t(noise).reshape(1, T, a.n_mel)
        cos, sin = rope_tables(a, T)
        outs = [noise_t, cos[None], sin[None], cos.t()[None].contiguous(), sin.t()[None].contiguous(),
                cat_c, cat_u]
        return [o.numpy() for o in outs] + [np.array([T_ref], dtype=np.int64)]


# =========================================================================================
# transformer graph  (core/tts_engine.py:148-174)
# =========================================================================================
class OracleTransformer:
    input_names = ["noise", "rope_cos_q", "rope_sin_q", "rope_cos_k", "rope_sin_k",
                   "cat_mel_text", "cat_mel_text_drop", "time_step"]
    output_names = ["noise_out", "time_step_out"]

    def __init__(self, arch: ArchConfig, W: Dict[str, np.ndarray], nfe: Optional[int] = None,
                 fuse_nfe: int = 1):
        self.arch = arch
        self.W = {k: _t(v) for k, v in W.items() if k.startswith("dit.")}
        self.nfe = nfe or arch.nfe
        self.fuse_nfe = fuse_nfe
        self.t = time_grid(arch, self.nfe)
        self.taps: Optional[dict] = None        # set to {} to record intermediates

    def time_embed(self, t: float) -> torch.Tensor:
        a, W = self.arch, self.W
        half = a.time_freq_dim // 2
        f = torch.exp(torch.arange(half, dtype=torch.float64) * (-math.log(10000.0) / (half - 1)))
        e = 1000.0 * t * f
        e = torch.cat([torch.sin(e), torch.cos(e)]).float()
        h = F.silu(e @ W["dit.time.l1.w"].t() + W["dit.time.l1.b"])
        return h @ W["dit.time.l2.w"].t() + W["dit.time.l2.b"]

    def velocity(self, x_in: torch.Tensor, t: float, cos: torch.Tensor, sin: torch.Tensor):
        """x_in [B,T,in_dim] -> v [B,T,n_mel].  SURVEY A.2 steps 1-5."""
        a, W = self.arch, self.W
        Bn, T, _ = x_in.shape
        tap = self.taps
        x = x_in @ W["dit.in.w"].t() + W["dit.in.b"]
        if tap is not None: tap["x0"] = x.clone()
        h = x.transpose(1, 2)
        pad = a.conv_pos_k // 2
        h = F.mish(F.conv1d(h, W["dit.pos.c1.w"], W["dit.pos.c1.b"], padding=pad, groups=a.conv_pos_groups))
        if tap is not None: tap["conv1"] = h.transpose(1, 2).clone()
        h = F.mish(F.conv1d(h, W["dit.pos.c2.w"], W["dit.pos.c2.b"], padding=pad, groups=a.conv_pos_groups))
        x = x + h.transpose(1, 2)
        if tap is not None: tap["x_embed"] = x.clone()
        temb = F.silu(self.time_embed(t))
        rd = a.rope_heads * a.head_dim
        cosr = cos.repeat(1, a.rope_heads)[None]
        sinr = sin.repeat(1, a.rope_heads)[None]
        for l in range(a.depth):
            p = f"dit.blocks.{l}"
            m = temb @ W[p + ".ada.w"].t() + W[p + ".ada.b"]
            sh_a, sc_a, g_a, sh_f, sc_f, g_f = m.chunk(6)
            hN = _ln(x, a.ln_eps) * (1 + sc_a) + sh_a
            if tap is not None and l == 0: tap["ln0"] = hN.clone()
            qkv = hN @ W[p + ".qkv.w"].t() + W[p + ".qkv.b"]
            q, k, v = qkv.split(a.dim, dim=-1)
            # upstream quirk: RoPE before the head split on the first rope_heads*head_dim channels
            q = torch.cat([q[..., :rd] * cosr + _rotate_half_interleaved(q[..., :rd]) * sinr, q[..., rd:]], -1)
            k = torch.cat([k[..., :rd] * cosr + _rotate_half_interleaved(k[..., :rd]) * sinr, k[..., rd:]], -1)
            if tap is not None and l == 0: tap["q0"], tap["k0"], tap["v0"] = q.clone(), k.clone(), v.clone()
            qh = q.view(Bn, T, a.heads, a.head_dim).transpose(1, 2)
            kh = k.view(Bn, T, a.heads, a.head_dim).transpose(1, 2)
            vh = v.view(Bn, T, a.heads, a.head_dim).transpose(1, 2)
            o = F.scaled_dot_product_attention(qh, kh, vh)
            o = o.transpose(1, 2).reshape(Bn, T, a.dim)
            if tap is not None and l == 0: tap["attn0"] = o.clone()
            o = o @ W[p + ".out.w"].t() + W[p + ".out.b"]
            x = x + g_a * o
            hN = _ln(x, a.ln_eps) * (1 + sc_f) + sh_f
            f = F.gelu(hN @ W[p + ".ff1.w"].t() + W[p + ".ff1.b"], approximate="tanh")
            f = f @ W[p + ".ff2.w"].t() + W[p + ".ff2.b"]
            x = x + g_f * f
            if tap is not None: tap[f"x_l{l}"] = x.clone()
        m = temb @ W["dit.final.ada.w"].t() + W["dit.final.ada.b"]
        sc, sh = m.chunk(2)
        x = _ln(x, a.ln_eps) * (1 + sc) + sh
        v = x @ W["dit.out.w"].t() + W["dit.out.b"]
        if tap is not None: tap["v"] = v.clone()
        return v

    def run(self, noise, rope_cos_q, rope_sin_q, rope_cos_k, rope_sin_k, cat_mel_text,
            cat_mel_text_drop, time_step):
        a = self.arch
        x = _t(noise)
        cos, sin = _t(rope_cos_q)[0], _t(rope_sin_q)[0]
        cat_c, cat_u = _t(cat_mel_text), _t(cat_mel_text_drop)
        i = int(np.asarray(time_step).reshape(-1)[0])
        for _ in range(self.fuse_nfe):
            t0, t1 = float(self.t[i]), float(self.t[i + 1])
            xin = torch.cat([torch.cat([x, cat_c], -1), torch.cat([x, cat_u], -1)], 0)
            v = self.velocity(xin, t0, cos, sin)
            vhat = v[0:1] + a.cfg_strength * (v[0:1] - v[1:2])
            x = x + np.float32(t1 - t0) * vhat
            i += 1
        return [x.numpy(), np.array([i], dtype=np.int32)]


# =========================================================================================
# decode graph  (core/tts_engine.py:176-187)
# =========================================================================================
class OracleDecode:
    input_names = ["denoised", "ref_signal_len"]
    output_names = ["output_audio"]

    def __init__(self, arch: ArchConfig, W: Dict[str, np.ndarray]):
        self.arch = arch
        self.W = {k: _t(v) for k, v in W.items() if k.startswith("voc.")}
        self.taps: Optional[dict] = None

    def backbone(self, mel_t: torch.Tensor) -> torch.Tensor:
        """mel_t [T_tgt, n_mel] -> head output [T_tgt, n_fft+2]."""
        a, W = self.arch, self.W
        pad = a.voc_k // 2
        x = F.conv1d(mel_t.t()[None], W["voc.embed.w"], W["voc.embed.b"], padding=pad)[0].t()
        x = _ln(x, a.ln_eps, W["voc.norm.g"], W["voc.norm.b"])
        for i in range(a.voc_layers):
            p = f"voc.blocks.{i}"
            r = x
            h = F.conv1d(x.t()[None], W[p + ".dw.w"][:, None, :], W[p + ".dw.b"], padding=pad,
                         groups=a.voc_dim)[0].t()
            h = _ln(h, a.ln_eps, W[p + ".ln.g"], W[p + ".ln.b"])
            h = F.gelu(h @ W[p + ".pw1.w"].t() + W[p + ".pw1.b"])
            h = h @ W[p + ".pw2.w"].t() + W[p + ".pw2.b"]
            x = r + W[p + ".gamma"] * h
        x = _ln(x, a.ln_eps, W["voc.final.g"], W["voc.final.b"])
        return x @ W["voc.head.w"].t() + W["voc.head.b"]

    def wave(self, head: torch.Tensor) -> torch.Tensor:
        a = self.arch
        nb = a.n_bins
        mag = torch.clamp(torch.exp(head[:, :nb]), max=a.mag_clip)
        ph = head[:, nb:]
        S = torch.complex(mag * torch.cos(ph), mag * torch.sin(ph)).t()      # [n_bins, T]
        win = torch.hann_window(a.n_fft, periodic=True)
        return torch.istft(S, a.n_fft, hop_length=a.hop, win_length=a.n_fft, window=win, center=True)

    def run(self, denoised, ref_signal_len):
        a = self.arch
        r = int(np.asarray(ref_signal_len).reshape(-1)[0])
        mel_t = _t(denoised)[0, r:, :]
        head = self.backbone(mel_t)
        if self.taps is not None: self.taps["head"] = head.clone()
        w = self.wave(head)
        if self.taps is not None: self.taps["wave"] = w.clone()
        pcm = torch.clamp(w * a.pcm_scale, -32768.0, 32767.0)
        return [pcm.numpy().astype(np.int16).reshape(1, 1, -1)]      # astype truncates toward zero


# =========================================================================================
# the reference's driver loop, restated (core/tts_engine.py:225-238)
# =========================================================================================
class OracleSessions:
    def __init__(self, arch: ArchConfig, W: Dict[str, np.ndarray], seed: int = 9527,
                 nfe: Optional[int] = None, fuse_nfe: int = 1):
        self.arch = arch
        self.nfe = nfe or arch.nfe
        self.fuse_nfe = fuse_nfe
        self.preprocess = OraclePreprocess(arch, W, seed)
        self.transformer = OracleTransformer(arch, W, self.nfe, fuse_nfe)
        self.decode = OracleDecode(arch, W)

    def synthesize_chunk(self, audio, text_ids, max_duration, noise=None, collect_steps=False):
        pre = self.preprocess.run(audio, text_ids, max_duration, noise)
        x, cq, sq, ck, sk, cat_c, cat_u, ref_len = pre
        ts = np.array([0], dtype=np.int32)
        steps = []
        for _ in range(0, self.nfe - 1, self.fuse_nfe):
            x, ts = self.transformer.run(x, cq, sq, ck, sk, cat_c, cat_u, ts)
            if collect_steps:
                steps.append(x.copy())
        wave = self.decode.run(x, ref_len)[0]
        return wave, x, steps, pre
