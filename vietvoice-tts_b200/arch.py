"""ArchConfig: every architectural constant of the three graphs, in ONE struct.

The reference ships the graphs as opaque ONNX bytes fetched at run time
(/root/reference/vietvoicetts/core/model.py:73-102), so none of these constants is
visible in the reference; they follow SURVEY.md Appendix A (F5-TTS-Base DiT + Vocos-mel-24k
as named by BASELINE.json north_star).  Oracle (oracle/) and the CUDA engine (csrc/) read
the SAME struct: the C mirror is `vv_arch` in include/vvb200.h, field for field.

What IS pinned by the reference: sample_rate 24 kHz / hop 256 / NFE 32 / seed 9527
(core/model_config.py:29-34) and the 3->8 / 8->2 / 2->1 positional I/O of the sessions
(core/tts_engine.py:133-187).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, asdict, fields


@dataclass(frozen=True)
class ArchConfig:
    # --- DiT (transformer graph) ---
    dim: int = 1024
    depth: int = 22
    heads: int = 16
    head_dim: int = 64
    ff_dim: int = 2048
    n_mel: int = 100
    text_dim: int = 512
    conv_pos_k: int = 31
    conv_pos_groups: int = 16
    time_freq_dim: int = 256
    rope_heads: int = 1          # upstream F5 rotates only the first head_dim channels (head 0)
    # --- text embedding (preprocess graph) ---
    vocab: int = 2545            # number of vocab.txt lines; embedding has vocab+1 rows
    text_layers: int = 4
    text_ff: int = 1024
    pos_table_len: int = 4096
    # --- mel front-end (preprocess graph) ---
    n_fft: int = 1024
    hop: int = 256
    sample_rate: int = 24000
    # --- Vocos (decode graph) ---
    voc_dim: int = 512
    voc_ff: int = 1536
    voc_layers: int = 8
    voc_k: int = 7
    # --- sampler ---
    nfe: int = 32                # -> nfe-1 transformer calls (core/tts_engine.py:157)
    # --- floats ---
    rope_theta: float = 10000.0
    cfg_strength: float = 2.0
    sway: float = -1.0
    ln_eps: float = 1e-6
    target_rms: float = 0.1
    mel_clamp: float = 1e-5
    mel_fmin: float = 0.0
    mel_fmax: float = 12000.0
    mag_clip: float = 100.0
    pcm_scale: float = 32767.0

    def __post_init__(self):
        # floats live as float32 in `vv_arch`; round here so a blob round-trip compares equal
        for n in _FLOAT_FIELDS:
            v = ctypes.c_float(getattr(self, n)).value
            object.__setattr__(self, n, float(v))

    # derived -----------------------------------------------------------------
    @property
    def n_bins(self) -> int:
        return self.n_fft // 2 + 1

    @property
    def in_dim(self) -> int:
        """Input-embedding fan-in: [noise | prompt mel | text]."""
        return 2 * self.n_mel + self.text_dim

    @property
    def cond_dim(self) -> int:
        """Width of cat_mel_text (core/tts_engine.py:229-230): [mel | text]."""
        return self.n_mel + self.text_dim

    def validate(self) -> None:
        assert self.heads * self.head_dim == self.dim
        assert self.head_dim == 64, "attention kernel is specialised for d_h = 64"
        assert self.dim % self.conv_pos_groups == 0
        assert self.dim // self.conv_pos_groups == 64, "conv_pos kernel: 64 channels per group"
        assert self.dim % 64 == 0 and self.ff_dim % 64 == 0
        assert self.text_dim % 64 == 0 and self.text_ff % 64 == 0
        assert self.voc_dim % 64 == 0 and self.voc_ff % 64 == 0
        assert self.conv_pos_k % 2 == 1 and self.voc_k % 2 == 1
        assert 1 <= self.rope_heads <= self.heads
        assert self.n_fft == 1024 and self.hop == 256, "mel/iSTFT kernels: n_fft 1024, hop 256"
        assert self.nfe >= 2

    def to_dict(self) -> dict:
        return asdict(self)

    @classmethod
    def from_dict(cls, d: dict) -> "ArchConfig":
        names = {f.name for f in fields(cls)}
        return cls(**{k: v for k, v in d.items() if k in names})

    def to_c(self) -> "VVArch":
        c = VVArch()
        for f in fields(self):
            setattr(c, f.name, getattr(self, f.name))
        return c


_INT_FIELDS = [
    "dim", "depth", "heads", "head_dim", "ff_dim", "n_mel", "text_dim", "conv_pos_k",
    "conv_pos_groups", "time_freq_dim", "rope_heads", "vocab", "text_layers", "text_ff",
    "pos_table_len", "n_fft", "hop", "sample_rate", "voc_dim", "voc_ff", "voc_layers",
    "voc_k", "nfe",
]
_FLOAT_FIELDS = [
    "rope_theta", "cfg_strength", "sway", "ln_eps", "target_rms", "mel_clamp", "mel_fmin",
    "mel_fmax", "mag_clip", "pcm_scale",
]


class VVArch(ctypes.Structure):
    """ctypes mirror of `vv_arch` (include/vvb200.h). Field order == dataclass order."""
    _fields_ = [(n, ctypes.c_int32) for n in _INT_FIELDS] + [(n, ctypes.c_float) for n in _FLOAT_FIELDS]


assert [f.name for f in fields(ArchConfig)] == _INT_FIELDS + _FLOAT_FIELDS


# The architecture named by BASELINE.json north_star.
FULL = ArchConfig()

# A small architecture of the same topology for CPU-speed parity tests.
TINY = ArchConfig(dim=256, depth=2, heads=4, ff_dim=512, text_dim=128, conv_pos_groups=4,
                  vocab=96, text_layers=2, text_ff=256, voc_dim=128, voc_ff=384, voc_layers=2,
                  nfe=8)
