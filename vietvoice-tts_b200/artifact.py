"""Weight artefact: the engine's own blob format and the model tarball layout.

The reference loads a tar (`model-bin.pt`) holding preprocess.onnx / transformer.onnx /
decode.onnx / vocab.txt / audio_metadata.json / cleaned_audios/*
(/root/reference/vietvoicetts/core/model.py:73-129, 204-211).  The HF checkpoint is not
reachable offline, so this module (a) defines a flat binary blob that stands where each
`.onnx` stood and (b) can build a tar of exactly that layout with seeded random-init weights
of the named architecture, which both the oracle and the CUDA engine load.

Blob layout (little endian):
    0    char[8]  magic "VVB200W1"
    8    u32      n_tensors
    12   u32      graph id (0 preprocess, 1 transformer, 2 decode, 3 all)
    16   vv_arch  (33 x 4 bytes, see arch.py)
    256  entries  n_tensors x 152 bytes {char name[96]; u32 dtype; u32 ndim; i64 shape[4];
                                         u64 offset; u64 nbytes}
    ...  tensor data, each 256-byte aligned, fp32
"""
from __future__ import annotations

import ctypes
import io
import json
import math
import struct
import tarfile
import wave
from typing import Dict, Iterable

import numpy as np

from .arch import ArchConfig, VVArch

MAGIC = b"VVB200W1"
HEADER_BYTES = 256
ENTRY_BYTES = 152
GRAPH_IDS = {"preprocess": 0, "transformer": 1, "decode": 2, "all": 3}
GRAPH_PREFIX = {"preprocess": ("pre.",), "transformer": ("dit.",), "decode": ("voc.",),
                "all": ("pre.", "dit.", "voc.")}

_ENTRY = np.dtype([("name", "S96"), ("dtype", "<u4"), ("ndim", "<u4"), ("shape", "<i8", (4,)),
                   ("offset", "<u8"), ("nbytes", "<u8")])
assert _ENTRY.itemsize == ENTRY_BYTES


def pack_blob(arch: ArchConfig, tensors: Dict[str, np.ndarray], graph: str = "all") -> bytes:
    names = [n for n in tensors if n.startswith(GRAPH_PREFIX[graph])]
    head = bytearray(HEADER_BYTES)
    head[0:8] = MAGIC
    struct.pack_into("<II", head, 8, len(names), GRAPH_IDS[graph])
    carch = bytes(arch.to_c())
    head[16:16 + len(carch)] = carch
    entries = np.zeros(len(names), dtype=_ENTRY)
    off = HEADER_BYTES + ENTRY_BYTES * len(names)
    off = (off + 255) // 256 * 256
    chunks = []
    for i, n in enumerate(names):
        a = np.ascontiguousarray(tensors[n], dtype=np.float32)
        assert a.ndim <= 4 and len(n) < 96
        entries[i]["name"] = n.encode()
        entries[i]["dtype"] = 0
        entries[i]["ndim"] = a.ndim
        entries[i]["shape"][: a.ndim] = a.shape
        entries[i]["offset"] = off
        entries[i]["nbytes"] = a.nbytes
        chunks.append((off, a))
        off = (off + a.nbytes + 255) // 256 * 256
    out = bytearray(off)
    out[0:HEADER_BYTES] = head
    eb = entries.tobytes()
    out[HEADER_BYTES:HEADER_BYTES + len(eb)] = eb
    for o, a in chunks:
        out[o:o + a.nbytes] = a.tobytes()
    return bytes(out)


def unpack_blob(blob: bytes) -> tuple[ArchConfig, Dict[str, np.ndarray], int]:
    if bytes(blob[0:8]) != MAGIC:
        raise ValueError("not a VVB200 weight blob (bad magic)")
    n, gid = struct.unpack_from("<II", blob, 8)
    carch = VVArch.from_buffer_copy(bytes(blob[16:16 + ctypes.sizeof(VVArch)]))
    arch = ArchConfig(**{f: getattr(carch, f) for f, _ in VVArch._fields_})
    # float fields come back as float32-rounded python floats; keep them as such on both sides
    entries = np.frombuffer(blob, dtype=_ENTRY, count=n, offset=HEADER_BYTES)
    out = {}
    for e in entries:
        shape = tuple(int(s) for s in e["shape"][: int(e["ndim"])])
        cnt = int(e["nbytes"]) // 4
        out[e["name"].decode()] = np.frombuffer(blob, dtype="<f4", count=cnt,
                                                offset=int(e["offset"])).reshape(shape)
    return arch, out, gid


# ---------------------------------------------------------------------------------------
# Fixed (non-learned) tables that ride in the blob so oracle and engine share them bit-exact
# ---------------------------------------------------------------------------------------
def mel_filterbank(arch: ArchConfig) -> np.ndarray:
    """HTK triangular filters, no area norm: [n_bins, n_mel]."""
    n_freqs, n_mels = arch.n_bins, arch.n_mel
    all_freqs = np.linspace(0.0, arch.sample_rate / 2.0, n_freqs)
    hz2mel = lambda f: 2595.0 * np.log10(1.0 + f / 700.0)
    mel2hz = lambda m: 700.0 * (10.0 ** (m / 2595.0) - 1.0)
    m_pts = np.linspace(hz2mel(arch.mel_fmin), hz2mel(arch.mel_fmax), n_mels + 2)
    f_pts = mel2hz(m_pts)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return np.maximum(0.0, np.minimum(down, up)).astype(np.float32)


# ---------------------------------------------------------------------------------------
# Seeded random-init weights of the named architecture
# ---------------------------------------------------------------------------------------
def make_random_weights(arch: ArchConfig, seed: int = 9527) -> Dict[str, np.ndarray]:
    """Random-init weights, scaled so activations stay O(1) through depth x (nfe-1) steps.

    AdaLN gates are NOT zero-initialised (a zero gate makes every block the identity and the
    parity test vacuous).  numpy's PCG64 stream is platform-independent,
    so the same seed yields the same bytes here and on the GPU box.
    """
    arch.validate()
    rng = np.random.Generator(np.random.PCG64(seed))
    W: Dict[str, np.ndarray] = {}

    def normal(shape, std):
        return (rng.standard_normal(size=shape, dtype=np.float32) * np.float32(std)).astype(np.float32)

    def linear(name, out_f, in_f, gain=1.0, bias_std=0.02):
        W[name + ".w"] = normal((out_f, in_f), gain / math.sqrt(in_f))
        W[name + ".b"] = normal((out_f,), bias_std)

    d, td = arch.dim, arch.text_dim
    # ---- preprocess graph
    W["pre.mel_fb"] = mel_filterbank(arch)
    emb = normal((arch.vocab + 1, td), 1.0)
    W["pre.text_embed"] = emb
    for i in range(arch.text_layers):
        p = f"pre.text_blocks.{i}"
        W[p + ".dw.w"] = normal((td, 7), 1.0 / math.sqrt(7))
        W[p + ".dw.b"] = normal((td,), 0.02)
        W[p + ".ln.g"] = (1.0 + normal((td,), 0.05)).astype(np.float32)
        W[p + ".ln.b"] = normal((td,), 0.02)
        linear(p + ".pw1", arch.text_ff, td)
        W[p + ".grn.g"] = normal((arch.text_ff,), 0.3)
        W[p + ".grn.b"] = normal((arch.text_ff,), 0.02)
        linear(p + ".pw2", td, arch.text_ff, gain=0.5)
    # ---- transformer graph
    linear("dit.time.l1", d, arch.time_freq_dim)
    linear("dit.time.l2", d, d)
    linear("dit.in", d, arch.in_dim, gain=0.7)
    cg = d // arch.conv_pos_groups
    for c in ("c1", "c2"):
        W[f"dit.pos.{c}.w"] = normal((d, cg, arch.conv_pos_k), 1.0 / math.sqrt(cg * arch.conv_pos_k))
        W[f"dit.pos.{c}.b"] = normal((d,), 0.02)
    for l in range(arch.depth):
        p = f"dit.blocks.{l}"
        linear(p + ".ada", 6 * d, d, gain=1.0)
        linear(p + ".qkv", 3 * d, d)
        linear(p + ".out", d, d)
        linear(p + ".ff1", arch.ff_dim, d)
        linear(p + ".ff2", d, arch.ff_dim)
    linear("dit.final.ada", 2 * d, d)
    linear("dit.out", arch.n_mel, d)
    # ---- decode graph
    vd = arch.voc_dim
    W["voc.embed.w"] = normal((vd, arch.n_mel, arch.voc_k), 0.3 / math.sqrt(arch.n_mel * arch.voc_k))
    W["voc.embed.b"] = normal((vd,), 0.02)
    W["voc.norm.g"] = (1.0 + normal((vd,), 0.05)).astype(np.float32)
    W["voc.norm.b"] = normal((vd,), 0.02)
    for i in range(arch.voc_layers):
        p = f"voc.blocks.{i}"
        W[p + ".dw.w"] = normal((vd, arch.voc_k), 1.0 / math.sqrt(arch.voc_k))
        W[p + ".dw.b"] = normal((vd,), 0.02)
        W[p + ".ln.g"] = (1.0 + normal((vd,), 0.05)).astype(np.float32)
        W[p + ".ln.b"] = normal((vd,), 0.02)
        linear(p + ".pw1", arch.voc_ff, vd)
        linear(p + ".pw2", vd, arch.voc_ff)
        W[p + ".gamma"] = normal((vd,), 0.3)
    W["voc.final.g"] = (1.0 + normal((vd,), 0.05)).astype(np.float32)
    W["voc.final.b"] = normal((vd,), 0.02)
    linear("voc.head", arch.n_fft + 2, vd, gain=0.7)
    return W


# ---------------------------------------------------------------------------------------
# Synthetic model tarball in the reference's layout
# ---------------------------------------------------------------------------------------
DEFAULT_VOCAB_CHARS = (
    " .,!?'@$%&/"
    "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789"
    "àáảãạăằắẳẵặâầấẩẫậèéẻẽẹêềếểễệđìíỉĩịòóỏõọôồốổỗộơờớởỡợùúủũụưừứửữựỳỵỷỹý"
)


def synthetic_vocab(arch: ArchConfig) -> list[str]:
    chars = list(dict.fromkeys(DEFAULT_VOCAB_CHARS))
    upper = [c.upper() for c in chars if c.upper() != c and c.upper() not in chars]
    chars = list(dict.fromkeys(chars + upper))
    i = 0
    while len(chars) < arch.vocab:        # filler symbols the cleaner never emits
        chars.append(f"<unused{i}>")
        i += 1
    return chars[: arch.vocab]


def synthetic_prompt_pcm(n_samples: int, seed: int = 9527) -> np.ndarray:
    """N(0,1) noise peak-normalised to 29 491 and truncated to int16 — the stand-in prompt of
    SURVEY 8(d) cfg 1 (same arithmetic as core/audio_processor.py:29-44)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    a = rng.standard_normal(n_samples).astype(np.float32)
    a = a - np.mean(a)
    a = a * (np.float32(29491.0) / np.max(np.abs(a)))
    return a.astype(np.int16)


def _wav_bytes(pcm: np.ndarray, sr: int) -> bytes:
    bio = io.BytesIO()
    with wave.open(bio, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(sr)
        w.writeframes(np.asarray(pcm, dtype="<i2").tobytes())
    return bio.getvalue()


def build_model_tar(path: str, arch: ArchConfig, seed: int = 9527,
                    voices: Iterable[dict] | None = None, prompt_seconds: float = 6.0,
                    weights: Dict[str, np.ndarray] | None = None) -> None:
    """Write a tar with the member names core/model.py:73-77,84,109,207 expects.  `weights`: reuse an already
    generated `make_random_weights(arch, seed)` dict instead of drawing it again."""
    W = weights if weights is not None else make_random_weights(arch, seed)
    if voices is None:
        voices = [
            {"gender": "female", "group": "audiobook", "area": "northern", "emotion": "neutral"},
            {"gender": "male", "group": "news", "area": "southern", "emotion": "serious"},
            {"gender": "female", "group": "story", "area": "central", "emotion": "happy"},
        ]
    meta = []
    members: list[tuple[str, bytes]] = []
    for graph, fname in (("preprocess", "preprocess.onnx"), ("transformer", "transformer.onnx"),
                         ("decode", "decode.onnx")):
        members.append((fname, pack_blob(arch, W, graph)))
    members.append(("vocab.txt", ("\n".join(synthetic_vocab(arch)) + "\n").encode("utf-8")))
    for i, v in enumerate(voices):
        fn = f"voice_{i:03d}.wav"
        pcm = synthetic_prompt_pcm(int(prompt_seconds * arch.sample_rate), seed + i)
        members.append(("cleaned_audios/" + fn, _wav_bytes(pcm, arch.sample_rate)))
        meta.append({"file_name": fn, "text": "xin chào, đây là giọng mẫu số %d." % i, **v})
    members.append(("audio_metadata.json", json.dumps(meta, ensure_ascii=False).encode("utf-8")))
    with tarfile.open(path, "w") as tar:
        for name, data in members:
            ti = tarfile.TarInfo(name)
            ti.size = len(data)
            tar.addfile(ti, io.BytesIO(data))
