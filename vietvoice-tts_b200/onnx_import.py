"""ONNX-initialiser importer: upstream `model-bin.pt` -> the engine's weight blobs.

The reference's artefact is a tar of three ONNX graphs (/root/reference/vietvoicetts/core/model.py:73-102) fetched
from Hugging Face at run time (core/model_config.py:26).  Neither the file nor the `onnx` package is available
offline, so this module reads the ONNX container itself — protobuf wire format, no third-party import — pulls the
initialisers (and Constant-node tensors) out of `preprocess.onnx` / `transformer.onnx` / `decode.onnx`, maps them to
the blob names of `artifact.py`, and writes a tar of the same layout whose three `.onnx` members are VVB200 blobs.
Everything else in the tar (vocab.txt, audio_metadata.json, cleaned_audios/*) is copied byte for byte, so
`ModelSessionManager` loads the result unchanged.

How a tensor finds its blob name, in this order:
  1. by name — the upstream PyTorch parameter names (F5-TTS `DiT`, Vocos `VocosBackbone`/`ISTFTHead`) survive export
     for every initialiser torch does not rewrite (`_NAME_RULES`);
  2. by edge — a `MatMul` weight is exported transposed under an anonymous name (`onnx::MatMul_123`); its output feeds
     the `Add` of the layer's bias, which keeps its name, so `X.bias` names the weight `X.weight` (stored [in, out],
     transposed back here);
  3. by shape, for the one fixed table (`mel_fb`, [n_bins, n_mel] in either orientation).
Separate `to_q/to_k/to_v` projections are fused into `qkv` in that order; depthwise conv kernels lose their singleton
channel axis; GRN parameters are flattened.  `ConversionReport` lists what was mapped, what was left over in the ONNX
files and which blob tensors are still missing — `strict=True` raises on the latter.  Architecture constants that the
weights determine (dim, depth, ff_dim, text_dim, vocab, layer counts, kernel sizes) are read off the shapes.

This has never seen the real checkpoint (unreachable offline): `tests/test_onnx_import_cpu.py` checks it against
ONNX containers this module's own writer produces from seeded weights under upstream-style names, including the
anonymous-MatMul case.
"""
from __future__ import annotations

import io
import re
import struct
import tarfile
from dataclasses import dataclass, field, replace
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np

from .arch import ArchConfig, FULL
from . import artifact

# --------------------------------------------------------------------------------------------------------------
# protobuf wire format (the subset ONNX uses: varint, 64-bit, length-delimited, 32-bit)
# --------------------------------------------------------------------------------------------------------------


def _varint(buf: memoryview, pos: int) -> Tuple[int, int]:
    out = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if b < 0x80:
            return out, pos
        shift += 7
        if shift > 70:
            raise ValueError("malformed varint")


def _fields(buf: memoryview) -> Iterable[Tuple[int, int, object]]:
    """Yield (field number, wire type, value); length-delimited values come back as memoryviews."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = bytes(buf[pos:pos + 8])
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            if pos + ln > n:
                raise ValueError("truncated length-delimited field")
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = bytes(buf[pos:pos + 4])
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield num, wt, v


def _packed_varints(v, wt) -> List[int]:
    if wt == 0:
        return [v]
    out, pos = [], 0
    while pos < len(v):
        x, pos = _varint(v, pos)
        out.append(x)
    return out


def _signed(x: int) -> int:
    return x - (1 << 64) if x >= (1 << 63) else x


# TensorProto.DataType -> numpy
_DTYPES = {1: "<f4", 2: "u1", 3: "i1", 4: "<u2", 5: "<i2", 6: "<i4", 7: "<i8", 9: "?", 10: "<f2", 11: "<f8",
           12: "<u4", 13: "<u8"}
_BF16 = 16


def _decode_tensor(buf: memoryview) -> Tuple[str, np.ndarray]:
    """TensorProto: dims=1, data_type=2, float_data=4, int32_data=5, int64_data=7, name=8, raw_data=9,
    double_data=10, data_location=14."""
    dims: List[int] = []
    dtype = 0
    name = ""
    raw = None
    f32: List[np.ndarray] = []
    i32: List[int] = []
    i64: List[int] = []
    f64: List[np.ndarray] = []
    for num, wt, v in _fields(buf):
        if num == 1:
            dims += [_signed(x) for x in _packed_varints(v, wt)]
        elif num == 2:
            dtype = v
        elif num == 4:
            f32.append(np.frombuffer(v, dtype="<f4"))
        elif num == 5:
            i32 += [_signed(x) for x in _packed_varints(v, wt)]
        elif num == 7:
            i64 += [_signed(x) for x in _packed_varints(v, wt)]
        elif num == 8:
            name = bytes(v).decode("utf-8")
        elif num == 9:
            raw = bytes(v)
        elif num == 10:
            f64.append(np.frombuffer(v, dtype="<f8"))
        elif num == 14 and v == 1:
            raise ValueError(f"tensor '{name}': external data is not supported (weights must be embedded)")
    shape = tuple(dims)
    if dtype == _BF16:
        bits = np.frombuffer(raw, dtype="<u2") if raw is not None else np.asarray(i32, dtype="<u2")
        arr = (bits.astype(np.uint32) << 16).view(np.float32)
    elif dtype not in _DTYPES:
        raise ValueError(f"tensor '{name}': unsupported ONNX data type {dtype}")
    elif raw is not None:
        arr = np.frombuffer(raw, dtype=_DTYPES[dtype])
    elif f32:
        arr = np.concatenate(f32)
    elif f64:
        arr = np.concatenate(f64)
    elif i64:
        arr = np.asarray(i64, dtype="<i8")
    elif dtype == 10:                      # fp16 rides in int32_data as raw bit patterns
        arr = np.asarray(i32, dtype="<u2").view("<f2")
    else:
        arr = np.asarray(i32).astype(_DTYPES[dtype])
    n = int(np.prod(shape)) if shape else 1
    if arr.size != n:
        raise ValueError(f"tensor '{name}': {arr.size} elements for shape {shape}")
    return name, arr.reshape(shape)


@dataclass
class OnnxNode:
    op_type: str
    inputs: List[str]
    outputs: List[str]
    name: str = ""


@dataclass
class OnnxGraph:
    initializers: Dict[str, np.ndarray] = field(default_factory=dict)
    nodes: List[OnnxNode] = field(default_factory=list)
    inputs: List[str] = field(default_factory=list)
    outputs: List[str] = field(default_factory=list)


def _value_info_name(buf: memoryview) -> str:
    for num, _, v in _fields(buf):
        if num == 1:
            return bytes(v).decode("utf-8")
    return ""


def parse_model(data: bytes) -> OnnxGraph:
    """ModelProto.graph=7; GraphProto: node=1, initializer=5, input=11, output=12; NodeProto: input=1, output=2,
    name=3, op_type=4, attribute=5; AttributeProto: name=1, t=5.  `Constant` nodes are folded into initialisers."""
    g = OnnxGraph()
    graph = None
    for num, wt, v in _fields(memoryview(data)):
        if num == 7 and wt == 2:
            graph = v
    if graph is None:
        raise ValueError("not an ONNX model: no GraphProto (field 7)")
    for num, wt, v in _fields(graph):
        if num == 5 and wt == 2:
            name, arr = _decode_tensor(v)
            g.initializers[name] = arr
        elif num == 11 and wt == 2:
            g.inputs.append(_value_info_name(v))
        elif num == 12 and wt == 2:
            g.outputs.append(_value_info_name(v))
        elif num == 1 and wt == 2:
            node = OnnxNode("", [], [])
            const = None
            for n2, w2, v2 in _fields(v):
                if n2 == 1:
                    node.inputs.append(bytes(v2).decode("utf-8"))
                elif n2 == 2:
                    node.outputs.append(bytes(v2).decode("utf-8"))
                elif n2 == 3:
                    node.name = bytes(v2).decode("utf-8")
                elif n2 == 4:
                    node.op_type = bytes(v2).decode("utf-8")
                elif n2 == 5 and w2 == 2:
                    aname, at = "", None
                    for n3, w3, v3 in _fields(v2):
                        if n3 == 1:
                            aname = bytes(v3).decode("utf-8")
                        elif n3 == 5 and w3 == 2:
                            at = v3
                    if aname == "value" and at is not None:
                        const = at
            if node.op_type == "Constant" and const is not None and node.outputs:
                _, arr = _decode_tensor(const)
                g.initializers[node.outputs[0]] = arr
            g.nodes.append(node)
    # graph inputs that are also initialisers (keep_initializers_as_inputs) are not feeds
    g.inputs = [n for n in g.inputs if n not in g.initializers]
    return g


# --------------------------------------------------------------------------------------------------------------
# writer (used by the tests and by `export_initializers`: the inverse direction, blob -> ONNX container)
# --------------------------------------------------------------------------------------------------------------


def _enc_varint(x: int) -> bytes:
    x &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = x & 0x7F
        x >>= 7
        out.append(b | (0x80 if x else 0))
        if not x:
            return bytes(out)


def _enc_ld(num: int, payload: bytes) -> bytes:
    return _enc_varint((num << 3) | 2) + _enc_varint(len(payload)) + payload


def _enc_tensor(name: str, arr: np.ndarray, raw: bool = True) -> bytes:
    arr = np.asarray(arr)
    code = {"float32": 1, "int64": 7, "int32": 6, "float16": 10, "float64": 11}[arr.dtype.name]
    out = b"".join(_enc_varint((1 << 3) | 0) + _enc_varint(int(d)) for d in arr.shape)
    out += _enc_varint((2 << 3) | 0) + _enc_varint(code)
    out += _enc_ld(8, name.encode("utf-8"))
    if raw or code not in (1, 7):
        out += _enc_ld(9, np.ascontiguousarray(arr).astype(arr.dtype.newbyteorder("<")).tobytes())
    elif code == 1:
        out += _enc_ld(4, np.ascontiguousarray(arr, dtype="<f4").tobytes())
    else:
        out += _enc_ld(7, b"".join(_enc_varint(int(x)) for x in arr.reshape(-1)))
    return out


def write_model(initializers: Dict[str, np.ndarray], nodes: Iterable[OnnxNode] = (), inputs: Iterable[str] = (),
                outputs: Iterable[str] = (), raw: bool = True) -> bytes:
    """Smallest ModelProto that carries `initializers` and a node list (ir_version 8, one opset import)."""
    g = b""
    for n in nodes:
        body = b"".join(_enc_ld(1, s.encode()) for s in n.inputs) + b"".join(_enc_ld(2, s.encode()) for s in n.outputs)
        body += _enc_ld(3, n.name.encode()) + _enc_ld(4, n.op_type.encode())
        g += _enc_ld(1, body)
    g += _enc_ld(2, b"vvb200")
    for k, a in initializers.items():
        g += _enc_ld(5, _enc_tensor(k, a, raw))
    for s in inputs:
        g += _enc_ld(11, _enc_ld(1, s.encode()))
    for s in outputs:
        g += _enc_ld(12, _enc_ld(1, s.encode()))
    opset = _enc_ld(1, b"") + _enc_varint((2 << 3) | 0) + _enc_varint(17)
    return _enc_varint((1 << 3) | 0) + _enc_varint(8) + _enc_ld(2, b"vietvoice-tts-b200") + _enc_ld(7, g) + \
        _enc_ld(8, opset)


# --------------------------------------------------------------------------------------------------------------
# name mapping: upstream parameter names -> blob names
# --------------------------------------------------------------------------------------------------------------
# (regex on the END of the initialiser name so that any wrapper prefix — "transformer.", "ema_model.", "vocos." —
#  is ignored; `{i}` groups carry the layer index; "@" marks pieces that are fused afterwards.)
_NAME_RULES: List[Tuple[str, str]] = [
    # --- preprocess graph: text embedding (F5-TTS TextEmbedding + ConvNeXtV2Block)
    (r"text_embed\.text_embed\.weight$", "pre.text_embed"),
    (r"text_embed\.text_blocks\.(\d+)\.dwconv\.weight$", "pre.text_blocks.{0}.dw.w"),
    (r"text_embed\.text_blocks\.(\d+)\.dwconv\.bias$", "pre.text_blocks.{0}.dw.b"),
    (r"text_embed\.text_blocks\.(\d+)\.norm\.weight$", "pre.text_blocks.{0}.ln.g"),
    (r"text_embed\.text_blocks\.(\d+)\.norm\.bias$", "pre.text_blocks.{0}.ln.b"),
    (r"text_embed\.text_blocks\.(\d+)\.pwconv1\.weight$", "pre.text_blocks.{0}.pw1.w"),
    (r"text_embed\.text_blocks\.(\d+)\.pwconv1\.bias$", "pre.text_blocks.{0}.pw1.b"),
    (r"text_embed\.text_blocks\.(\d+)\.grn\.gamma$", "pre.text_blocks.{0}.grn.g"),
    (r"text_embed\.text_blocks\.(\d+)\.grn\.beta$", "pre.text_blocks.{0}.grn.b"),
    (r"text_embed\.text_blocks\.(\d+)\.pwconv2\.weight$", "pre.text_blocks.{0}.pw2.w"),
    (r"text_embed\.text_blocks\.(\d+)\.pwconv2\.bias$", "pre.text_blocks.{0}.pw2.b"),
    (r"mel_scale\.fb$", "pre.mel_fb"),
    # --- transformer graph (F5-TTS DiT)
    (r"time_embed\.time_mlp\.0\.weight$", "dit.time.l1.w"),
    (r"time_embed\.time_mlp\.0\.bias$", "dit.time.l1.b"),
    (r"time_embed\.time_mlp\.2\.weight$", "dit.time.l2.w"),
    (r"time_embed\.time_mlp\.2\.bias$", "dit.time.l2.b"),
    (r"input_embed\.proj\.weight$", "dit.in.w"),
    (r"input_embed\.proj\.bias$", "dit.in.b"),
    (r"conv_pos_embed\.conv1d\.0\.weight$", "dit.pos.c1.w"),
    (r"conv_pos_embed\.conv1d\.0\.bias$", "dit.pos.c1.b"),
    (r"conv_pos_embed\.conv1d\.2\.weight$", "dit.pos.c2.w"),
    (r"conv_pos_embed\.conv1d\.2\.bias$", "dit.pos.c2.b"),
    (r"transformer_blocks\.(\d+)\.attn_norm\.linear\.weight$", "dit.blocks.{0}.ada.w"),
    (r"transformer_blocks\.(\d+)\.attn_norm\.linear\.bias$", "dit.blocks.{0}.ada.b"),
    (r"transformer_blocks\.(\d+)\.attn\.to_q\.weight$", "dit.blocks.{0}.qkv.w@0"),
    (r"transformer_blocks\.(\d+)\.attn\.to_k\.weight$", "dit.blocks.{0}.qkv.w@1"),
    (r"transformer_blocks\.(\d+)\.attn\.to_v\.weight$", "dit.blocks.{0}.qkv.w@2"),
    (r"transformer_blocks\.(\d+)\.attn\.to_q\.bias$", "dit.blocks.{0}.qkv.b@0"),
    (r"transformer_blocks\.(\d+)\.attn\.to_k\.bias$", "dit.blocks.{0}.qkv.b@1"),
    (r"transformer_blocks\.(\d+)\.attn\.to_v\.bias$", "dit.blocks.{0}.qkv.b@2"),
    (r"transformer_blocks\.(\d+)\.attn\.to_qkv\.weight$", "dit.blocks.{0}.qkv.w"),
    (r"transformer_blocks\.(\d+)\.attn\.to_qkv\.bias$", "dit.blocks.{0}.qkv.b"),
    (r"transformer_blocks\.(\d+)\.attn\.to_out\.0\.weight$", "dit.blocks.{0}.out.w"),
    (r"transformer_blocks\.(\d+)\.attn\.to_out\.0\.bias$", "dit.blocks.{0}.out.b"),
    (r"transformer_blocks\.(\d+)\.ff\.ff\.0\.0\.weight$", "dit.blocks.{0}.ff1.w"),
    (r"transformer_blocks\.(\d+)\.ff\.ff\.0\.0\.bias$", "dit.blocks.{0}.ff1.b"),
    (r"transformer_blocks\.(\d+)\.ff\.ff\.2\.weight$", "dit.blocks.{0}.ff2.w"),
    (r"transformer_blocks\.(\d+)\.ff\.ff\.2\.bias$", "dit.blocks.{0}.ff2.b"),
    (r"norm_out\.linear\.weight$", "dit.final.ada.w"),
    (r"norm_out\.linear\.bias$", "dit.final.ada.b"),
    (r"proj_out\.weight$", "dit.out.w"),
    (r"proj_out\.bias$", "dit.out.b"),
    # --- decode graph (Vocos backbone + ISTFT head)
    (r"backbone\.embed\.weight$", "voc.embed.w"),
    (r"backbone\.embed\.bias$", "voc.embed.b"),
    (r"backbone\.norm\.weight$", "voc.norm.g"),
    (r"backbone\.norm\.bias$", "voc.norm.b"),
    (r"backbone\.convnext\.(\d+)\.dwconv\.weight$", "voc.blocks.{0}.dw.w"),
    (r"backbone\.convnext\.(\d+)\.dwconv\.bias$", "voc.blocks.{0}.dw.b"),
    (r"backbone\.convnext\.(\d+)\.norm\.weight$", "voc.blocks.{0}.ln.g"),
    (r"backbone\.convnext\.(\d+)\.norm\.bias$", "voc.blocks.{0}.ln.b"),
    (r"backbone\.convnext\.(\d+)\.pwconv1\.weight$", "voc.blocks.{0}.pw1.w"),
    (r"backbone\.convnext\.(\d+)\.pwconv1\.bias$", "voc.blocks.{0}.pw1.b"),
    (r"backbone\.convnext\.(\d+)\.pwconv2\.weight$", "voc.blocks.{0}.pw2.w"),
    (r"backbone\.convnext\.(\d+)\.pwconv2\.bias$", "voc.blocks.{0}.pw2.b"),
    (r"backbone\.convnext\.(\d+)\.gamma$", "voc.blocks.{0}.gamma"),
    (r"backbone\.final_layer_norm\.weight$", "voc.final.g"),
    (r"backbone\.final_layer_norm\.bias$", "voc.final.b"),
    (r"head\.out\.weight$", "voc.head.w"),
    (r"head\.out\.bias$", "voc.head.b"),
]
_RULES = [(re.compile(p), t) for p, t in _NAME_RULES]


def map_name(onnx_name: str) -> Optional[str]:
    for rx, target in _RULES:
        m = rx.search(onnx_name)
        if m:
            return target.format(*m.groups())
    return None


@dataclass
class ConversionReport:
    mapped: Dict[str, str] = field(default_factory=dict)       # blob name (or piece) -> ONNX initialiser name
    by_edge: List[str] = field(default_factory=list)           # blob names recovered through MatMul -> Add(bias)
    leftover: List[str] = field(default_factory=list)          # float initialisers of >= 2 dims nobody claimed
    missing: List[str] = field(default_factory=list)           # blob tensors the engine needs and nothing supplied
    computed: List[str] = field(default_factory=list)          # fixed tables rebuilt from the architecture
    arch_from_shapes: Dict[str, int] = field(default_factory=dict)

    def ok(self) -> bool:
        return not self.missing


def _as_f32(a: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _collect(graph: OnnxGraph, out: Dict[str, np.ndarray], rep: ConversionReport) -> None:
    claimed = set()
    for name, arr in graph.initializers.items():
        t = map_name(name)
        if t is not None and arr.dtype.kind == "f":
            out[t] = _as_f32(arr)
            rep.mapped[t] = name
            claimed.add(name)
    # anonymous MatMul weights: MatMul(x, W[in,out]) -> Add(., X.bias)  =>  X.weight = W^T
    consumers: Dict[str, List[OnnxNode]] = {}
    for n in graph.nodes:
        for i in n.inputs:
            consumers.setdefault(i, []).append(n)
    for n in graph.nodes:
        if n.op_type != "MatMul" or len(n.inputs) != 2 or not n.outputs:
            continue
        w_name = n.inputs[1]
        w = graph.initializers.get(w_name)
        if w is None or w_name in claimed or w.ndim != 2 or w.dtype.kind != "f":
            continue
        for c in consumers.get(n.outputs[0], []):
            if c.op_type != "Add":
                continue
            for b_name in c.inputs:
                bt = map_name(b_name) if b_name in graph.initializers else None
                if bt is None:
                    continue
                base, _, piece = bt.partition("@")
                if not base.endswith(".b"):
                    continue
                wt = base[:-2] + ".w" + ("@" + piece if piece else "")
                if wt in out or w.shape[1] != graph.initializers[b_name].size:
                    continue
                out[wt] = _as_f32(w.T)
                rep.mapped[wt] = w_name
                rep.by_edge.append(wt)
                claimed.add(w_name)
    for name, arr in graph.initializers.items():
        if name not in claimed and arr.dtype.kind == "f" and arr.ndim >= 2 and arr.size >= 64:
            rep.leftover.append(name)


def _finish(t: Dict[str, np.ndarray], rep: ConversionReport) -> Dict[str, np.ndarray]:
    """Fuse q/k/v pieces, drop singleton axes, orient the fixed tables."""
    out: Dict[str, np.ndarray] = {}
    pieces: Dict[str, Dict[int, np.ndarray]] = {}
    for k, a in t.items():
        base, sep, idx = k.partition("@")
        if sep:
            pieces.setdefault(base, {})[int(idx)] = a
        else:
            out[k] = a
    for base, ps in pieces.items():
        if base in out:
            continue
        if sorted(ps) == [0, 1, 2]:
            out[base] = np.concatenate([ps[0], ps[1], ps[2]], axis=0)
    for k in list(out):
        a = out[k]
        if k.endswith(".dw.w") and a.ndim == 3 and a.shape[1] == 1:          # depthwise [C,1,k] -> [C,k]
            out[k] = np.ascontiguousarray(a[:, 0, :])
        elif (".grn." in k or k.endswith(".gamma")) and a.ndim > 1:          # [1,1,C] -> [C]
            out[k] = np.ascontiguousarray(a.reshape(-1))
        elif k.endswith((".pw1.w", ".pw2.w")) and a.ndim == 3 and a.shape[2] == 1:   # 1x1 conv form of a Linear
            out[k] = np.ascontiguousarray(a[:, :, 0])
    return out


def infer_arch(t: Dict[str, np.ndarray], base: ArchConfig = FULL) -> Tuple[ArchConfig, Dict[str, int]]:
    """Read the constants the weight shapes determine; everything else (heads via head_dim 64, RoPE, sampler,
    mel front-end) stays as in `base`."""
    got: Dict[str, int] = {}

    def layers(prefix: str) -> int:
        idx = [int(m.group(1)) for k in t for m in [re.match(re.escape(prefix) + r"\.(\d+)\.", k)] if m]
        return max(idx) + 1 if idx else 0

    if "dit.out.w" in t:
        got["n_mel"], got["dim"] = (int(x) for x in t["dit.out.w"].shape)
    if layers("dit.blocks"):
        got["depth"] = layers("dit.blocks")
    if "dit.blocks.0.ff1.w" in t:
        got["ff_dim"] = int(t["dit.blocks.0.ff1.w"].shape[0])
    if "dit.time.l1.w" in t:
        got["time_freq_dim"] = int(t["dit.time.l1.w"].shape[1])
    if "dit.pos.c1.w" in t:
        d, cg, k = (int(x) for x in t["dit.pos.c1.w"].shape)
        got["conv_pos_k"], got["conv_pos_groups"] = k, d // cg
    if "pre.text_embed" in t:
        got["vocab"], got["text_dim"] = int(t["pre.text_embed"].shape[0]) - 1, int(t["pre.text_embed"].shape[1])
    if layers("pre.text_blocks"):
        got["text_layers"] = layers("pre.text_blocks")
    if "pre.text_blocks.0.pw1.w" in t:
        got["text_ff"] = int(t["pre.text_blocks.0.pw1.w"].shape[0])
    if "voc.embed.w" in t:
        got["voc_dim"], _, got["voc_k"] = (int(x) for x in t["voc.embed.w"].shape)
    if layers("voc.blocks"):
        got["voc_layers"] = layers("voc.blocks")
    if "voc.blocks.0.pw1.w" in t:
        got["voc_ff"] = int(t["voc.blocks.0.pw1.w"].shape[0])
    if "voc.head.w" in t:
        got["n_fft"] = int(t["voc.head.w"].shape[0]) - 2
    if "dim" in got:
        got["heads"] = got["dim"] // base.head_dim
    arch = replace(base, **got)
    return arch, got


def convert_graphs(graphs: Dict[str, bytes], base: ArchConfig = FULL, strict: bool = True
                   ) -> Tuple[ArchConfig, Dict[str, np.ndarray], ConversionReport]:
    """`graphs` maps 'preprocess' / 'transformer' / 'decode' to the bytes of the ONNX files."""
    rep = ConversionReport()
    raw: Dict[str, np.ndarray] = {}
    for key in ("preprocess", "transformer", "decode"):
        if key in graphs:
            _collect(parse_model(graphs[key]), raw, rep)
    tensors = _finish(raw, rep)
    arch, rep.arch_from_shapes = infer_arch(tensors, base)
    # the mel filter bank: by name, else by shape, else rebuilt from the architecture
    if "pre.mel_fb" not in tensors:
        for key in ("preprocess",):
            if key not in graphs:
                continue
            for name, a in parse_model(graphs[key]).initializers.items():
                if a.dtype.kind == "f" and a.shape in ((arch.n_bins, arch.n_mel), (arch.n_mel, arch.n_bins)):
                    tensors["pre.mel_fb"] = _as_f32(a if a.shape[0] == arch.n_bins else a.T)
                    rep.mapped["pre.mel_fb"] = name
                    if name in rep.leftover:
                        rep.leftover.remove(name)
                    break
    if "pre.mel_fb" not in tensors:
        tensors["pre.mel_fb"] = artifact.mel_filterbank(arch)
        rep.computed.append("pre.mel_fb")
    elif tensors["pre.mel_fb"].shape == (arch.n_mel, arch.n_bins):
        tensors["pre.mel_fb"] = np.ascontiguousarray(tensors["pre.mel_fb"].T)
    # what the engine needs = the names (and shapes) `make_random_weights` produces for this architecture
    want = expected_shapes(arch)
    for k, shp in want.items():
        if k not in tensors:
            rep.missing.append(k)
        elif tuple(tensors[k].shape) != shp:
            raise ValueError(f"{k}: shape {tuple(tensors[k].shape)} from '{rep.mapped.get(k, '?')}', expected {shp}")
    if strict and rep.missing:
        raise ValueError("ONNX graphs do not supply: " + ", ".join(rep.missing[:8]) +
                         (" ..." if len(rep.missing) > 8 else ""))
    arch.validate()
    return arch, {k: tensors[k] for k in want if k in tensors}, rep


_SHAPE_CACHE: Dict[ArchConfig, Dict[str, Tuple[int, ...]]] = {}


def expected_shapes(arch: ArchConfig) -> Dict[str, Tuple[int, ...]]:
    """Blob tensor names and shapes for `arch` (one source of truth: artifact.make_random_weights)."""
    if arch not in _SHAPE_CACHE:
        small = replace(arch, depth=min(arch.depth, 2), text_layers=min(arch.text_layers, 2),
                        voc_layers=min(arch.voc_layers, 2))
        w = artifact.make_random_weights(small, seed=0)
        shapes: Dict[str, Tuple[int, ...]] = {}
        for k, a in w.items():
            m = re.match(r"(dit\.blocks|pre\.text_blocks|voc\.blocks)\.(\d+)\.(.*)", k)
            if m is None:
                shapes[k] = tuple(a.shape)
            elif m.group(2) == "0":
                n = {"dit.blocks": arch.depth, "pre.text_blocks": arch.text_layers, "voc.blocks": arch.voc_layers}
                for i in range(n[m.group(1)]):
                    shapes[f"{m.group(1)}.{i}.{m.group(3)}"] = tuple(a.shape)
        _SHAPE_CACHE[arch] = shapes
    return _SHAPE_CACHE[arch]


_GRAPH_FILES = {"preprocess": "preprocess.onnx", "transformer": "transformer.onnx", "decode": "decode.onnx"}


def convert_model_tar(src: str, dst: str, base: ArchConfig = FULL, strict: bool = True) -> ConversionReport:
    """Rewrite the reference's model tar (core/model.py:73-129): the three `.onnx` members become VVB200 blobs under
    the same member names, every other member is copied unchanged."""
    with tarfile.open(src, "r") as tin:
        members = tin.getmembers()
        graphs: Dict[str, bytes] = {}
        where: Dict[str, str] = {}
        for key, fn in _GRAPH_FILES.items():
            m = next((m for m in members if m.name.endswith(fn)), None)
            if m is None:
                raise FileNotFoundError(f"Model file '{fn}' not found in model archive")
            data = tin.extractfile(m).read()
            if data[:8] == artifact.MAGIC:
                raise ValueError(f"{m.name} already is a VVB200 blob")
            graphs[key] = data
            where[m.name] = key
        arch, tensors, rep = convert_graphs(graphs, base, strict)
        with tarfile.open(dst, "w") as tout:
            for m in members:
                if m.name in where:
                    blob = artifact.pack_blob(arch, tensors, where[m.name])
                    ti = tarfile.TarInfo(m.name)
                    ti.size = len(blob)
                    tout.addfile(ti, io.BytesIO(blob))
                elif m.isfile():
                    tout.addfile(m, tin.extractfile(m))
                else:
                    tout.addfile(m)
    return rep


# --------------------------------------------------------------------------------------------------------------
# inverse direction: blob tensors -> ONNX containers under upstream names (fixtures for the tests; also lets a
# maintainer with onnxruntime load OUR seeded weights into the upstream graphs' initialisers)
# --------------------------------------------------------------------------------------------------------------
_UPSTREAM = {
    "pre.text_embed": "text_embed.text_embed.weight",
    "pre.mel_fb": "mel_spec.mel_scale.fb",
    "dit.time.l1": "time_embed.time_mlp.0", "dit.time.l2": "time_embed.time_mlp.2",
    "dit.in": "input_embed.proj", "dit.pos.c1": "input_embed.conv_pos_embed.conv1d.0",
    "dit.pos.c2": "input_embed.conv_pos_embed.conv1d.2", "dit.final.ada": "norm_out.linear",
    "dit.out": "proj_out", "voc.embed": "backbone.embed", "voc.head": "head.out",
}
_BLOCK_PARTS = {
    "dit.blocks": ("transformer_blocks", {"ada": "attn_norm.linear", "out": "attn.to_out.0", "ff1": "ff.ff.0.0",
                                          "ff2": "ff.ff.2"}),
    "pre.text_blocks": ("text_embed.text_blocks", {"dw": "dwconv", "pw1": "pwconv1", "pw2": "pwconv2"}),
    "voc.blocks": ("backbone.convnext", {"dw": "dwconv", "pw1": "pwconv1", "pw2": "pwconv2"}),
}


def export_initializers(arch: ArchConfig, tensors: Dict[str, np.ndarray], prefix: str = "transformer.",
                        anonymous_matmul: bool = False, raw: bool = True) -> Dict[str, bytes]:
    """Blob tensors -> three ONNX containers with upstream-style names.  `anonymous_matmul` stores every Linear
    weight the way torch.onnx does (transposed, `onnx::MatMul_<n>`, wired MatMul -> Add(bias))."""
    files: Dict[str, Tuple[Dict[str, np.ndarray], List[OnnxNode]]] = {k: ({}, []) for k in _GRAPH_FILES}
    counter = [0]

    def put(graph: str, name: str, a: np.ndarray) -> None:
        files[graph][0][prefix + name if graph != "decode" else "vocos." + name] = np.asarray(a, dtype=np.float32)

    def linear(graph: str, up: str, w: np.ndarray, b: np.ndarray) -> None:
        if anonymous_matmul and w.ndim == 2:
            counter[0] += 1
            wn = f"onnx::MatMul_{counter[0]}"
            files[graph][0][wn] = np.ascontiguousarray(w.T, dtype=np.float32)
            put(graph, up + ".bias", b)
            bn = (prefix if graph != "decode" else "vocos.") + up + ".bias"
            y = f"/{up}/MatMul_output_0"
            files[graph][1].append(OnnxNode("MatMul", [f"/{up}/in", wn], [y], f"/{up}/MatMul"))
            files[graph][1].append(OnnxNode("Add", [bn, y], [f"/{up}/Add_output_0"], f"/{up}/Add"))
        else:
            put(graph, up + ".weight", w)
            put(graph, up + ".bias", b)

    def graph_of(k: str) -> str:
        return {"pre": "preprocess", "dit": "transformer", "voc": "decode"}[k[:3]]

    d = arch.dim
    done = set()
    for k in tensors:
        if k in done:
            continue
        g = graph_of(k)
        m = re.match(r"(dit\.blocks|pre\.text_blocks|voc\.blocks)\.(\d+)\.(\w+)(?:\.(\w+))?$", k)
        if m:
            grp, i, part, leaf = m.group(1), m.group(2), m.group(3), m.group(4)
            up_blk, parts = _BLOCK_PARTS[grp]
            stem = f"{up_blk}.{i}"
            if part == "qkv":
                w, b = tensors[f"{grp}.{i}.qkv.w"], tensors[f"{grp}.{i}.qkv.b"]
                for j, nm in enumerate(("to_q", "to_k", "to_v")):
                    linear(g, f"{stem}.attn.{nm}", w[j * d:(j + 1) * d], b[j * d:(j + 1) * d])
                done.update({f"{grp}.{i}.qkv.w", f"{grp}.{i}.qkv.b"})
            elif part == "dw":
                put(g, f"{stem}.dwconv.weight", tensors[f"{grp}.{i}.dw.w"][:, None, :])
                put(g, f"{stem}.dwconv.bias", tensors[f"{grp}.{i}.dw.b"])
                done.update({f"{grp}.{i}.dw.w", f"{grp}.{i}.dw.b"})
            elif part == "ln":
                put(g, f"{stem}.norm.weight", tensors[f"{grp}.{i}.ln.g"])
                put(g, f"{stem}.norm.bias", tensors[f"{grp}.{i}.ln.b"])
                done.update({f"{grp}.{i}.ln.g", f"{grp}.{i}.ln.b"})
            elif part == "grn":
                put(g, f"{stem}.grn.gamma", tensors[f"{grp}.{i}.grn.g"][None, None, :])
                put(g, f"{stem}.grn.beta", tensors[f"{grp}.{i}.grn.b"][None, None, :])
                done.update({f"{grp}.{i}.grn.g", f"{grp}.{i}.grn.b"})
            elif part == "gamma":
                put(g, f"{stem}.gamma", tensors[k])
                done.add(k)
            else:
                linear(g, f"{stem}.{parts[part]}", tensors[f"{grp}.{i}.{part}.w"], tensors[f"{grp}.{i}.{part}.b"])
                done.update({f"{grp}.{i}.{part}.w", f"{grp}.{i}.{part}.b"})
            continue
        if k in ("pre.text_embed", "pre.mel_fb"):
            put(g, _UPSTREAM[k], tensors[k])
        elif k in ("voc.norm.g", "voc.norm.b"):
            put(g, "backbone.norm." + ("weight" if k.endswith(".g") else "bias"), tensors[k])
        elif k in ("voc.final.g", "voc.final.b"):
            put(g, "backbone.final_layer_norm." + ("weight" if k.endswith(".g") else "bias"), tensors[k])
        else:
            stem = k[:-2]
            if stem in ("dit.pos.c1", "dit.pos.c2", "voc.embed"):          # real convolutions keep their names
                put(g, _UPSTREAM[stem] + ".weight", tensors[stem + ".w"])
                put(g, _UPSTREAM[stem] + ".bias", tensors[stem + ".b"])
            else:
                linear(g, _UPSTREAM[stem], tensors[stem + ".w"], tensors[stem + ".b"])
            done.update({stem + ".w", stem + ".b"})
        done.add(k)
    return {g: write_model(init, nodes, raw=raw) for g, (init, nodes) in files.items()}


def main(argv: Optional[List[str]] = None) -> int:
    import argparse
    ap = argparse.ArgumentParser(description="upstream model-bin.pt (tar of ONNX graphs) -> VVB200 model tar")
    ap.add_argument("src")
    ap.add_argument("dst")
    ap.add_argument("--lenient", action="store_true", help="write the tar even if blob tensors are missing")
    a = ap.parse_args(argv)
    rep = convert_model_tar(a.src, a.dst, strict=not a.lenient)
    print(f"mapped {len(rep.mapped)} tensors ({len(rep.by_edge)} through MatMul->Add edges), "
          f"{len(rep.leftover)} left over, {len(rep.missing)} missing; arch from shapes: {rep.arch_from_shapes}")
    for n in rep.leftover[:20]:
        print("  leftover:", n)
    return 0 if rep.ok() else 1


if __name__ == "__main__":
    raise SystemExit(main())
