"""An `onnxruntime`-shaped module backed by the B200 engine.

The reference touches exactly eight symbols of onnxruntime:
  get_available_providers, SessionOptions (+ add_session_config_entry), ExecutionMode, GraphOptimizationLevel,
  InferenceSession(model_bytes, sess_options=, providers=), session.get_inputs()/get_outputs() (-> .name),
  session.run(output_names, feeds), set_seed
(/root/reference/vietvoicetts/core/model.py:33,52-62,98-106,133; core/tts_engine.py:146,172,187;
deterministic.py:29).  Installing this module as `sys.modules["onnxruntime"]` (see `install()`) lets the reference's
own Python run unmodified on top of libvvb200.so; feeds are bound BY POSITION exactly as the reference builds them.

This is the slow, call-compatible path (numpy in / numpy out for each of the 33 session calls per chunk).  The fast
path is `host.tts_engine.TTSEngine`, which hands whole chunks to `Engine.synthesize_batch`.
"""
from __future__ import annotations

import os
import sys
import threading
from typing import Dict, List, Optional, Sequence

import numpy as np

from .arch import ArchConfig, VVArch
from .artifact import HEADER_BYTES, MAGIC
from .engine import Batch, Engine

__version__ = "1.20.2+vvb200"

_seed = 9527
_lock = threading.RLock()
_engines: Dict[tuple, "_Shared"] = {}


def get_available_providers() -> List[str]:
    return ["CUDAExecutionProvider", "CPUExecutionProvider"]


def get_device() -> str:
    return "GPU"


def set_seed(seed: int) -> None:
    """ort.set_seed: seeds the y0 draw of the preprocess graph (core/model.py:133)."""
    global _seed
    _seed = int(seed)


class ExecutionMode:
    ORT_SEQUENTIAL = 0
    ORT_PARALLEL = 1


class GraphOptimizationLevel:
    ORT_DISABLE_ALL = 0
    ORT_ENABLE_BASIC = 1
    ORT_ENABLE_EXTENDED = 2
    ORT_ENABLE_ALL = 99


class SessionOptions:
    def __init__(self):
        self.log_severity_level = 2
        self.log_verbosity_level = 0
        self.inter_op_num_threads = 0
        self.intra_op_num_threads = 0
        self.enable_cpu_mem_arena = True
        self.execution_mode = ExecutionMode.ORT_SEQUENTIAL
        self.graph_optimization_level = GraphOptimizationLevel.ORT_ENABLE_ALL
        self._entries: Dict[str, str] = {}

    def add_session_config_entry(self, key: str, value: str) -> None:
        self._entries[str(key)] = str(value)

    def get_session_config_entry(self, key: str) -> str:
        return self._entries[key]


class NodeArg:
    def __init__(self, name: str, type_: str, shape):
        self.name = name
        self.type = type_
        self.shape = shape


class _Shared:
    """One engine per (device, architecture), shared by the three sessions of a ModelSessionManager."""

    def __init__(self, arch: ArchConfig, device: int):
        self.arch = arch
        self.engine = Engine(arch, device=device)
        self.loaded: set = set()
        self.batches: Dict[int, Batch] = {}
        self.calls = 0

    def ensure_final(self):
        if not self.engine._finalized:
            missing = {0, 1, 2} - self.loaded
            if missing:
                names = {0: "preprocess", 1: "transformer", 2: "decode"}
                raise RuntimeError("B200 engine needs all three graphs loaded before the first run; missing: "
                                   + ", ".join(names[m] for m in sorted(missing)))
            self.engine.finalize()

    def batch_for(self, T: int) -> Batch:
        b = self.batches.get(T)
        if b is None:
            if len(self.batches) >= 8:                      # bound the cache
                old = next(iter(self.batches))
                self.batches.pop(old).close()
            b = self.engine.batch([T])
            b._prepped = False
            self.batches[T] = b
        return b


_GRAPHS = {
    0: ("preprocess",
        [("audio", "tensor(int16)", [1, 1, "N"]), ("text_ids", "tensor(int32)", [1, "L"]),
         ("max_duration", "tensor(int64)", [1])],
        [("noise", "tensor(float)"), ("rope_cos_q", "tensor(float)"), ("rope_sin_q", "tensor(float)"),
         ("rope_cos_k", "tensor(float)"), ("rope_sin_k", "tensor(float)"), ("cat_mel_text", "tensor(float)"),
         ("cat_mel_text_drop", "tensor(float)"), ("ref_signal_len", "tensor(int64)")]),
    1: ("transformer",
        [("noise", "tensor(float)", [1, "T", 100]), ("rope_cos_q", "tensor(float)", [1, "T", 64]),
         ("rope_sin_q", "tensor(float)", [1, "T", 64]), ("rope_cos_k", "tensor(float)", [1, 64, "T"]),
         ("rope_sin_k", "tensor(float)", [1, 64, "T"]), ("cat_mel_text", "tensor(float)", [1, "T", 612]),
         ("cat_mel_text_drop", "tensor(float)", [1, "T", 612]), ("time_step", "tensor(int32)", [1])],
        [("noise_out", "tensor(float)"), ("time_step_out", "tensor(int32)")]),
    2: ("decode",
        [("denoised", "tensor(float)", [1, "T", 100]), ("ref_signal_len", "tensor(int64)", [1])],
        [("output_audio", "tensor(int16)")]),
}


def _rope_tables(arch: ArchConfig, T: int):
    hd = arch.head_dim
    inv = 1.0 / (float(arch.rope_theta) ** (np.arange(0, hd, 2, dtype=np.float64) / hd))
    ang = np.repeat(np.arange(T, dtype=np.float64)[:, None] * inv[None, :], 2, axis=-1)
    cos, sin = np.cos(ang).astype(np.float32), np.sin(ang).astype(np.float32)
    return cos[None], sin[None], np.ascontiguousarray(cos.T)[None], np.ascontiguousarray(sin.T)[None]


class InferenceSession:
    def __init__(self, path_or_bytes, sess_options: Optional[SessionOptions] = None,
                 providers: Optional[Sequence] = None, provider_options=None, **kwargs):
        if isinstance(path_or_bytes, str):
            with open(path_or_bytes, "rb") as f:
                blob = f.read()
        else:
            blob = bytes(path_or_bytes)
        if blob[:8] != MAGIC:
            raise RuntimeError("InvalidProtobuf: not a VVB200 weight blob (this executor does not run ONNX graphs; "
                               "build the artefact with vietvoice_tts_b200.artifact)")
        self._gid = int(np.frombuffer(blob, dtype="<u4", count=1, offset=12)[0])
        if self._gid not in _GRAPHS:
            raise RuntimeError("a session must be built from a single-graph blob (preprocess / transformer / decode)")
        carch = VVArch.from_buffer_copy(blob[16:16 + np.dtype("<i4").itemsize * 33])
        arch = ArchConfig(**{f: getattr(carch, f) for f, _ in VVArch._fields_})
        self._options = sess_options or SessionOptions()
        self._providers = list(providers) if providers else get_available_providers()
        device = int(os.environ.get("VVB200_DEVICE", os.environ.get("LOCAL_RANK", "0")))   # one process per GPU
        for p in self._providers:
            if isinstance(p, tuple) and isinstance(p[1], dict) and "device_id" in p[1]:
                device = int(p[1]["device_id"])
        self._fuse = int(self._options._entries.get("vvb200.fuse_nfe", "1"))
        with _lock:
            key = (device, arch)
            sh = _engines.get(key)
            if sh is None or self._gid in sh.loaded:       # a second model set -> a fresh engine
                sh = _Shared(arch, device)
                _engines[key] = sh
            sh.engine.load_blob(blob)
            sh.loaded.add(self._gid)
            self._sh = sh
        name, ins, outs = _GRAPHS[self._gid]
        self._name = name
        self._inputs = [NodeArg(n, t, s) for n, t, s in ins]
        self._outputs = [NodeArg(n, t, None) for n, t in outs]

    # ---- introspection the reference uses (core/model.py:105-106)
    def get_inputs(self):
        return list(self._inputs)

    def get_outputs(self):
        return list(self._outputs)

    def get_providers(self):
        return list(self._providers)

    # ---- execution
    def run(self, output_names, input_feed: Dict[str, np.ndarray], run_options=None):
        sh = self._sh
        try:
            feeds = [input_feed[a.name] for a in self._inputs]
        except KeyError as exc:
            raise ValueError(f"Required inputs ({[a.name for a in self._inputs]}) are missing from input feed: {exc}")
        with _lock:
            sh.ensure_final()
            if self._gid == 0:
                outs = self._run_preprocess(*feeds)
            elif self._gid == 1:
                outs = self._run_transformer(*feeds)
            else:
                outs = self._run_decode(*feeds)
        if output_names:
            order = {a.name: i for i, a in enumerate(self._outputs)}
            return [outs[order[n]] for n in output_names]
        return outs

    def _run_preprocess(self, audio, text_ids, max_duration):
        sh, a = self._sh, self._sh.arch
        T = int(np.asarray(max_duration).reshape(-1)[0])
        b = sh.batch_for(T)
        ref_len = b.preprocess(0, np.asarray(audio), np.asarray(text_ids), None, seed=_seed, chunk_key=sh.calls)
        sh.calls += 1
        b._prepped = True
        b._cond_ids = None
        cq, sq, ck, sk = _rope_tables(a, T)
        cat_c, cat_u = b.get(0, "cat_mel_text")[None], b.get(0, "cat_mel_text_drop")[None]
        b._cond_ids = (id(cat_c), id(cat_u))
        b._cond_keep = (cat_c, cat_u)
        return [b.get(0, "noise")[None], cq, sq, ck, sk, cat_c, cat_u, np.array([ref_len], dtype=np.int64)]

    def _standalone(self, T: int) -> Batch:
        b = self._sh.batch_for(T)
        if not getattr(b, "_prepped", False):      # session used without a preceding preprocess call
            b.preprocess(0, np.zeros(self._sh.arch.n_fft, dtype=np.int16), np.zeros(0, dtype=np.int32),
                         np.zeros((T, self._sh.arch.n_mel), dtype=np.float32))
            b._prepped = True
            b._cond_ids = None
        return b

    def _run_transformer(self, noise, rcq, rsq, rck, rsk, cat_c, cat_u, time_step):
        noise = np.asarray(noise, dtype=np.float32)
        T = noise.shape[1]
        b = self._standalone(T)
        if getattr(b, "_cond_ids", None) != (id(cat_c), id(cat_u)):      # new conditioning arrays -> re-project
            b.set_cond(0, np.asarray(cat_c)[0], np.asarray(cat_u)[0])
            b._cond_ids = (id(cat_c), id(cat_u))
            b._cond_keep = (cat_c, cat_u)
        step = int(np.asarray(time_step).reshape(-1)[0])
        b.set_noise(0, noise[0])
        b.sample(first_step=step, n_steps=self._fuse)
        return [b.get(0, "noise")[None], np.array([step + self._fuse], dtype=np.int32)]

    def _run_decode(self, denoised, ref_signal_len):
        denoised = np.asarray(denoised, dtype=np.float32)
        T = denoised.shape[1]
        b = self._standalone(T)
        b.set_noise(0, denoised[0])
        b.set_ref_len(0, int(np.asarray(ref_signal_len).reshape(-1)[0]))
        return [b.decode(0).reshape(1, 1, -1)]


def install(audio: bool = True) -> None:
    """Make `import onnxruntime` resolve to this module (for running the reference's Python unmodified).  With
    `audio`, `pydub` and `soundfile` — imported by the reference at module load, absent on offline hosts — get
    WAV-only stand-ins (audio_shims.py) unless the real packages are importable."""
    sys.modules["onnxruntime"] = sys.modules[__name__]
    if audio:
        from . import audio_shims
        audio_shims.install()
