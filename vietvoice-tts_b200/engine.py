"""Host-side handles over the C ABI: `Engine` (weights + tables on one GPU) and `Batch` (B chunks in flight).

`Engine.run_chunks` is the device-resident fast path that stands in for the three `session.run` wrappers of
the reference (`TTSEngine._run_preprocess/_run_transformer_steps/_run_decode`,
/root/reference/vietvoicetts/core/tts_engine.py:133-187) for B independent chunks at once.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

from . import _lib
from .arch import ArchConfig
from .artifact import pack_blob


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def fade_tables(n_fade: int):
    """The float64 cos^2 / sin^2 fade tables exactly as the reference computes them
    (/root/reference/vietvoicetts/core/audio_processor.py:175-176); numpy evaluates them on the host because cos / sin
    differ in the last bit between math libraries and the device cross-fade must reproduce the host's bits."""
    theta = np.linspace(0, np.pi / 2, n_fade)
    return np.ascontiguousarray(np.cos(theta) ** 2), np.ascontiguousarray(np.sin(theta) ** 2)


class Engine:
    def __init__(self, arch: ArchConfig, device: int = 0, stream: Optional[int] = None):
        arch.validate()
        if stream == 0:
            raise ValueError("stream=0 is the legacy default stream (no CUDA-graph capture, implicit syncs): pass a "
                             "non-default cudaStream_t handle, or None for an engine-owned stream")
        self.lib = _lib.load()
        self.arch = arch
        self.device = device
        self._h = C.c_void_p()
        carch = arch.to_c()
        _lib.check(self.lib.vv_engine_create(C.byref(carch), device, C.c_void_p(stream) if stream else None,
                                             C.byref(self._h)))
        self._finalized = False

    # -- construction ----------------------------------------------------------------------------
    @classmethod
    def from_weights(cls, arch: ArchConfig, weights: Dict[str, np.ndarray], device: int = 0,
                     stream: Optional[int] = None) -> "Engine":
        e = cls(arch, device, stream)
        for graph in ("preprocess", "transformer", "decode"):
            e.load_blob(pack_blob(arch, weights, graph))
        e.finalize()
        return e

    @classmethod
    def from_blobs(cls, blobs: Iterable[bytes], device: int = 0, stream: Optional[int] = None) -> "Engine":
        blobs = list(blobs)
        from .artifact import MAGIC, HEADER_BYTES
        from .arch import VVArch
        head = blobs[0][:HEADER_BYTES]
        if head[:8] != MAGIC:
            raise ValueError("not a VVB200 weight blob")
        carch = VVArch.from_buffer_copy(head[16:16 + C.sizeof(VVArch)])
        arch = ArchConfig(**{f: getattr(carch, f) for f, _ in VVArch._fields_})
        e = cls(arch, device, stream)
        for b in blobs:
            e.load_blob(b)
        e.finalize()
        return e

    def load_blob(self, blob: bytes) -> None:
        buf = (C.c_char * len(blob)).from_buffer_copy(blob)
        _lib.check(self.lib.vv_engine_load_blob(self._h, C.cast(buf, C.c_void_p), len(blob)))

    def finalize(self) -> None:
        _lib.check(self.lib.vv_engine_finalize(self._h))
        self._finalized = True

    def close(self) -> None:
        if self._h:
            self.lib.vv_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- info -------------------------------------------------------------------------------------
    @property
    def launch_count(self) -> int:
        return int(self.lib.vv_engine_launch_count(self._h))

    @property
    def stream(self) -> int:
        return int(self.lib.vv_engine_stream(self._h) or 0)

    def sync(self) -> None:
        _lib.check(self.lib.vv_sync(self._h))

    def batch(self, total_frames: Sequence[int]) -> "Batch":
        return Batch(self, total_frames)

    # -- device-resident prompts -------------------------------------------------
    def prompt_put(self, audio: np.ndarray) -> int:
        """Make a prompt resident in HBM (PCM + log-mel + ref_signal_len) and return its content id.  A prompt that
        is already resident is not uploaded again."""
        a = np.ascontiguousarray(np.asarray(audio).reshape(-1), dtype=np.int16)
        pid = C.c_uint64(0)
        _lib.check(self.lib.vv_prompt_put(self._h, _ptr(a), a.size, C.byref(pid), None))
        return int(pid.value)

    def prompt_drop(self, prompt_id: int) -> None:
        _lib.check(self.lib.vv_prompt_drop(self._h, int(prompt_id)))

    def prompt_cache_clear(self) -> None:
        _lib.check(self.lib.vv_prompt_cache_clear(self._h))

    def prompt_cache_stats(self) -> Dict[str, int]:
        w = (C.c_int64 * 3)()
        _lib.check(self.lib.vv_prompt_cache_stats(self._h, w))
        return {"resident": int(w[0]), "uploads": int(w[1]), "hits": int(w[2])}

    # -- cross-fade on the device ------------------------------------------------
    def crossfade_ok(self, lengths: Sequence[int], n_fade: int) -> bool:
        """the device cross-fade covers the regular case: >= 2 chunks, each at least two fades long"""
        return len(lengths) >= 2 and 1 <= n_fade <= 6144 and all(int(n) >= 2 * n_fade for n in lengths)

    def crossfade_pcm(self, waves: Sequence[np.ndarray], n_fade: int) -> np.ndarray:
        """`concatenate_with_crossfade_improved` of host int16 chunks, computed on the GPU, bit-exact with numpy."""
        ws = [np.ascontiguousarray(np.asarray(w).reshape(-1), dtype=np.int16) for w in waves]
        n = len(ws)
        fo, fi = fade_tables(n_fade)
        ptrs = (C.c_void_p * n)(*[w.ctypes.data for w in ws])
        lens = (C.c_int64 * n)(*[w.size for w in ws])
        out = np.empty(sum(w.size for w in ws), dtype=np.int16)
        got = C.c_int64(0)
        _lib.check(self.lib.vv_crossfade_pcm(self._h, ptrs, lens, n, _ptr(fo), _ptr(fi), n_fade, _ptr(out), out.size,
                                             C.byref(got)))
        return out[: int(got.value)]

    # -- whole path, host buffers ------------------------------------------------------------------
    def synthesize_batch(self, audios: Sequence[np.ndarray], text_ids: Sequence[np.ndarray],
                         total_frames: Sequence[int], noises: Optional[Sequence[Optional[np.ndarray]]] = None,
                         nfe: int = 0, seed: int = 9527, chunk_keys: Optional[Sequence[int]] = None,
                         pcm_out: Optional[Sequence[np.ndarray]] = None,
                         prompt_ids: Optional[Sequence[int]] = None, join_fade: int = 0):
        """One call = preprocess -> (nfe-1) steps -> decode for B chunks; host arrays in, int16 PCM out.
        `prompt_ids[i]` (from `prompt_put`) names a resident prompt; without it the engine finds the prompt by hashing
        audios[i] — chunks that pass the SAME array object are hashed once.
        `join_fade` > 0: the chunks are those of ONE text, in order — they are clip-fixed and cross-faded over
        `join_fade` samples on the device and ONE joined int16 wave is returned (a single chunk comes back untouched)."""
        B = len(audios)
        reqs = (_lib.VVRequest * B)()
        keep = []
        outs = []
        seen: Dict[int, np.ndarray] = {}
        shared: Dict[int, int] = {}
        for i in range(B):
            a = seen.get(id(audios[i]))
            if a is None:
                a = np.ascontiguousarray(np.asarray(audios[i]).reshape(-1), dtype=np.int16)
                seen[id(audios[i])] = a
            elif prompt_ids is None and id(audios[i]) not in shared:
                # the same array object again (chunks of one text share their prompt): make it resident once and
                # hand its id to every chunk instead of having each request's audio hashed
                try:
                    shared[id(audios[i])] = self.prompt_put(a)
                except _lib.VVError:
                    shared[id(audios[i])] = 0          # prompt cache disabled: the plain path uploads per chunk
            t = np.ascontiguousarray(np.asarray(text_ids[i]).reshape(-1), dtype=np.int32)
            T = int(total_frames[i])
            ref_len = a.size // self.arch.hop + 1
            n_pcm = max(0, T - ref_len - 1) * self.arch.hop
            o = pcm_out[i] if pcm_out is not None else np.empty(max(n_pcm, 1), dtype=np.int16)
            nz = None
            if noises is not None and noises[i] is not None:
                nz = np.ascontiguousarray(noises[i], dtype=np.float32).reshape(T, self.arch.n_mel)
            keep += [a, t, o, nz]
            reqs[i].audio = a.ctypes.data
            reqs[i].n_samples = a.size
            reqs[i].text_ids = t.ctypes.data
            reqs[i].n_ids = t.size
            reqs[i].total_frames = T
            reqs[i].noise = nz.ctypes.data if nz is not None else None
            reqs[i].chunk_key = int(chunk_keys[i]) if chunk_keys is not None else i
            reqs[i].pcm_out = o.ctypes.data
            reqs[i].pcm_capacity = o.size
            if prompt_ids is not None:
                reqs[i].prompt_id = int(prompt_ids[i] or 0)
            else:
                reqs[i].prompt_id = shared.get(id(audios[i]), 0)
            outs.append(o)
        if join_fade > 0:
            fo, fi = fade_tables(join_fade)
            joined = np.empty(sum(o.size for o in outs), dtype=np.int16)
            got = C.c_int64(0)
            _lib.check(self.lib.vv_synthesize_joined(self._h, reqs, B, nfe, seed, _ptr(fo), _ptr(fi), join_fade,
                                                     _ptr(joined), joined.size, C.byref(got)))
            return joined[: int(got.value)]
        _lib.check(self.lib.vv_synthesize_batch(self._h, reqs, B, nfe, seed))
        return [outs[i][: int(reqs[i].n_out)] for i in range(B)]


class Batch:
    def __init__(self, engine: Engine, total_frames: Sequence[int]):
        self.e = engine
        self.lib = engine.lib
        self.T = [int(t) for t in total_frames]
        self.B = len(self.T)
        self._h = C.c_void_p()
        arr = (C.c_int64 * self.B)(*self.T)
        _lib.check(self.lib.vv_batch_create(engine._h, self.B, arr, C.byref(self._h)))
        self.ref_len = [0] * self.B

    def close(self) -> None:
        if self._h and self.e._h:
            self.lib.vv_batch_destroy(self._h)
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def preprocess(self, idx: int, audio: np.ndarray, text_ids: np.ndarray, noise: Optional[np.ndarray] = None,
                   seed: int = 9527, chunk_key: int = 0) -> int:
        a = np.ascontiguousarray(np.asarray(audio).reshape(-1), dtype=np.int16)
        t = np.ascontiguousarray(np.asarray(text_ids).reshape(-1), dtype=np.int32)
        nz = None
        if noise is not None:
            nz = np.ascontiguousarray(noise, dtype=np.float32).reshape(self.T[idx], self.e.arch.n_mel)
        rl = C.c_int64(0)
        _lib.check(self.lib.vv_preprocess(self._h, idx, _ptr(a), a.size, _ptr(t), t.size, _ptr(nz), seed, chunk_key,
                                          C.byref(rl)))
        self.ref_len[idx] = int(rl.value)
        return self.ref_len[idx]

    def preprocess_prompt(self, idx: int, prompt_id: int, text_ids: np.ndarray, noise: Optional[np.ndarray] = None,
                          seed: int = 9527, chunk_key: int = 0) -> int:
        """`preprocess` for a prompt that is already resident (Engine.prompt_put): nothing but the ids is uploaded."""
        t = np.ascontiguousarray(np.asarray(text_ids).reshape(-1), dtype=np.int32)
        nz = None
        if noise is not None:
            nz = np.ascontiguousarray(noise, dtype=np.float32).reshape(self.T[idx], self.e.arch.n_mel)
        rl = C.c_int64(0)
        _lib.check(self.lib.vv_preprocess_prompt(self._h, idx, int(prompt_id), _ptr(t), t.size, _ptr(nz), seed,
                                                 chunk_key, C.byref(rl)))
        self.ref_len[idx] = int(rl.value)
        return self.ref_len[idx]

    def sample(self, nfe: int = 0, first_step: int = 0, n_steps: Optional[int] = None) -> None:
        n = (nfe or self.e.arch.nfe) - 1 - first_step if n_steps is None else n_steps
        _lib.check(self.lib.vv_sample(self._h, nfe, first_step, n))

    def run_resident(self, nfe: int = 0) -> None:
        """whole path on inputs already resident in HBM; asynchronous on the engine stream"""
        _lib.check(self.lib.vv_run_resident(self._h, nfe))

    def profile_step(self, step: int = 0, nfe: int = 0) -> List[float]:
        ms = (C.c_float * 8)()
        _lib.check(self.lib.vv_profile_step(self._h, nfe, step, ms))
        return [float(x) for x in ms]

    def profile_stages(self, nfe: int = 0) -> Dict[str, float]:
        """resident path with events at the stage boundaries -> milliseconds per stage"""
        ms = (C.c_float * 4)()
        _lib.check(self.lib.vv_profile_stages(self._h, nfe, ms))
        return {"pre_ms": float(ms[0]), "loop_ms": float(ms[1]), "decode_ms": float(ms[2]), "total_ms": float(ms[3])}

    def debug_partial_step(self, step: int, n_layers: int, nfe: int = 0) -> None:
        _lib.check(self.lib.vv_debug_partial_step(self._h, nfe, step, n_layers))

    def crossfade(self, n_fade: int, order: Optional[Sequence[int]] = None) -> np.ndarray:
        """clip-fix + cross-fade of the decoded chunks (all, or `order`) straight from device PCM -> joined int16 wave"""
        idx = list(range(self.B)) if order is None else [int(i) for i in order]
        fo, fi = fade_tables(n_fade)
        cap = sum(int(self.lib.vv_batch_pcm_len(self._h, i)) for i in idx)
        out = np.empty(max(cap, 1), dtype=np.int16)
        arr = (C.c_int32 * len(idx))(*idx)
        got = C.c_int64(0)
        _lib.check(self.lib.vv_batch_crossfade(self._h, arr, len(idx), _ptr(fo), _ptr(fi), n_fade, _ptr(out), out.size,
                                               C.byref(got)))
        return out[: int(got.value)]

    def decode(self, idx: int) -> np.ndarray:
        n = int(self.lib.vv_batch_pcm_len(self._h, idx))
        out = np.empty(max(n, 1), dtype=np.int16)
        got = C.c_int64(0)
        _lib.check(self.lib.vv_decode(self._h, idx, _ptr(out), out.size, C.byref(got)))
        return out[: int(got.value)]

    def get(self, idx: int, name: str) -> np.ndarray:
        a = self.e.arch
        T = self.T[idx]
        shapes = {
            "noise": (T, a.n_mel), "mel": (min(self.ref_len[idx], T), a.n_mel),
            "cat_mel_text": (T, a.cond_dim), "cat_mel_text_drop": (T, a.cond_dim),
            "hidden": (2, T, a.dim), "x0": (2, T, a.dim), "cond_proj": (2, T, a.dim), "v": (2, T, a.n_mel),
            "qkv": (2, T, 3 * a.dim), "attn": (2, T, a.dim), "hb": (2, T, a.dim), "h1b": (2, T, a.dim),
            "ffb": (2, T, a.ff_dim), "voc_head": (max(0, T - self.ref_len[idx]), a.n_fft + 2),
        }
        shp = shapes[name]
        out = np.empty(max(int(np.prod(shp)), 1), dtype=np.float32)
        n = _lib.check(int(self.lib.vv_get_tensor(self._h, idx, name.encode(), _ptr(out), out.size)))
        return out[:n].reshape(shp)

    def set_noise(self, idx: int, noise: np.ndarray) -> None:
        nz = np.ascontiguousarray(noise, dtype=np.float32).reshape(self.T[idx], self.e.arch.n_mel)
        _lib.check(self.lib.vv_set_noise(self._h, idx, _ptr(nz)))

    def set_ref_len(self, idx: int, ref_len: int) -> None:
        _lib.check(self.lib.vv_set_ref_len(self._h, idx, int(ref_len)))
        self.ref_len[idx] = int(ref_len)

    def set_cond(self, idx: int, cat_c: np.ndarray, cat_u: np.ndarray) -> None:
        a = self.e.arch
        c = np.ascontiguousarray(cat_c, dtype=np.float32).reshape(self.T[idx], a.cond_dim)
        u = np.ascontiguousarray(cat_u, dtype=np.float32).reshape(self.T[idx], a.cond_dim)
        _lib.check(self.lib.vv_set_cond(self._h, idx, _ptr(c), _ptr(u)))
