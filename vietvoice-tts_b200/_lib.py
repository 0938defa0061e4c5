"""ctypes binding of libvvb200.so (C ABI: include/vvb200.h).

Fails loudly: if the shared library is missing, `load()` raises ImportError telling how to build it;
there is no Python/NumPy/PyTorch fallback for any compute entry point.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess
import threading

from .arch import VVArch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VVB200_LIB") or os.path.join(_HERE, "libvvb200.so")   # override: A/B kernel experiments
CSRC = os.path.join(_HERE, "csrc")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "vvb200.h")

_lock = threading.Lock()
_lib = None


class VVRequest(C.Structure):
    _fields_ = [("audio", C.c_void_p), ("n_samples", C.c_int64), ("text_ids", C.c_void_p), ("n_ids", C.c_int64),
                ("total_frames", C.c_int64), ("noise", C.c_void_p), ("chunk_key", C.c_uint64),
                ("pcm_out", C.c_void_p), ("pcm_capacity", C.c_int64), ("n_out", C.c_int64),
                ("prompt_id", C.c_uint64)]


class VVGemmEpilogue(C.Structure):
    _fields_ = [("bias", C.c_void_p), ("gate", C.c_void_p), ("resid", C.c_void_p), ("ld_resid", C.c_int32),
                ("out_f32", C.c_void_p), ("ld_f32", C.c_int32), ("out_bf16", C.c_void_p), ("ld_bf16", C.c_int32),
                ("row_mask", C.c_void_p), ("row_pos", C.c_void_p), ("rope_dim", C.c_int32),
                ("rope_off2", C.c_int32), ("act", C.c_int32)]


def declared_symbols() -> list[str]:
    """Every function the public header declares (used by the CPU test that checks the exports)."""
    with open(HEADER, "r", encoding="utf-8") as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vv_[a-z0-9_]+)\s*\(", text)))


def build(verbose: bool = False) -> str:
    """Compile csrc/ into libvvb200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libvvb200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(r.stdout[-2000:])
    return LIB_PATH


def load() -> C.CDLL:
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found. This engine has no CPU/PyTorch fallback: build the CUDA library first "
                f"(`make -C {CSRC}` or `python -c 'import __graft_entry__ as g; g.build()'`).")
        lib = C.CDLL(LIB_PATH)
        P, I, I64, U64, F = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float
        sig = {
            "vv_last_error": (C.c_char_p, []),
            "vv_version": (I, []),
            "vv_device_count": (I, []),
            "vv_engine_create": (I, [C.POINTER(VVArch), I, P, C.POINTER(P)]),
            "vv_engine_load_blob": (I, [P, P, C.c_size_t]),
            "vv_engine_finalize": (I, [P]),
            "vv_engine_destroy": (None, [P]),
            "vv_engine_launch_count": (I64, [P]),
            "vv_engine_stream": (P, [P]),
            "vv_batch_create": (I, [P, I, C.POINTER(I64), C.POINTER(P)]),
            "vv_batch_destroy": (None, [P]),
            "vv_preprocess": (I, [P, I, P, I64, P, I64, P, U64, U64, C.POINTER(I64)]),
            "vv_prompt_put": (I, [P, P, I64, C.POINTER(U64), C.POINTER(I64)]),
            "vv_prompt_drop": (I, [P, U64]),
            "vv_prompt_cache_clear": (I, [P]),
            "vv_prompt_cache_stats": (I, [P, C.POINTER(I64)]),
            "vv_preprocess_prompt": (I, [P, I, U64, P, I64, P, U64, U64, C.POINTER(I64)]),
            "vv_sample": (I, [P, I, I, I]),
            "vv_decode": (I, [P, I, P, I64, C.POINTER(I64)]),
            "vv_decode_all": (I, [P, C.POINTER(P), C.POINTER(I64)]),
            "vv_batch_pcm_len": (I64, [P, I]),
            "vv_get_tensor": (I64, [P, I, C.c_char_p, P, I64]),
            "vv_set_noise": (I, [P, I, P]),
            "vv_set_cond": (I, [P, I, P, P]),
            "vv_set_ref_len": (I, [P, I, I64]),
            "vv_sync": (I, [P]),
            "vv_debug_partial_step": (I, [P, I, I, I]),
            "vv_synthesize_batch": (I, [P, C.POINTER(VVRequest), I, I, U64]),
            "vv_crossfade_pcm": (I, [P, C.POINTER(P), C.POINTER(I64), I, P, P, I, P, I64, C.POINTER(I64)]),
            "vv_batch_crossfade": (I, [P, C.POINTER(C.c_int32), I, P, P, I, P, I64, C.POINTER(I64)]),
            "vv_synthesize_joined": (I, [P, C.POINTER(VVRequest), I, I, U64, P, P, I, P, I64, C.POINTER(I64)]),
            "vv_run_resident": (I, [P, I]),
            "vv_profile_step": (I, [P, I, I, C.POINTER(C.c_float)]),
            "vv_profile_stages": (I, [P, I, C.POINTER(C.c_float)]),
            "vv_gemm_bf16": (I, [P, P, I, P, I, I, I, I, C.POINTER(VVGemmEpilogue), I]),
            "vv_conv_rows_bf16": (I, [P, P, I, P, I, I, I, C.POINTER(VVGemmEpilogue)]),
            "vv_attention_bf16": (I, [P, P, P, I, C.POINTER(C.c_int32), C.POINTER(C.c_int32), I, I]),
            "vv_ln_modulate": (I, [P, P, I, I, P, P, F, P]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


class VVError(RuntimeError):
    pass


def check(rc: int) -> int:
    if rc < 0:
        msg = load().vv_last_error()
        raise VVError(f"vvb200 error {rc}: {msg.decode('utf-8', 'replace') if msg else ''}")
    return rc
