"""A RECORDING stand-in for the `onnxruntime` module: every call the session manager makes is appended to `EVENTS`.

Used twice, with identical inputs:
  * tests/golden/make_session_goldens.py installs it as `sys.modules["onnxruntime"]` and drives the REFERENCE's own
    `ModelSessionManager` (/root/reference/vietvoicetts/core/model.py:18-224) -> tests/golden/session_manager.json;
  * tests/test_session_goldens_cpu.py patches it into the mirror (vietvoice-tts_b200/host/model.py) and compares.
It executes nothing: sessions only remember what they were built from and report the three graphs' positional I/O
names.  It is deliberately NOT the product's ort_shim, so the comparison is independent of it.
"""
import hashlib

EVENTS = []
AVAILABLE = ["CUDAExecutionProvider", "CPUExecutionProvider"]

_IO = {
    0: (["audio", "text_ids", "max_duration"],
        ["noise", "rope_cos_q", "rope_sin_q", "rope_cos_k", "rope_sin_k", "cat_mel_text", "cat_mel_text_drop",
         "ref_signal_len"]),
    1: (["noise", "rope_cos_q", "rope_sin_q", "rope_cos_k", "rope_sin_k", "cat_mel_text", "cat_mel_text_drop",
         "time_step"], ["noise_out", "time_step_out"]),
    2: (["denoised", "ref_signal_len"], ["output_audio"]),
}


def reset(available=None):
    del EVENTS[:]
    AVAILABLE[:] = list(available if available is not None else ["CUDAExecutionProvider", "CPUExecutionProvider"])


def get_available_providers():
    EVENTS.append(["get_available_providers"])
    return list(AVAILABLE)


def set_seed(seed):
    EVENTS.append(["set_seed", int(seed)])


class ExecutionMode:
    ORT_SEQUENTIAL = "ORT_SEQUENTIAL"
    ORT_PARALLEL = "ORT_PARALLEL"


class GraphOptimizationLevel:
    ORT_DISABLE_ALL = "ORT_DISABLE_ALL"
    ORT_ENABLE_BASIC = "ORT_ENABLE_BASIC"
    ORT_ENABLE_EXTENDED = "ORT_ENABLE_EXTENDED"
    ORT_ENABLE_ALL = "ORT_ENABLE_ALL"


class SessionOptions:
    def __init__(self):
        object.__setattr__(self, "attrs", {})
        object.__setattr__(self, "entries", {})

    def __setattr__(self, k, v):
        self.attrs[k] = v

    def add_session_config_entry(self, k, v):
        self.entries[str(k)] = str(v)


class _Arg:
    def __init__(self, name):
        self.name = name


class InferenceSession:
    def __init__(self, model_bytes, sess_options=None, providers=None, **kw):
        data = bytes(model_bytes)
        self.gid = int.from_bytes(data[12:16], "little") if data[:8] == b"VVB200W1" else -1
        EVENTS.append(["InferenceSession", {"nbytes": len(data), "sha256": hashlib.sha256(data).hexdigest(),
                                            "graph_id": self.gid,
                                            "attrs": {k: (v if isinstance(v, (int, float, str, bool)) else str(v))
                                                      for k, v in sorted(sess_options.attrs.items())},
                                            "entries": dict(sorted(sess_options.entries.items())),
                                            "providers": list(providers or []), "extra_kwargs": sorted(kw)}])

    def get_inputs(self):
        return [_Arg(n) for n in _IO[self.gid][0]]

    def get_outputs(self):
        return [_Arg(n) for n in _IO[self.gid][1]]

    def run(self, *a, **k):
        raise RuntimeError("fake_ort sessions do not execute")
