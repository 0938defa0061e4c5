"""onnxruntime-shaped shim: the eight symbols the reference touches exist with the right shape, and a session
cannot be created from anything but a VVB200 blob / without a GPU (no silent fallback)."""
import numpy as np
import pytest

from vietvoice_tts_b200 import artifact, ort_shim
from vietvoice_tts_b200.arch import TINY
from vietvoice_tts_b200._lib import VVError, load


def test_surface():
    assert "CPUExecutionProvider" in ort_shim.get_available_providers()
    o = ort_shim.SessionOptions()
    for name in ("log_severity_level", "log_verbosity_level", "inter_op_num_threads", "intra_op_num_threads",
                 "enable_cpu_mem_arena", "execution_mode", "graph_optimization_level"):
        setattr(o, name, getattr(o, name))
    o.add_session_config_entry("session.set_denormal_as_zero", "1")
    assert ort_shim.ExecutionMode.ORT_SEQUENTIAL == 0 and ort_shim.GraphOptimizationLevel.ORT_ENABLE_ALL == 99
    ort_shim.set_seed(9527)


def test_rejects_onnx_bytes():
    with pytest.raises(RuntimeError):
        ort_shim.InferenceSession(b"\x08\x07\x12\x07pytorch", providers=["CPUExecutionProvider"])


def test_no_gpu_no_session():
    if load().vv_device_count() > 0:
        pytest.skip("a GPU is present")
    blob = artifact.pack_blob(TINY, artifact.make_random_weights(TINY), "preprocess")
    with pytest.raises(VVError):
        ort_shim.InferenceSession(blob, providers=["CPUExecutionProvider"])


def test_rope_tables_match_oracle():
    from oracle.graphs import rope_tables
    cq, sq, ck, sk = ort_shim._rope_tables(TINY, 77)
    cos, sin = rope_tables(TINY, 77)
    assert np.array_equal(cq[0], cos.numpy()) and np.array_equal(sq[0], sin.numpy())
    assert np.array_equal(ck[0], cos.numpy().T) and ck.shape == (1, 64, 77)
