"""Path-level checks on the B200 beyond test_engine_gpu.py: committed golden vectors, the onnxruntime-shaped session
API (what the reference's TTSEngine calls), the host TTSEngine mirror end to end from a model tarball, and
size-independent properties at the FULL architecture / BASELINE sizes (batch invariance, determinism, a single
full-size step against the oracle)."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from vietvoice_tts_b200 import artifact, ort_shim
from vietvoice_tts_b200.arch import FULL, TINY
from vietvoice_tts_b200.engine import Engine

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def snr_db(x, ref):
    x, ref = np.asarray(x, np.float64).reshape(-1), np.asarray(ref, np.float64).reshape(-1)
    return 10 * np.log10(np.sum(ref ** 2) / (np.sum((x - ref) ** 2) + 1e-30))


@pytest.fixture(scope="module")
def need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


@pytest.fixture(scope="module")
def tiny_engine(need_gpu):
    eng = Engine.from_weights(TINY, artifact.make_random_weights(TINY, 9527))
    yield eng
    eng.close()


def test_committed_golden_vectors(tiny_engine):
    """CUDA path vs tests/golden/oracle_tiny.npz (inputs + oracle outputs committed with their generator)."""
    g = np.load(os.path.join(G, "oracle_tiny.npz"))
    T = int(g["T"][0])
    b = tiny_engine.batch([T])
    ref_len = b.preprocess(0, g["audio"], g["ids"], g["noise"])
    assert ref_len == int(g["ref_signal_len"][0])
    assert np.abs(b.get(0, "cat_mel_text")[None] - g["cat_mel_text"].astype(np.float32)).max() < 6e-2
    b.sample(first_step=0, n_steps=1)
    assert rel(b.get(0, "noise")[None], g["step1"]) < 3e-3
    b.sample(first_step=1, n_steps=TINY.nfe - 2)
    assert rel(b.get(0, "noise")[None], g["final"]) < 2e-2
    assert snr_db(b.decode(0), g["wave"]) > 25.0
    b.close()


@pytest.fixture(scope="module")
def model_tar(tmp_path_factory, need_gpu):
    d = tmp_path_factory.mktemp("models")
    artifact.build_model_tar(str(d / "model-bin.pt"), TINY, seed=9527, prompt_seconds=1.5)
    return str(d)


def test_session_api_and_host_engine(model_tar):
    """the reference's call sequence (3 -> 8, 8 -> 2 x (nfe-1), 2 -> 1, positional feeds) through the shim, and the
    batched fast path of the host TTSEngine: both give the same waveform for the same seed"""
    from vietvoice_tts_b200.host.model_config import ModelConfig
    from vietvoice_tts_b200.host.tts_engine import TTSEngine

    cfg = ModelConfig(model_cache_dir=model_tar, nfe_step=TINY.nfe)
    text = "Xin chào Việt Nam. Hôm nay trời đẹp quá, tôi đi học!"
    with TTSEngine(cfg, use_sessions=True) as slow:
        m = slow.model_session_manager
        assert [len(m.input_names[k]) for k in ("preprocess", "transformer", "decode")] == [3, 8, 2]
        assert [len(m.output_names[k]) for k in ("preprocess", "transformer", "decode")] == [8, 2, 1]
        ort_shim.set_seed(cfg.random_seed)
        m.sessions["preprocess"]._sh.calls = 0
        w_slow, secs = slow.synthesize(text)
        assert w_slow.dtype == np.int16 and w_slow.ndim == 1 and w_slow.size > 24000 and secs > 0
        # session-level contract
        ref_audio, ref_text = m.select_sample()
        ins = slow._prepare_inputs(ref_audio, ref_text, text)
        pre = slow._run_preprocess(*ins[0][:3])
        assert len(pre) == 8 and pre[0].shape == (1, int(ins[0][2][0]), 100) and pre[3].shape[1] == 64
        x, ts = slow._run_transformer_steps(*pre[:7], ins[0][3])
        assert int(ts[0]) == TINY.nfe - 1 and x.shape == pre[0].shape
        wav = slow._run_decode(x, pre[7])
        assert wav.dtype == np.int16 and wav.reshape(-1).size == (int(ins[0][2][0]) - int(pre[7][0]) - 1) * 256
    with TTSEngine(cfg) as fast:
        w_fast, _ = fast.synthesize(text)
        assert w_fast.shape == w_slow.shape
        assert snr_db(w_fast, w_slow) > 30.0          # same Philox key (seed, chunk 0): only fp summation order differs
        with pytest.raises(ValueError):               # custom prompt + default voice filters
            fast.synthesize("a", reference_audio=os.path.join(model_tar, "model-bin.pt"), reference_text="x")
        long_text = " ".join(["Đây là một câu khá dài để kiểm tra việc chia đoạn văn bản thành nhiều phần nhỏ hơn."] * 12)
        w_long, _ = fast.synthesize(long_text)
        n_chunks = len(fast._prepare_inputs(*fast.model_session_manager.select_sample(), long_text))
        assert n_chunks > 1 and w_long.dtype == np.int16 and w_long.size > w_fast.size


def test_full_arch_single_step_vs_oracle(need_gpu):
    """FULL architecture (dim 1024, 22 layers, 16 heads): one Euler step from identical input vs the fp32 oracle"""
    from oracle.graphs import OracleSessions
    W = artifact.make_random_weights(FULL, 9527)
    eng = Engine.from_weights(FULL, W)
    ora = OracleSessions(FULL, W)
    n_samples, T = 24000, 94 + 106
    rng = np.random.default_rng(5)
    audio = artifact.synthetic_prompt_pcm(n_samples, 5)
    ids = rng.integers(0, FULL.vocab, size=(1, 60)).astype(np.int32)
    noise = rng.standard_normal((1, T, 100)).astype(np.float32)
    pre = ora.preprocess.run(audio.reshape(1, 1, -1), ids, np.array([T], dtype=np.int64), noise)
    b = eng.batch([T])
    b.preprocess(0, audio, ids, noise)
    assert rel(b.get(0, "cat_mel_text"), pre[5][0]) < 1e-2
    b.set_cond(0, pre[5][0], pre[6][0])
    ref, _ = ora.transformer.run(*pre[:7], np.array([0], dtype=np.int32))
    b.sample(first_step=0, n_steps=1)
    got = b.get(0, "noise")
    assert rel(got, ref[0]) < 3e-3
    assert rel(got - noise[0], ref[0] - noise[0]) < 3e-2
    b.close()
    # ---- BASELINE sizes: batch invariance + determinism at T = 1501 (properties; the oracle is too slow here)
    T = 1501
    audios = [artifact.synthetic_prompt_pcm(144000, 40 + i) for i in range(3)]
    idl = [rng.integers(0, FULL.vocab, size=270).astype(np.int32) for _ in range(3)]
    a1 = eng.synthesize_batch(audios, idl, [T, T, 1200], nfe=8, seed=11, chunk_keys=[0, 1, 2])
    a2 = eng.synthesize_batch(audios, idl, [T, T, 1200], nfe=8, seed=11, chunk_keys=[0, 1, 2])
    solo = eng.synthesize_batch(audios[1:2], idl[1:2], [T], nfe=8, seed=11, chunk_keys=[1])
    assert a1[0].shape[0] == (T - 563 - 1) * 256 and a1[2].shape[0] == (1200 - 563 - 1) * 256
    for x, y in zip(a1, a2):
        assert np.array_equal(x, y)                      # replay of the same batch is bit-identical
    assert snr_db(solo[0], a1[1]) > 25.0                 # chunk 1 alone == chunk 1 inside a ragged batch
    eng.close()
