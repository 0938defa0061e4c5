"""Path-level parity on the B200: preprocess -> DiT steps -> decode through the C ABI, against the CPU oracle on
the same seeded inputs (random-init weights of a small architecture of the same topology; injected noise).

Tolerances (bf16 GEMM operands / fp32 residual stream, vs the fp32 oracle):
  * preprocess outputs (fp32 kernels + bf16 text GEMMs)     rel-L2 <= 1e-2
  * mel after ONE step from identical input                 rel-L2 <= 3e-3   
  * mel after all nfe-1 steps                               rel-L2 <= 2e-2   
  * waveform                                                SNR >= 25 dB
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from vietvoice_tts_b200 import artifact
from vietvoice_tts_b200.arch import TINY
from vietvoice_tts_b200.engine import Engine
from oracle.graphs import OracleSessions


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def snr_db(x, ref):
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return 10 * np.log10(np.sum(ref ** 2) / (np.sum((x - ref) ** 2) + 1e-30))


@pytest.fixture(scope="module")
def setup():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    W = artifact.make_random_weights(TINY, 9527)
    eng = Engine.from_weights(TINY, W, device=0)
    ora = OracleSessions(TINY, W)
    yield eng, ora
    eng.close()


def make_inputs(n_samples, n_ids, T, seed):
    rng = np.random.default_rng(seed)
    audio = artifact.synthetic_prompt_pcm(n_samples, seed)
    ids = rng.integers(0, TINY.vocab, size=(1, n_ids)).astype(np.int32)
    noise = rng.standard_normal((1, T, TINY.n_mel)).astype(np.float32)
    return audio, ids, noise


def test_preprocess_matches_oracle(setup):
    eng, ora = setup
    n_samples, T = 30000, 30000 // 256 + 1 + 90
    audio, ids, noise = make_inputs(n_samples, 50, T, 1)
    pre = ora.preprocess.run(audio.reshape(1, 1, -1), ids, np.array([T], dtype=np.int64), noise)
    b = eng.batch([T])
    ref_len = b.preprocess(0, audio, ids, noise)
    assert ref_len == int(pre[7][0]) == n_samples // 256 + 1
    mel = b.get(0, "mel")
    assert np.abs(mel - pre[5][0, :ref_len, :TINY.n_mel]).max() < 2e-3
    cat_c = b.get(0, "cat_mel_text")
    cat_u = b.get(0, "cat_mel_text_drop")
    assert cat_c.shape == pre[5][0].shape
    assert rel(cat_c, pre[5][0]) < 1e-2
    assert rel(cat_u, pre[6][0]) < 1e-2
    assert np.array_equal(b.get(0, "noise"), noise[0])
    assert np.all(cat_u[:, :TINY.n_mel] == 0)
    assert np.all(cat_c[ref_len:, :TINY.n_mel] == 0)
    b.close()


def test_first_layers_match_oracle(setup):
    """stage-by-stage taps of step 0: input embedding, conv_pos_embed, first block"""
    eng, ora = setup
    n_samples, T = 20000, 20000 // 256 + 1 + 70
    audio, ids, noise = make_inputs(n_samples, 30, T, 2)
    pre = ora.preprocess.run(audio.reshape(1, 1, -1), ids, np.array([T], dtype=np.int64), noise)
    ora.transformer.taps = {}
    ora.transformer.run(*pre[:7], np.array([0], dtype=np.int32))
    taps = ora.transformer.taps
    ora.transformer.taps = None
    b = eng.batch([T])
    b.preprocess(0, audio, ids, noise)
    b.set_cond(0, pre[5][0], pre[6][0])          # identical conditioning: isolates the transformer
    b.debug_partial_step(0, 0)
    assert rel(b.get(0, "x0"), taps["x0"].numpy()) < 5e-3
    assert rel(b.get(0, "h1b"), taps["conv1"].numpy()) < 8e-3
    assert rel(b.get(0, "hidden"), taps["x_embed"].numpy()) < 5e-3
    b.debug_partial_step(0, 1)
    qkv = b.get(0, "qkv")
    d = TINY.dim
    assert rel(qkv[..., :d], taps["q0"].numpy()) < 8e-3
    assert rel(qkv[..., d:2 * d], taps["k0"].numpy()) < 8e-3
    assert rel(qkv[..., 2 * d:], taps["v0"].numpy()) < 8e-3
    assert rel(b.get(0, "attn"), taps["attn0"].numpy()) < 1.5e-2
    assert rel(b.get(0, "hidden"), taps["x_l0"].numpy()) < 8e-3
    b.close()


def test_single_step_matches_oracle(setup):
    eng, ora = setup
    n_samples, T = 24000, 24000 // 256 + 1 + 120
    audio, ids, noise = make_inputs(n_samples, 60, T, 3)
    pre = ora.preprocess.run(audio.reshape(1, 1, -1), ids, np.array([T], dtype=np.int64), noise)
    b = eng.batch([T])
    b.preprocess(0, audio, ids, noise)
    b.set_cond(0, pre[5][0], pre[6][0])
    for step in (0, 3, TINY.nfe - 2):
        x_in = np.random.default_rng(step).standard_normal((1, T, TINY.n_mel)).astype(np.float32)
        ref, ts = ora.transformer.run(x_in, *pre[1:7], np.array([step], dtype=np.int32))
        assert int(ts[0]) == step + 1
        b.set_noise(0, x_in[0])
        b.sample(first_step=step, n_steps=1)
        got = b.get(0, "noise")
        # the update itself (dt * v) is what carries the kernel error; compare the increments too
        assert rel(got, ref[0]) < 3e-3
        assert rel(got - x_in[0], ref[0] - x_in[0]) < 2e-2
    b.close()


def test_full_path_batch_matches_oracle(setup):
    """B=3 ragged chunks: graph-captured loop + decode, each chunk vs its own oracle run"""
    eng, ora = setup
    specs = [(24000, 40, 94 + 100), (30000, 10, 118 + 37), (15000, 80, 59 + 160)]
    ins = [make_inputs(n, l, T, 10 + i) for i, (n, l, T) in enumerate(specs)]
    b = eng.batch([s[2] for s in specs])
    for i, (audio, ids, noise) in enumerate(ins):
        b.preprocess(i, audio, ids, noise)
    b.sample()
    for i, ((n, l, T), (audio, ids, noise)) in enumerate(zip(specs, ins)):
        wave, x, steps, pre = ora.synthesize_chunk(audio.reshape(1, 1, -1), ids, np.array([T], dtype=np.int64), noise)
        got = b.get(i, "noise")
        assert rel(got, x[0]) < 2e-2, (i, rel(got, x[0]))
        pcm = b.decode(i)
        assert pcm.dtype == np.int16 and pcm.shape[0] == wave.size == (T - pre[7][0] - 1) * 256
        assert snr_db(pcm, wave.reshape(-1)) > 25.0, (i, snr_db(pcm, wave.reshape(-1)))
    b.close()


def test_decode_from_identical_mel(setup):
    eng, ora = setup
    n_samples, T = 16000, 63 + 140
    audio, ids, noise = make_inputs(n_samples, 20, T, 4)
    mel = np.random.default_rng(7).standard_normal((1, T, TINY.n_mel)).astype(np.float32) * 1.5
    ref_len = n_samples // 256 + 1
    wave = ora.decode.run(mel, np.array([ref_len], dtype=np.int64))[0].reshape(-1)
    ora.decode.taps = {}
    ora.decode.run(mel, np.array([ref_len], dtype=np.int64))
    head_ref = ora.decode.taps["head"].numpy()
    ora.decode.taps = None
    b = eng.batch([T])
    b.preprocess(0, audio, ids, noise)
    b.set_noise(0, mel[0])
    pcm = b.decode(0)
    assert rel(b.get(0, "voc_head"), head_ref) < 1e-2
    assert pcm.shape[0] == wave.shape[0] == (T - ref_len - 1) * 256
    assert snr_db(pcm, wave) > 30.0
    b.close()


def test_synthesize_batch_host_api_and_philox(setup):
    eng, ora = setup
    specs = [(24000, 40, 94 + 100), (24000, 25, 94 + 60)]
    audios = [artifact.synthetic_prompt_pcm(n, 5) for n, _, _ in specs]
    ids = [np.arange(l, dtype=np.int32) % TINY.vocab for _, l, _ in specs]
    Ts = [T for _, _, T in specs]
    out1 = eng.synthesize_batch(audios, ids, Ts, seed=9527)
    out2 = eng.synthesize_batch(audios, ids, Ts, seed=9527)
    out3 = eng.synthesize_batch(audios, ids, Ts, seed=1)
    for a, b_, c, T in zip(out1, out2, out3, Ts):
        assert a.shape[0] == (T - 94 - 1) * 256
        assert np.array_equal(a, b_)                 # same (seed, chunk_key) -> same noise -> same PCM
        assert not np.array_equal(a, c)
    # Philox noise is N(0,1)
    b = eng.batch([4000])
    b.preprocess(0, audios[0], ids[0], None, seed=3, chunk_key=5)
    z = b.get(0, "noise")
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01
    b.close()
