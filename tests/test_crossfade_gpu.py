"""Clip-fix + cross-fade on the GPU, BIT-EXACT against
  * tests/golden/host_audio.npz `xfr_*` — produced by RUNNING the reference's own
    AudioProcessor.concatenate_with_crossfade_improved (tests/golden/make_host_goldens.py;
    /root/reference/vietvoicetts/core/audio_processor.py:47-58,123-193), and
  * the host mirror (itself pinned to the reference by the same goldens) on seeded random chunk lists that hit every
    branch: clipped chunks (+32767 / -32768), quiet chunks (RMS < 100: no level matching), level ratios clipped to
    0.7 and 1.5 (with int16 wrap-around where 1.5 x overflows), odd lengths, several fade lengths.
Integer work: the bar is equality of every sample.
"""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from vietvoice_tts_b200 import _lib, artifact
from vietvoice_tts_b200.arch import TINY
from vietvoice_tts_b200.engine import Engine
from vietvoice_tts_b200.host.audio_processor import AudioProcessor

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def eng():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    e = Engine.from_weights(TINY, artifact.make_random_weights(TINY, 9527))
    yield e
    e.close()


def test_reference_generated_goldens(eng):
    g = np.load(os.path.join(G, "host_audio.npz"))
    waves = [g[f"xfr_wave{i}"] for i in range(7)]
    got = eng.crossfade_pcm(waves, int(0.1 * 24000))
    assert got.dtype == np.int16 and np.array_equal(got, g["xfr_out"])
    got2 = eng.crossfade_pcm(waves[3:5], int(0.05 * 24000))
    assert np.array_equal(got2, g["xfr_out_two"])


def _random_chunks(rng, n, n_fade):
    out = []
    for _ in range(n):
        length = int(rng.integers(2 * n_fade, 2 * n_fade + 20000))
        kind = rng.integers(0, 5)
        amp = [3000.0, 40.0, 20000.0, 600.0, 9000.0][kind]
        w = np.clip(rng.standard_normal(length) * amp, -32768, 32767).astype(np.int16)
        if kind == 2:
            w[int(rng.integers(0, length))] = 32767
        if kind == 4 and rng.random() < 0.5:
            w[int(rng.integers(0, length))] = -32768           # np.abs(int16) keeps it negative: no clip fix by itself
        out.append(w.reshape(1, 1, -1))
    return out


@pytest.mark.parametrize("seed", range(6))
def test_matches_host_mirror_sample_for_sample(eng, seed):
    rng = np.random.default_rng(seed)
    for n_fade in (2400, 1, 777, 4096):
        waves = _random_chunks(rng, int(rng.integers(2, 9)), n_fade)
        want = AudioProcessor.concatenate_with_crossfade_improved(waves, n_fade / 24000.0 + 1e-9, 24000)
        assert int((n_fade / 24000.0 + 1e-9) * 24000) == n_fade
        got = eng.crossfade_pcm(waves, n_fade)
        assert got.shape == want.shape and np.array_equal(got, want), (seed, n_fade, int(np.argmax(got != want)))


def test_irregular_cases_are_refused_not_approximated(eng):
    short = [np.zeros(3000, np.int16), np.zeros(9000, np.int16)]           # first chunk shorter than two fades
    assert not eng.crossfade_ok([3000, 9000], 2400) and eng.crossfade_ok([4800, 9000], 2400)
    with pytest.raises(_lib.VVError, match="fewer than two cross-fades"):
        eng.crossfade_pcm(short, 2400)
    with pytest.raises(_lib.VVError):
        eng.crossfade_pcm(short[1:], 2400)                                   # a single chunk is not a fold


def test_joined_synthesis_equals_chunks_plus_host_fold(eng):
    """vv_synthesize_joined == vv_synthesize_batch + the host's cross-fade, and Batch.crossfade on device PCM."""
    rng = np.random.default_rng(3)
    prompt = artifact.synthetic_prompt_pcm(12000, 5)
    ref_frames = 12000 // TINY.hop + 1
    frames = [ref_frames + 60, ref_frames + 45, ref_frames + 80]
    ids = [rng.integers(0, TINY.vocab, size=40).astype(np.int32) for _ in frames]
    audios = [prompt] * 3
    n_fade = 2400
    chunks = eng.synthesize_batch(audios, ids, frames, nfe=6, seed=4, chunk_keys=[0, 1, 2])
    assert all(c.size >= 2 * n_fade for c in chunks)
    want = AudioProcessor.concatenate_with_crossfade_improved([c.reshape(1, 1, -1) for c in chunks], 0.1, 24000)
    joined = eng.synthesize_batch(audios, ids, frames, nfe=6, seed=4, chunk_keys=[0, 1, 2], join_fade=n_fade)
    assert np.array_equal(joined, want)
    one = eng.synthesize_batch(audios[:1], ids[:1], frames[:1], nfe=6, seed=4, chunk_keys=[0], join_fade=n_fade)
    assert np.array_equal(one, chunks[0])                                    # single chunk: returned untouched
    b = eng.batch(frames)
    for i in range(3):
        b.preprocess(i, prompt, ids[i], None, seed=4, chunk_key=i)
    b.sample(nfe=6)
    assert np.array_equal(b.crossfade(n_fade), want)
    assert np.array_equal(b.crossfade(n_fade, order=[2, 0]), AudioProcessor.concatenate_with_crossfade_improved(
        [chunks[2].reshape(1, 1, -1), chunks[0].reshape(1, 1, -1)], 0.1, 24000))
    b.close()
