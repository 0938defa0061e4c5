"""Host-side mirror of the reference interface vs golden vectors produced by the reference's own modules
(tests/golden/make_host_goldens.py).  Text normalisation, chunking and token ids must match bit-exactly
(north_star); every int16 cast in the audio helpers truncates exactly as the reference does.  Also the reference's
own known-answer tests for these helpers."""
import json
import os
import wave

import numpy as np
import pytest

from vietvoice_tts_b200.host.text_processor import TextProcessor
from vietvoice_tts_b200.host.audio_processor import AudioProcessor

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def text_gold():
    with open(os.path.join(G, "host_text.json"), encoding="utf-8") as f:
        return json.load(f)


@pytest.fixture(scope="module")
def tp(tmp_path_factory, text_gold):
    p = tmp_path_factory.mktemp("v") / "vocab.txt"
    p.write_text("\n".join(text_gold["vocab"]) + "\n", encoding="utf-8")
    return TextProcessor(str(p))


def test_text_goldens_bit_exact(tp, text_gold):
    assert len(text_gold["cases"]) >= 50
    for c in text_gold["cases"]:
        cleaned = tp.clean_text(c["raw"])
        assert cleaned == c["clean"], c["raw"]
        assert tp.calculate_text_length(cleaned, r".,?!:") == c["len_default"]
        assert tp.calculate_text_length(cleaned, r"[,.]") == c["len_class"]
        ids = tp.text_to_indices([list(cleaned)])
        assert ids.dtype == np.int32 and ids.tolist() == c["ids"]
        for mc, chunks in c["chunks"].items():
            assert tp.chunk_text(cleaned, max_chars=int(mc)) == chunks, (c["raw"], mc)
        assert tp.chunk_text(c["raw"], max_chars=30) == c["chunks_raw_30"]


def test_reference_known_answers_text(tmp_path):
    # /root/reference/tests/test_text_processor_full.py:14-35
    p = tmp_path / "vocab.txt"
    p.write_text("a\nb\nc\n")
    t = TextProcessor(str(p))
    assert t.vocab_char_map == {"a": 0, "b": 1, "c": 2} and t.vocab_size == 3
    assert t.text_to_indices([["a", "b", "c"]]).tolist() == [[0, 1, 2]]
    assert t.text_to_indices([["a", "z"]]).tolist() == [[0, 0]]            # unknown -> 0
    assert t.calculate_text_length("a, b, c.", r"[,.]") == len("a, b, c.".encode()) + 3 * 3 == 17
    assert t.calculate_text_length("a, b, c.", r".,?!:") == 8              # default pattern is a regex (Appendix B)
    assert t.clean_text("  a;b:c(d)   efg! ") == "a,b,c,d, efg!"
    with pytest.raises(FileNotFoundError):
        TextProcessor(str(tmp_path / "missing.txt"))


def test_chunk_invariants(tmp_path):
    # /root/reference/tests/test_text_processor.py:24-136
    p = tmp_path / "vocab.txt"
    p.write_text("a\n")
    t = TextProcessor(str(p))
    assert t.chunk_text("") == [] and t.chunk_text("    ") == []
    text = "This is a long sentence. This is another long sentence. And a third one."
    chunks = t.chunk_text(text, max_chars=30)
    assert chunks == ["This is a long sentence.", "This is another long sentence.", "And a third one."]
    assert all(len(c) <= 30 for c in chunks)
    long_word = "x" * 200
    assert t.chunk_text(long_word, max_chars=50) == [long_word]           # a single long word is returned whole
    words = " ".join(["word"] * 100)
    for c in t.chunk_text(words, max_chars=40):
        assert len(c) <= 40 and all(w == "word" for w in c.split())     # never splits inside a word


def test_audio_goldens_bit_exact():
    g = np.load(os.path.join(G, "host_audio.npz"))
    out = AudioProcessor.normalize_to_int16(g["norm_in"])
    assert out.dtype == np.int16 and np.array_equal(out, g["norm_out"]) and np.abs(out).max() <= 29491
    assert np.array_equal(AudioProcessor.normalize_to_int16(np.zeros(16, dtype=np.float32)), g["norm_zero_out"])
    fixed = AudioProcessor.fix_clipped_audio(g["clip_in"])
    assert fixed.dtype == g["clip_out"].dtype and np.array_equal(fixed, g["clip_out"]) and np.abs(fixed).max() < 32767
    same = AudioProcessor.fix_clipped_audio(g["noclip_in"])
    assert same.dtype == g["noclip_out"].dtype and np.array_equal(same, g["noclip_out"])
    waves = [g[f"wave{i}"] for i in range(5)]
    for name, args in (("xf_improved", (waves, 0.1)), ("xf_improved_two", (waves[:2], 0.1)),
                       ("xf_improved_nofade", (waves[:3], 0.0)), ("xf_improved_single", (waves[:1], 0.1))):
        r = AudioProcessor.concatenate_with_crossfade_improved(args[0], args[1], 24000)
        assert r.dtype == g[name].dtype and np.array_equal(r, g[name]), name
    r = AudioProcessor.concatenate_with_crossfade(waves, 0.1, 24000)
    assert r.dtype == g["xf_plain"].dtype and np.array_equal(r, g["xf_plain"])
    r = AudioProcessor.concatenate_with_crossfade([waves[4], waves[2]], 0.1, 24000)
    assert np.array_equal(r, g["xf_plain_short"])
    # /root/reference/tests/test_audio_processor_full.py:51-61 : length = sum(len) - overlap*(k-1)
    total = waves[0].size
    for w in waves[1:]:
        total += w.size - min(2400, total, w.size)
    assert g["xf_improved"].shape[0] == total == 13200
    assert AudioProcessor.concatenate_with_crossfade_improved([], 0.1, 24000).size == 0


def test_wav_roundtrip_and_load(tmp_path):
    sr = 24000
    t = np.arange(sr) / sr
    pcm = (np.sin(2 * np.pi * 220 * t) * 12000 + 500).astype(np.int16)
    p = str(tmp_path / "a" / "x.wav")
    AudioProcessor.save_audio(pcm, p, sr)                                   # WAVEX
    with wave.open(p, "rb") as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (1, 2, sr, sr)
    loaded = AudioProcessor.load_audio(p, sr)
    assert loaded.dtype == np.int16 and loaded.shape == pcm.shape
    assert np.array_equal(loaded, AudioProcessor.normalize_to_int16(pcm.astype(np.float32)))
    with open(p, "rb") as f:
        assert np.array_equal(AudioProcessor.load_audio(f.read(), sr), loaded)
    assert AudioProcessor.load_audio(p, 12000).shape[0] == sr // 2          # resampled
    with pytest.raises(FileNotFoundError):
        AudioProcessor.load_audio(str(tmp_path / "none.wav"), sr)
    with pytest.raises(ValueError):
        AudioProcessor.save_audio(np.array([], dtype=np.int16), p, sr)
    with pytest.raises(RuntimeError):
        AudioProcessor.load_audio(b"\x00\x00\x00\x20ftypM4A not a wav", sr)


def test_wav_bytes_in_memory_equals_file(tmp_path):
    """the in-memory WAV is byte-identical to the saved file and parses as 24 kHz mono int16"""
    import io
    import wave as wave_mod
    from vietvoice_tts_b200.host.audio_processor import AudioProcessor
    pcm = (np.random.default_rng(3).standard_normal(5000) * 8000).astype(np.int16)
    path = tmp_path / "a.wav"
    AudioProcessor.save_audio(pcm, str(path), 24000)
    blob = AudioProcessor.to_wav_bytes(pcm.reshape(1, 1, -1), 24000)
    assert blob == path.read_bytes()
    back = AudioProcessor.load_audio(blob, 24000)
    assert back.dtype == np.int16 and back.size == pcm.size
    with pytest.raises(ValueError):
        AudioProcessor.to_wav_bytes(np.zeros(0, np.int16), 24000)


def test_prepare_inputs_matches_reference(tp):
    """TTSEngine._prepare_inputs: duration model, chunking and ids vs the reference's own method
    (/root/reference/vietvoicetts/core/tts_engine.py:43-131), goldens from make_host_goldens.py."""
    from vietvoice_tts_b200.host.tts_engine import TTSEngine
    from vietvoice_tts_b200.host.model_config import ModelConfig

    class FakeAudio:
        def __init__(self, n):
            self.n = n

        def load_audio(self, path, sr):
            return (np.arange(self.n) % 1000).astype(np.int16)

    with open(os.path.join(G, "host_prepare_inputs.json"), encoding="utf-8") as f:
        cases = json.load(f)
    assert sum(len(c["chunks"]) for c in cases) > 100
    for c in cases:
        cfg = ModelConfig.__new__(ModelConfig)
        for k, v in dict(sample_rate=24000, hop_length=256, speed=c["speed"], pause_punctuation=r".,?!:",
                         max_chunk_duration=c["max_chunk_duration"], min_target_duration=1.0).items():
            setattr(cfg, k, v)
        eng = TTSEngine.__new__(TTSEngine)
        eng.config, eng.text_processor, eng.audio_processor, eng.sample_cache = cfg, tp, FakeAudio(c["n_samples"]), {}
        got = eng._prepare_inputs("unused.wav", c["ref_text"], c["target"])
        assert len(got) == len(c["chunks"])
        for (audio, ids, max_dur, ts), ref in zip(got, c["chunks"]):
            assert audio.shape == (1, 1, c["n_samples"]) and audio.dtype == np.int16
            assert ids.dtype == np.int32 and ids.tolist() == ref["ids"]
            assert max_dur.dtype == np.int64 and int(max_dur[0]) == ref["max_duration"]
            assert ts.dtype == np.int32 and int(ts[0]) == ref["time_step"] == 0


def test_model_config_contract(tmp_path):
    from vietvoice_tts_b200.host.model_config import ModelConfig, TTSConfig
    (tmp_path / "model-bin.pt").write_bytes(b"x")
    c = ModelConfig(model_cache_dir=str(tmp_path))
    assert (c.nfe_step, c.fuse_nfe, c.sample_rate, c.hop_length, c.random_seed, c.speed) == (32, 1, 24000, 256, 9527, 0.9)
    assert c.max_chunk_duration == 20.0 and c.pause_punctuation == r".,?!:" and TTSConfig is ModelConfig
    assert ModelConfig.from_dict(c.to_dict()).to_dict() == c.to_dict()
    # /root/reference/tests/test_edge_cases.py:330-357
    for bad in (dict(speed=0.05), dict(speed=5.5), dict(nfe_step=0), dict(nfe_step=101)):
        with pytest.raises(ValueError):
            ModelConfig(model_cache_dir=str(tmp_path), **bad)
    with pytest.raises(RuntimeError):          # no network, no cached file -> "Model validation failed"
        ModelConfig(model_cache_dir=str(tmp_path / "empty"), model_url="http://127.0.0.1:9/none")


def test_select_sample_rules_and_messages(tmp_path):
    """Voice selection of the session manager (/root/reference/vietvoicetts/core/model.py:137-214): config defaults
    are merged before filtering, enum validation and its messages, the custom-prompt early-out, the silent fallback
    to voice 0, `sample_iteration`, and the prompt bytes read from `cleaned_audios/` of the tar (cached here)."""
    import tarfile
    from vietvoice_tts_b200 import artifact
    from vietvoice_tts_b200.arch import TINY
    from vietvoice_tts_b200.host.model import ModelSessionManager
    from vietvoice_tts_b200.host.model_config import ModelConfig, MODEL_GENDER

    voices = [{"gender": "female", "group": "audiobook", "area": "northern", "emotion": "neutral"},
              {"gender": "male", "group": "news", "area": "southern", "emotion": "serious"},
              {"gender": "male", "group": "news", "area": "southern", "emotion": "serious"}]
    artifact.build_model_tar(str(tmp_path / "model-bin.pt"), TINY, seed=1, voices=voices, prompt_seconds=0.2)
    cfg = ModelConfig(model_cache_dir=str(tmp_path), gender=None, group=None, area=None, emotion=None)
    m = ModelSessionManager(cfg)
    assert m.providers[-1] == "CPUExecutionProvider"
    with tarfile.open(cfg.model_path) as tar:
        m.sample_metadata = json.load(tar.extractfile("audio_metadata.json"))
        wav = [tar.extractfile("cleaned_audios/" + s["file_name"]).read() for s in m.sample_metadata]

    audio, text = m.select_sample()                                   # no filter at all: first voice
    assert audio == wav[0] and text == m.sample_metadata[0]["text"]
    audio, _ = m.select_sample(gender="male")
    assert audio == wav[1]
    audio, _ = m.select_sample(gender="male", sample_iteration=1)
    assert audio == wav[2]
    with pytest.raises(ValueError, match="sample_iteration 2 is out of range. Only 2 samples available"):
        m.select_sample(gender="male", sample_iteration=2)
    audio, _ = m.select_sample(emotion="happy")                       # nothing matches: silent fallback to voice 0
    assert audio == wav[0]
    with pytest.raises(ValueError, match=r"Invalid gender: robot. Must be one of \['male', 'female'\]"):
        m.select_sample(gender="robot")
    with pytest.raises(ValueError, match="Invalid area"):
        m.select_sample(area="western")
    assert MODEL_GENDER == ["male", "female"]

    # custom prompt: text required, file must exist, no voice options may be active
    prompt = tmp_path / "p.wav"
    prompt.write_bytes(wav[0])
    with pytest.raises(ValueError, match="Reference text is required"):
        m.select_sample(reference_audio=str(prompt))
    with pytest.raises(FileNotFoundError, match="Reference audio file not found"):
        m.select_sample(reference_audio=str(tmp_path / "nope.wav"), reference_text="x")
    with pytest.raises(ValueError, match=r"Cannot use reference audio and text with options: \['gender'\]"):
        m.select_sample(gender="male", reference_audio=str(prompt), reference_text="x")
    assert m.select_sample(reference_audio=str(prompt), reference_text="xin chào") == (str(prompt), "xin chào")

    # config defaults count as options: a custom prompt with the stock config raises
    m2 = ModelSessionManager(ModelConfig(model_cache_dir=str(tmp_path)))
    m2.sample_metadata = m.sample_metadata
    with pytest.raises(ValueError, match="Cannot use reference audio and text with options"):
        m2.select_sample(reference_audio=str(prompt), reference_text="x")
    audio, _ = m2.select_sample()                                     # female / audiobook / northern / neutral
    assert audio == wav[0]

    # a metadata row without the filtered key -> the reference's KeyError wrapper
    m.sample_metadata = [{"file_name": "voice_000.wav", "text": "t"}]
    with pytest.raises(ValueError, match="Sample not found for gender: male"):
        m.select_sample(gender="male")
