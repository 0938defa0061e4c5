"""INTEGRATION route A, end to end: THE REFERENCE'S OWN `TTSEngine.synthesize`
(/root/reference/vietvoicetts/core/tts_engine.py:189-257), unmodified, driven through `ort_shim.install()` — model tar
in the reference's layout, the reference's ModelSessionManager / TextProcessor / AudioProcessor, 33 session calls per
chunk, cross-fade — and compared sample for sample with this repo's mirror (`host.TTSEngine(use_sessions=True)`).

There is no GPU in the build container and the product has no CPU path, so the engine BEHIND the shim is swapped for
an oracle-backed stand-in (test infrastructure, defined here): what is under test is everything ABOVE the C ABI — the
shim's eight onnxruntime symbols, the positional feed binding, dtypes and shapes the reference's code relies on, the
audio stand-ins, and that the mirror is call-for-call the reference.  On the GPU the same session path runs on
libvvb200.so (tests/test_path_gpu.py::test_session_api_and_host_engine).  Skipped where /root/reference is absent.
"""
import importlib
import os
import sys

import numpy as np
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "vietvoicetts")),
                                reason="the reference checkout exists only in the build container")


class _OracleBatch:
    def __init__(self, eng, T):
        self.e, self.T, self.ref_len = eng, [int(T[0])], [0]
        self.pre = None
        self.x = None

    def preprocess(self, idx, audio, text_ids, noise=None, seed=9527, chunk_key=0):
        a = self.e.arch
        if noise is None:
            noise = np.random.default_rng([int(seed), int(chunk_key)]).standard_normal((1, self.T[0], a.n_mel)).astype(np.float32)
        self.pre = self.e.ora.preprocess.run(np.asarray(audio).reshape(1, 1, -1), np.asarray(text_ids).reshape(1, -1),
                                             np.array([self.T[0]], dtype=np.int64), noise)
        self.x = self.pre[0].copy()
        self.cat = [self.pre[5], self.pre[6]]
        self.ref_len[0] = int(self.pre[7][0])
        return self.ref_len[0]

    def get(self, idx, name):
        return {"noise": self.x[0], "cat_mel_text": self.cat[0][0], "cat_mel_text_drop": self.cat[1][0]}[name]

    def set_cond(self, idx, c, u):
        self.cat = [np.asarray(c, np.float32)[None], np.asarray(u, np.float32)[None]]

    def set_noise(self, idx, noise):
        self.x = np.asarray(noise, np.float32)[None].copy()

    def set_ref_len(self, idx, n):
        self.ref_len[0] = int(n)

    def sample(self, nfe=0, first_step=0, n_steps=None):
        ts = np.array([first_step], dtype=np.int32)
        for _ in range(n_steps):
            self.x, ts = self.e.ora.transformer.run(self.x, *self.pre[1:5], self.cat[0], self.cat[1], ts)

    def decode(self, idx):
        return self.e.ora.decode.run(self.x, np.array([self.ref_len[0]], dtype=np.int64))[0].reshape(-1)

    def close(self):
        pass


class _OracleEngine:
    """Engine-shaped object over the CPU oracle (stands where vietvoice_tts_b200.engine.Engine binds libvvb200.so)."""

    def __init__(self, arch, device=0, stream=None):
        self.arch, self._W, self._finalized, self.ora = arch, {}, False, None

    def load_blob(self, blob):
        from vietvoice_tts_b200 import artifact
        self._W.update(artifact.unpack_blob(blob)[1])

    def finalize(self):
        from oracle.graphs import OracleSessions
        self.ora = OracleSessions(self.arch, self._W, nfe=self.arch.nfe)
        self._finalized = True

    def batch(self, T):
        return _OracleBatch(self, T)

    def close(self):
        pass


@pytest.fixture()
def model_dir(tmp_path):
    from vietvoice_tts_b200 import artifact
    from vietvoice_tts_b200.arch import TINY
    voices = [{"gender": "female", "group": "audiobook", "area": "northern", "emotion": "neutral"},
              {"gender": "male", "group": "news", "area": "southern", "emotion": "serious"}]
    artifact.build_model_tar(str(tmp_path / "model-bin.pt"), TINY, seed=3, voices=voices, prompt_seconds=1.5)
    return tmp_path


def test_reference_synthesize_runs_unmodified_on_the_shim_and_equals_the_mirror(model_dir, monkeypatch):
    from vietvoice_tts_b200 import ort_shim
    monkeypatch.setattr(ort_shim, "Engine", _OracleEngine)
    monkeypatch.setattr(ort_shim, "_engines", {})
    saved = {k: sys.modules.get(k) for k in ("onnxruntime", "pydub", "pydub.exceptions", "soundfile")}
    ort_shim.install()                                    # BEFORE the reference is imported (INTEGRATION.md route A)
    sys.path.insert(0, REF)
    try:
        for k in [k for k in sys.modules if k.startswith("vietvoicetts")]:
            del sys.modules[k]
        ref_engine_mod = importlib.import_module("vietvoicetts.core.tts_engine")
        ref_cfg_mod = importlib.import_module("vietvoicetts.core.model_config")
        assert ref_engine_mod.__file__.startswith(REF)
        # TINY: nfe 8 (7 transformer calls per chunk); max_chunk_duration shortened so that the text splits into chunks
        kw = dict(model_cache_dir=str(model_dir), nfe_step=8, max_chunk_duration=4.0, speed=1.0)
        text = "Xin chào Việt Nam. Hôm nay trời đẹp quá! Chúng ta cùng đi dạo nhé? Một hai ba bốn năm sáu bảy."
        ref = ref_engine_mod.TTSEngine(ref_cfg_mod.ModelConfig(**kw))
        n_chunks = len(ref._prepare_inputs(*ref.model_session_manager.select_sample(), text))
        wave_ref, secs = ref.synthesize(text, output_path=str(model_dir / "ref.wav"))
        wave_ref_m, _ = ref.synthesize("Tôi đi học.", gender="male", group="news", area="southern", emotion="serious")
        with pytest.raises(ValueError, match="Cannot use reference audio and text with options"):
            ref.synthesize("x" * 10, reference_audio=str(model_dir / "ref.wav"), reference_text="abc")   # options clash
        ref.cleanup()

        ort_shim._engines.clear()
        ort_shim.set_seed(9527)
        from vietvoice_tts_b200.host.model_config import ModelConfig
        from vietvoice_tts_b200.host.tts_engine import TTSEngine
        mir = TTSEngine(ModelConfig(**kw), use_sessions=True)
        wave_mir, _ = mir.synthesize(text, output_path=str(model_dir / "mir.wav"))
        wave_mir_m, _ = mir.synthesize("Tôi đi học.", gender="male", group="news", area="southern", emotion="serious")
        mir.cleanup()
    finally:
        sys.path.remove(REF)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        for k in [k for k in sys.modules if k.startswith("vietvoicetts")]:
            del sys.modules[k]
    assert n_chunks >= 3
    assert wave_ref.dtype == np.int16 and wave_ref.ndim == 1 and wave_ref.size > 24000 and secs > 0
    assert np.abs(wave_ref.astype(np.int32)).max() > 0
    assert np.array_equal(wave_ref, wave_mir)             # multi-chunk: chunker, 33-call loop, cross-fade
    assert np.array_equal(wave_ref_m, wave_mir_m)         # another voice, single chunk (returned untouched)
    ref_bytes, mir_bytes = (model_dir / "ref.wav").read_bytes(), (model_dir / "mir.wav").read_bytes()
    assert ref_bytes == mir_bytes and ref_bytes[:4] == b"RIFF"      # the reference's sf.write -> the WAVEX stand-in
