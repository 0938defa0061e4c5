"""Session manager vs goldens produced by RUNNING THE REFERENCE'S OWN CLASS (tests/golden/make_session_goldens.py):
`ModelSessionManager._get_optimal_providers / _create_session_options / _load_models_from_file / load_models / cleanup`
(/root/reference/vietvoicetts/core/model.py:31-135,216-221) and `select_sample` (:137-214, 1 366 cases).

Both sides talk to the same recording stand-in for onnxruntime (tests/golden/fake_ort.py, which executes nothing), so
what is compared is the host logic: which bytes reach which session in which order with which options, the positional
I/O name lists, the vocab extraction, error wrapping, and the voice-selection outcome of every filter combination.
"""
import hashlib
import json
import os
import sys
import tarfile

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

import fake_ort  # noqa: E402
from vietvoice_tts_b200 import artifact  # noqa: E402
from vietvoice_tts_b200.arch import TINY  # noqa: E402
from vietvoice_tts_b200.host import model as host_model  # noqa: E402
from vietvoice_tts_b200.host.model_config import ModelConfig  # noqa: E402


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(HERE, "golden", "session_manager.json"), encoding="utf-8") as f:
        return json.load(f)


@pytest.fixture(scope="module")
def model_dir(tmp_path_factory, gold):
    d = tmp_path_factory.mktemp("sessgold")
    artifact.build_model_tar(str(d / "model-bin.pt"), TINY, seed=gold["tar_seed"], voices=gold["voices"],
                             prompt_seconds=0.2)
    return d


@pytest.fixture()
def fake(monkeypatch):
    monkeypatch.setattr(host_model, "onnxruntime", fake_ort)
    fake_ort.reset()
    return fake_ort


def test_provider_choice(gold, model_dir, fake):
    for key, avail in (("cuda+cpu", None), ("cpu", ["CPUExecutionProvider"]),
                       ("trt+cuda", ["TensorrtExecutionProvider", "CUDAExecutionProvider"])):
        fake.reset(avail)
        assert host_model.ModelSessionManager(ModelConfig(model_cache_dir=str(model_dir))).providers == gold["providers"][key]


def test_load_models_event_for_event(gold, model_dir, fake):
    want = gold["load"]
    m = host_model.ModelSessionManager(ModelConfig(model_cache_dir=str(model_dir), **want["config"]))
    m.load_models()
    got = list(fake.EVENTS)
    assert len(got) == len(want["events"])
    for g, w in zip(got, want["events"]):
        assert g[0] == w[0]
        if g[0] == "InferenceSession":
            # the mirror adds ONE config entry of its own (the fuse_nfe stride for the shim); everything the reference
            # sets must arrive unchanged: same bytes, same order, same attributes, same providers
            extra = {k: v for k, v in g[1]["entries"].items() if k not in w[1]["entries"]}
            assert set(extra) <= {"vvb200.fuse_nfe"}
            g[1]["entries"] = {k: v for k, v in g[1]["entries"].items() if k in w[1]["entries"]}
            assert g[1] == w[1]
        else:
            assert g == w
    assert m.input_names == want["input_names"] and m.output_names == want["output_names"]
    assert list(m.sessions) == want["session_keys"]                   # preprocess, transformer, decode: in that order
    with open(m.vocab_path, "rb") as f:
        assert hashlib.sha256(f.read()).hexdigest() == want["vocab_sha256"]
    assert os.path.basename(m.vocab_path) == want["vocab_basename"]
    assert os.path.basename(m.temp_dir)[:10] == want["temp_dir_prefix"]
    assert len(m.sample_metadata) == want["n_metadata"]
    td = m.temp_dir
    m.cleanup()
    assert (not os.path.exists(td)) == want["cleanup_removed_temp_dir"]
    assert m.vocab_path == want["vocab_path_after_cleanup"]


def test_load_errors_are_wrapped_like_the_reference(gold, model_dir, fake, tmp_path):
    for key, drop in (("no_decode", "decode.onnx"), ("no_vocab", "vocab.txt"), ("no_metadata", "audio_metadata.json")):
        d2 = tmp_path / key
        d2.mkdir()
        with tarfile.open(str(model_dir / "model-bin.pt")) as src, tarfile.open(str(d2 / "model-bin.pt"), "w") as dst:
            for mem in src.getmembers():
                if not mem.name.endswith(drop):
                    dst.addfile(mem, src.extractfile(mem) if mem.isfile() else None)
        fake.reset()
        m = host_model.ModelSessionManager(ModelConfig(model_cache_dir=str(d2)))
        kind, etype, msg, temp_dir, vocab_path = gold["load_errors"][key]
        assert kind == "err"
        with pytest.raises(Exception) as ei:
            m.load_models()
        assert type(ei.value).__name__ == etype and str(ei.value) == msg
        assert m.temp_dir == temp_dir and m.vocab_path == vocab_path


def test_select_sample_matches_the_reference_case_by_case(gold, model_dir, fake, tmp_path):
    with tarfile.open(str(model_dir / "model-bin.pt")) as tar:
        meta = json.load(tar.extractfile("audio_metadata.json"))
        wav = {tar.extractfile("cleaned_audios/" + s["file_name"]).read(): i for i, s in enumerate(meta)}
        prompt = tmp_path / "custom_prompt.wav"
        prompt.write_bytes(tar.extractfile("cleaned_audios/" + meta[0]["file_name"]).read())
    missing = str(tmp_path / "missing.wav")
    managers = {}
    for dkey, dflt in gold["defaults"].items():
        m = host_model.ModelSessionManager(ModelConfig(model_cache_dir=str(model_dir), **dflt))
        m.sample_metadata = meta
        managers[dkey] = m
    names = {**wav, str(prompt): "<custom prompt>"}
    n = 0
    for dkey, kw, want in gold["select"]:
        kw = {k: (str(prompt) if v == "<custom prompt>" else (missing if v == "<missing>" else v)) for k, v in kw.items()}
        try:
            audio, text = managers[dkey].select_sample(**kw)
            got = ["ok", names.get(audio, audio if isinstance(audio, str) else "<bytes?>"), text]
        except Exception as exc:
            got = ["err", type(exc).__name__, str(exc).replace(missing, "<missing>")]
        assert got == want, (dkey, kw, got, want)
        n += 1
    assert n == len(gold["select"]) > 1300
