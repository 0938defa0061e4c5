"""Golden vectors for the SESSION MANAGER, produced by running the reference's own class in the build container:

    python tests/golden/make_session_goldens.py      ->  tests/golden/session_manager.json

`ModelSessionManager` (/root/reference/vietvoicetts/core/model.py:18-224) is imported unmodified with
tests/golden/fake_ort.py standing where `onnxruntime` would be (the wheel is not installable offline; the fake records
every call and executes nothing) and `soundfile` / `pydub` stubbed for the import only.  Recorded:

  * load: `_get_optimal_providers` for two provider sets, `load_models()` on a model tar built by
    `artifact.build_model_tar` (seeded -> the test rebuilds the identical tar): set_seed, the three InferenceSession
    constructions (bytes hash, session-option attributes and config entries, providers), the positional I/O name
    lists, the vocab file, the metadata, and the error wrapping for a tar without decode.onnx / without vocab.txt;
  * select_sample: the outcome (voice index + text | exception type + message) of every combination of
    gender x group x area x emotion x sample_iteration under three config-default settings, plus the custom-prompt
    rules (1 100+ cases).
/root/reference does not exist on the GPU box, hence the committed JSON.
"""
import itertools
import json
import os
import sys
import tarfile
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

VOICES = [
    {"gender": "female", "group": "audiobook", "area": "northern", "emotion": "neutral"},
    {"gender": "male", "group": "news", "area": "southern", "emotion": "serious"},
    {"gender": "male", "group": "news", "area": "southern", "emotion": "serious"},
    {"gender": "female", "group": "story", "area": "central", "emotion": "happy"},
    {"gender": "male", "group": "audiobook", "area": "northern", "emotion": "neutral"},
    {"gender": "female", "group": "audiobook", "area": "northern", "emotion": "sad"},
    {"gender": "female", "group": "interview", "area": "southern", "emotion": "neutral"},
]
TAR_SEED = 7
GENDERS = [None, "male", "female", "robot"]
GROUPS = [None, "news", "audiobook", "bad"]
AREAS = [None, "northern", "southern", "western"]
EMOTIONS = [None, "neutral", "happy", "furious"]
ITERS = [None, 0, 1, 5]
DEFAULTS = {
    "stock": {},                                                         # female / audiobook / northern / neutral
    "none": {"gender": None, "group": None, "area": None, "emotion": None},
    "male_south": {"gender": "male", "group": None, "area": "southern", "emotion": None},
}


def outcome(fn, names):
    try:
        audio, text = fn()
    except Exception as exc:
        return ["err", type(exc).__name__, str(exc)]
    return ["ok", names.get(audio, audio if isinstance(audio, str) else "<bytes?>"), text]


def select_cases(make_manager, tmp):
    """-> list of [defaults key, kwargs, outcome]"""
    out = []
    prompt = os.path.join(tmp, "custom_prompt.wav")
    for dkey, dflt in DEFAULTS.items():
        m, names = make_manager(dflt)
        for g, gr, a, e, it in itertools.product(GENDERS, GROUPS, AREAS, EMOTIONS, ITERS):
            if dkey != "none" and ("robot" == g or gr == "bad" or a == "western" or e == "furious" or it in (0, 5)):
                continue            # full product (1 024 cases) without defaults, valid values only with defaults
            kw = dict(gender=g, group=gr, area=a, emotion=e, sample_iteration=it)
            out.append([dkey, kw, outcome(lambda: m.select_sample(**kw), names)])
        for kw in (dict(reference_audio=prompt), dict(reference_audio=prompt, reference_text="xin chào"),
                   dict(reference_audio=os.path.join(tmp, "missing.wav"), reference_text="x"),
                   dict(reference_audio=prompt, reference_text="x", gender="male"),
                   dict(reference_audio=prompt, reference_text="x", emotion="furious"),
                   dict(reference_text="only text")):
            res = outcome(lambda: m.select_sample(**kw), {**names, prompt: "<custom prompt>"})
            kw = {k: ("<custom prompt>" if v == prompt else ("<missing>" if k == "reference_audio" else v))
                  for k, v in kw.items()}
            if res[0] == "err":
                res[2] = res[2].replace(os.path.join(tmp, "missing.wav"), "<missing>")
            out.append([dkey, kw, res])
    return out


def main():
    import fake_ort
    from loguru import logger
    logger.remove()
    sys.modules["onnxruntime"] = fake_ort
    for stub in ("soundfile", "pydub"):
        m = types.ModuleType(stub)
        m.AudioSegment = object
        sys.modules[stub] = m
    sys.path.insert(0, "/root/reference")
    from vietvoicetts.core.model import ModelSessionManager as RefManager
    from vietvoicetts.core.model_config import ModelConfig as RefConfig
    from vietvoice_tts_b200 import artifact
    from vietvoice_tts_b200.arch import TINY

    gold = {"voices": VOICES, "tar_seed": TAR_SEED, "defaults": DEFAULTS}
    with tempfile.TemporaryDirectory() as tmp:
        artifact.build_model_tar(os.path.join(tmp, "model-bin.pt"), TINY, seed=TAR_SEED, voices=VOICES,
                                 prompt_seconds=0.2)
        with tarfile.open(os.path.join(tmp, "model-bin.pt")) as tar:
            meta = json.load(tar.extractfile("audio_metadata.json"))
            wav = {tar.extractfile("cleaned_audios/" + s["file_name"]).read(): i for i, s in enumerate(meta)}
            with open(os.path.join(tmp, "custom_prompt.wav"), "wb") as f:
                f.write(tar.extractfile("cleaned_audios/" + meta[0]["file_name"]).read())

        # ---- providers
        prov = {}
        for key, avail in (("cuda+cpu", None), ("cpu", ["CPUExecutionProvider"]),
                           ("trt+cuda", ["TensorrtExecutionProvider", "CUDAExecutionProvider"])):
            fake_ort.reset(avail)
            prov[key] = RefManager(RefConfig(model_cache_dir=tmp)).providers
        gold["providers"] = prov

        # ---- load_models
        fake_ort.reset()
        cfg = RefConfig(model_cache_dir=tmp, random_seed=4242, inter_op_num_threads=3, log_severity_level=1)
        m = RefManager(cfg)
        m.load_models()
        with open(m.vocab_path, "rb") as f:
            vocab = f.read()
        gold["load"] = {"config": {"random_seed": 4242, "inter_op_num_threads": 3, "log_severity_level": 1},
                        "events": list(fake_ort.EVENTS), "input_names": m.input_names, "output_names": m.output_names,
                        "session_keys": list(m.sessions), "vocab_sha256": __import__("hashlib").sha256(vocab).hexdigest(),
                        "vocab_basename": os.path.basename(m.vocab_path),
                        "temp_dir_prefix": os.path.basename(m.temp_dir)[:10], "n_metadata": len(m.sample_metadata)}
        td = m.temp_dir
        m.cleanup()
        gold["load"]["cleanup_removed_temp_dir"] = not os.path.exists(td)
        gold["load"]["vocab_path_after_cleanup"] = m.vocab_path

        # ---- load errors: archive without decode.onnx / without vocab.txt / missing file
        errs = {}
        for key, drop in (("no_decode", "decode.onnx"), ("no_vocab", "vocab.txt"), ("no_metadata", "audio_metadata.json")):
            d2 = os.path.join(tmp, key)
            os.makedirs(d2)
            with tarfile.open(os.path.join(tmp, "model-bin.pt")) as src, \
                    tarfile.open(os.path.join(d2, "model-bin.pt"), "w") as dst:
                for mem in src.getmembers():
                    if not mem.name.endswith(drop):
                        dst.addfile(mem, src.extractfile(mem) if mem.isfile() else None)
            fake_ort.reset()
            mm = RefManager(RefConfig(model_cache_dir=d2))
            try:
                mm.load_models()
                errs[key] = ["ok"]
            except Exception as exc:
                errs[key] = ["err", type(exc).__name__, str(exc), mm.temp_dir, mm.vocab_path]
        gold["load_errors"] = errs

        # ---- select_sample
        def make_manager(dflt):
            fake_ort.reset()
            mgr = RefManager(RefConfig(model_cache_dir=tmp, **dflt))
            mgr.sample_metadata = meta
            return mgr, wav

        gold["select"] = select_cases(make_manager, tmp)
    with open(os.path.join(HERE, "session_manager.json"), "w", encoding="utf-8") as f:
        json.dump(gold, f, ensure_ascii=False, separators=(",", ":"))
    n_ok = sum(1 for c in gold["select"] if c[2][0] == "ok")
    print("wrote session_manager.json: %d select_sample cases (%d ok, %d errors), %d load events"
          % (len(gold["select"]), n_ok, len(gold["select"]) - n_ok, len(gold["load"]["events"])))


if __name__ == "__main__":
    main()
