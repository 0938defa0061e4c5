"""Generates the host-side golden fixtures by RUNNING THE REFERENCE'S OWN MODULES in the build container
(/root/reference is not available on the GPU box, so the vectors are committed next to this script).

    python tests/golden/make_host_goldens.py

Outputs: tests/golden/host_text.json, tests/golden/host_audio.npz
The reference's `core` package cannot be imported as a package (core/__init__.py pulls onnxruntime), so the two
pure-Python modules are loaded by file path; `soundfile` / `pydub` (absent offline, unused by the static methods
exercised here) are stubbed for the import only.
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference/vietvoicetts/core"
HERE = os.path.dirname(os.path.abspath(__file__))


def load(name):
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    for stub in ("soundfile", "pydub"):
        if stub not in sys.modules:
            m = types.ModuleType(stub)
            m.AudioSegment = object
            sys.modules[stub] = m
    tp_mod = load("text_processor")
    ap_mod = load("audio_processor")

    # ------------------------------------------------------------------ text
    vocab_path = os.path.join(HERE, "vocab_small.txt")
    chars = list(" .,!?abcdefghijklmnopqrstuvwxyzàáảãạăâđêôơưABCĐ0123456789'")
    with open(vocab_path, "w", encoding="utf-8") as f:
        f.write("\n".join(chars) + "\n")
    tp = tp_mod.TextProcessor(vocab_path)
    rng = np.random.default_rng(9527)
    raw_texts = [
        "  a;b:c(d)   efg! ",
        "Xin chào! Tôi là trợ lý AI.\nHôm nay thế nào? 🇻🇳",
        'Giá 5$ - 10% [ok] "quote" a—b',
        "Xin chào Việt Nam.",
        "",
        "   ",
        "dòng một\n\n  dòng hai.  \ndòng ba",
        "a,,,,b....c;;;d",
        "Hà Nội là thủ đô của nước Cộng hòa Xã hội chủ nghĩa Việt Nam, nằm ở phía tây bắc của vùng đồng bằng "
        "châu thổ sông Hồng. Thành phố có lịch sử hơn một nghìn năm! Bạn đã đến đây chưa? Tôi thì rồi, nhiều lần.",
        "This is a long sentence. This is another long sentence. And a third one.",
        "supercalifragilisticexpialidociousandevenlongerthanthatbyfar",
        "one two three four five six seven eight nine ten eleven twelve thirteen fourteen fifteen sixteen",
        "Tab\there\r\nCRLF line",
        "kết thúc bằng dấu phẩy,",
        "ĐÂY LÀ CHỮ HOA VIỆT NAM: ỲỴỶỸÝ!",
    ]
    words = ["xin", "chào", "việt", "nam", "hôm", "nay", "trời", "đẹp", "quá", "tôi", "đi", "học", "a", "b.",
             "c,", "d!", "e?", "dài_lắm_luôn_đấy_nhé", "ok", "1", "22", "333,", "rồi.", "ư", "ơi!"]
    for _ in range(40):
        n = int(rng.integers(1, 60))
        raw_texts.append(" ".join(rng.choice(words, size=n)))
    cases = []
    for t in raw_texts:
        cleaned = tp.clean_text(t)
        entry = {"raw": t, "clean": cleaned,
                 "len_default": tp.calculate_text_length(cleaned, r".,?!:"),
                 "len_class": tp.calculate_text_length(cleaned, r"[,.]"),
                 "ids": tp.text_to_indices([list(cleaned)]).tolist(),
                 "chunks": {}}
        for mc in (10, 30, 60, 135):
            entry["chunks"][str(mc)] = tp.chunk_text(cleaned, max_chars=mc)
        entry["chunks_raw_30"] = tp.chunk_text(t, max_chars=30)
        cases.append(entry)
    with open(os.path.join(HERE, "host_text.json"), "w", encoding="utf-8") as f:
        json.dump({"vocab": chars, "cases": cases}, f, ensure_ascii=False, indent=0)

    # ------------------------------------------------------------------ audio
    AP = ap_mod.AudioProcessor
    out = {}
    sig = (rng.standard_normal(5000) * 3000 + 200).astype(np.float32)
    out["norm_in"] = sig
    out["norm_out"] = AP.normalize_to_int16(sig)
    out["norm_zero_out"] = AP.normalize_to_int16(np.zeros(16, dtype=np.float32))
    clip = (rng.standard_normal(3000) * 9000).astype(np.float32)
    clip[10] = 40000.0
    clip[11] = np.nan
    clip[12] = np.inf
    out["clip_in"] = clip
    out["clip_out"] = AP.fix_clipped_audio(clip)
    ok = (rng.standard_normal(100) * 1000).astype(np.int16)
    out["noclip_in"] = ok
    out["noclip_out"] = AP.fix_clipped_audio(ok)
    waves = [(rng.standard_normal(n) * a).astype(np.int16).reshape(1, 1, -1)
             for n, a in ((6000, 4000), (5000, 900), (2000, 50), (7000, 12000), (1000, 3000))]
    waves[3][0, 0, 5] = 32767
    for i, w in enumerate(waves):
        out[f"wave{i}"] = w
    out["xf_improved"] = AP.concatenate_with_crossfade_improved(waves, 0.1, 24000)
    out["xf_improved_two"] = AP.concatenate_with_crossfade_improved(waves[:2], 0.1, 24000)
    out["xf_improved_nofade"] = AP.concatenate_with_crossfade_improved(waves[:3], 0.0, 24000)
    out["xf_improved_single"] = AP.concatenate_with_crossfade_improved(waves[:1], 0.1, 24000)
    out["xf_plain"] = AP.concatenate_with_crossfade(waves, 0.1, 24000)
    out["xf_plain_short"] = AP.concatenate_with_crossfade([waves[4], waves[2]], 0.1, 24000)
    # regular regime of the fold (every chunk at least two fades long: what the device cross-fade covers): a loud, a
    # clipped (+32767 and -32768 present), a quiet (RMS < 100: no level matching), a much louder (ratio clipped to 0.7)
    # and a much softer chunk (ratio clipped to 1.5, wraps int16 where 1.5 x overflows), seven chunks in all
    rr = np.random.default_rng(424242)
    reg = []
    for n, a in ((9000, 4000), (12000, 9000), (7000, 30), (10000, 15000), (6000, 700), (8000, 21000), (5000, 2500)):
        reg.append(np.clip(rr.standard_normal(n) * a, -32768, 32767).astype(np.int16).reshape(1, 1, -1))
    reg[1][0, 0, 77] = 32767
    reg[1][0, 0, 78] = -32768
    reg[5][0, 0, :2400] = np.clip(reg[5][0, 0, :2400].astype(np.int32) * 3 // 2, -32768, 32767).astype(np.int16)
    for i, w in enumerate(reg):
        out[f"xfr_wave{i}"] = w
    out["xfr_out"] = AP.concatenate_with_crossfade_improved(reg, 0.1, 24000)
    out["xfr_out_two"] = AP.concatenate_with_crossfade_improved(reg[3:5], 0.05, 24000)
    np.savez_compressed(os.path.join(HERE, "host_audio.npz"), **out)

    # ------------------------------------------------------------------ TTSEngine._prepare_inputs (core/tts_engine.py:43-131)
    # the package imports onnxruntime at import time; our onnxruntime-shaped shim satisfies that import only
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from vietvoice_tts_b200 import ort_shim
    ort_shim.install()
    sys.path.insert(0, "/root/reference")
    from vietvoicetts.core.tts_engine import TTSEngine as RefEngine
    from vietvoicetts.core.model_config import ModelConfig as RefConfig

    class FakeAudio:
        def __init__(self, n):
            self.n = n

        def load_audio(self, path, sr):
            return (np.arange(self.n) % 1000).astype(np.int16)

    long_text = " ".join(raw_texts[8:9] * 6)
    prep = []
    for n_samples, ref_text, target, speed, max_dur in [
        (144000, "xin chào, đây là giọng mẫu.", "Xin chào Việt Nam.", 0.9, 20.0),
        (144000, "xin chào, đây là giọng mẫu.", long_text, 0.9, 20.0),
        (217689, "một câu tham chiếu khá dài để làm mẫu giọng nói, có dấu phẩy.", long_text, 1.0, 20.0),
        (72000, "ngắn.", "a", 0.5, 20.0),
        (100000, "câu mẫu thứ ba!", long_text + " " + long_text, 1.3, 15.0),
        (144000, "xin chào.", "supercalifragilisticexpialidocious " * 30, 0.9, 20.0),
    ]:
        cfg = RefConfig.__new__(RefConfig)
        for k, v in dict(sample_rate=24000, hop_length=256, speed=speed, pause_punctuation=r".,?!:",
                         max_chunk_duration=max_dur, min_target_duration=1.0).items():
            setattr(cfg, k, v)
        eng = RefEngine.__new__(RefEngine)
        eng.config = cfg
        eng.text_processor = tp
        eng.audio_processor = FakeAudio(n_samples)
        res = eng._prepare_inputs("unused.wav", ref_text, target)
        prep.append({"n_samples": n_samples, "ref_text": ref_text, "target": target, "speed": speed,
                     "max_chunk_duration": max_dur,
                     "chunks": [{"ids": r[1].tolist(), "max_duration": int(r[2][0]), "time_step": int(r[3][0])}
                                for r in res]})
    with open(os.path.join(HERE, "host_prepare_inputs.json"), "w", encoding="utf-8") as f:
        json.dump(prep, f, ensure_ascii=False)
    print("wrote host_prepare_inputs.json:", [len(p["chunks"]) for p in prep])
    print("wrote host_text.json (%d cases), host_audio.npz (%d arrays)" % (len(cases), len(out)))


if __name__ == "__main__":
    main()
