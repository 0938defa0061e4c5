"""Inference configuration read by the engine.

Mirror of the reference interface `ModelConfig` / `TTSConfig` (/root/reference/vietvoicetts/core/model_config.py:21-156):
same field names, defaults, range checks and error types, so code written against the reference's config keeps
working.  The engine itself only READS nfe_step, fuse_nfe, sample_rate, hop_length, random_seed, speed and the
chunking limits; the download logic is host glue kept for interface parity.
"""
from __future__ import annotations

import urllib.error
import urllib.request
from dataclasses import dataclass, fields
from pathlib import Path
from typing import Optional

from loguru import logger

MODEL_GENDER = ["male", "female"]
MODEL_GROUP = ["story", "news", "audiobook", "interview", "review"]
MODEL_AREA = ["northern", "southern", "central"]
MODEL_EMOTION = ["neutral", "serious", "monotone", "sad", "surprised", "happy", "angry"]


def _fetch(url: str, target: Path) -> None:
    """Download `url` to `target`; a partial file never survives a failure (core/model_config.py:83-98)."""
    logger.info(f"Downloading model from {url}")
    try:
        urllib.request.urlretrieve(url, target)
    except urllib.error.URLError as exc:
        raise RuntimeError(f"Failed to download model from {url}: {exc}")
    except Exception as exc:
        if target.exists():
            target.unlink()
        raise RuntimeError(f"Failed to download model: {exc}")


def _wav_seconds(path: str) -> float:
    import wave
    with wave.open(path, "rb") as w:
        return w.getnframes() / float(w.getframerate())


@dataclass
class ModelConfig:
    # ---- where the artefact lives (read by ModelSessionManager._load_models_from_file / select_sample)
    model_url: str = "https://huggingface.co/nguyenvulebinh/VietVoice-TTS/resolve/main/model-bin.pt"
    model_cache_dir: str = "models"
    model_filename: str = "model-bin.pt"
    # ---- sampler: nfe_step - 1 transformer calls of stride fuse_nfe (tts_engine.py:157); speed scales the duration
    #      model of _prepare_inputs; random_seed seeds the sessions and keys the engine's Philox y0
    nfe_step: int = 32
    fuse_nfe: int = 1
    sample_rate: int = 24000
    speed: float = 0.9
    random_seed: int = 9527
    hop_length: int = 256
    # ---- default voice: merged into every select_sample call BEFORE filtering (None = no constraint)
    gender: Optional[str] = "female"
    area: Optional[str] = "northern"
    emotion: Optional[str] = "neutral"
    group: Optional[str] = "audiobook"
    # ---- text: used as a REGEX by calculate_text_length, not as a character class
    pause_punctuation: str = r".,?!:"
    # ---- chunking and joining (seconds): cross-fade overlap, chunk budget incl. the prompt, shortest useful target
    cross_fade_duration: float = 0.1
    max_chunk_duration: float = 20.0
    min_target_duration: float = 1.0
    # ---- executor knobs of the reference: accepted and forwarded to SessionOptions, without effect on the B200 engine
    log_severity_level: int = 4
    log_verbosity_level: int = 4
    inter_op_num_threads: int = 0
    intra_op_num_threads: int = 0
    enable_cpu_mem_arena: bool = True

    def __post_init__(self):
        for ok, message in ((0.1 <= self.speed <= 5.0, "Speed must be between 0.1 and 5.0"),
                            (1 <= self.nfe_step <= 100, "NFE step must be between 1 and 100")):
            if not ok:
                raise ValueError(message)
        self.validate_paths()

    @property
    def model_path(self) -> str:
        return str(Path(self.model_cache_dir).expanduser() / self.model_filename)

    def ensure_model_downloaded(self) -> str:
        """Path of the cached artefact, downloading it first if it is not there."""
        target = Path(self.model_path)
        target.parent.mkdir(parents=True, exist_ok=True)
        if target.exists():
            logger.info(f"Using cached model: {target}")
        else:
            _fetch(self.model_url, target)
        return str(target)

    def validate_paths(self):
        try:
            self.ensure_model_downloaded()
        except Exception as exc:
            raise RuntimeError(f"Model validation failed: {exc}")

    def validate_with_reference_audio(self, reference_audio_path: str) -> bool:
        """True when prompt + 1 s safety margin + min_target_duration fits in max_chunk_duration (PCM WAV prompts;
        the reference decodes any container through pydub/ffmpeg, absent offline)."""
        try:
            need = _wav_seconds(reference_audio_path) + 1.0 + self.min_target_duration
        except Exception as exc:
            logger.error(f"Error validating reference audio: {exc}")
            return False
        if self.max_chunk_duration >= need:
            return True
        logger.error(f"Configuration Error: the reference audio needs max_chunk_duration > {need:.1f}s "
                     f"(current {self.max_chunk_duration:.1f}s)")
        return False

    @classmethod
    def from_dict(cls, config_dict: dict) -> "ModelConfig":
        return cls(**config_dict)

    def to_dict(self) -> dict:
        return {f.name: getattr(self, f.name) for f in fields(self)}


TTSConfig = ModelConfig   # backward-compatible alias, as in the reference
