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


@dataclass
class ModelConfig:
    # model artefact
    model_url: str = "https://huggingface.co/nguyenvulebinh/VietVoice-TTS/resolve/main/model-bin.pt"
    model_cache_dir: str = "models"
    model_filename: str = "model-bin.pt"
    # sampler
    nfe_step: int = 32
    fuse_nfe: int = 1
    sample_rate: int = 24000
    speed: float = 0.9
    random_seed: int = 9527
    hop_length: int = 256
    # voice selection
    gender: Optional[str] = "female"
    area: Optional[str] = "northern"
    emotion: Optional[str] = "neutral"
    group: Optional[str] = "audiobook"
    # text
    pause_punctuation: str = r".,?!:"
    # audio
    cross_fade_duration: float = 0.1
    max_chunk_duration: float = 20.0
    min_target_duration: float = 1.0
    # executor knobs of the reference (accepted, unused by the B200 engine)
    log_severity_level: int = 4
    log_verbosity_level: int = 4
    inter_op_num_threads: int = 0
    intra_op_num_threads: int = 0
    enable_cpu_mem_arena: bool = True

    def __post_init__(self):
        if not 0.1 <= self.speed <= 5.0:
            raise ValueError("Speed must be between 0.1 and 5.0")
        if not 1 <= self.nfe_step <= 100:
            raise ValueError("NFE step must be between 1 and 100")
        self.validate_paths()

    @property
    def model_path(self) -> str:
        return str(Path(self.model_cache_dir).expanduser() / self.model_filename)

    def ensure_model_downloaded(self) -> str:
        target = Path(self.model_path)
        target.parent.mkdir(parents=True, exist_ok=True)
        if target.exists():
            logger.info(f"Using cached model: {target}")
            return str(target)
        logger.info(f"Downloading model from {self.model_url}")
        try:
            urllib.request.urlretrieve(self.model_url, target)
        except urllib.error.URLError as exc:
            raise RuntimeError(f"Failed to download model from {self.model_url}: {exc}")
        except Exception as exc:
            if target.exists():
                target.unlink()          # never leave a partial download behind
            raise RuntimeError(f"Failed to download model: {exc}")
        return str(target)

    def validate_paths(self):
        try:
            self.ensure_model_downloaded()
        except Exception as exc:
            raise RuntimeError(f"Model validation failed: {exc}")

    def validate_with_reference_audio(self, reference_audio_path: str) -> bool:
        """True when prompt + 1 s safety margin + min_target_duration fits in max_chunk_duration."""
        try:
            import wave
            with wave.open(reference_audio_path, "rb") as w:
                ref_duration = w.getnframes() / float(w.getframerate())
        except Exception as exc:
            logger.error(f"Error validating reference audio: {exc}")
            return False
        need = ref_duration + 1.0 + self.min_target_duration
        if self.max_chunk_duration < need:
            logger.error(f"Configuration Error: reference audio {ref_duration:.1f}s needs max_chunk_duration > {need:.1f}s "
                         f"(current {self.max_chunk_duration:.1f}s)")
            return False
        return True

    @classmethod
    def from_dict(cls, config_dict: dict) -> "ModelConfig":
        return cls(**config_dict)

    def to_dict(self) -> dict:
        return {f.name: getattr(self, f.name) for f in fields(self)}


TTSConfig = ModelConfig   # backward-compatible alias, as in the reference
