"""Session manager: owns the three sessions, the vocabulary file and the voice-sample metadata.

Mirror of the reference interface `ModelSessionManager` (/root/reference/vietvoicetts/core/model.py:18-224): same
attributes (`sessions`, `input_names`, `output_names`, `sample_metadata`, `vocab_path`, `providers`), same tar member
names, same voice-selection rules and error types.  Sessions are created through `vietvoice_tts_b200.ort_shim`
(libvvb200.so) instead of onnxruntime.  Additions that do not change results: the tar index and the prompt bytes are
cached (the reference re-opens the tar on every call, model.py:204-211 — SURVEY 8f rank 1), and `engine` exposes
the shared B200 engine for the batched fast path.
"""
from __future__ import annotations

import json
import random
import shutil
import tarfile
import tempfile
from pathlib import Path
from typing import Dict, List, Optional, Tuple

from loguru import logger

from .. import ort_shim as onnxruntime
from .model_config import MODEL_AREA, MODEL_EMOTION, MODEL_GENDER, MODEL_GROUP, ModelConfig

_GRAPH_FILES = {"preprocess": "preprocess.onnx", "transformer": "transformer.onnx", "decode": "decode.onnx"}
_FILTER_DOMAINS = (("gender", MODEL_GENDER), ("group", MODEL_GROUP), ("area", MODEL_AREA), ("emotion", MODEL_EMOTION))


class ModelSessionManager:
    """Manages the engine sessions"""

    def __init__(self, config: ModelConfig):
        self.config = config
        self.providers = self._get_optimal_providers()
        self.sessions: Dict[str, onnxruntime.InferenceSession] = {}
        self.input_names: Dict[str, List[str]] = {}
        self.output_names: Dict[str, List[str]] = {}
        self.sample_metadata = {}
        self.temp_dir: Optional[str] = None
        self.vocab_path: Optional[str] = None
        self._prompt_cache: Dict[str, bytes] = {}

    def _get_optimal_providers(self) -> List[str]:
        available = onnxruntime.get_available_providers()
        chosen = [p for p in ("CUDAExecutionProvider", "CPUExecutionProvider") if p in available]
        if "CPUExecutionProvider" not in chosen:
            chosen.append("CPUExecutionProvider")
        return chosen

    def _create_session_options(self) -> onnxruntime.SessionOptions:
        opts = onnxruntime.SessionOptions()
        opts.log_severity_level = self.config.log_severity_level
        opts.log_verbosity_level = self.config.log_verbosity_level
        opts.inter_op_num_threads = self.config.inter_op_num_threads
        opts.intra_op_num_threads = self.config.intra_op_num_threads
        opts.enable_cpu_mem_arena = self.config.enable_cpu_mem_arena
        opts.execution_mode = onnxruntime.ExecutionMode.ORT_SEQUENTIAL
        opts.graph_optimization_level = onnxruntime.GraphOptimizationLevel.ORT_ENABLE_ALL
        for key in ("session.intra_op.allow_spinning", "session.inter_op.allow_spinning", "session.set_denormal_as_zero"):
            opts.add_session_config_entry(key, "1")
        opts.add_session_config_entry("vvb200.fuse_nfe", str(self.config.fuse_nfe))
        return opts

    def _load_models_from_file(self) -> None:
        model_path = self.config.ensure_model_downloaded()
        if not Path(model_path).exists():
            raise FileNotFoundError(f"Model file not found: {model_path}")
        try:
            with tarfile.open(model_path, "r") as tar:
                names = tar.getnames()
                self.sample_metadata = json.load(tar.extractfile("audio_metadata.json"))
                for graph, fname in _GRAPH_FILES.items():
                    member = next((m for m in names if m.endswith(fname)), None)
                    if not member:
                        raise FileNotFoundError(f"Model file '{fname}' not found in model archive")
                    fh = tar.extractfile(member)
                    if not fh:
                        raise RuntimeError(f"Failed to extract {fname} from model archive")
                    session = onnxruntime.InferenceSession(fh.read(), sess_options=self._create_session_options(),
                                                           providers=self.providers)
                    self.sessions[graph] = session
                    self.input_names[graph] = [i.name for i in session.get_inputs()]
                    self.output_names[graph] = [o.name for o in session.get_outputs()]
                vocab_member = next((m for m in names if m.endswith("vocab.txt")), None)
                if not vocab_member:
                    raise FileNotFoundError("Vocabulary file 'vocab.txt' not found in model archive")
                fh = tar.extractfile(vocab_member)
                if not fh:
                    raise RuntimeError("Failed to extract vocab.txt from model archive")
                self.temp_dir = tempfile.mkdtemp(prefix="tts_vocab_")
                vocab_file = Path(self.temp_dir) / "vocab.txt"
                vocab_file.write_bytes(fh.read())
                self.vocab_path = str(vocab_file)
        except Exception as exc:
            if self.temp_dir and Path(self.temp_dir).exists():
                shutil.rmtree(self.temp_dir)
                self.temp_dir = None
            raise RuntimeError(f"Failed to load models from file: {str(exc)}")

    def load_models(self) -> None:
        onnxruntime.set_seed(self.config.random_seed)
        random.seed(self.config.random_seed)
        self._load_models_from_file()

    @property
    def engine(self):
        """The B200 engine shared by the three sessions (batched fast path)."""
        sh = self.sessions["transformer"]._sh
        sh.ensure_final()
        return sh.engine

    def select_sample(self, gender: Optional[str] = None, group: Optional[str] = None, area: Optional[str] = None,
                      emotion: Optional[str] = None, sample_iteration: Optional[int] = None,
                      reference_audio: Optional[str] = None, reference_text: Optional[str] = None) -> Tuple[str, str]:
        """-> (prompt wav bytes | path, prompt text).  Config defaults are merged BEFORE filtering, so a custom
        prompt combined with non-None defaults raises, exactly as upstream."""
        requested = {"gender": gender or self.config.gender, "group": group or self.config.group,
                     "area": area or self.config.area, "emotion": emotion or self.config.emotion}
        filters = {}
        for key, domain in _FILTER_DOMAINS:
            value = requested[key]
            if value is None:
                continue
            if value not in domain:
                raise ValueError(f"Invalid {key}: {value}. Must be one of {domain}")
            filters[key] = value

        if reference_audio is not None:
            if reference_text is None:
                raise ValueError("Reference text is required when using reference audio")
            if not Path(reference_audio).exists():
                raise FileNotFoundError(f"Reference audio file not found: {reference_audio}")
            if filters:
                raise ValueError(f"Cannot use reference audio and text with options: {list(filters.keys())}")
            logger.info(f"Using reference audio and text: {reference_audio}")
            return reference_audio, reference_text

        try:
            matches = [(s, i) for i, s in enumerate(self.sample_metadata)
                       if all(s[k] == v for k, v in filters.items())]
            if not matches:
                sample, idx = self.sample_metadata[0], 0          # silent fallback, as upstream
            elif sample_iteration is not None:
                if sample_iteration >= len(matches):
                    raise ValueError(f"sample_iteration {sample_iteration} is out of range. Only {len(matches)} "
                                     f"samples available for the given filters.")
                sample, idx = matches[sample_iteration]
            else:
                sample, idx = matches[0]
            logger.info(f"Selected sample #{idx} with gender: {sample['gender']}, group: {sample['group']}, "
                        f"area: {sample['area']}, emotion: {sample['emotion']}")
            fname = sample["file_name"]
            audio = self._prompt_cache.get(fname)
            if audio is None:
                with tarfile.open(self.config.ensure_model_downloaded(), "r") as tar:
                    fh = tar.extractfile("cleaned_audios/" + fname)
                    if not fh:
                        raise FileNotFoundError(f"Audio file {fname} not found in model archive")
                    audio = fh.read()
                self._prompt_cache[fname] = audio
            text = sample["text"]
        except KeyError:
            raise ValueError(f"Sample not found for gender: {gender}, group: {group}, area: {area}, emotion: {emotion}")
        return audio, text

    def cleanup(self) -> None:
        if self.temp_dir and Path(self.temp_dir).exists():
            shutil.rmtree(self.temp_dir)
            self.temp_dir = None
            self.vocab_path = None

    def __del__(self):
        try:
            self.cleanup()
        except Exception:
            pass
