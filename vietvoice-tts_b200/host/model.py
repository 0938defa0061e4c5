"""Session manager: owns the three sessions, the vocabulary file and the voice-sample metadata.

Mirror of the reference interface `ModelSessionManager` (/root/reference/vietvoicetts/core/model.py:18-224): same
attributes (`sessions`, `input_names`, `output_names`, `sample_metadata`, `vocab_path`, `providers`), same tar member
names, same voice-selection rules, error types and messages (callers match on them).  Sessions are created through
`vietvoice_tts_b200.ort_shim` (libvvb200.so) instead of onnxruntime.  Additions that do not change results: the prompt
bytes are cached per file name (the reference re-opens the tar on every call, model.py:204-211 — SURVEY 8f rank 1),
and `engine` exposes the shared B200 engine for the batched fast path.
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

# graph key -> archive member suffix (model.py:73-77)
_GRAPH_FILES = (("preprocess", "preprocess.onnx"), ("transformer", "transformer.onnx"), ("decode", "decode.onnx"))
# voice attribute -> values the metadata table uses (model.py:151-167)
_VOICE_DOMAINS = (("gender", MODEL_GENDER), ("group", MODEL_GROUP), ("area", MODEL_AREA), ("emotion", MODEL_EMOTION))
# numeric / boolean session options copied verbatim from the config (model.py:52-57)
_COPIED_OPTIONS = ("log_severity_level", "log_verbosity_level", "inter_op_num_threads", "intra_op_num_threads",
                   "enable_cpu_mem_arena")
_SPIN_ENTRIES = ("session.intra_op.allow_spinning", "session.inter_op.allow_spinning", "session.set_denormal_as_zero")


def _member_bytes(tar: tarfile.TarFile, names: List[str], suffix: str, missing: str, unreadable: str) -> bytes:
    """Bytes of the first archive member whose name ends with `suffix`."""
    hit = [n for n in names if n.endswith(suffix)]
    if not hit:
        raise FileNotFoundError(missing)
    handle = tar.extractfile(hit[0])
    if not handle:
        raise RuntimeError(unreadable)
    return handle.read()


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

    # ------------------------------------------------------------------------------------------ sessions
    def _get_optimal_providers(self) -> List[str]:
        have = set(onnxruntime.get_available_providers())
        order = ["CUDAExecutionProvider"] if "CUDAExecutionProvider" in have else []
        return order + ["CPUExecutionProvider"]

    def _create_session_options(self) -> onnxruntime.SessionOptions:
        so = onnxruntime.SessionOptions()
        for attr in _COPIED_OPTIONS:
            setattr(so, attr, getattr(self.config, attr))
        so.execution_mode = onnxruntime.ExecutionMode.ORT_SEQUENTIAL
        so.graph_optimization_level = onnxruntime.GraphOptimizationLevel.ORT_ENABLE_ALL
        for entry in _SPIN_ENTRIES:
            so.add_session_config_entry(entry, "1")
        so.add_session_config_entry("vvb200.fuse_nfe", str(self.config.fuse_nfe))
        return so

    def _bind_session(self, graph: str, blob: bytes) -> None:
        sess = onnxruntime.InferenceSession(blob, sess_options=self._create_session_options(), providers=self.providers)
        self.sessions[graph] = sess
        self.input_names[graph] = [node.name for node in sess.get_inputs()]        # order is the ABI
        self.output_names[graph] = [node.name for node in sess.get_outputs()]

    def _drop_temp_dir(self) -> None:
        if self.temp_dir and Path(self.temp_dir).exists():
            shutil.rmtree(self.temp_dir)
        self.temp_dir = None

    def _load_models_from_file(self) -> None:
        archive = self.config.ensure_model_downloaded()
        if not Path(archive).exists():
            raise FileNotFoundError(f"Model file not found: {archive}")
        try:
            with tarfile.open(archive, "r") as tar:
                names = tar.getnames()
                self.sample_metadata = json.load(tar.extractfile("audio_metadata.json"))
                for graph, fname in _GRAPH_FILES:
                    self._bind_session(graph, _member_bytes(
                        tar, names, fname, f"Model file '{fname}' not found in model archive",
                        f"Failed to extract {fname} from model archive"))
                vocab = _member_bytes(tar, names, "vocab.txt", "Vocabulary file 'vocab.txt' not found in model archive",
                                      "Failed to extract vocab.txt from model archive")
            self.temp_dir = tempfile.mkdtemp(prefix="tts_vocab_")
            target = Path(self.temp_dir, "vocab.txt")
            target.write_bytes(vocab)
            self.vocab_path = str(target)
        except Exception as exc:
            self._drop_temp_dir()
            raise RuntimeError(f"Failed to load models from file: {str(exc)}")

    def load_models(self) -> None:
        seed = self.config.random_seed
        onnxruntime.set_seed(seed)
        random.seed(seed)
        self._load_models_from_file()

    @property
    def engine(self):
        """The B200 engine shared by the three sessions (batched fast path)."""
        shared = self.sessions["transformer"]._sh
        shared.ensure_final()
        return shared.engine

    # ------------------------------------------------------------------------------------------ voices
    def _voice_filters(self, asked: Dict[str, Optional[str]]) -> Dict[str, str]:
        """Config defaults are merged BEFORE validation and filtering."""
        out: Dict[str, str] = {}
        for key, domain in _VOICE_DOMAINS:
            value = asked[key] or getattr(self.config, key)
            if value is None:
                continue
            if value not in domain:
                raise ValueError(f"Invalid {key}: {value}. Must be one of {domain}")
            out[key] = value
        return out

    def _prompt_bytes(self, file_name: str) -> bytes:
        cached = self._prompt_cache.get(file_name)
        if cached is not None:
            return cached
        with tarfile.open(self.config.ensure_model_downloaded(), "r") as tar:
            handle = tar.extractfile("cleaned_audios/" + file_name)
            if not handle:
                raise FileNotFoundError(f"Audio file {file_name} not found in model archive")
            data = handle.read()
        self._prompt_cache[file_name] = data
        return data

    def select_sample(self, gender: Optional[str] = None, group: Optional[str] = None, area: Optional[str] = None,
                      emotion: Optional[str] = None, sample_iteration: Optional[int] = None,
                      reference_audio: Optional[str] = None, reference_text: Optional[str] = None) -> Tuple[str, str]:
        """-> (prompt wav bytes | path, prompt text).  A custom prompt combined with non-None config defaults raises,
        exactly as upstream."""
        filters = self._voice_filters({"gender": gender, "group": group, "area": area, "emotion": emotion})

        if reference_audio is not None:                     # custom prompt: early out (model.py:169-177)
            if reference_text is None:
                raise ValueError("Reference text is required when using reference audio")
            if not Path(reference_audio).exists():
                raise FileNotFoundError(f"Reference audio file not found: {reference_audio}")
            if filters:
                raise ValueError(f"Cannot use reference audio and text with options: {list(filters.keys())}")
            logger.info(f"Using reference audio and text: {reference_audio}")
            return reference_audio, reference_text

        try:
            rows = list(enumerate(self.sample_metadata))
            hits = [(i, row) for i, row in rows if all(row[k] == v for k, v in filters.items())]
            if not hits:
                pick = 0                                    # silent fallback to the first voice, as upstream
            elif sample_iteration is None:
                pick = hits[0][0]
            elif sample_iteration >= len(hits):
                raise ValueError(f"sample_iteration {sample_iteration} is out of range. Only {len(hits)} "
                                 f"samples available for the given filters.")
            else:
                pick = hits[sample_iteration][0]
            row = self.sample_metadata[pick]
            logger.info(f"Selected sample #{pick} with gender: {row['gender']}, group: {row['group']}, "
                        f"area: {row['area']}, emotion: {row['emotion']}")
            return self._prompt_bytes(row["file_name"]), row["text"]
        except KeyError:
            raise ValueError(f"Sample not found for gender: {gender or self.config.gender}, "
                             f"group: {group or self.config.group}, area: {area or self.config.area}, "
                             f"emotion: {emotion or self.config.emotion}")

    # ------------------------------------------------------------------------------------------ lifetime
    def cleanup(self) -> None:
        if self.temp_dir:
            self._drop_temp_dir()
            self.vocab_path = None

    def __del__(self):
        try:
            self.cleanup()
        except Exception:
            pass
