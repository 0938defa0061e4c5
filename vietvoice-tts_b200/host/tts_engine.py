"""TTSEngine: text + voice selection -> int16 waveform, on the B200 engine.

Mirror of the reference interface `TTSEngine` (/root/reference/vietvoicetts/core/tts_engine.py:17-267): same
constructor, `synthesize(...) -> (np.ndarray, seconds)`, `_prepare_inputs`, the three `_run_*` session wrappers
(:133-187, kept call-compatible — tests of the reference patch them) and the same error wrapping (:256-257).

What is different, by design: the reference runs the chunks of a long text one after another, 33 session calls each
(:225-238).  Chunks are independent, so `synthesize` hands ALL chunks to the engine as one batch
(`Engine.synthesize_batch`: one CUDA-graph replay for the whole sampling loop) and, when `shard` is given, deals them
across ranks.  Per-chunk y0 is keyed on (random_seed, chunk index) so the result does not depend on how
the chunks were batched or sharded.  `use_sessions=True` forces the reference's sequential session loop.
"""
from __future__ import annotations

import time
from typing import List, Optional, Tuple

import numpy as np
from loguru import logger

from .audio_processor import AudioProcessor
from .model import ModelSessionManager
from .model_config import ModelConfig
from .text_processor import TextProcessor


class TTSEngine:
    """Main TTS engine for inference"""

    def __init__(self, config: Optional[ModelConfig] = None, use_sessions: bool = False, shard=None):
        self.config = config or ModelConfig()
        self.model_session_manager = ModelSessionManager(self.config)
        self.model_session_manager.load_models()
        if not self.model_session_manager.vocab_path:
            raise RuntimeError("Vocabulary file not found in model tar archive")
        self.text_processor = TextProcessor(self.model_session_manager.vocab_path)
        self.audio_processor = AudioProcessor()
        self.sample_cache = {}
        self.use_sessions = use_sessions
        self.shard = shard               # optional vietvoice_tts_b200.shard.Sharder
        self.last_timing: dict = {}      # seconds per stage of the last synthesize() call (bench / logging only)
        self.gpu_crossfade = True        # unsharded multi-chunk texts: clip-fix + cross-fade on the device

    def cleanup(self) -> None:
        if self.model_session_manager:
            self.model_session_manager.cleanup()

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        self.cleanup()

    # ------------------------------------------------------------------------------------------ input preparation
    def _target_duration(self, text: str, rate: float, speed: Optional[float] = None) -> float:
        n = self.text_processor.calculate_text_length(text, self.config.pause_punctuation)
        return max(n / rate / (self.config.speed if speed is None else speed), self.config.min_target_duration)

    def _prepare_inputs(self, reference_audio_path_or_bytes, reference_text: str, target_text: str,
                        speed: Optional[float] = None) -> List[Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]]:
        """-> per chunk (audio int16 [1,1,N], text_ids int32 [1,L], max_duration int64 [1], time_step int32 [1]).
        `speed` overrides config.speed for this call only (the reference mutates the shared config per request,
        api/tts_engine.py:64-69; a per-call value has no such race)."""
        cfg = self.config
        speed = cfg.speed if speed is None else speed
        cache_key = reference_audio_path_or_bytes if isinstance(reference_audio_path_or_bytes, (str, bytes)) else None
        audio = self.sample_cache.get(cache_key) if cache_key is not None else None
        if audio is None:
            audio = self.audio_processor.load_audio(reference_audio_path_or_bytes, cfg.sample_rate).reshape(1, 1, -1)
            if cache_key is not None and len(self.sample_cache) < 64:
                self.sample_cache[cache_key] = audio
        reference_text = self.text_processor.clean_text(reference_text)
        target_text = self.text_processor.clean_text(target_text)

        n_samples = audio.shape[-1]
        ref_frames = n_samples // cfg.hop_length + 1
        ref_seconds = n_samples / cfg.sample_rate
        ref_units = self.text_processor.calculate_text_length(reference_text, cfg.pause_punctuation)
        rate = ref_units / ref_seconds if ref_seconds > 0 else 100
        total = ref_seconds + self._target_duration(target_text, rate, speed)

        if total <= cfg.max_chunk_duration:
            chunks = [target_text]
        else:
            budget = cfg.max_chunk_duration - ref_seconds - 1.0          # 1 s safety margin
            if budget <= 0:
                raise ValueError(f"Reference audio duration ({ref_seconds:.1f}s) exceeds max chunk duration "
                                 f"({cfg.max_chunk_duration}s)")
            chunks = []
            for piece in self.text_processor.chunk_text(target_text, max_chars=int(rate * budget * speed)):
                dur = self._target_duration(piece, rate, speed)
                if ref_seconds + dur <= cfg.max_chunk_duration:
                    chunks.append(piece)
                else:                                                    # still too long: split again, 90 % target
                    logger.warning(f"Chunk too long ({ref_seconds + dur:.1f}s), splitting further...")
                    chunks.extend(self.text_processor.chunk_text(piece, max_chars=int(len(piece) * budget / dur * 0.9)))
            logger.info(f"Long text detected (total estimated {total:.1f}s), split into {len(chunks)} chunks")

        inputs = []
        for piece in chunks:
            dur = self._target_duration(piece, rate, speed)
            frames = ref_frames + int(dur * cfg.sample_rate) // cfg.hop_length + 1
            ids = self.text_processor.text_to_indices([list(reference_text + piece)])
            inputs.append((audio, ids, np.array([frames], dtype=np.int64), np.array([0], dtype=np.int32)))
        return inputs

    # ------------------------------------------------------------------------------------------ session wrappers
    def _run_preprocess(self, audio: np.ndarray, text_ids: np.ndarray, max_duration: np.ndarray):
        m = self.model_session_manager
        names = m.input_names["preprocess"]
        return m.sessions["preprocess"].run(m.output_names["preprocess"],
                                            {names[0]: audio, names[1]: text_ids, names[2]: max_duration})

    def _run_transformer_steps(self, noise, rope_cos_q, rope_sin_q, rope_cos_k, rope_sin_k, cat_mel_text,
                               cat_mel_text_drop, time_step):
        m = self.model_session_manager
        names, outs, sess = m.input_names["transformer"], m.output_names["transformer"], m.sessions["transformer"]
        for _ in range(0, self.config.nfe_step - 1, self.config.fuse_nfe):
            noise, time_step = sess.run(outs, {names[0]: noise, names[1]: rope_cos_q, names[2]: rope_sin_q,
                                               names[3]: rope_cos_k, names[4]: rope_sin_k, names[5]: cat_mel_text,
                                               names[6]: cat_mel_text_drop, names[7]: time_step})
        return noise, time_step

    def _run_decode(self, noise: np.ndarray, ref_signal_len: np.ndarray) -> np.ndarray:
        m = self.model_session_manager
        names = m.input_names["decode"]
        return m.sessions["decode"].run(m.output_names["decode"], {names[0]: noise, names[1]: ref_signal_len})[0]

    # ------------------------------------------------------------------------------------------ synthesis
    def _synthesize_chunks_sessions(self, inputs_list) -> List[np.ndarray]:
        waves = []
        for audio, text_ids, max_duration, time_step in inputs_list:
            pre = self._run_preprocess(audio, text_ids, max_duration)
            noise, cq, sq, ck, sk, cat_c, cat_u, ref_len = pre
            noise, time_step = self._run_transformer_steps(noise, cq, sq, ck, sk, cat_c, cat_u, time_step)
            waves.append(self._run_decode(noise, ref_len))
        return waves

    def _synthesize_chunks_batched(self, inputs_list) -> List[np.ndarray]:
        eng = self.model_session_manager.engine
        n = len(inputs_list)
        if self.shard is not None:
            self.shard.arch = eng.arch                       # chunk costs from the loaded architecture
        mine = list(range(n)) if self.shard is None else self.shard.assign([int(i[2][0]) for i in inputs_list])
        local, failure = {}, None
        t0 = time.time()
        if mine:
            try:
                out = eng.synthesize_batch([inputs_list[i][0] for i in mine], [inputs_list[i][1] for i in mine],
                                           [int(inputs_list[i][2][0]) for i in mine], nfe=self.config.nfe_step,
                                           seed=self.config.random_seed, chunk_keys=mine)
                local = {i: w.reshape(1, 1, -1) for i, w in zip(mine, out)}
            except Exception as exc:
                if self.shard is None:
                    raise
                failure = exc            # still take part in the gather: the other ranks are waiting in it
        t1 = time.time()
        if self.shard is not None:
            local = self.shard.gather(local, n, error=None if failure is None else str(failure))
        self.last_timing.update(synth_s=t1 - t0, gather_s=time.time() - t1, n_chunks=n, local_chunks=len(mine))
        return [local[i] for i in range(n)]

    def _fold(self, waves: List[np.ndarray]) -> np.ndarray:
        """The reference's clip-fix / cross-fade fold over the ordered chunk list (core/tts_engine.py:244-246).  Waves
        gathered from several ranks are folded on this rank's GPU when the case is the regular one (bit-exact with
        the host code, ~10x faster than numpy for a 40-chunk text); otherwise on the host."""
        cfg = self.config
        n_fade = int(cfg.cross_fade_duration * cfg.sample_rate)
        if self.gpu_crossfade and not self.use_sessions and len(waves) >= 2:
            eng = self.model_session_manager.engine
            if eng.crossfade_ok([w.size for w in waves], n_fade) and all(w.dtype == np.int16 for w in waves):
                self.last_timing["crossfade"] = "device (gathered waves)"
                return eng.crossfade_pcm(waves, n_fade)
        self.last_timing["crossfade"] = "host"
        return self.audio_processor.concatenate_with_crossfade_improved(waves, cfg.cross_fade_duration, cfg.sample_rate)

    def _synthesize_joined(self, inputs_list) -> Optional[np.ndarray]:
        """All chunks of the text as one batch AND the reference's clip-fix / cross-fade fold
        (/root/reference/vietvoicetts/core/audio_processor.py:123-193) on the GPU, bit-exact with the host code: only
        the joined wave is copied back.  Returns None when the case is not the regular one (sharded over ranks, fewer
        than two chunks worth fading, a chunk shorter than two fades) — the caller then takes the host fold."""
        if (self.shard is not None and self.shard.world > 1) or not self.gpu_crossfade or len(inputs_list) < 2:
            return None
        cfg = self.config
        eng = self.model_session_manager.engine
        n_fade = int(cfg.cross_fade_duration * cfg.sample_rate)
        ref_frames = inputs_list[0][0].shape[-1] // cfg.hop_length + 1
        lengths = [max(0, int(i[2][0]) - ref_frames - 1) * cfg.hop_length for i in inputs_list]
        if not eng.crossfade_ok(lengths, n_fade):
            return None
        t0 = time.time()
        keys = list(range(len(inputs_list)))
        wave = eng.synthesize_batch([i[0] for i in inputs_list], [i[1] for i in inputs_list],
                                    [int(i[2][0]) for i in inputs_list], nfe=cfg.nfe_step, seed=cfg.random_seed,
                                    chunk_keys=keys, join_fade=n_fade)
        self.last_timing.update(synth_s=time.time() - t0, gather_s=0.0, n_chunks=len(keys), local_chunks=len(keys),
                                crossfade="device")
        return wave

    def synthesize(self, text: str, gender: Optional[str] = None, group: Optional[str] = None,
                   area: Optional[str] = None, emotion: Optional[str] = None, sample_iteration: Optional[int] = None,
                   output_path: Optional[str] = None, reference_audio: Optional[str] = None,
                   reference_text: Optional[str] = None) -> Tuple[np.ndarray, float]:
        """-> (int16 waveform, wall seconds)"""
        t0 = time.time()
        ref_audio, ref_text = self.model_session_manager.select_sample(
            gender, group, area, emotion, sample_iteration, reference_audio, reference_text)
        try:
            self.last_timing = {}
            inputs_list = self._prepare_inputs(ref_audio, ref_text, text)
            t1 = time.time()
            final = None
            if self.use_sessions or self.config.fuse_nfe != 1:
                waves = self._synthesize_chunks_sessions(inputs_list)
            else:
                final = self._synthesize_joined(inputs_list)          # batch + clip-fix + cross-fade on the device
                waves = None if final is not None else self._synthesize_chunks_batched(inputs_list)
            t2 = time.time()
            if final is None:
                final = self._fold(waves)
            elapsed = time.time() - t0
            self.last_timing.update(prepare_s=t1 - t0, crossfade_s=time.time() - t2, total_s=elapsed)
            if output_path:
                self.audio_processor.save_audio(final, output_path, self.config.sample_rate)
                logger.info(f"Audio saved to: {output_path}")
            return final, elapsed
        except Exception as exc:
            raise RuntimeError(f"Speech synthesis failed: {str(exc)}")

    def synthesize_to_bytes(self, text: str, **voice) -> Tuple[bytes, float]:
        """-> (WAV file bytes, wall seconds) without the temp-file round trip of the reference's client
        (/root/reference/vietvoicetts/client.py:149-172); same keyword arguments as `synthesize`."""
        voice.pop("output_path", None)
        wave, elapsed = self.synthesize(text, **voice)
        return self.audio_processor.to_wav_bytes(wave, self.config.sample_rate), elapsed

    def validate_configuration(self, reference_audio: Optional[str] = None) -> bool:
        if reference_audio is None:
            return True
        return self.config.validate_with_reference_audio(reference_audio)
