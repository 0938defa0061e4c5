"""Host-side audio helpers either side of the engine: prompt loading/normalisation before `preprocess`, and
clip-fix / cross-fade / WAV writing after `decode`.

Mirror of the reference interface `AudioProcessor`
(/root/reference/vietvoicetts/core/audio_processor.py: load_audio :16-26, normalize_to_int16 :29-44,
fix_clipped_audio :47-58, save_audio :61-67, concatenate_with_crossfade :70-120, ..._improved :123-193).
Same names, argument meaning and integer results (every float->int16 cast truncates toward zero); written
independently.  Parity is pinned by tests/golden/host_audio.npz, produced by running the reference's own static
methods (tests/golden/make_host_goldens.py).

Difference that is deliberate and documented ("out of scope"): the reference decodes any container
through pydub -> ffmpeg; neither exists offline, so `load_audio` reads RIFF/WAVE PCM itself and raises for
anything else.
"""
from __future__ import annotations

import io
import struct
import wave
from pathlib import Path
from typing import List, Union

import numpy as np


def _read_wav(fh) -> tuple[np.ndarray, int, int]:
    """-> (samples [n, channels] float32 on the int16 scale, sample_rate, channels)"""
    try:
        with wave.open(fh, "rb") as w:
            ch, width, sr, n = w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()
            raw = w.readframes(n)
    except wave.Error as exc:
        raise RuntimeError(f"unsupported audio container (only PCM WAV can be decoded without ffmpeg): {exc}")
    if width == 2:
        pcm = np.frombuffer(raw, dtype="<i2").astype(np.float32)
    elif width == 1:
        pcm = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) * 256.0
    elif width == 4:
        pcm = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 65536.0
    elif width == 3:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        v = np.where(v >= 1 << 23, v - (1 << 24), v)
        pcm = v.astype(np.float32) / 256.0
    else:
        raise RuntimeError(f"unsupported WAV sample width {width}")
    return pcm.reshape(-1, ch), sr, ch


def _resample(mono_i16: np.ndarray, sr_in: int, sr_out: int) -> np.ndarray:
    if sr_in == sr_out:
        return mono_i16
    try:  # pydub's set_frame_rate is audioop.ratecv; use it when the interpreter still ships audioop
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", DeprecationWarning)
            import audioop
        out, _ = audioop.ratecv(mono_i16.astype("<i2").tobytes(), 2, 1, sr_in, sr_out, None)
        return np.frombuffer(out, dtype="<i2")
    except ImportError:
        from math import gcd
        from scipy.signal import resample_poly
        g = gcd(sr_in, sr_out)
        y = resample_poly(mono_i16.astype(np.float64), sr_out // g, sr_in // g)
        return np.clip(np.rint(y), -32768, 32767).astype(np.int16)


class AudioProcessor:
    """Handles audio processing operations"""

    @staticmethod
    def load_audio(path_or_bytes: Union[str, bytes], sample_rate: int) -> np.ndarray:
        if isinstance(path_or_bytes, str):
            if not Path(path_or_bytes).exists():
                raise FileNotFoundError(f"Audio file not found: {path_or_bytes}")
            with open(path_or_bytes, "rb") as fh:
                pcm, sr, ch = _read_wav(fh)
        else:
            pcm, sr, ch = _read_wav(io.BytesIO(path_or_bytes))
        if ch > 1:      # pydub/audioop.tomono: equal-weight mix, truncated to int16
            pcm = np.trunc(pcm.sum(axis=1) / ch)
        mono = np.clip(pcm.reshape(-1), -32768, 32767).astype(np.int16)
        mono = _resample(mono, sr, sample_rate)
        return AudioProcessor.normalize_to_int16(mono.astype(np.float32))

    @staticmethod
    def normalize_to_int16(audio: np.ndarray) -> np.ndarray:
        """remove DC, scale the peak to 29 491 (90 % of full scale), truncate to int16"""
        centred = audio - np.mean(audio)
        peak = np.max(np.abs(centred))
        if peak > 0:
            centred = centred * (29491.0 / peak)
        return centred.astype(np.int16)

    @staticmethod
    def fix_clipped_audio(audio: np.ndarray) -> np.ndarray:
        """scrub NaN/Inf; if the peak touches full scale, rescale to 26 214 (80 %)"""
        audio = np.nan_to_num(audio, nan=0.0, posinf=0.0, neginf=0.0)
        peak = np.max(np.abs(audio))
        if peak >= 32767:
            return (audio * (26214.0 / peak)).astype(np.int16)
        return audio

    @staticmethod
    def to_wav_bytes(audio: np.ndarray, sample_rate: int) -> bytes:
        """The bytes `save_audio` would put in the file, built in memory: the reference serves
        `synthesize_to_bytes` by writing a temp file and reading it back
        (/root/reference/vietvoicetts/client.py:149-172)."""
        if audio.size == 0:
            raise ValueError("Cannot save empty audio.")
        flat = audio.reshape(-1)
        if flat.dtype != np.int16:
            if np.issubdtype(flat.dtype, np.floating):   # soundfile maps float [-1,1) to int16 full scale
                flat = np.clip(flat * 32768.0, -32768, 32767).astype(np.int16)
            else:
                flat = flat.astype(np.int16)
        data = flat.astype("<i2").tobytes()
        guid_pcm = b"\x01\x00\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71"
        fmt = struct.pack("<HHIIHHHHI", 0xFFFE, 1, sample_rate, sample_rate * 2, 2, 16, 22, 16, 0x4) + guid_pcm
        fact = struct.pack("<I", len(flat))
        body = (b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"fact" + struct.pack("<I", 4) + fact +
                b"data" + struct.pack("<I", len(data)) + data)
        return b"RIFF" + struct.pack("<I", len(body)) + body

    @staticmethod
    def save_audio(audio: np.ndarray, file_path: str, sample_rate: int) -> None:
        """WAVE_FORMAT_EXTENSIBLE ('WAVEX') file, as soundfile.write(..., format='WAVEX') produces for int16"""
        payload = AudioProcessor.to_wav_bytes(audio, sample_rate)
        Path(file_path).parent.mkdir(parents=True, exist_ok=True)
        with open(file_path, "wb") as fh:
            fh.write(payload)

    @staticmethod
    def concatenate_with_crossfade(generated_waves: List[np.ndarray], cross_fade_duration: float,
                                   sample_rate: int) -> np.ndarray:
        """linear cross-fade, no level matching (float64 result once a fade has been applied)"""
        if not generated_waves:
            return np.array([])
        if len(generated_waves) == 1:
            return generated_waves[0].reshape(-1)
        waves = [w.reshape(-1) for w in generated_waves]
        if cross_fade_duration <= 0:
            return np.concatenate(waves)
        acc = waves[0]
        for nxt in waves[1:]:
            n = min(int(cross_fade_duration * sample_rate), len(acc), len(nxt))
            if n <= 0:
                acc = np.concatenate([acc, nxt])
                continue
            down = np.linspace(1, 0, n)
            up = np.linspace(0, 1, n)
            seam = acc[-n:] * down + nxt[:n] * up
            acc = np.concatenate([acc[:-n], seam, nxt[n:]])
        return acc

    @staticmethod
    def concatenate_with_crossfade_improved(generated_waves: List[np.ndarray], cross_fade_duration: float,
                                            sample_rate: int) -> np.ndarray:
        """clip-fix each chunk, RMS-match the incoming chunk (ratio clipped to [0.7, 1.5], only when both seam
        RMS values exceed 100), cos^2/sin^2 fade; left fold, so the result depends on chunk order"""
        if not generated_waves:
            return np.array([])
        if len(generated_waves) == 1:
            return generated_waves[0].reshape(-1)
        waves = [AudioProcessor.fix_clipped_audio(w.reshape(-1)) for w in generated_waves]
        if cross_fade_duration <= 0:
            return np.concatenate(waves)
        acc = waves[0]
        for nxt in waves[1:]:
            n = min(int(cross_fade_duration * sample_rate), len(acc), len(nxt))
            if n <= 0:
                acc = np.concatenate([acc, nxt])
                continue
            tail = acc[-n:]
            head = nxt[:n]
            rms_tail = np.sqrt(np.mean(tail.astype(np.float32) ** 2))
            rms_head = np.sqrt(np.mean(head.astype(np.float32) ** 2))
            if rms_tail > 100 and rms_head > 100:
                ratio = np.clip(rms_tail / rms_head, 0.7, 1.5)
                nxt = (nxt.astype(np.float32) * ratio).astype(np.int16)
                head = nxt[:n]
            theta = np.linspace(0, np.pi / 2, n)
            seam = (tail.astype(np.float32) * np.cos(theta) ** 2 +
                    head.astype(np.float32) * np.sin(theta) ** 2).astype(np.int16)
            acc = np.concatenate([acc[:-n], seam, nxt[n:]])
        return acc
