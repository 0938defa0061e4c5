"""Stand-ins for the two audio packages the reference imports at module load — `pydub` (ffmpeg front-end) and
`soundfile` (libsndfile) — for hosts where neither is installable.

The reference touches exactly this much of them (/root/reference/vietvoicetts/core/audio_processor.py):
    AudioSegment.from_file(fh_or_BytesIO).set_channels(1).set_frame_rate(sr).get_array_of_samples()     :16-26
    soundfile.write(path, int16_array, sr, format='WAVEX')                                               :61-67
Only PCM WAV input can be decoded without ffmpeg; anything else raises `CouldntDecodeError` (pydub's own exception
name), which the reference wraps like any other load failure.  `install()` registers the stand-ins in `sys.modules`
ONLY IF the real packages are absent, so a host that has them keeps using them.
"""
from __future__ import annotations

import importlib.util
import sys
import types

import numpy as np

from .host.audio_processor import AudioProcessor, _read_wav, _resample


class CouldntDecodeError(Exception):
    pass


class AudioSegment:
    """The slice of pydub.AudioSegment the reference uses: an int16 PCM buffer with channel / rate conversion."""

    def __init__(self, samples: np.ndarray, frame_rate: int, channels: int):
        self._s = np.asarray(samples, dtype=np.int16).reshape(-1, channels)
        self.frame_rate = int(frame_rate)
        self.channels = int(channels)
        self.sample_width = 2

    @classmethod
    def from_file(cls, file, format=None, **kwargs) -> "AudioSegment":
        try:
            pcm, sr, ch = _read_wav(file if not isinstance(file, (str, bytes)) else open(file, "rb"))
        except RuntimeError as exc:
            raise CouldntDecodeError(str(exc))
        return cls(np.clip(pcm, -32768, 32767).astype(np.int16), sr, ch)

    from_wav = from_file

    def set_channels(self, channels: int) -> "AudioSegment":
        if channels == self.channels:
            return self
        if channels != 1:
            raise ValueError("only down-mixing to mono is supported")
        mono = np.trunc(self._s.astype(np.float32).sum(axis=1) / self.channels)       # audioop.tomono: equal weights
        return AudioSegment(np.clip(mono, -32768, 32767).astype(np.int16), self.frame_rate, 1)

    def set_frame_rate(self, frame_rate: int) -> "AudioSegment":
        if frame_rate == self.frame_rate:
            return self
        if self.channels != 1:
            raise ValueError("resample after set_channels(1)")
        return AudioSegment(_resample(self._s.reshape(-1), self.frame_rate, int(frame_rate)), frame_rate, 1)

    def get_array_of_samples(self):
        import array
        return array.array("h", self._s.reshape(-1).tobytes())

    def __len__(self) -> int:                       # milliseconds, as pydub
        return int(round(1000.0 * self._s.shape[0] / self.frame_rate))


def sf_write(file, data, samplerate, subtype=None, endian=None, format=None, closefd=True) -> None:
    """soundfile.write for what the reference writes: int16 (or float in [-1, 1)) mono, WAVE_FORMAT_EXTENSIBLE."""
    if format not in (None, "WAV", "WAVEX"):
        raise ValueError(f"only WAV / WAVEX can be written without libsndfile, got {format!r}")
    payload = AudioProcessor.to_wav_bytes(np.asarray(data), int(samplerate))
    if hasattr(file, "write"):
        file.write(payload)
    else:
        with open(file, "wb") as fh:
            fh.write(payload)


def install(force: bool = False) -> list:
    """-> names of the stand-ins that were registered (empty when the real packages are importable)."""
    done = []
    if force or ("pydub" not in sys.modules and importlib.util.find_spec("pydub") is None):
        m = types.ModuleType("pydub")
        m.AudioSegment = AudioSegment
        ex = types.ModuleType("pydub.exceptions")
        ex.CouldntDecodeError = CouldntDecodeError
        m.exceptions = ex
        sys.modules["pydub"], sys.modules["pydub.exceptions"] = m, ex
        done.append("pydub")
    if force or ("soundfile" not in sys.modules and importlib.util.find_spec("soundfile") is None):
        m = types.ModuleType("soundfile")
        m.write = sf_write
        sys.modules["soundfile"] = m
        done.append("soundfile")
    return done
