"""B200-native synthesis engine for the VietVoice-TTS hot path (preprocess -> DiT x (NFE-1) -> decode).

Layout:
  arch.py      ArchConfig (one struct for oracle and engine)
  artifact.py  weight blob + model tarball in the reference's layout
  _lib.py      ctypes binding of the C-ABI library (csrc/ -> libvvb200.so); fails loudly if absent
  engine.py    Engine / Batch: the host-side handle over the C-ABI
  ort_shim.py  onnxruntime-shaped module so the reference's own Python runs on this engine
  host/        host-side mirror of the reference interface (TTSEngine, ModelSessionManager, ...)
"""
__version__ = "0.1.0"
