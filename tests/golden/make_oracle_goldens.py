"""Pins the oracle (oracle/graphs.py) against drift: one TINY-architecture chunk, inputs and outputs of the three
graphs.  NOTE: these vectors come from OUR oracle, not from the reference — the reference's graphs/executor are not
runnable offline (oracle/graphs.py header: PARITY UNPINNED).  Run in the build container:

    python tests/golden/make_oracle_goldens.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from vietvoice_tts_b200 import artifact
from vietvoice_tts_b200.arch import TINY
from oracle.graphs import OracleSessions


def main():
    import torch
    torch.set_num_threads(1)
    W = artifact.make_random_weights(TINY, 9527)
    S = OracleSessions(TINY, W)
    rng = np.random.default_rng(123)
    n_samples, n_ids = 12000, 24
    T = n_samples // 256 + 1 + 50
    audio = artifact.synthetic_prompt_pcm(n_samples, 77)
    ids = rng.integers(0, TINY.vocab, size=(1, n_ids)).astype(np.int32)
    noise = rng.standard_normal((1, T, TINY.n_mel)).astype(np.float32)
    wave, x, steps, pre = S.synthesize_chunk(audio.reshape(1, 1, -1), ids, np.array([T], dtype=np.int64), noise, True)
    np.savez_compressed(os.path.join(HERE, "oracle_tiny.npz"), audio=audio, ids=ids, T=np.array([T]), noise=noise,
                        cat_mel_text=pre[5].astype(np.float16), cat_mel_text_drop=pre[6].astype(np.float16),
                        ref_signal_len=pre[7], step1=steps[0].astype(np.float32), final=x.astype(np.float32),
                        wave=wave)
    print("T", T, "wave", wave.shape, "final std", float(x.std()))


if __name__ == "__main__":
    main()
