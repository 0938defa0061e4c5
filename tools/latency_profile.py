"""Where a single utterance's latency goes: per-kernel-class times of one eager DiT evaluation (event pair around every
launch) against the per-step time of the graph loop, for B = 1, 2, 4 utterances of the bench shape (T = 1501).
usage: python tools/latency_profile.py [B ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from vietvoice_tts_b200 import artifact
from vietvoice_tts_b200.arch import FULL
from vietvoice_tts_b200.engine import Engine

Bs = [int(a) for a in sys.argv[1:]] or [1, 2, 4]
T, nfe = 1501, 32
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
eng = Engine.from_weights(FULL, artifact.make_random_weights(FULL, 9527), stream=st.cuda_stream)
names = ["qkv", "out", "ff1", "ff2", "attn", "ln", "conv", "other"]
for B in Bs:
    rng = np.random.default_rng(B)
    batch = eng.batch([T] * B)
    for i in range(B):
        batch.preprocess(i, artifact.synthetic_prompt_pcm(144000, i), rng.integers(0, FULL.vocab, 270).astype(np.int32),
                         None, chunk_key=i)
    for _ in range(3):
        batch.run_resident(nfe)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        batch.run_resident(nfe)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    for _ in range(2):
        cls = batch.profile_step(step=1, nfe=nfe)
    print(f"B={B}: whole path {ms:7.2f} ms  = {ms / (nfe - 1):6.3f} ms per step incl. pre/decode; eager step sum "
          f"{sum(cls):6.3f} ms: " + ", ".join(f"{n} {v:.3f}" for n, v in zip(names, cls)))
    batch.close()
