"""What a batch shape costs the first time it is seen (vv_batch creation from the memory pool + CUDA-graph capture of the
loop) against a cached shape: six new shapes of 8 chunks, three calls each.  usage: python tools/newshape_cost.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vietvoice_tts_b200 import artifact
from vietvoice_tts_b200.arch import FULL
from vietvoice_tts_b200.engine import Engine
eng = Engine.from_weights(FULL, artifact.make_random_weights(FULL, 9527))
rng = np.random.default_rng(0)
B = 8
audios = [artifact.synthetic_prompt_pcm(144000, i) for i in range(B)]
ids = [rng.integers(0, FULL.vocab, 200).astype(np.int32) for _ in range(B)]
eng.synthesize_batch(audios, ids, [1400] * B, nfe=32)   # mod table etc.
for k in range(6):
    T = [1000 + 37 * k + 11 * i for i in range(B)]
    t0 = time.perf_counter(); eng.synthesize_batch(audios, ids, T, nfe=32); t1 = time.perf_counter()
    eng.synthesize_batch(audios, ids, T, nfe=32); t2 = time.perf_counter()
    eng.synthesize_batch(audios, ids, T, nfe=32); t3 = time.perf_counter()
    print(f"shape {k}: first {1e3*(t1-t0):7.1f} ms, second {1e3*(t2-t1):7.1f} ms, third {1e3*(t3-t2):7.1f} ms")
