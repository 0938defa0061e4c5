"""Driver for ncu captures of the PREPROCESS and DECODE kernels at the bench shape (8 utterances x T 1501, FULL
architecture): one resident pass with a 2-point time grid (a single DiT evaluation), so that the launch list is
dominated by the mel front-end, the text ConvNeXt-V2 blocks, the Vocos backbone and the fused iSTFT/overlap-add.
usage: python tools/prof_stages.py [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from vietvoice_tts_b200 import artifact
from vietvoice_tts_b200.arch import FULL as arch
from vietvoice_tts_b200.engine import Engine

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
t_ref, t_tgt, T, audio_s = bench.workload_dims(arch)
eng = Engine.from_weights(arch, artifact.make_random_weights(arch, 9527))
audios, ids = bench.make_inputs(arch, bench.WORK["B"], T, 0)
b = eng.batch([T] * bench.WORK["B"])
for i in range(bench.WORK["B"]):
    b.preprocess(i, audios[i], ids[i], None, seed=9527, chunk_key=i)
for _ in range(reps):
    st = b.profile_stages(2)
print("stages (nfe=2):", st)
pcm = b.decode(0)
assert pcm.shape[0] == (t_tgt - 1) * arch.hop and np.abs(pcm.astype(np.int32)).max() > 0
b.close()
eng.close()
