"""Event-times the three stages of the resident path at the bench shape: preprocess+conditioning, the CUDA-graph
sampling loop, Vocos/iSTFT decode.  usage: python tools/stage_times.py [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from vietvoice_tts_b200 import artifact
from vietvoice_tts_b200.arch import FULL as arch
from vietvoice_tts_b200.engine import Engine
import bench

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B, nfe = 8, 32
t_ref, t_tgt, T, audio_s = bench.workload_dims(arch)
W = artifact.make_random_weights(arch, 9527)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
eng = Engine.from_weights(arch, W, device=0, stream=st.cuda_stream)
audios, ids = bench.make_inputs(arch, B, T, 0)
batch = eng.batch([T] * B)
for i in range(B):
    batch.preprocess(i, audios[i], ids[i], None, seed=9527, chunk_key=i)
for _ in range(2):
    batch.run_resident(nfe)
torch.cuda.synchronize()


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


tot, loop = [], []
for _ in range(reps):
    e0 = ev(); batch.run_resident(nfe); e1 = ev()
    torch.cuda.synchronize()
    tot.append(e0.elapsed_time(e1))
    e0 = ev(); batch.sample(nfe); e1 = ev()
    torch.cuda.synchronize()
    loop.append(e0.elapsed_time(e1))
print(f"run_resident (pre + loop + decode): {np.median(tot):8.2f} ms")
print(f"sample (graph loop only)          : {np.median(loop):8.2f} ms   -> pre + decode = {np.median(tot) - np.median(loop):.2f} ms")
