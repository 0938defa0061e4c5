"""Parity AT THE CONFIGURATION THE BENCHMARK RUNS (BASELINE configs[1]): FULL architecture (dim 1024, 22 layers,
16 heads), T = 1501 frames, a batch of B = 2 utterances (M = 6 068 rows >= 1024, so every DiT GEMM runs on the CTA-pair
kernel and attention on the 16-head T = 1501 shape), ALL 31 Euler steps of utterance 0 against the fp32 oracle.

What the reference does here: 31 `transformer` session calls per chunk, state handed from call to call
(/root/reference/vietvoicetts/core/tts_engine.py:157-172), then `decode` (:176-187).  The north star asks for "the mel
after each NFE step within a stated BF16 relative tolerance, with final waveform SNR and mel-L1 reported": the
tolerances are written below, every step is asserted, the report goes to gpurun_out/parity_full.json.

Two GPU paths are checked: the step-by-step path (one vv_sample call per step, what a `transformer` session call
maps to) and the CUDA-graph replay of the whole loop through vv_synthesize_batch (what bench.py times).
"""
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from vietvoice_tts_b200 import artifact
from vietvoice_tts_b200.arch import FULL
from vietvoice_tts_b200.engine import Engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# ---- stated tolerances (bf16 GEMM operands, fp32 residual stream / Euler state; oracle fp32 throughout)
TOL_STEP1 = 3e-3          # rel-L2 of the state after ONE step from identical input
TOL_STEP = 2e-2           # rel-L2 of the state after every one of the 31 steps
TOL_INCREMENT = 6e-2      # rel-L2 of the accumulated update (x_k - y0): the part the network actually produced
TOL_MEL_L1 = 5e-2         # mean |mel_gpu - mel_oracle| over the target frames of the final mel (log-mel units)
MIN_SNR_DB = 25.0         # final int16 waveform vs the oracle's


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def snr_db(x, ref):
    x, ref = np.asarray(x, np.float64).reshape(-1), np.asarray(ref, np.float64).reshape(-1)
    return float(10 * np.log10(np.sum(ref ** 2) / (np.sum((x - ref) ** 2) + 1e-30)))


def test_full_size_all_steps_vs_oracle():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle.graphs import OracleSessions
    torch.set_num_threads(os.cpu_count() or 1)
    arch = FULL
    W = artifact.make_random_weights(arch, 9527)
    T, n_samples, n_ids, B = 1501, 144000, 270, 2
    rng = np.random.default_rng(1501)
    audios = [artifact.synthetic_prompt_pcm(n_samples, 70 + i) for i in range(B)]
    ids = [rng.integers(0, arch.vocab, size=n_ids).astype(np.int32) for _ in range(B)]
    noises = [rng.standard_normal((T, arch.n_mel)).astype(np.float32) for _ in range(B)]

    # ---- oracle: the reference's driver loop on utterance 0 (about half a minute of CPU at T = 1501)
    ora = OracleSessions(arch, W)
    with torch.no_grad():
        wave_ref, x_ref, steps_ref, pre = ora.synthesize_chunk(audios[0].reshape(1, 1, -1), ids[0][None],
                                                               np.array([T], dtype=np.int64), noises[0][None],
                                                               collect_steps=True)
    assert len(steps_ref) == arch.nfe - 1 == 31
    ref_len = int(pre[7][0])

    eng = Engine.from_weights(arch, W)
    del W
    # ---- path 1: one engine call per step (a `transformer` session call each)
    b = eng.batch([T] * B)
    assert b.T == [T, T]
    for i in range(B):
        assert b.preprocess(i, audios[i], ids[i], noises[i]) == ref_len
    assert rel(b.get(0, "cat_mel_text"), pre[5][0]) < 1e-2
    assert rel(b.get(0, "cat_mel_text_drop"), pre[6][0]) < 1e-2
    per_step, per_incr = [], []
    for s in range(arch.nfe - 1):
        b.sample(first_step=s, n_steps=1)
        x = b.get(0, "noise")
        per_step.append(rel(x, steps_ref[s][0]))
        per_incr.append(rel(x - noises[0], steps_ref[s][0] - noises[0]))
    mel_gpu = b.get(0, "noise")
    mel_l1 = float(np.mean(np.abs(mel_gpu[ref_len:] - x_ref[0][ref_len:])))
    wave_steps = b.decode(0)
    snr_steps = snr_db(wave_steps, wave_ref)
    b.close()

    # ---- path 2: the whole loop as one CUDA-graph replay through the host-buffer call bench.py's e2e times
    out = eng.synthesize_batch(audios, ids, [T] * B, noises=noises, nfe=arch.nfe)
    snr_graph = snr_db(out[0], wave_ref)
    snr_paths = snr_db(out[0], wave_steps)
    eng.close()

    report = {
        "config": {"arch": "FULL", "T": T, "B": B, "M_rows": 2 * ((B * (T + 16) + 7) // 8 * 8), "nfe": arch.nfe,
                   "gemm_path": "gemm_pair_kernel (M >= 1024)", "checked_vs": "oracle (PyTorch CPU fp32, parity unpinned)"},
        "tolerances": {"step1_rel_l2": TOL_STEP1, "every_step_rel_l2": TOL_STEP, "increment_rel_l2": TOL_INCREMENT,
                       "mel_l1": TOL_MEL_L1, "min_snr_db": MIN_SNR_DB},
        "rel_l2_per_step": per_step, "rel_l2_increment_per_step": per_incr,
        "final": {"rel_l2": per_step[-1], "mel_l1_target_frames": mel_l1, "snr_db_stepwise": snr_steps,
                  "snr_db_graph_loop": snr_graph, "snr_db_graph_vs_stepwise": snr_paths},
    }
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_full.json"), "w") as f:
        json.dump(report, f, indent=1)
    print("parity FULL T=1501 B=2:", json.dumps(report["final"]),
          "max step rel-L2 %.2e" % max(per_step), "step1 %.2e" % per_step[0])

    assert per_step[0] < TOL_STEP1, per_step[0]
    for s, r in enumerate(per_step):
        assert r < TOL_STEP, (s, r)
    for s, r in enumerate(per_incr):
        assert r < TOL_INCREMENT, (s, r)
    assert mel_l1 < TOL_MEL_L1, mel_l1
    assert snr_steps > MIN_SNR_DB, snr_steps
    assert snr_graph > MIN_SNR_DB, snr_graph
