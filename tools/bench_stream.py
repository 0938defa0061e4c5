"""BASELINE config 5 as a measurement: a Poisson stream of API-style requests (voices swept over the metadata table,
NFE 16 / 32 / 64, texts of one to a few sentences) through the request scheduler on ONE GPU; with torchrun every
rank runs the scheduler for the requests it owns (request i -> rank i % world, no collective).

Prints one JSON line: aggregate audio-seconds per second over the makespan, p50 / p95 request latency
(submit -> waveform), micro-batch statistics.  Random-init weights of the FULL architecture, synthetic 6 s prompts.

usage: python tools/bench_stream.py [--requests 48] [--rate 12] [--seed 0]
"""
import argparse
import json
import os
import statistics
import sys
import tempfile
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from vietvoice_tts_b200 import artifact
from vietvoice_tts_b200.arch import FULL

SENTENCES = [
    "Xin chào Việt Nam.",
    "Hôm nay trời đẹp quá, chúng ta cùng đi dạo quanh hồ nhé!",
    "Tôi là trợ lý ảo, tôi có thể đọc văn bản tiếng Việt với nhiều giọng khác nhau.",
    "Một hai ba bốn năm sáu bảy tám chín mười, mười một mười hai mười ba mười bốn mười lăm.",
    "Cảm ơn bạn rất nhiều, hẹn gặp lại bạn vào một ngày gần nhất có thể.",
    "Thành phố Hồ Chí Minh là trung tâm kinh tế lớn nhất của cả nước, với hơn chín triệu dân.",
    "Bạn cần giúp gì không?",
    "Mùa thu Hà Nội có hương hoa sữa nồng nàn trên từng con phố nhỏ, và những cơn gió heo may se lạnh.",
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--requests", type=int, default=48)
    ap.add_argument("--rate", type=float, default=12.0, help="mean arrivals per second (Poisson), whole job")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--max-batch-chunks", type=int, default=8)
    ap.add_argument("--max-batch-frames", type=int, default=8 * 1800)
    args = ap.parse_args()

    import torch
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    from vietvoice_tts_b200.host.model_config import ModelConfig
    from vietvoice_tts_b200.host.scheduler import RequestScheduler
    from vietvoice_tts_b200.host.tts_engine import TTSEngine

    d = tempfile.mkdtemp(prefix=f"vvb200_stream_{rank}_")
    voices = [{"gender": g, "group": grp, "area": a, "emotion": e}
              for g, grp, a, e in (("female", "audiobook", "northern", "neutral"), ("male", "news", "southern", "serious"),
                                   ("female", "story", "central", "happy"), ("male", "interview", "northern", "neutral"),
                                   ("female", "news", "southern", "sad"), ("male", "story", "central", "angry"))]
    artifact.build_model_tar(os.path.join(d, "model-bin.pt"), FULL, seed=9527, voices=voices, prompt_seconds=6.0)
    cfg = ModelConfig(model_cache_dir=d, nfe_step=32)

    rng = np.random.default_rng(args.seed)                 # the same stream on every rank
    reqs = []
    t = 0.0
    for i in range(args.requests):
        t += rng.exponential(1.0 / args.rate)
        n_sent = int(rng.integers(1, 4))
        text = " ".join(SENTENCES[int(k)] for k in rng.integers(0, len(SENTENCES), n_sent))
        v = voices[int(rng.integers(0, len(voices)))]
        nfe = int(rng.choice([16, 32, 64], p=[0.25, 0.5, 0.25]))
        reqs.append((t, text, v, nfe))

    with TTSEngine(cfg) as tts:
        # warm-up: one request per NFE so that the modulation tables exist (model load is not part of the metric)
        with RequestScheduler(tts, max_batch_chunks=args.max_batch_chunks) as sch:
            for nfe in (16, 32, 64):
                sch.submit(SENTENCES[0], nfe=nfe).result(timeout=600)
        lat, audio_s, errs = [], [], []
        with RequestScheduler(tts, max_batch_chunks=args.max_batch_chunks, max_batch_frames=args.max_batch_frames,
                              rank=rank, world=world) as sch:
            lock = threading.Lock()
            t_start = time.time()

            def client(i):
                at, text, v, nfe = reqs[i]
                delay = t_start + at - time.time()
                if delay > 0:
                    time.sleep(delay)
                t0 = time.time()
                try:
                    res = sch.submit(text, gender=v["gender"], area=v["area"], emotion=v["emotion"], nfe=nfe).result(600)
                except Exception as ex:
                    with lock:
                        errs.append(repr(ex))
                    return
                if res is None:                             # another rank's request
                    return
                wave, _ = res
                with lock:
                    lat.append(time.time() - t0)
                    audio_s.append(wave.shape[0] / 24000.0)

            th = [threading.Thread(target=client, args=(i,)) for i in range(len(reqs))]
            [x.start() for x in th]
            [x.join() for x in th]
            makespan = time.time() - t_start
            batches, chunks = sch.batches_run, sch.chunks_run
    stats = [sum(audio_s), makespan, len(lat), batches, chunks]
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")
        allv = [None] * world
        dist.all_gather_object(allv, (stats, lat, errs))
        stats = [sum(a[0][0] for a in allv), max(a[0][1] for a in allv), sum(a[0][2] for a in allv),
                 sum(a[0][3] for a in allv), sum(a[0][4] for a in allv)]
        lat = [x for a in allv for x in a[1]]
        errs = [x for a in allv for x in a[2]]
    if rank == 0:
        lat.sort()
        line = {
            "metric": "request stream: synth audio-sec/sec over the makespan; request latency", "n_gpus": world,
            "value": stats[0] / stats[1], "unit": "audio-s/s", "audio_s_total": stats[0], "makespan_s": stats[1],
            "requests": stats[2], "errors": errs[:3], "arrival_rate_per_s": args.rate,
            "latency_p50_ms": 1e3 * statistics.median(lat) if lat else None,
            "latency_p95_ms": 1e3 * lat[min(len(lat) - 1, int(0.95 * len(lat)))] if lat else None,
            "micro_batches": stats[3], "chunks": stats[4],
            "config": {"workload": "configs[4]: Poisson request stream, 6 voices, NFE 16/32/64 (p 0.25/0.5/0.25), 1-3 "
                                   "sentences per request", "max_batch_chunks": args.max_batch_chunks, "max_batch_frames": args.max_batch_frames,
                       "weights": "random-init F5-TTS-Base/Vocos shapes, seed 9527"},
        }
        print(json.dumps(line))


if __name__ == "__main__":
    main()
