"""BASELINE config 5 as a measurement: a Poisson stream of API-style requests (voices swept over the metadata table,
NFE 16 / 32 / 64, texts of one to a few sentences) through the request scheduler on ONE GPU; with torchrun every
rank runs the scheduler for the requests it owns (owner = the rank with the least estimated work at the request's
arrival, computed identically on every rank; no collective).

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


VOICES = [{"gender": g, "group": grp, "area": a, "emotion": e}
          for g, grp, a, e in (("female", "audiobook", "northern", "neutral"), ("male", "news", "southern", "serious"),
                               ("female", "story", "central", "happy"), ("male", "interview", "northern", "neutral"),
                               ("female", "news", "southern", "sad"), ("male", "story", "central", "angry"))]
NFE_CHOICES, NFE_P = [16, 32, 64], [0.25, 0.5, 0.25]


def make_requests(n, rate, seed):
    """The same Poisson stream on every rank: (arrival time, text, voice, nfe) per request."""
    rng = np.random.default_rng(seed)
    reqs, t = [], 0.0
    for _ in range(n):
        t += rng.exponential(1.0 / rate)
        n_sent = int(rng.integers(1, 4))
        text = " ".join(SENTENCES[int(k)] for k in rng.integers(0, len(SENTENCES), n_sent))
        v = VOICES[int(rng.integers(0, len(VOICES)))]
        reqs.append((t, text, v, int(rng.choice(NFE_CHOICES, p=NFE_P))))
    return reqs


def request_owners(reqs, world):
    """Front-end balancing: estimated work of a request = characters of text x Euler steps; requests go, in arrival
    order, to the rank with the least work accepted so far (shard.dispatch_requests) — the same on every rank."""
    from vietvoice_tts_b200.shard import dispatch_requests
    return dispatch_requests([len(text) * (nfe - 1) for _, text, _, nfe in reqs], world)


def run_stream(tts, reqs, rank=0, world=1, max_batch_chunks=8, max_batch_frames=8 * 1800, warm=True):
    """Plays `reqs` through a RequestScheduler on this rank (owner of request i: request_owners(); VVB200_STREAM_RR=1
    falls back to i % world).
    -> dict(audio_s, work_s, makespan_s, lat, errs, batches, chunks); work_s = audio seconds weighted by
    (nfe - 1) / 31, i.e. in units of NFE-32 work (an NFE-64 second costs 63/31 of an NFE-32 second)."""
    from vietvoice_tts_b200.host.scheduler import RequestScheduler
    if warm:     # one request per NFE so that the modulation tables exist (model load is not part of the metric)
        with RequestScheduler(tts, max_batch_chunks=max_batch_chunks) as sch:
            for nfe in NFE_CHOICES:
                sch.submit(SENTENCES[0], nfe=nfe).result(timeout=600)
    lat, audio_s, work_s, errs = [], [], [], []
    owners = [i % world for i in range(len(reqs))] if os.environ.get("VVB200_STREAM_RR") == "1" else request_owners(reqs, world)
    with RequestScheduler(tts, max_batch_chunks=max_batch_chunks, max_batch_frames=max_batch_frames,
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
                res = sch.submit(text, gender=v["gender"], group=v["group"], area=v["area"], emotion=v["emotion"],
                                 nfe=nfe, owner=owners[i]).result(600)
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
                work_s.append(wave.shape[0] / 24000.0 * (nfe - 1) / 31.0)

        th = [threading.Thread(target=client, args=(i,)) for i in range(len(reqs))]
        [x.start() for x in th]
        [x.join() for x in th]
        makespan = time.time() - t_start
        batches, chunks = sch.batches_run, sch.chunks_run
    return {"audio_s": sum(audio_s), "work_s": sum(work_s), "makespan_s": makespan, "lat": lat, "errs": errs,
            "batches": batches, "chunks": chunks, "requests": len(lat)}


def merge_stream_stats(parts):
    """per-rank run_stream results -> whole-job summary (makespan = the slowest rank's)"""
    lat = sorted(x for p in parts for x in p["lat"])
    audio, work = sum(p["audio_s"] for p in parts), sum(p["work_s"] for p in parts)
    span = max(p["makespan_s"] for p in parts)
    return {
        "value": audio / span, "unit": "audio-s/s", "nfe32_equivalent_audio_s_per_s": work / span,
        "audio_s_total": audio, "makespan_s": span, "requests": sum(p["requests"] for p in parts),
        "errors": [e for p in parts for e in p["errs"]][:3],
        "latency_p50_ms": 1e3 * statistics.median(lat) if lat else None,
        "latency_p95_ms": 1e3 * lat[min(len(lat) - 1, int(0.95 * len(lat)))] if lat else None,
        "micro_batches": sum(p["batches"] for p in parts), "chunks": sum(p["chunks"] for p in parts),
    }


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
    from vietvoice_tts_b200.host.tts_engine import TTSEngine

    d = tempfile.mkdtemp(prefix=f"vvb200_stream_{rank}_")
    artifact.build_model_tar(os.path.join(d, "model-bin.pt"), FULL, seed=9527, voices=VOICES, prompt_seconds=6.0)
    cfg = ModelConfig(model_cache_dir=d, nfe_step=32)
    reqs = make_requests(args.requests, args.rate, args.seed)
    with TTSEngine(cfg) as tts:
        part = run_stream(tts, reqs, rank, world, args.max_batch_chunks, args.max_batch_frames)
    parts = [part]
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")
        parts = [None] * world
        dist.all_gather_object(parts, part)
    if rank == 0:
        line = {"metric": "request stream: synth audio-sec/sec over the makespan; request latency", "n_gpus": world,
                **merge_stream_stats(parts), "arrival_rate_per_s": args.rate,
                "config": {"workload": "configs[4]: Poisson request stream, 6 voices, NFE 16/32/64 (p 0.25/0.5/0.25), 1-3 "
                                       "sentences per request", "max_batch_chunks": args.max_batch_chunks,
                           "max_batch_frames": args.max_batch_frames,
                           "weights": "random-init F5-TTS-Base/Vocos shapes, seed 9527"}}
        print(json.dumps(line))


if __name__ == "__main__":
    main()
