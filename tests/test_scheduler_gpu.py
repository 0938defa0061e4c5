"""BASELINE config 5 on the B200: a concurrent request stream sweeping voices, speed and NFE through the request
scheduler.  Every request must come back exactly as if it had been synthesised alone by
`TTSEngine.synthesize` with that nfe / speed / seed — whatever else shared its micro-batches."""
import threading

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from vietvoice_tts_b200 import artifact
from vietvoice_tts_b200.arch import TINY


def snr_db(x, ref):
    x, ref = np.asarray(x, np.float64).reshape(-1), np.asarray(ref, np.float64).reshape(-1)
    return 10 * np.log10(np.sum(ref ** 2) / (np.sum((x - ref) ** 2) + 1e-30))


@pytest.fixture(scope="module")
def model_dir(tmp_path_factory):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    d = tmp_path_factory.mktemp("models")
    artifact.build_model_tar(str(d / "model-bin.pt"), TINY, seed=9527, prompt_seconds=1.5)
    return str(d)


def test_request_stream_matches_solo_synthesis(model_dir):
    from vietvoice_tts_b200.host.model_config import ModelConfig
    from vietvoice_tts_b200.host.scheduler import RequestScheduler
    from vietvoice_tts_b200.host.tts_engine import TTSEngine

    texts = ["Xin chào Việt Nam.",
             "Hôm nay trời đẹp quá, tôi đi học!",
             " ".join(["Đây là một câu khá dài để kiểm tra việc chia đoạn văn bản thành nhiều phần nhỏ hơn."] * 10),
             "Tôi là trợ lý ảo. Bạn cần giúp gì không?",
             "Một hai ba bốn năm sáu bảy tám chín mười.",
             "Cảm ơn bạn rất nhiều, hẹn gặp lại!"]
    nfes = [4, 8, 4, 6, 8, 4]
    speeds = [None, 1.2, None, 0.8, None, None]
    genders = ["female", "male", None, "female", None, "male"]
    cfg = ModelConfig(model_cache_dir=model_dir, nfe_step=TINY.nfe)
    with TTSEngine(cfg) as tts:
        got = {}
        with RequestScheduler(tts, max_batch_chunks=8, max_wait_s=0.25) as sch:
            def client(i):
                got[i] = sch.submit(texts[i], gender=genders[i], nfe=nfes[i], speed=speeds[i]).result(timeout=300)
            threads = [threading.Thread(target=client, args=(i,)) for i in range(len(texts))]
            [t.start() for t in threads]
            [t.join() for t in threads]
            assert sch.chunks_run > len(texts)            # the long text was chunked
            assert sch.batches_run < sch.chunks_run       # and chunks shared micro-batches
        # solo reference: the plain TTSEngine with the config set up for that one request
        for i, text in enumerate(texts):
            solo_cfg = ModelConfig(model_cache_dir=model_dir, nfe_step=nfes[i],
                                   speed=cfg.speed if speeds[i] is None else speeds[i])
            tts.config = solo_cfg
            want, _ = tts.synthesize(text, gender=genders[i])
            tts.config = cfg
            wave, secs = got[i]
            assert wave.dtype == np.int16 and wave.shape == want.shape, (i, wave.shape, want.shape)
            assert snr_db(wave, want) > 25.0, (i, snr_db(wave, want))
            assert secs > 0
