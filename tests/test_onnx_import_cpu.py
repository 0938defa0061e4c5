"""ONNX-initialiser importer: wire-format reader, name/edge mapping, tar rewrite.

The upstream checkpoint is unreachable offline, so the fixtures are ONNX containers written by the module's own
encoder from seeded weights under upstream (F5-TTS / Vocos) parameter names — including torch.onnx's habit of storing
Linear weights transposed under anonymous `onnx::MatMul_<n>` names.  The tar layout is the one
/root/reference/vietvoicetts/core/model.py:73-129 reads.
"""
import io
import json
import tarfile

import numpy as np
import pytest

from vietvoice_tts_b200 import artifact, onnx_import as oi
from vietvoice_tts_b200.arch import TINY, FULL


def _weights():
    return artifact.make_random_weights(TINY, 9527)


def test_wire_format_roundtrip_all_encodings():
    rng = np.random.default_rng(0)
    init = {
        "a.raw": rng.standard_normal((3, 5)).astype(np.float32),
        "b.i64": np.array([-1, 0, 7, 2 ** 40], dtype=np.int64),
        "c.f16": rng.standard_normal((4,)).astype(np.float16),
        "d.scalar": np.float32(2.5).reshape(()),
    }
    for raw in (True, False):      # raw_data vs typed repeated fields (float_data / int64_data)
        g = oi.parse_model(oi.write_model(init, [oi.OnnxNode("MatMul", ["x", "a.raw"], ["y"], "mm")],
                                          inputs=["x", "a.raw"], outputs=["y"], raw=raw))
        assert list(g.initializers) == list(init)
        for k, a in init.items():
            assert g.initializers[k].dtype == a.dtype and g.initializers[k].shape == a.shape
            np.testing.assert_array_equal(g.initializers[k], a)
        assert g.inputs == ["x"] and g.outputs == ["y"]         # initialisers listed as inputs are not feeds
        assert [(n.op_type, n.inputs, n.outputs, n.name) for n in g.nodes] == [("MatMul", ["x", "a.raw"], ["y"], "mm")]


def test_bf16_and_constant_nodes():
    # bfloat16 tensor (data_type 16) as raw bits, and a Constant node carrying a tensor attribute
    vals = np.array([1.0, -2.5, 0.15625], dtype=np.float32)
    bits = (vals.view(np.uint32) >> 16).astype("<u2").tobytes()
    t = oi._enc_varint(8) + oi._enc_varint(3) + oi._enc_varint(16) + oi._enc_varint(16) + oi._enc_ld(8, b"w.bf16") + \
        oi._enc_ld(9, bits)
    const_t = oi._enc_tensor("", np.arange(6, dtype=np.float32).reshape(2, 3))
    attr = oi._enc_ld(1, b"value") + oi._enc_ld(5, const_t)
    node = oi._enc_ld(2, b"c_out") + oi._enc_ld(4, b"Constant") + oi._enc_ld(5, attr)
    graph = oi._enc_ld(1, node) + oi._enc_ld(5, t)
    g = oi.parse_model(oi._enc_ld(7, graph))
    np.testing.assert_array_equal(g.initializers["w.bf16"], vals)
    np.testing.assert_array_equal(g.initializers["c_out"], np.arange(6, dtype=np.float32).reshape(2, 3))


def test_rejects_garbage_and_external_data():
    with pytest.raises(ValueError):
        oi.parse_model(b"\x08\x08")                       # a ModelProto without a graph
    with pytest.raises(ValueError):
        oi.parse_model(b"\x3a\xff\xff\x03abc")           # length runs past the end
    ext = oi._enc_ld(8, b"w") + oi._enc_varint((14 << 3) | 0) + oi._enc_varint(1)
    with pytest.raises(ValueError, match="external data"):
        oi.parse_model(oi._enc_ld(7, oi._enc_ld(5, ext)))


@pytest.mark.parametrize("anonymous", [False, True])
def test_convert_recovers_every_tensor_bit_exact(anonymous):
    W = _weights()
    graphs = oi.export_initializers(TINY, W, anonymous_matmul=anonymous)
    arch, T, rep = oi.convert_graphs(graphs, base=FULL)
    assert rep.ok() and not rep.leftover and not rep.computed
    assert set(T) == set(W)
    for k in W:
        assert T[k].shape == W[k].shape, k
        np.testing.assert_array_equal(T[k], W[k], err_msg=k)
    if anonymous:
        assert "dit.blocks.0.qkv.w@1" in rep.by_edge and "dit.out.w" in rep.by_edge and "voc.head.w" in rep.by_edge
    # every constant the shapes determine came back; the rest is FULL's (TINY differs from FULL only in nfe there)
    for f in ("dim", "depth", "heads", "ff_dim", "text_dim", "vocab", "text_layers", "text_ff", "conv_pos_groups",
              "conv_pos_k", "voc_dim", "voc_ff", "voc_layers", "voc_k", "n_mel", "n_fft", "time_freq_dim"):
        assert getattr(arch, f) == getattr(TINY, f), f


def test_missing_and_misshaped_tensors_are_reported():
    W = _weights()
    graphs = oi.export_initializers(TINY, {k: v for k, v in W.items() if k != "dit.blocks.1.ff2.w" and
                                           k != "dit.blocks.1.ff2.b" and k != "pre.mel_fb"})
    with pytest.raises(ValueError, match="dit.blocks.1.ff2"):
        oi.convert_graphs(graphs)
    arch, T, rep = oi.convert_graphs(graphs, strict=False)
    assert rep.missing == ["dit.blocks.1.ff2.w", "dit.blocks.1.ff2.b"]
    assert rep.computed == ["pre.mel_fb"]                 # fixed table rebuilt from the architecture
    np.testing.assert_array_equal(T["pre.mel_fb"], W["pre.mel_fb"])
    bad = dict(W)
    bad["dit.out.b"] = np.zeros(TINY.n_mel + 1, np.float32)
    with pytest.raises(ValueError, match="dit.out.b"):
        oi.convert_graphs(oi.export_initializers(TINY, bad))


def test_mel_filterbank_found_by_shape_in_either_orientation():
    W = _weights()
    graphs = oi.export_initializers(TINY, {k: v for k, v in W.items() if k != "pre.mel_fb"})
    g = oi.parse_model(graphs["preprocess"])
    g.initializers["/mel/Constant_7_output_0"] = np.ascontiguousarray(W["pre.mel_fb"].T)      # [n_mel, n_bins]
    graphs["preprocess"] = oi.write_model(g.initializers, g.nodes)
    _, T, rep = oi.convert_graphs(graphs)
    assert rep.mapped["pre.mel_fb"] == "/mel/Constant_7_output_0" and not rep.leftover
    np.testing.assert_array_equal(T["pre.mel_fb"], W["pre.mel_fb"])


def test_tar_rewrite_keeps_layout_and_blobs_unpack(tmp_path):
    W = _weights()
    graphs = oi.export_initializers(TINY, W, anonymous_matmul=True)
    src, dst = str(tmp_path / "model-bin.pt"), str(tmp_path / "model-b200.pt")
    meta = [{"file_name": "v.wav", "text": "xin chào.", "gender": "female", "group": "story", "area": "northern",
             "emotion": "neutral"}]
    wav = artifact._wav_bytes(artifact.synthetic_prompt_pcm(2400, 1), 24000)
    members = [("model/preprocess.onnx", graphs["preprocess"]), ("model/transformer.onnx", graphs["transformer"]),
               ("model/decode.onnx", graphs["decode"]), ("model/vocab.txt", "a\nb\n".encode()),
               ("model/audio_metadata.json", json.dumps(meta).encode()), ("model/cleaned_audios/v.wav", wav)]
    with tarfile.open(src, "w") as tar:
        for name, data in members:
            ti = tarfile.TarInfo(name)
            ti.size = len(data)
            tar.addfile(ti, io.BytesIO(data))
    rep = oi.convert_model_tar(src, dst)
    assert rep.ok()
    with tarfile.open(dst, "r") as tar:
        assert tar.getnames() == [n for n, _ in members]
        for name, data in members[3:]:
            assert tar.extractfile(name).read() == data                     # copied byte for byte
        for graph in ("preprocess", "transformer", "decode"):
            blob = tar.extractfile(f"model/{graph}.onnx").read()
            arch, T, gid = artifact.unpack_blob(blob)
            assert gid == artifact.GRAPH_IDS[graph] and arch.dim == TINY.dim and arch.depth == TINY.depth
            for k, a in T.items():
                np.testing.assert_array_equal(a, W[k], err_msg=k)
    with pytest.raises(ValueError, match="already is a VVB200 blob"):
        oi.convert_model_tar(dst, str(tmp_path / "again.pt"))
    assert oi.main([src, str(tmp_path / "cli.pt")]) == 0


def test_converted_weights_drive_the_oracle_identically():
    torch = pytest.importorskip("torch")
    from oracle.graphs import OracleSessions
    torch.set_num_threads(2)
    W = _weights()
    _, T, _ = oi.convert_graphs(oi.export_initializers(TINY, W, anonymous_matmul=True), base=TINY)
    audio = artifact.synthetic_prompt_pcm(6000, 5).reshape(1, 1, -1)
    ids = np.arange(1, 21, dtype=np.int32).reshape(1, -1)
    T_total = np.array([6000 // 256 + 1 + 24], dtype=np.int64)
    outs = []
    for weights in (W, T):
        S = OracleSessions(TINY, weights)
        pre = S.preprocess.run(audio, ids, T_total)
        x, _ = S.transformer.run(*pre[:7], np.array([0], dtype=np.int32))
        outs.append(list(pre) + [x, S.decode.run(x, pre[7])[0]])
    for a, b in zip(*outs):
        np.testing.assert_array_equal(np.asarray(a), np.asarray(b))


def test_full_architecture_parameter_count():
    """the blob layout at the named size: 14.69 M parameters per DiT block, 323 M in the 22 blocks"""
    shapes = oi.expected_shapes(FULL)
    n = lambda prefix: sum(int(np.prod(s)) for k, s in shapes.items() if k.startswith(prefix))
    d = FULL.dim
    assert n("dit.blocks.0.") == 14 * d * d + 13 * d == 14_693_376
    assert n("dit.blocks.") == 22 * 14_693_376
    assert shapes["dit.in.w"] == (d, FULL.in_dim) == (1024, 712)
    assert shapes["dit.pos.c1.w"] == (d, d // FULL.conv_pos_groups, FULL.conv_pos_k) == (1024, 64, 31)
    assert shapes["pre.text_embed"] == (FULL.vocab + 1, FULL.text_dim)
    assert shapes["voc.head.w"] == (FULL.n_fft + 2, FULL.voc_dim) == (1026, 512)
    assert 13_000_000 < n("voc.") < 14_500_000          # Vocos-mel-24k backbone + ISTFT head
