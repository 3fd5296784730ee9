"""CPU-side checks: the C-ABI library loads and exports every symbol include/cmpc_b200.h declares (no compute calls --
there is no GPU here), argument validation fails loudly, and the host-side packing / sharding logic is right."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


def _declared_symbols():
    text = (ROOT / "include" / "cmpc_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cmpc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    declared = _declared_symbols()
    assert len(declared) >= 20
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/cmpc_b200.h but not exported"
    from cmpc_refseg_b200 import _lib
    assert sorted(_lib.exported_symbols()) == declared, "ctypes binding and header disagree"


def test_version_and_error_string(lib):
    assert lib.cmpc_version() >= 100
    assert isinstance(lib.cmpc_last_error(), bytes)


def test_null_args_fail_loudly(lib):
    from cmpc_refseg_b200 import _lib
    rc = lib.cmpc_gemm_f16(None, None)
    assert rc == -1 and b"null args" in lib.cmpc_last_error()
    with pytest.raises(_lib.CmpcError):
        _lib.check(rc, "cmpc_gemm_f16")


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_gpu_means_error_not_fallback(lib):
    from cmpc_refseg_b200 import _lib
    g = _lib.GemmArgs()
    buf = (C.c_char * 64)()
    g.a1 = g.w = g.out = C.addressof(buf)
    g.m = g.n = g.k1 = 8; g.lda1 = g.ldw = g.ldo = 8; g.rows_per_sample = 8
    rc = lib.cmpc_gemm_f16(C.byref(g), None)
    assert rc == -3, "without an sm_100 device the call must return CMPC_ERR_ARCH, not compute anything"
    from cmpc_refseg_b200.CMPC_model import LSTM_model
    with pytest.raises((_lib.CmpcError, RuntimeError, AssertionError)):
        LSTM_model(batch_size=1, vf_h=2, vf_w=2, H=16, W=16, vf_dim=16, v_emb_dim=16, rnn_size=16, mlp_dim=8, device="cpu")


def test_product_path_never_imports_the_oracle():
    for f in (ROOT / "cmpc_refseg_b200").glob("*.py"):
        assert "oracle" not in re.sub(r'""".*?"""', "", f.read_text(), flags=re.S).replace("# ", ""), f"{f.name} references the oracle"


def test_mutan_weight_packing_layout():
    from cmpc_refseg_b200.weights import pack_mutan_weights
    C_, K = 56, 64
    dws = [torch.randn(1, 1, C_ + 8, C_) for _ in range(5)]
    w = pack_mutan_weights(dws, C_, 64).float()
    assert w.shape == (2 * 240, 64)
    for (j, k, cc) in [(0, 0, 0), (0, 3, 47), (1, 4, 7), (1, 2, 8)]:
        c = j * 48 + cc
        row = w[j * 240 + k * 48 + cc]
        if c < C_:
            assert torch.allclose(row[:C_ + 8], dws[k][0, 0, :, c].half().float())
            assert row[C_ + 8:].abs().sum() == 0
        else:
            assert row.abs().sum() == 0                   # channels beyond C are zero rows


def test_head_weight_packing_matches_reference_layouts():
    from cmpc_refseg_b200.CMPC_model import head_param_shapes, reference_init
    from cmpc_refseg_b200.weights import Dims, pack_head_weights
    kw = dict(vf_h=2, vf_w=3, vf_dim=32, v_emb_dim=24, rnn_size=24, mlp_dim=12, c4_dim=16, c3_dim=8, parse_hidden=10)
    params = reference_init(head_param_shapes(**kw), seed=3)
    for k in params:                                    # make biases / LN params non-trivial
        if k.endswith(("biases", "beta")):
            params[k] = torch.randn_like(params[k]) * 0.1
    d = Dims(C=24, R=24, Mm=12, T=5, HID=10, h=2, w=3, H=16, W=24, cin={"c5": 32, "c4": 16, "c3": 8})
    W = pack_head_weights(params, d, torch.device("cpu"))
    C_, R, Mm, GW = 24, 24, 12, d.GW
    # 1x1 conv: W16[cout, cin] = DW[0,0,cin,cout]
    assert torch.allclose(W["lat_w_c4"][:C_, :16].float(), params["c4_lateral/DW"][0, 0].t().half().float())
    # fusion conv: K segment 1 = vis_la_sp rows, segment 2 = spa_graph rows followed by the 8 spatial rows; language rows -> per-sample bias GEMM
    dw = params["fusion_c3/DW"][0, 0]
    k1p = 64
    fw = W["fusion_w_c3"].float()
    assert torch.allclose(fw[:Mm, :C_], dw[:C_].t().half().float())
    assert torch.allclose(fw[:Mm, k1p:k1p + C_], dw[C_:2 * C_].t().half().float())
    assert torch.allclose(fw[:Mm, k1p + C_:k1p + C_ + 8], dw[2 * C_ + R:].t().half().float())
    i3 = 2                                                # LEVELS = (c5, c4, c3)
    assert torch.allclose(W["fsb_w"][i3 * GW:i3 * GW + Mm, :R].float(), dw[2 * C_:2 * C_ + R].t().half().float())
    assert torch.allclose(W["fsb_b"][i3 * GW:i3 * GW + Mm], params["fusion_c3/biases"])
    # affinity re-association weight: rows = DW2[cin, :], extra row C = bias
    g = W["gt_w_c5"].float()
    assert torch.allclose(g[:C_, :R], params["spa_graph_trans2_c5/DW"][0, 0].half().float())
    assert torch.allclose(g[C_, :R], params["spa_graph_trans2_c5/biases"].half().float())
    # ConvLSTM kernel: gate g rows at g*GW, K segments [x | h]
    kern = params["rnn/conv_lstm_cell/kernel"][0, 0]
    lw = W["lstm_w"].float()
    assert torch.allclose(lw[2 * GW:2 * GW + Mm, :Mm], kern[:Mm, 2 * Mm:3 * Mm].t().half().float())
    assert torch.allclose(lw[3 * GW:3 * GW + Mm, 64:64 + Mm], kern[Mm:, 3 * Mm:].t().half().float())
    # score taps: row k = 3*dy + dx
    assert torch.allclose(W["score_w"][5, :Mm].float(), params["score/DW"][1, 2, :, 0].half().float())
    # key conv folded into the query: keyT[o, cin] = DW_key[cin, o]
    assert torch.allclose(W["keyT"][4], params["spa_graph_key_c4_2gv_f1/DW"][0, 0].t())


def test_shard_range_partitions_the_batch():
    from cmpc_refseg_b200.parallel import shard_range
    for world in (1, 2, 3, 8):
        for B in (1, 7, 32, 256):
            spans = [shard_range(r, world, B) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(2, 2, 8)


def test_iou_summary_matches_oracle_bookkeeping():
    from cmpc_refseg_b200.parallel import local_iou_stats, summarize
    from oracle.cmpc_head_ref import iou_stats
    I = torch.tensor([10, 0, 50, 90]); U = torch.tensor([20, 30, 60, 100])
    s = summarize(local_iou_stats(I, U)); o = iou_stats(I, U)
    assert s["cum_I"] == o["cum_I"] and s["cum_U"] == o["cum_U"]
    assert s["overall_iou"] == pytest.approx(o["overall_iou"]) and s["mean_iou"] == pytest.approx(o["mean_iou"])
    assert s["precision@0.5"] == pytest.approx(o["prec@0.5"] / 4) and s["precision@0.9"] == pytest.approx(o["prec@0.9"] / 4)


def test_synthetic_inputs_equal_the_oracle_generator():
    from cmpc_refseg_b200.synthetic import make_inputs
    from oracle.cmpc_head_ref import HeadConfig, make_inputs as oracle_inputs
    kw = dict(vf_h=3, vf_w=3, H=24, W=24, c3_dim=8, c4_dim=16, vf_dim=32, num_steps=5, rnn_size=16)
    a = make_inputs(3, seed=9, seq_len=[5, 2, 1], **kw)
    b = oracle_inputs(HeadConfig(batch_size=3, v_emb_dim=16, mlp_dim=8, **kw), 3, seed=9, seq_len=[5, 2, 1])
    assert all(torch.equal(a[k], b[k]) for k in a)


def test_drop_in_exposes_the_reference_method_signatures():
    """Method names and positional argument order of LSTM_model (reference CMPC_model.py, line of each `def` cited)."""
    import inspect
    from cmpc_refseg_b200.CMPC_model import LSTM_model
    reference = {                                                     # CMPC_model.py:<line>
        "valid_lang": ["words_parse", "words_feat"],                                                    # :166
        "nec_lang": ["words_parse", "words_feat"],                                                      # :180
        "lang_se": ["feat", "lang_feat", "level"],                                                      # :194
        "global_vec": ["feat", "lang_feat", "level"],                                                   # :212
        "gated_exchange_module": ["feat", "feat1", "feat2", "lang_feat", "level"],                      # :245
        "gated_exchange_fusion_lstm_2times": ["feat3", "feat4", "feat5", "lang_feat"],                  # :261
        "mutan_head": ["lang_feat", "spatial_feat", "visual_feat", "level"],                            # :295
        "mutan_fusion": ["lang_feat", "spatial_feat", "visual_feat", "level"],                          # :311
        "build_lang2vis": ["visual_feat", "words_feat", "lang_feat", "words_parse", "spatial", "level"],  # :330
        "build_lang_parser": ["words_feat"],                                                            # :347
        "graph_conv": ["graph_feat", "nodes_num", "nodes_dim", "adj_mat", "graph_name", "level"],       # :359
        "build_spa_graph": ["spa_graph", "words_feat", "spatial", "words_parse", "level"],              # :376
        "_conv": ["name", "x", "filter_size", "in_filters", "out_filters", "strides"],                  # :412
    }
    for name, args in reference.items():
        sig = list(inspect.signature(getattr(LSTM_model, name)).parameters)
        assert sig == ["self"] + args, (name, sig)
    for name in ("build_graph", "lstm", "train_op"):                                                   # :89, :144, :426
        assert callable(getattr(LSTM_model, name))


def test_checkpoint_npz_round_trip(tmp_path):
    """TF-variable-name keyed .npz exchange (cmpc_refseg_b200/checkpoint.py): scope handling, shape check, foreign variables ignored."""
    import numpy as np
    from cmpc_refseg_b200.CMPC_model import head_param_shapes, reference_init
    from cmpc_refseg_b200.checkpoint import TF_SCOPE, load_variables, save_variables
    kw = dict(vf_h=4, vf_w=4, vf_dim=32, v_emb_dim=16, rnn_size=16, mlp_dim=8, c4_dim=16, c3_dim=8, parse_hidden=12)
    shapes = head_param_shapes(**kw)
    params = reference_init(shapes, seed=3)
    f = str(tmp_path / "head.npz")
    save_variables(f, params)
    back = load_variables(f, shapes)
    assert set(back) == set(shapes) and all(torch.equal(back[k], params[k]) for k in shapes)
    # a reference-side dump also carries backbone / LSTM / Adam variables and ':0' suffixes
    extra = {TF_SCOPE + k + ":0": v.numpy() for k, v in params.items()}
    extra["res5c_branch2c/weights"] = np.zeros((1, 1, 4, 4), np.float32)
    extra[TF_SCOPE + "rnn/lstm_cell/kernel"] = np.zeros((4, 4), np.float32)
    np.savez(str(tmp_path / "dump.npz"), **extra)
    back = load_variables(str(tmp_path / "dump.npz"), shapes)
    # the word-encoder variables travel with the head's (LSTM_model.encode_words), the backbone's are dropped
    assert set(back) == set(shapes) | {"rnn/lstm_cell/kernel"} and len(load_variables.last_ignored) == 1
    bad = dict(extra)
    bad[TF_SCOPE + "score/DW:0"] = np.zeros((3, 3, 9, 1), np.float32)
    np.savez(str(tmp_path / "bad.npz"), **bad)
    with pytest.raises(ValueError):
        load_variables(str(tmp_path / "bad.npz"), shapes)
    del extra[TF_SCOPE + "score/biases:0"]
    np.savez(str(tmp_path / "missing.npz"), **extra)
    with pytest.raises(KeyError):
        load_variables(str(tmp_path / "missing.npz"), shapes)


def test_oracle_word_lstm_is_the_tf_lstm_cell_under_dynamic_rnn():
    """The oracle's word encoder (CMPC_model.py:144-157) against an independent LSTM (torch.nn.LSTM on packed sequences):
    TF gate order [i, j, f, o] with forget_bias 1 vs torch's [i, f, g, o]; zero output and frozen state past sequence_length."""
    from oracle.cmpc_head_ref import word_lstm
    g = torch.Generator().manual_seed(3)
    B, T, E, R, V = 3, 7, 10, 6, 17
    emb = torch.randn(V, E, generator=g, dtype=torch.float64)
    kern = torch.randn(E + R, 4 * R, generator=g, dtype=torch.float64) * 0.4
    bias = torch.randn(4 * R, generator=g, dtype=torch.float64) * 0.2
    words = torch.randint(0, V, (B, T), generator=g)
    seq_len = torch.tensor([7, 3, 1])
    got = word_lstm(words, seq_len, emb, kern, bias)
    ref = torch.nn.LSTM(E, R, batch_first=True).double()
    i, j, f, o = torch.split(kern, R, dim=1)
    bi, bj, bf, bo = torch.split(bias, R)
    w = torch.cat([i, f, j, o], 1)
    with torch.no_grad():
        ref.weight_ih_l0.copy_(w[:E].t()); ref.weight_hh_l0.copy_(w[E:].t())
        ref.bias_ih_l0.copy_(torch.cat([bi, bf + 1.0, bj, bo])); ref.bias_hh_l0.zero_()
        packed = torch.nn.utils.rnn.pack_padded_sequence(emb[words], seq_len, batch_first=True, enforce_sorted=True)
        want, _ = torch.nn.utils.rnn.pad_packed_sequence(ref(packed)[0], batch_first=True, total_length=T)
    assert torch.allclose(got, want, atol=1e-12)
    assert torch.all(got[1, 3:] == 0) and torch.all(got[2, 1:] == 0)


def test_resize_and_crop_oracle_and_host_arithmetic():
    """Post-processing oracle (util/im_processing.py:25-41 via the restated skimage resize): hand-checked values, the host-side
    integer arithmetic of the drop-in, and the reader / evaluator bookkeeping (trainval_model.py:266-296)."""
    from oracle.cmpc_head_ref import postprocess_iu, resize_and_crop_mask
    from cmpc_refseg_b200.postprocess import NpzBatchReader, SegEvaluator, resize_and_crop_meta
    m = (np.random.RandomState(0).rand(8, 8) > 0.5).astype(np.float32)
    assert np.array_equal(resize_and_crop_mask(m, 8, 8), m)                       # same size: identity
    r = resize_and_crop_mask(np.ones((2, 2), np.float32), 4, 4)                   # 2x: first sample at -0.25 blends with cval 0
    assert np.allclose(r[0], [0.5625, 0.75, 0.75, 0.5625]) and np.allclose(r[1], [0.75, 1, 1, 0.75])
    assert np.array_equal(resize_and_crop_mask(np.ones((2, 2), np.float32), 4, 4, "reflect"), np.ones((4, 4)))
    one = np.zeros((4, 4), np.float32); one[1, 2] = 1
    r = resize_and_crop_mask(one, 8, 8)                                           # a single pixel spreads over its bilinear footprint
    assert set(zip(*np.nonzero(r))) == {(y, x) for y in range(1, 5) for x in range(3, 7)}
    # aspect change: scale = max ratio, centre crop (320x320 -> 427x640 resizes to 640x640, drops 106 rows at the top)
    assert resize_and_crop_meta(320, 320, 427, 640) == (640, 640, 106, 0)
    assert resize_and_crop_meta(320, 320, 640, 480) == (640, 640, 0, 80)
    assert resize_and_crop_meta(320, 320, 333, 500) == (500, 500, 83, 0)
    up = np.random.RandomState(1).randn(1, 16, 16, 1)
    gt = np.random.RandomState(2).rand(21, 30) > 0.5
    pred, I, U = postprocess_iu(up, gt)
    assert pred.shape == gt.shape and 0 < I <= U <= gt.size
    ev = SegEvaluator()
    ev.update(torch.tensor([50, 10]), torch.tensor([100, 100]))
    s = ev.finish()
    assert s["overall_iou"] == 0.3 and abs(s["mean_iou"] - 0.3) < 1e-12 and s["precision@0.5"] == 0.5 and s["precision@0.6"] == 0.0
    assert SegEvaluator.report(s).splitlines()[1] == "precision@0.5 = 0.500000" and "overall IoU = 0.300000; mean IoU = 0.300000" in SegEvaluator.report(s)


def test_npz_batch_reader(tmp_path):
    from cmpc_refseg_b200.postprocess import NpzBatchReader
    for i in range(3):
        np.savez(str(tmp_path / f"unc_train_{i}.npz"), text_batch=np.full(20, i), im_batch=np.zeros((4, 4, 3), np.uint8),
                 mask_batch=np.ones((4, 4), bool), sent_batch=[f"sentence {i}"])
    rd = NpzBatchReader(str(tmp_path), "unc_train", shuffle=False, prefetch_num=2)
    seen = [int(rd.read_batch(is_log=False)["text_batch"][0]) for _ in range(7)]
    assert seen == [0, 1, 2, 0, 1, 2, 0] and rd.n_epoch == 2 and rd.n_batch == 1
    b = rd.read_batch(is_log=False)
    assert set(b) == {"text_batch", "im_batch", "mask_batch", "sent_batch"} and str(b["sent_batch"][0]) == "sentence 1"
    with pytest.raises(RuntimeError):
        (tmp_path / "empty").mkdir()
        NpzBatchReader(str(tmp_path / "empty"), "x")


def test_eval_loop_bookkeeping_matches_the_reference_loop():
    """cmpc_refseg_b200/evaluate.py::test against a literal restatement of the counters of trainval_model.py:161-166, 266-296, with a
    stand-in model (no GPU): per-sample I / U, cumulative IoU, mean IoU, precision@X, the report text, and the 2-way shard split."""
    import io
    import numpy as np
    from cmpc_refseg_b200 import evaluate
    rng = np.random.default_rng(0)
    H = W = 16

    class FakeModel:
        device, num_steps, H, W = torch.device("cpu"), 5, 16, 16

        def run(self, fetches, feed_dict):
            up = feed_dict["visual_feat_c5"][:, :, :, :1] - 0.5            # the "logits" are smuggled in through the c5 tap
            return [up[:, ::2, ::2], up, torch.sigmoid(up), torch.zeros(1, 1, 5, 4)]

    batches = []
    for i in range(23):
        batches.append(dict(visual_feat_c3=np.zeros((H, W, 1), np.float32), visual_feat_c4=np.zeros((H, W, 1), np.float32),
                            visual_feat_c5=rng.random((H, W, 1)).astype(np.float32), lstm_outputs=np.zeros((5, 4), np.float32),
                            mask_batch=(rng.random((H, W)) > 0.6), sent_batch=["s%d" % i]))

    class Reader:
        def __init__(self):
            self.num_batch, self.i = len(batches), 0

        def read_batch(self, is_log=True):
            b = batches[self.i % len(batches)]
            self.i += 1
            return b

    def iu_fn(up, mask):          # same-size masks: resize_and_crop is the identity
        pred = (up.reshape(H, W) >= 1e-9)
        m = mask.bool()
        return torch.tensor([int((pred & m).sum())]), torch.tensor([int((pred | m).sum())])
    res = evaluate.test(FakeModel(), Reader(), iu_fn=iu_fn, out=io.StringIO())
    # the reference's counters, literally
    cum_I = cum_U = 0
    mean_IoU, seg_correct, seg_total = 0.0, np.zeros(5, np.int32), 0.0
    for b in batches:
        pred_raw = (b["visual_feat_c5"][:, :, 0] - np.float32(0.5) >= 1e-9)
        I, U = np.sum(np.logical_and(pred_raw, b["mask_batch"])), np.sum(np.logical_or(pred_raw, b["mask_batch"]))
        mean_IoU += float(I) / U; cum_I += I; cum_U += U
        for n, th in enumerate([.5, .6, .7, .8, .9]):
            seg_correct[n] += (I / U >= th)
        seg_total += 1
    s = res["summary"]
    assert s["cum_I"] == cum_I and s["cum_U"] == cum_U and s["n"] == 23
    assert abs(s["mean_iou"] - mean_IoU / seg_total) < 1e-12 and abs(s["overall_iou"] - cum_I / cum_U) < 1e-12
    want = "".join("precision@%s = %f\n" % (str(th), seg_correct[n] / seg_total) for n, th in enumerate([.5, .6, .7, .8, .9]))
    want += "overall IoU = %f; mean IoU = %f\n" % (cum_I / cum_U, mean_IoU / seg_total)
    assert want in res["report"] and "Segmentation evaluation (without DenseCRF):" in res["report"]
    assert [r["batch_no"] for r in res["IU_result"]] == list(range(23))
    # two "ranks" (no process group: summaries are per shard) cover the batches exactly once
    parts = [evaluate.test(FakeModel(), Reader(), iu_fn=iu_fn, out=io.StringIO(), rank=r, world=2) for r in range(2)]
    assert sum(p["summary"]["cum_I"] for p in parts) == cum_I and sum(p["summary"]["n"] for p in parts) == 23
    # a DenseCRF stand-in is evaluated through the same resize / IU path
    res = evaluate.test(FakeModel(), Reader(), iu_fn=iu_fn, out=io.StringIO(), dcrf=lambda sigm, b: (sigm > 0.5).astype(np.float32))
    assert res["summary_dcrf"]["cum_I"] == cum_I and "with DenseCRF" in res["report"]
