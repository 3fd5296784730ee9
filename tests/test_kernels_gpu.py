"""Per-kernel GPU tests through the C ABI (ctypes) against plain torch fp32 math on the same fp16-rounded operands,
plus size-independent properties at the full BASELINE size (batch 32, N = 1600)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env(lib):
    from cmpc_refseg_b200 import _lib as L
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    return L, lib, dev, torch.cuda.current_stream(dev).cuda_stream


def _rup(x, m):
    return (x + m - 1) // m * m


@pytest.mark.parametrize("M,N,K1,K2,act,fp32,rps", [
    (128, 256, 64, 0, 0, False, None),            # one tile, one k-step
    (1000, 500, 1000, 0, 1, False, 250),          # ragged M, ragged N, ragged K, per-sample bias + gate, stats
    (3200, 1000, 1000, 1008, 0, True, 1600),      # two K segments (fusion-style), fp32 output, sample straddling a tile
    (51200, 1000, 2048, 0, 0, False, 1600),       # lateral c5 at full size
    (640, 32, 500, 0, 0, True, None),             # skinny (N <= 32) path: score taps / affinity shape class
    (1000, 256, 64, 0, 1, False, 100),            # samples shorter than a tile: no tile descriptors, warps straddling two samples
    (1000000, 256, 64, 0, 1, False, 250000),      # 53 tiles per CTA: more than the descriptor table holds -> per-tile index math
])
def test_gemm_matches_torch(env, M, N, K1, K2, act, fp32, rps):
    L, lib, dev, st = env
    rps = rps or M
    B = (M + rps - 1) // rps
    kp1, kp2 = _rup(K1, 64), (_rup(K2, 64) if K2 else 0)
    a1 = torch.full((M, kp1), float("nan"), device=dev, dtype=torch.float16); a1[:, :K1] = torch.randn(M, K1, device=dev) * 0.5
    a2 = None
    if K2:
        a2 = torch.full((M, kp2), float("nan"), device=dev, dtype=torch.float16); a2[:, :K2] = torch.randn(M, K2, device=dev) * 0.5
    w = torch.zeros(N, kp1 + kp2, device=dev, dtype=torch.float16)
    w[:, :K1] = torch.randn(N, K1, device=dev) * 0.05
    if K2:
        w[:, kp1:kp1 + K2] = torch.randn(N, K2, device=dev) * 0.05
    npad = _rup(N, 256)
    bias = torch.randn(npad, device=dev); sbias = torch.randn(B, npad, device=dev); gate = torch.rand(B, npad, device=dev)
    ldo = _rup(N, 32)
    out = torch.full((M, ldo), 777.0, device=dev, dtype=torch.float32 if fp32 else torch.float16)
    rs = torch.zeros(M, device=dev); stt = torch.zeros(B, 1, 2, device=dev, dtype=torch.float64)
    g = L.GemmArgs()
    g.a1, g.lda1, g.k1 = a1.data_ptr(), kp1, K1
    if K2:
        g.a2, g.lda2, g.k2 = a2.data_ptr(), kp2, K2
    g.w, g.ldw, g.m, g.n, g.rows_per_sample = w.data_ptr(), kp1 + kp2, M, N, rps
    g.bias, g.sbias, g.ld_sbias, g.gate, g.ld_gate, g.act = bias.data_ptr(), sbias.data_ptr(), npad, gate.data_ptr(), npad, act
    g.out, g.ldo, g.out_fp32, g.row_sumsq, g.stats = out.data_ptr(), ldo, int(fp32), rs.data_ptr(), stt.data_ptr()
    L.check(lib.cmpc_gemm_f16(C.byref(g), st), "gemm")
    torch.cuda.synchronize()
    ref = a1[:, :K1].float() @ w[:, :K1].float().t()
    if K2:
        ref = ref + a2[:, :K2].float() @ w[:, kp1:kp1 + K2].float().t()
    bidx = torch.arange(M, device=dev) // rps
    ref = ref + bias[:N] + sbias[bidx, :N]
    if act == 1:
        ref = torch.relu(ref)
    ref = ref * gate[bidx, :N]
    got = out[:, :N].float()
    tol = 2e-3 if fp32 else 2e-2                                # fp16 output rounding at |x| <= ~16
    assert (got - ref).abs().max() < tol
    assert (out[:, N:] == 0).all(), "padding columns must be written as zero"
    ss = (ref ** 2).sum(1)
    assert ((rs - ss).abs() / ss.clamp_min(1e-6)).max() < 1e-3
    s1 = torch.zeros(B, device=dev, dtype=torch.float64).index_add_(0, bidx, ref.double().sum(1))
    s2 = torch.zeros(B, device=dev, dtype=torch.float64).index_add_(0, bidx, (ref.double() ** 2).sum(1))
    assert ((stt[:, 0, 0] - s1).abs() / s1.abs().clamp_min(1.0)).max() < 1e-3
    assert ((stt[:, 0, 1] - s2).abs() / s2).max() < 1e-3


def test_gemm_rejects_bad_arguments(env):
    L, lib, dev, st = env
    a = torch.zeros(128, 64, device=dev, dtype=torch.float16)
    g = L.GemmArgs()
    g.a1, g.lda1, g.k1, g.w, g.ldw, g.m, g.n, g.rows_per_sample = a.data_ptr(), 64, 64, a.data_ptr(), 64, 128, 128, 128
    g.out, g.ldo = a.data_ptr() + 2, 64                        # misaligned output pointer
    assert lib.cmpc_gemm_f16(C.byref(g), st) == -2
    g.out, g.ldo = a.data_ptr(), 8                             # ldo smaller than n
    assert lib.cmpc_gemm_f16(C.byref(g), st) == -1
    with pytest.raises(L.CmpcError):
        L.check(-1, "cmpc_gemm_f16")


@pytest.mark.parametrize("M,Cc,K,rps", [(256, 96, 64, 128), (3200, 1000, 1008, 1600)])
def test_mutan_epilogue_matches_torch(env, M, Cc, K, rps):
    from cmpc_refseg_b200.weights import pack_mutan_weights
    L, lib, dev, st = env
    B = M // rps
    kp = _rup(K, 64)
    a = torch.full((M, kp), float("nan"), device=dev, dtype=torch.float16); a[:, :K] = torch.randn(M, K, device=dev) * 0.3
    dws = [torch.randn(1, 1, K, Cc, device=dev) * 0.05 for _ in range(5)]
    w = pack_mutan_weights(dws, Cc, kp)
    ldb = _rup(Cc, 64)
    bias = torch.randn(5, ldb, device=dev) * 0.1
    lang = torch.tanh(torch.randn(B, 5, ldb, device=dev))
    out = torch.full((M, ldb), 777.0, device=dev); rs = torch.zeros(M, device=dev)
    m = L.MutanArgs()
    m.a, m.lda, m.k, m.w, m.ldw, m.m, m.c, m.rows_per_sample = a.data_ptr(), kp, K, w.data_ptr(), kp, M, Cc, rps
    m.bias, m.ld_bias, m.lang, m.ld_lang, m.out, m.ldo, m.row_sumsq = bias.data_ptr(), ldb, lang.data_ptr(), ldb, out.data_ptr(), ldb, rs.data_ptr()
    L.check(lib.cmpc_mutan_f16(C.byref(m), st), "mutan")
    torch.cuda.synchronize()
    bidx = torch.arange(M, device=dev) // rps
    pre = torch.stack([a[:, :K].float() @ dw[0, 0].half().float() for dw in dws], 1)           # [M, 5, C]
    ref = torch.tanh((torch.tanh(pre + bias[None, :, :Cc]) * lang[bidx][:, :, :Cc]).sum(1))
    assert (out[:, :Cc] - ref).abs().max() < 1e-4
    assert ((rs - (ref ** 2).sum(1)).abs() / (ref ** 2).sum(1)).max() < 1e-3


@pytest.mark.parametrize("mode", [0, 2, 4, 5])
@pytest.mark.parametrize("B,N,Cc,T", [(1, 128, 256, 20), (2, 200, 64, 7), (2, 1600, 1000, 20), (1, 4096, 1000, 20), (40, 60, 72, 12),
                                      (9, 512, 520, 20), (5, 1000, 264, 3)])
def test_graph_reason_dense_adjacency(env, B, N, Cc, T, mode):
    """Y = (W V^T) X with the adjacency tiles dumped for inspection: ragged N, odd / even query-tile counts (the odd tile is
    paired across channel chunks), one to many units per persistent cluster (fewer than 3 key tiles per unit included), odd
    channel-chunk counts, the 4096-node high-resolution case of BASELINE config 4.  mode 0 = default kernel (skip-once ring order for
    6 or more key tiles per unit: N = 1000, 1600, 4096 here), 2 = 2-SM variant, 4 = round-robin ring + convert-first epilogue (the kernel
    of the start of round 2), 5 = skip-once ring + convert-first epilogue."""
    L, lib, dev, st = env
    lib.cmpc_graph_set_mode(mode)
    ldx = _rup(Cc + 8, 64)
    w = torch.zeros(B * N, 32, device=dev); v = torch.zeros(B * N, 32, device=dev)
    w[:, :T] = torch.softmax(torch.randn(B * N, T, device=dev) * 2, -1)
    vs = float(1 << (N - 1).bit_length())
    v[:, :T] = torch.softmax(torch.randn(B, N, T, device=dev) * 2, 1).reshape(B * N, T) * vs
    x = torch.full((B * N, ldx), 3.0, device=dev, dtype=torch.float16); x[:, :Cc] = (torch.randn(B * N, Cc, device=dev) * 0.03).half()
    w16, v16 = w.half(), v.half()
    y = torch.full((B * N, ldx), 7.0, device=dev, dtype=torch.float16)
    stats = torch.zeros(B, 2, device=dev, dtype=torch.float64)
    want_dbg = N <= 1600
    dbg = torch.zeros(B, N, N, device=dev) if want_dbg else None
    L.check(lib.cmpc_graph_reason_f16(w16.data_ptr(), v16.data_ptr(), x.data_ptr(), ldx, B, N, Cc, vs, y.data_ptr(), ldx, stats.data_ptr(),
                                      dbg.data_ptr() if want_dbg else None, st), "graph")
    torch.cuda.synchronize()
    lib.cmpc_graph_set_mode(0)
    Wf, Vf, Xf = w16.float().view(B, N, 32), v16.float().view(B, N, 32), x[:, :Cc].float().view(B, N, Cc)
    P = Wf @ Vf.transpose(1, 2)
    if want_dbg:
        assert ((dbg * vs) - P).abs().max() < 2e-3 * P.abs().max()
        assert ((dbg.sum(2) - (P / vs).sum(2)).abs().max()) < 1e-3                # rows of the adjacency
    Yref = (P.half().float() @ Xf) / vs
    assert (y[:, :Cc].float().view(B, N, Cc) - Yref).abs().max() < 2e-3 * Yref.abs().max() + 1e-6
    assert (y[:, Cc:_rup(Cc, 8)] == 0).all()
    s1, s2 = Yref.double().sum((1, 2)), (Yref.double() ** 2).sum((1, 2))
    assert ((stats[:, 1] - s2).abs() / s2).max() < 2e-3
    assert ((stats[:, 0] - s1).abs() / s1.abs().clamp_min(1e-3)).max() < 5e-2


def test_affinity_softmax_masks_and_normalises(env):
    L, lib, dev, st = env
    B, N, T = 3, 200, 20
    affi = torch.zeros(B * N, 32, device=dev); affi[:, :T] = torch.randn(B * N, T, device=dev) * 2
    mask = torch.ones(B, T, device=dev); mask[1, 7:] = 0; mask[2, 1:] = 0
    affi.view(B, N, 32)[1, :, 7:] = 0; affi.view(B, N, 32)[2, :, 1:] = 0            # masked words have R_t = 0 -> zero column
    w16 = torch.empty(B * N, 32, device=dev, dtype=torch.float16); v16 = torch.empty_like(w16)
    gw_w = torch.empty(B * N, T, device=dev); gw_v = torch.empty_like(gw_w)
    ws = torch.zeros(lib.cmpc_affinity_workspace_bytes(B), dtype=torch.uint8, device=dev)
    L.check(lib.cmpc_affinity_softmax(affi.data_ptr(), mask.data_ptr(), B, N, T, 256.0, w16.data_ptr(), v16.data_ptr(), gw_w.data_ptr(),
                                      gw_v.data_ptr(), ws.data_ptr(), ws.numel(), st), "affinity")
    torch.cuda.synchronize()
    a = affi.view(B, N, 32)[:, :, :T]
    m = mask.view(B, 1, T)
    rw = torch.softmax(m * a + (1 - m) * torch.finfo(torch.float32).min, dim=2)
    rv = m * torch.softmax(a, dim=1)
    assert (gw_w.view(B, N, T) - rw).abs().max() < 1e-5 and (gw_v.view(B, N, T) - rv).abs().max() < 1e-6
    assert (w16.view(B, N, 32)[:, :, T:] == 0).all() and (v16.view(B, N, 32)[:, :, T:] == 0).all()
    assert (v16.view(B, N, 32)[2, :, 1:T] == 0).all() and (w16.view(B, N, 32)[2, :, 1:T] == 0).all()
    assert ((v16.float().view(B, N, 32)[:, :, :T] / 256.0 - rv).abs() <= 1e-3 * rv + 1e-7).all()   # fp16 rounding of V * v_scale


@pytest.fixture(scope="module")
def full_size():
    """BASELINE config 2: batch 32 at 320x320 (N = 1600), reference initialisers; runs once for the property tests."""
    from cmpc_refseg_b200.CMPC_model import LSTM_model
    from cmpc_refseg_b200.synthetic import make_inputs
    dev = torch.device("cuda:0")
    model = LSTM_model(batch_size=32, mode="eval", device=dev, seed=0)
    inp = make_inputs(32, seed=1234, seq_len="unc")
    d = {k: inp[k].to(dev) for k in ("c3", "c4", "c5", "lstm_outputs", "target_fine")}
    out = model.forward(d["c3"], d["c4"], d["c5"], d["lstm_outputs"])
    torch.cuda.synchronize()
    return model, d, {k: v.clone() for k, v in out.items()}, inp


def test_full_size_properties(full_size):
    model, d, out, inp = full_size
    assert torch.isfinite(out["pred"]).all() and torch.isfinite(out["up"]).all()
    # word-type weights are a distribution on real words and exactly zero on padding
    parse = out["words_parse"]
    mask = out["seq_mask"]
    assert torch.allclose(parse.sum(3, keepdim=True), mask, atol=1e-5)
    sl = inp["seq_len"].to(mask.device)
    assert torch.equal(mask.view(32, 20).sum(1).long(), sl.long())
    # sigm = sigmoid(up); up is the x8 legacy-bilinear of pred: every 8th sample reproduces pred exactly
    assert (out["sigm"] - torch.sigmoid(out["up"])).abs().max() < 1e-6
    assert torch.equal(out["up"][:, ::8, ::8, :], out["pred"])
    # relation weights: rows of W sum to one, columns of V sum to the word mask; hence adjacency rows sum to one
    gw_w, gw_v = out["gw_w"], out["gw_v"]
    assert (gw_w.sum(2) - 1).abs().max() < 1e-4
    assert (gw_v.sum(1) - mask.view(32, 20)).abs().max() < 1e-4
    # integer I/U from the device kernel == torch on the same logits, bit for bit
    I, U = model.mIoU_counts(d["target_fine"])
    p = out["up"] > 0
    t = d["target_fine"] != 0
    assert torch.equal(I, (p & t).sum((1, 2, 3))) and torch.equal(U, (p | t).sum((1, 2, 3)))


def test_full_size_batch_invariance(full_size):
    """Sharding contract: a sample's logits do not depend on which batch / rank it is processed in.
    fp32 atomics (row sums of squares) are order-dependent at ~1e-7, which flips an occasional fp16 rounding upstream of the
    whole-map layer norms; measured run-to-run / batch-to-batch spread of the logits is ~1e-3, the same size as the
    fp16-operand error itself and 10x below the 1e-2 parity budget."""
    from cmpc_refseg_b200.CMPC_model import LSTM_model
    model, d, out, inp = full_size
    one = LSTM_model(batch_size=1, mode="eval", device=d["c3"].device, seed=0)
    for b in (0, 17, 31):
        o = one.forward(d["c3"][b:b + 1], d["c4"][b:b + 1], d["c5"][b:b + 1], d["lstm_outputs"][b:b + 1])
        torch.cuda.synchronize()
        assert (o["pred"] - out["pred"][b:b + 1]).abs().max() < 3e-3


@pytest.mark.parametrize("mode", ["constant", "reflect"])
def test_postprocess_resize_crop_iou(mode):
    """Threshold -> resize_and_crop to ragged ground-truth sizes -> (I, U) (trainval_model.py:243-245, 266) in one kernel vs
    the oracle's restated skimage resize: bit-exact masks and integer counts (up-scaling, down-scaling, both aspect orders,
    same size, a 1-pixel-high mask)."""
    from cmpc_refseg_b200.postprocess import postprocess
    from oracle.cmpc_head_ref import postprocess_iu
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(3)
    sizes = [(427, 640), (640, 480), (320, 320), (333, 500), (97, 211), (500, 375), (1, 50), (64, 64)]
    B, H, W = len(sizes), 320, 320
    # smooth random fields so that masks have real boundaries (blobs), plus exact zeros to exercise >= at the threshold
    low = torch.randn(B, 1, 10, 10, generator=g)
    up = torch.nn.functional.interpolate(low, size=(H, W), mode="bicubic", align_corners=False).reshape(B, H, W, 1).contiguous()
    up[:, ::7, ::5] = 1e-9
    rs = np.random.RandomState(5)
    gts = [rs.rand(h, w) > 0.6 for h, w in sizes]
    pred, I, U = postprocess(up.to(dev), gts, mode=mode)
    for b, gt in enumerate(gts):
        want, wi, wu = postprocess_iu(up[b].numpy(), gt, mode=mode)
        assert np.array_equal(pred[b].cpu().numpy() != 0, want != 0), (b, sizes[b])
        assert (int(I[b]), int(U[b])) == (wi, wu)
    _, I2, U2 = postprocess(up.to(dev), gts, mode=mode, return_masks=False)
    assert torch.equal(I, I2) and torch.equal(U, U2)
    with pytest.raises(Exception):
        postprocess(up.to(dev), gts[:-1])


@pytest.mark.parametrize("rows,rps,c,ldu,ldo,spatial,normalize", [
    (6400, 1600, 1000, 1024, 1024, (40, 40), 1),     # bench geometry: contiguous rows, one bulk copy per chunk
    (4916, 1229, 1000, 1024, 1024, (0, 0), 1),       # rows % 8 != 0 (partial last chunk), rows_per_sample odd, no spatial channels
    (4800, 1600, 1000, 1032, 1008, (40, 40), 0),     # strided input and output: per-row bulk copies; relu(LN) only
    (4096, 4096, 504, 512, 512, (64, 64), 1),        # narrow rows: register kernel whatever the mode
])
def test_ln_relu_l2norm_matches_torch(env, rows, rps, c, ldu, ldo, spatial, normalize):
    """l2_normalize_C(relu(LN(U))) (CMPC_model.py:370-372, :408) + appended spatial channels: the bulk-copy staged kernel (mode 0)
    and the register-file kernels (mode 1) against torch fp32 on the same fp16 input"""
    L, lib, dev, st = env
    B = rows // rps
    u = (torch.randn(rows, ldu, device=dev) * 2).half()
    gamma, beta = torch.rand(1024, device=dev) + 0.5, torch.randn(1024, device=dev) * 0.3
    mr = torch.stack([torch.randn(B, device=dev) * 0.2, torch.rand(B, device=dev) + 0.5], 1).contiguous()
    x = (u[:, :c].float() - mr[:, 0].repeat_interleave(rps)[:, None]) * mr[:, 1].repeat_interleave(rps)[:, None] * gamma[:c] + beta[:c]
    x = x.clamp_min(0)
    ss = (x * x).sum(1)
    ref = x * torch.rsqrt(ss.clamp_min(1e-12))[:, None] if normalize else x
    outs = []
    for mode in (0, 1):
        lib.cmpc_ln_relu_l2norm_set_mode(mode)
        out = torch.full((rows, ldo), 7.0, device=dev, dtype=torch.float16)
        rss = torch.zeros(rows, device=dev)
        L.check(lib.cmpc_ln_relu_l2norm_f16(u.data_ptr(), ldu, mr.data_ptr(), gamma.data_ptr(), beta.data_ptr(), out.data_ptr(), ldo, rows, c,
                                            spatial[0], spatial[1], rps, normalize, rss.data_ptr(), st), "ln_relu_l2norm")
        torch.cuda.synchronize()
        tol = 2e-3 if normalize else 2e-3 * float(ref.abs().max())
        assert (out[:, :c].float() - ref).abs().max() < tol
        assert ((rss - ss).abs() <= 1e-4 * ss + 1e-6).all()
        pad0 = c
        if spatial[0] > 0:
            from oracle.cmpc_head_ref import generate_spatial_batch
            sp = torch.from_numpy(generate_spatial_batch(1, spatial[0], spatial[1])).reshape(-1, 8).to(dev)
            assert torch.equal(out[:, c:c + 8].float(), sp.half().float().repeat(B, 1)[:rows])
            pad0 = c + 8
        assert (out[:, pad0:] == 0).all()
        outs.append(out)
    lib.cmpc_ln_relu_l2norm_set_mode(0)
    assert (outs[0].float() - outs[1].float()).abs().max() <= 1e-3          # same arithmetic up to the order of the row sum


def test_convlstm_gates2_from_gate_map(env):
    """cmpc_convlstm_gates2_y16 (o' recomputed from the fp16 gate map, util/cell.py:66-75) against gates1 -> fp32 o' -> gates2"""
    L, lib, dev, st = env
    B, N, Mm, GW = 2, 400, 500, 512
    M = B * N
    y = (torch.randn(M, 4 * GW, device=dev)).half()
    y.view(M, 4, GW)[:, :, Mm:] = 0
    mr_in = torch.stack([torch.randn(B, 4, device=dev) * 0.1, torch.rand(B, 4, device=dev) + 0.5], 2).contiguous()
    g, bt = torch.rand(5, GW, device=dev) + 0.5, torch.randn(5, GW, device=dev) * 0.2
    cprev, wco = torch.randn(M, GW, device=dev), torch.randn(N, GW, device=dev)
    res = []
    for lean in (False, True):
        cnew, opre = torch.zeros(M, GW, device=dev), torch.zeros(M, GW, device=dev)
        so = torch.zeros(B, 2, 2, device=dev, dtype=torch.float64)
        L.check(lib.cmpc_convlstm_gates1(y.data_ptr(), 1, 4 * GW, GW, Mm, mr_in.data_ptr(), g.data_ptr(), bt.data_ptr(), cprev.data_ptr(),
                                         wco.data_ptr(), cnew.data_ptr(), None if lean else opre.data_ptr(), so.data_ptr(), M, N, st), "gates1")
        mr = torch.empty(B, 2, 2, device=dev)
        L.check(lib.cmpc_ln_finalize(so.data_ptr(), 2 * B, float(N * Mm), mr.data_ptr(), st), "finalize")
        cst, h = torch.zeros(M, GW, device=dev), torch.zeros(M, GW, device=dev, dtype=torch.float16)
        if lean:
            L.check(lib.cmpc_convlstm_gates2_y16(y[:, 3 * GW:].data_ptr(), 4 * GW, wco.data_ptr(), cnew.data_ptr(), GW, Mm, mr.data_ptr(),
                                                 g.data_ptr(), bt.data_ptr(), cst.data_ptr(), h.data_ptr(), 0, M, N, st), "gates2_y16")
            h2 = torch.zeros_like(h)                      # last-step form: no cell state written
            L.check(lib.cmpc_convlstm_gates2_y16(y[:, 3 * GW:].data_ptr(), 4 * GW, wco.data_ptr(), cnew.data_ptr(), GW, Mm, mr.data_ptr(),
                                                 g.data_ptr(), bt.data_ptr(), None, h2.data_ptr(), 0, M, N, st), "gates2_y16")
            torch.cuda.synchronize()
            assert torch.equal(h, h2)
        else:
            L.check(lib.cmpc_convlstm_gates2(opre.data_ptr(), cnew.data_ptr(), GW, Mm, mr.data_ptr(), g.data_ptr(), bt.data_ptr(), cst.data_ptr(),
                                             h.data_ptr(), None, M, N, st), "gates2")
        torch.cuda.synchronize()
        res.append((cnew, cst, h, so.clone()))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][3], res[1][3])
    assert (res[0][1] - res[1][1]).abs().max() == 0
    assert (res[0][2].float() - res[1][2].float()).abs().max() <= 1e-3      # o' re-evaluated with the same fp32 expression (at most an fma contraction apart)
    assert (res[1][2][:, Mm:] == 0).all() and res[1][2].float().abs().max() > 0.1


def test_convlstm_fp16_state_path(env):
    """Inference form of the ConvLSTM gate passes with the cell state kept in fp16 (cmpc_convlstm_gates1_h16 + gates2_y16(state_f16=1))
    against the fp32-state passes: same statistics (taken from the fp32 values), state and h equal up to one fp16 rounding of c' / c."""
    L, lib, dev, st = env
    B, N, Mm, GW = 2, 400, 500, 512
    M = B * N
    y = (torch.randn(M, 4 * GW, device=dev)).half()
    y.view(M, 4, GW)[:, :, Mm:] = 0
    mr_in = torch.stack([torch.randn(B, 4, device=dev) * 0.1, torch.rand(B, 4, device=dev) + 0.5], 2).contiguous()
    g, bt = torch.rand(5, GW, device=dev) + 0.5, torch.randn(5, GW, device=dev) * 0.2
    cprev32, wco = torch.randn(M, GW, device=dev), torch.randn(N, GW, device=dev)
    cprev16 = cprev32.half()
    cprev32 = cprev16.float()                                  # both paths start from the same (fp16-representable) state
    # fp32 state
    cnew32, so32 = torch.zeros(M, GW, device=dev), torch.zeros(B, 2, 2, device=dev, dtype=torch.float64)
    L.check(lib.cmpc_convlstm_gates1(y.data_ptr(), 1, 4 * GW, GW, Mm, mr_in.data_ptr(), g.data_ptr(), bt.data_ptr(), cprev32.data_ptr(),
                                     wco.data_ptr(), cnew32.data_ptr(), None, so32.data_ptr(), M, N, st), "gates1")
    # fp16 state
    cnew16, so16 = torch.zeros(M, GW, device=dev, dtype=torch.float16), torch.zeros(B, 2, 2, device=dev, dtype=torch.float64)
    L.check(lib.cmpc_convlstm_gates1_h16(y.data_ptr(), 4 * GW, GW, Mm, mr_in.data_ptr(), g.data_ptr(), bt.data_ptr(), cprev16.data_ptr(),
                                         wco.data_ptr(), cnew16.data_ptr(), so16.data_ptr(), M, N, st), "gates1_h16")
    torch.cuda.synchronize()
    assert torch.equal(so32, so16)                             # statistics come from the fp32 values in both forms
    assert torch.equal(cnew32.half(), cnew16)
    mr = torch.empty(B, 2, 2, device=dev)
    L.check(lib.cmpc_ln_finalize(so32.data_ptr(), 2 * B, float(N * Mm), mr.data_ptr(), st), "finalize")
    c32, h32 = torch.zeros(M, GW, device=dev), torch.zeros(M, GW, device=dev, dtype=torch.float16)
    c16, h16 = torch.zeros(M, GW, device=dev, dtype=torch.float16), torch.zeros(M, GW, device=dev, dtype=torch.float16)
    L.check(lib.cmpc_convlstm_gates2_y16(y[:, 3 * GW:].data_ptr(), 4 * GW, wco.data_ptr(), cnew32.data_ptr(), GW, Mm, mr.data_ptr(), g.data_ptr(),
                                         bt.data_ptr(), c32.data_ptr(), h32.data_ptr(), 0, M, N, st), "gates2_y16")
    L.check(lib.cmpc_convlstm_gates2_y16(y[:, 3 * GW:].data_ptr(), 4 * GW, wco.data_ptr(), cnew16.data_ptr(), GW, Mm, mr.data_ptr(), g.data_ptr(),
                                         bt.data_ptr(), c16.data_ptr(), h16.data_ptr(), 1, M, N, st), "gates2_y16")
    torch.cuda.synchronize()
    scale = float(c32.abs().max())
    assert (c32 - c16.float()).abs().max() <= 2e-3 * scale      # c' rounded to fp16 before its layer norm, c after it
    assert (h32.float() - h16.float()).abs().max() <= 4e-3
    assert (c16[:, Mm:] == 0).all() and (h16[:, Mm:] == 0).all()
