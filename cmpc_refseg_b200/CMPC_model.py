"""Drop-in for the head path of the reference's ``CMPC_model.LSTM_model`` (CMPC_model.py:13-142).

Same constructor keywords and defaults (CMPC_model.py:15-40), same public attribute names
(``pred``, ``up``, ``sigm``, ``up_c3/4/5``, ``words_parse``, ``gw_w``, ``gw_v``, ``seq_mask``, hparams ``H``, ``W``,
``num_steps`` ...).  The TF placeholders + ``sess.run(feed_dict)`` of the reference become one call::

    model = LSTM_model(batch_size=32, mode='eval')
    pred, up, sigm = model.run([model.FETCH_PRED, model.FETCH_UP, model.FETCH_SIGM],
                               feed_dict=dict(visual_feat_c3=c3, visual_feat_c4=c4, visual_feat_c5=c5,
                                              lstm_outputs=words, seq_len=seq_len))

or simply ``model.forward(c3, c4, c5, lstm_outputs, seq_len)``.  The backbone (deeplab taps, :73-76) and the word
LSTM recurrence (:144-157) are upstream producers and out of scope: their outputs are the inputs here, exactly as
``CMPCv4_BERT_model.py:80-83,119-121`` already feeds word features through placeholders.

All arithmetic runs in the sm_100a kernels of libcmpc_b200 (C ABI, include/cmpc_b200.h).  No CPU fallback.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from .head import on_device

from . import _lib as L
from .head import CMPCHeadB200
from .methods import ReferenceMethods
from .weights import EXG, LEVELS
from .word_encoder import BIAS, EMB, KERNEL, WordEncoderB200

ENCODER_VARIABLES = (EMB, KERNEL, BIAS)


def reference_init(shapes: Dict[str, tuple], seed: int = 0) -> Dict[str, torch.Tensor]:
    """The reference's initialisers: xavier-uniform conv kernels and zero biases (CMPC_model.py:414-416),
    TF-default glorot-uniform for the ConvLSTM kernel / peepholes (util/cell.py:42,49-50,62), LN gamma=1, beta=0."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in shapes.items():
        leaf = name.rsplit("/", 1)[1]
        if leaf in ("biases", "beta"):
            t = torch.zeros(shape)
        elif leaf == "gamma":
            t = torch.ones(shape)
        else:
            if len(shape) == 2:
                fi, fo = shape
            else:
                rf = 1
                for s in shape[:-2]:
                    rf *= s
                fi, fo = shape[-2] * rf, shape[-1] * rf
            lim = math.sqrt(6.0 / (fi + fo))
            t = ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * lim).float()
        out[name] = t
    return out


def head_param_shapes(*, vf_h, vf_w, vf_dim, v_emb_dim, rnn_size, mlp_dim, c4_dim=1024, c3_dim=512, parse_hidden=500):
    """TF variable names / shapes of the head under scope text_objseg/ (SURVEY App. B)."""
    C, R, M = v_emb_dim, rnn_size, mlp_dim
    s: Dict[str, tuple] = {}

    def conv(name, k, cin, cout):
        s[name + "/DW"] = (k, k, cin, cout)
        s[name + "/biases"] = (cout,)

    conv("c5_lateral", 1, vf_dim, C); conv("c4_lateral", 1, c4_dim, C); conv("c3_lateral", 1, c3_dim, C)
    conv("words_parse_1", 1, R, parse_hidden); conv("words_parse_2", 1, parse_hidden, 4)
    for lvl in LEVELS:
        for k in range(1, 6):
            conv(f"vis_trans_{lvl}_head{k}", 1, C + 8, C)
            conv(f"lang_trans_{lvl}_head{k}", 1, R, C)
        conv(f"words_trans_{lvl}", 1, R, R)
        conv(f"spa_graph_trans2_{lvl}", 1, C, C)
        conv(f"gconv_update_spa_graph_{lvl}", 1, C, C)
        for ln in ("gconv_feat_ln_spa_graph", "gconv_update_ln_spa_graph"):
            s[f"{ln}_{lvl}/beta"] = (C,)
            s[f"{ln}_{lvl}/gamma"] = (C,)
        conv(f"fusion_{lvl}", 1, 2 * C + R + 8, M)
        conv(f"score_{lvl}", 3, M, 1)
    conv("score", 3, M, 1)
    for x in EXG:
        conv(f"spa_graph_key_{x}gv_f1", 1, M, M)
        conv(f"lang_query_{x}gv_f1", 1, R, M)
        conv(f"gv_lang_{x}gv_f1", 1, M + R, M)
        for f in ("_f1", "_f2"):
            conv(f"lang_feat_{x}{f}", 1, M, M)
            conv(f"trans_feat_{x}{f}", 1, M, M)
    s["rnn/conv_lstm_cell/kernel"] = (1, 1, 2 * M, 4 * M)
    for p in ("W_ci", "W_cf", "W_co"):
        s[f"rnn/conv_lstm_cell/{p}"] = (vf_h, vf_w, M)
    for i in range(5):
        nm = "LayerNorm" if i == 0 else f"LayerNorm_{i}"
        s[f"rnn/conv_lstm_cell/{nm}/beta"] = (M,)
        s[f"rnn/conv_lstm_cell/{nm}/gamma"] = (M,)
    return s


class LSTM_model(ReferenceMethods):
    FETCH_PRED, FETCH_UP, FETCH_SIGM = "pred", "up", "sigm"

    def __init__(self, batch_size=1,
                 num_steps=20,
                 vf_h=40,
                 vf_w=40,
                 H=320,
                 W=320,
                 vf_dim=2048,
                 vocab_size=12112,
                 w_emb_dim=1000,
                 v_emb_dim=1000,
                 mlp_dim=500,
                 start_lr=0.00025,
                 lr_decay_step=800000,
                 lr_decay_rate=1.0,
                 rnn_size=1000,
                 keep_prob_rnn=1.0,
                 keep_prob_emb=1.0,
                 keep_prob_mlp=1.0,
                 num_rnn_layers=1,
                 optimizer='adam',
                 weight_decay=0.0005,
                 mode='eval',
                 conv5=False,
                 glove_dim=300,
                 emb_name='Gref',
                 emb_dir='data',
                 *, params: Optional[Dict[str, torch.Tensor]] = None, device=None, seed: int = 0,
                 head_kwargs: Optional[dict] = None, cuda_graph: bool = False, gv_norm: Optional[str] = None):
        # hyper-parameters, stored under the reference's attribute names (CMPC_model.py:41-65)
        self.batch_size = batch_size
        self.num_steps = num_steps
        self.vf_h = vf_h
        self.vf_w = vf_w
        self.H = H
        self.W = W
        self.vf_dim = vf_dim
        self.start_lr = start_lr
        self.lr_decay_step = lr_decay_step
        self.lr_decay_rate = lr_decay_rate
        self.vocab_size = vocab_size
        self.w_emb_dim = w_emb_dim
        self.v_emb_dim = v_emb_dim
        self.glove_dim = glove_dim
        self.emb_name = emb_name
        self.mlp_dim = mlp_dim
        self.rnn_size = rnn_size
        self.keep_prob_rnn = keep_prob_rnn
        self.keep_prob_emb = keep_prob_emb
        self.keep_prob_mlp = keep_prob_mlp
        self.num_rnn_layers = num_rnn_layers
        self.optimizer = optimizer
        self.weight_decay = weight_decay
        self.mode = mode
        self.conv5 = conv5
        if optimizer != 'adam':
            raise ValueError("Unknown optimizer type %s!" % optimizer)       # CMPC_model.py:458
        if mode not in ('eval', 'train'):
            raise ValueError("mode must be 'eval' or 'train' (CMPC_model.py:84-87)")
        self.device = torch.device(device if device is not None else "cuda:0")
        # tf.nn.l2_normalize(gv_lang) has no axis (:241): at batch > 1 the literal graph couples the samples of a batch.  Training
        # reproduces that (the reference trains with -bs 8, trainval.sh); inference defaults to the per-sample form, i.e. the graph
        # the reference's own test drivers run (batch 1), which is also what makes a batch shard independent of its neighbours.
        self.gv_norm = gv_norm if gv_norm is not None else ("batch" if mode == 'train' else "sample")
        # the reference hard-codes c4 = 1024, c3 = 512 channels and a 500-wide parser (CMPC_model.py:110,112,349);
        # head_kwargs (c4_dim, c3_dim, parse_hidden) only exists so that tests can run scaled-down heads
        hk = dict(head_kwargs or {})
        if params is None:
            params = reference_init(head_param_shapes(vf_h=vf_h, vf_w=vf_w, vf_dim=vf_dim, v_emb_dim=v_emb_dim,
                                                      rnn_size=rnn_size, mlp_dim=mlp_dim, **hk), seed)
        # the word-encoder variables (lstm(), :144-157) are optional and frozen here; everything else belongs to the head
        self.encoder_params = {k: params[k] for k in ENCODER_VARIABLES if k in params}
        params = {k: v for k, v in params.items() if k not in ENCODER_VARIABLES}
        self.params = params
        self.cuda_graph = cuda_graph      # replay the pass from a CUDA graph (same input buffers every call)
        self._head = CMPCHeadB200(params, batch_size=batch_size, num_steps=num_steps, vf_h=vf_h, vf_w=vf_w, H=H, W=W,
                                  vf_dim=vf_dim, v_emb_dim=v_emb_dim, rnn_size=rnn_size, mlp_dim=mlp_dim,
                                  device=self.device, gv_norm=self.gv_norm, **hk)
        # "placeholders": set by forward()/run(); outputs: populated after each forward
        self.visual_feat_c3 = self.visual_feat_c4 = self.visual_feat_c5 = None
        self.lstm_outputs = None
        self.seq_len = None
        self.target_fine = None
        self.pred = self.up = self.sigm = None
        self.up_c3 = self.up_c4 = self.up_c5 = None
        self.words_parse = self.seq_mask = self.gw_w = self.gw_v = None

    # ---- the hot path -------------------------------------------------------------------------------------------
    @on_device
    def build_graph(self, aux: bool = False):
        """CMPC_model.py:89-142 on the currently fed inputs; populates pred / up / sigm and the aux attributes."""
        if self.visual_feat_c5 is None or self.lstm_outputs is None:
            raise L.CmpcError("feed visual_feat_c3/c4/c5 and lstm_outputs first (forward() or run(feed_dict=...))")
        fwd = self._head.forward_graphed if self.cuda_graph else self._head.forward
        out = fwd(self.visual_feat_c3, self.visual_feat_c4, self.visual_feat_c5, self.lstm_outputs, self.seq_len, aux=aux)
        for k in ("pred", "up", "sigm", "words_parse", "seq_mask", "gw_w", "gw_v"):
            setattr(self, k, out[k])
        if aux:
            self.up_c3, self.up_c4, self.up_c5 = out["up_c3"], out["up_c4"], out["up_c5"]
        self._out = out
        return out

    def forward(self, c3, c4, c5, lstm_outputs=None, seq_len=None, target_fine=None, aux: bool = False, words=None):
        """Either lstm_outputs (the dynamic_rnn outputs) or words [B, T] + seq_len [B] (the reference's placeholders, :67-71; needs the
        word-encoder variables `Variable`, `rnn/lstm_cell/kernel`, `rnn/lstm_cell/bias` among params)."""
        if lstm_outputs is None:
            if words is None or seq_len is None:
                raise L.CmpcError("feed lstm_outputs, or words and seq_len")
            lstm_outputs = self.encode_words(words, seq_len)
        self.visual_feat_c3, self.visual_feat_c4, self.visual_feat_c5 = c3, c4, c5
        self.lstm_outputs, self.seq_len, self.target_fine = lstm_outputs, seq_len, target_fine
        return self.build_graph(aux=aux)

    @on_device
    def encode_words(self, words, seq_len):
        """embedding lookup + word LSTM of lstm() (:144-157) on the device -> lstm_outputs [B, T, rnn_size]"""
        if getattr(self, "_encoder", None) is None:
            if len(self.encoder_params) != len(ENCODER_VARIABLES):
                raise L.CmpcError("the word encoder needs `Variable`, `rnn/lstm_cell/kernel` and `rnn/lstm_cell/bias` among params")
            self._encoder = WordEncoderB200(self._head, self.encoder_params)
        self.words = words
        return self._encoder.forward(words, seq_len)

    __call__ = forward

    def run(self, fetches, feed_dict):
        """sess.run-style entry (trainval_model.py:232, test.py:286): fetches are attribute names."""
        for k, v in feed_dict.items():
            if not hasattr(self, k):
                raise KeyError(f"unknown placeholder {k!r}")
            setattr(self, k, v)
        names = [fetches] if isinstance(fetches, str) else list(fetches)
        self.build_graph(aux=any(n.startswith("up_c") for n in names))
        vals = [getattr(self, n) for n in names]
        return vals[0] if isinstance(fetches, str) else vals

    @on_device
    def mIoU_counts(self, target_fine=None):
        """Per-sample integer (I, U) of (up > 0) vs target (CMPC_model.py:486-489)."""
        t = target_fine if target_fine is not None else self.target_fine
        return self._head.mask_iu(self.up, t)

    @on_device
    def postprocess(self, gt_masks, score_thresh: float = 1e-9, mode: str = "constant", return_masks: bool = True):
        """trainval_model.py:243-245, 266 on the last `up`: threshold, resize_and_crop to each ground-truth size, (I, U)."""
        from .postprocess import postprocess
        return postprocess(self.up, gt_masks, score_thresh=score_thresh, mode=mode, return_masks=return_masks)

    @on_device
    def losses(self, target_fine=None):
        """Forward value of the training objective (CMPC_model.py:439-447): the four sigmoid-CE terms (sum over pixels,
        mean over the batch, util/loss.py:6-16), their 0.7/0.1/0.1/0.1 combination, the L2 regulariser over every `DW`
        (util/loss.py:28-32) and the total cost.  Needs the aux heads: call forward(..., aux=True) first."""
        t = target_fine if target_fine is not None else self.target_fine
        if self.up is None or self.up_c3 is None:
            raise L.CmpcError("losses() needs forward(..., aux=True) (up_c3/up_c4/up_c5 feed three of the four terms)")
        h = self._head
        self.cls_loss = h.ce_sums(self.up, t).mean()
        self.cls_loss_c5 = h.ce_sums(self.up_c5, t).mean()
        self.cls_loss_c4 = h.ce_sums(self.up_c4, t).mean()
        self.cls_loss_c3 = h.ce_sums(self.up_c3, t).mean()
        self.cls_loss_all = 0.7 * self.cls_loss + 0.1 * self.cls_loss_c5 + 0.1 * self.cls_loss_c4 + 0.1 * self.cls_loss_c3
        if self.mode != 'eval' or not hasattr(self, "_l2"):      # weights are constant in eval mode only: cache sum ||DW||^2 / 2 there
            self._l2 = sum(float((v.double() ** 2).sum()) / 2 for k, v in self.params.items() if k.endswith("/DW"))
        self.reg_loss = self.weight_decay * self._l2
        self.cost = self.cls_loss_all + self.reg_loss
        return dict(cls_loss=self.cls_loss, cls_loss_c5=self.cls_loss_c5, cls_loss_c4=self.cls_loss_c4,
                    cls_loss_c3=self.cls_loss_c3, cls_loss_all=self.cls_loss_all, reg_loss=self.reg_loss, cost=self.cost)

    @on_device
    def train_op(self, process_group=None, reduce_groups=None):
        """CMPC_model.py:426-478: sets up the objective, the polynomial learning-rate decay and Adam (cmpc_refseg_b200/train.py).  The
        TF `train` / `train_step` / `learning_rate` / `cls_loss*` fetches become `train_step(...)` and the attributes it refreshes."""
        from .train import HeadTrainer
        if self.mode != 'train':
            raise L.CmpcError("train_op() needs LSTM_model(mode='train')")
        enc = None
        if len(self.encoder_params) == len(ENCODER_VARIABLES):      # the reference trains the embedding and the word LSTM too (:426-431)
            if getattr(self, "_encoder", None) is None:
                self._encoder = WordEncoderB200(self._head, self.encoder_params)
            enc = self._encoder
        self._trainer = HeadTrainer(self._head, start_lr=self.start_lr, lr_decay_step=self.lr_decay_step, weight_decay=self.weight_decay,
                                    process_group=process_group, encoder=enc, reduce_groups=reduce_groups)
        self.params = {k: v for k, v in self._trainer.params.items() if k not in ENCODER_VARIABLES}
        self.encoder_params = {k: v for k, v in self._trainer.params.items() if k in ENCODER_VARIABLES}
        self.train_step = 0
        return self._trainer

    @on_device
    def train(self, c3, c4, c5, lstm_outputs, target_fine, seq_len=None, words=None):
        """One optimizer step on a batch (what `sess.run([model.train, ...])` does at trainval_model.py:98-107); refreshes cls_loss,
        cls_loss_c3/4/5, cls_loss_all, learning_rate, train_step, pred / up / sigm.  With lstm_outputs=None and words / seq_len
        (and the word-encoder variables among params) the embedding and the word LSTM are trained as well."""
        if getattr(self, "_trainer", None) is None:
            self.train_op()
        out = self._trainer.train_step(c3, c4, c5, lstm_outputs, target_fine, seq_len, words=words)
        for k in ("pred", "up", "sigm", "words_parse", "seq_mask", "gw_w", "gw_v"):
            setattr(self, k, out[k])
        self.up_c3, self.up_c4, self.up_c5 = out["up_c3"], out["up_c4"], out["up_c5"]
        for k, v in self._trainer.last.items():
            setattr(self, k, v)
        self.train_step = self._trainer.step
        self.target_fine = target_fine
        return self._trainer.last
