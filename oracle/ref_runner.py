"""Runs the reference's OWN ``CMPC_model.py`` (unmodified, imported from the read-only checkout) through the eager
``tensorflow`` stand-in of ``oracle/tfshim`` -- TEST INFRASTRUCTURE ONLY.

This is what pins the oracle: ``LSTM_model.__init__`` -> ``build_graph`` (CMPC_model.py:89-142) -> ``train_op`` (:426-492)
execute line by line as the reference wrote them, with ``util/cell.py``, ``util/loss.py`` and ``util/processing_tools.py``
imported from the reference as well; only the ``tf.*`` ops underneath are restated (TensorFlow is third-party and cannot be
installed here).  The modules the head never calls into (``deeplab_resnet`` backbone, ``util.data_reader``,
``util.im_processing``, ``util.text_processing``, ``util.eval_tools``: skimage / nltk / pyximport imports) are stubbed.

Used by ``tests/golden/make_ref_golden.py`` (writes the committed fixtures) and, when the checkout is present, by the live
tests in ``tests/test_reference_pin.py``.  The reference checkout does not exist on the GPU box; nothing there calls this.
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys
import tempfile
import types
from pathlib import Path
from typing import Dict, Optional

import numpy as np
import torch

SHIM_DIR = Path(__file__).resolve().parent / "tfshim"
REF_ROOT = Path(os.environ.get("CMPC_REFERENCE_ROOT", "/root/reference"))
_STUBBED = ("util.data_reader", "util.im_processing", "util.text_processing", "util.eval_tools")
_OURS = ("tensorflow", "deeplab_resnet", "deeplab_resnet.model", "util", "util.cell", "util.loss", "util.processing_tools",
         "util.functions", "_ref_CMPC_model") + _STUBBED


def reference_available() -> bool:
    return (REF_ROOT / "CMPC_model.py").is_file() and (REF_ROOT / "util" / "cell.py").is_file()


@contextlib.contextmanager
def _reference_imports():
    """sys.path / sys.modules set up so that `import tensorflow`, `from deeplab_resnet import model`, `from util...`
    inside the reference resolve to the shim, the stub and the reference's own util package; undone on exit."""
    saved_mods = {k: sys.modules.get(k) for k in _OURS}
    saved_path = list(sys.path)
    saved_flag = sys.dont_write_bytecode
    sys.dont_write_bytecode = True           # the checkout is read-only
    for k in _OURS:
        sys.modules.pop(k, None)
    sys.path[:0] = [str(SHIM_DIR), str(REF_ROOT)]
    for name in _STUBBED:
        sys.modules[name] = types.ModuleType(name)
    try:
        import tensorflow as tf               # the shim
        assert getattr(tf, "__version__", "").endswith("-shim"), "a real tensorflow shadowed the stand-in"
        spec = importlib.util.spec_from_file_location("_ref_CMPC_model", str(REF_ROOT / "CMPC_model.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["_ref_CMPC_model"] = mod
        yield tf, spec, mod
    finally:
        sys.path[:] = saved_path
        sys.dont_write_bytecode = saved_flag
        for k in _OURS:
            sys.modules.pop(k, None)
            if saved_mods[k] is not None:
                sys.modules[k] = saved_mods[k]


def run_reference(model_kwargs: Dict, params: Optional[Dict[str, torch.Tensor]], c3, c4, c5, *, lstm_outputs=None,
                  words=None, seq_len=None, embedding=None, target_fine=None, float64: bool = False,
                  init_seed: int = 0, quiet: bool = True) -> Dict:
    """Constructs the reference's ``LSTM_model(**model_kwargs)`` (constructing it executes the graph eagerly) and returns
    its public attributes as torch tensors.

    params: initial values by TF variable name below ``text_objseg/`` (``"c5_lateral/DW"`` ...); variables not listed take
            the initializer the reference names (xavier / zeros / glorot default).
    lstm_outputs: when given, stands for the output of ``dynamic_rnn`` over the word LSTM (CMPC_model.py:153-156) -- the
            boundary of the head; otherwise ``words`` [B,T] + ``seq_len`` [B] + ``embedding`` run the reference's ``lstm()``.
    model_kwargs['mode'] != 'eval' also runs ``train_op`` (needs ``target_fine``): one optimizer step is applied, the
            gradients are returned as computed (``raw_grads``) and as applied (bias gradients doubled, ``applied_grads``).
    """
    kw = dict(model_kwargs)
    B, T = kw.get("batch_size", 1), kw.get("num_steps", 20)
    dt = torch.float64 if float64 else torch.float32
    with _reference_imports() as (tf, spec, mod), tempfile.TemporaryDirectory() as tmp:
        st = tf._shim
        st.reset()
        st.float_dtype = dt
        st.init_seed = init_seed
        st.params = params
        st.requires_grad = kw.get("mode", "eval") != "eval"
        st.layers = {"res5c_relu": tf.Tensor(torch.as_tensor(c5).to(dt)), "res4b22_relu": tf.Tensor(torch.as_tensor(c4).to(dt)),
                     "res3b3_relu": tf.Tensor(torch.as_tensor(c3).to(dt))}
        if lstm_outputs is not None:
            st.rnn_outputs_feed = torch.as_tensor(lstm_outputs).to(dt)
            emb = np.zeros((4, kw.get("glove_dim", 300)), np.float32) if embedding is None else np.asarray(embedding)
            words_feed = torch.zeros(B, T, dtype=torch.int32) if words is None else torch.as_tensor(words).to(torch.int32)
        else:
            assert words is not None and seq_len is not None and embedding is not None
            emb = embedding.detach().cpu().numpy() if isinstance(embedding, torch.Tensor) else np.asarray(embedding)
            words_feed = torch.as_tensor(words).to(torch.int32)
        sl = torch.full((B,), T, dtype=torch.int32) if seq_len is None else torch.as_tensor(seq_len).to(torch.int32)
        # placeholders in creation order (CMPC_model.py:67-71): words, im, target_fine, valid_idx, seq_len
        st.placeholder_feeds = [words_feed, None, None if target_fine is None else torch.as_tensor(target_fine).to(dt), None, sl]
        np.save(os.path.join(tmp, "Shim_emb.npy"), emb.astype(np.float64 if float64 else np.float32))
        kw.update(emb_name="Shim", emb_dir=tmp)
        sink = io.StringIO()
        with (contextlib.redirect_stdout(sink) if quiet else contextlib.nullcontext()):
            spec.loader.exec_module(mod)                 # the reference's module, unmodified
            model = mod.LSTM_model(**kw)                 # == build graph + sess.run
        out: Dict = {}
        for name in ("pred", "up", "sigm", "up_c3", "up_c4", "up_c5", "words_parse", "gw_w", "gw_v", "seq_mask",
                     "target", "cls_loss", "cls_loss_c3", "cls_loss_c4", "cls_loss_c5", "cls_loss_all", "reg_loss", "cost",
                     "learning_rate", "mIoU", "train_step"):
            if hasattr(model, name):
                out[name] = getattr(model, name)._v.detach().clone()
        out["variables"] = {tf._strip_root(k): v._v.detach().clone() for k, v in st.variables.items()}
        out["variable_order"] = [tf._strip_root(v._name) for v in st.var_order]
        out["trainable"] = [tf._strip_root(v._name) for v in st.var_order if v.trainable]
        if st.optimizers:
            opt = st.optimizers[-1]
            out["raw_grads"] = {tf._strip_root(v._name): (None if g is None else g._v.detach().clone()) for g, v in opt.raw_grads}
            out["applied_grads"] = {tf._strip_root(v._name): (None if g is None else g._v.detach().clone()) for g, v in opt.applied}
        out["log"] = sink.getvalue()
        return out
