"""Variable I/O for the head: the reference stores its weights as TF-1 checkpoints written by ``tf.train.Saver``
(trainval_model.py:46-63, 135-142, 185-190) under the variable scope ``text_objseg/`` (CMPC_model.py:83); the head's variables are
the 206 tensors of SURVEY App. B.  ``load_variables`` opens such a checkpoint DIRECTLY -- ``load_variables('.../model.ckpt-700000', shapes)``
reads the V2 tensor bundle (``.index`` + ``.data-0000x-of-0000y``) with the pure-Python reader in ``tf_bundle.py``, no TensorFlow
needed -- or a NumPy ``.npz`` keyed by the TF variable names.  It maps the variables onto the ``params`` dict the drop-in
``LSTM_model`` takes (scope and ``:0`` stripped, shapes checked, the three word-encoder variables passed through, variables of the
backbone / optimizer slots (``.../Adam``, ``.../Adam_1``, ``beta1_power`` ...) ignored).  ``save_variables`` writes the same two
formats back (a bundle written here restores through ``tf.train.Saver`` on the reference side).
``HeadTrainer.state_dict`` / ``load_state_dict`` (train.py) snapshot a training run.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch

TF_SCOPE = "text_objseg/"
ENCODER_VARIABLES = ("Variable", "rnn/lstm_cell/kernel", "rnn/lstm_cell/bias")


def _strip(name: str) -> str:
    name = name[:-2] if name.endswith(":0") else name
    return name[len(TF_SCOPE):] if name.startswith(TF_SCOPE) else name


class _nullctx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def load_variables(path: str, shapes: Dict[str, Tuple[int, ...]], *, strict: bool = True, verify: bool = False) -> Dict[str, torch.Tensor]:
    """path: a TF checkpoint prefix (V2 bundle) or an .npz keyed by TF variable names; shapes: cmpc_refseg_b200.CMPC_model.head_param_shapes(...).
    Returns {name below text_objseg/: float32 tensor}.  strict: every head variable must be present."""
    out, ignored = {}, []
    path = str(path)
    if path.endswith(".npz"):
        src = np.load(path)
        files = src.files
    else:                                          # a tf.train.Saver checkpoint prefix, e.g. .../model.ckpt-700000
        from .tf_bundle import read_bundle, read_index
        entries, _ = read_index(path)
        keep = [n for n in entries if _strip(n) in shapes or _strip(n) in ENCODER_VARIABLES]
        ignored += [n for n in entries if n not in keep]
        src = read_bundle(path, keep, verify=verify)
        files = keep
    with (src if hasattr(src, "__exit__") else _nullctx()) as z:
        z = src
        for raw in files:
            name = _strip(raw)
            if name in ENCODER_VARIABLES:          # word encoder (CMPC_model.py:144-157): optional, shapes checked by WordEncoderB200
                out[name] = torch.from_numpy(np.asarray(z[raw]).astype(np.float32, copy=True))
                continue
            if name not in shapes:
                ignored.append(raw)
                continue
            a = np.asarray(z[raw])
            if tuple(a.shape) != tuple(shapes[name]):
                raise ValueError(f"{raw}: shape {tuple(a.shape)} in the checkpoint, {tuple(shapes[name])} expected by the head")
            out[name] = torch.from_numpy(a.astype(np.float32, copy=True))
    missing = sorted(set(shapes) - set(out) - set(ENCODER_VARIABLES))
    if strict and missing:
        raise KeyError(f"{len(missing)} head variables missing from {path}, e.g. {missing[:3]}")
    load_variables.last_ignored = ignored          # backbone / LSTM / optimizer slots etc.
    return out


def save_variables(path: str, params: Dict[str, torch.Tensor], global_step: int = None) -> None:
    """.npz, or (any other path) a V2 tensor bundle `<path>.index` / `<path>.data-00000-of-00001` under scope text_objseg/"""
    arrs = {TF_SCOPE + k: v.detach().to("cpu", torch.float32).numpy() for k, v in params.items()}
    if str(path).endswith(".npz"):
        np.savez(path, **arrs)
        return
    from .tf_bundle import write_bundle
    if global_step is not None:
        arrs[TF_SCOPE + "Variable_1"] = np.array(global_step, np.int32)      # train_step, CMPC_model.py:450 (second tf.Variable of the scope)
    write_bundle(str(path), arrs)
