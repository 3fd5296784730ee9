"""Variable I/O for the head: the reference stores its weights as TF-1 checkpoints written by ``tf.train.Saver``
(trainval_model.py:46-63, 135-142, 185-190) under the variable scope ``text_objseg/`` (CMPC_model.py:83); the head's variables are
the 206 tensors of SURVEY App. B.  TensorFlow is not importable next to this package, so the exchange format is a NumPy ``.npz``
keyed by the TF variable names -- produced on the reference side with four lines::

    reader = tf.train.load_checkpoint(ckpt_path)                     # TF >= 1.13
    np.savez(out_path, **{n: reader.get_tensor(n) for n in reader.get_variable_to_shape_map()
                          if n.startswith('text_objseg/') and '/Adam' not in n})

``load_variables`` maps such a file onto the ``params`` dict the drop-in ``LSTM_model`` takes (scope and ``:0`` stripped, shapes
checked, the three word-encoder variables passed through, variables of the backbone / optimizer slots ignored); ``save_variables`` writes one back (to be assigned
with ``tf.assign`` on the reference side).  ``HeadTrainer.state_dict`` / ``load_state_dict`` (train.py) snapshot a training run.
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


def load_variables(path: str, shapes: Dict[str, Tuple[int, ...]], *, strict: bool = True) -> Dict[str, torch.Tensor]:
    """path: .npz keyed by TF variable names; shapes: cmpc_refseg_b200.CMPC_model.head_param_shapes(...).
    Returns {name below text_objseg/: float32 tensor}.  strict: every head variable must be present."""
    out, ignored = {}, []
    with np.load(path) as z:
        for raw in z.files:
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


def save_variables(path: str, params: Dict[str, torch.Tensor]) -> None:
    np.savez(path, **{TF_SCOPE + k: v.detach().to("cpu", torch.float32).numpy() for k, v in params.items()})
