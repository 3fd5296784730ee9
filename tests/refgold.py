"""Loader for the reference-executed fixtures tests/golden/ref_*.npz (written by tests/golden/make_ref_golden.py from the
reference's own CMPC_model.py run through oracle/tfshim).  Inputs / parameters are regenerated from their seeds and checked
against the float64 checksums stored in the fixture before anything is compared."""
import importlib.util
from pathlib import Path

import numpy as np
import torch

GOLD = Path(__file__).parent / "golden"
_spec = importlib.util.spec_from_file_location("_make_ref_golden", GOLD / "make_ref_golden.py")
mk = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mk)


def load(name):
    z = np.load(GOLD / (name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def _verify(fix, inp, params, name):
    cs = mk.checksums(inp, params)
    for k, v in cs.items():
        ref = float(fix[k])
        assert abs(v - ref) <= 1e-9 * max(1.0, abs(ref)), f"{name}: regenerated {k[3:]} does not reproduce the fixture's checksum"


def forward_case(name, dtype=torch.float64):
    """-> (model_kw incl. c4/c3/parse dims, B, params(dtype), inputs(fp32), fixture dict of torch tensors)"""
    model_kw, B, pseed, pkw, iseed, seq_len, keys = mk.FORWARD_CASES[name]
    fix = load(name)
    cfg, params, inp = mk.gen_inputs(model_kw, B, pseed, pkw, iseed, seq_len)
    _verify(fix, inp, params, name)
    params = {k: v.to(dtype) for k, v in params.items()}
    out = {k: torch.from_numpy(v) for k, v in fix.items() if not k.startswith("cs_")}
    return dict(model_kw, **mk.REF_FIXED), B, cfg, params, inp, out


def train_case(dtype=torch.float64, name="ref_tiny_train"):
    fix = load(name)
    cfg, params, inp = mk.gen_inputs(mk.TINY, 3, 7, mk.TRAIN_CASES[name], 4321, [20, 9, 2])
    _verify(fix, inp, params, name)
    assert abs(float(inp["target_fine"].double().sum()) - float(fix["cs_target"])) < 1e-6
    params = {k: v.to(dtype) for k, v in params.items()}
    return dict(mk.TINY, **mk.REF_FIXED), 3, cfg, params, inp, fix


def words_case(dtype=torch.float64):
    fix = load("ref_tiny_words")
    cfg, params, inp = mk.gen_inputs(mk.TINY, 3, 7, dict(sharp=40.0, bias_std=0.05, ln_jitter=0.1), 4321, [20, 9, 2])
    _verify(fix, inp, params, "ref_tiny_words")
    params = {k: v.to(dtype) for k, v in params.items()}
    return dict(mk.TINY, **mk.REF_FIXED), 3, cfg, params, inp, fix
