"""Importable as ``CMPC_model`` exactly like the reference module (get_model.py:15-17 does eval(name).LSTM_model)."""
from cmpc_refseg_b200.CMPC_model import LSTM_model, head_param_shapes, reference_init  # noqa: F401
