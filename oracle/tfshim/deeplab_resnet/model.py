"""``DeepLabResNetModel({'data': im}, is_training=False).layers[name]`` -> the fed backbone taps (CMPC_model.py:73-76)."""
import tensorflow as tf


class DeepLabResNetModel(object):
    def __init__(self, inputs, is_training=False, **_):
        self.inputs = inputs
        self.layers = dict(tf._shim.layers)
