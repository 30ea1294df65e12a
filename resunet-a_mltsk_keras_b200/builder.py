"""Model builder shared by the two drop-in ``Resunet_a`` classes (ResUnet_a/model2.py, model.py)."""
from __future__ import annotations

import os

from . import graph
from .keras_api import Model


def default_dtype():
    """'bf16' (performance mode, fp32 accumulate) or 'fp32' (validation mode); env RSA_DTYPE."""
    return os.environ.get("RSA_DTYPE", "bf16")


def build_model(input_shape, num_classes, multitask, variant, dtype=None, seed=1234, lib=None):
    dtype = dtype or default_dtype()
    net = graph.Net(tuple(input_shape), int(num_classes), bool(multitask), variant=variant, dtype=dtype, seed=seed,
                    lib=lib)
    cfg = dict(input_shape=list(input_shape), num_classes=int(num_classes), multitask=bool(multitask),
               variant=variant, dtype=dtype)
    return Model(net, cfg)


class _ResunetBase(object):
    """Same constructor and attributes as the reference class (model2.py:6-12 / model.py:6-12)."""
    VARIANT = "v2"

    def __init__(self, input_shape, num_classes, args, inputs=None, dtype=None, seed=1234):
        self.num_classes = num_classes
        self.img_height, self.img_width, self.img_channel = input_shape
        self.args = args
        self.inputs = inputs
        self._dtype = dtype
        self._seed = seed
        self.model = self.build_model_ResUneta()

    def build_model_ResUneta(self):
        return build_model((self.img_height, self.img_width, self.img_channel), self.num_classes,
                           bool(getattr(self.args, "multitasking", False)), self.VARIANT, dtype=self._dtype,
                           seed=self._seed)
