"""Drop-in for ResUnet_a/model.py (older graph used by amazon_py/main_tcc.py:9): no identity add,
no BatchNormalization after 1x1 convolutions, PSPPooling convolves before up-sampling, decoder is
Conv1x1(f) -> UpSampling2D.  Honours ``inputs=`` (model.py:72-74, kept as an attribute) and
``args.gpu_parallel`` (model.py:164-165: returns ``(inputs, [seg, bound, dist, color])``)."""
from ..builder import _ResunetBase


class _OutputHandle:
    """Stands for one output tensor of the un-compiled graph in the gpu_parallel return value."""

    def __init__(self, model, name):
        self.model, self.name = model, name

    def __repr__(self):
        return f"<ResUnet-a output '{self.name}'>"


class Resunet_a(_ResunetBase):
    VARIANT = "v1"

    def build_model_ResUneta(self):
        model = super().build_model_ResUneta()
        if getattr(self.args, "multitasking", False) and getattr(self.args, "gpu_parallel", False):
            return self.inputs, [_OutputHandle(model, n) for n in model.output_names]
        return model
