"""Drop-in for ResUnet_a/model2.py (the graph train_ISPRS.py:4 trains): ResBlock-a with identity
add, Conv2DN (1x1 conv + BN) in PSPPooling / UpSampling / combine, ReLU after PSPPooling.
``Resunet_a(input_shape, num_classes, args).model`` has the Keras model surface; the graph itself
is emitted by resuneta_b200.graph.define_network as B200 kernels."""
from ..builder import _ResunetBase


class Resunet_a(_ResunetBase):
    VARIANT = "v2"
