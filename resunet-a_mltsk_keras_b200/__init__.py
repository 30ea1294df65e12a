"""resuneta_b200 — B200-native ResUnet-a multitask hot path (drop-in for the Keras reference).

Importable as ``resuneta_b200`` (the on-disk directory is ``resunet-a_mltsk_keras_b200``).
"""
from .ResUnet_a.model2 import Resunet_a  # noqa: F401  (primary variant, train_ISPRS.py:4)
from .keras_api import (Adam, SGD, Tanimoto_dual_loss, weighted_categorical_crossentropy,  # noqa: F401
                        CategoricalCrossentropy, BinaryCrossentropy, MeanSquaredError, EarlyStopping,
                        ModelCheckpoint, load_model, Model)

__version__ = "0.1.0"
