"""Drop-in for the loss part of the reference's multitasking_utils.py (:38-85)."""
from .keras_api import Tanimoto_dual_loss  # noqa: F401
