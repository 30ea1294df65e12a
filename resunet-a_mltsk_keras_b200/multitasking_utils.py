"""Drop-in for the reference's multitasking_utils.py: the Tanimoto dual loss (:38-85) and the label generators
(:6-34), the latter running on the GPU (labels.py / csrc/labels.cu)."""
from .keras_api import Tanimoto_dual_loss  # noqa: F401
from .labels import get_boundary_label, get_distance_label  # noqa: F401
