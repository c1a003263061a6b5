"""vimoclip_b200: B200-native implementation of ViMoCLIP's per-frame video-encoding hot path.

Drop-in classes (same names, constructor arguments, call signatures and state_dict layouts as the
reference):

* ``FlowStudentModel`` / ``FrameDiffStudentModel``  (models/student_model*.py)
* ``AMO_CLIP``                                      (TFAM/models/AMO_CLIP.py)
* ``CLIPVisionFeatures.get_image_features``         (HF ``CLIPModel`` as used by extract_embeddings.py)
* ``CLIPImageProcessor``                            (HF processor call of extract_embeddings.py:91-93, on the GPU)
* ``distillation_loss`` / ``classification_loss``   (losses.py)

All arithmetic of the path runs in hand-written sm_100a CUDA kernels reached through the C-ABI of
``include/vimoclip_b200.h`` (``libvimoclip_b200.so``).  There is no CPU fallback.
"""
from . import _lib, graphs, indexing, ops, store  # noqa: F401
from .graphs import graphed  # noqa: F401
from .clip_hf import CLIPImageProcessor, CLIPVisionFeatures  # noqa: F401
from .losses import classification_loss, distillation_loss  # noqa: F401
from .pipeline import ViMoCLIPPipeline  # noqa: F401
from .student import FlowStudentModel, FrameDiffStudentModel, ResidualMLP  # noqa: F401
from .tfam import AMO_CLIP  # noqa: F401
from .vit import VIT_CONFIGS, VisionTower  # noqa: F401

__version__ = "0.1.0"
