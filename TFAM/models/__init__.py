"""``from models import AMO_CLIP`` inside TFAM/ (TFAM/train_and_eval.py:20) -> B200-native drop-in."""
from .AMO_CLIP import AMO_CLIP  # noqa: F401
