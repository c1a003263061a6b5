"""``from losses import distillation_loss, classification_loss`` (train.py:7) -> B200-native drop-in."""
from vimoclip_b200.losses import classification_loss, distillation_loss, reconstruction_loss  # noqa: F401
