"""Reference module path ``TFAM/models/AMO_CLIP.py`` resolved to the B200-native drop-in."""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _root not in sys.path:  # the reference runs TFAM scripts with TFAM/ as cwd
    sys.path.insert(0, _root)

from vimoclip_b200.tfam import AMO_CLIP, AttentionLayer  # noqa: E402,F401
