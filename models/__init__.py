"""Reference module path ``models.*`` (train.py:6, inference.py:11) resolved to the B200-native drop-ins."""
