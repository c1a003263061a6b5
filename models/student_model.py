"""``from models.student_model import FlowStudentModel`` (train.py:6, inference.py:11) -> B200-native drop-in."""
from vimoclip_b200.student import FlowStudentModel, ResidualMLP  # noqa: F401
