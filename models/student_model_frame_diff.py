"""``from models.student_model_frame_diff import FrameDiffStudentModel`` (train_frame_diff.py:6) -> drop-in."""
from vimoclip_b200.student import FrameDiffStudentModel, ResidualMLP  # noqa: F401
