"""Only the load-bearing constant of reference ``capstone/utils/miccai.py:14-24``."""
from ...losses import STRUCTURES  # noqa: F401
