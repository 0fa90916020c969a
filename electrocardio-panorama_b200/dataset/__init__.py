"""Device-side counterpart of the reference's ``codes/dataset`` preprocessing (dataset/tianchi.py)."""
from .tianchi import LEAD_THETA, prepare_segments  # noqa: F401
