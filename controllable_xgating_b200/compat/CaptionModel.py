"""`from CaptionModel import CaptionModel` shim (see compat/SAModel.py)."""
from controllable_xgating_b200.CaptionModel import CaptionModel  # noqa: F401
