"""model.py of the reference -> native EncVideo / EncTxt / LAVENDER_Base (same names, methods, state-dict keys)."""
from utils.lib import *  # noqa: F401,F403  (scripts rely on `from model import ...` re-exporting the hub's names)
from lavender_b200.model import EncTxt, EncVideo, LAVENDER_Base  # noqa: F401
