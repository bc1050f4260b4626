"""agent.py of the reference -> native train loop (same class names and methods)."""
from utils.lib import *  # noqa: F401,F403
from lavender_b200.agent import Agent_Base, NormSoftmaxLoss, WarmupLinearLR, humanbytes, move_to_cuda  # noqa: F401
