"""visbackbone/video_swin.py of the reference -> the native Video Swin (same class / function names)."""
from lavender_b200.video_swin import (BasicLayer, Mlp, PatchEmbed3D, PatchMerging, SwinTransformer3D,  # noqa: F401
                                      SwinTransformerBlock3D, WindowAttention3D, get_vidswin_model, get_window_size)
