"""Drop-in replacement of the reference's star-import hub (utils/lib.py:1-27): same names, but
  * `transformers` is a facade whose AutoModel* factories return lavender_b200's native BERT modules, and
  * optional third-party packages the hot path never needs (easydict, skimage, fairscale, toolz, cv2) degrade to small
    stand-ins when they are not installed."""
import argparse, sys, os, io, base64, pickle, json, math, random  # noqa: F401,E401
import os.path as op  # noqa: F401
import time, errno  # noqa: F401,E401
import inspect  # noqa: F401
from collections import defaultdict  # noqa: F401
from datetime import datetime, timedelta  # noqa: F401

import numpy as np  # noqa: F401
import torch as T  # noqa: F401
import torch.distributed as DIST  # noqa: F401
from packaging import version  # noqa: F401
from torch.utils.data import ConcatDataset  # noqa: F401
from tqdm import tqdm  # noqa: F401

try:
    import torchvision as TV  # noqa: F401
except ImportError:  # pragma: no cover
    TV = None
try:
    import cv2  # noqa: F401
except ImportError:  # pragma: no cover
    cv2 = None
try:
    from PIL import Image  # noqa: F401
except ImportError:  # pragma: no cover
    Image = None
try:
    from easydict import EasyDict as edict  # noqa: F401
except ImportError:
    from lavender_b200.config import Args as edict  # noqa: F401
try:
    from skimage.feature import hog as hog_feature  # noqa: F401
except ImportError:
    hog_feature = None
try:
    from toolz.sandbox import unzip  # noqa: F401
except ImportError:
    def unzip(seq):
        return zip(*seq)

if not hasattr(inspect, "getargspec"):       # removed in Python 3.11; agent.py:208 of the reference uses it
    inspect.getargspec = inspect.getfullargspec

os.environ["TOKENIZERS_PARALLELISM"] = "true"


def checkpoint_wrapper(module, offload_to_cpu=False, **unused):
    """fairscale.nn.misc.checkpoint_wrapper (model.py:167-169): activation checkpointing with CPU offload exists to
    fit 16-32 GB GPUs; with 180 GB of HBM3e the native path keeps activations resident, so this is the identity."""
    return module


from lavender_b200.hf_facade import transformers  # noqa: E402,F401
