"""Minimal `gym` shim: just what the reference uses (see ../README.md).  Not OpenAI gym."""
from . import spaces, utils  # noqa: F401
from .core import Env  # noqa: F401
from .envs.registration import make, register  # noqa: F401

__version__ = "0.0-grasp-shim"
