"""B200-native (sm_100a) GGNN role-graph stage of vFones/situation-recognition.

Python host side (drop-in for the reference's model.py / utils) over libsrggnn.so (include/srggnn.h).
"""
from .imsitu_encoder import imsitu_encoder, tables_from_encoder  # noqa: F401
from .model import FCGGNN, GGSNN, resnet  # noqa: F401
from . import _lib  # noqa: F401

__all__ = ["FCGGNN", "GGSNN", "resnet", "imsitu_encoder", "tables_from_encoder"]
