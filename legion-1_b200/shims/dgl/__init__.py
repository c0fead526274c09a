"""Minimal pure-torch stand-in for the part of DGL the Legion trainers use (see ../README.md)."""
from . import heterograph, heterograph_index, nn  # noqa: F401

__version__ = "0.0-legion-b200-shim"
