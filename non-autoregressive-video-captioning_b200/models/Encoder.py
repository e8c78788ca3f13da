"""Encoders (contract: reference models/Encoder.py).  Only Encoder_HighWay exists in the reference."""
from .modules import Encoder_HighWay, HighWay  # noqa: F401

__all__ = ("Encoder_HighWay",)
