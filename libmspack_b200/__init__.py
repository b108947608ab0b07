"""libmspack_b200 - B200-native batch decompressor for the CAB-folder codecs (MSZIP, Quantum, LZX)."""
from .units import *  # noqa: F401,F403
