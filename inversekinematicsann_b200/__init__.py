"""B200-native batched inverse kinematics -- drop-in for the hot path of lstar93/InverseKinematicsANN.

The product path is ``csrc/libikb200.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/ikb200.h``) plus the thin Python mirror of the reference's ``kinematics/`` and ``robot/``
packages found in ``inversekinematicsann_b200.kinematics`` / ``.robot``.  There is no CPU fallback:
solving without a CUDA device raises.
"""
from .engine import IkEngine, IkStats, NativeLibraryError  # noqa: F401

__version__ = "0.1.0"
