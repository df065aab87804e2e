"""hs-pose_b200 — B200-native (sm_100a) hot path of HS-Pose.

Python host layer over libhspose_b200.so (hand-written CUDA behind a C ABI,
include/hspose_b200.h).  There is NO CPU fallback: importing is cheap, but any
op raises if the shared library is missing or the device is not a B200.
"""
__version__ = "0.1.0"
