"""ndt_slam_b200 -- B200-native (sm_100a) implementation of the NDT hot path of hibikid39/ndt_slam.

Only what the path needs lives here: csrc/ (CUDA kernels + the C ABI of include/ndt_b200.h),
host/ (C++ mirror of the reference's PoseEstimator / ScanMatcher / PointCloudMap /
ScanPointResampler / PoseFuser classes above that ABI), capi.py (ctypes plumbing for tests and
bench.py) and synth.py (seeded synthetic scans). There is no CPU fallback.
"""
__version__ = "0.1.0"
