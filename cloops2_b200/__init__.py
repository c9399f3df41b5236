"""
cloops2_b200 -- B200 (sm_100a) implementation of the cLoops2 `callLoops` hot path: grid DBSCAN over .ixy files and the
per-candidate PET counting, behind the reference's own function / class interface.  Compute lives in libcl2.so
(hand-written CUDA, C ABI in include/cl2.h); this package is the host-side mirror of the reference interface.
"""
__version__ = "0.1.0"
