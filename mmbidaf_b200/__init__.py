"""mmbidaf_b200: B200-native (sm_100a) implementation of the MMBiDAF data-parallel hot path.

``layers`` mirrors the reference's ``layers`` package, ``models.MMBiDAF`` the reference's model class.
All heavy arithmetic runs in hand-written CUDA kernels reached through the C ABI in
``include/mmbidaf_b200.h`` (``libmmbidaf_b200.so``, built in-tree by ``python -m mmbidaf_b200.build``).
"""
__version__ = "0.1.0"
