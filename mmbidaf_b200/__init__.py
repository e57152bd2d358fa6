"""mmbidaf_b200: B200-native (sm_100a) implementation of the MMBiDAF data-parallel hot path.

``layers`` mirrors the reference's ``layers`` package, ``models.MMBiDAF`` the reference's model class.
All heavy arithmetic runs in hand-written CUDA kernels reached through the C ABI in
``include/mmbidaf_b200.h`` (``libmmbidaf_b200.so``, built in-tree by ``python -m mmbidaf_b200.build``).
"""
__version__ = "0.1.0"


def set_precision(tier: str) -> None:
    """Select the arithmetic tier of the contractions (north_star tolerances):

    "fp32"  fp32 FFMA BiDAF kernels, library GEMMs in fp32           -> rel <= 1e-5 against the reference
    "fast"  tcgen05 bf16 BiDAF kernels, library GEMMs in TF32        -> rel <= 2e-2
    Soft-maxes, the LSTM recurrences and all accumulation stay fp32 in both tiers."""
    import torch
    from .layers.attention import BiDAFAttention
    if tier not in ("fp32", "fast"):
        raise ValueError(f"unknown precision tier {tier!r}")
    BiDAFAttention.precision = "bf16" if tier == "fast" else "fp32"
    torch.backends.cuda.matmul.allow_tf32 = tier == "fast"
    torch.backends.cudnn.allow_tf32 = tier == "fast"
