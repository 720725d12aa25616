"""coregistrationgame_b200 - B200-native Fractional ICP (the hot path of Silviculturalist/CoRegistrationGame).

Only what the path needs: the CUDA kernels + C ABI (``csrc/``, ``libficp_b200.so``), the ctypes
binding (``_lib``), the drop-in ``FractionalICP`` class (``ficp``) and the batched / multi-GPU
search (``batch``, ``dist``).
"""
from .batch import (IcpBatch, TargetIndex, hypothesis_table, register_batch, translation_lattice)  # noqa: F401
from .ficp import FractionalICP  # noqa: F401

__all__ = ["FractionalICP", "TargetIndex", "IcpBatch", "register_batch", "hypothesis_table", "translation_lattice"]
