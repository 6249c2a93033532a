"""vsgpu — B200-native batched region-query engine for VariantStore indexes.

Host-side mirror of the reference's query operator API (include/query.h) over the C ABI of
libvsgpu (include/vsgpu.h).  The library has no CPU path: importing works anywhere, but opening an
index without a CUDA device raises.
"""
from .api import Batch, Router, VariantStoreIndex, VsgpuError, Variant, load_library, read_regions, read_sequences  # noqa: F401

__all__ = ["Batch", "Router", "VariantStoreIndex", "VsgpuError", "Variant", "load_library", "read_regions", "read_sequences"]
