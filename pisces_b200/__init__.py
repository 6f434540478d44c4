"""pisces_b200 — B200-native (sm_100a) implementation of the Illumina/Pisces per-locus variant-calling hot path.

The product is libpisces_b200.so (hand-written CUDA kernels behind the C ABI of include/pisces_b200.h). This package is the thin
host-side mirror of the reference's interfaces for that path (IStateManager / IAlleleCaller, see interfaces.py) used by the
tests, bench.py and Python hosts. Importing it never falls back to a CPU implementation.
"""
from . import _native  # noqa: F401
from .interfaces import (AlleleCategory, AlleleType, BamReadStager, CalledAllele, DirectionType, FilterType, Genotype, GpuAlleleCaller, GpuStateManager,
                         PiscesB200Error, Read, VariantCallerConfig, make_config)

__all__ = ["AlleleCategory", "AlleleType", "BamReadStager", "CalledAllele", "DirectionType", "FilterType", "Genotype", "GpuAlleleCaller", "GpuStateManager",
           "PiscesB200Error", "Read", "VariantCallerConfig", "make_config"]
