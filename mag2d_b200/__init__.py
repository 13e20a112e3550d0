"""mag2d_b200 — B200-native particle-in-cell / Monte-Carlo-collision hot path of rouckas/mag2d.

``mag2d_b200.api.Sim`` drives the CUDA library (csrc/, include/mag2d_b200.h); ``config``, ``geometry``
and ``decks`` are host-side readers/builders for the reference's input formats.
"""
__version__ = "0.1.0"
