"""
mdcraft_b200 — B200-native S(q) and g(r) for MDCraft's ``analysis.structure``.

Only the per-frame structure-analysis hot path of bbye98/mdcraft lives here
(SURVEY.md §8): ``mdcraft_b200.analysis.structure.StructureFactor`` and
``RadialDistributionFunction`` keep the reference's ``AnalysisBase`` surface and
run on hand-written sm_100a CUDA kernels behind a C ABI
(``include/mdcraft_b200.h``).  There is no CPU fallback: using the classes
without the built CUDA library raises.
"""

__version__ = "0.1.0"
