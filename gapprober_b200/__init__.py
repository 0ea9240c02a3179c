"""gapprober_b200 — the figbird EM hot path of GapProber on B200 (sm_100a), behind the reference's interface.

Scope: GapFiller::fillGap and what it calls.  csrc/ holds the CUDA kernels and the C ABI
(include/figbird_cuda.h); figbird.py mirrors the reference's GapFiller interface over that ABI;
synth.py drives the seeded workload generator.  Nothing here runs the EM on the CPU.
"""
from .figbird import (FigbirdContext, FigbirdError, FigbirdModel, GapBatch, GapFiller, gap_range, load_library,  # noqa: F401
                      read_gap_file, read_gap_info, read_scaffolds)

__all__ = ["FigbirdContext", "FigbirdError", "FigbirdModel", "GapBatch", "GapFiller", "gap_range", "load_library",
           "read_gap_file", "read_gap_info", "read_scaffolds"]
