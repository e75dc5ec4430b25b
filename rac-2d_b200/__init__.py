"""rac-2d_b200: B200-native (sm_100a) drop-in for RAC-2D's per-cell stiff-chemistry solve.

The product is the C-ABI shared library ``libracg.so`` (``include/racg.h``) built from
``csrc/``; this package is the thin host-side mirror of the reference's own interface for
the path (``chem_read_reactions`` ... ``chem_evol_solve``), used by the tests, the
benchmark and ``__graft_entry__``.  There is no CPU fallback: every compute call goes
through the CUDA library and raises if it is missing or no GPU is visible.
"""
from .chem import (NPAR, NSTAT, ChemNetwork, ChemSolver, RacgError, SolveParams, build, default_cfg,
                   lib, lib_path, read_chemical_data, write_chemical_data)
from . import synth

__all__ = ["NPAR", "NSTAT", "ChemNetwork", "ChemSolver", "RacgError", "SolveParams", "build",
           "default_cfg", "lib", "lib_path", "read_chemical_data", "write_chemical_data", "synth"]
