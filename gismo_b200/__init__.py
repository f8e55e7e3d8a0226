"""gismo_b200 — B200-native isogeometric system assembly behind G+Smo's assembler API.

The package holds only the hot path of SURVEY.md section 8: csrc/ (CUDA kernels + C ABI),
host/ (C++ drop-in shims over gismo), and this thin Python face used by tests and bench.py.
"""
from .capi import (Problem, PatchData, CompiledProgram, Gsb200Error, expr_compile, load_library,
                   FORM_POISSON, FORM_ELASTICITY, FORM_MASS)
from .assembler import DeviceAssembler, assemble_host, measure_peaks
from . import host

__all__ = ["Problem", "PatchData", "CompiledProgram", "Gsb200Error", "expr_compile", "load_library",
           "DeviceAssembler", "assemble_host", "measure_peaks", "host",
           "FORM_POISSON", "FORM_ELASTICITY", "FORM_MASS"]
