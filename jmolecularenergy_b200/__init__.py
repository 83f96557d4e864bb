"""jmolecularenergy_b200 - B200 (sm_100a) drop-in for JMolecularEnergy's MMFF94 energy / gradient hot path.

The compute lives in ``libjme_b200.so`` (hand-written CUDA behind the C ABI of ``include/jme_b200.h``);
this package is the host-side mirror of the reference's Java API plus marshalling helpers.  There is no
CPU fallback.
"""
from .system import MolecularSystem, enumerate_terms, pair_separation  # noqa: F401
from .forcefield import (EnergyComponent, EnergyUnit, ForceField, JmeError, MMFF94,  # noqa: F401
                         measure_fp64_peak)
from .minimizer import (AdamMinimizer, EnergyMinimizer, GradientDescentMinimizer,  # noqa: F401
                        GradientMinimizer, NativeAdamMinimizer, NativeGradientDescentMinimizer)

__version__ = "0.1.0"
