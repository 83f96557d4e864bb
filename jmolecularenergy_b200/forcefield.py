"""Host-side mirror of the reference's force-field API, backed by ``libjme_b200.so``.

Names, argument meaning and error behaviour follow the Java classes (paths under
``/root/reference/src/main/java/org/jme/forcefield``) so that parity tests read like the reference's
own: ``ForceField`` (``ForceField.java:49-212``), ``EnergyComponent`` (``EnergyComponent.java:40-57``),
``MMFF94`` (``mmff/MMFF94.java:57-397``).  The "atom container" is a :class:`MolecularSystem`: the
flat form of what ``assignParameters`` leaves on a CDK container.

What differs, by design: ``assignParameters`` does not type atoms (the CDK typer is out of scope,
SURVEY.md section 8f) - it marshals an already parameterised system to the GPU; and there is a
``calculateGradient`` (the reference has none).  No CPU path exists: without the CUDA library every
call raises.
"""
from __future__ import annotations

import ctypes as C
import enum
from typing import Dict, List, Optional

import numpy as np

from . import _lib
from ._abi import COMPONENTS, ENERGY_UNITS, PATHS, marshal, ptr
from .system import MolecularSystem

NOT_ASSIGNED = "MMFF94 parameters need to be assigned first"      # MMFF94.java:373


class EnergyUnit(enum.Enum):
    """``ForceField.EnergyUnit`` (``ForceField.java:133-160``)."""
    KCAL_PER_MOL = 4184.0
    KJ_PER_MOL = 1000.0
    JOULE_PER_MOL = 1.0

    def convertTo(self, value: float, toUnit: "EnergyUnit") -> float:
        if self is toUnit:
            return value
        return (value * self.value) / toUnit.value


class JmeError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{message} (jme status {code})")
        self.code = code


class EnergyComponent:
    """``EnergyComponent`` (``EnergyComponent.java:40-57``)."""

    def __init__(self):
        self.forceField: Optional["ForceField"] = None

    def setForceField(self, forceField) -> None:
        self.forceField = forceField

    def calculateEnergy(self, atomContainer: MolecularSystem) -> float:
        raise NotImplementedError


class ForceField:
    """``ForceField`` (``ForceField.java:49-212``): component registry and the three options."""

    def __init__(self):
        self._dielectricConstant = 1.0
        self._cutoffDistance = 0.0
        self.energyUnit = EnergyUnit.KCAL_PER_MOL
        self.components: List[EnergyComponent] = []

    def assignParameters(self, atomContainer) -> None:
        raise NotImplementedError

    def calculateEnergy(self, atomContainer) -> float:
        raise NotImplementedError

    def getEnergyComponents(self):
        return tuple(self.components)

    def addEnergyComponent(self, component: EnergyComponent) -> None:
        if component is None:
            raise TypeError("component must not be None")           # Objects.requireNonNull
        component.setForceField(self)
        self.components.append(component)

    def removeEnergyComponent(self, component: EnergyComponent) -> None:
        if component is None:
            raise TypeError("component must not be None")
        component.setForceField(None)
        if component in self.components:
            self.components.remove(component)

    def setEnergyUnit(self, energyUnit: EnergyUnit) -> None:
        if energyUnit is None:
            raise TypeError("energyUnit must not be None")
        self.energyUnit = energyUnit

    def getEnergyUnit(self) -> EnergyUnit:
        return self.energyUnit

    def setCutoffDistance(self, cutoffDistance: float) -> None:
        self._cutoffDistance = float(cutoffDistance)

    def getCutoffDistance(self) -> float:
        return self._cutoffDistance

    def setDielectricConstant(self, dielectricConstant: float) -> None:
        self._dielectricConstant = float(dielectricConstant)

    def getDielectricConstant(self) -> float:
        return self._dielectricConstant


class _GpuState:
    """Device-side twin of one parameterised container: the jme handle plus a host shadow of the
    coordinates last uploaded, so that a caller that moves ONE coordinate in place (both reference
    minimizers do, ``GradientDescentMinimizer.java:107-115``) costs an 8-byte update, not 3N doubles."""

    def __init__(self, system: MolecularSystem):
        L = _lib.lib()
        js, keep = marshal(system)
        h = C.c_void_p()
        rc = L.jme_create(C.byref(js), C.byref(h))
        if rc != 0:
            raise JmeError(rc, L.jme_last_create_error().decode())
        self.L = L
        self.h = h
        self.n = system.n_atoms
        self.shadow: Optional[np.ndarray] = None
        self.options = None
        self.path = None
        self.cached = None             # (options, components + total) of the last energy, for the incremental mode
        self.since_full = 0            # incremental answers since the last full evaluation

    def check(self, rc: int) -> None:
        if rc != 0:
            raise JmeError(rc, self.L.jme_last_error(self.h).decode())

    def sync_options(self, ff: ForceField, path: str) -> None:
        opt = (ff.getCutoffDistance(), ff.getDielectricConstant(), ff.getEnergyUnit().name)
        if opt != self.options:
            self.check(self.L.jme_set_options(self.h, opt[0], opt[1], ENERGY_UNITS[opt[2]]))
            self.options = opt
            self.cached = None
        if path != self.path:
            self.check(self.L.jme_set_path(self.h, PATHS[path]))
            self.path = path

    def sync_coords(self, xyz: np.ndarray) -> None:
        flat = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1)
        if flat.shape[0] != 3 * self.n:
            raise ValueError("coordinate count differs from the parameterised system")
        if self.shadow is not None:
            changed = np.flatnonzero(flat != self.shadow)
            if changed.size == 0:
                return
            self.cached = None         # (the incremental mode moves its one coordinate itself, before this is reached)
            if changed.size <= 3:
                for k in changed.tolist():
                    self.check(self.L.jme_set_coord1(self.h, k // 3, k % 3, float(flat[k])))
                    self.shadow[k] = flat[k]
                return
        self.check(self.L.jme_set_coords(self.h, flat.ctypes.data, self.n))
        self.shadow = flat.copy()
        self.cached = None

    def close(self) -> None:
        if self.h is not None and self.h.value:
            self.L.jme_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _MMFF94TermComponent(EnergyComponent):
    """One of the seven MMFF94 terms (``MMFF94.java:75-82``).  All seven are produced by a single GPU
    evaluation; ``calculateEnergy`` of one component returns its share of that evaluation."""

    def __init__(self, key: str):
        super().__init__()
        self.key = key

    def calculateEnergy(self, atomContainer: MolecularSystem) -> float:
        if self.forceField is None:
            raise TypeError("component is not attached to a force field")     # Objects.requireNonNull(forceField)
        return self.forceField._evaluate(atomContainer, gradient=False)[self.key]


class MMFF94(ForceField):
    """``MMFF94`` (``mmff/MMFF94.java:57-397``) on the GPU.

    ``MMFF94(mmff94s)``: MMFF94 vs MMFF94s differ only in out-of-plane / torsion parameter values
    (``MMFF94.java:265,660``), which arrive with the system; the flag is kept for API parity.
    """

    def __init__(self, mmff94s: bool = False, path: str = "auto"):
        super().__init__()
        self.mmff94s = bool(mmff94s)
        self.path = path
        self.bondStretchingComponent = _MMFF94TermComponent("bond")
        self.angleBendingComponent = _MMFF94TermComponent("angle")
        self.stretchBendComponent = _MMFF94TermComponent("stretch_bend")
        self.outOfPlaneComponent = _MMFF94TermComponent("oop")
        self.torsionComponent = _MMFF94TermComponent("torsion")
        self.vdwComponent = _MMFF94TermComponent("vdw")
        self.electrostaticComponent = _MMFF94TermComponent("ele")
        for c in (self.bondStretchingComponent, self.angleBendingComponent, self.stretchBendComponent,
                  self.outOfPlaneComponent, self.torsionComponent, self.vdwComponent,
                  self.electrostaticComponent):                         # order of MMFF94.java:98
            self.addEnergyComponent(c)
        self.lastEvaluation: Dict[str, float] = {}
        self.incremental = False

    def setIncremental(self, on: bool) -> None:
        """Opt in to incremental energies: when exactly ONE coordinate changed since the last ``calculateEnergy`` of a
        container - which is what the reference's minimizers do between two calls (``GradientDescentMinimizer.java:
        75-93``) - the answer is the previous energy plus ``jme_energy_delta`` (the moved atom's terms only, O(N))
        instead of a full evaluation (two coordinates - the undo of a rejected trial plus the next trial - cost two deltas);
        every 3 N such answers, and whenever anything else changed, a full evaluation resynchronises.  The unchanged minimizer loops get the speed-up through the unchanged ``calculateEnergy``."""
        self.incremental = bool(on)

    # -- marshalling -------------------------------------------------------------------------
    def assignParameters(self, atomContainer: MolecularSystem) -> None:
        """Marshal the parameterised container to the current CUDA device (``jme_create``).  Must be
        repeated whenever the molecular structure changes, like the reference's (``MMFF94.java:104-109``)."""
        old = atomContainer.__dict__.pop("_jme_gpu", None)
        if old is not None:
            old.close()
        atomContainer.validate()
        atomContainer.__dict__["_jme_gpu"] = _GpuState(atomContainer)

    @staticmethod
    def _state(atomContainer) -> _GpuState:
        st = getattr(atomContainer, "__dict__", {}).get("_jme_gpu")
        if st is None or st.h is None:
            raise RuntimeError(NOT_ASSIGNED)                           # MMFF94.java:371-375
        return st

    def release(self, atomContainer: MolecularSystem) -> None:
        st = atomContainer.__dict__.pop("_jme_gpu", None)
        if st is not None:
            st.close()

    # -- evaluation --------------------------------------------------------------------------
    def _evaluate(self, atomContainer: MolecularSystem, gradient: bool) -> Dict[str, object]:
        st = self._state(atomContainer)
        st.sync_options(self, self.path)
        if self.incremental and not gradient and st.cached is not None and st.shadow is not None \
                and st.since_full < 3 * st.n:
            flat = np.ascontiguousarray(atomContainer.xyz, dtype=np.float64).reshape(-1)
            changed = np.flatnonzero(flat != st.shadow) if flat.shape[0] == st.shadow.shape[0] else np.arange(2)
            if changed.size == 0:
                return dict(st.cached)
            if changed.size <= 2:
                # one coordinate moved - or two: the undo of a rejected trial and the next trial, which is what the
                # reference's loops leave between two calls (GradientDescentMinimizer.java:83-93) - one delta each
                out = dict(st.cached)
                for k in changed.tolist():
                    d = (C.c_double * 8)()
                    st.check(st.L.jme_energy_delta(st.h, k // 3, k % 3, float(flat[k]), d))
                    st.shadow[k] = flat[k]
                    st.since_full += 1
                    out["total"] += d[0]
                    for j, name in enumerate(COMPONENTS):
                        out[name] += d[1 + j]
                st.cached = dict(out)
                self.lastEvaluation = out
                return out
        st.sync_coords(atomContainer.xyz)
        total = C.c_double()
        comps = np.zeros(7)
        out: Dict[str, object] = {}
        if gradient:
            grad = np.empty((st.n, 3), dtype=np.float64)
            st.check(st.L.jme_energy_gradient(st.h, C.byref(total), ptr(comps), grad.ctypes.data))
            out["gradient"] = grad
        else:
            st.check(st.L.jme_energy(st.h, C.byref(total), ptr(comps)))
        out.update(zip(COMPONENTS, comps.tolist()))
        out["total"] = total.value
        pe, pc = C.c_uint64(), C.c_uint64()
        st.L.jme_pair_stats(st.h, C.byref(pe), C.byref(pc))
        out["pairs"] = pe.value
        out["candidates"] = pc.value
        self.lastEvaluation = out
        st.cached = {k: v for k, v in out.items() if k != "gradient"}
        st.since_full = 0
        return out

    def calculateEnergy(self, atomContainer: MolecularSystem) -> float:
        """``MMFF94.calculateEnergy`` (``MMFF94.java:385-397``): sum over the registered components in
        registration order.  One GPU evaluation yields all seven component sums."""
        if atomContainer.n_atoms == 0:
            return 0.0
        ev = self._evaluate(atomContainer, gradient=False)
        if len(self.components) == 7 and all(isinstance(c, _MMFF94TermComponent) for c in self.components):
            return ev["total"]
        energy = 0.0
        for c in self.components:
            energy += ev[c.key] if isinstance(c, _MMFF94TermComponent) and c.forceField is self \
                else c.calculateEnergy(atomContainer)
        return energy

    def calculateGradient(self, atomContainer: MolecularSystem) -> np.ndarray:
        """dE/dx, shape (N, 3), in energy-unit per Angstrom.  No counterpart in the reference."""
        return self._evaluate(atomContainer, gradient=True)["gradient"]

    def calculateEnergyAndGradient(self, atomContainer: MolecularSystem):
        ev = self._evaluate(atomContainer, gradient=True)
        return ev["total"], ev["gradient"]

    def calculateComponents(self, atomContainer: MolecularSystem, gradient: bool = False) -> Dict[str, object]:
        return dict(self._evaluate(atomContainer, gradient=gradient))

    def energyChangeOfMove(self, atomContainer: MolecularSystem, atom: int, axis: int, value: float) -> Dict[str, float]:
        """Move one coordinate of one atom to ``value`` and return what that changes: ``total`` and the seven components
        (``jme_energy_delta``: only the moved atom's pairs and bonded terms are evaluated, O(N)).  The container's
        coordinate is updated too.  No counterpart in the reference, whose minimizers take the difference of two full
        evaluations."""
        st = self._state(atomContainer)
        st.sync_options(self, self.path)
        st.sync_coords(atomContainer.xyz)
        d = (C.c_double * 8)()
        st.check(st.L.jme_energy_delta(st.h, int(atom), int(axis), float(value), d))
        atomContainer.xyz[atom, axis] = value
        if st.shadow is not None:
            st.shadow[3 * atom + axis] = value
        st.cached = None
        out = {"total": d[0]}
        for k, name in enumerate(COMPONENTS):
            out[name] = d[1 + k]
        return out

    def pairList(self, atomContainer: MolecularSystem):
        """Selected nonbonded pair set (i<j, is14) of the current coordinates and options - diagnostic
        twin of the reference's ``mmff.nonBondedInteraction`` map after the cutoff test."""
        st = self._state(atomContainer)
        st.sync_options(self, self.path)
        st.sync_coords(atomContainer.xyz)
        n = C.c_int64()
        st.check(st.L.jme_pair_list(st.h, None, None, None, 0, C.byref(n)))
        cap = max(1, n.value)
        pi, pj, f = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.uint8)
        st.check(st.L.jme_pair_list(st.h, ptr(pi), ptr(pj), ptr(f), cap, C.byref(n)))
        m = n.value
        order = np.lexsort((pj[:m], pi[:m]))
        return pi[:m][order], pj[:m][order], f[:m][order]

    def launchCount(self, atomContainer: MolecularSystem) -> int:
        st = self._state(atomContainer)
        return int(st.L.jme_launch_count(st.h))


def measure_fp64_peak():
    """(TFLOP/s, estimated SM clock MHz) of the current device's FP64 FMA pipe."""
    L = _lib.lib()
    t, c = C.c_double(), C.c_double()
    rc = L.jme_measure_fp64_peak(C.byref(t), C.byref(c))
    if rc != 0:
        raise JmeError(rc, "jme_measure_fp64_peak failed (no CUDA device?)")
    return t.value, c.value
