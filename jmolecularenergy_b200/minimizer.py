"""Host-side mirror of the reference's minimizers - the CALLERS of the hot path.

They are restated loop for loop (paths under ``/root/reference/src/main/java/org/jme/minimizer``):
``EnergyMinimizer`` (``EnergyMinimizer.java:38-95``), ``GradientDescentMinimizer``
(``GradientDescentMinimizer.java:49-115``) and ``AdamMinimizer`` (``AdamMinimizer.java:62-161``).
Both estimate a per-coordinate slope by moving ONE Cartesian coordinate and re-evaluating the full
energy, i.e. 3N (6N for Adam) ``calculateEnergy`` calls per step; with the GPU force field each of
those calls is an 8-byte coordinate update plus one energy evaluation on the device.

``GradientMinimizer`` is an addition with no counterpart in the reference: steepest descent on the
analytic gradient the GPU path provides (one evaluation per step instead of 3N).
"""
from __future__ import annotations

import math
from typing import Callable, Optional

import numpy as np


class EnergyMinimizer:
    def __init__(self):
        self.step = 0
        self.onStepConsumer: Optional[Callable[[int, float], None]] = None
        self.constraint = None

    def setOnStepConsumer(self, onStepConsumer) -> None:
        self.onStepConsumer = onStepConsumer

    def setConstraint(self, constraint) -> None:
        self.constraint = constraint

    def getConstraint(self):
        return self.constraint

    def minimize(self, container, maximumSteps: int, stepSize: float, threshold: float) -> None:
        raise NotImplementedError

    def getStep(self) -> int:
        return self.step


def _signum(v: float) -> float:
    return 0.0 if v == 0 else math.copysign(1.0, v)


class GradientDescentMinimizer(EnergyMinimizer):
    """``GradientDescentMinimizer.minimize`` (``GradientDescentMinimizer.java:49-105``).  ``energyCalls``
    counts the ``calculateEnergy`` calls made (the quantity BASELINE config 4 measures)."""

    def __init__(self, forcefield):
        super().__init__()
        self.forcefield = forcefield
        self.energyCalls = 0

    def _energy(self, container) -> float:
        self.energyCalls += 1
        return self.forcefield.calculateEnergy(container)

    def minimize(self, container, maximumSteps: int, stepSize: float, threshold: float) -> None:
        self.step = 0
        n = container.n_atoms
        xyz = container.xyz                      # mutated in place, like the CDK Point3d objects (:107-115)
        grads = np.full((n, 3), float(stepSize))
        initialized = False
        lastEnergy = self._energy(container)
        previousStepEnergy = 0.0
        if self.onStepConsumer is not None:
            self.onStepConsumer(self.step, lastEnergy)
        while self.step < maximumSteps:
            if self.step != 0 and abs(previousStepEnergy - lastEnergy) <= threshold:
                break
            previousStepEnergy = lastEnergy
            for a in range(n):
                for i in range(3):
                    g = grads[a, i]
                    if g == 0:
                        continue
                    distance = min(abs(stepSize * g), .3) * _signum(g)
                    xyz[a, i] += distance
                    if self.constraint is None or self.constraint.check(self.forcefield, container) or not initialized:
                        difference = self._energy(container) - lastEnergy
                        grads[a, i] = -difference / (1 if distance == 0 else distance)
                        if not initialized:
                            xyz[a, i] += -distance
                        elif difference > 0:
                            xyz[a, i] += -distance
                            grads[a, i] = grads[a, i] * 0.9
                        else:
                            lastEnergy += difference
                            grads[a, i] = grads[a, i] * 1.05
                    else:
                        xyz[a, i] += -distance
            if not initialized:
                initialized = True
            else:
                self.step += 1
                if self.onStepConsumer is not None:
                    self.onStepConsumer(self.step, lastEnergy)


def _incremental_mode(incremental) -> int:
    """False / 0: a full evaluation per trial (the reference's loop); True / 1: ``jme_energy_delta`` per trial;
    "device" / 2: the whole loop in one kernel (``minimize_kernel``: systems of up to 8 192 atoms - larger ones are
    run as mode 1)."""
    if incremental in ("device", 2):
        return 2
    return 1 if incremental else 0


class NativeGradientDescentMinimizer(EnergyMinimizer):
    """The same loop as :class:`GradientDescentMinimizer`, run inside ``libjme_b200.so`` (``jme_gd_minimize``): no
    per-call Python / ctypes overhead, which is what BASELINE config 4 (energy calls per second inside the
    optimizer) measures.  Constraints are not supported on this path.  ``incremental=True``: every trial move is
    answered by ``jme_energy_delta``; ``incremental="device"``: the whole loop runs in one kernel - (the energy change from the moved atom's terms only, O(N)) instead of a full
    evaluation - same loop, same decisions up to the rounding of the differences."""

    def __init__(self, forcefield, incremental=False):
        super().__init__()
        self.forcefield = forcefield
        self.incremental = _incremental_mode(incremental)
        self.energyCalls = 0
        self.stepEnergies = []

    def minimize(self, container, maximumSteps: int, stepSize: float, threshold: float) -> None:
        import ctypes as C
        if self.constraint is not None:
            raise NotImplementedError("constraints are evaluated on the host: use GradientDescentMinimizer")
        st = self.forcefield._state(container)
        st.sync_options(self.forcefield, self.forcefield.path)
        xyz = np.ascontiguousarray(container.xyz, dtype=np.float64)
        calls, steps = C.c_int64(), C.c_int()
        log = np.zeros(maximumSteps + 2)
        st.check(st.L.jme_set_incremental(st.h, self.incremental))
        st.check(st.L.jme_gd_minimize(st.h, int(maximumSteps), float(stepSize), float(threshold),
                                      xyz.ctypes.data_as(C.POINTER(C.c_double)), C.byref(calls), C.byref(steps),
                                      log.ctypes.data_as(C.POINTER(C.c_double)), log.shape[0]))
        container.xyz[:] = xyz
        st.shadow = xyz.reshape(-1).copy()
        self.energyCalls = int(calls.value)
        self.step = int(steps.value)
        self.stepEnergies = log[:self.step + 1].tolist()
        if self.onStepConsumer is not None:
            for k, e in enumerate(self.stepEnergies):
                self.onStepConsumer(k, e)


class _AdamOptimizer:
    """``AdamMinimizer.AdamOptimizer`` (``AdamMinimizer.java:131-161``)."""

    def __init__(self, learningRate, beta1, beta2, epsilon):
        self.learningRate, self.beta1, self.beta2, self.epsilon = learningRate, beta1, beta2, epsilon
        self.m = 0.0
        self.v = 0.0
        self.t = 1

    def optimize(self, x: float, gradient: float) -> float:
        self.t += 1
        self.m = self.beta1 * self.m + (1 - self.beta1) * gradient
        self.v = self.beta2 * self.v + (1 - self.beta2) * gradient * gradient
        m_hat = self.m / (1 - math.pow(self.beta1, self.t))
        v_hat = self.v / (1 - math.pow(self.beta2, self.t))
        return x - self.learningRate * m_hat / (math.sqrt(v_hat) + self.epsilon)


class AdamMinimizer(EnergyMinimizer):
    """``AdamMinimizer.minimize`` (``AdamMinimizer.java:62-120``): two energy calls per coordinate."""

    def __init__(self, forceField, beta1: float, beta2: float, epsilon: float, incremental: bool = False):
        super().__init__()
        self.forcefield = forceField
        self.beta1, self.beta2, self.epsilon = beta1, beta2, epsilon
        self.incremental = _incremental_mode(incremental)          # see NativeGradientDescentMinimizer
        self.energyCalls = 0

    def _energy(self, container) -> float:
        self.energyCalls += 1
        return self.forcefield.calculateEnergy(container)

    def minimize(self, container, maximumSteps: int, stepSize: float, threshold: float) -> None:
        self.step = 0
        n = container.n_atoms
        xyz = container.xyz
        grads = np.full((n, 3), float(stepSize))
        adams = [[_AdamOptimizer(stepSize, self.beta1, self.beta2, self.epsilon) for _ in range(3)] for _ in range(n)]
        initialized = False
        lastCalculatedEnergy = self._energy(container)
        energyPerStep = 0.0
        if self.onStepConsumer is not None:
            self.onStepConsumer(self.step, lastCalculatedEnergy)
        while self.step < maximumSteps:
            if self.step != 0 and abs(energyPerStep - lastCalculatedEnergy) <= threshold:
                break
            energyPerStep = lastCalculatedEnergy
            for a in range(n):
                for i in range(3):
                    g = grads[a, i]
                    if g == 0:
                        continue
                    initialEnergy = self._energy(container)
                    distance = adams[a][i].optimize(float(xyz[a, i]), g) - float(xyz[a, i])
                    xyz[a, i] += distance
                    if self.constraint is None or self.constraint.check(self.forcefield, container) or not initialized:
                        difference = self._energy(container) - initialEnergy
                        # Java double division (AdamMinimizer.java:96): a step that rounds to 0 gives +-Infinity / NaN
                        # and the loop goes on
                        with np.errstate(divide="ignore", invalid="ignore"):
                            grads[a, i] = np.float64(-difference) / np.float64(abs(distance))
                        if not initialized:
                            xyz[a, i] += -distance
                        elif difference > 0:
                            xyz[a, i] += -distance
                        else:
                            lastCalculatedEnergy += difference
                    else:
                        xyz[a, i] += -distance
            if not initialized:
                initialized = True
            else:
                self.step += 1
                if self.onStepConsumer is not None:
                    self.onStepConsumer(self.step, lastCalculatedEnergy)


class NativeAdamMinimizer(EnergyMinimizer):
    """:class:`AdamMinimizer`'s loop run inside ``libjme_b200.so`` (``jme_adam_minimize``); no constraints."""

    def __init__(self, forceField, beta1: float, beta2: float, epsilon: float, incremental: bool = False):
        super().__init__()
        self.forcefield = forceField
        self.beta1, self.beta2, self.epsilon = beta1, beta2, epsilon
        self.incremental = _incremental_mode(incremental)          # see NativeGradientDescentMinimizer
        self.energyCalls = 0
        self.stepEnergies = []

    def minimize(self, container, maximumSteps: int, stepSize: float, threshold: float) -> None:
        import ctypes as C
        if self.constraint is not None:
            raise NotImplementedError("constraints are evaluated on the host: use AdamMinimizer")
        st = self.forcefield._state(container)
        st.sync_options(self.forcefield, self.forcefield.path)
        xyz = np.ascontiguousarray(container.xyz, dtype=np.float64)
        calls, steps = C.c_int64(), C.c_int()
        log = np.zeros(maximumSteps + 2)
        st.check(st.L.jme_set_incremental(st.h, self.incremental))
        st.check(st.L.jme_adam_minimize(st.h, int(maximumSteps), float(stepSize), float(threshold), float(self.beta1),
                                        float(self.beta2), float(self.epsilon), xyz.ctypes.data_as(C.POINTER(C.c_double)),
                                        C.byref(calls), C.byref(steps), log.ctypes.data_as(C.POINTER(C.c_double)),
                                        log.shape[0]))
        container.xyz[:] = xyz
        st.shadow = xyz.reshape(-1).copy()
        self.energyCalls = int(calls.value)
        self.step = int(steps.value)
        self.stepEnergies = log[:self.step + 1].tolist()
        if self.onStepConsumer is not None:
            for k, e in enumerate(self.stepEnergies):
                self.onStepConsumer(k, e)


class GradientMinimizer(EnergyMinimizer):
    """Steepest descent with backtracking on the analytic gradient (NOT in the reference)."""

    def __init__(self, forcefield):
        super().__init__()
        self.forcefield = forcefield
        self.energyCalls = 0

    def minimize(self, container, maximumSteps: int, stepSize: float, threshold: float) -> None:
        self.step = 0
        energy, grad = self.forcefield.calculateEnergyAndGradient(container)
        self.energyCalls = 1
        if self.onStepConsumer is not None:
            self.onStepConsumer(0, energy)
        h = float(stepSize)
        while self.step < maximumSteps:
            gmax = float(np.abs(grad).max()) if grad.size else 0.0
            if gmax == 0:
                break
            move = -min(h, .3 / gmax) * grad
            container.xyz += move
            new_energy, new_grad = self.forcefield.calculateEnergyAndGradient(container)
            self.energyCalls += 1
            if new_energy > energy:
                container.xyz -= move
                h *= 0.5
                if h < 1e-12:
                    break
                continue
            converged = abs(energy - new_energy) <= threshold
            energy, grad = new_energy, new_grad
            h *= 1.2
            self.step += 1
            if self.onStepConsumer is not None:
                self.onStepConsumer(self.step, energy)
            if converged:
                break
