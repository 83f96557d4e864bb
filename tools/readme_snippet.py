import sys; sys.path.insert(0, ".")
from jmolecularenergy_b200 import MMFF94, NativeGradientDescentMinimizer, systems
s = systems.water_box(1667)
ff = MMFF94(False); ff.setCutoffDistance(0.0); ff.assignParameters(s)
e = ff.calculateEnergy(s)
m = NativeGradientDescentMinimizer(ff, incremental="device"); m.minimize(s, 2, 0.01, 1e-4)
print(e, m.stepEnergies, ff.calculateEnergy(s))
