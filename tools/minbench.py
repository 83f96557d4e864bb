import sys, time, os
sys.path.insert(0, "/root/repo")
from bench import build_system
from jmolecularenergy_b200 import MMFF94, NativeGradientDescentMinimizer
s = build_system("opt5k")
ff = MMFF94(False); ff.assignParameters(s)
m = NativeGradientDescentMinimizer(ff, incremental="device")
t0 = time.perf_counter(); m.minimize(s, 1, 0.01, 1e-4); dt = time.perf_counter() - t0
print(os.environ.get("JME_B200_LIB", "main").split("/")[-1], m.energyCalls, round(1e6 * dt / m.energyCalls, 3), "us/call", m.stepEnergies[-1], ff.calculateEnergy(s))
