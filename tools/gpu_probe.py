"""Quick GPU probe (dev tool): timing of the fused step on a few workloads."""
import sys, time, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pylbm_b200
from pylbm_b200 import cases, runtime as rt

def probe(name, kw, steps=20, warm=5, **simkw):
    t = time.time()
    sim = pylbm_b200.Simulation(cases.CASES[name](**kw), **simkw)
    tb = time.time() - t
    sim.run(warm)
    sim.synchronize()
    lib = rt.lib()
    lib.lbm_sim_timer_start(sim._handle)
    sim.run(steps)
    ms = ctypes.c_float()
    lib.lbm_sim_timer_stop(sim._handle, ctypes.byref(ms))
    cells = np.prod(sim.domain.shape_in)
    q = sim.container.nv
    mlups = cells * steps / (ms.value * 1e-3) / 1e6
    bw = mlups * 1e6 * 2 * q * sim.container.F.itemsize / 1e9
    print("%s %s %s: build %.1fs, %.3f ms/step, %.0f MLUPS, %.0f GB/s algorithmic (%.1f%% of 6552)" % (
        name, kw, simkw, tb, ms.value / steps, mlups, bw, 100 * bw / 6552), flush=True)
    return sim

if __name__ == "__main__":
    which = sys.argv[1:] or ["small"]
    if "small" in which:
        probe("lid_cavity_d3q19", dict(n=128))
        probe("karman_d2q9", dict(nx=1024, ny=256))
    if "big" in which:
        probe("lid_cavity_d3q19", dict(n=256))
        probe("karman_d2q9", dict(nx=4096, ny=1024))
        probe("shallow_water_d2q4", dict(n=4096))
        probe("lid_cavity_d2q9", dict(n=256), steps=200)
        probe("lid_cavity_d3q19", dict(n=512), steps=10, warm=3)
