"""diagnostic: where do in-place and two-array runs of D3Q19 differ?"""
import os, sys
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests")]
os.environ["PYLBM_B200_AA_LIBRARY"] = "1"
import numpy as np
import pylbm_b200
from pylbm_b200 import cases

def nat(sim):
    return np.stack([sim.F[k] for k in range(sim.container.nv)])

for nowalls in ("", "1"):
    if nowalls:
        os.environ["PYLBM_B200_NO_WALLS"] = "1"
    a = pylbm_b200.Simulation(cases.lid_cavity_d3q19(n=16, perturb=0), in_place=True)
    b = pylbm_b200.Simulation(cases.lid_cavity_d3q19(n=16, perturb=0))
    print("walls plan in b:", b.bc.walls is not None)
    for step in range(1, 5):
        a.one_time_step(); b.one_time_step()
        fa, fb = nat(a), nat(b)
        diff = np.abs(fa - fb)
        bad = np.argwhere(diff > 0)
        print("step", step, "swapped", a._swapped, "max diff %.3e" % diff.max(), "n differing", len(bad),
              "pops", sorted(set(bad[:, 0].tolist()))[:20])
        if len(bad):
            cells = bad[:, 1:]
            print("   x range", cells[:, 0].min(), cells[:, 0].max(), "y", cells[:, 1].min(), cells[:, 1].max(),
                  "z", cells[:, 2].min(), cells[:, 2].max(), "sample", bad[:5].tolist())
