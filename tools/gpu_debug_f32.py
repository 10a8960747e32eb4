import sys, os
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests")]
import numpy as np
import pylbm_b200
from pylbm_b200 import cases
from oracle.lbm_oracle import OracleSimulation
from conftest import PARITY_CASES

for idx in (4, 5, 3, 0, 1, 6):
    name, kw = PARITY_CASES[idx]
    for dtype, comp in (("float32", None), ("float32", "float32")):
        sim = pylbm_b200.Simulation(cases.CASES[name](perturb=0, **kw), dtype=dtype, compute_dtype=comp)
        ora = OracleSimulation(cases.CASES[name](perturb=0, **kw))
        fluid = ora.domain.in_or_out[tuple(slice(v, -v) for v in ora.domain.stencil.vmax)] == ora.domain.valin
        line = []
        for n in (0, 1, 2, 5, 20, 50):
            while sim.nt < n:
                sim.one_time_step(); ora.one_time_step()
            worst = 0
            for key in sim.scheme.consm:
                okey = [k for k in ora.scheme.consm if str(k) == str(key)][0]
                a, b = sim.m[key], ora.m[okey]
                worst = max(worst, np.abs(a[fluid] - b[fluid]).max() / np.abs(b[fluid]).max())
            line.append("%d:%.2e" % (n, worst))
        print(name, dtype, comp, " ".join(line), flush=True)
