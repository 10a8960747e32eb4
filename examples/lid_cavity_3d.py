#!/usr/bin/env python
"""
D3Q19 MRT lid-driven cavity on one B200 through the pylbm API (BASELINE config 4 at a chosen size),
with HDF5 + XDMF output every `period` steps -- the shape of the reference's demo/3D/lid_cavity.py
(`save()` is that demo's helper, unchanged), with `generator='cuda'`.

    python examples/lid_cavity_3d.py [n] [nsteps] [period]        # defaults: 128 2000 500
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import pylbm_b200 as pylbm                      # noqa: E402
from pylbm_b200 import cases                    # noqa: E402

RHO, QX, QY, QZ = cases.RHO, cases.QX, cases.QY, cases.QZ


def save(sol, im, path):
    x, y, z = sol.domain.x, sol.domain.y, sol.domain.z
    h5 = pylbm.H5File(sol.domain.mpi_topo, "lid_cavity", path, im)
    h5.set_grid(x, y, z)
    h5.add_scalar("mass", sol.m[RHO])
    h5.add_vector("velocity", [sol.m[QX], sol.m[QY], sol.m[QZ]])
    h5.save()


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    period = int(sys.argv[3]) if len(sys.argv) > 3 else 500
    sol = pylbm.Simulation(cases.lid_cavity_d3q19(n=n))        # a plain pylbm dictionary, generator='cuda'
    print(sol)
    done, im = 0, 0
    t0 = time.perf_counter()
    while done < nsteps:
        k = min(period, nsteps - done)
        sol.run(k)                       # k steps in one runtime call (or: for ...: sol.one_time_step())
        done += k
        save(sol, im, "./lid_cavity_out")
        im += 1
    sol.synchronize()
    wall = time.perf_counter() - t0
    print("%d steps of %d^3 in %.2f s incl. output: %.0f MLUPS" % (nsteps, n, wall, nsteps * n**3 / wall / 1e6))


if __name__ == "__main__":
    main()
