#!/usr/bin/env python
"""
The reference's Karman vortex street (demo/2D/Karman_vortex_street.py: D2Q9 with Geier's moments and a
relative velocity, Bouzidi bounce-back inlet / walls / cylinder, Neumann outlet) driven through the
UNMODIFIED `pylbm.Simulation` with `generator='cuda'`.  Needs the reference installed in oracle/_ref
(`tools/make_ref.sh`, done by `__graft_entry__.build()` in the build container) or any importable pylbm.

    python examples/karman_pylbm_cuda.py [nx] [ny] [nsteps]          # defaults: 2048 512 4000
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle", "_ref"), os.path.join(ROOT, "oracle", "_ref", "shims")]

import pylbm                                    # noqa: E402  the reference package
from pylbm_b200 import cases, plugin            # noqa: E402

plugin.register()                               # admits generator='cuda'


def main():
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    ny = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    nsteps = int(sys.argv[3]) if len(sys.argv) > 3 else 4000
    # a plain pylbm dictionary built with the reference's own classes (pylbm.Circle, pylbm.bc.*)
    dico = cases.karman_d2q9(nx=nx, ny=ny, mod=pylbm, generator="cuda")
    sol = pylbm.Simulation(dico)
    print(type(sol).__mro__[:3])
    rho, qx, qy = (k for k in sol.scheme.consm)
    t0 = time.perf_counter()
    for _ in range(nsteps):
        sol.one_time_step()                      # one enqueue-only runtime call per step
    vorticity_proxy = sol.m[qy]                  # device -> host, interior cells
    dt = time.perf_counter() - t0
    print("%d steps of %d x %d in %.2f s: %.0f MLUPS; max |qy| = %.3e"
          % (nsteps, nx, ny, dt, nsteps * nx * ny / dt / 1e6, abs(vorticity_proxy).max()))


if __name__ == "__main__":
    main()
