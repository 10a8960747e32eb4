"""
The north-star interface on several ranks, host side, on the CPU: two processes (gloo, world_size 2)
construct the SAME `pylbm.Simulation(dico, generator='cuda')`; the plugin must find the
`torch.distributed` process group (the role MPI.COMM_WORLD plays in the reference,
mpi_topology.py:74-105), give every rank its x-slab (interface faces labelled -2, no boundary entries
there), broadcast one NCCL id, exchange the CUDA-IPC blobs of the two ring neighbours and drive the
runtime once per step.  The runtime is the recording test double of tests/fake_runtime.py; the slabs'
numbers are checked on real GPUs (tests/test_gpu_multi.py, bench.py's parity block).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, paths, queue):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")] + list(paths)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    from _pytest.monkeypatch import MonkeyPatch

    import fake_runtime

    dist.init_process_group("gloo", rank=rank, world_size=world)
    fake = fake_runtime.install(MonkeyPatch())
    import pylbm
    from pylbm_b200 import cases, plugin

    plugin.register()
    sol = pylbm.Simulation(cases.karman_d2q9(nx=64, ny=16, mod=pylbm, generator="cuda"))
    for _ in range(3):
        sol.one_time_step()
    calls = [n for n, _ in fake.calls]
    ids = [a for n, a in fake.calls if n == "lbm_sim_comm_init"]
    out = {
        "rank": (sol.rank, sol.nranks), "region": [list(map(int, r)) for r in sol.domain.region],
        "labels": [int(v) for v in sol.domain.box_label], "shape_in": [int(v) for v in sol.domain.shape_in],
        "comm_init": len(ids), "comm_args": (int(ids[0][1]), int(ids[0][2])) if ids else None,
        "ipc": (calls.count("lbm_sim_ipc_export"), calls.count("lbm_sim_ipc_open")),
        "steps": calls.count("lbm_sim_step"),
        "entries": int(sum(m.istore.shape[0] for m in sol.bc.methods)),
        "xmin": int(min(m.istore[:, 1].min() for m in sol.bc.methods)),
        "xmax": int(max(m.istore[:, 1].max() for m in sol.bc.methods)),
    }
    queue.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_discover_their_slabs(pylbm):
    import multiprocessing as mp
    import socket

    from conftest import reference_paths

    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, reference_paths(), queue)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted((queue.get(timeout=300) for _ in range(2)), key=lambda d: d["rank"])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    a, b = results
    assert a["rank"] == (0, 2) and b["rank"] == (1, 2)
    assert a["region"][0] == [0, 32] and b["region"][0] == [32, 64] and a["shape_in"] == b["shape_in"] == [32, 16]
    # the inlet face belongs to rank 0, the outlet to rank 1; the cut faces carry label -2 (no boundary entries)
    assert a["labels"][0] == 0 and a["labels"][1] == -2 and b["labels"][0] == -2 and b["labels"][1] == 1
    for r in results:
        assert r["comm_init"] == 1 and r["comm_args"] == (r["rank"][0], 2)
        assert r["ipc"] == (1, 1) and r["steps"] == 3
    # both slabs carry boundary entries from their low to their high ghost plane (the walls' diagonal links
    # reach the corners of the cut faces); the inlet / outlet faces exist on one rank each (labels above)
    assert a["xmin"] == 0 and b["xmax"] == 33 and a["entries"] > 0 and b["entries"] > 0
