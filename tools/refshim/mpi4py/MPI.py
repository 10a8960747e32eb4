"""Single-rank mpi4py.MPI stand-in (dev tooling; see tools/refshim/README).

Surface = what the reference touches on one rank: COMM_WORLD, Cartesian
communicator, DOUBLE.Create_subarray, Irecv/Isend + Request.Waitall implemented
as tag-matched slice copies of the same array (periodic self-exchange)."""
import time
import numpy as np

SUM = "sum"


def Wtime():
    return time.perf_counter()


def Compute_dims(size, dim):
    return [1] * (dim if isinstance(dim, int) else len(dim))


class _Subarray:
    def __init__(self, sizes, subsizes, starts):
        self.slices = tuple(slice(s, s + n) for s, n in zip(starts, subsizes))

    def Commit(self):
        return self

    def Free(self):
        pass


class _Double:
    def Create_subarray(self, sizes, subsizes, starts):
        return _Subarray(sizes, subsizes, starts)


DOUBLE = _Double()


class _Req:
    def __init__(self, kind, array, sub, tag):
        self.kind, self.array, self.sub, self.tag = kind, array, sub, tag


class Request:
    @staticmethod
    def Waitall(reqs):
        sends = {r.tag: r.array[r.sub.slices].copy() for r in reqs if r.kind == "s"}
        for r in reqs:
            if r.kind == "r":
                r.array[r.sub.slices] = sends[r.tag]


class _Comm:
    rank = 0
    size = 1

    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def Barrier(self):
        pass

    def allreduce(self, sendobj=None, op=SUM):
        return sendobj

    def gather(self, obj, root=0):
        return [obj]

    def bcast(self, obj, root=0):
        return obj

    def Create_cart(self, dims, periods=None, reorder=False):
        return _Cart(list(dims), list(periods) if periods is not None else None)

    def Irecv(self, buf, source=0, tag=0):
        return _Req("r", buf[0], buf[1], tag)

    def Isend(self, buf, dest=0, tag=0):
        return _Req("s", buf[0], buf[1], tag)

    def Free(self):
        pass


class _Cart(_Comm):
    def __init__(self, dims, periods):
        self.dims = dims
        self.periods = periods

    def Get_topo(self):
        return self.dims, self.periods, [0] * len(self.dims)

    def Get_coords(self, rank):
        return [0] * len(self.dims)

    def Get_cart_rank(self, coords):
        return 0

    def Sub(self, remain):
        return self


COMM_WORLD = _Comm()
