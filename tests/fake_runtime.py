"""
Test double of the runtime library for CPU boxes: lets the HOST-SIDE plumbing of the CUDA backend
(constructor flow through the reference's `pylbm.Simulation`, boundary-list registration, argument
marshalling) run where there is no GPU.  It computes nothing: device memory is host memory, copies
and kernels are no-ops, every call is recorded.  Test infrastructure only -- the product never
imports it, and a real box without a GPU still fails loudly (runtime.ensure_gpu).
"""
import ctypes


class FakeRuntime:
    def __init__(self):
        self.calls = []
        self._buffers = {}
        self._next_bc = 0

    def __getattr__(self, name):
        if not name.startswith("lbm_"):
            raise AttributeError(name)

        def call(*args):
            self.calls.append((name, args))
            return 0

        return call

    def count(self, name):
        return sum(1 for n, _ in self.calls if n == name)

    def lbm_abi_version(self):
        return 1

    def lbm_last_error(self):
        return b"fake runtime"

    def lbm_device_count(self):
        return 1

    def lbm_malloc(self, out, nbytes):
        buf = ctypes.create_string_buffer(max(int(nbytes), 8))
        addr = ctypes.addressof(buf)
        self._buffers[addr] = buf
        ctypes.cast(out, ctypes.POINTER(ctypes.c_void_p))[0] = addr
        self.calls.append(("lbm_malloc", (nbytes,)))
        return 0

    lbm_host_alloc = lbm_malloc

    def lbm_free(self, ptr):
        self._buffers.pop(getattr(ptr, "value", ptr), None)
        return 0

    lbm_host_free = lbm_free

    def lbm_sim_create(self, desc):
        self.calls.append(("lbm_sim_create", ()))
        return 0xB200

    def lbm_sim_add_bc(self, *args):
        self.calls.append(("lbm_sim_add_bc", args))
        self._next_bc += 1
        return self._next_bc - 1

    def lbm_sim_stream(self, handle):
        return None

    def lbm_sim_launch_count(self, handle):
        return self.count("lbm_sim_step")


def install(monkeypatch):
    """route pylbm_b200.runtime through a FakeRuntime; kernel launches become recorded no-ops."""
    from pylbm_b200 import runtime

    fake = FakeRuntime()
    monkeypatch.setattr(runtime, "_lib", fake)
    monkeypatch.setattr(runtime, "lib", lambda: fake)

    def launch(self, name, fin, fout, grid, scalars=(), stream=None):
        fake.calls.append(("launch:" + name, (tuple(scalars),)))

    monkeypatch.setattr(runtime.KernelLibrary, "launch", launch)
    return fake
