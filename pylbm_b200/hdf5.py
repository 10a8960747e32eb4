"""
HDF5 + XDMF output with the interface of the reference's `pylbm.H5File`
(reference: pylbm/hdf5.py:19-325: `H5File(mpi_topo, filename, path, timestep)`,
`set_grid`, `add_scalar`, `add_vector`, `save`), fed from the device-resident
fields (`sol.m[...]` reads one conserved moment back through the
conserved-only kernel and a pipelined page-locked copy).

The reference writes through h5py; neither h5py nor libhdf5 exists in this
image, so the file is produced directly in the on-disk format the reference's
own files use (HDF5 1.8 defaults as written by h5py, e.g.
tests/reference/*.h5): superblock version 0, a root group stored as a symbol
table (version-1 B-tree + local heap + symbol-table nodes), one version-1
object header per dataset (dataspace v1, IEEE f64 little-endian datatype,
fill value v2, contiguous layout v3, modification time) and contiguous raw
data.  Datasets are `x_0..x_{dim-1}` plus one per scalar / vector field,
stored transposed (`[z, y, x]`, vectors `[z, y, x, 3]`) exactly like
hdf5.py:168, 204-206, 222-228; the XDMF text is the reference's.

Multi-GPU slabs: every rank passes its part; the parts are brought to rank 0
with the all-gather callable of the topology (`Simulation(..., gather=...)`),
where the reference uses mpi4py Send/Recv (hdf5.py:163-178).
"""

import os
import struct
import time

import numpy as np

__all__ = ["H5File", "write_hdf5"]

_UNDEF = 0xFFFFFFFFFFFFFFFF
_LEAF_K = 4         # group leaf node K: a symbol-table node holds up to 2K entries
_INTERNAL_K = 16    # group internal node K: a B-tree node holds up to 2K children
_SNOD_ENTRIES = 2 * _LEAF_K
_ENTRY_SIZE = 40    # symbol table entry with 8-byte offsets


def _pad8(n):
    return (n + 7) // 8 * 8


def _message(mtype, body, flags=0):
    body = body + b"\0" * (_pad8(len(body)) - len(body))
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _dataset_header(shape, data_address, nbytes, mtime):
    """version-1 object header of a contiguous little-endian float64 dataset."""
    rank = len(shape)
    # dataspace message v1: version, rank, flags (1: max dims present), reserved; dims; max dims
    dataspace = struct.pack("<BBB5x", 1, rank, 1) + b"".join(struct.pack("<Q", n) for n in shape) * 2
    # datatype message v1, class 1 (floating point): bit fields 0x20 0x3f 0x00 = little-endian,
    # implied mantissa msb, sign bit at 63; size 8; bit offset 0, precision 64, exponent at 52 (11 bits),
    # mantissa at 0 (52 bits), bias 1023
    datatype = struct.pack("<BBBBI", 0x11, 0x20, 0x3F, 0x00, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
    # fill value message v2: allocate late (2), write fill value if set (2), fill value defined (1), size 0
    fill = struct.pack("<BBBBI", 2, 2, 2, 1, 0)
    # data layout message v3, class 1 (contiguous): address, size
    layout = struct.pack("<BBQQ", 3, 1, data_address, nbytes)
    # modification time message v1
    mod = struct.pack("<B3xI", 1, int(mtime) & 0xFFFFFFFF)
    messages = [
        _message(0x0001, dataspace), _message(0x0003, datatype, flags=1), _message(0x0005, fill, flags=1),
        _message(0x0008, layout), _message(0x0012, mod),
    ]
    payload = b"".join(messages)
    # header prefix: version 1, reserved, number of messages, reference count, size of the message block
    return struct.pack("<BxHII4x", 1, len(messages), 1, len(payload)) + payload


def write_hdf5(filename, datasets, mtime=None):
    """
    Write `datasets` (name -> float64 ndarray, any rank >= 1) as the root-group datasets of a new
    HDF5 file.  Layout: superblock | root object header | B-tree node | local heap | symbol-table
    nodes | dataset object headers | raw data (8-byte aligned, contiguous).
    """
    mtime = time.time() if mtime is None else mtime
    names = sorted(datasets)                      # symbol tables are ordered by name
    if len(names) > _SNOD_ENTRIES * 2 * _INTERNAL_K:
        raise ValueError("too many datasets for a single-level group B-tree (%d)" % len(names))
    arrays = {k: np.ascontiguousarray(datasets[k], dtype="<f8") for k in names}
    for k, a in arrays.items():
        if a.ndim < 1:
            arrays[k] = a.reshape(1)
        if not k or "/" in k or "\0" in k:
            raise ValueError("invalid dataset name %r" % k)

    # ---- local heap: "" at offset 0, then the names, 8-byte aligned ------------------------------
    heap = bytearray(8)
    name_offset = {}
    for k in names:
        name_offset[k] = len(heap)
        raw = k.encode() + b"\0"
        heap += raw + b"\0" * (_pad8(len(raw)) - len(raw))
    free_head = 1                                   # H5HL_FREE_NULL: no free block
    if len(heap) < 88:                              # libhdf5's default first data segment is 88 bytes
        free_at = len(heap)
        size = 88 - free_at
        if size >= 16:
            heap += struct.pack("<QQ", 1, size) + b"\0" * (size - 16)
            free_head = free_at
        else:
            heap += b"\0" * size
    heap = bytes(heap)

    # ---- addresses -----------------------------------------------------------------------------------
    groups = [names[i: i + _SNOD_ENTRIES] for i in range(0, len(names), _SNOD_ENTRIES)] or [[]]
    pos = 96                                         # after the superblock
    root_header = pos
    pos += 16 + 24
    btree = pos
    pos += 24 + (2 * _INTERNAL_K + 1) * 8 + 2 * _INTERNAL_K * 8
    heap_header = pos
    pos += 32
    heap_data = pos
    pos += len(heap)
    snods = []
    for _ in groups:
        snods.append(pos)
        pos += 8 + _SNOD_ENTRIES * _ENTRY_SIZE
    header_addr, header_bytes = {}, {}
    for k in names:
        header_addr[k] = pos
        pos += len(_dataset_header(arrays[k].shape, 0, 0, mtime))
    data_addr = {}
    for k in names:
        pos = _pad8(pos)
        data_addr[k] = pos
        pos += arrays[k].nbytes
    eof = pos
    for k in names:
        header_bytes[k] = _dataset_header(arrays[k].shape, data_addr[k], arrays[k].nbytes, mtime)

    # ---- metadata blocks -----------------------------------------------------------------------------
    def entry(link_offset, header, cache_type=0, scratch=b""):
        return struct.pack("<QQI4x", link_offset, header, cache_type) + scratch + b"\0" * (16 - len(scratch))

    superblock = (
        b"\x89HDF\r\n\x1a\n"
        + struct.pack("<BBBBBBBx", 0, 0, 0, 0, 0, 8, 8)         # versions, size of offsets / lengths
        + struct.pack("<HHI", _LEAF_K, _INTERNAL_K, 0)           # group leaf / internal K, consistency flags
        + struct.pack("<QQQQ", 0, _UNDEF, eof, _UNDEF)           # base, free-space info, end of file, driver info
        + entry(0, root_header, 1, struct.pack("<QQ", btree, heap_header))
    )
    assert len(superblock) == 96
    root = struct.pack("<BxHII4x", 1, 1, 1, 24) + _message(0x0011, struct.pack("<QQ", btree, heap_header))
    keys = [0] + [name_offset[g[-1]] if g else 0 for g in groups]
    node = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(groups), _UNDEF, _UNDEF)
    for i, addr in enumerate(snods):
        node += struct.pack("<QQ", keys[i], addr)
    node += struct.pack("<Q", keys[len(groups)])
    node += b"\0" * (24 + (2 * _INTERNAL_K + 1) * 8 + 2 * _INTERNAL_K * 8 - len(node))
    heap_hdr = b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), free_head, heap_data)

    with open(filename, "wb") as fh:
        fh.write(superblock)
        fh.write(root)
        fh.write(node)
        fh.write(heap_hdr)
        fh.write(heap)
        for g in groups:
            block = b"SNOD" + struct.pack("<BxH", 1, len(g))
            for k in g:
                block += entry(name_offset[k], header_addr[k])
            block += b"\0" * (8 + _SNOD_ENTRIES * _ENTRY_SIZE - len(block))
            fh.write(block)
        for k in names:
            fh.write(header_bytes[k])
        for k in names:
            fh.write(b"\0" * (data_addr[k] - fh.tell()))
            fh.write(arrays[k].tobytes())
    return eof


class H5File:
    """
    Drop-in for `pylbm.H5File` (reference: hdf5.py:19-55).

    h5 = H5File(sol.domain.mpi_topo, "lid_cavity", "./lid_cavity", im)
    h5.set_grid(x, y, z); h5.add_scalar("mass", sol.m[mass]); h5.add_vector("velocity", [qx, qy, qz]); h5.save()
    """

    def __init__(self, mpi_topo, filename, path="", timestep=None, init_xdmf=False):
        prefix = "_{}".format(timestep) if timestep is not None else ""
        self.path = path
        name, _ = os.path.splitext(filename)
        self.filename = name + prefix
        self.h5filename = name + prefix + ".h5"
        self.mpi_topo = mpi_topo
        self.rank = int(getattr(mpi_topo, "rank", 0) or 0)
        self.size = int(getattr(mpi_topo, "size", 1) or 1)
        self._gather = getattr(mpi_topo, "gather", None)
        if self.size > 1 and self._gather is None:
            raise RuntimeError(
                "H5File on a decomposed run needs the all-gather callable of the topology "
                "(Simulation(..., gather=...)); the reference uses mpi4py Send/Recv here"
            )
        if self.rank == 0 and path and not os.path.exists(path):
            os.makedirs(path, exist_ok=True)
        self.origin = self.dx = self.dim = self.n = self.region = self.global_size = None
        self.scalars, self.vectors = {}, {}
        self._datasets = {}
        self._order = []
        self._init_xdmf = init_xdmf

    # ------------------------------------------------------------------
    def _all(self, obj):
        return [obj] if self.size == 1 else list(self._gather(obj))

    def set_grid(self, x, y=None, z=None):
        """(reference: hdf5.py:57-121) coordinates of the points owned by this rank."""
        coords = [np.asarray(c, dtype=float) for c in (x, y, z) if c is not None]
        self.dim = len(coords)
        self.origin = [c[0] for c in coords]
        self.dx = [c[1] - c[0] if c.size > 1 else 0.0 for c in coords]
        parts = self._all(coords)
        # slabs are cut along x only (SlabTopology): concatenate axis 0, the other axes are whole
        full = [np.concatenate([p[0] for p in parts])] + [parts[0][d] for d in range(1, self.dim)]
        self.region = [[0] + list(np.cumsum([p[0].size for p in parts]))] + [[0, full[d].size] for d in range(1, self.dim)]
        self.global_size = [int(c.size) for c in full]
        self.n = list(self.global_size)
        if self.rank == 0:
            for d, c in enumerate(full):
                self._put("x_%d" % d, c)

    def _put(self, name, array):
        if name not in self._datasets:
            self._order.append(name)
        self._datasets[name] = array

    def _assemble(self, data):
        """local part(s) -> global array on rank 0, stored transposed like hdf5.py:168."""
        parts = self._all(np.ascontiguousarray(data, dtype=float))
        if self.rank != 0:
            return None
        full = parts[0] if len(parts) == 1 else np.concatenate(parts, axis=0)
        if list(full.shape) != self.global_size:
            raise ValueError("field of shape %s does not match the grid %s" % (full.shape, self.global_size))
        return full.T

    def add_scalar(self, name, f, *fargs):
        """(reference: hdf5.py:180-213)"""
        if self.global_size is None:
            raise RuntimeError("set_grid must be called before add_scalar")
        data = f if isinstance(f, np.ndarray) else f(*fargs)
        full = self._assemble(data)
        if self.rank == 0:
            self._put(name, full)
            self.scalars[name] = self.h5filename + ":/" + name

    def add_vector(self, name, f, *fargs):
        """(reference: hdf5.py:215-248) components go to the last axis; always 3 wide."""
        if self.global_size is None:
            raise RuntimeError("set_grid must be called before add_vector")
        datas = f if isinstance(f, (list, tuple)) else f(*fargs)
        comps = [self._assemble(d) for d in datas]
        if self.rank == 0:
            out = np.zeros(tuple(self.global_size[::-1]) + (3,))
            for i, c in enumerate(comps):
                out[..., i] = c
            self._put(name, out)
            self.vectors[name] = self.h5filename + ":/" + name

    def save(self):
        """write <path>/<name>.h5 and <path>/<name>.xdmf on rank 0 (reference: hdf5.py:250-325)."""
        if self.rank != 0:
            return
        base = os.path.join(self.path, self.filename) if self.path else self.filename
        write_hdf5(base + ".h5", {k: self._datasets[k] for k in self._order})
        gs = self.global_size
        with open(base + ".xdmf", "w") as out:
            out.write('<?xml version="1.0" ?>\n<!DOCTYPE Xdmf SYSTEM "Xdmf.dtd" []>\n<Xdmf>\n <Domain>\n')
            topo, geo = ("2DRectMesh", "VXVY") if self.dim == 2 else ("3DRectMesh", "VXVYVZ")
            out.write('  <Grid Name="Structured Grid" GridType="Uniform">\n')
            out.write('   <Topology TopologyType="%s" NumberOfElements="%s"/>\n' % (topo, " ".join(map(str, gs))))
            out.write('   <Geometry GeometryType="%s">\n' % geo)
            for d in range(self.dim):
                out.write('    <DataItem Format="HDF" Dimensions="%d">\n     %s.h5:/x_%d\n    </DataItem>\n'
                          % (gs[d], self.filename, d))
            out.write("   </Geometry>\n")
            dims = " ".join(map(str, gs[::-1]))
            for k, v in self.scalars.items():
                out.write('   <Attribute Name="%s" AttributeType="Scalar" Center="Node">\n' % k)
                out.write('    <DataItem Format="HDF" Dimensions="%s">\n     %s\n    </DataItem>\n   </Attribute>\n' % (dims, v))
            for k, v in self.vectors.items():
                out.write('   <Attribute Name="%s" AttributeType="Vector" Center="Node">\n' % k)
                out.write('    <DataItem Format="HDF" Dimensions="%s %d">\n     %s\n    </DataItem>\n   </Attribute>\n'
                          % (dims, self.dim, v))
            out.write("  </Grid>\n </Domain>\n</Xdmf>\n")
