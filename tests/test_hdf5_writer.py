"""
Output path: `pylbm_b200.H5File` (reference: pylbm/hdf5.py:19-325) writes HDF5 + XDMF without h5py.
The files are checked with an independent minimal reader (tools/make_golden.py:H5Lite, the one that
reads the reference's golden files) and, message by message, against a file written by libhdf5 through
the reference (tests/reference/test1D_advection.h5) when the reference checkout is present.  CPU only.
"""
import os
import struct
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def _reader():
    from make_golden import H5Lite

    return H5Lite


def test_round_trip_many_datasets(tmp_path):
    from pylbm_b200.hdf5 import write_hdf5

    rng = np.random.default_rng(0)
    data = {"x_0": rng.uniform(size=7), "rho": rng.uniform(size=(5, 7)), "velocity": rng.uniform(size=(3, 5, 7, 3))}
    for i in range(20):                      # more than one symbol-table node (8 entries each)
        data["field_%02d" % i] = rng.uniform(size=(i + 1, 3))
    path = str(tmp_path / "many.h5")
    eof = write_hdf5(path, data)
    assert os.path.getsize(path) == eof
    back = _reader()(path).datasets()
    assert sorted(back) == sorted(data)
    for key, val in data.items():
        assert back[key].shape == val.shape and np.array_equal(back[key], val), key


def test_h5file_interface_matches_reference_layout(tmp_path):
    import pylbm_b200
    from pylbm_b200.hdf5 import H5File

    x, y, z = np.linspace(0, 1, 6), np.linspace(0, 2, 4), np.linspace(-1, 1, 5)
    rng = np.random.default_rng(1)
    mass = rng.uniform(size=(6, 4, 5))
    q = [rng.uniform(size=(6, 4, 5)) for _ in range(3)]
    h5 = H5File(None, "lid_cavity", str(tmp_path / "out"), 12)
    h5.set_grid(x, y, z)
    h5.add_scalar("mass", mass)
    h5.add_vector("velocity", lambda a, b, c: [a, b, c], *q)
    h5.save()
    assert pylbm_b200.H5File is H5File
    back = _reader()(str(tmp_path / "out" / "lid_cavity_12.h5")).datasets()
    # fields are stored transposed ([z, y, x], vectors [z, y, x, 3]): reference hdf5.py:168, 204-206
    assert np.array_equal(back["mass"], mass.T)
    for i in range(3):
        assert np.array_equal(back["velocity"][..., i], q[i].T)
    assert np.array_equal(back["x_0"], x) and np.array_equal(back["x_1"], y) and np.array_equal(back["x_2"], z)
    xdmf = open(str(tmp_path / "out" / "lid_cavity_12.xdmf")).read()
    assert 'TopologyType="3DRectMesh" NumberOfElements="6 4 5"' in xdmf
    assert "lid_cavity_12.h5:/velocity" in xdmf and 'Dimensions="5 4 6 3"' in xdmf


def _structure(path):
    """superblock constants + per dataset the list of (message type, flags, body) with addresses,
    sizes of the file and times blanked: what must be equal between two writers."""
    b = open(path, "rb").read()
    assert b[:8] == b"\x89HDF\r\n\x1a\n"
    sb = {"versions": b[8:13], "sizes": b[13:15], "k": b[16:20], "flags": b[20:24], "base": b[24:32],
          "root_cache_type": b[56 + 16: 56 + 20]}
    reader = _reader()(path)
    names = {}
    reader._walk(reader.btree, reader.heap, names)
    out = {}
    for name, header in names.items():
        version, nmsg, refcount = b[header], struct.unpack_from("<H", b, header + 2)[0], struct.unpack_from("<I", b, header + 4)[0]
        msgs = []
        for mtype, body, msize in reader._messages(header):
            raw = bytearray(b[body: body + msize])
            flags = b[body - 4]
            if mtype == 0x0008:          # layout: blank the data address
                raw[2:10] = b"\0" * 8
            if mtype == 0x0012:          # modification time
                raw[4:8] = b"\0" * 4
            if mtype == 0x0000:          # NIL padding of libhdf5's pre-allocated header block
                continue
            msgs.append((mtype, flags, bytes(raw)))
        out[name] = (version, refcount, msgs)
    return sb, out


@pytest.mark.skipif(not os.path.exists("/root/reference/tests/reference/test1D_advection.h5"),
                    reason="needs the reference checkout (a file written by libhdf5)")
def test_same_structure_as_a_file_written_by_libhdf5(tmp_path):
    from pylbm_b200.hdf5 import write_hdf5

    golden = "/root/reference/tests/reference/test1D_advection.h5"
    data = _reader()(golden).datasets()
    mine = str(tmp_path / "mine.h5")
    write_hdf5(mine, data)
    sb_ref, ds_ref = _structure(golden)
    sb_mine, ds_mine = _structure(mine)
    assert sb_ref == sb_mine
    assert sorted(ds_ref) == sorted(ds_mine) == ["u", "x_0"]
    for name in ds_ref:
        assert ds_ref[name] == ds_mine[name], name
    # 2-D golden as well (rank-2 dataspace)
    golden2 = "/root/reference/tests/reference/test2D_lid_driven_cavity.h5"
    data2 = _reader()(golden2).datasets()
    mine2 = str(tmp_path / "mine2.h5")
    write_hdf5(mine2, data2)
    assert _structure(golden2)[1] == _structure(mine2)[1]
    for key, val in data2.items():
        assert np.array_equal(_reader()(mine2).datasets()[key], val)


def test_h5file_gathers_slabs_on_rank_0(tmp_path):
    """two ranks (threads here) own x-slabs; the parts reach rank 0 through the topology's all-gather
    callable, like the reference's Send/Recv to rank 0 (hdf5.py:163-178)."""
    import threading

    from pylbm_b200.domain import SlabTopology
    from pylbm_b200.hdf5 import H5File

    world = 2
    x = np.linspace(0.0, 1.0, 10)
    y = np.linspace(0.0, 0.5, 4)
    rng = np.random.default_rng(3)
    field = rng.uniform(size=(10, 4))
    cuts = [0, 6, 10]
    box, lock, barrier = {}, threading.Lock(), threading.Barrier(world)
    errors = []

    def make_gather(rank):
        def gather(obj):
            with lock:
                box[rank] = obj
            barrier.wait()
            out = [box[r] for r in range(world)]
            barrier.wait()
            return out
        return gather

    def worker(rank):
        try:
            topo = SlabTopology(2, rank, world)
            topo.gather = make_gather(rank)
            h5 = H5File(topo, "slabs", str(tmp_path))
            sl = slice(cuts[rank], cuts[rank + 1])
            h5.set_grid(x[sl], y)
            h5.add_scalar("rho", field[sl])
            h5.add_vector("q", [field[sl], 2 * field[sl]])
            h5.save()
        except Exception as exc:      # pragma: no cover
            errors.append(exc)
            barrier.abort()

    threads = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(60)
    assert not errors, errors
    back = _reader()(str(tmp_path / "slabs.h5")).datasets()
    assert np.array_equal(back["x_0"], x) and np.array_equal(back["x_1"], y)
    assert np.array_equal(back["rho"], field.T)
    assert np.array_equal(back["q"][..., 1], 2 * field.T) and not back["q"][..., 2].any()
