"""Host front-end: velocity numbering, lean sparse domain, boundary-list scheduling."""
import numpy as np
import pytest


def test_velocity_numbering_convention():
    from pylbm_b200.stencil import Velocity

    # reference convention (pylbm/stencil.py:285-373): D1, D2Q9 shell, D3Q19/27 order
    assert [Velocity(dim=1, num=i).v for i in range(5)] == [[0], [1], [-1], [2], [-2]]
    d2 = [(0, 0), (1, 0), (0, 1), (-1, 0), (0, -1), (1, 1), (-1, 1), (-1, -1), (1, -1)]
    assert [tuple(Velocity(dim=2, num=i).v) for i in range(9)] == d2
    assert tuple(Velocity(dim=2, num=9).v) == (2, 0) and tuple(Velocity(dim=2, num=17).v) == (2, 1)
    d3 = [(0, 0, 0), (0, 0, 1), (0, 0, -1), (0, 1, 0), (0, -1, 0), (1, 0, 0), (-1, 0, 0),
          (0, 1, 1), (0, 1, -1), (0, -1, 1), (0, -1, -1), (1, 0, 1), (1, 0, -1), (-1, 0, 1), (-1, 0, -1),
          (1, 1, 0), (1, -1, 0), (-1, 1, 0), (-1, -1, 0), (1, 1, 1)]
    assert [tuple(Velocity(dim=3, num=i).v) for i in range(20)] == d3
    for dim, count in ((1, 30), (2, 120), (3, 200)):
        for num in range(count):
            v = Velocity(dim=dim, num=num)
            full = v.v + [None] * (3 - dim)
            assert Velocity(vx=full[0], vy=full[1], vz=full[2]).num == num
            assert v.get_symmetric().get_symmetric().num == num


def test_stencil_symmetric_index_per_scheme():
    from pylbm_b200.stencil import Stencil

    st = Stencil({"box": {"x": [0, 1], "y": [0, 1]},
                  "schemes": [{"velocities": list(range(1, 5))}, {"velocities": [0, 5, 6, 7, 8, 1, 3, 2, 4]}]})
    vel = st.get_all_velocities()
    ksym = st.get_symmetric()
    assert np.array_equal(vel[ksym], -vel)
    assert list(st.nv_ptr) == [0, 4, 13] and st.unvtot == 9 and list(st.vmax) == [1, 1]
    assert np.array_equal(vel[st.get_symmetric(axis=0)][:, 0], vel[:, 0])


def test_lean_domain_matches_dense_semantics():
    import pylbm_b200 as lb

    dico = {"box": {"x": [0, 2], "y": [0, 1], "label": [0, 1, 0, -1]},
            "elements": [lb.Circle([0.5, 0.5], 0.2, label=2), lb.Parallelogram([1.2, 0.2], [0.3, 0.0], [0.0, 0.4], label=3)],
            "space_step": 1 / 32, "schemes": [{"velocities": list(range(9))}]}
    dom = lb.Domain(dico)
    assert dom.shape_halo == [66, 34] and dom.shape_in == [64, 32]
    flag, dist = dom.flag, dom.distance
    ioo = dom.in_or_out
    # records only on fluid cells, distances in (0, 1]
    for k in range(dom.stencil.unvtot):
        touched = flag[k] != dom.valin
        assert np.all(ioo[touched] == dom.valin)
        assert np.all((dist[k][touched] > 0) & (dist[k][touched] <= 1))
    # a link is cut exactly when the neighbour cell is solid / outside (periodic faces carry label -1)
    vel = dom.stencil.uvel
    for k in range(1, 9):
        cut = flag[k][1:-1, 1:-1] != dom.valin
        nb = ioo[1 + vel[k][0]: 65 + vel[k][0], 1 + vel[k][1]: 33 + vel[k][1]] == dom.valout
        fluid = ioo[1:-1, 1:-1] == dom.valin
        assert np.array_equal(cut, nb & fluid)
    # sparse query == np.where on the dense view, in C order
    for label in (0, 1, 2, 3):
        for k in range(9):
            cells, d = dom.cells_with_flag(k, label)
            ref = np.where(flag[k] == label)
            assert all(np.array_equal(a, b) for a, b in zip(cells, ref))
            assert np.array_equal(d, dist[k][ref])


def test_slab_topology_regions_and_interface_labels():
    import pylbm_b200 as lb
    from pylbm_b200.domain import SlabTopology

    dico = {"box": {"x": [0, 1], "y": [0, 1], "label": [0, 1, 2, 3]}, "space_step": 1 / 10,
            "schemes": [{"velocities": list(range(9))}]}
    regions = [SlabTopology(2, r, 3).get_region(10, 10) for r in range(3)]
    assert [r[0] for r in regions] == [[0, 4], [4, 7], [7, 10]]     # n // P + (n % P > i), reference mpi_topology.py:82-105
    doms = [lb.Domain(dico, topology=SlabTopology(2, r, 3)) for r in range(3)]
    assert doms[0].box_label == [0, -2, 2, 3] and doms[1].box_label == [-2, -2, 2, 3] and doms[2].box_label == [-2, 1, 2, 3]
    full = lb.Domain(dico)
    x = np.concatenate([d.coords[0] for d in doms])
    np.testing.assert_allclose(x, full.coords[0], rtol=0, atol=1e-15)


def test_schedule_levels_reproduce_sequential_semantics():
    from pylbm_b200.boundary import schedule

    rng = np.random.default_rng(0)
    for trial in range(200):
        n = int(rng.integers(1, 40))
        npos = int(rng.integers(4, 30))
        store = rng.integers(0, npos, size=n)
        l0 = rng.integers(0, npos, size=n)
        l1 = rng.integers(0, npos, size=n)
        snapshot = bool(trial % 2)
        f0 = rng.uniform(size=npos)
        # sequential reference loop
        f = f0.copy()
        src = f0.copy() if snapshot else f
        for i in range(n):
            f[store[i]] = 0.25 * src[l0[i]] + 0.5 * src[l1[i]] + i
        # levelled execution, modelled on the runtime (k_bc): a two-phase level gathers from the
        # CURRENT array and then scatters; a one-phase level works in place, in any order
        order, ptr, two = schedule(store, [l0, l1], snapshot=snapshot)
        g = f0.copy()
        for lev in range(len(ptr) - 1):
            sel = order[ptr[lev]: ptr[lev + 1]]
            if two[lev]:
                vals = 0.25 * g[l0[sel]] + 0.5 * g[l1[sel]] + sel
                g[store[sel]] = vals
            else:
                for i in sel[::-1]:
                    g[store[i]] = 0.25 * g[l0[i]] + 0.5 * g[l1[i]] + i
        assert np.array_equal(f, g), trial


def test_merge_groups_keep_sequential_semantics():
    """consecutive methods are merged into one launch only when the order of the methods cannot
    matter; executing each group's entries in any order must reproduce the method-by-method loop."""
    from pylbm_b200.boundary import merge_groups

    rng = np.random.default_rng(1)
    merged = 0
    for trial in range(300):
        npos = int(rng.integers(20, 80))
        methods = []
        for _ in range(int(rng.integers(1, 6))):
            n = int(rng.integers(1, 6))
            store = rng.choice(npos, size=n, replace=False)
            load = rng.integers(0, npos, size=n)
            single = not np.isin(load, store).any()
            methods.append((store, [load], single))
        ptr = merge_groups(methods)
        assert ptr[0] == 0 and ptr[-1] == len(methods) and np.all(np.diff(ptr) > 0)
        f0 = rng.uniform(size=npos)
        f = f0.copy()
        for store, (load,), _ in methods:          # method after method; a method reads a snapshot
            f[store] = 2.0 * f[load] + 1.0          # (gather, then scatter)
        g = f0.copy()
        for a, b in zip(ptr[:-1], ptr[1:]):
            if b - a == 1:
                store, (load,), _ = methods[a]
                g[store] = 2.0 * g[load] + 1.0
                continue
            merged += 1
            entries = [(s_, l_) for store, (load,), single in methods[a:b] for s_, l_ in zip(store, load)]
            assert all(single for _, _, single in methods[a:b])
            for k in rng.permutation(len(entries)):    # one launch: entries in any order, in place
                g[entries[k][0]] = 2.0 * g[entries[k][1]] + 1.0
        assert np.array_equal(f, g), trial
    assert merged > 20


def test_boundary_lists_without_aliasing_are_single_level():
    import pylbm_b200 as lb
    from pylbm_b200 import cases
    from pylbm_b200.boundary import Boundary, schedule

    dico = cases.lid_cavity_d3q19(n=12)
    dom = lb.Domain(dico)
    bc = Boundary(dom, None, dico)
    (method,) = bc.methods
    method.set_iload()
    method.fix_iload()
    shape = dom.shape_halo
    pos = lambda a: np.ravel_multi_index((a[:, 0], a[:, 1], a[:, 2], a[:, 3]), [19] + shape)
    order, ptr, two = schedule(pos(method.istore), [pos(method.iload[0])])
    assert len(ptr) == 2 and not two[0] and np.array_equal(order, np.arange(order.size))
    assert method.istore.shape == (6 * 12 * 12 * 5 - 0, 4) or method.istore.shape[1] == 4


def test_plan_walls_on_the_parity_workloads():
    """host proof for the fused-kernel walls (boundary.plan_walls) on real boundary lists: the D3Q19
    cavity and the D3Q27 channel qualify (entries with another right-hand side or owned by the outlet
    stay in the list), Bouzidi / Neumann / periodic cases do not."""
    import pylbm_b200 as lb
    from pylbm_b200 import cases
    from pylbm_b200.boundary import Boundary, plan_walls, schedule
    from pylbm_b200.scheme import Scheme
    from pylbm_b200.storage import Layout

    def plan(dico):
        dom, sch = lb.Domain(dico), Scheme(dico)
        bc = Boundary(dom, None, dico)
        nv = int(sch.stencil.nv_ptr[-1])
        lay = Layout(nv, dom.shape_halo, list(dom.stencil.vmax))
        methods = []
        for m in bc.methods:
            m.set_iload()
            m.fix_iload()
            store = lay.positions(m.istore.T)
            loads = [lay.positions(l.T) for l in m.iload]
            order, ptr, two = schedule(store, loads, snapshot=m.snapshot)
            rhs = 0.25 * np.asarray(m.ilabel, dtype=float)[order]     # label-dependent stand-in
            methods.append({"kind": m.kind, "store": store[order], "loads": [l[order] for l in loads],
                            "rhs": rhs, "eligible": len(ptr) == 2 and not two[0]})
        return plan_walls(methods, lay, sch.stencil.get_all_velocities(), sch.stencil.get_symmetric()), methods

    res, methods = plan(cases.lid_cavity_d3q19(n=12))
    walls, masks = res
    # two faces x 12^2 cells x 5 populations, minus the edge entries labelled by the side walls
    assert 2 * 144 * 5 - 2 * 48 <= int(masks[0].sum()) < 2 * 144 * 5
    assert walls["lo_plane"] == 1 and walls["hi_plane"] == 12 and walls["neg_lo"] == walls["neg_hi"] == 0
    assert walls["rhs"][:19].max() == 0.25 and walls["rhs"][:19].min() == 0.0
    res, methods = plan(cases.channel_sphere_d3q27(nx=16, ny=8, nz=8))
    assert res is not None and int(res[1][0].sum()) > 0 and not res[1][1].any() and not res[1][2].any()
    for dico in (cases.karman_d2q9(nx=64, ny=32), cases.heat_d2q5(n=24), cases.advection_d2q13(nx=23, ny=19)):
        assert plan(dico)[0] is None
    # a time-dependent (not eligible) owner refuses the plan
    res, methods = plan(cases.lid_cavity_d3q19(n=12))
    for m in methods:
        m["eligible"] = False
    lay = Layout(19, [14, 14, 14], [1, 1, 1])
    sch = Scheme(cases.lid_cavity_d3q19(n=12))
    assert plan_walls(methods, lay, sch.stencil.get_all_velocities(), sch.stencil.get_symmetric()) is None
