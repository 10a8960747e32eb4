"""
Boundary conditions: index lists (host) and their device application.

The lists are part of the parity contract and reproduce the reference's
construction order exactly (reference: pylbm/boundary.py:23-47 per label and
velocity, :78-160 concatenation order = labels ascending -> method dictionary
order -> velocities of the scheme -> C order of the cells; :411-419, 508-534,
777-784, 834-842, 858-866, 882-890 load indices; :226-235 final int32 layout;
:421-427, 561-567, 637-643, 698-704 right-hand sides; :238-321 wall
equilibrium).  They are built from the sparse records of domain.py instead of
`np.where` over dense [unvtot, nx, ny, nz] arrays.

Application differs by design: the reference runs one generated *sequential*
loop per method (boundary.py:462-464 ...).  Here each method is one (or two)
CUDA kernels over device-resident position lists.  To keep the sequential
semantics, `schedule()` finds the (rare) entries that read or overwrite what
another entry of the same method stores and splits the list into levels;
levels with read/write aliasing gather into a scratch buffer before they
scatter.  Bouzidi bounce-back reads a snapshot taken when the method starts
(the reference's `fcopy`, boundary.py:545-549) = one gather-then-scatter level.
"""

import collections
import ctypes

import numpy as np

from . import runtime as rt
from .domain import stencil_uvel
from .storage import HostArray

__all__ = [
    "Boundary", "BoundaryMethod", "BounceBack", "BouzidiBounceBack", "AntiBounceBack",
    "BouzidiAntiBounceBack", "Neumann", "NeumannX", "NeumannY", "NeumannZ", "schedule", "merge_groups", "plan_walls", "plan_tasks", "plan_aa",
]


def _member(sorted_values, x):
    """boolean mask: x[i] is in `sorted_values` (ascending).  One binary search per query instead of
    np.isin's sort of both arrays: the boundary lists of a 512^3 lattice have 8 M entries."""
    x = np.asarray(x)
    if sorted_values.size == 0 or x.size == 0:
        return np.zeros(x.shape, dtype=bool)
    idx = np.searchsorted(sorted_values, x)
    idx[idx == sorted_values.size] = sorted_values.size - 1
    return sorted_values[idx] == x


def merge_groups(methods, max_group=12):
    """
    Greedy grouping of consecutive boundary methods into merged launches (lbm_sim_bc_groups).
    `methods`: list of (store positions, [load position arrays], single_level) in application order.
    A method joins the current group when it is a single gather-free level and no position it stores
    is read or stored by the group, and no position it reads is stored by the group: the sequential
    order of the methods is then irrelevant.  Returns group_ptr (len = ngroups + 1).
    """
    group_ptr = [0]
    stores = loads = None
    size = 0
    for i, (st, lds, single) in enumerate(methods):
        st = np.asarray(st, dtype=np.int64)
        ld = np.concatenate([np.asarray(l, dtype=np.int64) for l in lds]) if lds else np.zeros(0, dtype=np.int64)
        joinable = False
        if i > 0 and single and prev_single and size < max_group:
            joinable = not (np.isin(st, stores).any() or np.isin(st, loads).any() or np.isin(ld, stores).any())
        if i > 0 and not joinable:
            group_ptr.append(i)
            stores = loads = None
            size = 0
        stores = st if stores is None else np.concatenate([stores, st])
        loads = ld if loads is None else np.concatenate([loads, ld])
        size += 1
        prev_single = single if size == 1 else (prev_single and single)
    group_ptr.append(len(methods))
    return np.asarray(group_ptr, dtype=np.int32)


def plan_walls(methods, array, velocities, symmetric):
    """
    Decide whether the bounce-back walls normal to the FASTEST axis can be applied by the fused kernel
    itself (include/lbmk.h: lbmk_walls) and which list entries that replaces.

    methods    [{"kind", "store", "loads": [positions, ...], "rhs", "eligible"}] in application order
               (device positions of `array`, device order; `eligible` is False for methods whose
               right-hand side changes in time or that are not a single gather-free level)
    array      layout of the populations (storage.Layout / DeviceArray)
    velocities integer lattice velocities [Q, dim];  symmetric: index of the opposite population

    With walls on, every interior cell c of the two planes next to the faces stores, for every
    population k moving towards the wall, `+-f_k(c) + rhs[k]` at W(c, k) = (sym k, c + v_k) right after
    it has computed f_k(c), and the periodic images along the axis are no longer produced.  That equals
    the reference (periodic update, then the methods in order, boundary.py / simulation.py:373-390) iff,
    for each face:
      * the entries `f[sym k](c + v_k) = +-f[k](c) + rhs` with c on the plane are all of one kind
        (bounce-back or anti-bounce-back) and belong to eligible methods; the kernel uses the most
        frequent right-hand side per population, and the entries that have it are REPLACED (removed
        from the list);
      * every other position W(c, k) the kernel writes (an entry with another right-hand side, e.g. an
        edge cell labelled by the neighbouring face, or no bounce-back entry at all, e.g. next to an
        outlet) is stored by a remaining list entry, which overwrites the kernel's value at the next
        step, and is not read by any list entry;
      * a replaced position is read only by entries of LATER methods (they saw the stored value in the
        reference too), is stored by no other entry, and what a replaced entry reads is stored by nobody;
      * no remaining entry reads a ghost cell of that side other than a replaced position.
    Returns None, or (walls dict, [boolean mask of the replaced entries, one per method]).
    """
    dim = array.dim
    n = array.canonical_n
    w = array.canonical_vmax
    if w[2] != 1 or n[2] < 4 or len(velocities) > 64:
        return None
    vel = np.zeros((len(velocities), 3), dtype=np.int64)
    vel[:, 3 - dim:] = np.asarray(velocities, dtype=np.int64)[:, :dim]
    if np.abs(vel[:, 2]).max() > 1:
        return None
    pstride, pitch, lead = array.pstride, array.pitch, array.lead

    def decode(pos):
        pos = np.asarray(pos, dtype=np.int64)
        k = pos // pstride
        r = pos - k * pstride - lead
        row = r // pitch
        return k, row // n[1], row % n[1], r - row * pitch

    planes = {"lo": w[2], "hi": n[2] - w[2] - 1}
    toward = {"lo": np.nonzero(vel[:, 2] < 0)[0], "hi": np.nonzero(vel[:, 2] > 0)[0]}
    if toward["lo"].size == 0 or toward["hi"].size == 0:
        return None
    sym = np.asarray(symmetric, dtype=np.int64)
    voff = (vel[:, 0] * n[1] + vel[:, 1]) * pitch + vel[:, 2]          # element offset of c + v_k
    i0, i1 = np.meshgrid(np.arange(w[0], n[0] - w[0]), np.arange(w[1], n[1] - w[1]), indexing="ij")
    rows = (i0 * n[1] + i1).ravel() * pitch + lead

    masks = [np.zeros(len(m["store"]), dtype=bool) for m in methods]
    walls = {"lo_plane": int(planes["lo"]), "hi_plane": int(planes["hi"]), "rhs": np.zeros(64)}
    written = []
    for side in ("lo", "hi"):
        found_kind, per_method = None, []
        for im, m in enumerate(methods):
            if m["kind"] not in (rt.BC_BOUNCE_BACK, rt.BC_ANTI_BOUNCE_BACK) or len(m["store"]) == 0:
                continue
            store, load, rhs = np.asarray(m["store"]), np.asarray(m["loads"][0]), np.asarray(m["rhs"])
            k, c0, c1, c2 = decode(load)
            cand = (c2 == planes[side]) & np.isin(k, toward[side])
            cand &= (c0 >= w[0]) & (c0 < n[0] - w[0]) & (c1 >= w[1]) & (c1 < n[1] - w[1])
            if not cand.any():
                continue
            if not m["eligible"]:
                return None
            kc = k[cand]                     # the store must be the bounced population at c + v_k
            if not np.array_equal(sym[kc] * pstride + (load[cand] - kc * pstride) + voff[kc], store[cand]):
                return None
            if found_kind is None:
                found_kind = m["kind"]
            elif found_kind != m["kind"]:
                return None
            per_method.append((im, cand, k, rhs))
        if found_kind is None:
            return None
        walls["neg_" + side] = 1 if found_kind == rt.BC_ANTI_BOUNCE_BACK else 0
        for kk in toward[side]:
            vals = [rhs[cand & (k == kk)] for _, cand, k, rhs in per_method]
            vals = np.concatenate(vals) if vals else np.zeros(0)
            if vals.size == 0:
                return None
            uniq, cnt = np.unique(vals, return_counts=True)
            walls["rhs"][int(kk)] = uniq[np.argmax(cnt)]
            for im, cand, k, rhs in per_method:
                masks[im] |= cand & (k == kk) & (rhs == walls["rhs"][int(kk)])
            written.append(sym[kk] * pstride + rows + planes[side] + voff[kk])
    written = np.concatenate(written)

    empty = np.zeros(0, dtype=np.int64)
    mine_store = np.concatenate([np.asarray(m["store"])[mask] for m, mask in zip(methods, masks)] + [empty])
    mine_owner = np.concatenate([np.full(int(mask.sum()), im) for im, mask in enumerate(masks)] + [empty])
    mine_load = np.concatenate([np.asarray(m["loads"][0])[mask] for m, mask in zip(methods, masks)] + [empty])
    other_store = np.sort(np.concatenate([np.asarray(m["store"])[~mask] for m, mask in zip(methods, masks)] + [empty]))
    if mine_store.size == 0:
        return None
    order = np.argsort(mine_store)
    sorted_store, sorted_owner = mine_store[order], mine_owner[order]
    if (sorted_store[1:] == sorted_store[:-1]).any():
        return None
    if _member(other_store, mine_store).any():
        return None
    if _member(other_store, mine_load).any() or _member(sorted_store, mine_load).any():
        return None
    # what the kernel writes without a replaced entry behind it must be overwritten by the list
    loose = np.sort(written[~_member(sorted_store, written)])
    if not _member(other_store, loose).all():
        return None
    for im, (m, mask) in enumerate(zip(methods, masks)):
        for j, l in enumerate(m["loads"]):
            l = np.asarray(l)
            l = l[~mask] if j == 0 else l
            if l.size == 0:
                continue
            if _member(loose, l).any():
                return None
            hit = _member(sorted_store, l)
            if hit.any():                      # only later methods may read a replaced position
                owner = sorted_owner[np.searchsorted(sorted_store, l[hit])]
                if (owner >= im).any():
                    return None
            _, _, _, c2 = decode(l[~hit])
            if ((c2 < w[2]) | (c2 >= n[2] - w[2])).any():
                return None
    return walls, masks


def plan_aa(methods, array, velocities, symmetric):
    """
    Boundary lists of the ODD steps of in-place streaming (AA pattern, include/lbm_b200.h:
    lbm_sim_set_aa).  After an even step the array holds the new population k of cell x in the slot
    (kbar, x + v_k); every position (k, y) of a list is therefore read / written at (kbar, y + v_k) -- a
    bijection of the slots, so levels, snapshots and the order of the methods carry over unchanged.

    methods    [{"store", "loads": [positions, ...], ...}] device positions, device order
    Returns the list of transformed methods (same dictionaries with "store" / "loads" replaced), or None
    when a transformed position would leave the array (an entry that touches the outermost ghost layer
    against its own velocity: not produced by the reference's list builders).
    """
    dim = array.dim
    n = array.canonical_n
    pstride, pitch, lead = array.pstride, array.pitch, array.lead
    vel = np.zeros((len(velocities), 3), dtype=np.int64)
    vel[:, 3 - dim:] = np.asarray(velocities, dtype=np.int64)[:, :dim]
    sym = np.asarray(symmetric, dtype=np.int64)
    voff = (vel[:, 0] * n[1] + vel[:, 1]) * pitch + vel[:, 2]

    def transform(pos):
        pos = np.asarray(pos, dtype=np.int64)
        k = pos // pstride
        r = pos - k * pstride - lead
        row = r // pitch
        c = [row // n[1], row % n[1], r - row * pitch]
        for a in range(3):
            t = c[a] + vel[k, a]
            if ((t < 0) | (t >= n[a])).any():
                return None
        return sym[k] * pstride + (pos - k * pstride) + voff[k]

    out = []
    for m in methods:
        store = transform(m["store"])
        loads = [transform(l) for l in m["loads"]]
        if store is None or any(l is None for l in loads):
            return None
        odd = dict(m)
        odd["store"], odd["loads"] = store, loads
        out.append(odd)
    return out


def plan_tasks(methods, array, velocities):
    """
    Task table for the fused kernel (include/lbmk.h: lbmk_tasks): every boundary entry is evaluated by the
    128-thread block that owns the cell pulling its value, instead of by a list kernel.

    methods    [{"kind", "store", "loads": [positions, ...], "dist" (or None), "single": bool}] in
               application order (device positions of `array`, device order of the entries)
    array      layout of the populations (storage.Layout / DeviceArray), with `tx` threads of a block
               along the fastest axis
    velocities integer lattice velocities [Q, dim]

    The reference applies the methods one after the other on F, then pulls (simulation.py:373-420).  An
    entry stores population k of an outside cell c_out, which exactly one pull reads: cell c_out + v_k.
    Evaluating the entries from the INPUT array at pull time gives the same populations iff
      * every method is one gather-free level (boundary.schedule), and
      * no entry -- of any method -- reads a position that an entry stores (then every value only
        depends on populations the previous fused launch wrote), except a Neumann entry copying a
        position stored EARLIER in the sequence (the edge entries of an outlet next to a wall): it
        inherits the definition of the entry it copies;
    of several entries storing the same position the last one in application order wins, and an entry
    whose value no interior cell pulls is dropped.  Returns None when the conditions do not hold, else
    a dict of arrays sorted by block.
    """
    dim = array.dim
    n = array.canonical_n
    w = array.canonical_vmax
    if len(velocities) > 64 or any(not m["single"] for m in methods):
        return None
    sizes = [len(m["store"]) for m in methods]
    if sum(sizes) == 0 or sum(sizes) >= 2 ** 31:
        return None
    store = np.concatenate([np.asarray(m["store"], dtype=np.int64) for m in methods])
    ibc = np.concatenate([np.full(sz, i, dtype=np.int32) for i, sz in enumerate(sizes)])
    entry = np.concatenate([np.arange(sz, dtype=np.int64) for sz in sizes])
    kind = np.concatenate([np.full(sz, m["kind"], dtype=np.int64) for m, sz in zip(methods, sizes)])
    l0 = np.concatenate([np.asarray(m["loads"][0], dtype=np.int64) for m in methods])
    l1 = np.concatenate([np.asarray(m["loads"][1], dtype=np.int64) if len(m["loads"]) > 1
                         else np.zeros(sz, dtype=np.int64) for m, sz in zip(methods, sizes)])
    dist = np.concatenate([np.asarray(m["dist"], dtype=np.float64) if m.get("dist") is not None
                           else np.zeros(sz) for m, sz in zip(methods, sizes)])
    two = (kind == rt.BC_BOUZIDI_BOUNCE_BACK) | (kind == rt.BC_BOUZIDI_ANTI_BOUNCE_BACK)
    # positions stored, with their LAST writer in application order (global entry index)
    upos, last_rev = np.unique(store[::-1], return_index=True)
    last_writer = store.size - 1 - last_rev

    def reads_a_store(idx):
        hit = _member(upos, l0[idx])
        hit |= two[idx] & _member(upos, l1[idx])
        return hit

    chained = reads_a_store(np.arange(store.size))
    # A Neumann entry copying a position that an EARLIER entry stored (an outlet next to a wall) pulls
    # what that entry computes: it inherits its definition.  Anything else that reads a stored position
    # (an interpolation through a stored value, a position stored LATER in the sequence, a Bouzidi
    # snapshot) keeps the list kernels.
    pending = np.nonzero(chained)[0]
    while pending.size:
        if (kind[pending] != rt.BC_NEUMANN).any():
            return None
        writer = last_writer[np.searchsorted(upos, l0[pending])]
        if (writer >= pending).any():
            return None
        ready = ~chained[writer]
        if not ready.any():
            return None
        dst, src = pending[ready], writer[ready]
        for arr in (kind, l0, l1, dist, ibc, entry, two):
            arr[dst] = arr[src]
        chained[dst] = False
        pending = pending[~ready]
    # last writer of every position
    _, last = np.unique(store[::-1], return_index=True)
    keep = np.sort(store.size - 1 - last)
    # the cell that pulls (k, c_out) is c_out + v_k
    pstride, pitch, lead = array.pstride, array.pitch, array.lead
    pos = store[keep]
    k = pos // pstride
    r = pos - k * pstride - lead
    row = r // pitch
    c = [row // n[1], row % n[1], r - row * pitch]
    vel = np.zeros((len(velocities), 3), dtype=np.int64)
    vel[:, 3 - dim:] = np.asarray(velocities, dtype=np.int64)[:, :dim]
    c = [c[a] + vel[k, a] for a in range(3)]
    inside = np.ones(pos.size, dtype=bool)
    for a in range(3):
        inside &= (c[a] >= w[a]) & (c[a] < n[a] - w[a])
    keep, k = keep[inside], k[inside]
    c = [ca[inside] for ca in c]
    tx = int(array.tx)
    ty = 128 // tx
    ngx = -(-(n[2] - 2 * w[2]) // tx)
    ngy = -(-(n[1] - 2 * w[1]) // ty)
    r2, r1 = c[2] - w[2], c[1] - w[1]
    block = ((c[0] - w[0]) * ngy + r1 // ty) * ngx + r2 // tx
    thread = (r1 % ty) * tx + r2 % tx
    nblocks = int((n[0] - 2 * w[0]) * ngy * ngx)
    if nblocks + 1 >= 2 ** 31:
        return None
    order = np.argsort(block, kind="stable")
    keep, k, block, thread = keep[order], k[order], block[order], thread[order]
    block_ptr = np.zeros(nblocks + 1, dtype=np.int32)
    np.cumsum(np.bincount(block, minlength=nblocks), out=block_ptr[1:])
    code = (thread | (k << 8) | (kind[keep] << 16)).astype(np.uint32)
    return {
        "ntasks": int(keep.size), "nblocks": nblocks, "block_ptr": block_ptr, "code": code,
        "l0": np.ascontiguousarray(l0[keep]), "l1": np.ascontiguousarray(l1[keep]),
        "dist": np.ascontiguousarray(dist[keep]), "ibc": np.ascontiguousarray(ibc[keep]),
        "entry": np.ascontiguousarray(entry[keep]), "ngroups_y": int(ngy), "ngroups_x": int(ngx), "tx": tx,
        "nentries": int(store.size),
    }


def schedule(store, loads, snapshot=False):
    """
    Split a boundary list into parallel levels that reproduce the sequential loop.

    store : positions written, in list order;  loads : list of arrays of positions read.
    Returns (order, level_ptr, two_phase): `order` is a stable permutation grouping the
    entries by level; level l is order[level_ptr[l]:level_ptr[l+1]]; two_phase[l] tells
    whether level l must gather before it scatters.

    Sequential semantics for entries i < j:
      read-after-write  (store_i in loads_j)  -> level_j >  level_i   (not for snapshot reads)
      write-after-write (store_i == store_j)  -> level_j >  level_i
      write-after-read  (store_j in loads_i)  -> level_j >= level_i, two-phase if equal
    """
    store = np.asarray(store, dtype=np.int64)
    n = store.size
    loads = [np.asarray(l, dtype=np.int64) for l in loads]
    level = np.zeros(n, dtype=np.int64)
    if n == 0:
        return np.arange(0), np.array([0, 0], dtype=np.int64), np.array([0], dtype=np.int32)

    if snapshot:
        # every read sees the state before the method started, so of several entries storing the same
        # position only the LAST one matters (nobody reads the intermediate values): drop the others.
        # One level; it gathers before it scatters only if some entry reads a position the method stores.
        _, last = np.unique(store[::-1], return_index=True)
        order = np.sort(n - 1 - last)
        st = np.sort(store[order])
        alias = any(_member(st, l[order]).any() for l in loads)
        return order, np.array([0, order.size], dtype=np.int64), np.array([1 if alias else 0], dtype=np.int32)

    sorted_store = np.sort(store)
    aliased = np.zeros(n, dtype=bool)          # entry reads something some entry stores
    for l in loads:
        aliased |= _member(sorted_store, l)
    dup_pos = np.unique(sorted_store[1:][sorted_store[1:] == sorted_store[:-1]])
    duplicated = np.isin(store, dup_pos) if dup_pos.size else np.zeros(n, dtype=bool)

    if aliased.any() or duplicated.any():
        # positions involved in any hazard
        hazard_pos = set(dup_pos.tolist())
        for l in loads:
            hazard_pos.update(l[aliased].tolist())
        involved = np.isin(store, np.fromiter(hazard_pos, dtype=np.int64, count=len(hazard_pos)))
        involved |= duplicated
        involved |= aliased
        idx = np.nonzero(involved)[0]
        writers = collections.defaultdict(list)   # position -> earlier entries that store it
        readers = collections.defaultdict(list)   # position -> earlier entries that read it
        for i in idx:
            lev = 0
            si = int(store[i])
            for j in writers.get(si, ()):                    # write-after-write
                lev = max(lev, level[j] + 1)
            for l in loads:                                  # read-after-write
                for j in writers.get(int(l[i]), ()):
                    lev = max(lev, level[j] + 1)
            for j in readers.get(si, ()):                    # write-after-read
                lev = max(lev, level[j])
            level[i] = lev
            writers[si].append(i)
            for l in loads:
                readers[int(l[i])].append(i)

    order = np.argsort(level, kind="stable")
    nlev = int(level.max()) + 1
    level_ptr = np.zeros(nlev + 1, dtype=np.int64)
    np.cumsum(np.bincount(level, minlength=nlev), out=level_ptr[1:])
    two_phase = np.zeros(nlev, dtype=np.int32)
    if nlev == 1:
        two_phase[0] = int(aliased.any())      # one level: an entry reads a position the level stores
    else:
        for l in range(nlev):
            sel = order[level_ptr[l] : level_ptr[l + 1]]
            st = np.sort(store[sel])
            two_phase[l] = int(any(_member(st, ld[sel]).any() for ld in loads))
    return order, level_ptr, two_phase


class Boundary:
    """
    Builds, for every boundary method class used in the dictionary, the ordered
    list of (population, cell) entries to set (reference: boundary.py:78-160).
    `methods` is the list of BoundaryMethod instances in application order.
    """

    def __init__(self, domain, generator, dico):
        self.domain = domain
        stencil = domain.stencil
        dico_bound = dico.get("boundary_conditions", {}) or {}
        uvel = stencil_uvel(stencil)

        def entries(label, ku):
            # cells whose link along the symmetric of unique velocity ku is cut by `label`,
            # shifted to the outside cell where population ku is stored
            vsym = stencil.unique_velocities[ku].get_symmetric()
            num = int(stencil.unum2index[vsym.num])
            cells, dist = domain.cells_with_flag(num, label)
            indices = np.array(cells)
            if indices.size != 0:
                indices = indices + uvel[num][:, np.newaxis]
            return indices, np.array(dist)

        istore = collections.OrderedDict()
        ilabel, distance = {}, {}
        value_bc, time_bc = {}, {}
        for label in domain.list_of_labels():
            if label in [-1, -2]:
                continue
            if label not in dico_bound:
                raise KeyError("no boundary condition given for the label %s" % label)
            value_bc[label] = dico_bound[label].get("value", None)
            time_bc[label] = dico_bound[label].get("time_bc", False)
            for k, method in dico_bound[label]["method"].items():
                for inumk, numk in enumerate(stencil.num[k]):
                    indices, dist = entries(label, int(stencil.unum2index[numk]))
                    if indices.size == 0:
                        continue
                    ncell = indices.shape[1]
                    velocity = (inumk + stencil.nv_ptr[k]) * np.ones(ncell, dtype=np.int32)[np.newaxis, :]
                    block = np.concatenate([velocity, indices])
                    lab = label * np.ones(ncell, dtype=np.int32)
                    if method not in istore:
                        istore[method], ilabel[method], distance[method] = block, lab, dist
                    else:
                        istore[method] = np.concatenate([istore[method], block], axis=1)
                        ilabel[method] = np.concatenate([ilabel[method], lab])
                        distance[method] = np.concatenate([distance[method], dist])

        self.methods = [
            resolve_method(method)(istore[method], ilabel[method], distance[method], None, stencil, value_bc,
                                   time_bc, tuple([stencil.unvtot] + list(domain.shape_halo)), generator)
            for method in istore
        ]


def resolve_method(cls):
    """boundary-method class of this package for a class given in a dictionary: itself, or -- for the
    classes of an importable reference pylbm (`pylbm.bc.BounceBack`, ...) -- the class of the same name
    (a user-defined subclass resolves through its first known base)."""
    if isinstance(cls, type) and issubclass(cls, BoundaryMethod):
        return cls
    known = {c.__name__: c for c in (BounceBack, BouzidiBounceBack, AntiBounceBack, BouzidiAntiBounceBack,
                                     Neumann, NeumannX, NeumannY, NeumannZ)}
    for base in getattr(cls, "__mro__", ()):
        if base.__name__ in known:
            if base is not cls and any(name in vars(cls) for name in ("set_iload", "set_rhs", "generate", "update")):
                break       # a user subclass that changes the lists or the kernel: not expressible here
            return known[base.__name__]
    raise NotImplementedError("boundary method %r has no CUDA implementation" % (cls,))


class BoundaryMethod:
    """
    Attributes mirrored from the reference: `istore`, `iload` (list), `ilabel`,
    `distance`, `rhs`, `feq`, `value_bc`, `time_bc`, and `s` for Bouzidi types.
    """

    kind = None
    snapshot = False

    def __init__(self, istore, ilabel, distance, normal, stencil, value_bc, time_bc, nspace, generator=None):
        self.istore = istore
        self.feq = np.zeros((int(stencil.nv_ptr[-1]), istore.shape[1]))
        self.rhs = np.zeros(istore.shape[1])
        self.ilabel = ilabel
        self.distance = distance
        self.normal = normal
        self.stencil = stencil
        self.value_bc, self.time_bc = {}, {}
        for k in np.unique(self.ilabel):
            self.value_bc[k] = value_bc[k]
            self.time_bc[k] = time_bc[k]
        self.iload = []
        self.nspace = nspace
        self.generator = generator
        self.func, self.args, self.f, self.m, self.indices = [], [], [], [], []
        self.device_index = None     # index of the method inside the runtime's lbm_sim
        self._order = None

    # ---- lists -------------------------------------------------------------
    def set_iload(self):
        raise NotImplementedError

    def generate(self, sorder):
        """the boundary kernels are part of the static runtime (k_bc<KIND>): nothing to generate
        (reference: boundary.py:323-339 adds one routine per method)."""

    def fix_iload(self):
        """final (ncond, dim+1) int32 C-contiguous layout (reference: boundary.py:226-235)."""
        self.iload = [np.ascontiguousarray(l.T, dtype=np.int32) for l in self.iload]
        self.istore = np.ascontiguousarray(self.istore.T, dtype=np.int32)

    def set_rhs(self):
        raise NotImplementedError

    def _sym_difference(self, sign):
        k = self.istore[:, 0]
        ksym = self.stencil.get_symmetric()[k]
        cols = np.arange(k.size)
        self.rhs[:] = self.feq[k, cols] + sign * self.feq[ksym, cols]

    # ---- wall equilibrium ----------------------------------------------------
    def prepare_rhs(self, simulation):
        """equilibrium populations at the wall points for the prescribed moments
        (reference: boundary.py:238-305).  Runs before fix_iload (istore is (dim+1, n))."""
        nv = simulation.container.nv
        dim = simulation.domain.dim
        v = self.stencil.get_all_velocities()
        for key, value in self.value_bc.items():
            if value is None:
                continue
            indices = np.where(self.ilabel == key)
            ncond = indices[0].size
            k = self.istore[0, indices]
            s = 1 - self.distance[indices]
            coords = tuple()
            for i in range(dim):
                x = simulation.domain.coords_halo[i][self.istore[i + 1, indices]]
                x += s * v[k, i] * simulation.domain.dx
                x = x.ravel()
                for _ in range(1, dim):
                    x = x[:, np.newaxis]
                coords += (x,)
            nspace = [ncond] + [1] * (dim - 1)
            m = HostArray(nv, nspace, consm=simulation.scheme.consm)
            f = HostArray(nv, nspace, consm=simulation.scheme.consm)
            args, func = coords, value
            if isinstance(value, tuple):
                func = value[0]
                args = args + tuple(value[1])
            if self.time_bc[key]:
                func(f, m, 0, *args)
            else:
                func(f, m, *args)
            simulation.equilibrium(m)
            simulation.m2f(m, f)
            self.feq[:, indices[0]] = f.array.reshape((nv, ncond))
            if self.time_bc[key]:
                self.func.append(func)
                self.args.append(args)
                self.f.append(f)
                self.m.append(m)
                self.indices.append(indices[0])

    def update_feq(self, simulation):
        """time-dependent boundary values (reference: boundary.py:307-321)."""
        nv = simulation.container.nv
        for i, func in enumerate(self.func):
            func(self.f[i], self.m[i], simulation.t, *self.args[i])
            simulation.equilibrium(self.m[i])
            simulation.m2f(self.m[i], self.f[i])
            self.feq[:, self.indices[i]] = self.f[i].array.reshape((nv, self.indices[i].size))

    @property
    def is_time_dependent(self):
        return len(self.func) > 0

    rhs_sign = None          # rhs = feq[k] + rhs_sign * feq[ksym] (None: the kind has no right-hand side)

    def prepare_time_bc(self, simulation):
        """device side of the time-dependent labels (after move2gpu): per label, scratch arrays for the
        moments / equilibrium populations at the wall points and the index lists of
        `rhs = feq[k] -/+ feq[ksym]` in the order of the device list."""
        from .storage import DeviceArray

        self._time_plans = []
        if not self.func or self.rhs_sign is None or self.device_index is None:
            return
        nv = simulation.container.nv
        sym = np.asarray(self.stencil.get_symmetric())
        inverse = np.empty(self.istore.shape[0], dtype=np.int64)
        inverse[:] = -1
        inverse[self._order] = np.arange(self._order.size)
        for idx in self.indices:
            ncond = idx.size
            dm = DeviceArray(nv, (ncond,), [0], "f64")
            df = DeviceArray(nv, (ncond,), [0], "f64")
            k = self.istore[idx, 0].astype(np.int64)
            col = np.arange(ncond, dtype=np.int64)
            dst = inverse[idx]
            live = dst >= 0                     # entries dropped from the device list (duplicates)
            a = df.positions(np.stack([k, col]))[live]
            b = df.positions(np.stack([sym[k], col]))[live]
            lists = []
            for arr in (a, b, dst[live]):
                arr = np.ascontiguousarray(arr, dtype=np.int64)
                ptr = ctypes.c_void_p()
                rt.check(rt.lib().lbm_malloc(ctypes.byref(ptr), max(8, arr.nbytes)), "lbm_malloc")
                rt.check(rt.lib().lbm_memcpy_h2d(ptr, arr.ctypes.data, arr.nbytes), "h2d")
                lists.append(ptr.value)
            self._time_plans.append((dm, df, lists, int(live.sum())))

    def update_feq_device(self, simulation):
        """time-dependent boundary values (reference: boundary.py:307-321 + set_rhs 421-427) with only
        the user's callback on the host: its moments are uploaded asynchronously, equilibrium + m2f run
        on the simulation's stream and the right-hand sides are recomputed in the device list -- no
        device-to-host copy, no synchronisation."""
        lib = rt.lib()
        handle = simulation._handle
        kernels = simulation._kernels_f64()
        stream = lib.lbm_sim_stream(handle)
        for i, func in enumerate(self.func):
            dm, df, lists, count = self._time_plans[i]
            m = self.m[i]
            func(self.f[i], m, simulation.t, *self.args[i])
            host = m.array.reshape(dm.nv, -1)
            rt.check(lib.lbm_sim_upload_rows(handle, dm.ptr + dm.lead * 8, dm.pstride * 8, host.ctypes.data,
                                             host.strides[0], host.shape[1] * 8, dm.nv), "lbm_sim_upload_rows")
            kernels.launch("equilibrium", dm.ptr, dm.ptr, dm.grid, simulation._scalar_values("equilibrium"), stream)
            kernels.launch("m2f", dm.ptr, df.ptr, dm.grid, simulation._scalar_values("m2f"), stream)
            rt.check(lib.lbm_sim_rhs_update(handle, self.device_index, count, lists[0], lists[1], lists[2],
                                            float(self.rhs_sign), df.ptr), "lbm_sim_rhs_update")

    def __del__(self):
        for plan in getattr(self, "_time_plans", ()):
            for ptr in plan[2]:
                try:
                    rt.lib().lbm_free(ptr)
                except Exception:
                    pass

    # ---- device side ---------------------------------------------------------
    def device_lists(self, array):
        """positions in the padded device layout + level schedule (after fix_iload)."""
        store = array.positions(self.istore.T)
        loads = [array.positions(l.T) for l in self.iload]
        order, level_ptr, two_phase = schedule(store, loads, snapshot=self.snapshot)
        self._order = order
        return store[order], [l[order] for l in loads], level_ptr, two_phase

    def prepare_device(self, array):
        """device-order lists (positions in the padded layout, level schedule) kept in `_keep`."""
        store, loads, level_ptr, two_phase = self.device_lists(array)
        store = np.ascontiguousarray(store)
        l0 = np.ascontiguousarray(loads[0])
        l1 = np.ascontiguousarray(loads[1]) if len(loads) > 1 else None
        rhs = np.ascontiguousarray(self.rhs[self._order])
        dist = np.ascontiguousarray(self.s[self._order]) if hasattr(self, "s") else None
        self._keep = (store, l0, l1, rhs, dist, level_ptr, two_phase)
        self._wall_mask = None
        return self._keep

    def move2gpu(self, sim_handle, array, wall_mask=None):
        """register the method in the runtime (reference: boundary.py:378-397 move2gpu).  Entries
        selected by `wall_mask` are applied by the fused kernel (plan_walls): they are left out of the
        method's list and registered by the caller as a separate stale-only method."""
        if getattr(self, "_keep", None) is None:
            self.prepare_device(array)
        store, l0, l1, rhs, dist, level_ptr, two_phase = self._keep
        if wall_mask is not None and wall_mask.any():
            keep = ~wall_mask                      # only single-level, gather-free methods get here
            self._wall_mask = wall_mask
            self._wall_lists = (np.ascontiguousarray(store[wall_mask]), np.ascontiguousarray(l0[wall_mask]),
                                np.ascontiguousarray(rhs[wall_mask]))
            self._order = self._order[keep]
            store, l0, rhs = (np.ascontiguousarray(store[keep]), np.ascontiguousarray(l0[keep]),
                              np.ascontiguousarray(rhs[keep]))
            level_ptr = np.array([0, store.size], dtype=np.int64)
            self._keep = (store, l0, l1, rhs, dist, level_ptr, two_phase)
        idx = rt.lib().lbm_sim_add_bc(
            sim_handle, self.kind, store.size, store.ctypes.data, l0.ctypes.data,
            l1.ctypes.data if l1 is not None else None, rhs.ctypes.data,
            dist.ctypes.data if dist is not None else None,
            two_phase.size, level_ptr.ctypes.data, two_phase.ctypes.data,
        )
        rt.check(idx, "lbm_sim_add_bc(%s)" % type(self).__name__)
        self.device_index = idx

    def push_rhs(self, sim_handle):
        rhs = np.ascontiguousarray(self.rhs[self._order])
        rt.check(rt.lib().lbm_sim_set_rhs(sim_handle, self.device_index, rhs.ctypes.data), "lbm_sim_set_rhs")


class BounceBack(BoundaryMethod):
    """f_k(store) = f_ksym(store + v_k) + rhs (reference: boundary.py:400-469)."""

    kind = rt.BC_BOUNCE_BACK

    def set_iload(self):
        k = self.istore[0]
        ksym = self.stencil.get_symmetric()[k][np.newaxis, :]
        v = self.stencil.get_all_velocities()
        self.iload.append(np.concatenate([ksym, self.istore[1:] + v[k].T]))

    rhs_sign = -1

    def set_rhs(self):
        self._sym_difference(-1)


class BouzidiBounceBack(BoundaryMethod):
    """Bouzidi-Firdaouss-Lallemand interpolated bounce-back (reference: boundary.py:472-623)."""

    kind = rt.BC_BOUZIDI_BOUNCE_BACK
    snapshot = True

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.s = np.empty(self.istore.shape[1])

    def set_iload(self):
        k = self.istore[0]
        ksym = self.stencil.get_symmetric()[k]
        v = self.stencil.get_all_velocities()
        iload1 = np.zeros(self.istore.shape, dtype=np.int32)
        iload2 = np.zeros(self.istore.shape, dtype=np.int32)

        near = self.distance < 0.5
        iload1[0, near] = ksym[near]
        iload2[0, near] = ksym[near]
        iload1[1:, near] = self.istore[1:, near] + v[k[near]].T
        iload2[1:, near] = self.istore[1:, near] + 2 * v[k[near]].T
        self.s[near] = 2.0 * self.distance[near]

        far = np.logical_not(near)
        iload1[0, far] = ksym[far]
        iload2[0, far] = k[far]
        iload1[1:, far] = self.istore[1:, far] + v[k[far]].T
        iload2[1:, far] = self.istore[1:, far] + v[k[far]].T
        self.s[far] = 0.5 / self.distance[far]

        self.iload.append(iload1)
        self.iload.append(iload2)

    rhs_sign = -1

    def set_rhs(self):
        self._sym_difference(-1)


class AntiBounceBack(BounceBack):
    """f_k(store) = -f_ksym(store + v_k) + rhs (reference: boundary.py:626-684)."""

    kind = rt.BC_ANTI_BOUNCE_BACK
    rhs_sign = +1

    def set_rhs(self):
        self._sym_difference(+1)


class BouzidiAntiBounceBack(BouzidiBounceBack):
    """interpolated anti bounce-back, in place (reference: boundary.py:687-760)."""

    kind = rt.BC_BOUZIDI_ANTI_BOUNCE_BACK
    snapshot = False
    rhs_sign = +1

    def set_rhs(self):
        self._sym_difference(+1)


class Neumann(BoundaryMethod):
    """f_k(store) = f_k(store + v_k) (reference: boundary.py:763-823)."""

    kind = rt.BC_NEUMANN
    name = "neumann"
    axis = None

    def set_rhs(self):
        pass

    def set_iload(self):
        k = self.istore[0]
        v = self.stencil.get_all_velocities()
        if self.axis is None:
            indices = self.istore[1:] + v[k].T
        else:
            indices = self.istore[1:].copy()
            indices[self.axis] += v[k].T[self.axis]
        self.iload.append(np.concatenate([k[np.newaxis, :], indices]))


class NeumannX(Neumann):
    name = "neumannx"
    axis = 0


class NeumannY(Neumann):
    name = "neumanny"
    axis = 1


class NeumannZ(Neumann):
    name = "neumannz"
    axis = 2
