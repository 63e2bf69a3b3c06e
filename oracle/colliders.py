"""numpy restatement of the naive, cell-list and Verlet neighbour-list colliders.  Oracle only.

jaxdem/colliders/__init__.py:225-243 (mask), naive.py:73-235,
_partition.py:23-150, cell_list.py:35-96,99-174,187-261,374-595.
"""

from __future__ import annotations

import numpy as np

from .state import int_dtype_for

from . import linalg as la
from .forces import LAWS

PAIR_UNROLL = 4  # cell_list.py:184


def valid_interaction_mask(clump_i, clump_j, bond_id_i, index_j, interact_same_bond_id=False):
    """jaxdem/colliders/__init__.py:225-243 (argument names as in the reference;
    the cell list passes the CANDIDATE's clump/bond row first and the OWNER's
    index last, cell_list.py:240-246)."""
    is_bonded = np.any(bond_id_i == np.asarray(index_j)[..., None], axis=-1)
    return ((clump_i != clump_j) & (~is_bonded | bool(interact_same_bond_id))).astype(np.int64)


# --------------------------------------------------------------------------
# naive O(N^2)
# --------------------------------------------------------------------------


def naive_compute_force(state, system):
    """jaxdem/colliders/naive.py:187-235."""
    force_fn = LAWS[system.force_model][0]
    N = state.N
    pos = state.pos
    iota = np.arange(N)
    F = np.zeros_like(state.force)
    T = np.zeros_like(state.torque)
    for i in range(N):
        ii = np.full(N, i)
        f, t = force_fn(ii, iota, pos, state, system)
        mask = valid_interaction_mask(
            state.clump_id[i], state.clump_id, state.bond_id[i], iota, system.interact_same_bond_id
        )[:, None]
        f = np.where(mask > 0, f, 0.0).astype(F.dtype)
        t = np.where(mask > 0, t, 0.0).astype(T.dtype)
        F[i] = f.sum(axis=0)
        T[i] = t.sum(axis=0) + la.cross(state._pos_p_rot[i], F[i])
    state.force, state.torque = F, T


def naive_compute_potential_energy(state, system):
    """jaxdem/colliders/naive.py:73-113."""
    energy_fn = LAWS[system.force_model][1]
    N = state.N
    pos = state.pos
    iota = np.arange(N)
    tot = state.fdtype.type(0.0)
    for i in range(N):
        e = energy_fn(np.full(N, i), iota, pos, state, system)
        mask = valid_interaction_mask(
            state.clump_id[i], state.clump_id, state.bond_id[i], iota, system.interact_same_bond_id
        )
        tot = tot + state.fdtype.type(0.5) * np.where(mask > 0, e, 0.0).astype(e.dtype).sum()
    return tot


# --------------------------------------------------------------------------
# cell list
# --------------------------------------------------------------------------


class OCellList:
    """Fields of DynamicCellList (cell_list.py:368-372) + Collider.overflow."""

    kind = "celllist"

    def __init__(self, neighbor_mask, cell_size):
        self.neighbor_mask = neighbor_mask
        self.cell_size = cell_size
        self.overflow = False


class OMultiCellList(OCellList):
    """DynamicMultiCellList (multi_cell_list.py:283-328): the cell list's fields; traversal prunes by cell AABB."""

    kind = "multicelllist"


def multicelllist_create(state, cell_size=None, search_range=None, box_size=None, max_hashes=None) -> OMultiCellList:
    """DynamicMultiCellList.Create (multi_cell_list.py:330-392): like the cell list's, with the default cell
    2 r_max whatever the polydispersity."""
    if cell_size is None:
        cell_size = state.fdtype.type(2.0) * np.max(state._rad)
    base = celllist_create(state, cell_size=cell_size, search_range=search_range, box_size=box_size)
    return OMultiCellList(base.neighbor_mask, base.cell_size)


def loose_cell_aabbs(member_min, member_max, sorted_hash):
    """_loose_cell_aabbs (multi_cell_list.py:46-73): per-cell union of the member boxes, a segmented min / max over
    the sorted runs, broadcast back to every member: (centre, half extent) indexed by sorted slot."""
    n = member_min.shape[0]
    if n == 0:
        return member_min, member_min
    seg_start = np.concatenate([[True], sorted_hash[1:] != sorted_hash[:-1]])
    seg_id = np.cumsum(seg_start) - 1
    starts = np.nonzero(seg_start)[0]
    cmin = np.minimum.reduceat(member_min, starts, axis=0)[seg_id]
    cmax = np.maximum.reduceat(member_max, starts, axis=0)[seg_id]
    half = member_min.dtype.type(0.5)
    return half * (cmin + cmax), half * (cmax - cmin)


def float_to_int(x: np.ndarray, idtype) -> np.ndarray:
    """float -> int conversion with the saturating semantics XLA and CUDA share
    (NaN -> 0); numpy's own astype is undefined out of range."""
    info = np.iinfo(idtype)
    y = np.where(np.isnan(x), 0.0, x)
    y = np.clip(y, float(info.min), float(info.max))
    out = y.astype(np.float64).astype(idtype) if y.dtype != np.float64 else y.astype(idtype)
    # float(info.max) rounds up to 2**31 / 2**63: clamp after conversion too
    out = np.where(y >= float(info.max), info.max, out).astype(idtype)
    return out


def celllist_create(state, cell_size=None, search_range=None, box_size=None) -> OCellList:
    """DynamicCellList.Create (cell_list.py:374-432)."""
    F, I = state.fdtype, state.idtype
    min_rad, max_rad = np.min(state._rad), np.max(state._rad)
    alpha = max_rad / min_rad
    if cell_size is None:
        cell_size = np.where(alpha < 2.5, F.type(2.0) * max_rad, F.type(0.5) * max_rad).astype(F)
    cell_size = np.asarray(cell_size, dtype=F)
    if box_size is not None:
        box_size = np.asarray(box_size, dtype=F)
        for _ in range(2):
            if search_range is None:
                sr = max(1, int(np.ceil(F.type(2) * max_rad / cell_size)))
            else:
                sr = int(search_range)
            min_grids = 2 * sr + 1
            gd = np.maximum(np.floor(box_size / cell_size).astype(I), min_grids)
            cell_size = np.min(box_size / gd.astype(F)).astype(F)
    if search_range is None:
        search_range = max(1, int(np.ceil(F.type(2) * max_rad / cell_size)))
    search_range = int(search_range)
    r = np.arange(-search_range, search_range + 1, dtype=I)
    mesh = np.meshgrid(*([r] * state.dim), indexing="ij")
    mask = np.stack([m.ravel() for m in mesh], axis=1).astype(I)  # last axis fastest
    return OCellList(mask, np.asarray(cell_size, dtype=F))


def grid_params(box_size, cell_size, periodic, idtype):
    """_grid_params (_partition.py:54-99)."""
    F = box_size.dtype
    if periodic:
        gd = float_to_int(np.floor(box_size / cell_size), idtype)
        gd = np.maximum(gd, idtype.type(1))
        cell_size = box_size / gd.astype(F)
    else:
        gd = float_to_int(np.ceil(box_size / cell_size), idtype)
        gd = np.maximum(gd, idtype.type(1))
    with np.errstate(over="ignore"):
        strides = np.concatenate([np.array([1], dtype=idtype), np.cumprod(gd[:-1], dtype=idtype)])
    total = np.prod(gd.astype(F))
    overflow = bool(total > F.type(float(np.iinfo(idtype).max)))
    return gd, strides, cell_size, overflow


def get_spatial_partition(pos, system, cell_size, neighbor_mask, idtype, order_id=None):
    """_get_spatial_partition (cell_list.py:35-87).

    Returns (perm, sorted_hash, neighbor_cell_hashes (N, M), hash_overflow).  ``order_id`` (not in the
    reference): the iota of the stable sort, for a State whose rows are a subset / permutation of a larger
    system's particles (slab decomposition: the global particle ids)."""
    dom = system.domain
    F = pos.dtype
    gd, strides, cs_eff, overflow = grid_params(dom.box_size, cell_size, dom.periodic, idtype)
    with np.errstate(over="ignore", invalid="ignore"):
        if dom.periodic:
            u = (pos - dom.anchor) / dom.box_size
            # jnp.remainder(u, 1): fmod then sign fix-up
            r = np.fmod(u, F.type(1.0))
            r = np.where((r != 0) & (r < 0), r + F.type(1.0), r)
            coords = float_to_int(np.floor(r * gd.astype(F)), idtype)
        else:
            # scalar cell_size as passed: _grid_params rewrites it only when periodic
            coords = float_to_int(np.floor((pos - dom.anchor) / cell_size), idtype)
        hashes = (coords * strides).sum(axis=-1, dtype=idtype)  # wrapping integer dot
        N = pos.shape[0]
        if order_id is None:
            perm = np.argsort(hashes, kind="stable").astype(idtype)  # lax.sort(num_keys=1) is stable
        else:
            perm = np.lexsort((np.asarray(order_id), hashes)).astype(idtype)  # (hash, global iota)
        sorted_hash = hashes[perm]

        ncoords = coords[:, None, :] + neighbor_mask[None, :, :]
        if dom.periodic:
            ncoords = ncoords - gd * float_to_int(np.floor(ncoords.astype(F) / gd.astype(F)), idtype)
            nh = (ncoords * strides).sum(axis=-1, dtype=idtype)
        else:
            oob = np.any((ncoords < 0) | (ncoords >= gd), axis=-1)
            nh = (ncoords * strides).sum(axis=-1, dtype=idtype)
            nh = np.where(oob, idtype.type(-1), nh)
    return perm, sorted_hash, nh, overflow, hashes


def dedup_stencil_hashes(nh: np.ndarray) -> np.ndarray:
    """_dedup_stencil_hashes (cell_list.py:90-96), vectorised over particles:
    later duplicates in a row become -1."""
    M = nh.shape[1]
    upper = np.triu(np.ones((M, M), bool), k=1)  # a < b
    out = np.empty_like(nh)
    for s in range(0, nh.shape[0], 8192):  # chunked: the (n, M, M) compare is large
        blk = nh[s : s + 8192]
        eq = blk[:, :, None] == blk[:, None, :]  # (n, a, b)
        is_dup = np.any(eq & upper[None], axis=1)  # column b: any earlier a equal
        out[s : s + 8192] = np.where(is_dup, nh.dtype.type(-1), blk)
    return out


def _partition_for(state, system, cell_size):
    col = system.collider
    pos = state.pos
    perm, sh, nh, ovf, hashes = get_spatial_partition(pos, system, cell_size, col.neighbor_mask, state.idtype,
                                                      order_id=getattr(state, "order_id", None))
    if system.domain.periodic:
        nh = dedup_stencil_hashes(nh)
    return pos, perm, sh, nh, ovf


def _traverse(state, system, visit):
    """_traverse_pairs (cell_list.py:187-261): for every particle i, every
    stencil row m, every k in the sorted run of that cell (from lower_bound)
    call ``visit(i_idx, j_idx, m)``.  Vectorised over particles; ordered by
    (m, offset in run) so accumulation order is deterministic."""
    col = system.collider
    pos, perm, sh, nh, ovf = _partition_for(state, system, col.cell_size)
    N = state.N
    iota = np.arange(N)
    prune = getattr(col, "kind", "") == "multicelllist"
    if prune:  # multi_cell_list.py:170-180: expandable AABB per loose cell, query box = the particle's own box
        rad = state._rad[:, None]
        xmin, xmax = pos - rad, pos + rad
        cell_center, cell_half = loose_cell_aabbs(xmin[perm], xmax[perm], sh)
        half = state.fdtype.type(0.5)
        qc, qh = half * (xmin + xmax), half * (xmax - xmin)
    for m in range(nh.shape[1]):
        target = nh[:, m]
        start = np.searchsorted(sh, target, side="left")
        end = np.searchsorted(sh, target, side="right")
        cnt = end - start
        if prune:  # :186-204: a cell whose box does not reach the query box is skipped wholesale
            ss = np.minimum(start, max(N, 1) - 1)
            dr = system.domain.displacement(qc, cell_center[ss])
            cnt = np.where(np.all(np.abs(dr) <= qh + cell_half[ss], axis=-1), cnt, 0)
        for t in range(int(cnt.max()) if N else 0):
            sel = np.nonzero(cnt > t)[0]
            kj = start[sel] + t
            visit(iota[sel], perm[kj].astype(np.int64), pos)
    return ovf


def celllist_compute_force(state, system):
    """DynamicCellList.compute_force (cell_list.py:434-464) with _force_pair_fn
    (_partition.py:23-36)."""
    force_fn = LAWS[system.force_model][0]
    F = np.zeros_like(state.force)
    T = np.zeros_like(state.torque)

    def visit(i, j, pos):
        valid = valid_interaction_mask(
            state.clump_id[j], state.clump_id[i], state.bond_id[j], i, system.interact_same_bond_id
        )
        f, t = force_fn(i, j, pos, state, system)
        f = np.where((valid > 0)[:, None], f, 0.0).astype(F.dtype)
        t = np.where((valid > 0)[:, None], t, 0.0).astype(T.dtype)
        F[i] += f  # i is unique within one visit
        T[i] += t

    ovf = _traverse(state, system, visit)
    state.force = F
    state.torque = T + la.cross(state._pos_p_rot, F)
    system.collider.overflow = ovf


def celllist_compute_potential_energy(state, system):
    """DynamicCellList.compute_potential_energy (cell_list.py:466-496)."""
    energy_fn = LAWS[system.force_model][1]
    E = np.zeros(state.N, state.fdtype)

    def visit(i, j, pos):
        valid = valid_interaction_mask(
            state.clump_id[j], state.clump_id[i], state.bond_id[j], i, system.interact_same_bond_id
        )
        e = energy_fn(i, j, pos, state, system)
        E[i] += state.fdtype.type(0.5) * np.where(valid > 0, e, 0.0).astype(E.dtype)

    ovf = _traverse(state, system, visit)
    system.collider.overflow = ovf
    return E.sum()


def celllist_create_neighbor_list(state, system, cutoff, max_neighbors):
    """create_neighbor_list (cell_list.py:498-595) incl. the PAIR_UNROLL chunking
    of _make_stencil_body (:99-174) and _pack_stencil_lists (_partition.py:102-150).

    Returns (neighbor_list (N, K) int, overflow bool)."""
    col = system.collider
    I = state.idtype
    N = state.N
    F = state.fdtype
    if max_neighbors == 0:
        return np.empty((N, 0), I), False
    cutoff = F.type(cutoff)
    cutoff_sq = cutoff**2
    search_range = max(int(np.max(col.neighbor_mask)), 1)
    cell_size = np.maximum(col.cell_size, cutoff / F.type(search_range))
    pos, perm, sh, nh, hash_ovf = _partition_for(state, system, cell_size)
    M = nh.shape[1]
    cap = max_neighbors
    dom = system.domain

    nl = np.full((N, max_neighbors), -1, dtype=I)
    row_off = np.zeros(N, np.int64)  # running sum of raw per-cell counts
    any_stencil_ovf = False
    total = np.zeros(N, np.int64)
    iota = np.arange(N)
    prune = getattr(col, "kind", "") == "multicelllist"
    if prune:  # multi_cell_list.py:466-468: AABBs of the bare positions
        cell_center, cell_half = loose_cell_aabbs(pos[perm], pos[perm], sh)
    for m in range(M):
        target = nh[:, m]
        k = np.searchsorted(sh, target, side="left").astype(np.int64)
        if prune:  # :475-481: non-overlapping cells are looked up as hash -1, which no particle carries
            ss = np.minimum(k, max(1, N) - 1)
            dr = dom.displacement(pos, cell_center[ss])
            target = np.where(np.all(np.abs(dr) <= cutoff + cell_half[ss], axis=-1), target, I.type(-1))
        c = np.zeros(N, np.int64)
        while True:
            safe_k = np.minimum(k, max(1, N) - 1)
            alive = (k < N) & (sh[safe_k] == target) & (c < cap + 1)  # cond_fun :137-142
            if not alive.any():
                break
            sel = np.nonzero(alive)[0]
            for u in range(PAIR_UNROLL):  # body_fun :144-165, lanes in order
                ak = k[sel] + u
                sk = np.minimum(ak, max(1, N) - 1)
                in_cell = (ak < N) & (sh[sk] == target[sel])
                j = perm[sk].astype(np.int64)
                dr = dom.displacement(pos[sel], pos[j])  # division form
                d2 = la.norm2(dr)
                valid = (
                    valid_interaction_mask(
                        state.clump_id[j], state.clump_id[sel], state.bond_id[j], iota[sel],
                        system.interact_same_bond_id,
                    )
                    > 0
                ) & (d2 <= cutoff_sq) & in_cell
                w = sel[valid]
                slot = c[w]  # position inside this cell's local buffer
                keep = slot < cap  # mode="drop" beyond local capacity
                dest = row_off[w] + slot  # _pack_stencil_lists target index
                ok = keep & (dest < max_neighbors)
                nl[w[ok], dest[ok]] = j[valid][ok].astype(I)
                c[w] += 1
            any_stencil_ovf |= bool(np.any(c[sel] > cap))  # overflow flag :163
            k[sel] += PAIR_UNROLL
        row_off += c
        total += c
    count_ovf = bool(np.any(total > max_neighbors))
    return nl, bool(any_stencil_ovf or count_ovf or hash_ovf)


def celllist_create_cross_neighbor_list(pos_a, pos_b, system, cutoff, max_neighbors, idtype=None):
    """create_cross_neighbor_list (cell_list.py:600-715): for every query point of pos_a the
    points of pos_b within ``cutoff``; pos_b is partitioned into cells, the stencil comes from
    the query point's own cell; no clump / bond mask; rows in stencil x sorted-run order, mapped
    back to original B indices, padded with -1.  Returns (list (N_A, K) int, overflow bool)."""
    col = system.collider
    pos_a, pos_b = np.asarray(pos_a), np.asarray(pos_b)
    F = pos_b.dtype
    I = np.dtype(idtype) if idtype is not None else int_dtype_for(F)
    n_a, n_b = pos_a.shape[0], pos_b.shape[0]
    if n_a == 0:
        return np.empty((0, max_neighbors), I), False
    if n_b == 0:
        return np.full((n_a, max_neighbors), -1, I), False
    if max_neighbors == 0:
        return np.empty((n_a, 0), I), False
    cutoff = F.type(cutoff)
    cutoff_sq = cutoff**2
    search_range = max(int(np.max(col.neighbor_mask)), 1)
    cell_size = np.maximum(col.cell_size, cutoff / F.type(search_range))
    perm_b, sh, _, ovf_b, _ = get_spatial_partition(pos_b, system, cell_size, col.neighbor_mask, I)
    _, _, nh, ovf_a, _ = get_spatial_partition(pos_a, system, cell_size, col.neighbor_mask, I)
    if system.domain.periodic:
        nh = dedup_stencil_hashes(nh)
    dom = system.domain
    cap = max_neighbors
    nl = np.full((n_a, max_neighbors), -1, dtype=I)
    row_off = np.zeros(n_a, np.int64)
    total = np.zeros(n_a, np.int64)
    any_stencil_ovf = False
    for m in range(nh.shape[1]):
        target = nh[:, m]
        k = np.searchsorted(sh, target, side="left").astype(np.int64)
        c = np.zeros(n_a, np.int64)
        while True:
            safe_k = np.minimum(k, n_b - 1)
            alive = (k < n_b) & (sh[safe_k] == target) & (c < cap + 1)  # cond_fun of _make_stencil_body
            if not alive.any():
                break
            sel = np.nonzero(alive)[0]
            for u in range(PAIR_UNROLL):
                ak = k[sel] + u
                sk = np.minimum(ak, n_b - 1)
                in_cell = (ak < n_b) & (sh[sk] == target[sel])
                j = perm_b[sk].astype(np.int64)
                dr = dom.displacement(pos_a[sel], pos_b[j])  # division form
                valid = (la.norm2(dr) <= cutoff_sq) & in_cell
                w = sel[valid]
                slot = c[w]
                dest = row_off[w] + slot
                ok = (slot < cap) & (dest < max_neighbors)
                nl[w[ok], dest[ok]] = j[valid][ok].astype(I)
                c[w] += 1
            any_stencil_ovf |= bool(np.any(c[sel] > cap))
            k[sel] += PAIR_UNROLL
        row_off += c
        total += c
    count_ovf = bool(np.any(total > max_neighbors))
    return nl, bool(any_stencil_ovf or count_ovf or ovf_a or ovf_b)

# --------------------------------------------------------------------------
# Verlet NeighborList collider (jaxdem/colliders/neighbor_list.py:57-131, 286-480, 542-727)
# --------------------------------------------------------------------------


class ONeighborList:
    """Fields of NeighborList (neighbor_list.py:262-284)."""

    kind = "neighborlist"

    def __init__(self, secondary, neighbor_list, old_pos, cutoff, skin, max_neighbors):
        self.secondary_collider = secondary
        self.neighbor_list = neighbor_list
        self.old_pos = old_pos
        self.n_build_times = 0
        self.cutoff = cutoff
        self.skin = skin
        self.max_neighbors = int(max_neighbors)
        self.overflow = False


def neighborlist_max_neighbors(state, cutoff, skin_val, max_neighbors=None, number_density=1.0, safety_factor=1.2):
    """The buffer-size heuristics of NeighborList.Create (neighbor_list.py:346-392)."""
    list_cutoff = cutoff + skin_val
    dim, N = state.dim, state.N
    max_rad = float(np.max(state._rad))
    pos = state.pos
    box = np.maximum(pos.max(axis=0) - pos.min(axis=0) + 2.0 * max_rad, 1.0)
    density_est = float(N / np.prod(box.astype(np.float64)))
    eff_density = max(number_density, density_est)
    r_eff = 0.9 * float(np.min(state._rad))
    upper = ((list_cutoff + r_eff) / r_eff) ** dim
    max_possible = int(np.ceil((0.91 if dim == 2 else 0.74) * upper))
    r_eff_mean = 0.9 * float(np.mean(state._rad))
    typical = int(np.ceil(((list_cutoff + r_eff_mean) / r_eff_mean) ** dim))
    if max_neighbors is None:
        vol = np.pi * list_cutoff**dim * (1.0 if dim == 2 else 4.0 / 3.0)
        max_neighbors = max(int(np.ceil(safety_factor * vol * eff_density)), typical)
    return max(min(max_neighbors, max_possible, N), 0)


def neighborlist_create(state, cutoff, skin=None, skin_fraction=None, max_neighbors=None, number_density=1.0,
                        safety_factor=1.2, secondary_collider_kw=None) -> ONeighborList:
    """NeighborList.Create (neighbor_list.py:286-402), secondary collider = CellList with
    cell_size = cutoff + skin unless given."""
    if skin is not None and skin_fraction is not None:
        raise ValueError("Pass either `skin` or `skin_fraction`, not both.")
    if skin is None:
        skin_val = float(0.05 if skin_fraction is None else skin_fraction) * cutoff
    else:
        skin_val = float(skin)
    K = neighborlist_max_neighbors(state, cutoff, skin_val, max_neighbors, number_density, safety_factor)
    kw = dict(secondary_collider_kw or {})
    kw.pop("state", None)
    kw.setdefault("cell_size", cutoff + skin_val)
    sec = celllist_create(state, **kw)
    F = state.fdtype
    return ONeighborList(sec, np.full((state.N, K), -1, state.idtype), state.pos.copy(), F.type(cutoff),
                         F.type(skin_val), K)


def neighborlist_check_and_rebuild(state, system):
    """_check_and_rebuild (neighbor_list.py:57-131), no history."""
    col = system.collider
    disp = state.pos - col.old_pos  # deliberately not a periodic displacement (:81-83)
    max_disp_sq = np.max(la.norm2(disp)) if state.N else state.fdtype.type(0)
    trigger = col.skin**2 / state.fdtype.type(4)
    if (max_disp_sq > trigger) or col.n_build_times == 0:
        list_cutoff = col.cutoff + col.skin  # _rebuild :463
        system.collider = col.secondary_collider
        try:
            nl, ovf = celllist_create_neighbor_list(state, system, list_cutoff, col.max_neighbors)
        finally:
            system.collider = col
        col.neighbor_list = nl
        col.old_pos = state.pos.copy()
        col.n_build_times += 1
        col.overflow = ovf
    return col.neighbor_list


def _nl_pairs(state, system):
    """(i, j) of every list entry that passes the -1 test and valid_interaction_mask as NeighborList calls it
    (neighbor_list.py:588-598: clump of i, clump of j, bond row of i, index j), in list order."""
    nl = neighborlist_check_and_rebuild(state, system)
    N, K = nl.shape
    ii = np.repeat(np.arange(N), K)
    jj = nl.reshape(-1).astype(np.int64)
    keep = jj != -1
    ii, jj = ii[keep], jj[keep]
    valid = valid_interaction_mask(state.clump_id[ii], state.clump_id[jj], state.bond_id[ii], jj,
                                   system.interact_same_bond_id) > 0
    return ii[valid], jj[valid]


def neighborlist_compute_force(state, system):
    """NeighborList.compute_force (neighbor_list.py:542-632)."""
    force_fn = LAWS[system.force_model][0]
    ii, jj = _nl_pairs(state, system)
    F = np.zeros_like(state.force)
    T = np.zeros_like(state.torque)
    if len(ii):
        f, t = force_fn(ii, jj, state.pos, state, system)
        np.add.at(F, ii, f.astype(F.dtype))  # list order within a row
        np.add.at(T, ii, t.astype(T.dtype))
    state.force = F
    state.torque = T + la.cross(state._pos_p_rot, F)


def neighborlist_compute_potential_energy(state, system):
    """NeighborList.compute_potential_energy (neighbor_list.py:634-727)."""
    energy_fn = LAWS[system.force_model][1]
    ii, jj = _nl_pairs(state, system)
    E = np.zeros(state.N, state.fdtype)
    if len(ii):
        np.add.at(E, ii, energy_fn(ii, jj, state.pos, state, system).astype(E.dtype))
    return (state.fdtype.type(0.5) * E).sum()


def compute_force(state, system):
    if system.collider_type == "naive":
        naive_compute_force(state, system)
    elif system.collider_type in ("celllist", "multicelllist"):
        celllist_compute_force(state, system)  # _traverse prunes when the collider is the loose grid
    elif system.collider_type == "neighborlist":
        neighborlist_compute_force(state, system)
    else:  # "" no-op collider (colliders/__init__.py:56-88)
        state.force = state.force * 0
        state.torque = state.torque * 0


def compute_potential_energy(state, system):
    if system.collider_type == "naive":
        return naive_compute_potential_energy(state, system)
    if system.collider_type in ("celllist", "multicelllist"):
        return celllist_compute_potential_energy(state, system)
    if system.collider_type == "neighborlist":
        return neighborlist_compute_potential_energy(state, system)
    return state.fdtype.type(0.0)
