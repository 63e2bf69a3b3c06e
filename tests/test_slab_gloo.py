"""Slab decomposition (jaxdem_b200/slab.py) on CPU: world_size-2 and -3 ``gloo`` groups.

The host logic under test — slab layout, message protocol, capacity and count book-keeping —
is device-agnostic; the exchange kernels are played by a numpy/torch restatement on the same
message layout.  The per-rank compute hooks are played by the
numpy oracle here (an *engine* plugged in by this test; the product engine is CUDA-only), so
the decomposed trajectory can be compared with the oracle's single-system trajectory:
identical particle sets per slab, forces / velocities / positions to rounding."""

import os
import socket
import sys
import traceback

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from oracle import colliders as ocol, force_manager as ofm, integrators as oint
from helpers import MATS, make_inputs


class OracleEngine:
    """The reference's hooks (oracle restatement) on the rank-local rows of a SlabSystem."""

    def __init__(self, slab, *, box, law, lin, rot, dt, dtype, nmat=1, cell_size=None):
        self.slab, self.box, self.law, self.lin, self.rot, self.dt, self.dtype = slab, box, law, lin, rot, dt, dtype
        self.nmat, self.cell_size = nmat, cell_size

    def _oracle(self, st):
        np_ = lambda t: t.numpy().copy()
        ost = oracle.create_state(np_(st.pos_c), vel=np_(st.vel), force=np_(st.force), ang_vel=np_(st.ang_vel),
                                  torque=np_(st.torque), rad=np_(st.rad), mass=np_(st.mass),
                                  inertia=np_(st.inertia), mat_id=np_(st.mat_id), fixed=np_(st.fixed),
                                  q=np.concatenate([np_(st.q.w), np_(st.q.xyz)], axis=1), dtype=self.dtype)
        oid = getattr(st, "order_id", None)
        if oid is not None:  # in-cell order by global particle id, as the undecomposed system's stable sort gives it
            ost.order_id = np_(oid)
        osy = oracle.create_system(ost, dt=self.dt, linear_integrator_type=self.lin, rotation_integrator_type=self.rot,
                                   collider_type="celllist", collider_kw=dict(cell_size=self.cell_size),
                                   domain_type="periodic", domain_kw=dict(box_size=self.box),
                                   force_model_type=self.law,
                                   mat_table=oracle.make_material_table(MATS[:self.nmat], "harmonic"))
        return ost, osy

    @staticmethod
    def _back(ost, st, fields):
        for f in fields:
            if f == "q":
                st.q.w.copy_(torch.as_tensor(ost.q_w))
                st.q.xyz.copy_(torch.as_tensor(ost.q_xyz))
            else:
                getattr(st, f).copy_(torch.as_tensor(getattr(ost, f)))

    def before_force(self, st):
        if st.N == 0:
            return
        ost, osy = self._oracle(st)
        oint.LINEAR[self.lin][0](ost, osy)
        oint.ROTATION[self.rot][0](ost, osy)
        self._back(ost, st, ("pos_c", "vel", "ang_vel", "q"))

    def compute_force(self, st):
        if st.N == 0:
            return
        ost, osy = self._oracle(st)
        ocol.celllist_compute_force(ost, osy)
        self._back(ost, st, ("force", "torque"))

    def after_force(self, st):
        if st.N == 0:
            return
        ost, osy = self._oracle(st)
        ofm.apply(ost, osy)
        oint.LINEAR[self.lin][1](ost, osy)
        oint.ROTATION[self.rot][1](ost, osy)
        self._back(ost, st, ("force", "torque", "vel", "ang_vel"))

    # ---- CPU restatement of the exchange kernels (csrc/slab.cu) on the same message layout ----
    @staticmethod
    def _cat(bufs, idx, fields, widths):
        return torch.cat([bufs[k][idx].reshape(idx.numel(), widths[k]) for k in fields], dim=1)

    def pack(self, slab):
        from jaxdem_b200.slab import GHOST_FLOAT_FIELDS, ROW_FLOAT_FIELDS, cell_layer, message_views
        n, lay, R, G = slab.n_own, slab.layout, slab.layout.R, slab.layout.n_layers
        lo, up = lay.bounds[slab.rank], lay.bounds[slab.rank + 1]
        b = slab.buf
        layer = cell_layer(b["pos_c"][:n, -1], slab.anchor_last, slab.box_last, G)
        inside = (layer >= lo) & (layer < up)
        below, above = (lo - layer) % G, (layer - (up - 1)) % G
        leave_lo = ~inside & (below >= 1) & (below <= R)
        leave_up = ~inside & ~leave_lo & (above >= 1) & (above <= R)
        stray = ~inside & ~leave_lo & ~leave_up
        stay = inside | stray
        halo_lo, halo_up = inside & (layer < lo + R), inside & (layer >= up - R)
        nz = lambda m: torch.nonzero(m).flatten()
        kept = message_views(slab.kept, slab.kept_layout, slab.dtype)
        for msg, mi, hi, koff in ((slab.send_lo, nz(leave_lo), nz(halo_lo), 0),
                                  (slab.send_up, nz(leave_up), nz(halo_up), slab.migrant_cap)):
            v = message_views(msg, slab.msg_layout, slab.dtype)
            cm, ch = min(mi.numel(), slab.migrant_cap), min(hi.numel(), slab.ghost_cap)
            v["mig_f"][:cm] = self._cat(b, mi[:cm], ROW_FLOAT_FIELDS, slab.widths)
            v["mig_i"][:cm] = torch.stack([b["gid"][mi[:cm]], b["mat_id"][mi[:cm]].long(), b["fixed"][mi[:cm]].long()], 1)
            v["gh_f"][:ch] = self._cat(b, hi[:ch], GHOST_FLOAT_FIELDS, slab.widths)
            v["gh_i"][:ch] = torch.stack([b["gid"][hi[:ch]], b["mat_id"][hi[:ch]].long()], 1)
            v["header"][:3] = torch.tensor([mi.numel(), hi.numel(), int(stray.sum())])
            kept["gh_f"][koff:koff + cm] = self._cat(b, mi[:cm], GHOST_FLOAT_FIELDS, slab.widths)
            kept["gh_i"][koff:koff + cm] = torch.stack([b["gid"][mi[:cm]], b["mat_id"][mi[:cm]].long()], 1)
            slab.holes[koff:koff + cm] = mi[:cm].to(torch.int32)
        slab.header_local[:4] = torch.tensor([int(stay.sum()), int(leave_lo.sum()), int(leave_up.sum()),
                                              int(stray.sum())])

    def unpack(self, slab, counts):
        from jaxdem_b200.slab import GHOST_FLOAT_FIELDS, ROW_FLOAT_FIELDS, message_views
        n_old, a_lo, a_up, k_lo, k_up, g_lo, g_up = counts
        lo = message_views(slab.recv_lo, slab.msg_layout, slab.dtype)
        up = message_views(slab.recv_up, slab.msg_layout, slab.dtype)
        kp = message_views(slab.kept, slab.kept_layout, slab.dtype)
        d, w, cm = slab.buf, slab.widths, slab.migrant_cap
        l, a = k_lo + k_up, a_lo + a_up
        holes = torch.sort(torch.cat([slab.holes[:k_lo], slab.holes[cm:cm + k_up]]).long()).values
        slab.holes[2 * cm:2 * cm + l] = holes.to(torch.int32)
        n_new = n_old - l + a
        names = (*ROW_FLOAT_FIELDS, "gid", "mat_id", "fixed")
        if l > a:  # rows of the tail that are not holes move, in order, into the holes below n_new
            rest = holes[a:]
            tail = torch.arange(n_new, n_old)
            movers = tail[~torch.isin(tail, rest)]
            low = rest[:movers.numel()]
            assert movers.numel() == int((rest < n_new).sum())
            for k in names:
                d[k][low] = d[k][movers]

        def put(fields, f, iv, rows, full):
            c = 0
            cnt = rows.numel()
            for k in fields:
                d[k][rows] = f[:cnt, c:c + w[k]].reshape(cnt, *d[k].shape[1:])
                c += w[k]
            d["gid"][rows] = iv[:cnt, 0]
            d["mat_id"][rows] = iv[:cnt, 1].to(d["mat_id"].dtype)
            d["fixed"][rows] = iv[:cnt, 2].bool() if full else False
            if not full:
                for k, val in (("force", 0), ("torque", 0), ("inertia", 1), ("q_w", 1), ("q_xyz", 0)):
                    d[k][rows] = val

        dst = torch.cat([holes[:min(a, l)], torch.arange(n_old, n_old + max(a - l, 0))])
        put(ROW_FLOAT_FIELDS, lo["mig_f"], lo["mig_i"], dst[:a_lo], True)
        put(ROW_FLOAT_FIELDS, up["mig_f"], up["mig_i"], dst[a_lo:], True)
        row = n_new
        for f, iv, cnt in ((kp["gh_f"], kp["gh_i"], k_lo), (kp["gh_f"][cm:], kp["gh_i"][cm:], k_up),
                           (lo["gh_f"], lo["gh_i"], g_lo), (up["gh_f"], up["gh_i"], g_up)):
            put(GHOST_FLOAT_FIELDS, f, iv, torch.arange(row, row + cnt), False)
            row += cnt


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, cfg, out):
    try:
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        torch.set_num_threads(1)
        from jaxdem_b200.slab import SlabSystem
        dtype = cfg["dtype"]
        F = torch.float32 if dtype == np.float32 else torch.float64
        inp = cfg["inp"]
        dim = inp["pos"].shape[1]
        slab = SlabSystem(dim=dim, dtype=F, device="cpu", capacity=cfg["capacity"], box=inp["box"],
                          anchor=np.zeros(dim), n_layers=cfg["n_layers"], search_range=1,
                          ghost_capacity=cfg["capacity"], migrant_capacity=cfg["capacity"])
        slab.load_global(dict(pos=inp["pos"], vel=inp["vel"], ang_vel=inp["ang_vel"], rad=inp["rad"],
                              mass=inp["mass"], mat_id=inp.get("mat_id", np.zeros(len(inp["rad"]), np.int64))))
        slab.engine = OracleEngine(slab, box=inp["box"], law=cfg["law"], lin="verlet", rot=cfg["rot"], dt=cfg["dt"],
                                   dtype=dtype, nmat=cfg["nmat"], cell_size=cfg["cell_size"])
        owned0 = slab.n_own
        slab.compute_force()  # loop-carried state.force of the first step
        if cfg.get("expect_error"):
            try:
                slab.step(cfg["steps"])
                out.put(("ok", "no error raised", owned0, False))
            except RuntimeError as e:  # every rank sees the stray flag (its own or a neighbour's header)
                out.put(("ok", str(e), owned0, True))
            dist.barrier()
            dist.destroy_process_group()
            return
        slab.step(cfg["steps"])
        res = slab.gather(("pos_c", "vel", "force", "torque", "ang_vel"))
        # every owned particle lies in this rank's layers; ghosts only within the halo
        from jaxdem_b200.slab import cell_layer
        lay = slab.layout
        layer = cell_layer(slab.buf["pos_c"][:slab.n_own, -1], 0.0, float(inp["box"][-1]), lay.n_layers).numpy()
        ok_own = bool((lay.owner[layer] == rank).all())
        if rank == 0:
            out.put(("ok", res, owned0, ok_own))
        else:
            out.put(("aux", None, owned0, ok_own))
        dist.barrier()
        dist.destroy_process_group()
    except Exception:  # surface the traceback in the parent
        out.put(("err", traceback.format_exc(), 0, False))


def _run(world, cfg):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cfg, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = [out.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for g in got:
        assert g[0] != "err", g[1]
    res = [g for g in got if g[0] == "ok"][0][1]
    assert sum(g[2] for g in got) == cfg["inp"]["pos"].shape[0]
    assert all(g[3] for g in got)
    return res


def _reference(cfg):
    inp, dtype = cfg["inp"], cfg["dtype"]
    ost = oracle.create_state(inp["pos"], vel=inp["vel"], ang_vel=inp["ang_vel"], rad=inp["rad"], mass=inp["mass"],
                              mat_id=inp.get("mat_id"), dtype=dtype)
    osy = oracle.create_system(ost, dt=cfg["dt"], linear_integrator_type="verlet", rotation_integrator_type=cfg["rot"],
                               collider_type="celllist", collider_kw=dict(cell_size=cfg["cell_size"]),
                               domain_type="periodic", domain_kw=dict(box_size=inp["box"]),
                               force_model_type=cfg["law"],
                               mat_table=oracle.make_material_table(MATS[:cfg["nmat"]], "harmonic"))
    ocol.celllist_compute_force(ost, osy)
    oracle.step(ost, osy, cfg["steps"])
    return ost


@pytest.mark.parametrize("world,dim,law,rot,dtype", [
    (2, 3, "spring", "", np.float64),
    (2, 3, "cundallstrack", "verletspiral", np.float64),
    (3, 2, "hertz", "verletspiral", np.float32),
])
def test_slab_decomposition_matches_single_system(world, dim, law, rot, dtype):
    n = 600 if dim == 3 else 400
    nmat = 2 if law != "spring" else 1
    inp = make_inputs(n, dim, seed=21, dtype=dtype, phi=0.5 if dim == 3 else 0.6, nmat=nmat)
    inp["vel"] = inp["vel"] * 40.0  # fast enough for particles to change slab within the run
    inp["pos"][::7, -1] -= inp["box"][-1]  # un-wrapped coordinates: ownership goes by the wrapped cell
    cs = dtype(2.0 * inp["rad"].max())
    n_layers = int(np.floor(dtype(inp["box"][-1]) / cs))
    cfg = dict(inp=inp, dtype=dtype, law=law, rot=rot, dt=2e-3, steps=12, nmat=nmat, cell_size=cs,
               n_layers=n_layers, capacity=n + 64)
    got = _run(world, cfg)
    ref = _reference(cfg)
    assert np.array_equal(got["gid"], np.arange(n))
    for f in ("pos_c", "vel", "force", "torque", "ang_vel"):
        a, b = got[f], getattr(ref, f)
        # same arithmetic per pair AND the same order of every particle's contact sum: ties inside a cell are
        # broken by the global particle id (State.order_id), i.e. each rank's perm is the undecomposed system's
        # stable sort of (hash, iota) (colliders/_partition.py:91-93) restricted to its rows => bit for bit
        assert np.array_equal(a, b), (f, np.abs(a - b).max())
    # particles did migrate
    z0 = inp["pos"][:, -1]
    moved = np.abs(ref.pos_c[:, -1] - z0).max()
    assert moved > 0.5 * inp["box"][-1] / n_layers


def test_slab_step_too_large_raises():
    """A particle that moves further than the halo in one step cannot be handed over by the
    single neighbour exchange: the exchange reports it (every rank raises) instead of losing it."""
    dtype = np.float64
    inp = make_inputs(1200, 3, seed=3, dtype=dtype, phi=0.4)
    cs = dtype(2.0 * inp["rad"].max())
    n_layers = int(np.floor(dtype(inp["box"][-1]) / cs))
    inp["vel"][:] = 0.0
    lz = float(inp["box"][-1])
    j = int(np.argmin(np.abs(inp["pos"][:, -1] - 0.25 * lz)))  # a particle in the middle of rank 0's slab ...
    inp["vel"][j, -1] = 0.5 * lz / 1e-2                        # ... jumps half a box in one step of dt = 1e-2
    cfg = dict(inp=inp, dtype=dtype, law="spring", rot="", dt=1e-2, steps=1, nmat=1, cell_size=cs,
               n_layers=n_layers, capacity=1264, expect_error=True)
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, cfg, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = [out.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    for g in got:
        assert g[0] != "err", g[1]
    # the rank that owned the particle and its neighbour both see the flag
    assert sum(1 for g in got if g[3] and "further than the halo" in g[1]) == 2, got
