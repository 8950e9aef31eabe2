"""World-size-2 (and 4) gloo tests of the N>1 path on CPU: the product's SlabDriver
(plane exchange, energy reduction, seismogram ownership) driving numpy slabs must
reproduce the whole-grid oracle -- the same code drives the GPU slabs over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, outdir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch.distributed as dist
    import refcfg
    from np_slab import NumpySlab
    from seismic_cpml_b200.slab import SlabDriver, owner_of_plane
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c = refcfg.cfg3d(nx=22, ny=26, nz=24, npml=4, nstep=40)
    slab = NumpySlab(c, world, rank)
    drv = SlabDriver(slab, rank, world, slab.nzl)
    drv.run(1, 25)
    drv.run(26, c["nstep"])
    owner = owner_of_plane(c["nz"] // 2, c["nz"], world)
    sx, sy = drv.seismograms(owner)
    e = drv.total_energy()
    vn = drv.maxnorm()
    own_sx = slab.get_seismograms()[0]
    np.savez(os.path.join(outdir, f"r{rank}.npz"), sx=sx, sy=sy, e=e, vn=vn, owner=owner,
             own_nonzero=bool(own_sx.any()), vx=slab.f["vx"][1:slab.nzl + 1, 1:-1, 1:-1], sent=drv.bytes_sent)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_slab_driver_gloo_matches_oracle(world, tmp_path):
    import refcfg
    from oracle import oracle as O
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    c = refcfg.cfg3d(nx=22, ny=26, nz=24, npml=4, nstep=40)
    o = O.run_3d_iso(**c, nproc=world, want_fields=True)
    res = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    nzl = c["nz"] // world
    for r, d in enumerate(res):
        assert int(d["owner"]) == world // 2 - 1                  # rank_cut_plane, :346
        assert np.array_equal(d["sx"], o["sisvx"]) and np.array_equal(d["sy"], o["sisvy"])
        assert bool(d["own_nonzero"]) == (r == world // 2 - 1)
        assert refcfg.rel_l2(d["e"], o["total_energy"]) < 1e-13
        assert float(d["vn"]) == pytest.approx(o["vnorm"], rel=1e-15)
        assert np.array_equal(d["vx"], o["vx"][r * nzl:(r + 1) * nzl])
        # 40 steps x 3 planes per interior interface direction
        n_if = (1 if r > 0 else 0) + (1 if r < world - 1 else 0)
        assert int(d["sent"]) == 40 * 3 * n_if * (c["nx"] + 2) * (c["ny"] + 2) * 8


# ---- viscoelastic exchange plan (fourth order: two planes one way, one the other) --------------

class _FakeViscoSlab:
    """Planes tagged with (rank, field, klocal): checks WHERE the driver moves data, not the physics."""

    def __init__(self, rank, nzl, npts=7):
        import torch
        self.rank, self.nzl = rank, nzl
        self.f = {}
        for field in (0, 1, 2, 5, 7, 8):
            t = torch.full((nzl + 4, npts), -1.0, dtype=torch.float64)
            for k in range(1, nzl + 1):
                t[k + 1] = rank * 10000 + field * 100 + k
            self.f[field] = t

    def plane(self, field, klocal, nplanes=1):
        return self.f[field][klocal + 1:klocal + 1 + nplanes].reshape(-1)


def _visco_plan_worker(rank, world, port, outdir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch.distributed as dist
    from seismic_cpml_b200.slab import PHASE_S, PHASE_V, SlabDriver
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nzl = 6
    slab = _FakeViscoSlab(rank, nzl)
    drv = SlabDriver(slab, rank, world, nzl, visco=True)
    drv.exchange(PHASE_V)
    drv.exchange(PHASE_S)
    np.savez(os.path.join(outdir, f"v{rank}.npz"), sent=drv.bytes_sent,
             **{f"f{k}": v.numpy() for k, v in slab.f.items()})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_visco_exchange_plan_gloo(world, tmp_path):
    """3D-visco :962-975 / :1229-1242 plus the planes the reference never sends (quirk B6)."""
    mp.spawn(_visco_plan_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    nzl = 6
    res = [np.load(tmp_path / f"v{r}.npz") for r in range(world)]

    def tag(r, field, k):
        return r * 10000 + field * 100 + k

    for r, d in enumerate(res):
        for field, two_planes_go in ((0, "left"), (1, "left"), (5, "left"), (2, "right"), (7, "right"), (8, "right")):
            a = d[f"f{field}"][:, 0]          # index = klocal + 1
            hi = [a[nzl + 2], a[nzl + 3]]     # planes NZ_LOCAL+1, NZ_LOCAL+2
            lo = [a[0], a[1]]                 # planes -1, 0
            if two_planes_go == "left":
                want_hi = [tag(r + 1, field, 1), tag(r + 1, field, 2)] if r < world - 1 else [-1, -1]
                want_lo = [-1, tag(r - 1, field, nzl)] if r > 0 else [-1, -1]
            else:
                want_hi = [tag(r + 1, field, 1), -1] if r < world - 1 else [-1, -1]
                want_lo = [tag(r - 1, field, nzl - 1), tag(r - 1, field, nzl)] if r > 0 else [-1, -1]
            assert hi == want_hi and lo == want_lo, (r, field, hi, want_hi, lo, want_lo)
            assert list(a[2:nzl + 2]) == [tag(r, field, k) for k in range(1, nzl + 1)]   # owned planes untouched
        n_if = (1 if r > 0 else 0) + (1 if r < world - 1 else 0)
        # per interface and rank: 3 fields x (2 + 1 planes) over the two directions = 9 planes sent
        assert int(d["sent"]) == n_if * 9 * 7 * 8
