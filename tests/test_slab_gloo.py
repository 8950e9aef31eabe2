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
