"""Multi-GPU slab decomposition (needs >= 2 GPUs; skipped otherwise): N ranks, one GPU each,
exchanging halo planes either by direct peer stores from inside the update kernels (CUDA IPC,
halo="p2p") or with torch.distributed point-to-point operations over NCCL (halo="sendrecv"),
must reproduce the whole-grid oracle bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, outdir, shape, halo):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch
    import torch.distributed as dist
    import refcfg
    from seismic_cpml_b200 import lib as L
    from seismic_cpml_b200.slab import GpuSlab, SlabDriver, owner_of_plane
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    nx, ny, nz, npml, nstep = shape
    c = refcfg.cfg3d(nx=nx, ny=ny, nz=nz, npml=npml, nstep=nstep)
    s = L.Solver(ndim=3, order=2, nx=nx, ny=ny, nz=nz, nstep=nstep, npoints_pml=npml, nrec=len(c["ix_rec"]),
                 isource=c["isource"], jsource=c["jsource"], nslabs=world, slab_rank=rank, device=rank,
                 deltax=c["deltax"], deltay=c["deltay"], deltaz=c["deltaz"], deltat=c["deltat"], lam=c["lam"],
                 mu=c["mu"], lambdaplustwomu=c["lambdaplustwomu"], rho=c["rho"], cp=3300.0)
    s.set_profiles(0, c["prof_x"]); s.set_profiles(1, c["prof_y"]); s.set_profiles(2, c["prof_z"])
    s.set_source_series(c["force_x"], c["force_y"])
    s.set_receivers(c["ix_rec"], c["iy_rec"])
    drv = SlabDriver(GpuSlab(s), rank, world, s.nzl, halo=halo)
    drv.run(1, nstep)
    owner = owner_of_plane(nz // 2, nz, world)
    sx, sy = drv.seismograms(owner)
    e = drv.total_energy()
    vn = drv.maxnorm()
    np.savez(os.path.join(outdir, f"r{rank}.npz"), sx=sx, sy=sy, e=e, vn=vn, vz=s.get_field(2))
    dist.barrier()
    s.close()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("halo", ["p2p", "sendrecv"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_slabs_match_oracle(world, halo, tmp_path):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    import refcfg
    from oracle import oracle as O
    shape = (40, 37, 48, 5, 80)
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), shape, halo), nprocs=world, join=True)
    nx, ny, nz, npml, nstep = shape
    c = refcfg.cfg3d(nx=nx, ny=ny, nz=nz, npml=npml, nstep=nstep)
    o = O.run_3d_iso(**c, nproc=2, want_fields=True)
    nzl = nz // world
    for r in range(world):
        d = np.load(tmp_path / f"r{r}.npz")
        assert np.array_equal(d["sx"], o["sisvx"]) and np.array_equal(d["sy"], o["sisvy"])
        assert refcfg.rel_l2(d["e"], o["total_energy"]) <= 1e-11
        assert float(d["vn"]) == pytest.approx(o["vnorm"], rel=1e-15)
        assert np.array_equal(d["vz"], o["vz"][r * nzl:(r + 1) * nzl])


# ---- viscoelastic slabs: NCCL send/recv of the complete fourth-order halo ----------------------

def _visco_worker(rank, world, port, outdir, shape, emulate):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch
    import torch.distributed as dist
    import refcfg
    import test_gpu_visco as TV
    from seismic_cpml_b200.slab import GpuSlab, SlabDriver, owner_of_plane
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    nx, ny, nz, npml, nstep = shape
    c = refcfg.cfgv3d(nx=nx, ny=ny, nz=nz, npml=npml, nstep=nstep)
    s = TV.solver_visco(c, emulate_nproc=emulate, nslabs=world, slab_rank=rank, device=rank)
    drv = SlabDriver(GpuSlab(s), rank, world, s.nzl, visco=True)
    drv.run(1, nstep)
    owner = owner_of_plane(nz // 2, nz, world)
    sx, sy = drv.seismograms(owner)
    e = drv.total_energy()
    np.savez(os.path.join(outdir, f"r{rank}.npz"), sx=sx, sy=sy, e=e, vz=s.get_field(2), sxy_r=s.get_field(12))
    dist.barrier()
    s.close()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("emulate", [1, 4])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_visco_slabs_match_oracle(world, emulate, tmp_path):
    """The result must depend on emulate_nproc (the reference's NPROC) only, never on the number of GPUs."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    import refcfg
    from oracle import oracle as O
    shape = (40, 37, 48, 5, 80)
    mp.spawn(_visco_worker, args=(world, _free_port(), str(tmp_path), shape, emulate), nprocs=world, join=True)
    nx, ny, nz, npml, nstep = shape
    c = refcfg.cfgv3d(nx=nx, ny=ny, nz=nz, npml=npml, nstep=nstep)
    o = O.run_3d_visco(**c, nproc=emulate, want_fields=True)
    nzl = nz // world
    for r in range(world):
        d = np.load(tmp_path / f"r{r}.npz")
        assert np.array_equal(d["sx"], o["sisvx"]) and np.array_equal(d["sy"], o["sisvy"])
        assert refcfg.rel_l2(d["e"], o["total_energy"]) <= 1e-11
        assert np.array_equal(d["vz"], o["vz"][r * nzl:(r + 1) * nzl])
        assert np.array_equal(d["sxy_r"], o["sigmaxy_R"][r * nzl:(r + 1) * nzl])
