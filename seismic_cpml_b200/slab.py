"""z-slab decomposition of the 3-D solver across ranks: one process per GPU, the plane
exchange of the reference's MPI_SENDRECV calls done with torch.distributed point-to-point
operations (NCCL over NVLink between GPUs; gloo in the CPU tests) -- or, on GPUs and by
default, not done by the driver at all: with halo="p2p" the slabs are attached to each other
through CUDA IPC (cpml_p2p_export / cpml_p2p_attach_ipc) and the update kernels store the
boundary planes straight into the neighbour GPU's halo planes over NVLink, ordered by
device-side flags; torch.distributed then only carries the 64-byte IPC blobs at start-up and
the reductions of the results.

Reference layout (seismic_CPML_3D_isotropic_MPI_OpenMP.f90):
  * rank r owns global planes r*NZ_LOCAL+1 .. (r+1)*NZ_LOCAL (:131,:397), arrays carry
    halo planes 0 and NZ_LOCAL+1 (:273);
  * before the stress update (:811-823): vx(:,:,1), vy(:,:,1) go to rank-1's plane
    NZ_LOCAL+1 ("left shift"), vz(:,:,NZ_LOCAL) goes to rank+1's plane 0 ("right shift");
  * before the velocity update (:951-963): sigmazz(:,:,1) left, sigmayz(:,:,NZ_LOCAL) and
    sigmaxz(:,:,NZ_LOCAL) right;
  * total_energy(it) is MPI_REDUCE(SUM) of the ranks' shares (:1179); seismograms live on
    the rank that owns the cut plane (:346,1124).
The end ranks have MPI_PROC_NULL neighbours (:775-790): their outer halo planes are never
written and stay zero.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

F_VX, F_VY, F_VZ, F_SZZ, F_SXZ, F_SYZ = 0, 1, 2, 5, 7, 8

# (field, direction) of the two exchange phases
PHASE_V = ((F_VX, "left"), (F_VY, "left"), (F_VZ, "right"))        # :811-823
PHASE_S = ((F_SZZ, "left"), (F_SYZ, "right"), (F_SXZ, "right"))    # :951-963


def exchange_plan(phase, nzl, visco=False):
    """Messages of one exchange phase as (field, direction, first plane sent, plane count, first
    plane received into), local k.

    Isotropic (second order): one plane per field -- "left": my plane 1 -> left's NZ_LOCAL+1,
    "right": my plane NZ_LOCAL -> right's 0.
    Viscoelastic (fourth order, seismic_CPML_3D_viscoelastic_MPI.f90:962-975, :1229-1242): two
    planes in the reference's direction (planes 1:2 -> NZ_LOCAL+1:NZ_LOCAL+2, or
    NZ_LOCAL-1:NZ_LOCAL -> -1:0) PLUS the one plane the other way that the reference never sends
    although its stencils read it.  The kernels decide per plane which taps
    read zero (cpml_config.emulate_nproc), so GPU slabs always exchange the complete halo and the
    result does not depend on the number of GPUs."""
    plan = []
    for field, direction in phase:
        if not visco:
            plan.append((field, "left", 1, 1, nzl + 1) if direction == "left" else (field, "right", nzl, 1, 0))
        elif direction == "left":
            plan.append((field, "left", 1, 2, nzl + 1))
            plan.append((field, "right", nzl, 1, 0))
        else:
            plan.append((field, "right", nzl - 1, 2, -1))
            plan.append((field, "left", 1, 1, nzl + 1))
    return plan


class _DevicePlane:
    """Zero-copy torch view of a device plane owned by libcpml_b200 (__cuda_array_interface__)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes // 8,), "typestr": "<f8",
                                         "data": (ptr, False), "version": 2}


def plane_tensor(solver, field: int, klocal: int, nplanes: int = 1) -> torch.Tensor:
    """View of `nplanes` consecutive planes starting at klocal (planes are contiguous in k)."""
    ptr, nbytes = solver.halo_plane(field, klocal)
    if nplanes > 1:
        ptr2, _ = solver.halo_plane(field, klocal + nplanes - 1)
        assert ptr2 - ptr == (nplanes - 1) * nbytes
    return torch.as_tensor(_DevicePlane(ptr, nbytes * nplanes), device=torch.device("cuda", torch.cuda.current_device()))


def owner_of_plane(kglobal: int, nz: int, nslabs: int) -> int:
    """Rank that holds global plane kglobal; for kglobal = NZ/2 this is nb_procs/2 - 1 (:346)."""
    return (kglobal - 1) // (nz // nslabs)


class SlabDriver:
    """Drives one slab (`backend`) of an nslabs-way decomposition.

    backend: an object with step_stress(it), step_velocity(it), step_finish(it),
    synchronize(), get_seismograms(), get_energy() and `plane(field, klocal) -> torch.Tensor`
    (a view, so that receives land in place) -- seismic_cpml_b200.lib.Solver wrapped by
    GpuSlab below, or the numpy slab of tests/ on CPU.
    """

    def __init__(self, backend, rank: int, nslabs: int, nzl: int, group=None, halo: str = "sendrecv",
                 visco: bool = False):
        self.b, self.rank, self.nslabs, self.nzl, self.group = backend, rank, nslabs, nzl, group
        self.visco = visco
        self.plan_v = exchange_plan(PHASE_V, nzl, visco)
        self.plan_s = exchange_plan(PHASE_S, nzl, visco)
        self.left = rank - 1 if rank > 0 else None          # MPI_PROC_NULL at the ends
        self.right = rank + 1 if rank < nslabs - 1 else None
        self._planes = {}
        self.bytes_sent = 0
        if halo not in ("sendrecv", "p2p"):
            raise ValueError("halo must be 'sendrecv' or 'p2p'")
        self.halo = halo
        if halo == "p2p" and nslabs > 1:
            self._attach_peers()

    def _attach_peers(self):
        """Every rank publishes the IPC blob of its slab; each attaches its two neighbours."""
        blobs = [None] * self.nslabs
        dist.all_gather_object(blobs, self.b.p2p_export(), group=self.group)
        if self.left is not None:
            self.b.p2p_attach_ipc(0, blobs[self.left])
        if self.right is not None:
            self.b.p2p_attach_ipc(1, blobs[self.right])
        dist.barrier(group=self.group)      # nobody steps before every neighbour is mapped

    def _plane(self, field, klocal, nplanes=1):
        key = (field, klocal, nplanes)
        if key not in self._planes:
            self._planes[key] = self.b.plane(field, klocal) if nplanes == 1 else self.b.plane(field, klocal, nplanes)
        return self._planes[key]

    def exchange(self, plan):
        """One group of MPI_SENDRECV calls: all sends and receives of the phase in flight at once."""
        if plan is PHASE_V:
            plan = self.plan_v
        elif plan is PHASE_S:
            plan = self.plan_s
        ops = []
        for field, direction, k_send, n, k_recv in plan:
            if direction == "left":      # my low planes -> left's high halo ; right's low planes -> my high halo
                if self.left is not None:
                    ops.append(dist.P2POp(dist.isend, self._plane(field, k_send, n), self.left, self.group))
                if self.right is not None:
                    ops.append(dist.P2POp(dist.irecv, self._plane(field, k_recv, n), self.right, self.group))
            else:                        # my high planes -> right's low halo ; left's high planes -> my low halo
                if self.right is not None:
                    ops.append(dist.P2POp(dist.isend, self._plane(field, k_send, n), self.right, self.group))
                if self.left is not None:
                    ops.append(dist.P2POp(dist.irecv, self._plane(field, k_recv, n), self.left, self.group))
        if not ops:
            return
        for op in ops:
            if op.op is dist.isend:
                self.bytes_sent += op.tensor.numel() * 8
        for req in dist.batch_isend_irecv(ops):
            req.wait()

    def step(self, it: int):
        """One pass of the loop body :804-1180 for this slab."""
        if self.halo == "p2p":              # the kernels exchange the planes themselves
            self.b.step_stress(it)
            self.b.step_velocity(it)
            self.b.step_finish(it)
            return
        self.exchange(PHASE_V)
        self.b.step_stress(it)
        self.exchange(PHASE_S)
        self.b.step_velocity(it)
        self.b.step_finish(it)

    def reset(self):
        """cpml_reset on every slab for a second run.  Nobody resets while a neighbour may still be storing into its
        halo planes, and nobody steps before every slab is reset (its first step writes into the neighbours' halos):
        barriers on both sides, like the synchronisation around the zeroing of :720-756 in an MPI run."""
        self.b.synchronize()
        if self.nslabs > 1:
            dist.barrier(group=self.group)
        self.b.reset()
        if self.nslabs > 1:
            dist.barrier(group=self.group)

    def run(self, it_begin: int, it_end: int):
        for it in range(it_begin, it_end + 1):
            self.step(it)
        self.b.synchronize()

    # ---- results, as rank_cut_plane sees them in the reference
    def total_energy(self) -> np.ndarray:
        e = torch.from_numpy(np.ascontiguousarray(self.b.get_energy()[0]))
        if self.nslabs > 1:
            dev = self._reduce_device()
            e = e.to(dev)
            dist.all_reduce(e, op=dist.ReduceOp.SUM, group=self.group)   # MPI_REDUCE(SUM), :1179
            e = e.cpu()
        return e.numpy()

    def seismograms(self, owner: int):
        sx, sy = self.b.get_seismograms()
        if self.nslabs > 1:
            dev = self._reduce_device()
            t = torch.from_numpy(np.stack([sx, sy])).to(dev)
            dist.broadcast(t, src=owner, group=self.group)
            t = t.cpu().numpy()
            sx, sy = t[0], t[1]
        return sx, sy

    def maxnorm(self) -> float:
        v = torch.tensor([self.b.get_maxnorm()], dtype=torch.float64)
        if self.nslabs > 1:
            v = v.to(self._reduce_device())
            dist.all_reduce(v, op=dist.ReduceOp.MAX, group=self.group)   # MPI_REDUCE(MAX), :1185
        return float(v.cpu()[0])

    def _reduce_device(self):
        return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(self.group) == "nccl" \
            else torch.device("cpu")


class GpuSlab:
    """lib.Solver with the `plane` view SlabDriver needs; kernels run on torch's current stream so
    that they are ordered with the NCCL sends/receives."""

    def __init__(self, solver):
        self.s = solver
        solver.set_stream(torch.cuda.current_stream().cuda_stream)

    def plane(self, field, klocal, nplanes=1):
        return plane_tensor(self.s, field, klocal, nplanes)

    def __getattr__(self, name):
        return getattr(self.s, name)
