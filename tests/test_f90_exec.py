"""oracle/f90_exec.py, the transliterator that executes a reference program from its Fortran source text.

Two layers: (1) the Fortran semantics the reference programs rely on, on small synthetic programs (integer division,
`**`, array lower bounds, sections, SUM order, if / else if / else, one-line if, continuation lines, comments, case
insensitivity, MPI_SENDRECV / MPI_REDUCE between threads); (2) where the reference is present (this container; not the
GPU box) one program is re-run from its source on a tiny grid and compared with the C oracle bit for bit -- the same
check tests/test_reference_vectors.py does against the committed vectors, live.
"""
import os
import textwrap

import numpy as np
import pytest

import refcfg
from oracle import f90_exec as F
from oracle import oracle as O

REF = os.environ.get("CPML_REFERENCE_DIR", "/root/reference")


def run(tmp_path, src, nproc=1, **kw):
    p = tmp_path / "t.f90"
    p.write_text(textwrap.dedent(src))
    return F.run_program(str(p), nproc=nproc, **kw)


def test_scalar_and_array_semantics(tmp_path):
    sp = run(tmp_path, """
      program t
      implicit none
      integer, parameter :: N = 5
      double precision, parameter :: HALF = 1.d0 / 2, TINY = 1.5d-3
      double precision, dimension(0:N+1) :: a      ! lower bound 0
      double precision, dimension(N,2) :: b
      integer, dimension(2) :: idx
      integer :: i, k, m
      double precision :: s, x, lambda
      k = 7 / 2                  ! integer division truncates
      m = -7 / 2
      x = 7.d0 / 2
      lambda = 2.d0              ! a Python keyword
      do i = 0, N+1
        a(i) = dble(i)**2 + &    ! continuation
               lambda * HALF
      enddo
      b(:,:) = 1.d0
      B(2,:) = 3.D0              ! case-insensitive
      idx(1) = 2
      idx(2) = 1
      s = sum(b(1:3,1)) + b(idx(1),idx(2))
      if (k == 3 .and. x > 3.d0) then
        s = s + TINY
      else if (k /= 3) then
        s = -1.d0
      else
        s = -2.d0
      endif
      if (s > 0.d0) k = k + mod(17,5)
      if (s < 0.d0) stop 'never'
      print *, 'skipped', s
      end program t
      """)[0]
    assert sp["k"] == 5 and sp["m"] == -3 and sp["x"] == 3.5
    assert np.array_equal(sp["a"], np.arange(7.0) ** 2 + 1.0)
    assert sp["s"] == (1.0 + 3.0 + 1.0) + 3.0 + 1.5e-3
    assert isinstance(sp["k"], int)


def test_sum_is_sequential_in_array_element_order(tmp_path):
    sp = run(tmp_path, """
      program t
      double precision, dimension(3,2) :: b
      double precision :: s
      b(1,1) = 1.d16
      b(2,1) = 1.d0
      b(3,1) = -1.d16
      b(1,2) = 1.d0
      b(2,2) = 1.d0
      b(3,2) = 1.d0
      s = sum(b)
      end program t
      """)[0]
    # ((((1e16 + 1) - 1e16) + 1) + 1) + 1: the first 1 is absorbed
    assert sp["s"] == 3.0


def test_stop_raises(tmp_path):
    with pytest.raises(RuntimeError, match="too large"):
        run(tmp_path, """
          program t
          double precision :: c
          c = 2.d0
          if (c > 1.d0) stop 'time step is too large'
          end program t
          """)


def test_mpi_sendrecv_and_reduce_between_rank_threads(tmp_path):
    sp = run(tmp_path, """
      program t
      implicit none
      include 'mpif.h'
      integer, parameter :: NX = 3, NZ_LOCAL = 2
      double precision, dimension(NX,0:NZ_LOCAL+1) :: v
      double precision :: total, biggest
      integer :: nb_procs, rank, code, i, k, up, down
      integer, dimension(MPI_STATUS_SIZE) :: message_status
      call MPI_INIT(code)
      call MPI_COMM_SIZE(MPI_COMM_WORLD, nb_procs, code)
      call MPI_COMM_RANK(MPI_COMM_WORLD, rank, code)
      up = rank + 1
      down = rank - 1
      if (rank == 0) down = MPI_PROC_NULL
      if (rank == nb_procs - 1) up = MPI_PROC_NULL
      do k = 1,NZ_LOCAL
        do i = 1,NX
          v(i,k) = 100*rank + 10*k + i
        enddo
      enddo
      call MPI_SENDRECV(v(:,NZ_LOCAL),NX,MPI_DOUBLE_PRECISION,up,0,v(:,0),NX,MPI_DOUBLE_PRECISION,down,0, &
                        MPI_COMM_WORLD,message_status,code)
      call MPI_SENDRECV(v(:,1),NX,MPI_DOUBLE_PRECISION,down,0,v(:,NZ_LOCAL+1),NX,MPI_DOUBLE_PRECISION,up,0, &
                        MPI_COMM_WORLD,message_status,code)
      call MPI_REDUCE(sum(v(:,1:NZ_LOCAL)),total,1,MPI_DOUBLE_PRECISION,MPI_SUM,0,MPI_COMM_WORLD,code)
      call MPI_REDUCE(maxval(v),biggest,1,MPI_DOUBLE_PRECISION,MPI_MAX,0,MPI_COMM_WORLD,code)
      call MPI_FINALIZE(code)
      end program t
      """, nproc=3)
    assert [s["rank"] for s in sp] == [0, 1, 2]
    assert np.array_equal(sp[1]["v"][:, 0], sp[0]["v"][:, 2]) and np.array_equal(sp[1]["v"][:, 3], sp[2]["v"][:, 1])
    assert not sp[0]["v"][:, 0].any() and not sp[2]["v"][:, 3].any()          # MPI_PROC_NULL: nothing arrives
    assert sp[0]["total"] == sum(100 * r + 10 * k + i for r in range(3) for k in (1, 2) for i in (1, 2, 3))
    assert sp[0]["biggest"] == 223.0 and sp[1]["total"] == 0.0


def test_parameter_overrides_and_statement_edits(tmp_path):
    sp = run(tmp_path, """
      program t
      integer, parameter :: NX = 101
      integer, parameter :: NHALF = NX / 2
      double precision, dimension(NX) :: a
      double precision :: x
      x = 500.d0
      end program t
      """, overrides={"NX": "9"}, edits=[(r"500\.d0", "20.d0")])[0]
    assert sp["nx"] == 9 and sp["nhalf"] == 4 and sp["a"].shape == (9,) and sp["x"] == 20.0


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "seismic_CPML_3D_isotropic_MPI_OpenMP.f90")),
                    reason="the reference tree is not on this machine")
def test_live_reference_run_equals_the_oracle():
    """seismic_CPML_3D_isotropic_MPI_OpenMP.f90 from its source, 12 x 14 x 8 points on two ranks, 12 steps."""
    nx, ny, nz, npml, nstep = 12, 14, 8, 2, 12
    ov = {"NX": nx, "NY": ny, "NZ": nz, "NPROC": 2, "NSTEP": nstep, "NPOINTS_PML": npml, "ydeb": f"{(ny // 3) * 10}.d0", "yfin": "30.d0"}
    sp = F.run_program(os.path.join(REF, "seismic_CPML_3D_isotropic_MPI_OpenMP.f90"), {k: str(v) for k, v in ov.items()}, nproc=2)
    r = sp[sp[0]["rank_cut_plane"]]
    c = refcfg.cfg3d(nx=nx, ny=ny, nz=nz, nstep=nstep, npml=npml)
    o = O.run_3d_iso(**c, nproc=2, want_fields=True)
    assert np.array_equal(r["a_x_half"], np.asarray(c["prof_x"]["a_half"])) and np.array_equal(r["b_z"], np.asarray(c["prof_z"]["b"]))
    assert np.abs(o["sisvx"]).max() > 0
    assert np.array_equal(r["sisvx"].T, o["sisvx"]) and np.array_equal(r["sisvy"].T, o["sisvy"])
    assert np.array_equal(r["total_energy"], o["total_energy"])
    for f in ("vx", "vz", "sigmaxx", "sigmayz"):
        g = np.concatenate([q[f][:, :, 1:nz // 2 + 1] for q in sp], axis=2).transpose(2, 1, 0)
        assert np.array_equal(g, o[f]), f


VEC_SRC = """
  program t
  implicit none
  integer, parameter :: NX = 7, NY = 5
  double precision, dimension(0:NX+1,0:NY+1) :: u, w
  double precision, dimension(NX) :: bx
  double precision, dimension(NY) :: acc
  double precision, dimension(NX,NY,2) :: r
  double precision, dimension(2) :: phi
  double precision :: d, total, chain, ssum
  integer :: i, j, m
  do j = 0,NY+1
    do i = 0,NX+1
      u(i,j) = 1.d0 / dble(3 + i + 10*j)
    enddo
  enddo
  do i = 1,NX
    bx(i) = 0.5d0 + 0.1d0*dble(i)
  enddo
  total = 0.d0
  acc(:) = 0.d0
  do j = 1,NY
    do i = 1,NX
      d = (u(i+1,j) - u(i-1,j)) * bx(i)           ! private scalar, neighbours of an array that is not written
      w(i,j) = bx(i) * w(i,j) + d / 3.d0          ! read-modify-write at (i,j) itself
      d = d + u(i,j+1)
      total = total + d * 1.d16                   ! ordered accumulation into a scalar
      acc(3) = acc(3) + d                         ! ... and into one array element
    enddo
  enddo
  phi(1) = 0.25d0
  phi(2) = 0.75d0
  r(:,:,:) = 0.d0
  do j = 1,NY
    do i = 1,NX
      ssum = 0.d0
      do m = 1,2                                  ! a loop over another variable inside the nest stays a loop
        r(i,j,m) = (r(i,j,m) + u(i,j) * phi(m)) / (1.d0 + phi(m))
        ssum = ssum + r(i,j,m)
      enddo
      w(i,j) = w(i,j) + ssum
    enddo
  enddo
  chain = 0.d0
  do i = 1,NX
    w(i,1) = w(i-1,1) + u(i,1)                    ! a true dependence between iterations: must stay a loop
  enddo
  end program t
  """


def test_vectorised_loops_give_the_same_bits_and_keep_true_dependences_sequential(tmp_path):
    a = run(tmp_path, VEC_SRC)[0]
    b = run(tmp_path, VEC_SRC, vectorize=True)[0]
    for k in ("u", "w", "bx", "acc", "r"):
        assert np.array_equal(a[k], b[k]), k
    # (the private scalar d holds an array after a vectorised loop: like the PRIVATE variables of the reference's OpenMP
    # loops it is undefined afterwards, and no program reads one)
    assert a["total"] == b["total"] and a["total"] != 0.0
    assert a["w"][7, 1] == sum(a["u"][1:8, 1].tolist()) or abs(a["w"][7, 1] - a["u"][1:8, 1].sum()) < 1e-15
    p = tmp_path / "t.f90"
    tr = F.Translator(vectorize=True)
    src = tr.translate(F.logical_lines(str(p)))
    assert src.count("np.arange(") == 7            # the u, w and r nests (2 each), the bx loop (1); not the chain loop
    assert "for m in _range(1, 2):" in src
    assert src.count("_accumulate(") == 2


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "seismic_CPML_3D_isotropic_MPI_OpenMP.f90")),
                    reason="the reference tree is not on this machine")
def test_live_vectorised_run_equals_the_statement_by_statement_run_on_every_array():
    ov = {"NX": "12", "NY": "14", "NZ": "8", "NPROC": "2", "NSTEP": "10", "NPOINTS_PML": "2", "ydeb": "40.d0", "yfin": "30.d0"}
    path = os.path.join(REF, "seismic_CPML_3D_isotropic_MPI_OpenMP.f90")
    a = F.run_program(path, ov, nproc=2)
    b = F.run_program(path, ov, nproc=2, vectorize=True)
    for r in range(2):
        for k in a[r]["_bounds"]:
            assert np.array_equal(a[r][k], b[r][k]), (r, k)
    assert np.abs(a[0]["sisvx"]).max() > 0
