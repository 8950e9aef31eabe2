"""Work decomposition of the persistent 3-D isotropic kernels (kernels_3d_ws.cu), checked on the CPU.

The library exports two test hooks that are the DEVICE functions compiled for the host (`decode_item`: work item ->
x tile, y tile, planes kb..ke, with the finer items at the end of the list; `map_lane_packed`: the per-item thread map
of the single-precision kernels).  The properties the kernels rely on:
  * every (tile, plane) belongs to exactly one item, whatever the chunking and the split of the tail;
  * the first 2 x tiles items are the two boundary chunks, whole (the in-kernel slab ordering counts on that:
    SlabSync.n_boundary = tiles per plane, boundary items are claimed first);
  * the thread map is one-to-one onto the pairs of a tile for ANY shell bounds, and on the reference's default grid it
    packs all x-shell lanes into 3 of the 13 consumer warps.
(The reference has no counterpart: its loops are `do k / do j / do i` over the whole slab,
seismic_CPML_3D_isotropic_MPI_OpenMP.f90:836-1052.)
"""
import ctypes as C

import numpy as np
import pytest

from seismic_cpml_b200 import lib as L


@pytest.fixture(scope="module")
def hooks():
    lib = C.CDLL(L.LIB_PATH)
    lib.cpml_debug_work_item.argtypes = [C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.POINTER(C.c_int32)]
    lib.cpml_debug_work_item.restype = C.c_int32
    lib.cpml_debug_lane.argtypes = [C.c_int32] * 7 + [C.POINTER(C.c_int32)]
    lib.cpml_debug_lane.restype = C.c_int32
    return lib


def host_decomposition(nzl, tiles, nzc_req, resident, split):
    """The host side of the decomposition (cpml_api.cu: build_tile), restated: chunk length, chunk count, which coarse
    items are split."""
    nzc = max(1, min(nzc_req, nzl))
    kchunk = -(-nzl // nzc)
    nzc = -(-nzl // kchunk)
    coarse = tiles * nzc
    fine_from, sp = coarse, 1
    if nzc >= 3 and 2 <= split <= 8 and (split - 1) * (-(-kchunk // split)) < kchunk:
        sp = split
        fine_from = coarse - min(coarse - 2 * tiles, resident * 3 // 2)
    return kchunk, nzc, fine_from, sp, fine_from + (coarse - fine_from) * sp


@pytest.mark.parametrize("nzl", [640, 128, 42, 24, 17, 9])
@pytest.mark.parametrize("ntx,nty", [(1, 1), (1, 6), (3, 5), (1, 81)])
@pytest.mark.parametrize("nzc_req", [1, 2, 3, 5, 20, 40])
@pytest.mark.parametrize("split", [1, 2, 3, 4])
def test_items_cover_every_tile_plane_once_and_boundary_chunks_come_first(hooks, nzl, ntx, nty, nzc_req, split):
    tiles = ntx * nty
    kchunk, nzc, fine_from, sp, nitems = host_decomposition(nzl, tiles, nzc_req, 148, split)
    tile6 = (C.c_int32 * 6)(ntx, nty, kchunk, nzc, fine_from, sp)
    out = (C.c_int32 * 4)()
    seen = np.zeros((ntx, nty, nzl + 1), dtype=np.int32)
    for item in range(nitems):
        assert hooks.cpml_debug_work_item(tile6, nzl, item, out) == 0
        tix, tiy, kb, ke = out[0], out[1], out[2], out[3]
        assert 0 <= tix < ntx and 0 <= tiy < nty and 1 <= kb <= ke <= nzl, (item, tix, tiy, kb, ke)
        seen[tix, tiy, kb:ke + 1] += 1
        if item < min(2 * tiles, nitems) and nzc >= 2:
            # the two boundary chunks, whole: the bottom one first, then the top one
            if item < tiles:
                assert kb == 1 and ke == min(nzl, kchunk)
            else:
                assert ke == nzl and kb == 1 + (nzc - 1) * kchunk
        elif nzc >= 3:
            assert kb > 1 and ke < nzl          # no other item touches a halo plane
    assert np.all(seen[:, :, 1:] == 1)
    if sp > 1:
        assert nitems > tiles * nzc


@pytest.mark.parametrize("tx,ty", [(64, 8), (104, 8), (128, 7)])
@pytest.mark.parametrize("nx,xlo,xhi", [(101, 10, 91), (101, 0, 102), (101, 101, 102), (37, 6, 32), (130, 8, 123), (300, 6, 295),
                                        (1024, 10, 1015), (20, 10, 11)])
def test_packed_thread_map_is_one_to_one(hooks, tx, ty, nx, xlo, xhi):
    out = (C.c_int32 * 3)()
    hp = tx // 2
    nthreads = (hp * ty + 31) // 32 * 32
    for tix in range((nx + tx - 1) // tx):
        i0 = 1 + tix * tx
        seen = set()
        for tid in range(nthreads):
            assert hooks.cpml_debug_lane(tx, ty, tid, i0, xlo, xhi, nx, out) == 0
            assert 0 <= out[0] < hp and 0 <= out[1] < ty
            if out[2]:
                assert (out[0], out[1]) not in seen
                seen.add((out[0], out[1]))
        assert len(seen) == hp * ty


def test_packed_thread_map_on_the_default_grid(hooks):
    """NX = 101, shell [1..10] u [91..101], one 104 x 8 tile per row: ten warps without a single x-shell lane, three
    warps that hold nothing else (shell and pad pairs)."""
    out = (C.c_int32 * 3)()
    shell_lanes = np.zeros(13, dtype=int)
    for tid in range(416):
        hooks.cpml_debug_lane(104, 8, tid, 1, 10, 91, 101, out)
        i = 1 + 2 * out[0]
        in_shell_or_pad = i <= 10 or i + 1 >= 91
        shell_lanes[tid // 32] += int(in_shell_or_pad)
    assert list(shell_lanes) == [0] * 10 + [32] * 3
