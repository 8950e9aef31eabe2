// Device helpers shared by the 3-D viscoelastic kernels (kernels_3d_visco.cu: register-marching stress and velocity
// kernels; kernels_3d_visco_ws.cu: the TMA-staged velocity kernel with a producer warp): streaming accesses, the L2
// prefetch, the in-kernel halo stores, the C-PML recursion and the fourth-order difference of
// seismic_CPML_3D_viscoelastic_MPI.f90 (:989-999), written once so that both kernels produce the same bits.
#pragma once
#include "cpml_internal.h"

namespace cpml {

__device__ __forceinline__ double vld(const double *p) { return __ldcs(p); }
__device__ __forceinline__ void vst(double *p, double v) { __stcs(p, v); }
__device__ __forceinline__ double2 vld2(const double2 *p) { return __ldcs(p); }
__device__ __forceinline__ void vst2(double2 *p, double2 v) { __stcs(p, v); }

// L2 prefetch of the words a thread will stream a little later.  The stress kernel is bound by memory
// latency, not bandwidth (13.7 long-scoreboard stall cycles per issued instruction at 16 warps per SM,
// profiles/r01_v7_ncu_cfg5d.txt): every nest of a plane waits a full DRAM round trip for the
// read-modify-write words it loads at its head.  A prefetch costs no register and turns those round
// trips into L2 hits; the LSU merges the 32 addresses of a warp into the 2-4 lines they touch.
// Measured on B200 (profiles/r01_v8_vpf_sweep.txt), 1024 x 1024 x 128 slab, stress kernel: no prefetch
// 14.97 ms (58 % of measured HBM); everything plane k+1 streams, at the head of plane k: 11.81 ms; the
// same two planes ahead: 14.87 ms (the prefetched lines of 296 resident blocks evict each other);
// staggered by half a plane (default): 11.16 ms (78 %); staggered by one nest: 11.87 ms; one bulk
// prefetch per row by lane 0 (cp.async.bulk.prefetch.L2): the same as the per-lane form.  The velocity
// kernel (3 streamed words, all loads already hoisted to the head of the plane) gains at most 2 % from
// any of four prefetch placements, less than the code costs it when switched off, and has none.
__device__ __forceinline__ void pf_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Fourth-order halo exchange by direct stores into the neighbour slabs (replaces the MPI_SENDRECV calls of :962-975
// and :1229-1242).  A field that travels "left" in the reference (vx, vy, sigmazz: planes 1:2 -> the lower slab's
// NZ_LOCAL+1:NZ_LOCAL+2) ALSO sends its plane NZ_LOCAL to the upper slab's plane 0, and a field that travels "right"
// (vz, sigmaxz, sigmayz: planes NZ_LOCAL-1:NZ_LOCAL -> the upper slab's -1:0) ALSO sends its plane 1 to the lower
// slab's NZ_LOCAL+1: the reference never sends those two planes although its stencils read them (quirk B6); GPU slabs
// always exchange the complete halo and the kernels decide per plane which taps read zero (nzl_e).
// rel = offset of the point inside its plane, pl = plane pitch.
template <bool LEFT>
__device__ __forceinline__ void peer_put(double *lo, double *hi, int k, int nzl, int rel, int pl, double v)
{
    if (LEFT) {
        if (lo && k <= 2) __stcs(lo + (nzl + k) * pl + rel, v);
        if (hi && k == nzl) __stcs(hi + rel, v);
    } else {
        if (hi && k >= nzl - 1) __stcs(hi + (k - nzl) * pl + rel, v);
        if (lo && k == 1) __stcs(lo + (nzl + 1) * pl + rel, v);
    }
}

// memory_x = b * memory_x + a * value ; value / K + memory_x   (e.g. :993-999); rK = RN(1/K)
__device__ __forceinline__ double vcpml(double *__restrict__ mem, int q, double b, double a, double K, double rK, double value)
{
    double m = mem[q];
    m = b * m + a * value;
    mem[q] = m;
    return div_exact(value, K, rK) + m;
}

__device__ __forceinline__ int vshell(int i, int lo, int hi) { return i <= lo ? i - 1 : lo + (i - hi); }

// (27 a - 27 b - c + d) * ONE_OVER_DELTA / 24   (:989-991): the numerator; the three (or two) differences of
// a nest are then divided by 24 behind one shared range test, with the single correction that is enough
// for a divisor with a small odd part (div_small3 / div_small2, cpml_internal.h)
__device__ __forceinline__ double d4n(double a, double b, double c, double d, double od)
{
    return (27.0 * a - 27.0 * b - c + d) * od;
}
#define DIV24_3(x, y, z) div_small3(x, y, z, 24.0, 1.0 / 24.0)
#define DIV24_2(x, y) div_small2(x, y, 24.0, 1.0 / 24.0)

}  // namespace cpml
