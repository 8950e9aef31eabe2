// Device helpers shared by the TMA-staged 3-D isotropic kernels (kernels_3d_tma.cu: every warp computes, thread 0
// issues the loads; kernels_3d_ws.cu: a dedicated producer warp issues them): mbarrier / TMA wrappers, the tile
// geometry, the shared-memory ring position, and the per-point updates of the reference loop nests
//   stress_point    sigmaxx/yy/zz (:836-863), sigmaxy (:877-894), sigmaxz/yz (:908-943)
//   velocity_point  vx/vy (:976-1017), vz (:1031-1052), source (:1055-1083), Dirichlet (:1087-1121), energy (:1131-1177)
// (line numbers: seismic_CPML_3D_isotropic_MPI_OpenMP.f90), written once so that both kernel families produce the
// same bits.
#pragma once
#include <cuda.h>

#include "cpml_internal.h"

namespace cpml {

namespace {

constexpr int kBarBytes = 128;   // room for up to 16 mbarriers ahead of the stages

__host__ __device__ constexpr int round128(int v) { return (v + 127) / 128 * 128; }

// HX: cells a halo box starts before the tile when the operator reaches back -- 16 bytes (the x start of a tiled load
// must be 16-byte aligned): 2 doubles or 4 floats; the box is TX + HX wide either way.
template <typename T, int TX, int TY>
struct TileGeomT {
    static constexpr int ES = (int)sizeof(T);
    static constexpr int HX = 16 / ES;
    static constexpr int W = TX + HX;                      // halo box row length
    static constexpr int HROWS = TY + 1;                   // halo box rows
    static constexpr int HALO_BOX_BYTES = W * HROWS * ES;
    static constexpr int PLAIN_BOX_BYTES = TX * TY * ES;
    static constexpr int HALO_BYTES = round128(HALO_BOX_BYTES);
    static constexpr int PLAIN_BYTES = round128(PLAIN_BOX_BYTES);
};
template <int TX, int TY>
using TileGeom = TileGeomT<double, TX, TY>;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a TMA that never lands (bad descriptor) traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity))
        if (++spins > (1u << 24)) __trap();
}
// Work items of the persistent kernels are handed out through a queue in global memory (queue[0]: items claimed
// beyond the first one of every CTA, queue[1]: CTAs that have finished): an SM that gets more of the HBM bandwidth
// takes more items, so the CTAs finish within one item of each other instead of idling behind the slowest static
// share (ncu, static shares: the SMs were idle 6-19 % of a kernel).  Items are claimed in increasing order, which
// keeps "boundary chunks first".  queue == nullptr: static shares (item += gridDim.x).
__device__ __forceinline__ int claim_item(unsigned int *queue, int static_next)
{
    return queue ? (int)(gridDim.x + atomicAdd(queue, 1u)) : static_next;
}
// One thread per CTA, after the CTA's last claim: the last CTA to retire re-arms the queue for the next launch
// (kernels sharing a queue never overlap).
__device__ __forceinline__ void retire_queue(unsigned int *queue)
{
    if (!queue) return;
    const unsigned int done = atomicAdd(queue + 1, 1u);
    if (done == gridDim.x - 1) { queue[0] = 0u; queue[1] = 0u; }
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int x, int y, int z, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"((unsigned long long)map), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int x, int y, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"((unsigned long long)map), "r"(x), "r"(y), "r"(bar) : "memory");
}

// 1-D bulk copy (TMA engine, no descriptor): contiguous, 16-byte aligned, size % 16 == 0
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <typename T>
__device__ __forceinline__ void st_stream(T *p, T v) { __stcs(p, v); }

// memory_x = b * memory_x + a * value ; value = value / K + memory_x   (e.g. :845-851)
template <bool KUNIT, typename T>
__device__ __forceinline__ T cpml_apply(T *__restrict__ mem, long long q, T m,
                                        T b, T a, T K, T value)
{
    m = b * m + a * value;
    mem[q] = m;
    return KUNIT ? value + m : value / K + m;
}

// The same recursion step, branch-free across the lanes of a warp: lanes outside the shell carry
// a = 0 and m = 0, so their m' is 0 and the derivative comes back as value + 0 (exact; only the
// sign of a zero can differ); only shell lanes store.  This keeps the C-PML code of a warp that
// holds a few shell lanes straight-line instead of nine divergent blocks per point.
template <bool KUNIT, typename T>
__device__ __forceinline__ T cpml_step(T *__restrict__ mem, long long q, bool store, T m,
                                       T b, T a, T K, T value)
{
    m = b * m + a * value;
    if (store) mem[q] = m;
    return KUNIT ? value + m : value / K + m;
}

__device__ __forceinline__ int shell_index(int i, int lo, int hi) { return i <= lo ? i - 1 : lo + (i - hi); }

template <int NT>
__device__ __forceinline__ void block_sum2(double &a, double &b, double *red /* 2*NT/32 + 2 */)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, o);
        b += __shfl_down_sync(0xffffffffu, b, o);
    }
    constexpr int NW = (NT + 31) / 32;
    const int t = threadIdx.y * blockDim.x + threadIdx.x;
    const int w = t >> 5, l = t & 31;
    if (l == 0) { red[w] = a; red[NW + w] = b; }
    __syncthreads();
    if (w == 0) {
        a = (l < NW) ? red[l] : 0.0;
        b = (l < NW) ? red[NW + l] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_down_sync(0xffffffffu, a, o);
            b += __shfl_down_sync(0xffffffffu, b, o);
        }
    }
    __syncthreads();     // red[] is reused by the next work item
}

}  // namespace

// Threads of a CTA: one per pair of points, rounded up to whole warps.  Tiles whose pair count is not
// a multiple of 32 (104 x 7: 364 pairs on 384 threads = 12 warps, three per SM sub-partition, which lifts
// the register cap from 128 to 168) are launched one-dimensional with idle threads at the end.
__host__ __device__ constexpr int tile_pairs(int tx, int ty) { return tx / 2 * ty; }
__host__ __device__ constexpr int tile_threads(int tx, int ty) { return (tile_pairs(tx, ty) + 31) / 32 * 32; }

// Position in a ring of shared-memory stages: stage index and the mbarrier phase parity of
// its current use.  Every load goes through the stages cyclically, so the parity flips
// exactly when the index wraps.
struct RingPos {
    uint32_t s, par;
    __device__ __forceinline__ void advance(uint32_t depth) { if (++s == depth) { s = 0; par ^= 1u; } }
};

// ------------------------------------------------------------------------------- stress
// ring N (planes n and n+1 are needed):  vx (halo box at (0,0)), vy (halo, (-2,-1))
// ring C (plane n only):                 vz (halo, (-2,0)), sxx syy szz sxy sxz syz (plain boxes)
// maps: 0 vx 1 vy 2 vz 3..8 sigma
//
// One thread updates TWO x-adjacent points (16-byte shared-memory loads and global stores,
// half the address / predicate / barrier instructions per point, two independent dependency
// chains): with one point per thread the kernels were bound by instruction latency, not by
// HBM (profiles/r01_v4_*).  The point update itself is written once, per point.
template <typename T>
struct StressValsT { T sxx, syy, szz, sxy, sxz, syz; };   // in: old values, out: new values
using StressVals = StressValsT<double>;

// Coefficient table of the tile's columns in shared memory: rows a, b, K, a_half, b_half, K_half
// (CXW doubles each).  ux / uy / uz: the warp holds x- / y-shell lanes, the plane lies in the z
// shell (warp-uniform); in_* : this point does (stores its memory variables).
template <bool PML, bool KUNIT, int CXW, typename T>
__device__ __forceinline__ void stress_point(
    const Params3DT<T> &p, const T *__restrict__ Cx, const int c, const int j, const int kg,
    const bool do_n, const bool do_xy, const bool do_xz, const bool do_yz,
    const bool ux, const bool uy, const bool uz, const bool in_x, const bool in_y, const bool in_z,
    const long long qx, const long long qy, const long long qz, const T (&mv)[9],
    const T vx_c, const T vx_ip, const T vx_jp, const T vx_n,
    const T vy_c, const T vy_im, const T vy_jm, const T vy_n,
    const T vz_c, const T vz_im, const T vz_jp, const T vz_m, StressValsT<T> &s)
{
    const T odx = p.odx, ody = p.ody, odz = p.odz;
    const T dt_l = p.dt_lambda, dt_m = p.dt_mu, dt_l2m = p.dt_lambdaplus2mu;
    // ---- sigmaxx, sigmayy, sigmazz  (:836-863)
    if (do_n && kg >= 2) {                                  // k2begin, :792-793
        T value_dvx_dx = (vx_ip - vx_c) * odx;
        T value_dvy_dy = (vy_c - vy_jm) * ody;
        T value_dvz_dz = (vz_c - vz_m) * odz;
        if (PML) {
            if (ux) value_dvx_dx = cpml_step<KUNIT>(p.mx[0], qx, in_x, mv[0], Cx[4 * CXW + c], Cx[3 * CXW + c], KUNIT ? T(1) : Cx[5 * CXW + c], value_dvx_dx);
            if (uy) value_dvy_dy = cpml_step<KUNIT>(p.my[0], qy, in_y, mv[3], p.cy.b[j], p.cy.a[j], KUNIT ? T(1) : p.cy.K[j], value_dvy_dy);
            if (uz) value_dvz_dz = cpml_step<KUNIT>(p.mz[0], qz, in_z, mv[6], p.cz.b[kg], p.cz.a[kg], KUNIT ? T(1) : p.cz.K[kg], value_dvz_dz);
        }
        s.sxx = dt_l2m * value_dvx_dx + dt_l * (value_dvy_dy + value_dvz_dz) + s.sxx;
        s.syy = dt_l * (value_dvx_dx + value_dvz_dz) + dt_l2m * value_dvy_dy + s.syy;
        s.szz = dt_l * (value_dvx_dx + value_dvy_dy) + dt_l2m * value_dvz_dz + s.szz;
    }
    // ---- sigmaxy  (:877-894)
    if (do_xy) {
        T value_dvy_dx = (vy_c - vy_im) * odx;
        T value_dvx_dy = (vx_jp - vx_c) * ody;
        if (PML) {
            if (ux) value_dvy_dx = cpml_step<KUNIT>(p.mx[1], qx, in_x, mv[1], Cx[1 * CXW + c], Cx[0 * CXW + c], KUNIT ? T(1) : Cx[2 * CXW + c], value_dvy_dx);
            if (uy) value_dvx_dy = cpml_step<KUNIT>(p.my[1], qy, in_y, mv[4], p.cy.b_half[j], p.cy.a_half[j], KUNIT ? T(1) : p.cy.K_half[j], value_dvx_dy);
        }
        s.sxy = dt_m * (value_dvy_dx + value_dvx_dy) + s.sxy;
    }
    // ---- sigmaxz, sigmayz  (:908-943)
    if (kg <= p.nz - 1) {                                   // kminus1end, :795-796
        if (do_xz) {
            T value_dvz_dx = (vz_c - vz_im) * odx;
            T value_dvx_dz = (vx_n - vx_c) * odz;
            if (PML) {
                if (ux) value_dvz_dx = cpml_step<KUNIT>(p.mx[2], qx, in_x, mv[2], Cx[1 * CXW + c], Cx[0 * CXW + c], KUNIT ? T(1) : Cx[2 * CXW + c], value_dvz_dx);
                if (uz) value_dvx_dz = cpml_step<KUNIT>(p.mz[1], qz, in_z, mv[7], p.cz.b_half[kg], p.cz.a_half[kg], KUNIT ? T(1) : p.cz.K_half[kg], value_dvx_dz);
            }
            s.sxz = dt_m * (value_dvz_dx + value_dvx_dz) + s.sxz;
        }
        if (do_yz) {
            T value_dvz_dy = (vz_jp - vz_c) * ody;
            T value_dvy_dz = (vy_n - vy_c) * odz;
            if (PML) {
                if (uy) value_dvz_dy = cpml_step<KUNIT>(p.my[2], qy, in_y, mv[5], p.cy.b_half[j], p.cy.a_half[j], KUNIT ? T(1) : p.cy.K_half[j], value_dvz_dy);
                if (uz) value_dvy_dz = cpml_step<KUNIT>(p.mz[2], qz, in_z, mv[8], p.cz.b_half[kg], p.cz.a_half[kg], KUNIT ? T(1) : p.cz.K_half[kg], value_dvy_dz);
            }
            s.syz = dt_m * (value_dvz_dy + value_dvy_dz) + s.syz;
        }
    }
}

// Fills the column coefficient table of a tile (columns beyond NX: a = b = 0, K = 1).
template <int TX, int NT, typename T>
__device__ __forceinline__ void fill_cx(const Params3DT<T> &p, T *Cx, int i0, int tid)
{
    for (int e = tid; e < 6 * TX; e += NT) {
        const int f = e / TX, i = i0 + (e - f * TX);
        const T *src = f == 0 ? p.cx.a : f == 1 ? p.cx.b : f == 2 ? p.cx.K : f == 3 ? p.cx.a_half : f == 4 ? p.cx.b_half : p.cx.K_half;
        Cx[e] = (i <= p.nx) ? src[i] : ((f == 2 || f == 5) ? T(1) : T(0));
    }
}

__device__ __forceinline__ double2 lds2(const double *t, int e) { return *reinterpret_cast<const double2 *>(t + e); }
__device__ __forceinline__ float2 lds2(const float *t, int e) { return *reinterpret_cast<const float2 *>(t + e); }
__device__ __forceinline__ void st_stream2(double *p, double a, double b) { __stcs(reinterpret_cast<double2 *>(p), make_double2(a, b)); }
__device__ __forceinline__ void st_stream2(float *p, float a, float b) { __stcs(reinterpret_cast<float2 *>(p), make_float2(a, b)); }
__device__ __forceinline__ double2 ldg2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ float2 ldg2(const float *p) { return *reinterpret_cast<const float2 *>(p); }

// Loads the nine C-PML memory variables of one point (group g0 = 0: stress kernel, 3: velocity).
template <typename T>
__device__ __forceinline__ void load_memvars(const Params3DT<T> &p, int g0, bool in_x, bool in_y, bool in_z,
                                             long long qx, long long qy, long long qz, T (&mv)[9])
{
    if (in_x) { mv[0] = p.mx[g0 + 0][qx]; mv[1] = p.mx[g0 + 1][qx]; mv[2] = p.mx[g0 + 2][qx]; }
    if (in_y) { mv[3] = p.my[g0 + 0][qy]; mv[4] = p.my[g0 + 1][qy]; mv[5] = p.my[g0 + 2][qy]; }
    if (in_z) { mv[6] = p.mz[g0 + 0][qz]; mv[7] = p.mz[g0 + 1][qz]; mv[8] = p.mz[g0 + 2][qz]; }
}

// ----------------------------------------------------------------------------- velocity
// ring N (planes n and n+1): szz (plain box)
// ring C (plane n only):     sxx (halo box at (-2,0)), syy (halo, (0,0)), sxy (halo, (0,-1)),
//                            sxz (halo, (0,0)), syz (halo, (0,-1)), vx vy vz (plain boxes)
// maps: 0 sxx 1 syy 2 sxy 3 sxz 4 syz 5 szz 6 vx 7 vy 8 vz
template <typename T>
struct VelValsT { T vx, vy, vz; };                           // in: old values, out: new values
using VelVals = VelValsT<double>;

template <bool PML, bool KUNIT, int CXW, typename T>
__device__ __forceinline__ void velocity_point(
    const Params3DT<T> &p, const T *__restrict__ Cx, const int c, const int j, const int k, const int kg,
    const bool do_vx, const bool do_vy, const bool do_vz, const bool edge_ij, const bool ebox_ij, const bool src_ij,
    const bool ux, const bool uy, const bool uz, const bool in_x, const bool in_y, const bool in_z,
    const long long qx, const long long qy, const long long qz, const T (&mv)[9],
    const T sxx_c, const T sxx_im, const T syy_c, const T syy_jp,
    const T sxy_c, const T sxy_jm, const T sxy_ip, const T sxz_c, const T sxz_ip,
    const T sxz_m, const T syz_c, const T syz_jm, const T syz_m, const T szz_c,
    const T szz_n, VelValsT<T> &v, double &ekin, double &epot)
{
    const T odx = p.odx, ody = p.ody, odz = p.odz, dt_r = p.dt_over_rho;
    T vx = v.vx, vy = v.vy, vz = v.vz;
    if (kg >= 2) {                                           // k2begin
        if (do_vx) {                                         // :976-996
            T value_dsigmaxx_dx = (sxx_c - sxx_im) * odx;
            T value_dsigmaxy_dy = (sxy_c - sxy_jm) * ody;
            T value_dsigmaxz_dz = (sxz_c - sxz_m) * odz;
            if (PML) {
                if (ux) value_dsigmaxx_dx = cpml_step<KUNIT>(p.mx[3], qx, in_x, mv[0], Cx[1 * CXW + c], Cx[0 * CXW + c], KUNIT ? T(1) : Cx[2 * CXW + c], value_dsigmaxx_dx);
                if (uy) value_dsigmaxy_dy = cpml_step<KUNIT>(p.my[3], qy, in_y, mv[3], p.cy.b[j], p.cy.a[j], KUNIT ? T(1) : p.cy.K[j], value_dsigmaxy_dy);
                if (uz) value_dsigmaxz_dz = cpml_step<KUNIT>(p.mz[3], qz, in_z, mv[6], p.cz.b[kg], p.cz.a[kg], KUNIT ? T(1) : p.cz.K[kg], value_dsigmaxz_dz);
            }
            vx = dt_r * (value_dsigmaxx_dx + value_dsigmaxy_dy + value_dsigmaxz_dz) + vx;
        }
        if (do_vy) {                                         // :998-1016
            T value_dsigmaxy_dx = (sxy_ip - sxy_c) * odx;
            T value_dsigmayy_dy = (syy_jp - syy_c) * ody;
            T value_dsigmayz_dz = (syz_c - syz_m) * odz;
            if (PML) {
                if (ux) value_dsigmaxy_dx = cpml_step<KUNIT>(p.mx[4], qx, in_x, mv[1], Cx[4 * CXW + c], Cx[3 * CXW + c], KUNIT ? T(1) : Cx[5 * CXW + c], value_dsigmaxy_dx);
                if (uy) value_dsigmayy_dy = cpml_step<KUNIT>(p.my[4], qy, in_y, mv[4], p.cy.b_half[j], p.cy.a_half[j], KUNIT ? T(1) : p.cy.K_half[j], value_dsigmayy_dy);
                if (uz) value_dsigmayz_dz = cpml_step<KUNIT>(p.mz[4], qz, in_z, mv[7], p.cz.b[kg], p.cz.a[kg], KUNIT ? T(1) : p.cz.K[kg], value_dsigmayz_dz);
            }
            vy = dt_r * (value_dsigmaxy_dx + value_dsigmayy_dy + value_dsigmayz_dz) + vy;
        }
    }
    if (do_vz && kg <= p.nz - 1) {                           // kminus1end, :1031-1052
        T value_dsigmaxz_dx = (sxz_ip - sxz_c) * odx;
        T value_dsigmayz_dy = (syz_c - syz_jm) * ody;
        T value_dsigmazz_dz = (szz_n - szz_c) * odz;
        if (PML) {
            if (ux) value_dsigmaxz_dx = cpml_step<KUNIT>(p.mx[5], qx, in_x, mv[2], Cx[4 * CXW + c], Cx[3 * CXW + c], KUNIT ? T(1) : Cx[5 * CXW + c], value_dsigmaxz_dx);
            if (uy) value_dsigmayz_dy = cpml_step<KUNIT>(p.my[5], qy, in_y, mv[5], p.cy.b[j], p.cy.a[j], KUNIT ? T(1) : p.cy.K[j], value_dsigmayz_dy);
            if (uz) value_dsigmazz_dz = cpml_step<KUNIT>(p.mz[5], qz, in_z, mv[8], p.cz.b_half[kg], p.cz.a_half[kg], KUNIT ? T(1) : p.cz.K_half[kg], value_dsigmazz_dz);
        }
        vz = dt_r * (value_dsigmaxz_dx + value_dsigmayz_dy + value_dsigmazz_dz) + vz;
    }

    // source, :1080-1081 (after the update of step it, before Dirichlet; quirk B10)
    if (src_ij && k == p.ksrc) {
        vx = vx + (T)p.src_x[p.it - 1];
        vy = vy + (T)p.src_y[p.it - 1];
    }
    // Dirichlet on the six faces, :1087-1121
    if (edge_ij || kg == 1 || kg == p.nz) { vx = 0.0; vy = 0.0; vz = 0.0; }
    v.vx = vx; v.vy = vy; v.vz = vz;

    // energy over the PML-free box, :1131-1177.  The trace is a reduction whose order differs
    // from the reference anyway (quirk B11), so this part uses reciprocals and fused
    // multiply-adds (half the FP64 instructions); it agrees with the oracle to ~1e-14.
    if (ebox_ij && kg >= p.npml + 1 && kg <= p.nz - p.npml) {
        const double lam = p.lambda, c2lm = p.c2lm, inv_den = p.inv_den, inv_mu = p.inv_mu;
        // (accumulated in double in the single-precision build too: the conversions are no-ops for T = double)
        const double dvx = vx, dvy = vy, dvz = vz;
        const double esxx = sxx_c, esyy = syy_c, eszz = szz_c, esxy = sxy_c, esxz = sxz_c, esyz = syz_c;
        ekin = __fma_rn(p.half_rho, __fma_rn(dvz, dvz, __fma_rn(dvy, dvy, dvx * dvx)), ekin);
        const double epsilon_xx = __fma_rn(-lam, eszz, __fma_rn(-lam, esyy, c2lm * esxx)) * inv_den;
        const double epsilon_yy = __fma_rn(-lam, eszz, __fma_rn(-lam, esxx, c2lm * esyy)) * inv_den;
        // quirk B2 (:1169-1172): the reference adds epsilon_yy*sigmayy twice and never
        // epsilon_zz*sigmazz
        double third;
        if (p.energy_bug_compat) third = epsilon_yy * esyy;
        else third = __fma_rn(-lam, esyy, __fma_rn(-lam, esxx, c2lm * eszz)) * inv_den * eszz;
        // 2 * epsilon_ij * sigma_ij = sigma_ij^2 / mu
        double acc = __fma_rn(epsilon_xx, esxx, __fma_rn(epsilon_yy, esyy, third));
        acc = __fma_rn(esxy * inv_mu, esxy, acc);
        acc = __fma_rn(esxz * inv_mu, esxz, acc);
        acc = __fma_rn(esyz * inv_mu, esyz, acc);
        epot = __fma_rn(0.5, acc, epot);
    }
}


}  // namespace cpml
