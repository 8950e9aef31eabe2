// Internal declarations of libcpml_b200 (not part of the C ABI).
//
// Data layout in HBM (DESIGN.md "Data layout"):
//   * every wavefield is one padded array, x fastest, row pitch a multiple of 16
//     doubles (128 B) so that a warp row starts on a cache line;
//   * 3-D fields carry the two z halo planes of the reference (k = 0 and
//     k = NZ_LOCAL+1, 3D-iso :273); 2-D fields carry a 2-cell zero ghost ring (the
//     fourth-order program's (0:NX+1,0:NY+1) arrays, 2D-4th :205, widened to 2 so
//     that discarded lanes never read out of bounds);
//   * the 18 (3-D) / 8 (2-D) C-PML memory variables exist only inside the thin
//     shells where their damping profile is non-trivial: x-shell arrays
//     [k][j][sx], y-shell arrays [k][sy][i], z-shell arrays [sz][j][i].
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/cpml_b200.h"

namespace cpml {

// Index set of one axis where C-PML coefficients are non-trivial: [1..lo] U [hi..n].
struct Shell {
    int lo;   // last index of the low shell (0: none)
    int hi;   // first index of the high shell (n+1: none)
    int n;
    int size() const { return lo + (n - hi + 1); }
};

// 1-based device views of the six coefficient arrays of one axis, plus the correctly rounded
// reciprocals of K and K_half (for div_exact).
template <typename T>
struct AxisCoefT {
    const T *a, *b, *K, *a_half, *b_half, *K_half;
    const T *rK, *rK_half;
};
using AxisCoef = AxisCoefT<double>;

#ifdef __CUDACC__
// a / c, correctly rounded, for a divisor whose correctly rounded reciprocal y = RN(1/c) is known
// (a constant, or a profile value with a host-computed reciprocal).  q = a*y followed by two
// residual corrections r = fma(-c, q, a), q += r*y: after the first correction q is a faithful
// rounding of a/c, and then Markstein's theorem (IBM J. Res. Dev. 34, 1990; Muller et al., Handbook
// of Floating-Point Arithmetic, "division with a correctly rounded reciprocal") makes the second one
// the correctly rounded quotient, provided nothing underflows or overflows on the way and the
// significand of c is not all ones (checked on the host).  Outside a wide exponent window of the
// dividend the generic IEEE division runs instead, so the result equals a / c for every input
// (for a zero dividend possibly with the other sign of zero); zero dividends stay on the short path.  Five dependent FP64 operations instead of
// the ~60-instruction generic sequence (whose slow path is also taken for every zero dividend,
// i.e. on the whole quiescent part of the grid).
// The generic division, kept out of line so that the 20-odd call sites of a kernel share one copy
// (inlined, the fallbacks made up a third of the viscoelastic stress kernel's code).
static __device__ __noinline__ double div_generic(double a, double c) { return a / c; }

// the five operations alone (exact inside the window tested by div_slow)
__device__ __forceinline__ double div_fast(double a, double c, double y)
{
    const double q0 = a * y;
    double r = __fma_rn(-c, q0, a);
    double q = __fma_rn(r, y, q0);
    r = __fma_rn(-c, q, a);
    return __fma_rn(r, y, q);
}

// one unsigned compare on the exponent field (sign shifted out): inside 2^-822 < |a| < 2^817 the
// five operations above are exact; outside, a zero keeps their result (+-0, equal in value to
// the quotient) and everything else takes the generic division
// (one predicate, no short-circuit: zero dividends -- the quiescent part of the grid -- fall
// straight through like ordinary ones)
__device__ __forceinline__ bool div_slow(double a)
{
    const unsigned u = ((unsigned)__double2hiint(a) << 1) - (0x0c9u << 21);
    return (u >= ((0x730u - 0x0c9u) << 21)) & (a != 0.0);
}

__device__ __forceinline__ double div_exact(double a, double c, double y)
{
    double q = div_fast(a, c, y);
    if (div_slow(a)) q = div_generic(a, c);
    return q;
}

// Divisors with a SMALL ODD PART (24 = 3 * 8, 3): one correction is enough, provably.  With c = d * 2^k,
// d odd, the quotient a / c is never closer to a rounding midpoint than 1 / (2 d) ulp: in units of half an
// ulp, a / c minus the odd integer of a midpoint is a non-zero integer over d (non-zero because a midpoint
// times c would need more than 53 bits).  q0 = RN(a y) is within 1.5 ulp of a / c, so r = a - c q0 is a small
// multiple of a common unit and exact in the FMA, and q0 + r y differs from a / c by |r / c| 2^-53 < 2^-52 ulp --
// far inside the 1 / (2 d) ulp margin, so RN(q0 + r y) is the correctly rounded quotient.
__device__ __forceinline__ double div_fast_small(double a, double c, double y)
{
    const double q0 = a * y;
    const double r = __fma_rn(-c, q0, a);
    return __fma_rn(r, y, q0);
}
__device__ __forceinline__ double div_small(double a, double c, double y)
{
    double q = div_fast_small(a, c, y);
    if (div_slow(a)) q = div_generic(a, c);
    return q;
}
__device__ __forceinline__ void div_small2(double &a0, double &a1, double c, double y)
{
    const double q0 = div_fast_small(a0, c, y), q1 = div_fast_small(a1, c, y);
    if (div_slow(a0) | div_slow(a1)) { a0 = div_generic(a0, c); a1 = div_generic(a1, c); }
    else { a0 = q0; a1 = q1; }
}
__device__ __forceinline__ void div_small3(double &a0, double &a1, double &a2, double c, double y)
{
    const double q0 = div_fast_small(a0, c, y), q1 = div_fast_small(a1, c, y), q2 = div_fast_small(a2, c, y);
    if (div_slow(a0) | div_slow(a1) | div_slow(a2)) {
        a0 = div_generic(a0, c); a1 = div_generic(a1, c); a2 = div_generic(a2, c);
    } else { a0 = q0; a1 = q1; a2 = q2; }
}

// Two / three independent quotients behind ONE range test (in place): the generic division returns the
// same correctly rounded value for an in-range dividend, so when any of the group leaves the window all of
// them are redone.  Saves the branch / reconvergence instructions of the other tests (a third of the
// non-arithmetic instructions of a fourth-order difference).
__device__ __forceinline__ void div_exact2(double &a0, double &a1, double c0, double y0, double c1, double y1)
{
    const double q0 = div_fast(a0, c0, y0), q1 = div_fast(a1, c1, y1);
    if (div_slow(a0) | div_slow(a1)) { a0 = div_generic(a0, c0); a1 = div_generic(a1, c1); }
    else { a0 = q0; a1 = q1; }
}
__device__ __forceinline__ void div_exact3(double &a0, double &a1, double &a2, double c0, double y0, double c1, double y1,
                                           double c2, double y2)
{
    const double q0 = div_fast(a0, c0, y0), q1 = div_fast(a1, c1, y1), q2 = div_fast(a2, c2, y2);
    if (div_slow(a0) | div_slow(a1) | div_slow(a2)) {
        a0 = div_generic(a0, c0); a1 = div_generic(a1, c1); a2 = div_generic(a2, c2);
    } else { a0 = q0; a1 = q1; a2 = q2; }
}
#endif

// ------------------------------------------------------------------ 3-D
// T = double: the reference's precision.  T = float: the single-precision build the reference endorses ("significantly
// faster", 3D-iso :114-116): fields, memory variables, profiles and update constants in single precision; the energy
// is still accumulated in double.
template <typename T>
struct Params3DT {
    int nx, ny, nzl;          // local slab extent
    int nz;                   // global NZ
    int koff;                 // offset_k = rank * NZ_LOCAL (3D-iso :397)
    int pitch;                // row pitch in doubles
    long long plane;          // plane pitch in doubles (= pitch * ny)
    // fields: pointer to element (i=1, j=1, k=0)
    T *vx, *vy, *vz, *sxx, *syy, *szz, *sxy, *sxz, *syz;
    // shells
    int xlo, xhi, sxp;        // sxp: padded x-shell row length
    int ylo, yhi, sy;         // sy: number of y-shell rows
    int zlo, zhi;             // global z shell
    int zbase;                // global shell index of this slab's first stored z-shell plane
    // memory variables, see kernels_3d.cu for the order inside each group
    T *mx[6], *my[6], *mz[6];
    AxisCoefT<T> cx, cy, cz;  // cz is indexed by GLOBAL k
    T odx, ody, odz;          // ONE_OVER_DELTAX.. (:134-136)
    T dt_lambda, dt_mu, dt_lambdaplus2mu, dt_over_rho;   // :296-300
    // velocity kernel extras
    int it;                   // time step (1-based)
    int isrc, jsrc, ksrc;     // source point, ksrc local (0: not on this slab)
    const double *src_x, *src_y;   // force*DELTAT/rho per step (:1080-1081)
    int npml;                 // energy box (:1135-1145)
    int energy_bug_compat;
    double rho, lambda, mu;
    double inv_den, inv_2mu;  // 1/(2 mu (3 lambda + 2 mu)), 1/(2 mu): energy (:1159-1167)
    double inv_mu, c2lm, half_rho;   // 1/mu, 2 (lambda + mu), rho/2
    double *partials;         // [2][nblocks] kinetic / potential per block
    int nblocks;
    int kunit;                // every K profile == 1: the value/K division is dropped (exact)
    // slab neighbours' halo planes, element (1,1), mapped peer memory (null: no neighbour /
    // exchange done by the driver).  lo = slab rank-1, its plane NZ_LOCAL+1: vx, vy, sigmazz;
    // hi = slab rank+1, its plane 0: vz, sigmaxz, sigmayz  (3D-iso :811-823, :951-963)
    T *peer_lo[3];
    T *peer_hi[3];
};
using Params3D = Params3DT<double>;
using Params3DF = Params3DT<float>;

// TMA descriptors of one kernel's nine plane tiles (kernels_3d_tma.cu lists the order).
struct TmaMaps {
    CUtensorMap m[9];
};

// Work decomposition of the TMA kernels: persistent CTAs, static round-robin over
// (x-tile, y-tile, z-chunk) items.
struct Tile3D {
    int tx, ty;               // thread tile = box size
    int ntx, nty;             // tiles per plane
    int kchunk, nzc;          // planes per item, z chunks
    int nitems;               // work items (= energy partial slots): ntx * nty * nzc, more with a finer tail
    int fine_from, split;     // kernels_3d_ws.cu: coarse items >= fine_from are `split` items each (split = 1: none)
    int stages;               // shared-memory ring depth (planes)
    int minb;                 // resident CTAs per SM the kernel variant is compiled for
    int xm_bytes;             // per stage and variable: x-shell memory variables of the tile rows (0: no x shell)
    int grid_stress, grid_velocity;
    unsigned int *queue;      // [2] work queue of the persistent kernels (tma_common.cuh: claim_item); null: static shares
};

// Slab-to-slab ordering done INSIDE the update kernels (kernels_3d_ws.cu): the work items that read a halo plane
// poll this slab's flag words, the CTA that finishes the last boundary item of a side publishes to that neighbour.
// All pointers null: single slab, or the exchange is done by the driver.
struct SlabSync {
    const unsigned long long *wait_lo, *wait_hi;   // this slab's flag words written by slab rank-1 / rank+1
    unsigned long long wait_value;                 // epoch << 32 | time step the planes must belong to
    unsigned long long *pub_lo, *pub_hi;           // the neighbours' flag words this slab writes (mapped peer memory)
    unsigned long long pub_value;
    unsigned int *count;                           // [2] boundary items completed on the lo / hi side (self-resetting)
    int n_boundary;                                // work items per boundary chunk (= tiles per plane)
    unsigned int *timeout;                         // set when a poll gives up
};

// Work decomposition of the TMA-staged 2-D kernels (kernels_2d_ws.cu): strips of tx columns, marched along y in blocks
// of rb rows; a work item is (strip, y chunk of `rows` rows).
struct Tile2D {
    int tx, rb;
    int ntx, rows, nchunks, nitems;
    int grid_stress, grid_velocity;
    unsigned int *queue;      // work queue of the persistent kernels (see Tile3D)
};

// One launch region of a 3-D kernel: the box [i0,i1] x [j0,j1] x [k0,k1] (1-based, k local).
// The thread grid starts at ia <= i0 (ia-1 a multiple of 4, so warp rows stay 32-byte
// sector aligned); lanes with i < i0 idle.
struct Box3D {
    int i0, i1, j0, j1, k0, k1;
    int ia;
    int kchunk;               // planes marched by one block
    int pbase;                // offset of this region's blocks in the energy partials
    int pml;                  // 1: region touches a PML shell -> <PML=true> instantiation
    int tx, ty;               // thread tile
    int gx, gy, gz;           // grid
};

struct Post3D {
    const double *partials;
    int nblocks;                   // kinetic partials [0, nblocks)
    int npot;                      // potential partials [nblocks, nblocks + npot)
    double *energy_k, *energy_p;   // traces, slot it-1 is written
    double *step_out;              // [4] this step's kinetic, potential, sisvx(it,1), sisvy(it,1) side by side (cpml_fetch_step)
    int it, nstep, nrec;
    const int *ix_rec, *iy_rec;
    const double *vx, *vy;         // element (1,1,0)
    const double *vz;              // 3-D only (null in 2-D): the Vz extension, quirk B7
    int pitch;
    long long plane;
    int krec;                      // local k of the receiver plane, 0: not here
    int f32;                       // 1: vx, vy, vz point to single-precision fields
    double *sisvx, *sisvy, *sisvz;
};

// ------------------------------------------------------------------ 3-D viscoelastic
// seismic_CPML_3D_viscoelastic_MPI.f90: fourth order, N_SLS = 2.  Fields carry TWO z halo planes
// per side (k = -1..NZ_LOCAL+2, 3D-visco :301) and a two-cell zero ghost ring in x and y (the
// reference's (0:NX+1,0:NY+1) extents, widened so that discarded lanes stay in bounds).
struct ParamsV3D {
    int nx, ny, nzl, nz, koff, pitch;
    long long plane;
    double *vx, *vy, *vz, *sxx, *syy, *szz, *sxy, *sxz, *syz;   // element (1,1,0)
    double *rxx, *ryy, *rzz, *rxy, *rxz, *ryz;                  // sigma*_R (:302)
    double2 *e1, *e11, *e22, *e12, *e13, *e23;                  // (N_SLS = 2, ...) : both mechanisms of a point
    int xlo, xhi, sxp, ylo, yhi, sy, zlo, zhi, zbase;
    double *mx[6], *my[6], *mz[6];                              // same order as the isotropic kernels
    AxisCoef cx, cy, cz;
    double odx, ody, odz, dt, dt_over_rho;
    // constants of :982-987 and :458-477, evaluated on the host in the reference's order
    double lam, mu, l2m_r, lam23mu, two_mu, two_thirds_mu, lam_u, mu_u, l2m_u;
    double szz_e1, szz_dev;   // coefficients of the sigmazz memory-variable term: (l2m_r, two_thirds_mu) as in the
                              // reference (:1058-1060, quirk B14) or (lam23mu, two_mu) with cfg.sigmazz_isotropic
    double phi1[2], phi2[2], tauinv1[2], tauinv2[2], den1[2], den2[2];   // den = 1 - dt*0.5*tauinv
    double rden1[2], rden2[2];                                           // RN(1/den) for div_exact
    int nzl_e;                // plane count of one EMULATED reference slab (quirk B6); nz: none
    int it, isrc, jsrc, ksrc;
    const double *src_x, *src_y;
    int npml;
    double half_rho, c2lm, inv_den, inv_2mu;
    double *partials;         // [0, nb) kinetic (velocity kernel), [nb, 2nb) and [2nb, 3nb) potential (stress launches)
    int nblocks;
    int kchunk;               // planes marched by one block
    // slab neighbours' fields, element (1,1,0), mapped peer memory (null: no neighbour / exchange done by the driver):
    // [0] vx [1] vy [2] sigmazz go "left" (3D-visco :963-969, :1230-1232), [3] vz [4] sigmaxz [5] sigmayz "right"
    double *peer_lo[6], *peer_hi[6];
    int pf;                   // L2 prefetch (set by the dispatcher, CPML_VPF).  Bits 0-1, stress kernel, streamed words:
                              // 0 off, 1 one plane ahead, 2 staggered by half a plane; bit 2, both kernels: the C-PML
                              // memory variables of the next plane
};

// ------------------------------------------------------------------ 2-D
struct Params2D {
    int nx, ny, pitch;
    int order;                // 2 or 4
    double *vx, *vy, *sxx, *syy, *sxy;       // pointer to element (i=1, j=1)
    const double *lambda, *mu, *rho;         // same layout, zero ghost ring
    int rho_exact;                           // 1: no density has an all-ones significand (div_rho)
    int xlo, xhi, sxp;
    int ylo, yhi, sy;
    double *mx[4], *my[4];
    AxisCoef cx, cy;
    double deltax, deltay, deltat;
    double denx, rdenx, deny, rdeny;         // DELTAX (2nd order) or 24*DELTAX (4th), and RN(1/.)
    // viscoelastic programs (kernels_2d_visco.cu): 9/(8 DELTA) and 1/(24 DELTA) (2D-visco-4th :210-213;
    // second order: c98 = 1/DELTA), the N_SLS = 3 memory variables and the constants of :386-399
    double c98x, c24x, c98y, c24y;
    double *e1[3], *e11[3], *e13[3];
    double half1[3], half2[3], mul1[3], mul2[3], dt_phi1[3], dt_phi2[3];
    int it;
    int isrc, jsrc;
    const double *force_x, *force_y;         // raw force series (2D-2nd :656-657)
    int npml;
    double *partials;
    int nblocks;
};

struct Post2D {
    const double *partials;
    int nblocks;
    double *energy_k, *energy_p;
    int it, nstep, nrec;
    const int *ix_rec, *iy_rec;
    const double *vx, *vy;
    int pitch;
    double *sisvx, *sisvy;
};

// launchers (kernels_3d.cu / kernels_2d.cu)
void launch_stress3d(const Params3D &p, const Box3D &b, cudaStream_t s);
void launch_velocity3d(const Params3D &p, const Box3D &b, cudaStream_t s);
bool tile_supported(int tx, int ty);
void launch_post3d(const Post3D &p, cudaStream_t s);
bool tma_tile_supported(int tx, int ty);
cudaError_t tma_occupancy(const Params3D &p, const Tile3D &t, bool stress, int *occ);
cudaError_t launch_stress3d_tma(const Params3D &p, const TmaMaps &tm, const Tile3D &t, cudaStream_t s);
cudaError_t launch_velocity3d_tma(const Params3D &p, const TmaMaps &tm, const Tile3D &t, cudaStream_t s);
bool ws_tile_supported(int tx, int ty, bool f32);
cudaError_t ws_occupancy(const Params3D &p, const Tile3D &t, bool stress, int *occ, bool f32);
cudaError_t launch_stress3d_ws(const Params3D &p, const TmaMaps &tm, const Tile3D &t, const SlabSync &ss, cudaStream_t s);
cudaError_t launch_velocity3d_ws(const Params3D &p, const TmaMaps &tm, const Tile3D &t, const SlabSync &ss, cudaStream_t s);
cudaError_t launch_stress3d_ws(const Params3DF &p, const TmaMaps &tm, const Tile3D &t, const SlabSync &ss, cudaStream_t s);
cudaError_t launch_velocity3d_ws(const Params3DF &p, const TmaMaps &tm, const Tile3D &t, const SlabSync &ss, cudaStream_t s);
void launch_f2d(const float *src, double *dst, long long n, cudaStream_t s);
void launch_maxnorm_f(const float *vx, const float *vy, const float *vz, long long n, unsigned long long *out_bits, cudaStream_t s);
void launch_signal(unsigned long long *flag_lo, unsigned long long *flag_hi, unsigned long long value, cudaStream_t s);
void launch_wait(const unsigned long long *flag_a, const unsigned long long *flag_b, unsigned long long value,
                 unsigned int *timeout_flag, cudaStream_t s);
bool vws_tile_supported(int tx, int ty);
void vws_boxes(int tx, int ty, int (*box)[2]);
cudaError_t vws_occupancy(const Tile3D &t, int *occ);
cudaError_t launch_vvelocity3d_ws(const ParamsV3D &p, const TmaMaps &tm, const Tile3D &t, cudaStream_t s);
void launch_vstress3d(const ParamsV3D &p, dim3 grid, cudaStream_t s);
void launch_vvelocity3d(const ParamsV3D &p, dim3 grid, cudaStream_t s);
void visco_tile(int *tx, int *ty);
int visco_stress_launches();
void ws2_geometry(int *tx, int *rb, int (*box_stress)[2], int (*box_velocity)[2]);
cudaError_t ws2_occupancy(int order, bool stress, int *occ);
cudaError_t launch_stress2d_ws(const Params2D &p, const TmaMaps &tm, const Tile2D &t, cudaStream_t s);
cudaError_t launch_velocity2d_ws(const Params2D &p, const TmaMaps &tm, const Tile2D &t, cudaStream_t s);
void launch_stress2d(const Params2D &p, dim3 grid, dim3 block, cudaStream_t s);
void launch_velocity2d(const Params2D &p, dim3 grid, dim3 block, cudaStream_t s);
void launch_post2d(const Post2D &p, cudaStream_t s);
void launch_vstress2d(const Params2D &p, dim3 grid, cudaStream_t s);
void launch_vvelocity2d(const Params2D &p, dim3 grid, cudaStream_t s);
void launch_vpressure2d(const Params2D &p, const int *ix_rec, const int *iy_rec, int nrec, int nstep, double *sispressure, cudaStream_t s);
void launch_venergy2d(const Params2D &p, dim3 grid, cudaStream_t s);
void launch_maxnorm(const double *vx, const double *vy, const double *vz, long long n,
                    unsigned long long *out_bits, cudaStream_t s);

}  // namespace cpml
