// 3-D isotropic C-PML kernels for sm_100a, TMA-staged (the default path).
//
// Same two fused kernels per time step as kernels_3d.cu (the register-marching variant kept
// for A/B measurements):
//
//   k_stress3d_tma    sigmaxx/yy/zz (:836-863), sigmaxy (:877-894), sigmaxz/yz (:908-943)
//   k_velocity3d_tma  vx/vy (:976-1017), vz (:1031-1052), source (:1055-1083), Dirichlet
//                     faces (:1087-1121), energy partials (:1131-1177)
// (line numbers: seismic_CPML_3D_isotropic_MPI_OpenMP.f90)
//
// Why TMA: ncu on the register-marching kernels (profiles/r01_v3_ncu_cfg3.txt) shows 124-128
// registers per thread, 24 % occupancy and 6-14 long-scoreboard stalls per issue: the bytes in
// flight per SM are capped by the registers that receive them.  Here every wavefield plane
// tile is fetched by the TMA unit (cp.async.bulk.tensor.3d) into a ring of shared-memory
// stages, S planes deep, completion signalled on mbarriers; the registers only hold the
// values carried along z and the C-PML memory variables of the shell points.
//
// Mapping: persistent CTAs (grid = SMs x resident CTAs), static round-robin over work items
// (x-tile, y-tile, z-chunk); a CTA owns a TX x TY tile and marches kchunk planes.  Per plane
// one stage holds nine tiles: the fields that need an in-plane neighbour come as
// (TX+2) x (TY+1) boxes shifted by (-2|0, -1|0), the others as TX x TY boxes.  Boxes that
// hang over the grid are zero-filled by the TMA (the loop bounds never use those values).
// The x shift is -2, not -1: measured with tools/tma_probe.cu on B200, a tiled FP64 load whose
// first element is not 16-byte aligned (odd x) raises "illegal instruction"; negative
// coordinates, boxes larger than the tensor and zero fill all behave as documented.
// Thread (tx,ty) updates point (i0+tx, j0+ty) of the plane from shared memory and stores the
// results with streaming stores.  Plane k+1 of the fields differenced forward in z
// (vx, vy / sigmazz) is read from the next stage; plane k-1 (vz / sigmaxz, sigmayz) is
// carried in registers.
//
// Slab decomposition: the boundary planes every neighbour needs (:811-823, :951-963) are
// stored by the same kernels straight into the neighbour GPU's halo plane over NVLink
// (peer pointers in Params3D, null without a neighbour), replacing MPI_SENDRECV.
//
// Arithmetic: compiled with -fmad=false, same expressions in the same order as the
// reference, so fields are bit-identical to an IEEE (non-FMA) build of the Fortran loops.
#include <cuda.h>

#include "cpml_internal.h"

namespace cpml {

namespace {

constexpr int kBarBytes = 128;   // room for up to 16 mbarriers ahead of the stages

__host__ __device__ constexpr int round128(int v) { return (v + 127) / 128 * 128; }

template <int TX, int TY>
struct TileGeom {
    static constexpr int W = TX + 2;                       // halo box row length
    static constexpr int HROWS = TY + 1;                   // halo box rows
    static constexpr int HALO_BOX_BYTES = W * HROWS * 8;
    static constexpr int PLAIN_BOX_BYTES = TX * TY * 8;
    static constexpr int HALO_BYTES = round128(HALO_BOX_BYTES);
    static constexpr int PLAIN_BYTES = round128(PLAIN_BOX_BYTES);
};

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
// Bounded wait: a TMA that never lands (bad descriptor) traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity))
        if (++spins > (1u << 24)) __trap();
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int x, int y, int z, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"((unsigned long long)map), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}

__device__ __forceinline__ void st_stream(double *p, double v) { __stcs(p, v); }

// memory_x = b * memory_x + a * value ; value = value / K + memory_x   (e.g. :845-851)
template <bool KUNIT>
__device__ __forceinline__ double cpml_apply(double *__restrict__ mem, long long q, double m,
                                             double b, double a, double K, double value)
{
    m = b * m + a * value;
    mem[q] = m;
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

// ------------------------------------------------------------------------------- stress
// maps: 0 vx (halo box at (0,0))  1 vy (halo, (-2,-1))  2 vz (halo, (-2,0))
//       3 sxx 4 syy 5 szz 6 sxy 7 sxz 8 syz (plain boxes)
template <bool KUNIT, int TX, int TY, int MINB>
__global__ void __launch_bounds__(TX *TY, MINB)
k_stress3d_tma(const __grid_constant__ Params3D p, const __grid_constant__ TmaMaps tm, const __grid_constant__ Tile3D t)
{
    using G = TileGeom<TX, TY>;
    constexpr int W = G::W;
    constexpr int STAGE_BYTES = 3 * G::HALO_BYTES + 6 * G::PLAIN_BYTES;
    constexpr uint32_t TX_FULL = 3 * G::HALO_BOX_BYTES + 6 * G::PLAIN_BOX_BYTES;
    constexpr uint32_t TX_NEXT = 2 * G::HALO_BOX_BYTES;          // vx, vy of plane ke+1

    extern __shared__ unsigned char smem_dyn[];
    const uint32_t sbase = (smem_u32(smem_dyn) + 127u) & ~127u;
    unsigned char *gbase = smem_dyn + (sbase - smem_u32(smem_dyn));
    const uint32_t bar0 = sbase;                                  // S mbarriers, 8 bytes each
    const uint32_t stage0 = sbase + kBarBytes;
    const unsigned char *gstage0 = gbase + kBarBytes;
    const int S = t.stages;

    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tid = ty * TX + tx;
    if (tid == 0) {
        for (int s = 0; s < S; s++) mbar_init(bar0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int pitch = p.pitch;
    const long long pl = p.plane;
    const double odx = p.odx, ody = p.ody, odz = p.odz;
    const double dt_l = p.dt_lambda, dt_m = p.dt_mu, dt_l2m = p.dt_lambdaplus2mu;

    uint32_t g = 0;     // running count of plane loads of this CTA: stage = g % S, parity = (g / S) & 1
    for (int item = blockIdx.x; item < t.nitems; item += gridDim.x) {
        const int tix = item % t.ntx;
        const int rest = item / t.ntx;
        const int tiy = rest % t.nty;
        const int zc = rest / t.nty;
        const int i0 = 1 + tix * TX, j0 = 1 + tiy * TY;
        const int kb = 1 + zc * t.kchunk;
        const int ke = min(p.nzl, kb + t.kchunk - 1);
        const int np = ke - kb + 1;

        // plane load l of this item: planes kb .. ke complete, plane ke+1 only vx, vy
        auto issue = [&](int l) {
            const uint32_t gl = g + (uint32_t)l;
            const uint32_t s = gl % (uint32_t)S;
            const uint32_t bar = bar0 + 8 * s;
            const uint32_t dst = stage0 + s * STAGE_BYTES;
            const int k = kb + l;
            const int x0 = i0 - 1, y0 = j0 - 1;
            if (l < np) {
                mbar_expect_tx(bar, TX_FULL);
                tma_load_3d(dst + 0 * G::HALO_BYTES, &tm.m[0], x0, y0, k, bar);
                tma_load_3d(dst + 1 * G::HALO_BYTES, &tm.m[1], x0 - 2, y0 - 1, k, bar);
                tma_load_3d(dst + 2 * G::HALO_BYTES, &tm.m[2], x0 - 2, y0, k, bar);
#pragma unroll
                for (int f = 0; f < 6; f++)
                    tma_load_3d(dst + 3 * G::HALO_BYTES + f * G::PLAIN_BYTES, &tm.m[3 + f], x0, y0, k, bar);
            } else {
                mbar_expect_tx(bar, TX_NEXT);
                tma_load_3d(dst + 0 * G::HALO_BYTES, &tm.m[0], x0, y0, k, bar);
                tma_load_3d(dst + 1 * G::HALO_BYTES, &tm.m[1], x0 - 2, y0 - 1, k, bar);
            }
        };
        if (tid == 0) {
            const int npro = min(S, np + 1);
            for (int l = 0; l < npro; l++) issue(l);
        }

        const int i = i0 + tx, j = j0 + ty;
        const bool valid = (i <= p.nx) && (j <= p.ny);
        long long q = (long long)kb * pl + (long long)(j - 1) * pitch + (i - 1);

        const bool in_x = valid && ((i <= p.xlo) || (i >= p.xhi));
        const bool in_y = valid && ((j <= p.ylo) || (j >= p.yhi));
        const int sx = in_x ? shell_index(i, p.xlo, p.xhi) : 0;
        const int sy = in_y ? shell_index(j, p.ylo, p.yhi) : 0;

        // loop bounds of the four nests (i, j part; the k part is tested per plane)
        const bool do_n = valid && (i <= p.nx - 1) && (j >= 2);     // :838-839
        const bool do_xy = valid && (i >= 2) && (j <= p.ny - 1);    // :878-879
        const bool do_xz = valid && (i >= 2);                       // :910-911
        const bool do_yz = valid && (j <= p.ny - 1);                // :927-928

        double ax = 0, bxc = 0, Kx = 1, axh = 0, bxh = 0, Kxh = 1, ay = 0, by = 0, Ky = 1, ayh = 0, byh = 0, Kyh = 1;
        if (in_x) { ax = p.cx.a[i]; bxc = p.cx.b[i]; axh = p.cx.a_half[i]; bxh = p.cx.b_half[i];
                    if (!KUNIT) { Kx = p.cx.K[i]; Kxh = p.cx.K_half[i]; } }
        if (in_y) { ay = p.cy.a[j]; by = p.cy.b[j]; ayh = p.cy.a_half[j]; byh = p.cy.b_half[j];
                    if (!KUNIT) { Ky = p.cy.K[j]; Kyh = p.cy.K_half[j]; } }

        // x / y shell memory variables are fetched one plane ahead of their use
        long long qx = in_x ? ((long long)(kb - 1) * p.ny + (j - 1)) * p.sxp + sx : 0;
        long long qy = in_y ? ((long long)(kb - 1) * p.sy + sy) * pitch + (i - 1) : 0;
        const long long qx_step = (long long)p.ny * p.sxp, qy_step = (long long)p.sy * pitch;
        double m_x0 = 0, m_x1 = 0, m_x2 = 0, m_y0 = 0, m_y1 = 0, m_y2 = 0;
        if (in_x) { m_x0 = p.mx[0][qx]; m_x1 = p.mx[1][qx]; m_x2 = p.mx[2][qx]; }
        if (in_y) { m_y0 = p.my[0][qy]; m_y1 = p.my[1][qy]; m_y2 = p.my[2][qy]; }

        double vz_m = valid ? p.vz[q - pl] : 0.0;                   // plane kb-1, carried along z

        for (int n = 0; n < np; ++n, q += pl, qx += qx_step, qy += qy_step) {
            const int k = kb + n;
            const int kg = k + p.koff;                              // :837
            const uint32_t gc = g + (uint32_t)n, gn = gc + 1;
            const uint32_t sc = gc % (uint32_t)S, sn = gn % (uint32_t)S;

            // next plane's x / y memory variables
            double n_x0 = 0, n_x1 = 0, n_x2 = 0, n_y0 = 0, n_y1 = 0, n_y2 = 0;
            if (n + 1 < np) {
                if (in_x) { n_x0 = p.mx[0][qx + qx_step]; n_x1 = p.mx[1][qx + qx_step]; n_x2 = p.mx[2][qx + qx_step]; }
                if (in_y) { n_y0 = p.my[0][qy + qy_step]; n_y1 = p.my[1][qy + qy_step]; n_y2 = p.my[2][qy + qy_step]; }
            }
            const bool in_z = valid && ((kg <= p.zlo) || (kg >= p.zhi));
            long long qz = 0;
            double m_z0 = 0, m_z1 = 0, m_z2 = 0, az = 0, bz = 0, Kz = 1, azh = 0, bzh = 0, Kzh = 1;
            if (in_z) {
                qz = ((long long)(shell_index(kg, p.zlo, p.zhi) - p.zbase) * p.ny + (j - 1)) * pitch + (i - 1);
                m_z0 = p.mz[0][qz]; m_z1 = p.mz[1][qz]; m_z2 = p.mz[2][qz];
                az = p.cz.a[kg]; bz = p.cz.b[kg]; azh = p.cz.a_half[kg]; bzh = p.cz.b_half[kg];
                if (!KUNIT) { Kz = p.cz.K[kg]; Kzh = p.cz.K_half[kg]; }
            }

            if (n == 0) mbar_wait(bar0 + 8 * sc, (gc / (uint32_t)S) & 1u);
            mbar_wait(bar0 + 8 * sn, (gn / (uint32_t)S) & 1u);

            const unsigned char *st = gstage0 + (size_t)sc * STAGE_BYTES;
            const unsigned char *stn = gstage0 + (size_t)sn * STAGE_BYTES;
            const double *Tvx = (const double *)(st + 0 * G::HALO_BYTES);
            const double *Tvy = (const double *)(st + 1 * G::HALO_BYTES);
            const double *Tvz = (const double *)(st + 2 * G::HALO_BYTES);
            const double *Ts = (const double *)(st + 3 * G::HALO_BYTES);
            constexpr int PD = G::PLAIN_BYTES / 8;
            const int c = ty * TX + tx;

            const double vx_c = Tvx[ty * W + tx], vx_ip = Tvx[ty * W + tx + 1], vx_jp = Tvx[(ty + 1) * W + tx];
            const double vy_c = Tvy[(ty + 1) * W + tx + 2], vy_im = Tvy[(ty + 1) * W + tx + 1], vy_jm = Tvy[ty * W + tx + 2];
            const double vz_c = Tvz[ty * W + tx + 2], vz_im = Tvz[ty * W + tx + 1], vz_jp = Tvz[(ty + 1) * W + tx + 2];
            const double vx_n = ((const double *)(stn + 0 * G::HALO_BYTES))[ty * W + tx];
            const double vy_n = ((const double *)(stn + 1 * G::HALO_BYTES))[(ty + 1) * W + tx + 2];
            const double sxx = Ts[0 * PD + c], syy = Ts[1 * PD + c], szz = Ts[2 * PD + c];
            const double sxy = Ts[3 * PD + c], sxz = Ts[4 * PD + c], syz = Ts[5 * PD + c];

            double szz_out = szz, sxz_out = sxz, syz_out = syz;

            // ---- sigmaxx, sigmayy, sigmazz  (:836-863)
            if (do_n && kg >= 2) {                                  // k2begin, :792-793
                double value_dvx_dx = (vx_ip - vx_c) * odx;
                double value_dvy_dy = (vy_c - vy_jm) * ody;
                double value_dvz_dz = (vz_c - vz_m) * odz;
                if (in_x) value_dvx_dx = cpml_apply<KUNIT>(p.mx[0], qx, m_x0, bxh, axh, Kxh, value_dvx_dx);
                if (in_y) value_dvy_dy = cpml_apply<KUNIT>(p.my[0], qy, m_y0, by, ay, Ky, value_dvy_dy);
                if (in_z) value_dvz_dz = cpml_apply<KUNIT>(p.mz[0], qz, m_z0, bz, az, Kz, value_dvz_dz);
                st_stream(p.sxx + q, dt_l2m * value_dvx_dx + dt_l * (value_dvy_dy + value_dvz_dz) + sxx);
                st_stream(p.syy + q, dt_l * (value_dvx_dx + value_dvz_dz) + dt_l2m * value_dvy_dy + syy);
                szz_out = dt_l * (value_dvx_dx + value_dvy_dy) + dt_l2m * value_dvz_dz + szz;
                st_stream(p.szz + q, szz_out);
            }
            // ---- sigmaxy  (:877-894)
            if (do_xy) {
                double value_dvy_dx = (vy_c - vy_im) * odx;
                double value_dvx_dy = (vx_jp - vx_c) * ody;
                if (in_x) value_dvy_dx = cpml_apply<KUNIT>(p.mx[1], qx, m_x1, bxc, ax, Kx, value_dvy_dx);
                if (in_y) value_dvx_dy = cpml_apply<KUNIT>(p.my[1], qy, m_y1, byh, ayh, Kyh, value_dvx_dy);
                st_stream(p.sxy + q, dt_m * (value_dvy_dx + value_dvx_dy) + sxy);
            }
            // ---- sigmaxz, sigmayz  (:908-943)
            if (kg <= p.nz - 1) {                                   // kminus1end, :795-796
                if (do_xz) {
                    double value_dvz_dx = (vz_c - vz_im) * odx;
                    double value_dvx_dz = (vx_n - vx_c) * odz;
                    if (in_x) value_dvz_dx = cpml_apply<KUNIT>(p.mx[2], qx, m_x2, bxc, ax, Kx, value_dvz_dx);
                    if (in_z) value_dvx_dz = cpml_apply<KUNIT>(p.mz[1], qz, m_z1, bzh, azh, Kzh, value_dvx_dz);
                    sxz_out = dt_m * (value_dvz_dx + value_dvx_dz) + sxz;
                    st_stream(p.sxz + q, sxz_out);
                }
                if (do_yz) {
                    double value_dvz_dy = (vz_jp - vz_c) * ody;
                    double value_dvy_dz = (vy_n - vy_c) * odz;
                    if (in_y) value_dvz_dy = cpml_apply<KUNIT>(p.my[2], qy, m_y2, byh, ayh, Kyh, value_dvz_dy);
                    if (in_z) value_dvy_dz = cpml_apply<KUNIT>(p.mz[2], qz, m_z2, bzh, azh, Kzh, value_dvy_dz);
                    syz_out = dt_m * (value_dvz_dy + value_dvy_dz) + syz;
                    st_stream(p.syz + q, syz_out);
                }
            }
            // ---- boundary planes go straight into the neighbour slabs' halo planes (:951-963)
            if (valid) {
                const long long qp = (long long)(j - 1) * pitch + (i - 1);
                if (k == 1 && p.peer_lo[2]) p.peer_lo[2][qp] = szz_out;                  // sigmazz(:,:,1) -> left
                if (k == p.nzl && p.peer_hi[1]) { p.peer_hi[1][qp] = sxz_out; p.peer_hi[2][qp] = syz_out; }   // -> right
            }

            vz_m = vz_c;
            m_x0 = n_x0; m_x1 = n_x1; m_x2 = n_x2; m_y0 = n_y0; m_y1 = n_y1; m_y2 = n_y2;

            __syncthreads();                                        // stage sc is free
            if (tid == 0 && n + S <= np) issue(n + S);
        }
        g += (uint32_t)(np + 1);
    }
}

// ----------------------------------------------------------------------------- velocity
// maps: 0 sxx (halo box at (-2,0))  1 syy (halo, (0,0))  2 sxy (halo, (0,-1))
//       3 sxz (halo, (0,0))  4 syz (halo, (0,-1))  5 szz  6 vx  7 vy  8 vz (plain boxes)
template <bool KUNIT, int TX, int TY, int MINB>
__global__ void __launch_bounds__(TX *TY, MINB)
k_velocity3d_tma(const __grid_constant__ Params3D p, const __grid_constant__ TmaMaps tm, const __grid_constant__ Tile3D t)
{
    using G = TileGeom<TX, TY>;
    constexpr int W = G::W;
    constexpr int STAGE_BYTES = 5 * G::HALO_BYTES + 4 * G::PLAIN_BYTES;
    constexpr uint32_t TX_FULL = 5 * G::HALO_BOX_BYTES + 4 * G::PLAIN_BOX_BYTES;
    constexpr uint32_t TX_NEXT = G::PLAIN_BOX_BYTES;              // sigmazz of plane ke+1
    constexpr int NT = TX * TY;

    __shared__ double red[2 * ((NT + 31) / 32)];
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t sbase = (smem_u32(smem_dyn) + 127u) & ~127u;
    unsigned char *gbase = smem_dyn + (sbase - smem_u32(smem_dyn));
    const uint32_t bar0 = sbase;
    const uint32_t stage0 = sbase + kBarBytes;
    const unsigned char *gstage0 = gbase + kBarBytes;
    const int S = t.stages;

    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tid = ty * TX + tx;
    if (tid == 0) {
        for (int s = 0; s < S; s++) mbar_init(bar0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int pitch = p.pitch;
    const long long pl = p.plane;
    const double odx = p.odx, ody = p.ody, odz = p.odz, dt_r = p.dt_over_rho;
    // energy constants (:1159-1167); reciprocals instead of the reference's divisions -- the
    // energy sum is reduction-order dependent anyway (quirk B11)
    const double lam = p.lambda, mu = p.mu;
    const double c2lm = 2.0 * (lam + mu);
    const double inv_den = 1.0 / (2.0 * mu * (3.0 * lam + 2.0 * mu));
    const double inv_2mu = 1.0 / (2.0 * mu);
    const double half_rho = 0.5 * p.rho;

    uint32_t g = 0;
    for (int item = blockIdx.x; item < t.nitems; item += gridDim.x) {
        const int tix = item % t.ntx;
        const int rest = item / t.ntx;
        const int tiy = rest % t.nty;
        const int zc = rest / t.nty;
        const int i0 = 1 + tix * TX, j0 = 1 + tiy * TY;
        const int kb = 1 + zc * t.kchunk;
        const int ke = min(p.nzl, kb + t.kchunk - 1);
        const int np = ke - kb + 1;

        auto issue = [&](int l) {
            const uint32_t gl = g + (uint32_t)l;
            const uint32_t s = gl % (uint32_t)S;
            const uint32_t bar = bar0 + 8 * s;
            const uint32_t dst = stage0 + s * STAGE_BYTES;
            const int k = kb + l;
            const int x0 = i0 - 1, y0 = j0 - 1;
            if (l < np) {
                mbar_expect_tx(bar, TX_FULL);
                tma_load_3d(dst + 0 * G::HALO_BYTES, &tm.m[0], x0 - 2, y0, k, bar);
                tma_load_3d(dst + 1 * G::HALO_BYTES, &tm.m[1], x0, y0, k, bar);
                tma_load_3d(dst + 2 * G::HALO_BYTES, &tm.m[2], x0, y0 - 1, k, bar);
                tma_load_3d(dst + 3 * G::HALO_BYTES, &tm.m[3], x0, y0, k, bar);
                tma_load_3d(dst + 4 * G::HALO_BYTES, &tm.m[4], x0, y0 - 1, k, bar);
#pragma unroll
                for (int f = 0; f < 4; f++)
                    tma_load_3d(dst + 5 * G::HALO_BYTES + f * G::PLAIN_BYTES, &tm.m[5 + f], x0, y0, k, bar);
            } else {
                mbar_expect_tx(bar, TX_NEXT);
                tma_load_3d(dst + 5 * G::HALO_BYTES, &tm.m[5], x0, y0, k, bar);
            }
        };
        if (tid == 0) {
            const int npro = min(S, np + 1);
            for (int l = 0; l < npro; l++) issue(l);
        }

        const int i = i0 + tx, j = j0 + ty;
        const bool valid = (i <= p.nx) && (j <= p.ny);
        long long q = (long long)kb * pl + (long long)(j - 1) * pitch + (i - 1);

        const bool in_x = valid && ((i <= p.xlo) || (i >= p.xhi));
        const bool in_y = valid && ((j <= p.ylo) || (j >= p.yhi));
        const int sx = in_x ? shell_index(i, p.xlo, p.xhi) : 0;
        const int sy = in_y ? shell_index(j, p.ylo, p.yhi) : 0;

        const bool do_vx = valid && (i >= 2) && (j >= 2);                    // :978-979
        const bool do_vy = valid && (i <= p.nx - 1) && (j <= p.ny - 1);      // :998-999
        const bool do_vz = valid && (i <= p.nx - 1) && (j >= 2);             // :1033-1034
        const bool edge_ij = (i == 1) || (i == p.nx) || (j == 1) || (j == p.ny);   // :1089-1106
        const bool ebox_ij = valid && (i >= p.npml + 1) && (i <= p.nx - p.npml) &&
                             (j >= p.npml + 1) && (j <= p.ny - p.npml);            // :1144-1145
        const bool src_ij = (i == p.isrc) && (j == p.jsrc);

        double ax = 0, bxc = 0, Kx = 1, axh = 0, bxh = 0, Kxh = 1, ay = 0, by = 0, Ky = 1, ayh = 0, byh = 0, Kyh = 1;
        if (in_x) { ax = p.cx.a[i]; bxc = p.cx.b[i]; axh = p.cx.a_half[i]; bxh = p.cx.b_half[i];
                    if (!KUNIT) { Kx = p.cx.K[i]; Kxh = p.cx.K_half[i]; } }
        if (in_y) { ay = p.cy.a[j]; by = p.cy.b[j]; ayh = p.cy.a_half[j]; byh = p.cy.b_half[j];
                    if (!KUNIT) { Ky = p.cy.K[j]; Kyh = p.cy.K_half[j]; } }

        long long qx = in_x ? ((long long)(kb - 1) * p.ny + (j - 1)) * p.sxp + sx : 0;
        long long qy = in_y ? ((long long)(kb - 1) * p.sy + sy) * pitch + (i - 1) : 0;
        const long long qx_step = (long long)p.ny * p.sxp, qy_step = (long long)p.sy * pitch;
        double m_x3 = 0, m_x4 = 0, m_x5 = 0, m_y3 = 0, m_y4 = 0, m_y5 = 0;
        if (in_x) { m_x3 = p.mx[3][qx]; m_x4 = p.mx[4][qx]; m_x5 = p.mx[5][qx]; }
        if (in_y) { m_y3 = p.my[3][qy]; m_y4 = p.my[4][qy]; m_y5 = p.my[5][qy]; }

        double sxz_m = valid ? p.sxz[q - pl] : 0.0, syz_m = valid ? p.syz[q - pl] : 0.0;   // plane kb-1
        double ekin = 0.0, epot = 0.0;

        for (int n = 0; n < np; ++n, q += pl, qx += qx_step, qy += qy_step) {
            const int k = kb + n;
            const int kg = k + p.koff;
            const uint32_t gc = g + (uint32_t)n, gn = gc + 1;
            const uint32_t sc = gc % (uint32_t)S, sn = gn % (uint32_t)S;

            double n_x3 = 0, n_x4 = 0, n_x5 = 0, n_y3 = 0, n_y4 = 0, n_y5 = 0;
            if (n + 1 < np) {
                if (in_x) { n_x3 = p.mx[3][qx + qx_step]; n_x4 = p.mx[4][qx + qx_step]; n_x5 = p.mx[5][qx + qx_step]; }
                if (in_y) { n_y3 = p.my[3][qy + qy_step]; n_y4 = p.my[4][qy + qy_step]; n_y5 = p.my[5][qy + qy_step]; }
            }
            const bool in_z = valid && ((kg <= p.zlo) || (kg >= p.zhi));
            long long qz = 0;
            double m_z3 = 0, m_z4 = 0, m_z5 = 0, az = 0, bz = 0, Kz = 1, azh = 0, bzh = 0, Kzh = 1;
            if (in_z) {
                qz = ((long long)(shell_index(kg, p.zlo, p.zhi) - p.zbase) * p.ny + (j - 1)) * pitch + (i - 1);
                m_z3 = p.mz[3][qz]; m_z4 = p.mz[4][qz]; m_z5 = p.mz[5][qz];
                az = p.cz.a[kg]; bz = p.cz.b[kg]; azh = p.cz.a_half[kg]; bzh = p.cz.b_half[kg];
                if (!KUNIT) { Kz = p.cz.K[kg]; Kzh = p.cz.K_half[kg]; }
            }

            if (n == 0) mbar_wait(bar0 + 8 * sc, (gc / (uint32_t)S) & 1u);
            mbar_wait(bar0 + 8 * sn, (gn / (uint32_t)S) & 1u);

            const unsigned char *st = gstage0 + (size_t)sc * STAGE_BYTES;
            const unsigned char *stn = gstage0 + (size_t)sn * STAGE_BYTES;
            const double *Txx = (const double *)(st + 0 * G::HALO_BYTES);
            const double *Tyy = (const double *)(st + 1 * G::HALO_BYTES);
            const double *Txy = (const double *)(st + 2 * G::HALO_BYTES);
            const double *Txz = (const double *)(st + 3 * G::HALO_BYTES);
            const double *Tyz = (const double *)(st + 4 * G::HALO_BYTES);
            const double *Tp = (const double *)(st + 5 * G::HALO_BYTES);
            constexpr int PD = G::PLAIN_BYTES / 8;
            const int c = ty * TX + tx;

            const double sxx_c = Txx[ty * W + tx + 2], sxx_im = Txx[ty * W + tx + 1];
            const double syy_c = Tyy[ty * W + tx], syy_jp = Tyy[(ty + 1) * W + tx];
            const double sxy_c = Txy[(ty + 1) * W + tx], sxy_jm = Txy[ty * W + tx], sxy_ip = Txy[(ty + 1) * W + tx + 1];
            const double sxz_c = Txz[ty * W + tx], sxz_ip = Txz[ty * W + tx + 1];
            const double syz_c = Tyz[(ty + 1) * W + tx], syz_jm = Tyz[ty * W + tx];
            const double szz_c = Tp[0 * PD + c];
            const double szz_n = ((const double *)(stn + 5 * G::HALO_BYTES))[c];
            double vx = Tp[1 * PD + c], vy = Tp[2 * PD + c], vz = Tp[3 * PD + c];

            if (kg >= 2) {                                           // k2begin
                if (do_vx) {                                         // :976-996
                    double value_dsigmaxx_dx = (sxx_c - sxx_im) * odx;
                    double value_dsigmaxy_dy = (sxy_c - sxy_jm) * ody;
                    double value_dsigmaxz_dz = (sxz_c - sxz_m) * odz;
                    if (in_x) value_dsigmaxx_dx = cpml_apply<KUNIT>(p.mx[3], qx, m_x3, bxc, ax, Kx, value_dsigmaxx_dx);
                    if (in_y) value_dsigmaxy_dy = cpml_apply<KUNIT>(p.my[3], qy, m_y3, by, ay, Ky, value_dsigmaxy_dy);
                    if (in_z) value_dsigmaxz_dz = cpml_apply<KUNIT>(p.mz[3], qz, m_z3, bz, az, Kz, value_dsigmaxz_dz);
                    vx = dt_r * (value_dsigmaxx_dx + value_dsigmaxy_dy + value_dsigmaxz_dz) + vx;
                }
                if (do_vy) {                                         // :998-1016
                    double value_dsigmaxy_dx = (sxy_ip - sxy_c) * odx;
                    double value_dsigmayy_dy = (syy_jp - syy_c) * ody;
                    double value_dsigmayz_dz = (syz_c - syz_m) * odz;
                    if (in_x) value_dsigmaxy_dx = cpml_apply<KUNIT>(p.mx[4], qx, m_x4, bxh, axh, Kxh, value_dsigmaxy_dx);
                    if (in_y) value_dsigmayy_dy = cpml_apply<KUNIT>(p.my[4], qy, m_y4, byh, ayh, Kyh, value_dsigmayy_dy);
                    if (in_z) value_dsigmayz_dz = cpml_apply<KUNIT>(p.mz[4], qz, m_z4, bz, az, Kz, value_dsigmayz_dz);
                    vy = dt_r * (value_dsigmaxy_dx + value_dsigmayy_dy + value_dsigmayz_dz) + vy;
                }
            }
            if (do_vz && kg <= p.nz - 1) {                           // kminus1end, :1031-1052
                double value_dsigmaxz_dx = (sxz_ip - sxz_c) * odx;
                double value_dsigmayz_dy = (syz_c - syz_jm) * ody;
                double value_dsigmazz_dz = (szz_n - szz_c) * odz;
                if (in_x) value_dsigmaxz_dx = cpml_apply<KUNIT>(p.mx[5], qx, m_x5, bxh, axh, Kxh, value_dsigmaxz_dx);
                if (in_y) value_dsigmayz_dy = cpml_apply<KUNIT>(p.my[5], qy, m_y5, by, ay, Ky, value_dsigmayz_dy);
                if (in_z) value_dsigmazz_dz = cpml_apply<KUNIT>(p.mz[5], qz, m_z5, bzh, azh, Kzh, value_dsigmazz_dz);
                vz = dt_r * (value_dsigmaxz_dx + value_dsigmayz_dy + value_dsigmazz_dz) + vz;
            }

            // source, :1080-1081 (after the update of step it, before Dirichlet; quirk B10)
            if (src_ij && k == p.ksrc) {
                vx = vx + p.src_x[p.it - 1];
                vy = vy + p.src_y[p.it - 1];
            }
            // Dirichlet on the six faces, :1087-1121
            if (edge_ij || kg == 1 || kg == p.nz) { vx = 0.0; vy = 0.0; vz = 0.0; }

            if (valid) {
                st_stream(p.vx + q, vx);
                st_stream(p.vy + q, vy);
                st_stream(p.vz + q, vz);
                // boundary planes go straight into the neighbour slabs' halo planes (:811-823)
                const long long qp = (long long)(j - 1) * pitch + (i - 1);
                if (k == 1 && p.peer_lo[0]) { p.peer_lo[0][qp] = vx; p.peer_lo[1][qp] = vy; }   // -> left
                if (k == p.nzl && p.peer_hi[0]) p.peer_hi[0][qp] = vz;                          // -> right
            }

            // energy over the PML-free box, :1131-1177
            if (ebox_ij && kg >= p.npml + 1 && kg <= p.nz - p.npml) {
                ekin += half_rho * (vx * vx + vy * vy + vz * vz);
                const double epsilon_xx = (c2lm * sxx_c - lam * syy_c - lam * szz_c) * inv_den;
                const double epsilon_yy = (c2lm * syy_c - lam * sxx_c - lam * szz_c) * inv_den;
                const double epsilon_zz = (c2lm * szz_c - lam * sxx_c - lam * syy_c) * inv_den;
                const double epsilon_xy = sxy_c * inv_2mu;
                const double epsilon_xz = sxz_c * inv_2mu;
                const double epsilon_yz = syz_c * inv_2mu;
                // quirk B2 (:1169-1172): the reference adds epsilon_yy*sigmayy twice and never
                // epsilon_zz*sigmazz
                const double third = p.energy_bug_compat ? epsilon_yy * syy_c : epsilon_zz * szz_c;
                epot += 0.5 * (epsilon_xx * sxx_c + epsilon_yy * syy_c + third +
                               2.0 * epsilon_xy * sxy_c + 2.0 * epsilon_xz * sxz_c +
                               2.0 * epsilon_yz * syz_c);
            }
            sxz_m = sxz_c; syz_m = syz_c;
            m_x3 = n_x3; m_x4 = n_x4; m_x5 = n_x5; m_y3 = n_y3; m_y4 = n_y4; m_y5 = n_y5;

            __syncthreads();
            if (tid == 0 && n + S <= np) issue(n + S);
        }
        g += (uint32_t)(np + 1);

        block_sum2<NT>(ekin, epot, red);
        if (tid == 0) {
            p.partials[item] = ekin;
            p.partials[p.nblocks + item] = epot;
        }
    }
}

// ---- launch dispatch ---------------------------------------------------------------

template <int TX, int TY>
static size_t smem_need(bool stress, int stages)
{
    using G = TileGeom<TX, TY>;
    const size_t stage = stress ? 3 * G::HALO_BYTES + 6 * G::PLAIN_BYTES : 5 * G::HALO_BYTES + 4 * G::PLAIN_BYTES;
    return kBarBytes + 128 + stage * (size_t)stages;
}

template <bool KUNIT, int TX, int TY, int MINB>
static cudaError_t launch_tile(const Params3D &p, const TmaMaps &tm, const Tile3D &t, cudaStream_t s, bool stress, int *occ)
{
    const size_t smem = smem_need<TX, TY>(stress, t.stages);
    const void *fn = stress ? (const void *)k_stress3d_tma<KUNIT, TX, TY, MINB> : (const void *)k_velocity3d_tma<KUNIT, TX, TY, MINB>;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (occ) {
        if (stress) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k_stress3d_tma<KUNIT, TX, TY, MINB>, TX * TY, smem);
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k_velocity3d_tma<KUNIT, TX, TY, MINB>, TX * TY, smem);
    }
    const dim3 grid(stress ? t.grid_stress : t.grid_velocity), block(TX, TY);
    if (stress) k_stress3d_tma<KUNIT, TX, TY, MINB><<<grid, block, smem, s>>>(p, tm, t);
    else        k_velocity3d_tma<KUNIT, TX, TY, MINB><<<grid, block, smem, s>>>(p, tm, t);
    return cudaGetLastError();
}

// MINB (minimum resident CTAs per SM) caps the registers: 65536 / (threads * MINB).
template <bool KUNIT>
static cudaError_t dispatch_tile(const Params3D &p, const TmaMaps &tm, const Tile3D &t, cudaStream_t s, bool stress, int *occ)
{
    switch (t.tx * 1000 + t.ty * 10 + t.minb) {
    case 32081:  case 32082:  return launch_tile<KUNIT, 32, 8, 2>(p, tm, t, s, stress, occ);
    case 64041:  case 64042:  return launch_tile<KUNIT, 64, 4, 2>(p, tm, t, s, stress, occ);
    case 128021: case 128022: return launch_tile<KUNIT, 128, 2, 2>(p, tm, t, s, stress, occ);
    case 32083:  return launch_tile<KUNIT, 32, 8, 3>(p, tm, t, s, stress, occ);
    case 64043:  return launch_tile<KUNIT, 64, 4, 3>(p, tm, t, s, stress, occ);
    case 128023: return launch_tile<KUNIT, 128, 2, 3>(p, tm, t, s, stress, occ);
    case 64081:  return launch_tile<KUNIT, 64, 8, 1>(p, tm, t, s, stress, occ);
    case 64082:  return launch_tile<KUNIT, 64, 8, 2>(p, tm, t, s, stress, occ);
    case 104041: return launch_tile<KUNIT, 104, 4, 1>(p, tm, t, s, stress, occ);
    case 104042: return launch_tile<KUNIT, 104, 4, 2>(p, tm, t, s, stress, occ);
    case 128041: return launch_tile<KUNIT, 128, 4, 1>(p, tm, t, s, stress, occ);
    case 128042: return launch_tile<KUNIT, 128, 4, 2>(p, tm, t, s, stress, occ);
    default: return cudaErrorInvalidValue;
    }
}

bool tma_tile_supported(int tx, int ty)
{
    switch (tx * 100 + ty) {
    case 3208: case 6404: case 6408: case 10404: case 12802: case 12804: return true;
    default: return false;
    }
}

cudaError_t tma_occupancy(const Params3D &p, const Tile3D &t, bool stress, int *occ)
{
    TmaMaps dummy{};
    return p.kunit ? dispatch_tile<true>(p, dummy, t, nullptr, stress, occ) : dispatch_tile<false>(p, dummy, t, nullptr, stress, occ);
}

cudaError_t launch_stress3d_tma(const Params3D &p, const TmaMaps &tm, const Tile3D &t, cudaStream_t s)
{
    return p.kunit ? dispatch_tile<true>(p, tm, t, s, true, nullptr) : dispatch_tile<false>(p, tm, t, s, true, nullptr);
}
cudaError_t launch_velocity3d_tma(const Params3D &p, const TmaMaps &tm, const Tile3D &t, cudaStream_t s)
{
    return p.kunit ? dispatch_tile<true>(p, tm, t, s, false, nullptr) : dispatch_tile<false>(p, tm, t, s, false, nullptr);
}

// ---- slab-to-slab flags ---------------------------------------------------------------
// k_signal: after a kernel whose boundary planes went into the neighbours' halo planes,
// publish "step `value` of this phase is there" in each neighbour's flag word (system scope:
// the stores of the preceding kernel on this stream are complete, the fence orders them
// before the flag for the peer).  k_wait: spin until both neighbours have published `value`.
__global__ void k_signal(unsigned long long *flag_lo, unsigned long long *flag_hi, unsigned long long value)
{
    __threadfence_system();
    if (flag_lo) *(volatile unsigned long long *)flag_lo = value;
    if (flag_hi) *(volatile unsigned long long *)flag_hi = value;
    __threadfence_system();
}

__global__ void k_wait(const unsigned long long *flag_a, const unsigned long long *flag_b, unsigned long long value,
                       unsigned int *timeout_flag)
{
    const unsigned long long *f[2] = {flag_a, flag_b};
    for (int q = 0; q < 2; q++) {
        if (!f[q]) continue;
        unsigned long long spins = 0;
        while (true) {
            unsigned long long v;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f[q]) : "memory");
            if (v >= value) break;
            if (++spins > (1ull << 28)) { *timeout_flag = 1u; return; }   // ~ tens of seconds: neighbour is gone
            __nanosleep(64);
        }
    }
    __threadfence_system();
}

void launch_signal(unsigned long long *flag_lo, unsigned long long *flag_hi, unsigned long long value, cudaStream_t s)
{
    k_signal<<<1, 1, 0, s>>>(flag_lo, flag_hi, value);
}
void launch_wait(const unsigned long long *flag_a, const unsigned long long *flag_b, unsigned long long value,
                 unsigned int *timeout_flag, cudaStream_t s)
{
    k_wait<<<1, 1, 0, s>>>(flag_a, flag_b, value, timeout_flag);
}

}  // namespace cpml
