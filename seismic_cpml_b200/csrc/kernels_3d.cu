// 3-D isotropic C-PML kernels for sm_100a.
//
// Two fused kernels per time step replace the five OpenMP loop nests, the Dirichlet
// pass and the energy pass of seismic_CPML_3D_isotropic_MPI_OpenMP.f90:825-1177:
//
//   k_stress3d    sigmaxx/yy/zz (:836-863), sigmaxy (:877-894), sigmaxz/yz (:908-943)
//   k_velocity3d  vx/vy (:976-1017), vz (:1031-1052), source (:1055-1083),
//                 Dirichlet faces (:1087-1121), per-block energy partials (:1131-1177)
//   k_post3d      energy sum of the step (:1179) + seismogram sample (:1124-1129)
//
// Mapping: one thread per (i,j) column, x across the warp (coalesced 256 B rows), each
// block marches a chunk of z planes.  Values reused along z (vx,vy at k / k+1, vz at
// k-1 / k; sigmaxz, sigmayz at k-1 / k, sigmazz at k / k+1) stay in registers, so every
// field plane is fetched from HBM once per kernel; in-plane neighbours (radius 1) are
// re-read through L1, which holds them from the loads of the neighbouring threads.
// Streamed read-modify-write traffic (the six stresses in k_stress3d, the three
// velocities in k_velocity3d) uses evict-first loads/stores so that it does not push
// the reused planes out of L1.  C-PML memory variables are touched only by threads
// inside a PML shell; elsewhere the reference recursion is the identity (a = 0, K = 1,
// memory variable == 0), so skipping it is exact.
//
// Compiled with -fmad=false: every product and sum is rounded separately, in the order
// the Fortran source writes them, so fields are bit-identical to an IEEE (non-FMA)
// build of the reference loops.
#include "cpml_internal.h"

namespace cpml {

__device__ __forceinline__ double ld_stream(const double *p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(double *p, double v) { __stcs(p, v); }

// memory_x = b * memory_x + a * value ; value = value / K + memory_x   (e.g. :845-851)
__device__ __forceinline__ double cpml_apply(double *__restrict__ mem, long long q,
                                             double b, double a, double K, double value)
{
    double m = mem[q];
    m = b * m + a * value;
    mem[q] = m;
    return value / K + m;
}

__device__ __forceinline__ int shell_index(int i, int lo, int hi)
{
    return i <= lo ? i - 1 : lo + (i - hi);
}

// mx: 0 dvx_dx  1 dvy_dx  2 dvz_dx  3 dsigmaxx_dx  4 dsigmaxy_dx  5 dsigmaxz_dx
// my: 0 dvy_dy  1 dvx_dy  2 dvz_dy  3 dsigmaxy_dy  4 dsigmayy_dy  5 dsigmayz_dy
// mz: 0 dvz_dz  1 dvx_dz  2 dvy_dz  3 dsigmaxz_dz  4 dsigmayz_dz  5 dsigmazz_dz

template <int TX, int TY>
__global__ void __launch_bounds__(TX *TY)
k_stress3d(const __grid_constant__ Params3D p)
{
    const int i = blockIdx.x * TX + threadIdx.x + 1;
    const int j = blockIdx.y * TY + threadIdx.y + 1;
    if (i > p.nx || j > p.ny) return;
    const int kb = 1 + blockIdx.z * p.kchunk;
    const int ke = min(p.nzl, kb + p.kchunk - 1);
    const int pitch = p.pitch;
    const long long pl = p.plane;
    long long q = (long long)kb * pl + (long long)(j - 1) * pitch + (i - 1);

    const bool in_x = (i <= p.xlo) || (i >= p.xhi);
    const bool in_y = (j <= p.ylo) || (j >= p.yhi);
    const int sx = in_x ? shell_index(i, p.xlo, p.xhi) : 0;
    const int sy = in_y ? shell_index(j, p.ylo, p.yhi) : 0;

    // loop bounds of the four nests (i, j part; the k part is tested per plane)
    const bool do_n = (i <= p.nx - 1) && (j >= 2);     // :838-839
    const bool do_xy = (i >= 2) && (j <= p.ny - 1);    // :878-879
    const bool do_xz = (i >= 2);                       // :910-911
    const bool do_yz = (j <= p.ny - 1);                // :927-928

    const double odx = p.odx, ody = p.ody, odz = p.odz;
    const double dt_l = p.dt_lambda, dt_m = p.dt_mu, dt_l2m = p.dt_lambdaplus2mu;

    double vx_c = p.vx[q], vy_c = p.vy[q];
    double vz_m = p.vz[q - pl], vz_c = p.vz[q];

    for (int k = kb; k <= ke; ++k, q += pl) {
        const int kg = k + p.koff;                      // :837
        // next plane (kept in registers for the next iteration)
        const double vx_n = p.vx[q + pl];
        const double vy_n = p.vy[q + pl];
        const double vz_n = p.vz[q + pl];
        // in-plane neighbours (L1)
        const double vx_ip = p.vx[q + 1], vx_jp = p.vx[q + pitch];
        const double vy_im = p.vy[q - 1], vy_jm = p.vy[q - pitch];
        const double vz_im = p.vz[q - 1], vz_jp = p.vz[q + pitch];

        const bool in_z = (kg <= p.zlo) || (kg >= p.zhi);
        const long long qx = ((long long)(k - 1) * p.ny + (j - 1)) * p.sxp + sx;
        const long long qy = ((long long)(k - 1) * p.sy + sy) * pitch + (i - 1);
        const long long qz = in_z ? ((long long)(shell_index(kg, p.zlo, p.zhi) - p.zbase) * p.ny + (j - 1)) * pitch + (i - 1) : 0;

        if (do_n && kg >= 2) {                          // k2begin, :792-793
            double value_dvx_dx = (vx_ip - vx_c) * odx;
            double value_dvy_dy = (vy_c - vy_jm) * ody;
            double value_dvz_dz = (vz_c - vz_m) * odz;
            if (in_x) value_dvx_dx = cpml_apply(p.mx[0], qx, p.cx.b_half[i], p.cx.a_half[i], p.cx.K_half[i], value_dvx_dx);
            if (in_y) value_dvy_dy = cpml_apply(p.my[0], qy, p.cy.b[j], p.cy.a[j], p.cy.K[j], value_dvy_dy);
            if (in_z) value_dvz_dz = cpml_apply(p.mz[0], qz, p.cz.b[kg], p.cz.a[kg], p.cz.K[kg], value_dvz_dz);
            const double sxx = ld_stream(p.sxx + q), syy = ld_stream(p.syy + q), szz = ld_stream(p.szz + q);
            st_stream(p.sxx + q, dt_l2m * value_dvx_dx + dt_l * (value_dvy_dy + value_dvz_dz) + sxx);
            st_stream(p.syy + q, dt_l * (value_dvx_dx + value_dvz_dz) + dt_l2m * value_dvy_dy + syy);
            st_stream(p.szz + q, dt_l * (value_dvx_dx + value_dvy_dy) + dt_l2m * value_dvz_dz + szz);
        }
        if (do_xy) {
            double value_dvy_dx = (vy_c - vy_im) * odx;
            double value_dvx_dy = (vx_jp - vx_c) * ody;
            if (in_x) value_dvy_dx = cpml_apply(p.mx[1], qx, p.cx.b[i], p.cx.a[i], p.cx.K[i], value_dvy_dx);
            if (in_y) value_dvx_dy = cpml_apply(p.my[1], qy, p.cy.b_half[j], p.cy.a_half[j], p.cy.K_half[j], value_dvx_dy);
            const double sxy = ld_stream(p.sxy + q);
            st_stream(p.sxy + q, dt_m * (value_dvy_dx + value_dvx_dy) + sxy);
        }
        if (kg <= p.nz - 1) {                           // kminus1end, :795-796
            if (do_xz) {
                double value_dvz_dx = (vz_c - vz_im) * odx;
                double value_dvx_dz = (vx_n - vx_c) * odz;
                if (in_x) value_dvz_dx = cpml_apply(p.mx[2], qx, p.cx.b[i], p.cx.a[i], p.cx.K[i], value_dvz_dx);
                if (in_z) value_dvx_dz = cpml_apply(p.mz[1], qz, p.cz.b_half[kg], p.cz.a_half[kg], p.cz.K_half[kg], value_dvx_dz);
                const double sxz = ld_stream(p.sxz + q);
                st_stream(p.sxz + q, dt_m * (value_dvz_dx + value_dvx_dz) + sxz);
            }
            if (do_yz) {
                double value_dvz_dy = (vz_jp - vz_c) * ody;
                double value_dvy_dz = (vy_n - vy_c) * odz;
                if (in_y) value_dvz_dy = cpml_apply(p.my[2], qy, p.cy.b_half[j], p.cy.a_half[j], p.cy.K_half[j], value_dvz_dy);
                if (in_z) value_dvy_dz = cpml_apply(p.mz[2], qz, p.cz.b_half[kg], p.cz.a_half[kg], p.cz.K_half[kg], value_dvy_dz);
                const double syz = ld_stream(p.syz + q);
                st_stream(p.syz + q, dt_m * (value_dvz_dy + value_dvy_dz) + syz);
            }
        }
        vx_c = vx_n; vy_c = vy_n; vz_m = vz_c; vz_c = vz_n;
    }
}

template <int NT>
__device__ __forceinline__ void block_sum2(double &a, double &b, double *smem /* 2*NT/32 */)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, o);
        b += __shfl_down_sync(0xffffffffu, b, o);
    }
    const int t = threadIdx.y * blockDim.x + threadIdx.x;
    const int w = t >> 5, l = t & 31;
    if (l == 0) { smem[w] = a; smem[NT / 32 + w] = b; }
    __syncthreads();
    if (w == 0) {
        a = (l < NT / 32) ? smem[l] : 0.0;
        b = (l < NT / 32) ? smem[NT / 32 + l] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_down_sync(0xffffffffu, a, o);
            b += __shfl_down_sync(0xffffffffu, b, o);
        }
    }
}

template <int TX, int TY>
__global__ void __launch_bounds__(TX *TY)
k_velocity3d(const __grid_constant__ Params3D p)
{
    __shared__ double red[2 * TX * TY / 32];
    const int i = blockIdx.x * TX + threadIdx.x + 1;
    const int j = blockIdx.y * TY + threadIdx.y + 1;
    const bool active = (i <= p.nx) && (j <= p.ny);
    double ekin = 0.0, epot = 0.0;

    if (active) {
        const int kb = 1 + blockIdx.z * p.kchunk;
        const int ke = min(p.nzl, kb + p.kchunk - 1);
        const int pitch = p.pitch;
        const long long pl = p.plane;
        long long q = (long long)kb * pl + (long long)(j - 1) * pitch + (i - 1);

        const bool in_x = (i <= p.xlo) || (i >= p.xhi);
        const bool in_y = (j <= p.ylo) || (j >= p.yhi);
        const int sx = in_x ? shell_index(i, p.xlo, p.xhi) : 0;
        const int sy = in_y ? shell_index(j, p.ylo, p.yhi) : 0;

        const bool do_vx = (i >= 2) && (j >= 2);                    // :978-979
        const bool do_vy = (i <= p.nx - 1) && (j <= p.ny - 1);      // :998-999
        const bool do_vz = (i <= p.nx - 1) && (j >= 2);             // :1033-1034
        const bool edge_ij = (i == 1) || (i == p.nx) || (j == 1) || (j == p.ny);   // :1089-1106
        const bool ebox_ij = (i >= p.npml + 1) && (i <= p.nx - p.npml) &&
                             (j >= p.npml + 1) && (j <= p.ny - p.npml);            // :1144-1145
        const bool src_ij = (i == p.isrc) && (j == p.jsrc);

        const double odx = p.odx, ody = p.ody, odz = p.odz, dt_r = p.dt_over_rho;
        // energy constants (:1159-1167); reciprocals instead of the reference's
        // divisions -- the energy sum is reduction-order dependent anyway (quirk B11)
        const double lam = p.lambda, mu = p.mu;
        const double c2lm = 2.0 * (lam + mu);
        const double inv_den = 1.0 / (2.0 * mu * (3.0 * lam + 2.0 * mu));
        const double inv_2mu = 1.0 / (2.0 * mu);
        const double half_rho = 0.5 * p.rho;

        double sxz_m = p.sxz[q - pl], syz_m = p.syz[q - pl];
        double szz_c = p.szz[q];

        for (int k = kb; k <= ke; ++k, q += pl) {
            const int kg = k + p.koff;
            const double szz_n = p.szz[q + pl];
            const double sxx_c = p.sxx[q], sxx_im = p.sxx[q - 1];
            const double syy_c = p.syy[q], syy_jp = p.syy[q + pitch];
            const double sxy_c = p.sxy[q], sxy_jm = p.sxy[q - pitch], sxy_ip = p.sxy[q + 1];
            const double sxz_c = p.sxz[q], sxz_ip = p.sxz[q + 1];
            const double syz_c = p.syz[q], syz_jm = p.syz[q - pitch];
            double vx = ld_stream(p.vx + q), vy = ld_stream(p.vy + q), vz = ld_stream(p.vz + q);

            const bool in_z = (kg <= p.zlo) || (kg >= p.zhi);
            const long long qx = ((long long)(k - 1) * p.ny + (j - 1)) * p.sxp + sx;
            const long long qy = ((long long)(k - 1) * p.sy + sy) * pitch + (i - 1);
            const long long qz = in_z ? ((long long)(shell_index(kg, p.zlo, p.zhi) - p.zbase) * p.ny + (j - 1)) * pitch + (i - 1) : 0;

            if (kg >= 2) {                                           // k2begin
                if (do_vx) {
                    double value_dsigmaxx_dx = (sxx_c - sxx_im) * odx;
                    double value_dsigmaxy_dy = (sxy_c - sxy_jm) * ody;
                    double value_dsigmaxz_dz = (sxz_c - sxz_m) * odz;
                    if (in_x) value_dsigmaxx_dx = cpml_apply(p.mx[3], qx, p.cx.b[i], p.cx.a[i], p.cx.K[i], value_dsigmaxx_dx);
                    if (in_y) value_dsigmaxy_dy = cpml_apply(p.my[3], qy, p.cy.b[j], p.cy.a[j], p.cy.K[j], value_dsigmaxy_dy);
                    if (in_z) value_dsigmaxz_dz = cpml_apply(p.mz[3], qz, p.cz.b[kg], p.cz.a[kg], p.cz.K[kg], value_dsigmaxz_dz);
                    vx = dt_r * (value_dsigmaxx_dx + value_dsigmaxy_dy + value_dsigmaxz_dz) + vx;
                }
                if (do_vy) {
                    double value_dsigmaxy_dx = (sxy_ip - sxy_c) * odx;
                    double value_dsigmayy_dy = (syy_jp - syy_c) * ody;
                    double value_dsigmayz_dz = (syz_c - syz_m) * odz;
                    if (in_x) value_dsigmaxy_dx = cpml_apply(p.mx[4], qx, p.cx.b_half[i], p.cx.a_half[i], p.cx.K_half[i], value_dsigmaxy_dx);
                    if (in_y) value_dsigmayy_dy = cpml_apply(p.my[4], qy, p.cy.b_half[j], p.cy.a_half[j], p.cy.K_half[j], value_dsigmayy_dy);
                    if (in_z) value_dsigmayz_dz = cpml_apply(p.mz[4], qz, p.cz.b[kg], p.cz.a[kg], p.cz.K[kg], value_dsigmayz_dz);
                    vy = dt_r * (value_dsigmaxy_dx + value_dsigmayy_dy + value_dsigmayz_dz) + vy;
                }
            }
            if (do_vz && kg <= p.nz - 1) {                           // kminus1end
                double value_dsigmaxz_dx = (sxz_ip - sxz_c) * odx;
                double value_dsigmayz_dy = (syz_c - syz_jm) * ody;
                double value_dsigmazz_dz = (szz_n - szz_c) * odz;
                if (in_x) value_dsigmaxz_dx = cpml_apply(p.mx[5], qx, p.cx.b_half[i], p.cx.a_half[i], p.cx.K_half[i], value_dsigmaxz_dx);
                if (in_y) value_dsigmayz_dy = cpml_apply(p.my[5], qy, p.cy.b[j], p.cy.a[j], p.cy.K[j], value_dsigmayz_dy);
                if (in_z) value_dsigmazz_dz = cpml_apply(p.mz[5], qz, p.cz.b_half[kg], p.cz.a_half[kg], p.cz.K_half[kg], value_dsigmazz_dz);
                vz = dt_r * (value_dsigmaxz_dx + value_dsigmayz_dy + value_dsigmazz_dz) + vz;
            }

            // source, :1080-1081 (after the update of step it, before Dirichlet; quirk B10)
            if (src_ij && k == p.ksrc) {
                vx = vx + p.src_x[p.it - 1];
                vy = vy + p.src_y[p.it - 1];
            }
            // Dirichlet on the six faces, :1087-1121
            if (edge_ij || kg == 1 || kg == p.nz) { vx = 0.0; vy = 0.0; vz = 0.0; }

            st_stream(p.vx + q, vx);
            st_stream(p.vy + q, vy);
            st_stream(p.vz + q, vz);

            // energy over the PML-free box, :1131-1177
            if (ebox_ij && kg >= p.npml + 1 && kg <= p.nz - p.npml) {
                ekin += half_rho * (vx * vx + vy * vy + vz * vz);
                const double epsilon_xx = (c2lm * sxx_c - lam * syy_c - lam * szz_c) * inv_den;
                const double epsilon_yy = (c2lm * syy_c - lam * sxx_c - lam * szz_c) * inv_den;
                const double epsilon_zz = (c2lm * szz_c - lam * sxx_c - lam * syy_c) * inv_den;
                const double epsilon_xy = sxy_c * inv_2mu;
                const double epsilon_xz = sxz_c * inv_2mu;
                const double epsilon_yz = syz_c * inv_2mu;
                // quirk B2 (:1169-1172): the reference adds epsilon_yy*sigmayy twice and
                // never epsilon_zz*sigmazz
                const double third = p.energy_bug_compat ? epsilon_yy * syy_c : epsilon_zz * szz_c;
                epot += 0.5 * (epsilon_xx * sxx_c + epsilon_yy * syy_c + third +
                               2.0 * epsilon_xy * sxy_c + 2.0 * epsilon_xz * sxz_c +
                               2.0 * epsilon_yz * syz_c);
            }
            sxz_m = sxz_c; syz_m = syz_c; szz_c = szz_n;
        }
    }

    block_sum2<TX * TY>(ekin, epot, red);
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        const int b = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        p.partials[b] = ekin;
        p.partials[p.nblocks + b] = epot;
    }
}

// One block: fixed-order sum of the per-block energy partials (:1179) and the
// seismogram sample of this step (:1124-1129).
__global__ void __launch_bounds__(256) k_post3d(const __grid_constant__ Post3D p)
{
    __shared__ double red[16];
    double a = 0.0, b = 0.0;
    for (int q = threadIdx.x; q < p.nblocks; q += 256) {
        a += p.partials[q];
        b += p.partials[p.nblocks + q];
    }
    block_sum2<256>(a, b, red);
    if (threadIdx.x == 0) {
        p.energy_k[p.it - 1] = a;
        p.energy_p[p.it - 1] = b;
    }
    if (p.krec > 0) {
        for (int r = threadIdx.x; r < p.nrec; r += 256) {
            const long long q = (long long)p.krec * p.plane + (long long)(p.iy_rec[r] - 1) * p.pitch + (p.ix_rec[r] - 1);
            p.sisvx[(long long)r * p.nstep + (p.it - 1)] = p.vx[q];
            p.sisvy[(long long)r * p.nstep + (p.it - 1)] = p.vy[q];
        }
    }
}

// max over n contiguous padded elements of sqrt(vx^2+vy^2(+vz^2)) (:1185); the padding
// holds zeros.  Non-negative doubles order like their bit patterns.
__global__ void __launch_bounds__(256) k_maxnorm(const double *vx, const double *vy, const double *vz,
                                                  long long n, unsigned long long *out_bits)
{
    double m = 0.0;
    for (long long q = (long long)blockIdx.x * 256 + threadIdx.x; q < n; q += (long long)gridDim.x * 256) {
        const double a = vx[q], b = vy[q], c = vz ? vz[q] : 0.0;
        const double v = sqrt(a * a + b * b + c * c);
        m = v > m ? v : m;     // NaN never wins; the driver's threshold test sees Inf
        if (v != v) m = __longlong_as_double(0x7ff0000000000000LL);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double t = __shfl_down_sync(0xffffffffu, m, o);
        m = t > m ? t : m;
    }
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, (unsigned long long)__double_as_longlong(m));
}

template <int TX, int TY>
static void launch_pair(const Params3D &p, dim3 grid, cudaStream_t s, bool stress)
{
    if (stress) k_stress3d<TX, TY><<<grid, dim3(TX, TY), 0, s>>>(p);
    else        k_velocity3d<TX, TY><<<grid, dim3(TX, TY), 0, s>>>(p);
}

static void dispatch3d(const Params3D &p, dim3 grid, dim3 block, cudaStream_t s, bool stress)
{
    const int key = block.x * 100 + block.y;
    switch (key) {
    case 32 * 100 + 4:  launch_pair<32, 4>(p, grid, s, stress); break;
    case 32 * 100 + 8:  launch_pair<32, 8>(p, grid, s, stress); break;
    case 32 * 100 + 16: launch_pair<32, 16>(p, grid, s, stress); break;
    case 64 * 100 + 2:  launch_pair<64, 2>(p, grid, s, stress); break;
    case 64 * 100 + 4:  launch_pair<64, 4>(p, grid, s, stress); break;
    case 64 * 100 + 8:  launch_pair<64, 8>(p, grid, s, stress); break;
    case 128 * 100 + 1: launch_pair<128, 1>(p, grid, s, stress); break;
    case 128 * 100 + 2: launch_pair<128, 2>(p, grid, s, stress); break;
    case 128 * 100 + 4: launch_pair<128, 4>(p, grid, s, stress); break;
    case 16 * 100 + 16: launch_pair<16, 16>(p, grid, s, stress); break;
    case 16 * 100 + 8:  launch_pair<16, 8>(p, grid, s, stress); break;
    default:            launch_pair<32, 8>(p, dim3((p.nx + 31) / 32, (p.ny + 7) / 8, grid.z), s, stress); break;
    }
}

void launch_stress3d(const Params3D &p, dim3 grid, dim3 block, cudaStream_t s) { dispatch3d(p, grid, block, s, true); }
void launch_velocity3d(const Params3D &p, dim3 grid, dim3 block, cudaStream_t s) { dispatch3d(p, grid, block, s, false); }
void launch_post3d(const Post3D &p, cudaStream_t s) { k_post3d<<<1, 256, 0, s>>>(p); }
void launch_maxnorm(const double *vx, const double *vy, const double *vz, long long n,
                    unsigned long long *out_bits, cudaStream_t s)
{
    long long nb = (n + 255) / 256;
    if (nb > 148 * 16) nb = 148 * 16;
    k_maxnorm<<<(int)nb, 256, 0, s>>>(vx, vy, vz, n, out_bits);
}

}  // namespace cpml
