// 2-D isotropic C-PML kernels for sm_100a, TMA-staged with y-marching and a producer warp (the default 2-D path;
// kernels_2d.cu keeps the one-launch-per-point kernels for A/B runs, CPML_2D_KERNEL=pair).
//
//   k_stress2d_ws<ORDER>    sigma_xx/yy (2D-2nd :556-580, 2D-4th :557-581), sigma_xy (:582-600 / :583-601), potential energy
//   k_velocity2d_ws<ORDER>  vx, vy (2D-2nd :606-641), source (:643-667), Dirichlet (:669-680), kinetic energy (:688-713)
//
// Why.  The pair kernels of kernels_2d.cu read every fourth-order tap through L1 (19 16-byte loads per pair of points)
// and run at 50-64 % of the measured HBM bandwidth at 4096 x 4096, bound by dependent address / FP64 instructions.
// Here a CTA owns a strip of TX columns and MARCHES along y in blocks of RB rows.  A ring of SLOTS shared-memory stages
// holds consecutive row blocks, filled by a dedicated producer warp through the TMA unit (one box per field and block;
// the tensors are the padded arrays, so the ghost ring's zeros come from memory): the y taps of block n are simply the
// rows of blocks n-1 and n+1 that are resident anyway, so EVERY FIELD ROW IS FETCHED FROM HBM EXACTLY ONCE; only the x
// halo (two columns each side of a 64-column strip) is fetched twice.  Consumers never meet in a CTA barrier inside
// the marching loop: they wait on the block's "full" mbarrier, read their taps, ARRIVE (bar.arrive) on a named barrier
// and compute; the producer SYNCs on that barrier and refills the stage block n-1 occupied.
//
// The arithmetic of a point is kernels_2d.cu's stress_point2 / velocity_point2 / epot_point2 (same operations, same
// order): fields are bit-identical to the pair kernels and to the oracle.
#include <cstdlib>

#include "kernels_2d_point.cuh"
#include "tma_common.cuh"

namespace cpml {

namespace {

constexpr int k2ConsBar = 1;
constexpr int k2RelBar0 = 2;         // named barriers k2RelBar0 + (iteration & 3)

__device__ __forceinline__ void bar2_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar2_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

template <int TX, int RB>
struct Geom2 {
    static constexpr int W = TX + 4;                               // tap boxes start two columns before the strip
    static constexpr int TAP = round128(W * RB * 8), PLAIN = round128(TX * RB * 8);
    static constexpr int TAP_BOX = W * RB * 8, PLAIN_BOX = TX * RB * 8;
};

template <int NC>
__device__ __forceinline__ double cons2_sum(double a, double *red, int tid)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
    constexpr int NW = NC / 32;
    const int w = tid >> 5, l = tid & 31;
    if (l == 0) red[w] = a;
    bar2_sync(k2ConsBar, NC);
    if (w == 0) {
        a = (l < NW) ? red[l] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
    }
    bar2_sync(k2ConsBar, NC);
    return a;
}

// Row rr (may be -2 .. RB+1) of a boxed field in the three resident stages (previous, current, next block).
struct Stages3 { const unsigned char *m, *c, *n; };       // previous, current, next block

template <int RB>
__device__ __forceinline__ const double *row3(const Stages3 &st, int off, int W, int rr)
{
    const unsigned char *b = rr < 0 ? st.m : (rr >= RB ? st.n : st.c);      // selects, no local-memory indexing
    const int r = rr < 0 ? rr + RB : (rr >= RB ? rr - RB : rr);
    return reinterpret_cast<const double *>(b + off) + r * W;
}

}  // namespace

// Work item = (x strip, y chunk).  Blocks -1 .. nb of a chunk are loaded (block -1 and block nb only feed y taps, their
// read-modify-write boxes are skipped); ring position and barrier counter run on across items.
//
// ---------------------------------------------------------------------------------------------------- stress
// stage: vx, vy, lambda, mu as (TX+4) x RB tap boxes; sigma_xx, sigma_yy, sigma_xy as TX x RB boxes
// maps: 0 vx 1 vy 2 lambda 3 mu 4 sxx 5 syy 6 sxy
template <int ORDER, int TX, int RB, int SLOTS, int MINB>
__global__ void __launch_bounds__((TX / 2) * RB + 32, MINB)
k_stress2d_ws(const __grid_constant__ Params2D p, const __grid_constant__ TmaMaps tm, const __grid_constant__ Tile2D t)
{
    using G = Geom2<TX, RB>;
    constexpr int W = G::W;
    constexpr int NC = (TX / 2) * RB, NALL = NC + 32;
    constexpr int STAGE = 4 * G::TAP + 3 * G::PLAIN;
    constexpr int O_VX = 0, O_VY = G::TAP, O_LAM = 2 * G::TAP, O_MU = 3 * G::TAP, O_SXX = 4 * G::TAP, O_SYY = O_SXX + G::PLAIN,
                  O_SXY = O_SYY + G::PLAIN;
    __shared__ double red[NC / 32 + 1];
    __shared__ unsigned long long item_bar;
    __shared__ int item_slot[2];          // the producer posts the CTA's work items here (tma_common.cuh: claim_item)
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t sbase = (smem_u32(smem_dyn) + 127u) & ~127u;
    const unsigned char *gbase = smem_dyn + (sbase - smem_u32(smem_dyn));
    const uint32_t bar = sbase, ring = sbase + kBarBytes;
    const unsigned char *gring = gbase + kBarBytes;

    const uint32_t barI = smem_u32(&item_bar);

    const int tid = (int)threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < SLOTS; s++) mbar_init(bar + 8 * s, 1);
        mbar_init(barI, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= NC) {        // ================================================================ producer warp
        const bool lead = tid == NC;
        uint32_t slot = 0, cnt = 0, ip = 0;
        int item = (int)blockIdx.x;
        while (true) {
            if (lead) { item_slot[ip & 1u] = item; mbar_arrive(barI); }     // slot ip & 1 was last read two items ago
            ++ip;
            if (item >= t.nitems) break;
            const int tix = item % t.ntx, yc = item / t.ntx;
            const int xt = 16 + tix * TX - 2;                     // tensor x of column i0 - 2 (padded arrays: x offset 16)
            const int jb = 1 + yc * t.rows;
            const int je = min(p.ny, jb + t.rows - 1);
            const int nb = (je - jb + 1 + RB - 1) / RB;
            auto issue = [&](uint32_t s, int b) {                 // block b: rows jb + b RB ...
                const uint32_t br = bar + 8 * s, dst = ring + s * STAGE;
                const int yt = 2 + (jb - 1) + b * RB;
                const bool rmw = b >= 0 && b < nb;
                mbar_expect_tx(br, 4 * G::TAP_BOX + (rmw ? 3 * G::PLAIN_BOX : 0));
                tma_load_2d(dst + O_VX, &tm.m[0], xt, yt, br);
                tma_load_2d(dst + O_VY, &tm.m[1], xt, yt, br);
                tma_load_2d(dst + O_LAM, &tm.m[2], xt, yt, br);
                tma_load_2d(dst + O_MU, &tm.m[3], xt, yt, br);
                if (rmw) {
                    tma_load_2d(dst + O_SXX, &tm.m[4], xt + 2, yt, br);
                    tma_load_2d(dst + O_SYY, &tm.m[5], xt + 2, yt, br);
                    tma_load_2d(dst + O_SXY, &tm.m[6], xt + 2, yt, br);
                }
            };
            if (lead) {
                uint32_t s = slot;
                for (int l = 0; l < min(SLOTS, nb + 2); l++) { issue(s, l - 1); if (++s == SLOTS) s = 0; }
            }
            int next = item + (int)gridDim.x;
            if (lead) next = claim_item(t.queue, next);       // the answer is needed after the row-block loop
            // iteration n frees the stage of block n-1; iterations nb, nb+1 free the last two
            for (int n = 0; n < nb + 2; ++n, ++cnt) {
                bar2_sync(k2RelBar0 + (int)(cnt & 3u), NALL);
                if (lead && n - 1 + SLOTS <= nb) issue(slot, n - 1 + SLOTS);
                if (++slot == SLOTS) slot = 0;
            }
            item = __shfl_sync(0xffffffffu, next, 0);
        }
        if (lead) retire_queue(t.queue);
        return;
    }

    // ================================================================ consumer warps
    const int tx = tid % (TX / 2), r = tid / (TX / 2);
    const int cW = 2 * tx + 2, cP = 2 * tx;
    RingPos pos{0, 0};                    // stage of block -1 of the current item
    uint32_t cnt = 0;
    for (uint32_t ip = 0;; ++ip) {
        mbar_wait(barI, ip & 1u);
        const int item = item_slot[ip & 1u];
        if (item >= t.nitems) break;
        const int tix = item % t.ntx, yc = item / t.ntx;
        const int i = 1 + tix * TX + 2 * tx;                      // A; B = i + 1
        const int jb = 1 + yc * t.rows;
        const int je = min(p.ny, jb + t.rows - 1);
        const int nb = (je - jb + 1 + RB - 1) / RB;
        const bool colA = i <= p.nx, validBx = i + 1 <= p.nx;
        const bool in_xA = (i <= p.xlo) || (i >= p.xhi), in_xB = validBx && ((i + 1 <= p.xlo) || (i + 1 >= p.xhi));
        const int pitch = p.pitch;
        double epot = 0.0;

        RingPos pm = pos, pc = pos;
        pc.advance(SLOTS);
        mbar_wait(bar + 8 * pm.s, pm.par);                        // block -1
        mbar_wait(bar + 8 * pc.s, pc.par);                        // block 0
        for (int n = 0; n < nb; ++n, ++cnt) {
            RingPos pn = pc;
            pn.advance(SLOTS);
            mbar_wait(bar + 8 * pn.s, pn.par);                    // block n+1
            const Stages3 st{gring + (size_t)pm.s * STAGE, gring + (size_t)pc.s * STAGE, gring + (size_t)pn.s * STAGE};
            const int j = jb + n * RB + r;
            const bool valid = colA && j <= je;

            const double *vx0 = row3<RB>(st, O_VX, W, r), *vy0 = row3<RB>(st, O_VY, W, r);
            const double *la0 = row3<RB>(st, O_LAM, W, r), *mu0 = row3<RB>(st, O_MU, W, r);
            const double2 lam_c = lds2(la0, cW), mu_c = lds2(mu0, cW), mu_jp = lds2(row3<RB>(st, O_MU, W, r + 1), cW);
            const double lam_r = la0[cW + 2], mu_r = mu0[cW + 2];
            const double2 vx_c = lds2(vx0, cW), vx_r = lds2(vx0, cW + 2), vx_jp = lds2(row3<RB>(st, O_VX, W, r + 1), cW);
            const double2 vy_c = lds2(vy0, cW), vy_l = lds2(vy0, cW - 2), vy_jm = lds2(row3<RB>(st, O_VY, W, r - 1), cW);
            double2 vx_l = make_double2(0.0, 0.0), vx_jpp = vx_l, vx_jm = vx_l, vy_jp = vx_l, vy_jmm = vx_l;
            double vy_r = 0.0;
            if (ORDER == 4) {
                vx_l = lds2(vx0, cW - 2); vx_jpp = lds2(row3<RB>(st, O_VX, W, r + 2), cW); vx_jm = lds2(row3<RB>(st, O_VX, W, r - 1), cW);
                vy_jp = lds2(row3<RB>(st, O_VY, W, r + 1), cW); vy_jmm = lds2(row3<RB>(st, O_VY, W, r - 2), cW); vy_r = vy0[cW + 2];
            }
            const double *ps = reinterpret_cast<const double *>(st.c);
            double2 sxx = lds2(ps + O_SXX / 8, r * TX + cP), syy = lds2(ps + O_SYY / 8, r * TX + cP), sxy = lds2(ps + O_SXY / 8, r * TX + cP);

            bar2_arrive(k2RelBar0 + (int)(cnt & 3u), NALL);      // block n-1 is no longer needed

            if (valid) {
                const long long q = (long long)(j - 1) * pitch + (i - 1);
                const bool in_y = (j <= p.ylo) || (j >= p.yhi);
                const long long qxA = in_xA ? (long long)(j - 1) * p.sxp + shell_index2(i, p.xlo, p.xhi) : 0;
                const long long qxB = in_xB ? (long long)(j - 1) * p.sxp + shell_index2(i + 1, p.xlo, p.xhi) : 0;
                const long long qy = in_y ? (long long)shell_index2(j, p.ylo, p.yhi) * pitch + (i - 1) : 0;
                stress_point2<ORDER>(p, i, j, in_xA, in_y, qxA, qy, lam_c.x, lam_c.y, mu_c.x, mu_c.y, mu_jp.x,
                                     vx_c.y, vx_c.x, vx_r.x, vx_l.y, vx_jp.x, vx_jpp.x, vx_jm.x,
                                     vy_c.x, vy_l.y, vy_c.y, vy_l.x, vy_jm.x, vy_jp.x, vy_jmm.x, sxx.x, syy.x, sxy.x);
                if (validBx)
                    stress_point2<ORDER>(p, i + 1, j, in_xB, in_y, qxB, qy + 1, lam_c.y, lam_r, mu_c.y, mu_r, mu_jp.y,
                                         vx_r.x, vx_c.y, vx_r.y, vx_c.x, vx_jp.y, vx_jpp.y, vx_jm.y,
                                         vy_c.y, vy_c.x, vy_r, vy_l.y, vy_jm.y, vy_jp.y, vy_jmm.y, sxx.y, syy.y, sxy.y);
                st_stream2(p.sxx + q, sxx.x, sxx.y);
                st_stream2(p.syy + q, syy.x, syy.y);
                st_stream2(p.sxy + q, sxy.x, sxy.y);
                epot += epot_point2<ORDER>(p, i, j, lam_c.x, mu_c.x, sxx.x, syy.x, sxy.x);
                if (validBx) epot += epot_point2<ORDER>(p, i + 1, j, lam_c.y, mu_c.y, sxx.y, syy.y, sxy.y);
            }
            pm = pc; pc = pn;
        }
        // the stages of blocks nb-1 and nb are free as well
        bar2_arrive(k2RelBar0 + (int)(cnt & 3u), NALL); ++cnt;
        bar2_arrive(k2RelBar0 + (int)(cnt & 3u), NALL); ++cnt;
        pos = pc;
        pos.advance(SLOTS);               // past block nb

        epot = cons2_sum<NC>(epot, red, tid);
        if (tid == 0) p.partials[p.nblocks + item] = epot;
    }
}

// -------------------------------------------------------------------------------------------------- velocity
// stage: sigma_xx, sigma_xy, rho as (TX+4) x RB tap boxes, sigma_yy as a TX x RB box (y taps); vx, vy as TX x RB boxes
// maps: 0 sxx 1 sxy 2 rho 3 syy 4 vx 5 vy
template <int ORDER, int TX, int RB, int SLOTS, int MINB>
__global__ void __launch_bounds__((TX / 2) * RB + 32, MINB)
k_velocity2d_ws(const __grid_constant__ Params2D p, const __grid_constant__ TmaMaps tm, const __grid_constant__ Tile2D t)
{
    using G = Geom2<TX, RB>;
    constexpr int W = G::W;
    constexpr int NC = (TX / 2) * RB, NALL = NC + 32;
    constexpr int STAGE = 3 * G::TAP + 3 * G::PLAIN;
    constexpr int O_SXX = 0, O_SXY = G::TAP, O_RHO = 2 * G::TAP, O_SYY = 3 * G::TAP, O_VX = O_SYY + G::PLAIN, O_VY = O_VX + G::PLAIN;
    __shared__ double red[NC / 32 + 1];
    __shared__ unsigned long long item_bar;
    __shared__ int item_slot[2];          // the producer posts the CTA's work items here (tma_common.cuh: claim_item)
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t sbase = (smem_u32(smem_dyn) + 127u) & ~127u;
    const unsigned char *gbase = smem_dyn + (sbase - smem_u32(smem_dyn));
    const uint32_t bar = sbase, ring = sbase + kBarBytes;
    const unsigned char *gring = gbase + kBarBytes;

    const uint32_t barI = smem_u32(&item_bar);

    const int tid = (int)threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < SLOTS; s++) mbar_init(bar + 8 * s, 1);
        mbar_init(barI, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= NC) {        // ================================================================ producer warp
        const bool lead = tid == NC;
        uint32_t slot = 0, cnt = 0, ip = 0;
        int item = (int)blockIdx.x;
        while (true) {
            if (lead) { item_slot[ip & 1u] = item; mbar_arrive(barI); }     // slot ip & 1 was last read two items ago
            ++ip;
            if (item >= t.nitems) break;
            const int tix = item % t.ntx, yc = item / t.ntx;
            const int xt = 16 + tix * TX - 2;
            const int jb = 1 + yc * t.rows;
            const int je = min(p.ny, jb + t.rows - 1);
            const int nb = (je - jb + 1 + RB - 1) / RB;
            auto issue = [&](uint32_t s, int b) {
                const uint32_t br = bar + 8 * s, dst = ring + s * STAGE;
                const int yt = 2 + (jb - 1) + b * RB;
                const bool rmw = b >= 0 && b < nb;
                mbar_expect_tx(br, 3 * G::TAP_BOX + G::PLAIN_BOX + (rmw ? 2 * G::PLAIN_BOX : 0));
                tma_load_2d(dst + O_SXX, &tm.m[0], xt, yt, br);
                tma_load_2d(dst + O_SXY, &tm.m[1], xt, yt, br);
                tma_load_2d(dst + O_RHO, &tm.m[2], xt, yt, br);
                tma_load_2d(dst + O_SYY, &tm.m[3], xt + 2, yt, br);
                if (rmw) {
                    tma_load_2d(dst + O_VX, &tm.m[4], xt + 2, yt, br);
                    tma_load_2d(dst + O_VY, &tm.m[5], xt + 2, yt, br);
                }
            };
            if (lead) {
                uint32_t s = slot;
                for (int l = 0; l < min(SLOTS, nb + 2); l++) { issue(s, l - 1); if (++s == SLOTS) s = 0; }
            }
            int next = item + (int)gridDim.x;
            if (lead) next = claim_item(t.queue, next);       // the answer is needed after the row-block loop
            for (int n = 0; n < nb + 2; ++n, ++cnt) {
                bar2_sync(k2RelBar0 + (int)(cnt & 3u), NALL);
                if (lead && n - 1 + SLOTS <= nb) issue(slot, n - 1 + SLOTS);
                if (++slot == SLOTS) slot = 0;
            }
            item = __shfl_sync(0xffffffffu, next, 0);
        }
        if (lead) retire_queue(t.queue);
        return;
    }

    // ================================================================ consumer warps
    const int tx = tid % (TX / 2), r = tid / (TX / 2);
    const int cW = 2 * tx + 2, cP = 2 * tx;
    RingPos pos{0, 0};
    uint32_t cnt = 0;
    for (uint32_t ip = 0;; ++ip) {
        mbar_wait(barI, ip & 1u);
        const int item = item_slot[ip & 1u];
        if (item >= t.nitems) break;
        const int tix = item % t.ntx, yc = item / t.ntx;
        const int i = 1 + tix * TX + 2 * tx;
        const int jb = 1 + yc * t.rows;
        const int je = min(p.ny, jb + t.rows - 1);
        const int nb = (je - jb + 1 + RB - 1) / RB;
        const bool colA = i <= p.nx, validBx = i + 1 <= p.nx;
        const bool in_xA = (i <= p.xlo) || (i >= p.xhi), in_xB = validBx && ((i + 1 <= p.xlo) || (i + 1 >= p.xhi));
        const int pitch = p.pitch;
        double ekin = 0.0;

        RingPos pm = pos, pc = pos;
        pc.advance(SLOTS);
        mbar_wait(bar + 8 * pm.s, pm.par);
        mbar_wait(bar + 8 * pc.s, pc.par);
        for (int n = 0; n < nb; ++n, ++cnt) {
            RingPos pn = pc;
            pn.advance(SLOTS);
            mbar_wait(bar + 8 * pn.s, pn.par);
            const Stages3 st{gring + (size_t)pm.s * STAGE, gring + (size_t)pc.s * STAGE, gring + (size_t)pn.s * STAGE};
            const int j = jb + n * RB + r;
            const bool valid = colA && j <= je;

            const double *xx0 = row3<RB>(st, O_SXX, W, r), *xy0 = row3<RB>(st, O_SXY, W, r);
            const double *rh0 = row3<RB>(st, O_RHO, W, r), *rh1 = row3<RB>(st, O_RHO, W, r + 1);
            const double2 rho_c = lds2(rh0, cW), rho_jp = lds2(rh1, cW);
            const double rho_r = rh0[cW + 2], rho_jpr = rh1[cW + 2];
            const double2 sxx_c = lds2(xx0, cW), sxx_l = lds2(xx0, cW - 2);
            const double2 sxy_c = lds2(xy0, cW), sxy_r = lds2(xy0, cW + 2), sxy_jm = lds2(row3<RB>(st, O_SXY, W, r - 1), cW);
            const double2 syy_c = lds2(row3<RB>(st, O_SYY, TX, r), cP), syy_jp = lds2(row3<RB>(st, O_SYY, TX, r + 1), cP);
            double2 z2 = make_double2(0.0, 0.0), sxy_l = z2, sxy_jp = z2, sxy_jmm = z2, syy_jpp = z2, syy_jm = z2;
            double sxx_r = 0.0;
            if (ORDER == 4) {
                sxx_r = xx0[cW + 2]; sxy_l = lds2(xy0, cW - 2); sxy_jp = lds2(row3<RB>(st, O_SXY, W, r + 1), cW);
                sxy_jmm = lds2(row3<RB>(st, O_SXY, W, r - 2), cW);
                syy_jpp = lds2(row3<RB>(st, O_SYY, TX, r + 2), cP); syy_jm = lds2(row3<RB>(st, O_SYY, TX, r - 1), cP);
            }
            const double *pv = reinterpret_cast<const double *>(st.c);
            double2 v_x = lds2(pv + O_VX / 8, r * TX + cP), v_y = lds2(pv + O_VY / 8, r * TX + cP);

            bar2_arrive(k2RelBar0 + (int)(cnt & 3u), NALL);

            if (valid) {
                const long long q = (long long)(j - 1) * pitch + (i - 1);
                const bool in_y = (j <= p.ylo) || (j >= p.yhi);
                const long long qxA = in_xA ? (long long)(j - 1) * p.sxp + shell_index2(i, p.xlo, p.xhi) : 0;
                const long long qxB = in_xB ? (long long)(j - 1) * p.sxp + shell_index2(i + 1, p.xlo, p.xhi) : 0;
                const long long qy = in_y ? (long long)shell_index2(j, p.ylo, p.yhi) * pitch + (i - 1) : 0;
                const double rho_hA = 0.25 * (rho_c.x + rho_c.y + rho_jp.y + rho_jp.x);          // 2D-2nd :627
                const double rho_hB = 0.25 * (rho_c.y + rho_r + rho_jpr + rho_jp.y);
                velocity_point2<ORDER>(p, i, j, in_xA, in_y, qxA, qy, rho_c.x, rho_hA,
                                       sxx_c.x, sxx_l.y, sxx_c.y, sxx_l.x,
                                       sxy_c.x, sxy_jm.x, sxy_jp.x, sxy_jmm.x, sxy_c.y, sxy_r.x, sxy_l.y,
                                       syy_c.x, syy_jp.x, syy_jpp.x, syy_jm.x, v_x.x, v_y.x, ekin);
                if (validBx)
                    velocity_point2<ORDER>(p, i + 1, j, in_xB, in_y, qxB, qy + 1, rho_c.y, rho_hB,
                                           sxx_c.y, sxx_c.x, sxx_r, sxx_l.y,
                                           sxy_c.y, sxy_jm.y, sxy_jp.y, sxy_jmm.y, sxy_r.x, sxy_r.y, sxy_c.x,
                                           syy_c.y, syy_jp.y, syy_jpp.y, syy_jm.y, v_x.y, v_y.y, ekin);
                st_stream2(p.vx + q, v_x.x, v_x.y);
                st_stream2(p.vy + q, v_y.x, v_y.y);
            }
            pm = pc; pc = pn;
        }
        bar2_arrive(k2RelBar0 + (int)(cnt & 3u), NALL); ++cnt;
        bar2_arrive(k2RelBar0 + (int)(cnt & 3u), NALL); ++cnt;
        pos = pc;
        pos.advance(SLOTS);

        ekin = cons2_sum<NC>(ekin, red, tid);
        if (tid == 0) p.partials[item] = ekin;
    }
}

// ---- launch dispatch ---------------------------------------------------------------
// Geometry: 64-column strips in 4-row blocks (default) or 128-column strips in 2-row blocks (CPML_2D_WS_TX=128: longer
// contiguous row segments per TMA box, twice the marching iterations); four ring stages, three CTAs per SM either way.

constexpr int k2SLOTS = 4;

static int ws2_tx()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("CPML_2D_WS_TX"); v = e ? atoi(e) : 64; if (v != 64 && v != 128) v = 64; }
    return v;
}

static int ws2_minb()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("CPML_2D_WS_MINB"); v = e ? atoi(e) : 3; if (v != 2 && v != 3) v = 3; }
    return v;
}

template <int ORDER, int TX, int RB, int MINB>
static cudaError_t ws2_launch(const Params2D &p, const TmaMaps &tm, const Tile2D &t, cudaStream_t s, bool stress, int *occ)
{
    using G = Geom2<TX, RB>;
    const size_t stage = stress ? 4 * G::TAP + 3 * G::PLAIN : 3 * G::TAP + 3 * G::PLAIN;
    const size_t smem = kBarBytes + 128 + stage * k2SLOTS;
    constexpr int NT = (TX / 2) * RB + 32;
    const void *fn = stress ? (const void *)k_stress2d_ws<ORDER, TX, RB, k2SLOTS, MINB> : (const void *)k_velocity2d_ws<ORDER, TX, RB, k2SLOTS, MINB>;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (occ) {
        if (stress) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k_stress2d_ws<ORDER, TX, RB, k2SLOTS, MINB>, NT, smem);
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k_velocity2d_ws<ORDER, TX, RB, k2SLOTS, MINB>, NT, smem);
    }
    const int grid = stress ? t.grid_stress : t.grid_velocity;
    if (stress) k_stress2d_ws<ORDER, TX, RB, k2SLOTS, MINB><<<grid, NT, smem, s>>>(p, tm, t);
    else        k_velocity2d_ws<ORDER, TX, RB, k2SLOTS, MINB><<<grid, NT, smem, s>>>(p, tm, t);
    return cudaGetLastError();
}

template <int ORDER>
static cudaError_t ws2_launch(const Params2D &p, const TmaMaps &tm, const Tile2D &t, cudaStream_t s, bool stress, int *occ)
{
    if (ws2_tx() == 128)
        return ws2_minb() == 3 ? ws2_launch<ORDER, 128, 2, 3>(p, tm, t, s, stress, occ) : ws2_launch<ORDER, 128, 2, 2>(p, tm, t, s, stress, occ);
    return ws2_minb() == 3 ? ws2_launch<ORDER, 64, 4, 3>(p, tm, t, s, stress, occ) : ws2_launch<ORDER, 64, 4, 2>(p, tm, t, s, stress, occ);
}

void ws2_geometry(int *tx, int *rb, int (*box_stress)[2], int (*box_velocity)[2])
{
    const int TX = ws2_tx(), RB = TX == 128 ? 2 : 4;
    *tx = TX; *rb = RB;
    const int bs[7][2] = {{TX + 4, RB}, {TX + 4, RB}, {TX + 4, RB}, {TX + 4, RB}, {TX, RB}, {TX, RB}, {TX, RB}};
    const int bv[6][2] = {{TX + 4, RB}, {TX + 4, RB}, {TX + 4, RB}, {TX, RB}, {TX, RB}, {TX, RB}};
    for (int m = 0; m < 7; m++) { box_stress[m][0] = bs[m][0]; box_stress[m][1] = bs[m][1]; }
    for (int m = 0; m < 6; m++) { box_velocity[m][0] = bv[m][0]; box_velocity[m][1] = bv[m][1]; }
}

cudaError_t ws2_occupancy(int order, bool stress, int *occ)
{
    Params2D p{};
    TmaMaps dummy{};
    Tile2D t{};
    return order == 4 ? ws2_launch<4>(p, dummy, t, nullptr, stress, occ) : ws2_launch<2>(p, dummy, t, nullptr, stress, occ);
}

cudaError_t launch_stress2d_ws(const Params2D &p, const TmaMaps &tm, const Tile2D &t, cudaStream_t s)
{
    return p.order == 4 ? ws2_launch<4>(p, tm, t, s, true, nullptr) : ws2_launch<2>(p, tm, t, s, true, nullptr);
}
cudaError_t launch_velocity2d_ws(const Params2D &p, const TmaMaps &tm, const Tile2D &t, cudaStream_t s)
{
    return p.order == 4 ? ws2_launch<4>(p, tm, t, s, false, nullptr) : ws2_launch<2>(p, tm, t, s, false, nullptr);
}

}  // namespace cpml
