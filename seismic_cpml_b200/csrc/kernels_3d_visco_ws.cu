// 3-D viscoelastic velocity kernel for sm_100a, TMA-staged with a producer warp (the default velocity kernel of the
// viscoelastic solver; kernels_3d_visco.cu keeps the register-marching one for A/B runs, CPML_VKERNEL=reg).
//
//   k_vvelocity3d_ws  vx/vy (:1244-1285), vz (:1287-1308), source (:1310-1335), Dirichlet on two planes per face
//                     (:1337-1371), the kinetic part of the energy (:1396-1397)
// (line numbers: seismic_CPML_3D_viscoelastic_MPI.f90)
//
// Why.  The register-marching kernel reads its 24 fourth-order taps per point through L1 and runs at 48-56 % of the
// measured HBM bandwidth (profiles/r01_v9_ncu_cfg5d.txt: 3.85 long-scoreboard + 2.0 wait stall cycles per issued
// instruction, 25 % of the warps active, DRAM at 42 % of peak): latency bound.  Here, as in kernels_3d_ws.cu, every
// plane tile comes through the TMA unit into a ring of shared-memory stages filled by a dedicated producer warp, the
// compute warps never meet in a barrier inside the plane loop, and each thread updates two x-adjacent points, so the
// x taps of both points are three 16-byte shared-memory loads instead of eight global ones.
//
// Tiles per plane (TX x TY points; boxes start two cells before the tile where the operator reaches back: the x start
// must stay 16-byte aligned, see tma_common.cuh):
//   ring C (plane n):      sigmaxx (TX+4) x TY        x taps i-2..i+2 of the pair
//                          sigmaxy (TX+4) x (TY+4)    x taps (vy) and y taps j-2..j+1 (vx)
//                          sigmayy TX x (TY+4)        y taps j-1..j+2
//                          vx vy vz TX x TY, sigmazz of plane n+2 (TX x TY), x-shell memory variables of the rows
//   ring N (planes n, n+1): sigmaxz (TX+4) x TY, sigmayz TX x (TY+4)   in-plane taps of plane n, centre of plane n+1
// The rest of the z windows (sigmaxz, sigmayz at k-2, k-1; sigmazz at k-1, k, k+1) is carried in registers.  The
// tensors are the PADDED arrays (two-cell zero ghost ring in x and y, two halo planes per side in z), so edge taps read
// the zeros the reference's (0:NX+1,0:NY+1,-1:NZ_LOCAL+2) arrays hold, from memory.
//
// Arithmetic: the same functions as the register-marching kernel (d4n, the shared-range-test divisions by 24, div_exact
// for /K), same order: bit-identical fields (tests/test_gpu_visco.py runs both).
#include "tma_common.cuh"
#include "visco_common.cuh"

namespace cpml {

namespace {

constexpr int kVConsBar = 1;
constexpr int kVRelBar0 = 2;

__device__ __forceinline__ void vbar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void vbar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

template <int TX, int TY>
struct VGeom {
    static constexpr int W4 = TX + 4, H4 = TY + 4;
    static constexpr int SXX = round128(W4 * TY * 8), SXY = round128(W4 * H4 * 8), SYY = round128(TX * H4 * 8);
    static constexpr int SXZ = SXX, SYZ = SYY, PL = round128(TX * TY * 8);
    // byte offsets inside a ring-C stage
    static constexpr int O_SXX = 0, O_SXY = O_SXX + SXX, O_SYY = O_SXY + SXY, O_VX = O_SYY + SYY, O_VY = O_VX + PL, O_VZ = O_VY + PL,
                         O_SZZ = O_VZ + PL, CBYTES = O_SZZ + PL;
    static constexpr int O_SXZ = 0, O_SYZ = SXZ, NBYTES = SXZ + SYZ;
    static constexpr uint32_t TX_C = (W4 * TY + W4 * H4 + TX * H4 + 4 * TX * TY) * 8;
    static constexpr uint32_t TX_N = (W4 * TY + TX * H4) * 8;
};

// C-PML recursion with the old memory variable already in a register (staged x shell): memory = b*memory + a*value,
// value/K + memory (:1253-1259)
__device__ __forceinline__ double vcpml_m(double *__restrict__ mem, int q, double m, double b, double a, double K, double rK, double value)
{
    m = b * m + a * value;
    mem[q] = m;
    return div_exact(value, K, rK) + m;
}

template <int NC>
__device__ __forceinline__ double vcons_sum(double a, double *red, int tid)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
    constexpr int NW = NC / 32;
    const int w = tid >> 5, l = tid & 31;
    if (l == 0) red[w] = a;
    vbar_sync(kVConsBar, NC);
    if (w == 0) {
        a = (l < NW) ? red[l] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
    }
    vbar_sync(kVConsBar, NC);
    return a;
}

}  // namespace

template <int TX, int TY>
__global__ void __launch_bounds__(tile_threads(TX, TY) + 32, 1)
k_vvelocity3d_ws(const __grid_constant__ ParamsV3D p, const __grid_constant__ TmaMaps tm, const __grid_constant__ Tile3D t)
{
    using G = VGeom<TX, TY>;
    constexpr int W4 = G::W4;
    constexpr int NC = tile_threads(TX, TY), NALL = NC + 32;
    const uint32_t XMB = (uint32_t)t.xm_bytes, XM_TX = (uint32_t)(TY * p.sxp * 8);
    const uint32_t CSTAGE = G::CBYTES + 3 * XMB;

    __shared__ double red[NC / 32];
    __shared__ double Cx[8 * TX];        // a, b, K, 1/K, a_half, b_half, K_half, 1/K_half of the tile's columns
    __shared__ unsigned long long item_bar;
    __shared__ int item_slot[2];         // the producer posts the CTA's work items here (tma_common.cuh: claim_item)
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t sbase = (smem_u32(smem_dyn) + 127u) & ~127u;
    const unsigned char *gbase = smem_dyn + (sbase - smem_u32(smem_dyn));
    const uint32_t SC = (uint32_t)t.stages, SN = SC + 1;
    const uint32_t barN = sbase, barC = sbase + 64;
    const uint32_t ringN = sbase + kBarBytes, ringC = ringN + SN * G::NBYTES;
    const unsigned char *gN = gbase + kBarBytes, *gC = gN + (size_t)SN * G::NBYTES;

    const uint32_t barI = smem_u32(&item_bar);

    const int tid = (int)threadIdx.x;
    if (tid == 0) {
        for (uint32_t s = 0; s < SN; s++) mbar_init(barN + 8 * s, 1);
        for (uint32_t s = 0; s < SC; s++) mbar_init(barC + 8 * s, 1);
        mbar_init(barI, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // ================================================================ producer warp
    if (tid >= NC) {
        const bool lead = tid == NC;
        uint32_t sn = 0, sc = 0, ip = 0;
        int item = (int)blockIdx.x;
        while (true) {
            if (lead) { item_slot[ip & 1u] = item; mbar_arrive(barI); }     // slot ip & 1 was last read two items ago
            ++ip;
            if (item >= t.nitems) break;
            const int tix = item % t.ntx, rest = item / t.ntx, tiy = rest % t.nty, zc = rest / t.nty;
            // tensor coordinates of element (i0, j0, k): x = 16 + (i0-1), y = 2 + (j0-1), z = k + 1 (padded arrays)
            const int xt = 16 + tix * TX, yt = 2 + tiy * TY, j0 = 1 + tiy * TY;
            const int kb = 1 + zc * t.kchunk;
            const int ke = min(p.nzl, kb + t.kchunk - 1);
            const int np = ke - kb + 1;
            const bool tile_xpml = XMB != 0 && ((tix * TX + 1 <= p.xlo) || (tix * TX + TX >= p.xhi));
            auto issue_n = [&](uint32_t s, int kk) {
                const uint32_t bar = barN + 8 * s, dst = ringN + s * G::NBYTES;
                mbar_expect_tx(bar, G::TX_N);
                tma_load_3d(dst + G::O_SXZ, &tm.m[3], xt - 2, yt, kk + 1, bar);
                tma_load_3d(dst + G::O_SYZ, &tm.m[4], xt, yt - 2, kk + 1, bar);
            };
            auto issue_c = [&](uint32_t s, int kk) {
                const uint32_t bar = barC + 8 * s, dst = ringC + s * CSTAGE;
                mbar_expect_tx(bar, G::TX_C + (tile_xpml ? 3 * XM_TX : 0u));
                if (tile_xpml) {
                    const long long row0 = ((long long)(kk - 1) * p.ny + (j0 - 1)) * p.sxp;
#pragma unroll
                    for (int f = 0; f < 3; f++) bulk_load(dst + G::CBYTES + f * XMB, p.mx[3 + f] + row0, XM_TX, bar);
                }
                tma_load_3d(dst + G::O_SXX, &tm.m[0], xt - 2, yt, kk + 1, bar);
                tma_load_3d(dst + G::O_SXY, &tm.m[1], xt - 2, yt - 2, kk + 1, bar);
                tma_load_3d(dst + G::O_SYY, &tm.m[2], xt, yt - 2, kk + 1, bar);
                tma_load_3d(dst + G::O_VX, &tm.m[6], xt, yt, kk + 1, bar);
                tma_load_3d(dst + G::O_VY, &tm.m[7], xt, yt, kk + 1, bar);
                tma_load_3d(dst + G::O_VZ, &tm.m[8], xt, yt, kk + 1, bar);
                tma_load_3d(dst + G::O_SZZ, &tm.m[5], xt, yt, kk + 3, bar);       // sigmazz of plane kk+2
            };
            if (lead) {
                uint32_t s = sn;
                for (int l = 0; l < min((int)SN, np + 1); l++) { issue_n(s, kb + l); if (++s == SN) s = 0; }
                s = sc;
                for (int l = 0; l < min((int)SC, np); l++) { issue_c(s, kb + l); if (++s == SC) s = 0; }
            }
            int next = item + (int)gridDim.x;
            if (lead) next = claim_item(t.queue, next);      // the answer is needed after the plane loop
            for (int n = 0; n < np; ++n) {
                vbar_sync(kVRelBar0 + (int)sc, NALL);
                if (lead) {
                    if (n + (int)SN <= np) issue_n(sn, kb + n + (int)SN);
                    if (n + (int)SC < np) issue_c(sc, kb + n + (int)SC);
                }
                if (++sn == SN) sn = 0;
                if (++sc == SC) sc = 0;
            }
            if (++sn == SN) sn = 0;
            item = __shfl_sync(0xffffffffu, next, 0);
        }
        if (lead) retire_queue(t.queue);
        return;
    }

    // ================================================================ consumer warps
    const int tx = tid % (TX / 2);
    const int ty_raw = tid / (TX / 2);
    const bool lane_ok = ty_raw < TY;
    const int ty = min(ty_raw, TY - 1);
    const int c = 2 * tx;
    const int pitch = p.pitch;
    const int pl = (int)p.plane;             // < 2^31 elements per field (checked by finalize)
    const double odx = p.odx, ody = p.ody, odz = p.odz, dt_r = p.dt_over_rho;

    RingPos rn{0, 0}, rc{0, 0};
    for (uint32_t ip = 0;; ++ip) {
        mbar_wait(barI, ip & 1u);
        const int item = item_slot[ip & 1u];
        if (item >= t.nitems) break;
        const int tix = item % t.ntx, rest = item / t.ntx, tiy = rest % t.nty, zc = rest / t.nty;
        const int i0 = 1 + tix * TX, j0 = 1 + tiy * TY;
        const int kb = 1 + zc * t.kchunk;
        const int ke = min(p.nzl, kb + t.kchunk - 1);
        const int np = ke - kb + 1;
        const bool tile_xpml = XMB != 0 && ((i0 <= p.xlo) || (i0 + TX - 1 >= p.xhi));

        const int i = i0 + c, j = j0 + ty;
        const bool row = lane_ok && (j <= p.ny);
        const bool validA = row && (i <= p.nx), validB = row && (i + 1 <= p.nx);
        int q = kb * pl + (j - 1) * pitch + (i - 1);

        const bool in_xA = validA && ((i <= p.xlo) || (i >= p.xhi));
        const bool in_xB = validB && ((i + 1 <= p.xlo) || (i + 1 >= p.xhi));
        const bool in_y = validA && ((j <= p.ylo) || (j >= p.yhi));
        const int sxA = in_xA ? vshell(i, p.xlo, p.xhi) : 0, sxB = in_xB ? vshell(i + 1, p.xlo, p.xhi) : 0;
        const int sy = in_y ? vshell(j, p.ylo, p.yhi) : 0;
        if (tile_xpml) {
            for (int e = tid; e < 8 * TX; e += NC) {
                const int m = e / TX, ii = i0 + (e - m * TX);
                const double *src = m == 0 ? p.cx.a : m == 1 ? p.cx.b : m == 2 ? p.cx.K : m == 3 ? p.cx.rK : m == 4 ? p.cx.a_half
                                  : m == 5 ? p.cx.b_half : m == 6 ? p.cx.K_half : p.cx.rK_half;
                Cx[e] = (ii <= p.nx) ? src[ii] : ((m == 0 || m == 4) ? 0.0 : 1.0);
            }
        }
        vbar_sync(kVConsBar, NC);

        const bool do_vxA = validA && (i >= 2) && (j >= 2), do_vxB = validB && (j >= 2);                                          // :1246-1247
        const bool do_vyA = validA && (i <= p.nx - 1) && (j <= p.ny - 1), do_vyB = validB && (i + 1 <= p.nx - 1) && (j <= p.ny - 1);    // :1266-1267
        const bool do_vzA = validA && (i <= p.nx - 1) && (j >= 2), do_vzB = validB && (i + 1 <= p.nx - 1) && (j >= 2);            // :1289-1290
        const bool edge_j = (j <= 1) || (j >= p.ny);
        const bool edgeA = (i <= 1) || (i >= p.nx) || edge_j, edgeB = (i + 1 >= p.nx) || edge_j;                                  // :1340-1358
        const bool ebox_j = (j >= p.npml) && (j <= p.ny - p.npml + 1);
        const bool eboxA = validA && ebox_j && (i >= p.npml) && (i <= p.nx - p.npml + 1);
        const bool eboxB = validB && ebox_j && (i + 1 >= p.npml) && (i + 1 <= p.nx - p.npml + 1);
        const bool srcA = validA && (i == p.isrc) && (j == p.jsrc), srcB = validB && (i + 1 == p.isrc) && (j == p.jsrc);

        // z windows carried in registers: sigmaxz, sigmayz at k-2, k-1; sigmazz at k-1, k, k+1
        double2 sxz_mm = {0, 0}, sxz_m = {0, 0}, syz_mm = {0, 0}, syz_m = {0, 0}, szz_m = {0, 0}, szz_c = {0, 0}, szz_p = {0, 0};
        if (validA) {
            sxz_mm = *reinterpret_cast<const double2 *>(p.sxz + q - 2 * pl); sxz_m = *reinterpret_cast<const double2 *>(p.sxz + q - pl);
            syz_mm = *reinterpret_cast<const double2 *>(p.syz + q - 2 * pl); syz_m = *reinterpret_cast<const double2 *>(p.syz + q - pl);
            szz_m = *reinterpret_cast<const double2 *>(p.szz + q - pl); szz_c = *reinterpret_cast<const double2 *>(p.szz + q);
            szz_p = *reinterpret_cast<const double2 *>(p.szz + q + pl);
        }
        double ekin = 0.0;
        int kmod = (kb + p.koff) % p.nzl_e;                      // position inside the emulated reference slab (quirk B6)

        mbar_wait(barN + 8 * rn.s, rn.par);
        for (int n = 0; n < np; ++n, q += pl, kmod = (kmod + 1 == p.nzl_e) ? 0 : kmod + 1) {
            const int k = kb + n;
            const int kg = k + p.koff;
            const bool in_z = (kg <= p.zlo) || (kg >= p.zhi);    // uniform
            const bool cut_up = (kmod == 0), cut_dn = (kmod == 1);
            // y / z shell memory variables of the next plane into L2 (each recursion is a dependent load -> update -> store)
            if ((p.pf & 4) && k + 1 <= p.nzl) {
                if (in_y) {
                    const int qn = (k * p.sy + sy) * pitch + (i - 1);
                    pf_l2(p.my[3] + qn); pf_l2(p.my[4] + qn); pf_l2(p.my[5] + qn);
                }
                const int kgn = kg + 1;
                if (validA && ((kgn <= p.zlo) || (kgn >= p.zhi))) {
                    const int qn = ((vshell(kgn, p.zlo, p.zhi) - p.zbase) * p.ny + (j - 1)) * pitch + (i - 1);
                    pf_l2(p.mz[3] + qn); pf_l2(p.mz[4] + qn); pf_l2(p.mz[5] + qn);
                }
            }
            RingPos rn1 = rn;
            rn1.advance(SN);
            mbar_wait(barN + 8 * rn1.s, rn1.par);
            mbar_wait(barC + 8 * rc.s, rc.par);

            const unsigned char *Cb = gC + (size_t)rc.s * CSTAGE;
            const double *Txx = (const double *)(Cb + G::O_SXX), *Txy = (const double *)(Cb + G::O_SXY), *Tyy = (const double *)(Cb + G::O_SYY);
            const double *Txz = (const double *)(gN + (size_t)rn.s * G::NBYTES + G::O_SXZ), *Tyz = (const double *)(gN + (size_t)rn.s * G::NBYTES + G::O_SYZ);
            const double *Txz1 = (const double *)(gN + (size_t)rn1.s * G::NBYTES + G::O_SXZ), *Tyz1 = (const double *)(gN + (size_t)rn1.s * G::NBYTES + G::O_SYZ);
            const int oW = ty * W4 + c, oP = ty * TX + c;

            // ---- vx numerators (:1248-1250)
            const double2 sxz_c = lds2(Txz, oW + 2), sxz_p = lds2(Txz1, oW + 2);
            const double2 sxy_c = lds2(Txy, oW + 2 * W4 + 2);
            double d1A, d1B, d2A, d2B, d3A, d3B;                 // vx
            {
                const double2 s01 = lds2(Txx, oW), s23 = lds2(Txx, oW + 2);
                const double s4 = Txx[oW + 4];
                d1A = d4n(s23.x, s01.y, s23.y, s01.x, odx);
                d1B = d4n(s23.y, s23.x, s4, s01.y, odx);
                const double2 jm1 = lds2(Txy, oW + W4 + 2), jp1 = lds2(Txy, oW + 3 * W4 + 2), jm2 = lds2(Txy, oW + 2);
                d2A = d4n(sxy_c.x, jm1.x, jp1.x, jm2.x, ody);
                d2B = d4n(sxy_c.y, jm1.y, jp1.y, jm2.y, ody);
                d3A = d4n(sxz_c.x, sxz_m.x, cut_up ? 0.0 : sxz_p.x, sxz_mm.x, odz);
                d3B = d4n(sxz_c.y, sxz_m.y, cut_up ? 0.0 : sxz_p.y, sxz_mm.y, odz);
            }
            // ---- vy numerators (:1268-1270)
            const double2 syz_c = lds2(Tyz, oP + 2 * TX), syz_p = lds2(Tyz1, oP + 2 * TX);
            double e1A, e1B, e2A, e2B, e3A, e3B;                 // vy
            {
                const double x1 = Txy[oW + 2 * W4 + 1];
                const double2 x45 = lds2(Txy, oW + 2 * W4 + 4);
                e1A = d4n(sxy_c.y, sxy_c.x, x45.x, x1, odx);
                e1B = d4n(x45.x, sxy_c.y, x45.y, sxy_c.x, odx);
                const double2 jm1 = lds2(Tyy, oP + TX), cc = lds2(Tyy, oP + 2 * TX), jp1 = lds2(Tyy, oP + 3 * TX), jp2 = lds2(Tyy, oP + 4 * TX);
                e2A = d4n(jp1.x, cc.x, jp2.x, jm1.x, ody);
                e2B = d4n(jp1.y, cc.y, jp2.y, jm1.y, ody);
                e3A = d4n(syz_c.x, syz_m.x, cut_up ? 0.0 : syz_p.x, syz_mm.x, odz);
                e3B = d4n(syz_c.y, syz_m.y, cut_up ? 0.0 : syz_p.y, syz_mm.y, odz);
            }
            // ---- vz numerators (:1291-1293)
            const double2 szz_pp = lds2((const double *)(Cb + G::O_SZZ), oP);
            double f1A, f1B, f2A, f2B, f3A, f3B;                 // vz
            {
                const double x1 = Txz[oW + 1];
                const double2 x45 = lds2(Txz, oW + 4);
                f1A = d4n(sxz_c.y, sxz_c.x, x45.x, x1, odx);
                f1B = d4n(x45.x, sxz_c.y, x45.y, sxz_c.x, odx);
                const double2 jm2 = lds2(Tyz, oP), jm1 = lds2(Tyz, oP + TX), jp1 = lds2(Tyz, oP + 3 * TX);
                f2A = d4n(syz_c.x, jm1.x, jp1.x, jm2.x, ody);
                f2B = d4n(syz_c.y, jm1.y, jp1.y, jm2.y, ody);
                f3A = d4n(szz_p.x, szz_c.x, szz_pp.x, cut_dn ? 0.0 : szz_m.x, odz);
                f3B = d4n(szz_p.y, szz_c.y, szz_pp.y, cut_dn ? 0.0 : szz_m.y, odz);
            }
            double2 vx = lds2((const double *)(Cb + G::O_VX), oP), vy = lds2((const double *)(Cb + G::O_VY), oP),
                    vz = lds2((const double *)(Cb + G::O_VZ), oP);
            // x-shell memory variables of this plane, staged with the tiles
            double mxA[3] = {0, 0, 0}, mxB[3] = {0, 0, 0};
            if (in_xA | in_xB) {
                const double *Tm = (const double *)(Cb + G::CBYTES);
                const int xd = (int)(XMB / 8), r0 = ty * p.sxp;
                if (in_xA) { mxA[0] = Tm[r0 + sxA]; mxA[1] = Tm[xd + r0 + sxA]; mxA[2] = Tm[2 * xd + r0 + sxA]; }
                if (in_xB) { mxB[0] = Tm[r0 + sxB]; mxB[1] = Tm[xd + r0 + sxB]; mxB[2] = Tm[2 * xd + r0 + sxB]; }
            }
            vbar_arrive(kVRelBar0 + (int)rc.s, NALL);            // the stages of this plane have been read

            // ---- divisions by 24 (one range test per nest) and the C-PML recursions (:1251-1259, :1271-1279, :1294-1302)
            DIV24_3(d1A, d2A, d3A); DIV24_3(d1B, d2B, d3B);
            DIV24_3(e1A, e2A, e3A); DIV24_3(e1B, e2B, e3B);
            DIV24_3(f1A, f2A, f3A); DIV24_3(f1B, f2B, f3B);
            // (a recursion runs -- and its memory variable moves -- only where its nest does, like in the reference)
            const bool kv = kg >= 2, kw = kg <= p.nz - 1;        // k2begin (vx, vy) / kminus1end (vz)
            if (in_xA | in_xB | in_y | in_z) {
                const int qxr = ((k - 1) * p.ny + (j - 1)) * p.sxp;
                if (in_xA) {
                    const double *cx = Cx + c;
                    if (do_vxA && kv) d1A = vcpml_m(p.mx[3], qxr + sxA, mxA[0], cx[1 * TX], cx[0 * TX], cx[2 * TX], cx[3 * TX], d1A);
                    if (do_vyA && kv) e1A = vcpml_m(p.mx[4], qxr + sxA, mxA[1], cx[5 * TX], cx[4 * TX], cx[6 * TX], cx[7 * TX], e1A);
                    if (do_vzA && kw) f1A = vcpml_m(p.mx[5], qxr + sxA, mxA[2], cx[5 * TX], cx[4 * TX], cx[6 * TX], cx[7 * TX], f1A);
                }
                if (in_xB) {
                    const double *cx = Cx + c + 1;
                    if (do_vxB && kv) d1B = vcpml_m(p.mx[3], qxr + sxB, mxB[0], cx[1 * TX], cx[0 * TX], cx[2 * TX], cx[3 * TX], d1B);
                    if (do_vyB && kv) e1B = vcpml_m(p.mx[4], qxr + sxB, mxB[1], cx[5 * TX], cx[4 * TX], cx[6 * TX], cx[7 * TX], e1B);
                    if (do_vzB && kw) f1B = vcpml_m(p.mx[5], qxr + sxB, mxB[2], cx[5 * TX], cx[4 * TX], cx[6 * TX], cx[7 * TX], f1B);
                }
                if (in_y) {
                    const int qy = ((k - 1) * p.sy + sy) * pitch + (i - 1);
                    const double ay = p.cy.a[j], by = p.cy.b[j], Ky = p.cy.K[j], rKy = p.cy.rK[j];
                    const double ayh = p.cy.a_half[j], byh = p.cy.b_half[j], Kyh = p.cy.K_half[j], rKyh = p.cy.rK_half[j];
                    if (do_vxA && kv) d2A = vcpml(p.my[3], qy, by, ay, Ky, rKy, d2A);
                    if (do_vyA && kv) e2A = vcpml(p.my[4], qy, byh, ayh, Kyh, rKyh, e2A);
                    if (do_vzA && kw) f2A = vcpml(p.my[5], qy, by, ay, Ky, rKy, f2A);
                    if (do_vxB && kv) d2B = vcpml(p.my[3], qy + 1, by, ay, Ky, rKy, d2B);
                    if (do_vyB && kv) e2B = vcpml(p.my[4], qy + 1, byh, ayh, Kyh, rKyh, e2B);
                    if (do_vzB && kw) f2B = vcpml(p.my[5], qy + 1, by, ay, Ky, rKy, f2B);
                }
                if (in_z && validA) {
                    const int qz = ((vshell(kg, p.zlo, p.zhi) - p.zbase) * p.ny + (j - 1)) * pitch + (i - 1);
                    const double az = p.cz.a[kg], bz = p.cz.b[kg], Kz = p.cz.K[kg], rKz = p.cz.rK[kg];
                    const double azh = p.cz.a_half[kg], bzh = p.cz.b_half[kg], Kzh = p.cz.K_half[kg], rKzh = p.cz.rK_half[kg];
                    if (do_vxA && kv) d3A = vcpml(p.mz[3], qz, bz, az, Kz, rKz, d3A);
                    if (do_vyA && kv) e3A = vcpml(p.mz[4], qz, bz, az, Kz, rKz, e3A);
                    if (do_vzA && kw) f3A = vcpml(p.mz[5], qz, bzh, azh, Kzh, rKzh, f3A);
                    if (do_vxB && kv) d3B = vcpml(p.mz[3], qz + 1, bz, az, Kz, rKz, d3B);
                    if (do_vyB && kv) e3B = vcpml(p.mz[4], qz + 1, bz, az, Kz, rKz, e3B);
                    if (do_vzB && kw) f3B = vcpml(p.mz[5], qz + 1, bzh, azh, Kzh, rKzh, f3B);
                }
            }
            // ---- updates
            if (kv) {
                if (do_vxA) vx.x = dt_r * (d1A + d2A + d3A) + vx.x;
                if (do_vxB) vx.y = dt_r * (d1B + d2B + d3B) + vx.y;
                if (do_vyA) vy.x = dt_r * (e1A + e2A + e3A) + vy.x;
                if (do_vyB) vy.y = dt_r * (e1B + e2B + e3B) + vy.y;
            }
            if (kw) {
                if (do_vzA) vz.x = dt_r * (f1A + f2A + f3A) + vz.x;
                if (do_vzB) vz.y = dt_r * (f1B + f2B + f3B) + vz.y;
            }
            // source (:1332-1333), after the update of step it and before Dirichlet
            if (k == p.ksrc && (srcA || srcB)) {
                const double fx = p.src_x[p.it - 1], fy = p.src_y[p.it - 1];
                if (srcA) { vx.x = vx.x + fx; vy.x = vy.x + fy; }
                if (srcB) { vx.y = vx.y + fx; vy.y = vy.y + fy; }
            }
            // Dirichlet, two planes per face (:1337-1371); ghost cells and outer halo planes are never written
            const bool edge_k = kg <= 1 || kg >= p.nz;
            if (edgeA || edge_k) { vx.x = 0.0; vy.x = 0.0; vz.x = 0.0; }
            if (edgeB || edge_k) { vx.y = 0.0; vy.y = 0.0; vz.y = 0.0; }

            if (validA) {
                // B beyond NX is a ghost cell: outside every nest, it keeps the zero it was loaded with
                if (!validB) { vx.y = 0.0; vy.y = 0.0; vz.y = 0.0; }
                __stcs(reinterpret_cast<double2 *>(p.vx + q), vx);
                __stcs(reinterpret_cast<double2 *>(p.vy + q), vy);
                __stcs(reinterpret_cast<double2 *>(p.vz + q), vz);
                if ((k <= 2) | (k >= p.nzl - 1)) {               // planes a neighbour slab needs (:962-975)
                    const int rel = q - k * pl;
                    peer_put<true>(p.peer_lo[0], p.peer_hi[0], k, p.nzl, rel, pl, vx.x);
                    peer_put<true>(p.peer_lo[1], p.peer_hi[1], k, p.nzl, rel, pl, vy.x);
                    peer_put<false>(p.peer_lo[3], p.peer_hi[3], k, p.nzl, rel, pl, vz.x);
                    if (validB) {
                        peer_put<true>(p.peer_lo[0], p.peer_hi[0], k, p.nzl, rel + 1, pl, vx.y);
                        peer_put<true>(p.peer_lo[1], p.peer_hi[1], k, p.nzl, rel + 1, pl, vy.y);
                        peer_put<false>(p.peer_lo[3], p.peer_hi[3], k, p.nzl, rel + 1, pl, vz.y);
                    }
                }
            }
            // kinetic energy over the PML-free box (:1387-1397)
            if (kg >= p.npml && kg <= p.nz - p.npml + 1) {
                if (eboxA) ekin += p.half_rho * (vx.x * vx.x + vy.x * vy.x + vz.x * vz.x);
                if (eboxB) ekin += p.half_rho * (vx.y * vx.y + vy.y * vy.y + vz.y * vz.y);
            }

            sxz_mm = sxz_m; sxz_m = sxz_c;
            syz_mm = syz_m; syz_m = syz_c;
            szz_m = szz_c; szz_c = szz_p; szz_p = szz_pp;
            rn = rn1;
            rc.advance(SC);
        }
        rn.advance(SN);

        ekin = vcons_sum<NC>(ekin, red, tid);
        if (tid == 0) p.partials[item] = ekin;
    }
}

// ---- launch dispatch ---------------------------------------------------------------

template <int TX, int TY>
static size_t vws_smem(int stages, int xm_bytes)
{
    using G = VGeom<TX, TY>;
    return kBarBytes + 128 + ((size_t)G::CBYTES + 3 * (size_t)xm_bytes) * (size_t)stages + (size_t)G::NBYTES * (size_t)(stages + 1);
}

template <int TX, int TY>
static cudaError_t vws_launch(const ParamsV3D &p, const TmaMaps &tm, const Tile3D &t, cudaStream_t s, int *occ)
{
    const size_t smem = vws_smem<TX, TY>(t.stages, t.xm_bytes);
    constexpr int NT = tile_threads(TX, TY) + 32;
    cudaError_t e = cudaFuncSetAttribute(k_vvelocity3d_ws<TX, TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (occ) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k_vvelocity3d_ws<TX, TY>, NT, smem);
    k_vvelocity3d_ws<TX, TY><<<t.grid_velocity, NT, smem, s>>>(p, tm, t);
    return cudaGetLastError();
}

static cudaError_t vws_dispatch(const ParamsV3D &p, const TmaMaps &tm, const Tile3D &t, cudaStream_t s, int *occ)
{
    switch (t.tx * 100 + t.ty) {
    case 6408:  return vws_launch<64, 8>(p, tm, t, s, occ);        // 256 + 32 threads
    case 10408: return vws_launch<104, 8>(p, tm, t, s, occ);       // 416 + 32
    case 10808: return vws_launch<108, 8>(p, tm, t, s, occ);       // 432 (448) + 32
    default: return cudaErrorInvalidValue;
    }
}

bool vws_tile_supported(int tx, int ty) { return ty == 8 && (tx == 64 || tx == 104 || tx == 108); }

// box extents (x, y) of the nine tensor maps of the kernel: sxx sxy syy sxz syz szz vx vy vz
void vws_boxes(int tx, int ty, int (*box)[2])
{
    const int b[9][2] = {{tx + 4, ty}, {tx + 4, ty + 4}, {tx, ty + 4}, {tx + 4, ty}, {tx, ty + 4}, {tx, ty}, {tx, ty}, {tx, ty}, {tx, ty}};
    for (int m = 0; m < 9; m++) { box[m][0] = b[m][0]; box[m][1] = b[m][1]; }
}

cudaError_t vws_occupancy(const Tile3D &t, int *occ)
{
    ParamsV3D p{};
    TmaMaps dummy{};
    return vws_dispatch(p, dummy, t, nullptr, occ);
}

cudaError_t launch_vvelocity3d_ws(const ParamsV3D &p_in, const TmaMaps &tm, const Tile3D &t, cudaStream_t s)
{
    ParamsV3D p = p_in;
    p.pf = 4;            // memory variables of the next plane into L2 (the streamed words come through the TMA ring)
    return vws_dispatch(p, tm, t, s, nullptr);
}

}  // namespace cpml
