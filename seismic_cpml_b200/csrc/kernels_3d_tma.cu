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
#include "tma_common.cuh"

namespace cpml {

template <bool KUNIT, int TX, int TY, int MINB>
__global__ void __launch_bounds__(tile_threads(TX, TY), MINB)
k_stress3d_tma(const __grid_constant__ Params3D p, const __grid_constant__ TmaMaps tm, const __grid_constant__ Tile3D t)
{
    using G = TileGeom<TX, TY>;
    constexpr int W = G::W;
    constexpr bool EARLY_RELEASE = tile_threads(TX, TY) < 512;             // see release_and_refill below
    constexpr int NBYTES = 2 * G::HALO_BYTES;                       // ring N stage: vx, vy
    constexpr int CBYTES = G::HALO_BYTES + 6 * G::PLAIN_BYTES;      // ring C stage: vz, 6 sigma
    constexpr uint32_t TX_N = 2 * G::HALO_BOX_BYTES;
    constexpr uint32_t TX_C = G::HALO_BOX_BYTES + 6 * G::PLAIN_BOX_BYTES;
    constexpr int PD = G::PLAIN_BYTES / 8;
    // the x-shell memory variables of the tile's rows ride in ring C too: rows of sxp doubles,
    // contiguous over j, one bulk copy per variable (t.xm_bytes each, 0 without an x shell)
    const uint32_t XMB = (uint32_t)t.xm_bytes, XM_TX = (uint32_t)(TY * p.sxp * 8);
    const uint32_t CSTAGE = CBYTES + 3 * XMB;

    __shared__ double Cx[6 * TX];                                   // x coefficients of the tile's columns
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t sbase = (smem_u32(smem_dyn) + 127u) & ~127u;
    const unsigned char *gbase = smem_dyn + (sbase - smem_u32(smem_dyn));
    const uint32_t SC = (uint32_t)t.stages, SN = SC + 1;
    const uint32_t barN = sbase, barC = sbase + 64;                 // up to 8 mbarriers per ring
    const uint32_t ringN = sbase + kBarBytes, ringC = ringN + SN * NBYTES;
    const unsigned char *gN = gbase + kBarBytes, *gC = gN + (size_t)SN * NBYTES;

    constexpr bool PADDED = tile_pairs(TX, TY) % 32 != 0;           // one-dimensional launch with idle threads
    const int tid = PADDED ? (int)threadIdx.x : (int)(threadIdx.y * (TX / 2) + threadIdx.x);
    const int tx = PADDED ? tid % (TX / 2) : (int)threadIdx.x;
    const int ty_raw = PADDED ? tid / (TX / 2) : (int)threadIdx.y;
    const bool lane_ok = !PADDED || ty_raw < TY;
    const int ty = PADDED ? min(ty_raw, TY - 1) : ty_raw;           // idle threads read row TY-1 and store nothing
    if (tid == 0) {
        for (uint32_t s = 0; s < SN; s++) mbar_init(barN + 8 * s, 1);
        for (uint32_t s = 0; s < SC; s++) mbar_init(barC + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int pitch = p.pitch;
    const long long pl = p.plane;
    const int oh = ty * W + 2 * tx;      // halo tile, box origin (0,0): first point of the pair
    const int oc = ty * TX + 2 * tx;     // plain tile

    RingPos rn{0, 0}, rc{0, 0};          // stage of the current plane in each ring
    for (int item = blockIdx.x; item < t.nitems; item += gridDim.x) {
        const int tix = item % t.ntx;
        const int rest = item / t.ntx;
        const int tiy = rest % t.nty;
        const int zc = rest / t.nty;
        const int i0 = 1 + tix * TX, j0 = 1 + tiy * TY;
        const int kb = 1 + zc * t.kchunk;
        const int ke = min(p.nzl, kb + t.kchunk - 1);
        const int np = ke - kb + 1;
        const int x0 = i0 - 1, y0 = j0 - 1;

        // does the tile touch the x shell?  (uniform)
        const bool tile_xpml = XMB != 0 && ((i0 <= p.xlo) || (i0 + TX - 1 >= p.xhi));
        auto issue_c = [&](uint32_t s, int kk) {
            const uint32_t bar = barC + 8 * s, dst = ringC + s * CSTAGE;
            mbar_expect_tx(bar, TX_C + (tile_xpml ? 3 * XM_TX : 0u));
            tma_load_3d(dst, &tm.m[2], x0 - 2, y0, kk, bar);
#pragma unroll
            for (int f = 0; f < 6; f++)
                tma_load_3d(dst + G::HALO_BYTES + f * G::PLAIN_BYTES, &tm.m[3 + f], x0, y0, kk, bar);
            if (tile_xpml) {
                const long long row0 = ((long long)(kk - 1) * p.ny + (j0 - 1)) * p.sxp;
#pragma unroll
                for (int f = 0; f < 3; f++) bulk_load(dst + CBYTES + f * XMB, p.mx[f] + row0, XM_TX, bar);
            }
        };
        // ring N load l = plane kb+l (l = 0..np: the last one only feeds the z differences of
        // plane ke); ring C load l = plane kb+l (l = 0..np-1)
        if (tid == 0) {
            uint32_t s = rn.s;
            for (int l = 0; l < min((int)SN, np + 1); l++) {
                mbar_expect_tx(barN + 8 * s, TX_N);
                tma_load_3d(ringN + s * NBYTES, &tm.m[0], x0, y0, kb + l, barN + 8 * s);
                tma_load_3d(ringN + s * NBYTES + G::HALO_BYTES, &tm.m[1], x0 - 2, y0 - 1, kb + l, barN + 8 * s);
                if (++s == SN) s = 0;
            }
            s = rc.s;
            for (int l = 0; l < min((int)SC, np); l++) {
                issue_c(s, kb + l);
                if (++s == SC) s = 0;
            }
        }

        // the pair: points A = (i, j) and B = (i+1, j); i-1 is even, so B shares A's 16 bytes
        const int i = i0 + 2 * tx, j = j0 + ty;
        const bool row = lane_ok && (j <= p.ny);
        const bool validA = row && (i <= p.nx), validB = row && (i + 1 <= p.nx);
        long long q = (long long)kb * pl + (long long)(j - 1) * pitch + (i - 1);

        const bool in_xA = validA && ((i <= p.xlo) || (i >= p.xhi));
        const bool in_xB = validB && ((i + 1 <= p.xlo) || (i + 1 >= p.xhi));
        const bool in_y = validA && ((j <= p.ylo) || (j >= p.yhi));
        const bool ux = __any_sync(0xffffffffu, in_xA || in_xB), uy = __any_sync(0xffffffffu, in_y);   // warp-uniform
        const bool warp_pml = ux || uy;
        const int jc = min(j, p.ny);                                // y coefficients are read unconditionally
        if (tile_xpml) fill_cx<TX, tile_threads(TX, TY)>(p, Cx, i0, tid);
        __syncthreads();
        // loop bounds of the four nests (i, j part; the k part is tested per plane)
        const bool do_nA = validA && (i <= p.nx - 1) && (j >= 2), do_nB = validB && (i + 1 <= p.nx - 1) && (j >= 2);   // :838-839
        const bool do_xyA = validA && (i >= 2) && (j <= p.ny - 1), do_xyB = validB && (j <= p.ny - 1);                 // :878-879
        const bool do_xzA = validA && (i >= 2), do_xzB = validB;                                                       // :910-911
        const bool do_yzA = validA && (j <= p.ny - 1), do_yzB = validB && (j <= p.ny - 1);                             // :927-928

        const int sxA = in_xA ? shell_index(i, p.xlo, p.xhi) : 0, sxB = in_xB ? shell_index(i + 1, p.xlo, p.xhi) : 0;
        long long qxr = ((long long)(kb - 1) * p.ny + (j - 1)) * p.sxp;                         // x-shell row of this (j, k)
        long long qy = in_y ? ((long long)(kb - 1) * p.sy + shell_index(j, p.ylo, p.yhi)) * pitch + (i - 1) : 0;
        const long long qx_step = (long long)p.ny * p.sxp, qy_step = (long long)p.sy * pitch;

        double vz_mA = 0.0, vz_mB = 0.0;                            // plane kb-1, carried along z
        if (validA) { const double2 t2 = *reinterpret_cast<const double2 *>(p.vz + q - pl); vz_mA = t2.x; vz_mB = t2.y; }

        mbar_wait(barN + 8 * rn.s, rn.par);
        for (int n = 0; n < np; ++n, q += pl, qxr += qx_step, qy += qy_step) {
            const int k = kb + n;
            const int kg = k + p.koff;                              // :837
            // C-PML memory variables of this plane: every load is issued before the waits (and
            // before any store of the recursion, which the compiler must assume to alias)
            const bool z_pml = (kg <= p.zlo) || (kg >= p.zhi);      // uniform
            const bool pml = warp_pml || z_pml;                     // warp-uniform
            const bool in_zA = validA && z_pml, in_zB = validB && z_pml;
            long long qz = 0;
            double mvA[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, mvB[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            if (pml) {
                if (z_pml) qz = ((long long)(shell_index(kg, p.zlo, p.zhi) - p.zbase) * p.ny + (j - 1)) * pitch + (i - 1);
                load_memvars(p, 0, false, uy && in_y, in_zA, 0, qy, qz, mvA);
                load_memvars(p, 0, false, uy && in_y && validB, in_zB, 0, qy + 1, qz + 1, mvB);
            }
            RingPos rn1 = rn;
            rn1.advance(SN);
            mbar_wait(barN + 8 * rn1.s, rn1.par);
            mbar_wait(barC + 8 * rc.s, rc.par);

            const double *Tvx = (const double *)(gN + (size_t)rn.s * NBYTES);
            const double *Tvy = (const double *)(gN + (size_t)rn.s * NBYTES + G::HALO_BYTES);
            const double *Tvxn = (const double *)(gN + (size_t)rn1.s * NBYTES);
            const double *Tvyn = (const double *)(gN + (size_t)rn1.s * NBYTES + G::HALO_BYTES);
            const double *Tvz = (const double *)(gC + (size_t)rc.s * CSTAGE);
            const double *Ts = (const double *)(gC + (size_t)rc.s * CSTAGE + G::HALO_BYTES);
            if (pml) {      // x-shell memory variables of this plane, staged with the tiles
                const double *Tm = (const double *)(gC + (size_t)rc.s * CSTAGE + CBYTES);
                const int xd = (int)(XMB / 8), r0 = ty * p.sxp;
                if (in_xA) { mvA[0] = Tm[r0 + sxA]; mvA[1] = Tm[xd + r0 + sxA]; mvA[2] = Tm[2 * xd + r0 + sxA]; }
                if (in_xB) { mvB[0] = Tm[r0 + sxB]; mvB[1] = Tm[xd + r0 + sxB]; mvB[2] = Tm[2 * xd + r0 + sxB]; }
            }

            const double2 vx_c = lds2(Tvx, oh), vx_jp = lds2(Tvx, oh + W), vx_n = lds2(Tvxn, oh);
            const double vx_ipB = Tvx[oh + 2];
            const double2 vy_c = lds2(Tvy, oh + W + 2), vy_jm = lds2(Tvy, oh + 2), vy_n = lds2(Tvyn, oh + W + 2);
            const double vy_imA = Tvy[oh + W + 1];
            const double2 vz_c = lds2(Tvz, oh + 2), vz_jp = lds2(Tvz, oh + W + 2);
            const double vz_imA = Tvz[oh + 1];
            const double2 sxx = lds2(Ts, 0 * PD + oc), syy = lds2(Ts, 1 * PD + oc), szz = lds2(Ts, 2 * PD + oc);
            const double2 sxy = lds2(Ts, 3 * PD + oc), sxz = lds2(Ts, 4 * PD + oc), syz = lds2(Ts, 5 * PD + oc);

            // Everything this plane needs from the stages rn.s / rc.s is in registers now: hand them
            // back to the TMA unit BEFORE the arithmetic, so that the loads of the planes that reuse
            // them are in flight during this plane's update and not only during the next one's
            // (+2.5 % on the 104-wide tile; the 512-thread 128-wide tile has no registers to spare
            // for it and releases after the stores, profiles/r01_v6_early_release.txt).
            auto release_and_refill = [&]() {
                __syncthreads();                                    // stages rn.s / rc.s are free
                if (tid == 0) {
                    if (n + (int)SN <= np) {
                        const uint32_t s = rn.s, bar = barN + 8 * s;
                        const int kk = kb + n + (int)SN;
                        mbar_expect_tx(bar, TX_N);
                        tma_load_3d(ringN + s * NBYTES, &tm.m[0], x0, y0, kk, bar);
                        tma_load_3d(ringN + s * NBYTES + G::HALO_BYTES, &tm.m[1], x0 - 2, y0 - 1, kk, bar);
                    }
                    if (n + (int)SC < np) issue_c(rc.s, kb + n + (int)SC);
                }
            };
            if (EARLY_RELEASE) release_and_refill();

            StressVals a{sxx.x, syy.x, szz.x, sxy.x, sxz.x, syz.x}, b{sxx.y, syy.y, szz.y, sxy.y, sxz.y, syz.y};
            if (pml) {
                stress_point<true, KUNIT, TX>(p, Cx, 2 * tx, jc, kg, do_nA, do_xyA, do_xzA, do_yzA, ux, uy, z_pml, in_xA, in_y, in_zA,
                                              qxr + sxA, qy, qz, mvA,
                                              vx_c.x, vx_c.y, vx_jp.x, vx_n.x, vy_c.x, vy_imA, vy_jm.x, vy_n.x, vz_c.x, vz_imA, vz_jp.x, vz_mA, a);
                stress_point<true, KUNIT, TX>(p, Cx, 2 * tx + 1, jc, kg, do_nB, do_xyB, do_xzB, do_yzB, ux, uy, z_pml, in_xB, in_y && validB, in_zB,
                                              qxr + sxB, qy + 1, qz + 1, mvB,
                                              vx_c.y, vx_ipB, vx_jp.y, vx_n.y, vy_c.y, vy_c.x, vy_jm.y, vy_n.y, vz_c.y, vz_c.x, vz_jp.y, vz_mB, b);
            } else {
                stress_point<false, KUNIT, TX>(p, Cx, 0, jc, kg, do_nA, do_xyA, do_xzA, do_yzA, false, false, false, false, false, false, 0, 0, 0, mvA,
                                               vx_c.x, vx_c.y, vx_jp.x, vx_n.x, vy_c.x, vy_imA, vy_jm.x, vy_n.x, vz_c.x, vz_imA, vz_jp.x, vz_mA, a);
                stress_point<false, KUNIT, TX>(p, Cx, 0, jc, kg, do_nB, do_xyB, do_xzB, do_yzB, false, false, false, false, false, false, 0, 0, 0, mvB,
                                               vx_c.y, vx_ipB, vx_jp.y, vx_n.y, vy_c.y, vy_c.x, vy_jm.y, vy_n.y, vz_c.y, vz_c.x, vz_jp.y, vz_mB, b);
            }
            vz_mA = vz_c.x; vz_mB = vz_c.y;

            // 16-byte streaming stores; where a nest does not update a point (grid edges, the pad
            // lane of an odd NX) the value loaded from this plane is written back unchanged
            if (validA) {
                st_stream2(p.sxx + q, a.sxx, b.sxx);
                st_stream2(p.syy + q, a.syy, b.syy);
                st_stream2(p.szz + q, a.szz, b.szz);
                st_stream2(p.sxy + q, a.sxy, b.sxy);
                st_stream2(p.sxz + q, a.sxz, b.sxz);
                st_stream2(p.syz + q, a.syz, b.syz);
                // boundary planes go straight into the neighbour slabs' halo planes (:951-963)
                const long long qp = (long long)(j - 1) * pitch + (i - 1);
                if (k == 1 && p.peer_lo[2]) st_stream2(p.peer_lo[2] + qp, a.szz, b.szz);          // sigmazz(:,:,1) -> left
                if (k == p.nzl && p.peer_hi[1]) {                                                 // -> right
                    st_stream2(p.peer_hi[1] + qp, a.sxz, b.sxz);
                    st_stream2(p.peer_hi[2] + qp, a.syz, b.syz);
                }
            }

            if (!EARLY_RELEASE) release_and_refill();
            rn = rn1;
            rc.advance(SC);
        }
        rn.advance(SN);        // plane ke+1 of ring N has been consumed as "next" only
    }
}

template <bool KUNIT, int TX, int TY, int MINB>
__global__ void __launch_bounds__(tile_threads(TX, TY), MINB)
k_velocity3d_tma(const __grid_constant__ Params3D p, const __grid_constant__ TmaMaps tm, const __grid_constant__ Tile3D t)
{
    using G = TileGeom<TX, TY>;
    constexpr int W = G::W;
    constexpr bool EARLY_RELEASE = tile_threads(TX, TY) < 512;
    constexpr int NBYTES = G::PLAIN_BYTES;                          // ring N stage: szz
    constexpr int CBYTES = 5 * G::HALO_BYTES + 3 * G::PLAIN_BYTES;  // ring C stage: 5 sigma (halo), vx vy vz
    constexpr uint32_t TX_N = G::PLAIN_BOX_BYTES;
    constexpr uint32_t TX_C = 5 * G::HALO_BOX_BYTES + 3 * G::PLAIN_BOX_BYTES;
    constexpr int PD = G::PLAIN_BYTES / 8, HD = G::HALO_BYTES / 8;
    constexpr int NT = tile_threads(TX, TY);
    const uint32_t XMB = (uint32_t)t.xm_bytes, XM_TX = (uint32_t)(TY * p.sxp * 8);    // x-shell memory variables, see k_stress3d_tma
    const uint32_t CSTAGE = CBYTES + 3 * XMB;

    __shared__ double red[2 * ((NT + 31) / 32)];
    __shared__ double Cx[6 * TX];                                   // x coefficients of the tile's columns
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t sbase = (smem_u32(smem_dyn) + 127u) & ~127u;
    const unsigned char *gbase = smem_dyn + (sbase - smem_u32(smem_dyn));
    const uint32_t SC = (uint32_t)t.stages, SN = SC + 1;
    const uint32_t barN = sbase, barC = sbase + 64;
    const uint32_t ringN = sbase + kBarBytes, ringC = ringN + SN * NBYTES;
    const unsigned char *gN = gbase + kBarBytes, *gC = gN + (size_t)SN * NBYTES;

    constexpr bool PADDED = tile_pairs(TX, TY) % 32 != 0;           // one-dimensional launch with idle threads
    const int tid = PADDED ? (int)threadIdx.x : (int)(threadIdx.y * (TX / 2) + threadIdx.x);
    const int tx = PADDED ? tid % (TX / 2) : (int)threadIdx.x;
    const int ty_raw = PADDED ? tid / (TX / 2) : (int)threadIdx.y;
    const bool lane_ok = !PADDED || ty_raw < TY;
    const int ty = PADDED ? min(ty_raw, TY - 1) : ty_raw;           // idle threads read row TY-1 and store nothing
    if (tid == 0) {
        for (uint32_t s = 0; s < SN; s++) mbar_init(barN + 8 * s, 1);
        for (uint32_t s = 0; s < SC; s++) mbar_init(barC + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int pitch = p.pitch;
    const long long pl = p.plane;
    const int oh = ty * W + 2 * tx;
    const int oc = ty * TX + 2 * tx;

    RingPos rn{0, 0}, rc{0, 0};
    for (int item = blockIdx.x; item < t.nitems; item += gridDim.x) {
        const int tix = item % t.ntx;
        const int rest = item / t.ntx;
        const int tiy = rest % t.nty;
        const int zc = rest / t.nty;
        const int i0 = 1 + tix * TX, j0 = 1 + tiy * TY;
        const int kb = 1 + zc * t.kchunk;
        const int ke = min(p.nzl, kb + t.kchunk - 1);
        const int np = ke - kb + 1;
        const int x0 = i0 - 1, y0 = j0 - 1;

        const bool tile_xpml = XMB != 0 && ((i0 <= p.xlo) || (i0 + TX - 1 >= p.xhi));
        auto issue_c = [&](uint32_t s, int kk) {
            const uint32_t bar = barC + 8 * s, dst = ringC + s * CSTAGE;
            mbar_expect_tx(bar, TX_C + (tile_xpml ? 3 * XM_TX : 0u));
            if (tile_xpml) {
                const long long row0 = ((long long)(kk - 1) * p.ny + (j0 - 1)) * p.sxp;
#pragma unroll
                for (int f = 0; f < 3; f++) bulk_load(dst + CBYTES + f * XMB, p.mx[3 + f] + row0, XM_TX, bar);
            }
            tma_load_3d(dst + 0 * G::HALO_BYTES, &tm.m[0], x0 - 2, y0, kk, bar);
            tma_load_3d(dst + 1 * G::HALO_BYTES, &tm.m[1], x0, y0, kk, bar);
            tma_load_3d(dst + 2 * G::HALO_BYTES, &tm.m[2], x0, y0 - 1, kk, bar);
            tma_load_3d(dst + 3 * G::HALO_BYTES, &tm.m[3], x0, y0, kk, bar);
            tma_load_3d(dst + 4 * G::HALO_BYTES, &tm.m[4], x0, y0 - 1, kk, bar);
#pragma unroll
            for (int f = 0; f < 3; f++)
                tma_load_3d(dst + 5 * G::HALO_BYTES + f * G::PLAIN_BYTES, &tm.m[6 + f], x0, y0, kk, bar);
        };
        if (tid == 0) {
            uint32_t s = rn.s;
            for (int l = 0; l < min((int)SN, np + 1); l++) {
                mbar_expect_tx(barN + 8 * s, TX_N);
                tma_load_3d(ringN + s * NBYTES, &tm.m[5], x0, y0, kb + l, barN + 8 * s);
                if (++s == SN) s = 0;
            }
            s = rc.s;
            for (int l = 0; l < min((int)SC, np); l++) {
                issue_c(s, kb + l);
                if (++s == SC) s = 0;
            }
        }

        const int i = i0 + 2 * tx, j = j0 + ty;
        const bool row = lane_ok && (j <= p.ny);
        const bool validA = row && (i <= p.nx), validB = row && (i + 1 <= p.nx);
        long long q = (long long)kb * pl + (long long)(j - 1) * pitch + (i - 1);

        const bool in_xA = validA && ((i <= p.xlo) || (i >= p.xhi));
        const bool in_xB = validB && ((i + 1 <= p.xlo) || (i + 1 >= p.xhi));
        const bool in_y = validA && ((j <= p.ylo) || (j >= p.yhi));
        const bool ux = __any_sync(0xffffffffu, in_xA || in_xB), uy = __any_sync(0xffffffffu, in_y);   // warp-uniform
        const bool warp_pml = ux || uy;
        const int jc = min(j, p.ny);
        if (tile_xpml) fill_cx<TX, NT>(p, Cx, i0, tid);
        __syncthreads();
        const bool do_vxA = validA && (i >= 2) && (j >= 2), do_vxB = validB && (j >= 2);                                         // :978-979
        const bool do_vyA = validA && (i <= p.nx - 1) && (j <= p.ny - 1), do_vyB = validB && (i + 1 <= p.nx - 1) && (j <= p.ny - 1);   // :998-999
        const bool do_vzA = validA && (i <= p.nx - 1) && (j >= 2), do_vzB = validB && (i + 1 <= p.nx - 1) && (j >= 2);           // :1033-1034
        const bool edge_j = (j == 1) || (j == p.ny);
        const bool edgeA = (i == 1) || (i == p.nx) || edge_j, edgeB = (i + 1 == p.nx) || edge_j;                                 // :1089-1106
        const bool ebox_j = (j >= p.npml + 1) && (j <= p.ny - p.npml);
        const bool eboxA = validA && ebox_j && (i >= p.npml + 1) && (i <= p.nx - p.npml);                                        // :1144-1145
        const bool eboxB = validB && ebox_j && (i + 1 >= p.npml + 1) && (i + 1 <= p.nx - p.npml);
        const bool srcA = (i == p.isrc) && (j == p.jsrc), srcB = (i + 1 == p.isrc) && (j == p.jsrc);

        const int sxA = in_xA ? shell_index(i, p.xlo, p.xhi) : 0, sxB = in_xB ? shell_index(i + 1, p.xlo, p.xhi) : 0;
        long long qxr = ((long long)(kb - 1) * p.ny + (j - 1)) * p.sxp;
        long long qy = in_y ? ((long long)(kb - 1) * p.sy + shell_index(j, p.ylo, p.yhi)) * pitch + (i - 1) : 0;
        const long long qx_step = (long long)p.ny * p.sxp, qy_step = (long long)p.sy * pitch;

        double sxz_mA = 0.0, sxz_mB = 0.0, syz_mA = 0.0, syz_mB = 0.0;     // plane kb-1
        if (validA) {
            const double2 t2 = *reinterpret_cast<const double2 *>(p.sxz + q - pl), u2 = *reinterpret_cast<const double2 *>(p.syz + q - pl);
            sxz_mA = t2.x; sxz_mB = t2.y; syz_mA = u2.x; syz_mB = u2.y;
        }
        double ekin = 0.0, epot = 0.0;

        mbar_wait(barN + 8 * rn.s, rn.par);
        for (int n = 0; n < np; ++n, q += pl, qxr += qx_step, qy += qy_step) {
            const int k = kb + n;
            const int kg = k + p.koff;
            const bool z_pml = (kg <= p.zlo) || (kg >= p.zhi);      // uniform
            const bool pml = warp_pml || z_pml;                     // warp-uniform
            const bool in_zA = validA && z_pml, in_zB = validB && z_pml;
            long long qz = 0;
            double mvA[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, mvB[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            if (pml) {
                if (z_pml) qz = ((long long)(shell_index(kg, p.zlo, p.zhi) - p.zbase) * p.ny + (j - 1)) * pitch + (i - 1);
                load_memvars(p, 3, false, uy && in_y, in_zA, 0, qy, qz, mvA);
                load_memvars(p, 3, false, uy && in_y && validB, in_zB, 0, qy + 1, qz + 1, mvB);
            }
            RingPos rn1 = rn;
            rn1.advance(SN);
            mbar_wait(barN + 8 * rn1.s, rn1.par);
            mbar_wait(barC + 8 * rc.s, rc.par);

            const double *Tc = (const double *)(gC + (size_t)rc.s * CSTAGE);
            if (pml) {      // x-shell memory variables of this plane, staged with the tiles
                const double *Tm = (const double *)(gC + (size_t)rc.s * CSTAGE + CBYTES);
                const int xd = (int)(XMB / 8), r0 = ty * p.sxp;
                if (in_xA) { mvA[0] = Tm[r0 + sxA]; mvA[1] = Tm[xd + r0 + sxA]; mvA[2] = Tm[2 * xd + r0 + sxA]; }
                if (in_xB) { mvB[0] = Tm[r0 + sxB]; mvB[1] = Tm[xd + r0 + sxB]; mvB[2] = Tm[2 * xd + r0 + sxB]; }
            }
            const double *Txx = Tc, *Tyy = Tc + HD, *Txy = Tc + 2 * HD, *Txz = Tc + 3 * HD, *Tyz = Tc + 4 * HD;
            const double *Tp = Tc + 5 * HD;

            const double2 sxx_c = lds2(Txx, oh + 2);
            const double sxx_imA = Txx[oh + 1];
            const double2 syy_c = lds2(Tyy, oh), syy_jp = lds2(Tyy, oh + W);
            const double2 sxy_c = lds2(Txy, oh + W), sxy_jm = lds2(Txy, oh);
            const double sxy_ipB = Txy[oh + W + 2];
            const double2 sxz_c = lds2(Txz, oh);
            const double sxz_ipB = Txz[oh + 2];
            const double2 syz_c = lds2(Tyz, oh + W), syz_jm = lds2(Tyz, oh);
            const double2 szz_c = lds2((const double *)(gN + (size_t)rn.s * NBYTES), oc);
            const double2 szz_n = lds2((const double *)(gN + (size_t)rn1.s * NBYTES), oc);
            const double2 vx = lds2(Tp, 0 * PD + oc), vy = lds2(Tp, 1 * PD + oc), vz = lds2(Tp, 2 * PD + oc);

            // stages rn.s / rc.s are in registers: free them before the arithmetic (see k_stress3d_tma)
            auto release_and_refill = [&]() {
                __syncthreads();
                if (tid == 0) {
                    if (n + (int)SN <= np) {
                        const uint32_t s = rn.s, bar = barN + 8 * s;
                        mbar_expect_tx(bar, TX_N);
                        tma_load_3d(ringN + s * NBYTES, &tm.m[5], x0, y0, kb + n + (int)SN, bar);
                    }
                    if (n + (int)SC < np) issue_c(rc.s, kb + n + (int)SC);
                }
            };
            if (EARLY_RELEASE) release_and_refill();

            VelVals a{vx.x, vy.x, vz.x}, b{vx.y, vy.y, vz.y};
            if (pml) {
                velocity_point<true, KUNIT, TX>(p, Cx, 2 * tx, jc, k, kg, do_vxA, do_vyA, do_vzA, edgeA, eboxA, srcA, ux, uy, z_pml, in_xA, in_y, in_zA,
                                                qxr + sxA, qy, qz, mvA,
                                                sxx_c.x, sxx_imA, syy_c.x, syy_jp.x, sxy_c.x, sxy_jm.x, sxy_c.y, sxz_c.x, sxz_c.y, sxz_mA,
                                                syz_c.x, syz_jm.x, syz_mA, szz_c.x, szz_n.x, a, ekin, epot);
                velocity_point<true, KUNIT, TX>(p, Cx, 2 * tx + 1, jc, k, kg, do_vxB, do_vyB, do_vzB, edgeB, eboxB, srcB, ux, uy, z_pml, in_xB, in_y && validB, in_zB,
                                                qxr + sxB, qy + 1, qz + 1, mvB,
                                                sxx_c.y, sxx_c.x, syy_c.y, syy_jp.y, sxy_c.y, sxy_jm.y, sxy_ipB, sxz_c.y, sxz_ipB, sxz_mB,
                                                syz_c.y, syz_jm.y, syz_mB, szz_c.y, szz_n.y, b, ekin, epot);
            } else {
                velocity_point<false, KUNIT, TX>(p, Cx, 0, jc, k, kg, do_vxA, do_vyA, do_vzA, edgeA, eboxA, srcA, false, false, false, false, false, false,
                                                 0, 0, 0, mvA,
                                                 sxx_c.x, sxx_imA, syy_c.x, syy_jp.x, sxy_c.x, sxy_jm.x, sxy_c.y, sxz_c.x, sxz_c.y, sxz_mA,
                                                 syz_c.x, syz_jm.x, syz_mA, szz_c.x, szz_n.x, a, ekin, epot);
                velocity_point<false, KUNIT, TX>(p, Cx, 0, jc, k, kg, do_vxB, do_vyB, do_vzB, edgeB, eboxB, srcB, false, false, false, false, false, false,
                                                 0, 0, 0, mvB,
                                                 sxx_c.y, sxx_c.x, syy_c.y, syy_jp.y, sxy_c.y, sxy_jm.y, sxy_ipB, sxz_c.y, sxz_ipB, sxz_mB,
                                                 syz_c.y, syz_jm.y, syz_mB, szz_c.y, szz_n.y, b, ekin, epot);
            }
            sxz_mA = sxz_c.x; sxz_mB = sxz_c.y; syz_mA = syz_c.x; syz_mB = syz_c.y;

            if (validA) {
                // the pad lane of an odd NX keeps its zero: B is then outside every nest, not an
                // edge point, and its loaded value is the TMA's zero fill
                if (!validB) { b.vx = 0.0; b.vy = 0.0; b.vz = 0.0; }
                st_stream2(p.vx + q, a.vx, b.vx);
                st_stream2(p.vy + q, a.vy, b.vy);
                st_stream2(p.vz + q, a.vz, b.vz);
                // boundary planes go straight into the neighbour slabs' halo planes (:811-823)
                const long long qp = (long long)(j - 1) * pitch + (i - 1);
                if (k == 1 && p.peer_lo[0]) {                                                   // -> left
                    st_stream2(p.peer_lo[0] + qp, a.vx, b.vx);
                    st_stream2(p.peer_lo[1] + qp, a.vy, b.vy);
                }
                if (k == p.nzl && p.peer_hi[0]) st_stream2(p.peer_hi[0] + qp, a.vz, b.vz);      // -> right
            }

            if (!EARLY_RELEASE) release_and_refill();
            rn = rn1;
            rc.advance(SC);
        }
        rn.advance(SN);

        block_sum2<NT>(ekin, epot, red);
        if (tid == 0) {
            p.partials[item] = ekin;
            p.partials[p.nblocks + item] = epot;
        }
    }
}

// ---- launch dispatch ---------------------------------------------------------------

template <int TX, int TY>
static size_t smem_need(bool stress, int stages, int xm_bytes)
{
    using G = TileGeom<TX, TY>;
    // ring C holds `stages` planes, ring N one more (kernels above)
    const size_t c = stress ? G::HALO_BYTES + 6 * G::PLAIN_BYTES : 5 * G::HALO_BYTES + 3 * G::PLAIN_BYTES;
    const size_t n = stress ? 2 * G::HALO_BYTES : G::PLAIN_BYTES;
    return kBarBytes + 128 + (c + 3 * (size_t)xm_bytes) * (size_t)stages + n * (size_t)(stages + 1);
}

template <bool KUNIT, int TX, int TY, int MINB>
static cudaError_t launch_tile(const Params3D &p, const TmaMaps &tm, const Tile3D &t, cudaStream_t s, bool stress, int *occ)
{
    const size_t smem = smem_need<TX, TY>(stress, t.stages, t.xm_bytes);
    constexpr int NT = tile_threads(TX, TY);        // one thread per pair of points, whole warps
    const void *fn = stress ? (const void *)k_stress3d_tma<KUNIT, TX, TY, MINB> : (const void *)k_velocity3d_tma<KUNIT, TX, TY, MINB>;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (occ) {
        if (stress) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k_stress3d_tma<KUNIT, TX, TY, MINB>, NT, smem);
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k_velocity3d_tma<KUNIT, TX, TY, MINB>, NT, smem);
    }
    const dim3 grid(stress ? t.grid_stress : t.grid_velocity);
    const dim3 block = tile_pairs(TX, TY) % 32 != 0 ? dim3(NT) : dim3(TX / 2, TY);
    if (stress) k_stress3d_tma<KUNIT, TX, TY, MINB><<<grid, block, smem, s>>>(p, tm, t);
    else        k_velocity3d_tma<KUNIT, TX, TY, MINB><<<grid, block, smem, s>>>(p, tm, t);
    return cudaGetLastError();
}

// Tiles (TX x TY points, TX/2 x TY threads, a multiple of 32); MINB (minimum resident CTAs per
// SM) caps the registers at 65536 / (threads * MINB).
template <bool KUNIT>
static cudaError_t dispatch_tile(const Params3D &p, const TmaMaps &tm, const Tile3D &t, cudaStream_t s, bool stress, int *occ)
{
    switch (t.tx * 1000 + t.ty * 10 + t.minb) {
    case 64041:  case 64042:  return launch_tile<KUNIT, 64, 4, 2>(p, tm, t, s, stress, occ);      // 128 threads
    case 64043:  case 64044:  return launch_tile<KUNIT, 64, 4, 4>(p, tm, t, s, stress, occ);
    case 64081:  case 64082:  return launch_tile<KUNIT, 64, 8, 2>(p, tm, t, s, stress, occ);      // 256 threads
    case 128041: case 128042: return launch_tile<KUNIT, 128, 4, 2>(p, tm, t, s, stress, occ);     // 256 threads
    case 128043: return launch_tile<KUNIT, 128, 4, 3>(p, tm, t, s, stress, occ);
    case 128061: return launch_tile<KUNIT, 128, 6, 1>(p, tm, t, s, stress, occ);                  // 384 threads: 168-register cap
    case 128081: return launch_tile<KUNIT, 128, 8, 1>(p, tm, t, s, stress, occ);                  // 512 threads
    case 128082: return launch_tile<KUNIT, 128, 8, 2>(p, tm, t, s, stress, occ);
    case 104041: case 104042: return launch_tile<KUNIT, 104, 4, 2>(p, tm, t, s, stress, occ);     // 208 threads
    case 104043: case 104044: return launch_tile<KUNIT, 104, 4, 3>(p, tm, t, s, stress, occ);
    case 104071: return launch_tile<KUNIT, 104, 7, 1>(p, tm, t, s, stress, occ);                  // 364 pairs on 384 threads
    case 104081: return launch_tile<KUNIT, 104, 8, 1>(p, tm, t, s, stress, occ);                  // 416 threads
    case 104082: return launch_tile<KUNIT, 104, 8, 2>(p, tm, t, s, stress, occ);
    default: return cudaErrorInvalidValue;
    }
}

bool tma_tile_supported(int tx, int ty)
{
    switch (tx * 100 + ty) {
    case 6404: case 6408: case 12804: case 12808: case 10408: case 10404: case 10407: case 12806: return true;     // TMA boxes hold at most 256 elements per dimension
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
