// 3-D isotropic C-PML kernels for sm_100a, TMA-staged with a dedicated PRODUCER WARP (the default path).
//
//   k_stress3d_ws    sigmaxx/yy/zz (:836-863), sigmaxy (:877-894), sigmaxz/yz (:908-943)
//   k_velocity3d_ws  vx/vy (:976-1017), vz (:1031-1052), source (:1055-1083), Dirichlet faces (:1087-1121),
//                    energy partials (:1131-1177)
// (line numbers: seismic_CPML_3D_isotropic_MPI_OpenMP.f90; the per-point arithmetic is tma_common.cuh's
// stress_point / velocity_point, shared with kernels_3d_tma.cu, so both families give the same bits.)
//
// Why a producer warp.  In kernels_3d_tma.cu thread 0 issues the nine tensor loads of a plane after a CTA-wide
// barrier; ncu (profiles/r01_v9_ncu_cfg3.txt) shows 1.65-1.78 barrier-stall cycles per issued instruction out of
// ~7: warp 0 executes ~170 extra instructions per plane (expect_tx, nine UTMALDG with their uniform-datapath
// election loops, three UBLKCP) and every other warp waits for it at the next barrier, plane after plane.  Here
// the compute warps ("consumers") never meet in a barrier inside the plane loop:
//   * consumers wait on the stage's "full" mbarrier, move the plane's values from shared memory to registers and
//     ARRIVE (bar.arrive, non-blocking) on the stage's named barrier, then update and store;
//   * the producer warp SYNCs (bar.sync) on that named barrier -- it completes when every consumer has read the
//     stage -- and refills the stage through the TMA unit.  It also starts the first loads of the next work item
//     while the consumers finish the current one, so item boundaries no longer drain the pipeline.
// The C-PML work is split three ways per warp and plane (all warp-uniform): no shell / x shell only (memory
// variables staged in the ring with the tiles, coefficient table in shared memory, branch-free over the lanes) /
// general (y or z shell too: those memory variables come from global memory, loads hoisted above the waits).
//
// Work distribution (persistent CTAs, one per SM): the producer's lead lane claims work items from a queue in global
// memory (tma_common.cuh: claim_item) and posts them to the consumers through item_slot[] / item_bar -- SMs that get
// more of the HBM bandwidth take more items, instead of idling behind the slowest static share -- and the items at the
// end of the list are shorter (decode_item), so the launch ends within a short item on every SM.  The
// single-precision kernels also re-deal the threads of a tile per item so that the x-shell lanes sit in their own
// warps (map_lane_packed).
//
// Slab decomposition: the boundary planes every neighbour needs (:811-823, :951-963) are stored by the same
// kernels straight into the neighbour GPU's halo planes over NVLink, AND the ordering between slabs is done
// inside the kernels as well (SlabSync): only the work items that read a halo plane poll the neighbour's flag
// (the producer warp for plane NZ_LOCAL+1, thread 0 of the consumers for plane 0), and the CTA that completes
// the last boundary item of a side publishes this slab's flag to that neighbour.  Boundary chunks are processed
// first, so a neighbour's planes have a whole kernel's time to arrive before they are needed: the exchange
// overlaps the interior update, and no wait / signal launches remain.
#include "tma_common.cuh"

namespace cpml {

namespace {

constexpr int kConsBar = 1;      // named barrier of the consumer warps
constexpr int kRelBar0 = 2;      // named barriers kRelBar0 + stage: "stage has been read" (consumers arrive, producer syncs)

__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// Sum of (a, b) over the NC consumer threads; result valid in thread 0.
template <int NC>
__device__ __forceinline__ void cons_sum2(double &a, double &b, double *red, int tid)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, o);
        b += __shfl_down_sync(0xffffffffu, b, o);
    }
    constexpr int NW = NC / 32;
    const int w = tid >> 5, l = tid & 31;
    if (l == 0) { red[w] = a; red[NW + w] = b; }
    bar_sync(kConsBar, NC);
    if (w == 0) {
        a = (l < NW) ? red[l] : 0.0;
        b = (l < NW) ? red[NW + l] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_down_sync(0xffffffffu, a, o);
            b += __shfl_down_sync(0xffffffffu, b, o);
        }
    }
    bar_sync(kConsBar, NC);      // red[] is reused by the next work item
}

// Spins until the neighbour slab has published `value` (or later) in this slab's flag word.
__device__ __forceinline__ void poll_flag(const unsigned long long *flag, unsigned long long value, unsigned int *timeout_flag)
{
    unsigned int spins = 0;
    while (true) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
        if (v >= value) break;
        if (++spins > (1u << 25)) { *timeout_flag = 1u; break; }     // tens of seconds: the neighbour is gone
        __nanosleep(100);
    }
}

// Called by ONE thread after a barrier that follows the CTA's last store of a boundary item: counts the item;
// the CTA that completes the side's last boundary item publishes the flag to the neighbour.
__device__ __forceinline__ void publish_side(unsigned int *counter, int n_items, unsigned long long *peer_flag, unsigned long long value)
{
    __threadfence_system();                       // this CTA's peer stores (ordered before by the barrier) precede the count
    const unsigned int done = atomicAdd(counter, 1u);
    if (done == (unsigned int)(n_items - 1)) {
        __threadfence_system();                   // every counted CTA's stores precede the flag
        *counter = 0u;                            // ready for the next launch (kernels of a handle never overlap)
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peer_flag), "l"(value) : "memory");
    }
}

// Work item -> (x tile, y tile, planes kb..ke).  Chunk order 0, nzc-1, 1, 2, ...: both boundary chunks first (see
// above).  The items at the END of the list are finer: every coarse item from t.fine_from on is t.split items of
// kchunk / split planes.  With the queue the CTAs finish within one item of each other, so the idle tail of a launch is
// about half an item per SM; short items at the end cut it without paying their start-up cost everywhere.
struct Item { int tix, tiy, kb, ke; };
__host__ __device__ __forceinline__ Item decode_item(int item, const Tile3D &t, int nzl)
{
    int ci = item, part = 0;
    const bool fine = item >= t.fine_from;
    if (fine) {
        const int r = item - t.fine_from, c = r / t.split;
        ci = t.fine_from + c;
        part = r - c * t.split;
    }
    Item it;
    it.tix = ci % t.ntx;
    const int rest = ci / t.ntx;
    it.tiy = rest % t.nty;
    const int o = rest / t.nty;
    const int zc = (o == 0) ? 0 : (o == 1) ? t.nzc - 1 : o - 1;
    it.kb = 1 + zc * t.kchunk;
    it.ke = min(nzl, it.kb + t.kchunk - 1);
    if (fine) {     // (never a boundary chunk, never the short last chunk: the host keeps those whole)
        const int sub = (t.kchunk + t.split - 1) / t.split;
        it.kb += part * sub;
        it.ke = min(it.ke, it.kb + sub - 1);
    }
    return it;
}

// Single precision only: consumer thread -> (pair of the tile row, row of the tile), recomputed per work item.  The
// pairs of a row are [0, nl) in the left x shell, [nl, nr) interior, [nr, TX/2) in the right x shell or beyond NX.
// Threads are dealt to the interior pairs first, row by row, then to the others: a warp then holds x-shell lanes or
// interior lanes, rarely both, and only the warps that hold shell lanes run the x part of the C-PML code (warp-uniform
// `ux`): on the 101-wide default grid 3 of the 13 consumer warps instead of all of them (velocity kernel 0.610 ->
// 0.555 ms, profiles/r02_n_bench_lane_map.txt).  Any 0 <= nl <= nr <= TX/2 gives a one-to-one map, so the
// classification is a matter of speed only: the per-point predicates are computed from (i, j) as before.  The
// double-precision kernels keep the fixed row-major map: at their 128-register cap the per-item map costs more in
// spills than it saves, and their stress kernel wants whole sectors per warp (same file).
struct Lane { int tx, ty; bool ok; };
template <int TX, int TY>
__host__ __device__ __forceinline__ Lane map_lane_packed(int tid, int i0, int xlo, int xhi, int nx)
{
    constexpr int HP = TX / 2;
    const int nl = (xlo >= i0) ? min(HP, (xlo - i0) / 2 + 1) : 0;            // pairs whose first column is <= xlo
    const int nr = max(nl, min(HP, max(0, min(xhi, nx + 1) - i0) / 2));      // first pair whose last column is >= xhi or > NX
    const int ni = nr - nl, ns = HP - ni;
    int tx = 0, tyr = TY;                                                    // (idle thread)
    if (tid < ni * TY) {
        tyr = tid / ni;
        tx = nl + (tid - tyr * ni);
    } else if (ns > 0) {
        const int u = tid - ni * TY;
        tyr = u / ns;
        const int e = u - tyr * ns;
        tx = e < nl ? e : nr + (e - nl);
    }
    Lane l;
    l.ok = tyr < TY;
    l.ty = min(tyr, TY - 1);
    l.tx = tx;
    return l;
}

}  // namespace

// ------------------------------------------------------------------------------- stress
// ring N (planes n and n+1 are needed):  vx (halo box at (0,0)), vy (halo, (-2,-1))
// ring C (plane n only):                 vz (halo, (-2,0)), sxx syy szz sxy sxz syz (plain boxes),
//                                        x-shell memory variables of the tile rows (three bulk copies)
// maps: 0 vx 1 vy 2 vz 3..8 sigma
template <typename T, bool KUNIT, int TX, int TY>
__global__ void __launch_bounds__(tile_threads(TX, TY) + 32, 1)
k_stress3d_ws(const __grid_constant__ Params3DT<T> p, const __grid_constant__ TmaMaps tm, const __grid_constant__ Tile3D t,
              const __grid_constant__ SlabSync ss)
{
    using G = TileGeomT<T, TX, TY>;
    constexpr int W = G::W, HX = G::HX, ES = G::ES;
    constexpr int NC = tile_threads(TX, TY), NALL = NC + 32;
    constexpr int NBYTES = 2 * G::HALO_BYTES;                       // ring N stage: vx, vy
    constexpr int CBYTES = G::HALO_BYTES + 6 * G::PLAIN_BYTES;      // ring C stage: vz, 6 sigma
    constexpr uint32_t TX_N = 2 * G::HALO_BOX_BYTES;
    constexpr uint32_t TX_C = G::HALO_BOX_BYTES + 6 * G::PLAIN_BOX_BYTES;
    constexpr int PD = G::PLAIN_BYTES / ES;
    const uint32_t XMB = (uint32_t)t.xm_bytes, XM_TX = (uint32_t)(TY * p.sxp * ES);
    const uint32_t CSTAGE = CBYTES + 3 * XMB;

    __shared__ T Cx[6 * TX];                                   // x coefficients of the tile's columns
    __shared__ unsigned long long item_bar;
    __shared__ int item_slot[2];                               // the producer posts the CTA's work items here
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t sbase = (smem_u32(smem_dyn) + 127u) & ~127u;
    const unsigned char *gbase = smem_dyn + (sbase - smem_u32(smem_dyn));
    const uint32_t SC = (uint32_t)t.stages, SN = SC + 1;
    const uint32_t barN = sbase, barC = sbase + 64;                 // up to 8 mbarriers per ring
    const uint32_t ringN = sbase + kBarBytes, ringC = ringN + SN * NBYTES;
    const unsigned char *gN = gbase + kBarBytes, *gC = gN + (size_t)SN * NBYTES;

    const uint32_t barI = smem_u32(&item_bar);                      // "work item it & 1 has been posted"

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
            // post the item (or the end mark) to the consumers; slot ip & 1 was last read two items ago
            if (lead) { item_slot[ip & 1u] = item; mbar_arrive(barI); }
            ++ip;
            if (item >= t.nitems) break;
            const Item it = decode_item(item, t, p.nzl);
            const int x0 = it.tix * TX, y0 = it.tiy * TY, j0 = 1 + y0;
            const int kb = it.kb, ke = it.ke;
            const int np = ke - kb + 1;
            const bool tile_xpml = XMB != 0 && ((x0 + 1 <= p.xlo) || (x0 + TX >= p.xhi));
            auto issue_n = [&](uint32_t s, int kk) {
                const uint32_t bar = barN + 8 * s;
                mbar_expect_tx(bar, TX_N);
                tma_load_3d(ringN + s * NBYTES, &tm.m[0], x0, y0, kk, bar);
                tma_load_3d(ringN + s * NBYTES + G::HALO_BYTES, &tm.m[1], x0 - HX, y0 - 1, kk, bar);
            };
            auto issue_c = [&](uint32_t s, int kk) {
                const uint32_t bar = barC + 8 * s, dst = ringC + s * CSTAGE;
                mbar_expect_tx(bar, TX_C + (tile_xpml ? 3 * XM_TX : 0u));
                tma_load_3d(dst, &tm.m[2], x0 - HX, y0, kk, bar);
#pragma unroll
                for (int f = 0; f < 6; f++)
                    tma_load_3d(dst + G::HALO_BYTES + f * G::PLAIN_BYTES, &tm.m[3 + f], x0, y0, kk, bar);
                if (tile_xpml) {
                    const long long row0 = ((long long)(kk - 1) * p.ny + (j0 - 1)) * p.sxp;
#pragma unroll
                    for (int f = 0; f < 3; f++) bulk_load(dst + CBYTES + f * XMB, p.mx[f] + row0, XM_TX, bar);
                }
            };
            if (lead) {
                // the item's last ring-N load is the upper neighbour's plane NZ_LOCAL+1 (vx, vy): it must have arrived
                if (ss.wait_hi && ke == p.nzl) poll_flag(ss.wait_hi, ss.wait_value, ss.timeout);
                uint32_t s = sn;
                for (int l = 0; l < min((int)SN, np + 1); l++) { issue_n(s, kb + l); if (++s == SN) s = 0; }
                s = sc;
                for (int l = 0; l < min((int)SC, np); l++) { issue_c(s, kb + l); if (++s == SC) s = 0; }
            }
            int next = item + (int)gridDim.x;
            if (lead) next = claim_item(t.queue, next);      // the answer is needed after the plane loop
            for (int n = 0; n < np; ++n) {
                bar_sync(kRelBar0 + (int)sc, NALL);                 // every consumer has read plane kb+n out of the stages
                if (lead) {
                    if (n + (int)SN <= np) issue_n(sn, kb + n + (int)SN);
                    if (n + (int)SC < np) issue_c(sc, kb + n + (int)SC);
                }
                if (++sn == SN) sn = 0;
                if (++sc == SC) sc = 0;
            }
            if (++sn == SN) sn = 0;        // plane ke+1 of ring N has been consumed as "next" only
            item = __shfl_sync(0xffffffffu, next, 0);
        }
        if (lead) retire_queue(t.queue);
        return;
    }

    // ================================================================ consumer warps
    constexpr bool PACK = sizeof(T) == 4;                           // per-item thread map (map_lane_packed)
    const int tx_rm = tid % (TX / 2);                               // the fixed row-major map
    const int ty_raw = tid / (TX / 2);
    const bool ok_rm = ty_raw < TY;                                 // tiles whose pair count is not a multiple of 32
    const int ty_rm = min(ty_raw, TY - 1);                          // idle threads read row TY-1 and store nothing
    const int pitch = p.pitch;
    const unsigned pl = (unsigned)p.plane;      // element offsets fit 32 bits (checked on the host)

    RingPos rn{0, 0}, rc{0, 0};          // stage of the current plane in each ring
    for (uint32_t ip = 0;; ++ip) {
        mbar_wait(barI, ip & 1u);
        const int item = item_slot[ip & 1u];
        if (item >= t.nitems) break;
        const Item itm = decode_item(item, t, p.nzl);
        const int i0 = 1 + itm.tix * TX, j0 = 1 + itm.tiy * TY;
        const int kb = itm.kb, ke = itm.ke;
        const int np = ke - kb + 1;
        const bool tile_xpml = XMB != 0 && ((i0 <= p.xlo) || (i0 + TX - 1 >= p.xhi));
        // plane 0 (vz of the lower neighbour) is read straight from global memory below
        if (ss.wait_lo && kb == 1 && tid == 0) poll_flag(ss.wait_lo, ss.wait_value, ss.timeout);

        // the pair: points A = (i, j) and B = (i+1, j); i-1 is even, so B shares A's 16 bytes
        Lane ln{tx_rm, ty_rm, ok_rm};
        if constexpr (PACK) ln = map_lane_packed<TX, TY>(tid, i0, p.xlo, p.xhi, p.nx);
        const int tx = ln.tx, ty = ln.ty;
        const int oh = ty * W + 2 * tx;      // halo tile, box origin (0,0): first point of the pair
        const int oc = ty * TX + 2 * tx;     // plain tile
        const int i = i0 + 2 * tx, j = j0 + ty;
        const bool row = ln.ok && (j <= p.ny);
        const bool validA = row && (i <= p.nx), validB = row && (i + 1 <= p.nx);
        const bool storeA = row && (i + 1 <= pitch);                // validA or a pad pair of the row
        unsigned q = (unsigned)kb * pl + (unsigned)((j - 1) * pitch + (i - 1));

        const bool in_xA = validA && ((i <= p.xlo) || (i >= p.xhi));
        const bool in_xB = validB && ((i + 1 <= p.xlo) || (i + 1 >= p.xhi));
        const bool in_y = validA && ((j <= p.ylo) || (j >= p.yhi));
        const bool ux = __any_sync(0xffffffffu, in_xA || in_xB), uy = __any_sync(0xffffffffu, in_y);   // warp-uniform
        const int jc = min(j, p.ny);                                // y coefficients are read unconditionally
        if (tile_xpml) fill_cx<TX, NC>(p, Cx, i0, tid);
        bar_sync(kConsBar, NC);
        // loop bounds of the four nests (i, j part; the k part is tested per plane)
        const bool do_nA = validA && (i <= p.nx - 1) && (j >= 2), do_nB = validB && (i + 1 <= p.nx - 1) && (j >= 2);   // :838-839
        const bool do_xyA = validA && (i >= 2) && (j <= p.ny - 1), do_xyB = validB && (j <= p.ny - 1);                 // :878-879
        const bool do_xzA = validA && (i >= 2), do_xzB = validB;                                                       // :910-911
        const bool do_yzA = validA && (j <= p.ny - 1), do_yzB = validB && (j <= p.ny - 1);                             // :927-928

        const int sxA = in_xA ? shell_index(i, p.xlo, p.xhi) : 0, sxB = in_xB ? shell_index(i + 1, p.xlo, p.xhi) : 0;
        unsigned qxr = (unsigned)(((kb - 1) * p.ny + (j - 1)) * p.sxp);                         // x-shell row of this (j, k)
        unsigned qy = in_y ? (unsigned)(((kb - 1) * p.sy + shell_index(j, p.ylo, p.yhi)) * pitch + (i - 1)) : 0u;
        const unsigned qx_step = (unsigned)(p.ny * p.sxp), qy_step = (unsigned)(p.sy * pitch);

        T vz_mA = 0, vz_mB = 0;                                     // plane kb-1, carried along z
        if (validA) { const auto t2 = ldg2(p.vz + q - pl); vz_mA = t2.x; vz_mB = t2.y; }

        mbar_wait(barN + 8 * rn.s, rn.par);
        for (int n = 0; n < np; ++n, q += pl, qxr += qx_step, qy += qy_step) {
            const int k = kb + n;
            const int kg = k + p.koff;                              // :837
            const bool z_pml = (kg <= p.zlo) || (kg >= p.zhi);      // uniform
            const bool gen = uy || z_pml;                           // warp-uniform: y / z shell memory variables needed
            const bool in_zA = validA && z_pml, in_zB = validB && z_pml;
            unsigned qz = 0;
            T mvA[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, mvB[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            if (gen) {      // issued before the waits (and before any store of the recursion, which may alias)
                if (z_pml) qz = (unsigned)(((shell_index(kg, p.zlo, p.zhi) - p.zbase) * p.ny + (j - 1)) * pitch + (i - 1));
                load_memvars(p, 0, false, uy && in_y, in_zA, 0, qy, qz, mvA);
                load_memvars(p, 0, false, uy && in_y && validB, in_zB, 0, qy + 1, qz + 1, mvB);
            }
            RingPos rn1 = rn;
            rn1.advance(SN);
            mbar_wait(barN + 8 * rn1.s, rn1.par);
            mbar_wait(barC + 8 * rc.s, rc.par);

            const T *Tvx = (const T *)(gN + (size_t)rn.s * NBYTES);
            const T *Tvy = (const T *)(gN + (size_t)rn.s * NBYTES + G::HALO_BYTES);
            const T *Tvxn = (const T *)(gN + (size_t)rn1.s * NBYTES);
            const T *Tvyn = (const T *)(gN + (size_t)rn1.s * NBYTES + G::HALO_BYTES);
            const T *Tvz = (const T *)(gC + (size_t)rc.s * CSTAGE);
            const T *Ts = (const T *)(gC + (size_t)rc.s * CSTAGE + G::HALO_BYTES);
            if (ux) {       // x-shell memory variables of this plane, staged with the tiles
                const T *Tm = (const T *)(gC + (size_t)rc.s * CSTAGE + CBYTES);
                const int xd = (int)(XMB / ES), r0 = ty * p.sxp;
                if (in_xA) { mvA[0] = Tm[r0 + sxA]; mvA[1] = Tm[xd + r0 + sxA]; mvA[2] = Tm[2 * xd + r0 + sxA]; }
                if (in_xB) { mvB[0] = Tm[r0 + sxB]; mvB[1] = Tm[xd + r0 + sxB]; mvB[2] = Tm[2 * xd + r0 + sxB]; }
            }

            // (boxes that start HX cells before the tile hold point A at offset +HX)
            const auto vx_c = lds2(Tvx, oh), vx_jp = lds2(Tvx, oh + W), vx_n = lds2(Tvxn, oh);
            const T vx_ipB = Tvx[oh + 2];
            const auto vy_c = lds2(Tvy, oh + W + HX), vy_jm = lds2(Tvy, oh + HX), vy_n = lds2(Tvyn, oh + W + HX);
            const T vy_imA = Tvy[oh + W + HX - 1];
            const auto vz_c = lds2(Tvz, oh + HX), vz_jp = lds2(Tvz, oh + W + HX);
            const T vz_imA = Tvz[oh + HX - 1];
            const auto sxx = lds2(Ts, 0 * PD + oc), syy = lds2(Ts, 1 * PD + oc), szz = lds2(Ts, 2 * PD + oc);
            const auto sxy = lds2(Ts, 3 * PD + oc), sxz = lds2(Ts, 4 * PD + oc), syz = lds2(Ts, 5 * PD + oc);

            // everything this plane needs from stages rn.s / rc.s is in registers: hand them back to the producer
            bar_arrive(kRelBar0 + (int)rc.s, NALL);

            StressValsT<T> a{sxx.x, syy.x, szz.x, sxy.x, sxz.x, syz.x}, b{sxx.y, syy.y, szz.y, sxy.y, sxz.y, syz.y};
            if (gen) {
                stress_point<true, KUNIT, TX>(p, Cx, 2 * tx, jc, kg, do_nA, do_xyA, do_xzA, do_yzA, ux, uy, z_pml, in_xA, in_y, in_zA,
                                              qxr + sxA, qy, qz, mvA,
                                              vx_c.x, vx_c.y, vx_jp.x, vx_n.x, vy_c.x, vy_imA, vy_jm.x, vy_n.x, vz_c.x, vz_imA, vz_jp.x, vz_mA, a);
                stress_point<true, KUNIT, TX>(p, Cx, 2 * tx + 1, jc, kg, do_nB, do_xyB, do_xzB, do_yzB, ux, uy, z_pml, in_xB, in_y && validB, in_zB,
                                              qxr + sxB, qy + 1, qz + 1, mvB,
                                              vx_c.y, vx_ipB, vx_jp.y, vx_n.y, vy_c.y, vy_c.x, vy_jm.y, vy_n.y, vz_c.y, vz_c.x, vz_jp.y, vz_mB, b);
            } else if (ux) {
                stress_point<true, KUNIT, TX>(p, Cx, 2 * tx, jc, kg, do_nA, do_xyA, do_xzA, do_yzA, true, false, false, in_xA, false, false,
                                              qxr + sxA, 0, 0, mvA,
                                              vx_c.x, vx_c.y, vx_jp.x, vx_n.x, vy_c.x, vy_imA, vy_jm.x, vy_n.x, vz_c.x, vz_imA, vz_jp.x, vz_mA, a);
                stress_point<true, KUNIT, TX>(p, Cx, 2 * tx + 1, jc, kg, do_nB, do_xyB, do_xzB, do_yzB, true, false, false, in_xB, false, false,
                                              qxr + sxB, 0, 0, mvB,
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
            // (pad pairs between NX and the row pitch are stored too -- their values are the TMA's zero fill, no nest
            // touches them -- so that every row, hence every 32-byte sector of the plane, is written whole)
            if (storeA) {
                st_stream2(p.sxx + q, a.sxx, b.sxx);
                st_stream2(p.syy + q, a.syy, b.syy);
                st_stream2(p.szz + q, a.szz, b.szz);
                st_stream2(p.sxy + q, a.sxy, b.sxy);
                st_stream2(p.sxz + q, a.sxz, b.sxz);
                st_stream2(p.syz + q, a.syz, b.syz);
            }
            if (validA) {
                // boundary planes go straight into the neighbour slabs' halo planes (:951-963)
                const unsigned qp = (unsigned)((j - 1) * pitch + (i - 1));
                if (k == 1 && p.peer_lo[2]) st_stream2(p.peer_lo[2] + qp, a.szz, b.szz);          // sigmazz(:,:,1) -> left
                if (k == p.nzl && p.peer_hi[1]) {                                                 // -> right
                    st_stream2(p.peer_hi[1] + qp, a.sxz, b.sxz);
                    st_stream2(p.peer_hi[2] + qp, a.syz, b.syz);
                }
            }
            rn = rn1;
            rc.advance(SC);
        }
        rn.advance(SN);        // plane ke+1 of ring N has been consumed as "next" only

        // boundary item done: count it; the last one of a side publishes this slab's sigma planes to that neighbour
        const bool pub_lo = ss.pub_lo && kb == 1, pub_hi = ss.pub_hi && ke == p.nzl;
        if (pub_lo || pub_hi) {
            bar_sync(kConsBar, NC);
            if (tid == 0) {
                if (pub_lo) publish_side(ss.count, ss.n_boundary, ss.pub_lo, ss.pub_value);
                if (pub_hi) publish_side(ss.count + 1, ss.n_boundary, ss.pub_hi, ss.pub_value);
            }
        }
    }
}

// ----------------------------------------------------------------------------- velocity
// ring N (planes n and n+1): szz (plain box)
// ring C (plane n only):     sxx (halo box at (-2,0)), syy (halo, (0,0)), sxy (halo, (0,-1)),
//                            sxz (halo, (0,0)), syz (halo, (0,-1)), vx vy vz (plain boxes), x-shell memory variables
// maps: 0 sxx 1 syy 2 sxy 3 sxz 4 syz 5 szz 6 vx 7 vy 8 vz
template <typename T, bool KUNIT, int TX, int TY>
__global__ void __launch_bounds__(tile_threads(TX, TY) + 32, 1)
k_velocity3d_ws(const __grid_constant__ Params3DT<T> p, const __grid_constant__ TmaMaps tm, const __grid_constant__ Tile3D t,
                const __grid_constant__ SlabSync ss)
{
    using G = TileGeomT<T, TX, TY>;
    constexpr int W = G::W, HX = G::HX, ES = G::ES;
    constexpr int NC = tile_threads(TX, TY), NALL = NC + 32;
    constexpr int NBYTES = G::PLAIN_BYTES;                          // ring N stage: szz
    constexpr int CBYTES = 5 * G::HALO_BYTES + 3 * G::PLAIN_BYTES;  // ring C stage: 5 sigma (halo), vx vy vz
    constexpr uint32_t TX_N = G::PLAIN_BOX_BYTES;
    constexpr uint32_t TX_C = 5 * G::HALO_BOX_BYTES + 3 * G::PLAIN_BOX_BYTES;
    constexpr int PD = G::PLAIN_BYTES / ES, HD = G::HALO_BYTES / ES;
    const uint32_t XMB = (uint32_t)t.xm_bytes, XM_TX = (uint32_t)(TY * p.sxp * ES);
    const uint32_t CSTAGE = CBYTES + 3 * XMB;

    __shared__ double red[2 * (NC / 32)];
    __shared__ T Cx[6 * TX];                                   // x coefficients of the tile's columns
    __shared__ unsigned long long item_bar;
    __shared__ int item_slot[2];                               // the producer posts the CTA's work items here
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t sbase = (smem_u32(smem_dyn) + 127u) & ~127u;
    const unsigned char *gbase = smem_dyn + (sbase - smem_u32(smem_dyn));
    const uint32_t SC = (uint32_t)t.stages, SN = SC + 1;
    const uint32_t barN = sbase, barC = sbase + 64;
    const uint32_t ringN = sbase + kBarBytes, ringC = ringN + SN * NBYTES;
    const unsigned char *gN = gbase + kBarBytes, *gC = gN + (size_t)SN * NBYTES;

    const uint32_t barI = smem_u32(&item_bar);                      // "work item it & 1 has been posted"

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
            // post the item (or the end mark) to the consumers; slot ip & 1 was last read two items ago
            if (lead) { item_slot[ip & 1u] = item; mbar_arrive(barI); }
            ++ip;
            if (item >= t.nitems) break;
            const Item it = decode_item(item, t, p.nzl);
            const int x0 = it.tix * TX, y0 = it.tiy * TY, j0 = 1 + y0;
            const int kb = it.kb, ke = it.ke;
            const int np = ke - kb + 1;
            const bool tile_xpml = XMB != 0 && ((x0 + 1 <= p.xlo) || (x0 + TX >= p.xhi));
            auto issue_n = [&](uint32_t s, int kk) {
                const uint32_t bar = barN + 8 * s;
                mbar_expect_tx(bar, TX_N);
                tma_load_3d(ringN + s * NBYTES, &tm.m[5], x0, y0, kk, bar);
            };
            auto issue_c = [&](uint32_t s, int kk) {
                const uint32_t bar = barC + 8 * s, dst = ringC + s * CSTAGE;
                mbar_expect_tx(bar, TX_C + (tile_xpml ? 3 * XM_TX : 0u));
                if (tile_xpml) {
                    const long long row0 = ((long long)(kk - 1) * p.ny + (j0 - 1)) * p.sxp;
#pragma unroll
                    for (int f = 0; f < 3; f++) bulk_load(dst + CBYTES + f * XMB, p.mx[3 + f] + row0, XM_TX, bar);
                }
                tma_load_3d(dst + 0 * G::HALO_BYTES, &tm.m[0], x0 - HX, y0, kk, bar);
                tma_load_3d(dst + 1 * G::HALO_BYTES, &tm.m[1], x0, y0, kk, bar);
                tma_load_3d(dst + 2 * G::HALO_BYTES, &tm.m[2], x0, y0 - 1, kk, bar);
                tma_load_3d(dst + 3 * G::HALO_BYTES, &tm.m[3], x0, y0, kk, bar);
                tma_load_3d(dst + 4 * G::HALO_BYTES, &tm.m[4], x0, y0 - 1, kk, bar);
#pragma unroll
                for (int f = 0; f < 3; f++)
                    tma_load_3d(dst + 5 * G::HALO_BYTES + f * G::PLAIN_BYTES, &tm.m[6 + f], x0, y0, kk, bar);
            };
            if (lead) {
                // the item's last ring-N load is the upper neighbour's plane NZ_LOCAL+1 (sigmazz)
                if (ss.wait_hi && ke == p.nzl) poll_flag(ss.wait_hi, ss.wait_value, ss.timeout);
                uint32_t s = sn;
                for (int l = 0; l < min((int)SN, np + 1); l++) { issue_n(s, kb + l); if (++s == SN) s = 0; }
                s = sc;
                for (int l = 0; l < min((int)SC, np); l++) { issue_c(s, kb + l); if (++s == SC) s = 0; }
            }
            int next = item + (int)gridDim.x;
            if (lead) next = claim_item(t.queue, next);      // the answer is needed after the plane loop
            for (int n = 0; n < np; ++n) {
                bar_sync(kRelBar0 + (int)sc, NALL);
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
    constexpr bool PACK = sizeof(T) == 4;       // per-item thread map (map_lane_packed), see k_stress3d_ws
    const int tx_rm = tid % (TX / 2);
    const int ty_raw = tid / (TX / 2);
    const bool ok_rm = ty_raw < TY;
    const int ty_rm = min(ty_raw, TY - 1);
    const int pitch = p.pitch;
    const unsigned pl = (unsigned)p.plane;      // element offsets fit 32 bits (checked on the host)

    RingPos rn{0, 0}, rc{0, 0};
    for (uint32_t ip = 0;; ++ip) {
        mbar_wait(barI, ip & 1u);
        const int item = item_slot[ip & 1u];
        if (item >= t.nitems) break;
        const Item itm = decode_item(item, t, p.nzl);
        const int i0 = 1 + itm.tix * TX, j0 = 1 + itm.tiy * TY;
        const int kb = itm.kb, ke = itm.ke;
        const int np = ke - kb + 1;
        const bool tile_xpml = XMB != 0 && ((i0 <= p.xlo) || (i0 + TX - 1 >= p.xhi));
        // plane 0 (sigmaxz, sigmayz of the lower neighbour) is read straight from global memory below
        if (ss.wait_lo && kb == 1 && tid == 0) poll_flag(ss.wait_lo, ss.wait_value, ss.timeout);

        Lane ln{tx_rm, ty_rm, ok_rm};
        if constexpr (PACK) ln = map_lane_packed<TX, TY>(tid, i0, p.xlo, p.xhi, p.nx);
        const int tx = ln.tx, ty = ln.ty;
        const int oh = ty * W + 2 * tx;
        const int oc = ty * TX + 2 * tx;
        const int i = i0 + 2 * tx, j = j0 + ty;
        const bool row = ln.ok && (j <= p.ny);
        const bool validA = row && (i <= p.nx), validB = row && (i + 1 <= p.nx);
        const bool storeA = row && (i + 1 <= pitch);                // validA or a pad pair of the row
        unsigned q = (unsigned)kb * pl + (unsigned)((j - 1) * pitch + (i - 1));

        const bool in_xA = validA && ((i <= p.xlo) || (i >= p.xhi));
        const bool in_xB = validB && ((i + 1 <= p.xlo) || (i + 1 >= p.xhi));
        const bool in_y = validA && ((j <= p.ylo) || (j >= p.yhi));
        const bool ux = __any_sync(0xffffffffu, in_xA || in_xB), uy = __any_sync(0xffffffffu, in_y);   // warp-uniform
        const int jc = min(j, p.ny);
        if (tile_xpml) fill_cx<TX, NC>(p, Cx, i0, tid);
        bar_sync(kConsBar, NC);
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
        unsigned qxr = (unsigned)(((kb - 1) * p.ny + (j - 1)) * p.sxp);
        unsigned qy = in_y ? (unsigned)(((kb - 1) * p.sy + shell_index(j, p.ylo, p.yhi)) * pitch + (i - 1)) : 0u;
        const unsigned qx_step = (unsigned)(p.ny * p.sxp), qy_step = (unsigned)(p.sy * pitch);

        T sxz_mA = 0, sxz_mB = 0, syz_mA = 0, syz_mB = 0;                  // plane kb-1
        if (validA) {
            const auto t2 = ldg2(p.sxz + q - pl), u2 = ldg2(p.syz + q - pl);
            sxz_mA = t2.x; sxz_mB = t2.y; syz_mA = u2.x; syz_mB = u2.y;
        }
        double ekin = 0.0, epot = 0.0;

        mbar_wait(barN + 8 * rn.s, rn.par);
        for (int n = 0; n < np; ++n, q += pl, qxr += qx_step, qy += qy_step) {
            const int k = kb + n;
            const int kg = k + p.koff;
            const bool z_pml = (kg <= p.zlo) || (kg >= p.zhi);      // uniform
            const bool gen = uy || z_pml;                           // warp-uniform
            const bool in_zA = validA && z_pml, in_zB = validB && z_pml;
            unsigned qz = 0;
            T mvA[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, mvB[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            if (gen) {
                if (z_pml) qz = (unsigned)(((shell_index(kg, p.zlo, p.zhi) - p.zbase) * p.ny + (j - 1)) * pitch + (i - 1));
                load_memvars(p, 3, false, uy && in_y, in_zA, 0, qy, qz, mvA);
                load_memvars(p, 3, false, uy && in_y && validB, in_zB, 0, qy + 1, qz + 1, mvB);
            }
            RingPos rn1 = rn;
            rn1.advance(SN);
            mbar_wait(barN + 8 * rn1.s, rn1.par);
            mbar_wait(barC + 8 * rc.s, rc.par);

            const T *Tc = (const T *)(gC + (size_t)rc.s * CSTAGE);
            if (ux) {
                const T *Tm = (const T *)(gC + (size_t)rc.s * CSTAGE + CBYTES);
                const int xd = (int)(XMB / ES), r0 = ty * p.sxp;
                if (in_xA) { mvA[0] = Tm[r0 + sxA]; mvA[1] = Tm[xd + r0 + sxA]; mvA[2] = Tm[2 * xd + r0 + sxA]; }
                if (in_xB) { mvB[0] = Tm[r0 + sxB]; mvB[1] = Tm[xd + r0 + sxB]; mvB[2] = Tm[2 * xd + r0 + sxB]; }
            }
            const T *Txx = Tc, *Tyy = Tc + HD, *Txy = Tc + 2 * HD, *Txz = Tc + 3 * HD, *Tyz = Tc + 4 * HD;
            const T *Tp = Tc + 5 * HD;

            const auto sxx_c = lds2(Txx, oh + HX);                  // box starts HX cells before the tile
            const T sxx_imA = Txx[oh + HX - 1];
            const auto syy_c = lds2(Tyy, oh), syy_jp = lds2(Tyy, oh + W);
            const auto sxy_c = lds2(Txy, oh + W), sxy_jm = lds2(Txy, oh);
            const T sxy_ipB = Txy[oh + W + 2];
            const auto sxz_c = lds2(Txz, oh);
            const T sxz_ipB = Txz[oh + 2];
            const auto syz_c = lds2(Tyz, oh + W), syz_jm = lds2(Tyz, oh);
            const auto szz_c = lds2((const T *)(gN + (size_t)rn.s * NBYTES), oc);
            const auto szz_n = lds2((const T *)(gN + (size_t)rn1.s * NBYTES), oc);
            const auto vx = lds2(Tp, 0 * PD + oc), vy = lds2(Tp, 1 * PD + oc), vz = lds2(Tp, 2 * PD + oc);

            bar_arrive(kRelBar0 + (int)rc.s, NALL);                 // stages rn.s / rc.s are in registers

            VelValsT<T> a{vx.x, vy.x, vz.x}, b{vx.y, vy.y, vz.y};
            if (gen) {
                velocity_point<true, KUNIT, TX>(p, Cx, 2 * tx, jc, k, kg, do_vxA, do_vyA, do_vzA, edgeA, eboxA, srcA, ux, uy, z_pml, in_xA, in_y, in_zA,
                                                qxr + sxA, qy, qz, mvA,
                                                sxx_c.x, sxx_imA, syy_c.x, syy_jp.x, sxy_c.x, sxy_jm.x, sxy_c.y, sxz_c.x, sxz_c.y, sxz_mA,
                                                syz_c.x, syz_jm.x, syz_mA, szz_c.x, szz_n.x, a, ekin, epot);
                velocity_point<true, KUNIT, TX>(p, Cx, 2 * tx + 1, jc, k, kg, do_vxB, do_vyB, do_vzB, edgeB, eboxB, srcB, ux, uy, z_pml, in_xB, in_y && validB, in_zB,
                                                qxr + sxB, qy + 1, qz + 1, mvB,
                                                sxx_c.y, sxx_c.x, syy_c.y, syy_jp.y, sxy_c.y, sxy_jm.y, sxy_ipB, sxz_c.y, sxz_ipB, sxz_mB,
                                                syz_c.y, syz_jm.y, syz_mB, szz_c.y, szz_n.y, b, ekin, epot);
            } else if (ux) {
                velocity_point<true, KUNIT, TX>(p, Cx, 2 * tx, jc, k, kg, do_vxA, do_vyA, do_vzA, edgeA, eboxA, srcA, true, false, false, in_xA, false, false,
                                                qxr + sxA, 0, 0, mvA,
                                                sxx_c.x, sxx_imA, syy_c.x, syy_jp.x, sxy_c.x, sxy_jm.x, sxy_c.y, sxz_c.x, sxz_c.y, sxz_mA,
                                                syz_c.x, syz_jm.x, syz_mA, szz_c.x, szz_n.x, a, ekin, epot);
                velocity_point<true, KUNIT, TX>(p, Cx, 2 * tx + 1, jc, k, kg, do_vxB, do_vyB, do_vzB, edgeB, eboxB, srcB, true, false, false, in_xB, false, false,
                                                qxr + sxB, 0, 0, mvB,
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

            // pad lanes (beyond NX, inside the row pitch) hold and keep zero: they are outside every nest and their
            // loaded values are the TMA's zero fill; they are stored so that whole rows / sectors are written
            if (!validA) { a.vx = 0; a.vy = 0; a.vz = 0; }
            if (!validB) { b.vx = 0; b.vy = 0; b.vz = 0; }
            if (storeA) {
                st_stream2(p.vx + q, a.vx, b.vx);
                st_stream2(p.vy + q, a.vy, b.vy);
                st_stream2(p.vz + q, a.vz, b.vz);
            }
            if (validA) {
                // boundary planes go straight into the neighbour slabs' halo planes (:811-823)
                const unsigned qp = (unsigned)((j - 1) * pitch + (i - 1));
                if (k == 1 && p.peer_lo[0]) {                                                   // -> left
                    st_stream2(p.peer_lo[0] + qp, a.vx, b.vx);
                    st_stream2(p.peer_lo[1] + qp, a.vy, b.vy);
                }
                if (k == p.nzl && p.peer_hi[0]) st_stream2(p.peer_hi[0] + qp, a.vz, b.vz);      // -> right
            }
            rn = rn1;
            rc.advance(SC);
        }
        rn.advance(SN);

        cons_sum2<NC>(ekin, epot, red, tid);       // (its barriers also order this item's peer stores before the count below)
        if (tid == 0) {
            p.partials[item] = ekin;
            p.partials[p.nblocks + item] = epot;
            if (ss.pub_lo && kb == 1) publish_side(ss.count, ss.n_boundary, ss.pub_lo, ss.pub_value);
            if (ss.pub_hi && ke == p.nzl) publish_side(ss.count + 1, ss.n_boundary, ss.pub_hi, ss.pub_value);
        }
    }
}

// ---- launch dispatch ---------------------------------------------------------------

template <typename T, int TX, int TY>
static size_t ws_smem_need(bool stress, int stages, int xm_bytes)
{
    using G = TileGeomT<T, TX, TY>;
    const size_t c = stress ? G::HALO_BYTES + 6 * G::PLAIN_BYTES : 5 * G::HALO_BYTES + 3 * G::PLAIN_BYTES;
    const size_t n = stress ? 2 * G::HALO_BYTES : G::PLAIN_BYTES;
    return kBarBytes + 128 + (c + 3 * (size_t)xm_bytes) * (size_t)stages + n * (size_t)(stages + 1);
}

template <typename T, bool KUNIT, int TX, int TY>
static cudaError_t ws_launch_tile(const Params3DT<T> &p, const TmaMaps &tm, const Tile3D &t, const SlabSync &ss, cudaStream_t s,
                                  bool stress, int *occ)
{
    const size_t smem = ws_smem_need<T, TX, TY>(stress, t.stages, t.xm_bytes);
    constexpr int NT = tile_threads(TX, TY) + 32;     // consumer warps + the producer warp
    const void *fn = stress ? (const void *)k_stress3d_ws<T, KUNIT, TX, TY> : (const void *)k_velocity3d_ws<T, KUNIT, TX, TY>;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (occ) {
        if (stress) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k_stress3d_ws<T, KUNIT, TX, TY>, NT, smem);
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k_velocity3d_ws<T, KUNIT, TX, TY>, NT, smem);
    }
    const dim3 grid(stress ? t.grid_stress : t.grid_velocity);
    if (stress) k_stress3d_ws<T, KUNIT, TX, TY><<<grid, NT, smem, s>>>(p, tm, t, ss);
    else        k_velocity3d_ws<T, KUNIT, TX, TY><<<grid, NT, smem, s>>>(p, tm, t, ss);
    return cudaGetLastError();
}

// Tiles (TX x TY points, one consumer thread per pair of x-adjacent points + 32 producer threads).  With the
// producer warp a 128 x 8 tile would be 17 warps (five on one SM sub-partition: 96 registers), so wide grids
// take 128 x 7 (15 warps) or 104 x 8 (14 warps).  The single-precision build has 64 x 8, 104 x 8 and 128 x 7.
template <bool KUNIT>
static cudaError_t ws_dispatch(const Params3D &p, const TmaMaps &tm, const Tile3D &t, const SlabSync &ss, cudaStream_t s,
                               bool stress, int *occ)
{
    switch (t.tx * 100 + t.ty) {
    case 6404:  return ws_launch_tile<double, KUNIT, 64, 4>(p, tm, t, ss, s, stress, occ);      // 128 + 32 threads (tests)
    case 6408:  return ws_launch_tile<double, KUNIT, 64, 8>(p, tm, t, ss, s, stress, occ);      // 256 + 32
    case 10407: return ws_launch_tile<double, KUNIT, 104, 7>(p, tm, t, ss, s, stress, occ);     // 384 + 32
    case 10408: return ws_launch_tile<double, KUNIT, 104, 8>(p, tm, t, ss, s, stress, occ);     // 416 + 32
    case 12806: return ws_launch_tile<double, KUNIT, 128, 6>(p, tm, t, ss, s, stress, occ);     // 384 + 32
    case 12807: return ws_launch_tile<double, KUNIT, 128, 7>(p, tm, t, ss, s, stress, occ);     // 448 + 32
    default: return cudaErrorInvalidValue;
    }
}
template <bool KUNIT>
static cudaError_t ws_dispatch(const Params3DF &p, const TmaMaps &tm, const Tile3D &t, const SlabSync &ss, cudaStream_t s,
                               bool stress, int *occ)
{
    switch (t.tx * 100 + t.ty) {
    case 6408:  return ws_launch_tile<float, KUNIT, 64, 8>(p, tm, t, ss, s, stress, occ);
    case 10408: return ws_launch_tile<float, KUNIT, 104, 8>(p, tm, t, ss, s, stress, occ);
    case 12807: return ws_launch_tile<float, KUNIT, 128, 7>(p, tm, t, ss, s, stress, occ);
    default: return cudaErrorInvalidValue;
    }
}

bool ws_tile_supported(int tx, int ty, bool f32)
{
    switch (tx * 100 + ty) {
    case 6408: case 10408: case 12807: return true;
    case 6404: case 10407: case 12806: return !f32;
    default: return false;
    }
}

cudaError_t ws_occupancy(const Params3D &p, const Tile3D &t, bool stress, int *occ, bool f32)
{
    TmaMaps dummy{};
    SlabSync none{};
    if (f32) {
        Params3DF pf{};
        return p.kunit ? ws_dispatch<true>(pf, dummy, t, none, nullptr, stress, occ) : ws_dispatch<false>(pf, dummy, t, none, nullptr, stress, occ);
    }
    return p.kunit ? ws_dispatch<true>(p, dummy, t, none, nullptr, stress, occ) : ws_dispatch<false>(p, dummy, t, none, nullptr, stress, occ);
}

cudaError_t launch_stress3d_ws(const Params3D &p, const TmaMaps &tm, const Tile3D &t, const SlabSync &ss, cudaStream_t s)
{
    return p.kunit ? ws_dispatch<true>(p, tm, t, ss, s, true, nullptr) : ws_dispatch<false>(p, tm, t, ss, s, true, nullptr);
}
cudaError_t launch_velocity3d_ws(const Params3D &p, const TmaMaps &tm, const Tile3D &t, const SlabSync &ss, cudaStream_t s)
{
    return p.kunit ? ws_dispatch<true>(p, tm, t, ss, s, false, nullptr) : ws_dispatch<false>(p, tm, t, ss, s, false, nullptr);
}
cudaError_t launch_stress3d_ws(const Params3DF &p, const TmaMaps &tm, const Tile3D &t, const SlabSync &ss, cudaStream_t s)
{
    return p.kunit ? ws_dispatch<true>(p, tm, t, ss, s, true, nullptr) : ws_dispatch<false>(p, tm, t, ss, s, true, nullptr);
}
cudaError_t launch_velocity3d_ws(const Params3DF &p, const TmaMaps &tm, const Tile3D &t, const SlabSync &ss, cudaStream_t s)
{
    return p.kunit ? ws_dispatch<true>(p, tm, t, ss, s, false, nullptr) : ws_dispatch<false>(p, tm, t, ss, s, false, nullptr);
}

}  // namespace cpml

// Test hooks, not part of the ABI of include/cpml_b200.h (tests/test_work_items.py, no GPU needed): the work
// decomposition of the persistent kernels exactly as the device code computes it -- the same functions, compiled for
// the host.  tile6 = ntx, nty, kchunk, nzc, fine_from, split; out4 = tix, tiy, kb, ke.
extern "C" int32_t cpml_debug_work_item(const int32_t *tile6, int32_t nzl, int32_t item, int32_t *out4)
{
    if (!tile6 || !out4) return -1;
    cpml::Tile3D t{};
    t.ntx = tile6[0]; t.nty = tile6[1]; t.kchunk = tile6[2]; t.nzc = tile6[3]; t.fine_from = tile6[4]; t.split = tile6[5];
    const cpml::Item it = cpml::decode_item(item, t, nzl);
    out4[0] = it.tix; out4[1] = it.tiy; out4[2] = it.kb; out4[3] = it.ke;
    return 0;
}

// The packed thread map of the single-precision kernels for tile tx x ty: out3 = pair of the row, row, active.
extern "C" int32_t cpml_debug_lane(int32_t tx, int32_t ty, int32_t tid, int32_t i0, int32_t xlo, int32_t xhi, int32_t nx, int32_t *out3)
{
    if (!out3) return -1;
    cpml::Lane l{};
    switch (tx * 100 + ty) {
    case 6408:  l = cpml::map_lane_packed<64, 8>(tid, i0, xlo, xhi, nx); break;
    case 10408: l = cpml::map_lane_packed<104, 8>(tid, i0, xlo, xhi, nx); break;
    case 12807: l = cpml::map_lane_packed<128, 7>(tid, i0, xlo, xhi, nx); break;
    default: return -1;
    }
    out3[0] = l.tx; out3[1] = l.ty; out3[2] = l.ok ? 1 : 0;
    return 0;
}
