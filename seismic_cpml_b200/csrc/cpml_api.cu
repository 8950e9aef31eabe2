// C ABI of libcpml_b200 (include/cpml_b200.h): handle life cycle, setters, the step
// loop and the getters.  Host logic only; the device code is in kernels_{2d,3d}.cu.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "cpml_internal.h"

using namespace cpml;

namespace {
thread_local std::string g_create_error;

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

int env_int(const char *name, int dflt)
{
    const char *s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}
}  // namespace

struct cpml_handle {
    cpml_config cfg{};
    std::string err;
    cudaStream_t stream = nullptr;
    int device = 0;
    int sm_count = 148;

    // geometry of this slab
    int nzl = 1, koff = 0, pitch = 0, ksrc_global = 0;
    long long plane = 0;       // doubles per z plane (3-D) / whole padded field (2-D)
    size_t field_doubles = 0;  // allocation size of one field
    long long origin = 0;      // offset of element (1,1,0) / (1,1) inside the allocation
    int nfields = 0;
    double *arena = nullptr;   // ONE allocation: nfields * field_doubles + the slab flags (so that a
                               // neighbour process maps everything with a single IPC handle)
    size_t arena_doubles = 0;
    double *field_alloc[15] = {};
    double *f0[15] = {};       // element (1,1,0) / (1,1)
    size_t flags_offset = 0;   // doubles from the arena start to the slab flags

    // viscoelastic (rheology == 1): six (N_SLS = 2) strain memory variables, constants of 3D-visco :458-477
    bool visco = false;        // 3-D viscoelastic
    bool visco2d = false;      // 2-D viscoelastic (N_SLS = 3): fields 5..13 are e1(1..3), e11(1..3), e13(1..3)
    double2 *e0[6] = {};       // element (1,1,0) of e1, e11, e22, e12, e13, e23
    bool have_attenuation = false;
    double tau[4][3] = {};     // tau_epsilon_nu1, tau_sigma_nu1, tau_epsilon_nu2, tau_sigma_nu2
    double *d_sisp = nullptr;  // 2-D viscoelastic: sispressure
    double *d_sisvz = nullptr; // 3-D: Vz seismograms (extension, quirk B7)
    dim3 vgrid;
    int vkchunk = 1, vtx = 32, vty = 8;
    bool ws2 = false;          // 2-D isotropic: TMA-staged y-marching kernels (kernels_2d_ws.cu, the default)
    Tile2D tile2{};
    TmaMaps maps2_stress{}, maps2_velocity{};
    bool v_ws = false;         // TMA-staged velocity kernel with a producer warp (kernels_3d_visco_ws.cu, the default)
    Tile3D vtile{};
    TmaMaps vmaps{};

    // TMA path (3-D): descriptors of the two kernels' plane tiles, work decomposition
    bool use_tma = false;          // TMA-staged kernels (either family)
    bool f32 = false;              // cfg.precision == 1: single-precision wavefields (3-D isotropic, one GPU)
    float *dprof_f[3][6] = {};     // single-precision copies of the profiles
    double *d_scratch = nullptr;   // one padded plane in double: single-precision planes are converted here for the getters
    bool use_ws = false;           // ... with a producer warp and in-kernel slab ordering (kernels_3d_ws.cu, the default)
    unsigned int *d_bcount = nullptr;   // [0..1] boundary-item counters of the in-kernel slab ordering, [2..5] two work queues
    TmaMaps maps_stress{}, maps_velocity{};
    Tile3D tile{}, tile_stress{};        // velocity kernel (also: energy partial slots) / stress kernel

    // slab neighbours reached by direct peer stores (cpml_p2p_*): side 0 = rank-1, 1 = rank+1
    bool peer_on[2] = {false, false};
    double *peer_arena[2] = {nullptr, nullptr};
    bool peer_ipc[2] = {false, false};
    unsigned long long *flags = nullptr;     // [0] v from lo, [1] v from hi, [2] sigma from lo, [3] sigma from hi
    unsigned int *d_timeout = nullptr;
    unsigned long long epoch = 0;            // run number (cpml_reset increments it): flag values are epoch << 32 | it, so a
                                             // neighbour's signal survives this slab's reset and stale values never satisfy a wait

    // profiles
    bool have_prof[3] = {false, false, false};
    std::vector<double> hprof[3][6];
    double *dprof[3][8] = {};  // device copies, 0-based; [6], [7] = RN(1/K), RN(1/K_half) for div_exact
    Shell shell[3] = {};
    int nz_own[3][2] = {};     // per axis: count of a != 0 (integer, half) -- algorithmic bytes

    // memory variables
    bool finalized = false;
    int sxp = 0, sy = 0, zbase = 0, sz_local = 0;
    double *mx[6] = {}, *my[6] = {}, *mz[6] = {};
    size_t mx_doubles = 0, my_doubles = 0, mz_doubles = 0;

    // 2-D material
    double *mat[3] = {};       // lambda, mu, rho allocations (same layout as fields)
    bool have_material = false;
    bool rho_exact = true;     // no density / interpolated density with an all-ones significand

    // source / receivers / traces
    double *d_src_x = nullptr, *d_src_y = nullptr;      // one allocation [2][nstep] (d_src_y = d_src_x + nstep)
    double *d_step_out = nullptr;  // [nstep][4] kinetic, potential, sisvx(it,1), sisvy(it,1): what cpml_fetch_step pulls
    bool have_source = false;
    int *d_ix_rec = nullptr, *d_iy_rec = nullptr;
    bool have_receivers = false;
    double *d_sisvx = nullptr, *d_sisvy = nullptr, *d_ek = nullptr, *d_ep = nullptr;
    double *d_partials = nullptr;
    unsigned long long *d_maxbits = nullptr;
    double *pin_src = nullptr;     // pinned host staging: [2][nstep] per-step source increments
    double *pin_out = nullptr;     // pinned host staging: [nstep][4] kinetic, potential, sisvx(it,1), sisvy(it,1)

    // asynchronous snapshot planes (cpml_snapshot_begin / _end): device-side copy, then D2H into pinned memory on a side stream
    cudaStream_t snap_stream = nullptr;
    cudaEvent_t snap_ready[2] = {nullptr, nullptr}, snap_done[2] = {nullptr, nullptr};
    double *snap_dev[2] = {nullptr, nullptr}, *snap_pin[2] = {nullptr, nullptr};
    bool snap_pending[2] = {false, false};

    // launch geometry: 3-D kernels run once per region (interior box + PML shell boxes)
    std::vector<Box3D> regions;
    dim3 grid, block;          // 2-D kernels
    int nblocks = 0;           // energy partial slots

    // kernel timing
    bool timing = false;
    // fixed pool of event pairs created by cpml_enable_kernel_timing (nothing is created inside a
    // timed loop); slot q of the ring: ev[2q] before / ev[2q+1] after launch number ev_used+q
    std::vector<cudaEvent_t> ev;
    std::vector<int> ev_kind;      // per slot: 0 stress, 1 velocity, -1 free
    size_t ev_head = 0, ev_count = 0;
    double ms_stress = 0, ms_velocity = 0;
    long long n_launches = 0;
};

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e__);                 \
            return CPML_ECUDA;                                                            \
        }                                                                                 \
    } while (0)

#define FAIL(code, msg)            \
    do {                           \
        h->err = (msg);            \
        return (code);             \
    } while (0)

// Markstein's correctly-rounded division needs a divisor whose significand is not all ones.
static bool all_ones_significand(double c)
{
    unsigned long long b;
    memcpy(&b, &c, sizeof(b));
    return (b & 0x000fffffffffffffULL) == 0x000fffffffffffffULL;
}

static Shell find_shell(const std::vector<double> *p, int n)
{
    // indices (1-based) where any coefficient is non-trivial: a != 0 or K != 1
    std::vector<char> nt(n + 2, 0);
    bool any = false;
    for (int i = 1; i <= n; i++) {
        nt[i] = (p[0][i - 1] != 0.0) || (p[3][i - 1] != 0.0) || (p[2][i - 1] != 1.0) || (p[5][i - 1] != 1.0);
        any = any || nt[i];
    }
    Shell s{0, n + 1, n};
    if (!any) return s;
    // the longest run of trivial indices separates the low shell from the high shell
    int best_len = -1, best_start = 1, run_start = -1;
    for (int i = 1; i <= n + 1; i++) {
        const bool triv = (i <= n) && !nt[i];
        if (triv && run_start < 0) run_start = i;
        if (!triv && run_start >= 0) {
            if (i - run_start > best_len) { best_len = i - run_start; best_start = run_start; }
            run_start = -1;
        }
    }
    if (best_len <= 0) { s.lo = n; s.hi = n + 1; return s; }   // no trivial index at all
    s.lo = best_start - 1;
    s.hi = best_start + best_len;
    return s;
}

extern "C" int32_t cpml_abi_version(void) { return CPML_B200_ABI_VERSION; }

extern "C" const char *cpml_last_error(const cpml_handle *h)
{
    return h ? h->err.c_str() : g_create_error.c_str();
}

static int32_t create_impl(cpml_handle *h)
{
    const cpml_config &c = h->cfg;
    if (c.ndim != 2 && c.ndim != 3) FAIL(CPML_EINVAL, "ndim must be 2 or 3");
    if (c.rheology != 0 && c.rheology != 1) FAIL(CPML_EINVAL, "rheology must be 0 (elastic) or 1 (viscoelastic)");
    h->visco = c.rheology == 1 && c.ndim == 3;
    h->visco2d = c.rheology == 1 && c.ndim == 2;
    if (h->visco && c.order != 4) FAIL(CPML_EINVAL, "the 3-D viscoelastic solver is fourth order (order = 4)");
    if (c.ndim == 3 && !h->visco && c.order != 2) FAIL(CPML_EINVAL, "3-D isotropic solver is second order (order must be 2)");
    if (c.precision != 0 && c.precision != 1) FAIL(CPML_EINVAL, "precision must be 0 (double) or 1 (single)");
    h->f32 = c.precision == 1;
    if (h->f32 && (c.ndim != 3 || c.rheology != 0)) FAIL(CPML_EINVAL, "single precision is implemented for the 3-D isotropic solver");
    if (h->f32 && c.nslabs != 1) FAIL(CPML_ETOPOLOGY, "single precision runs on one GPU (nslabs = 1)");
    if (c.emulate_nproc < 0) FAIL(CPML_EINVAL, "emulate_nproc must be >= 0");
    if (c.emulate_nproc > 1) {
        if (!h->visco) FAIL(CPML_EINVAL, "emulate_nproc applies to the viscoelastic solver only (the second-order exchange is complete)");
        if (c.emulate_nproc % 2 != 0) FAIL(CPML_ETOPOLOGY, "nb_procs must be even (3D-visco :522-523)");
        if (c.nz % c.emulate_nproc != 0 || c.nz / c.emulate_nproc < std::max(2, c.npoints_pml))
            FAIL(CPML_ETOPOLOGY, "emulate_nproc must divide NZ into slabs of at least NPOINTS_PML planes (3D-visco :525-528)");
    }
    if (c.ndim == 2 && c.order != 2 && c.order != 4) FAIL(CPML_EINVAL, "order must be 2 or 4");
    if (c.nx < 4 || c.ny < 4 || (c.ndim == 3 && c.nz < 4)) FAIL(CPML_EINVAL, "grid too small");
    if (c.nstep < 1 || c.nrec < 0 || c.npoints_pml < 0) FAIL(CPML_EINVAL, "bad nstep / nrec / npoints_pml");
    if (c.isource < 1 || c.isource > c.nx || c.jsource < 1 || c.jsource > c.ny) FAIL(CPML_EINVAL, "source outside the grid");
    if (!(c.deltax > 0) || !(c.deltay > 0) || !(c.deltat > 0)) FAIL(CPML_EINVAL, "DELTAX, DELTAY, DELTAT must be positive");

    if (c.ndim == 3) {
        if (!(c.deltaz > 0)) FAIL(CPML_EINVAL, "DELTAZ must be positive");
        if (!(c.rho > 0) || !(c.mu > 0)) FAIL(CPML_EINVAL, "rho and mu must be positive");
        // topology checks of 3D-iso :381-394 (evenness only matters for the default
        // cut plane, so it is required only when ksource is left to NZ/2)
        if (c.nslabs < 1 || c.slab_rank < 0 || c.slab_rank >= c.nslabs) FAIL(CPML_ETOPOLOGY, "bad nslabs / slab_rank");
        if (c.nz % c.nslabs != 0) FAIL(CPML_ETOPOLOGY, "NZ must be a multiple of nb_procs");
        h->nzl = c.nz / c.nslabs;
        if (h->nzl < c.npoints_pml) FAIL(CPML_ETOPOLOGY, "NZ_LOCAL must be greater than NPOINTS_PML");
        if (h->visco && h->nzl < 2) FAIL(CPML_ETOPOLOGY, "viscoelastic slabs need at least two planes");
        if (c.ksource == 0 && c.nslabs > 1 && c.nslabs % 2 != 0) FAIL(CPML_ETOPOLOGY, "nb_procs must be even");
        h->ksrc_global = c.ksource == 0 ? c.nz / 2 : c.ksource;
        if (h->ksrc_global < 1 || h->ksrc_global > c.nz) FAIL(CPML_EINVAL, "ksource outside the grid");
        h->koff = c.slab_rank * h->nzl;
    } else {
        if (c.nslabs > 1) FAIL(CPML_ETOPOLOGY, "2-D solvers are not decomposed");
        h->nzl = 1;
        h->koff = 0;
    }
    if (c.ndim == 2 && (all_ones_significand(24.0 * c.deltax) || all_ones_significand(c.deltax) ||
                        all_ones_significand(24.0 * c.deltay) || all_ones_significand(c.deltay)))
        FAIL(CPML_EINVAL, "grid spacing with an all-ones significand is not supported");
    if (c.cp > 0) {   // Courant check, 3D-iso :712-717 / 2D-2nd :513-516
        const double cn = cpml_host_courant(c.cp, c.deltat, c.deltax, c.deltay, c.ndim == 3 ? c.deltaz : 0.0);
        if (cn > 1.0) FAIL(CPML_ECFL, "time step is too large, simulation will be unstable");
    }

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        FAIL(CPML_ECUDA, "no CUDA device: libcpml_b200 has no CPU fallback");
    if (c.device >= 0) { CK(cudaSetDevice(c.device)); }
    CK(cudaGetDevice(&h->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, h->device));
    h->sm_count = prop.multiProcessorCount;

    // ---- field layout
    int ne = 0;           // double2 arrays (viscoelastic strain memory variables)
    if (h->visco) {
        // (0:NX+1, 0:NY+1, -1:NZ_LOCAL+2) of 3D-visco :301, ghost ring widened to two cells
        const int xo = 16, gy = 2;
        h->pitch = round_up(xo + c.nx + 2, 16);
        h->plane = (long long)h->pitch * (c.ny + 2 * gy);
        h->origin = h->plane + (long long)gy * h->pitch + xo;      // element (1,1,0); plane k = -1 comes first
        h->field_doubles = (size_t)h->plane * (h->nzl + 4) + 16;
        h->nfields = 15;
        ne = 6;
    } else if (c.ndim == 3) {
        // Row pitch: NX rounded up to 4 doubles (one 32-byte sector), NOT to a 128-byte line.  The TMA kernels do not
        // care where a row starts, and with NX = 101 a 112-double pitch left 11 pad doubles per row that the 128-byte
        // L2 promotion of the tensor loads fetched anyway and that turned the last sector of every stored row into a
        // partial write (ncu: 1.09x the algorithmic DRAM bytes, profiles/r02_a_ncu_cfg3_ws.txt); with 104 the planes
        // are dense -- the kernels also write the (zero) pad pairs -- and nothing but whole sectors moves.
        // CPML_PITCH_ALIGN=16 restores the old layout for A/B runs.
        int align = env_int("CPML_PITCH_ALIGN", 4);
        if (align != 2 && align != 4 && align != 8 && align != 16) align = 4;
        h->pitch = round_up(c.nx, align);
        h->plane = (long long)h->pitch * c.ny;
        h->origin = 16;   // leading pad so that (i-1) at the first element stays inside
        h->field_doubles = (size_t)h->plane * (h->nzl + 2) + 32 + (size_t)h->pitch;
        h->nfields = 9;
    } else {
        const int xo = 16, gy = 2;
        h->pitch = round_up(xo + c.nx + 2, 16);
        h->plane = (long long)h->pitch * (c.ny + 2 * gy);
        h->origin = (long long)gy * h->pitch + xo;
        h->field_doubles = (size_t)h->plane + 16;
        h->nfields = h->visco2d ? 14 : 5;
    }
    h->field_doubles = (h->field_doubles + 15) / 16 * 16;      // every field starts on a 128-byte line
    h->flags_offset = (size_t)(h->nfields + 2 * ne) * h->field_doubles;
    h->arena_doubles = h->flags_offset + 16;
    CK(cudaMalloc(&h->arena, h->arena_doubles * sizeof(double)));
    for (int f = 0; f < h->nfields; f++) {
        h->field_alloc[f] = h->arena + (size_t)f * h->field_doubles;
        h->f0[f] = h->field_alloc[f] + h->origin;
    }
    for (int m = 0; m < ne; m++)
        h->e0[m] = (double2 *)(h->arena + (size_t)(h->nfields + 2 * m) * h->field_doubles) + h->origin;
    h->flags = (unsigned long long *)(h->arena + h->flags_offset);
    CK(cudaMemset(h->flags, 0, 16 * sizeof(double)));          // zeroed once: cpml_reset leaves the flag words alone
    CK(cudaMalloc(&h->d_timeout, sizeof(unsigned int)));
    CK(cudaMemset(h->d_timeout, 0, sizeof(unsigned int)));
    CK(cudaMalloc(&h->d_bcount, 8 * sizeof(unsigned int)));
    CK(cudaMemset(h->d_bcount, 0, 8 * sizeof(unsigned int)));
    if (c.ndim == 2)
        for (int m = 0; m < 3; m++) CK(cudaMalloc(&h->mat[m], h->field_doubles * sizeof(double)));

    const size_t nt = (size_t)c.nstep;
    // (x and y series in one allocation, like the pinned staging: cpml_set_source_step is ONE strided copy)
    CK(cudaMalloc(&h->d_src_x, 2 * nt * sizeof(double)));
    h->d_src_y = h->d_src_x + nt;
    CK(cudaMemset(h->d_src_x, 0, 2 * nt * sizeof(double)));  // steps beyond a short series inject nothing
    CK(cudaMalloc(&h->d_step_out, 4 * nt * sizeof(double)));
    CK(cudaMemset(h->d_step_out, 0, 4 * nt * sizeof(double)));
    CK(cudaMalloc(&h->d_ek, nt * sizeof(double)));
    CK(cudaMalloc(&h->d_ep, nt * sizeof(double)));
    const size_t ns = std::max<size_t>(1, nt * (size_t)c.nrec);
    CK(cudaMalloc(&h->d_sisvx, ns * sizeof(double)));
    CK(cudaMalloc(&h->d_sisvy, ns * sizeof(double)));
    if (h->visco2d) CK(cudaMalloc(&h->d_sisp, ns * sizeof(double)));
    if (c.ndim == 3) CK(cudaMalloc(&h->d_sisvz, ns * sizeof(double)));
    CK(cudaMalloc(&h->d_ix_rec, std::max(1, c.nrec) * sizeof(int)));
    CK(cudaMalloc(&h->d_iy_rec, std::max(1, c.nrec) * sizeof(int)));
    CK(cudaMalloc(&h->d_maxbits, sizeof(unsigned long long)));
    CK(cudaMallocHost(&h->pin_src, 2 * nt * sizeof(double)));
    CK(cudaMallocHost(&h->pin_out, 4 * nt * sizeof(double)));
    memset(h->pin_out, 0, 4 * nt * sizeof(double));

    // ---- launch geometry of the 2-D kernels (the 3-D regions need the shells: finalize())
    if (c.ndim == 2) {
        h->block = dim3(32, 8, 1);
        h->grid = dim3((c.nx + 31) / 32, (c.ny + 7) / 8, 1);
        h->nblocks = h->grid.x * h->grid.y;
        CK(cudaMalloc(&h->d_partials, 2 * (size_t)h->nblocks * sizeof(double)));
    }
    return cpml_reset(h);
}

// Adds the box [i0,i1] x [j0,j1] x [k0,k1] (k local) to the region list, if non-empty.
static void add_region(cpml_handle *h, int i0, int i1, int j0, int j1, int k0, int k1, bool pml, int tx, int ty)
{
    if (i0 > i1 || j0 > j1 || k0 > k1) return;
    Box3D b{};
    b.i0 = i0; b.i1 = i1; b.j0 = j0; b.j1 = j1; b.k0 = k0; b.k1 = k1;
    b.ia = ((i0 - 1) / 16) * 16 + 1;        // 128-byte line aligned start of the thread grid
    b.pml = pml ? 1 : 0;
    b.tx = tx; b.ty = ty;
    b.gx = (i1 - b.ia + 1 + tx - 1) / tx;
    b.gy = (j1 - j0 + 1 + ty - 1) / ty;
    // z chunks: enough blocks for ~4 waves of resident CTAs (the kernels are latency-bound
    // with fewer), but chunks of >= 8 planes so that the register-carried z reuse pays for
    // the extra plane fetched at every chunk start.  CPML_ZCHUNKS overrides.
    const int nk = k1 - k0 + 1;
    int zch = env_int("CPML_ZCHUNKS", 0);
    if (zch <= 0) {
        const int resident = h->sm_count * std::max(1, 2048 / (tx * ty));
        zch = (4 * resident + b.gx * b.gy - 1) / (b.gx * b.gy);
        zch = std::min(zch, std::max(1, nk / 8));
    }
    zch = std::max(1, std::min(zch, nk));
    b.kchunk = (nk + zch - 1) / zch;
    b.gz = (nk + b.kchunk - 1) / b.kchunk;
    b.pbase = h->nblocks;
    h->nblocks += b.gx * b.gy * b.gz;
    h->regions.push_back(b);
}

// Splits the slab into the PML-free interior box and up to six shell boxes; every grid
// point falls in exactly one region.
static void build_regions(cpml_handle *h)
{
    const cpml_config &c = h->cfg;
    const Shell &sx = h->shell[0], &sy = h->shell[1], &sz = h->shell[2];
    // thread tile: half-warp rows waste fewer lanes on narrow grids (NX = 101: 112 of 128
    // lanes instead of 101 of 128); measured in profiles/r01_v2_tile_sweep.txt
    int tx = env_int("CPML_TX", c.nx <= 160 ? 16 : 32), ty = env_int("CPML_TY", 8);
    if (!tile_supported(tx, ty)) { tx = 32; ty = 8; }
    int ptx = env_int("CPML_PML_TX", 16), pty = env_int("CPML_PML_TY", 8);   // thin x shells: half-warp rows
    if (!tile_supported(ptx, pty)) { ptx = 16; pty = 8; }
    h->regions.clear();
    h->nblocks = 0;
    // local k range outside the z shells
    const int kz0 = std::max(1, sz.lo + 1 - h->koff), kz1 = std::min(h->nzl, sz.hi - 1 - h->koff);
    const int jy0 = sy.lo + 1, jy1 = sy.hi - 1;
    const int ix0 = sx.lo + 1, ix1 = sx.hi - 1;
    // Default: ONE launch over the whole slab.  Splitting into an interior launch plus six
    // shell launches (CPML_REGIONS=1) was measured slower on B200 (profiles/r01_v2_region_sweep.txt:
    // 12.1 vs 16.5 Gpts/s on 101x641x640): the <PML=true> kernel is as fast per byte as the lean
    // one once its loads are hoisted, and seven serial launches add tails.
    if (!env_int("CPML_REGIONS", 0)) {
        const bool any_shell = sx.size() + sy.size() + sz.size() > 0;
        add_region(h, 1, c.nx, 1, c.ny, 1, h->nzl, any_shell, tx, ty);
        return;
    }
    add_region(h, ix0, ix1, jy0, jy1, kz0, kz1, false, tx, ty);                    // interior
    add_region(h, 1, c.nx, 1, c.ny, 1, std::min(h->nzl, kz0 - 1), true, tx, ty);   // z-
    add_region(h, 1, c.nx, 1, c.ny, std::max(1, kz1 + 1), h->nzl, true, tx, ty);   // z+
    if (kz0 <= kz1) {
        add_region(h, 1, c.nx, 1, std::min(c.ny, jy0 - 1), kz0, kz1, true, tx, ty);    // y-
        add_region(h, 1, c.nx, std::max(1, jy1 + 1), c.ny, kz0, kz1, true, tx, ty);    // y+
        if (jy0 <= jy1) {
            add_region(h, 1, std::min(c.nx, ix0 - 1), jy0, jy1, kz0, kz1, true, ptx, pty);   // x-
            add_region(h, std::max(1, ix1 + 1), c.nx, jy0, jy1, kz0, kz1, true, ptx, pty);   // x+
        }
    }
}

extern "C" int32_t cpml_create(const cpml_config *cfg, cpml_handle **out)
{
    if (out) *out = nullptr;
    if (!cfg || !out) { g_create_error = "null argument"; return CPML_EINVAL; }
    cpml_handle *h = new (std::nothrow) cpml_handle();
    if (!h) { g_create_error = "out of host memory"; return CPML_ENOMEM; }
    h->cfg = *cfg;
    const int32_t rc = create_impl(h);
    if (rc != CPML_OK) {
        g_create_error = h->err;
        cpml_destroy(h);
        return rc;
    }
    *out = h;
    return CPML_OK;
}

extern "C" int32_t cpml_destroy(cpml_handle *h)
{
    if (!h) return CPML_OK;
    cudaSetDevice(h->device);
    cpml_p2p_detach(h);
    cudaFree(h->arena);
    cudaFree(h->d_timeout);
    cudaFree(h->d_bcount);
    cudaFree(h->d_scratch);
    for (auto &ax : h->dprof_f) for (auto &p : ax) cudaFree(p);
    for (auto &p : h->mat) cudaFree(p);
    for (auto &ax : h->dprof) for (auto &p : ax) cudaFree(p);
    for (auto &p : h->mx) cudaFree(p);
    for (auto &p : h->my) cudaFree(p);
    for (auto &p : h->mz) cudaFree(p);
    cudaFree(h->d_src_x); cudaFree(h->d_step_out); cudaFree(h->d_ix_rec); cudaFree(h->d_iy_rec);
    cudaFree(h->d_sisvx); cudaFree(h->d_sisvy); cudaFree(h->d_sisp); cudaFree(h->d_sisvz); cudaFree(h->d_ek); cudaFree(h->d_ep);
    cudaFree(h->d_partials); cudaFree(h->d_maxbits);
    cudaFreeHost(h->pin_src); cudaFreeHost(h->pin_out);
    for (int q = 0; q < 2; q++) {
        if (h->snap_done[q]) cudaEventSynchronize(h->snap_done[q]);
        cudaFree(h->snap_dev[q]); cudaFreeHost(h->snap_pin[q]);
        if (h->snap_ready[q]) cudaEventDestroy(h->snap_ready[q]);
        if (h->snap_done[q]) cudaEventDestroy(h->snap_done[q]);
    }
    if (h->snap_stream) cudaStreamDestroy(h->snap_stream);
    for (auto e : h->ev) cudaEventDestroy(e);
    delete h;
    return CPML_OK;
}

extern "C" int32_t cpml_reset(cpml_handle *h)
{
    if (!h) return CPML_EINVAL;
    const cpml_config &c = h->cfg;
    CK(cudaSetDevice(h->device));
    // the fields -- not the slab flags behind them: a neighbour that has already reset and stepped may have
    // published there (slab drivers put a barrier between the last step of a run, the resets and the first step)
    CK(cudaMemsetAsync(h->arena, 0, h->flags_offset * sizeof(double), h->stream));
    CK(cudaMemsetAsync(h->d_timeout, 0, sizeof(unsigned int), h->stream));
    CK(cudaMemsetAsync(h->d_bcount, 0, 8 * sizeof(unsigned int), h->stream));
    h->epoch++;
    if (h->finalized) {
        for (int m = 0; m < 6; m++) {
            if (h->mx[m]) CK(cudaMemsetAsync(h->mx[m], 0, h->mx_doubles * sizeof(double), h->stream));
            if (h->my[m]) CK(cudaMemsetAsync(h->my[m], 0, h->my_doubles * sizeof(double), h->stream));
            if (h->mz[m]) CK(cudaMemsetAsync(h->mz[m], 0, h->mz_doubles * sizeof(double), h->stream));
        }
    }
    const size_t nt = (size_t)c.nstep;
    CK(cudaMemsetAsync(h->d_ek, 0, nt * sizeof(double), h->stream));
    CK(cudaMemsetAsync(h->d_ep, 0, nt * sizeof(double), h->stream));
    CK(cudaMemsetAsync(h->d_step_out, 0, 4 * nt * sizeof(double), h->stream));
    const size_t ns = std::max<size_t>(1, nt * (size_t)c.nrec);
    CK(cudaMemsetAsync(h->d_sisvx, 0, ns * sizeof(double), h->stream));
    CK(cudaMemsetAsync(h->d_sisvy, 0, ns * sizeof(double), h->stream));
    if (h->d_sisp) CK(cudaMemsetAsync(h->d_sisp, 0, ns * sizeof(double), h->stream));
    if (h->d_sisvz) CK(cudaMemsetAsync(h->d_sisvz, 0, ns * sizeof(double), h->stream));
    // 2-D: the paired kernels write fewer partial slots than the one-point geometry allocates, and the
    // 2-D viscoelastic kernels none unless compute_energy is set: the unused slots must read zero
    if (c.ndim == 2 && h->d_partials) CK(cudaMemsetAsync(h->d_partials, 0, 2 * (size_t)h->nblocks * sizeof(double), h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return CPML_OK;
}

extern "C" int32_t cpml_set_stream(cpml_handle *h, void *cuda_stream)
{
    if (!h) return CPML_EINVAL;
    h->stream = (cudaStream_t)cuda_stream;
    return CPML_OK;
}

extern "C" int32_t cpml_set_profiles(cpml_handle *h, int32_t axis, const double *a, const double *b,
                                     const double *K, const double *a_half, const double *b_half,
                                     const double *K_half, int32_t n)
{
    if (!h) return CPML_EINVAL;
    const cpml_config &c = h->cfg;
    if (axis < 0 || axis > 2 || (axis == 2 && c.ndim == 2)) FAIL(CPML_EINVAL, "bad axis");
    const int want = axis == 0 ? c.nx : axis == 1 ? c.ny : c.nz;
    if (n != want) FAIL(CPML_EINVAL, "profile length must be NX, NY or the global NZ");
    if (!a || !b || !K || !a_half || !b_half || !K_half) FAIL(CPML_EINVAL, "null profile");
    if (h->finalized) FAIL(CPML_ESTATE, "profiles cannot change after the first time step (cpml_reset keeps them)");
    CK(cudaSetDevice(h->device));
    const double *src[6] = {a, b, K, a_half, b_half, K_half};
    for (int q = 0; q < 6; q++) {
        h->hprof[axis][q].assign(src[q], src[q] + n);
        for (int i = 0; i < n; i++)
            if (!std::isfinite(src[q][i])) FAIL(CPML_EINVAL, "non-finite profile value");
        if (q == 2 || q == 5)
            for (int i = 0; i < n; i++)
                if (src[q][i] == 0.0) FAIL(CPML_EINVAL, "K profile contains zero");
        if (!h->dprof[axis][q]) CK(cudaMalloc(&h->dprof[axis][q], (size_t)n * sizeof(double)));
        CK(cudaMemcpy(h->dprof[axis][q], src[q], (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
    }
    // correctly rounded reciprocals of the K profiles (div_exact, cpml_internal.h)
    for (int q = 0; q < 2; q++) {
        std::vector<double> r(n);
        for (int i = 0; i < n; i++) {
            const double K = src[q == 0 ? 2 : 5][i];
            if (all_ones_significand(K)) FAIL(CPML_EINVAL, "K profile value with an all-ones significand is not supported");
            r[i] = 1.0 / K;
        }
        if (!h->dprof[axis][6 + q]) CK(cudaMalloc(&h->dprof[axis][6 + q], (size_t)n * sizeof(double)));
        CK(cudaMemcpy(h->dprof[axis][6 + q], r.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
    }
    if (h->f32)
        for (int q = 0; q < 6; q++) {
            std::vector<float> pf(src[q], src[q] + n);
            if (!h->dprof_f[axis][q]) CK(cudaMalloc(&h->dprof_f[axis][q], (size_t)n * sizeof(float)));
            CK(cudaMemcpy(h->dprof_f[axis][q], pf.data(), (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
        }
    h->shell[axis] = find_shell(h->hprof[axis], n);
    h->have_prof[axis] = true;
    return CPML_OK;
}

extern "C" int32_t cpml_set_material_2d(cpml_handle *h, const double *lambda, const double *mu, const double *rho)
{
    if (!h) return CPML_EINVAL;
    const cpml_config &c = h->cfg;
    if (c.ndim != 2) FAIL(CPML_EINVAL, "cpml_set_material_2d is for the 2-D solvers");
    if (!lambda || !mu || !rho) FAIL(CPML_EINVAL, "null material array");
    CK(cudaSetDevice(h->device));
    // div_rho (kernels_2d.cu) needs divisors whose significand is not all ones: rho(i,j) and the
    // interpolated rho_half_x_half_y of 2D-2nd :627, evaluated here exactly as the kernels do
    h->rho_exact = true;
    for (int j = 1; j <= c.ny && h->rho_exact; j++)
        for (int i = 1; i <= c.nx; i++) {
            auto R = [&](int ii, int jj) { return (ii <= c.nx && jj <= c.ny) ? rho[(size_t)(jj - 1) * c.nx + (ii - 1)] : 0.0; };
            const double r0 = R(i, j), rh = 0.25 * (r0 + R(i + 1, j) + R(i + 1, j + 1) + R(i, j + 1));
            if (all_ones_significand(r0) || all_ones_significand(rh) || !(r0 > 0.0)) { h->rho_exact = false; break; }
        }
    const double *src[3] = {lambda, mu, rho};
    for (int m = 0; m < 3; m++) {
        CK(cudaMemset(h->mat[m], 0, h->field_doubles * sizeof(double)));
        CK(cudaMemcpy2D(h->mat[m] + h->origin, (size_t)h->pitch * sizeof(double), src[m],
                        (size_t)c.nx * sizeof(double), (size_t)c.nx * sizeof(double), c.ny,
                        cudaMemcpyHostToDevice));
    }
    h->have_material = true;
    return CPML_OK;
}

extern "C" int32_t cpml_set_attenuation(cpml_handle *h, int32_t n_sls, const double *tau_epsilon_nu1,
                                        const double *tau_sigma_nu1, const double *tau_epsilon_nu2,
                                        const double *tau_sigma_nu2)
{
    if (!h) return CPML_EINVAL;
    if (!h->visco && !h->visco2d) FAIL(CPML_EINVAL, "cpml_set_attenuation is for the viscoelastic solvers (rheology = 1)");
    if (h->visco && n_sls != 2) FAIL(CPML_EINVAL, "the 3-D viscoelastic loop is written for N_SLS = 2 (3D-visco :189, :1003-1049)");
    if (h->visco2d && n_sls != 3) FAIL(CPML_EINVAL, "the 2-D viscoelastic programs use N_SLS = 3 (2D-visco-4th :329)");
    if (!tau_epsilon_nu1 || !tau_sigma_nu1 || !tau_epsilon_nu2 || !tau_sigma_nu2) FAIL(CPML_EINVAL, "null relaxation times");
    const double *src[4] = {tau_epsilon_nu1, tau_sigma_nu1, tau_epsilon_nu2, tau_sigma_nu2};
    for (int q = 0; q < 4; q++)
        for (int l = 0; l < n_sls; l++) {
            if (!(src[q][l] > 0.0) || !std::isfinite(src[q][l])) FAIL(CPML_EINVAL, "relaxation times must be positive");
            h->tau[q][l] = src[q][l];
        }
    if (h->visco) {
        // div_exact divides by den = 1 - DELTAT/2 * tauinv (make_pv): Markstein's argument excludes divisors with
        // an all-ones significand, like K, rho and the grid spacings
        for (int q : {1, 3})
            for (int l = 0; l < n_sls; l++) {
                const double tauinv = -(1.0 / h->tau[q][l]);
                if (all_ones_significand(1.0 - h->cfg.deltat * 0.5 * tauinv))
                    FAIL(CPML_EINVAL, "1 - DELTAT/2 * tauinv has an all-ones significand: not supported");
            }
    }
    h->have_attenuation = true;
    return CPML_OK;
}

extern "C" int32_t cpml_set_source_series(cpml_handle *h, const double *force_x, const double *force_y, int32_t n)
{
    if (!h) return CPML_EINVAL;
    const cpml_config &c = h->cfg;
    if (!force_x || !force_y || n < 1 || n > c.nstep) FAIL(CPML_EINVAL, "bad source series");
    CK(cudaSetDevice(h->device));
    std::vector<double> sx(force_x, force_x + n), sy(force_y, force_y + n);
    if (c.ndim == 3) {
        // the increment of :1080-1081, force * DELTAT / rho, in the reference's order
        for (int q = 0; q < n; q++) {
            sx[q] = force_x[q] * c.deltat / c.rho;
            sy[q] = force_y[q] * c.deltat / c.rho;
        }
    }
    CK(cudaMemcpyAsync(h->d_src_x, sx.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_src_y, sy.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->have_source = true;
    return CPML_OK;
}

extern "C" int32_t cpml_set_receivers(cpml_handle *h, const int32_t *ix_rec, const int32_t *iy_rec, int32_t n)
{
    if (!h) return CPML_EINVAL;
    const cpml_config &c = h->cfg;
    if (n != c.nrec) FAIL(CPML_EINVAL, "n must equal NREC");
    if (n > 0 && (!ix_rec || !iy_rec)) FAIL(CPML_EINVAL, "null receiver arrays");
    for (int r = 0; r < n; r++)
        if (ix_rec[r] < 1 || ix_rec[r] > c.nx || iy_rec[r] < 1 || iy_rec[r] > c.ny)
            FAIL(CPML_EINVAL, "receiver outside the grid");
    CK(cudaSetDevice(h->device));
    if (n > 0) {
        CK(cudaMemcpy(h->d_ix_rec, ix_rec, (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(h->d_iy_rec, iy_rec, (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
    }
    h->have_receivers = true;
    return CPML_OK;
}


// Per-step form of the source and of the loop's per-step outputs, for drivers that keep the
// reference's structure (the source term is evaluated inside `do it`, :1058-1071, and
// total_energy(it) / the seismogram samples are produced every step): pinned staging, copies
// ordered on the handle's stream, no host synchronisation.
extern "C" int32_t cpml_set_source_step(cpml_handle *h, int32_t it, double force_x, double force_y)
{
    if (!h) return CPML_EINVAL;
    const cpml_config &c = h->cfg;
    if (it < 1 || it > c.nstep) FAIL(CPML_EINVAL, "time step outside 1..NSTEP");
    CK(cudaSetDevice(h->device));
    double *sx = h->pin_src + (it - 1), *sy = h->pin_src + c.nstep + (it - 1);
    *sx = c.ndim == 3 ? force_x * c.deltat / c.rho : force_x;      // :1080-1081
    *sy = c.ndim == 3 ? force_y * c.deltat / c.rho : force_y;
    // both components in one copy operation: two 8-byte rows NSTEP doubles apart, on the host and on the device
    CK(cudaMemcpy2DAsync(h->d_src_x + (it - 1), (size_t)c.nstep * sizeof(double), sx, (size_t)c.nstep * sizeof(double),
                         sizeof(double), 2, cudaMemcpyHostToDevice, h->stream));
    h->have_source = true;
    return CPML_OK;
}

extern "C" int32_t cpml_fetch_step(cpml_handle *h, int32_t it)
{
    if (!h) return CPML_EINVAL;
    const cpml_config &c = h->cfg;
    if (it < 1 || it > c.nstep) FAIL(CPML_EINVAL, "time step outside 1..NSTEP");
    CK(cudaSetDevice(h->device));
    // the finisher kernel of the step (k_post3d) left the four values side by side: one 32-byte copy
    CK(cudaMemcpyAsync(h->pin_out + 4 * (size_t)(it - 1), h->d_step_out + 4 * (size_t)(it - 1), 4 * sizeof(double),
                       cudaMemcpyDeviceToHost, h->stream));
    return CPML_OK;
}

extern "C" int32_t cpml_get_fetched_step(cpml_handle *h, int32_t it, double *out4)
{
    if (!h || !out4) return CPML_EINVAL;
    if (it < 1 || it > h->cfg.nstep) FAIL(CPML_EINVAL, "time step outside 1..NSTEP");
    memcpy(out4, h->pin_out + 4 * (size_t)(it - 1), 4 * sizeof(double));
    return CPML_OK;
}

// ---- TMA path: descriptors and work decomposition -------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int32_t encode_plane_map(cpml_handle *h, EncodeTiledFn enc, CUtensorMap *out, int field, int bx, int by)
{
    const cpml_config &c = h->cfg;
    // tensor = the field as (x, y, plane) with the grid's own extents: whatever a box covers
    // beyond NX / NY (or before index 1) is zero-filled by the TMA unit, never read
    const size_t es = h->f32 ? sizeof(float) : sizeof(double);
    void *base = h->f32 ? (void *)((float *)h->field_alloc[field] + h->origin) : (void *)h->f0[field];
    const cuuint64_t dims[3] = {(cuuint64_t)c.nx, (cuuint64_t)c.ny, (cuuint64_t)(h->nzl + 2)};
    const cuuint64_t strides[2] = {(cuuint64_t)h->pitch * es, (cuuint64_t)h->plane * es};
    const cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    // L2 promotion of the tensor loads: 128 B by default; CPML_L2PROMO = 0 none, 1 64 B, 2 128 B, 3 256 B (A/B runs)
    static const CUtensorMapL2promotion promo[4] = {CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_64B,
                                                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B};
    const int pm = std::max(0, std::min(3, env_int("CPML_L2PROMO", 2)));
    const CUresult r = enc(out, h->f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo[pm],
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        h->err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r);
        return CPML_ECUDA;
    }
    return CPML_OK;
}

// One kernel's tile, TMA descriptors and work decomposition.  The two kernels get their own: on the
// 101-wide default grid the stress kernel is fastest on 104 x 7 tiles (384 threads = 12 warps, three per SM
// sub-partition: 168 registers instead of 128, no spills; 1.02 ms against 1.12 ms), the velocity kernel on
// 104 x 8 (1.04 ms against 1.10 ms) -- profiles/r01_v8_tile_104x7.txt.
// Work queue `which` (0 / 1) of the persistent kernels, or null for static shares (CPML_SCHED=static).
static unsigned int *work_queue(const cpml_handle *h, int which)
{
    const char *e = getenv("CPML_SCHED");
    if (e && !strcmp(e, "static")) return nullptr;
    return h->d_bcount + 2 + 2 * which;
}

static int32_t build_tile(cpml_handle *h, EncodeTiledFn enc, bool stress, Tile3D &t, TmaMaps &maps)
{
    const cpml_config &c = h->cfg;
    t.queue = h->use_ws ? work_queue(h, stress ? 1 : 0) : nullptr;
    // thread tile = TMA box.  Narrow grids (the reference's NX = 101) take one 104-wide tile per row, wide
    // grids 128 x 8: the width that wastes the fewest columns (ties: the wider one), measured in
    // profiles/r01_v5_tile_sweep.txt; one CTA per SM, two-plane ring.  CPML_TX / CPML_TY / CPML_STAGES /
    // CPML_ZCHUNKS override both kernels, CPML_TY_STRESS / CPML_ZCHUNKS_STRESS the stress kernel alone.
    int best_tx = 128;
    for (int cand : {104, 64}) {
        const int w_best = (c.nx + best_tx - 1) / best_tx * best_tx, w = (c.nx + cand - 1) / cand * cand;
        if (w < w_best) best_tx = cand;
    }
    // producer-warp kernels: 104 x 8 (14 warps, halo rows 9/8) beats 128 x 7 (15 warps, 8/7) unless it wastes more than
    // 4 % more columns -- 1024-wide slabs: 23.7 against 23.3 Gpts/s (profiles/r02_a_bench.txt)
    if (h->use_ws && best_tx == 128 && (c.nx + 103) / 104 * 104 <= 1.04 * ((c.nx + 127) / 128 * 128)) best_tx = 104;
    t.tx = env_int("CPML_TX", best_tx);
    if (h->use_ws) {
        // with the producer warp a CTA is tile/2 + 32 threads: 104 x 8 -> 14 warps, 128 x 7 -> 15 warps (at most four per
        // SM sub-partition: 128 registers); both kernels use the same tile unless CPML_TY_STRESS says otherwise
        t.ty = env_int("CPML_TY", t.tx == 128 ? 7 : 8);
        if (stress) t.ty = env_int("CPML_TY_STRESS", t.ty);
        if (!ws_tile_supported(t.tx, t.ty, h->f32)) FAIL(CPML_EINVAL, "unsupported CPML_TX x CPML_TY tile for the producer-warp kernels");
    } else {
        t.ty = env_int("CPML_TY", 8);
        if (stress) t.ty = env_int("CPML_TY_STRESS", (t.tx == 104 && t.ty == 8) ? 7 : t.ty);
        if (!tma_tile_supported(t.tx, t.ty)) FAIL(CPML_EINVAL, "unsupported CPML_TX x CPML_TY tile");
    }
    // ring depth: two planes in double precision (three evict each other's lines from L2 before they are consumed:
    // 20.1 against 22.1 Gpts/s); the single-precision stages are half the size and three are 1.4 % faster
    t.stages = std::max(1, std::min(h->use_ws ? 4 : 7, env_int("CPML_STAGES", h->f32 ? 3 : 2)));
    t.minb = std::max(1, std::min(4, env_int("CPML_MINB", t.tx == 64 ? 2 : 1)));
    if (t.ty == 7 || t.ty == 6) t.minb = 1;
    t.xm_bytes = h->shell[0].size() > 0 ? round_up(t.ty * h->sxp * (h->f32 ? 4 : 8), 128) : 0;
    t.ntx = (c.nx + t.tx - 1) / t.tx;
    t.nty = (c.ny + t.ty - 1) / t.ty;

    // descriptors: stress 0 vx 1 vy 2 vz 3 sxx 4 syy 5 szz 6 sxy 7 sxz 8 syz (three halo boxes);
    // velocity 0 sxx 1 syy 2 sxy 3 sxz 4 syz 5 szz 6 vx 7 vy 8 vz (five halo boxes)
    const int hx = t.tx + (h->f32 ? 4 : 2), hy = t.ty + 1;      // halo boxes start 16 bytes before the tile
    const int stress_field[9] = {0, 1, 2, 3, 4, 5, 6, 7, 8};
    const int velocity_field[9] = {3, 4, 6, 7, 8, 5, 0, 1, 2};
    const int nhalo = stress ? 3 : 5;
    for (int m = 0; m < 9; m++) {
        const int32_t rc = encode_plane_map(h, enc, &maps.m[m], stress ? stress_field[m] : velocity_field[m],
                                            m < nhalo ? hx : t.tx, m < nhalo ? hy : t.ty);
        if (rc) return rc;
    }

    // resident CTAs per SM; shrink the ring if a stage set does not fit at all
    Params3D p{};
    p.kunit = 1;
    for (int ax = 0; ax < 3; ax++)
        for (int q : {2, 5})
            for (double K : h->hprof[ax][q])
                if (K != 1.0) p.kunit = 0;
    int occ = 0;
    while (true) {
        const cudaError_t e = h->use_ws ? ws_occupancy(p, t, stress, &occ, h->f32) : tma_occupancy(p, t, stress, &occ);
        if (e == cudaSuccess && occ >= 1) break;
        cudaGetLastError();
        if (t.stages <= 1) FAIL(CPML_ECUDA, "TMA kernels do not fit on this device");
        t.stages--;
    }
    const int cap = env_int("CPML_CTAS_PER_SM", 0);
    if (cap > 0) occ = std::min(occ, cap);

    // z chunks: minimise rounds x (planes per item + pipeline fill) over the persistent grid; among the
    // decompositions within 5 % of the best take the finest one (measured: 18 chunks against 9 on the default
    // grid +1.3 %, 16 against 8 with 104 x 7 tiles +3 %, 2 against 1 on 1024 x 1024 x 128 slabs +4 %)
    const int tiles = t.ntx * t.nty;
    const int resident = h->sm_count * occ;
    const int cmax = std::max(1, h->nzl / 8);
    auto cost_of = [&](int nc, int *ncr_out) {
        const int kc = (h->nzl + nc - 1) / nc;
        const int ncr = (h->nzl + kc - 1) / kc;
        const long long items = (long long)tiles * ncr;
        const double rounds = (double)((items + resident - 1) / resident);
        *ncr_out = ncr;
        return rounds * (kc + 3.0);
    };
    double best_cost = 1e300;
    for (int nc = 1; nc <= cmax; nc++) { int ncr; best_cost = std::min(best_cost, cost_of(nc, &ncr)); }
    int best = 1;
    for (int nc = 1; nc <= cmax; nc++) {
        int ncr;
        if (cost_of(nc, &ncr) <= 1.05 * best_cost) best = std::max(best, ncr);
    }
    int nzc = env_int("CPML_ZCHUNKS", 0);
    if (stress) nzc = env_int("CPML_ZCHUNKS_STRESS", nzc);
    // producer-warp stress kernel: twice as many chunks again (default grid: 40 chunks of 16 planes 0.967 ms against
    // 0.995 ms with 20; 64 or 80 chunks no better -- profiles/r02_c_bench.txt); its item-boundary cost is small because
    // the producer starts the next item's loads while the consumers finish the current one
    if (nzc <= 0) nzc = (stress && h->use_ws) ? std::min(cmax, 2 * best) : best;
    nzc = std::max(1, std::min(nzc, h->nzl));
    t.kchunk = (h->nzl + nzc - 1) / nzc;
    t.nzc = (h->nzl + t.kchunk - 1) / t.kchunk;
    // finer tail (producer-warp kernels with the work queue): the last 1.5 x resident-CTAs coarse items (measured:
    // profiles/r02_k_bench_finer_tail.txt; CPML_TAIL_ITEMS_PCT, CPML_TAIL_SPLIT for A/B runs) -- never the two
    // boundary chunks, which are first in the list -- are split into items of >= 8 planes (decode_item)
    const int coarse = tiles * t.nzc;
    t.split = 1;
    t.fine_from = coarse;
    if (h->use_ws && t.queue && t.nzc >= 3) {
        const int sp = env_int("CPML_TAIL_SPLIT", std::min(4, t.kchunk / 8));       // (1: no finer tail)
        if (sp >= 2 && sp <= 8 && (sp - 1) * ((t.kchunk + sp - 1) / sp) < t.kchunk) {  // every part holds a plane
            t.split = sp;
            const int pct = std::max(0, env_int("CPML_TAIL_ITEMS_PCT", 150));      // coarse items split, % of the resident CTAs
            t.fine_from = coarse - std::min(coarse - 2 * tiles, (int)((long long)resident * pct / 100));
        }
    }
    t.nitems = t.fine_from + (coarse - t.fine_from) * t.split;
    t.grid_stress = t.grid_velocity = std::min(t.nitems, h->sm_count * occ);
    return CPML_OK;
}

static int32_t setup_tma(cpml_handle *h)
{
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) FAIL(CPML_ECUDA, "driver does not export cuTensorMapEncodeTiled");
    const EncodeTiledFn enc = (EncodeTiledFn)fn;
    int32_t rc = build_tile(h, enc, true, h->tile_stress, h->maps_stress);
    if (rc) return rc;
    rc = build_tile(h, enc, false, h->tile, h->maps_velocity);
    if (rc) return rc;
    h->nblocks = h->tile.nitems;        // energy partial slots: one per item of the velocity kernel
    return CPML_OK;
}

// Tile, tensor maps and work decomposition of the TMA-staged viscoelastic velocity kernel.  The tensors are the padded
// arrays (x offset 16, two ghost rows, planes -1 .. NZ_LOCAL+2), so every edge tap reads the zeros the reference reads.
static int32_t setup_visco_ws(cpml_handle *h)
{
    const cpml_config &c = h->cfg;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) FAIL(CPML_ECUDA, "driver does not export cuTensorMapEncodeTiled");
    const EncodeTiledFn enc = (EncodeTiledFn)fn;
    Tile3D &t = h->vtile;
    t.queue = work_queue(h, 0);
    int best_tx = 104;
    for (int cand : {108, 64}) {      // the width that wastes the fewest columns
        const int w_best = (c.nx + best_tx - 1) / best_tx * best_tx, w = (c.nx + cand - 1) / cand * cand;
        if (w < w_best) best_tx = cand;
    }
    t.tx = env_int("CPML_VWS_TX", best_tx);
    t.ty = 8;
    if (!vws_tile_supported(t.tx, t.ty)) FAIL(CPML_EINVAL, "unsupported CPML_VWS_TX tile");
    t.stages = std::max(1, std::min(3, env_int("CPML_VWS_STAGES", 2)));
    t.minb = 1;
    t.xm_bytes = h->shell[0].size() > 0 ? round_up(t.ty * h->sxp * 8, 128) : 0;
    t.ntx = (c.nx + t.tx - 1) / t.tx;
    t.nty = (c.ny + t.ty - 1) / t.ty;
    // maps: 0 sxx 1 sxy 2 syy 3 sxz 4 syz 5 szz 6 vx 7 vy 8 vz
    const int field[9] = {3, 6, 4, 7, 8, 5, 0, 1, 2};
    int box[9][2];
    vws_boxes(t.tx, t.ty, box);
    for (int m = 0; m < 9; m++) {
        const cuuint64_t dims[3] = {(cuuint64_t)h->pitch, (cuuint64_t)(c.ny + 4), (cuuint64_t)(h->nzl + 4)};
        const cuuint64_t strides[2] = {(cuuint64_t)h->pitch * sizeof(double), (cuuint64_t)h->plane * sizeof(double)};
        const cuuint32_t bx[3] = {(cuuint32_t)box[m][0], (cuuint32_t)box[m][1], 1u};
        const cuuint32_t estr[3] = {1u, 1u, 1u};
        const CUresult r = enc(&h->vmaps.m[m], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void *)h->field_alloc[field[m]], dims, strides, bx, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) FAIL(CPML_ECUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    }
    int occ = 0;
    while (true) {
        const cudaError_t e = vws_occupancy(t, &occ);
        if (e == cudaSuccess && occ >= 1) break;
        cudaGetLastError();
        if (t.stages <= 1) FAIL(CPML_ECUDA, "the TMA-staged viscoelastic velocity kernel does not fit on this device");
        t.stages--;
    }
    // z chunks: rounds x (planes per item + pipeline fill and window preload) over the persistent grid, finest within 5 %
    const int tiles = t.ntx * t.nty, resident = h->sm_count * occ, cmax = std::max(1, h->nzl / 8);
    auto cost_of = [&](int nc, int *ncr_out) {
        const int kc = (h->nzl + nc - 1) / nc, ncr = (h->nzl + kc - 1) / kc;
        *ncr_out = ncr;
        return (double)(((long long)tiles * ncr + resident - 1) / resident) * (kc + 4.0);
    };
    double best_cost = 1e300;
    for (int nc = 1; nc <= cmax; nc++) { int ncr; best_cost = std::min(best_cost, cost_of(nc, &ncr)); }
    int best = 1;
    for (int nc = 1; nc <= cmax; nc++) { int ncr; if (cost_of(nc, &ncr) <= 1.05 * best_cost) best = std::max(best, ncr); }
    int nzc = env_int("CPML_VWS_ZCHUNKS", 0);
    if (nzc <= 0) nzc = best;
    nzc = std::max(1, std::min(nzc, h->nzl));
    t.kchunk = (h->nzl + nzc - 1) / nzc;
    t.nzc = (h->nzl + t.kchunk - 1) / t.kchunk;
    t.nitems = tiles * t.nzc;
    t.grid_stress = t.grid_velocity = std::min(t.nitems, h->sm_count * occ);
    return CPML_OK;
}

// Tensor maps and work decomposition of the TMA-staged 2-D kernels.  The tensors are the padded arrays (x offset 16, two
// ghost rows), so every edge tap reads the zeros of the reference's (0:NX+1,0:NY+1) arrays from memory.
static int32_t setup_2d_ws(cpml_handle *h)
{
    const cpml_config &c = h->cfg;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) FAIL(CPML_ECUDA, "driver does not export cuTensorMapEncodeTiled");
    const EncodeTiledFn enc = (EncodeTiledFn)fn;
    Tile2D &t = h->tile2;
    t.queue = work_queue(h, 0);
    int bs[7][2], bv[6][2];
    ws2_geometry(&t.tx, &t.rb, bs, bv);
    auto encode = [&](CUtensorMap *out, double *base, const int (&box)[2]) -> int32_t {
        const cuuint64_t dims[2] = {(cuuint64_t)h->pitch, (cuuint64_t)(c.ny + 4)};
        const cuuint64_t strides[1] = {(cuuint64_t)h->pitch * sizeof(double)};
        const cuuint32_t bx[2] = {(cuuint32_t)box[0], (cuuint32_t)box[1]};
        const cuuint32_t estr[2] = {1u, 1u};
        const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)base, dims, strides, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) FAIL(CPML_ECUDA, "cuTensorMapEncodeTiled (2-D) failed with CUresult " + std::to_string((int)r));
        return CPML_OK;
    };
    // stress: vx vy lambda mu sxx syy sxy ; velocity: sxx sxy rho syy vx vy
    double *sb[7] = {h->field_alloc[0], h->field_alloc[1], h->mat[0], h->mat[1], h->field_alloc[2], h->field_alloc[3], h->field_alloc[4]};
    double *vb[6] = {h->field_alloc[2], h->field_alloc[4], h->mat[2], h->field_alloc[3], h->field_alloc[0], h->field_alloc[1]};
    for (int m = 0; m < 7; m++) { const int32_t rc = encode(&h->maps2_stress.m[m], sb[m], bs[m]); if (rc) return rc; }
    for (int m = 0; m < 6; m++) { const int32_t rc = encode(&h->maps2_velocity.m[m], vb[m], bv[m]); if (rc) return rc; }
    int occ_s = 0, occ_v = 0;
    if (ws2_occupancy(c.order, true, &occ_s) != cudaSuccess || ws2_occupancy(c.order, false, &occ_v) != cudaSuccess || occ_s < 1 || occ_v < 1) {
        cudaGetLastError();
        FAIL(CPML_ECUDA, "the TMA-staged 2-D kernels do not fit on this device");
    }
    t.ntx = (c.nx + t.tx - 1) / t.tx;
    // y chunks: SHORT ones, about ten row blocks each.  Measured at 4096 x 4096 (profiles/r02_g_bench_2d.txt): 6 chunks (one
    // round of long items) 16.9 Gpts/s, 27 chunks 25.3, 64 chunks 26.8, 100 chunks 27.6 -- although every chunk re-reads two
    // blocks of taps.  With few long items the CTAs of an SM start together and stay in phase (all waiting for the ring,
    // then all computing); many short items put them out of phase and spread the rows in flight over the DRAM channels.
    const int resident = h->sm_count * std::min(occ_s, occ_v);
    const int cmax = std::max(1, c.ny / (8 * t.rb));
    const int best = std::max(1, std::min(cmax, (c.ny + 10 * t.rb - 1) / (10 * t.rb)));
    (void)resident;
    int nc = env_int("CPML_2D_CHUNKS", 0);
    if (nc <= 0) nc = best;
    nc = std::max(1, std::min(nc, cmax));
    t.rows = ((c.ny + nc - 1) / nc + t.rb - 1) / t.rb * t.rb;
    t.nchunks = (c.ny + t.rows - 1) / t.rows;
    t.nitems = t.ntx * t.nchunks;
    if (t.nitems > h->nblocks) FAIL(CPML_EINVAL, "internal: more 2-D work items than energy partial slots");
    t.grid_stress = std::min(t.nitems, h->sm_count * occ_s);
    t.grid_velocity = std::min(t.nitems, h->sm_count * occ_v);
    return CPML_OK;
}

// Allocates the shell-only memory variables once every input is known.
static int32_t finalize(cpml_handle *h)
{
    if (h->finalized) return CPML_OK;
    const cpml_config &c = h->cfg;
    const int naxes = c.ndim;
    for (int ax = 0; ax < naxes; ax++)
        if (!h->have_prof[ax]) FAIL(CPML_ESTATE, "cpml_set_profiles has not been called for every axis");
    if (!h->have_source) FAIL(CPML_ESTATE, "cpml_set_source_series has not been called");
    if (c.nrec > 0 && !h->have_receivers) FAIL(CPML_ESTATE, "cpml_set_receivers has not been called");
    if (c.ndim == 2 && !h->have_material) FAIL(CPML_ESTATE, "cpml_set_material_2d has not been called");
    if ((h->visco || h->visco2d) && !h->have_attenuation) FAIL(CPML_ESTATE, "cpml_set_attenuation has not been called");
    CK(cudaSetDevice(h->device));

    const Shell &sx = h->shell[0], &sy = h->shell[1];
    h->sxp = std::max(4, round_up(sx.size(), 4));
    h->sy = std::max(1, sy.size());
    const int nmem = c.ndim == 3 ? 6 : 4;
    h->mx_doubles = (size_t)h->sxp * c.ny * h->nzl + (size_t)h->sxp * 16;   // + slack: bulk copies of a ragged last y tile
    h->my_doubles = (size_t)h->pitch * h->sy * h->nzl;
    for (int m = 0; m < nmem; m++) {
        CK(cudaMalloc(&h->mx[m], h->mx_doubles * sizeof(double)));
        CK(cudaMalloc(&h->my[m], h->my_doubles * sizeof(double)));
        CK(cudaMemset(h->mx[m], 0, h->mx_doubles * sizeof(double)));
        CK(cudaMemset(h->my[m], 0, h->my_doubles * sizeof(double)));
    }
    if (c.ndim == 3) {
        const Shell &sz = h->shell[2];
        int first = -1, last = -2;
        for (int k = 1; k <= h->nzl; k++) {
            const int kg = k + h->koff;
            if (kg <= sz.lo || kg >= sz.hi) {
                const int s = kg <= sz.lo ? kg - 1 : sz.lo + (kg - sz.hi);
                if (first < 0) first = s;
                last = s;
            }
        }
        h->zbase = std::max(0, first);
        h->sz_local = first < 0 ? 0 : last - first + 1;
        h->mz_doubles = (size_t)h->pitch * c.ny * std::max(1, h->sz_local);
        for (int m = 0; m < 6; m++) {
            CK(cudaMalloc(&h->mz[m], h->mz_doubles * sizeof(double)));
            CK(cudaMemset(h->mz[m], 0, h->mz_doubles * sizeof(double)));
        }
    }
    // own non-zero counts per axis (for cpml_algorithmic_bytes)
    for (int ax = 0; ax < naxes; ax++) {
        int n0 = 0, n1 = 0;
        const int n = (int)h->hprof[ax][0].size();
        int lo = 1, hi = n;
        if (ax == 2) { lo = h->koff + 1; hi = h->koff + h->nzl; }
        for (int i = lo; i <= hi; i++) {
            n0 += (h->hprof[ax][0][i - 1] != 0.0) || (h->hprof[ax][2][i - 1] != 1.0);
            n1 += (h->hprof[ax][3][i - 1] != 0.0) || (h->hprof[ax][5][i - 1] != 1.0);
        }
        h->nz_own[ax][0] = n0;
        h->nz_own[ax][1] = n1;
    }
    if (h->visco) {
        // the viscoelastic kernels index with 32-bit element offsets
        if (h->field_doubles >= (1ull << 31) || h->mx_doubles >= (1ull << 31) || h->my_doubles >= (1ull << 31) || h->mz_doubles >= (1ull << 31))
            FAIL(CPML_EINVAL, "viscoelastic slab too large: a field must hold fewer than 2^31 points (use more z-slabs)");
        visco_tile(&h->vtx, &h->vty);
        // z chunks: short enough for several waves of blocks, long enough that the three extra
        // velocity planes fetched to fill the z windows stay small against 87 words per point
        int kc = env_int("CPML_VKCHUNK", 16);
        kc = std::max(1, std::min(kc, h->nzl));
        const int nzc = (h->nzl + kc - 1) / kc;
        h->vkchunk = (h->nzl + nzc - 1) / nzc;
        h->vgrid = dim3((c.nx + h->vtx - 1) / h->vtx, (c.ny + h->vty - 1) / h->vty, (h->nzl + h->vkchunk - 1) / h->vkchunk);
        h->nblocks = (int)(h->vgrid.x * h->vgrid.y * h->vgrid.z);
        // CPML_VKERNEL=reg keeps the register-marching velocity kernel (A/B runs)
        const char *vk = getenv("CPML_VKERNEL");
        h->v_ws = !(vk && std::string(vk) == "reg");
        if (h->v_ws) { const int32_t rc = setup_visco_ws(h); if (rc) return rc; }
        // energy partial slots: [0, nblocks) kinetic (velocity kernel: one per block or per work item), then two
        // regions of nblocks for the potential parts of the stress launches; slots a kernel never writes stay zero
        if (h->v_ws) h->nblocks = std::max(h->nblocks, h->vtile.nitems);
        CK(cudaMalloc(&h->d_partials, 3 * (size_t)h->nblocks * sizeof(double)));
        CK(cudaMemset(h->d_partials, 0, 3 * (size_t)h->nblocks * sizeof(double)));
    } else if (c.ndim == 3) {
        // CPML_KERNEL=reg selects the register-marching kernels of kernels_3d.cu, CPML_KERNEL=tma the TMA-staged
        // kernels without a producer warp (kernels_3d_tma.cu); both are kept for A/B runs
        const char *kv = getenv("CPML_KERNEL");
        const std::string kvs = kv ? kv : "";
        h->use_tma = kvs != "reg";
        h->use_ws = h->use_tma && kvs != "tma";
        if (h->use_ws && (h->field_doubles >= (1ull << 32) || h->mx_doubles >= (1ull << 32) || h->my_doubles >= (1ull << 32) ||
                          h->mz_doubles >= (1ull << 32)))
            h->use_ws = false;      // the producer-warp kernels index with 32-bit element offsets (fields of < 2^32 points)
        if (h->f32 && !h->use_ws) FAIL(CPML_EINVAL, "single precision needs the producer-warp kernels (unset CPML_KERNEL)");
        if (h->f32) CK(cudaMalloc(&h->d_scratch, (size_t)h->plane * sizeof(double)));
        if (h->use_tma) { const int32_t rc = setup_tma(h); if (rc) return rc; }
        else build_regions(h);
        CK(cudaMalloc(&h->d_partials, 2 * (size_t)std::max(1, h->nblocks) * sizeof(double)));
        CK(cudaMemset(h->d_partials, 0, 2 * (size_t)std::max(1, h->nblocks) * sizeof(double)));
    }
    if (c.ndim == 2 && !h->visco2d) {
        // CPML_2D_KERNEL=pair keeps the one-launch-per-point kernels of kernels_2d.cu (A/B runs)
        const char *k2 = getenv("CPML_2D_KERNEL");
        h->ws2 = !(k2 && std::string(k2) == "pair");
        if (h->ws2) {
            const int32_t rc = setup_2d_ws(h); if (rc) return rc;
            // energy partials: one slot per work item (the one-point geometry of create has 65536 slots at 4096 x 4096,
            // which the one-block finisher k_post3d spent 43 us per step summing: 6 % of the step)
            h->nblocks = h->tile2.nitems;
        }
    }
    h->finalized = true;
    return CPML_OK;
}

static AxisCoef coef_view(cpml_handle *h, int ax)
{
    // 1-based views: index i addresses element i-1 of the device array
    AxisCoef c;
    c.a = h->dprof[ax][0] - 1; c.b = h->dprof[ax][1] - 1; c.K = h->dprof[ax][2] - 1;
    c.a_half = h->dprof[ax][3] - 1; c.b_half = h->dprof[ax][4] - 1; c.K_half = h->dprof[ax][5] - 1;
    c.rK = h->dprof[ax][6] - 1; c.rK_half = h->dprof[ax][7] - 1;
    return c;
}

static Params3D make_p3(cpml_handle *h, int it)
{
    const cpml_config &c = h->cfg;
    Params3D p{};
    p.nx = c.nx; p.ny = c.ny; p.nzl = h->nzl; p.nz = c.nz; p.koff = h->koff;
    p.pitch = h->pitch; p.plane = h->plane;
    p.vx = h->f0[0]; p.vy = h->f0[1]; p.vz = h->f0[2];
    p.sxx = h->f0[3]; p.syy = h->f0[4]; p.szz = h->f0[5];
    p.sxy = h->f0[6]; p.sxz = h->f0[7]; p.syz = h->f0[8];
    p.xlo = h->shell[0].lo; p.xhi = h->shell[0].hi; p.sxp = h->sxp;
    p.ylo = h->shell[1].lo; p.yhi = h->shell[1].hi; p.sy = h->sy;
    p.zlo = h->shell[2].lo; p.zhi = h->shell[2].hi; p.zbase = h->zbase;
    for (int m = 0; m < 6; m++) { p.mx[m] = h->mx[m]; p.my[m] = h->my[m]; p.mz[m] = h->mz[m]; }
    p.cx = coef_view(h, 0); p.cy = coef_view(h, 1); p.cz = coef_view(h, 2);
    p.odx = 1.0 / c.deltax; p.ody = 1.0 / c.deltay; p.odz = 1.0 / c.deltaz;   // :134-136
    p.dt_lambda = c.deltat * c.lambda;                                        // :296-300
    p.dt_mu = c.deltat * c.mu;
    p.dt_lambdaplus2mu = c.deltat * c.lambdaplustwomu;
    p.dt_over_rho = c.deltat / c.rho;
    p.it = it;
    p.isrc = c.isource; p.jsrc = c.jsource;
    const int kl = h->ksrc_global - h->koff;
    p.ksrc = (kl >= 1 && kl <= h->nzl) ? kl : 0;
    p.src_x = h->d_src_x; p.src_y = h->d_src_y;
    p.npml = c.npoints_pml; p.energy_bug_compat = c.energy_bug_compat;
    p.rho = c.rho; p.lambda = c.lambda; p.mu = c.mu;
    p.inv_den = 1.0 / (2.0 * c.mu * (3.0 * c.lambda + 2.0 * c.mu));
    p.inv_2mu = 1.0 / (2.0 * c.mu);
    p.inv_mu = 1.0 / c.mu;
    p.c2lm = 2.0 * (c.lambda + c.mu);
    p.half_rho = 0.5 * c.rho;
    p.partials = h->d_partials; p.nblocks = h->nblocks;
    p.kunit = 1;
    for (int ax = 0; ax < 3; ax++)
        for (int q : {2, 5})
            for (double K : h->hprof[ax][q])
                if (K != 1.0) p.kunit = 0;
    // neighbour halo planes (3D-iso :811-823, :951-963): lo gets vx, vy, sigmazz of my plane 1
    // in its plane NZ_LOCAL+1; hi gets vz, sigmaxz, sigmayz of my plane NZ_LOCAL in its plane 0
    const int lo_f[3] = {0, 1, 5}, hi_f[3] = {2, 7, 8};
    for (int q = 0; q < 3; q++) {
        p.peer_lo[q] = h->peer_on[0] ? h->peer_arena[0] + (size_t)lo_f[q] * h->field_doubles + h->origin + (long long)(h->nzl + 1) * h->plane : nullptr;
        p.peer_hi[q] = h->peer_on[1] ? h->peer_arena[1] + (size_t)hi_f[q] * h->field_doubles + h->origin : nullptr;
    }
    return p;
}

// The same block for the single-precision kernels: update constants rounded ONCE from their double values, fields
// and memory variables viewed as float arrays (same element counts and offsets), profiles from the float copies.
static Params3DF make_p3f(cpml_handle *h, int it)
{
    const Params3D d = make_p3(h, it);
    Params3DF p{};
    p.nx = d.nx; p.ny = d.ny; p.nzl = d.nzl; p.nz = d.nz; p.koff = d.koff; p.pitch = d.pitch; p.plane = d.plane;
    float **fld[9] = {&p.vx, &p.vy, &p.vz, &p.sxx, &p.syy, &p.szz, &p.sxy, &p.sxz, &p.syz};
    for (int f = 0; f < 9; f++) *fld[f] = (float *)h->field_alloc[f] + h->origin;
    p.xlo = d.xlo; p.xhi = d.xhi; p.sxp = d.sxp; p.ylo = d.ylo; p.yhi = d.yhi; p.sy = d.sy;
    p.zlo = d.zlo; p.zhi = d.zhi; p.zbase = d.zbase;
    for (int m = 0; m < 6; m++) { p.mx[m] = (float *)h->mx[m]; p.my[m] = (float *)h->my[m]; p.mz[m] = (float *)h->mz[m]; }
    AxisCoefT<float> *ax[3] = {&p.cx, &p.cy, &p.cz};
    for (int a = 0; a < 3; a++) {
        ax[a]->a = h->dprof_f[a][0] - 1; ax[a]->b = h->dprof_f[a][1] - 1; ax[a]->K = h->dprof_f[a][2] - 1;
        ax[a]->a_half = h->dprof_f[a][3] - 1; ax[a]->b_half = h->dprof_f[a][4] - 1; ax[a]->K_half = h->dprof_f[a][5] - 1;
        ax[a]->rK = nullptr; ax[a]->rK_half = nullptr;
    }
    p.odx = (float)d.odx; p.ody = (float)d.ody; p.odz = (float)d.odz;
    p.dt_lambda = (float)d.dt_lambda; p.dt_mu = (float)d.dt_mu; p.dt_lambdaplus2mu = (float)d.dt_lambdaplus2mu;
    p.dt_over_rho = (float)d.dt_over_rho;
    p.it = d.it; p.isrc = d.isrc; p.jsrc = d.jsrc; p.ksrc = d.ksrc; p.src_x = d.src_x; p.src_y = d.src_y;
    p.npml = d.npml; p.energy_bug_compat = d.energy_bug_compat;
    p.rho = d.rho; p.lambda = d.lambda; p.mu = d.mu; p.inv_den = d.inv_den; p.inv_2mu = d.inv_2mu;
    p.inv_mu = d.inv_mu; p.c2lm = d.c2lm; p.half_rho = d.half_rho;
    p.partials = d.partials; p.nblocks = d.nblocks; p.kunit = d.kunit;
    return p;
}

static Params2D make_p2(cpml_handle *h, int it)
{
    const cpml_config &c = h->cfg;
    Params2D p{};
    p.nx = c.nx; p.ny = c.ny; p.pitch = h->pitch; p.order = c.order;
    p.vx = h->f0[0]; p.vy = h->f0[1]; p.sxx = h->f0[2]; p.syy = h->f0[3]; p.sxy = h->f0[4];
    p.lambda = h->mat[0] + h->origin; p.mu = h->mat[1] + h->origin; p.rho = h->mat[2] + h->origin;
    p.rho_exact = h->rho_exact ? 1 : 0;
    p.xlo = h->shell[0].lo; p.xhi = h->shell[0].hi; p.sxp = h->sxp;
    p.ylo = h->shell[1].lo; p.yhi = h->shell[1].hi; p.sy = h->sy;
    for (int m = 0; m < 4; m++) { p.mx[m] = h->mx[m]; p.my[m] = h->my[m]; }
    p.cx = coef_view(h, 0); p.cy = coef_view(h, 1);
    p.deltax = c.deltax; p.deltay = c.deltay; p.deltat = c.deltat;
    p.denx = c.order == 4 ? 24.0 * c.deltax : c.deltax;      // 2D-4th :565 / 2D-2nd :564
    p.deny = c.order == 4 ? 24.0 * c.deltay : c.deltay;
    p.rdenx = 1.0 / p.denx; p.rdeny = 1.0 / p.deny;
    if (h->visco2d) {
        // 2D-visco-4th :210-213 (second-order file :209-210) and :386-399, in the reference's order
        const double DELTAT = c.deltat;
        p.c98x = c.order == 4 ? 9.0 / (8.0 * c.deltax) : 1.0 / c.deltax;
        p.c98y = c.order == 4 ? 9.0 / (8.0 * c.deltay) : 1.0 / c.deltay;
        p.c24x = 1.0 / (24.0 * c.deltax); p.c24y = 1.0 / (24.0 * c.deltay);
        const double *te1 = h->tau[0], *ts1 = h->tau[1], *te2 = h->tau[2], *ts2 = h->tau[3];
        double sum1 = 0.0, sum2 = 0.0;
        for (int l = 0; l < 3; l++) { sum1 = sum1 + te1[l] / ts1[l]; sum2 = sum2 + te2[l] / ts2[l]; }
        for (int l = 0; l < 3; l++) {
            const double one_over_tau_sigma_nu1 = 1.0 / ts1[l], one_over_tau_sigma_nu2 = 1.0 / ts2[l];
            p.half1[l] = 0.5 * DELTAT / ts1[l];
            p.half2[l] = 0.5 * DELTAT / ts2[l];
            p.mul1[l] = 1.0 / (1.0 + 0.5 * DELTAT * one_over_tau_sigma_nu1);
            p.mul2[l] = 1.0 / (1.0 + 0.5 * DELTAT * one_over_tau_sigma_nu2);
            p.dt_phi1[l] = DELTAT * (1.0 - te1[l] / ts1[l]) / ts1[l] / sum1;
            p.dt_phi2[l] = DELTAT * (1.0 - te2[l] / ts2[l]) / ts2[l] / sum2;
            p.e1[l] = h->f0[5 + l]; p.e11[l] = h->f0[8 + l]; p.e13[l] = h->f0[11 + l];
        }
    }
    p.it = it; p.isrc = c.isource; p.jsrc = c.jsource;
    p.force_x = h->d_src_x; p.force_y = h->d_src_y;
    p.npml = c.npoints_pml;
    p.partials = h->d_partials; p.nblocks = h->nblocks;
    return p;
}

static ParamsV3D make_pv(cpml_handle *h, int it)
{
    const cpml_config &c = h->cfg;
    ParamsV3D p{};
    p.nx = c.nx; p.ny = c.ny; p.nzl = h->nzl; p.nz = c.nz; p.koff = h->koff;
    p.pitch = h->pitch; p.plane = h->plane;
    p.vx = h->f0[0]; p.vy = h->f0[1]; p.vz = h->f0[2];
    p.sxx = h->f0[3]; p.syy = h->f0[4]; p.szz = h->f0[5];
    p.sxy = h->f0[6]; p.sxz = h->f0[7]; p.syz = h->f0[8];
    p.rxx = h->f0[9]; p.ryy = h->f0[10]; p.rzz = h->f0[11];
    p.rxy = h->f0[12]; p.rxz = h->f0[13]; p.ryz = h->f0[14];
    p.e1 = h->e0[0]; p.e11 = h->e0[1]; p.e22 = h->e0[2]; p.e12 = h->e0[3]; p.e13 = h->e0[4]; p.e23 = h->e0[5];
    p.xlo = h->shell[0].lo; p.xhi = h->shell[0].hi; p.sxp = h->sxp;
    p.ylo = h->shell[1].lo; p.yhi = h->shell[1].hi; p.sy = h->sy;
    p.zlo = h->shell[2].lo; p.zhi = h->shell[2].hi; p.zbase = h->zbase;
    for (int m = 0; m < 6; m++) { p.mx[m] = h->mx[m]; p.my[m] = h->my[m]; p.mz[m] = h->mz[m]; }
    p.cx = coef_view(h, 0); p.cy = coef_view(h, 1); p.cz = coef_view(h, 2);
    p.odx = 1.0 / c.deltax; p.ody = 1.0 / c.deltay; p.odz = 1.0 / c.deltaz;   // 3D-visco :162-164
    p.dt = c.deltat;
    p.dt_over_rho = c.deltat / c.rho;                                          // :337
    // :458-477 and :982-987, same expressions and order as the reference
    const double ONE = 1.0, TWO = 2.0, DIM = 3.0;
    const double *te1 = h->tau[0], *ts1 = h->tau[1], *te2 = h->tau[2], *ts2 = h->tau[3];
    for (int l = 0; l < 2; l++) {
        p.tauinv1[l] = -(ONE / ts1[l]);
        p.tauinv2[l] = -(ONE / ts2[l]);
        p.phi1[l] = (ONE - te1[l] / ts1[l]) / ts1[l];
        p.phi2[l] = (ONE - te2[l] / ts2[l]) / ts2[l];
        p.den1[l] = 1.0 - c.deltat * 0.5 * p.tauinv1[l];
        p.den2[l] = 1.0 - c.deltat * 0.5 * p.tauinv2[l];
        p.rden1[l] = 1.0 / p.den1[l];
        p.rden2[l] = 1.0 / p.den2[l];
    }
    const double Mu_nu1 = ONE - (ONE - te1[0] / ts1[0]) - (ONE - te1[1] / ts1[1]);
    const double Mu_nu2 = ONE - (ONE - te2[0] / ts2[0]) - (ONE - te2[1] / ts2[1]);
    const double mul_relaxed = c.mu, lambdal_relaxed = c.lambda;
    p.lam = lambdal_relaxed; p.mu = mul_relaxed;
    p.l2m_r = lambdal_relaxed + TWO * mul_relaxed;
    p.lam_u = (lambdal_relaxed + 2.0 / DIM * mul_relaxed) * Mu_nu1 - 2.0 / DIM * mul_relaxed * Mu_nu2;
    p.mu_u = mul_relaxed * Mu_nu2;
    p.l2m_u = p.lam_u + TWO * p.mu_u;
    p.lam23mu = lambdal_relaxed + 2.0 / DIM * mul_relaxed;
    p.two_mu = TWO * mul_relaxed;
    p.two_thirds_mu = TWO / DIM * mul_relaxed;
    p.szz_e1 = c.sigmazz_isotropic ? p.lam23mu : p.l2m_r;
    p.szz_dev = c.sigmazz_isotropic ? p.two_mu : p.two_thirds_mu;
    p.nzl_e = c.emulate_nproc > 1 ? c.nz / c.emulate_nproc : c.nz;
    p.it = it;
    p.isrc = c.isource; p.jsrc = c.jsource;
    const int kl = h->ksrc_global - h->koff;
    p.ksrc = (kl >= 1 && kl <= h->nzl) ? kl : 0;
    p.src_x = h->d_src_x; p.src_y = h->d_src_y;
    p.npml = c.npoints_pml;
    p.half_rho = 0.5 * c.rho;
    p.c2lm = 2.0 * (c.lambda + c.mu);
    p.inv_den = 1.0 / (2.0 * c.mu * (3.0 * c.lambda + 2.0 * c.mu));
    p.inv_2mu = 1.0 / (2.0 * c.mu);
    p.partials = h->d_partials; p.nblocks = h->nblocks;
    p.kchunk = h->vkchunk;
    // neighbour slabs' fields (element (1,1,0)) for the in-kernel halo stores: vx vy sigmazz / vz sigmaxz sigmayz
    const int pf[6] = {0, 1, 5, 2, 7, 8};
    for (int q = 0; q < 6; q++) {
        p.peer_lo[q] = h->peer_on[0] ? h->peer_arena[0] + (size_t)pf[q] * h->field_doubles + h->origin : nullptr;
        p.peer_hi[q] = h->peer_on[1] ? h->peer_arena[1] + (size_t)pf[q] * h->field_doubles + h->origin : nullptr;
    }
    return p;
}

static unsigned long long *peer_flags(cpml_handle *h, int side)
{
    return (unsigned long long *)(h->peer_arena[side] + h->flags_offset);
}

static const size_t kEventSlots = 4096;      // launches in flight before the oldest pair is harvested

// Adds the elapsed time of the oldest recorded pair to its kernel's total and frees the slot.
static int32_t harvest_oldest(cpml_handle *h)
{
    const size_t q = h->ev_head;
    float ms = 0.f;
    CK(cudaEventSynchronize(h->ev[2 * q + 1]));
    CK(cudaEventElapsedTime(&ms, h->ev[2 * q], h->ev[2 * q + 1]));
    (h->ev_kind[q] == 0 ? h->ms_stress : h->ms_velocity) += ms;
    h->ev_kind[q] = -1;
    h->ev_head = (h->ev_head + 1) % kEventSlots;
    h->ev_count--;
    return CPML_OK;
}

static int32_t time_begin(cpml_handle *h, int kind)
{
    if (!h->timing) return CPML_OK;
    if (h->ev_count == kEventSlots) { const int32_t rc = harvest_oldest(h); if (rc) return rc; }
    const size_t q = (h->ev_head + h->ev_count) % kEventSlots;
    h->ev_kind[q] = kind;
    h->ev_count++;
    CK(cudaEventRecord(h->ev[2 * q], h->stream));
    return CPML_OK;
}
static int32_t time_end(cpml_handle *h)
{
    if (!h->timing) return CPML_OK;
    const size_t q = (h->ev_head + h->ev_count - 1) % kEventSlots;
    CK(cudaEventRecord(h->ev[2 * q + 1], h->stream));
    return CPML_OK;
}

static int32_t check_it(cpml_handle *h, int it)
{
    if (it < 1 || it > h->cfg.nstep) FAIL(CPML_EINVAL, "time step outside 1..NSTEP");
    return CPML_OK;
}

// One half step: [wait for the neighbours' planes] kernel [publish my planes].
// phase 0 = stress update (needs the velocity planes of step it-1, produces sigma planes of
// step it), phase 1 = velocity update (needs the sigma planes of step it, produces velocity
// planes of step it).  Flag words of a slab: [0] v from lo, [1] v from hi, [2] sigma from lo,
// [3] sigma from hi; a slab writes into its lo neighbour's "from hi" word and vice versa.
static int32_t half_step(cpml_handle *h, int32_t it, int phase)
{
    int32_t rc = check_it(h, it); if (rc) return rc;
    rc = finalize(h); if (rc) return rc;
    CK(cudaSetDevice(h->device));
    const bool peers = h->cfg.ndim == 3 && (h->peer_on[0] || h->peer_on[1]);
    SlabSync ss{};
    if (peers && h->use_ws && !h->visco) {
        // the kernels order the slabs themselves: boundary items poll flag words [w], [w+1] of this slab and the
        // last boundary item of a side publishes into the neighbour's words (same words as k_wait / k_signal)
        const int w = phase == 0 ? 0 : 2, wp = phase == 0 ? 2 : 0;
        const bool nothing_to_wait_for = phase == 0 && it == 1;      // step 1 reads the zero halo planes of the reset
        if (!nothing_to_wait_for) {
            ss.wait_lo = h->peer_on[0] ? h->flags + w : nullptr;
            ss.wait_hi = h->peer_on[1] ? h->flags + w + 1 : nullptr;
        }
        ss.wait_value = (h->epoch << 32) | (unsigned long long)(phase == 0 ? it - 1 : it);
        ss.pub_lo = h->peer_on[0] ? peer_flags(h, 0) + wp + 1 : nullptr;
        ss.pub_hi = h->peer_on[1] ? peer_flags(h, 1) + wp : nullptr;
        ss.pub_value = (h->epoch << 32) | (unsigned long long)it;
        ss.count = h->d_bcount;
        ss.n_boundary = (phase == 0 ? h->tile_stress : h->tile).ntx * (phase == 0 ? h->tile_stress : h->tile).nty;
        ss.timeout = h->d_timeout;
    } else if (peers) {
        if (!h->use_tma && !h->visco) FAIL(CPML_ESTATE, "peer stores need the TMA kernels (unset CPML_KERNEL=reg)");
        const int w = phase == 0 ? 0 : 2;
        // the stress update of step 1 reads the zero halo planes of the reset: nothing to wait for
        if (!(phase == 0 && it == 1)) {
            launch_wait(h->peer_on[0] ? h->flags + w : nullptr, h->peer_on[1] ? h->flags + w + 1 : nullptr,
                        (h->epoch << 32) | (unsigned long long)(phase == 0 ? it - 1 : it), h->d_timeout, h->stream);
            h->n_launches++;
        }
    }
    rc = time_begin(h, phase); if (rc) return rc;
    if (h->visco) {
        const ParamsV3D p = make_pv(h, it);
        if (phase == 0) launch_vstress3d(p, h->vgrid, h->stream);
        else if (h->v_ws) CK(launch_vvelocity3d_ws(p, h->vmaps, h->vtile, h->stream));
        else launch_vvelocity3d(p, h->vgrid, h->stream);
        h->n_launches += phase == 0 ? visco_stress_launches() : 1;
    } else if (h->cfg.ndim == 3) {
        const Params3D p = make_p3(h, it);
        if (h->f32) {
            const Params3DF pf = make_p3f(h, it);
            if (phase == 0) CK(launch_stress3d_ws(pf, h->maps_stress, h->tile_stress, ss, h->stream));
            else CK(launch_velocity3d_ws(pf, h->maps_velocity, h->tile, ss, h->stream));
            h->n_launches++;
        } else if (h->use_ws) {
            if (phase == 0) CK(launch_stress3d_ws(p, h->maps_stress, h->tile_stress, ss, h->stream));
            else CK(launch_velocity3d_ws(p, h->maps_velocity, h->tile, ss, h->stream));
            h->n_launches++;
        } else if (h->use_tma) {
            if (phase == 0) CK(launch_stress3d_tma(p, h->maps_stress, h->tile_stress, h->stream));
            else CK(launch_velocity3d_tma(p, h->maps_velocity, h->tile, h->stream));
            h->n_launches++;
        } else {
            for (const Box3D &b : h->regions) {
                if (phase == 0) launch_stress3d(p, b, h->stream); else launch_velocity3d(p, b, h->stream);
                h->n_launches++;
            }
        }
    } else if (h->visco2d) {
        if (phase == 0) launch_vstress2d(make_p2(h, it), h->grid, h->stream);
        else launch_vvelocity2d(make_p2(h, it), h->grid, h->stream);
        h->n_launches++;
    } else if (h->ws2) {
        if (phase == 0) CK(launch_stress2d_ws(make_p2(h, it), h->maps2_stress, h->tile2, h->stream));
        else CK(launch_velocity2d_ws(make_p2(h, it), h->maps2_velocity, h->tile2, h->stream));
        h->n_launches++;
    } else {
        if (phase == 0) launch_stress2d(make_p2(h, it), h->grid, h->block, h->stream);
        else launch_velocity2d(make_p2(h, it), h->grid, h->block, h->stream);
        h->n_launches++;
    }
    rc = time_end(h); if (rc) return rc;
    if (peers && (!h->use_ws || h->visco)) {
        const int w = phase == 0 ? 2 : 0;
        launch_signal(h->peer_on[0] ? peer_flags(h, 0) + w + 1 : nullptr, h->peer_on[1] ? peer_flags(h, 1) + w : nullptr,
                      (h->epoch << 32) | (unsigned long long)it, h->stream);
        h->n_launches++;
    }
    CK(cudaGetLastError());
    return CPML_OK;
}

extern "C" int32_t cpml_step_stress(cpml_handle *h, int32_t it)
{
    if (!h) return CPML_EINVAL;
    return half_step(h, it, 0);
}

extern "C" int32_t cpml_step_velocity(cpml_handle *h, int32_t it)
{
    if (!h) return CPML_EINVAL;
    return half_step(h, it, 1);
}

extern "C" int32_t cpml_step_finish(cpml_handle *h, int32_t it)
{
    if (!h) return CPML_EINVAL;
    int32_t rc = check_it(h, it); if (rc) return rc;
    rc = finalize(h); if (rc) return rc;
    CK(cudaSetDevice(h->device));
    const cpml_config &c = h->cfg;
    Post3D p{};
    p.partials = h->d_partials; p.nblocks = h->nblocks;
    p.npot = h->visco ? 2 * h->nblocks : h->nblocks;
    p.energy_k = h->d_ek; p.energy_p = h->d_ep;
    p.step_out = h->d_step_out + 4 * (size_t)(it - 1);
    p.it = it; p.nstep = c.nstep; p.nrec = c.nrec;
    p.ix_rec = h->d_ix_rec; p.iy_rec = h->d_iy_rec;
    p.vx = h->f0[0]; p.vy = h->f0[1];
    p.pitch = h->pitch;
    if (c.ndim == 3) {
        p.plane = h->plane;
        const int kl = h->ksrc_global - h->koff;
        p.krec = (kl >= 1 && kl <= h->nzl) ? kl : 0;
    } else {
        p.plane = 0;
        p.krec = 1;
    }
    p.sisvx = h->d_sisvx; p.sisvy = h->d_sisvy;
    p.vz = c.ndim == 3 ? h->f0[2] : nullptr; p.sisvz = h->d_sisvz;
    if (h->f32) {       // the float views start at the same ELEMENT offset of each field slot
        p.f32 = 1;
        p.vx = (const double *)((float *)h->field_alloc[0] + h->origin);
        p.vy = (const double *)((float *)h->field_alloc[1] + h->origin);
        p.vz = (const double *)((float *)h->field_alloc[2] + h->origin);
    }
    if (h->visco2d) {
        const Params2D p2 = make_p2(h, it);
        if (c.compute_energy) { launch_venergy2d(p2, h->grid, h->stream); h->n_launches++; }   // COMPUTE_ENERGY, :1037
        launch_vpressure2d(p2, h->d_ix_rec, h->d_iy_rec, c.nrec, c.nstep, h->d_sisp, h->stream);
        h->n_launches += c.nrec > 0 ? 1 : 0;
    }
    launch_post3d(p, h->stream);
    h->n_launches++;
    CK(cudaGetLastError());
    return CPML_OK;
}

// Waits for the handle's stream; with peer stores also checks that no wait kernel gave up on a neighbour
// (results computed from stale halo planes must not be handed out without an error).
static int32_t sync_checked(cpml_handle *h)
{
    CK(cudaStreamSynchronize(h->stream));
    if (h->peer_on[0] || h->peer_on[1]) {
        unsigned int timed_out = 0;
        CK(cudaMemcpy(&timed_out, h->d_timeout, sizeof(timed_out), cudaMemcpyDeviceToHost));
        if (timed_out) FAIL(CPML_ESTATE, "a neighbour slab never published its boundary planes (wait kernel timed out)");
    }
    return CPML_OK;
}

extern "C" int32_t cpml_synchronize(cpml_handle *h)
{
    if (!h) return CPML_EINVAL;
    CK(cudaSetDevice(h->device));
    return sync_checked(h);
}

extern "C" int32_t cpml_run(cpml_handle *h, int32_t it_begin, int32_t it_end)
{
    if (!h) return CPML_EINVAL;
    if (h->cfg.nslabs != 1) FAIL(CPML_ESTATE, "cpml_run needs the whole grid on one device; slab drivers interleave cpml_step_* with their plane exchange");
    if (it_begin < 1 || it_end > h->cfg.nstep || it_begin > it_end) FAIL(CPML_EINVAL, "bad time step range");
    for (int it = it_begin; it <= it_end; it++) {
        int32_t rc = cpml_step_stress(h, it); if (rc) return rc;
        rc = cpml_step_velocity(h, it); if (rc) return rc;
        rc = cpml_step_finish(h, it); if (rc) return rc;
    }
    return cpml_synchronize(h);
}

extern "C" int32_t cpml_halo_plane(cpml_handle *h, int32_t field, int32_t klocal, void **device_ptr, int64_t *nbytes)
{
    if (!h) return CPML_EINVAL;
    if (h->cfg.ndim != 3) FAIL(CPML_EINVAL, "halo planes exist only in 3-D");
    if (h->f32) FAIL(CPML_EINVAL, "single precision runs on one GPU: no halo planes");
    const int hz = h->visco ? 2 : 1;
    if (field < 0 || field >= 9 || klocal < 1 - hz || klocal > h->nzl + hz || !device_ptr || !nbytes) FAIL(CPML_EINVAL, "bad halo plane request");
    if (h->visco) {
        // the whole padded plane including its ghost rows: start of the row block of plane klocal
        *device_ptr = (void *)(h->field_alloc[field] + (long long)(klocal + 1) * h->plane);
    } else {
        *device_ptr = (void *)(h->f0[field] + (long long)klocal * h->plane);
    }
    *nbytes = (int64_t)h->plane * (int64_t)sizeof(double);
    return CPML_OK;
}

extern "C" int32_t cpml_copy_plane(cpml_handle *dst, int32_t klocal_dst, cpml_handle *src, int32_t klocal_src, int32_t field)
{
    if (!dst || !src) return CPML_EINVAL;
    cpml_handle *h = dst;
    if (dst->cfg.ndim != 3 || src->cfg.ndim != 3) FAIL(CPML_EINVAL, "plane copies exist only in 3-D");
    if (dst->visco != src->visco) FAIL(CPML_EINVAL, "slabs of different solvers");
    if (dst->visco) {
        if (field < 0 || field >= 9 || klocal_dst < -1 || klocal_dst > dst->nzl + 2 || klocal_src < -1 || klocal_src > src->nzl + 2 ||
            dst->plane != src->plane)
            FAIL(CPML_EINVAL, "bad plane copy request");
        CK(cudaSetDevice(src->device));
        CK(cudaStreamSynchronize(src->stream));
        CK(cudaSetDevice(dst->device));
        CK(cudaMemcpyPeerAsync(dst->field_alloc[field] + (long long)(klocal_dst + 1) * dst->plane, dst->device,
                               src->field_alloc[field] + (long long)(klocal_src + 1) * src->plane, src->device,
                               (size_t)dst->plane * sizeof(double), dst->stream));
        return CPML_OK;
    }
    if (dst->plane != src->plane || dst->cfg.nx != src->cfg.nx || dst->cfg.ny != src->cfg.ny) FAIL(CPML_EINVAL, "slabs of different grids");
    if (field < 0 || field >= 9 || klocal_dst < 0 || klocal_dst > dst->nzl + 1 || klocal_src < 0 || klocal_src > src->nzl + 1)
        FAIL(CPML_EINVAL, "bad plane copy request");
    CK(cudaSetDevice(src->device));
    CK(cudaStreamSynchronize(src->stream));
    CK(cudaSetDevice(dst->device));
    CK(cudaMemcpyPeerAsync(dst->f0[field] + (long long)klocal_dst * dst->plane, dst->device,
                           src->f0[field] + (long long)klocal_src * src->plane, src->device,
                           (size_t)dst->plane * sizeof(double), dst->stream));
    return CPML_OK;
}

// ---- direct slab-to-slab stores (replaces MPI_SENDRECV, 3D-iso :811-823, :951-963) ------

extern "C" int32_t cpml_p2p_export(cpml_handle *h, void *blob, int64_t blob_capacity, int64_t *nbytes)
{
    if (!h || !blob || !nbytes) return CPML_EINVAL;
    if (h->cfg.ndim != 3) FAIL(CPML_EINVAL, "slabs exist only in 3-D");
    if (blob_capacity < (int64_t)sizeof(cudaIpcMemHandle_t)) FAIL(CPML_EINVAL, "blob too small (need 64 bytes)");
    CK(cudaSetDevice(h->device));
    cudaIpcMemHandle_t mh;
    CK(cudaIpcGetMemHandle(&mh, h->arena));
    memcpy(blob, &mh, sizeof(mh));
    *nbytes = (int64_t)sizeof(mh);
    return CPML_OK;
}

static int32_t check_side(cpml_handle *h, int32_t side)
{
    if (h->cfg.ndim != 3) FAIL(CPML_EINVAL, "slabs exist only in 3-D");
    if (side != 0 && side != 1) FAIL(CPML_EINVAL, "side must be 0 (slab rank-1) or 1 (slab rank+1)");
    if ((side == 0 && h->cfg.slab_rank == 0) || (side == 1 && h->cfg.slab_rank == h->cfg.nslabs - 1))
        FAIL(CPML_ETOPOLOGY, "no neighbour on that side (MPI_PROC_NULL, 3D-iso :775-790)");
    if (h->peer_on[side]) FAIL(CPML_ESTATE, "neighbour already attached");
    return CPML_OK;
}

extern "C" int32_t cpml_p2p_attach_ipc(cpml_handle *h, int32_t side, const void *blob, int64_t nbytes)
{
    if (!h || !blob) return CPML_EINVAL;
    int32_t rc = check_side(h, side); if (rc) return rc;
    if (nbytes != (int64_t)sizeof(cudaIpcMemHandle_t)) FAIL(CPML_EINVAL, "bad blob size");
    CK(cudaSetDevice(h->device));
    cudaIpcMemHandle_t mh;
    memcpy(&mh, blob, sizeof(mh));
    void *ptr = nullptr;
    CK(cudaIpcOpenMemHandle(&ptr, mh, cudaIpcMemLazyEnablePeerAccess));
    h->peer_arena[side] = (double *)ptr;
    h->peer_ipc[side] = true;
    h->peer_on[side] = true;
    return CPML_OK;
}

extern "C" int32_t cpml_p2p_attach_local(cpml_handle *h, int32_t side, cpml_handle *neighbour)
{
    if (!h || !neighbour) return CPML_EINVAL;
    int32_t rc = check_side(h, side); if (rc) return rc;
    const cpml_config &a = h->cfg, &b = neighbour->cfg;
    if (b.ndim != 3 || a.nx != b.nx || a.ny != b.ny || a.nz != b.nz || a.nslabs != b.nslabs ||
        b.slab_rank != a.slab_rank + (side == 0 ? -1 : 1))
        FAIL(CPML_ETOPOLOGY, "that handle is not the neighbouring slab of the same grid");
    CK(cudaSetDevice(h->device));
    if (neighbour->device != h->device) {
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, h->device, neighbour->device));
        if (!can) FAIL(CPML_ECUDA, "no peer access between the two devices");
        const cudaError_t e = cudaDeviceEnablePeerAccess(neighbour->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
        cudaGetLastError();
    }
    h->peer_arena[side] = neighbour->arena;
    h->peer_ipc[side] = false;
    h->peer_on[side] = true;
    return CPML_OK;
}

extern "C" int32_t cpml_p2p_detach(cpml_handle *h)
{
    if (!h) return CPML_EINVAL;
    cudaSetDevice(h->device);
    for (int side = 0; side < 2; side++) {
        if (h->peer_on[side] && h->peer_ipc[side]) {
            cudaStreamSynchronize(h->stream);
            cudaIpcCloseMemHandle(h->peer_arena[side]);
        }
        h->peer_on[side] = h->peer_ipc[side] = false;
        h->peer_arena[side] = nullptr;
    }
    return CPML_OK;
}

extern "C" int32_t cpml_get_launch_info(cpml_handle *h, int32_t *info, int32_t n)
{
    if (!h || !info) return CPML_EINVAL;
    int32_t rc = finalize(h); if (rc) return rc;
    if (h->visco) {
        const int32_t w[10] = {h->v_ws ? 2 : 0, h->vtx, h->vty, 0, h->vkchunk, (int32_t)h->vgrid.z, h->nblocks, h->nblocks, h->nblocks,
                               (h->peer_on[0] ? 1 : 0) + (h->peer_on[1] ? 2 : 0)};
        for (int q = 0; q < n && q < 10; q++) info[q] = w[q];
        return CPML_OK;
    }
    const Tile3D &t = h->tile, &ts = h->tile_stress;
    const int32_t v[14] = {h->use_ws ? 2 : h->use_tma ? 1 : 0, t.tx, t.ty, t.stages, t.kchunk, t.nzc, t.nitems, ts.grid_stress, t.grid_velocity,
                           (h->peer_on[0] ? 1 : 0) + (h->peer_on[1] ? 2 : 0), ts.ty, ts.kchunk, ts.nzc, ts.nitems};
    for (int q = 0; q < n && q < 14; q++) info[q] = v[q];
    return CPML_OK;
}

extern "C" int32_t cpml_get_seismograms(cpml_handle *h, double *sisvx, double *sisvy)
{
    if (!h || !sisvx || !sisvy) return CPML_EINVAL;
    CK(cudaSetDevice(h->device));
    const size_t n = (size_t)h->cfg.nstep * (size_t)h->cfg.nrec;
    if (n == 0) return CPML_OK;
    CK(cudaMemcpyAsync(sisvx, h->d_sisvx, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(sisvy, h->d_sisvy, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    { const int32_t rc_sync = sync_checked(h); if (rc_sync) return rc_sync; }
    return CPML_OK;
}

extern "C" int32_t cpml_get_seismograms_vz(cpml_handle *h, double *sisvz)
{
    if (!h || !sisvz) return CPML_EINVAL;
    if (h->cfg.ndim != 3) FAIL(CPML_EINVAL, "Vz seismograms exist in the 3-D programs only");
    CK(cudaSetDevice(h->device));
    const size_t n = (size_t)h->cfg.nstep * (size_t)h->cfg.nrec;
    if (n == 0) return CPML_OK;
    CK(cudaMemcpyAsync(sisvz, h->d_sisvz, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    { const int32_t rc_sync = sync_checked(h); if (rc_sync) return rc_sync; }
    return CPML_OK;
}

extern "C" int32_t cpml_get_pressure_seismograms(cpml_handle *h, double *sispressure)
{
    if (!h || !sispressure) return CPML_EINVAL;
    if (!h->visco2d) FAIL(CPML_EINVAL, "pressure seismograms exist in the 2-D viscoelastic programs only");
    CK(cudaSetDevice(h->device));
    const size_t n = (size_t)h->cfg.nstep * (size_t)h->cfg.nrec;
    if (n == 0) return CPML_OK;
    CK(cudaMemcpyAsync(sispressure, h->d_sisp, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    { const int32_t rc_sync = sync_checked(h); if (rc_sync) return rc_sync; }
    return CPML_OK;
}

extern "C" int32_t cpml_get_energy(cpml_handle *h, double *total, double *kinetic, double *potential)
{
    if (!h) return CPML_EINVAL;
    CK(cudaSetDevice(h->device));
    const size_t n = (size_t)h->cfg.nstep;
    std::vector<double> ek(n), ep(n);
    CK(cudaMemcpyAsync(ek.data(), h->d_ek, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(ep.data(), h->d_ep, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    { const int32_t rc_sync = sync_checked(h); if (rc_sync) return rc_sync; }
    for (size_t q = 0; q < n; q++) {
        if (total) total[q] = ek[q] + ep[q];       // :1179 sums kinetic + potential
        if (kinetic) kinetic[q] = ek[q];
        if (potential) potential[q] = ep[q];
    }
    return CPML_OK;
}

// Device pointer to plane `off` (element offset from element (1,1,0)) of a field as DOUBLE values: the field itself, or --
// single precision -- the scratch plane the values were just converted into (on the handle's stream).
static const double *plane_as_double(cpml_handle *h, int field, long long off)
{
    if (!h->f32) return h->f0[field] + off;
    launch_f2d((const float *)h->field_alloc[field] + h->origin + off, h->d_scratch, (long long)h->pitch * h->cfg.ny, h->stream);
    return h->d_scratch;
}

extern "C" int32_t cpml_get_plane(cpml_handle *h, int32_t field, int32_t kglobal, double *out)
{
    if (!h || !out) return CPML_EINVAL;
    const cpml_config &c = h->cfg;
    if (field < 0 || field >= h->nfields) FAIL(CPML_EINVAL, "bad field id");
    CK(cudaSetDevice(h->device));
    long long off = 0;
    if (c.ndim == 3) {
        const int kl = kglobal - h->koff;
        if (kl < 1 || kl > h->nzl) FAIL(CPML_EINVAL, "this slab does not hold that plane");
        off = (long long)kl * h->plane;
    }
    CK(cudaMemcpy2DAsync(out, (size_t)c.nx * sizeof(double), plane_as_double(h, field, off), (size_t)h->pitch * sizeof(double),
                         (size_t)c.nx * sizeof(double), c.ny, cudaMemcpyDeviceToHost, h->stream));
    { const int32_t rc_sync = sync_checked(h); if (rc_sync) return rc_sync; }
    return CPML_OK;
}

// ---- asynchronous snapshot planes ------------------------------------------------------
// The display phase of the reference hands vx(:,:,NZ_LOCAL) and vy(:,:,NZ_LOCAL) to create_color_image every IT_DISPLAY
// steps (3D-iso :1236-1239; the 2-D programs their whole 134 MB fields at 4096 x 4096).  cpml_get_plane copies
// synchronously into the caller's pageable buffer and stalls the time loop for the whole transfer;
// cpml_snapshot_begin instead copies the plane device-side (a dense NX x NY image, so later time steps cannot touch it)
// and starts the device-to-host transfer into PINNED memory on a side stream: the loop goes on, and
// cpml_snapshot_end collects the plane whenever the driver wants it (typically at the next display step).

extern "C" int32_t cpml_snapshot_begin(cpml_handle *h, int32_t slot, int32_t field, int32_t kglobal)
{
    if (!h) return CPML_EINVAL;
    const cpml_config &c = h->cfg;
    if (slot < 0 || slot > 1) FAIL(CPML_EINVAL, "snapshot slot must be 0 or 1");
    if (field < 0 || field >= h->nfields) FAIL(CPML_EINVAL, "bad field id");
    if (h->snap_pending[slot]) FAIL(CPML_ESTATE, "snapshot slot still holds a plane that was not collected (cpml_snapshot_end)");
    CK(cudaSetDevice(h->device));
    long long off = 0;
    if (c.ndim == 3) {
        const int kl = kglobal - h->koff;
        if (kl < 1 || kl > h->nzl) FAIL(CPML_EINVAL, "this slab does not hold that plane");
        off = (long long)kl * h->plane;
    }
    const size_t bytes = (size_t)c.nx * c.ny * sizeof(double);
    if (!h->snap_stream) CK(cudaStreamCreateWithFlags(&h->snap_stream, cudaStreamNonBlocking));
    if (!h->snap_dev[slot]) {
        CK(cudaMalloc(&h->snap_dev[slot], bytes));
        CK(cudaMallocHost(&h->snap_pin[slot], bytes));
        CK(cudaEventCreateWithFlags(&h->snap_ready[slot], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->snap_done[slot], cudaEventDisableTiming));
        CK(cudaEventRecord(h->snap_done[slot], h->snap_stream));
    }
    CK(cudaStreamWaitEvent(h->stream, h->snap_done[slot], 0));          // the previous transfer out of snap_dev is over
    CK(cudaMemcpy2DAsync(h->snap_dev[slot], (size_t)c.nx * sizeof(double), plane_as_double(h, field, off), (size_t)h->pitch * sizeof(double),
                         (size_t)c.nx * sizeof(double), c.ny, cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaEventRecord(h->snap_ready[slot], h->stream));
    CK(cudaStreamWaitEvent(h->snap_stream, h->snap_ready[slot], 0));
    CK(cudaMemcpyAsync(h->snap_pin[slot], h->snap_dev[slot], bytes, cudaMemcpyDeviceToHost, h->snap_stream));
    CK(cudaEventRecord(h->snap_done[slot], h->snap_stream));
    h->snap_pending[slot] = true;
    return CPML_OK;
}

extern "C" int32_t cpml_snapshot_end(cpml_handle *h, int32_t slot, double *out, const double **pinned)
{
    if (!h) return CPML_EINVAL;
    if (slot < 0 || slot > 1) FAIL(CPML_EINVAL, "snapshot slot must be 0 or 1");
    if (!h->snap_pending[slot]) FAIL(CPML_ESTATE, "no snapshot was started in this slot (cpml_snapshot_begin)");
    CK(cudaSetDevice(h->device));
    CK(cudaEventSynchronize(h->snap_done[slot]));
    if (out) memcpy(out, h->snap_pin[slot], (size_t)h->cfg.nx * h->cfg.ny * sizeof(double));
    if (pinned) *pinned = h->snap_pin[slot];        // valid until the next cpml_snapshot_begin on this slot
    h->snap_pending[slot] = false;
    return CPML_OK;
}

extern "C" int32_t cpml_get_field(cpml_handle *h, int32_t field, double *out)
{
    if (!h || !out) return CPML_EINVAL;
    const cpml_config &c = h->cfg;
    if (field < 0 || field >= h->nfields) FAIL(CPML_EINVAL, "bad field id");
    CK(cudaSetDevice(h->device));
    if (c.ndim == 2) return cpml_get_plane(h, field, 0, out);
    if (h->visco) {   // ghost rows between the planes: one 2-D copy per plane
        for (int k = 1; k <= h->nzl; k++)
            CK(cudaMemcpy2DAsync(out + (size_t)(k - 1) * c.nx * c.ny, (size_t)c.nx * sizeof(double),
                                 h->f0[field] + (long long)k * h->plane, (size_t)h->pitch * sizeof(double),
                                 (size_t)c.nx * sizeof(double), c.ny, cudaMemcpyDeviceToHost, h->stream));
        { const int32_t rc_sync = sync_checked(h); if (rc_sync) return rc_sync; }
        return CPML_OK;
    }
    if (h->f32) {     // plane by plane through the conversion scratch
        for (int k = 1; k <= h->nzl; k++)
            CK(cudaMemcpy2DAsync(out + (size_t)(k - 1) * c.nx * c.ny, (size_t)c.nx * sizeof(double),
                                 plane_as_double(h, field, (long long)k * h->plane), (size_t)h->pitch * sizeof(double),
                                 (size_t)c.nx * sizeof(double), c.ny, cudaMemcpyDeviceToHost, h->stream));
        { const int32_t rc_sync = sync_checked(h); if (rc_sync) return rc_sync; }
        return CPML_OK;
    }
    // rows of all owned planes are equally spaced (plane = pitch * ny): one 2-D copy
    CK(cudaMemcpy2DAsync(out, (size_t)c.nx * sizeof(double), h->f0[field] + h->plane, (size_t)h->pitch * sizeof(double),
                         (size_t)c.nx * sizeof(double), (size_t)c.ny * h->nzl, cudaMemcpyDeviceToHost, h->stream));
    { const int32_t rc_sync = sync_checked(h); if (rc_sync) return rc_sync; }
    return CPML_OK;
}

extern "C" int32_t cpml_get_maxnorm(cpml_handle *h, double *out)
{
    if (!h || !out) return CPML_EINVAL;
    const cpml_config &c = h->cfg;
    CK(cudaSetDevice(h->device));
    CK(cudaMemsetAsync(h->d_maxbits, 0, sizeof(unsigned long long), h->stream));
    if (h->f32)
        launch_maxnorm_f((const float *)h->field_alloc[0] + h->origin + h->plane, (const float *)h->field_alloc[1] + h->origin + h->plane,
                         (const float *)h->field_alloc[2] + h->origin + h->plane, h->plane * h->nzl, h->d_maxbits, h->stream);
    else if (c.ndim == 3)
        launch_maxnorm(h->f0[0] + h->plane, h->f0[1] + h->plane, h->f0[2] + h->plane, h->plane * h->nzl, h->d_maxbits, h->stream);
    else
        launch_maxnorm(h->field_alloc[0], h->field_alloc[1], nullptr, (long long)h->field_doubles, h->d_maxbits, h->stream);
    h->n_launches++;
    unsigned long long bits = 0;
    CK(cudaMemcpyAsync(&bits, h->d_maxbits, sizeof(bits), cudaMemcpyDeviceToHost, h->stream));
    { const int32_t rc_sync = sync_checked(h); if (rc_sync) return rc_sync; }
    memcpy(out, &bits, sizeof(double));
    return CPML_OK;
}

extern "C" int32_t cpml_enable_kernel_timing(cpml_handle *h, int32_t on)
{
    if (!h) return CPML_EINVAL;
    if (on && h->ev.empty()) {      // the event pool is created here, outside any timed loop
        CK(cudaSetDevice(h->device));
        h->ev.resize(2 * kEventSlots);
        h->ev_kind.assign(kEventSlots, -1);
        for (auto &e : h->ev) CK(cudaEventCreate(&e));
    }
    h->timing = on != 0;
    return CPML_OK;
}

extern "C" int32_t cpml_get_kernel_times(cpml_handle *h, double *ms_stress, double *ms_velocity, int64_t *n_launches, int32_t reset)
{
    if (!h) return CPML_EINVAL;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    while (h->ev_count > 0) { const int32_t rc = harvest_oldest(h); if (rc) return rc; }
    if (ms_stress) *ms_stress = h->ms_stress;
    if (ms_velocity) *ms_velocity = h->ms_velocity;
    if (n_launches) *n_launches = h->n_launches;
    if (reset) { h->ms_stress = h->ms_velocity = 0; h->n_launches = 0; }
    return CPML_OK;
}

extern "C" int32_t cpml_algorithmic_bytes(cpml_handle *h, double *bytes_stress, double *bytes_velocity)
{
    if (!h) return CPML_EINVAL;
    int32_t rc = finalize(h); if (rc) return rc;
    const cpml_config &c = h->cfg;
    const double N = (double)c.nx * c.ny * h->nzl;
    const int (*nz)[2] = h->nz_own;
    double ws, wv;   // words
    if (h->visco) {
        // stress: read vx,vy,vz; read+write 6 sigma, 6 sigma_R, 12 strain memory variables -> 3 + 2*24;
        // velocity: read 6 sigma; read+write vx,vy,vz -> 6 + 2*3 (the energy is fused: potential part in
        // the stress kernel, kinetic part in the velocity kernel); C-PML memory variables as isotropic
        const double px = (double)c.ny * h->nzl, py = (double)c.nx * h->nzl, pz = (double)c.nx * c.ny;
        ws = 51.0 * N + 2.0 * (px * (nz[0][1] + 2 * nz[0][0]) + py * (nz[1][0] + 2 * nz[1][1]) + pz * (nz[2][0] + 2 * nz[2][1]));
        wv = 12.0 * N + 2.0 * (px * (nz[0][0] + 2 * nz[0][1]) + py * (2 * nz[1][0] + nz[1][1]) + pz * (2 * nz[2][0] + nz[2][1]));
    } else if (c.ndim == 3) {
        // stress: read vx,vy,vz; read+write 6 sigma; memory variables per direction:
        //   x: dvx_dx (half), dvy_dx, dvz_dx (integer)   y: dvy_dy (int), dvx_dy, dvz_dy (half)
        //   z: dvz_dz (int), dvx_dz, dvy_dz (half)        -- each read + written
        const double px = (double)c.ny * h->nzl, py = (double)c.nx * h->nzl, pz = (double)c.nx * c.ny;
        ws = 15.0 * N + 2.0 * (px * (nz[0][1] + 2 * nz[0][0]) + py * (nz[1][0] + 2 * nz[1][1]) + pz * (nz[2][0] + 2 * nz[2][1]));
        // velocity: read 6 sigma; read+write vx,vy,vz; memory variables:
        //   x: dsxx_dx (int), dsxy_dx, dsxz_dx (half)    y: dsxy_dy, dsyz_dy (int), dsyy_dy (half)
        //   z: dsxz_dz, dsyz_dz (int), dszz_dz (half)
        wv = 12.0 * N + 2.0 * (px * (nz[0][0] + 2 * nz[0][1]) + py * (2 * nz[1][0] + nz[1][1]) + pz * (2 * nz[2][0] + nz[2][1]));
    } else if (h->visco2d) {
        // stress: read vx,vy, lambda, mu; read+write 3 sigma and the 9 memory variables; velocity: read 3 sigma, rho; read+write vx,vy
        const double px = (double)c.ny, py = (double)c.nx;
        ws = 28.0 * N + 2.0 * (px * (nz[0][0] + nz[0][1]) + py * (nz[1][0] + nz[1][1]));
        wv = 8.0 * N + 2.0 * (px * (nz[0][0] + nz[0][1]) + py * (nz[1][0] + nz[1][1]));
    } else {
        // stress: read vx,vy, lambda, mu; read+write 3 sigma; x: dvx_dx (half), dvy_dx (int); y: dvy_dy (int), dvx_dy (half)
        const double px = (double)c.ny, py = (double)c.nx;
        ws = 10.0 * N + 2.0 * (px * (nz[0][0] + nz[0][1]) + py * (nz[1][0] + nz[1][1]));
        // velocity: read 3 sigma, rho; read+write vx,vy (the potential energy is summed in the stress kernel, which
        // already holds lambda and mu: 18 words per point-update)
        wv = 8.0 * N + 2.0 * (px * (nz[0][0] + nz[0][1]) + py * (nz[1][0] + nz[1][1]));
    }
    const double es = h->f32 ? 4.0 : 8.0;       // bytes per word
    if (bytes_stress) *bytes_stress = es * ws;
    if (bytes_velocity) *bytes_velocity = es * wv;
    return CPML_OK;
}
