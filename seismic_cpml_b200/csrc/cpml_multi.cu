// cpml_multi_*: the whole z-slab decomposition behind ONE handle, driven from ONE host thread.
//
// The reference distributes its 3-D grids over MPI ranks from inside the Fortran program
// (seismic_CPML_3D_isotropic_MPI_OpenMP.f90:337-346 rank set-up, :770-796 neighbours, :811-823 / :951-963 plane
// exchange, :1179 energy reduction; seismic_CPML_3D_viscoelastic_MPI.f90:384-390, :922-948, :962-975, :1229-1242).
// A Fortran / C / C++ host that has no MPI (none exists on the B200 box) uses this instead: N slab handles on N
// devices of one node, neighbours attached to each other (cpml_p2p_attach_local: the update kernels store their
// boundary planes straight into the neighbour GPU's halo planes over NVLink and order themselves through
// device-side flags), the loop body launched on every device per time step, the energy summed and the seismograms
// taken from the slab that owns the cut plane -- what rank_cut_plane does in the reference.
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "cpml_internal.h"

struct cpml_multi {
    cpml_config cfg{};
    int n = 0;
    std::vector<cpml_handle *> slab;
    std::vector<int> device;
    std::vector<cudaStream_t> stream;        // one per DISTINCT device (slabs sharing a device share its stream)
    std::vector<cudaStream_t> owned;
    std::string err;
    int nzl = 0, ksrc = 0;
};

namespace {
thread_local std::string g_multi_create_error;

int32_t fail(cpml_multi *m, int32_t code, const std::string &msg)
{
    m->err = msg;
    return code;
}

// error of slab r -> error of the multi handle
int32_t fwd(cpml_multi *m, int r, int32_t rc)
{
    if (rc != CPML_OK) m->err = "slab " + std::to_string(r) + ": " + cpml_last_error(m->slab[r]);
    return rc;
}
}  // namespace

#define MFWD(r, call)                                  \
    do {                                               \
        const int32_t rc__ = fwd(m, (r), (call));      \
        if (rc__ != CPML_OK) return rc__;              \
    } while (0)

extern "C" const char *cpml_multi_last_error(const cpml_multi *m)
{
    return m ? m->err.c_str() : g_multi_create_error.c_str();
}

extern "C" int32_t cpml_multi_destroy(cpml_multi *m)
{
    if (!m) return CPML_OK;
    for (cpml_handle *h : m->slab)
        if (h) cpml_synchronize(h);
    for (cpml_handle *h : m->slab)
        if (h) cpml_p2p_detach(h);
    for (cpml_handle *h : m->slab) cpml_destroy(h);
    for (size_t q = 0; q < m->owned.size(); q++) cudaStreamDestroy(m->owned[q]);
    delete m;
    return CPML_OK;
}

extern "C" int32_t cpml_multi_create(const cpml_config *cfg, int32_t ngpus, const int32_t *devices, cpml_multi **out)
{
    if (out) *out = nullptr;
    if (!cfg || !out || ngpus < 1) { g_multi_create_error = "null argument or ngpus < 1"; return CPML_EINVAL; }
    if (cfg->ndim != 3) { g_multi_create_error = "only the 3-D solvers are decomposed into z-slabs"; return CPML_ETOPOLOGY; }
    cpml_multi *m = new (std::nothrow) cpml_multi();
    if (!m) { g_multi_create_error = "out of host memory"; return CPML_ENOMEM; }
    m->cfg = *cfg;
    m->n = ngpus;
    m->slab.assign(ngpus, nullptr);
    m->device.resize(ngpus);
    m->stream.assign(ngpus, nullptr);
    auto bail = [&](int32_t rc, const std::string &msg) {
        g_multi_create_error = msg;
        cpml_multi_destroy(m);
        return rc;
    };
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return bail(CPML_ECUDA, "no CUDA device: libcpml_b200 has no CPU fallback");
    for (int r = 0; r < ngpus; r++) {
        m->device[r] = devices ? devices[r] : r;
        if (m->device[r] < 0 || m->device[r] >= ndev) return bail(CPML_EINVAL, "device ordinal outside 0..deviceCount-1");
    }
    for (int r = 0; r < ngpus; r++) {
        cpml_config c = *cfg;
        c.nslabs = ngpus;
        c.slab_rank = r;
        c.device = m->device[r];
        const int32_t rc = cpml_create(&c, &m->slab[r]);
        if (rc != CPML_OK) return bail(rc, std::string("slab ") + std::to_string(r) + ": " + cpml_last_error(nullptr));
        // slabs on the same device share one stream: their kernels then run in launch order, which the
        // in-kernel slab ordering relies on when two slabs cannot be resident at the same time
        int first = r;
        for (int q = 0; q < r; q++) if (m->device[q] == m->device[r]) { first = q; break; }
        if (first == r) {
            if (cudaSetDevice(m->device[r]) != cudaSuccess || cudaStreamCreateWithFlags(&m->stream[r], cudaStreamNonBlocking) != cudaSuccess)
                return bail(CPML_ECUDA, "cudaStreamCreate failed");
            m->owned.push_back(m->stream[r]);
        } else m->stream[r] = m->stream[first];
        cpml_set_stream(m->slab[r], (void *)m->stream[r]);
    }
    m->nzl = cfg->nz / ngpus;
    m->ksrc = cfg->ksource == 0 ? cfg->nz / 2 : cfg->ksource;
    // neighbours, :770-796 (MPI_PROC_NULL at the two ends)
    for (int r = 0; r + 1 < ngpus; r++) {
        int32_t rc = cpml_p2p_attach_local(m->slab[r], 1, m->slab[r + 1]);
        if (rc == CPML_OK) rc = cpml_p2p_attach_local(m->slab[r + 1], 0, m->slab[r]);
        if (rc != CPML_OK)
            return bail(rc, std::string("attaching slabs ") + std::to_string(r) + " and " + std::to_string(r + 1) + ": " +
                                cpml_last_error(m->slab[r]) + " / " + cpml_last_error(m->slab[r + 1]));
    }
    *out = m;
    return CPML_OK;
}

extern "C" int32_t cpml_multi_ngpus(const cpml_multi *m) { return m ? m->n : 0; }

extern "C" int32_t cpml_multi_slab(cpml_multi *m, int32_t rank, cpml_handle **out)
{
    if (!m || !out) return CPML_EINVAL;
    if (rank < 0 || rank >= m->n) return fail(m, CPML_EINVAL, "slab rank outside 0..ngpus-1");
    *out = m->slab[rank];
    return CPML_OK;
}

extern "C" int32_t cpml_multi_synchronize(cpml_multi *m)
{
    if (!m) return CPML_EINVAL;
    for (int r = 0; r < m->n; r++) MFWD(r, cpml_synchronize(m->slab[r]));
    return CPML_OK;
}

extern "C" int32_t cpml_multi_reset(cpml_multi *m)
{
    if (!m) return CPML_EINVAL;
    // nobody resets while a neighbour may still be storing into its halo planes, nobody steps before every slab is reset
    int32_t rc = cpml_multi_synchronize(m);
    if (rc) return rc;
    for (int r = 0; r < m->n; r++) MFWD(r, cpml_reset(m->slab[r]));
    return cpml_multi_synchronize(m);
}

extern "C" int32_t cpml_multi_set_profiles(cpml_multi *m, int32_t axis, const double *a, const double *b, const double *K,
                                           const double *a_half, const double *b_half, const double *K_half, int32_t n)
{
    if (!m) return CPML_EINVAL;
    for (int r = 0; r < m->n; r++) MFWD(r, cpml_set_profiles(m->slab[r], axis, a, b, K, a_half, b_half, K_half, n));
    return CPML_OK;
}

extern "C" int32_t cpml_multi_set_attenuation(cpml_multi *m, int32_t n_sls, const double *tau_epsilon_nu1, const double *tau_sigma_nu1,
                                              const double *tau_epsilon_nu2, const double *tau_sigma_nu2)
{
    if (!m) return CPML_EINVAL;
    for (int r = 0; r < m->n; r++)
        MFWD(r, cpml_set_attenuation(m->slab[r], n_sls, tau_epsilon_nu1, tau_sigma_nu1, tau_epsilon_nu2, tau_sigma_nu2));
    return CPML_OK;
}

extern "C" int32_t cpml_multi_set_source_series(cpml_multi *m, const double *force_x, const double *force_y, int32_t n)
{
    if (!m) return CPML_EINVAL;
    for (int r = 0; r < m->n; r++) MFWD(r, cpml_set_source_series(m->slab[r], force_x, force_y, n));
    return CPML_OK;
}

extern "C" int32_t cpml_multi_set_receivers(cpml_multi *m, const int32_t *ix_rec, const int32_t *iy_rec, int32_t n)
{
    if (!m) return CPML_EINVAL;
    for (int r = 0; r < m->n; r++) MFWD(r, cpml_set_receivers(m->slab[r], ix_rec, iy_rec, n));
    return CPML_OK;
}

// One pass of the loop body (:804-1180) on every slab; asynchronous (nothing waits for the devices).
extern "C" int32_t cpml_multi_step(cpml_multi *m, int32_t it)
{
    if (!m) return CPML_EINVAL;
    for (int r = 0; r < m->n; r++) MFWD(r, cpml_step_stress(m->slab[r], it));
    for (int r = 0; r < m->n; r++) MFWD(r, cpml_step_velocity(m->slab[r], it));
    for (int r = 0; r < m->n; r++) MFWD(r, cpml_step_finish(m->slab[r], it));
    return CPML_OK;
}

extern "C" int32_t cpml_multi_run(cpml_multi *m, int32_t it_begin, int32_t it_end)
{
    if (!m) return CPML_EINVAL;
    if (it_begin < 1 || it_end > m->cfg.nstep || it_begin > it_end) return fail(m, CPML_EINVAL, "bad time step range");
    for (int it = it_begin; it <= it_end; it++) {
        const int32_t rc = cpml_multi_step(m, it);
        if (rc) return rc;
    }
    return cpml_multi_synchronize(m);
}

static int owner_of(const cpml_multi *m, int kglobal) { return (kglobal - 1) / m->nzl; }

extern "C" int32_t cpml_multi_get_seismograms(cpml_multi *m, double *sisvx, double *sisvy)
{
    if (!m) return CPML_EINVAL;
    const int r = owner_of(m, m->ksrc);      // rank_cut_plane, :346
    MFWD(r, cpml_get_seismograms(m->slab[r], sisvx, sisvy));
    return CPML_OK;
}

extern "C" int32_t cpml_multi_get_seismograms_vz(cpml_multi *m, double *sisvz)
{
    if (!m) return CPML_EINVAL;
    const int r = owner_of(m, m->ksrc);
    MFWD(r, cpml_get_seismograms_vz(m->slab[r], sisvz));
    return CPML_OK;
}

// MPI_REDUCE(SUM) of :1179 / 3D-visco :1425-1430, in rank order (deterministic)
extern "C" int32_t cpml_multi_get_energy(cpml_multi *m, double *total, double *kinetic, double *potential)
{
    if (!m) return CPML_EINVAL;
    const size_t n = (size_t)m->cfg.nstep;
    std::vector<double> ek(n, 0.0), ep(n, 0.0), k1(n), p1(n);
    for (int r = 0; r < m->n; r++) {
        MFWD(r, cpml_get_energy(m->slab[r], nullptr, k1.data(), p1.data()));
        for (size_t q = 0; q < n; q++) { ek[q] += k1[q]; ep[q] += p1[q]; }
    }
    for (size_t q = 0; q < n; q++) {
        if (total) total[q] = ek[q] + ep[q];
        if (kinetic) kinetic[q] = ek[q];
        if (potential) potential[q] = ep[q];
    }
    return CPML_OK;
}

extern "C" int32_t cpml_multi_get_plane(cpml_multi *m, int32_t field, int32_t kglobal, double *out)
{
    if (!m) return CPML_EINVAL;
    if (kglobal < 1 || kglobal > m->cfg.nz) return fail(m, CPML_EINVAL, "plane outside 1..NZ");
    const int r = owner_of(m, kglobal);
    MFWD(r, cpml_get_plane(m->slab[r], field, kglobal, out));
    return CPML_OK;
}

extern "C" int32_t cpml_multi_get_field(cpml_multi *m, int32_t field, double *out)
{
    if (!m || !out) return CPML_EINVAL;
    const size_t per = (size_t)m->cfg.nx * m->cfg.ny * m->nzl;
    for (int r = 0; r < m->n; r++) MFWD(r, cpml_get_field(m->slab[r], field, out + (size_t)r * per));
    return CPML_OK;
}

// MPI_REDUCE(MAX) of :1185-1186
extern "C" int32_t cpml_multi_get_maxnorm(cpml_multi *m, double *out)
{
    if (!m || !out) return CPML_EINVAL;
    double v = 0.0;
    for (int r = 0; r < m->n; r++) {
        double x = 0.0;
        MFWD(r, cpml_get_maxnorm(m->slab[r], &x));
        v = std::max(v, x);
    }
    *out = v;
    return CPML_OK;
}
