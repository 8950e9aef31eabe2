// Which global->shared staging mechanism sustains HBM bandwidth for the z-marching pattern?
// (run on the GPU box)  Same traffic as k_stress3d (3 arrays read, 6 read-modify-written), a CTA
// owns CH contiguous doubles of a plane and marches kchunk planes; the nine chunks of a plane go
// to a shared-memory ring D deep by
//   M=1  cp.async 16 B per thread (LDGSTS), commit/wait groups
//   M=2  one cp.async.bulk (TMA, UBLKCP) per array, mbarrier complete_tx
// then every thread reads its two doubles of each array from shared memory, updates, stores.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok = 0, spins = 0;
    while (!ok) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (++spins > (1u << 24)) __trap();
    }
}

template <int M, int CH>   // threads = CH / 2
__global__ void __launch_bounds__(CH / 2) staged(double *const *arr, long long plane, int nz, int kchunk, int D)
{
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int NT = CH / 2, STAGE = 9 * CH * 8;
    const uint32_t sb = smem_u32(smem);
    const uint32_t bar0 = sb, ring = sb + 128;
    const int tid = threadIdx.x;
    if (M == 2 && tid == 0) {
        for (int s = 0; s < D; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * s) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int ntile = (int)(plane / CH), nzc = (nz + kchunk - 1) / kchunk, nitems = ntile * nzc;
    uint32_t g = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int tile = item % ntile, zc = item / ntile;
        const int kb = zc * kchunk, np = min(nz, kb + kchunk) - kb;
        const long long off = (long long)tile * CH;
        auto issue = [&](int l) {
            const uint32_t s = (g + l) % D;
            const long long q = (long long)(kb + l) * plane + off;
            if (M == 1) {
#pragma unroll
                for (int a = 0; a < 9; a++)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring + s * STAGE + a * CH * 8 + tid * 16), "l"(arr[a] + q + 2 * tid) : "memory");
                asm volatile("cp.async.commit_group;" ::: "memory");
            } else if (tid < 9) {
                const uint32_t bar = bar0 + 8 * s;
                if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)STAGE) : "memory");
                __syncwarp(0x1ff);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(ring + s * STAGE + tid * CH * 8), "l"(arr[tid] + q), "r"((uint32_t)(CH * 8)), "r"(bar) : "memory");
            }
        };
        for (int l = 0; l < D - 1; l++) { if (l < np) issue(l); else if (M == 1) asm volatile("cp.async.commit_group;" ::: "memory"); }
        for (int n = 0; n < np; n++) {
            if (n + D - 1 < np) issue(n + D - 1); else if (M == 1) asm volatile("cp.async.commit_group;" ::: "memory");
            const uint32_t s = (g + n) % D;
            if (M == 1) {
                // D-1 groups may stay in flight
                if (D == 2) asm volatile("cp.async.wait_group 1;" ::: "memory");
                else if (D == 3) asm volatile("cp.async.wait_group 2;" ::: "memory");
                else if (D == 4) asm volatile("cp.async.wait_group 3;" ::: "memory");
                else asm volatile("cp.async.wait_group 5;" ::: "memory");
                __syncthreads();
            } else {
                mbar_wait(bar0 + 8 * s, ((g + n) / D) & 1u);
            }
            const double2 *st = (const double2 *)(smem + 128 + (size_t)s * STAGE);
            const long long q = (long long)(kb + n) * plane + off + 2 * tid;
            double2 r0 = st[0 * NT + tid], r1 = st[1 * NT + tid], r2 = st[2 * NT + tid];
            const double ax = 1e-9 * (r0.x + r1.x + r2.x), ay = 1e-9 * (r0.y + r1.y + r2.y);
#pragma unroll
            for (int a = 0; a < 6; a++) {
                double2 w = st[(3 + a) * NT + tid];
                __stcs((double2 *)(arr[3 + a] + q), make_double2(w.x + ax, w.y + ay));
            }
            __syncthreads();     // the stage may be refilled by the next issue
        }
        g += np;
    }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int M, int CH>
void run(double **d, long long plane, int nz, int kchunk, int D, int cps, double bytes)
{
    const size_t smem = 128 + (size_t)D * 9 * CH * 8;
    if (smem > 220 * 1024) return;
    CK(cudaFuncSetAttribute(staged<M, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, staged<M, CH>, CH / 2, smem));
    if (occ < 1) return;
    if (cps > occ) return;
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int it = 0; it < 2; it++) staged<M, CH><<<148 * cps, CH / 2, smem>>>(d, plane, nz, kchunk, D);
    CK(cudaEventRecord(a));
    for (int it = 0; it < 5; it++) staged<M, CH><<<148 * cps, CH / 2, smem>>>(d, plane, nz, kchunk, D);
    CK(cudaEventRecord(b));
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("M %d CH %d D %d: %s\n", M, CH, D, cudaGetErrorString(e)); exit(1); }
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); ms /= 5;
    printf("%s chunk %4d thr %3d D %d ctas/SM %d (max %d) kchunk %3d : %.3f ms  %.0f GB/s\n", M == 1 ? "cp.async " : "bulk TMA ", CH, CH / 2, D, cps, occ, kchunk, ms, bytes / ms / 1e6);
}

int main(int argc, char **argv)
{
    const int pitch = argc > 1 ? atoi(argv[1]) : 112, ny = argc > 2 ? atoi(argv[2]) : 641, nz = argc > 3 ? atoi(argv[3]) : 640;
    const long long plane = (long long)pitch * ny / 1024 * 1024, n = plane * nz;
    double *h[9], **d;
    for (int a = 0; a < 9; a++) { CK(cudaMalloc(&h[a], n * 8 + 4096)); CK(cudaMemset(h[a], 0, n * 8)); }
    CK(cudaMalloc(&d, sizeof(h))); CK(cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice));
    const double bytes = 15.0 * 8 * n;
    printf("plane %lld doubles x %d planes, 3R + 6RMW, %.2f GB per pass\n", plane, nz, bytes / 1e9);
    for (int kchunk : {64, 16})
    for (int D : {2, 3, 4, 6})
    for (int cps : {1, 2, 3, 4, 6, 8}) {
        run<1, 256>(d, plane, nz, kchunk, D, cps, bytes);
        run<2, 256>(d, plane, nz, kchunk, D, cps, bytes);
        run<1, 512>(d, plane, nz, kchunk, D, cps, bytes);
        run<2, 512>(d, plane, nz, kchunk, D, cps, bytes);
        run<1, 1024>(d, plane, nz, kchunk, D, cps, bytes);
        run<2, 1024>(d, plane, nz, kchunk, D, cps, bytes);
    }
    return 0;
}
