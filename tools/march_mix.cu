// Does the z-marching access pattern itself cap HBM bandwidth?  (run on the GPU box)
// Same traffic mix as k_stress3d (3 arrays read, 6 read-modify-written), no stencil, but the
// work is cut the way the C-PML kernels cut it: a CTA owns `chunk` contiguous doubles of a plane
// and marches `kchunk` planes (stride = plane), persistent CTAs, static round-robin.
// Compare with tools/stream_mix.cu (same mix, linear sweep).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

template <int V>   // V doubles per thread per plane (1, 2 or 4), contiguous
__global__ void __launch_bounds__(256) march(double *const *arr, long long plane, int nz, int chunk, int kchunk, int order)
{
    const int ntile = (int)((plane + chunk - 1) / chunk);
    const int nzc = (nz + kchunk - 1) / kchunk;
    const int nitems = ntile * nzc;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        int tile, zc;
        if (order == 0) { tile = item % ntile; zc = item / ntile; }      // tiles fastest (like the kernels)
        else            { zc = item % nzc; tile = item / nzc; }          // z chunks fastest
        const int kb = zc * kchunk, ke = min(nz, kb + kchunk);
        for (int e0 = threadIdx.x * V; e0 < chunk; e0 += 256 * V) {
            const long long off = (long long)tile * chunk + e0;
            if (off + V > plane) continue;
            for (int k = kb; k < ke; k++) {
                const long long q = (long long)k * plane + off;
                double r[3][V], w[6][V];
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    if (V == 4) { double2 t = __ldcs((const double2 *)(arr[a] + q)), u = __ldcs((const double2 *)(arr[a] + q) + 1); r[a][0] = t.x; r[a][1] = t.y; r[a][2] = u.x; r[a][3] = u.y; }
                    else if (V == 2) { double2 t = __ldcs((const double2 *)(arr[a] + q)); r[a][0] = t.x; r[a][1] = t.y; }
                    else r[a][0] = __ldcs(arr[a] + q);
                }
#pragma unroll
                for (int a = 0; a < 6; a++) {
                    if (V == 4) { double2 t = __ldcs((const double2 *)(arr[3 + a] + q)), u = __ldcs((const double2 *)(arr[3 + a] + q) + 1); w[a][0] = t.x; w[a][1] = t.y; w[a][2] = u.x; w[a][3] = u.y; }
                    else if (V == 2) { double2 t = __ldcs((const double2 *)(arr[3 + a] + q)); w[a][0] = t.x; w[a][1] = t.y; }
                    else w[a][0] = __ldcs(arr[3 + a] + q);
                }
#pragma unroll
                for (int a = 0; a < 6; a++) {
#pragma unroll
                    for (int v = 0; v < V; v++) w[a][v] += 1e-9 * (r[0][v] + r[1][v] + r[2][v]);
                    if (V == 4) { __stcs((double2 *)(arr[3 + a] + q), make_double2(w[a][0], w[a][1])); __stcs((double2 *)(arr[3 + a] + q) + 1, make_double2(w[a][2], w[a][3])); }
                    else if (V == 2) __stcs((double2 *)(arr[3 + a] + q), make_double2(w[a][0], w[a][1]));
                    else __stcs(arr[3 + a] + q, w[a][0]);
                }
            }
        }
    }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

int main(int argc, char **argv)
{
    const int pitch = argc > 1 ? atoi(argv[1]) : 112, ny = argc > 2 ? atoi(argv[2]) : 641, nz = argc > 3 ? atoi(argv[3]) : 640;
    const long long plane = (long long)pitch * ny, n = plane * nz;
    double *h[9], **d;
    for (int a = 0; a < 9; a++) { CK(cudaMalloc(&h[a], n * 8 + 4096)); CK(cudaMemset(h[a], 0, n * 8)); }
    CK(cudaMalloc(&d, sizeof(h))); CK(cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    printf("grid %d x %d x %d (pitch), 3R + 6RMW, %.2f GB per pass\n", pitch, ny, nz, 15.0 * 8 * n / 1e9);
    const int chunks[] = {256, 512, 1024, 2048, 4096, 16384};
    for (int order = 0; order < 2; order++)
    for (int kchunk : {640, 64, 16})
    for (int chunk : chunks)
    for (int cps : {4, 8}) {
        if (chunk % 4 || plane % 4) continue;
        auto launch = [&]() {
            const int V = chunk >= 1024 ? 4 : chunk >= 512 ? 2 : 1;
            if (V == 4) march<4><<<148 * cps, 256>>>(d, plane, nz, chunk, kchunk, order);
            else if (V == 2) march<2><<<148 * cps, 256>>>(d, plane, nz, chunk, kchunk, order);
            else march<1><<<148 * cps, 256>>>(d, plane, nz, chunk, kchunk, order);
        };
        launch(); launch();
        CK(cudaEventRecord(a));
        for (int it = 0; it < 5; it++) launch();
        CK(cudaEventRecord(b)); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, a, b)); ms /= 5;
        printf("order %d kchunk %3d chunk %5d doubles  ctas/SM %d : %.3f ms  %.0f GB/s\n", order, kchunk, chunk, cps, ms, 15.0 * 8 * n / ms / 1e6);
    }
    return 0;
}
