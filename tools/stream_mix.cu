// Practical HBM ceiling for the traffic mix of the C-PML kernels (run on the GPU box):
// an element-wise kernel that reads NR arrays and read-modify-writes NW arrays of doubles,
// no stencil, fully coalesced.  stress = 3 reads + 6 RMW, velocity = 6 reads + 3 RMW,
// copy = 1 read + 1 write.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_mix stream_mix.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

template <int NR, int NW, bool RMW>
__global__ void __launch_bounds__(256) mix(double *const *r, double *const *w, long long n2)
{
    for (long long q = (long long)blockIdx.x * 256 + threadIdx.x; q < n2; q += (long long)gridDim.x * 256) {
        double2 acc = make_double2(0.0, 0.0);
        double2 rv[NR + 1];
#pragma unroll
        for (int a = 0; a < NR; a++) rv[a] = __ldcs(((const double2 *)r[a]) + q);
        double2 wv[NW + 1];
        if (RMW) {
#pragma unroll
            for (int a = 0; a < NW; a++) wv[a] = __ldcs(((const double2 *)w[a]) + q);
        }
#pragma unroll
        for (int a = 0; a < NR; a++) { acc.x += rv[a].x; acc.y += rv[a].y; }
        if (NW == 0 && acc.x == 123.456) ((double2 *)r[0])[q] = acc;
#pragma unroll
        for (int a = 0; a < NW; a++) {
            double2 o = RMW ? make_double2(wv[a].x + 1e-9 * acc.x, wv[a].y + 1e-9 * acc.y) : acc;
            __stcs(((double2 *)w[a]) + q, o);
        }
    }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int NR, int NW, bool RMW>
void run(const char *name, double **dr, double **dw, long long n, int ctas_per_sm)
{
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    const int grid = 148 * ctas_per_sm;
    for (int it = 0; it < 3; it++) mix<NR, NW, RMW><<<grid, 256>>>(dr, dw, n / 2);
    CK(cudaEventRecord(a));
    const int reps = 10;
    for (int it = 0; it < reps; it++) mix<NR, NW, RMW><<<grid, 256>>>(dr, dw, n / 2);
    CK(cudaEventRecord(b)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); ms /= reps;
    const double bytes = 8.0 * n * (NR + (RMW ? 2 : 1) * NW);
    printf("%-28s grid %4d  %.3f ms  %.0f GB/s\n", name, grid, ms, bytes / ms / 1e6);
}

int main()
{
    const long long n = 101LL * 641 * 640 / 2 * 2;     // points of the default 3-D grid
    double *h[12], **d;
    for (int a = 0; a < 12; a++) { CK(cudaMalloc(&h[a], n * 8)); CK(cudaMemset(h[a], 0, n * 8)); }
    CK(cudaMalloc(&d, sizeof(h))); CK(cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice));
    for (int c : {2, 4, 8, 16}) {
        run<1, 1, false>("copy 1R + 1W", d, d + 6, n, c);
        run<3, 6, true>("stress mix 3R + 6RMW", d, d + 6, n, c);
        run<6, 3, true>("velocity mix 6R + 3RMW", d, d + 3 + 3, n, c);
        run<9, 0, false>("read only 9R", d, d + 9, n, c);
        run<0, 6, false>("write only 6W", d, d + 6, n, c);
    }
    return 0;
}
