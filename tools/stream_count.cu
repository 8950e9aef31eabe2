// How many independent read-modify-write streams can a z-marching kernel sustain on B200?
// (run on the GPU box)  The viscoelastic stress kernel walks 3 read + 18 read-modify-write arrays
// (24 words) plus up to 9 shell arrays; this probe separates "number of arrays" from everything
// else: every thread owns an (i,j) column of a 32 x 8 tile, marches 16 planes, and for each of
// NV variables loads a word, scales it and stores it back.
//   layout 0: NV separate arrays [k][j][i]                      (what libcpml_b200 did first)
//   layout 1: one array [k][j][v][i]   (rows of the NV variables of a (j,k) line are adjacent)
//   layout 2: one array [k][v][j][i]   (planes of the NV variables of a k are adjacent)
// Prints GB/s (read + write) per configuration.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

struct P {
    double *base[48];
    long long sv, sj, sk;      // strides (doubles) of variable, row, plane inside one base (layouts 1, 2)
    int nx, ny, nz, nv, layout, kchunk;
};

template <int NV>
__global__ void __launch_bounds__(256) k(const __grid_constant__ P p)
{
    const int i = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 8 + threadIdx.y;
    if (i >= p.nx || j >= p.ny) return;
    const int kb = blockIdx.z * p.kchunk, ke = min(p.nz, kb + p.kchunk);
    for (int kk = kb; kk < ke; kk++) {
        double x[NV];
#pragma unroll
        for (int v = 0; v < NV; v++) {
            const double *a = p.layout == 0 ? p.base[v] + ((long long)kk * p.ny + j) * p.nx + i
                                            : p.base[0] + kk * p.sk + j * p.sj + v * p.sv + i;
            x[v] = __ldcs(a);
        }
#pragma unroll
        for (int v = 0; v < NV; v++) {
            double *a = p.layout == 0 ? p.base[v] + ((long long)kk * p.ny + j) * p.nx + i
                                      : p.base[0] + kk * p.sk + j * p.sj + v * p.sv + i;
            __stcs(a, x[v] * 1.0000001);
        }
    }
}

template <int NV>
static float run(P p, int reps)
{
    dim3 g((p.nx + 31) / 32, (p.ny + 7) / 8, (p.nz + p.kchunk - 1) / p.kchunk), b(32, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NV><<<g, b>>>(p);
    cudaEventRecord(e0);
    for (int r = 0; r < reps; r++) k<NV><<<g, b>>>(p);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

int main(int argc, char **argv)
{
    const int nx = argc > 1 ? atoi(argv[1]) : 1056, ny = argc > 2 ? atoi(argv[2]) : 1028, nz = argc > 3 ? atoi(argv[3]) : 64;
    const long long n = (long long)nx * ny * nz;
    const int nvs[] = {6, 12, 24, 36, 48};
    for (int nv : nvs) {
        for (int layout = 0; layout < 3; layout++) {
            P p{};
            p.nx = nx; p.ny = ny; p.nz = nz; p.nv = nv; p.layout = layout; p.kchunk = 16;
            std::vector<double *> owned;
            if (layout == 0) {
                for (int v = 0; v < nv; v++) { double *a; if (cudaMalloc(&a, n * 8) != cudaSuccess) { printf("alloc failed\n"); return 1; } cudaMemset(a, 0, n * 8); p.base[v] = a; owned.push_back(a); }
            } else {
                double *a; if (cudaMalloc(&a, n * 8 * nv) != cudaSuccess) { printf("alloc failed\n"); return 1; }
                cudaMemset(a, 0, n * 8 * nv); p.base[0] = a; owned.push_back(a);
                if (layout == 1) { p.sv = nx; p.sj = (long long)nx * nv; p.sk = (long long)nx * nv * ny; }
                else             { p.sj = nx; p.sv = (long long)nx * ny; p.sk = (long long)nx * ny * nv; }
            }
            float ms = 0;
            switch (nv) {
            case 6: ms = run<6>(p, 5); break;
            case 12: ms = run<12>(p, 5); break;
            case 24: ms = run<24>(p, 5); break;
            case 36: ms = run<36>(p, 5); break;
            default: ms = run<48>(p, 5); break;
            }
            printf("nv %2d layout %d : %8.3f ms  %7.1f GB/s\n", nv, layout, ms, 2.0 * n * 8 * nv / (ms * 1e-3) / 1e9);
            for (double *a : owned) cudaFree(a);
        }
    }
    return 0;
}
