// Probe of the TMA assumptions used by kernels_3d_tma.cu (run on the GPU box):
//   FP64 3-D tiled tensor maps, boxes larger than the tensor, negative / overhanging
//   coordinates (zero fill), transaction bytes = full box.
// nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu && ./tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap tm, int x, int y, int z, int box_doubles, uint32_t tx_bytes,
                      double *out, int *status)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t sbase = (smem_u32(smem) + 127u) & ~127u;
    unsigned char *g = smem + (sbase - smem_u32(smem));
    const uint32_t bar = sbase, dst = sbase + 128;
    double *tile = (double *)(g + 128);
    for (int q = threadIdx.x; q < box_doubles; q += blockDim.x) tile[q] = -777.0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tx_bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(dst), "l"((unsigned long long)&tm), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
    }
    int ok = 0;
    for (int spin = 0; spin < (1 << 20) && !ok; spin++) {
        uint32_t r;
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(r) : "r"(bar), "r"(0u) : "memory");
        ok = r;
    }
    if (threadIdx.x == 0) *status = ok;
    __syncthreads();
    for (int q = threadIdx.x; q < box_doubles; q += blockDim.x) out[q] = tile[q];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

int main(int argc, char **argv)
{
    const int nx = 37, ny = 45, nzp = 10, pitch = 48;
    const long long plane = (long long)pitch * ny;
    std::vector<double> h((size_t)plane * nzp + 64, 0.0);
    for (int z = 0; z < nzp; z++) for (int y = 0; y < ny; y++) for (int x = 0; x < pitch; x++)
        h[16 + z * plane + (long long)y * pitch + x] = 1000.0 * z + 10.0 * y + 0.01 * x + 1.0;   // never 0; pad lanes too
    double *d; CK(cudaMalloc(&d, h.size() * 8)); CK(cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
    double *dout; CK(cudaMalloc(&dout, 1 << 20)); int *dst; CK(cudaMalloc(&dst, 4));
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)fn;
    if (argc < 6) { printf("usage: tma_probe bx by x y z\n"); return 3; }
    const int b[2] = {atoi(argv[1]), atoi(argv[2])};
    const int c[3] = {atoi(argv[3]), atoi(argv[4]), atoi(argv[5])};
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    int bad = 0;
    {
        CUtensorMap tm;
        const cuuint64_t dims[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nzp};
        const cuuint64_t strides[2] = {(cuuint64_t)pitch * 8, (cuuint64_t)plane * 8};
        const cuuint32_t box[3] = {(cuuint32_t)b[0], (cuuint32_t)b[1], 1u};
        const cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void *)(d + 16), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("box %dx%d: encode -> %d\n", b[0], b[1], (int)r); return 4; }
        {
            const int nd = b[0] * b[1];
            CK(cudaMemset(dst, 0, 4));
            probe<<<1, 128, 128 + 128 + nd * 8>>>(tm, c[0], c[1], c[2], nd, (uint32_t)nd * 8, dout, dst);
            cudaError_t e = cudaDeviceSynchronize();
            int st = -1; std::vector<double> o(nd);
            if (e == cudaSuccess) { CK(cudaMemcpy(&st, dst, 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(o.data(), dout, nd * 8, cudaMemcpyDeviceToHost)); }
            int wrong = 0;
            for (int yy = 0; yy < b[1]; yy++) for (int xx = 0; xx < b[0]; xx++) {
                const int gx = c[0] + xx, gy = c[1] + yy;
                const double want = (gx >= 0 && gx < nx && gy >= 0 && gy < ny && c[2] >= 0 && c[2] < nzp) ? h[16 + c[2] * plane + (long long)gy * pitch + gx] : 0.0;
                if (e == cudaSuccess && o[yy * b[0] + xx] != want) wrong++;
            }
            printf("box %3dx%d coord (%3d,%3d,%2d): err=%s completed=%d wrong=%d of %d\n", b[0], b[1], c[0], c[1], c[2], cudaGetErrorString(e), st, wrong, nd);
            if (e != cudaSuccess || st != 1 || wrong) bad++;
        }
    }
    return bad ? 1 : 0;
}
