// Stand-alone probe of the TMA row copy used by k_threshold_tma (run on the GPU box): tma_probe <case>
//   case 0: box <= tensor, x >= 0     case 1: box <= tensor, x = -2 (left OOB)     case 2: box > tensor width
//   case 3: row pitch in shared memory not 128-byte aligned (784)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
__global__ void k(const __grid_constant__ CUtensorMap map, int x4, int y, int f, int pitch, int rows, int rowbytes, uint8_t* out) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    uint32_t sb = (uint32_t)__cvta_generic_to_shared(&bar), ss = (uint32_t)__cvta_generic_to_shared(sm);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sb), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sb), "r"(rows * rowbytes) : "memory");
        for (int r = 0; r < rows; r++)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(ss + r * pitch),
                         "l"(&map), "r"(x4), "r"(y + r), "r"(f), "r"(sb)
                         : "memory");
    }
    __syncthreads();
    asm volatile(
        "{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D1;\nbra W1;\nD1:\n}\n" ::"r"(sb), "r"(0) : "memory");
    for (int i = threadIdx.x; i < rows * rowbytes; i += blockDim.x) out[i] = sm[(i / rowbytes) * pitch + i % rowbytes];
}
typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                           const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
    int c = argc > 1 ? atoi(argv[1]) : 0;
    int W = 640, H = 480, B = 2;
    int boxw = c == 2 ? 672 : 272, x4 = c == 0 ? 4 : (c == 4 ? 100 : -4), pitch = (boxw + 127) & ~127, rows = 3;
    if (c == 3) { boxw = 272; pitch = 272 + 16; }
    std::vector<uint8_t> h((size_t)W * H * B);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 7 + (i >> 9));
    uint8_t *d, *o;
    cudaMalloc(&d, h.size());
    cudaMalloc(&o, rows * boxw);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    CUtensorMap map;
    cuuint64_t gd[3] = {(cuuint64_t)W / 4, (cuuint64_t)H, (cuuint64_t)B}, gs[2] = {(cuuint64_t)W, (cuuint64_t)W * H};
    cuuint32_t box[3] = {(cuuint32_t)boxw / 4, 1, 1}, es[3] = {1, 1, 1};
    CUresult r = ((enc_fn)p)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("case %d: encode rc=%d boxw=%d x4=%d pitch=%d\n", c, (int)r, boxw, x4, pitch);
    if (r) return 1;
    k<<<1, 128, rows * pitch + 128>>>(map, x4, 5, 1, pitch, rows, boxw, o);
    cudaError_t e = cudaDeviceSynchronize();
    printf("  kernel: %s\n", cudaGetErrorString(e));
    if (e) return 2;
    std::vector<uint8_t> g(rows * boxw);
    cudaMemcpy(g.data(), o, g.size(), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int rr = 0; rr < rows; rr++)
        for (int i = 0; i < boxw; i++) {
            int x = 4 * x4 + i;
            uint8_t exp = (x < 0 || x >= W) ? 0 : h[(size_t)1 * W * H + (size_t)(5 + rr) * W + x];
            bad += g[rr * boxw + i] != exp;
        }
    printf("  mismatches: %d\n", bad);
    return 0;
}
