// Developer probe: which 2-D tiled tensor copies (cp.async.bulk.tensor.2d) does this GPU / driver accept?
// nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu   (no -lcuda: the encoder is looked up at run time)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void probe(const __grid_constant__ CUtensorMap tm, float *out, int c, int r, int boxw, int boxh)
{
    extern __shared__ unsigned char dyn[];
    __shared__ __align__(8) uint64_t bar;
    float *tile = reinterpret_cast<float *>(dyn + ((128u - ((unsigned)__cvta_generic_to_shared(dyn) & 127u)) & 127u));
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(&bar)), "r"(boxw * boxh * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
                         (unsigned)__cvta_generic_to_shared(tile)), "l"(&tm), "r"(c), "r"(r), "r"((unsigned)__cvta_generic_to_shared(&bar)) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(
                     (unsigned)__cvta_generic_to_shared(&bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < boxw * boxh; i += blockDim.x) out[i] = tile[i];
}
int main()
{
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)p;
    printf("encoder %p status %d\n", p, (int)q);
    const int G = 512;
    std::vector<float> h((size_t)G * G);
    for (int i = 0; i < G * G; ++i) h[i] = (float)i;
    float *d, *out;
    cudaMalloc(&d, h.size() * 4); cudaMalloc(&out, 256 * 64 * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
    const int cases[][5] = {{32, 8, 0, 0, 0}, {128, 8, 0, 0, 0}, {160, 8, 0, 0, 0}, {160, 8, 36, 21, 0}, {160, 8, 400, 508, 0}, {160, 8, -4, -2, 0}, {160, 8, 36, 21, 1}, {256, 8, 8, 5, 0}, {160, 64, 32, 17, 0}, {160, 8, 38, 21, 0}, {160, 8, 37, 21, 0}};
    for (auto &cs : cases) {
        const int bw = cs[0], bh = cs[1], c = cs[2], r = cs[3];
        CUtensorMap tm;
        const cuuint64_t dims[2] = {(cuuint64_t)G, (cuuint64_t)G};
        const cuuint64_t strides[1] = {(cuuint64_t)G * 4};
        const cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh};
        const cuuint32_t estr[2] = {1, 1};
        CUresult rc = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, cs[4] ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cudaMemset(out, 0xff, 256 * 64 * 4);
        probe<<<1, 256, bw * bh * 4 + 128>>>(tm, out, c, r, bw, bh);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> o((size_t)bw * bh);
        cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int y = 0; y < bh; ++y)
            for (int x = 0; x < bw; ++x) {
                const int gx = c + x, gy = r + y;
                const float want = (gx >= 0 && gx < G && gy >= 0 && gy < G) ? (float)(gy * G + gx) : 0.0f;
                bad += (o[(size_t)y * bw + x] != want);
            }
        printf("box %dx%d at (%d,%d) l2promo %d: encode %d run %s mismatches %d\n", bw, bh, c, r, cs[4], (int)rc, cudaGetErrorString(e), bad);
        if (e != cudaSuccess) return 1;
    }
    return 0;
}
