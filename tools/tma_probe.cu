// tma_probe.cu — which cp.async.bulk.tensor (u8, SWIZZLE_NONE) box configurations does this GPU/driver accept?
// usage: tma_probe <xstart> <boxw> <by_version 0|1> <desc_in_global 0|1>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <vector>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int BW>
__global__ void k(const __grid_constant__ CUtensorMap tmp, const CUtensorMap* tmg, int use_g, int x, int y, uint8_t* out) {
    __shared__ __align__(128) uint8_t s[BW * 8];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
        asm volatile("fence.proxy.async.shared::cta;");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const CUtensorMap* tm = use_g ? tmg : &tmp;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(BW * 8));
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(s32(s)), "l"((uint64_t)tm), "r"(x), "r"(y), "r"(0), "r"(s32(&bar)) : "memory");
    }
    uint32_t ok = 0;
    for (int i = 0; !ok && i < 1000000; i++)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(s32(&bar)) : "memory");
    for (int i = threadIdx.x; i < BW * 8; i += blockDim.x) out[i] = ok ? s[i] : 0xEE;
}
int main(int argc, char** argv) {
    const int xs = atoi(argv[1]), bw = atoi(argv[2]), byver = atoi(argv[3]), useg = atoi(argv[4]);
    const int W = 320, H = 64;
    std::vector<uint8_t> h(W * H);
    for (int i = 0; i < W * H; i++) h[i] = (uint8_t)((i % W) ^ (i / W));
    uint8_t *d, *o; cudaMalloc(&d, W * H); cudaMalloc(&o, 4096); cudaMemcpy(d, h.data(), W * H, cudaMemcpyHostToDevice);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    if (byver) cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &q);
    else cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto enc = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    CUtensorMap tm;
    cuuint64_t gd[3] = {(cuuint64_t)W, (cuuint64_t)H, 1}, gs[2] = {(cuuint64_t)W, (cuuint64_t)W * H};
    cuuint32_t box[3] = {(cuuint32_t)bw, 8, 1}, es[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("xs=%d bw=%d byver=%d global_desc=%d encode=%d ", xs, bw, byver, useg, (int)r);
    if (r != CUDA_SUCCESS) { printf("\n"); return 1; }
    CUtensorMap* dtm; cudaMalloc(&dtm, sizeof(tm)); cudaMemcpy(dtm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
    if (bw == 80) k<80><<<1, 128>>>(tm, dtm, useg, xs, 5, o); else if (bw == 96) k<96><<<1, 128>>>(tm, dtm, useg, xs, 5, o); else k<64><<<1, 128>>>(tm, dtm, useg, xs, 5, o);
    cudaError_t e = cudaDeviceSynchronize();
    printf("run=%s ", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<uint8_t> res(bw * 8); cudaMemcpy(res.data(), o, bw * 8, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int yy = 0; yy < 8; yy++) for (int xx = 0; xx < bw; xx++) {
            int gx = xs + xx, gy = 5 + yy; uint8_t exp = (gx >= 0 && gx < W && gy < H) ? h[gy * W + gx] : 0;
            bad += res[yy * bw + xx] != exp;
        }
        printf("mismatches=%d", bad);
    }
    printf("\n");
    return 0;
}
