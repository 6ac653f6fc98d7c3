#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
struct Maps { CUtensorMap m[16]; };
struct Big { int pad[600]; };
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int MODE>
__global__ void k(const __grid_constant__ CUtensorMap one, const __grid_constant__ Maps maps, const __grid_constant__ Big big, int l, int x, int y, int z, int bytes, unsigned* out) {
    __shared__ __align__(128) uint8_t tile[80 * 76];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned barAddr = smem_u32(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(barAddr));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const CUtensorMap* mp = MODE == 0 ? &one : &maps.m[l];
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barAddr), "r"((unsigned)bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(smem_u32(tile)), "l"(mp), "r"(x), "r"(y), "r"(z), "r"(barAddr) : "memory");
    }
    unsigned done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(barAddr), "r"(0u) : "memory");
    }
    if (threadIdx.x < 8) out[threadIdx.x] = tile[threadIdx.x] + 256 * tile[48 + threadIdx.x] + big.pad[0];
}
typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    Fn fn = (Fn)p;
    int w = 752, h = 480, pitch = 768, B = 2;
    uint8_t* d; cudaMalloc(&d, (size_t)pitch * h * B);
    std::vector<uint8_t> hbuf((size_t)pitch * h * B);
    for (size_t i = 0; i < hbuf.size(); ++i) hbuf[i] = (uint8_t)(i % 251);
    cudaMemcpy(d, hbuf.data(), hbuf.size(), cudaMemcpyHostToDevice);
    Maps maps; CUtensorMap one;
    cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * h};
    cuuint32_t box[3] = {48, 44, 1}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = fn(&one, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d\n", (int)r);
    for (int i = 0; i < 16; ++i) maps.m[i] = one;
    Big big{}; unsigned* out; cudaMalloc(&out, 64);
    unsigned ho[8];
    for (int mode = 0; mode < 2; ++mode) {
        if (mode == 0) k<0><<<1, 128>>>(one, maps, big, 3, 15, 16, 1, 48 * 44, out); else k<1><<<1, 128>>>(one, maps, big, 3, 15, 16, 1, 48 * 44, out);
        cudaError_t e = cudaDeviceSynchronize();
        printf("mode %d: %s\n", mode, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        cudaMemcpy(ho, out, 32, cudaMemcpyDeviceToHost);
        size_t base = (size_t)pitch * h * 1 + 16 * pitch + 15;
        printf("got %u %u expect %u %u\n", ho[0] & 255, ho[0] >> 8, hbuf[base], hbuf[base + pitch]);
    }
    return 0;
}
