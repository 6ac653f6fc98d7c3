#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int RANK>
__global__ void k(const __grid_constant__ CUtensorMap one, int x, int y, int z, int bytes, unsigned* out) {
    extern __shared__ __align__(1024) uint8_t tile[];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned barAddr = smem_u32(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(barAddr));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barAddr), "r"((unsigned)bytes) : "memory");
        if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(smem_u32(tile)), "l"(&one), "r"(x), "r"(y), "r"(z), "r"(barAddr) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_u32(tile)), "l"(&one), "r"(x), "r"(y), "r"(barAddr) : "memory");
    }
    unsigned done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(barAddr), "r"(0u) : "memory");
    }
    if (threadIdx.x < 8) out[threadIdx.x] = tile[threadIdx.x];
}
typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
    int rank = atoi(argv[1]), boxw = atoi(argv[2]), boxh = atoi(argv[3]), x = atoi(argv[4]), l2 = atoi(argv[5]);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    Fn fn = (Fn)p;
    int w = 752, h = 480, pitch = 768, B = 2;
    uint8_t* d; cudaMalloc(&d, (size_t)pitch * h * B);
    std::vector<uint8_t> hbuf((size_t)pitch * h * B);
    for (size_t i = 0; i < hbuf.size(); ++i) hbuf[i] = (uint8_t)(i % 251);
    cudaMemcpy(d, hbuf.data(), hbuf.size(), cudaMemcpyHostToDevice);
    CUtensorMap one;
    cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * h};
    cuuint32_t box[3] = {(cuuint32_t)boxw, (cuuint32_t)boxh, 1}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = fn(&one, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    unsigned* out; cudaMalloc(&out, 64);
    unsigned ho[8];
    if (rank == 3) k<3><<<1, 128, 16384>>>(one, x, 16, 1, boxw * boxh, out); else k<2><<<1, 128, 16384>>>(one, x, 16, 0, boxw * boxh, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("rank %d box %dx%d x=%d l2=%d encode=%d: %s\n", rank, boxw, boxh, x, l2, (int)r, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    cudaMemcpy(ho, out, 32, cudaMemcpyDeviceToHost);
    size_t base = (size_t)pitch * h * (rank == 3 ? 1 : 0) + 16 * pitch + x;
    printf("  got %u %u expect %u %u\n", ho[0], ho[1], hbuf[base], hbuf[base + 1]);
    return 0;
}
