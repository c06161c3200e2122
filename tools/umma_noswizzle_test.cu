// Bring-up probe: tcgen05.mma with NO-SWIZZLE K-major operands laid out as [K/8 groups][rows][8 bf16] (16 bytes per
// (row, group), rows contiguous), which is what an implicit-GEMM causal convolution wants: a tap is then just a
// row shift of the descriptor start address.  Tries both assignments of the LBO / SBO descriptor fields and a
// shifted start, and prints the max error against a CPU reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/umma_probe tools/umma_noswizzle_test.cu && /tmp/umma_probe
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

constexpr int M = 128, ROWS = 160, N = 32, K = 32;   // A buffer has ROWS rows; the MMA reads rows [shift, shift + 128)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;    // descriptor version (Blackwell)
    // layout type 0 = no swizzle
    return d;
}
__device__ __forceinline__ uint32_t make_idesc(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__global__ void probe(const __nv_bfloat16* __restrict__ a_img, const __nv_bfloat16* __restrict__ b_img, float* __restrict__ out,
                      int variant, int shift) {
    __shared__ __align__(128) __nv_bfloat16 sa[(K / 8) * ROWS * 8];
    __shared__ __align__(128) __nv_bfloat16 sb[(K / 8) * N * 8];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (K / 8) * ROWS * 8; i += blockDim.x) sa[i] = a_img[i];
    for (int i = tid; i < (K / 8) * N * 8; i += blockDim.x) sb[i] = b_img[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_slot)), "r"(32u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic smem writes -> tensor core reads
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint32_t a_addr = smem_u32(sa) + shift * 16, b_addr = smem_u32(sb);
        const uint32_t a_kstride = ROWS * 16, b_kstride = N * 16, mn_stride = 128;   // K-group stride, 8-row-group stride
        const uint32_t idesc = make_idesc(M, N);
        for (int ks = 0; ks < K / 16; ++ks) {
            uint64_t da, db;
            if (variant == 0) {   // LBO = stride between core matrices along K, SBO = along M/N
                da = make_desc(a_addr + ks * 2 * a_kstride, a_kstride, mn_stride);
                db = make_desc(b_addr + ks * 2 * b_kstride, b_kstride, mn_stride);
            } else {              // swapped
                da = make_desc(a_addr + ks * 2 * a_kstride, mn_stride, a_kstride);
                db = make_desc(b_addr + ks * 2 * b_kstride, mn_stride, b_kstride);
            }
            const uint32_t acc = ks ? 1u : 0u;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
    }
    // wait
    {
        uint32_t ok = 0;
        const long long t0 = clock64();
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
            if (clock64() - t0 > 2000000000LL) break;
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if (warp < 4) {
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        for (int c0 = 0; c0 < N; c0 += 16) {
            uint32_t r[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(taddr + c0));
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * N + c0 + i] = __uint_as_float(r[i]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(32u));
}

int main() {
    std::vector<float> A((size_t)ROWS * K), B((size_t)N * K);
    srand(1);
    for (auto& v : A) v = (float)(rand() % 17 - 8) / 8.f;
    for (auto& v : B) v = (float)(rand() % 13 - 6) / 4.f;
    std::vector<__nv_bfloat16> a_img((size_t)(K / 8) * ROWS * 8), b_img((size_t)(K / 8) * N * 8);
    for (int r = 0; r < ROWS; ++r)
        for (int k = 0; k < K; ++k) a_img[((size_t)(k / 8) * ROWS + r) * 8 + k % 8] = __float2bfloat16(A[(size_t)r * K + k]);
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) b_img[((size_t)(k / 8) * N + n) * 8 + k % 8] = __float2bfloat16(B[(size_t)n * K + k]);
    __nv_bfloat16 *da, *db;
    float* dout;
    cudaMalloc(&da, a_img.size() * 2); cudaMalloc(&db, b_img.size() * 2); cudaMalloc(&dout, M * N * 4);
    cudaMemcpy(da, a_img.data(), a_img.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b_img.data(), b_img.size() * 2, cudaMemcpyHostToDevice);
    for (int variant = 0; variant < 2; ++variant)
        for (int shift : {0, 1, 5, 8, 27}) {
            cudaMemset(dout, 0, M * N * 4);
            probe<<<1, 128>>>(da, db, dout, variant, shift);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<float> out((size_t)M * N);
            cudaMemcpy(out.data(), dout, M * N * 4, cudaMemcpyDeviceToHost);
            double maxerr = 0;
            for (int m = 0; m < M; ++m)
                for (int n = 0; n < N; ++n) {
                    double ref = 0;
                    for (int k = 0; k < K; ++k) ref += (double)A[(size_t)(m + shift) * K + k] * B[(size_t)n * K + k];
                    maxerr = fmax(maxerr, fabs(ref - out[(size_t)m * N + n]));
                }
            printf("variant %d (%s) shift %2d: max err %.4f  [%s]\n", variant, variant == 0 ? "LBO=K stride, SBO=MN stride" : "swapped",
                   shift, maxerr, cudaGetErrorString(e));
        }
    return 0;
}
