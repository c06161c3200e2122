// Bring-up probe: tcgen05.mma with the A operand in TENSOR MEMORY (written with tcgen05.st, thread = row, 2 bf16 per
// 32-bit column) and B in shared memory (K-major SWIZZLE_128B, as in the recurrent kernel).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o build_tmp/umma_tmem_probe tools/umma_tmem_a_test.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

constexpr int M = 128, N = 64, K = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__global__ void probe(const __nv_bfloat16* __restrict__ a, const unsigned char* __restrict__ b_img, float* __restrict__ out, int variant) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t base = (smem_u32(sm) + 1023u) & ~1023u;
    unsigned char* sb = sm + (base - smem_u32(sm));
    for (int i = tid; i < N * 128; i += blockDim.x) sb[i] = b_img[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_slot)), "r"(128u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const uint32_t a_col = 64;                       // D: columns 0..63, A: columns 64..95 (K = 64 bf16 = 32 words per row)
    if (warp < 4) {
        // thread = row: K = 64 bf16 -> 32 words, two x16 stores
        const int row = warp * 32 + lane;
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + a_col;
        uint32_t w[32];
        for (int i = 0; i < 32; ++i) {
            const uint16_t lo = __bfloat16_as_ushort(a[row * K + 2 * i]), hi = __bfloat16_as_ushort(a[row * K + 2 * i + 1]);
            w[i] = (uint32_t)lo | ((uint32_t)hi << 16);
        }
        for (int h = 0; h < 2; ++h)
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n" ::"r"(taddr + 16 * h),
                "r"(w[16 * h + 0]), "r"(w[16 * h + 1]), "r"(w[16 * h + 2]), "r"(w[16 * h + 3]), "r"(w[16 * h + 4]), "r"(w[16 * h + 5]),
                "r"(w[16 * h + 6]), "r"(w[16 * h + 7]), "r"(w[16 * h + 8]), "r"(w[16 * h + 9]), "r"(w[16 * h + 10]), "r"(w[16 * h + 11]),
                "r"(w[16 * h + 12]), "r"(w[16 * h + 13]), "r"(w[16 * h + 14]), "r"(w[16 * h + 15])
                : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint64_t db0 = make_desc_sw128(base);
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint32_t a_t = tmem + a_col + (variant == 0 ? 8 * ks : 16 * ks);   // variant 0: 8 columns per k16 step
            const uint32_t acc = ks ? 1u : 0u;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem), "r"(a_t), "l"(db0 + 2 * ks), "r"(idesc), "r"(acc)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
    }
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
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(128u));
}

int main() {
    std::vector<float> A((size_t)M * K), B((size_t)N * K);
    srand(2);
    for (auto& v : A) v = (float)(rand() % 17 - 8) / 8.f;
    for (auto& v : B) v = (float)(rand() % 13 - 6) / 4.f;
    std::vector<__nv_bfloat16> a_bf(A.size());
    for (size_t i = 0; i < A.size(); ++i) a_bf[i] = __float2bfloat16(A[i]);
    std::vector<unsigned char> b_img((size_t)N * 128, 0);
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) {
            const __nv_bfloat16 v = __float2bfloat16(B[(size_t)n * K + k]);
            const size_t off = (size_t)n * 128 + (((size_t)(k >> 3) ^ (size_t)(n & 7)) << 4) + (k & 7) * 2;
            memcpy(&b_img[off], &v, 2);
        }
    __nv_bfloat16* da; unsigned char* db; float* dout;
    cudaMalloc(&da, a_bf.size() * 2); cudaMalloc(&db, b_img.size()); cudaMalloc(&dout, M * N * 4);
    cudaMemcpy(da, a_bf.data(), a_bf.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b_img.data(), b_img.size(), cudaMemcpyHostToDevice);
    for (int variant = 0; variant < 2; ++variant) {
        cudaMemset(dout, 0, M * N * 4);
        probe<<<1, 128, N * 128 + 1024>>>(da, db, dout, variant);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> out((size_t)M * N);
        cudaMemcpy(out.data(), dout, M * N * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                double ref = 0;
                for (int k = 0; k < K; ++k) ref += (double)A[(size_t)m * K + k] * B[(size_t)n * K + k];
                maxerr = fmax(maxerr, fabs(ref - out[(size_t)m * N + n]));
            }
        printf("A in TMEM, variant %d (%d columns per k16 step): max err %.4f [%s]\n", variant, variant == 0 ? 8 : 16, maxerr, cudaGetErrorString(e));
    }
    return 0;
}
