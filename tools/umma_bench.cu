// Micro-benchmarks behind the round-2 redesign of the recurrent kernel's MMA stage (run on a B200):
//   1. tcgen05.mma issue/throughput, M = 128, K = 16, N in {16..256}: A from shared memory (SS) against A from tensor
//      memory (TS), for the three-product split-bf16 step and its "stacked" two-MMA form
//   2. tcgen05.cp 128x256b shared -> tensor memory: cycles per 32 KiB activation chunk, and a correctness check that a
//      copy with the MMA's own K-major SWIZZLE_128B descriptor puts the operand where a TS MMA expects it
//   3. copy + MMA pipelined chunk by chunk (what a phase of the recurrent kernel would issue)
//   4. the split-K reduce-scatter over a 4-CTA cluster: st.async from registers (today) against bulk DSMEM copies
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build_tmp/umma_bench tools/umma_bench.cu && build_tmp/umma_bench
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {     // K-major SWIZZLE_128B, SBO = 1024
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
                 "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tcp_128x256(uint32_t taddr, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;\n" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
    const long long t0 = clock64();
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (clock64() - t0 > 2000000000LL) return false;
    }
    return true;
}

constexpr int ACT_PART = 128 * 128;          // 128 rows x 64 bf16
constexpr int ACT_CHUNK = 2 * ACT_PART;      // hi + lo
constexpr int SM_A = 0, SM_B = 4 * ACT_CHUNK;   // B: up to 256 rows hi + 256 rows lo = 64 KiB
constexpr int SM_TOTAL = SM_B + 64 * 1024 + 1024;

// mode: 0 SS 3-product, 1 SS stacked (2N + N), 2 TS 3-product, 3 TS stacked, 4 SS one MMA, 5 TS one MMA
template <int mode>
__global__ void __launch_bounds__(128, 1) k_mma(int N, int reps, long long* out) {
    extern __shared__ __align__(1024) unsigned char sm_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(sm_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5;
    unsigned char* sm = sm_raw + (base - smem_u32(sm_raw));
    for (int i = tid; i < (SM_B + 64 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u + (i & 0xff);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const uint32_t a_hi_t = tmem + 256, a_lo_t = tmem + 384;
    if (warp == 0) {
        const uint64_t dA = make_desc(base + SM_A), dB = make_desc(base + SM_B);
        const uint32_t id1 = make_idesc(128, N), id2 = make_idesc(128, 2 * N);
        uint32_t lane_pred;
        asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(lane_pred));
        long long t0 = clock64();
        if (lane_pred) {
            for (int r = 0; r < reps; ++r)
                for (int c = 0; c < 4; ++c) {
                    const uint64_t dah = dA + (uint64_t)((c * ACT_CHUNK) >> 4), dal = dah + (ACT_PART >> 4);
                    const uint64_t dwh = dB, dwl = dB + (uint64_t)((N * 128) >> 4);
                    const uint32_t th = a_hi_t + c * 32, tl = a_lo_t + c * 32;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint32_t acc = (r | c | ks) ? 1u : 0u;
                        if (mode == 0) {
                            mma_ss(tmem, dal + 2 * ks, dwh + 2 * ks, id1, acc);
                            mma_ss(tmem, dah + 2 * ks, dwl + 2 * ks, id1, 1u);
                            mma_ss(tmem, dah + 2 * ks, dwh + 2 * ks, id1, 1u);
                        } else if (mode == 1) {
                            mma_ss(tmem, dah + 2 * ks, dwh + 2 * ks, id2, acc);
                            mma_ss(tmem, dal + 2 * ks, dwh + 2 * ks, id1, 1u);
                        } else if (mode == 2) {
                            mma_ts(tmem, tl + 8 * ks, dwh + 2 * ks, id1, acc);
                            mma_ts(tmem, th + 8 * ks, dwl + 2 * ks, id1, 1u);
                            mma_ts(tmem, th + 8 * ks, dwh + 2 * ks, id1, 1u);
                        } else if (mode == 3) {
                            mma_ts(tmem, th + 8 * ks, dwh + 2 * ks, id2, acc);
                            mma_ts(tmem, tl + 8 * ks, dwh + 2 * ks, id1, 1u);
                        } else if (mode == 4) {
                            mma_ss(tmem, dah + 2 * ks, dwh + 2 * ks, id1, acc);
                        } else {
                            mma_ts(tmem, th + 8 * ks, dwh + 2 * ks, id1, acc);
                        }
                    }
                }
            commit(&bar);
        }
        __syncwarp();
        long long t_issued = clock64();
        const bool ok = mbar_wait_bounded(&bar, 0);
        long long t1 = clock64();
        if (tid == 0) {
            out[blockIdx.x * 3 + 0] = ok ? (t1 - t0) : -1;
            out[blockIdx.x * 3 + 1] = t_issued - t0;
            out[blockIdx.x * 3 + 2] = 0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512u));
}

// tcgen05.cp of whole activation chunks; mode 0: copies only, 1: copy + TS MMAs of N = 64 chunk by chunk (stacked),
// 2: copy + TS 3-product
template <int mode>
__global__ void __launch_bounds__(128, 1) k_cp(int N, int reps, long long* out) {
    extern __shared__ __align__(1024) unsigned char sm_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(sm_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5;
    unsigned char* sm = sm_raw + (base - smem_u32(sm_raw));
    for (int i = tid; i < (SM_B + 64 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u + (i & 0xff);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const uint32_t a_hi_t = tmem + 256, a_lo_t = tmem + 384;
    if (warp == 0) {
        const uint64_t dA = make_desc(base + SM_A), dB = make_desc(base + SM_B);
        const uint32_t id1 = make_idesc(128, N), id2 = make_idesc(128, 2 * N);
        uint32_t lane_pred;
        asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(lane_pred));
        long long t0 = clock64();
        if (lane_pred) {
            for (int r = 0; r < reps; ++r)
                for (int c = 0; c < 4; ++c) {
                    const uint64_t dah = dA + (uint64_t)((c * ACT_CHUNK) >> 4), dal = dah + (ACT_PART >> 4);
                    const uint64_t dwh = dB, dwl = dB + (uint64_t)((N * 128) >> 4);
                    const uint32_t th = a_hi_t + c * 32, tl = a_lo_t + c * 32;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        tcp_128x256(th + 8 * ks, dah + 2 * ks);
                        tcp_128x256(tl + 8 * ks, dal + 2 * ks);
                    }
                    if (mode == 1) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            mma_ts(tmem, th + 8 * ks, dwh + 2 * ks, id2, (r | c | ks) ? 1u : 0u);
                            mma_ts(tmem, tl + 8 * ks, dwh + 2 * ks, id1, 1u);
                        }
                    } else if (mode == 2) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            mma_ts(tmem, tl + 8 * ks, dwh + 2 * ks, id1, (r | c | ks) ? 1u : 0u);
                            mma_ts(tmem, th + 8 * ks, dwl + 2 * ks, id1, 1u);
                            mma_ts(tmem, th + 8 * ks, dwh + 2 * ks, id1, 1u);
                        }
                    }
                }
            commit(&bar);
        }
        __syncwarp();
        long long t_issued = clock64();
        const bool ok = mbar_wait_bounded(&bar, 0);
        long long t1 = clock64();
        if (tid == 0) {
            out[blockIdx.x * 3 + 0] = ok ? (t1 - t0) : -1;
            out[blockIdx.x * 3 + 1] = t_issued - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512u));
}

// Correctness of copy + TS: D_ss (cols 0..63) = A.B^T with A from smem, D_ts (cols 64..127) with A copied to TMEM first.
__global__ void __launch_bounds__(128, 1) k_check(const unsigned char* a_img, const unsigned char* b_img, float* out_ss, float* out_ts) {
    extern __shared__ __align__(1024) unsigned char sm_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(sm_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char* sm = sm_raw + (base - smem_u32(sm_raw));
    for (int i = tid; i < ACT_PART / 16; i += blockDim.x) reinterpret_cast<uint4*>(sm)[i] = reinterpret_cast<const uint4*>(a_img)[i];
    for (int i = tid; i < 64 * 128 / 16; i += blockDim.x) reinterpret_cast<uint4*>(sm + SM_B)[i] = reinterpret_cast<const uint4*>(b_img)[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint64_t dA = make_desc(base + SM_A), dB = make_desc(base + SM_B);
        const uint32_t id1 = make_idesc(128, 64);
        for (int ks = 0; ks < 4; ++ks) mma_ss(tmem, dA + 2 * ks, dB + 2 * ks, id1, ks ? 1u : 0u);
        for (int ks = 0; ks < 4; ++ks) tcp_128x256(tmem + 256 + 8 * ks, dA + 2 * ks);
        for (int ks = 0; ks < 4; ++ks) mma_ts(tmem + 64, tmem + 256 + 8 * ks, dB + 2 * ks, id1, ks ? 1u : 0u);
        commit(&bar);
    }
    mbar_wait_bounded(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    {
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        for (int c0 = 0; c0 < 128; c0 += 16) {
            uint32_t r[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                           "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                         : "r"(taddr + c0));
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            float* dst = c0 < 64 ? out_ss : out_ts;
            for (int i = 0; i < 16; ++i) dst[(warp * 32 + lane) * 64 + (c0 & 63) + i] = __uint_as_float(r[i]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512u));
}

// ---- split-K reduce-scatter over a 4-CTA cluster: 256 threads, each CTA sends 3 x 8 KiB (128 rows x 16 fp32) ----
constexpr int STG = 4 * 128 * 16;     // bytes per sender: [4 column quads][128 rows][16 B]
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) {
    uint32_t o;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(o) : "r"(a), "r"(r));
    return o;
}
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(256, 1) k_xchg(int mode, int iters, long long* out) {
    extern __shared__ __align__(1024) unsigned char xsm[];
    unsigned char* recv = xsm;
    unsigned char* send = xsm + 3 * STG;
    __shared__ __align__(8) uint64_t full;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(rank));
    const int tid = threadIdx.x, row = tid & 127, hf = tid >> 7;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&full)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    float acc[32];
    for (int i = 0; i < 32; ++i) acc[i] = (float)(tid * 32 + i) * 1e-3f + (float)rank;
    float sum = 0.f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (tid == 0 && mode <= 1) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&full)), "r"(3 * STG) : "memory");
        if (mode == 0) {
            for (uint32_t p = 0; p < 4; ++p) {
                if (p == rank) continue;
                const uint32_t ss = rank < p ? rank : rank - 1;
                const uint32_t dst = mapa(smem_u32(recv) + ss * STG + (2 * hf) * 2048 + row * 16, p);
                const uint32_t bar = mapa(smem_u32(&full), p);
                asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];\n" ::"r"(dst),
                             "f"(acc[8 * p]), "f"(acc[8 * p + 1]), "f"(acc[8 * p + 2]), "f"(acc[8 * p + 3]), "r"(bar) : "memory");
                asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];\n" ::"r"(dst + 2048),
                             "f"(acc[8 * p + 4]), "f"(acc[8 * p + 5]), "f"(acc[8 * p + 6]), "f"(acc[8 * p + 7]), "r"(bar) : "memory");
            }
        } else if (mode == 1) {
            for (uint32_t p = 0; p < 4; ++p) {
                if (p == rank) continue;
                const uint32_t ss = p < rank ? p : p - 1;     // my send buffer for peer p
                unsigned char* d = send + ss * STG + (2 * hf) * 2048 + row * 16;
                *reinterpret_cast<float4*>(d) = make_float4(acc[8 * p], acc[8 * p + 1], acc[8 * p + 2], acc[8 * p + 3]);
                *reinterpret_cast<float4*>(d + 2048) = make_float4(acc[8 * p + 4], acc[8 * p + 5], acc[8 * p + 6], acc[8 * p + 7]);
            }
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            __syncthreads();
            if (tid < 3) {
                const uint32_t p = tid < (int)rank ? tid : tid + 1;          // peer index
                const uint32_t ss_mine = rank < p ? rank : rank - 1;          // my slot in the peer's recv buffer
                const uint32_t dst = mapa(smem_u32(recv) + ss_mine * STG, p);
                const uint32_t bar = mapa(smem_u32(&full), p);
                asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
                             "r"(smem_u32(send) + (uint32_t)tid * STG), "r"((uint32_t)STG), "r"(bar) : "memory");
            }
        }
        if (mode <= 1) {
            const long long tw = clock64();
            uint32_t ok = 0;
            while (!ok) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                             : "=r"(ok) : "r"(smem_u32(&full)), "r"((uint32_t)(it & 1)) : "memory");
                if (clock64() - tw > 1000000000LL) break;
            }
            for (int s = 0; s < 3; ++s) {
                const float4 a = *reinterpret_cast<const float4*>(recv + s * STG + (2 * hf) * 2048 + row * 16);
                const float4 b = *reinterpret_cast<const float4*>(recv + s * STG + (2 * hf) * 2048 + 2048 + row * 16);
                sum += a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
            }
        }
        // all four CTAs have read their buffers before anyone overwrites them
        __syncthreads();
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    }
    const long long t1 = clock64();
    if (tid == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = (long long)sum; }
}

static uint16_t f2bf(float f) {
    uint32_t u; memcpy(&u, &f, 4); u += 0x7fffu + ((u >> 16) & 1u); return (uint16_t)(u >> 16);
}

template <int mode>
void run_mma(int grid, int N, int reps, long long* out) {
    CK(cudaFuncSetAttribute(k_mma<mode>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
    k_mma<mode><<<grid, 128, SM_TOTAL>>>(N, reps, out);
    CK(cudaDeviceSynchronize());
}
void run_mma_mode(int mode, int grid, int N, int reps, long long* out) {
    switch (mode) {
        case 0: run_mma<0>(grid, N, reps, out); break;
        case 1: run_mma<1>(grid, N, reps, out); break;
        case 2: run_mma<2>(grid, N, reps, out); break;
        case 3: run_mma<3>(grid, N, reps, out); break;
        case 4: run_mma<4>(grid, N, reps, out); break;
        default: run_mma<5>(grid, N, reps, out); break;
    }
}
template <int mode>
void run_cp(int N, int reps, long long* out) {
    CK(cudaFuncSetAttribute(k_cp<mode>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
    k_cp<mode><<<1, 128, SM_TOTAL>>>(N, reps, out);
    CK(cudaDeviceSynchronize());
}

int main(int argc, char** argv) {
    const bool only_xchg = argc > 1;
    long long* d_out;
    CK(cudaMalloc(&d_out, 4096));
    long long h[64];
    CK(cudaFuncSetAttribute(k_check, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
    if (!only_xchg) {
    const char* names[6] = {"SS 3-product", "SS stacked 2N+N", "TS 3-product", "TS stacked 2N+N", "SS single", "TS single"};
    const int reps = 8, ksteps = reps * 16;
    printf("== tcgen05.mma, M=128 K=16, clk per k16 step (and per MMA), %d k-steps back to back, 1 CTA ==\n", ksteps);
    for (int mode = 0; mode < 6; ++mode)
        for (int N : {16, 32, 48, 64, 96, 128, 256}) {
            if ((mode == 1 || mode == 3) && 2 * N > 256) continue;
            for (int pass = 0; pass < 2; ++pass) run_mma_mode(mode, 1, N, reps, d_out);
            CK(cudaMemcpy(h, d_out, 24, cudaMemcpyDeviceToHost));
            const int per = (mode == 0 || mode == 2) ? 3 : (mode == 1 || mode == 3) ? 2 : 1;
            printf("%-18s N=%3d : total %7lld clk, %6.1f clk/k-step, %6.1f clk/MMA, issue %6.1f clk/k-step\n", names[mode], N, h[0],
                   (double)h[0] / ksteps, (double)h[0] / ksteps / per, (double)h[1] / ksteps);
        }
    printf("== the same on 148 CTAs at once (SS stacked / TS stacked, N=64): max over CTAs ==\n");
    for (int mode : {1, 3}) {
        long long* d_big; CK(cudaMalloc(&d_big, 148 * 24));
        for (int pass = 0; pass < 2; ++pass) run_mma_mode(mode, 148, 64, reps, d_big);
        std::vector<long long> hb(148 * 3);
        CK(cudaMemcpy(hb.data(), d_big, 148 * 24, cudaMemcpyDeviceToHost));
        long long mx = 0, mn = 1LL << 60;
        for (int i = 0; i < 148; ++i) { mx = hb[3 * i] > mx ? hb[3 * i] : mx; mn = hb[3 * i] < mn ? hb[3 * i] : mn; }
        printf("%-18s min %lld max %lld clk -> %.1f .. %.1f clk/k-step\n", names[mode], mn, mx, (double)mn / ksteps, (double)mx / ksteps);
        cudaFree(d_big);
    }
    printf("== tcgen05.cp 128x256b: 8 copies = one 32 KiB chunk (hi + lo) ==\n");
    for (int mode = 0; mode < 3; ++mode)
        for (int N : {48, 64}) {
            if (mode == 0 && N != 64) continue;
            for (int pass = 0; pass < 2; ++pass) { if (mode == 0) run_cp<0>(N, reps, d_out); else if (mode == 1) run_cp<1>(N, reps, d_out); else run_cp<2>(N, reps, d_out); }
            CK(cudaMemcpy(h, d_out, 24, cudaMemcpyDeviceToHost));
            printf("%-28s N=%3d: total %7lld clk, %7.1f clk per chunk (copy%s), issue %6.1f\n",
                   mode == 0 ? "copy only" : mode == 1 ? "copy + TS stacked" : "copy + TS 3-product", N, h[0], (double)h[0] / (reps * 4),
                   mode ? " + 4 k-steps" : "", (double)h[1] / (reps * 4));
        }
    {   // correctness of copy + TS
        std::vector<float> A(128 * 64), B(64 * 64);
        srand(3);
        for (auto& v : A) v = (float)(rand() % 17 - 8) / 8.f;
        for (auto& v : B) v = (float)(rand() % 13 - 6) / 4.f;
        std::vector<unsigned char> ai(128 * 128, 0), bi(64 * 128, 0);
        for (int r = 0; r < 128; ++r)
            for (int k = 0; k < 64; ++k) {
                const uint16_t v = f2bf(A[r * 64 + k]);
                memcpy(&ai[(size_t)r * 128 + ((((size_t)k >> 3) ^ (size_t)(r & 7)) << 4) + (k & 7) * 2], &v, 2);
            }
        for (int r = 0; r < 64; ++r)
            for (int k = 0; k < 64; ++k) {
                const uint16_t v = f2bf(B[r * 64 + k]);
                memcpy(&bi[(size_t)r * 128 + ((((size_t)k >> 3) ^ (size_t)(r & 7)) << 4) + (k & 7) * 2], &v, 2);
            }
        unsigned char *da, *db; float *o1, *o2;
        CK(cudaMalloc(&da, ai.size())); CK(cudaMalloc(&db, bi.size())); CK(cudaMalloc(&o1, 128 * 64 * 4)); CK(cudaMalloc(&o2, 128 * 64 * 4));
        CK(cudaMemcpy(da, ai.data(), ai.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(db, bi.data(), bi.size(), cudaMemcpyHostToDevice));
        CK(cudaMemset(o1, 0, 128 * 64 * 4)); CK(cudaMemset(o2, 0, 128 * 64 * 4));
        k_check<<<1, 128, SM_TOTAL>>>(da, db, o1, o2);
        CK(cudaDeviceSynchronize());
        std::vector<float> h1(128 * 64), h2(128 * 64);
        CK(cudaMemcpy(h1.data(), o1, h1.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(h2.data(), o2, h2.size() * 4, cudaMemcpyDeviceToHost));
        double e1 = 0, e2 = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 64; ++n) {
                double ref = 0;
                for (int k = 0; k < 64; ++k) ref += (double)A[m * 64 + k] * B[n * 64 + k];
                e1 = fmax(e1, fabs(ref - h1[m * 64 + n]));
                e2 = fmax(e2, fabs(ref - h2[m * 64 + n]));
            }
        printf("== copy + TS correctness: max err SS %.4f, tcgen05.cp + TS %.4f ==\n", e1, e2);
    }
    }
    printf("== reduce-scatter over a 4-CTA cluster, 256 threads, 24 KiB out / in per CTA and round ==\n");
    const int iters = 200;
    CK(cudaFuncSetAttribute(k_xchg, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * STG));
    for (int mode : {2, 0, 1}) {
        for (int pass = 0; pass < 2; ++pass) { k_xchg<<<4, 256, 6 * STG>>>(mode, iters, d_out); CK(cudaDeviceSynchronize()); }
        CK(cudaMemcpy(h, d_out, 64, cudaMemcpyDeviceToHost));
        printf("%-34s %7.1f clk per round (CTA 0)\n", mode == 2 ? "cluster barrier only" : mode == 0 ? "st.async from registers + barrier" : "st.shared + bulk DSMEM copy + barrier",
               (double)h[0] / iters);
    }
    return 0;
}
