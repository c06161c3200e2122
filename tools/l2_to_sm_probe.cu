// Probe: how fast can ONE SM pull L2-resident data into shared memory, as a function of how many SMs do it at once?
//   mode 0: cp.async.bulk (4 x 32 KiB in flight, mbarrier complete_tx)      mode 1: ld.global.v4 by 256 threads (kept in registers)
//   mode 2: cp.async (LDGSTS.128) by 256 threads
//   nvcc -gencode arch=compute_100a,code=sm_100a -o build_tmp/l2probe tools/l2_to_sm_probe.cu && build_tmp/l2probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void fill(uint4* dst, size_t n, unsigned v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = make_uint4(v, v + 1, v + 2, (unsigned)i);
}

__global__ void __launch_bounds__(256, 1) probe(const uint4* __restrict__ src, size_t region_bytes, int iters, int mode, uint4* sink, int share, int rot, int stride) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ __align__(8) uint64_t bar[4];
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(s32(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    const size_t chunk = 32768, n_chunks = region_bytes / chunk;
    // share > 0: groups of `share` CTAs walk the SAME chunks in the same order (like the clusters of one row tile reading one
    // activation image); rot: each CTA of a group starts its walk `rot * rank` chunks later
    size_t pos = share > 0 ? ((size_t)(blockIdx.x / share) * 37 + (size_t)(blockIdx.x % share) * rot) % n_chunks : (size_t)blockIdx.x * 37 % n_chunks;
    uint4 acc = make_uint4(0, 0, 0, 0);
    if (mode == 0) {
        if (tid == 0) {
            for (int it = 0; it < iters; ++it) {
                const int slot = it & 3;
                if (it >= 4) {
                    uint32_t ok = 0;
                    while (!ok)
                        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                                     : "=r"(ok) : "r"(s32(&bar[slot])), "r"((uint32_t)(((it >> 2) - 1) & 1)) : "memory");
                }
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(s32(&bar[slot])), "r"((uint32_t)chunk) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(s32(sm + slot * chunk)),
                             "l"(reinterpret_cast<const unsigned char*>(src) + pos * chunk), "r"((uint32_t)chunk), "r"(s32(&bar[slot])) : "memory");
                pos = (pos + stride) % n_chunks;
            }
            for (int it = iters; it < iters + 4; ++it) {
                const int slot = it & 3;
                uint32_t ok = 0;
                while (!ok)
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                                 : "=r"(ok) : "r"(s32(&bar[slot])), "r"((uint32_t)(((it >> 2) - 1) & 1)) : "memory");
            }
        }
    } else if (mode == 1) {
        for (int it = 0; it < iters; ++it) {
            const uint4* p = src + pos * (chunk / 16);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const uint4 v = __ldg(p + q * 256 + tid);
                acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
            }
            pos = (pos + stride) % n_chunks;
        }
    } else {
        for (int it = 0; it < iters; ++it) {
            const uint4* p = src + pos * (chunk / 16);
            const int slot = it & 3;
#pragma unroll
            for (int q = 0; q < 8; ++q)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s32(sm + slot * chunk + (q * 256 + tid) * 16)), "l"(p + q * 256 + tid) : "memory");
            asm volatile("cp.async.commit_group;\n" ::: "memory");
            asm volatile("cp.async.wait_group 3;\n" ::: "memory");
            pos = (pos + stride) % n_chunks;
        }
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    }
    __syncthreads();
    if (acc.x == 0x12345678 && sink) sink[tid] = acc;
}


// Persistent write -> grid barrier -> read loop, like one phase of the recurrent kernel: every CTA (re)writes its 8 KiB
// slice of a 1 MiB activation image with generic stores, all CTAs meet at a counter, then every CTA bulk-copies 128 KiB of
// the image (groups of 16 CTAs read the same quarter).  Reports the time from passing the barrier to the last byte landing.
__global__ void __launch_bounds__(256, 1) phase_probe(uint4* img, unsigned* counter, int phases, int do_write, int use_ldg, float* out_us) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ __align__(8) uint64_t bar[4];
    const int tid = threadIdx.x, cta = blockIdx.x, n = gridDim.x;
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(s32(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    unsigned long long t_acc = 0;
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (int ph = 0; ph < phases; ++ph) {
        uint4* image = img + (size_t)(ph & 1) * (1u << 20) / 16;          // two images, alternating
        if (do_write == 1) {
            uint4* mine = image + (size_t)cta * 512;                         // 8 KiB slice, coalesced
            for (int i = tid; i < 512; i += 256) mine[i] = make_uint4(ph, cta, i, 7);
        } else if (do_write == 2) {
            // the recurrent kernel's epilogue pattern: CTA = (row tile, n-tile, K-quarter rank); thread = (row, column half)
            // writes one 16-byte piece of the hi part and one of the lo part of a swizzled 128-byte row
            const int mt = cta / 64, c = cta % 64, chunk = c / 4, rank = c % 4, row = tid & 127, hf = tid >> 7;
            unsigned char* base = reinterpret_cast<unsigned char*>(image) + (size_t)mt * 524288 + (size_t)chunk * 32768 + row * 128 +
                                  (((rank * 2 + hf) ^ (row & 7)) << 4);
            *reinterpret_cast<uint4*>(base) = make_uint4(ph, cta, tid, 7);
            *reinterpret_cast<uint4*>(base + 16384) = make_uint4(ph, cta, tid, 9);
        }
        __syncthreads();
        if (tid == 0) {
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(counter) : "memory");
            const unsigned target = (unsigned)(ph + 1) * n;
            unsigned v;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(counter) : "memory"); } while (v < target);
        }
        __syncthreads();
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t0));
        const unsigned char* src = reinterpret_cast<const unsigned char*>(image) + (size_t)(cta / 64) * 524288 + (size_t)(cta % 4) * 131072;
        if (!use_ldg) {
            if (tid == 0) {
                asm volatile("fence.proxy.async.global;\n" ::: "memory");
                for (int c = 0; c < 4; ++c) {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(s32(&bar[c])), "r"(32768u) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(s32(sm + c * 32768)),
                                 "l"(src + c * 32768), "r"(32768u), "r"(s32(&bar[c])) : "memory");
                }
                for (int c = 0; c < 4; ++c) {
                    uint32_t ok = 0;
                    while (!ok)
                        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                                     : "=r"(ok) : "r"(s32(&bar[c])), "r"((uint32_t)(ph & 1)) : "memory");
                }
            }
        } else {
            const uint4* p = reinterpret_cast<const uint4*>(src);
#pragma unroll 8
            for (int q = 0; q < 32; ++q) {
                uint4 v;
                asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + q * 256 + tid));
                acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
            }
        }
        __syncthreads();
        asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t1));
        if (ph >= 4) t_acc += t1 - t0;
    }
    if (acc.x == 0x12345678) img[tid] = acc;
    if (tid == 0) out_us[cta] = (float)t_acc / (phases - 4) / 1e3f;
}

int main() {
    const size_t region = 48u << 20;   // fits in L2 (126 MB)
    uint4* src;
    cudaMalloc(&src, region);
    cudaMemset(src, 1, region);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 2000;
    const char* names[3] = {"cp.async.bulk 4x32K", "ld.global.v4 x256thr", "cp.async 16B x256thr"};
    for (int mode = 0; mode < 3; ++mode)
        for (int ctas : {1, 64, 148}) {
            probe<<<ctas, 256, 131072>>>(src, region, 200, mode, nullptr, 0, 0, 149);   // warm L2
            cudaEventRecord(e0);
            probe<<<ctas, 256, 131072>>>(src, region, iters, mode, nullptr, 0, 0, 149);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            const double bytes = (double)ctas * iters * 32768.0;
            printf("%-22s ctas %3d: %7.1f GB/s per SM, %7.2f TB/s total  [%s]\n", names[mode], ctas, bytes / ctas / ms / 1e6, bytes / ms / 1e9,
                   cudaGetErrorString(cudaGetLastError()));
        }
    // shared walks: 128 CTAs, groups of `share` CTAs read the same chunks in lock step; a small hot region like the activations
    for (int share : {16})
        for (int rot : {0, 1, 7}) {
            const size_t hot = 4u << 20;
            probe<<<128, 256, 131072>>>(src, hot, 200, 0, nullptr, share, rot, 1);
            cudaEventRecord(e0);
            probe<<<128, 256, 131072>>>(src, hot, iters, 0, nullptr, share, rot, 1);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            printf("bulk, 128 CTAs, 4 MB region, groups of %2d share a walk, rotation %d: %7.1f GB/s per SM\n", share, rot, iters * 32768.0 / ms / 1e6);
        }
    // freshly written data: a fill kernel rewrites 16 MB, then every CTA reads 128 KB of it once (4 chunks), 16 CTAs share each chunk set
    for (int rep = 0; rep < 0; ++rep)
        for (int fresh : {0, 1}) {
            const size_t hot = 16u << 20;
            float tot = 0;
            for (int r = 0; r < 20; ++r) {
                if (fresh) fill<<<592, 256>>>(src, hot / 16, r);
                cudaEventRecord(e0);
                probe<<<128, 256, 131072>>>(src, hot, 4, 0, nullptr, 16, 0, 1);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                tot += ms;
            }
            printf("one-shot 128 KB per CTA (128 CTAs, 16 share), %s: %.2f us per launch (incl. ~2-3 us launch)\n", fresh ? "freshly written" : "resident", tot / 20 * 1e3);
        }
    {
        uint4* img;
        unsigned* counter;
        float* out;
        cudaMalloc(&img, 2u << 20);
        cudaMalloc(&counter, 4);
        cudaMalloc(&out, 148 * 4);
        cudaFuncSetAttribute(phase_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
        for (int use_ldg : {0, 1})
            for (int do_write : {0, 1, 2})
                for (int ctas : {64, 128}) {
                    cudaMemset(counter, 0, 4);
                    void* args[] = {&img, &counter, nullptr, &do_write, &use_ldg, &out};
                    int phases = 200;
                    args[2] = &phases;
                    cudaLaunchCooperativeKernel((void*)phase_probe, dim3(ctas), dim3(256), args, 131072, 0);
                    cudaDeviceSynchronize();
                    float h[148];
                    cudaMemcpy(h, out, ctas * 4, cudaMemcpyDeviceToHost);
                    float mx = 0, mean = 0;
                    for (int i = 0; i < ctas; ++i) { mx = h[i] > mx ? h[i] : mx; mean += h[i] / ctas; }
                    printf("phase loop, %s, %3d CTAs, image %s: 128 KiB per CTA lands in %.2f us mean, %.2f us max  [%s]\n",
                           use_ldg ? "ld.global.cg" : "cp.async.bulk", ctas, do_write == 2 ? "rewritten in scattered 16-byte pieces" : do_write ? "rewritten every phase" : "read-only", mean, mx,
                           cudaGetErrorString(cudaGetLastError()));
                }
    }
    return 0;
}
