// Hoisted Linear layers of the BVRNN coder on the 5th-gen tensor cores:  C = epilogue(A . W^T) over all B*T frames.
//
// These are the state-independent layers of reference bvrnn.py:178,189 (encode: phi_x(y) and the phi_x half of
// enc.0) and :222-224 (decode: phi_z(z) and the phi_z halves of dec.0 and the GRU's W_ih).  Operands are the same
// ready-made shared-memory images the persistent recurrent kernel uses (recurrent.cuh: 64-wide K chunks, split bf16
// hi + lo, SWIZZLE_128B), so a chain of layers never leaves that format: the epilogue of one GEMM writes the
// activation images the next one bulk-copies.
//   tile       128 rows x 256 columns, K streamed in 64-wide chunks: 32 KiB of activations + 64 KiB of weights per stage,
//              2 stages; PERSISTENT: one CTA per SM walks the tiles (n fastest: the CTAs that share an activation tile run
//              together and it is read from HBM once), two TMEM accumulators of 256 columns
//   roles      warp 1: cp.async.bulk producer; warp 0: tcgen05.mma issuer (M = 128, N = 256, three MMAs per k16 step for
//              the split-bf16 product); warps 4-7: TMEM -> registers -> bias / ELU -> output of tile i under the main loop
//              of tile i + 1
//   output     activation images for the next layer, or fp32 rows (a thread writes 64 contiguous bytes of its row per step)
#include <cuda_bf16.h>

#include "common.cuh"
#include "recurrent.cuh"

namespace bvc {

namespace {

using rec::ACT_CHUNK_BYTES;
using rec::ACT_PART_BYTES;

constexpr int kThreads = 256;
constexpr int TILE_N = 256;
static_assert(kThreads == TILE_N, "one thread stages one bias value");
constexpr int W_PART_BYTES = TILE_N * 128;
constexpr int W_CHUNK_BYTES = 2 * W_PART_BYTES;
constexpr int STAGE_BYTES = ACT_CHUNK_BYTES + W_CHUNK_BYTES;     // 96 KiB
constexpr int STAGES = 2;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;

struct GemmArgs {
    const unsigned char* a_img;    // [m_tiles][k_chunks] activation images
    const unsigned char* w_img;    // [n_tiles][k_chunks] weight images, bn = 256
    const float* bias;             // [n_tiles * 256] or null
    float* out_f;                  // fp32 [M][ldo] or null
    unsigned char* out_img;        // [m_tiles][out_kchunks] or null
    int k_chunks, out_kchunks, ldo, M, N, act;
    int n_tiles_n, total_tiles;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded: a protocol bug traps (reported as a launch failure by the next CUDA call) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t b = smem_u32(bar);
    if (mbar_try(b, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try(b, parity)) {
        if (clock64() - t0 > 4000000000LL) asm volatile("trap;\n");
    }
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {   // K-major, SWIZZLE_128B, SBO = 1024 B
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {        // D = f32, A = B = bf16, K-major, dense
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void split_pair(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
    const __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0));
    const __nv_bfloat16 l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
    hi = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    lo = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
}
// 16 consecutive values of activation row `row` (columns col0 .. col0+15) -> image of the consumer layer
__device__ __forceinline__ void store_img16(unsigned char* img, size_t m_tile, int kchunks, int row, int col0, const float* v) {
    unsigned char* base = img + (m_tile * kchunks + (col0 >> 6)) * ACT_CHUNK_BYTES + row * 128;
    const int ch = (col0 & 63) >> 3;
    uint4 h0, l0, h1, l1;
    split_pair(v[0], v[1], h0.x, l0.x);   split_pair(v[2], v[3], h0.y, l0.y);
    split_pair(v[4], v[5], h0.z, l0.z);   split_pair(v[6], v[7], h0.w, l0.w);
    split_pair(v[8], v[9], h1.x, l1.x);   split_pair(v[10], v[11], h1.y, l1.y);
    split_pair(v[12], v[13], h1.z, l1.z); split_pair(v[14], v[15], h1.w, l1.w);
    const int p0 = (ch ^ (row & 7)) << 4, p1 = ((ch + 1) ^ (row & 7)) << 4;
    *reinterpret_cast<uint4*>(base + p0) = h0;
    *reinterpret_cast<uint4*>(base + p1) = h1;
    *reinterpret_cast<uint4*>(base + ACT_PART_BYTES + p0) = l0;
    *reinterpret_cast<uint4*>(base + ACT_PART_BYTES + p1) = l1;
}

// Persistent: one CTA per SM walks the tiles (n fastest, so the CTAs that share an activation tile run together); two TMEM
// accumulators, so the epilogue of tile i (tensor memory -> bias / ELU -> images or fp32 rows) runs under the main loop of
// tile i + 1, and TMEM allocation, barrier set-up and the first-chunk latency are paid once per CTA instead of once per
// tile (the one-tile-per-CTA version spent ~35 % of a tile outside the MMAs).
__global__ void __launch_bounds__(kThreads, 1) linear_umma_kernel(GemmArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t full[STAGES], empty[STAGES], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float bias_s[2][TILE_N];    // per accumulator: the tile's bias (no global load between two TMEM reads)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    const int KC = a.k_chunks;
    const int n_tiles_n = a.n_tiles_n, total = a.total_tiles;

    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_slot)),
                     "r"((uint32_t)(2 * TILE_N)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp == 1) {
        if (lane == 0) {
            uint32_t kcnt = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
                const size_t m_tile = (size_t)(tile / n_tiles_n);
                const int n_tile = tile % n_tiles_n;
                const unsigned char* a_src = a.a_img + m_tile * KC * ACT_CHUNK_BYTES;
                const unsigned char* w_src = a.w_img + (size_t)n_tile * KC * W_CHUNK_BYTES;
                for (int kc = 0; kc < KC; ++kc, ++kcnt) {
                    const int s = kcnt % STAGES;
                    const uint32_t round = kcnt / STAGES;
                    if (round >= 1) mbar_wait(&empty[s], (round - 1) & 1);
                    mbar_expect_tx(&full[s], STAGE_BYTES);
                    bulk_g2s(smem_base + s * STAGE_BYTES, a_src + (size_t)kc * ACT_CHUNK_BYTES, ACT_CHUNK_BYTES, &full[s]);
                    bulk_g2s(smem_base + s * STAGE_BYTES + ACT_CHUNK_BYTES, w_src + (size_t)kc * W_CHUNK_BYTES, W_CHUNK_BYTES, &full[s]);
                }
            }
        }
        __syncwarp();
    } else if (warp == 0) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc(128, TILE_N);
            const uint64_t d0 = make_desc(smem_base);
            uint32_t kcnt = 0, tcnt = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++tcnt) {
                const uint32_t acc = tcnt & 1u;
                if (tcnt >= 2) mbar_wait(&acc_empty[acc], ((tcnt >> 1) - 1) & 1);     // the epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem + acc * TILE_N;
                for (int kc = 0; kc < KC; ++kc, ++kcnt) {
                    const int s = kcnt % STAGES;
                    mbar_wait(&full[s], (kcnt / STAGES) & 1);
                    tc_fence_after();
                    const uint64_t dah = d0 + (uint64_t)((s * STAGE_BYTES) >> 4);
                    const uint64_t dal = dah + (ACT_PART_BYTES >> 4);
                    const uint64_t dwh = dah + (ACT_CHUNK_BYTES >> 4);
                    const uint64_t dwl = dwh + (W_PART_BYTES >> 4);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {   // small terms first
                        umma(d_tmem, dal + 2 * ks, dwh + 2 * ks, idesc, (kc | ks) != 0 ? 1u : 0u);
                        umma(d_tmem, dah + 2 * ks, dwl + 2 * ks, idesc, 1u);
                        umma(d_tmem, dah + 2 * ks, dwh + 2 * ks, idesc, 1u);
                    }
                    umma_commit(&empty[s]);
                }
                umma_commit(&acc_full[acc]);
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ---- epilogue: warps 4-7 own one accumulator row each ----
        const int quad = warp & 3, row = quad * 32 + lane, et = tid - 128;
        uint32_t tcnt = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++tcnt) {
            const size_t m_tile = (size_t)(tile / n_tiles_n);
            const int n_tile = tile % n_tiles_n;
            const uint32_t acc = tcnt & 1u;
            bias_s[acc][et] = a.bias ? __ldg(a.bias + n_tile * TILE_N + et) : 0.f;
            bias_s[acc][et + 128] = a.bias ? __ldg(a.bias + n_tile * TILE_N + et + 128) : 0.f;
            asm volatile("bar.sync 1, 128;\n" ::: "memory");
            mbar_wait(&acc_full[acc], (tcnt >> 1) & 1);
            tc_fence_after();
            const size_t m = m_tile * 128 + row;
            const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + acc * TILE_N;
#pragma unroll 1
            for (int c0 = 0; c0 < TILE_N; c0 += 16) {
                float v[16];
                tmem_ld16(taddr + c0, v);
                const int col0 = n_tile * TILE_N + c0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 b = *(reinterpret_cast<const float4*>(&bias_s[acc][c0]) + i);
                    v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
                }
                if (a.act == 1) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = elu_fast(v[i]);
                } else if (a.act == 2) {      // Bernoulli probabilities of the prior head (bvrnn.py:68-73)
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = sigmoidf_(v[i]);
                }
                if (a.out_img && col0 < a.N) store_img16(a.out_img, m_tile, a.out_kchunks, row, col0, v);
                if (a.out_f && m < (size_t)a.M && col0 < a.N) {      // 64 contiguous bytes of the row: two full sectors
                    float4* dst = reinterpret_cast<float4*>(a.out_f + m * a.ldo + col0);
#pragma unroll
                    for (int i = 0; i < 4; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(&acc_empty[acc])) : "memory");
            }
        }
    }
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"((uint32_t)(2 * TILE_N)));
    }
}

// fp32 rows [M][K_in] (optionally normalised (x - mean) / std per column) -> activation images with K padded to `kchunks` * 64
__global__ void __launch_bounds__(256) to_image_kernel(const float* __restrict__ x, int M, int K_in, const float* __restrict__ mean,
                                                       const float* __restrict__ sd, unsigned char* __restrict__ img, int kchunks) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per 8 consecutive columns
    const int per_row = kchunks * 8;
    const size_t rows = ((size_t)M + 127) / 128 * 128;
    if (idx >= rows * per_row) return;
    const size_t m = idx / per_row;
    const int k0 = (int)(idx - m * per_row) * 8;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int k = k0 + i;
        float t = (m < (size_t)M && k < K_in) ? x[m * K_in + k] : 0.f;
        if (mean && m < (size_t)M && k < K_in) t = (t - mean[k]) / sd[k];
        v[i] = t;
    }
    uint4 hi, lo;
    split_pair(v[0], v[1], hi.x, lo.x); split_pair(v[2], v[3], hi.y, lo.y);
    split_pair(v[4], v[5], hi.z, lo.z); split_pair(v[6], v[7], hi.w, lo.w);
    const int row = (int)(m % 128);
    unsigned char* base = img + ((m / 128) * kchunks + (k0 >> 6)) * ACT_CHUNK_BYTES + row * 128 + ((((k0 & 63) >> 3) ^ (row & 7)) << 4);
    *reinterpret_cast<uint4*>(base) = hi;
    *reinterpret_cast<uint4*>(base + ACT_PART_BYTES) = lo;
}

}  // namespace

int to_image(const float* x, int M, int K_in, const float* mean, const float* sd, unsigned char* img, int kchunks,
             cudaStream_t stream) {
    const size_t total = ((size_t)M + 127) / 128 * 128 * kchunks * 8;
    to_image_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(x, M, K_in, mean, sd, img, kchunks);
    BVC_CHECK_LAUNCH();
    return BVC_OK;
}

int linear_umma(const unsigned char* a_img, int M, const WImg& w, const float* bias, int act, float* out_f, int ldo,
                unsigned char* out_img, cudaStream_t stream) {
    if (w.bn != TILE_N || M <= 0) { set_error("linear_umma: weight image must be packed with bn = 256"); return BVC_ERR_INVALID; }
    static bool attr_set[kMaxDevices] = {};
    const int dslot = device_slot();
    if (!attr_set[dslot]) {
        BVC_CUDA(cudaFuncSetAttribute(linear_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_set[dslot] = true;
    }
    GemmArgs a;
    a.a_img = a_img; a.w_img = w.img; a.bias = bias; a.out_f = out_f; a.out_img = out_img;
    a.k_chunks = (w.K + 63) / 64;
    a.out_kchunks = (w.N + 63) / 64;
    a.ldo = ldo; a.M = M; a.N = w.N; a.act = act;
    a.n_tiles_n = (w.N + TILE_N - 1) / TILE_N;
    a.total_tiles = a.n_tiles_n * ((M + 127) / 128);
    static int sms_of[kMaxDevices] = {};
    if (!sms_of[dslot]) {
        int dev = 0, n = 0;
        BVC_CUDA(cudaGetDevice(&dev));
        BVC_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
        sms_of[dslot] = n;
    }
    const int grid = a.total_tiles < sms_of[dslot] ? a.total_tiles : sms_of[dslot];
    linear_umma_kernel<<<grid, kThreads, SMEM_BYTES, stream>>>(a);
    BVC_CHECK_LAUNCH();
    return BVC_OK;
}

}  // namespace bvc
