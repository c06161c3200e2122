// Persistent recurrent kernel for the BVRNN time loop (5th-gen tensor cores, accumulators in TMEM,
// split-K over 4-CTA clusters with a distributed-shared-memory reduce-scatter).
//
// Replaces the per-frame Python loop of reference bvrnn.py:186-206 (encode) and :222-227 (decode),
// i.e. ~15 nn.Linear launches + ~250 small ATen kernels per frame, with ONE kernel that stays resident
// for all T frames.
//
// Why this shape.  A frame is a chain of 13 (encode) / 7 (decode) dependent [B x 1024] . [1024 x N] layers.
// The weights (80 MB as split bf16) only fit in L2, so every layer streams its operands L2 -> SM and the
// chain is bound by that traffic and by the barrier between layers, not by the tensor pipe.  An output
// tile that owns its full K reads (m + n) * K operand elements; splitting K four ways over a cluster
// and keeping the activation quarter resident while the CTA walks its n-tiles cuts the bytes per CTA
// and layer from 384 KB (previous 64 x 32 full-K tiles) to 128 KB of activations + 64 KB per n-tile:
//   cluster (4 CTAs)   one 128-row m-tile x a list of n-tiles per phase; CTA rank r owns K quarter r
//   operands           global memory holds ready-made shared-memory images (recurrent.cuh), moved by
//                      cp.async.bulk from ONE thread; completion lands on mbarriers (complete_tx)
//   MMA                one thread issues tcgen05.mma.cta_group::1.kind::f16, M = 128, N = 64 / 48 / 16,
//                      three MMAs per k16 step for the split-bf16 product a.w ~= a_lo.w_hi + a_hi.w_lo + a_hi.w_hi;
//                      8 accumulator slots of 64 TMEM columns let the MMAs of the next n-tile run under
//                      the epilogue of the previous one
//   reduce-scatter     epilogue thread = one row; it stores the 3 foreign column quarters of its partial
//                      row into the peers' staging buffers (st.shared::cluster), signals their mbarrier,
//                      then sums the 4 partials of its own quarter in fixed K order (deterministic)
//   fused epilogues    bias, hoisted addend, ELU, Bernoulli bottleneck (threshold, bit-budget mask,
//                      16 packed bits per thread), GRU gates; outputs are written straight into the
//                      next layer's activation image (split bf16, swizzled) and / or as fp32
//   phase barrier      one counter per m-tile (batch rows never interact): the epilogue leader does
//                      red.release.gpu, the copy thread of each consumer polls ld.acquire.gpu before it
//                      loads activations; weight chunks of the next phase are prefetched before the wait
#include <cuda_bf16.h>

#include "common.cuh"
#include "recurrent.cuh"

namespace bvc {
namespace rec {

namespace {

constexpr int kThreads = 256;
constexpr int A_SLOTS = 4;                                   // chunks of the resident activation quarter
constexpr int W_SLOTS = 3;
constexpr int W_SLOT_BYTES = 2 * 64 * 128;                   // bn <= 64
constexpr int ACC_SLOTS = 8;
constexpr int ACC_COLS = 64;
constexpr int TMEM_COLS = ACC_SLOTS * ACC_COLS;              // 512: the whole tensor memory of the SM
constexpr int STG_SENDER_BYTES = 4 * TILE_M * 16;            // [4 column quads][128 rows][16 bytes]
constexpr int STG_BUF_BYTES = (CLUSTER - 1) * STG_SENDER_BYTES;
constexpr int SMEM_A = 0;
constexpr int SMEM_W = SMEM_A + A_SLOTS * ACT_CHUNK_BYTES;
constexpr int SMEM_STG = SMEM_W + W_SLOTS * W_SLOT_BYTES;
constexpr int SMEM_TOTAL = SMEM_STG + 2 * STG_BUF_BYTES;     // 224 KiB
constexpr int SMEM_DYNAMIC = SMEM_TOTAL + 1024;              // manual 1024-byte alignment

struct Bars {
    uint64_t fullA[A_SLOTS];       // activation chunk landed (expect_tx)
    uint64_t fullW[W_SLOTS];       // weight chunk landed (expect_tx)
    uint64_t emptyW[W_SLOTS];      // weight chunk consumed (tcgen05.commit)
    uint64_t accFull[ACC_SLOTS];   // accumulator complete (tcgen05.commit)
    uint64_t accEmpty[ACC_SLOTS];  // accumulator drained (4 epilogue warps)
    uint64_t aFree;                // all MMAs of the job retired: activation buffer reusable
    uint64_t stgFull[2];           // the 3 peers wrote their partials into my staging buffer
    uint64_t stgEmpty[2];          // the 3 peers finished reading their staging buffer
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
__device__ __forceinline__ bool mbar_try_cluster(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded waits: a protocol bug or a lost peer becomes an abort flag instead of a hung GPU.  Every waiter
// also watches the flag, so one failing role releases all the others.
template <bool kCluster>
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int* abort_flag, int code) {
    const uint32_t b = smem_u32(bar);
    if (kCluster ? mbar_try_cluster(b, parity) : mbar_try(b, parity)) return true;
    const long long t0 = clock64();
    int spins = 0;
    while (!(kCluster ? mbar_try_cluster(b, parity) : mbar_try(b, parity))) {
        if (((++spins) & 63) == 0) {
            if (*(volatile int*)abort_flag) return false;
            if (clock64() - t0 > 3000000000LL) {
                atomicCAS(abort_flag, 0, code);
                return false;
            }
        }
    }
    return true;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void st_cluster_f4(uint32_t cluster_addr, float a, float b, float c, float d) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(cluster_addr), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}
__device__ __forceinline__ void fence_cluster() { asm volatile("fence.acq_rel.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;\n" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}

// global -> shared bulk copy (TMA engine, no tensor map); completion is counted in bytes on `bar`
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
// shared-memory matrix descriptor: K-major, SWIZZLE_128B (layout type 2), version 1 (Blackwell);
// LBO is unused for swizzled K-major operands (encoded 1), SBO = 1024 B = 8 rows x 128 B
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor: D = f32, A = B = bf16, both K-major, dense
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
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

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// The entries of (phase, cluster) this CTA executes: index e0 + i * stride, i < n.
struct JobView {
    int e0, n, stride;
};
__device__ __forceinline__ JobView my_entries(const Program* p, int ph, int cluster, int rank) {
    const int* es = p->entry_start + ph * (p->n_clusters + 1) + cluster;
    const int eb = es[0], cnt = es[1] - es[0];
    JobView v;
    if (p->phases[ph].split) {
        v.e0 = eb; v.n = cnt; v.stride = 1;
    } else {
        v.e0 = eb + rank; v.n = cnt > rank ? (cnt - rank + CLUSTER - 1) / CLUSTER : 0; v.stride = CLUSTER;
    }
    return v;
}

// 16 consecutive values of activation row `row` (columns col0 .. col0+15, col0 % 16 == 0) -> image of the
// consumer layer: two 16-byte pieces of the hi part and of the lo part
__device__ __forceinline__ void store_img16(unsigned char* img, int m_tile, int kchunks, int row, int col0, const float* v,
                                            bool zero_lo) {
    unsigned char* base = img + ((size_t)m_tile * kchunks + (col0 >> 6)) * ACT_CHUNK_BYTES + row * 128;
    const int ch = (col0 & 63) >> 3;
    uint4 h0, l0, h1, l1;
    split_pair(v[0], v[1], h0.x, l0.x);   split_pair(v[2], v[3], h0.y, l0.y);
    split_pair(v[4], v[5], h0.z, l0.z);   split_pair(v[6], v[7], h0.w, l0.w);
    split_pair(v[8], v[9], h1.x, l1.x);   split_pair(v[10], v[11], h1.y, l1.y);
    split_pair(v[12], v[13], h1.z, l1.z); split_pair(v[14], v[15], h1.w, l1.w);
    if (zero_lo) { l0 = make_uint4(0, 0, 0, 0); l1 = l0; }
    const int p0 = (ch ^ (row & 7)) << 4, p1 = ((ch + 1) ^ (row & 7)) << 4;
    *reinterpret_cast<uint4*>(base + p0) = h0;
    *reinterpret_cast<uint4*>(base + p1) = h1;
    *reinterpret_cast<uint4*>(base + ACT_PART_BYTES + p0) = l0;
    *reinterpret_cast<uint4*>(base + ACT_PART_BYTES + p1) = l1;
}

__device__ __forceinline__ void load16(const float* p, float* v, bool through_l2) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 a = through_l2 ? __ldcg(reinterpret_cast<const float4*>(p) + i) : __ldg(reinterpret_cast<const float4*>(p) + i);
        v[4 * i] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = a.z; v[4 * i + 3] = a.w;
    }
}
__device__ __forceinline__ void store16(float* p, const float* v) {
#pragma unroll
    for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}

// Epilogue of 16 output columns (col0 .. col0+15) of row m.  v = A.W^T summed over the whole K.
__device__ __forceinline__ void finalize16(const Op& op, const Frame& fr, int t, int m, int row, int m_tile, int col0,
                                           float* v) {
    if (m >= fr.M) return;
    float tmp[16];
    if (op.bias) {
        load16(op.bias + col0, tmp, false);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += tmp[i];
    }
    if (op.addend) {   // written by another CTA in an earlier phase or by an earlier kernel: read through L2
        load16(op.addend + (size_t)t * op.add_tstride + (size_t)m * op.ldadd + col0, tmp, true);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += tmp[i];
    }
    if (op.kind == KIND_BOTTLENECK) {
        // z = round(sigmoid(logit)), masked to 0.5 beyond the frame's bit budget (bvrnn.py:191-196)
        const float budget = fr.bits ? __ldg(fr.bits + (size_t)m * fr.T + t) : fr.bits_scalar;
        float code[16];
        uint32_t word = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const bool active = !fr.var_bit || (budget > (float)(col0 + i));
            const bool bit = active && (sigmoidf_(v[i]) > 0.5f);
            code[i] = active ? (bit ? 1.f : 0.f) : 0.5f;
            if (bit) word |= 1u << i;
        }
        const size_t o = ((size_t)m * fr.T + t) * fr.Z + col0;
        store16(fr.codes + o, code);
        if (fr.logits) store16(fr.logits + o, v);
        if (fr.packed) reinterpret_cast<unsigned short*>(fr.packed)[((size_t)m * fr.T + t) * 4 + (col0 >> 4)] = (unsigned short)word;
        store_img16(op.out_img, m_tile, op.out_kchunks, row, col0, code, true);   // {0, .5, 1} are exact in bf16
        return;
    }
    if (op.act) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = elu1(v[i]);
    }
    if (op.kind == KIND_MEL) {
        if (fr.mel_out) {
            float* mo = fr.mel_out + ((size_t)m * fr.T + t) * fr.X;
            if (col0 + 16 <= fr.X) store16(mo + col0, v);
            else
                for (int i = 0; i < 16; ++i)
                    if (col0 + i < fr.X) mo[col0 + i] = v[i];
        }
        return;
    }
    if (op.out_f && col0 < op.N) store16(op.out_f + (size_t)m * op.ldo + col0, v);
    if (op.out_img && col0 < op.N) store_img16(op.out_img, m_tile, op.out_kchunks, row, col0, v, false);
}

// GRU epilogue: 12 columns = [r(4) | z(4) | n(4)] of the hidden units u0 .. u0+3 (PyTorch gate order r,z,n;
// reference bvrnn.py:83,206):  r = s(gi_r + gh_r), z = s(gi_z + gh_z), n = tanh(gi_n + r * gh_n), h' = (h - n) z + n
__device__ __forceinline__ void finalize_gru(const Op& op, const Frame& fr, int t, int m, int row, int m_tile, int col0,
                                             int u0, const float* v) {
    if (m >= fr.M) return;
    const int H = fr.H;
    const float* gz = op.addend + (size_t)t * op.add_tstride + (size_t)m * op.ldadd + col0;
    const float* gh = fr.gh + (size_t)m * 3 * H + col0;
    float* hp = fr.h + (size_t)m * H + u0;
    const float4 zr = __ldcg(reinterpret_cast<const float4*>(gz)), zz = __ldcg(reinterpret_cast<const float4*>(gz + 4));
    const float4 zn = __ldcg(reinterpret_cast<const float4*>(gz + 8));
    const float4 hr = __ldcg(reinterpret_cast<const float4*>(gh)), hz = __ldcg(reinterpret_cast<const float4*>(gh + 4));
    const float4 hn4 = __ldcg(reinterpret_cast<const float4*>(gh + 8));
    const float4 hv4 = *reinterpret_cast<const float4*>(hp);     // only this thread ever touches these 4 states
    const float gzr[4] = {zr.x, zr.y, zr.z, zr.w}, gzz[4] = {zz.x, zz.y, zz.z, zz.w}, gzn[4] = {zn.x, zn.y, zn.z, zn.w};
    const float ghr[4] = {hr.x, hr.y, hr.z, hr.w}, ghz[4] = {hz.x, hz.y, hz.z, hz.w}, ghn[4] = {hn4.x, hn4.y, hn4.z, hn4.w};
    const float hv[4] = {hv4.x, hv4.y, hv4.z, hv4.w};
    float hn[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float rr = sigmoidf_(v[e] + gzr[e] + ghr[e]);
        const float zg = sigmoidf_(v[4 + e] + gzz[e] + ghz[e]);
        const float nn = tanhf(v[8 + e] + gzn[e] + rr * ghn[e]);
        hn[e] = (hv[e] - nn) * zg + nn;
    }
    if (fr.all_h)      // state entering frame t (bvrnn.py:205)
        *reinterpret_cast<float4*>(fr.all_h + ((size_t)m * fr.T + t) * H + u0) = hv4;
    *reinterpret_cast<float4*>(hp) = make_float4(hn[0], hn[1], hn[2], hn[3]);
    uint32_t h0, l0, h1, l1;
    split_pair(hn[0], hn[1], h0, l0);
    split_pair(hn[2], hn[3], h1, l1);
    unsigned char* base = fr.h_img + ((size_t)m_tile * (H >> 6) + (u0 >> 6)) * ACT_CHUNK_BYTES + row * 128 +
                          ((((u0 & 63) >> 3) ^ (row & 7)) << 4) + (u0 & 7) * 2;
    *reinterpret_cast<uint2*>(base) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(base + ACT_PART_BYTES) = make_uint2(l0, l1);
}

__global__ void __cluster_dims__(CLUSTER, 1, 1) __launch_bounds__(kThreads, 1)
recurrent_cluster_kernel(const Program* __restrict__ prog, unsigned* sync_words) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) Bars bars;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = (int)cluster_ctarank();
    const int cluster = blockIdx.x / CLUSTER;
    int* abort_flag = reinterpret_cast<int*>(sync_words);
    // identical in every CTA of the cluster (same static layout), so mapa() of a local address names the peer's copy
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;

    if (tid == 0) {
        for (int i = 0; i < A_SLOTS; ++i) mbar_init(&bars.fullA[i], 1);
        for (int i = 0; i < W_SLOTS; ++i) { mbar_init(&bars.fullW[i], 1); mbar_init(&bars.emptyW[i], 1); }
        for (int i = 0; i < ACC_SLOTS; ++i) { mbar_init(&bars.accFull[i], 1); mbar_init(&bars.accEmpty[i], 4); }
        mbar_init(&bars.aFree, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&bars.stgFull[i], 4 * (CLUSTER - 1)); mbar_init(&bars.stgEmpty[i], 4 * (CLUSTER - 1)); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_slot)),
                     "r"((uint32_t)TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // peers' mbarriers are initialised before anyone signals them
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    const Frame& fr = prog->frame;
    const int T = fr.T, n_phases = prog->n_phases;
    const int m_tile = prog->cluster_mtile[cluster];
    const unsigned dom_ctas = (unsigned)prog->mtile_ctas[m_tile];
    unsigned* counter = sync_words + 32 * (1 + m_tile);

    if (warp == 1) {
        // =========================== copy thread ===========================
        if (lane == 0) {
            uint32_t wIt = 0, aIt = 0;
            bool dead = false;
            for (int t = 0; t < T && !dead; ++t) {
                for (int ph = 0; ph < n_phases && !dead; ++ph) {
                    const Phase& phs = prog->phases[ph];
                    const JobView jv = my_entries(prog, ph, cluster, rank);
                    const int nck = phs.split ? phs.k_chunks / CLUSTER : phs.k_chunks;
                    const int kc0 = phs.split ? rank * nck : 0;
                    const int nW = jv.n * nck;
                    auto issue_w = [&](int i) {
                        const int j = i / nck, c = i - j * nck;
                        const uint32_t e = prog->entries[jv.e0 + j * jv.stride];
                        const Op& op = prog->ops[e >> 16];
                        const int nt = (int)(e & 0xFFFF);
                        const int slot = wIt % W_SLOTS, round = wIt / W_SLOTS;
                        if (round >= 1 && !mbar_wait<false>(&bars.emptyW[slot], (round - 1) & 1, abort_flag, 11)) { dead = true; return; }
                        const uint32_t bytes = 2u * op.bn * 128u;
                        const unsigned char* src = op.w_img + ((size_t)nt * phs.k_chunks + kc0 + c) * bytes;
                        mbar_expect_tx(&bars.fullW[slot], bytes);
                        bulk_g2s(smem_base + SMEM_W + slot * W_SLOT_BYTES, src, bytes, &bars.fullW[slot]);
                        ++wIt;
                    };
                    // weights do not depend on the previous phase: start them before waiting for it
                    const int pre = nW < W_SLOTS ? nW : W_SLOTS;
                    for (int i = 0; i < pre && !dead; ++i) issue_w(i);
                    if (jv.n > 0 && !dead) {
                        const unsigned target = (unsigned)(t * n_phases + ph) * dom_ctas;
                        if (ld_acquire(counter) < target) {
                            const long long t0 = clock64();
                            int spins = 0;
                            while (ld_acquire(counter) < target) {
                                if (((++spins) & 15) == 0) {
                                    if (*(volatile int*)abort_flag) { dead = true; break; }
                                    if (clock64() - t0 > 4000000000LL) { atomicCAS(abort_flag, 0, 1); dead = true; break; }
                                }
                            }
                        }
                        if (prog->debug_flags & 16) { const long long d0 = clock64(); while (clock64() - d0 < 100000) {} }
                        fence_proxy_async_all();   // the producers' generic-proxy stores -> this thread's async-proxy reads
                        if (!dead && aIt >= 1 && !mbar_wait<false>(&bars.aFree, (aIt - 1) & 1, abort_flag, 12)) dead = true;
                        if (!dead) {
                            const unsigned char* src = phs.a_img + ((size_t)m_tile * phs.k_chunks + kc0) * ACT_CHUNK_BYTES;
                            for (int c = 0; c < nck; ++c) {
                                mbar_expect_tx(&bars.fullA[c], ACT_CHUNK_BYTES);
                                bulk_g2s(smem_base + SMEM_A + c * ACT_CHUNK_BYTES, src + (size_t)c * ACT_CHUNK_BYTES, ACT_CHUNK_BYTES,
                                         &bars.fullA[c]);
                            }
                            ++aIt;
                        }
                    }
                    for (int i = pre; i < nW && !dead; ++i) issue_w(i);
                }
            }
        }
        __syncwarp();
    } else if (warp == 0) {
        // =========================== MMA thread ===========================
        if (lane == 0) {
            uint32_t wIt = 0, accIt = 0;
            uint32_t aUse[A_SLOTS] = {0, 0, 0, 0};
            bool dead = false;
            const uint64_t descA = make_desc(smem_base + SMEM_A), descW = make_desc(smem_base + SMEM_W);
            for (int t = 0; t < T && !dead; ++t) {
                for (int ph = 0; ph < n_phases && !dead; ++ph) {
                    const Phase& phs = prog->phases[ph];
                    const JobView jv = my_entries(prog, ph, cluster, rank);
                    const int nck = phs.split ? phs.k_chunks / CLUSTER : phs.k_chunks;
                    for (int j = 0; j < jv.n && !dead; ++j) {
                        const uint32_t e = prog->entries[jv.e0 + j * jv.stride];
                        const int bn = prog->ops[e >> 16].bn;
                        const int slot = accIt % ACC_SLOTS, round = accIt / ACC_SLOTS;
                        if (round >= 1 && !mbar_wait<false>(&bars.accEmpty[slot], (round - 1) & 1, abort_flag, 21)) { dead = true; break; }
                        tc_fence_after();
                        const uint32_t idesc = make_idesc(TILE_M, bn);
                        const uint32_t d_tmem = tmem + slot * ACC_COLS;
                        for (int c = 0; c < nck; ++c) {
                            if (j == 0) {
                                if (!mbar_wait<false>(&bars.fullA[c], aUse[c] & 1, abort_flag, 22)) { dead = true; break; }
                                ++aUse[c];
                            }
                            const int ws = wIt % W_SLOTS, wr = wIt / W_SLOTS;
                            if (!mbar_wait<false>(&bars.fullW[ws], wr & 1, abort_flag, 23)) { dead = true; break; }
                            ++wIt;
                            tc_fence_after();
                            const uint64_t dah = descA + (uint64_t)((c * ACT_CHUNK_BYTES) >> 4);
                            const uint64_t dal = dah + (ACT_PART_BYTES >> 4);
                            const uint64_t dwh = descW + (uint64_t)((ws * W_SLOT_BYTES) >> 4);
                            const uint64_t dwl = dwh + (uint64_t)((bn * 128) >> 4);
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {   // small terms first
                                umma(d_tmem, dal + 2 * ks, dwh + 2 * ks, idesc, (c | ks) != 0 ? 1u : 0u);
                                umma(d_tmem, dah + 2 * ks, dwl + 2 * ks, idesc, 1u);
                                umma(d_tmem, dah + 2 * ks, dwh + 2 * ks, idesc, 1u);
                            }
                            umma_commit(&bars.emptyW[ws]);
                        }
                        if (dead) break;
                        umma_commit(&bars.accFull[slot]);
                        ++accIt;
                    }
                    if (jv.n > 0 && !dead) umma_commit(&bars.aFree);
                }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // =========================== epilogue warps ===========================
        const int quad = warp & 3;                     // TMEM lane quadrant this warp may access
        const int row = quad * 32 + lane;              // row of the m-tile this thread owns
        const int m = m_tile * TILE_M + row;
        const uint32_t t_lane = tmem + ((uint32_t)(quad * 32) << 16);
        const uint32_t stg_local = smem_base + SMEM_STG;
        uint32_t accIt = 0, sIt = 0;
        bool dead = false;
        for (int t = 0; t < T && !dead; ++t) {
            for (int ph = 0; ph < n_phases && !dead; ++ph) {
                const Phase& phs = prog->phases[ph];
                const JobView jv = my_entries(prog, ph, cluster, rank);
                for (int j = 0; j < jv.n && !dead; ++j) {
                    const uint32_t e = prog->entries[jv.e0 + j * jv.stride];
                    const Op& op = prog->ops[e >> 16];
                    const int nt = (int)(e & 0xFFFF);
                    const int slot = accIt % ACC_SLOTS, round = accIt / ACC_SLOTS;
                    if (!mbar_wait<false>(&bars.accFull[slot], round & 1, abort_flag, 31)) { dead = true; break; }
                    ++accIt;
                    tc_fence_after();
                    const uint32_t taddr = t_lane + slot * ACC_COLS;
                    float v[16];
                    int col0;
                    if (phs.split) {
                        const int qc = op.bn >> 2;                       // columns per CTA after the reduce-scatter: 16 or 12
                        const int buf = sIt & 1, sr = sIt >> 1;
                        ++sIt;
                        // my staging slot at every peer is free once all peers have read their copy of two tiles ago
                        if (sr >= 1 && !mbar_wait<false>(&bars.stgEmpty[buf], (sr - 1) & 1, abort_flag, 32)) { dead = true; break; }
                        float own[16];
#pragma unroll
                        for (int p = 0; p < CLUSTER; ++p) {
                            if (p == rank) {
                                tmem_ld16(taddr + qc * p, own);
                            } else {
                                float tmp[16];
                                tmem_ld16(taddr + qc * p, tmp);
                                const int ss = rank < p ? rank : rank - 1;
                                const uint32_t dst =
                                    map_to_cta(stg_local + buf * STG_BUF_BYTES + ss * STG_SENDER_BYTES + row * 16, (uint32_t)p);
                                st_cluster_f4(dst, tmp[0], tmp[1], tmp[2], tmp[3]);
                                st_cluster_f4(dst + 2048, tmp[4], tmp[5], tmp[6], tmp[7]);
                                st_cluster_f4(dst + 4096, tmp[8], tmp[9], tmp[10], tmp[11]);
                                if (qc == 16) st_cluster_f4(dst + 6144, tmp[12], tmp[13], tmp[14], tmp[15]);
                            }
                        }
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            mbar_arrive(&bars.accEmpty[slot]);
                            fence_cluster();
#pragma unroll
                            for (int p = 0; p < CLUSTER; ++p)
                                if (p != rank) mbar_arrive_remote(map_to_cta(smem_u32(&bars.stgFull[buf]), (uint32_t)p));
                        }
                        if (!mbar_wait<true>(&bars.stgFull[buf], sr & 1, abort_flag, 33)) { dead = true; break; }
                        // sum the four K quarters in fixed order 0,1,2,3 (bit-reproducible)
#pragma unroll
                        for (int p = 0; p < CLUSTER; ++p) {
                            float part[16];
                            if (p == rank) {
#pragma unroll
                                for (int i = 0; i < 16; ++i) part[i] = own[i];
                            } else {
                                const int ss = p < rank ? p : p - 1;
                                const unsigned char* src = smem_dyn + (smem_base - smem_u32(smem_dyn)) + SMEM_STG + buf * STG_BUF_BYTES +
                                                           ss * STG_SENDER_BYTES + row * 16;
#pragma unroll
                                for (int q4 = 0; q4 < 4; ++q4) {
                                    if (q4 < 3 || qc == 16) {
                                        const float4 a = *reinterpret_cast<const float4*>(src + q4 * 2048);
                                        part[4 * q4] = a.x; part[4 * q4 + 1] = a.y; part[4 * q4 + 2] = a.z; part[4 * q4 + 3] = a.w;
                                    } else {
                                        part[12] = part[13] = part[14] = part[15] = 0.f;
                                    }
                                }
                            }
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] = (p == 0) ? part[i] : v[i] + part[i];
                        }
                        __syncwarp();
                        if (lane == 0) {
#pragma unroll
                            for (int p = 0; p < CLUSTER; ++p)
                                if (p != rank) mbar_arrive_remote(map_to_cta(smem_u32(&bars.stgEmpty[buf]), (uint32_t)p));
                        }
                        col0 = nt * op.bn + qc * rank;
                    } else {
                        tmem_ld16(taddr, v);
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bars.accEmpty[slot]);
                        col0 = nt * op.bn;
                    }
                    if (op.kind == KIND_GRU) finalize_gru(op, fr, t, m, row, m_tile, col0, nt * 16 + 4 * rank, v);
                    else finalize16(op, fr, t, m, row, m_tile, col0, v);
                }
                // ---- end of phase: publish this CTA's outputs to the m-tile's barrier domain ----
                if (prog->debug_flags & 32) __threadfence();
                fence_proxy_async_all();                                  // they are read by other CTAs' bulk copies
                {   // barrier over the 4 epilogue warps; a failed wait anywhere retires all of them together
                    uint32_t any;
                    asm volatile(
                        "{\n\t.reg .pred p, q;\n\t"
                        "setp.ne.u32 q, %1, 0;\n\t"
                        "bar.red.or.pred p, 1, 128, q;\n\t"
                        "selp.u32 %0, 1, 0, p;\n\t}\n"
                        : "=r"(any)
                        : "r"((uint32_t)dead)
                        : "memory");
                    dead = any != 0;
                }
                if (tid == 128 && !(prog->debug_flags & 8)) {
                    // The counter is monotonic over all phases, so nobody may arrive for phase p before every CTA
                    // of the domain has arrived for phase p - 1.  A CTA with work in phase p got that from its
                    // copy thread (which waited for it before loading activations); an idle CTA waits here.
                    if (jv.n == 0) {
                        const unsigned target = (unsigned)(t * n_phases + ph) * dom_ctas;
                        const long long t0 = clock64();
                        int spins = 0;
                        while (ld_acquire(counter) < target) {
                            if (((++spins) & 15) == 0 &&
                                (*(volatile int*)abort_flag || clock64() - t0 > 4000000000LL)) {
                                atomicCAS(abort_flag, 0, 2);
                                break;
                            }
                        }
                    }
                    asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(counter) : "memory");
                }
            }
        }
    }

    __syncthreads();
    cluster_sync_all();          // no CTA exits (and frees its shared memory) while a peer may still write to it
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"((uint32_t)TMEM_COLS));
    }
}

// h0 (or zeros) -> fp32 state + activation images of the state; rows M .. m_tiles*128 are zero
__global__ void init_state_kernel(const float* __restrict__ h0, float* __restrict__ h, unsigned char* __restrict__ h_img,
                                  int M, int H, int rows) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // one thread per 8 consecutive columns
    const int per_row = H / 8;
    if (idx >= rows * per_row) return;
    const int m = idx / per_row, k0 = (idx - m * per_row) * 8;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (m < M && h0) ? h0[(size_t)m * H + k0 + i] : 0.f;
    if (m < M) {
#pragma unroll
        for (int i = 0; i < 8; ++i) h[(size_t)m * H + k0 + i] = v[i];
    }
    uint4 hi, lo;
    split_pair(v[0], v[1], hi.x, lo.x); split_pair(v[2], v[3], hi.y, lo.y);
    split_pair(v[4], v[5], hi.z, lo.z); split_pair(v[6], v[7], hi.w, lo.w);
    const int row = m % TILE_M;
    unsigned char* base = h_img + ((size_t)(m / TILE_M) * (H >> 6) + (k0 >> 6)) * ACT_CHUNK_BYTES + row * 128 +
                          ((((k0 & 63) >> 3) ^ (row & 7)) << 4);
    *reinterpret_cast<uint4*>(base) = hi;
    *reinterpret_cast<uint4*>(base + ACT_PART_BYTES) = lo;
}

int set_attrs() {
    static bool done = false;
    if (!done) {
        BVC_CUDA(cudaFuncSetAttribute(recurrent_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DYNAMIC));
        done = true;
    }
    return BVC_OK;
}

}  // namespace

size_t smem_bytes() { return SMEM_DYNAMIC; }

int init_state(const float* h0, float* h, unsigned char* h_img, int M, int H, cudaStream_t stream) {
    const int rows = ((M + TILE_M - 1) / TILE_M) * TILE_M;
    const int total = rows * (H / 8);
    init_state_kernel<<<(total + 255) / 256, 256, 0, stream>>>(h0, h, h_img, M, H, rows);
    BVC_CHECK_LAUNCH();
    return BVC_OK;
}

// clusters that can be co-resident (the phase barrier needs all of them running)
int max_clusters(int device, int* out) {
    static int cached[64] = {0};
    if (device >= 0 && device < 64 && cached[device]) { *out = cached[device]; return BVC_OK; }
    int rc = set_attrs();
    if (rc) return rc;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(MAX_CLUSTERS * CLUSTER);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = SMEM_DYNAMIC;
    int n = 0;
    BVC_CUDA(cudaOccupancyMaxActiveClusters(&n, recurrent_cluster_kernel, &cfg));
    if (n < 1) { set_error("recurrent kernel: no 4-CTA cluster fits on this device"); return BVC_ERR_DEVICE; }
    if (n > MAX_CLUSTERS) n = MAX_CLUSTERS;
    *out = n;
    if (device >= 0 && device < 64) cached[device] = n;
    return BVC_OK;
}

int launch(const Program* prog_dev, int n_clusters, unsigned* sync_words, cudaStream_t stream) {
    int rc = set_attrs();
    if (rc) return rc;
    BVC_CUDA(cudaMemsetAsync(sync_words, 0, SYNC_WORDS * sizeof(unsigned), stream));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_clusters * CLUSTER);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = SMEM_DYNAMIC;
    cfg.stream = stream;
    // Cooperative launch: all clusters co-resident or the launch fails.  Where the driver refuses the
    // cooperative + cluster combination the kernel is launched plainly: the grid (<= 32 clusters, sized by
    // cudaOccupancyMaxActiveClusters) fits the device at once, and every wait in the kernel is bounded.
    static int cooperative_ok = 1;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = cooperative_ok ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, recurrent_cluster_kernel, prog_dev, sync_words);
    if (e != cudaSuccess && cooperative_ok) {
        cudaGetLastError();
        cooperative_ok = 0;
        cfg.numAttrs = 0;
        e = cudaLaunchKernelEx(&cfg, recurrent_cluster_kernel, prog_dev, sync_words);
    }
    if (e != cudaSuccess) {
        set_error(std::string("recurrent kernel launch failed: ") + cudaGetErrorString(e));
        return BVC_ERR_DEVICE;
    }
    if (g_launch_counter) ++*g_launch_counter;
    return BVC_OK;
}

}  // namespace rec
}  // namespace bvc
