// Persistent recurrent kernel for the BVRNN time loop (5th-gen tensor cores, accumulators in TMEM).
//
// Replaces the per-frame Python loop of reference bvrnn.py:186-206 (encode) and :222-227 (decode),
// i.e. ~15 nn.Linear launches + ~250 small ATen kernels per frame, with ONE cooperative kernel that
// stays resident on every SM for all T frames:
//   * a frame is a fixed program of "phases"; a phase is a set of GEMM tiles that only depend on
//     earlier phases; a device-wide barrier separates phases
//   * the host builds the program (bvrnn.cu): critical-path layers form the phases, layers off the
//     critical path (W_hh h, dec.0_h h, W_ih_z phi_z) are scheduled as background tiles into the
//     slack of other phases so that every SM has work in every phase
//   * activations travel between phases as split bf16 (hi, lo) rows through L2; weights are split
//     bf16 too
// Tile engine:
//   * tile = 64 rows x BN columns (BN = 32, or 48 for the GRU layer), K streamed in 64-wide stages
//   * operands are copied global -> shared with cp.async (16-byte chunks, 8 lanes per 128-byte row:
//     fully coalesced) straight into the UMMA canonical K-major SWIZZLE_128B layout: row r of a stage
//     occupies 128 contiguous bytes and its 16-byte chunk c is stored at chunk position c ^ (r & 7)
//     (the layout TMA would produce); SBO = 1024 B (next 8 rows), a k16 step advances the start by 32 B
//   * one thread issues tcgen05.mma.cta_group::1.kind::f16 (M=64, N=BN, K=16); the split-bf16 product
//     a.w ~= a_lo.w_hi + a_hi.w_lo + a_hi.w_hi is three MMAs into the same fp32 TMEM accumulator
//   * tcgen05.commit -> mbarrier tells the copy threads when a shared-memory stage may be refilled and
//     the epilogue warps when the accumulator is complete
//   * epilogue: warps 0-3 read their TMEM lane quadrant (tcgen05.ld.32x32b); with M=64 the rows of a
//     quadrant sit in its lower 16 lanes, one thread owns one output row and all BN columns of it, which
//     makes the fused bottleneck (32 code bits of a row -> one word) and GRU gate math (r, z, n of 16
//     hidden units in one thread) straight-line code
#include <cuda_bf16.h>

#include "common.cuh"
#include "recurrent.cuh"

namespace bvc {
namespace rec {

namespace {

constexpr int BM = 64, BK = 128, STAGES = 3, MAX_BN = 48;   // a stage = two 64-wide (128-byte) K sub-blocks
constexpr int A_SUB = BM * 128, W_SUB = MAX_BN * 128;         // bytes of one sub-block of one of hi / lo
constexpr int A_BYTES = 2 * A_SUB;                            // one of hi / lo
constexpr int W_BYTES = 2 * W_SUB;
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * W_BYTES;
constexpr int EPI_PITCH = MAX_BN + 1;                         // fp32 staging tile [BM][EPI_PITCH] for the epilogue
constexpr int kThreads = 256;
constexpr int TMEM_COLS = 64;

struct Ctx {
    unsigned char* smem;      // STAGES stages
    float* epi;               // [BM][EPI_PITCH] accumulator staging for the epilogue
    uint64_t* full;           // STAGES mbarriers: stage landed (kProducerThreads arrivals)
    uint64_t* empty;          // STAGES mbarriers: stage consumed (tcgen05.commit)
    uint64_t* acc;            // accumulator complete (tcgen05.commit)
    uint32_t tmem;            // TMEM base address
    uint32_t it;              // running k-iteration counter (uniform), selects stage and mbarrier parity
    uint32_t n_tiles;         // tiles processed so far (parity of the acc barrier)
    int* abort_flag;
    bool aborted;
    int dbg;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_addr), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug becomes an abort flag instead of a hung GPU
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int* abort_flag) {
    if (mbar_try(bar, parity)) return true;
    const long long t0 = clock64();
    while (!mbar_try(bar, parity)) {
        if (clock64() - t0 > 2000000000LL) {
            atomicExch(abort_flag, 2);
            return false;
        }
    }
    return true;
}

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

// store 8 consecutive values of a row as split bf16 (16 bytes hi, 16 bytes lo)
__device__ __forceinline__ void store_split8(__nv_bfloat16* hi_ptr, __nv_bfloat16* lo_ptr, const float* v) {
    uint4 h, l;
    split_pair(v[0], v[1], h.x, l.x);
    split_pair(v[2], v[3], h.y, l.y);
    split_pair(v[4], v[5], h.z, l.z);
    split_pair(v[6], v[7], h.w, l.w);
    *reinterpret_cast<uint4*>(hi_ptr) = h;
    if (lo_ptr) *reinterpret_cast<uint4*>(lo_ptr) = l;
}

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ bool grid_barrier(unsigned* counter, unsigned target, int* abort_flag) {
    __syncthreads();
    __shared__ int s_abort;
    if (threadIdx.x == 0) {
        // release: orders every write of this CTA that happened before the bar.sync above (cumulativity)
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(counter) : "memory");
        int ab = 0;
        const long long t_start = clock64();
        while (ld_acquire(counter) < target) {
            if (clock64() - t_start > 4000000000LL) {
                atomicExch(abort_flag, 1);
                ab = 1;
                break;
            }
            if (*(volatile int*)abort_flag) { ab = 1; break; }
        }
        s_abort = ab;   // ld.acquire above + the bar.sync below order every later read of this CTA;
                        // cross-CTA data is only read through L2 (cp.async.cg / ld.cg), so no L1 invalidate is needed
    }
    __syncthreads();
    return s_abort != 0;
}

// ---------------------------------------------------------------------------------------------
// Warp roles inside a tile:
//   warp 0      : MMA issuer (one elected lane): waits full[slot], issues the UMMAs of the stage,
//                 tcgen05.commit -> empty[slot]; after the last stage commit -> acc barrier
//   warps 1..7  : producers: wait empty[slot], cp.async the stage, and D stages later (when their copies
//                 have landed) fence.proxy.async + arrive on full[slot]
//   warps 4..7  : additionally the epilogue (TMEM lane quadrant = warp % 4)
// ---------------------------------------------------------------------------------------------
constexpr int kProducerThreads = kThreads - 32;
constexpr int LOOKAHEAD = STAGES - 1;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}

template <int BN>
__device__ __forceinline__ void umma_tile(const Op& op, const Frame& fr, int tile, int t, Ctx& cx) {
    const int M = fr.M;
    const int m_tiles = (M + BM - 1) / BM;
    const int m_tile = tile % m_tiles, n_tile = tile / m_tiles;
    const int m0 = m_tile * BM, n0 = n_tile * BN;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = op.K, kiters = (K + BK - 1) / BK;
    const __nv_bfloat16* a_hi = op.a_hi + (size_t)t * op.a_tstride;
    const __nv_bfloat16* a_lo = op.a_lo ? op.a_lo + (size_t)t * op.a_tstride : nullptr;
    const uint32_t smem_base = smem_u32(cx.smem);
    const uint32_t it0 = cx.it;

    if (warp == 0) {
        // ------------------------------- MMA issuer -------------------------------
        const uint32_t idesc = make_idesc(BM, BN);
        const uint64_t d0 = make_desc(smem_base);                    // descriptor of (slot 0, Ah, sub 0, ks 0)
        const bool leader = elect_one();
#pragma unroll 1
        for (int kt = 0; kt < kiters; ++kt) {
            const uint32_t it = it0 + kt;
            const int slot = it % STAGES;
            const int subs = (K - kt * BK) >= BK ? 2 : 1;
            if (!mbar_wait(&cx.full[slot], (it / STAGES) & 1, cx.abort_flag)) cx.aborted = true;
            tc_fence_after();
            if (leader) {
                if (!(cx.dbg & 2)) {
                    const uint64_t dah0 = d0 + (uint64_t)((slot * STAGE_BYTES) >> 4);
                    for (int sub = 0; sub < subs; ++sub) {
                        const uint64_t dah = dah0 + (uint64_t)((sub * A_SUB) >> 4);
                        const uint64_t dal = dah + (A_BYTES >> 4);
                        const uint64_t dwh = dah0 + (uint64_t)((2 * A_BYTES + sub * W_SUB) >> 4);
                        const uint64_t dwl = dwh + (W_BYTES >> 4);
                        if (a_lo) {
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                umma(cx.tmem, dal + 2 * ks, dwh + 2 * ks, idesc, (kt | sub | ks) != 0 ? 1u : 0u);
                                umma(cx.tmem, dah + 2 * ks, dwl + 2 * ks, idesc, 1u);
                                umma(cx.tmem, dah + 2 * ks, dwh + 2 * ks, idesc, 1u);
                            }
                        } else {
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                umma(cx.tmem, dah + 2 * ks, dwl + 2 * ks, idesc, (kt | sub | ks) != 0 ? 1u : 0u);
                                umma(cx.tmem, dah + 2 * ks, dwh + 2 * ks, idesc, 1u);
                            }
                        }
                    }
                }
                umma_commit(&cx.empty[slot]);                        // stage may be refilled when these retire
                if (kt == kiters - 1) umma_commit(cx.acc);           // accumulator complete
            }
            __syncwarp();
        }
    } else {
        // ------------------------------- producers -------------------------------
        const int ptid = tid - 32;
        auto load_stage = [&](int slot, int k0, int subs) {
            const uint32_t st = smem_base + slot * STAGE_BYTES;
            for (int c = ptid; c < subs * BM * 8; c += kProducerThreads) {   // A: subs x 64 rows x 8 chunks
                const int ch = c & 7, r = (c >> 3) % BM, sub = c / (BM * 8);
                const int row = min(m0 + r, M - 1);
                const size_t go = (size_t)row * op.lda + k0 + sub * 64 + ch * 8;
                const uint32_t so = sub * A_SUB + r * 128 + ((ch ^ (r & 7)) << 4);
                cp_async16(st + so, a_hi + go);
                if (a_lo) cp_async16(st + A_BYTES + so, a_lo + go);
            }
            for (int c = ptid; c < subs * BN * 8; c += kProducerThreads) {   // W: subs x BN rows x 8 chunks
                const int ch = c & 7, r = (c >> 3) % BN, sub = c / (BN * 8);
                const int row = min(n0 + r, op.N - 1);
                const size_t go = (size_t)row * K + k0 + sub * 64 + ch * 8;
                const uint32_t so = sub * W_SUB + r * 128 + ((ch ^ (r & 7)) << 4);
                cp_async16(st + 2 * A_BYTES + so, op.w_hi + go);
                cp_async16(st + 2 * A_BYTES + W_BYTES + so, op.w_lo + go);
            }
        };
#pragma unroll 1
        for (int j = 0; j < kiters + LOOKAHEAD; ++j) {
            if (j < kiters) {
                const uint32_t it = it0 + j;
                const int slot = it % STAGES;
                if (it >= (uint32_t)STAGES)      // previous use of this slot must have been consumed
                    if (!mbar_wait(&cx.empty[slot], ((it / STAGES) - 1) & 1, cx.abort_flag)) cx.aborted = true;
                if (!(cx.dbg & 1)) load_stage(slot, j * BK, (K - j * BK) >= BK ? 2 : 1);
            }
            cp_async_commit();
            if (j >= LOOKAHEAD) {
                cp_async_wait<LOOKAHEAD>();       // the copies of stage j - LOOKAHEAD have landed
                fence_proxy_async();              // ... and are visible to the tensor-core (async) proxy
                mbar_arrive(&cx.full[(it0 + j - LOOKAHEAD) % STAGES]);
            }
        }
    }
    cx.it = it0 + kiters;

    // ---- accumulator: TMEM -> registers (warps 4..7, one row per lane) -> shared staging tile ----
    float* epi = cx.epi;
    if (warp >= 4) {
        if (!mbar_wait(cx.acc, cx.n_tiles & 1, cx.abort_flag)) cx.aborted = true;
        tc_fence_after();
        float v[BN];
        const int quad = warp & 3;                        // TMEM lane quadrant this warp may access
        const uint32_t taddr = cx.tmem + ((uint32_t)(quad * 32) << 16);
#pragma unroll
        for (int c = 0; c < BN; c += 16) tmem_ld16(taddr + c, v + c);
        if (lane < 16) {                                  // M = 64: row 16 q + i lives in lane 32 q + i, i < 16
            float* dst = epi + (quad * 16 + lane) * EPI_PITCH;
#pragma unroll
            for (int c = 0; c < BN; ++c) dst[c] = v[c];
        }
    }
    cx.n_tiles++;
    tc_fence_before();   // TMEM reads done before the next tile's first MMA overwrites the accumulator
    __syncthreads();

    // ------------------------------- epilogue: all 256 threads -------------------------------
    if (!(cx.dbg & 4)) {
        if (op.kind == KIND_LINEAR || op.kind == KIND_MEL) {
            // item = (row, group of 8 columns)
            constexpr int G8 = BN / 8;
            for (int item = tid; item < BM * G8; item += kThreads) {
                const int r = item / G8, c8 = (item - r * G8) * 8;
                const int m = m0 + r, nb = n0 + c8;
                if (m >= M || nb >= op.N) continue;
                const float* src = epi + r * EPI_PITCH + c8;
                float o[8];
                float bb[8];
                if (op.bias) {
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(op.bias + nb));
                    const float4 b1 = __ldg(reinterpret_cast<const float4*>(op.bias + nb + 4));
                    bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w; bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) bb[e] = 0.f;
                }
                if (op.addend) {
                    const float* ap = op.addend + (size_t)t * op.add_tstride + (size_t)m * op.ldadd + nb;
                    const float4 a0 = __ldcg(reinterpret_cast<const float4*>(ap));
                    const float4 a1 = __ldcg(reinterpret_cast<const float4*>(ap + 4));
                    bb[0] += a0.x; bb[1] += a0.y; bb[2] += a0.z; bb[3] += a0.w; bb[4] += a1.x; bb[5] += a1.y; bb[6] += a1.z; bb[7] += a1.w;
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    float x = src[e] + bb[e];
                    if (op.act) x = elu1(x);
                    o[e] = x;
                }
                if (op.kind == KIND_MEL) {
                    if (fr.mel_out) {
                        float* mo = fr.mel_out + ((size_t)m * fr.T + t) * fr.X + nb;
                        *reinterpret_cast<float4*>(mo) = make_float4(o[0], o[1], o[2], o[3]);
                        *reinterpret_cast<float4*>(mo + 4) = make_float4(o[4], o[5], o[6], o[7]);
                    }
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] = (o[e] - __ldg(fr.mean + nb + e)) / __ldg(fr.std + nb + e);
                }
                if (op.out_f) {
                    float* of = op.out_f + (size_t)m * op.ldo + nb;
                    *reinterpret_cast<float4*>(of) = make_float4(o[0], o[1], o[2], o[3]);
                    *reinterpret_cast<float4*>(of + 4) = make_float4(o[4], o[5], o[6], o[7]);
                }
                if (op.out_hi)
                    store_split8(op.out_hi + (size_t)m * op.ldos + nb, op.out_lo + (size_t)m * op.ldos + nb, o);
            }
        } else if (op.kind == KIND_BOTTLENECK) {
            // z = round(sigmoid(logit)), masked to 0.5 beyond the frame's bit budget (bvrnn.py:191-196).
            // BN = 32: 4 consecutive lanes hold the 4 x 8 code bits of one row.
            if constexpr (BN == 32) {
                const int r = tid >> 2, c8 = (tid & 3) * 8;
                const int m = m0 + r, nb = n0 + c8;
                const bool valid = m < M;
                uint32_t word = 0;
                if (valid) {
                    const float budget = fr.bits ? __ldg(fr.bits + (size_t)m * fr.T + t) : fr.bits_scalar;
                    const float* src = epi + r * EPI_PITCH + c8;
                    float code[8], lg[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        lg[e] = src[e] + __ldg(op.bias + nb + e);
                        const bool active = !fr.var_bit || (budget > (float)(nb + e));
                        const bool bit = active && (sigmoidf_(lg[e]) > 0.5f);
                        code[e] = active ? (bit ? 1.f : 0.f) : 0.5f;
                        if (bit) word |= 1u << (c8 + e);
                    }
                    float* co = fr.codes + ((size_t)m * fr.T + t) * fr.Z + nb;
                    *reinterpret_cast<float4*>(co) = make_float4(code[0], code[1], code[2], code[3]);
                    *reinterpret_cast<float4*>(co + 4) = make_float4(code[4], code[5], code[6], code[7]);
                    if (fr.logits) {
                        float* lo_out = fr.logits + ((size_t)m * fr.T + t) * fr.Z + nb;
                        *reinterpret_cast<float4*>(lo_out) = make_float4(lg[0], lg[1], lg[2], lg[3]);
                        *reinterpret_cast<float4*>(lo_out + 4) = make_float4(lg[4], lg[5], lg[6], lg[7]);
                    }
                    store_split8(op.out_hi + (size_t)m * op.ldos + nb, nullptr, code);   // {0,.5,1} exact in bf16
                }
                word |= __shfl_xor_sync(0xffffffffu, word, 1);
                word |= __shfl_xor_sync(0xffffffffu, word, 2);
                if (valid && fr.packed && (tid & 3) == 0)
                    reinterpret_cast<uint32_t*>(fr.packed)[((size_t)m * fr.T + t) * 2 + n_tile] = word;
            }
        } else if (op.kind == KIND_GRU) {
            // columns of this tile: [r(16) | z(16) | n(16)] of hidden units j0 .. j0+15; a thread owns 4 units of a row
            if constexpr (BN == 48) {
                const int H = fr.H, j0 = n_tile * 16;
                const int r = tid >> 2, u = (tid & 3) * 4;
                const int m = m0 + r;
                if (m < M) {
                    const float* src = epi + r * EPI_PITCH;
                    const float* gz = op.addend + (size_t)t * op.add_tstride + (size_t)m * op.ldadd + n0;
                    const float* gh = fr.gh + (size_t)m * 3 * H + n0;
                    float* hp = fr.h + (size_t)m * H + j0 + u;
                    const float4 zr = __ldcg(reinterpret_cast<const float4*>(gz + u));
                    const float4 zz = __ldcg(reinterpret_cast<const float4*>(gz + 16 + u));
                    const float4 zn = __ldcg(reinterpret_cast<const float4*>(gz + 32 + u));
                    const float4 hr = __ldcg(reinterpret_cast<const float4*>(gh + u));
                    const float4 hz = __ldcg(reinterpret_cast<const float4*>(gh + 16 + u));
                    const float4 hn4 = __ldcg(reinterpret_cast<const float4*>(gh + 32 + u));
                    const float4 hv4 = __ldcg(reinterpret_cast<const float4*>(hp));
                    const float gzr[4] = {zr.x, zr.y, zr.z, zr.w}, gzz[4] = {zz.x, zz.y, zz.z, zz.w};
                    const float gzn[4] = {zn.x, zn.y, zn.z, zn.w}, ghr[4] = {hr.x, hr.y, hr.z, hr.w};
                    const float ghz[4] = {hz.x, hz.y, hz.z, hz.w}, ghn[4] = {hn4.x, hn4.y, hn4.z, hn4.w};
                    const float hv[4] = {hv4.x, hv4.y, hv4.z, hv4.w};
                    float hn[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float rr = sigmoidf_(src[u + e] + gzr[e] + ghr[e]);
                        const float zg = sigmoidf_(src[16 + u + e] + gzz[e] + ghz[e]);
                        const float nn = tanhf(src[32 + u + e] + gzn[e] + rr * ghn[e]);
                        hn[e] = (hv[e] - nn) * zg + nn;
                    }
                    if (fr.all_h)      // state entering frame t (bvrnn.py:205)
                        *reinterpret_cast<float4*>(fr.all_h + ((size_t)m * fr.T + t) * H + j0 + u) = hv4;
                    *reinterpret_cast<float4*>(hp) = make_float4(hn[0], hn[1], hn[2], hn[3]);
                    uint32_t h0, l0, h1, l1;
                    split_pair(hn[0], hn[1], h0, l0);
                    split_pair(hn[2], hn[3], h1, l1);
                    *reinterpret_cast<uint2*>(op.out_hi + (size_t)m * op.ldos + j0 + u) = make_uint2(h0, h1);
                    *reinterpret_cast<uint2*>(op.out_lo + (size_t)m * op.ldos + j0 + u) = make_uint2(l0, l1);
                }
            }
        }
    }
    __syncthreads();   // staging tile and pipeline stages are reused by the next tile
}

__global__ void __launch_bounds__(kThreads, 1)
recurrent_umma_kernel(const Program* __restrict__ prog, unsigned* barrier_counter, int* abort_flag) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar[2 * STAGES + 1];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&mbar[i], kProducerThreads);     // full
            mbar_init(&mbar[STAGES + i], 1);           // empty
        }
        mbar_init(&mbar[2 * STAGES], 1);               // accumulator
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_slot)),
                     "r"((uint32_t)TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    Ctx cx;
    cx.smem = smem_raw;
    cx.epi = reinterpret_cast<float*>(smem_raw + STAGES * STAGE_BYTES);
    cx.full = mbar;
    cx.empty = mbar + STAGES;
    cx.acc = mbar + 2 * STAGES;
    cx.tmem = tmem_slot;
    cx.it = 0;
    cx.n_tiles = 0;
    cx.abort_flag = abort_flag;
    cx.aborted = false;
    cx.dbg = prog->debug_flags;

    const Frame& fr = prog->frame;
    const int G = gridDim.x, cta = blockIdx.x;
    const int n_phases = prog->n_phases;
    unsigned bar = 0;
    bool stop = false;
    for (int t = 0; t < fr.T && !stop; ++t) {
        for (int ph = 0; ph < n_phases; ++ph) {
            const int lb = prog->list_start[ph * G + cta], le = prog->list_start[ph * G + cta + 1];
            for (int i = lb; i < le; ++i) {
                const uint32_t e = prog->tiles[i];
                const Op& op = prog->ops[e >> 20];
                const int tile = (int)(e & 0xFFFFF);
                if (op.bn == 48) umma_tile<48>(op, fr, tile, t, cx);
                else umma_tile<32>(op, fr, tile, t, cx);
            }
            bar += (unsigned)G;
            if (!(cx.dbg & 8) && grid_barrier(barrier_counter, bar, abort_flag)) { stop = true; break; }
        }
    }
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(cx.tmem), "r"((uint32_t)TMEM_COLS));
    }
}

// h0 (or zeros) -> fp32 state + split-bf16 state; also clears the zero-padded columns of the mel buffer
__global__ void init_state_kernel(const float* __restrict__ h0, float* __restrict__ h, __nv_bfloat16* __restrict__ h_hi,
                                  __nv_bfloat16* __restrict__ h_lo, int n, __nv_bfloat16* __restrict__ mn_hi,
                                  __nv_bfloat16* __restrict__ mn_lo, int n_mn) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float v = h0 ? h0[i] : 0.f;
        h[i] = v;
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        h_hi[i] = hi;
        h_lo[i] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
    if (i < n_mn) {
        mn_hi[i] = __float2bfloat16_rn(0.f);
        mn_lo[i] = __float2bfloat16_rn(0.f);
    }
}

}  // namespace

size_t umma_smem_bytes() { return (size_t)STAGES * STAGE_BYTES + BM * EPI_PITCH * sizeof(float) + 1024; }

int init_state(const float* h0, float* h, __nv_bfloat16* h_hi, __nv_bfloat16* h_lo, int n, __nv_bfloat16* mn_hi,
               __nv_bfloat16* mn_lo, int n_mn, cudaStream_t stream) {
    const int total = n > n_mn ? n : n_mn;
    init_state_kernel<<<(total + 255) / 256, 256, 0, stream>>>(h0, h, h_hi, h_lo, n, mn_hi, mn_lo, n_mn);
    BVC_CHECK_LAUNCH();
    return BVC_OK;
}

// one CTA per SM; fails if the kernel cannot be co-resident (cooperative launch requirement)
int max_grid(int device, int* out) {
    static int cached[64] = {0};
    if (device >= 0 && device < 64 && cached[device]) { *out = cached[device]; return BVC_OK; }
    BVC_CUDA(cudaFuncSetAttribute(recurrent_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)umma_smem_bytes()));
    int per_sm = 0, sms = 0;
    BVC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, recurrent_umma_kernel, kThreads, umma_smem_bytes()));
    BVC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    if (per_sm < 1) { set_error("recurrent kernel does not fit on an SM"); return BVC_ERR_DEVICE; }
    *out = sms;
    if (device >= 0 && device < 64) cached[device] = sms;
    return BVC_OK;
}

int umma_launch(const Program* prog_dev, int grid, unsigned* sync_words, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        BVC_CUDA(cudaFuncSetAttribute(recurrent_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)umma_smem_bytes()));
        attr_set = true;
    }
    BVC_CUDA(cudaMemsetAsync(sync_words, 0, 2 * sizeof(unsigned), stream));
    unsigned* barrier_counter = sync_words;
    int* abort_flag = reinterpret_cast<int*>(sync_words + 1);
    void* args[] = {(void*)&prog_dev, (void*)&barrier_counter, (void*)&abort_flag};
    BVC_CUDA(cudaLaunchCooperativeKernel((const void*)recurrent_umma_kernel, dim3(grid), dim3(kThreads), args,
                                         umma_smem_bytes(), stream));
    if (g_launch_counter) ++*g_launch_counter;
    return BVC_OK;
}

}  // namespace rec
}  // namespace bvc
