// Causal BigVGAN-tiny vocoder.
//
// Replaces reference third_party/BigVGAN/models.py:207-238 (BigVGAN.forward), :103-121
// (AMPBlock1.forward), activations.py:107-120 (SnakeBeta) with weight-norm folded at load.
//
// Data flow (per utterance, n_0 = T mel frames, n_{i+1} = u_i (n_i + 1)); every intermediate in HBM
// is channel-last [B, n, C]:
//   conv_pre      : one GEMM over 7 contiguous channel-last mel frames -> [T, 128]
//   stage i (x4)  : one kernel per resblock kernel size k in {3,7,11}.  A CTA owns a time tile of the
//                   stage's output rate plus a left halo of 12 (k-1) samples, and keeps the whole
//                   chain  ConvTranspose1d -> 3 x (snake, dilated conv, snake, conv, residual)
//                   in shared memory: the stage input is read once (at the lower rate) and the
//                   resblock output written once.  The three resblocks of a stage write separate
//                   partial tensors; their mean (models.py:219-225) is taken by the consumer on load.
//   post          : mean -> snake -> conv_post(k=7) -> tanh -> / SCALING -> [:length]
// Causal left padding is implicit: every conv input at global time < 0 is zero (models.py:110,117).
//
// Two arithmetic paths for the stage kernel:
//   precision 0  fp32 FFMA, activations [C][time] in shared memory
//   precision 1  tensor cores (mma.sync m16n8k16 bf16, fp32 accumulate), implicit GEMM with
//                M = time, N = C_out, K = taps x C_in.  Conv inputs live in shared memory as split
//                bf16 (hi, lo) rows [time][C]; a tap is a row offset of the A fragment, so no im2col
//                copy exists.  x.w ~= x_hi.w_hi + x_hi.w_lo + x_lo.w_hi.  The residual stream stays fp32.
#include <algorithm>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"

#define BVC_TRY(x) do { int rc_try_ = (x); if (rc_try_ != BVC_OK) return rc_try_; } while (0)

namespace bvc {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float snake(float x, float ea, float inv_eb) {
    const float s = sinf(x * ea);
    return x + inv_eb * (s * s);
}
// tensor-core path: two-constant Cody-Waite reduction to [-pi, pi] followed by MUFU.SIN (abs error ~4e-7 for the
// argument range of the vocoder, |x e^alpha| << 1e4); the fp32 path keeps sinf
__device__ __forceinline__ float snake_fast(float x, float ea, float inv_eb) {
    const float y = x * ea;
    const float k = rintf(y * 0.15915494309189535f);
    float r = fmaf(k, -6.2831854820251465f, y);
    r = fmaf(k, 1.7484555e-7f, r);
    const float s = __sinf(r);
    return x + inv_eb * (s * s);
}

__global__ void pad_mel_kernel(const float* __restrict__ mel, float* __restrict__ out, int B, int T, int X) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)B * (T + 6) * X;
    if (idx >= total) return;
    const int c = (int)(idx % X);
    const size_t row = idx / X;
    const int t = (int)(row % (T + 6)) - 6;
    const size_t b = row / (T + 6);
    out[idx] = t >= 0 ? mel[(b * T + t) * X + c] : 0.f;
}

struct StageArgs {
    // stage input: n_parts channel-last tensors [B, rows, CIN] whose mean is the input (1 for conv_pre's output)
    const float* in_p[3];
    int n_parts;
    long long in_bstride;     // floats per utterance
    int n_in;
    int n_out;
    const float* w_up;        // fp32 [tap][ci][co]
    const float* b_up;
    const uint2* upf_h;       // fragment-packed per phase
    const uint2* upf_l;
    const float* w1[3];
    const float* b1[3];
    const float* w2[3];
    const float* b2[3];
    const uint2 *f1h[3], *f1l[3], *f2h[3], *f2l[3];
    const float* ea[6];
    const float* ieb[6];
    int dil[3];
    float* out;               // [B, n_out, C] channel-last
    int TT;
    // fused tail of the vocoder (last stage, last resblock kernel only): mean with the two partials written by the other
    // resblock kernels -> SnakeBeta -> conv_post -> tanh -> / inv_scale_div -> wav[:n_wav]  (models.py:228-238)
    const float* post_p1;     // partials [B, n_out, C] of the other two resblocks, summed in the reference's order
    const float* post_p2;
    const float* post_ea;
    const float* post_ieb;
    const float* post_w;      // [ci][7]
    const float* post_bias;
    float post_inv_scale_div;
    float* wav;               // [B, n_wav]
    int n_wav;
};

__device__ __forceinline__ float load_mean(const StageArgs& a, size_t o) {
    if (a.n_parts == 1) return __ldg(a.in_p[0] + o);
    return ((__ldg(a.in_p[0] + o) + __ldg(a.in_p[1] + o)) + __ldg(a.in_p[2] + o)) / 3.0f;
}

// ===========================================================================
// precision 0: fp32 FFMA stage kernel
// ===========================================================================
// in [C][WP] (conv input, already activated and zero where global time < 0)
// w  [ci][tap][co], out positions p in [p_begin, W).  f(co, p, value) consumes the result.
template <int C, int K, typename F>
__device__ __forceinline__ void conv_layer(const float* __restrict__ in, int WP, const float* __restrict__ w,
                                           const float* __restrict__ bias, int d, int p_begin, int W, F f) {
    constexpr int CG = C < 8 ? C : 8;
    constexpr int NCG = C / CG;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_chunks = (W - p_begin + 127) / 128;
    for (int item = warp; item < n_chunks * NCG; item += kThreads / 32) {
        const int cg = item / n_chunks, chunk = item - cg * n_chunks;
        const int p0 = p_begin + chunk * 128 + lane;
        float acc[CG][4];
#pragma unroll
        for (int a = 0; a < CG; ++a) {
            const float bv = __ldg(bias + cg * CG + a);
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = bv;
        }
        for (int ci = 0; ci < C; ++ci) {
            const float* row = in + ci * WP;
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const int off = (K - 1 - j) * d;
                float x[4];
#pragma unroll
                for (int b = 0; b < 4; ++b) x[b] = row[min(p0 + 32 * b - off, WP - 1)];
                const float* wp = w + ((size_t)(ci * K + j) * C + cg * CG);
                float wv[CG];
#pragma unroll
                for (int a = 0; a < CG; a += 4) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(wp + a));
                    wv[a] = v.x; wv[a + 1] = v.y; wv[a + 2] = v.z; wv[a + 3] = v.w;
                }
#pragma unroll
                for (int a = 0; a < CG; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(wv[a], x[b], acc[a][b]);
            }
        }
#pragma unroll
        for (int a = 0; a < CG; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int p = p0 + 32 * b;
                if (p < W) f(cg * CG + a, p, acc[a][b]);
            }
    }
}

template <int C, int U, int K>
__global__ void __launch_bounds__(kThreads, 1) stage_fp32_kernel(StageArgs a) {
    constexpr int CIN = 2 * C;
    constexpr int HALO = 12 * (K - 1);
    extern __shared__ __align__(16) float smem[];
    const int TT = a.TT;
    const int W = TT + HALO;
    const int WP = W + 4;
    float* cur = smem;                 // [C][WP] residual stream
    float* s1 = cur + C * WP;          // [C][WP] conv1 input
    float* s2 = s1 + C * WP;           // [C][WP] conv2 input
    float* xin = s1;                   // [CIN][NJP] stage input tile (aliases s1/s2 until the upsample is done)
    const int NJ = W / U + 2;
    const int NJP = NJ | 1;

    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * TT;
    const int tg0 = t0 - HALO;               // global time of position 0 (multiple of U)
    const int j_base = tg0 / U - 1;          // xin[.][0] <-> input sample j_base (exact: tg0 % U == 0)

    {
        const size_t boff = (size_t)b * a.in_bstride;
        for (int i = tid; i < NJ * CIN; i += kThreads) {
            const int jj = i / CIN, ci = i - jj * CIN;
            const int j = j_base + jj;
            xin[ci * NJP + jj] = (j >= 0 && j < a.n_in) ? load_mean(a, boff + (size_t)j * CIN + ci) : 0.f;
        }
    }
    __syncthreads();

    // ConvTranspose1d(k = 2U, stride U): y[co, U j + r] = b + sum_ci x[ci,j] W[ci,co,r] + x[ci,j-1] W[ci,co,r+U]
    {
        constexpr int CG = C < 8 ? C : 8;
        constexpr int NCG = C / CG;
        for (int item = tid; item < W * NCG; item += kThreads) {
            const int cg = item / W, p = item - cg * W;
            const int jj = p / U + 1, r = p - (p / U) * U;
            float acc[CG];
#pragma unroll
            for (int c = 0; c < CG; ++c) acc[c] = __ldg(a.b_up + cg * CG + c);
            const float* w0 = a.w_up + ((size_t)r * CIN) * C + cg * CG;
            const float* w1 = a.w_up + ((size_t)(r + U) * CIN) * C + cg * CG;
            for (int ci = 0; ci < CIN; ++ci) {
                const float xa = xin[ci * NJP + jj], xb = xin[ci * NJP + jj - 1];
#pragma unroll
                for (int c = 0; c < CG; c += 4) {
                    const float4 va = __ldg(reinterpret_cast<const float4*>(w0 + (size_t)ci * C + c));
                    const float4 vb = __ldg(reinterpret_cast<const float4*>(w1 + (size_t)ci * C + c));
                    acc[c] = fmaf(xa, va.x, acc[c]);         acc[c] = fmaf(xb, vb.x, acc[c]);
                    acc[c + 1] = fmaf(xa, va.y, acc[c + 1]); acc[c + 1] = fmaf(xb, vb.y, acc[c + 1]);
                    acc[c + 2] = fmaf(xa, va.z, acc[c + 2]); acc[c + 2] = fmaf(xb, vb.z, acc[c + 2]);
                    acc[c + 3] = fmaf(xa, va.w, acc[c + 3]); acc[c + 3] = fmaf(xb, vb.w, acc[c + 3]);
                }
            }
#pragma unroll
            for (int c = 0; c < CG; ++c) cur[(cg * CG + c) * WP + p] = acc[c];
        }
    }
    __syncthreads();   // xin (aliasing s1/s2) is dead from here on

    int lo = 0;
#pragma unroll 1
    for (int l = 0; l < 3; ++l) {
        const int d = a.dil[l];
        {
            const float* ea = a.ea[2 * l];
            const float* ieb = a.ieb[2 * l];
            for (int i = tid; i < C * (W - lo); i += kThreads) {
                const int c = i / (W - lo), p = lo + i - c * (W - lo);
                const float v = cur[c * WP + p];
                s1[c * WP + p] = (tg0 + p >= 0) ? snake(v, __ldg(ea + c), __ldg(ieb + c)) : 0.f;
            }
        }
        __syncthreads();
        const int lo1 = lo + (K - 1) * d;
        {
            const float* ea = a.ea[2 * l + 1];
            const float* ieb = a.ieb[2 * l + 1];
            conv_layer<C, K>(s1, WP, a.w1[l], a.b1[l], d, lo1, W, [&](int co, int p, float v) {
                s2[co * WP + p] = (tg0 + p >= 0) ? snake(v, __ldg(ea + co), __ldg(ieb + co)) : 0.f;
            });
        }
        __syncthreads();
        const int lo2 = lo1 + (K - 1);
        conv_layer<C, K>(s2, WP, a.w2[l], a.b2[l], 1, lo2, W, [&](int co, int p, float v) {
            cur[co * WP + p] += v;
        });
        __syncthreads();
        lo = lo2;
    }

    float* dst = a.out + (size_t)b * a.n_out * C;
    for (int i = tid; i < C * TT; i += kThreads) {
        const int tt = i / C, c = i - tt * C;
        const int tg = t0 + tt;
        if (tg < a.n_out) dst[(size_t)tg * C + c] = cur[c * WP + HALO + tt];
    }
}

// ===========================================================================
// precision 1: tensor-core stage kernel, streaming over time
// ===========================================================================
// A CTA owns (utterance b, resblock kernel size K, a contiguous time range) and walks that range in tiles of TT
// output samples.  All convolutions are causal, so the only coupling between consecutive tiles is the left
// context of each conv input: (K-1) d rows for the dilated conv of layer l, K-1 rows for its second conv.  Those
// rows are kept in shared memory from one tile to the next (6 small context buffers), so no sample is ever
// computed twice; a range that starts inside the utterance warms its contexts up on the 12 (K-1) samples before it.
// Time ranges per utterance for the streaming stage kernels.  Every CTA walks one range tile by tile, all CTAs of a launch
// take about the same time, and `slots` of them are resident at once, so the launch lasts ceil(CTAs / slots) waves of
// (tiles per range + warm-up tiles) tile times: pick the range count that minimises that product (a launch of 5.2 waves
// costs 6).  `units` = CTAs per range index (utterances x resblock CTAs).
static int pick_ranges(int units, int tiles_total, int warm_tiles, int slots) {
    int best = 1;
    long long best_cost = -1;
    const int max_ranges = tiles_total / (4 * (warm_tiles > 0 ? warm_tiles : 1)) > 0 ? tiles_total / (4 * (warm_tiles > 0 ? warm_tiles : 1)) : 1;
    for (int r = 1; r <= max_ranges && r <= 64; ++r) {
        const int tiles_per = (tiles_total + r - 1) / r;
        const int r_eff = (tiles_total + tiles_per - 1) / tiles_per;
        const long long waves = ((long long)units * r_eff + slots - 1) / slots;
        const long long cost = waves * (tiles_per + (r_eff > 1 ? warm_tiles : 0));
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = r_eff; }
    }
    return best;
}

template <int C>
struct RowLayout {
    static constexpr int PW = C >= 16 ? C / 2 + 4 : C / 2;   // uint32 (bf16 pair) words per activation row; rows are
                                                             // 16-byte aligned and 8 consecutive rows hit 8 distinct
                                                             // 16-byte bank groups (conflict-free ldmatrix)
    static constexpr int PF = C + 8;                         // floats per residual row
};

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], const uint2& b) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t smem_addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_addr));
}

__device__ __forceinline__ void split_pair(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
    const __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0));
    const __nv_bfloat16 l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
    hi = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    lo = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
}

// One warp computes a (16 MT)-row x COUT tile of  out[r][co] = sum_{tap,ci} in[r + lead - (NTAPS-1-tap) d][ci] W[tap][ci][co]
// for r = r0 .. r0 + 16 MT - 1.  inh/inl: shared-memory byte addresses of the split-bf16 activation rows (pitch
// RowLayout<CIN>::PW words); `lead` = index of the buffer row that holds output row 0's undelayed input.
// wh/wl: fragment-packed weights.  epi(r, nt, co, v0, v1) receives channels (co, co+1) = (8 nt + 2 q, +1) of output row r.
template <int CIN, int COUT, int NTAPS, int MT, typename Epi>
__device__ __forceinline__ void mma_rows(uint32_t inh, uint32_t inl, int lead, const uint2* __restrict__ wh,
                                         const uint2* __restrict__ wl, int d, int r0, Epi epi) {
    constexpr int PWB = RowLayout<CIN>::PW * 4;     // row pitch in bytes
    constexpr int NT = COUT / 8;
    constexpr int KC = (NTAPS * CIN + 15) / 16;
    const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    // ldmatrix: lane -> (matrix mi = lane / 8, row rr = lane % 8); matrices 0..3 = (rows 0-7 | 8-15) x (k 0-7 | 8-15)
    const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lkh = lane >> 4;
    float acc[MT][NT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[mt][nt][c] = 0.f;

    // weight fragments are fetched one k16 chunk ahead (they come from L1 / L2: long-scoreboard latency), and the
    // three split-bf16 terms are issued term-major so that consecutive MMAs never target the same accumulator
    uint2 bh[NT], bl[NT];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        bh[nt] = __ldg(wh + nt * 32 + lane);
        bl[nt] = __ldg(wl + nt * 32 + lane);
    }
#pragma unroll 2
    for (int kc = 0; kc < KC; ++kc) {
        const int kk = 16 * kc + 8 * lkh;           // first GEMM-K index of this lane's 8x8 matrix
        const int tap = kk / CIN, ci = kk % CIN;
        const int off = tap < NTAPS ? (NTAPS - 1 - tap) * d : 0;   // K padding (zero weights): any valid row
        const int rbase = r0 + lead + lrow - off;
        uint32_t ah[MT][4], al[MT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            const uint32_t o = (uint32_t)((rbase + mt * 16) * PWB + ci * 2);
            ldmatrix_x4(ah[mt], inh + o);
            ldmatrix_x4(al[mt], inl + o);
        }
        uint2 nh[NT], nl[NT];
        const int kn = kc + 1 < KC ? kc + 1 : kc;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            nh[nt] = __ldg(wh + (kn * NT + nt) * 32 + lane);
            nl[nt] = __ldg(wl + (kn * NT + nt) * 32 + lane);
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) mma16816(acc[mt][nt], al[mt], bh[nt]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) mma16816(acc[mt][nt], ah[mt], bl[nt]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) mma16816(acc[mt][nt], ah[mt], bh[nt]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) { bh[nt] = nh[nt]; bl[nt] = nl[nt]; }
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int row = r0 + mt * 16 + g, co = nt * 8 + 2 * q;
            epi(row, nt, co, acc[mt][nt][0], acc[mt][nt][1]);
            epi(row + 8, nt, co, acc[mt][nt][2], acc[mt][nt][3]);
        }
}

template <int C, int K>
struct StreamLayout {
    static constexpr int PW = RowLayout<C>::PW, PF = RowLayout<C>::PF;
    static constexpr int CTX1 = (K - 1) * 5;       // rows in front of the dilated conv's input tile (largest dilation)
    static constexpr int CTX2 = K - 1;
    static constexpr int CTXSUM = (K - 1) * 9;     // (K-1)(1+3+5) saved rows for the three dilated convs
    // shared memory in 32-bit words for a tile of TT rows (+16 spare rows so that partial warp tiles stay in bounds)
    static constexpr size_t words(int TT, bool post = false) {
        return (size_t)(TT + 16) * PF + 2 * (size_t)(CTX1 + TT + 16) * PW + 2 * (size_t)(CTX2 + TT + 16) * PW +
               2 * (size_t)CTXSUM * PW + 2 * (size_t)3 * CTX2 * PW + (post ? (size_t)(TT + 6) * C : 0);
    }
};

template <int C, int U, int K, int MT, bool POST = false>
__global__ void __launch_bounds__(kThreads) stage_stream_kernel(StageArgs a) {
    constexpr int CIN = 2 * C;
    constexpr int HALO = 12 * (K - 1);
    constexpr int HALOR = ((HALO + U - 1) / U) * U;
    using L = StreamLayout<C, K>;
    constexpr int PW = L::PW, PF = L::PF, PWI = RowLayout<CIN>::PW, CTX1 = L::CTX1, CTX2 = L::CTX2;
    constexpr int RW = 16 * MT;                    // rows per warp tile
    extern __shared__ __align__(16) float smem[];
    const int TT = a.TT;
    float* cur = smem;                                                   // [TT+16][PF] fp32 residual stream
    uint32_t* s1h = reinterpret_cast<uint32_t*>(cur + (TT + 16) * PF);   // [CTX1+TT+16][PW] dilated conv input (hi)
    uint32_t* s1l = s1h + (CTX1 + TT + 16) * PW;
    uint32_t* s2h = s1l + (CTX1 + TT + 16) * PW;                         // [CTX2+TT+16][PW] second conv input
    uint32_t* s2l = s2h + (CTX2 + TT + 16) * PW;
    uint32_t* c1h = s2l + (CTX2 + TT + 16) * PW;                         // saved contexts of the 3 dilated convs
    uint32_t* c1l = c1h + L::CTXSUM * PW;
    uint32_t* c2h = c1l + L::CTXSUM * PW;                                // [3][CTX2] saved contexts of the second convs
    uint32_t* c2l = c2h + 3 * CTX2 * PW;
    float* pact = reinterpret_cast<float*>(c2l + 3 * CTX2 * PW);        // POST: [6 + TT][C] activated mean, 6 rows of history
    const int NJ = TT / U + 1;                                           // low-rate rows j0-1 .. j0+TT/U-1
    uint32_t* xh = s1h;                                                  // [NJ+16][PWI] stage input tile (aliases s1)
    uint32_t* xl = xh + (NJ + 16) * PWI;

    const int tid = threadIdx.x, warp = tid >> 5;
    const int b = blockIdx.y;
    // this CTA's time range [t_begin, t_end) of the stage output, tile-aligned
    const int tiles_total = (a.n_out + TT - 1) / TT;
    const int tiles_per = (tiles_total + gridDim.x - 1) / gridDim.x;
    const int t_begin = blockIdx.x * tiles_per * TT;
    const int t_end = min(a.n_out, t_begin + tiles_per * TT);
    if (t_begin >= t_end) return;
    const int t_first = max(0, t_begin - ((HALOR + TT - 1) / TT) * TT);  // warm-up tiles (outputs discarded)

    for (int i = tid; i < 2 * (L::CTXSUM + 3 * CTX2) * PW; i += kThreads) c1h[i] = 0u;   // causal zero history
    if (POST)
        for (int i = tid; i < 6 * C; i += kThreads) pact[i] = 0.f;
    __syncthreads();

    const size_t boff = (size_t)b * a.in_bstride;
    float* dst = a.out + (size_t)b * a.n_out * C;

#pragma unroll 1
    for (int t0 = t_first; t0 < t_end; t0 += TT) {
        // ---- stage input rows j0-1 .. j0+TT/U-1: mean of the producer's partials, split to bf16 hi/lo ----
        {
            const int j_base = t0 / U - 1;
            constexpr int V = CIN / 4;
            for (int i = tid; i < NJ * V; i += kThreads) {
                const int jj = i / V, c4 = i - jj * V;
                const int j = j_base + jj;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j >= 0 && j < a.n_in) {
                    const size_t o = boff + (size_t)j * CIN + c4 * 4;
                    v = __ldg(reinterpret_cast<const float4*>(a.in_p[0] + o));
                    if (a.n_parts == 3) {
                        const float4 v1 = __ldg(reinterpret_cast<const float4*>(a.in_p[1] + o));
                        const float4 v2 = __ldg(reinterpret_cast<const float4*>(a.in_p[2] + o));
                        v.x = ((v.x + v1.x) + v2.x) / 3.0f; v.y = ((v.y + v1.y) + v2.y) / 3.0f;
                        v.z = ((v.z + v1.z) + v2.z) / 3.0f; v.w = ((v.w + v1.w) + v2.w) / 3.0f;
                    }
                }
                uint32_t h0, l0, h1, l1;
                split_pair(v.x, v.y, h0, l0);
                split_pair(v.z, v.w, h1, l1);
                *reinterpret_cast<uint2*>(xh + jj * PWI + c4 * 2) = make_uint2(h0, h1);
                *reinterpret_cast<uint2*>(xl + jj * PWI + c4 * 2) = make_uint2(l0, l1);
            }
        }
        __syncthreads();

        // ---- ConvTranspose1d as U phase convolutions: phase r is a 2-tap conv over (x[j-1], x[j]) ----
        {
            const int rows = TT / U;                    // low-rate rows m: output position p = U m + r
            const int tiles = (rows + RW - 1) / RW;
            const uint32_t xh_a = (uint32_t)__cvta_generic_to_shared(xh), xl_a = (uint32_t)__cvta_generic_to_shared(xl);
            constexpr int KC_UP = (2 * CIN) / 16, NT = C / 8;
            float2 bup[NT];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) bup[nt] = __ldg(reinterpret_cast<const float2*>(a.b_up + nt * 8 + 2 * (tid & 3)));
            for (int item = warp; item < U * tiles; item += kThreads / 32) {
                const int r = item / tiles, tile = item - r * tiles;
                const uint2* wh = a.upf_h + (size_t)r * KC_UP * NT * 32;
                const uint2* wl = a.upf_l + (size_t)r * KC_UP * NT * 32;
                mma_rows<CIN, C, 2, MT>(xh_a, xl_a, 1, wh, wl, 1, tile * RW, [&](int row, int nt, int co, float v0, float v1) {
                    if (row < rows)
                        *reinterpret_cast<float2*>(cur + (U * row + r) * PF + co) = make_float2(v0 + bup[nt].x, v1 + bup[nt].y);
                });
            }
        }
        __syncthreads();   // xh/xl (aliasing s1) are dead from here on

        // ---- AMP block: 3 x (snake, dilated conv, snake, conv, residual) ----
        const uint32_t s1h_a = (uint32_t)__cvta_generic_to_shared(s1h), s1l_a = (uint32_t)__cvta_generic_to_shared(s1l);
        const uint32_t s2h_a = (uint32_t)__cvta_generic_to_shared(s2h), s2l_a = (uint32_t)__cvta_generic_to_shared(s2l);
        const int tiles = (TT + RW - 1) / RW;
        int c1off = 0;
#pragma unroll 1
        for (int l = 0; l < 3; ++l) {
            const int d = a.dil[l];
            const int ctx = (K - 1) * d;
            {   // restore the saved contexts in front of the tiles, then x -> snake -> s1 tile
                for (int i = tid; i < ctx * PW; i += kThreads) {
                    s1h[(CTX1 - ctx) * PW + i] = c1h[c1off * PW + i];
                    s1l[(CTX1 - ctx) * PW + i] = c1l[c1off * PW + i];
                }
                for (int i = tid; i < CTX2 * PW; i += kThreads) {
                    s2h[i] = c2h[l * CTX2 * PW + i];
                    s2l[i] = c2l[l * CTX2 * PW + i];
                }
                constexpr int CP = C / 2;                       // kThreads % CP == 0: a thread keeps its channel pair
                const int cp = tid % CP;
                const float2 ea = __ldg(reinterpret_cast<const float2*>(a.ea[2 * l] + 2 * cp));
                const float2 ieb = __ldg(reinterpret_cast<const float2*>(a.ieb[2 * l] + 2 * cp));
                for (int row = tid / CP; row < TT; row += kThreads / CP) {
                    const float2 v = *reinterpret_cast<const float2*>(cur + row * PF + 2 * cp);
                    uint32_t hi, lw;
                    split_pair(snake_fast(v.x, ea.x, ieb.x), snake_fast(v.y, ea.y, ieb.y), hi, lw);
                    s1h[(CTX1 + row) * PW + cp] = hi;
                    s1l[(CTX1 + row) * PW + cp] = lw;
                }
            }
            __syncthreads();
            for (int i = tid; i < ctx * PW; i += kThreads) {      // history for the next tile
                c1h[c1off * PW + i] = s1h[(CTX1 + TT - ctx) * PW + i];
                c1l[c1off * PW + i] = s1l[(CTX1 + TT - ctx) * PW + i];
            }
            {
                constexpr int NT = C / 8;
                float2 bv[NT], ea[NT], ieb[NT];                 // per-thread constants: channels 8 nt + 2 q, +1
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const int co = nt * 8 + 2 * (tid & 3);
                    bv[nt] = __ldg(reinterpret_cast<const float2*>(a.b1[l] + co));
                    ea[nt] = __ldg(reinterpret_cast<const float2*>(a.ea[2 * l + 1] + co));
                    ieb[nt] = __ldg(reinterpret_cast<const float2*>(a.ieb[2 * l + 1] + co));
                }
                for (int tile = warp; tile < tiles; tile += kThreads / 32) {
                    mma_rows<C, C, K, MT>(s1h_a, s1l_a, CTX1, a.f1h[l], a.f1l[l], d, tile * RW,
                                          [&](int row, int nt, int co, float v0, float v1) {
                                              uint32_t hi, lw;
                                              split_pair(snake_fast(v0 + bv[nt].x, ea[nt].x, ieb[nt].x),
                                                         snake_fast(v1 + bv[nt].y, ea[nt].y, ieb[nt].y), hi, lw);
                                              s2h[(CTX2 + row) * PW + (co >> 1)] = hi;
                                              s2l[(CTX2 + row) * PW + (co >> 1)] = lw;
                                          });
                }
            }
            __syncthreads();
            for (int i = tid; i < CTX2 * PW; i += kThreads) {
                c2h[l * CTX2 * PW + i] = s2h[TT * PW + i];
                c2l[l * CTX2 * PW + i] = s2l[TT * PW + i];
            }
            {
                constexpr int NT = C / 8;
                float2 bv[NT];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) bv[nt] = __ldg(reinterpret_cast<const float2*>(a.b2[l] + nt * 8 + 2 * (tid & 3)));
                for (int tile = warp; tile < tiles; tile += kThreads / 32) {
                    mma_rows<C, C, K, MT>(s2h_a, s2l_a, CTX2, a.f2h[l], a.f2l[l], 1, tile * RW,
                                          [&](int row, int nt, int co, float v0, float v1) {
                                              float2* p = reinterpret_cast<float2*>(cur + row * PF + co);
                                              float2 c = *p;
                                              c.x += v0 + bv[nt].x;
                                              c.y += v1 + bv[nt].y;
                                              *p = c;
                                          });
                }
            }
            __syncthreads();
            c1off += ctx;
        }

        if (POST) {
            // ---- fused tail: the other two resblocks' partials are complete (their kernels ran before this one) ----
            constexpr int V = C / 4;
            const float* p1 = a.post_p1 + (size_t)b * a.n_out * C;
            const float* p2 = a.post_p2 + (size_t)b * a.n_out * C;
            for (int i = tid; i < TT * V; i += kThreads) {
                const int tt = i / V, c4 = i - tt * V;
                const int tg = t0 + tt;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (tg < a.n_out) {
                    const float4 x0 = *reinterpret_cast<const float4*>(cur + tt * PF + c4 * 4);
                    const float4 x1 = __ldg(reinterpret_cast<const float4*>(p1 + (size_t)tg * C) + c4);
                    const float4 x2 = __ldg(reinterpret_cast<const float4*>(p2 + (size_t)tg * C) + c4);
                    const float4 ea = __ldg(reinterpret_cast<const float4*>(a.post_ea) + c4);
                    const float4 ieb = __ldg(reinterpret_cast<const float4*>(a.post_ieb) + c4);
                    v.x = snake(((x0.x + x1.x) + x2.x) / 3.0f, ea.x, ieb.x);
                    v.y = snake(((x0.y + x1.y) + x2.y) / 3.0f, ea.y, ieb.y);
                    v.z = snake(((x0.z + x1.z) + x2.z) / 3.0f, ea.z, ieb.z);
                    v.w = snake(((x0.w + x1.w) + x2.w) / 3.0f, ea.w, ieb.w);
                }
                *reinterpret_cast<float4*>(pact + (6 + tt) * C + c4 * 4) = v;
            }
            __syncthreads();
            const float bias = __ldg(a.post_bias);
            for (int tt = tid; tt < TT; tt += kThreads) {
                const int tg = t0 + tt;
                if (t0 < t_begin || tg >= t_end || tg >= a.n_wav) continue;
                float acc = bias;
#pragma unroll
                for (int c = 0; c < C; ++c)
#pragma unroll
                    for (int j = 0; j < 7; ++j) acc = fmaf(__ldg(a.post_w + c * 7 + j), pact[(tt + j) * C + c], acc);
                a.wav[(size_t)b * a.n_wav + tg] = tanhf(acc) / a.post_inv_scale_div;
            }
            __syncthreads();
            for (int i = tid; i < 6 * C; i += kThreads) pact[i] = pact[TT * C + i];      // history for the next tile
        }
        // ---- write the tile, channel-last, coalesced (warm-up tiles are not written) ----
        if (!POST && t0 >= t_begin) {
            constexpr int V = C / 4;
            for (int i = tid; i < TT * V; i += kThreads) {
                const int tt = i / V, c4 = i - tt * V;
                const int tg = t0 + tt;
                if (tg < t_end)
                    *reinterpret_cast<float4*>(dst + (size_t)tg * C + c4 * 4) =
                        *reinterpret_cast<const float4*>(cur + tt * PF + c4 * 4);
            }
        }
        // the next tile's input load overwrites xh/xl = s1, which the last conv no longer reads; cur is rewritten
        // only after the barrier that follows the input load
    }
}

// ===========================================================================
// precision 1, C <= 32: stage kernel on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM)
// ===========================================================================
// One CTA owns (utterance, time range) and runs ALL THREE resblocks (k = 11, 7, 3) of the stage on tiles of 128
// output samples, streaming over time like stage_stream_kernel (causal contexts carried in shared memory).
//   implicit GEMM   M = 128 time rows, N = C_out (16 for C = 8, zero padded), K = taps x C_in in k16 steps.
//                   Conv inputs live in shared memory as split-bf16 NO-SWIZZLE K-major operands laid out
//                   [hi | lo][8-channel group][row][8 channels]: rows are 16 bytes apart, so a tap is nothing but a
//                   shift of the descriptor's start address (LBO = group stride, SBO = 128 B).  For C = 8 one k16
//                   step covers two taps: the second K group is the same buffer d rows later (LBO = 16 d).
//   residual        the fp32 residual stream x of each resblock lives in TMEM; the second conv of a layer accumulates
//                   straight onto it (x += conv2(...)); biases of the second convs are added when x is read
//   roles           warp 0 issues the MMAs of the 18 jobs of a tile (3 resblocks x 3 layers x 2 convs) round-robin over
//                   the resblocks; warp 1 streams the weights of the jobs in issue order (cp.async.bulk, 16 KiB
//                   chunks, 3-slot ring); warps 4-7 are the epilogue (thread = time row): TMEM -> bias -> SnakeBeta ->
//                   split bf16 -> next conv's operand buffer.  While the epilogue of one resblock runs, the tensor core
//                   works on the other two.
//   tile prologue   input tile (mean of the producer's partials) -> ConvTranspose1d with mma.sync (3 % of the MACs) ->
//                   x of the three resblocks (tcgen05.st) and their first conv inputs
constexpr int UM_ROWS = 128;                                    // rows of one MMA (UMMA M)
constexpr int UM_PADR = 8;
constexpr int UM_WSLOTS = 3;
constexpr int UM_WSLOT_BYTES = 16384;
__host__ __device__ constexpr int um_K(int c) { return c == 0 ? 11 : c == 1 ? 7 : 3; }
__host__ __device__ constexpr int um_R1(int c, int TT) { return (um_K(c) - 1) * 5 + TT + UM_PADR; }
__host__ __device__ constexpr int um_R2(int c, int TT) { return (um_K(c) - 1) + TT + UM_PADR; }

// MT  = 128-row MMA tiles per CTA tile: a conv job issues MT MMAs per k16 step and term, and its epilogue covers 128 MT
//       rows with MT times the threads, so the per-job latency (~1 us: TMEM read, barrier, SnakeBeta, proxy fence,
//       mbarrier round trips) is spread over more samples.  C = 32 is MMA-bound already at MT = 1; C = 16 needs MT = 2.
// NCH = resblocks ("chains") per CTA: 3 = all of them, interleaved (their mean is written); 1 = the resblock blockIdx.z
//       (C = 64: three chains do not fit in shared memory; each CTA then writes its resblock's partial tensor).
template <int C, int U, int MT, int NCH>
struct UmLayout {
    static constexpr int TT = UM_ROWS * MT;
    static constexpr int G = C / 8, N = C < 16 ? 16 : C;
    static constexpr int CPT = C >= 16 ? 16 : 8;                // channels per epilogue thread
    static constexpr int GE = C / CPT;                           // sets of 4 epilogue warps per 128-row tile
    // NG epilogue groups work on different resblocks at the same time: group 0 owns the k = 11 resblock, group 1 the other
    // two (6 and 12 jobs per tile; one group for all 18 made the epilogue, not the tensor core, the limiter of a tile)
    static constexpr int NG = NCH == 3 ? 2 : 1;
    // byte offsets; with NCH = 1 the single chain is sized for the largest kernel (local index 0 = k 11)
    __host__ __device__ static constexpr int a1(int c) { return c == 0 ? 0 : a1(c - 1) + 2 * G * um_R1(c - 1, TT) * 16; }
    __host__ __device__ static constexpr int a2(int c) { return c == 0 ? a1(NCH) : a2(c - 1) + 2 * G * um_R2(c - 1, TT) * 16; }
    // saved contexts: per chain [layer][part][group][row][16 B]; layer l of the dilated conv has (K-1) d_l rows
    __host__ __device__ static constexpr int c1(int c) { return c == 0 ? a2(NCH) : c1(c - 1) + 2 * G * (um_K(c - 1) - 1) * 9 * 16; }
    __host__ __device__ static constexpr int c2(int c) { return c == 0 ? c1(NCH) : c2(c - 1) + 2 * G * (um_K(c - 1) - 1) * 3 * 16; }
    static constexpr int wring = (c2(NCH) + 1023) / 1024 * 1024;
    static constexpr int x0 = wring + UM_WSLOTS * UM_WSLOT_BYTES;          // fp32 [TT][C + 4]: the upsampled tile (bulk-copied)
    static constexpr int cst = x0 + TT * (C + 4) * 4;                      // per-channel constants of the jobs: [chain][20][C] fp32
    static constexpr int total = cst + NCH * 20 * C * 4;
    // TMEM columns per 128-row tile and chain: [x | x aux | D1 | D1 aux], N each.  A product is a_hi [w_hi | w_lo] (one MMA
    // of width 2 N: main and aux accumulator) + a_lo w_hi (width N): the activation operand, whose fetch from shared
    // memory bounds these narrow MMAs, is read twice per k16 step instead of three times; the epilogue adds main + aux.
    static constexpr int tile_cols = 4 * N * NCH;
    static constexpr int scratch_col = tile_cols * MT;                     // N columns per row tile: group 0's output sum for group 1
    static constexpr int cols_used = scratch_col + (NG == 2 ? N * MT : 0);
    static constexpr int tmem_cols = cols_used <= 128 ? 128 : cols_used <= 256 ? 256 : 512;
    static_assert(cols_used <= 512, "tensor memory");
    static constexpr int threads = 128 + NG * 128 * GE * MT;
};

// The transposed convolution that opens a stage runs as its own kernel (upsample_kernel) and leaves x0 [B, n_out, C] in
// global memory; the stage kernel bulk-copies its tile of x0 one tile ahead.
struct UpsampleArgs {
    const float* in_p[3];       // the producer's partials [B, n_in, 2 C] channel-last (n_parts = 1: a single tensor)
    int n_parts;
    long long in_bstride;
    int n_in, n_out;
    const float* b_up;
    const uint2* upf_h;         // per output phase r: fragment-packed 2-tap weights, hi / lo parts
    const uint2* upf_l;
    float* x0;                  // [B, n_out, C]
};

struct UmmaStageArgs {
    const float* x0;            // [B, n_out, C] (+ one tile of readable slack)
    int n_out;
    const float* ea[3][6];
    const float* ieb[3][6];
    float* out[3];              // NCH = 3: out[0] = [B, n_out, C] mean of the resblocks; NCH = 1: out[chain] = its partial
    unsigned long long* trace;  // bring-up: CTA (0,0), tile 2: per job 4 %globaltimer stamps (+ 2 for the tile prologue)
    int dbg;                    // BVC_VOC_DEBUG timing probes (wrong results; only in builds with -DBVC_VOC_PROBES, the checks cost ~1 ms per
                                // vocoder pass): 1 no MMAs, 2 no weight copies, 4 no SnakeBeta / split maths
    UmmaStageWeights w;
};

#ifdef BVC_VOC_PROBES
#define UM_DBG(a, bit) ((a).dbg & (bit))
#else
#define UM_DBG(a, bit) 0
#endif
__device__ __forceinline__ uint32_t um_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool um_try(uint32_t bar, uint32_t parity) {
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
// bounded: a protocol bug traps (surfaces as a CUDA error at the next API call) instead of hanging the GPU
__device__ __forceinline__ void um_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t b = um_smem_u32(bar);
    if (um_try(b, parity)) return;
    const long long t0 = clock64();
    while (!um_try(b, parity)) {
        if (clock64() - t0 > 4000000000LL) asm volatile("trap;\n");
    }
}
__device__ __forceinline__ unsigned long long um_ns() {
    unsigned long long v;
    asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(v));
    return v;
}
// job timeline of one tile (BVC_VOC_TRACE=<file>): compiled in only with -DBVC_VOC_TRACING, the run-time checks sit in the
// MMA-issue and epilogue loops of every job
#ifdef BVC_VOC_TRACING
#define UM_TRACE(slot)                                                                         \
    do {                                                                                       \
        if (a.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tile == 2) a.trace[slot] = um_ns(); \
    } while (0)
#else
#define UM_TRACE(slot) do { } while (0)
#endif
__device__ __forceinline__ void um_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(um_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool um_elect() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}
// no-swizzle K-major descriptor: LBO = byte stride between the two 8-wide K groups of a k16 step, SBO = 128 B (8 rows)
__device__ __forceinline__ uint64_t um_desc(uint32_t addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(128 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ void um_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void um_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(um_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void um_ld8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void um_st8(uint32_t taddr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(taddr), "r"(__float_as_uint(v[0])),
                 "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
                 "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}

// 8 channels of one row -> 16 bytes of the hi part and of the lo part
__device__ __forceinline__ void um_store8(unsigned char* part_hi, int part_stride, int offset16, const float* y) {
    uint4 h, l;
    split_pair(y[0], y[1], h.x, l.x); split_pair(y[2], y[3], h.y, l.y);
    split_pair(y[4], y[5], h.z, l.z); split_pair(y[6], y[7], h.w, l.w);
    *reinterpret_cast<uint4*>(part_hi + (size_t)offset16 * 16) = h;
    *reinterpret_cast<uint4*>(part_hi + part_stride + (size_t)offset16 * 16) = l;
}

// ConvTranspose1d(2 C -> C, kernel 2 U, stride U) as U phase convolutions with 2 taps each (mma.sync, split bf16):
// out[U j + r] = W[r + U] in[j - 1] + W[r] in[j].  A CTA owns (utterance, JT input rows); a warp keeps one phase so
// that its weight fragments stay in L1 over its row tiles.
template <int C, int U>
__global__ void __launch_bounds__(256) upsample_kernel(UpsampleArgs a) {
    constexpr int CIN = 2 * C, PWI = RowLayout<CIN>::PW, JT = 512 / U, MTU = 2, NT = C / 8, KC_UP = (2 * CIN) / 16;
    __shared__ __align__(16) uint32_t xh[(JT + 1) * PWI];
    __shared__ __align__(16) uint32_t xl[(JT + 1) * PWI];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int b = blockIdx.y, j0 = blockIdx.x * JT;
    const size_t boff = (size_t)b * a.in_bstride;
    constexpr int V = CIN / 4;
    for (int i = tid; i < (JT + 1) * V; i += 256) {
        const int jj = i / V, c4 = i - jj * V;
        const int j = j0 - 1 + jj;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j >= 0 && j < a.n_in) {
            const size_t o = boff + (size_t)j * CIN + c4 * 4;
            v = __ldg(reinterpret_cast<const float4*>(a.in_p[0] + o));
            if (a.n_parts == 3) {
                const float4 v1 = __ldg(reinterpret_cast<const float4*>(a.in_p[1] + o));
                const float4 v2 = __ldg(reinterpret_cast<const float4*>(a.in_p[2] + o));
                v.x = ((v.x + v1.x) + v2.x) / 3.0f; v.y = ((v.y + v1.y) + v2.y) / 3.0f;
                v.z = ((v.z + v1.z) + v2.z) / 3.0f; v.w = ((v.w + v1.w) + v2.w) / 3.0f;
            }
        }
        uint32_t h0, l0, h1, l1;
        split_pair(v.x, v.y, h0, l0);
        split_pair(v.z, v.w, h1, l1);
        *reinterpret_cast<uint2*>(xh + jj * PWI + c4 * 2) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(xl + jj * PWI + c4 * 2) = make_uint2(l0, l1);
    }
    __syncthreads();
    float2 bup[NT];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) bup[nt] = __ldg(reinterpret_cast<const float2*>(a.b_up + nt * 8 + 2 * (tid & 3)));
    const uint32_t xh_a = (uint32_t)__cvta_generic_to_shared(xh), xl_a = (uint32_t)__cvta_generic_to_shared(xl);
    constexpr int RT = JT / (16 * MTU);                 // row tiles per phase
    float* dst = a.x0 + (size_t)b * a.n_out * C;
    for (int item = warp; item < U * RT; item += 8) {   // U = 8: phase = warp; U = 2: phase = warp & 1
        const int r = item % U, tl = item / U;
        const uint2* wh = a.upf_h + (size_t)r * KC_UP * NT * 32;
        const uint2* wl = a.upf_l + (size_t)r * KC_UP * NT * 32;
        mma_rows<CIN, C, 2, MTU>(xh_a, xl_a, 1, wh, wl, 1, tl * 16 * MTU, [&](int row, int nt, int co, float v0, float v1) {
            const long long t = (long long)U * (j0 + row) + r;
            if (t < a.n_out) *reinterpret_cast<float2*>(dst + t * C + co) = make_float2(v0 + bup[nt].x, v1 + bup[nt].y);
        });
    }
}

template <int C, int U>
int launch_upsample(const UpsampleArgs& a, int B, cudaStream_t stream) {
    constexpr int JT = 512 / U;
    upsample_kernel<C, U><<<dim3((a.n_in + 1 + JT - 1) / JT, B), 256, 0, stream>>>(a);
    BVC_CHECK_LAUNCH();
    return BVC_OK;
}

template <int C, int U, int MT, int NCH>
__global__ void __launch_bounds__(UmLayout<C, U, MT, NCH>::threads, 1) stage_umma_kernel(UmmaStageArgs a) {
    using L = UmLayout<C, U, MT, NCH>;
    constexpr int UM_TT = L::TT;
    constexpr int G = L::G, N = L::N, PX = C + 4, CPT = L::CPT, GE = L::GE;
    constexpr int S8 = CPT / 8;                        // 8-channel operand groups per epilogue thread
    constexpr int HALO = 12 * (11 - 1);
    constexpr int UM_THREADS = L::threads;             // warps 0-3: MMA issue, weight stream, 2 idle; then GE MT epilogue groups of 4 warps
    extern __shared__ __align__(1024) unsigned char um_smem[];
    __shared__ __align__(8) uint64_t a_ready[3], d_ready[3], w_full[UM_WSLOTS], w_empty[UM_WSLOTS], x_full, x_empty, o_ready, o_free;
    __shared__ uint32_t tmem_slot;
    unsigned char* sm = um_smem;
    const uint32_t sm_a = um_smem_u32(um_smem);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y;
    const int kc0 = NCH == 1 ? (int)blockIdx.z : 0;    // first resblock of this CTA; local chain lc <-> resblock kc0 + lc
    const int tiles_total = (a.n_out + UM_TT - 1) / UM_TT;
    const int tiles_per = (tiles_total + gridDim.x - 1) / gridDim.x;
    const int t_begin = blockIdx.x * tiles_per * UM_TT;
    const int t_end = min(a.n_out, t_begin + tiles_per * UM_TT);
    if (t_begin >= t_end) return;
    const int t_first = max(0, t_begin - ((HALO + UM_TT - 1) / UM_TT) * UM_TT);   // warm-up tiles (outputs discarded)
    const int n_tiles = (t_end - t_first + UM_TT - 1) / UM_TT;

    // zero everything once: causal zero history, and the pad rows that zero-weight taps may touch must be finite
    for (int i = tid; i < L::wring / 16; i += UM_THREADS) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
    // Per-channel constants (biases, SnakeBeta exp(alpha), 1 / exp(beta)) of every job, once per CTA: vector v of chain c at
    // cst[(c * 20 + v) * C].  v = 0, 1: first activation; 2 + 6 l + {0, 1, 2}: bias / ea / ieb behind the dilated conv of
    // layer l; 2 + 6 l + {3, 4, 5}: bias sum / ea / ieb behind its second conv.  Read with ld.shared in the job loop: a
    // global load in flight would hold up the epilogue's next shared- and tensor-memory reads.
    {
        float* cst = reinterpret_cast<float*>(sm + L::cst);
        for (int i = tid; i < NCH * 20 * C; i += UM_THREADS) {
            const int c = i / (20 * C), r = i - c * 20 * C, v = r / C, ch = r - v * C, kc = kc0 + c;
            const float* src = nullptr;
            if (v == 0) src = a.ea[kc][0];
            else if (v == 1) src = a.ieb[kc][0];
            else {
                const int l = (v - 2) / 6, w = (v - 2) % 6;
                src = w == 0 ? a.w.b1[kc][l] : w == 1 ? a.ea[kc][2 * l + 1] : w == 2 ? a.ieb[kc][2 * l + 1]
                    : w == 3 ? a.w.bsum[kc][l] : l < 2 ? (w == 4 ? a.ea[kc][2 * (l + 1)] : a.ieb[kc][2 * (l + 1)]) : nullptr;
            }
            cst[i] = src ? __ldg(src + ch) : 0.f;
        }
    }
    if (tid == 0) {
        for (int i = 0; i < 3; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(um_smem_u32(&a_ready[i])), "r"(4 * GE * MT));   // the owning group's warps
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(um_smem_u32(&d_ready[i])), "r"(1));
        }
        for (int i = 0; i < UM_WSLOTS; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(um_smem_u32(&w_full[i])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(um_smem_u32(&w_empty[i])), "r"(1));
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(um_smem_u32(&x_full)), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(um_smem_u32(&x_empty)), "r"(L::NG * 4 * GE * MT));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(um_smem_u32(&o_ready)), "r"(4 * GE * MT));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(um_smem_u32(&o_free)), "r"(4 * GE * MT));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(um_smem_u32(&tmem_slot)),
                     "r"((uint32_t)L::tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = tmem_slot;

    if (warp == 1) {
        // =========================== weight stream ===========================
        uint32_t it = 0;
        for (int tile = 0; tile < n_tiles; ++tile) {
#pragma unroll 1
            for (int ji = 0; ji < 18; ++ji) {
                const UmmaJob jb = a.w.jobs[ji];
                if (jb.chain < kc0 || jb.chain >= kc0 + NCH) continue;
                constexpr int SPCW = UM_WSLOT_BYTES / (N * 64);
                for (int s0 = 0; s0 < jb.steps; s0 += SPCW, ++it) {
                    const int slot = it % UM_WSLOTS, round = it / UM_WSLOTS;
                    if (round >= 1) um_wait(&w_empty[slot], (round - 1) & 1);
                    if (UM_DBG(a, 2)) {
                        if (um_elect()) um_arrive(&w_full[slot]);
                    } else if (um_elect()) {
                        const uint32_t bytes = (uint32_t)(min(SPCW, jb.steps - s0) * N * 64);
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(um_smem_u32(&w_full[slot])), "r"(bytes)
                                     : "memory");
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                                         sm_a + L::wring + slot * UM_WSLOT_BYTES),
                                     "l"(a.w.wstream + jb.off + s0 * N * 64), "r"(bytes), "r"(um_smem_u32(&w_full[slot]))
                                     : "memory");
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 2) {
        // =========================== x0 tiles: one bulk copy per row (rows are padded in shared memory), one tile ahead ===========================
        const float* src = a.x0 + (size_t)b * a.n_out * C;
        for (int tile = 0; tile < n_tiles; ++tile) {
            const int t0 = t_first + tile * UM_TT;
            if (tile >= 1) um_wait(&x_empty, (tile - 1) & 1);
            if (lane == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(um_smem_u32(&x_full)), "r"((uint32_t)(UM_TT * C * 4))
                             : "memory");
            __syncwarp();
            for (int r = lane; r < UM_TT; r += 32)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                                 sm_a + L::x0 + r * PX * 4),
                             "l"(src + (size_t)(t0 + r) * C), "r"((uint32_t)(C * 4)), "r"(um_smem_u32(&x_full))
                             : "memory");
        }
    } else if (warp == 0) {
        // =========================== MMA issue ===========================
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(UM_ROWS >> 4) << 24);
        constexpr uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((2 * N) >> 3) << 17) | ((uint32_t)(UM_ROWS >> 4) << 24);
        constexpr int SPC = UM_WSLOT_BYTES / (N * 64);       // k16 steps per weight chunk
        uint32_t wit = 0;
        uint32_t jobs_done[3] = {0, 0, 0};                    // per chain: selects the parity of a_ready
        for (int tile = 0; tile < n_tiles; ++tile) {
#pragma unroll 1
            for (int ji = 0; ji < 18; ++ji) {
                const UmmaJob jb = a.w.jobs[ji];
                if (jb.chain < kc0 || jb.chain >= kc0 + NCH) continue;
                const int c = jb.chain - kc0;
                um_wait(&a_ready[c], jobs_done[c] & 1);
                ++jobs_done[c];
                if (lane == 0) UM_TRACE(ji * 4 + 0);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const int K = jb.K, d = jb.d;
                const int R = jb.conv2 ? (K - 1) + UM_TT + UM_PADR : (K - 1) * 5 + UM_TT + UM_PADR;
                const int lead = jb.conv2 ? (K - 1) : (K - 1) * 5;
                const uint32_t abase = sm_a + (jb.conv2 ? (c == 0 ? L::a2(0) : c == 1 ? L::a2(1) : L::a2(2))
                                                        : (c == 0 ? L::a1(0) : c == 1 ? L::a1(1) : L::a1(2)));
                const uint32_t d_tmem = tmem + (uint32_t)(c * 4 * N + (jb.conv2 ? 0 : 2 * N));     // row tile mt: + mt * tile_cols
                // Descriptors advance by plain additions on their 16-byte-unit address field:
                //   C >= 16: step (tap, channel pair gq): +d rows per tap, +2 R rows per channel pair (two 8-channel groups)
                //   C == 8 : step s covers taps 2s, 2s+1: +2 d rows per step, LBO = d rows
                constexpr int SPT = C >= 16 ? C / 16 : 1;
                const uint64_t dA0 = um_desc(abase + (uint32_t)((lead - (K - 1) * d) * 16), C >= 16 ? (uint32_t)(R * 16) : (uint32_t)(d * 16));
                const uint32_t a_lo16 = (uint32_t)(G * R);                       // lo part, in 16-byte units
                const uint32_t step_tap16 = C >= 16 ? (uint32_t)d : (uint32_t)(2 * d);
                uint32_t tap_off16 = 0, gq = 0;                                  // running position of the next step
                uint32_t first = jb.conv2 ? 1u : 0u;                             // the second conv accumulates onto x in TMEM
                for (int s0 = 0; s0 < jb.steps; s0 += SPC) {
                    const int slot = wit % UM_WSLOTS, round = wit / UM_WSLOTS;
                    um_wait(&w_full[slot], round & 1);
                    ++wit;
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                    const int s1 = min(jb.steps, s0 + SPC);
                    if (um_elect()) {
                        // weight step: [k group][hi rows 0..N-1 | lo rows N..2N-1][8]: one descriptor serves the stacked and the hi-only MMA
                        uint64_t dwh = um_desc(sm_a + L::wring + slot * UM_WSLOT_BYTES, N * 32);
#pragma unroll 2
                        for (int s = s0; s < s1; ++s) {
                            const uint64_t dah0 = dA0 + (uint64_t)(tap_off16 + gq * (uint32_t)(2 * R));
#pragma unroll
                            for (int mt = 0; mt < MT; ++mt) {       // the weight step is fetched once per 128-row tile
                                const uint64_t dah = dah0 + (uint64_t)(mt * UM_ROWS), dal = dah + a_lo16;
                                const uint32_t dt = d_tmem + (uint32_t)(mt * L::tile_cols);
                                if (!UM_DBG(a, 1)) {
                                    um_mma(dt, dah, dwh, idesc2, first);
                                    um_mma(dt, dal, dwh, idesc, 1u);
                                }
                            }
                            first = 1u;
                            dwh += (uint64_t)(N * 4);
                            if (SPT == 1 || ++gq == SPT) { gq = 0; tap_off16 += step_tap16; }
                        }
                        um_commit(&w_empty[slot]);
                        if (s1 == jb.steps) um_commit(&d_ready[c]);
                    }
                    __syncwarp();
                    {   // every lane tracks the position the elected lane has advanced to
                        const uint32_t steps_done = (uint32_t)s1;
                        gq = steps_done % SPT;
                        tap_off16 = (steps_done / SPT) * step_tap16;
                        first = 1u;
                    }
                }
                if (lane == 0) UM_TRACE(ji * 4 + 1);
            }
        }
    } else if (warp >= 4) {
        // =========================== epilogue (thread = time row x CPT channels) ===========================
        // GE MT groups of 4 warps; a group owns CPT channels (S8 operand groups of 8) of the 128 rows of one MMA tile (a
        // warp may only touch the TMEM lane quadrant warp % 4, so each group is one full set of quadrants)
        constexpr int NE = 128 * GE * MT, NG = L::NG;  // threads per epilogue group, groups
        const int grp = (tid - 128) / NE;
        const int et = (tid - 128) % NE, wg = et >> 7, ge = wg % GE, mt = wg / GE, quad = warp & 3;
        auto owner = [&](int c) { return NG == 1 ? 0 : (c == 0 ? 0 : 1); };   // epilogue group of local resblock c
        const int g0 = ge * S8;                        // first 8-channel operand group of this thread
        const int pl = quad * 32 + lane;               // TMEM lane = row of the 128-row MMA tile
        const int p = mt * UM_ROWS + pl;               // row of the CTA tile
        const uint32_t t_lane = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(mt * L::tile_cols);
        (void)pl;
        float* x0 = reinterpret_cast<float*>(sm + L::x0);
        uint32_t d_seen[3] = {0, 0, 0};
        auto epi_bar = [&]() { asm volatile("bar.sync %0, %1;\n" ::"r"(1 + grp), "n"(NE) : "memory"); };   // this group only

        // writes this thread's CPT channels of its row (already activated) into operand buffer `buf` (rows R, tile starts
        // at row `lead`), and the tail rows additionally into the saved context `ctx` (ctx_rows rows)
        auto write_rows = [&](int buf_off, int R, int lead, const float* y, int ctx_off, int ctx_rows) {
            const int cr = p - (UM_TT - ctx_rows);
#pragma unroll
            for (int s8 = 0; s8 < S8; ++s8) {
                um_store8(sm + buf_off, G * R * 16, (g0 + s8) * R + lead + p, y + 8 * s8);
                if (cr >= 0) um_store8(sm + ctx_off, G * ctx_rows * 16, (g0 + s8) * ctx_rows + cr, y + 8 * s8);
            }
        };
        // copies the saved context (ctx_rows rows) in front of the tile rows of an operand buffer
        auto restore = [&](int buf_off, int R, int lead, int ctx_off, int ctx_rows) {
            if (p < ctx_rows) {
#pragma unroll
                for (int part = 0; part < 2; ++part)
#pragma unroll
                    for (int s8 = 0; s8 < S8; ++s8)
                        *reinterpret_cast<uint4*>(sm + buf_off + part * G * R * 16 + ((g0 + s8) * R + lead - ctx_rows + p) * 16) =
                            *reinterpret_cast<const uint4*>(sm + ctx_off + part * G * ctx_rows * 16 + ((g0 + s8) * ctx_rows + p) * 16);
            }
        };
        auto publish = [&](int c) {   // operand buffer of chain c is complete: hand it to the MMA warp
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            __syncwarp();
            if (lane == 0) um_arrive(&a_ready[c]);
        };
        const float* cst = reinterpret_cast<const float*>(sm + L::cst);
        auto loadc = [&](int c, int v, float* o) {          // this thread's CPT per-channel constants (vector v of chain c)
#pragma unroll
            for (int q = 0; q < CPT / 4; ++q) {
                const float4 u = *(reinterpret_cast<const float4*>(cst + (c * 20 + v) * C + CPT * ge) + q);
                o[4 * q] = u.x; o[4 * q + 1] = u.y; o[4 * q + 2] = u.z; o[4 * q + 3] = u.w;
            }
        };
        auto tld = [&](uint32_t taddr, float* o) {          // CPT accumulator columns: main + aux
#pragma unroll
            for (int s8 = 0; s8 < S8; ++s8) {
                float aux[8];
                um_ld8(taddr + 8 * s8, o + 8 * s8);
                um_ld8(taddr + N + 8 * s8, aux);
#pragma unroll
                for (int i = 0; i < 8; ++i) o[8 * s8 + i] += aux[i];
            }
        };
        auto a1_off = [&](int c) { return c == 0 ? L::a1(0) : c == 1 ? L::a1(1) : L::a1(2); };
        auto a2_off = [&](int c) { return c == 0 ? L::a2(0) : c == 1 ? L::a2(1) : L::a2(2); };
        // context of the dilated conv of layer l: rows (K-1) d_l, stored after those of the earlier layers
        auto c1_off = [&](int c, int K, int l) {
            const int rows_before = (K - 1) * (l == 0 ? 0 : l == 1 ? 1 : 4);
            return (c == 0 ? L::c1(0) : c == 1 ? L::c1(1) : L::c1(2)) + 2 * G * rows_before * 16;
        };
        auto c2_off = [&](int c, int K, int l) { return (c == 0 ? L::c2(0) : c == 1 ? L::c2(1) : L::c2(2)) + 2 * G * (K - 1) * l * 16; };

        // Tile prologue of one resblock, second half: x = x0 row (TMEM, next to a zeroed aux accumulator), first conv
        // input = SnakeBeta(x).  The first half (restore of the layer-0 context) must be separated from it by an epi_bar.
        auto chain_begin = [&](int c) {
            const int kc = kc0 + c, K = um_K(kc);
            float xr[CPT];
            const float zero8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int q = 0; q < CPT / 4; ++q) {
                const float4 u = *reinterpret_cast<const float4*>(x0 + p * PX + CPT * ge + 4 * q);
                xr[4 * q] = u.x; xr[4 * q + 1] = u.y; xr[4 * q + 2] = u.z; xr[4 * q + 3] = u.w;
            }
#pragma unroll
            for (int s8 = 0; s8 < S8; ++s8) {
                um_st8(t_lane + c * 4 * N + CPT * ge + 8 * s8, xr + 8 * s8);
                um_st8(t_lane + c * 4 * N + N + CPT * ge + 8 * s8, zero8);      // the second convs accumulate onto [x | 0]
            }
            float ea[CPT], ieb[CPT], y[CPT];
            loadc(c, 0, ea);
            loadc(c, 1, ieb);
#pragma unroll
            for (int i = 0; i < CPT; ++i) y[i] = snake_fast(xr[i], ea[i], ieb[i]);
            write_rows(a1_off(c), (K - 1) * 5 + UM_TT + UM_PADR, (K - 1) * 5, y, c1_off(c, K, 0), (K - 1) * 1);
            asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
            publish(c);
        };
        auto chain_restore0 = [&](int c) {
            const int K = um_K(kc0 + c);
            restore(a1_off(c), (K - 1) * 5 + UM_TT + UM_PADR, (K - 1) * 5, c1_off(c, K, 0), (K - 1) * 1);
        };

        // first tile: all resblocks start together.  Later tiles are started resblock by resblock from inside the previous
        // tile's job loop (right after the resblock's last epilogue), so the tensor core never waits for a tile boundary.
#pragma unroll
        for (int c = 0; c < NCH; ++c)
            if (owner(c) == grp) chain_restore0(c);
        um_wait(&x_full, 0);
        epi_bar();                                      // all context rows restored before this tile's tails replace them
#pragma unroll
        for (int c = 0; c < NCH; ++c)
            if (owner(c) == grp) chain_begin(c);
        __syncwarp();
        if (lane == 0) um_arrive(&x_empty);             // the loader may fetch the next tile (once both groups have arrived)

        for (int tile = 0; tile < n_tiles; ++tile) {
            const int t0 = t_first + tile * UM_TT;
            if (et == 0) UM_TRACE(72);
            if (et == 0) UM_TRACE(73);
            float osum[CPT];
#pragma unroll
            for (int i = 0; i < CPT; ++i) osum[i] = 0.f;
#pragma unroll 1
            for (int ji = 0; ji < 18; ++ji) {
                const UmmaJob jb = a.w.jobs[ji];
                if (jb.chain < kc0 || jb.chain >= kc0 + NCH) continue;
                const int kc = jb.chain, c = kc - kc0, l = jb.layer, K = jb.K;
                if (owner(c) != grp) continue;
                float v[CPT], pa[CPT], ea[CPT], ieb[CPT];
                // per-channel constants of this job are fetched while the MMAs are still running
                if (!jb.conv2) {
                    loadc(c, 2 + 6 * l, pa); loadc(c, 3 + 6 * l, ea); loadc(c, 4 + 6 * l, ieb);
                } else {
                    loadc(c, 5 + 6 * l, pa);
                    if (l < 2) { loadc(c, 6 + 6 * l, ea); loadc(c, 7 + 6 * l, ieb); }
                }
                um_wait(&d_ready[c], d_seen[c] & 1);
                ++d_seen[c];
                if (et == 0) UM_TRACE(ji * 4 + 2);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                if (!jb.conv2) {
                    // dilated conv done: + bias -> SnakeBeta -> input of the second conv
                    restore(a2_off(c), (K - 1) + UM_TT + UM_PADR, K - 1, c2_off(c, K, l), K - 1);
                    tld(t_lane + c * 4 * N + 2 * N + CPT * ge, v);
                    epi_bar();
                    if (!UM_DBG(a, 4)) {
#pragma unroll
                        for (int i = 0; i < CPT; ++i) v[i] = snake_fast(v[i] + pa[i], ea[i], ieb[i]);
                        write_rows(a2_off(c), (K - 1) + UM_TT + UM_PADR, K - 1, v, c2_off(c, K, l), K - 1);
                    }
                    publish(c);
                } else {
                    // residual stream updated in TMEM: x_true = x + (sum of the second convs' biases so far)
                    const int ctx = (K - 1) * (l == 0 ? 3 : 5);
                    const bool start_next = l == 2 && tile + 1 < n_tiles;     // this resblock is done with the tile
                    if (l < 2) restore(a1_off(c), (K - 1) * 5 + UM_TT + UM_PADR, (K - 1) * 5, c1_off(c, K, l + 1), ctx);
                    else if (start_next) chain_restore0(c);
                    tld(t_lane + c * 4 * N + CPT * ge, v);
                    if (start_next && (c == 0 || (NG == 2 && c == 1))) um_wait(&x_full, (tile + 1) & 1);   // the next x0 tile (loaded
                                                                                                           // during this one); first use per group
                    epi_bar();
                    if (l < 2) {
                        if (!UM_DBG(a, 4)) {
#pragma unroll
                            for (int i = 0; i < CPT; ++i) v[i] = snake_fast(v[i] + pa[i], ea[i], ieb[i]);
                            write_rows(a1_off(c), (K - 1) * 5 + UM_TT + UM_PADR, (K - 1) * 5, v, c1_off(c, K, l + 1), ctx);
                        }
                        publish(c);
                    } else {
#pragma unroll
                        for (int i = 0; i < CPT; ++i) osum[i] += v[i] + pa[i];
                        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
                        if (start_next) {
                            chain_begin(c);
                            if (c == NCH - 1 || (NG == 2 && c == 0)) {       // this group's last read of the x0 tile
                                __syncwarp();
                                if (lane == 0) um_arrive(&x_empty);
                            }
                        }
                    }
                }
                if (et == 0) UM_TRACE(ji * 4 + 3);
            }
            // ---- channel-last output (warm-up tiles are not written): the mean of the three resblocks, or this one's partial ----
            const uint32_t t_scr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(L::scratch_col + mt * N + CPT * ge);
            if (NG == 2 && grp == 0) {
                // group 0 (k = 11 resblock) finishes first: its sum goes to group 1 through tensor memory (same lane, same columns)
                if (tile >= 1) um_wait(&o_free, (tile - 1) & 1);
#pragma unroll
                for (int s8 = 0; s8 < S8; ++s8) um_st8(t_scr + 8 * s8, osum + 8 * s8);
                asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
                __syncwarp();
                if (lane == 0) um_arrive(&o_ready);
                continue;
            }
            if (NG == 2) {
                um_wait(&o_ready, tile & 1);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
                for (int s8 = 0; s8 < S8; ++s8) {
                    float o0[8];
                    um_ld8(t_scr + 8 * s8, o0);
#pragma unroll
                    for (int i = 0; i < 8; ++i) osum[8 * s8 + i] += o0[i];
                }
                asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
                __syncwarp();
                if (lane == 0) um_arrive(&o_free);
            }
            const int tg = t0 + p;
            if (t0 >= t_begin && tg < t_end) {
                float* dst = a.out[NCH == 3 ? 0 : kc0] + (size_t)b * a.n_out * C + (size_t)tg * C + CPT * ge;
#pragma unroll
                for (int q = 0; q < CPT / 4; ++q)
                    reinterpret_cast<float4*>(dst)[q] = NCH == 3 ? make_float4(osum[4 * q] / 3.0f, osum[4 * q + 1] / 3.0f, osum[4 * q + 2] / 3.0f, osum[4 * q + 3] / 3.0f)
                                                                 : make_float4(osum[4 * q], osum[4 * q + 1], osum[4 * q + 2], osum[4 * q + 3]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"((uint32_t)L::tmem_cols));
    }
}

template <int C, int U, int MT, int NCH>
int launch_stage_umma(const UmmaStageArgs& a, int B, cudaStream_t stream) {
    using L = UmLayout<C, U, MT, NCH>;
    constexpr int UM_TT = L::TT;
    static int sms_of[kMaxDevices] = {};
    const int dslot = device_slot();
    if (!sms_of[dslot]) {
        BVC_CUDA(cudaFuncSetAttribute(stage_umma_kernel<C, U, MT, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total));
        int dev = 0, n = 0;
        BVC_CUDA(cudaGetDevice(&dev));
        BVC_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
        sms_of[dslot] = n;
    }
    const int sms = sms_of[dslot];
    // time ranges per utterance: ~4 waves of CTAs, but ranges long enough that their warm-up (120 samples) stays small
    const int ctas_per_range = 3 / NCH;
    const int tiles_total = (a.n_out + UM_TT - 1) / UM_TT;
    // one CTA per SM; with one resblock per CTA the three kinds of CTAs cost 11 : 7 : 3, which the wave model below only
    // approximates (the heavy ones are launched first)
    const int ranges = pick_ranges(B * ctas_per_range, tiles_total, (120 + UM_TT - 1) / UM_TT, sms);
    stage_umma_kernel<C, U, MT, NCH><<<dim3(ranges, B, ctas_per_range), L::threads, L::total, stream>>>(a);
    BVC_CHECK_LAUNCH();
    return BVC_OK;
}

// ===========================================================================
// Anti-aliased activation (reference alias_free_torch/act.py:8-28) and the layer-by-layer stages that use it
// ===========================================================================
// Activation1d = 2x FIR up-sampling -> SnakeBeta -> 2x FIR down-sampling, per channel, with replicate padding at the
// utterance edges (resample.py:24-32, filter.py:88-96).  With the 12-tap filter f, for a sequence x[0..n):
//   u[2t]   = 2 sum_q f[2q+1] x[t+2-q],   u[2t+1] = 2 sum_q f[2q] x[t+3-q]      (q = 0..5, x index clamped to [0, n))
//   a[i]    = snakebeta(u[i])                                                    (i in [0, 2n))
//   y[t]    = sum_k g[k] a[clamp(2t + k - 5, 0, 2n - 1)]                         (k = 0..11)
// so y[t] looks 6 samples ahead: the activation is not causal, which is why these stages cannot use the time-streaming
// kernels (they carry left contexts only) and run layer by layer through HBM instead.  One kernel does the whole
// activation in shared memory: x tile (+-6 rows) -> u, a at the doubled rate -> y; HBM traffic = one read + one write.
struct AaArgs {
    const float* in_p[3];   // [B, n, C] channel-last; n_parts = 3: the mean of three tensors is activated
    int n_parts;
    int n;
    const float* ea;        // exp(alpha), 1 / (exp(beta) + 1e-9) per channel
    const float* ieb;
    float fu[12], fd[12];
    float* out;             // [B, n, C]
};

template <int C>
struct AaLayout {
    static constexpr int TT = 4096 / C;          // output rows per CTA
    static constexpr int S1 = (TT + 6 + 3) / 4;  // strips of 4 low-rate positions t = t0 - 3 + 4 s + j: a[2t], a[2t+1] each
    static constexpr int XR = 4 * S1 + 6;        // x rows t0 - 6 ..
    static constexpr int AR = 8 * S1;            // a rows 2 t0 - 6 .. (row 0 is one more than the 2 t0 - 5 the filters reach)
    static constexpr int P = C + 4;              // row pitch in floats: float4-aligned, consecutive strips hit different banks
    static constexpr size_t smem = (size_t)(XR + AR) * P * 4;
};

__device__ __forceinline__ float4 fma4(float s, const float4& x, const float4& acc) {
    return make_float4(fmaf(s, x.x, acc.x), fmaf(s, x.y, acc.y), fmaf(s, x.z, acc.z), fmaf(s, x.w, acc.w));
}

template <int C>
__global__ void __launch_bounds__(256) aa_act_kernel(AaArgs a) {
    // Work per element: 12 + 12 FIR taps and two SnakeBeta evaluations against 8 bytes of HBM traffic: on B200 the kernel is
    // bound by instruction issue, not by HBM (~37 instructions per 8 bytes is the balance point).  A thread owns 4
    // channels (128-bit shared-memory accesses) and strips of 4 consecutive rows, so the FIR windows slide through
    // registers: 10 loads per 8 up-sampled rows and 18 per 4 output rows instead of 6 and 12 per row.
    using L = AaLayout<C>;
    constexpr int TT = L::TT, S1 = L::S1, XR = L::XR, P = L::P, C4 = C / 4, RS = 256 / C4;
    extern __shared__ __align__(16) float aa_sm[];
    float* xs = aa_sm;
    float* as = aa_sm + XR * P;
    const int tid = threadIdx.x, c = 4 * (tid % C4), r0 = tid / C4;
    const int b = blockIdx.y, t0 = blockIdx.x * TT, n = a.n;
    const size_t boff = (size_t)b * n * C;
    const bool interior = t0 - 6 >= 0 && t0 - 6 + XR <= n;           // no replicate padding inside this tile
    for (int r = r0; r < XR; r += RS) {
        int t = t0 - 6 + r;
        if (!interior) t = t < 0 ? 0 : (t > n - 1 ? n - 1 : t);      // replicate padding of x
        const size_t o = boff + (size_t)t * C + c;
        float4 v = __ldg(reinterpret_cast<const float4*>(a.in_p[0] + o));
        if (a.n_parts == 3) {
            const float4 v1 = __ldg(reinterpret_cast<const float4*>(a.in_p[1] + o));
            const float4 v2 = __ldg(reinterpret_cast<const float4*>(a.in_p[2] + o));
            v.x = ((v.x + v1.x) + v2.x) / 3.0f; v.y = ((v.y + v1.y) + v2.y) / 3.0f;
            v.z = ((v.z + v1.z) + v2.z) / 3.0f; v.w = ((v.w + v1.w) + v2.w) / 3.0f;
        }
        *reinterpret_cast<float4*>(xs + r * P + c) = v;
    }
    __syncthreads();
    const float4 ea = __ldg(reinterpret_cast<const float4*>(a.ea + c)), ieb = __ldg(reinterpret_cast<const float4*>(a.ieb + c));
    for (int s = r0; s < S1; s += RS) {
        // positions t = t0 - 3 + 4 s + j (j = 0..3) need x[t - 3 .. t + 3]: rows 4 s .. 4 s + 9 of xs
        float4 x[10];
#pragma unroll
        for (int q = 0; q < 10; ++q) x[q] = *reinterpret_cast<const float4*>(xs + (4 * s + q) * P + c);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float4 ue = make_float4(0.f, 0.f, 0.f, 0.f), uo = ue;    // x[t + d] = x[j + 3 + d]
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                ue = fma4(2.0f * a.fu[2 * q + 1], x[j + 5 - q], ue); // u[2t]   = 2 sum f[2q+1] x[t + 2 - q]
                uo = fma4(2.0f * a.fu[2 * q], x[j + 6 - q], uo);     // u[2t+1] = 2 sum f[2q]   x[t + 3 - q]
            }
            *reinterpret_cast<float4*>(as + (8 * s + 2 * j) * P + c) =
                make_float4(snake_fast(ue.x, ea.x, ieb.x), snake_fast(ue.y, ea.y, ieb.y), snake_fast(ue.z, ea.z, ieb.z),
                            snake_fast(ue.w, ea.w, ieb.w));
            *reinterpret_cast<float4*>(as + (8 * s + 2 * j + 1) * P + c) =
                make_float4(snake_fast(uo.x, ea.x, ieb.x), snake_fast(uo.y, ea.y, ieb.y), snake_fast(uo.z, ea.z, ieb.z),
                            snake_fast(uo.w, ea.w, ieb.w));
        }
    }
    __syncthreads();
    const int abase = 2 * t0 - 6;                                    // doubled-rate index of as row 0
    for (int s = r0; s < TT / 4; s += RS) {
        const int t = t0 + 4 * s;
        if (t >= n) break;
        float4 av[18];                                               // a[2t - 5 .. 2t + 12]
        if (interior) {
#pragma unroll
            for (int q = 0; q < 18; ++q) av[q] = *reinterpret_cast<const float4*>(as + (8 * s + 1 + q) * P + c);
        } else {
#pragma unroll
            for (int q = 0; q < 18; ++q) {
                int i = 2 * t - 5 + q;
                i = i < 0 ? 0 : (i > 2 * n - 1 ? 2 * n - 1 : i);     // replicate padding of a
                av[q] = *reinterpret_cast<const float4*>(as + (i - abase) * P + c);
            }
        }
        float* dst = a.out + boff + (size_t)t * C + c;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (t + j >= n) break;
            float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k < 12; ++k) y = fma4(a.fd[k], av[2 * j + k], y);
            *reinterpret_cast<float4*>(dst + j * C) = y;
        }
    }
}

// One causal dilated Conv1d over channel-last rows on the tensor cores (split bf16, mma.sync): out = conv(in) + bias (+ res).
struct ConvArgs {
    const float* in;        // [B, n, C]
    const uint2* wh;        // fragment-packed weights (hi / lo), as for the streaming kernels
    const uint2* wl;
    const float* bias;
    const float* res;       // residual [B, n, C] or null; may alias out
    float* out;
    int n, d;
};

template <int C, int K>
__global__ void __launch_bounds__(256) conv_cl_kernel(ConvArgs a) {
    constexpr int TT = 128, PW = RowLayout<C>::PW, NT = C / 8;
    extern __shared__ __align__(16) uint32_t conv_sm[];
    const int lead = (K - 1) * a.d, rows = lead + TT;
    uint32_t* sh = conv_sm;
    uint32_t* sl = conv_sm + ((K - 1) * 5 + TT + 16) * PW;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int b = blockIdx.y, t0 = blockIdx.x * TT;
    const size_t boff = (size_t)b * a.n * C;
    constexpr int V = C / 2;
    for (int i = tid; i < (rows + 16) * V; i += 256) {
        const int r = i / V, cp = i - r * V;
        const int t = t0 - lead + r;
        float2 v = make_float2(0.f, 0.f);                            // causal zero padding (models.py:110,117); spare rows zero
        if (r < rows && t >= 0 && t < a.n) v = __ldg(reinterpret_cast<const float2*>(a.in + boff + (size_t)t * C) + cp);
        uint32_t hi, lo;
        split_pair(v.x, v.y, hi, lo);
        sh[r * PW + cp] = hi;
        sl[r * PW + cp] = lo;
    }
    __syncthreads();
    float2 bv[NT];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) bv[nt] = __ldg(reinterpret_cast<const float2*>(a.bias + nt * 8 + 2 * (tid & 3)));
    const uint32_t sh_a = (uint32_t)__cvta_generic_to_shared(sh), sl_a = (uint32_t)__cvta_generic_to_shared(sl);
    mma_rows<C, C, K, 1>(sh_a, sl_a, lead, a.wh, a.wl, a.d, warp * 16, [&](int row, int nt, int co, float v0, float v1) {
        const int t = t0 + row;
        if (t < a.n) {
            const size_t o = boff + (size_t)t * C + co;
            float2 r = make_float2(0.f, 0.f);
            if (a.res) r = *reinterpret_cast<const float2*>(a.res + o);
            *reinterpret_cast<float2*>(a.out + o) = make_float2(v0 + bv[nt].x + r.x, v1 + bv[nt].y + r.y);
        }
    });
}

template <int C>
int launch_aa_act(const AaArgs& a, int B, cudaStream_t stream) {
    using L = AaLayout<C>;
    static bool attr[kMaxDevices] = {};
    const int dslot = device_slot();
    if (!attr[dslot]) {
        BVC_CUDA(cudaFuncSetAttribute(aa_act_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::smem));
        attr[dslot] = true;
    }
    aa_act_kernel<C><<<dim3((a.n + L::TT - 1) / L::TT, B), 256, L::smem, stream>>>(a);
    BVC_CHECK_LAUNCH();
    return BVC_OK;
}
static int launch_aa_act_c(int C, const AaArgs& a, int B, cudaStream_t stream) {
    switch (C) {
        case 64: return launch_aa_act<64>(a, B, stream);
        case 32: return launch_aa_act<32>(a, B, stream);
        case 16: return launch_aa_act<16>(a, B, stream);
        default: return launch_aa_act<8>(a, B, stream);
    }
}

template <int C, int K>
int launch_conv_cl(const ConvArgs& a, int B, cudaStream_t stream) {
    constexpr int TT = 128;
    constexpr size_t smem = 2 * (size_t)((K - 1) * 5 + TT + 16) * RowLayout<C>::PW * 4;
    static bool attr[kMaxDevices] = {};
    const int dslot = device_slot();
    if (!attr[dslot]) {
        BVC_CUDA(cudaFuncSetAttribute(conv_cl_kernel<C, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr[dslot] = true;
    }
    conv_cl_kernel<C, K><<<dim3((a.n + TT - 1) / TT, B), 256, smem, stream>>>(a);
    BVC_CHECK_LAUNCH();
    return BVC_OK;
}
template <int C>
int launch_conv_cl_k(int K, const ConvArgs& a, int B, cudaStream_t stream) {
    switch (K) {
        case 3: return launch_conv_cl<C, 3>(a, B, stream);
        case 7: return launch_conv_cl<C, 7>(a, B, stream);
        case 11: return launch_conv_cl<C, 11>(a, B, stream);
        default: set_error("vocoder: unsupported resblock kernel size"); return BVC_ERR_INVALID;
    }
}
static int launch_conv_cl_ck(int C, int K, const ConvArgs& a, int B, cudaStream_t stream) {
    switch (C) {
        case 64: return launch_conv_cl_k<64>(K, a, B, stream);
        case 32: return launch_conv_cl_k<32>(K, a, B, stream);
        case 16: return launch_conv_cl_k<16>(K, a, B, stream);
        default: return launch_conv_cl_k<8>(K, a, B, stream);
    }
}

struct PostArgs {
    int n_parts;            // 3: mean of three partial tensors, 1: a single tensor
    const float* in_p[3];   // [B, n, C] channel-last
    int n;
    int n_out;              // min(length, n)
    int skip_act;           // 1: the input is already activated (anti-aliased activation_post ran as its own kernel)
    const float* ea;
    const float* ieb;
    const float* w;         // [ci][7]
    const float* bias;
    float inv_scale_div;
    float* wav;             // [B, n_out]
};

template <int C>
__global__ void __launch_bounds__(kThreads) post_kernel(PostArgs a) {
    constexpr int TT = 1024, K = 7;
    __shared__ float s[C][TT + K - 1];
    __shared__ float w[C * K];
    const int tid = threadIdx.x, b = blockIdx.y, t0 = blockIdx.x * TT;
    if (tid < C * K) w[tid] = a.w[tid];
    const size_t boff = (size_t)b * a.n * C;
    for (int i = tid; i < C * (TT + K - 1); i += kThreads) {
        const int p = i / C, c = i - p * C;
        const int tg = t0 - (K - 1) + p;
        float v = 0.f;
        if (tg >= 0 && tg < a.n) {
            const size_t o = boff + (size_t)tg * C + c;
            const float x = a.n_parts == 1 ? __ldg(a.in_p[0] + o)
                                           : ((__ldg(a.in_p[0] + o) + __ldg(a.in_p[1] + o)) + __ldg(a.in_p[2] + o)) / 3.0f;
            v = a.skip_act ? x : snake(x, __ldg(a.ea + c), __ldg(a.ieb + c));
        }
        s[c][p] = v;
    }
    __syncthreads();
    const float bias = __ldg(a.bias);
    for (int tt = tid; tt < TT; tt += kThreads) {
        const int tg = t0 + tt;
        if (tg >= a.n_out) break;
        float acc = bias;
#pragma unroll
        for (int c = 0; c < C; ++c)
#pragma unroll
            for (int j = 0; j < K; ++j) acc = fmaf(w[c * K + j], s[c][tt + j], acc);
        a.wav[(size_t)b * a.n_out + tg] = tanhf(acc) / a.inv_scale_div;
    }
}

// tile rows and m16 tiles per warp of the streaming tensor-core kernel, per channel count
template <int C> struct StreamTile { static constexpr int TT = 128, MT = 1; };
template <> struct StreamTile<16> { static constexpr int TT = 256, MT = 2; };
template <> struct StreamTile<8> { static constexpr int TT = 256, MT = 2; };

template <int C, int U, int K, bool POST>
int launch_stream(StageArgs a, int B, cudaStream_t stream) {
    constexpr int HALO = 12 * (K - 1);
    constexpr int TT = StreamTile<C>::TT, MT = StreamTile<C>::MT;
    a.TT = TT;
    const size_t smem = StreamLayout<C, K>::words(TT, POST) * 4;
    static int ctas_of[kMaxDevices] = {}, sms_of[kMaxDevices] = {};
    const int dslot = device_slot();
    if (!ctas_of[dslot]) {
        BVC_CUDA(cudaFuncSetAttribute(stage_stream_kernel<C, U, K, MT, POST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
        int dev = 0, n_sm = 0, n_cta = 0;
        BVC_CUDA(cudaGetDevice(&dev));
        BVC_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        BVC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n_cta, stage_stream_kernel<C, U, K, MT, POST>, kThreads, smem));
        if (n_cta < 1) { set_error("vocoder stage kernel does not fit on an SM"); return BVC_ERR_DEVICE; }
        sms_of[dslot] = n_sm;
        ctas_of[dslot] = n_cta;
    }
    const int ctas_per_sm = ctas_of[dslot], sms = sms_of[dslot];
    // time ranges per utterance: whole waves of CTAs (pick_ranges), ranges long enough that the warm-up of a range
    // (12 (K-1) samples) stays small against its length
    const int tiles_total = (a.n_out + TT - 1) / TT;
    const int ranges = pick_ranges(B, tiles_total, (HALO + TT - 1) / TT, sms * ctas_per_sm);
    dim3 grid(ranges, B);
    stage_stream_kernel<C, U, K, MT, POST><<<grid, kThreads, smem, stream>>>(a);
    BVC_CHECK_LAUNCH();
    return BVC_OK;
}

template <int C, int U, int K>
int launch_stage(const StageArgs& a_in, int B, int precision, cudaStream_t stream) {
    constexpr int HALO = 12 * (K - 1);
    StageArgs a = a_in;
    if (precision >= 1) {
        if constexpr (C == 8 && K == 3) {
            if (a.wav) return launch_stream<C, U, K, true>(a, B, stream);
        }
        return launch_stream<C, U, K, false>(a, B, stream);
    } else {
        dim3 grid((a.n_out + a.TT - 1) / a.TT, B);
        const int WP = a.TT + HALO + 4;
        const size_t smem = (size_t)3 * C * WP * sizeof(float);
        static bool attr_set[kMaxDevices] = {};
        const int dslot = device_slot();
        if (!attr_set[dslot]) {
            BVC_CUDA(cudaFuncSetAttribute(stage_fp32_kernel<C, U, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
            attr_set[dslot] = true;
        }
        stage_fp32_kernel<C, U, K><<<grid, kThreads, smem, stream>>>(a);
    }
    BVC_CHECK_LAUNCH();
    return BVC_OK;
}

template <int C, int U>
int launch_stage_k(int k, const StageArgs& a, int B, int precision, cudaStream_t stream) {
    switch (k) {
        case 3: return launch_stage<C, U, 3>(a, B, precision, stream);
        case 7: return launch_stage<C, U, 7>(a, B, precision, stream);
        case 11: return launch_stage<C, U, 11>(a, B, precision, stream);
    }
    set_error("unsupported resblock kernel size (build covers 3, 7, 11)");
    return BVC_ERR_INVALID;
}

}  // namespace

static void vocoder_dims(const VocoderWeights& w, int T, int64_t* n, int* C) {
    n[0] = T;
    C[0] = w.c0;
    for (int i = 0; i < w.n_stages; ++i) {
        n[i + 1] = (n[i] + 1) * w.rates[i];
        C[i + 1] = C[i] / 2;
    }
}

struct StageIn {            // input of a stage's transposed convolution: n_parts channel-last tensors [B, n_in, C_in] (their mean)
    const float* p[3];
    int n_parts;
    long long bstride;      // floats per sequence
    int n_in;
};
struct PostFuse {           // tail fused into the last launch of stage 3 (SnakeBeta -> conv_post -> tanh -> / scale -> [:length])
    bool enable;
    float inv_scale_div;
    float* wav;
    int length;
};

// Stage i without anti-aliasing: ConvTranspose1d + the three resblocks on B sequences of in.n_in input rows -> out[0..2]
// ([B, n_out, C] channel-last partial tensors, or one mean tensor in out[0] when *single), n_out = U (n_in + 1).
// Shared by the offline vocoder_forward and the hop-by-hop vocoder_stream_step (which feeds it short "virtual utterances").
static int run_plain_stage(const VocoderWeights& w, int i, const StageIn& in, int n_out, float* x0, float* const out[3], int B,
                           int precision, const PostFuse& pf, bool* single, bool* post_done, cudaStream_t stream) {
    static const int kTT[4] = {128, 256, 256, 512};
    static const int umma_mask = getenv("BVC_VOC_UMMA") ? atoi(getenv("BVC_VOC_UMMA")) : 0x7;
    const int Cs = w.c0 >> (i + 1);
    *single = false;
    auto run_upsample = [&]() -> int {
        UpsampleArgs up;
        up.n_parts = in.n_parts;
        for (int q = 0; q < 3; ++q) up.in_p[q] = in.p[q];
        up.in_bstride = in.bstride;
        up.n_in = in.n_in;
        up.n_out = n_out;
        up.b_up = w.b_up[i];
        up.upf_h = w.upf_h[i];
        up.upf_l = w.upf_l[i];
        up.x0 = x0;
        switch (Cs) {
            case 64: return launch_upsample<64, 8>(up, B, stream);
            case 32: return launch_upsample<32, 8>(up, B, stream);
            case 16: return launch_upsample<16, 2>(up, B, stream);
            default: return launch_upsample<8, 2>(up, B, stream);
        }
    };
        if (precision >= 1 && ((umma_mask >> i) & 1) && w.umma[i].ready) {
            // tcgen05 stage kernel: C <= 32: all three resblocks per CTA, the output is their mean; C = 64: one resblock per CTA
            const bool all_chains = Cs <= 32;
            { const int rc_up = run_upsample(); if (rc_up != BVC_OK) return rc_up; }
            int rc;
            UmmaStageArgs ua;
            ua.x0 = x0;
            ua.n_out = n_out;
            for (int cc = 0; cc < 3; ++cc) {
                const AmpBlockWeights& bw = w.blocks[i * 3 + (2 - cc)];
                for (int q = 0; q < 6; ++q) { ua.ea[cc][q] = bw.act[q].ea; ua.ieb[cc][q] = bw.act[q].inv_eb; }
                ua.out[cc] = all_chains ? out[0] : out[2 - cc];     // chain cc = resblock kernel index 2 - cc
            }
            ua.w = w.umma[i];
            ua.trace = nullptr;
            { static const int voc_dbg = getenv("BVC_VOC_DEBUG") ? atoi(getenv("BVC_VOC_DEBUG")) : 0; ua.dbg = voc_dbg; }
            static unsigned long long* trace_dev = nullptr;
            const bool tracing = getenv("BVC_VOC_TRACE") && atoi(getenv("BVC_VOC_TRACE")) == i;
            if (tracing) {
                if (!trace_dev) BVC_CUDA(cudaMalloc(&trace_dev, 80 * sizeof(unsigned long long)));
                BVC_CUDA(cudaMemsetAsync(trace_dev, 0, 80 * sizeof(unsigned long long), stream));
                ua.trace = trace_dev;
            }
            switch (Cs) {
                case 64: rc = launch_stage_umma<64, 8, 1, 1>(ua, B, stream); break;
                case 32: rc = launch_stage_umma<32, 8, 1, 3>(ua, B, stream); break;
                case 16: rc = launch_stage_umma<16, 2, 2, 3>(ua, B, stream); break;
                default: rc = launch_stage_umma<8, 2, 2, 3>(ua, B, stream); break;
            }
            if (rc != BVC_OK) return rc;
            if (tracing) {   // per job: MMA warp got the operand / issued all MMAs, epilogue saw the result / finished
                unsigned long long hbuf[80];
                BVC_CUDA(cudaStreamSynchronize(stream));
                BVC_CUDA(cudaMemcpy(hbuf, trace_dev, sizeof(hbuf), cudaMemcpyDeviceToHost));
                const unsigned long long z = hbuf[72];
                fprintf(stderr, "stage %d tile trace (us): prologue %.2f\n", i, (hbuf[73] - z) / 1e3);
                for (int ji = 0; ji < 18; ++ji)
                    if (hbuf[ji * 4])
                        fprintf(stderr, "  job %2d chain %d layer %d conv%d steps %2d: mma_start %6.2f mma_issued %6.2f epi_start %6.2f epi_done %6.2f\n", ji,
                                ua.w.jobs[ji].chain, ua.w.jobs[ji].layer, ua.w.jobs[ji].conv2 + 1, ua.w.jobs[ji].steps, (hbuf[ji * 4] - z) / 1e3,
                                (hbuf[ji * 4 + 1] - z) / 1e3, (hbuf[ji * 4 + 2] - z) / 1e3, (hbuf[ji * 4 + 3] - z) / 1e3);
            }
            if (all_chains) *single = true;
            return BVC_OK;
        }
        for (int jj = 0; jj < 3; ++jj) {
            const int j = 2 - jj;   // largest kernel first
            const AmpBlockWeights& bw = w.blocks[i * 3 + j];
            StageArgs a;
            a.n_parts = in.n_parts;
            for (int q = 0; q < 3; ++q) a.in_p[q] = in.p[q];
            a.in_bstride = in.bstride;
            a.n_in = in.n_in;
            a.n_out = n_out;
            a.w_up = w.w_up[i];
            a.b_up = w.b_up[i];
            a.upf_h = w.upf_h[i];
            a.upf_l = w.upf_l[i];
            for (int l = 0; l < 3; ++l) {
                a.w1[l] = bw.w1[l]; a.b1[l] = bw.b1[l];
                a.w2[l] = bw.w2[l]; a.b2[l] = bw.b2[l];
                a.f1h[l] = bw.f1h[l]; a.f1l[l] = bw.f1l[l];
                a.f2h[l] = bw.f2h[l]; a.f2l[l] = bw.f2l[l];
                a.dil[l] = w.dil[l];
            }
            for (int q = 0; q < 6; ++q) { a.ea[q] = bw.act[q].ea; a.ieb[q] = bw.act[q].inv_eb; }
            a.out = out[j];
            a.TT = kTT[i];
            a.wav = nullptr;
            if (i == 3 && j == 0 && precision >= 1 && !w.antialias_post && pf.enable) {
                // last launch of the vocoder: the k = 3 resblock kernel of the last stage also does the tail
                // (mean with the k = 7 and k = 11 partials -> SnakeBeta -> conv_post -> tanh), post_kernel is skipped
                a.post_p1 = out[1];
                a.post_p2 = out[2];
                a.post_ea = w.act_post.ea;
                a.post_ieb = w.act_post.inv_eb;
                a.post_w = w.w_post;
                a.post_bias = w.b_post;
                a.post_inv_scale_div = pf.inv_scale_div;
                a.wav = pf.wav;
                a.n_wav = pf.length < n_out ? pf.length : n_out;
                *post_done = a.n_wav > 0;
            }
            int rc;
            switch (i) {
                case 0: rc = launch_stage_k<64, 8>(bw.k, a, B, precision, stream); break;
                case 1: rc = launch_stage_k<32, 8>(bw.k, a, B, precision, stream); break;
                case 2: rc = launch_stage_k<16, 2>(bw.k, a, B, precision, stream); break;
                default: rc = launch_stage_k<8, 2>(bw.k, a, B, precision, stream); break;
            }
            if (rc != BVC_OK) return rc;
        }
    return BVC_OK;
}

size_t vocoder_workspace_floats(const VocoderWeights& w, int B, int T) {
    int64_t n[5];
    int C[5];
    vocoder_dims(w, T, n, C);
    size_t total = (size_t)B * (T + 6) * (w.n_mels + w.c0) + 256;
    size_t x0 = 0;
    for (int i = 0; i < w.n_stages; ++i) {
        total += 3 * ((size_t)B * C[i + 1] * n[i + 1] + 64);
        x0 = std::max(x0, (size_t)B * C[i + 1] * n[i + 1] + (size_t)512 * C[i + 1] + 64);
    }
    if (w.antialias[0] || w.antialias[1] || w.antialias[2] || w.antialias[3] || w.antialias_post) total += 2 * (x0 + 64);
    return total + x0;
}

int vocoder_forward(const VocoderWeights& w, Workspace& ws, VocoderBuffers& vb, const float* mel, int B, int T,
                    int length, float inv_scale_div, float* wav, int precision, cudaStream_t stream) {
    if (w.n_stages != 4 || w.n_kernels != 3 || w.c0 != 128) {
        set_error("vocoder: this build covers the shipped 4-stage / 3-kernel / 128-channel configuration");
        return BVC_ERR_INVALID;
    }
    vocoder_dims(w, T, vb.n, vb.C);
    vb.B = B;
    vb.T = T;
    const int X = w.n_mels;
    vb.mel_pad = ws.take((size_t)B * (T + 6) * X);
    vb.pre = ws.take((size_t)B * (T + 6) * w.c0);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 3; ++j) vb.part[i][j] = ws.take((size_t)B * vb.C[i + 1] * vb.n[i + 1]);
    {   // upsampled stage input of the tcgen05 stage kernels (+ one tile of slack: whole tiles are copied)
        size_t x0 = 0;
        for (int i = 0; i < 4; ++i) x0 = std::max(x0, (size_t)B * vb.C[i + 1] * vb.n[i + 1] + (size_t)512 * vb.C[i + 1]);
        vb.x0 = ws.take(x0);
        vb.aa_a = vb.aa_b = nullptr;
        if (w.antialias[0] || w.antialias[1] || w.antialias[2] || w.antialias[3] || w.antialias_post) {
            vb.aa_a = ws.take(x0);
            vb.aa_b = ws.take(x0);
        }
    }
    if (int rc = ws_check(ws, "vocoder_forward")) return rc;

    {
        const size_t total = (size_t)B * (T + 6) * X;
        pad_mel_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(mel, vb.mel_pad, B, T, X);
        BVC_CHECK_LAUNCH();
        LinearEpilogue ep;
        ep.bias = w.b_pre;
        ep.out = vb.pre;
        ep.ldo = w.c0;
        const int M = B * (T + 6) - 6;
        int rc = linear_forward(vb.mel_pad, X, M, w.pre, ep, precision, stream);
        if (rc != BVC_OK) return rc;
    }

    static const int kTT[4] = {128, 256, 256, 512};
    // Stages that run on the tcgen05 kernel (bit i = stage i).  Its MMAs fetch the 128 x 16 activation operand from
    // shared memory for every instruction (~64 clk measured), so with N = C_out it sustains ~40 C_out MAC/clk: twice the
    // mma.sync kernel at C = 32, on par at 16, behind at 8 where the per-job epilogue latency dominates.  Default: stage 1.
    static const int umma_mask = getenv("BVC_VOC_UMMA") ? atoi(getenv("BVC_VOC_UMMA")) : 0x7;
    static const bool fuse_post = !(getenv("BVC_VOC_FUSE_POST") && atoi(getenv("BVC_VOC_FUSE_POST")) == 0);
    bool post_done = false;
    bool single[4] = {false, false, false, false};     // stage i wrote one tensor (the mean) instead of three partials
    // ConvTranspose1d of stage i (reads the previous stage's partials or their mean) -> vb.x0
    auto run_upsample = [&](int i) -> int {
        UpsampleArgs up;
        if (i == 0) {
            up.n_parts = 1;
            up.in_p[0] = up.in_p[1] = up.in_p[2] = vb.pre;
            up.in_bstride = (long long)(T + 6) * w.c0;
        } else {
            up.n_parts = single[i - 1] ? 1 : 3;
            for (int q = 0; q < 3; ++q) up.in_p[q] = vb.part[i - 1][single[i - 1] ? 0 : q];
            up.in_bstride = (long long)vb.n[i] * vb.C[i];
        }
        up.n_in = (int)vb.n[i];
        up.n_out = (int)vb.n[i + 1];
        up.b_up = w.b_up[i];
        up.upf_h = w.upf_h[i];
        up.upf_l = w.upf_l[i];
        up.x0 = vb.x0;
        int rc;
        switch (vb.C[i + 1]) {
            case 64: rc = launch_upsample<64, 8>(up, B, stream); break;
            case 32: rc = launch_upsample<32, 8>(up, B, stream); break;
            case 16: rc = launch_upsample<16, 2>(up, B, stream); break;
            default: rc = launch_upsample<8, 2>(up, B, stream); break;
        }
        return rc;
    };
    for (int i = 0; i < 4; ++i) {
        if (w.antialias[i]) {
            // Anti-aliased stage, layer by layer: transposed conv, then per resblock 3 x (Activation1d, dilated conv,
            // Activation1d, conv + residual).  Always split-bf16 tensor-core convolutions, whatever `precision` says.
            { const int rc_up = run_upsample(i); if (rc_up != BVC_OK) return rc_up; }
            const int Cc = vb.C[i + 1], nn = (int)vb.n[i + 1];
            for (int jj = 0; jj < 3; ++jj) {
                const int j = 2 - jj;
                const AmpBlockWeights& bw = w.blocks[i * 3 + j];
                const float* cur = vb.x0;
                for (int l = 0; l < 3; ++l) {
                    AaArgs aa;
                    aa.n_parts = 1; aa.n = nn;
                    for (int q = 0; q < 12; ++q) { aa.fu[q] = w.aa_up[q]; aa.fd[q] = w.aa_down[q]; }
                    ConvArgs cv;
                    cv.n = nn;
                    // xt = c1(a1(x))
                    aa.in_p[0] = aa.in_p[1] = aa.in_p[2] = cur;
                    aa.ea = bw.act[2 * l].ea; aa.ieb = bw.act[2 * l].inv_eb; aa.out = vb.aa_a;
                    BVC_TRY(launch_aa_act_c(Cc, aa, B, stream));
                    cv.in = vb.aa_a; cv.wh = bw.f1h[l]; cv.wl = bw.f1l[l]; cv.bias = bw.b1[l]; cv.res = nullptr; cv.out = vb.aa_b;
                    cv.d = w.dil[l];
                    BVC_TRY(launch_conv_cl_ck(Cc, bw.k, cv, B, stream));
                    // x = c2(a2(xt)) + x
                    aa.in_p[0] = aa.in_p[1] = aa.in_p[2] = vb.aa_b;
                    aa.ea = bw.act[2 * l + 1].ea; aa.ieb = bw.act[2 * l + 1].inv_eb;
                    BVC_TRY(launch_aa_act_c(Cc, aa, B, stream));
                    cv.in = vb.aa_a; cv.wh = bw.f2h[l]; cv.wl = bw.f2l[l]; cv.bias = bw.b2[l]; cv.res = cur; cv.out = vb.part[i][j];
                    cv.d = 1;
                    BVC_TRY(launch_conv_cl_ck(Cc, bw.k, cv, B, stream));
                    cur = vb.part[i][j];
                }
            }
            continue;
        }
        {
            StageIn in;
            if (i == 0) {
                in.n_parts = 1;
                in.p[0] = in.p[1] = in.p[2] = vb.pre;
                in.bstride = (long long)(T + 6) * w.c0;
            } else {
                in.n_parts = single[i - 1] ? 1 : 3;
                for (int q = 0; q < 3; ++q) in.p[q] = vb.part[i - 1][single[i - 1] ? 0 : q];
                in.bstride = (long long)vb.n[i] * vb.C[i];
            }
            in.n_in = (int)vb.n[i];
            PostFuse pf;
            pf.enable = fuse_post;
            pf.inv_scale_div = inv_scale_div;
            pf.wav = wav;
            pf.length = length;
            float* outs[3] = {vb.part[i][0], vb.part[i][1], vb.part[i][2]};
            const int rc = run_plain_stage(w, i, in, (int)vb.n[i + 1], vb.x0, outs, B, precision, pf, &single[i], &post_done, stream);
            if (rc != BVC_OK) return rc;
            if (single[i]) vb.part[i][1] = vb.part[i][2] = vb.part[i][0];   // parity taps read "three partials": all the same mean
        }
    }
    if (!post_done) {
        PostArgs p;
        p.n_parts = single[3] ? 1 : 3;
        for (int q = 0; q < 3; ++q) p.in_p[q] = vb.part[3][q];
        p.n = (int)vb.n[4];
        p.n_out = length < p.n ? length : p.n;
        p.skip_act = 0;
        if (w.antialias_post) {     // Activation1d(activation_post) as its own kernel (models.py:189-190,228), then conv_post + tanh
            AaArgs aa;
            aa.n_parts = p.n_parts; aa.n = p.n;
            for (int q = 0; q < 3; ++q) aa.in_p[q] = p.in_p[q];
            for (int q = 0; q < 12; ++q) { aa.fu[q] = w.aa_up[q]; aa.fd[q] = w.aa_down[q]; }
            aa.ea = w.act_post.ea; aa.ieb = w.act_post.inv_eb; aa.out = vb.aa_a;
            BVC_TRY(launch_aa_act_c(8, aa, B, stream));
            p.n_parts = 1;
            p.in_p[0] = p.in_p[1] = p.in_p[2] = vb.aa_a;
            p.skip_act = 1;
        }
        p.ea = w.act_post.ea;
        p.ieb = w.act_post.inv_eb;
        p.w = w.w_post;
        p.bias = w.b_post;
        p.inv_scale_div = inv_scale_div;
        p.wav = wav;
        if (p.n_out > 0) {
            dim3 grid((p.n_out + 1023) / 1024, B);
            post_kernel<8><<<grid, kThreads, 0, stream>>>(p);
            BVC_CHECK_LAUNCH();
        }
    }
    return BVC_OK;
}


// =============================================================================================
// Hop-by-hop vocoder for the streaming sessions (BASELINE configs[4])
// =============================================================================================
// Per stream and stage the session keeps a ring of the stage's INPUT rows (the previous stage's output, before the
// transposed convolution): H_i history rows + the rows of the new hop.  A hop runs every stage on these short "virtual
// utterances" with the same kernels as the offline path and keeps only the rows of the new hop:
//     rows of a stage output that are exact = those whose receptive field (transposed conv: 1 input row back; resblocks:
//     12 (k - 1) = 120 rows of x0; tail: 6 more rows) lies inside the ring  ->  H = {16, 16, 61, 64} input rows.
// Work per hop and stream: 2 MMA tiles per stage (the warm-up rows ride in the 128-row tile the new rows need anyway),
// instead of re-synthesising the last 29 frames (75 tiles) as round 1's chunked decoder did.  Rings start at zero
// ("zero-state start"): the first 27 frames of a stream differ from the offline decode, which zero-pads every conv input
// at t < 0 instead (stated edge policy; from frame 27 on the outputs are those of the offline path).
namespace {

constexpr int kStreamHist[4] = {16, 16, 61, 64};

// one block per stream: ring <- shift left by n_new, append mean(src parts)[src_row0 .. src_row0 + n_new)
__global__ void __launch_bounds__(256) ring_push_kernel(float* __restrict__ ring, int R, int C, const float* __restrict__ s0,
                                                        const float* __restrict__ s1, const float* __restrict__ s2, int n_parts,
                                                        long long src_bstride, int src_row0, int n_new,
                                                        const unsigned char* __restrict__ active) {
    extern __shared__ float keep[];
    const int s = blockIdx.x, tid = threadIdx.x;
    if (active && !active[s]) return;
    float* r = ring + (size_t)s * R * C;
    const int n_keep = (R - n_new) * C;
    for (int i = tid; i < n_keep; i += 256) keep[i] = r[(size_t)n_new * C + i];
    __syncthreads();
    for (int i = tid; i < n_keep; i += 256) r[i] = keep[i];
    const size_t so = (size_t)s * src_bstride + (size_t)src_row0 * C;
    for (int i = tid; i < n_new * C; i += 256) {
        float v = s0[so + i];
        if (n_parts == 3) v = ((v + s1[so + i]) + s2[so + i]) / 3.0f;
        r[n_keep + i] = v;
    }
}

__global__ void __launch_bounds__(256) stream_tail_kernel(const float* __restrict__ wav_v, int n_v, int row0, int n,
                                                          const unsigned char* __restrict__ active, float* __restrict__ out) {
    const int s = blockIdx.x;
    const bool on = !active || active[s];
    for (int i = threadIdx.x; i < n; i += 256) out[(size_t)s * n + i] = on ? wav_v[(size_t)s * n_v + row0 + i] : 0.f;
}

}  // namespace

size_t vocoder_stream_state_floats(const VocoderWeights& w, int S) {
    size_t n = (size_t)S * 7 * w.n_mels;
    int new_rows = 1;
    for (int i = 0; i < 4; ++i) {
        n += (size_t)S * (kStreamHist[i] + new_rows) * (w.c0 >> i) + 64;
        new_rows *= w.rates[i];
    }
    return n + 64;
}

// zero the rings of the streams marked in `which` (nullptr: all)
int vocoder_stream_reset(const VocoderWeights& w, float* state, const unsigned char* which, int S, cudaStream_t stream) {
    float* p = state;
    int rc = stream_reset_rows(p, which, S, (size_t)7 * w.n_mels, stream);
    if (rc) return rc;
    p += (size_t)S * 7 * w.n_mels;
    int new_rows = 1;
    for (int i = 0; i < 4; ++i) {
        const size_t per = (size_t)(kStreamHist[i] + new_rows) * (w.c0 >> i);
        if ((rc = stream_reset_rows(p, which, S, per, stream))) return rc;
        p += (size_t)S * per + 64;
        new_rows *= w.rates[i];
    }
    return BVC_OK;
}

size_t vocoder_stream_workspace_floats(const VocoderWeights& w, int S) {
    size_t n = (size_t)S * w.c0 + 256, x0 = 0;
    int new_rows = 1;
    for (int i = 0; i < 4; ++i) {
        const int n_in = kStreamHist[i] + new_rows, C = w.c0 >> (i + 1);
        const size_t n_out = (size_t)w.rates[i] * (n_in + 1);
        n += 3 * ((size_t)S * n_out * C + 64);
        x0 = std::max(x0, (size_t)S * n_out * C + (size_t)512 * C + 64);
        new_rows *= w.rates[i];
    }
    return n + x0 + (size_t)S * 512 + 256;
}

// state: [mel ring S x 7 x X][ring 0][ring 1][ring 2][ring 3] (vocoder_stream_state_floats), zero = fresh streams.
// mel_new [S, X]: the frame decoded in this hop; active [S] (nullable): streams that advance; wav_out [S, hop samples].
int vocoder_stream_step(const VocoderWeights& w, Workspace& ws, float* state, const float* mel_new, int S,
                        const unsigned char* active, float inv_scale_div, float* wav_out, int precision, cudaStream_t stream) {
    if (w.n_stages != 4 || w.n_kernels != 3 || w.c0 != 128 || w.antialias[0] || w.antialias[1] || w.antialias[2] ||
        w.antialias[3] || w.antialias_post) {
        set_error("streaming vocoder: needs the shipped causal configuration without anti-aliased activations");
        return BVC_ERR_INVALID;
    }
    if (precision < 1) { set_error("streaming vocoder: tensor-core mode only"); return BVC_ERR_INVALID; }
    const int X = w.n_mels;
    float* mel_ring = state;
    float* ring[4];
    int R[4], n_new[5];
    {
        float* p = state + (size_t)S * 7 * X;
        n_new[0] = 1;
        for (int i = 0; i < 4; ++i) {
            R[i] = kStreamHist[i] + n_new[i];
            ring[i] = p;
            p += (size_t)S * R[i] * (w.c0 >> i) + 64;
            n_new[i + 1] = n_new[i] * w.rates[i];
        }
    }
    float* pre_new = ws.take((size_t)S * w.c0);
    float* outs[4][3];
    int n_out[4];
    size_t x0n = 0;
    for (int i = 0; i < 4; ++i) {
        const int C = w.c0 >> (i + 1);
        n_out[i] = w.rates[i] * (R[i] + 1);
        for (int j = 0; j < 3; ++j) outs[i][j] = ws.take((size_t)S * n_out[i] * C);
        x0n = std::max(x0n, (size_t)S * n_out[i] * C + (size_t)512 * C);
    }
    float* x0 = ws.take(x0n);
    float* wav_v = ws.take((size_t)S * n_out[3]);
    if (int rc = ws_check(ws, "vocoder_stream_step")) return rc;

    auto push = [&](float* rg, int Rr, int C, const float* a, const float* b, const float* c, int parts, long long bstride, int row0,
                    int nn) -> int {
        ring_push_kernel<<<S, 256, (size_t)(Rr - nn) * C * sizeof(float), stream>>>(rg, Rr, C, a, b, c, parts, bstride, row0, nn, active);
        BVC_CHECK_LAUNCH();
        return BVC_OK;
    };
    // conv_pre (models.py:209-213): 7-tap causal conv = Linear over the 7 frames of the mel ring
    int rc = push(mel_ring, 7, X, mel_new, mel_new, mel_new, 1, X, 0, 1);
    if (rc) return rc;
    {
        LinearEpilogue ep;
        ep.bias = w.b_pre;
        ep.out = pre_new;
        ep.ldo = w.c0;
        if ((rc = linear_forward(mel_ring, 7 * X, S, w.pre, ep, precision, stream))) return rc;
    }
    if ((rc = push(ring[0], R[0], w.c0, pre_new, pre_new, pre_new, 1, w.c0, 0, 1))) return rc;
    bool single_prev = true, post_done = false;
    for (int i = 0; i < 4; ++i) {
        const int C_in = w.c0 >> i, C = C_in / 2;
        StageIn in;
        in.n_parts = 1;
        in.p[0] = in.p[1] = in.p[2] = ring[i];
        in.bstride = (long long)R[i] * C_in;
        in.n_in = R[i];
        PostFuse pf;
        pf.enable = true;
        pf.inv_scale_div = inv_scale_div;
        pf.wav = wav_v;
        pf.length = n_out[3];
        bool single = false;
        if ((rc = run_plain_stage(w, i, in, n_out[i], x0, outs[i], S, precision, pf, &single, &post_done, stream))) return rc;
        const int row0 = w.rates[i] * kStreamHist[i];          // first output row of the new hop
        if (i < 3) {
            if ((rc = push(ring[i + 1], R[i + 1], C, outs[i][0], outs[i][single ? 0 : 1], outs[i][single ? 0 : 2], single ? 1 : 3,
                           (long long)n_out[i] * C, row0, n_new[i + 1])))
                return rc;
        }
        single_prev = single;
    }
    (void)single_prev;
    if (!post_done) { set_error("streaming vocoder: the fused tail did not run"); return BVC_ERR_STATE; }
    stream_tail_kernel<<<S, 256, 0, stream>>>(wav_v, n_out[3], w.rates[3] * kStreamHist[3], n_new[4], active, wav_out);
    BVC_CHECK_LAUNCH();
    return BVC_OK;
}

}  // namespace bvc
