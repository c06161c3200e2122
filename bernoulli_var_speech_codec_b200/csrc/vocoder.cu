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
#include <cuda_bf16.h>

#include "common.cuh"

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
    static constexpr size_t words(int TT) {
        return (size_t)(TT + 16) * PF + 2 * (size_t)(CTX1 + TT + 16) * PW + 2 * (size_t)(CTX2 + TT + 16) * PW +
               2 * (size_t)CTXSUM * PW + 2 * (size_t)3 * CTX2 * PW;
    }
};

template <int C, int U, int K, int MT>
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

        // ---- write the tile, channel-last, coalesced (warm-up tiles are not written) ----
        if (t0 >= t_begin) {
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

struct PostArgs {
    const float* in_p[3];   // [B, n, C] channel-last
    int n;
    int n_out;              // min(length, n)
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
            const float x = ((__ldg(a.in_p[0] + o) + __ldg(a.in_p[1] + o)) + __ldg(a.in_p[2] + o)) / 3.0f;
            v = snake(x, __ldg(a.ea + c), __ldg(a.ieb + c));
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
template <> struct StreamTile<8> { static constexpr int TT = 512, MT = 2; };

template <int C, int U, int K>
int launch_stage(const StageArgs& a_in, int B, int precision, cudaStream_t stream) {
    constexpr int HALO = 12 * (K - 1);
    StageArgs a = a_in;
    if (precision >= 1) {
        constexpr int TT = StreamTile<C>::TT, MT = StreamTile<C>::MT;
        a.TT = TT;
        const size_t smem = StreamLayout<C, K>::words(TT) * 4;
        static int ctas_per_sm = 0, sms = 0;
        if (!ctas_per_sm) {
            BVC_CUDA(cudaFuncSetAttribute(stage_stream_kernel<C, U, K, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
            int dev = 0;
            BVC_CUDA(cudaGetDevice(&dev));
            BVC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            BVC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, stage_stream_kernel<C, U, K, MT>, kThreads, smem));
            if (ctas_per_sm < 1) { set_error("vocoder stage kernel does not fit on an SM"); ctas_per_sm = 0; return BVC_ERR_DEVICE; }
        }
        // time ranges per utterance: enough CTAs for ~4 waves (tail effect), but ranges long enough that the
        // warm-up of a range (12 (K-1) samples) stays small against its length
        const int tiles_total = (a.n_out + TT - 1) / TT;
        const int want = (4 * sms * ctas_per_sm + B - 1) / B;
        const int max_ranges = tiles_total / (4 * ((HALO + TT - 1) / TT)) > 0 ? tiles_total / (4 * ((HALO + TT - 1) / TT)) : 1;
        int ranges = want < max_ranges ? want : max_ranges;
        if (ranges < 1) ranges = 1;
        const int tiles_per = (tiles_total + ranges - 1) / ranges;
        ranges = (tiles_total + tiles_per - 1) / tiles_per;
        dim3 grid(ranges, B);
        stage_stream_kernel<C, U, K, MT><<<grid, kThreads, smem, stream>>>(a);
    } else {
        dim3 grid((a.n_out + a.TT - 1) / a.TT, B);
        const int WP = a.TT + HALO + 4;
        const size_t smem = (size_t)3 * C * WP * sizeof(float);
        static bool attr_set = false;
        if (!attr_set) {
            BVC_CUDA(cudaFuncSetAttribute(stage_fp32_kernel<C, U, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
            attr_set = true;
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

size_t vocoder_workspace_floats(const VocoderWeights& w, int B, int T) {
    int64_t n[5];
    int C[5];
    vocoder_dims(w, T, n, C);
    size_t total = (size_t)B * (T + 6) * (w.n_mels + w.c0) + 256;
    for (int i = 0; i < w.n_stages; ++i) total += 3 * ((size_t)B * C[i + 1] * n[i + 1] + 64);
    return total;
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
    for (int i = 0; i < 4; ++i) {
        for (int jj = 0; jj < 3; ++jj) {
            const int j = 2 - jj;   // largest kernel first
            const AmpBlockWeights& bw = w.blocks[i * 3 + j];
            StageArgs a;
            if (i == 0) {
                a.n_parts = 1;
                a.in_p[0] = a.in_p[1] = a.in_p[2] = vb.pre;
                a.in_bstride = (long long)(T + 6) * w.c0;
            } else {
                a.n_parts = 3;
                for (int q = 0; q < 3; ++q) a.in_p[q] = vb.part[i - 1][q];
                a.in_bstride = (long long)vb.n[i] * vb.C[i];
            }
            a.n_in = (int)vb.n[i];
            a.n_out = (int)vb.n[i + 1];
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
            a.out = vb.part[i][j];
            a.TT = kTT[i];
            int rc;
            switch (i) {
                case 0: rc = launch_stage_k<64, 8>(bw.k, a, B, precision, stream); break;
                case 1: rc = launch_stage_k<32, 8>(bw.k, a, B, precision, stream); break;
                case 2: rc = launch_stage_k<16, 2>(bw.k, a, B, precision, stream); break;
                default: rc = launch_stage_k<8, 2>(bw.k, a, B, precision, stream); break;
            }
            if (rc != BVC_OK) return rc;
        }
    }
    {
        PostArgs p;
        for (int q = 0; q < 3; ++q) p.in_p[q] = vb.part[3][q];
        p.n = (int)vb.n[4];
        p.n_out = length < p.n ? length : p.n;
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

}  // namespace bvc
