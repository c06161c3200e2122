// Causal BigVGAN-tiny vocoder.
//
// Replaces reference third_party/BigVGAN/models.py:207-238 (BigVGAN.forward), :103-121
// (AMPBlock1.forward), activations.py:107-120 (SnakeBeta) with weight-norm folded at load.
//
// Data flow (per utterance, n_0 = T mel frames, n_{i+1} = u_i (n_i + 1)):
//   conv_pre      : one GEMM over 7 contiguous channel-last mel frames -> [T, 128] channel-last
//   stage i (x4)  : one kernel per resblock kernel size k in {3,7,11}.  A CTA owns a time tile of the
//                   stage's output rate plus a left halo of 12 (k-1) samples, and keeps the whole
//                   chain  ConvTranspose1d -> 3 x (snake, dilated conv, snake, conv, residual)
//                   in shared memory: the stage input is read once (at the lower rate) and the
//                   resblock output written once.  The three resblocks of a stage write separate
//                   partial tensors; their mean (models.py:219-225) is taken by the consumer on load.
//   post          : mean -> snake -> conv_post(k=7) -> tanh -> / SCALING -> [:length]
// Causal left padding is implicit: every conv input at global time < 0 is zero (models.py:110,117).
#include "common.cuh"

namespace bvc {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float snake(float x, float ea, float inv_eb) {
    const float s = sinf(x * ea);
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

struct StageArgs {
    // input: either channel-last single tensor (stage 0) or three channel-first partials
    const float* in_cl;       // [B, rows_per_b, Cin] channel-last, or null
    long long in_cl_bstride;  // floats per utterance
    const float* in_p[3];     // [B, Cin, n_in]
    int n_in;
    int n_out;
    const float* w_up;        // [tap][ci][co]
    const float* b_up;
    const float* w1[3];
    const float* b1[3];
    const float* w2[3];
    const float* b2[3];
    const float* ea[6];
    const float* ieb[6];
    int dil[3];
    float* out;               // [B, C, n_out]
    int TT;
};

template <int C, int U, int K>
__global__ void __launch_bounds__(kThreads, 1) stage_kernel(StageArgs a) {
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

    // ---- load the stage input tile (mean of the producer's three resblock partials) ----
    if (a.in_cl) {
        const float* src = a.in_cl + (size_t)b * a.in_cl_bstride;
        for (int i = tid; i < NJ * CIN; i += kThreads) {
            const int jj = i / CIN, ci = i - jj * CIN;
            const int j = j_base + jj;
            xin[ci * NJP + jj] = (j >= 0 && j < a.n_in) ? __ldg(src + (size_t)j * CIN + ci) : 0.f;
        }
    } else {
        const size_t boff = (size_t)b * CIN * a.n_in;
        for (int i = tid; i < NJ * CIN; i += kThreads) {
            const int ci = i / NJ, jj = i - ci * NJ;
            const int j = j_base + jj;
            float v = 0.f;
            if (j >= 0 && j < a.n_in) {
                const size_t o = boff + (size_t)ci * a.n_in + j;
                v = ((__ldg(a.in_p[0] + o) + __ldg(a.in_p[1] + o)) + __ldg(a.in_p[2] + o)) / 3.0f;
            }
            xin[ci * NJP + jj] = v;
        }
    }
    __syncthreads();

    // ---- ConvTranspose1d(k = 2U, stride U): y[co, U j + r] = b + sum_ci x[ci,j] W[ci,co,r] + x[ci,j-1] W[ci,co,r+U]
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

    // ---- AMP block: 3 x (snake -> dilated conv -> snake -> conv -> residual) ----
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

    // ---- write the tile (positions [HALO, W)) ----
    float* dst = a.out + (size_t)b * C * a.n_out;
    for (int i = tid; i < C * TT; i += kThreads) {
        const int c = i / TT, tt = i - c * TT;
        const int tg = t0 + tt;
        if (tg < a.n_out) dst[(size_t)c * a.n_out + tg] = cur[c * WP + HALO + tt];
    }
}

struct PostArgs {
    const float* in_p[3];   // [B, C, n]
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
    const size_t boff = (size_t)b * C * a.n;
    for (int i = tid; i < C * (TT + K - 1); i += kThreads) {
        const int c = i / (TT + K - 1), p = i - c * (TT + K - 1);
        const int tg = t0 - (K - 1) + p;
        float v = 0.f;
        if (tg >= 0 && tg < a.n) {
            const size_t o = boff + (size_t)c * a.n + tg;
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

template <int C, int U, int K>
int launch_stage(const StageArgs& a, int B, cudaStream_t stream) {
    constexpr int HALO = 12 * (K - 1);
    const int WP = a.TT + HALO + 4;
    const size_t smem = (size_t)3 * C * WP * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        BVC_CUDA(cudaFuncSetAttribute(stage_kernel<C, U, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr_set = true;
    }
    dim3 grid((a.n_out + a.TT - 1) / a.TT, B);
    stage_kernel<C, U, K><<<grid, kThreads, smem, stream>>>(a);
    BVC_CHECK_LAUNCH();
    return BVC_OK;
}

template <int C, int U>
int launch_stage_k(int k, const StageArgs& a, int B, cudaStream_t stream) {
    switch (k) {
        case 3: return launch_stage<C, U, 3>(a, B, stream);
        case 7: return launch_stage<C, U, 7>(a, B, stream);
        case 11: return launch_stage<C, U, 11>(a, B, stream);
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
                a.in_cl = vb.pre;
                a.in_cl_bstride = (long long)(T + 6) * w.c0;
                a.in_p[0] = a.in_p[1] = a.in_p[2] = nullptr;
            } else {
                a.in_cl = nullptr;
                a.in_cl_bstride = 0;
                for (int q = 0; q < 3; ++q) a.in_p[q] = vb.part[i - 1][q];
            }
            a.n_in = (int)vb.n[i];
            a.n_out = (int)vb.n[i + 1];
            a.w_up = w.w_up[i];
            a.b_up = w.b_up[i];
            for (int l = 0; l < 3; ++l) {
                a.w1[l] = bw.w1[l]; a.b1[l] = bw.b1[l];
                a.w2[l] = bw.w2[l]; a.b2[l] = bw.b2[l];
                a.dil[l] = w.dil[l];
            }
            for (int q = 0; q < 6; ++q) { a.ea[q] = bw.act[q].ea; a.ieb[q] = bw.act[q].inv_eb; }
            a.out = vb.part[i][j];
            a.TT = kTT[i];
            int rc;
            switch (i) {
                case 0: rc = launch_stage_k<64, 8>(bw.k, a, B, stream); break;
                case 1: rc = launch_stage_k<32, 8>(bw.k, a, B, stream); break;
                case 2: rc = launch_stage_k<16, 2>(bw.k, a, B, stream); break;
                default: rc = launch_stage_k<8, 2>(bw.k, a, B, stream); break;
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
