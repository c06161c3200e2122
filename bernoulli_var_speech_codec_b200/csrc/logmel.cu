// Fused log-mel front end: reflect framing + Hann window + 1024-point real FFT +
// magnitude + sparse mel filterbank + log, one pass over the waveform.
//
// Replaces reference third_party/BigVGAN/meldataset.py:60-95 (mel_spectrogram:
// F.pad(reflect) -> torch.stft -> sqrt(re^2+im^2+1e-9) -> mel matmul ->
// log(clamp(., 1e-5))) and the x*SCALING of bvrnn_codec_model.py:49.
//
// HBM traffic: reads 4 B/sample (each sample is touched by 4 overlapping frames,
// the re-reads hit L1/L2), writes 320 B/frame.  Nothing else is materialised
// (the reference's un-fused path writes the [B,513,T] complex spectrum).
//
// Two real frames are packed into one complex 1024-point Stockham radix-4 FFT
// (z = f0 + i f1) and separated with the conjugate-symmetry identities.
#include "common.cuh"

namespace bvc {

namespace {

constexpr int kN = 1024;
constexpr int kThreads = 256;
constexpr int kFramesPerCta = 8;
constexpr int kMaxBins = 513;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// One radix-4 Stockham pass over 1024 points, thread i handles butterfly i (0..255).
__device__ __forceinline__ void radix4_pass(const float2* __restrict__ src, float2* __restrict__ dst,
                                            const float2* __restrict__ tw, int p, int i) {
    const int t = kN / 4;
    const int k = i & (p - 1);
    const int j = ((i - k) << 2) + k;
    const int q = k * (256 / p);
    float2 u0 = src[i];
    float2 u1 = cmul(src[i + t], tw[q]);
    float2 u2 = cmul(src[i + 2 * t], tw[2 * q]);
    float2 u3 = cmul(src[i + 3 * t], tw[3 * q]);
    float2 v0 = make_float2(u0.x + u2.x, u0.y + u2.y);
    float2 v1 = make_float2(u0.x - u2.x, u0.y - u2.y);
    float2 v2 = make_float2(u1.x + u3.x, u1.y + u3.y);
    float2 d = make_float2(u1.x - u3.x, u1.y - u3.y);
    float2 v3 = make_float2(d.y, -d.x);  // -i * d
    dst[j] = make_float2(v0.x + v2.x, v0.y + v2.y);
    dst[j + p] = make_float2(v1.x + v3.x, v1.y + v3.y);
    dst[j + 2 * p] = make_float2(v0.x - v2.x, v0.y - v2.y);
    dst[j + 3 * p] = make_float2(v1.x - v3.x, v1.y - v3.y);
}

__global__ void __launch_bounds__(kThreads)
logmel_kernel(const float* __restrict__ x, int L, int T, int hop, int pad_left, float scale,
              const float* __restrict__ g_window, const float2* __restrict__ g_tw,
              const int* __restrict__ mel_start, const int* __restrict__ mel_count,
              const float* __restrict__ mel_taps, int mel_width, int n_mels, int n_bins,
              float* __restrict__ mel) {
    __shared__ float2 buf0[kN];
    __shared__ float2 buf1[kN];
    __shared__ float2 tw[kN];
    __shared__ float win[kN];
    __shared__ float mag[2][kMaxBins + 3];

    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const float* xb = x + (size_t)b * L;
    for (int n = tid; n < kN; n += kThreads) {
        tw[n] = g_tw[n];
        win[n] = g_window[n];
    }
    __syncthreads();

    const int t_first = blockIdx.x * kFramesPerCta;
    for (int pair = 0; pair < kFramesPerCta / 2; ++pair) {
        const int t0 = t_first + 2 * pair;
        if (t0 >= T) break;
        const bool has1 = (t0 + 1) < T;
        // framing with reflect indexing (meldataset.py:76-80), window applied on load
        for (int n = tid; n < kN; n += kThreads) {
            int i0 = hop * t0 + n - pad_left;
            int i1 = i0 + hop;
            i0 = i0 < 0 ? -i0 : i0;
            i0 = i0 >= L ? 2 * (L - 1) - i0 : i0;
            i1 = i1 < 0 ? -i1 : i1;
            i1 = i1 >= L ? 2 * (L - 1) - i1 : i1;
            const float w = win[n];
            const float f0 = (__ldg(xb + i0) * scale) * w;
            const float f1 = has1 ? (__ldg(xb + i1) * scale) * w : 0.f;
            buf0[n] = make_float2(f0, f1);
        }
        __syncthreads();
        radix4_pass(buf0, buf1, tw, 1, tid);
        __syncthreads();
        radix4_pass(buf1, buf0, tw, 4, tid);
        __syncthreads();
        radix4_pass(buf0, buf1, tw, 16, tid);
        __syncthreads();
        radix4_pass(buf1, buf0, tw, 64, tid);
        __syncthreads();
        radix4_pass(buf0, buf1, tw, 256, tid);
        __syncthreads();
        // separate the two real spectra, magnitude (meldataset.py:86-87)
        for (int k = tid; k < n_bins; k += kThreads) {
            const float2 zk = buf1[k];
            const float2 zn = buf1[(kN - k) & (kN - 1)];
            const float r0 = 0.5f * (zk.x + zn.x), i0 = 0.5f * (zk.y - zn.y);
            const float r1 = 0.5f * (zk.y + zn.y), i1 = -0.5f * (zk.x - zn.x);
            mag[0][k] = sqrtf(r0 * r0 + i0 * i0 + 1e-9f);
            mag[1][k] = sqrtf(r1 * r1 + i1 * i1 + 1e-9f);
        }
        __syncthreads();
        // sparse mel rows + log (meldataset.py:89-90, :38-39)
        for (int o = tid; o < 2 * n_mels; o += kThreads) {
            const int f = o / n_mels, m = o - f * n_mels;
            if (f == 1 && !has1) continue;
            const int s = mel_start[m], c = mel_count[m];
            const float* taps = mel_taps + (size_t)m * mel_width;
            float acc = 0.f;
            for (int i = 0; i < c; ++i) acc = fmaf(__ldg(taps + i), mag[f][s + i], acc);
            mel[((size_t)b * T + t0 + f) * n_mels + m] = logf(fmaxf(acc, 1e-5f));
        }
        __syncthreads();
    }
}

}  // namespace

int logmel_forward(const FrontendTables& ft, const float* x, int B, int L, int hop, int pad_left, float scale,
                   float* mel, cudaStream_t stream) {
    const int T = L / hop;
    if (T <= 0) return BVC_OK;
    dim3 grid((T + kFramesPerCta - 1) / kFramesPerCta, B);
    logmel_kernel<<<grid, kThreads, 0, stream>>>(x, L, T, hop, pad_left, scale, ft.window, ft.twiddle,
                                                 ft.mel_start, ft.mel_count, ft.mel_taps, ft.mel_width,
                                                 ft.n_mels, ft.n_bins_used, mel);
    BVC_CHECK_LAUNCH();
    return BVC_OK;
}

}  // namespace bvc
