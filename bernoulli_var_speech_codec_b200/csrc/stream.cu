// Streaming sessions (BASELINE configs[4]): S concurrent streams advance hop by hop (256 new samples =
// 11.61 ms per stream and step).  Not in the reference, which only has the ingredients: the BVRNN takes and returns its
// hidden state (bvrnn.py:163,211), every vocoder convolution is causal (third_party/BigVGAN/models.py:19-20,110,117), and
// the log-mel frame t looks 768 samples ahead (meldataset.py:76-80) -- the codec's 34.8 ms algorithmic latency
// (README.md:19).
//
// Per-stream state (device, owned by the session):
//   encoder   the last 1024 samples (frame window FIFO), samples received, the BVRNN encoder state h
//   decoder   the BVRNN decoder state h, the vocoder's stage-input rings (vocoder.cu, vocoder_stream_step)
// Streams are independent and may be active or idle in any hop (`active` masks): every kernel runs over all S streams
// with uniform shapes, and the state of an idle stream is simply not committed.
#include "common.cuh"

namespace bvc {

namespace {

// one block per stream: push the hop into the FIFO, emit the analysis window of the frame it completes (if any).
// Frame t covers samples [256 t - 256, 256 t + 768): complete once 256 t + 768 samples have arrived, i.e. with the third
// hop for t = 0, whose left part is the reference's reflect padding of the first samples (meldataset.py:80).
__global__ void __launch_bounds__(256) stream_push_kernel(float* __restrict__ fifo, int* __restrict__ n_samples,
                                                          const float* __restrict__ x_new, const unsigned char* __restrict__ active,
                                                          int hop, int n_fft, int pad_left, float* __restrict__ win,
                                                          unsigned char* __restrict__ valid) {
    const int s = blockIdx.x, tid = threadIdx.x;
    const bool on = !active || active[s];
    float* f = fifo + (size_t)s * n_fft;
    __shared__ float buf[1024];
    if (!on) {
        if (tid == 0) valid[s] = 0;
        return;
    }
    for (int i = tid; i < n_fft; i += 256) buf[i] = i + hop < n_fft ? f[i + hop] : x_new[(size_t)s * hop + (i + hop - n_fft)];
    __syncthreads();
    for (int i = tid; i < n_fft; i += 256) f[i] = buf[i];
    const int n = n_samples[s] + hop;
    __syncthreads();
    if (tid == 0) n_samples[s] = n;
    const int lookahead = n_fft - pad_left;                 // 768
    const bool ok = n >= lookahead;
    if (tid == 0) valid[s] = ok ? 1 : 0;
    if (!ok) return;
    const bool first = n == lookahead;                      // frame 0: x[0 .. 768) sits at FIFO positions [pad_left, n_fft)
    for (int i = tid; i < n_fft; i += 256) {
        float v = buf[i];
        if (first && i < pad_left) v = buf[pad_left + (pad_left - i)];     // reflect without edge repeat: x[256 - i]
        win[(size_t)s * n_fft + i] = v;
    }
}

// dst[s] <- src[s] for the streams with mask[s] != 0 (state commit of the active streams)
__global__ void commit_rows_kernel(float* __restrict__ dst, const float* __restrict__ src, const unsigned char* __restrict__ mask,
                                   int S, int n) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)S * n) return;
    const int s = (int)(idx / n);
    if (!mask || mask[s]) dst[idx] = src[idx];
}
__global__ void mask_words_kernel(unsigned long long* __restrict__ out, const unsigned long long* __restrict__ in,
                                  const unsigned char* __restrict__ mask, int S) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < S) out[s] = (!mask || mask[s]) ? in[s] : 0ull;
}
__global__ void reset_rows_kernel(float* __restrict__ p, const unsigned char* __restrict__ which, int S, size_t n) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)S * n) return;
    if (!which || which[idx / n]) p[idx] = 0.f;
}
__global__ void reset_ints_kernel(int* __restrict__ p, const unsigned char* __restrict__ which, int S) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < S && (!which || which[s])) p[s] = 0;
}

}  // namespace

int stream_push(float* fifo, int* n_samples, const float* x_new, const unsigned char* active, int S, int hop, int n_fft,
                int pad_left, float* win, unsigned char* valid, cudaStream_t s) {
    if (n_fft != 1024) { set_error("streaming: n_fft must be 1024"); return BVC_ERR_INVALID; }
    stream_push_kernel<<<S, 256, 0, s>>>(fifo, n_samples, x_new, active, hop, n_fft, pad_left, win, valid);
    BVC_CHECK_LAUNCH();
    return BVC_OK;
}
int stream_commit(float* dst, const float* src, const unsigned char* mask, int S, int n, cudaStream_t s) {
    const size_t total = (size_t)S * n;
    commit_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(dst, src, mask, S, n);
    BVC_CHECK_LAUNCH();
    return BVC_OK;
}
int stream_mask_words(unsigned long long* out, const unsigned long long* in, const unsigned char* mask, int S, cudaStream_t s) {
    mask_words_kernel<<<(S + 255) / 256, 256, 0, s>>>(out, in, mask, S);
    BVC_CHECK_LAUNCH();
    return BVC_OK;
}
int stream_reset_rows(float* p, const unsigned char* which, int S, size_t n, cudaStream_t s) {
    const size_t total = (size_t)S * n;
    reset_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(p, which, S, n);
    BVC_CHECK_LAUNCH();
    return BVC_OK;
}
int stream_reset_ints(int* p, const unsigned char* which, int S, cudaStream_t s) {
    reset_ints_kernel<<<(S + 255) / 256, 256, 0, s>>>(p, which, S);
    BVC_CHECK_LAUNCH();
    return BVC_OK;
}

}  // namespace bvc
