// Shared declarations for libbvc (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "bvc.h"

namespace bvc {

void set_error(const std::string& msg);
extern thread_local int64_t* g_launch_counter;   // points at the active handle's counter

#define BVC_CUDA(expr)                                                                   \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            ::bvc::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));        \
            return BVC_ERR_DEVICE;                                                       \
        }                                                                                \
    } while (0)

#define BVC_CHECK_LAUNCH()                                                               \
    do {                                                                                 \
        if (::bvc::g_launch_counter) ++*::bvc::g_launch_counter;                         \
        cudaError_t _e = cudaGetLastError();                                             \
        if (_e != cudaSuccess) {                                                         \
            ::bvc::set_error(std::string("kernel launch failed at ") + __FILE__ + ":" +  \
                             std::to_string(__LINE__) + ": " + cudaGetErrorString(_e));  \
            return BVC_ERR_DEVICE;                                                       \
        }                                                                                \
    } while (0)

// ---------------------------------------------------------------------------
// Linear layer  C = epilogue(A . W^T)          (gemm.cu)
// A [M, K] row stride lda, W [N, K] row stride K (nn.Linear layout), fp32.
// ---------------------------------------------------------------------------
struct LinearEpilogue {
    const float* bias = nullptr;     // [N]
    const float* addend = nullptr;   // [M, ldadd], added for columns n < n_add
    int ldadd = 0;
    int n_add = 0;
    int n_act = 0;                   // ELU on columns n < n_act
    float* out = nullptr;            // [M, ldo]
    int ldo = 0;
    // optional second output: (v - nmean[n]) / nstd[n]
    const float* nmean = nullptr;
    const float* nstd = nullptr;
    float* out2 = nullptr;
    int ldo2 = 0;
};

struct LinearWeights {
    const float* w = nullptr;        // fp32 [N, K]
    const uint32_t* w_hi = nullptr;  // split-bf16 copies packed for the tensor-core path (may be null)
    const uint32_t* w_lo = nullptr;
    int N = 0, K = 0;
};

int linear_forward(const float* A, int lda, int M, const LinearWeights& W, const LinearEpilogue& ep,
                   int precision, cudaStream_t stream);

// tcgen05 path for the hoisted layers (gemm_umma.cu): operands are shared-memory images (recurrent.cuh)
struct WImg;
int to_image(const float* x, int M, int K_in, const float* mean, const float* sd, unsigned char* img, int kchunks,
             cudaStream_t stream);
int linear_umma(const unsigned char* a_img, int M, const WImg& w, const float* bias, int act, float* out_f, int ldo,
                unsigned char* out_img, cudaStream_t stream);

// ---------------------------------------------------------------------------
// log-mel front end (logmel.cu)
// ---------------------------------------------------------------------------
struct FrontendTables {
    float* window = nullptr;      // [1024]
    float2* twiddle = nullptr;    // [1024] exp(-2 pi i q / 1024)
    int* mel_start = nullptr;     // [n_mels]
    int* mel_count = nullptr;     // [n_mels]
    float* mel_taps = nullptr;    // [n_mels, mel_width]
    int mel_width = 0;
    int n_mels = 0;
    int n_bins_used = 0;          // highest non-zero FFT bin + 1
};
int logmel_forward(const FrontendTables& ft, const float* x, int B, int L, int hop, int pad_left, float scale,
                   float* mel, cudaStream_t stream);

// ---------------------------------------------------------------------------
// BVRNN coder (bvrnn.cu)
// ---------------------------------------------------------------------------
// Weight matrix [N][K] (nn.Linear layout) as split-bf16 shared-memory images for the persistent recurrent
// kernel: [n_tile][k_chunk] chunks of bn rows x 64 columns, hi part then lo part, SWIZZLE_128B (recurrent.cuh)
struct WImg {
    const unsigned char* img = nullptr;
    int N = 0, K = 0, bn = 0;
};
namespace rec { struct Program; }
struct RecurrentWeights {
    bool ready = false;
    // *_q: GRU rows gate-interleaved in groups of 12 = [r(4) z(4) n(4)] (KIND_GRU epilogue: after the cluster's
    // reduce-scatter a CTA owns 12 of the 48 columns of a tile, i.e. all three gates of 4 hidden units)
    // x1f: phi_x.0 . diag(1/std) . dec.6 -- dec.6, the mel normalisation and phi_x.0 are all affine, so the
    // reconstructed mel never has to exist on the recurrence's critical path (bvrnn.py:202-204)
    WImg e0h, d0h, whh_q, e2, e4, pz0, pz2, pz4, d0z, ihz_q, d2, d4, d6, x1f, px2, px4, ihx_q;
    float *b_e0 = nullptr, *b_hh_q = nullptr, *b_d0 = nullptr, *b_ih_q = nullptr, *b_x1f = nullptr, *b_d6p = nullptr;
    // hoisted layers (gemm_umma.cu), weight images with bn = 256; g_zcat = [dec.0[:, :H]; W_ih[:, H:] gate-interleaved]
    WImg g_px0, g_px2, g_px4, g_e0x, g_pz0, g_pz2, g_pz4, g_zcat;
    float* b_zcat_q = nullptr;
    WImg g_pr0, g_pr2, g_pr4;        // prior head (bvrnn.py:68-73), evaluated over all frames from all_h
    float* b_pr4p = nullptr;         // prior.4 bias padded to the 256-wide tile
    unsigned* sync_words = nullptr;  // device: abort flag + one barrier counter per m-tile
    rec::Program* prog_dev = nullptr;
    // Launches are asynchronous: a ring of pinned staging slots (program image + the kernel's abort flag read back after
    // the launch + timing events).  A slot is reused only after its `done` event; the abort flag of a finished launch is
    // looked at lazily (next entry point on the handle, bvc_check).
    static constexpr int PROG_SLOTS = 8;
    struct ProgSlot {
        rec::Program* host = nullptr;    // pinned
        int* flag_host = nullptr;        // pinned: sync_words[0] of the launch, copied back in stream order
        cudaEvent_t done = nullptr, ev_begin = nullptr, ev_end = nullptr;
        bool used = false, checked = true;
        int kind = 0;                    // 0 encode, 1 decode
        long long call_id = 0;
    } slots[PROG_SLOTS];
    int next_slot = 0, cur_slot = -1;
    long long call_seq = 0;
    int deferred_abort = 0;              // sticky until reported: abort code of an earlier launch
    // parity taps of the last encode call (views into the workspace): dec.0_h . h [B,H], W_hh h + b [B,3H] and
    // W_ih_z phi_z + b [B,3H] of the LAST frame, the latter two with gate-interleaved columns
    const float *tap_dh = nullptr, *tap_gh = nullptr, *tap_giz = nullptr;
    int tap_B = 0;
};
// host-side services of the launch ring (bvrnn.cu)
int rec_poll_aborts(RecurrentWeights& rw, bool wait);                 // BVC_ERR_DEVICE (+ message) if an earlier launch aborted
float rec_launch_ms(RecurrentWeights& rw, int kind, int age);         // device time of the age-th latest call of `kind` (-1: any); waits for it; -1 if none
void rec_free_slots(RecurrentWeights& rw);

struct BvrnnWeights {
    int X = 0, H = 0, Z = 0, var_bit = 0;
    float *mean = nullptr, *std = nullptr;
    LinearWeights px0, px2, px4;      // phi_x
    LinearWeights pz0, pz2, pz4;      // phi_z
    LinearWeights e0x;                // enc.0[:, :H]      (hoisted over all frames, no bias)
    LinearWeights e2, e4;
    LinearWeights d2, d4, d6;
    LinearWeights hcat_enc;           // [enc.0[:,H:]; dec.0[:,H:]; W_hh]   N = 5H   (input h)
    LinearWeights hcat_dec;           // [dec.0[:,H:]; W_hh]                N = 4H   (input h)
    LinearWeights zcat;               // [dec.0[:,:H]; W_ih[:,H:]]          N = 4H   (input phi_z)
    LinearWeights ihx;                // W_ih[:, :H]                        N = 3H   (input phi_x_gen)
    LinearWeights pr0, pr2, pr4;      // prior head
    float *b_pr0 = nullptr, *b_pr2 = nullptr, *b_pr4 = nullptr;
    float *b_px0, *b_px2, *b_px4, *b_pz0, *b_pz2, *b_pz4, *b_e2, *b_e4, *b_d2, *b_d4, *b_d6;
    float *b_hcat_enc, *b_hcat_dec, *b_zcat;
    RecurrentWeights rw;
};

struct Workspace {
    float* base = nullptr;
    size_t bytes = 0;
    size_t used = 0;
    bool overflow = false;     // sticky: a take() went past the allocation (sizing formula out of step with the takes)
    // Bounds-checked bump allocation.  On overflow the pointer returned still lies inside the allocation (its start),
    // so nothing is written out of bounds before the entry point sees `overflow` (checked by ws_check() after the
    // takes of every path) and returns BVC_ERR_NOMEM.
    float* take(size_t n_floats) {
        size_t off = (used + 63) & ~size_t(63);
        if ((off + n_floats) * sizeof(float) > bytes) {
            overflow = true;
            return base;
        }
        used = off + n_floats;
        return base + off;
    }
};
int ws_check(const Workspace& ws, const char* where);

// cudaFuncSetAttribute / occupancy / SM count are per device: one flag (and cached values) per device ordinal.
constexpr int kMaxDevices = 64;
inline int device_slot() {
    int d = 0;
    cudaGetDevice(&d);
    return (d >= 0 && d < kMaxDevices) ? d : 0;
}

size_t bvrnn_workspace_floats(const BvrnnWeights& w, int B, int T);
int unpack_codes(const unsigned long long* packed, const float* bits, float bits_scalar, int var_bit, size_t n_frames,
                 int Z, float* codes, cudaStream_t s);
// packed wire format (wire.cu)
size_t bitstream_bytes(int T, int Z, int per_frame);
int pack_bitstream(Workspace& ws, const unsigned long long* words, const float* bits, float bits_scalar, int var_bit, int Z,
                   int B, int T, unsigned char* out, size_t stride, cudaStream_t s);
int unpack_bitstream(Workspace& ws, const unsigned char* in, size_t stride, int B, int T, int Z, unsigned long long* words,
                     float* bits_out, cudaStream_t s);
int words_to_image(const unsigned long long* words, const float* bits, float bits_scalar, int var_bit, int M,
                   unsigned char* img, cudaStream_t s);
// uniforms [B,T,Z] (nullable): sampled bits z = round(u - 0.5 + p) instead of round(p) (bvrnn.py:123-126)
int bvrnn_encode(BvrnnWeights& w, Workspace& ws, const float* mel, const float* bits, float bits_scalar,
                 const float* h0, int B, int T, float* codes, unsigned long long* packed, float* logits,
                 float* all_h, float* h_final, float* mel_hat, int precision, cudaStream_t stream,
                 const float* uniforms = nullptr);
// prior(h_t) for all frames: all_h [B,T,H] -> probabilities [B,T,Z] (bvrnn.py:68-73,115-120)
int bvrnn_prior(BvrnnWeights& w, Workspace& ws, const float* all_h, int B, int T, float* prior, int precision,
                cudaStream_t stream);
// codes [B,T,Z] floats, or (codes == nullptr) packed words [B,T] + budgets (bits [B,T] or bits_scalar)
int bvrnn_decode(BvrnnWeights& w, Workspace& ws, const float* codes, const float* h0, int B, int T,
                 float* mel, float* h_final, int precision, cudaStream_t stream,
                 const unsigned long long* packed = nullptr, const float* bits = nullptr, float bits_scalar = 0.f);

// ---------------------------------------------------------------------------
// causal BigVGAN-tiny vocoder (vocoder.cu)
// ---------------------------------------------------------------------------
struct SnakeParams {
    float* ea = nullptr;      // exp(alpha)            [C]
    float* inv_eb = nullptr;  // 1 / (exp(beta) + 1e-9) [C]
};
struct AmpBlockWeights {
    int k = 0;
    // per layer l in 0..2: conv1 (dilated) and conv2
    float* w1[3];   // fp32 [ci][tap][co]                       (FFMA path)
    float* b1[3];
    float* w2[3];
    float* b2[3];
    // split-bf16 copies in mma.m16n8k16 B-fragment order: [k16 chunk][n8 tile][lane] -> uint2 (b0, b1),
    // GEMM-K index = tap * C + ci                              (tensor-core path)
    uint2 *f1h[3], *f1l[3], *f2h[3], *f2l[3];
    SnakeParams act[6];
};
// Stage on the 5th-gen tensor cores (vocoder.cu, stage_umma_kernel): one tile = 18 convolution "jobs" (3 resblocks x
// 3 layers x 2 convs) whose weights are streamed in MMA issue order as chunks of <= 16 KiB
struct UmmaJob {
    short chain, layer, conv2, K, d, steps;
    int off;                                  // byte offset of the job's weight chunks in the stream (16 KB each, the last one shorter)
};
struct UmmaStageWeights {
    bool ready = false;
    const unsigned char* wstream = nullptr;   // all chunks of one tile, concatenated
    UmmaJob jobs[18];
    const float* b1[3][3];                    // [chain][layer] bias of the dilated conv
    const float* bsum[3][3];                  // [chain][layer] sum of the second convs' biases of layers 0..layer
};
struct VocoderWeights {
    int n_mels = 0, c0 = 0, n_stages = 0, n_kernels = 0;
    int rates[4], dil[3], rks[3];
    LinearWeights pre;        // conv_pre as a Linear over 7 contiguous mel frames: [c0, 7*n_mels], index j*n_mels+ci
    float* b_pre = nullptr;
    float* w_up[4];           // ConvTranspose weights repacked [tap][ci][co]
    float* b_up[4];
    uint2 *upf_h[4], *upf_l[4];   // per output phase r: 2-tap fragment-packed weights (taps r+U, r), phases concatenated
    AmpBlockWeights blocks[12];
    UmmaStageWeights umma[4];
    bool antialias[4] = {false, false, false, false};   // stage i: Activation1d around every SnakeBeta of its resblocks
    bool antialias_post = false;                        // ... and around activation_post
    float aa_up[12], aa_down[12];                       // the checkpoint's 12-tap Kaiser-sinc filters (up / down sampling)
    SnakeParams act_post;
    float* w_post = nullptr;  // [ci][tap]
    float* b_post = nullptr;  // [1]
};
struct VocoderBuffers {       // views into the workspace, valid after the last vocoder_forward
    float* mel_pad = nullptr; // [B, T+6, n_mels]
    float* pre = nullptr;     // [B, T+6, c0] channel-last (rows t >= T of each utterance are scratch)
    float* aa_a = nullptr;    // [B, n, C] scratch of the layer-by-layer anti-aliased stages
    float* aa_b = nullptr;
    float* x0 = nullptr;      // [B, n, C] output of a stage's transposed convolution (tcgen05 stage kernels; reused by every stage)
    float* part[4][3];        // per stage, per resblock: [B, n, C] channel-last
    int64_t n[5];             // n[0] = T, n[i+1] = length after stage i
    int C[5];
    int B = 0, T = 0;
};
size_t vocoder_workspace_floats(const VocoderWeights& w, int B, int T);
int vocoder_forward(const VocoderWeights& w, Workspace& ws, VocoderBuffers& vb, const float* mel, int B, int T,
                    int length, float inv_scale_div, float* wav, int precision, cudaStream_t stream);

// hop-by-hop vocoder of the streaming sessions (vocoder.cu): per-stream stage-input rings in `state`
size_t vocoder_stream_state_floats(const VocoderWeights& w, int S);
size_t vocoder_stream_workspace_floats(const VocoderWeights& w, int S);
int vocoder_stream_reset(const VocoderWeights& w, float* state, const unsigned char* which, int S, cudaStream_t stream);
int vocoder_stream_step(const VocoderWeights& w, Workspace& ws, float* state, const float* mel_new, int S,
                        const unsigned char* active, float inv_scale_div, float* wav_out, int precision, cudaStream_t stream);

// streaming session kernels (stream.cu)
int stream_push(float* fifo, int* n_samples, const float* x_new, const unsigned char* active, int S, int hop, int n_fft,
                int pad_left, float* win, unsigned char* valid, cudaStream_t s);
int stream_commit(float* dst, const float* src, const unsigned char* mask, int S, int n, cudaStream_t s);
int stream_mask_words(unsigned long long* out, const unsigned long long* in, const unsigned char* mask, int S, cudaStream_t s);
int stream_reset_rows(float* p, const unsigned char* which, int S, size_t n, cudaStream_t s);
int stream_reset_ints(int* p, const unsigned char* which, int S, cudaStream_t s);

// ---------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ float elu1(float v) { return v > 0.f ? v : expm1f(v); }
// tensor-core epilogues: ex2.approx based; absolute error <= ~1.2e-7 for v <= 0 (expm1f costs ~0.7 us per layer on the
// recurrence's critical path: 16 serial evaluations per epilogue thread)
// (ex2.approx.ftz directly: __expf adds a denormal-range rescue of ~10 predicated instructions per value, and exp(v) - 1 is
// -1 to fp32 precision long before that range)
__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
#ifdef BVC_EPI_OLD
__device__ __forceinline__ float elu_fast(float v) { return v > 0.f ? v : __expf(v) - 1.f; }
#else
__device__ __forceinline__ float elu_fast(float v) { return v > 0.f ? v : ex2_ftz(v * 1.4426950408889634f) - 1.f; }
#endif
__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }
__device__ __forceinline__ float sigmoid_fast(float v) { return __fdividef(1.f, 1.f + __expf(-v)); }
// tanh(v) = 1 - 2 / (1 + e^{2v}); saturates correctly: e^{2v} -> inf gives 1, -> 0 gives -1
__device__ __forceinline__ float tanh_fast(float v) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * v)); }

}  // namespace bvc
