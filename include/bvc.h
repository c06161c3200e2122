/*
 * bvc.h -- C ABI of libbvc.so, the B200 (sm_100a) implementation of the batched
 * BVRNN speech-codec encode -> decode hot path.
 *
 * The reference (BenjSta/bernoulli-var-speech-codec) has no FFI layer: its
 * boundary is the Python nn.Module API (bvrnn_codec_model.py:19-76).  This
 * header is the boundary a binding for that API talks to; each entry point
 * names the reference function it replaces.  The Python facade
 * (bernoulli_var_speech_codec_b200/codec.py) binds it with ctypes, see
 * INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success, a negative bvc_status otherwise;
 *     bvc_last_error() returns a thread-local message for the last failure.
 *     No exceptions cross the ABI.
 *   - "dev" pointers are CUDA device pointers on the handle's device, "host"
 *     pointers are ordinary (ideally page-locked) host memory.  The caller
 *     owns every input/output buffer; the handle owns packed weights and
 *     workspace.
 *   - `stream` is a cudaStream_t passed as void*.  Device entry points only
 *     enqueue work on it: no host synchronisation, no blocking copy (ABI 3; the
 *     persistent recurrent kernel's program image goes through a ring of pinned
 *     staging slots and its abort flag is read back in stream order).  A
 *     protocol failure inside that kernel (bounded waits, see
 *     recurrent_cluster.cu) is therefore reported LAZILY: by the next
 *     bvc_encode* / bvc_decode_mel on the handle that finds the failed launch
 *     finished, by bvc_check(), or by the *_host entry points (which
 *     synchronise anyway).  Exceptions are marked "synchronises".
 *   - one handle per device; calls on one handle must be serialised by the
 *     caller.
 *   - there is no CPU path: bvc_create fails with BVC_ERR_DEVICE when the
 *     device is not compute capability 10.x.
 *   - layouts are row-major, innermost dimension last, float32 unless noted.
 */
#ifndef BVC_H_
#define BVC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BVC_ABI_VERSION 3

typedef enum bvc_status {
    BVC_OK = 0,
    BVC_ERR_INVALID = -1,   /* bad argument / shape                                  */
    BVC_ERR_SCHEMA = -2,    /* checkpoint tensor missing, unexpected or mis-shaped    */
    BVC_ERR_DEVICE = -3,    /* no sm_100 device, or a CUDA runtime error              */
    BVC_ERR_STATE = -4,     /* weights / front-end tables not loaded yet              */
    BVC_ERR_NOMEM = -5
} bvc_status;

typedef struct bvc_handle bvc_handle;

/* Model hyper-parameters: the keys the reference reads from its TOML config
 * (configs/config_varBitRate.toml:21-29,35-37,39-56; bvrnn_codec_model.py:27-36). */
typedef struct bvc_config {
    int32_t device;            /* CUDA device ordinal                                  */
    int32_t x_dim;             /* num_mels, 80                                         */
    int32_t h_dim;             /* 1024                                                 */
    int32_t z_dim;             /* 64                                                   */
    int32_t var_bit;           /* 1: bits >= bits-per-frame are masked to 0.5          */
    int32_t n_fft;             /* winsize, 1024 (must be 1024 in this build)           */
    int32_t hop;               /* hopsize, 256                                         */
    int32_t pad_left;          /* mel_pad_left, 256                                    */
    int32_t voc_initial_channel;       /* 128                                          */
    int32_t voc_num_stages;            /* 4                                            */
    int32_t voc_up_rates[4];           /* 8,8,2,2                                      */
    int32_t voc_up_kernels[4];         /* 16,16,4,4 (must equal 2*rate)                */
    int32_t voc_num_kernels;           /* 3                                            */
    int32_t voc_res_kernels[3];        /* 3,7,11                                       */
    int32_t voc_res_dilations[3];      /* 1,3,5 (same for every resblock)              */
    /* ABI 2: anti-aliased activations (third_party/BigVGAN/models.py:66-88,172-190; alias_free_torch/act.py:8-28).
     * A stage with voc_antialias[i] = 1 expects the Activation1d checkpoint schema for its resblocks
     * (activations.N.act.alpha / .act.beta / .upsample.filter / .downsample.lowpass.filter) and runs layer by layer
     * (FIR up -> SnakeBeta -> FIR down fused in one shared-memory kernel); both shipped configs leave these 0. */
    int32_t voc_antialias[4];          /* vocoder_config.layers_antialias               */
    int32_t voc_antialias_post;        /* vocoder_config.antialias_post                 */
} bvc_config;

/* One named checkpoint tensor, float32, contiguous, in HOST memory. */
typedef struct bvc_tensor {
    const char* name;          /* state-dict key, e.g. "phi_x.0.weight"               */
    const float* data;         /* host pointer                                         */
    int32_t ndim;
    int64_t shape[4];
} bvc_tensor;

int bvc_abi_version(void);
const char* bvc_last_error(void);

/* bvrnn_codec_model.py:20-36 (module construction). */
int bvc_create(bvc_handle** out, const bvc_config* cfg);
int bvc_destroy(bvc_handle* h);

/* bvrnn_codec_model.py:38,41 -- strict load of the 'vrnn' state dict (39 tensors,
 * schema SURVEY.md 3.1).  Unknown, missing or mis-shaped tensors fail with
 * BVC_ERR_SCHEMA.  prior.* feeds bvc_encode_ex's prior output; log_sigma is accepted and unused. */
int bvc_load_bvrnn(bvc_handle* h, const bvc_tensor* tensors, int32_t n);

/* bvrnn_codec_model.py:39,42 -- strict load of the 'generator' state dict
 * (weight_g / weight_v / bias triplets, snake alpha / beta).  weight-norm is
 * folded here once (third_party/BigVGAN/models.py:47-62,140,164,200). */
int bvc_load_vocoder(bvc_handle* h, const bvc_tensor* tensors, int32_t n);

/* Front-end constant tables (meldataset.py:67-70): analysis window [n_fft] and
 * the mel filterbank [n_mels, n_fft/2+1] (dense, host memory). */
int bvc_set_frontend(bvc_handle* h, const float* window, const float* mel_basis);

/* third_party/BigVGAN/meldataset.py:60-95 (mel_spectrogram, incl. the x*SCALING
 * of bvrnn_codec_model.py:49 through `scale`).
 * x_dev [B, L] -> mel_dev [B, T, x_dim] with T = L / hop.  Requires L > n_fft - pad_left - hop. */
int bvc_logmel(bvc_handle* h, const float* x_dev, int32_t B, int32_t L, float scale,
               float* mel_dev, void* stream);

/* bvrnn.py:163-209 (BVRNN.encode).
 * mel_dev [B,T,x_dim]; bits_dev [B,T] bits per frame (float, like the reference's
 * varBitrate) or NULL -> bits_scalar for every frame; h0_dev [B,h_dim] or NULL -> zeros.
 * Outputs (each nullable except codes_dev):
 *   codes_dev  [B,T,z_dim]  {0,1} active bits, 0.5 masked bits (bvrnn.py:191-196)
 *   packed_dev [B,T] uint64 bit i of word = code i (masked bits 0)
 *   logits_dev [B,T,z_dim]  pre-sigmoid encoder output (parity tap)
 *   all_h_dev  [B,T,h_dim]  state entering frame t (bvrnn.py:205)
 *   h_final_dev[B,h_dim]    state after the last frame */
int bvc_encode(bvc_handle* h, const float* mel_dev, const float* bits_dev, float bits_scalar,
               const float* h0_dev, int32_t B, int32_t T,
               float* codes_dev, uint64_t* packed_dev, float* logits_dev,
               float* all_h_dev, float* h_final_dev, void* stream);

/* bvc_encode that also returns the decoder's mel (fused forward, new).  The reference's encoder runs the decoder inside
 * its loop (analysis by synthesis, bvrnn.py:198-206): the dec_t it forms is exactly what BVRNN.decode (bvrnn.py:222-227)
 * recomputes from the same codes and h0, so forward() = decode(encode(x)) (bvrnn_codec_model.py:73-76) needs only one
 * recurrence.  mel_hat_dev [B,T,x_dim] or NULL (then identical to bvc_encode). */
int bvc_encode_mel(bvc_handle* h, const float* mel_dev, const float* bits_dev, float bits_scalar,
                   const float* h0_dev, int32_t B, int32_t T,
                   float* codes_dev, uint64_t* packed_dev, float* logits_dev,
                   float* all_h_dev, float* h_final_dev, float* mel_hat_dev, void* stream);

/* ABI 3: bvc_encode_mel plus the two remaining pieces of the reference's BVRNN.forward (bvrnn.py:86-160):
 *   uniforms_dev [B,T,z_dim] or NULL   sampled bits z = round(u - 0.5 + p) with caller-supplied uniforms u (bvrnn.py:123-126;
 *                                      the reference draws u with torch.rand_like) instead of the greedy round(p)
 *   prior_dev    [B,T,z_dim] or NULL   Bernoulli probabilities of the `prior` head, prior(h_t) (bvrnn.py:68-73,115-120), for
 *                                      entropy / KL estimates (bvrnn.py:148-158 against sigmoid(logits_dev))
 * Encoder and decoder state stay in lock-step as in BVRNN.encode (the reference's forward with p_use_gen = 1). */
int bvc_encode_ex(bvc_handle* h, const float* mel_dev, const float* bits_dev, float bits_scalar,
                  const float* h0_dev, const float* uniforms_dev, int32_t B, int32_t T,
                  float* codes_dev, uint64_t* packed_dev, float* logits_dev,
                  float* all_h_dev, float* h_final_dev, float* mel_hat_dev, float* prior_dev, void* stream);

/* Wire format (new; the reference only has the float layout): one uint64 per frame, bit i = code i, masked bits 0.
 * bvc_encode fills packed_dev; this expands it back to the reference's float codes ({0,1}, 0.5 for bits >= budget,
 * bvrnn.py:191-196) so that bvc_decode_mel can consume it.  bits_dev [B,T] or NULL -> bits_scalar. */
int bvc_unpack_codes(bvc_handle* h, const uint64_t* packed_dev, const float* bits_dev, float bits_scalar,
                     int32_t B, int32_t T, float* codes_dev, void* stream);

/* ABI 3: BVRNN.decode (bvrnn.py:211-229) straight from the wire words: packed_dev [B,T] uint64 + the bit budgets (bits_dev
 * [B,T] or NULL -> bits_scalar; ignored by a fixed-rate model).  The {0, 1, 0.5} code vector is formed directly as the
 * operand of phi_z.0; the float code tensor of the reference (bvrnn.py:191-196, 256 bytes per frame) is never materialised.
 * Same result as bvc_unpack_codes + bvc_decode_mel. */
int bvc_decode_packed(bvc_handle* h, const uint64_t* packed_dev, const float* bits_dev, float bits_scalar,
                      const float* h0_dev, int32_t B, int32_t T, float* mel_dev, float* h_final_dev, void* stream);

/* ABI 3: n-bit-per-frame bit-stream (new).  Per utterance, little-endian:
 *    0 "BVC1"   4 uint16 z_dim   6 uint8 mode (0: n bits in every frame, 1: per-frame budgets)   7 uint8 n (mode 0)
 *    8 uint32 T   12 uint32 payload bits   16 [mode 1: T budget bytes, padded to 4]   then the payload: the n_t ACTIVE
 *    bits of frame 0, 1, ... back to back, LSB first, in 32-bit words.
 * 3 kbps = 35 bits per frame -> 4.4 bytes per frame on the wire (the word format: 8, the reference's floats: 256).
 * bvc_bitstream_bytes: size of one utterance's stream (upper bound for mode 1), a valid `stride`.
 * bvc_pack_bitstream:   packed_dev [B,T] words (+ budgets as for bvc_encode) -> out_dev [B][stride] bytes.
 * bvc_unpack_bitstream: in_dev [B][stride] -> packed_dev [B,T] words and (nullable) bits_out_dev [B,T] budgets as floats,
 *                       ready for bvc_decode_packed / bvc_unpack_codes.  T must be the header's T. */
size_t bvc_bitstream_bytes(const bvc_handle* h, int32_t T, int32_t per_frame_budgets);
int bvc_pack_bitstream(bvc_handle* h, const uint64_t* packed_dev, const float* bits_dev, float bits_scalar,
                       int32_t B, int32_t T, uint8_t* out_dev, size_t stride, void* stream);
int bvc_unpack_bitstream(bvc_handle* h, const uint8_t* in_dev, size_t stride, int32_t B, int32_t T,
                         uint64_t* packed_dev, float* bits_out_dev, void* stream);

/* bvrnn.py:211-229 (BVRNN.decode).  codes_dev [B,T,z_dim] arbitrary floats. */
int bvc_decode_mel(bvc_handle* h, const float* codes_dev, const float* h0_dev,
                   int32_t B, int32_t T, float* mel_dev, float* h_final_dev, void* stream);

/* third_party/BigVGAN/models.py:207-238 (BigVGAN.forward) followed by the
 * division by SCALING of bvrnn_codec_model.py:71 (`inv_scale_div`: the output is
 * tanh(.) / inv_scale_div; pass 1.0f for the bare vocoder).
 * mel_dev [B,T,x_dim] (channel-last) -> wav_dev [B, n], n = min(length, 256*T+294). */
int bvc_vocode(bvc_handle* h, const float* mel_dev, int32_t B, int32_t T, int32_t length,
               float inv_scale_div, float* wav_dev, void* stream);
int64_t bvc_vocoder_out_len(const bvc_handle* h, int32_t T);

/* bvrnn_codec_model.py:44-62 / :64-71 with HOST buffers: host->device copy of the
 * input, the whole stage chain, device->host copy of the result, and a stream
 * synchronise before returning.  x_host [B,L] -> codes_host [B, L/hop, z_dim];
 * codes_host [B,T,z_dim] -> wav_host [B, length]. */
int bvc_encode_host(bvc_handle* h, const float* x_host, int32_t B, int32_t L, float scale,
                    float bits_scalar, float* codes_host);
int bvc_decode_host(bvc_handle* h, const float* codes_host, int32_t B, int32_t T, int32_t length,
                    float inv_scale_div, float* wav_host);

/* ABI 3: streaming sessions (new; SURVEY.md 8f-1, BASELINE configs[4]).  n_streams independent streams advance hop by hop:
 * one step = `hop` (256) new samples per stream = 11.61 ms of audio.  The reference has only the ingredients (h in / out at
 * bvrnn.py:163,211; causal convolutions at third_party/BigVGAN/models.py:19-20,110,117; 768 samples of look-ahead in
 * meldataset.py:76-80 = the 34.8 ms algorithmic latency of README.md:19).  The session owns the per-stream state:
 * encoder: window FIFO (1024 samples), sample count, BVRNN state; decoder: BVRNN state, the vocoder's stage-input rings
 * (15 k floats per stream).  Every step only enqueues work on `stream`.
 *   active_dev / valid_dev  [n_streams] bytes or NULL (= all): streams that take part in this hop; idle streams keep their
 *                           state untouched (ragged stream activity).
 *   encode_step   x_new_dev [n_streams, hop] -> packed_out_dev [n_streams] wire words of the frame each stream completed in
 *                 this hop, valid_out_dev [n_streams] = 1 where there is one (frame t needs 256 t + 768 samples: the first
 *                 frame comes with the third hop and uses the reference's reflect padding of its own first samples;
 *                 the right reflect padding of the offline call's last two frames is an end-of-utterance notion with no
 *                 streaming counterpart: a stream simply ends).  bits_dev [n_streams] per-stream budgets or NULL -> scalar.
 *   decode_step   packed_dev [n_streams] + valid_dev -> wav_out_dev [n_streams, hop] (zeros for idle streams).
 *                 Edge policy: zero-state start -- the vocoder rings start at zero, so the first 27 frames of a stream
 *                 (its 6 719-sample receptive field) differ from the offline decode, which zero-pads every convolution
 *                 input at t < 0; from frame 27 on the samples are those of the offline bvc_decode_mel + bvc_vocode.
 *   reset         zeroes the state of the streams marked in which_dev [n_streams] (NULL = all): a new stream starts. */
typedef struct bvc_stream bvc_stream;
int bvc_stream_create(bvc_handle* h, int32_t n_streams, bvc_stream** out);
int bvc_stream_destroy(bvc_stream* st);
int bvc_stream_reset(bvc_stream* st, const uint8_t* which_dev, void* stream);
int bvc_stream_encode_step(bvc_stream* st, const float* x_new_dev, const uint8_t* active_dev, const float* bits_dev,
                           float bits_scalar, float scale, uint64_t* packed_out_dev, uint8_t* valid_out_dev, void* stream);
int bvc_stream_decode_step(bvc_stream* st, const uint64_t* packed_dev, const uint8_t* valid_dev, const float* bits_dev,
                           float bits_scalar, float inv_scale_div, float* wav_out_dev, void* stream);

/* Page-locked host memory for the host-buffer entry points above (cudaHostAlloc / cudaFreeHost).  Pinning a
 * 226 MB result buffer costs ~37 ms, so a binding should pool these blocks instead of allocating per call
 * (codec.py does: blocks return to its pool when the last tensor viewing them dies). */
int bvc_host_alloc(void** out, size_t bytes);
int bvc_host_free(void* p);

/* Bytes of device workspace the handle holds for a (B, T) job (grown on demand). */
size_t bvc_workspace_bytes(const bvc_handle* h);
/* Kernels launched by this library since the handle was created (bench evidence). */
int64_t bvc_kernel_launches(const bvc_handle* h);
/* Device time in ms (CUDA events on the launching stream) of the persistent recurrent kernel launched by the last
 * bvc_encode / bvc_decode_mel (or their _host variants): the dominant kernel of the path, used for bench.py's roofline. */
float bvc_last_recurrent_ms(const bvc_handle* h);
/* ABI 3: the same per kind of call (0 = bvc_encode*, 1 = bvc_decode_mel, -1 = either) and age (0 = the latest such call,
 * 1 = the one before, ... as long as it is still in the ring of 8 launches), summed over the launches of that call;
 * -1 if there is no such call.  Both functions wait (host) for that launch to finish; neither is needed for correctness. */
float bvc_recurrent_ms(bvc_handle* h, int32_t kind, int32_t age);
/* ABI 3: waits for every recurrent-kernel launch enqueued through this handle and returns BVC_ERR_DEVICE (with
 * bvc_last_error) if one of them raised its abort flag, BVC_OK otherwise.  Synchronises (only those launches). */
int bvc_check(bvc_handle* h);
/* Arithmetic mode of the GEMM/conv inner products: 1 = split-bf16 tensor core (default; the benchmarked path),
 * 0 = fp32 FFMA kernels (slow cross-check path with reference-grade rounding). */
int bvc_set_precision(bvc_handle* h, int32_t mode);

/* Debug/parity taps: copy an internal device buffer of the LAST call to host.
 * Names: "voc_pre" [B,T+6,128] channel-last, "voc_stage{0..3}_{0..2}" [B,n,C] channel-last partial sums.
 * Synchronises the device. */
int bvc_debug_read(bvc_handle* h, const char* name, float* dst_host, size_t n_floats);

#ifdef __cplusplus
}
#endif
#endif /* BVC_H_ */
