// Packed wire format of the codes (not in the reference, which moves 256 bytes of floats per frame for
// <= 8 bytes of information, bvrnn.py:191-196) and the packed entry into the decoder.
//
//   word      uint64 per frame, bit i = code i, masked bits 0 (what the encode kernel's bottleneck epilogue emits)
//   stream    per utterance: 16-byte header + optional per-frame budgets + the ACTIVE bits of every frame back to back
//             (n_t bits for frame t, LSB first, little-endian 32-bit words):
//               0  "BVC1"        4  uint16 z_dim      6  uint8 mode (0: n bits in every frame, 1: per-frame budgets)
//               7  uint8 n       8  uint32 T          12 uint32 payload bits
//               16 [mode 1: T bytes of budgets, padded to a multiple of 4]     then the payload words
//   decode    bvrnn_decode takes the words + budgets directly: the {0, 1, 0.5} code vector is formed as the bf16 operand
//             image of phi_z.0 (exact in bf16), the float code tensor never exists
#include <string.h>

#include "common.cuh"
#include "recurrent.cuh"

namespace bvc {

namespace {

__device__ __forceinline__ int budget_of(const float* bits, float bits_scalar, int var_bit, int Z, size_t f) {
    if (!var_bit) return Z;
    const float b = bits ? bits[f] : bits_scalar;        // bit i is active iff b > i  (bvrnn.py:180-182)
    const int n = (int)ceilf(b);
    return n < 0 ? 0 : (n > Z ? Z : n);
}

// one block per utterance: off[b][t] = sum of the budgets of frames < t (exclusive scan), off[b][T] = payload bits
__global__ void __launch_bounds__(256) budget_offsets_kernel(const float* __restrict__ bits, float bits_scalar, int var_bit, int Z,
                                                             int T, unsigned* __restrict__ off) {
    __shared__ unsigned part[256];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int per = (T + 255) / 256, t0 = tid * per, t1 = min(T, t0 + per);
    unsigned s = 0;
    for (int t = t0; t < t1; ++t) s += (unsigned)budget_of(bits, bits_scalar, var_bit, Z, (size_t)b * T + t);
    part[tid] = s;
    __syncthreads();
    if (tid == 0) {
        unsigned acc = 0;
        for (int i = 0; i < 256; ++i) { const unsigned v = part[i]; part[i] = acc; acc += v; }
        off[(size_t)b * (T + 1) + T] = acc;
    }
    __syncthreads();
    unsigned acc = part[tid];
    for (int t = t0; t < t1; ++t) {
        off[(size_t)b * (T + 1) + t] = acc;
        acc += (unsigned)budget_of(bits, bits_scalar, var_bit, Z, (size_t)b * T + t);
    }
}

__device__ __forceinline__ size_t payload_offset(int mode, int T) { return 16 + (mode == 1 ? (size_t)((T + 3) / 4) * 4 : 0); }

// one thread per frame: header (thread of frame 0), budget byte, and the frame's n bits OR-ed into the payload words
__global__ void __launch_bounds__(256) pack_stream_kernel(const unsigned long long* __restrict__ words, const float* __restrict__ bits,
                                                          float bits_scalar, int var_bit, int Z, int B, int T, int mode,
                                                          const unsigned* __restrict__ off, unsigned char* __restrict__ out,
                                                          size_t stride) {
    const size_t f = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= (size_t)B * T) return;
    const int b = (int)(f / T), t = (int)(f - (size_t)b * T);
    unsigned char* u = out + (size_t)b * stride;
    const int n = budget_of(bits, bits_scalar, var_bit, Z, f);
    const unsigned o = off[(size_t)b * (T + 1) + t];
    if (t == 0) {
        u[0] = 'B'; u[1] = 'V'; u[2] = 'C'; u[3] = '1';
        u[4] = (unsigned char)(Z & 0xff); u[5] = (unsigned char)(Z >> 8);
        u[6] = (unsigned char)mode; u[7] = (unsigned char)(mode == 0 ? n : 0);
        const unsigned pb = off[(size_t)b * (T + 1) + T];
        for (int i = 0; i < 4; ++i) { u[8 + i] = (unsigned char)((unsigned)T >> (8 * i)); u[12 + i] = (unsigned char)(pb >> (8 * i)); }
    }
    if (mode == 1) u[16 + t] = (unsigned char)n;
    if (n == 0) return;
    unsigned* pw = reinterpret_cast<unsigned*>(u + payload_offset(mode, T));
    const unsigned long long w = n >= 64 ? words[f] : (words[f] & ((1ull << n) - 1ull));
    const unsigned sh = o & 31u;
    unsigned wi = o >> 5;
    atomicOr(pw + wi, (unsigned)(w << sh));
    int done = 32 - (int)sh;
    while (done < n) {
        ++wi;
        atomicOr(pw + wi, (unsigned)(w >> done));
        done += 32;
    }
}

// the same exclusive scan over the budget bytes of a mode-1 stream (mode 0: off = n t needs no table)
__global__ void __launch_bounds__(256) stream_offsets_kernel(const unsigned char* __restrict__ in, size_t stride, int T,
                                                             unsigned* __restrict__ off) {
    __shared__ unsigned part[256];
    const int b = blockIdx.x, tid = threadIdx.x;
    const unsigned char* u = in + (size_t)b * stride;
    if (u[6] != 1) return;
    const int per = (T + 255) / 256, t0 = tid * per, t1 = min(T, t0 + per);
    unsigned s = 0;
    for (int t = t0; t < t1; ++t) s += u[16 + t];
    part[tid] = s;
    __syncthreads();
    if (tid == 0) {
        unsigned acc = 0;
        for (int i = 0; i < 256; ++i) { const unsigned v = part[i]; part[i] = acc; acc += v; }
    }
    __syncthreads();
    unsigned acc = part[tid];
    for (int t = t0; t < t1; ++t) {
        off[(size_t)b * (T + 1) + t] = acc;
        acc += u[16 + t];
    }
}

// one thread per frame: the frame's n bits back into a word (+ the budget as float, for bvc_unpack_codes / bvc_decode_packed)
__global__ void __launch_bounds__(256) unpack_stream_kernel(const unsigned char* __restrict__ in, size_t stride, int B, int T, int Z,
                                                            const unsigned* __restrict__ off,
                                                            unsigned long long* __restrict__ words, float* __restrict__ bits_out) {
    const size_t f = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= (size_t)B * T) return;
    const int b = (int)(f / T), t = (int)(f - (size_t)b * T);
    const unsigned char* u = in + (size_t)b * stride;
    const int mode = u[6];
    int n;
    unsigned o;
    if (mode == 1) {
        n = u[16 + t];
        o = off[(size_t)b * (T + 1) + t];
    } else {
        n = u[7];
        o = (unsigned)n * (unsigned)t;
    }
    if (n > Z) n = Z;
    unsigned long long w = 0;
    if (n > 0) {
        const unsigned* pw = reinterpret_cast<const unsigned*>(u + payload_offset(mode, T));
        const unsigned sh = o & 31u;
        unsigned wi = o >> 5;
        w = (unsigned long long)(pw[wi] >> sh);
        int got = 32 - (int)sh;
        while (got < n) {
            ++wi;
            w |= (unsigned long long)pw[wi] << got;
            got += 32;
        }
        if (n < 64) w &= (1ull << n) - 1ull;
    }
    words[f] = w;
    if (bits_out) bits_out[f] = (float)n;
}

// words -> the activation image of phi_z.0 (K = 64 = one chunk): hi part {0, 1, 0.5} as bf16, lo part zero
__global__ void __launch_bounds__(256) words_to_image_kernel(const unsigned long long* __restrict__ words, const float* __restrict__ bits,
                                                             float bits_scalar, int var_bit, int M, unsigned char* __restrict__ img) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;    // one thread per (row, 8 columns)
    const size_t rows = ((size_t)M + 127) / 128 * 128;
    if (idx >= rows * 8) return;
    const size_t m = idx >> 3;
    const int g = (int)(idx & 7);
    uint4 hi = make_uint4(0, 0, 0, 0);
    if (m < (size_t)M) {
        const unsigned long long w = words[m];
        const float budget = bits ? bits[m] : bits_scalar;
        unsigned short h[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = 8 * g + i;
            const bool active = !var_bit || (budget > (float)c);
            h[i] = active ? (((w >> c) & 1ull) ? 0x3f80 : 0x0000) : 0x3f00;      // bf16 of 1.0 / 0.0 / 0.5
        }
        hi = make_uint4(h[0] | ((unsigned)h[1] << 16), h[2] | ((unsigned)h[3] << 16), h[4] | ((unsigned)h[5] << 16),
                        h[6] | ((unsigned)h[7] << 16));
    }
    const int row = (int)(m % 128);
    unsigned char* base = img + (m / 128) * rec::ACT_CHUNK_BYTES + row * 128 + ((g ^ (row & 7)) << 4);
    *reinterpret_cast<uint4*>(base) = hi;
    *reinterpret_cast<uint4*>(base + rec::ACT_PART_BYTES) = make_uint4(0, 0, 0, 0);
}

}  // namespace

size_t bitstream_bytes(int T, int Z, int per_frame) {
    const size_t payload_words = ((size_t)T * (size_t)Z + 31) / 32;
    return 16 + (per_frame ? (size_t)((T + 3) / 4) * 4 : 0) + 4 * payload_words;
}

int pack_bitstream(Workspace& ws, const unsigned long long* words, const float* bits, float bits_scalar, int var_bit, int Z,
                   int B, int T, unsigned char* out, size_t stride, cudaStream_t s) {
    if (Z > 64) { set_error("pack_bitstream: z_dim > 64 does not fit a 64-bit word"); return BVC_ERR_INVALID; }
    const int mode = (var_bit && bits) ? 1 : 0;
    if (stride < bitstream_bytes(T, Z, mode) || (stride & 3)) {
        set_error("pack_bitstream: stride must be a multiple of 4 and >= bvc_bitstream_bytes()");
        return BVC_ERR_INVALID;
    }
    unsigned* off = reinterpret_cast<unsigned*>(ws.take((size_t)B * (T + 1)));
    if (int rc = ws_check(ws, "pack_bitstream")) return rc;
    BVC_CUDA(cudaMemsetAsync(out, 0, (size_t)B * stride, s));
    budget_offsets_kernel<<<B, 256, 0, s>>>(bits, bits_scalar, var_bit, Z, T, off);
    BVC_CHECK_LAUNCH();
    const size_t n = (size_t)B * T;
    pack_stream_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(words, bits, bits_scalar, var_bit, Z, B, T, mode, off, out, stride);
    BVC_CHECK_LAUNCH();
    return BVC_OK;
}

int unpack_bitstream(Workspace& ws, const unsigned char* in, size_t stride, int B, int T, int Z, unsigned long long* words,
                     float* bits_out, cudaStream_t s) {
    if (Z > 64 || (stride & 3) || stride < 16) { set_error("unpack_bitstream: bad z_dim / stride"); return BVC_ERR_INVALID; }
    const size_t n = (size_t)B * T;
    if (n == 0) return BVC_OK;
    unsigned* off = reinterpret_cast<unsigned*>(ws.take((size_t)B * (T + 1)));
    if (int rc = ws_check(ws, "unpack_bitstream")) return rc;
    stream_offsets_kernel<<<B, 256, 0, s>>>(in, stride, T, off);
    BVC_CHECK_LAUNCH();
    unpack_stream_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(in, stride, B, T, Z, off, words, bits_out);
    BVC_CHECK_LAUNCH();
    return BVC_OK;
}

int words_to_image(const unsigned long long* words, const float* bits, float bits_scalar, int var_bit, int M,
                   unsigned char* img, cudaStream_t s) {
    const size_t total = ((size_t)M + 127) / 128 * 128 * 8;
    words_to_image_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(words, bits, bits_scalar, var_bit, M, img);
    BVC_CHECK_LAUNCH();
    return BVC_OK;
}

}  // namespace bvc
