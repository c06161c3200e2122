// Linear layers of the BVRNN coder:  C = epilogue(A . W^T)  with A [M,K] activations and
// W [N,K] in nn.Linear layout (reference bvrnn.py:44-83: phi_x, phi_z, enc, dec, and the
// GRU's weight_ih / weight_hh).  The epilogue fuses bias, an optional per-row addend
// (the hoisted half of a concatenated-input layer), ELU on a column prefix and an
// optional normalised second output ((v - mean) / std, bvrnn.py:204).
//
// precision 0: fp32 FFMA tiles (exact-grade arithmetic, the parity baseline)
// precision 1: split-bf16 tensor-core tiles, a = a_hi + a_lo, w = w_hi + w_lo,
//              a.w ~= a_hi.w_hi + a_hi.w_lo + a_lo.w_hi with fp32 accumulation
//              (~16 mantissa bits per operand)
#include <cuda_bf16.h>

#include "common.cuh"

namespace bvc {

namespace {

constexpr int kBK = 16;

__device__ __forceinline__ void epilogue_store(const LinearEpilogue& ep, int m, int n, float v) {
    if (ep.bias) v += __ldg(ep.bias + n);
    if (ep.addend && n < ep.n_add) v += __ldg(ep.addend + (size_t)m * ep.ldadd + n);
    if (n < ep.n_act) v = elu1(v);
    ep.out[(size_t)m * ep.ldo + n] = v;
    if (ep.out2) ep.out2[(size_t)m * ep.ldo2 + n] = (v - __ldg(ep.nmean + n)) / __ldg(ep.nstd + n);
}

// ---------------------------------------------------------------------------
// fp32 FFMA path
// ---------------------------------------------------------------------------
template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(256)
linear_fp32_kernel(const float* __restrict__ A, int lda, int M, const float* __restrict__ W, int N, int K,
                   LinearEpilogue ep) {
    static_assert((BM / TM) * (BN / TN) == 256, "256 threads");
    __shared__ __align__(16) float As[2][kBK][BM + 4];
    __shared__ __align__(16) float Bs[2][kBK][BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);

    // global->smem staging: one float4 (4 consecutive k) per (row, kq)
    constexpr int A_F4 = BM * 4, B_F4 = BN * 4;
    const int a_row = tid >> 2, a_kq = tid & 3;
    float4 a_stage[(A_F4 + 255) / 256], b_stage[(B_F4 + 255) / 256];

    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int i = 0; i < (A_F4 + 255) / 256; ++i) {
            const int r = a_row + i * 64;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < BM && m0 + r < M) v = *reinterpret_cast<const float4*>(A + (size_t)(m0 + r) * lda + k0 + a_kq * 4);
            a_stage[i] = v;
        }
#pragma unroll
        for (int i = 0; i < (B_F4 + 255) / 256; ++i) {
            const int r = a_row + i * 64;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < BN && n0 + r < N) v = __ldg(reinterpret_cast<const float4*>(W + (size_t)(n0 + r) * K + k0 + a_kq * 4));
            b_stage[i] = v;
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int i = 0; i < (A_F4 + 255) / 256; ++i) {
            const int r = a_row + i * 64;
            if (r < BM) {
                As[buf][a_kq * 4 + 0][r] = a_stage[i].x;
                As[buf][a_kq * 4 + 1][r] = a_stage[i].y;
                As[buf][a_kq * 4 + 2][r] = a_stage[i].z;
                As[buf][a_kq * 4 + 3][r] = a_stage[i].w;
            }
        }
#pragma unroll
        for (int i = 0; i < (B_F4 + 255) / 256; ++i) {
            const int r = a_row + i * 64;
            if (r < BN) {
                Bs[buf][a_kq * 4 + 0][r] = b_stage[i].x;
                Bs[buf][a_kq * 4 + 1][r] = b_stage[i].y;
                Bs[buf][a_kq * 4 + 2][r] = b_stage[i].z;
                Bs[buf][a_kq * 4 + 3][r] = b_stage[i].w;
            }
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    const int nk = K / kBK;
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load_tiles((kt + 1) * kBK);
#pragma unroll
        for (int k = 0; k < kBK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[buf][k][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[buf][k][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) store_tiles(buf ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + ty * TM + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + tx * TN + j;
            if (n < N) epilogue_store(ep, m, n, acc[i][j]);
        }
    }
}

// ---------------------------------------------------------------------------
// split-bf16 tensor-core path (mma.sync m16n8k16, fp32 accumulate)
// CTA tile BM x BN, 8 warps, BK = 32.  A is split into hi/lo on the fly while staging
// to shared memory; W hi/lo are pre-split at load time (bf16 pairs along K).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
    const __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0));
    const __nv_bfloat16 l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
    hi = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    lo = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
}

constexpr int kTK = 32;          // K per smem stage
constexpr int kTPitch = 20;      // uint32 (bf16 pair) pitch per row: 16 pairs + 4 pad -> conflict-free fragments

template <int BM, int BN, int WM, int WN>   // warp grid WM x WN = 8 warps; warp tile (BM/WM) x (BN/WN)
__global__ void __launch_bounds__(256)
linear_bf16x3_kernel(const float* __restrict__ A, int lda, int M, const uint32_t* __restrict__ Whi,
                     const uint32_t* __restrict__ Wlo, int N, int K, LinearEpilogue ep) {
    static_assert(WM * WN == 8, "8 warps");
    constexpr int WTM = BM / WM, WTN = BN / WN;
    constexpr int MI = WTM / 16, NI = WTN / 8;
    extern __shared__ __align__(16) uint32_t smem_u32[];
    typedef uint32_t (*TileA)[BM][kTPitch];
    typedef uint32_t (*TileB)[BN][kTPitch];
    TileA Ah = reinterpret_cast<TileA>(smem_u32);
    TileA Al = reinterpret_cast<TileA>(smem_u32 + 2 * BM * kTPitch);
    TileB Bh = reinterpret_cast<TileB>(smem_u32 + 4 * BM * kTPitch);
    TileB Bl = reinterpret_cast<TileB>(smem_u32 + 4 * BM * kTPitch + 2 * BN * kTPitch);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / WN, wn = warp % WN;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int g = lane >> 2, q = lane & 3;

    // staging: A tile BM x 32 floats = BM*8 float4 ; W tile BN x 16 pairs = BN*4 uint4 (hi) + same (lo)
    constexpr int A_IT = (BM * 8 + 255) / 256, B_IT = (BN * 4 + 255) / 256;
    float4 a_st[A_IT];
    uint4 bh_st[B_IT], bl_st[B_IT];
    const int Kp = K >> 1;  // pairs per W row

    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int i = 0; i < A_IT; ++i) {
            const int idx = tid + i * 256, r = idx >> 3, c4 = idx & 7;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < BM && m0 + r < M) v = *reinterpret_cast<const float4*>(A + (size_t)(m0 + r) * lda + k0 + c4 * 4);
            a_st[i] = v;
        }
#pragma unroll
        for (int i = 0; i < B_IT; ++i) {
            const int idx = tid + i * 256, r = idx >> 2, c4 = idx & 3;
            uint4 vh = make_uint4(0, 0, 0, 0), vl = make_uint4(0, 0, 0, 0);
            if (r < BN && n0 + r < N) {
                const size_t off = (size_t)(n0 + r) * Kp + (k0 >> 1) + c4 * 4;
                vh = __ldg(reinterpret_cast<const uint4*>(Whi + off));
                vl = __ldg(reinterpret_cast<const uint4*>(Wlo + off));
            }
            bh_st[i] = vh;
            bl_st[i] = vl;
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_IT; ++i) {
            const int idx = tid + i * 256, r = idx >> 3, c4 = idx & 7;
            if (r < BM) {
                uint32_t h0, l0, h1, l1;
                split2(a_st[i].x, a_st[i].y, h0, l0);
                split2(a_st[i].z, a_st[i].w, h1, l1);
                *reinterpret_cast<uint2*>(&Ah[buf][r][c4 * 2]) = make_uint2(h0, h1);
                *reinterpret_cast<uint2*>(&Al[buf][r][c4 * 2]) = make_uint2(l0, l1);
            }
        }
#pragma unroll
        for (int i = 0; i < B_IT; ++i) {
            const int idx = tid + i * 256, r = idx >> 2, c4 = idx & 3;
            if (r < BN) {
                *reinterpret_cast<uint4*>(&Bh[buf][r][c4 * 4]) = bh_st[i];
                *reinterpret_cast<uint4*>(&Bl[buf][r][c4 * 4]) = bl_st[i];
            }
        }
    };

    float acc[MI][NI][4];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[i][j][c] = 0.f;

    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    const int nk = K / kTK;
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load_tiles((kt + 1) * kTK);
#pragma unroll
        for (int ks = 0; ks < kTK / 16; ++ks) {
            uint32_t ah[MI][4], al[MI][4], bh[NI][2], bl[NI][2];
#pragma unroll
            for (int i = 0; i < MI; ++i) {
                const int r = wm * WTM + i * 16 + g;
                const int c = ks * 8 + q;
                ah[i][0] = Ah[buf][r][c];     ah[i][1] = Ah[buf][r + 8][c];
                ah[i][2] = Ah[buf][r][c + 4]; ah[i][3] = Ah[buf][r + 8][c + 4];
                al[i][0] = Al[buf][r][c];     al[i][1] = Al[buf][r + 8][c];
                al[i][2] = Al[buf][r][c + 4]; al[i][3] = Al[buf][r + 8][c + 4];
            }
#pragma unroll
            for (int j = 0; j < NI; ++j) {
                const int r = wn * WTN + j * 8 + g;
                const int c = ks * 8 + q;
                bh[j][0] = Bh[buf][r][c]; bh[j][1] = Bh[buf][r][c + 4];
                bl[j][0] = Bl[buf][r][c]; bl[j][1] = Bl[buf][r][c + 4];
            }
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) {
                    mma_bf16_16816(acc[i][j], al[i], bh[j]);   // small terms first
                    mma_bf16_16816(acc[i][j], ah[i], bl[j]);
                    mma_bf16_16816(acc[i][j], ah[i], bh[j]);
                }
        }
        if (kt + 1 < nk) store_tiles(buf ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int m = m0 + wm * WTM + i * 16 + g + (c >> 1) * 8;
                const int n = n0 + wn * WTN + j * 8 + q * 2 + (c & 1);
                if (m < M && n < N) epilogue_store(ep, m, n, acc[i][j][c]);
            }
}

}  // namespace

int linear_forward(const float* A, int lda, int M, const LinearWeights& W, const LinearEpilogue& ep,
                   int precision, cudaStream_t stream) {
    if (M <= 0) return BVC_OK;
    if (W.K % 16 != 0 || (lda & 3) != 0) {
        set_error("linear_forward: K must be a multiple of 16 and lda of 4");
        return BVC_ERR_INVALID;
    }
    const int N = W.N, K = W.K;
    if (precision >= 1 && W.w_hi && (K % kTK) == 0) {
        if (M > 512) {
            constexpr int smem = (4 * 128 + 4 * 128) * kTPitch * 4;
            static bool attr_set[kMaxDevices] = {};
            const int dslot = device_slot();
            if (!attr_set[dslot]) {
                BVC_CUDA(cudaFuncSetAttribute(linear_bf16x3_kernel<128, 128, 2, 4>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
                attr_set[dslot] = true;
            }
            dim3 grid((N + 127) / 128, (M + 127) / 128);
            linear_bf16x3_kernel<128, 128, 2, 4><<<grid, 256, smem, stream>>>(A, lda, M, W.w_hi, W.w_lo, N, K, ep);
        } else {
            constexpr int smem = (4 * 32 + 4 * 64) * kTPitch * 4;
            dim3 grid((N + 63) / 64, (M + 31) / 32);
            linear_bf16x3_kernel<32, 64, 2, 4><<<grid, 256, smem, stream>>>(A, lda, M, W.w_hi, W.w_lo, N, K, ep);
        }
    } else if (M > 512) {
        dim3 grid((N + 63) / 64, (M + 63) / 64);
        linear_fp32_kernel<64, 64, 4, 4><<<grid, 256, 0, stream>>>(A, lda, M, W.w, N, K, ep);
    } else {
        dim3 grid((N + 63) / 64, (M + 31) / 32);
        linear_fp32_kernel<32, 64, 2, 4><<<grid, 256, 0, stream>>>(A, lda, M, W.w, N, K, ep);
    }
    BVC_CHECK_LAUNCH();
    return BVC_OK;
}

}  // namespace bvc
