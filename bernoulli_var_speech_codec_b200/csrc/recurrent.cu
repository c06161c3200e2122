// Persistent recurrent kernel for the BVRNN time loop (tensor-core path).
//
// Replaces the per-frame Python loop of reference bvrnn.py:186-206 (encode) and :222-227 (decode),
// i.e. ~15 nn.Linear launches + ~250 small ATen kernels per frame, with ONE cooperative kernel that
// stays resident on every SM for all T frames:
//   * a frame is a fixed program of "phases"; a phase is a set of 32-row x (64|96)-column GEMM tiles
//     that only depend on earlier phases; a device-wide barrier separates phases
//   * the host builds the program (recurrent_host.cu): critical-path layers form the phases, layers off
//     the critical path (W_hh h, dec.0_h h, W_ih_z phi_z) are scheduled as background tiles into the
//     slack of other phases so that every SM has work in every phase
//   * activations travel between phases as split bf16 (hi, lo) rows through L2; weights are split bf16
//     too; a.w ~= a_hi.w_hi + a_hi.w_lo + a_lo.w_hi on mma.sync.m16n8k16 with fp32 accumulation
//   * operands are streamed global->shared with a 4-stage cp.async pipeline (16-byte chunks)
//   * epilogues fuse bias, the hoisted addend, ELU, the Bernoulli bottleneck (threshold, bit-budget
//     mask, code / packed-word / logit outputs), the mel normalisation and the GRU gate math
//     (gate-interleaved weight rows keep r, z, n of one hidden unit in one thread)
#include <cuda_bf16.h>

#include "common.cuh"
#include "recurrent.cuh"

namespace bvc {
namespace rec {

namespace {

constexpr int BM = 32, BK = 64, STAGES = 4, PITCH = 72;    // bf16 elements; 144-byte rows: conflict-free fragments
constexpr int MAX_BN = 96;
constexpr int STAGE_ELEMS = (2 * BM + 2 * MAX_BN) * PITCH; // Ah, Al, Wh, Wl
constexpr int kThreads = 256;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ void split_pair(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
    const __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0));
    const __nv_bfloat16 l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
    hi = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    lo = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
}

__device__ __forceinline__ float ld_cg(const float* p) { return __ldcg(p); }

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Device-wide barrier: monotonically increasing arrival counter (zeroed by the host before launch).
// All CTAs are co-resident (cooperative launch).  A watchdog turns a would-be hang into an error flag.
__device__ __forceinline__ bool grid_barrier(unsigned* counter, unsigned target, int* abort_flag) {
    __syncthreads();
    __shared__ int s_abort;
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        int ab = 0;
        const long long t_start = clock64();
        while (ld_acquire(counter) < target) {
            if (clock64() - t_start > 4000000000LL) {   // ~2 s: something is badly wrong
                atomicExch(abort_flag, 1);
                ab = 1;
                break;
            }
            if (*(volatile int*)abort_flag) { ab = 1; break; }
        }
        s_abort = ab;
        __threadfence();
    }
    __syncthreads();
    return s_abort != 0;
}

// ---------------------------------------------------------------------------------------------
// One GEMM tile: rows [m0, m0+32), columns [n0, n0 + 32 NI), K streamed in 64-wide stages.
// 8 warps as 2 (m) x 4 (n); a warp owns 16 rows x 8 NI columns.
// ---------------------------------------------------------------------------------------------
template <int NI>
__device__ __forceinline__ void gemm_tile(const Op& op, const Frame& fr, int tile, int t, __nv_bfloat16* smem) {
    constexpr int BN = 32 * NI;
    const int M = fr.M;
    const int m_tiles = (M + BM - 1) / BM;
    const int m_tile = tile % m_tiles, n_tile = tile / m_tiles;
    const int m0 = m_tile * BM, n0 = n_tile * BN;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    const int g = lane >> 2, q = lane & 3;
    const int K = op.K, kiters = K / BK;

    const __nv_bfloat16* a_hi = op.a_hi + (size_t)t * op.a_tstride;
    const __nv_bfloat16* a_lo = op.a_lo ? op.a_lo + (size_t)t * op.a_tstride : nullptr;

    auto load_stage = [&](int slot, int k0) {
        __nv_bfloat16* st = smem + (size_t)slot * STAGE_ELEMS;
        __nv_bfloat16* Ah = st;
        __nv_bfloat16* Al = st + BM * PITCH;
        __nv_bfloat16* Wh = st + 2 * BM * PITCH;
        __nv_bfloat16* Wl = Wh + MAX_BN * PITCH;
        {
            const int r = tid >> 3, ch = tid & 7;
            const int row = min(m0 + r, M - 1);
            const size_t go = (size_t)row * op.lda + k0 + ch * 8;
            cp_async16(Ah + r * PITCH + ch * 8, a_hi + go);
            if (a_lo) cp_async16(Al + r * PITCH + ch * 8, a_lo + go);
        }
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const int c = tid + i * kThreads;
            const int r = c >> 3, ch = c & 7;
            const int row = min(n0 + r, op.N - 1);
            const size_t go = (size_t)row * K + k0 + ch * 8;
            cp_async16(Wh + r * PITCH + ch * 8, op.w_hi + go);
            cp_async16(Wl + r * PITCH + ch * 8, op.w_lo + go);
        }
    };

    float acc[NI][4];
#pragma unroll
    for (int j = 0; j < NI; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[j][c] = 0.f;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < kiters) load_stage(s, s * BK);
        cp_async_commit();
    }
    for (int kt = 0; kt < kiters; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        if (kt + STAGES - 1 < kiters) load_stage((kt + STAGES - 1) % STAGES, (kt + STAGES - 1) * BK);
        cp_async_commit();
        const __nv_bfloat16* st = smem + (size_t)(kt % STAGES) * STAGE_ELEMS;
        const uint32_t* Ah = reinterpret_cast<const uint32_t*>(st);
        const uint32_t* Al = reinterpret_cast<const uint32_t*>(st + BM * PITCH);
        const uint32_t* Wh = reinterpret_cast<const uint32_t*>(st + 2 * BM * PITCH);
        const uint32_t* Wl = reinterpret_cast<const uint32_t*>(st + 2 * BM * PITCH + MAX_BN * PITCH);
        constexpr int PW = PITCH / 2;
#pragma unroll
        for (int ks = 0; ks < BK / 16; ++ks) {
            uint32_t ah[4], al[4];
            const int ar = (wm * 16 + g) * PW + ks * 8 + q;
            ah[0] = Ah[ar]; ah[1] = Ah[ar + 8 * PW]; ah[2] = Ah[ar + 4]; ah[3] = Ah[ar + 8 * PW + 4];
            if (a_lo) { al[0] = Al[ar]; al[1] = Al[ar + 8 * PW]; al[2] = Al[ar + 4]; al[3] = Al[ar + 8 * PW + 4]; }
#pragma unroll
            for (int j = 0; j < NI; ++j) {
                const int br = (wn * 8 * NI + j * 8 + g) * PW + ks * 8 + q;
                uint32_t bh[2], bl[2];
                bh[0] = Wh[br]; bh[1] = Wh[br + 4];
                bl[0] = Wl[br]; bl[1] = Wl[br + 4];
                if (a_lo) mma16816(acc[j], al, bh);
                mma16816(acc[j], ah, bl);
                mma16816(acc[j], ah, bh);
            }
        }
    }
    cp_async_wait<0>();

    // ------------------------------- epilogues -------------------------------
    const int row0 = m0 + wm * 16 + g;                 // rows row0 and row0 + 8
    if (op.kind == KIND_LINEAR || op.kind == KIND_MEL) {
#pragma unroll
        for (int j = 0; j < NI; ++j) {
            const int n = n0 + wn * 8 * NI + j * 8 + 2 * q;
            if (n >= op.N) continue;
            const float b0 = op.bias ? __ldg(op.bias + n) : 0.f, b1 = op.bias ? __ldg(op.bias + n + 1) : 0.f;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int m = row0 + hh * 8;
                if (m >= M) continue;
                float v0 = acc[j][hh * 2] + b0, v1 = acc[j][hh * 2 + 1] + b1;
                if (op.addend) {
                    const float* ap = op.addend + (size_t)t * op.add_tstride + (size_t)m * op.ldadd + n;
                    v0 += ld_cg(ap);
                    v1 += ld_cg(ap + 1);
                }
                if (op.act) { v0 = elu1(v0); v1 = elu1(v1); }
                if (op.kind == KIND_MEL) {
                    if (fr.mel_out) {
                        float* mo = fr.mel_out + ((size_t)m * fr.T + t) * fr.X + n;
                        mo[0] = v0; mo[1] = v1;
                    }
                    v0 = (v0 - __ldg(fr.mean + n)) / __ldg(fr.std + n);
                    v1 = (v1 - __ldg(fr.mean + n + 1)) / __ldg(fr.std + n + 1);
                }
                if (op.out_f) {
                    float* o = op.out_f + (size_t)m * op.ldo + n;
                    *reinterpret_cast<float2*>(o) = make_float2(v0, v1);
                }
                if (op.out_hi) {
                    uint32_t hi, lo;
                    split_pair(v0, v1, hi, lo);
                    *reinterpret_cast<uint32_t*>(op.out_hi + (size_t)m * op.ldos + n) = hi;
                    *reinterpret_cast<uint32_t*>(op.out_lo + (size_t)m * op.ldos + n) = lo;
                }
            }
        }
    } else if (op.kind == KIND_BOTTLENECK) {
        // N = Z = 64 (one n-tile).  z = round(sigmoid(logit)), masked to 0.5 beyond the bit budget (bvrnn.py:191-196)
        __shared__ unsigned long long rowbits[BM];
        if (tid < BM) rowbits[tid] = 0ull;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < NI; ++j) {
            const int n = n0 + wn * 8 * NI + j * 8 + 2 * q;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int m = row0 + hh * 8;
                if (m >= M || n >= op.N) continue;
                const float budget = fr.bits ? __ldg(fr.bits + (size_t)m * fr.T + t) : fr.bits_scalar;
                float code[2];
                unsigned long long word = 0ull;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float lg = acc[j][hh * 2 + e] + __ldg(op.bias + n + e);
                    const bool active = !fr.var_bit || (budget > (float)(n + e));
                    const bool bit = active && (sigmoidf_(lg) > 0.5f);
                    code[e] = active ? (bit ? 1.f : 0.f) : 0.5f;
                    if (bit) word |= 1ull << (n + e);
                    if (fr.logits) fr.logits[((size_t)m * fr.T + t) * fr.Z + n + e] = lg;
                }
                float* co = fr.codes + ((size_t)m * fr.T + t) * fr.Z + n;
                *reinterpret_cast<float2*>(co) = make_float2(code[0], code[1]);
                // {0, 0.5, 1} are exact in bf16: hi = value, lo = 0
                uint32_t hi, lo;
                split_pair(code[0], code[1], hi, lo);
                *reinterpret_cast<uint32_t*>(op.out_hi + (size_t)m * op.ldos + n) = hi;
                if (word) atomicOr(&rowbits[m - m0], word);
            }
        }
        __syncthreads();
        if (fr.packed && tid < BM && m0 + tid < M) fr.packed[(size_t)(m0 + tid) * fr.T + t] = rowbits[tid];
    } else if (op.kind == KIND_GRU) {
        // columns are gate-interleaved in groups of 24: [r(8) z(8) n(8)]; NI == 3 -> acc[gate][.]
        if (NI == 3) {
            const int H = fr.H;
            const int ncol0 = n0 + wn * 24 + 2 * q;              // column of gate r, element e = 0
            const int jh0 = (n0 / 24 + wn) * 8 + 2 * q;          // hidden unit of element e = 0
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int m = row0 + hh * 8;
                if (m >= M) continue;
                const float* gz = op.addend + (size_t)t * op.add_tstride + (size_t)m * op.ldadd + ncol0;
                const float* gh = fr.gh + (size_t)m * 3 * H + ncol0;
                float hn[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float gi_r = acc[0][hh * 2 + e] + ld_cg(gz + e);
                    const float gi_z = acc[NI > 1 ? 1 : 0][hh * 2 + e] + ld_cg(gz + 8 + e);
                    const float gi_n = acc[NI > 2 ? 2 : 0][hh * 2 + e] + ld_cg(gz + 16 + e);
                    const float r = sigmoidf_(gi_r + ld_cg(gh + e));
                    const float z = sigmoidf_(gi_z + ld_cg(gh + 8 + e));
                    const float nn = tanhf(gi_n + r * ld_cg(gh + 16 + e));
                    const float hv = ld_cg(fr.h + (size_t)m * H + jh0 + e);
                    if (fr.all_h) fr.all_h[((size_t)m * fr.T + t) * H + jh0 + e] = hv;   // state entering frame t
                    hn[e] = (hv - nn) * z + nn;
                }
                *reinterpret_cast<float2*>(fr.h + (size_t)m * H + jh0) = make_float2(hn[0], hn[1]);
                uint32_t hi, lo;
                split_pair(hn[0], hn[1], hi, lo);
                *reinterpret_cast<uint32_t*>(op.out_hi + (size_t)m * op.ldos + jh0) = hi;
                *reinterpret_cast<uint32_t*>(op.out_lo + (size_t)m * op.ldos + jh0) = lo;
            }
        }
    }
    __syncthreads();   // shared-memory stages are reused by the next tile
}

__global__ void __launch_bounds__(kThreads, 1)
recurrent_kernel(const Program* __restrict__ prog, unsigned* barrier_counter, int* abort_flag) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __nv_bfloat16* smem = reinterpret_cast<__nv_bfloat16*>(smem_raw);
    const Frame& fr = prog->frame;
    const int G = gridDim.x, cta = blockIdx.x;
    const int n_phases = prog->n_phases;
    unsigned bar = 0;
    for (int t = 0; t < fr.T; ++t) {
        for (int ph = 0; ph < n_phases; ++ph) {
            const int lb = prog->list_start[ph * G + cta], le = prog->list_start[ph * G + cta + 1];
            for (int i = lb; i < le; ++i) {
                const uint32_t e = prog->tiles[i];
                const Op& op = prog->ops[e >> 20];
                const int tile = (int)(e & 0xFFFFF);
                if (op.ni == 3) gemm_tile<3>(op, fr, tile, t, smem);
                else gemm_tile<2>(op, fr, tile, t, smem);
            }
            bar += (unsigned)G;
            if (grid_barrier(barrier_counter, bar, abort_flag)) return;
        }
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

size_t smem_bytes() { return (size_t)STAGES * STAGE_ELEMS * sizeof(__nv_bfloat16); }

int init_state(const float* h0, float* h, __nv_bfloat16* h_hi, __nv_bfloat16* h_lo, int n, __nv_bfloat16* mn_hi,
               __nv_bfloat16* mn_lo, int n_mn, cudaStream_t stream) {
    const int total = n > n_mn ? n : n_mn;
    init_state_kernel<<<(total + 255) / 256, 256, 0, stream>>>(h0, h, h_hi, h_lo, n, mn_hi, mn_lo, n_mn);
    BVC_CHECK_LAUNCH();
    return BVC_OK;
}

int max_grid(int device, int* out) {
    static int cached[64] = {0};
    if (device >= 0 && device < 64 && cached[device]) { *out = cached[device]; return BVC_OK; }
    BVC_CUDA(cudaFuncSetAttribute(recurrent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes()));
    int per_sm = 0, sms = 0;
    BVC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, recurrent_kernel, kThreads, smem_bytes()));
    BVC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    if (per_sm < 1) { set_error("recurrent kernel does not fit on an SM"); return BVC_ERR_DEVICE; }
    *out = sms;                      // one CTA per SM
    if (device >= 0 && device < 64) cached[device] = sms;
    return BVC_OK;
}

int launch(const Program* prog_dev, int grid, unsigned* sync_words, cudaStream_t stream) {
    BVC_CUDA(cudaMemsetAsync(sync_words, 0, 2 * sizeof(unsigned), stream));
    unsigned* barrier_counter = sync_words;
    int* abort_flag = reinterpret_cast<int*>(sync_words + 1);
    void* args[] = {(void*)&prog_dev, (void*)&barrier_counter, (void*)&abort_flag};
    BVC_CUDA(cudaLaunchCooperativeKernel((const void*)recurrent_kernel, dim3(grid), dim3(kThreads), args, smem_bytes(),
                                         stream));
    if (g_launch_counter) ++*g_launch_counter;
    return BVC_OK;
}

}  // namespace rec
}  // namespace bvc
