// Program format of the persistent recurrent kernel (recurrent_umma.cu) and its host-side builder (bvrnn.cu).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace bvc {
namespace rec {

enum Kind { KIND_LINEAR = 0, KIND_BOTTLENECK = 1, KIND_MEL = 2, KIND_GRU = 3 };

constexpr int MAX_OPS = 24;
constexpr int MAX_PHASES = 16;
constexpr int MAX_GRID = 160;
constexpr int MAX_TILES = 6144;

// One Linear layer evaluated as 64-row x bn-column tiles:  out = epilogue(A . W^T)
struct Op {
    const __nv_bfloat16* a_hi;     // activations, split bf16, [M][lda]
    const __nv_bfloat16* a_lo;     // may be null (exactly representable inputs)
    const __nv_bfloat16* w_hi;     // weights, split bf16, [N][K] (nn.Linear layout), K % 64 == 0
    const __nv_bfloat16* w_lo;
    const float* bias;             // [N] or null
    const float* addend;           // fp32 [M][ldadd] (+ t * add_tstride) or null
    float* out_f;                  // fp32 [M][ldo] or null
    __nv_bfloat16* out_hi;         // split bf16 [M][ldos] or null
    __nv_bfloat16* out_lo;
    long long a_tstride;           // elements added to a_hi / a_lo per frame
    long long add_tstride;
    int lda, ldadd, ldo, ldos;
    int N, K;
    int kind;
    int act;                       // ELU
    int bn;                        // tile width: 32, or 48 for the GRU layer
    int pad_;
};

// per-call constants shared by the epilogues
struct Frame {
    int M, T, X, Z, H, var_bit;
    float bits_scalar;
    int pad_;
    const float* bits;             // [M][T] or null
    float* codes;                  // [M][T][Z]
    unsigned long long* packed;    // [M][T] or null
    float* logits;                 // [M][T][Z] or null
    float* all_h;                  // [M][T][H] or null
    float* h;                      // [M][H] fp32 state (updated in place)
    const float* gh;               // [M][3H] W_hh h + b_hh, gate-interleaved columns
    const float* mean;
    const float* std;
    float* mel_out;                // [M][T][X] or null
};

// A frame = n_phases phases; phase p, CTA c executes tiles[list_start[p*G + c] .. list_start[p*G + c + 1]).
// A tile entry is (op index << 20) | tile index.
struct Program {
    Frame frame;
    int n_phases;
    int n_ops;
    int grid;
    int debug_flags;               // experiments only: 1 skip copies, 2 skip MMAs, 4 skip epilogue, 8 skip grid barrier
    Op ops[MAX_OPS];
    int list_start[MAX_PHASES * MAX_GRID + 1];
    uint32_t tiles[MAX_TILES];
};

int max_grid(int device, int* out);
int init_state(const float* h0, float* h, __nv_bfloat16* h_hi, __nv_bfloat16* h_lo, int n, __nv_bfloat16* mn_hi,
               __nv_bfloat16* mn_lo, int n_mn, cudaStream_t stream);
size_t umma_smem_bytes();
int umma_launch(const Program* prog_dev, int grid, unsigned* sync_words /* [0] barrier counter, [1] abort flag */,
                cudaStream_t stream);

}  // namespace rec
}  // namespace bvc
