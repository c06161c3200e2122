// Program format of the persistent recurrent kernel (recurrent_cluster.cu) and its host-side builder (bvrnn.cu).
//
// Operand "images".  Every matrix the kernel reads is stored in global memory as the exact shared-memory
// image the tensor core wants, so that one bulk copy (cp.async.bulk, no tensor map) moves a whole operand
// chunk and no thread ever computes a swizzled address while loading:
//   chunk      = R rows x 64 bf16 (128 bytes per row), hi part followed by lo part (split bf16, x = hi + lo)
//   row r      = 128 contiguous bytes at r * 128; its 16-byte piece c is stored at piece position c ^ (r & 7)
//                (the UMMA canonical K-major SWIZZLE_128B layout, SBO = 1024 bytes)
//   activation = [m_tile][k_chunk] chunks of R = 128 rows            (32 KiB each)
//   weight     = [n_tile][k_chunk] chunks of R = bn rows (bn = 64, 48 or 16; nn.Linear row = output feature)
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace bvc {
namespace rec {

enum Kind { KIND_LINEAR = 0, KIND_BOTTLENECK = 1, KIND_MEL = 2, KIND_GRU = 3 };

constexpr int CLUSTER = 4;          // CTAs per cluster = K split of a layer
constexpr int TILE_M = 128;         // batch rows per m-tile (UMMA M)
constexpr int CHUNK_K = 64;         // K elements per chunk
constexpr int ACT_PART_BYTES = TILE_M * 128;
constexpr int ACT_CHUNK_BYTES = 2 * ACT_PART_BYTES;

constexpr int MAX_OPS = 20;
constexpr int MAX_PHASES = 16;
constexpr int MAX_CLUSTERS = 32;
constexpr int MAX_MTILES = 32;
constexpr int MAX_ENTRIES = 8192;

// One Linear layer:  out = epilogue(A . W^T), evaluated as 128-row x bn-column tiles.
struct Op {
    const unsigned char* w_img;    // weight images [n_tiles][k_chunks], chunk = 2 * bn * 128 bytes
    const float* bias;             // [n_tiles * bn] or null
    const float* addend;           // fp32 [M][ldadd] (+ t * add_tstride) or null
    float* out_f;                  // fp32 [M][ldo] or null
    unsigned char* out_img;        // activation images of the output [m_tiles][out_kchunks] or null
    long long add_tstride;
    int ldadd, ldo;
    int out_kchunks;
    int N;                         // valid output columns
    int bn;                        // tile width: 64 (split-K), 48 (GRU, split-K) or 16 (K = 64 layers, no split)
    int kind;
    int act;                       // ELU
    int stack;                     // 1: a_hi x [w_hi | w_lo] as one MMA of width 2 bn into [main | aux] (two adjacent accumulator
                                   // slots) + a_lo x w_hi; 0: three MMAs into one slot.  A property of the LAYER, not of the
                                   // schedule, so that a row's arithmetic does not depend on batch size or cluster count.
};

// All ops of a phase read the same activation matrix A.
struct Phase {
    const unsigned char* a_img;    // activation images [m_tiles][k_chunks]
    int k_chunks;                  // K / 64
    int split;                     // 1: the 4 CTAs of a cluster each take K/4 and reduce through DSMEM
                                   // 0: each CTA takes every 4th entry of its cluster's list with the full K
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
    unsigned char* h_img;          // activation images of h
    const float* gh;               // [M][3H] W_hh h + b_hh, gate-interleaved columns
    float* mel_out;                // [M][T][X] or null
    const float* uniforms;         // [M][T][Z] or null: sampled bits round(u - 0.5 + p) (bvrnn.py:126)
};

// A frame = n_phases phases.  Cluster c works on m-tile cluster_mtile[c]; in phase p it executes the entries
// entries[entry_start[p * (n_clusters + 1) + c] .. entry_start[p * (n_clusters + 1) + c + 1]).
// An entry is (op index << 16) | n-tile index.  Phases are separated by a barrier over the CTAs of one m-tile.
struct Program {
    Frame frame;
    int n_phases;
    int n_ops;
    int n_clusters;
    int n_mtiles;
    int debug_flags;
    int trace_frames;              // bring-up: per CTA, frame < trace_frames, phase: TRACE_EVENTS %globaltimer stamps
    unsigned long long* trace;
    Op ops[MAX_OPS];
    Phase phases[MAX_PHASES];
    int cluster_mtile[MAX_CLUSTERS];
    int mtile_ctas[MAX_MTILES];    // CTAs per barrier domain
    int entry_start[MAX_PHASES * (MAX_CLUSTERS + 1)];
    uint32_t entries[MAX_ENTRIES];
};

int max_clusters(int device, int* out);
size_t smem_bytes();
// h0 (or zeros) -> fp32 state + activation images of the state
int init_state(const float* h0, float* h, unsigned char* h_img, int M, int H, cudaStream_t stream);
int launch(const Program* prog_dev, int n_clusters, unsigned* sync_words /* [0] abort flag, [32 (1 + m)] barrier of m-tile m */,
           cudaStream_t stream);
constexpr int SYNC_WORDS = 32 * (1 + MAX_MTILES);
constexpr int TRACE_EVENTS = 40;

}  // namespace rec
}  // namespace bvc
