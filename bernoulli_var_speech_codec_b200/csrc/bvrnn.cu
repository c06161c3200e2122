// BVRNN variational recurrent coder: frame-sequential encode / decode.
//
// Replaces reference bvrnn.py:163-209 (BVRNN.encode) and :211-229 (BVRNN.decode).
// Work that does not depend on the recurrent state is hoisted out of the time loop and
// run as large GEMMs over all B*T frames:
//   encode: phi_x(y) (bvrnn.py:178, hoisted in the reference too) and the phi_x half of
//           enc.0 (enc.0.weight[:, :H], bvrnn.py:189)
//   decode: phi_z(z) for all frames, the phi_z half of dec.0 and of the GRU's W_ih
// Inside the loop the layers that read h are merged into one GEMM per frame
// ([enc.0_h; dec.0_h; W_hh] . h), and the layers that read phi_z likewise
// ([dec.0_z; W_ih_z] . phi_z), so a frame is 13 (encode) / 8 (decode) dependent GEMMs
// plus the Bernoulli bottleneck and the GRU gate kernel.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "common.cuh"
#include "recurrent.cuh"

namespace bvc {

namespace {

// z = round(sigmoid(logit)) (half-to-even => 0.5 -> 0), masked to 0.5 beyond the frame's bit
// budget (bvrnn.py:191-196).  One thread per (b, i); a warp ballots 32 code bits.
__global__ void __launch_bounds__(256)
bottleneck_kernel(const float* __restrict__ logit, int B, int Z, int T, int t, const float* __restrict__ bits,
                  float bits_scalar, int var_bit, float* __restrict__ codes, uint32_t* __restrict__ packed,
                  float* __restrict__ logits_out, const float* __restrict__ uniforms) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = idx / Z, i = idx - b * Z;
    const bool valid = b < B;
    float lg = 0.f, code = 0.f;
    bool bit = false;
    if (valid) {
        lg = logit[idx];
        const float p = sigmoidf_(lg);
        bit = uniforms ? (rintf((uniforms[((size_t)b * T + t) * Z + i] - 0.5f) + p) == 1.f) : (p > 0.5f);   // bvrnn.py:126 / :191
        const float budget = bits ? bits[(size_t)b * T + t] : bits_scalar;
        const bool active = !var_bit || (budget > (float)i);
        bit = bit && active;
        code = active ? (bit ? 1.f : 0.f) : 0.5f;
        const size_t o = ((size_t)b * T + t) * Z + i;
        codes[o] = code;
        if (logits_out) logits_out[o] = lg;
    }
    const uint32_t word = __ballot_sync(0xffffffffu, bit);
    if (packed && valid && (threadIdx.x & 31) == 0 && Z == 64)
        packed[((size_t)b * T + t) * 2 + (i >> 5)] = word;
}

// GRU cell gates (PyTorch order r,z,n; reference bvrnn.py:83,206):
//   r = s(gi_r + gh_r), z = s(gi_z + gh_z), n = tanh(gi_n + r * gh_n), h' = (h - n) * z + n
__global__ void __launch_bounds__(256)
gru_gate_kernel(const float* __restrict__ gi, int ldgi, const float* __restrict__ gh, int ldgh,
                const float* __restrict__ h, float* __restrict__ h_next, int B, int H,
                float* __restrict__ all_h, int T, int t) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * H) return;
    const int b = idx / H, j = idx - b * H;
    const float* gib = gi + (size_t)b * ldgi;
    const float* ghb = gh + (size_t)b * ldgh;
    const float r = sigmoidf_(gib[j] + ghb[j]);
    const float z = sigmoidf_(gib[H + j] + ghb[H + j]);
    const float n = tanhf(gib[2 * H + j] + r * ghb[2 * H + j]);
    const float hv = h[idx];
    if (all_h) all_h[((size_t)b * T + t) * H + j] = hv;   // state entering frame t (bvrnn.py:205)
    h_next[idx] = (hv - n) * z + n;
}

__global__ void sigmoid_kernel(float* __restrict__ x, size_t n) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n) x[idx] = sigmoidf_(x[idx]);
}

__global__ void normalize_kernel(const float* __restrict__ y, const float* __restrict__ mean,
                                 const float* __restrict__ std, float* __restrict__ out, size_t n, int X) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const int c = (int)(idx % X);
    out[idx] = (y[idx] - mean[c]) / std[c];
}

// packed word -> the reference's float code vector: bit i < budget -> {0, 1}, masked bits -> 0.5 (bvrnn.py:191-196)
__global__ void __launch_bounds__(256)
unpack_codes_kernel(const unsigned long long* __restrict__ packed, const float* __restrict__ bits, float bits_scalar,
                    int var_bit, size_t n_frames, int Z, float* __restrict__ codes) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_frames * (size_t)Z) return;
    const size_t f = idx / Z;
    const int i = (int)(idx - f * Z);
    const float budget = bits ? bits[f] : bits_scalar;
    const bool active = !var_bit || (budget > (float)i);
    codes[idx] = active ? (float)((packed[f] >> i) & 1ull) : 0.5f;
}

struct Lin {
    const LinearWeights* w;
    const float* bias;
};

int run_linear(const float* A, int lda, int M, const LinearWeights& w, const float* bias, int n_act, float* out,
               int ldo, int precision, cudaStream_t s, const float* addend = nullptr, int ldadd = 0, int n_add = 0) {
    LinearEpilogue ep;
    ep.bias = bias;
    ep.addend = addend;
    ep.ldadd = ldadd;
    ep.n_add = n_add;
    ep.n_act = n_act;
    ep.out = out;
    ep.ldo = ldo;
    return linear_forward(A, lda, M, w, ep, precision, s);
}

#define BVC_TRY(expr)                 \
    do {                              \
        int _rc = (expr);             \
        if (_rc != BVC_OK) return _rc; \
    } while (0)

}  // namespace

int unpack_codes(const unsigned long long* packed, const float* bits, float bits_scalar, int var_bit, size_t n_frames,
                 int Z, float* codes, cudaStream_t s) {
    if (Z > 64) { set_error("unpack_codes: z_dim > 64 does not fit a 64-bit word"); return BVC_ERR_INVALID; }
    const size_t n = n_frames * (size_t)Z;
    if (n == 0) return BVC_OK;
    unpack_codes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(packed, bits, bits_scalar, var_bit, n_frames, Z, codes);
    BVC_CHECK_LAUNCH();
    return BVC_OK;
}

size_t bvrnn_workspace_floats(const BvrnnWeights& w, int B, int T) {
    const size_t BT = (size_t)B * T, H = w.H;
    const size_t Bp = ((size_t)B + 127) / 128 * 128;   // activation images cover whole 128-row m-tiles
    size_t n = 0;
    const size_t BTp = (BT + 127) / 128 * 128;
    n += BTp * 128 + BT * w.X + 64;     // input images (normalised mel padded to K = 128, or codes) / fp32 path: normalised mel
    n += 2 * (BTp * H + 64);            // two hoisted activation buffers (images: hi + lo bf16 = 4 bytes per element)
    n += BT * 4 * H + 64;               // hoisted fp32 output: encode [B*T, H], decode [dec.0_z ; W_ih_z] . phi_z [B*T, 4H]
    n += Bp * (40 * H + 2 * w.X + 2 * w.Z + 256) + 64 * 64;
    n += BT * w.Z + BT + (size_t)B + 128;   // packed decode on the layer path (expanded codes); bit-stream offset tables
    return n;
}

static int bvrnn_encode_layers(BvrnnWeights& w, Workspace& ws, const float* mel, const float* bits, float bits_scalar,
                 const float* h0, int B, int T, float* codes, unsigned long long* packed, float* logits,
                 float* all_h, float* h_final, int precision, cudaStream_t s, const float* uniforms) {
    const int H = w.H, X = w.X, Z = w.Z;
    const size_t BT = (size_t)B * T;
    if (BT > (size_t)INT32_MAX / 4) { set_error("B*T too large"); return BVC_ERR_INVALID; }
    float* yn = ws.take(BT * X);
    float* PA = ws.take(BT * H);
    float* PB = ws.take(BT * H);
    float* hg = ws.take((size_t)B * 5 * H);
    float* e2 = ws.take((size_t)B * H);
    float* lg = ws.take((size_t)B * Z);
    float* z1 = ws.take((size_t)B * H);
    float* z2 = ws.take((size_t)B * H);
    float* pz = ws.take((size_t)B * H);
    float* dg = ws.take((size_t)B * 4 * H);
    float* d2 = ws.take((size_t)B * H);
    float* d3 = ws.take((size_t)B * H);
    float* md = ws.take((size_t)B * X);
    float* mn = ws.take((size_t)B * X);
    float* x1 = ws.take((size_t)B * H);
    float* x2 = ws.take((size_t)B * H);
    float* px = ws.take((size_t)B * H);
    float* gi = ws.take((size_t)B * 3 * H);
    float* hA = ws.take((size_t)B * H);
    float* hB = ws.take((size_t)B * H);
    BVC_TRY(ws_check(ws, "BVRNN.encode (layer path)"));

    // hoisted: yn -> phi_x(yn) for all frames -> enc.0[:, :H] . phi_x    (bvrnn.py:173,178,189)
    {
        const size_t n = BT * X;
        normalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(mel, w.mean, w.std, yn, n, X);
        BVC_CHECK_LAUNCH();
    }
    BVC_TRY(run_linear(yn, X, (int)BT, w.px0, w.b_px0, H, PA, H, precision, s));
    BVC_TRY(run_linear(PA, H, (int)BT, w.px2, w.b_px2, H, PB, H, precision, s));
    BVC_TRY(run_linear(PB, H, (int)BT, w.px4, w.b_px4, H, PA, H, precision, s));
    BVC_TRY(run_linear(PA, H, (int)BT, w.e0x, nullptr, 0, PB, H, precision, s));
    float* E0x = PB;  // [B, T, H]

    if (h0) BVC_CUDA(cudaMemcpyAsync(hA, h0, sizeof(float) * B * H, cudaMemcpyDeviceToDevice, s));
    else BVC_CUDA(cudaMemsetAsync(hA, 0, sizeof(float) * B * H, s));
    float *hc = hA, *hn = hB;

    const int TH = T * H, TZ = T * Z;
    for (int t = 0; t < T; ++t) {
        // [e1 | W_d0h h | W_hh h + b_hh] = [enc.0_h; dec.0_h; W_hh] . h ; e1 = ELU(. + E0x_t + b_e0)
        BVC_TRY(run_linear(hc, H, B, w.hcat_enc, w.b_hcat_enc, H, hg, 5 * H, precision, s,
                           E0x + (size_t)t * H, TH, H));
        BVC_TRY(run_linear(hg, 5 * H, B, w.e2, w.b_e2, H, e2, H, precision, s));
        BVC_TRY(run_linear(e2, H, B, w.e4, w.b_e4, 0, lg, Z, precision, s));
        bottleneck_kernel<<<(B * Z + 255) / 256, 256, 0, s>>>(lg, B, Z, T, t, bits, bits_scalar, w.var_bit, codes,
                                                            reinterpret_cast<uint32_t*>(packed), logits, uniforms);
        BVC_CHECK_LAUNCH();
        // phi_z(z_t)                                                        (bvrnn.py:198)
        BVC_TRY(run_linear(codes + (size_t)t * Z, TZ, B, w.pz0, w.b_pz0, H, z1, H, precision, s));
        BVC_TRY(run_linear(z1, H, B, w.pz2, w.b_pz2, H, z2, H, precision, s));
        BVC_TRY(run_linear(z2, H, B, w.pz4, w.b_pz4, H, pz, H, precision, s));
        // [d1 | W_ih_z phi_z + b_ih] ; d1 = ELU(dec.0_z phi_z + dec.0_h h + b)   (bvrnn.py:202,206)
        BVC_TRY(run_linear(pz, H, B, w.zcat, w.b_zcat, H, dg, 4 * H, precision, s, hg + H, 5 * H, H));
        BVC_TRY(run_linear(dg, 4 * H, B, w.d2, w.b_d2, H, d2, H, precision, s));
        BVC_TRY(run_linear(d2, H, B, w.d4, w.b_d4, H, d3, H, precision, s));
        {
            LinearEpilogue ep;
            ep.bias = w.b_d6;
            ep.out = md;
            ep.ldo = X;
            ep.nmean = w.mean;
            ep.nstd = w.std;
            ep.out2 = mn;
            ep.ldo2 = X;
            BVC_TRY(linear_forward(d3, H, B, w.d6, ep, precision, s));
        }
        // phi_x of the reconstruction                                       (bvrnn.py:204)
        BVC_TRY(run_linear(mn, X, B, w.px0, w.b_px0, H, x1, H, precision, s));
        BVC_TRY(run_linear(x1, H, B, w.px2, w.b_px2, H, x2, H, precision, s));
        BVC_TRY(run_linear(x2, H, B, w.px4, w.b_px4, H, px, H, precision, s));
        // gi = W_ih_x phi_x_gen + (W_ih_z phi_z + b_ih)
        BVC_TRY(run_linear(px, H, B, w.ihx, nullptr, 0, gi, 3 * H, precision, s, dg + H, 4 * H, 3 * H));
        gru_gate_kernel<<<(B * H + 255) / 256, 256, 0, s>>>(gi, 3 * H, hg + 2 * H, 5 * H, hc, hn, B, H, all_h, T, t);
        BVC_CHECK_LAUNCH();
        float* tmp = hc; hc = hn; hn = tmp;
    }
    if (h_final) BVC_CUDA(cudaMemcpyAsync(h_final, hc, sizeof(float) * B * H, cudaMemcpyDeviceToDevice, s));
    return BVC_OK;
}

static int bvrnn_decode_layers(BvrnnWeights& w, Workspace& ws, const float* codes, const float* h0, int B, int T,
                 float* mel, float* h_final, int precision, cudaStream_t s) {
    const int H = w.H, X = w.X, Z = w.Z;
    const size_t BT = (size_t)B * T;
    if (BT > (size_t)INT32_MAX / 4) { set_error("B*T too large"); return BVC_ERR_INVALID; }
    float* PA = ws.take(BT * H);
    float* PB = ws.take(BT * H);
    float* DZ = ws.take(BT * 4 * H);
    float* hd = ws.take((size_t)B * 4 * H);
    float* d2 = ws.take((size_t)B * H);
    float* d3 = ws.take((size_t)B * H);
    float* mn = ws.take((size_t)B * X);
    float* x1 = ws.take((size_t)B * H);
    float* x2 = ws.take((size_t)B * H);
    float* px = ws.take((size_t)B * H);
    float* gi = ws.take((size_t)B * 3 * H);
    float* hA = ws.take((size_t)B * H);
    float* hB = ws.take((size_t)B * H);
    BVC_TRY(ws_check(ws, "BVRNN.decode (layer path)"));

    // hoisted over all frames: phi_z(z), then [dec.0_z ; W_ih_z] . phi_z + [b_d0 ; b_ih]
    BVC_TRY(run_linear(codes, Z, (int)BT, w.pz0, w.b_pz0, H, PA, H, precision, s));
    BVC_TRY(run_linear(PA, H, (int)BT, w.pz2, w.b_pz2, H, PB, H, precision, s));
    BVC_TRY(run_linear(PB, H, (int)BT, w.pz4, w.b_pz4, H, PA, H, precision, s));
    BVC_TRY(run_linear(PA, H, (int)BT, w.zcat, w.b_zcat, 0, DZ, 4 * H, precision, s));

    if (h0) BVC_CUDA(cudaMemcpyAsync(hA, h0, sizeof(float) * B * H, cudaMemcpyDeviceToDevice, s));
    else BVC_CUDA(cudaMemsetAsync(hA, 0, sizeof(float) * B * H, s));
    float *hc = hA, *hn = hB;
    const int T4H = T * 4 * H;
    for (int t = 0; t < T; ++t) {
        const float* DZt = DZ + (size_t)t * 4 * H;
        // [d1 | W_hh h + b_hh] ; d1 = ELU(dec.0_h h + DZ_t[:, :H])          (bvrnn.py:224)
        BVC_TRY(run_linear(hc, H, B, w.hcat_dec, w.b_hcat_dec, H, hd, 4 * H, precision, s, DZt, T4H, H));
        BVC_TRY(run_linear(hd, 4 * H, B, w.d2, w.b_d2, H, d2, H, precision, s));
        BVC_TRY(run_linear(d2, H, B, w.d4, w.b_d4, H, d3, H, precision, s));
        {
            LinearEpilogue ep;
            ep.bias = w.b_d6;
            ep.out = mel + (size_t)t * X;
            ep.ldo = T * X;
            ep.nmean = w.mean;
            ep.nstd = w.std;
            ep.out2 = mn;
            ep.ldo2 = X;
            BVC_TRY(linear_forward(d3, H, B, w.d6, ep, precision, s));
        }
        BVC_TRY(run_linear(mn, X, B, w.px0, w.b_px0, H, x1, H, precision, s));
        BVC_TRY(run_linear(x1, H, B, w.px2, w.b_px2, H, x2, H, precision, s));
        BVC_TRY(run_linear(x2, H, B, w.px4, w.b_px4, H, px, H, precision, s));
        BVC_TRY(run_linear(px, H, B, w.ihx, nullptr, 0, gi, 3 * H, precision, s, DZt + H, T4H, 3 * H));
        gru_gate_kernel<<<(B * H + 255) / 256, 256, 0, s>>>(gi, 3 * H, hd + H, 4 * H, hc, hn, B, H, nullptr, T, t);
        BVC_CHECK_LAUNCH();
        float* tmp = hc; hc = hn; hn = tmp;
    }
    if (h_final) BVC_CUDA(cudaMemcpyAsync(h_final, hc, sizeof(float) * B * H, cudaMemcpyDeviceToDevice, s));
    return BVC_OK;
}


// =============================================================================================
// Persistent-kernel path (precision 1): host-side program builder
// =============================================================================================
namespace {

struct PhaseSpec {
    const unsigned char* a_img;
    int k_chunks;
    int split;
    std::vector<int> ops;
};

struct ProgramBuilder {
    rec::Program* p;
    int n_clusters, M;
    std::vector<PhaseSpec> phases;

    int add_op(const rec::Op& op) {
        p->ops[p->n_ops] = op;
        return p->n_ops++;
    }
    void add_phase(const unsigned char* a_img, int K, int split, std::vector<int> ops) {
        phases.push_back({a_img, K / rec::CHUNK_K, split, std::move(ops)});
    }

    // Clusters are dealt to m-tiles in contiguous, equal blocks (an m-tile is a barrier domain); the n-tiles
    // of a phase's ops are dealt round-robin to the clusters of each m-tile.
    bool finish() {
        if (const char* e = getenv("BVC_REC_STACK_ALL")) {       // experiment: every non-GRU layer as [main | aux]
            if (atoi(e))
                for (int i = 0; i < p->n_ops; ++i)
                    if (p->ops[i].kind != rec::KIND_GRU && p->ops[i].bn == 64) p->ops[i].stack = 1;
        }
        const int n_mt = (M + rec::TILE_M - 1) / rec::TILE_M;
        if (n_mt > n_clusters || n_mt > rec::MAX_MTILES || (int)phases.size() > rec::MAX_PHASES) return false;
        p->n_clusters = n_clusters;
        p->n_mtiles = n_mt;
        p->n_phases = (int)phases.size();
        std::vector<std::vector<int>> of_mtile(n_mt);
        for (int c = 0; c < n_clusters; ++c) {
            const int mt = (int)((long long)c * n_mt / n_clusters);
            p->cluster_mtile[c] = mt;
            of_mtile[mt].push_back(c);
        }
        for (int mt = 0; mt < n_mt; ++mt) p->mtile_ctas[mt] = rec::CLUSTER * (int)of_mtile[mt].size();
        int n_e = 0;
        for (int ph = 0; ph < p->n_phases; ++ph) {
            p->phases[ph].a_img = phases[ph].a_img;
            p->phases[ph].k_chunks = phases[ph].k_chunks;
            p->phases[ph].split = phases[ph].split;
            std::vector<std::vector<uint32_t>> lists(n_clusters);
            for (int mt = 0; mt < n_mt; ++mt) {
                int rr = 0;
                const std::vector<int>& cl = of_mtile[mt];
                for (int oi : phases[ph].ops) {
                    const rec::Op& op = p->ops[oi];
                    const int n_tiles = (op.N + op.bn - 1) / op.bn;
                    for (int nt = 0; nt < n_tiles; ++nt) {
                        lists[cl[rr]].push_back(((uint32_t)oi << 16) | (uint32_t)nt);
                        rr = (rr + 1) % (int)cl.size();
                    }
                }
            }
            for (int c = 0; c < n_clusters; ++c) {
                p->entry_start[ph * (n_clusters + 1) + c] = n_e;
                for (uint32_t e : lists[c]) {
                    if (n_e >= rec::MAX_ENTRIES) return false;
                    p->entries[n_e++] = e;
                }
            }
            p->entry_start[ph * (n_clusters + 1) + n_clusters] = n_e;
        }
        return true;
    }
};

rec::Op linear_op(const WImg& w, const float* bias, int act, unsigned char* out_img, int out_kchunks) {
    rec::Op o;
    memset(&o, 0, sizeof(o));
    o.w_img = w.img; o.N = w.N; o.bn = w.bn;
    o.bias = bias; o.act = act;
    o.out_img = out_img; o.out_kchunks = out_kchunks;
    o.kind = rec::KIND_LINEAR;
    o.stack = 1;      // default for the single-layer phases; the builders clear it for layers that share a phase
    return o;
}

// activation images for an [m_tiles * 128, K] matrix
unsigned char* take_img(Workspace& ws, int M, int K) {
    const size_t bytes = (size_t)((M + rec::TILE_M - 1) / rec::TILE_M) * (K / rec::CHUNK_K) * rec::ACT_CHUNK_BYTES;
    return reinterpret_cast<unsigned char*>(ws.take(bytes / sizeof(float)));
}

std::string abort_message(int flag) {
    return "recurrent kernel aborted: " +
           std::string(flag == 1 ? "phase barrier timed out" : "pipeline wait timed out, code " + std::to_string(flag));
}

// Next staging slot of the launch ring; waits (host) only if the launch that used it 8 launches ago is still running.
int acquire_slot(RecurrentWeights& rw, int kind, rec::Program** out) {
    RecurrentWeights::ProgSlot& sl = rw.slots[rw.next_slot];
    if (!sl.host) {
        BVC_CUDA(cudaMallocHost((void**)&sl.host, sizeof(rec::Program)));
        BVC_CUDA(cudaMallocHost((void**)&sl.flag_host, sizeof(int)));
        BVC_CUDA(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
        BVC_CUDA(cudaEventCreate(&sl.ev_begin));
        BVC_CUDA(cudaEventCreate(&sl.ev_end));
    }
    if (sl.used) {
        BVC_CUDA(cudaEventSynchronize(sl.done));
        if (!sl.checked && *sl.flag_host != 0 && !rw.deferred_abort) rw.deferred_abort = *sl.flag_host;
    }
    sl.used = false;
    sl.checked = true;
    sl.kind = kind;
    sl.call_id = rw.call_seq;
    rw.cur_slot = rw.next_slot;
    rw.next_slot = (rw.next_slot + 1) % RecurrentWeights::PROG_SLOTS;
    *out = sl.host;
    return BVC_OK;
}

// Enqueues program upload, the persistent kernel and the read-back of its abort flag; no host synchronisation.
int run_program(BvrnnWeights& w, ProgramBuilder& pb, cudaStream_t s) {
    RecurrentWeights::ProgSlot& sl = w.rw.slots[w.rw.cur_slot];
    if (const char* e = getenv("BVC_REC_DEBUG")) pb.p->debug_flags = atoi(e);
    if (!pb.finish()) {
        set_error("recurrent program does not fit the static limits (m-tiles / entries)");
        return BVC_ERR_INVALID;
    }
    // bring-up: BVC_REC_TRACE=<file> dumps per-CTA, per-phase %globaltimer stamps of the first frames (synchronous)
    const char* trace_path = getenv("BVC_REC_TRACE");
    const int trace_frames = 6, n_ctas = pb.n_clusters * rec::CLUSTER;
    const size_t trace_n = (size_t)n_ctas * trace_frames * rec::MAX_PHASES * rec::TRACE_EVENTS;
    unsigned long long* trace_dev = nullptr;
    if (trace_path) {
        BVC_CUDA(cudaMalloc(&trace_dev, trace_n * sizeof(unsigned long long)));
        BVC_CUDA(cudaMemsetAsync(trace_dev, 0, trace_n * sizeof(unsigned long long), s));
        pb.p->trace = trace_dev;
        pb.p->trace_frames = trace_frames;
    }
    BVC_CUDA(cudaMemcpyAsync(w.rw.prog_dev, sl.host, sizeof(rec::Program), cudaMemcpyHostToDevice, s));
    BVC_CUDA(cudaEventRecord(sl.ev_begin, s));
    int rc = rec::launch(w.rw.prog_dev, pb.n_clusters, w.rw.sync_words, s);
    if (rc) return rc;
    BVC_CUDA(cudaEventRecord(sl.ev_end, s));
    *sl.flag_host = 0;
    BVC_CUDA(cudaMemcpyAsync(sl.flag_host, w.rw.sync_words, sizeof(int), cudaMemcpyDeviceToHost, s));
    BVC_CUDA(cudaEventRecord(sl.done, s));
    sl.used = true;
    sl.checked = false;
    if (trace_dev) {
        BVC_CUDA(cudaStreamSynchronize(s));
        std::vector<unsigned long long> hbuf(trace_n);
        BVC_CUDA(cudaMemcpy(hbuf.data(), trace_dev, trace_n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        cudaFree(trace_dev);
        static int trace_no = 0;     // one file per launch: <path>.0 (first encode), <path>.1 (first decode), ...
        const std::string path = std::string(trace_path) + "." + std::to_string(trace_no++);
        if (FILE* f = fopen(path.c_str(), "wb")) {
            const int hdr[4] = {n_ctas, trace_frames, rec::MAX_PHASES, rec::TRACE_EVENTS};
            fwrite(hdr, sizeof(int), 4, f);
            fwrite(hbuf.data(), sizeof(unsigned long long), trace_n, f);
            fclose(f);
        }
    }
    return BVC_OK;
}

int cluster_count(int* out) {
    int dev = 0;
    BVC_CUDA(cudaGetDevice(&dev));
    int rc = rec::max_clusters(dev, out);
    if (rc) return rc;
    // experiment / tuning: BVC_REC_CLUSTERS caps the clusters of the persistent kernel (the time loop is latency-bound: at
    // B = 256 it hardly slows down on half of the SMs)
    const char* cap_env = getenv("BVC_REC_CLUSTERS");      // read per call: the chunked-batch test flips it at run time
    const int cap = cap_env ? atoi(cap_env) : 0;
    if (cap > 0 && cap < *out) *out = cap;
    return BVC_OK;
}

}  // namespace

// ---- launch ring services ----
int rec_poll_aborts(RecurrentWeights& rw, bool wait) {
    for (int i = 0; i < RecurrentWeights::PROG_SLOTS; ++i) {
        RecurrentWeights::ProgSlot& sl = rw.slots[i];
        if (!sl.used || sl.checked) continue;
        if (wait) BVC_CUDA(cudaEventSynchronize(sl.done));
        else {
            const cudaError_t q = cudaEventQuery(sl.done);
            if (q == cudaErrorNotReady) continue;
            if (q != cudaSuccess) { set_error(std::string("recurrent kernel: ") + cudaGetErrorString(q)); return BVC_ERR_DEVICE; }
        }
        sl.checked = true;
        if (*sl.flag_host != 0 && !rw.deferred_abort) rw.deferred_abort = *sl.flag_host;
    }
    if (rw.deferred_abort) {
        set_error(abort_message(rw.deferred_abort) + " (reported by the first call after the failed launch)");
        rw.deferred_abort = 0;
        return BVC_ERR_DEVICE;
    }
    return BVC_OK;
}

float rec_launch_ms(RecurrentWeights& rw, int kind, int age) {
    // distinct call ids of this kind still in the ring, newest first
    long long ids[RecurrentWeights::PROG_SLOTS];
    int n = 0;
    for (int i = 0; i < RecurrentWeights::PROG_SLOTS; ++i) {
        const RecurrentWeights::ProgSlot& sl = rw.slots[i];
        if (!sl.used || (kind >= 0 && sl.kind != kind)) continue;
        bool seen = false;
        for (int j = 0; j < n; ++j) seen = seen || ids[j] == sl.call_id;
        if (!seen) ids[n++] = sl.call_id;
    }
    std::sort(ids, ids + n, [](long long a, long long b) { return a > b; });
    if (age < 0 || age >= n) return -1.f;
    float total = 0.f;
    for (int i = 0; i < RecurrentWeights::PROG_SLOTS; ++i) {
        const RecurrentWeights::ProgSlot& sl = rw.slots[i];
        if (!sl.used || sl.call_id != ids[age] || (kind >= 0 && sl.kind != kind)) continue;
        float ms = 0.f;
        if (cudaEventSynchronize(sl.ev_end) == cudaSuccess && cudaEventElapsedTime(&ms, sl.ev_begin, sl.ev_end) == cudaSuccess)
            total += ms;
    }
    return total;
}

void rec_free_slots(RecurrentWeights& rw) {
    for (int i = 0; i < RecurrentWeights::PROG_SLOTS; ++i) {
        RecurrentWeights::ProgSlot& sl = rw.slots[i];
        if (sl.host) cudaFreeHost(sl.host);
        if (sl.flag_host) cudaFreeHost(sl.flag_host);
        if (sl.done) cudaEventDestroy(sl.done);
        if (sl.ev_begin) cudaEventDestroy(sl.ev_begin);
        if (sl.ev_end) cudaEventDestroy(sl.ev_end);
        sl = RecurrentWeights::ProgSlot();
    }
}

// rows per persistent-kernel call: every m-tile needs a cluster, and the entry table is finite
static int persistent_max_rows(int n_clusters) {
    // at least two clusters per m-tile keep a CTA's per-frame entry list within the kernel's shared-memory table
    int mt = n_clusters / 2;
    if (mt > 16) mt = 16;
    if (mt < 1) mt = 1;
    return mt * rec::TILE_M;
}

static int bvrnn_encode_persistent(BvrnnWeights& w, Workspace& ws, const float* mel, const float* bits,
                                   float bits_scalar, const float* h0, int B, int T, float* codes,
                                   unsigned long long* packed, float* logits, float* all_h, float* h_final,
                                   float* mel_hat, cudaStream_t s, const float* uniforms) {
    const int H = w.H, X = w.X, Z = w.Z;
    const size_t BT = (size_t)B * T;
    RecurrentWeights& rw = w.rw;
    int G = 0;
    BVC_TRY(cluster_count(&G));

    unsigned char* ynI = take_img(ws, (int)BT, 128);
    unsigned char* PAi = take_img(ws, (int)BT, H);
    unsigned char* PBi = take_img(ws, (int)BT, H);
    float* E0x = ws.take(BT * H);
    float* hf = ws.take((size_t)B * H);
    float* dh = ws.take((size_t)B * H);
    float* gh = ws.take((size_t)B * 3 * H);
    float* giz = ws.take((size_t)B * 3 * H);
    rw.tap_dh = dh; rw.tap_gh = gh; rw.tap_giz = giz; rw.tap_B = B;
    const int KH = H / rec::CHUNK_K;
    unsigned char *hI = take_img(ws, B, H), *e1I = take_img(ws, B, H), *e2I = take_img(ws, B, H);
    unsigned char *zI = take_img(ws, B, Z), *z1I = take_img(ws, B, H), *z2I = take_img(ws, B, H);
    unsigned char *pzI = take_img(ws, B, H), *d1I = take_img(ws, B, H), *d2I = take_img(ws, B, H);
    unsigned char *d3I = take_img(ws, B, H), *x1I = take_img(ws, B, H), *x2I = take_img(ws, B, H);
    unsigned char* pxI = take_img(ws, B, H);
    BVC_TRY(ws_check(ws, "BVRNN.encode"));

    // hoisted over all frames (tcgen05 GEMMs, gemm_umma.cu): yn = (y - mean) / std -> phi_x -> enc.0[:, :H] . phi_x
    BVC_TRY(to_image(mel, (int)BT, X, w.mean, w.std, ynI, 2, s));
    BVC_TRY(linear_umma(ynI, (int)BT, rw.g_px0, w.b_px0, 1, nullptr, 0, PAi, s));
    BVC_TRY(linear_umma(PAi, (int)BT, rw.g_px2, w.b_px2, 1, nullptr, 0, PBi, s));
    BVC_TRY(linear_umma(PBi, (int)BT, rw.g_px4, w.b_px4, 1, nullptr, 0, PAi, s));
    BVC_TRY(linear_umma(PAi, (int)BT, rw.g_e0x, nullptr, 0, E0x, H, nullptr, s));
    BVC_TRY(rec::init_state(h0, hf, hI, B, H, s));

    rec::Program* p = nullptr;
    BVC_TRY(acquire_slot(rw, 0, &p));
    memset(p, 0, sizeof(rec::Program));
    rec::Frame& fr = p->frame;
    fr.M = B; fr.T = T; fr.X = X; fr.Z = Z; fr.H = H; fr.var_bit = w.var_bit;
    fr.bits_scalar = bits_scalar; fr.bits = bits; fr.codes = codes; fr.packed = packed; fr.logits = logits;
    fr.all_h = all_h; fr.h = hf; fr.h_img = hI; fr.gh = gh; fr.mel_out = mel_hat; fr.uniforms = uniforms;

    ProgramBuilder pb;
    pb.p = p; pb.n_clusters = G; pb.M = B;
    rec::Op o;
    // layers that read h: enc.0_h (critical), dec.0_h and W_hh (consumed later in the frame)
    o = linear_op(rw.e0h, rw.b_e0, 1, e1I, KH);
    o.addend = E0x; o.ldadd = T * H; o.add_tstride = H; o.stack = 0;
    const int op_e1 = pb.add_op(o);
    o = linear_op(rw.d0h, nullptr, 0, nullptr, 0);
    o.out_f = dh; o.ldo = H; o.stack = 0;
    const int op_dh = pb.add_op(o);
    o = linear_op(rw.whh_q, rw.b_hh_q, 0, nullptr, 0);
    o.out_f = gh; o.ldo = 3 * H; o.stack = 0;
    const int op_gh = pb.add_op(o);
    const int op_e2 = pb.add_op(linear_op(rw.e2, w.b_e2, 1, e2I, KH));
    o = linear_op(rw.e4, w.b_e4, 0, zI, Z / rec::CHUNK_K);
    o.kind = rec::KIND_BOTTLENECK;
    const int op_e4 = pb.add_op(o);
    const int op_z1 = pb.add_op(linear_op(rw.pz0, w.b_pz0, 1, z1I, KH));
    const int op_z2 = pb.add_op(linear_op(rw.pz2, w.b_pz2, 1, z2I, KH));
    const int op_pz = pb.add_op(linear_op(rw.pz4, w.b_pz4, 1, pzI, KH));
    // layers that read phi_z: dec.0_z (critical) and W_ih_z (consumed by the GRU)
    o = linear_op(rw.d0z, rw.b_d0, 1, d1I, KH);
    o.addend = dh; o.ldadd = H; o.stack = 0;
    const int op_d1 = pb.add_op(o);
    o = linear_op(rw.ihz_q, rw.b_ih_q, 0, nullptr, 0);
    o.out_f = giz; o.ldo = 3 * H; o.stack = 0;
    const int op_giz = pb.add_op(o);
    const int op_d2 = pb.add_op(linear_op(rw.d2, w.b_d2, 1, d2I, KH));
    const int op_d3 = pb.add_op(linear_op(rw.d4, w.b_d4, 1, d3I, KH));
    // phi_x.0 of the reconstruction, fused with dec.6 and the mel normalisation (bvrnn.py:202-204)
    const int op_x1 = pb.add_op(linear_op(rw.x1f, rw.b_x1f, 1, x1I, KH));
    // The encoder runs the decoder inside its loop (analysis by synthesis, bvrnn.py:198-206): dec_t is exactly the mel
    // BVRNN.decode would compute from these codes and the same h0.  Encoding alone never forms it (x1f skips it); the
    // fused forward asks for it as a side output, like the decode program does.
    int op_mel = -1;
    if (mel_hat) {
        o = linear_op(rw.d6, rw.b_d6p, 0, nullptr, 0);
        o.kind = rec::KIND_MEL;
        op_mel = pb.add_op(o);
    }
    const int op_x2 = pb.add_op(linear_op(rw.px2, w.b_px2, 1, x2I, KH));
    const int op_px = pb.add_op(linear_op(rw.px4, w.b_px4, 1, pxI, KH));
    o = linear_op(rw.ihx_q, nullptr, 0, hI, KH);
    o.kind = rec::KIND_GRU; o.stack = 0;
    o.addend = giz; o.ldadd = 3 * H;
    const int op_gru = pb.add_op(o);

    pb.add_phase(hI, H, 1, {op_e1, op_dh, op_gh});
    pb.add_phase(e1I, H, 1, {op_e2});
    pb.add_phase(e2I, H, 1, {op_e4});
    pb.add_phase(zI, Z, 0, {op_z1});
    pb.add_phase(z1I, H, 1, {op_z2});
    pb.add_phase(z2I, H, 1, {op_pz});
    pb.add_phase(pzI, H, 1, {op_d1, op_giz});
    pb.add_phase(d1I, H, 1, {op_d2});
    pb.add_phase(d2I, H, 1, {op_d3});
    if (op_mel >= 0) pb.add_phase(d3I, H, 1, {op_x1, op_mel});
    else pb.add_phase(d3I, H, 1, {op_x1});
    pb.add_phase(x1I, H, 1, {op_x2});
    pb.add_phase(x2I, H, 1, {op_px});
    pb.add_phase(pxI, H, 1, {op_gru});
    BVC_TRY(run_program(w, pb, s));
    if (h_final) BVC_CUDA(cudaMemcpyAsync(h_final, hf, sizeof(float) * B * H, cudaMemcpyDeviceToDevice, s));
    return BVC_OK;
}

static int bvrnn_decode_persistent(BvrnnWeights& w, Workspace& ws, const float* codes, const float* h0, int B, int T,
                                   float* mel, float* h_final, cudaStream_t s, const unsigned long long* packed,
                                   const float* bits, float bits_scalar) {
    const int H = w.H, X = w.X, Z = w.Z;
    const size_t BT = (size_t)B * T;
    RecurrentWeights& rw = w.rw;
    int G = 0;
    BVC_TRY(cluster_count(&G));

    unsigned char* zI_all = take_img(ws, (int)BT, Z);
    unsigned char* PAi = take_img(ws, (int)BT, H);
    unsigned char* PBi = take_img(ws, (int)BT, H);
    float* DZ = ws.take(BT * 4 * H);
    float* hf = ws.take((size_t)B * H);
    float* gh = ws.take((size_t)B * 3 * H);
    const int KH = H / rec::CHUNK_K;
    unsigned char *hI = take_img(ws, B, H), *d1I = take_img(ws, B, H), *d2I = take_img(ws, B, H);
    unsigned char *d3I = take_img(ws, B, H), *x1I = take_img(ws, B, H), *x2I = take_img(ws, B, H);
    unsigned char* pxI = take_img(ws, B, H);
    BVC_TRY(ws_check(ws, "BVRNN.decode"));

    // hoisted over all frames (tcgen05 GEMMs): phi_z(z), then [dec.0_z ; W_ih_z (gate-interleaved)] . phi_z + [b_d0 ; b_ih]
    if (codes) BVC_TRY(to_image(codes, (int)BT, Z, nullptr, nullptr, zI_all, Z / rec::CHUNK_K, s));
    else BVC_TRY(words_to_image(packed, bits, bits_scalar, w.var_bit, (int)BT, zI_all, s));    // wire words straight to the operand image
    BVC_TRY(linear_umma(zI_all, (int)BT, rw.g_pz0, w.b_pz0, 1, nullptr, 0, PAi, s));
    BVC_TRY(linear_umma(PAi, (int)BT, rw.g_pz2, w.b_pz2, 1, nullptr, 0, PBi, s));
    BVC_TRY(linear_umma(PBi, (int)BT, rw.g_pz4, w.b_pz4, 1, nullptr, 0, PAi, s));
    BVC_TRY(linear_umma(PAi, (int)BT, rw.g_zcat, rw.b_zcat_q, 0, DZ, 4 * H, nullptr, s));
    BVC_TRY(rec::init_state(h0, hf, hI, B, H, s));

    rec::Program* p = nullptr;
    BVC_TRY(acquire_slot(rw, 1, &p));
    memset(p, 0, sizeof(rec::Program));
    rec::Frame& fr = p->frame;
    fr.M = B; fr.T = T; fr.X = X; fr.Z = Z; fr.H = H; fr.var_bit = w.var_bit;
    fr.h = hf; fr.h_img = hI; fr.gh = gh; fr.mel_out = mel;

    ProgramBuilder pb;
    pb.p = p; pb.n_clusters = G; pb.M = B;
    rec::Op o;
    o = linear_op(rw.d0h, nullptr, 1, d1I, KH);
    o.addend = DZ; o.ldadd = T * 4 * H; o.add_tstride = 4 * H;       // includes dec.0 bias
    o.stack = 0;
    const int op_d1 = pb.add_op(o);
    o = linear_op(rw.whh_q, rw.b_hh_q, 0, nullptr, 0);
    o.out_f = gh; o.ldo = 3 * H; o.stack = 0;
    const int op_gh = pb.add_op(o);
    const int op_d2 = pb.add_op(linear_op(rw.d2, w.b_d2, 1, d2I, KH));
    const int op_d3 = pb.add_op(linear_op(rw.d4, w.b_d4, 1, d3I, KH));
    o = linear_op(rw.d6, rw.b_d6p, 0, nullptr, 0);                    // the decoder's output (bvrnn.py:225)
    o.kind = rec::KIND_MEL;
    const int op_mel = pb.add_op(o);
    const int op_x1 = pb.add_op(linear_op(rw.x1f, rw.b_x1f, 1, x1I, KH));
    const int op_x2 = pb.add_op(linear_op(rw.px2, w.b_px2, 1, x2I, KH));
    const int op_px = pb.add_op(linear_op(rw.px4, w.b_px4, 1, pxI, KH));
    o = linear_op(rw.ihx_q, nullptr, 0, hI, KH);
    o.kind = rec::KIND_GRU; o.stack = 0;
    o.addend = DZ + H; o.ldadd = T * 4 * H; o.add_tstride = 4 * H;    // W_ih_z phi_z + b_ih, gate-interleaved
    const int op_gru = pb.add_op(o);

    pb.add_phase(hI, H, 1, {op_d1, op_gh});
    pb.add_phase(d1I, H, 1, {op_d2});
    pb.add_phase(d2I, H, 1, {op_d3});
    pb.add_phase(d3I, H, 1, {op_x1, op_mel});
    pb.add_phase(x1I, H, 1, {op_x2});
    pb.add_phase(x2I, H, 1, {op_px});
    pb.add_phase(pxI, H, 1, {op_gru});
    BVC_TRY(run_program(w, pb, s));
    if (h_final) BVC_CUDA(cudaMemcpyAsync(h_final, hf, sizeof(float) * B * H, cudaMemcpyDeviceToDevice, s));
    return BVC_OK;
}

int bvrnn_encode(BvrnnWeights& w, Workspace& ws, const float* mel, const float* bits, float bits_scalar,
                 const float* h0, int B, int T, float* codes, unsigned long long* packed, float* logits,
                 float* all_h, float* h_final, float* mel_hat, int precision, cudaStream_t s, const float* uniforms) {
    if (!(precision >= 1 && w.rw.ready)) {
        const size_t mark = ws.used;
        BVC_TRY(bvrnn_encode_layers(w, ws, mel, bits, bits_scalar, h0, B, T, codes, packed, logits, all_h, h_final,
                                    precision, s, uniforms));
        if (mel_hat) {      // the layer-by-layer path keeps the second recurrence (same result, bvrnn.py:198-206 vs 222-227)
            ws.used = mark;
            BVC_TRY(bvrnn_decode_layers(w, ws, codes, h0, B, T, mel_hat, nullptr, precision, s));
        }
        return BVC_OK;
    }
    int G = 0;
    BVC_TRY(cluster_count(&G));
    const int max_rows = persistent_max_rows(G);
    const size_t mark = ws.used, T_ = (size_t)T;
    ++w.rw.call_seq;
    for (int b0 = 0; b0 < B; b0 += max_rows) {      // utterances are independent: chunk very large batches
        const int nb = B - b0 < max_rows ? B - b0 : max_rows;
        const size_t r = (size_t)b0;
        ws.used = mark;
        BVC_TRY(bvrnn_encode_persistent(w, ws, mel + r * T_ * w.X, bits ? bits + r * T_ : nullptr, bits_scalar,
                                        h0 ? h0 + r * w.H : nullptr, nb, T, codes + r * T_ * w.Z,
                                        packed ? packed + r * T_ : nullptr, logits ? logits + r * T_ * w.Z : nullptr,
                                        all_h ? all_h + r * T_ * w.H : nullptr, h_final ? h_final + r * w.H : nullptr,
                                        mel_hat ? mel_hat + r * T_ * w.X : nullptr, s,
                                        uniforms ? uniforms + r * T_ * w.Z : nullptr));
    }
    return BVC_OK;
}

// prior(h) = sigmoid(W4 ELU(W2 ELU(W0 h + b0) + b2) + b4) over all B*T states: three hoisted GEMMs (no recurrence)
int bvrnn_prior(BvrnnWeights& w, Workspace& ws, const float* all_h, int B, int T, float* prior, int precision, cudaStream_t s) {
    const size_t BT = (size_t)B * T;
    const int H = w.H, Z = w.Z;
    const size_t mark = ws.used;
    int rc = BVC_OK;
    if (precision >= 1 && w.rw.ready) {
        unsigned char* hI = take_img(ws, (int)BT, H);
        unsigned char* aI = take_img(ws, (int)BT, H);
        rc = ws_check(ws, "BVRNN prior");
        if (!rc) rc = to_image(all_h, (int)BT, H, nullptr, nullptr, hI, H / rec::CHUNK_K, s);
        if (!rc) rc = linear_umma(hI, (int)BT, w.rw.g_pr0, w.b_pr0, 1, nullptr, 0, aI, s);
        if (!rc) rc = linear_umma(aI, (int)BT, w.rw.g_pr2, w.b_pr2, 1, nullptr, 0, hI, s);
        if (!rc) rc = linear_umma(hI, (int)BT, w.rw.g_pr4, w.rw.b_pr4p, 2, prior, Z, nullptr, s);
    } else {
        float* a = ws.take(BT * H);
        float* b = ws.take(BT * H);
        rc = ws_check(ws, "BVRNN prior (layer path)");
        if (!rc) rc = run_linear(all_h, H, (int)BT, w.pr0, w.b_pr0, H, a, H, precision, s);
        if (!rc) rc = run_linear(a, H, (int)BT, w.pr2, w.b_pr2, H, b, H, precision, s);
        if (!rc) rc = run_linear(b, H, (int)BT, w.pr4, w.b_pr4, 0, prior, Z, precision, s);
        if (!rc) {
            const size_t n = BT * Z;
            sigmoid_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(prior, n);
            if (g_launch_counter) ++*g_launch_counter;
            if (cudaGetLastError() != cudaSuccess) { set_error("sigmoid kernel launch failed"); rc = BVC_ERR_DEVICE; }
        }
    }
    ws.used = mark;
    return rc;
}

int bvrnn_decode(BvrnnWeights& w, Workspace& ws, const float* codes, const float* h0, int B, int T, float* mel,
                 float* h_final, int precision, cudaStream_t s, const unsigned long long* packed, const float* bits,
                 float bits_scalar) {
    if (!(precision >= 1 && w.rw.ready)) {
        const size_t mark0 = ws.used;
        if (!codes) {      // the layer-by-layer path works on float codes: expand the words first
            float* tmp = ws.take((size_t)B * T * w.Z);
            BVC_TRY(ws_check(ws, "BVRNN.decode (packed)"));
            BVC_TRY(unpack_codes(packed, bits, bits_scalar, w.var_bit, (size_t)B * T, w.Z, tmp, s));
            codes = tmp;
        }
        const int rc = bvrnn_decode_layers(w, ws, codes, h0, B, T, mel, h_final, precision, s);
        ws.used = mark0;
        return rc;
    }
    int G = 0;
    BVC_TRY(cluster_count(&G));
    const int max_rows = persistent_max_rows(G);
    const size_t mark = ws.used, T_ = (size_t)T;
    ++w.rw.call_seq;
    for (int b0 = 0; b0 < B; b0 += max_rows) {
        const int nb = B - b0 < max_rows ? B - b0 : max_rows;
        const size_t r = (size_t)b0;
        ws.used = mark;
        BVC_TRY(bvrnn_decode_persistent(w, ws, codes ? codes + r * T_ * w.Z : nullptr, h0 ? h0 + r * w.H : nullptr, nb, T,
                                        mel + r * T_ * w.X, h_final ? h_final + r * w.H : nullptr, s,
                                        packed ? packed + r * T_ : nullptr, bits ? bits + r * T_ : nullptr, bits_scalar));
    }
    return BVC_OK;
}

}  // namespace bvc
