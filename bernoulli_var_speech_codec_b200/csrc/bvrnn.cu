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
#include <stdlib.h>
#include <string.h>

#include <algorithm>
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
                  float* __restrict__ logits_out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = idx / Z, i = idx - b * Z;
    const bool valid = b < B;
    float lg = 0.f, code = 0.f;
    bool bit = false;
    if (valid) {
        lg = logit[idx];
        const float p = sigmoidf_(lg);
        bit = p > 0.5f;
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

__global__ void normalize_kernel(const float* __restrict__ y, const float* __restrict__ mean,
                                 const float* __restrict__ std, float* __restrict__ out, size_t n, int X) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const int c = (int)(idx % X);
    out[idx] = (y[idx] - mean[c]) / std[c];
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

size_t bvrnn_workspace_floats(const BvrnnWeights& w, int B, int T) {
    const size_t BT = (size_t)B * T, H = w.H;
    size_t n = 0;
    n += BT * w.X + 64;                 // normalised mel
    n += 2 * (BT * H + 64);             // two hoisted activation buffers
    n += BT * 4 * H + 64;               // decode: hoisted [dec.0_z ; W_ih_z] . phi_z
    n += (size_t)B * (40 * H + 2 * w.X + 2 * w.Z + 256) + 64 * 64;
    return n;
}

static int bvrnn_encode_layers(BvrnnWeights& w, Workspace& ws, const float* mel, const float* bits, float bits_scalar,
                 const float* h0, int B, int T, float* codes, unsigned long long* packed, float* logits,
                 float* all_h, float* h_final, int precision, cudaStream_t s) {
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
                                                            reinterpret_cast<uint32_t*>(packed), logits);
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
// Persistent-kernel path (precision 1): host-side program builder and scheduler
// =============================================================================================
namespace {

struct BgWork {
    int op, first_phase, last_phase, n_tiles, next_tile;
};

struct ProgramBuilder {
    rec::Program* p;
    int G, M;
    std::vector<std::vector<int>> crit;   // per phase: critical op indices
    std::vector<BgWork> bg;

    int tiles_of(const rec::Op& op) const {
        return ((M + 63) / 64) * ((op.N + op.bn - 1) / op.bn);
    }
    double cost_of(const rec::Op& op) const {
        return (double)op.K * (64 + op.bn) / (1024.0 * 96.0);
    }

    int add_op(const rec::Op& op) {
        p->ops[p->n_ops] = op;
        return p->n_ops++;
    }

    // Greedy list scheduler: critical tiles round-robin over the CTAs; background tiles fill the slack of a
    // phase (earliest deadline first) and whatever is left at an op's deadline is spread over the least-loaded CTAs.
    bool schedule() {
        const int n_ph = (int)crit.size();
        p->n_phases = n_ph;
        p->grid = G;
        int n_list = 0;
        std::vector<std::vector<uint32_t>> lists(G);
        std::vector<double> load(G);
        const double full = 1.0;
        for (int ph = 0; ph < n_ph; ++ph) {
            for (auto& l : lists) l.clear();
            std::fill(load.begin(), load.end(), 0.0);
            int rr = 0;
            for (int oi : crit[ph]) {
                const int nt = tiles_of(p->ops[oi]);
                const double c = cost_of(p->ops[oi]);
                for (int t = 0; t < nt; ++t) {
                    lists[rr].push_back(((uint32_t)oi << 20) | (uint32_t)t);
                    load[rr] += c;
                    rr = (rr + 1) % G;
                }
            }
            double lmax = *std::max_element(load.begin(), load.end());
            if (lmax < full) lmax = full;
            std::vector<BgWork*> ready;
            for (auto& b : bg)
                if (b.first_phase <= ph && ph <= b.last_phase && b.next_tile < b.n_tiles) ready.push_back(&b);
            std::sort(ready.begin(), ready.end(), [](BgWork* a, BgWork* b) { return a->last_phase < b->last_phase; });
            for (BgWork* b : ready) {
                const double c = cost_of(p->ops[b->op]);
                const bool deadline = (b->last_phase == ph);
                while (b->next_tile < b->n_tiles) {
                    const int cta = (int)(std::min_element(load.begin(), load.end()) - load.begin());
                    if (!deadline && load[cta] + c > lmax + 1e-9) break;
                    lists[cta].push_back(((uint32_t)b->op << 20) | (uint32_t)b->next_tile++);
                    load[cta] += c;
                }
            }
            for (int c = 0; c < G; ++c) {
                p->list_start[ph * G + c] = n_list;
                for (uint32_t e : lists[c]) {
                    if (n_list >= rec::MAX_TILES) return false;
                    p->tiles[n_list++] = e;
                }
            }
        }
        p->list_start[n_ph * G] = n_list;
        for (auto& b : bg)
            if (b.next_tile < b.n_tiles) return false;
        return true;
    }
};

rec::Op linear_op(const __nv_bfloat16* a_hi, const __nv_bfloat16* a_lo, int lda, const SplitW& w, const float* bias,
                  int act, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, int ldos) {
    rec::Op o;
    memset(&o, 0, sizeof(o));
    o.a_hi = a_hi; o.a_lo = a_lo; o.lda = lda;
    o.w_hi = w.hi; o.w_lo = w.lo; o.N = w.N; o.K = w.K;
    o.bias = bias; o.act = act;
    o.out_hi = out_hi; o.out_lo = out_lo; o.ldos = ldos;
    o.bn = 32;
    o.kind = rec::KIND_LINEAR;
    return o;
}

struct SplitBuf {
    __nv_bfloat16 *hi, *lo;
};
SplitBuf take_split(Workspace& ws, size_t n) {
    SplitBuf b;
    b.hi = reinterpret_cast<__nv_bfloat16*>(ws.take((n + 1) / 2));
    b.lo = reinterpret_cast<__nv_bfloat16*>(ws.take((n + 1) / 2));
    return b;
}

int run_program(BvrnnWeights& w, ProgramBuilder& pb, cudaStream_t s) {
    if (const char* e = getenv("BVC_REC_DEBUG")) pb.p->debug_flags = atoi(e);
    if (!pb.schedule()) {
        set_error("recurrent program does not fit the static limits (MAX_TILES)");
        return BVC_ERR_INVALID;
    }
    // the pinned staging copy is reused by the next call: wait until the previous upload has been consumed
    BVC_CUDA(cudaMemcpyAsync(w.rw.prog_dev, w.rw.prog_host, sizeof(rec::Program), cudaMemcpyHostToDevice, s));
    int rc = rec::umma_launch(w.rw.prog_dev, pb.G, w.rw.sync_words, s);
    if (rc) return rc;
    // prog_host is overwritten by the next call; the copy above must have been issued from a stable buffer
    BVC_CUDA(cudaStreamSynchronize(s));
    unsigned flags[2];
    BVC_CUDA(cudaMemcpy(flags, w.rw.sync_words, sizeof(flags), cudaMemcpyDeviceToHost));
    if (flags[1] != 0) {
        set_error(flags[1] == 2 ? "recurrent kernel aborted: tensor-core pipeline (mbarrier) timed out"
                                : "recurrent kernel aborted: device-wide barrier timed out");
        return BVC_ERR_DEVICE;
    }
    return BVC_OK;
}

}  // namespace

static int bvrnn_encode_persistent(BvrnnWeights& w, Workspace& ws, const float* mel, const float* bits,
                                   float bits_scalar, const float* h0, int B, int T, float* codes,
                                   unsigned long long* packed, float* logits, float* all_h, float* h_final,
                                   cudaStream_t s) {
    const int H = w.H, X = w.X, Z = w.Z;
    const size_t BT = (size_t)B * T;
    RecurrentWeights& rw = w.rw;
    int dev = 0, G = 0;
    BVC_CUDA(cudaGetDevice(&dev));
    BVC_TRY(rec::max_grid(dev, &G));
    if (G > rec::MAX_GRID) G = rec::MAX_GRID;

    float* yn = ws.take(BT * X);
    float* PA = ws.take(BT * H);
    float* PB = ws.take(BT * H);
    float* hf = ws.take((size_t)B * H);
    float* dh = ws.take((size_t)B * H);
    float* gh = ws.take((size_t)B * 3 * H);
    float* giz = ws.take((size_t)B * 3 * H);
    const size_t BH = (size_t)B * H;
    SplitBuf hS = take_split(ws, BH), e1S = take_split(ws, BH), e2S = take_split(ws, BH);
    SplitBuf zS = take_split(ws, (size_t)B * Z), z1S = take_split(ws, BH), z2S = take_split(ws, BH);
    SplitBuf pzS = take_split(ws, BH), d1S = take_split(ws, BH), d2S = take_split(ws, BH), d3S = take_split(ws, BH);
    SplitBuf mnS = take_split(ws, (size_t)B * 128), x1S = take_split(ws, BH), x2S = take_split(ws, BH);
    SplitBuf pxS = take_split(ws, BH);

    // hoisted over all frames (large GEMMs): phi_x(yn), then enc.0[:, :H] . phi_x
    {
        const size_t n = BT * X;
        normalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(mel, w.mean, w.std, yn, n, X);
        BVC_CHECK_LAUNCH();
    }
    BVC_TRY(run_linear(yn, X, (int)BT, w.px0, w.b_px0, H, PA, H, 1, s));
    BVC_TRY(run_linear(PA, H, (int)BT, w.px2, w.b_px2, H, PB, H, 1, s));
    BVC_TRY(run_linear(PB, H, (int)BT, w.px4, w.b_px4, H, PA, H, 1, s));
    BVC_TRY(run_linear(PA, H, (int)BT, w.e0x, nullptr, 0, PB, H, 1, s));
    float* E0x = PB;
    BVC_TRY(rec::init_state(h0, hf, hS.hi, hS.lo, (int)BH, mnS.hi, mnS.lo, B * 128, s));

    rec::Program* p = rw.prog_host;
    memset(p, 0, sizeof(rec::Program));
    rec::Frame& fr = p->frame;
    fr.M = B; fr.T = T; fr.X = X; fr.Z = Z; fr.H = H; fr.var_bit = w.var_bit;
    fr.bits_scalar = bits_scalar; fr.bits = bits; fr.codes = codes; fr.packed = packed; fr.logits = logits;
    fr.all_h = all_h; fr.h = hf; fr.gh = gh; fr.mean = w.mean; fr.std = w.std; fr.mel_out = nullptr;

    ProgramBuilder pb;
    pb.p = p; pb.G = G; pb.M = B;
    rec::Op o;
    o = linear_op(hS.hi, hS.lo, H, rw.e0h, rw.b_e0, 1, e1S.hi, e1S.lo, H);
    o.addend = E0x; o.ldadd = T * H; o.add_tstride = H;
    const int op_e1 = pb.add_op(o);
    o = linear_op(hS.hi, hS.lo, H, rw.d0h, nullptr, 0, nullptr, nullptr, 0);
    o.out_f = dh; o.ldo = H;
    const int op_dh = pb.add_op(o);
    o = linear_op(hS.hi, hS.lo, H, rw.whh_q, rw.b_hh_q, 0, nullptr, nullptr, 0);
    o.out_f = gh; o.ldo = 3 * H;
    const int op_gh = pb.add_op(o);
    const int op_e2 = pb.add_op(linear_op(e1S.hi, e1S.lo, H, rw.e2, w.b_e2, 1, e2S.hi, e2S.lo, H));
    o = linear_op(e2S.hi, e2S.lo, H, rw.e4, w.b_e4, 0, zS.hi, zS.lo, Z);
    o.kind = rec::KIND_BOTTLENECK;
    const int op_e4 = pb.add_op(o);
    const int op_z1 = pb.add_op(linear_op(zS.hi, nullptr, Z, rw.pz0, w.b_pz0, 1, z1S.hi, z1S.lo, H));
    const int op_z2 = pb.add_op(linear_op(z1S.hi, z1S.lo, H, rw.pz2, w.b_pz2, 1, z2S.hi, z2S.lo, H));
    const int op_pz = pb.add_op(linear_op(z2S.hi, z2S.lo, H, rw.pz4, w.b_pz4, 1, pzS.hi, pzS.lo, H));
    o = linear_op(pzS.hi, pzS.lo, H, rw.d0z, rw.b_d0, 1, d1S.hi, d1S.lo, H);
    o.addend = dh; o.ldadd = H;
    const int op_d1 = pb.add_op(o);
    o = linear_op(pzS.hi, pzS.lo, H, rw.ihz_q, rw.b_ih_q, 0, nullptr, nullptr, 0);
    o.out_f = giz; o.ldo = 3 * H;
    const int op_giz = pb.add_op(o);
    const int op_d2 = pb.add_op(linear_op(d1S.hi, d1S.lo, H, rw.d2, w.b_d2, 1, d2S.hi, d2S.lo, H));
    const int op_d3 = pb.add_op(linear_op(d2S.hi, d2S.lo, H, rw.d4, w.b_d4, 1, d3S.hi, d3S.lo, H));
    o = linear_op(d3S.hi, d3S.lo, H, rw.d6, w.b_d6, 0, mnS.hi, mnS.lo, 128);
    o.kind = rec::KIND_MEL;
    const int op_mel = pb.add_op(o);
    const int op_x1 = pb.add_op(linear_op(mnS.hi, mnS.lo, 128, rw.px0p, w.b_px0, 1, x1S.hi, x1S.lo, H));
    const int op_x2 = pb.add_op(linear_op(x1S.hi, x1S.lo, H, rw.px2, w.b_px2, 1, x2S.hi, x2S.lo, H));
    const int op_px = pb.add_op(linear_op(x2S.hi, x2S.lo, H, rw.px4, w.b_px4, 1, pxS.hi, pxS.lo, H));
    o = linear_op(pxS.hi, pxS.lo, H, rw.ihx_q, nullptr, 0, hS.hi, hS.lo, H);
    o.kind = rec::KIND_GRU; o.bn = 48;
    o.addend = giz; o.ldadd = 3 * H;
    const int op_gru = pb.add_op(o);

    pb.crit = {{op_e1}, {op_e2}, {op_e4}, {op_z1}, {op_z2}, {op_pz}, {op_d1}, {op_d2}, {op_d3},
               {op_mel}, {op_x1}, {op_x2}, {op_px}, {op_gru}};
    pb.bg.push_back({op_dh, 0, 5, pb.tiles_of(p->ops[op_dh]), 0});
    pb.bg.push_back({op_giz, 6, 12, pb.tiles_of(p->ops[op_giz]), 0});
    pb.bg.push_back({op_gh, 0, 12, pb.tiles_of(p->ops[op_gh]), 0});
    BVC_TRY(run_program(w, pb, s));
    if (h_final) BVC_CUDA(cudaMemcpyAsync(h_final, hf, sizeof(float) * BH, cudaMemcpyDeviceToDevice, s));
    return BVC_OK;
}

static int bvrnn_decode_persistent(BvrnnWeights& w, Workspace& ws, const float* codes, const float* h0, int B, int T,
                                   float* mel, float* h_final, cudaStream_t s) {
    const int H = w.H, X = w.X, Z = w.Z;
    const size_t BT = (size_t)B * T;
    RecurrentWeights& rw = w.rw;
    int dev = 0, G = 0;
    BVC_CUDA(cudaGetDevice(&dev));
    BVC_TRY(rec::max_grid(dev, &G));
    if (G > rec::MAX_GRID) G = rec::MAX_GRID;

    float* PA = ws.take(BT * H);
    float* PB = ws.take(BT * H);
    float* DZ = ws.take(BT * 4 * H);
    float* hf = ws.take((size_t)B * H);
    float* gh = ws.take((size_t)B * 3 * H);
    const size_t BH = (size_t)B * H;
    SplitBuf hS = take_split(ws, BH), d1S = take_split(ws, BH), d2S = take_split(ws, BH), d3S = take_split(ws, BH);
    SplitBuf mnS = take_split(ws, (size_t)B * 128), x1S = take_split(ws, BH), x2S = take_split(ws, BH);
    SplitBuf pxS = take_split(ws, BH);

    // hoisted over all frames: phi_z(z), then [dec.0_z ; W_ih_z (gate-interleaved)] . phi_z + [b_d0 ; b_ih]
    BVC_TRY(run_linear(codes, Z, (int)BT, w.pz0, w.b_pz0, H, PA, H, 1, s));
    BVC_TRY(run_linear(PA, H, (int)BT, w.pz2, w.b_pz2, H, PB, H, 1, s));
    BVC_TRY(run_linear(PB, H, (int)BT, w.pz4, w.b_pz4, H, PA, H, 1, s));
    BVC_TRY(run_linear(PA, H, (int)BT, rw.zcat_q, rw.b_zcat_q, 0, DZ, 4 * H, 1, s));
    BVC_TRY(rec::init_state(h0, hf, hS.hi, hS.lo, (int)BH, mnS.hi, mnS.lo, B * 128, s));

    rec::Program* p = rw.prog_host;
    memset(p, 0, sizeof(rec::Program));
    rec::Frame& fr = p->frame;
    fr.M = B; fr.T = T; fr.X = X; fr.Z = Z; fr.H = H; fr.var_bit = w.var_bit;
    fr.h = hf; fr.gh = gh; fr.mean = w.mean; fr.std = w.std; fr.mel_out = mel;

    ProgramBuilder pb;
    pb.p = p; pb.G = G; pb.M = B;
    rec::Op o;
    o = linear_op(hS.hi, hS.lo, H, rw.d0h, nullptr, 1, d1S.hi, d1S.lo, H);
    o.addend = DZ; o.ldadd = T * 4 * H; o.add_tstride = 4 * H;       // includes dec.0 bias
    const int op_d1 = pb.add_op(o);
    o = linear_op(hS.hi, hS.lo, H, rw.whh_q, rw.b_hh_q, 0, nullptr, nullptr, 0);
    o.out_f = gh; o.ldo = 3 * H;
    const int op_gh = pb.add_op(o);
    const int op_d2 = pb.add_op(linear_op(d1S.hi, d1S.lo, H, rw.d2, w.b_d2, 1, d2S.hi, d2S.lo, H));
    const int op_d3 = pb.add_op(linear_op(d2S.hi, d2S.lo, H, rw.d4, w.b_d4, 1, d3S.hi, d3S.lo, H));
    o = linear_op(d3S.hi, d3S.lo, H, rw.d6, w.b_d6, 0, mnS.hi, mnS.lo, 128);
    o.kind = rec::KIND_MEL;
    const int op_mel = pb.add_op(o);
    const int op_x1 = pb.add_op(linear_op(mnS.hi, mnS.lo, 128, rw.px0p, w.b_px0, 1, x1S.hi, x1S.lo, H));
    const int op_x2 = pb.add_op(linear_op(x1S.hi, x1S.lo, H, rw.px2, w.b_px2, 1, x2S.hi, x2S.lo, H));
    const int op_px = pb.add_op(linear_op(x2S.hi, x2S.lo, H, rw.px4, w.b_px4, 1, pxS.hi, pxS.lo, H));
    o = linear_op(pxS.hi, pxS.lo, H, rw.ihx_q, nullptr, 0, hS.hi, hS.lo, H);
    o.kind = rec::KIND_GRU; o.bn = 48;
    o.addend = DZ + H; o.ldadd = T * 4 * H; o.add_tstride = 4 * H;    // W_ih_z phi_z + b_ih, gate-interleaved
    const int op_gru = pb.add_op(o);

    pb.crit = {{op_d1}, {op_d2}, {op_d3}, {op_mel}, {op_x1}, {op_x2}, {op_px}, {op_gru}};
    pb.bg.push_back({op_gh, 0, 6, pb.tiles_of(p->ops[op_gh]), 0});
    BVC_TRY(run_program(w, pb, s));
    if (h_final) BVC_CUDA(cudaMemcpyAsync(h_final, hf, sizeof(float) * BH, cudaMemcpyDeviceToDevice, s));
    return BVC_OK;
}

int bvrnn_encode(BvrnnWeights& w, Workspace& ws, const float* mel, const float* bits, float bits_scalar,
                 const float* h0, int B, int T, float* codes, unsigned long long* packed, float* logits,
                 float* all_h, float* h_final, int precision, cudaStream_t s) {
    if (precision >= 1 && w.rw.ready && w.Z == 64 && w.X <= 128)
        return bvrnn_encode_persistent(w, ws, mel, bits, bits_scalar, h0, B, T, codes, packed, logits, all_h, h_final, s);
    return bvrnn_encode_layers(w, ws, mel, bits, bits_scalar, h0, B, T, codes, packed, logits, all_h, h_final,
                               precision, s);
}

int bvrnn_decode(BvrnnWeights& w, Workspace& ws, const float* codes, const float* h0, int B, int T, float* mel,
                 float* h_final, int precision, cudaStream_t s) {
    if (precision >= 1 && w.rw.ready && w.Z == 64 && w.X <= 128)
        return bvrnn_decode_persistent(w, ws, codes, h0, B, T, mel, h_final, s);
    return bvrnn_decode_layers(w, ws, codes, h0, B, T, mel, h_final, precision, s);
}

}  // namespace bvc
