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
#include "common.cuh"

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
    n += (size_t)B * (28 * H + 2 * w.X + 2 * w.Z) + 64 * 32;
    return n;
}

int bvrnn_encode(const BvrnnWeights& w, Workspace& ws, const float* mel, const float* bits, float bits_scalar,
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

int bvrnn_decode(const BvrnnWeights& w, Workspace& ws, const float* codes, const float* h0, int B, int T,
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

}  // namespace bvc
