// C ABI of libbvc.so (see include/bvc.h): handle, strict checkpoint loading and weight
// preparation (weight-norm fold, concatenations, split-bf16 packing), workspace, entry points.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "recurrent.cuh"

namespace bvc {

static thread_local std::string g_error;
thread_local int64_t* g_launch_counter = nullptr;
void set_error(const std::string& msg) { g_error = msg; }

int ws_check(const Workspace& ws, const char* where) {
    if (!ws.overflow) return BVC_OK;
    set_error(std::string(where) + ": workspace takes exceed the sized allocation (" + std::to_string(ws.bytes >> 20) +
              " MiB): sizing formula out of step with the buffers taken");
    return BVC_ERR_NOMEM;
}

}  // namespace bvc

using namespace bvc;

struct bvc_handle {
    bvc_config cfg;
    BvrnnWeights bw;
    VocoderWeights vw;
    FrontendTables ft;
    VocoderBuffers vb;
    bool have_bvrnn = false, have_voc = false, have_frontend = false;
    Workspace ws;
    int64_t launches = 0;
    int precision = 1;   // 1: split-bf16 tensor-core kernels (default, the measured path); 0: fp32 FFMA kernels
    std::vector<void*> allocs;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // device-to-host copies of the host entry points that overlap later kernels
    cudaEvent_t copy_ready = nullptr, copy_done = nullptr;
    cudaEvent_t ws_event = nullptr;   // end of the last job that used the workspace (any stream)
    bool ws_event_valid = false;
};

namespace {

struct HostTensor {
    const float* data;
    std::vector<int64_t> shape;
    size_t numel() const {
        size_t n = 1;
        for (auto s : shape) n *= (size_t)s;
        return n;
    }
};
typedef std::map<std::string, HostTensor> TensorMap;

struct Guard {   // selects the device and the launch counter for the duration of a call
    int prev = -1;
    bool ok = true;
    explicit Guard(bvc_handle* h) {
        cudaGetDevice(&prev);
        if (cudaSetDevice(h->cfg.device) != cudaSuccess) { ok = false; set_error("cudaSetDevice failed"); }
        g_launch_counter = &h->launches;
    }
    ~Guard() {
        g_launch_counter = nullptr;
        if (prev >= 0) cudaSetDevice(prev);
    }
};

template <typename T>
T* dev_upload(bvc_handle* h, const std::vector<T>& v) {
    T* p = nullptr;
    if (cudaMalloc(&p, v.size() * sizeof(T) + 16) != cudaSuccess) return nullptr;
    if (cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
    h->allocs.push_back(p);
    return p;
}

uint16_t f2bf(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
float bf2f(uint16_t b) {
    uint32_t u = (uint32_t)b << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

bool make_linear(bvc_handle* h, const std::vector<float>& w, int N, int K, LinearWeights* out) {
    out->N = N;
    out->K = K;
    out->w = dev_upload(h, w);
    if (!out->w) return false;
    if (K % 2 == 0) {
        std::vector<uint32_t> hi((size_t)N * K / 2), lo((size_t)N * K / 2);
        for (size_t i = 0; i < hi.size(); ++i) {
            const float a = w[2 * i], b = w[2 * i + 1];
            const uint16_t ah = f2bf(a), bh = f2bf(b);
            const uint16_t al = f2bf(a - bf2f(ah)), bl = f2bf(b - bf2f(bh));
            hi[i] = (uint32_t)ah | ((uint32_t)bh << 16);
            lo[i] = (uint32_t)al | ((uint32_t)bl << 16);
        }
        out->w_hi = dev_upload(h, hi);
        out->w_lo = dev_upload(h, lo);
        if (!out->w_hi || !out->w_lo) return false;
    }
    return true;
}

// Split-bf16 shared-memory images of W [N][K] for the persistent recurrent kernel (layout: recurrent.cuh).
// N is zero padded to a multiple of bn, K to a multiple of 64.
bool make_wimg(bvc_handle* h, const std::vector<float>& w, int N, int K, int bn, WImg* out) {
    const int n_tiles = (N + bn - 1) / bn, kch = (K + 63) / 64;
    const size_t chunk = (size_t)2 * bn * 128;
    std::vector<unsigned char> img((size_t)n_tiles * kch * chunk, 0);
    for (int nt = 0; nt < n_tiles; ++nt)
        for (int kc = 0; kc < kch; ++kc) {
            unsigned char* base = img.data() + ((size_t)nt * kch + kc) * chunk;
            for (int r = 0; r < bn; ++r) {
                const int n = nt * bn + r;
                if (n >= N) continue;
                for (int kk = 0; kk < 64; ++kk) {
                    const int k = kc * 64 + kk;
                    if (k >= K) continue;
                    const float v = w[(size_t)n * K + k];
                    const uint16_t vh = f2bf(v), vl = f2bf(v - bf2f(vh));
                    const size_t off = (size_t)r * 128 + ((((size_t)kk >> 3) ^ (size_t)(r & 7)) << 4) + (kk & 7) * 2;
                    memcpy(base + off, &vh, 2);
                    memcpy(base + (size_t)bn * 128 + off, &vl, 2);
                }
            }
        }
    out->img = dev_upload(h, img);
    out->N = N;
    out->K = K;
    out->bn = bn;
    return out->img != nullptr;
}
std::vector<float> pad_to(const std::vector<float>& v, size_t n) {
    std::vector<float> o(v);
    o.resize(n, 0.f);
    return o;
}
// GRU gate interleave: natural row gate*H + j  ->  3 grp (j / grp) + grp gate + (j % grp), so that a group of
// 3 grp columns holds r, z, n of grp hidden units (recurrent_cluster.cu, KIND_GRU epilogue; grp = 4)
std::vector<float> gate_interleave_rows(const std::vector<float>& w, int H, int K, int grp) {
    std::vector<float> out(w.size());
    for (int gate = 0; gate < 3; ++gate)
        for (int j = 0; j < H; ++j)
            memcpy(&out[(size_t)(3 * grp * (j / grp) + grp * gate + (j % grp)) * K], &w[(size_t)(gate * H + j) * K],
                   sizeof(float) * K);
    return out;
}

int collect(const bvc_tensor* tensors, int n, TensorMap* m) {
    for (int i = 0; i < n; ++i) {
        const bvc_tensor& t = tensors[i];
        if (!t.name || !t.data || t.ndim < 1 || t.ndim > 4) {
            set_error("malformed tensor descriptor at index " + std::to_string(i));
            return BVC_ERR_SCHEMA;
        }
        HostTensor ht;
        ht.data = t.data;
        ht.shape.assign(t.shape, t.shape + t.ndim);
        if (!m->emplace(t.name, ht).second) {
            set_error(std::string("duplicate tensor: ") + t.name);
            return BVC_ERR_SCHEMA;
        }
    }
    return BVC_OK;
}

// strict: every expected tensor present with the right shape, nothing extra
int check_schema(const TensorMap& m, const std::vector<std::pair<std::string, std::vector<int64_t>>>& expect) {
    for (auto& e : expect) {
        auto it = m.find(e.first);
        if (it == m.end()) {
            set_error("missing key in state_dict: " + e.first);
            return BVC_ERR_SCHEMA;
        }
        if (it->second.shape != e.second) {
            std::string s = "size mismatch for " + e.first + ": got [";
            for (auto d : it->second.shape) s += std::to_string(d) + ",";
            s += "] expected [";
            for (auto d : e.second) s += std::to_string(d) + ",";
            set_error(s + "]");
            return BVC_ERR_SCHEMA;
        }
    }
    if (m.size() != expect.size()) {
        for (auto& kv : m) {
            bool found = false;
            for (auto& e : expect) found = found || e.first == kv.first;
            if (!found) {
                set_error("unexpected key in state_dict: " + kv.first);
                return BVC_ERR_SCHEMA;
            }
        }
    }
    return BVC_OK;
}

std::vector<float> to_vec(const HostTensor& t) { return std::vector<float>(t.data, t.data + t.numel()); }

// rows [r0, r1) x cols [c0, c1) of a row-major [R, C] matrix
std::vector<float> slice(const HostTensor& t, int64_t r0, int64_t r1, int64_t c0, int64_t c1) {
    const int64_t C = t.shape[1];
    std::vector<float> out((size_t)(r1 - r0) * (c1 - c0));
    for (int64_t r = r0; r < r1; ++r)
        memcpy(&out[(size_t)(r - r0) * (c1 - c0)], t.data + r * C + c0, sizeof(float) * (c1 - c0));
    return out;
}
void append(std::vector<float>& a, const std::vector<float>& b) { a.insert(a.end(), b.begin(), b.end()); }

// w = v * (g / ||v||), norm over all dims but 0 (old-style torch weight_norm, dim=0)
std::vector<float> fold_weight_norm(const HostTensor& g, const HostTensor& v) {
    const int64_t d0 = v.shape[0];
    const size_t inner = v.numel() / (size_t)d0;
    std::vector<float> w(v.numel());
    for (int64_t i = 0; i < d0; ++i) {
        double ss = 0.0;
        for (size_t j = 0; j < inner; ++j) ss += (double)v.data[i * inner + j] * v.data[i * inner + j];
        const float scale = g.data[i] / (float)sqrt(ss);
        for (size_t j = 0; j < inner; ++j) w[i * inner + j] = v.data[i * inner + j] * scale;
    }
    return w;
}

// Split-bf16 weights in mma.sync.m16n8k16 B-fragment order.  B[kk][co] with kk = tap * CIN + ci, zero padded
// to a multiple of 16 rows; entry ((kc * NT + nt) * 32 + lane) holds (b0, b1) of that lane.
template <typename Get>
void pack_fragments(int ntaps, int CIN, int COUT, Get get, std::vector<uint2>* hi, std::vector<uint2>* lo) {
    const int KT = ntaps * CIN, KC = (KT + 15) / 16, NT = COUT / 8;
    hi->assign((size_t)KC * NT * 32, make_uint2(0, 0));
    lo->assign((size_t)KC * NT * 32, make_uint2(0, 0));
    for (int kc = 0; kc < KC; ++kc)
        for (int nt = 0; nt < NT; ++nt)
            for (int lane = 0; lane < 32; ++lane) {
                const int g = lane >> 2, q = lane & 3, n = nt * 8 + g;
                uint32_t h2[2], l2[2];
                for (int reg = 0; reg < 2; ++reg) {
                    uint16_t hh[2], ll[2];
                    for (int e = 0; e < 2; ++e) {
                        const int kk = 16 * kc + 2 * q + 8 * reg + e;
                        const float v = kk < KT ? get(kk / CIN, kk % CIN, n) : 0.f;
                        hh[e] = f2bf(v);
                        ll[e] = f2bf(v - bf2f(hh[e]));
                    }
                    h2[reg] = (uint32_t)hh[0] | ((uint32_t)hh[1] << 16);
                    l2[reg] = (uint32_t)ll[0] | ((uint32_t)ll[1] << 16);
                }
                (*hi)[((size_t)kc * NT + nt) * 32 + lane] = make_uint2(h2[0], h2[1]);
                (*lo)[((size_t)kc * NT + nt) * 32 + lane] = make_uint2(l2[0], l2[1]);
            }
}

int ensure_workspace(bvc_handle* h, size_t floats) {
    const size_t bytes = floats * sizeof(float) + 4096;
    if (h->ws.bytes < bytes) {
        if (h->ws.base) {
            cudaDeviceSynchronize();
            cudaFree(h->ws.base);
            h->ws.base = nullptr;
            h->ws.bytes = 0;
        }
        void* p = nullptr;
        if (cudaMalloc(&p, bytes) != cudaSuccess) {
            cudaGetLastError();
            set_error("workspace allocation of " + std::to_string(bytes >> 20) + " MiB failed");
            return BVC_ERR_NOMEM;
        }
        h->ws.base = (float*)p;
        h->ws.bytes = bytes;
    }
    h->ws.used = 0;
    h->ws.overflow = false;
    return BVC_OK;
}

// The workspace is shared by every entry point.  Work is stream-ordered, so a job on stream s must
// not start before the previous job (possibly on another stream) has finished with the workspace.
int ws_acquire(bvc_handle* h, cudaStream_t s) {
    if (h->ws_event_valid) BVC_CUDA(cudaStreamWaitEvent(s, h->ws_event, 0));
    return BVC_OK;
}
int ws_release(bvc_handle* h, cudaStream_t s) {
    BVC_CUDA(cudaEventRecord(h->ws_event, s));
    h->ws_event_valid = true;
    return BVC_OK;
}

#define REQUIRE(cond, code, msg)      \
    do {                              \
        if (!(cond)) {                \
            set_error(msg);           \
            return code;              \
        }                             \
    } while (0)

}  // namespace

extern "C" {

int bvc_abi_version(void) { return BVC_ABI_VERSION; }
const char* bvc_last_error(void) { return g_error.c_str(); }

int bvc_create(bvc_handle** out, const bvc_config* cfg) {
    REQUIRE(out && cfg, BVC_ERR_INVALID, "bvc_create: null argument");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("no CUDA device: libbvc has no CPU path");
        return BVC_ERR_DEVICE;
    }
    REQUIRE(cfg->device >= 0 && cfg->device < ndev, BVC_ERR_DEVICE, "bvc_create: bad device ordinal");
    cudaDeviceProp prop;
    BVC_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) {
        set_error(std::string("device ") + prop.name + " is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                  "; libbvc is built for sm_100a only");
        return BVC_ERR_DEVICE;
    }
    REQUIRE(cfg->n_fft == 1024, BVC_ERR_INVALID, "n_fft must be 1024 in this build");
    REQUIRE(cfg->hop > 0 && cfg->pad_left >= 0 && cfg->n_fft - cfg->pad_left - cfg->hop >= 0, BVC_ERR_INVALID,
            "bad hop / pad_left");
    REQUIRE(cfg->x_dim % 16 == 0 && cfg->h_dim % 32 == 0 && cfg->z_dim % 16 == 0, BVC_ERR_INVALID,
            "x_dim, z_dim must be multiples of 16 and h_dim of 32");
    REQUIRE(cfg->voc_num_stages == 4 && cfg->voc_num_kernels == 3 && cfg->voc_initial_channel == 128, BVC_ERR_INVALID,
            "vocoder: this build covers the shipped 4-stage / 3-kernel / 128-channel configuration");
    static const int up_ok[4] = {8, 8, 2, 2};
    for (int i = 0; i < 4; ++i)
        REQUIRE(cfg->voc_up_rates[i] == up_ok[i] && cfg->voc_up_kernels[i] == 2 * up_ok[i], BVC_ERR_INVALID,
                "vocoder: upsample rates must be [8,8,2,2] with kernels [16,16,4,4]");
    bvc_handle* h = new bvc_handle();
    h->cfg = *cfg;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(cfg->device);
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ws_event, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->copy_ready, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->copy_done, cudaEventDisableTiming);
    cudaSetDevice(prev);
    if (e != cudaSuccess) {
        delete h;
        set_error("cudaStreamCreate failed");
        return BVC_ERR_DEVICE;
    }
    *out = h;
    return BVC_OK;
}

int bvc_destroy(bvc_handle* h) {
    if (!h) return BVC_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(h->cfg.device);
    cudaDeviceSynchronize();
    for (void* p : h->allocs) cudaFree(p);
    if (h->ws.base) cudaFree(h->ws.base);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->copy_ready) cudaEventDestroy(h->copy_ready);
    if (h->copy_done) cudaEventDestroy(h->copy_done);
    if (h->ws_event) cudaEventDestroy(h->ws_event);
    rec_free_slots(h->bw.rw);
    cudaSetDevice(prev);
    delete h;
    return BVC_OK;
}

int bvc_set_precision(bvc_handle* h, int32_t mode) {
    REQUIRE(h && (mode == 0 || mode == 1), BVC_ERR_INVALID, "precision mode must be 0 or 1");
    h->precision = mode;
    return BVC_OK;
}

int bvc_load_bvrnn(bvc_handle* h, const bvc_tensor* tensors, int32_t n) {
    REQUIRE(h && tensors, BVC_ERR_INVALID, "bvc_load_bvrnn: null argument");
    Guard g(h);
    if (!g.ok) return BVC_ERR_DEVICE;
    TensorMap m;
    int rc = collect(tensors, n, &m);
    if (rc) return rc;
    const int64_t X = h->cfg.x_dim, H = h->cfg.h_dim, Z = h->cfg.z_dim;
    std::vector<std::pair<std::string, std::vector<int64_t>>> ex = {
        {"mean_mel", {X}}, {"std_mel", {X}}, {"log_sigma", {1}},
        {"rnn.weight_ih_l0", {3 * H, 2 * H}}, {"rnn.weight_hh_l0", {3 * H, H}},
        {"rnn.bias_ih_l0", {3 * H}}, {"rnn.bias_hh_l0", {3 * H}}};
    auto lin = [&](const std::string& name, int64_t o, int64_t i) {
        ex.push_back({name + ".weight", {o, i}});
        ex.push_back({name + ".bias", {o}});
    };
    lin("phi_x.0", H, X); lin("phi_x.2", H, H); lin("phi_x.4", H, H);
    lin("phi_z.0", H, Z); lin("phi_z.2", H, H); lin("phi_z.4", H, H);
    lin("enc.0", H, 2 * H); lin("enc.2", H, H); lin("enc.4", Z, H);
    lin("prior.0", H, H); lin("prior.2", H, H); lin("prior.4", Z, H);
    lin("dec.0", H, 2 * H); lin("dec.2", H, H); lin("dec.4", H, H); lin("dec.6", X, H);
    rc = check_schema(m, ex);
    if (rc) return rc;

    BvrnnWeights& w = h->bw;
    w.X = (int)X; w.H = (int)H; w.Z = (int)Z; w.var_bit = h->cfg.var_bit;
    bool ok = true;
    auto up = [&](const std::vector<float>& v) { float* p = dev_upload(h, v); ok = ok && p; return p; };
    auto L = [&](const std::string& name, LinearWeights* lw, float** bias) {
        const HostTensor& t = m[name + ".weight"];
        ok = ok && make_linear(h, to_vec(t), (int)t.shape[0], (int)t.shape[1], lw);
        *bias = up(to_vec(m[name + ".bias"]));
    };
    w.mean = up(to_vec(m["mean_mel"]));
    w.std = up(to_vec(m["std_mel"]));
    L("phi_x.0", &w.px0, &w.b_px0); L("phi_x.2", &w.px2, &w.b_px2); L("phi_x.4", &w.px4, &w.b_px4);
    L("phi_z.0", &w.pz0, &w.b_pz0); L("phi_z.2", &w.pz2, &w.b_pz2); L("phi_z.4", &w.pz4, &w.b_pz4);
    L("enc.2", &w.e2, &w.b_e2); L("enc.4", &w.e4, &w.b_e4);
    L("dec.2", &w.d2, &w.b_d2); L("dec.4", &w.d4, &w.b_d4); L("dec.6", &w.d6, &w.b_d6);
    L("prior.0", &w.pr0, &w.b_pr0); L("prior.2", &w.pr2, &w.b_pr2); L("prior.4", &w.pr4, &w.b_pr4);

    const HostTensor &e0 = m["enc.0.weight"], &d0 = m["dec.0.weight"];
    const HostTensor &wih = m["rnn.weight_ih_l0"], &whh = m["rnn.weight_hh_l0"];
    const std::vector<float> zerosH((size_t)H, 0.f);
    ok = ok && make_linear(h, slice(e0, 0, H, 0, H), (int)H, (int)H, &w.e0x);
    {   // [enc.0[:, H:]; dec.0[:, H:]; W_hh]   (concat order of bvrnn.py:189,202: [phi ; h])
        std::vector<float> c = slice(e0, 0, H, H, 2 * H);
        append(c, slice(d0, 0, H, H, 2 * H));
        append(c, to_vec(whh));
        ok = ok && make_linear(h, c, (int)(5 * H), (int)H, &w.hcat_enc);
        std::vector<float> b = to_vec(m["enc.0.bias"]);
        append(b, zerosH);
        append(b, to_vec(m["rnn.bias_hh_l0"]));
        w.b_hcat_enc = up(b);
    }
    {
        std::vector<float> c = slice(d0, 0, H, H, 2 * H);
        append(c, to_vec(whh));
        ok = ok && make_linear(h, c, (int)(4 * H), (int)H, &w.hcat_dec);
        std::vector<float> b = zerosH;
        append(b, to_vec(m["rnn.bias_hh_l0"]));
        w.b_hcat_dec = up(b);
    }
    {   // [dec.0[:, :H]; W_ih[:, H:]]   (GRU input is [phi_x_gen ; phi_z], bvrnn.py:206)
        std::vector<float> c = slice(d0, 0, H, 0, H);
        append(c, slice(wih, 0, 3 * H, H, 2 * H));
        ok = ok && make_linear(h, c, (int)(4 * H), (int)H, &w.zcat);
        std::vector<float> b = to_vec(m["dec.0.bias"]);
        append(b, to_vec(m["rnn.bias_ih_l0"]));
        w.b_zcat = up(b);
    }
    ok = ok && make_linear(h, slice(wih, 0, 3 * H, 0, H), (int)(3 * H), (int)H, &w.ihx);
    {   // persistent recurrent kernel operands
        RecurrentWeights& rw = w.rw;
        const int Hi = (int)H, Xi = (int)X, Zi = (int)Z, G = 4;
        auto W = [&](const std::vector<float>& v, int N, int K, int bn, WImg* o) { ok = ok && make_wimg(h, v, N, K, bn, o); };
        W(slice(e0, 0, H, H, 2 * H), Hi, Hi, 64, &rw.e0h);
        W(slice(d0, 0, H, H, 2 * H), Hi, Hi, 64, &rw.d0h);
        W(slice(d0, 0, H, 0, H), Hi, Hi, 64, &rw.d0z);
        W(to_vec(m["enc.2.weight"]), Hi, Hi, 64, &rw.e2);
        W(to_vec(m["enc.4.weight"]), Zi, Hi, 64, &rw.e4);
        W(to_vec(m["phi_z.0.weight"]), Hi, Zi, 16, &rw.pz0);
        W(to_vec(m["phi_z.2.weight"]), Hi, Hi, 64, &rw.pz2);
        W(to_vec(m["phi_z.4.weight"]), Hi, Hi, 64, &rw.pz4);
        W(to_vec(m["dec.2.weight"]), Hi, Hi, 64, &rw.d2);
        W(to_vec(m["dec.4.weight"]), Hi, Hi, 64, &rw.d4);
        W(to_vec(m["dec.6.weight"]), Xi, Hi, 64, &rw.d6);
        W(to_vec(m["phi_x.2.weight"]), Hi, Hi, 64, &rw.px2);
        W(to_vec(m["phi_x.4.weight"]), Hi, Hi, 64, &rw.px4);
        rw.b_e0 = up(to_vec(m["enc.0.bias"]));
        rw.b_d0 = up(to_vec(m["dec.0.bias"]));
        rw.b_d6p = up(pad_to(to_vec(m["dec.6.bias"]), (size_t)((Xi + 63) / 64) * 64));
        {   // x1f = phi_x.0 . diag(1/std) . dec.6 ; bias = phi_x.0 . ((b_dec6 - mean) / std) + b_phi_x0   (float64)
            const float* px0 = m["phi_x.0.weight"].data;      // [H, X]
            const float* d6 = m["dec.6.weight"].data;          // [X, H]
            const float* b6 = m["dec.6.bias"].data;
            const float* bp = m["phi_x.0.bias"].data;
            const float* mean = m["mean_mel"].data;
            const float* sd = m["std_mel"].data;
            std::vector<float> wf((size_t)H * H), bf((size_t)H);
            std::vector<double> row((size_t)H);
            for (int64_t i = 0; i < H; ++i) {
                std::fill(row.begin(), row.end(), 0.0);
                double bacc = bp[i];
                for (int64_t j = 0; j < X; ++j) {
                    const double c = (double)px0[i * X + j] / (double)sd[j];
                    bacc += c * ((double)b6[j] - (double)mean[j]);
                    const float* d6r = d6 + j * H;
                    for (int64_t k = 0; k < H; ++k) row[k] += c * (double)d6r[k];
                }
                for (int64_t k = 0; k < H; ++k) wf[(size_t)i * H + k] = (float)row[k];
                bf[i] = (float)bacc;
            }
            W(wf, Hi, Hi, 64, &rw.x1f);
            rw.b_x1f = up(bf);
        }
        {   // GRU rows gate-interleaved in groups of 12 = [r(4) z(4) n(4)]
            W(gate_interleave_rows(to_vec(whh), Hi, Hi, G), 3 * Hi, Hi, 64, &rw.whh_q);
            const std::vector<float> ihz_q = gate_interleave_rows(slice(wih, 0, 3 * H, H, 2 * H), Hi, Hi, G);
            W(ihz_q, 3 * Hi, Hi, 64, &rw.ihz_q);
            W(gate_interleave_rows(slice(wih, 0, 3 * H, 0, H), Hi, Hi, G), 3 * Hi, Hi, 48, &rw.ihx_q);
            const std::vector<float> bih_q = gate_interleave_rows(to_vec(m["rnn.bias_ih_l0"]), Hi, 1, G);
            rw.b_hh_q = up(gate_interleave_rows(to_vec(m["rnn.bias_hh_l0"]), Hi, 1, G));
            rw.b_ih_q = up(bih_q);
            std::vector<float> zq = slice(d0, 0, H, 0, H);
            append(zq, ihz_q);
            W(zq, 4 * Hi, Hi, 256, &rw.g_zcat);
            std::vector<float> bq = to_vec(m["dec.0.bias"]);
            append(bq, bih_q);
            rw.b_zcat_q = up(bq);
        }
        // hoisted layers over all B*T frames
        W(to_vec(m["phi_x.0.weight"]), Hi, Xi, 256, &rw.g_px0);
        W(to_vec(m["phi_x.2.weight"]), Hi, Hi, 256, &rw.g_px2);
        W(to_vec(m["phi_x.4.weight"]), Hi, Hi, 256, &rw.g_px4);
        W(slice(e0, 0, H, 0, H), Hi, Hi, 256, &rw.g_e0x);
        W(to_vec(m["phi_z.0.weight"]), Hi, Zi, 256, &rw.g_pz0);
        W(to_vec(m["phi_z.2.weight"]), Hi, Hi, 256, &rw.g_pz2);
        W(to_vec(m["phi_z.4.weight"]), Hi, Hi, 256, &rw.g_pz4);
        W(to_vec(m["prior.0.weight"]), Hi, Hi, 256, &rw.g_pr0);
        W(to_vec(m["prior.2.weight"]), Hi, Hi, 256, &rw.g_pr2);
        W(to_vec(m["prior.4.weight"]), Zi, Hi, 256, &rw.g_pr4);
        rw.b_pr4p = up(pad_to(to_vec(m["prior.4.bias"]), 256));
        void* p1 = nullptr; void* p2 = nullptr;
        ok = ok && cudaMalloc(&p1, rec::SYNC_WORDS * sizeof(unsigned)) == cudaSuccess &&
             cudaMalloc(&p2, sizeof(rec::Program)) == cudaSuccess;
        if (p1) h->allocs.push_back(p1);
        if (p2) h->allocs.push_back(p2);
        rw.sync_words = (unsigned*)p1;
        rw.prog_dev = (rec::Program*)p2;
        rw.ready = ok && (H % 256 == 0) && Z == 64 && X <= 128;
    }
    REQUIRE(ok, BVC_ERR_NOMEM, "device allocation failed while loading BVRNN weights");
    h->have_bvrnn = true;
    return BVC_OK;
}

int bvc_load_vocoder(bvc_handle* h, const bvc_tensor* tensors, int32_t n) {
    REQUIRE(h && tensors, BVC_ERR_INVALID, "bvc_load_vocoder: null argument");
    Guard g(h);
    if (!g.ok) return BVC_ERR_DEVICE;
    TensorMap m;
    int rc = collect(tensors, n, &m);
    if (rc) return rc;
    const bvc_config& c = h->cfg;
    const int64_t X = c.x_dim, C0 = c.voc_initial_channel;
    std::vector<std::pair<std::string, std::vector<int64_t>>> ex;
    auto wn = [&](const std::string& name, int64_t d0, int64_t d1, int64_t k, int64_t nb) {
        ex.push_back({name + ".weight_g", {d0, 1, 1}});
        ex.push_back({name + ".weight_v", {d0, d1, k}});
        ex.push_back({name + ".bias", {nb}});
    };
    // SnakeBeta, or Activation1d wrapping it (models.py:66-88): `.act.alpha/.act.beta` + two registered filter buffers
    std::string aa_first;
    auto act_schema = [&](const std::string& name, int64_t C, bool aa) {
        const std::string inner = aa ? name + ".act" : name;
        ex.push_back({inner + ".alpha", {C}});
        ex.push_back({inner + ".beta", {C}});
        if (aa) {
            ex.push_back({name + ".upsample.filter", {1, 1, 12}});
            ex.push_back({name + ".downsample.lowpass.filter", {1, 1, 12}});
            if (aa_first.empty()) aa_first = name;
        }
    };
    wn("conv_pre", C0, X, 7, C0);
    int64_t ch = C0;
    for (int i = 0; i < c.voc_num_stages; ++i) {
        const int64_t cin = C0 >> i, cout = C0 >> (i + 1);
        wn("ups." + std::to_string(i) + ".1", cin, cout, c.voc_up_kernels[i], cout);
        ch = cout;
        for (int j = 0; j < c.voc_num_kernels; ++j) {
            const std::string rb = "resblocks." + std::to_string(i * c.voc_num_kernels + j);
            for (int l = 0; l < 3; ++l) {
                wn(rb + ".convs1." + std::to_string(l), ch, ch, c.voc_res_kernels[j], ch);
                wn(rb + ".convs2." + std::to_string(l), ch, ch, c.voc_res_kernels[j], ch);
            }
            for (int a = 0; a < 6; ++a) act_schema(rb + ".activations." + std::to_string(a), ch, c.voc_antialias[i] != 0);
        }
    }
    act_schema("activation_post", ch, c.voc_antialias_post != 0);
    wn("conv_post", 1, ch, 7, 1);
    rc = check_schema(m, ex);
    if (rc) return rc;

    VocoderWeights& w = h->vw;
    w.n_mels = (int)X; w.c0 = (int)C0; w.n_stages = c.voc_num_stages; w.n_kernels = c.voc_num_kernels;
    for (int i = 0; i < 4; ++i) w.rates[i] = c.voc_up_rates[i];
    for (int i = 0; i < 3; ++i) { w.dil[i] = c.voc_res_dilations[i]; w.rks[i] = c.voc_res_kernels[i]; }
    bool ok = true;
    auto up = [&](const std::vector<float>& v) { float* p = dev_upload(h, v); ok = ok && p; return p; };
    auto folded = [&](const std::string& name) { return fold_weight_norm(m[name + ".weight_g"], m[name + ".weight_v"]); };
    auto snake = [&](const std::string& name, int64_t C, SnakeParams* sp) {
        const HostTensor &al = m[name + ".alpha"], &be = m[name + ".beta"];
        std::vector<float> ea((size_t)C), ieb((size_t)C);
        for (int64_t i = 0; i < C; ++i) {
            ea[i] = expf(al.data[i]);
            ieb[i] = 1.0f / (expf(be.data[i]) + 1e-9f);
        }
        sp->ea = up(ea);
        sp->inv_eb = up(ieb);
    };
    {   // conv_pre [co, ci, 7] -> Linear over 7 contiguous channel-last frames: [co, j*X + ci]
        const std::vector<float> f = folded("conv_pre");
        std::vector<float> p((size_t)C0 * 7 * X);
        for (int64_t co = 0; co < C0; ++co)
            for (int64_t ci = 0; ci < X; ++ci)
                for (int64_t j = 0; j < 7; ++j) p[(co * 7 + j) * X + ci] = f[(co * X + ci) * 7 + j];
        ok = ok && make_linear(h, p, (int)C0, (int)(7 * X), &w.pre);
        w.b_pre = up(to_vec(m["conv_pre.bias"]));
    }
    ch = C0;
    for (int i = 0; i < w.n_stages; ++i) {
        const int64_t cin = C0 >> i, cout = C0 >> (i + 1), k = c.voc_up_kernels[i];
        const std::string un = "ups." + std::to_string(i) + ".1";
        {   // ConvTranspose1d weight [ci, co, tap] -> [tap][ci][co]
            const std::vector<float> f = folded(un);
            std::vector<float> p(f.size());
            for (int64_t ci = 0; ci < cin; ++ci)
                for (int64_t co = 0; co < cout; ++co)
                    for (int64_t t = 0; t < k; ++t) p[(t * cin + ci) * cout + co] = f[(ci * cout + co) * k + t];
            w.w_up[i] = up(p);
            w.b_up[i] = up(to_vec(m[un + ".bias"]));
            // tensor-core copy: output phase r is a 2-tap conv over (x[j-1], x[j]) with taps (W[.,.,r+U], W[.,.,r])
            const int64_t U = c.voc_up_rates[i];
            std::vector<uint2> allh, alll;
            for (int64_t r = 0; r < U; ++r) {
                std::vector<uint2> fh, fl;
                pack_fragments(2, (int)cin, (int)cout,
                               [&](int tap, int ci, int co) { return f[((size_t)ci * cout + co) * k + (tap == 0 ? r + U : r)]; },
                               &fh, &fl);
                allh.insert(allh.end(), fh.begin(), fh.end());
                alll.insert(alll.end(), fl.begin(), fl.end());
            }
            w.upf_h[i] = dev_upload(h, allh);
            w.upf_l[i] = dev_upload(h, alll);
            ok = ok && w.upf_h[i] && w.upf_l[i];
        }
        ch = cout;
        for (int j = 0; j < w.n_kernels; ++j) {
            AmpBlockWeights& bw = w.blocks[i * w.n_kernels + j];
            const std::string rb = "resblocks." + std::to_string(i * w.n_kernels + j);
            const int64_t kk = c.voc_res_kernels[j];
            bw.k = (int)kk;
            auto pack = [&](const std::string& name, uint2** fh_out, uint2** fl_out) {   // [co, ci, tap] -> [ci][tap][co]
                const std::vector<float> f = folded(name);
                std::vector<float> p(f.size());
                for (int64_t co = 0; co < ch; ++co)
                    for (int64_t ci = 0; ci < ch; ++ci)
                        for (int64_t t = 0; t < kk; ++t) p[(ci * kk + t) * ch + co] = f[(co * ch + ci) * kk + t];
                std::vector<uint2> fh, fl;
                pack_fragments((int)kk, (int)ch, (int)ch,
                               [&](int tap, int ci, int co) { return f[((size_t)co * ch + ci) * kk + tap]; }, &fh, &fl);
                *fh_out = dev_upload(h, fh);
                *fl_out = dev_upload(h, fl);
                ok = ok && *fh_out && *fl_out;
                return up(p);
            };
            for (int l = 0; l < 3; ++l) {
                bw.w1[l] = pack(rb + ".convs1." + std::to_string(l), &bw.f1h[l], &bw.f1l[l]);
                bw.b1[l] = up(to_vec(m[rb + ".convs1." + std::to_string(l) + ".bias"]));
                bw.w2[l] = pack(rb + ".convs2." + std::to_string(l), &bw.f2h[l], &bw.f2l[l]);
                bw.b2[l] = up(to_vec(m[rb + ".convs2." + std::to_string(l) + ".bias"]));
            }
            for (int a = 0; a < 6; ++a)
                snake(rb + ".activations." + std::to_string(a) + (c.voc_antialias[i] ? ".act" : ""), ch, &bw.act[a]);
        }
    }
    // ---- tcgen05 stage kernel (vocoder.cu, stage_umma_kernel): weight stream in MMA issue order ----
    for (int i = 0; i < w.n_stages; ++i) {
        UmmaStageWeights& um = w.umma[i];
        um.ready = false;
        const int Cs = (int)(C0 >> (i + 1));
        const bool cfg_ok = w.n_kernels == 3 && c.voc_res_kernels[0] == 3 && c.voc_res_kernels[1] == 7 && c.voc_res_kernels[2] == 11 &&
                            c.voc_res_dilations[0] == 1 && c.voc_res_dilations[1] == 3 && c.voc_res_dilations[2] == 5;
        if (!cfg_ok || Cs > 64 || Cs < 8) continue;
        const int N = Cs < 16 ? 16 : Cs;
        std::vector<unsigned char> stream;
        int n_jobs = 0;
        for (int l = 0; l < 3; ++l)
            for (int conv2 = 0; conv2 < 2; ++conv2)
                for (int cc = 0; cc < 3; ++cc) {
                    const int j = 2 - cc;                       // chain 0 = k 11, 1 = k 7, 2 = k 3
                    const int K = c.voc_res_kernels[j];
                    const std::string rb = "resblocks." + std::to_string(i * 3 + j);
                    const std::vector<float> f = folded(rb + (conv2 ? ".convs2." : ".convs1.") + std::to_string(l));   // [co, ci, tap]
                    const int steps = Cs >= 16 ? K * Cs / 16 : (K + 1) / 2;
                    UmmaJob& jb = um.jobs[n_jobs++];
                    jb.chain = (short)cc; jb.layer = (short)l; jb.conv2 = (short)conv2; jb.K = (short)K;
                    jb.d = (short)(conv2 ? 1 : c.voc_res_dilations[l]);
                    jb.steps = (short)steps; jb.off = (int)stream.size();
                    {
                        for (int st = 0; st < steps; ++st) {
                            std::vector<uint16_t> blk((size_t)N * 32, 0);   // [k group][hi | lo][N][8]
                            for (int gq = 0; gq < 2; ++gq)
                                for (int n = 0; n < Cs; ++n)
                                    for (int e = 0; e < 8; ++e) {
                                        int tap, ci;
                                        if (Cs >= 16) { const int spt = Cs / 16; tap = st / spt; ci = (st % spt) * 16 + gq * 8 + e; }
                                        else { tap = 2 * st + gq; ci = e; }
                                        const float v = tap < K ? f[((size_t)n * Cs + ci) * K + tap] : 0.f;
                                        const uint16_t vh = f2bf(v), vl = f2bf(v - bf2f(vh));
                                        blk[((size_t)(gq * 2 + 0) * N + n) * 8 + e] = vh;
                                        blk[((size_t)(gq * 2 + 1) * N + n) * 8 + e] = vl;
                                    }
                            const unsigned char* bp = reinterpret_cast<const unsigned char*>(blk.data());
                            stream.insert(stream.end(), bp, bp + blk.size() * 2);
                        }
                    }
                }
        um.wstream = dev_upload(h, stream);
        ok = ok && um.wstream;
        for (int cc = 0; cc < 3; ++cc) {
            const AmpBlockWeights& bw = w.blocks[i * 3 + (2 - cc)];
            const std::string rb = "resblocks." + std::to_string(i * 3 + (2 - cc));
            std::vector<float> acc((size_t)Cs, 0.f);
            for (int l = 0; l < 3; ++l) {
                um.b1[cc][l] = bw.b1[l];
                const std::vector<float> b2 = to_vec(m[rb + ".convs2." + std::to_string(l) + ".bias"]);
                for (int q = 0; q < Cs; ++q) acc[q] += b2[q];
                um.bsum[cc][l] = up(acc);
            }
        }
        um.ready = ok;
    }
    snake(c.voc_antialias_post ? "activation_post.act" : "activation_post", ch, &w.act_post);
    for (int i = 0; i < 4; ++i) w.antialias[i] = i < c.voc_num_stages && c.voc_antialias[i] != 0;
    w.antialias_post = c.voc_antialias_post != 0;
    if (!aa_first.empty()) {
        // every Activation1d of the reference is built with the same filter (act.py:10-21); the buffers of the first one
        // are used for all, after checking that the others carry the same values
        const HostTensor &fu = m[aa_first + ".upsample.filter"], &fd = m[aa_first + ".downsample.lowpass.filter"];
        for (int q = 0; q < 12; ++q) { w.aa_up[q] = fu.data[q]; w.aa_down[q] = fd.data[q]; }
        for (const auto& kv : m) {
            const std::string& nm = kv.first;
            const bool is_up = nm.size() > 16 && nm.compare(nm.size() - 16, 16, ".upsample.filter") == 0;
            const bool is_dn = nm.size() > 26 && nm.compare(nm.size() - 26, 26, ".downsample.lowpass.filter") == 0;
            if (!is_up && !is_dn) continue;
            for (int q = 0; q < 12; ++q)
                REQUIRE(kv.second.data[q] == (is_up ? w.aa_up[q] : w.aa_down[q]), BVC_ERR_SCHEMA,
                        "anti-aliasing filters differ between activations; this build expects one shared filter");
        }
    }
    w.w_post = up(folded("conv_post"));   // [1, ci, 7] is already [ci][tap]
    w.b_post = up(to_vec(m["conv_post.bias"]));
    REQUIRE(ok, BVC_ERR_NOMEM, "device allocation failed while loading vocoder weights");
    h->have_voc = true;
    return BVC_OK;
}

int bvc_set_frontend(bvc_handle* h, const float* window, const float* mel_basis) {
    REQUIRE(h && window && mel_basis, BVC_ERR_INVALID, "bvc_set_frontend: null argument");
    Guard g(h);
    if (!g.ok) return BVC_ERR_DEVICE;
    const int N = h->cfg.n_fft, nb = N / 2 + 1, M = h->cfg.x_dim;
    std::vector<float> win(window, window + N);
    std::vector<float2> tw(N);
    for (int q = 0; q < N; ++q) {
        const double a = -2.0 * M_PI * q / N;
        tw[q] = make_float2((float)cos(a), (float)sin(a));
    }
    std::vector<int> start(M, 0), count(M, 0);
    int width = 1, used = 1;
    for (int m = 0; m < M; ++m) {
        int first = -1, last = -1;
        for (int k = 0; k < nb; ++k)
            if (mel_basis[(size_t)m * nb + k] != 0.f) { if (first < 0) first = k; last = k; }
        if (first >= 0) {
            start[m] = first;
            count[m] = last - first + 1;
            if (count[m] > width) width = count[m];
            if (last + 1 > used) used = last + 1;
        }
    }
    std::vector<float> taps((size_t)M * width, 0.f);
    for (int m = 0; m < M; ++m)
        for (int i = 0; i < count[m]; ++i) taps[(size_t)m * width + i] = mel_basis[(size_t)m * nb + start[m] + i];
    FrontendTables& ft = h->ft;
    ft.window = dev_upload(h, win);
    ft.twiddle = dev_upload(h, tw);
    ft.mel_start = dev_upload(h, start);
    ft.mel_count = dev_upload(h, count);
    ft.mel_taps = dev_upload(h, taps);
    ft.mel_width = width;
    ft.n_mels = M;
    ft.n_bins_used = used;
    REQUIRE(ft.window && ft.twiddle && ft.mel_start && ft.mel_count && ft.mel_taps, BVC_ERR_NOMEM,
            "device allocation failed in bvc_set_frontend");
    h->have_frontend = true;
    return BVC_OK;
}

int bvc_logmel(bvc_handle* h, const float* x_dev, int32_t B, int32_t L, float scale, float* mel_dev, void* stream) {
    REQUIRE(h && x_dev && mel_dev, BVC_ERR_INVALID, "bvc_logmel: null argument");
    REQUIRE(h->have_frontend, BVC_ERR_STATE, "bvc_logmel: front-end tables not set");
    const int need = h->cfg.n_fft - h->cfg.pad_left - h->cfg.hop;
    REQUIRE(B > 0 && L > need && L > h->cfg.pad_left, BVC_ERR_INVALID,
            "bvc_logmel: reflect padding needs L > " + std::to_string(need > h->cfg.pad_left ? need : h->cfg.pad_left));
    Guard g(h);
    if (!g.ok) return BVC_ERR_DEVICE;
    return logmel_forward(h->ft, x_dev, B, L, h->cfg.hop, h->cfg.pad_left, scale, mel_dev, (cudaStream_t)stream);
}

int bvc_encode(bvc_handle* h, const float* mel_dev, const float* bits_dev, float bits_scalar, const float* h0_dev,
               int32_t B, int32_t T, float* codes_dev, uint64_t* packed_dev, float* logits_dev, float* all_h_dev,
               float* h_final_dev, void* stream) {
    return bvc_encode_mel(h, mel_dev, bits_dev, bits_scalar, h0_dev, B, T, codes_dev, packed_dev, logits_dev, all_h_dev,
                          h_final_dev, nullptr, stream);
}

int bvc_encode_mel(bvc_handle* h, const float* mel_dev, const float* bits_dev, float bits_scalar, const float* h0_dev,
                   int32_t B, int32_t T, float* codes_dev, uint64_t* packed_dev, float* logits_dev, float* all_h_dev,
                   float* h_final_dev, float* mel_hat_dev, void* stream) {
    return bvc_encode_ex(h, mel_dev, bits_dev, bits_scalar, h0_dev, nullptr, B, T, codes_dev, packed_dev, logits_dev, all_h_dev,
                         h_final_dev, mel_hat_dev, nullptr, stream);
}

int bvc_encode_ex(bvc_handle* h, const float* mel_dev, const float* bits_dev, float bits_scalar, const float* h0_dev,
                  const float* uniforms_dev, int32_t B, int32_t T, float* codes_dev, uint64_t* packed_dev, float* logits_dev,
                  float* all_h_dev, float* h_final_dev, float* mel_hat_dev, float* prior_dev, void* stream) {
    REQUIRE(h && mel_dev && codes_dev, BVC_ERR_INVALID, "bvc_encode: null argument");
    REQUIRE(h->have_bvrnn, BVC_ERR_STATE, "bvc_encode: BVRNN weights not loaded");
    REQUIRE(B > 0 && T >= 0, BVC_ERR_INVALID, "bvc_encode: bad B/T");
    if (T == 0) return BVC_OK;
    Guard g(h);
    if (!g.ok) return BVC_ERR_DEVICE;
    int rc = rec_poll_aborts(h->bw.rw, false);      // a finished earlier launch that aborted is reported here
    if (rc) return rc;
    if (packed_dev && h->bw.Z != 64) {
        set_error("bvc_encode: packed output needs z_dim == 64 (one uint64 word per frame)");
        return BVC_ERR_INVALID;
    }
    const size_t n_allh = (prior_dev && !all_h_dev) ? (size_t)B * T * h->bw.H + 64 : 0;
    if ((rc = ensure_workspace(h, bvrnn_workspace_floats(h->bw, B, T) + n_allh))) return rc;
    if ((rc = ws_acquire(h, (cudaStream_t)stream))) return rc;
    float* all_h = all_h_dev;
    if (n_allh) all_h = h->ws.take(n_allh);       // the prior head reads the state entering every frame
    const size_t mark = h->ws.used;
    rc = bvrnn_encode(h->bw, h->ws, mel_dev, bits_dev, bits_scalar, h0_dev, B, T, codes_dev,
                      (unsigned long long*)packed_dev, logits_dev, all_h, h_final_dev, mel_hat_dev, h->precision,
                      (cudaStream_t)stream, uniforms_dev);
    if (!rc && prior_dev) {
        h->ws.used = mark;                         // the encoder's temporaries are dead in stream order
        rc = bvrnn_prior(h->bw, h->ws, all_h, B, T, prior_dev, h->precision, (cudaStream_t)stream);
    }
    const int rc2 = ws_release(h, (cudaStream_t)stream);   // also on errors: work may already be enqueued
    return rc ? rc : rc2;
}

int bvc_unpack_codes(bvc_handle* h, const uint64_t* packed_dev, const float* bits_dev, float bits_scalar, int32_t B,
                     int32_t T, float* codes_dev, void* stream) {
    REQUIRE(h && packed_dev && codes_dev, BVC_ERR_INVALID, "bvc_unpack_codes: null argument");
    REQUIRE(B > 0 && T >= 0, BVC_ERR_INVALID, "bvc_unpack_codes: bad B/T");
    Guard g(h);
    if (!g.ok) return BVC_ERR_DEVICE;
    return unpack_codes((const unsigned long long*)packed_dev, bits_dev, bits_scalar, h->cfg.var_bit, (size_t)B * T,
                        h->cfg.z_dim, codes_dev, (cudaStream_t)stream);
}

int bvc_decode_mel(bvc_handle* h, const float* codes_dev, const float* h0_dev, int32_t B, int32_t T, float* mel_dev,
                   float* h_final_dev, void* stream) {
    REQUIRE(h && codes_dev && mel_dev, BVC_ERR_INVALID, "bvc_decode_mel: null argument");
    REQUIRE(h->have_bvrnn, BVC_ERR_STATE, "bvc_decode_mel: BVRNN weights not loaded");
    REQUIRE(B > 0 && T >= 0, BVC_ERR_INVALID, "bvc_decode_mel: bad B/T");
    if (T == 0) return BVC_OK;
    Guard g(h);
    if (!g.ok) return BVC_ERR_DEVICE;
    int rc = rec_poll_aborts(h->bw.rw, false);
    if (rc) return rc;
    if ((rc = ensure_workspace(h, bvrnn_workspace_floats(h->bw, B, T)))) return rc;
    if ((rc = ws_acquire(h, (cudaStream_t)stream))) return rc;
    rc = bvrnn_decode(h->bw, h->ws, codes_dev, h0_dev, B, T, mel_dev, h_final_dev, h->precision,
                      (cudaStream_t)stream);
    const int rc2 = ws_release(h, (cudaStream_t)stream);
    return rc ? rc : rc2;
}

int bvc_decode_packed(bvc_handle* h, const uint64_t* packed_dev, const float* bits_dev, float bits_scalar, const float* h0_dev,
                      int32_t B, int32_t T, float* mel_dev, float* h_final_dev, void* stream) {
    REQUIRE(h && packed_dev && mel_dev, BVC_ERR_INVALID, "bvc_decode_packed: null argument");
    REQUIRE(h->have_bvrnn, BVC_ERR_STATE, "bvc_decode_packed: BVRNN weights not loaded");
    REQUIRE(B > 0 && T >= 0, BVC_ERR_INVALID, "bvc_decode_packed: bad B/T");
    REQUIRE(h->bw.Z <= 64, BVC_ERR_INVALID, "bvc_decode_packed: z_dim > 64 does not fit a 64-bit word");
    if (T == 0) return BVC_OK;
    Guard g(h);
    if (!g.ok) return BVC_ERR_DEVICE;
    int rc = rec_poll_aborts(h->bw.rw, false);
    if (rc) return rc;
    if ((rc = ensure_workspace(h, bvrnn_workspace_floats(h->bw, B, T)))) return rc;
    if ((rc = ws_acquire(h, (cudaStream_t)stream))) return rc;
    rc = bvrnn_decode(h->bw, h->ws, nullptr, h0_dev, B, T, mel_dev, h_final_dev, h->precision, (cudaStream_t)stream,
                      (const unsigned long long*)packed_dev, bits_dev, bits_scalar);
    const int rc2 = ws_release(h, (cudaStream_t)stream);
    return rc ? rc : rc2;
}

size_t bvc_bitstream_bytes(const bvc_handle* h, int32_t T, int32_t per_frame_budgets) {
    if (!h || T < 0) return 0;
    return bitstream_bytes(T, h->cfg.z_dim, per_frame_budgets && h->cfg.var_bit);
}

int bvc_pack_bitstream(bvc_handle* h, const uint64_t* packed_dev, const float* bits_dev, float bits_scalar, int32_t B, int32_t T,
                       uint8_t* out_dev, size_t stride, void* stream) {
    REQUIRE(h && packed_dev && out_dev, BVC_ERR_INVALID, "bvc_pack_bitstream: null argument");
    REQUIRE(B > 0 && T > 0, BVC_ERR_INVALID, "bvc_pack_bitstream: bad B/T");
    Guard g(h);
    if (!g.ok) return BVC_ERR_DEVICE;
    int rc = ensure_workspace(h, (size_t)B * (T + 1) + 256);
    if (rc) return rc;
    if ((rc = ws_acquire(h, (cudaStream_t)stream))) return rc;
    rc = pack_bitstream(h->ws, (const unsigned long long*)packed_dev, bits_dev, bits_scalar, h->cfg.var_bit, h->cfg.z_dim, B, T,
                        out_dev, stride, (cudaStream_t)stream);
    const int rc2 = ws_release(h, (cudaStream_t)stream);
    return rc ? rc : rc2;
}

int bvc_unpack_bitstream(bvc_handle* h, const uint8_t* in_dev, size_t stride, int32_t B, int32_t T, uint64_t* packed_dev,
                         float* bits_out_dev, void* stream) {
    REQUIRE(h && in_dev && packed_dev, BVC_ERR_INVALID, "bvc_unpack_bitstream: null argument");
    REQUIRE(B > 0 && T > 0, BVC_ERR_INVALID, "bvc_unpack_bitstream: bad B/T");
    Guard g(h);
    if (!g.ok) return BVC_ERR_DEVICE;
    int rc = ensure_workspace(h, (size_t)B * (T + 1) + 256);
    if (rc) return rc;
    if ((rc = ws_acquire(h, (cudaStream_t)stream))) return rc;
    rc = unpack_bitstream(h->ws, in_dev, stride, B, T, h->cfg.z_dim, (unsigned long long*)packed_dev, bits_out_dev,
                          (cudaStream_t)stream);
    const int rc2 = ws_release(h, (cudaStream_t)stream);
    return rc ? rc : rc2;
}

int64_t bvc_vocoder_out_len(const bvc_handle* h, int32_t T) {
    if (!h) return -1;
    int64_t n = T;
    for (int i = 0; i < h->cfg.voc_num_stages; ++i) n = (n + 1) * h->cfg.voc_up_rates[i];
    return n;
}

int bvc_vocode(bvc_handle* h, const float* mel_dev, int32_t B, int32_t T, int32_t length, float inv_scale_div,
               float* wav_dev, void* stream) {
    REQUIRE(h && mel_dev && wav_dev, BVC_ERR_INVALID, "bvc_vocode: null argument");
    REQUIRE(h->have_voc, BVC_ERR_STATE, "bvc_vocode: vocoder weights not loaded");
    REQUIRE(B > 0 && T > 0 && length >= 0, BVC_ERR_INVALID, "bvc_vocode: bad B/T/length");
    Guard g(h);
    if (!g.ok) return BVC_ERR_DEVICE;
    int rc = ensure_workspace(h, vocoder_workspace_floats(h->vw, B, T));
    if (rc) return rc;
    if ((rc = ws_acquire(h, (cudaStream_t)stream))) return rc;
    rc = vocoder_forward(h->vw, h->ws, h->vb, mel_dev, B, T, length, inv_scale_div, wav_dev, h->precision,
                         (cudaStream_t)stream);
    const int rc2 = ws_release(h, (cudaStream_t)stream);
    return rc ? rc : rc2;
}

int bvc_encode_host(bvc_handle* h, const float* x_host, int32_t B, int32_t L, float scale, float bits_scalar,
                    float* codes_host) {
    REQUIRE(h && x_host && codes_host, BVC_ERR_INVALID, "bvc_encode_host: null argument");
    REQUIRE(h->have_bvrnn && h->have_frontend, BVC_ERR_STATE, "bvc_encode_host: weights / front-end not loaded");
    const int need = h->cfg.n_fft - h->cfg.pad_left - h->cfg.hop;
    REQUIRE(B > 0 && L > need && L > h->cfg.pad_left, BVC_ERR_INVALID, "bvc_encode_host: L too short for reflect padding");
    Guard g(h);
    if (!g.ok) return BVC_ERR_DEVICE;
    const int T = L / h->cfg.hop, X = h->cfg.x_dim, Z = h->cfg.z_dim;
    const size_t nx = (size_t)B * L, nmel = (size_t)B * T * X, ncodes = (size_t)B * T * Z;
    int rc = ensure_workspace(h, nx + nmel + ncodes + 256 + bvrnn_workspace_floats(h->bw, B, T));
    if (rc) return rc;
    float* x_dev = h->ws.take(nx);
    float* mel = h->ws.take(nmel);
    float* codes = h->ws.take(ncodes);
    if ((rc = ws_check(h->ws, "bvc_encode_host"))) return rc;
    cudaStream_t s = h->stream;
    if ((rc = ws_acquire(h, s))) return rc;
    rc = [&]() -> int {
        BVC_CUDA(cudaMemcpyAsync(x_dev, x_host, nx * sizeof(float), cudaMemcpyHostToDevice, s));
        int r = logmel_forward(h->ft, x_dev, B, L, h->cfg.hop, h->cfg.pad_left, scale, mel, s);
        if (r) return r;
        if (T > 0) {
            r = bvrnn_encode(h->bw, h->ws, mel, nullptr, bits_scalar, nullptr, B, T, codes, nullptr, nullptr, nullptr,
                             nullptr, nullptr, h->precision, s);
            if (r) return r;
            BVC_CUDA(cudaMemcpyAsync(codes_host, codes, ncodes * sizeof(float), cudaMemcpyDeviceToHost, s));
        }
        return BVC_OK;
    }();
    const int rc2 = ws_release(h, s);
    if (rc || rc2) return rc ? rc : rc2;
    BVC_CUDA(cudaStreamSynchronize(s));
    return rec_poll_aborts(h->bw.rw, true);
}

int bvc_decode_host(bvc_handle* h, const float* codes_host, int32_t B, int32_t T, int32_t length, float inv_scale_div,
                    float* wav_host) {
    REQUIRE(h && codes_host && wav_host, BVC_ERR_INVALID, "bvc_decode_host: null argument");
    REQUIRE(h->have_bvrnn && h->have_voc, BVC_ERR_STATE, "bvc_decode_host: weights not loaded");
    REQUIRE(B > 0 && T > 0 && length >= 0, BVC_ERR_INVALID, "bvc_decode_host: bad B/T/length");
    Guard g(h);
    if (!g.ok) return BVC_ERR_DEVICE;
    const int X = h->cfg.x_dim, Z = h->cfg.z_dim;
    const int64_t n_full = bvc_vocoder_out_len(h, T);
    const size_t n_out = (size_t)(length < n_full ? length : n_full);
    const size_t ncodes = (size_t)B * T * Z, nmel = (size_t)B * T * X, nwav = (size_t)B * n_out;
    int rc = ensure_workspace(h, ncodes + nmel + nwav + 256 + bvrnn_workspace_floats(h->bw, B, T) +
                                     vocoder_workspace_floats(h->vw, B, T));
    if (rc) return rc;
    float* codes = h->ws.take(ncodes);
    float* mel = h->ws.take(nmel);
    float* wav = h->ws.take(nwav);
    if ((rc = ws_check(h->ws, "bvc_decode_host"))) return rc;
    cudaStream_t s = h->stream;
    if ((rc = ws_acquire(h, s))) return rc;
    rc = [&]() -> int {
        BVC_CUDA(cudaMemcpyAsync(codes, codes_host, ncodes * sizeof(float), cudaMemcpyHostToDevice, s));
        int r = bvrnn_decode(h->bw, h->ws, codes, nullptr, B, T, mel, nullptr, h->precision, s);
        if (r) return r;
        // Large batches are synthesised in two halves so that the device-to-host copy of the first half (PCIe, ~2 ms per
        // 100 MB) runs on the copy stream while the vocoder works on the second half (utterances are independent).
        const int Ba = (B >= 16 && nwav * sizeof(float) >= ((size_t)32 << 20)) ? B / 2 : B;
        const size_t mark = h->ws.used;
        r = vocoder_forward(h->vw, h->ws, h->vb, mel, Ba, T, length, inv_scale_div, wav, h->precision, s);
        if (r) return r;
        if (Ba < B) {
            const size_t na = (size_t)Ba * n_out;
            BVC_CUDA(cudaEventRecord(h->copy_ready, s));
            BVC_CUDA(cudaStreamWaitEvent(h->copy_stream, h->copy_ready, 0));
            if (na) BVC_CUDA(cudaMemcpyAsync(wav_host, wav, na * sizeof(float), cudaMemcpyDeviceToHost, h->copy_stream));
            BVC_CUDA(cudaEventRecord(h->copy_done, h->copy_stream));
            h->ws.used = mark;      // the first half's intermediates are dead in stream order; its output lies outside them
            r = vocoder_forward(h->vw, h->ws, h->vb, mel + (size_t)Ba * T * X, B - Ba, T, length, inv_scale_div, wav + na,
                                h->precision, s);
            if (r) return r;
            if (nwav - na)
                BVC_CUDA(cudaMemcpyAsync(wav_host + na, wav + na, (nwav - na) * sizeof(float), cudaMemcpyDeviceToHost, s));
            BVC_CUDA(cudaStreamWaitEvent(s, h->copy_done, 0));   // the workspace (it holds wav) is released after both copies
        } else if (nwav) {
            BVC_CUDA(cudaMemcpyAsync(wav_host, wav, nwav * sizeof(float), cudaMemcpyDeviceToHost, s));
        }
        return BVC_OK;
    }();
    const int rc2 = ws_release(h, s);
    if (rc || rc2) return rc ? rc : rc2;
    BVC_CUDA(cudaStreamSynchronize(s));
    return rec_poll_aborts(h->bw.rw, true);
}

// ---------------------------------------------------------------------------------------------
// streaming sessions
// ---------------------------------------------------------------------------------------------
struct bvc_stream {
    bvc_handle* h = nullptr;
    int S = 0;
    float *fifo = nullptr, *h_enc = nullptr, *h_dec = nullptr, *voc = nullptr;   // per-stream state
    int* n_samples = nullptr;
    float *win = nullptr, *mel = nullptr, *codes = nullptr, *h_tmp = nullptr, *dmel = nullptr;   // per-hop scratch
    unsigned long long* words = nullptr;
    unsigned char* valid = nullptr;
    size_t voc_floats = 0;
    std::vector<void*> allocs;
};

int bvc_stream_create(bvc_handle* h, int32_t n_streams, bvc_stream** out) {
    REQUIRE(h && out && n_streams > 0, BVC_ERR_INVALID, "bvc_stream_create: bad argument");
    REQUIRE(h->have_bvrnn && h->have_voc && h->have_frontend, BVC_ERR_STATE, "bvc_stream_create: weights / front-end not loaded");
    REQUIRE(h->bw.Z == 64 && h->bw.rw.ready, BVC_ERR_INVALID, "bvc_stream_create: needs z_dim = 64 and the persistent-kernel weights");
    *out = nullptr;
    Guard g(h);
    if (!g.ok) return BVC_ERR_DEVICE;
    bvc_stream* st = new bvc_stream();
    st->h = h;
    st->S = n_streams;
    const size_t S = (size_t)n_streams, H = h->bw.H, X = h->bw.X, Z = h->bw.Z, N = h->cfg.n_fft;
    st->voc_floats = vocoder_stream_state_floats(h->vw, n_streams);
    bool ok = true;
    auto alloc = [&](size_t bytes) -> void* {
        void* p = nullptr;
        if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); ok = false; return nullptr; }
        st->allocs.push_back(p);
        return p;
    };
    st->fifo = (float*)alloc(S * N * 4);        st->h_enc = (float*)alloc(S * H * 4);   st->h_dec = (float*)alloc(S * H * 4);
    st->voc = (float*)alloc(st->voc_floats * 4); st->n_samples = (int*)alloc(S * 4);
    st->win = (float*)alloc(S * N * 4);         st->mel = (float*)alloc(S * X * 4);     st->codes = (float*)alloc(S * Z * 4);
    st->h_tmp = (float*)alloc(S * H * 4);       st->dmel = (float*)alloc(S * X * 4);
    st->words = (unsigned long long*)alloc(S * 8); st->valid = (unsigned char*)alloc(S);
    if (!ok) {
        for (void* p : st->allocs) cudaFree(p);
        delete st;
        set_error("bvc_stream_create: device allocation failed");
        return BVC_ERR_NOMEM;
    }
    *out = st;
    return bvc_stream_reset(st, nullptr, nullptr);
}

int bvc_stream_destroy(bvc_stream* st) {
    if (!st) return BVC_OK;
    Guard g(st->h);
    cudaDeviceSynchronize();
    for (void* p : st->allocs) cudaFree(p);
    delete st;
    return BVC_OK;
}

int bvc_stream_reset(bvc_stream* st, const uint8_t* which_dev, void* stream) {
    REQUIRE(st, BVC_ERR_INVALID, "bvc_stream_reset: null session");
    bvc_handle* h = st->h;
    Guard g(h);
    if (!g.ok) return BVC_ERR_DEVICE;
    cudaStream_t s = (cudaStream_t)stream;
    const int S = st->S;
    int rc;
    if ((rc = stream_reset_rows(st->fifo, which_dev, S, h->cfg.n_fft, s))) return rc;
    if ((rc = stream_reset_rows(st->h_enc, which_dev, S, h->bw.H, s))) return rc;
    if ((rc = stream_reset_rows(st->h_dec, which_dev, S, h->bw.H, s))) return rc;
    if ((rc = stream_reset_ints(st->n_samples, which_dev, S, s))) return rc;
    if ((rc = vocoder_stream_reset(h->vw, st->voc, which_dev, S, s))) return rc;
    return BVC_OK;
}

int bvc_stream_encode_step(bvc_stream* st, const float* x_new_dev, const uint8_t* active_dev, const float* bits_dev,
                           float bits_scalar, float scale, uint64_t* packed_out_dev, uint8_t* valid_out_dev, void* stream) {
    REQUIRE(st && x_new_dev && packed_out_dev, BVC_ERR_INVALID, "bvc_stream_encode_step: null argument");
    bvc_handle* h = st->h;
    Guard g(h);
    if (!g.ok) return BVC_ERR_DEVICE;
    cudaStream_t s = (cudaStream_t)stream;
    const int S = st->S;
    int rc = rec_poll_aborts(h->bw.rw, false);
    if (rc) return rc;
    if ((rc = ensure_workspace(h, bvrnn_workspace_floats(h->bw, S, 1)))) return rc;
    if ((rc = ws_acquire(h, s))) return rc;
    rc = [&]() -> int {
        int r = stream_push(st->fifo, st->n_samples, x_new_dev, active_dev, S, h->cfg.hop, h->cfg.n_fft, h->cfg.pad_left, st->win,
                            st->valid, s);
        if (r) return r;
        // one frame per stream, already framed: L = hop = n_fft and pad_left = 0 make the front end take each row as one frame
        if ((r = logmel_forward(h->ft, st->win, S, h->cfg.n_fft, h->cfg.n_fft, 0, scale, st->mel, s))) return r;
        if ((r = bvrnn_encode(h->bw, h->ws, st->mel, bits_dev, bits_scalar, st->h_enc, S, 1, st->codes, st->words, nullptr, nullptr,
                              st->h_tmp, nullptr, h->precision, s)))
            return r;
        if ((r = stream_commit(st->h_enc, st->h_tmp, st->valid, S, h->bw.H, s))) return r;
        if ((r = stream_mask_words((unsigned long long*)packed_out_dev, st->words, st->valid, S, s))) return r;
        if (valid_out_dev) BVC_CUDA(cudaMemcpyAsync(valid_out_dev, st->valid, S, cudaMemcpyDeviceToDevice, s));
        return BVC_OK;
    }();
    const int rc2 = ws_release(h, s);
    return rc ? rc : rc2;
}

int bvc_stream_decode_step(bvc_stream* st, const uint64_t* packed_dev, const uint8_t* valid_dev, const float* bits_dev,
                           float bits_scalar, float inv_scale_div, float* wav_out_dev, void* stream) {
    REQUIRE(st && packed_dev && wav_out_dev, BVC_ERR_INVALID, "bvc_stream_decode_step: null argument");
    bvc_handle* h = st->h;
    Guard g(h);
    if (!g.ok) return BVC_ERR_DEVICE;
    cudaStream_t s = (cudaStream_t)stream;
    const int S = st->S;
    int rc = rec_poll_aborts(h->bw.rw, false);
    if (rc) return rc;
    const size_t need = std::max(bvrnn_workspace_floats(h->bw, S, 1), vocoder_stream_workspace_floats(h->vw, S));
    if ((rc = ensure_workspace(h, need))) return rc;
    if ((rc = ws_acquire(h, s))) return rc;
    rc = [&]() -> int {
        int r = bvrnn_decode(h->bw, h->ws, nullptr, st->h_dec, S, 1, st->dmel, st->h_tmp, h->precision, s,
                             (const unsigned long long*)packed_dev, bits_dev, bits_scalar);
        if (r) return r;
        if ((r = stream_commit(st->h_dec, st->h_tmp, valid_dev, S, h->bw.H, s))) return r;
        h->ws.used = 0;
        return vocoder_stream_step(h->vw, h->ws, st->voc, st->dmel, S, valid_dev, inv_scale_div, wav_out_dev, h->precision, s);
    }();
    const int rc2 = ws_release(h, s);
    return rc ? rc : rc2;
}

float bvc_last_recurrent_ms(const bvc_handle* h) {
    const float ms = bvc_recurrent_ms(const_cast<bvc_handle*>(h), -1, 0);
    return ms < 0.f ? 0.f : ms;
}

float bvc_recurrent_ms(bvc_handle* h, int32_t kind, int32_t age) {
    if (!h) return -1.f;
    Guard g(h);
    if (!g.ok) return -1.f;
    return rec_launch_ms(h->bw.rw, kind, age);
}

int bvc_check(bvc_handle* h) {
    REQUIRE(h, BVC_ERR_INVALID, "bvc_check: null handle");
    Guard g(h);
    if (!g.ok) return BVC_ERR_DEVICE;
    return rec_poll_aborts(h->bw.rw, true);
}

int bvc_host_alloc(void** out, size_t bytes) {
    REQUIRE(out && bytes > 0, BVC_ERR_INVALID, "bvc_host_alloc: bad argument");
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        set_error("cudaHostAlloc of " + std::to_string(bytes >> 20) + " MiB failed");
        return BVC_ERR_NOMEM;
    }
    *out = p;
    return BVC_OK;
}
int bvc_host_free(void* p) {
    if (p) BVC_CUDA(cudaFreeHost(p));
    return BVC_OK;
}

size_t bvc_workspace_bytes(const bvc_handle* h) { return h ? h->ws.bytes : 0; }
int64_t bvc_kernel_launches(const bvc_handle* h) { return h ? h->launches : 0; }

int bvc_debug_read(bvc_handle* h, const char* name, float* dst_host, size_t n_floats) {
    REQUIRE(h && name && dst_host, BVC_ERR_INVALID, "bvc_debug_read: null argument");
    Guard g(h);
    if (!g.ok) return BVC_ERR_DEVICE;
    const VocoderBuffers& vb = h->vb;
    const float* src = nullptr;
    size_t avail = 0;
    const std::string nm(name);
    if (nm == "voc_pre") {
        src = vb.pre;
        avail = (size_t)vb.B * (vb.T + 6) * h->vw.c0;
    } else if (nm.rfind("voc_stage", 0) == 0 && nm.size() == 12 && nm[10] == '_') {
        const int i = nm[9] - '0', j = nm[11] - '0';
        if (i >= 0 && i < 4 && j >= 0 && j < 3) {
            src = vb.part[i][j];
            avail = (size_t)vb.B * vb.C[i + 1] * vb.n[i + 1];
        }
    }
    const RecurrentWeights& rw = h->bw.rw;
    if (nm == "rec_dh") { src = rw.tap_dh; avail = (size_t)rw.tap_B * h->bw.H; }
    if (nm == "rec_gh") { src = rw.tap_gh; avail = (size_t)rw.tap_B * 3 * h->bw.H; }
    if (nm == "rec_giz") { src = rw.tap_giz; avail = (size_t)rw.tap_B * 3 * h->bw.H; }
    REQUIRE(src && avail > 0, BVC_ERR_INVALID, "bvc_debug_read: unknown buffer or no call has filled it yet: " + nm);
    REQUIRE(n_floats <= avail, BVC_ERR_INVALID, "bvc_debug_read: buffer holds fewer floats than requested");
    BVC_CUDA(cudaDeviceSynchronize());
    BVC_CUDA(cudaMemcpy(dst_host, src, n_floats * sizeof(float), cudaMemcpyDeviceToHost));
    return BVC_OK;
}

}  // extern "C"
