// DiT handle: packed fp16 parameter store keyed by the reference's state_dict names, the per-image
// conditioning hoist, and the per-step forward schedule (one stream, ~12 launches per block, no host sync).
//   reference: models/dit_crossattn.py:111-213, models/attention.py, models/utils.py
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/tpx.h"
#include "kernels.cuh"

using namespace tpx;

struct DitLayer {
    __half *Wq, *bq, *Wkv, *bkv, *Wcp, *bcp, *Wqkv, *bqkv, *Wsp, *bsp, *W1, *b1, *W2, *b2;
};

struct tpx_dit {
    int N, Cin, Cout, Dc, D, L, H, Dm, Dh, DhP, Ltot;
    __half* store = nullptr;
    size_t store_halves = 0;
    std::vector<DitLayer> layers;
    __half *Wada, *bada;                  // [L*9D + 2D, D], [L*9D + 2D]
    __half *null16, *Wx, *bx, *Wt0, *bt0, *Wt2, *bt2, *Wfl, *bfl;
    __half *uconst;                       // [L, D] : proj(to_v(null)) per block
    __half *tmp_v;                        // [D]
    std::vector<uint8_t> seen;            // per required key
    std::vector<std::string> required;
    bool finalized = false;
    bool has_null = false;                // null_cond_embedding given (cond_drop_prob > 0 models only; dit_crossattn.py:143-146)
    // conditioning store
    __half *ck = nullptr, *cv = nullptr, *y16 = nullptr;
    int cond_n = 0, cond_M = 0, cond_MP = 0;
    bool tc_attn = false;                 // tcgen05 attention (Dh == 72, N % 8 == 0); V kept transposed
    // timestep table (tpx_dit_set_timesteps): modulation rows [K, Ltot] of a sampling schedule's timesteps, in caller memory
    __half* ts_table = nullptr;
    std::vector<long long> ts_values;
};

static size_t al8(size_t n) { return (n + 7) & ~static_cast<size_t>(7); }

namespace {
struct Carver {
    __half* base;
    size_t off = 0;
    __half* take(size_t n) {
        __half* p = base == nullptr ? nullptr : base + off;
        off += al8(n);
        return p;
    }
};

void carve_store(tpx_dit* h, __half* base) {
    Carver c{base};
    const size_t D = h->D, Dc = h->Dc, Dm = h->Dm;
    h->layers.resize(h->L);
    for (int i = 0; i < h->L; ++i) {
        DitLayer& l = h->layers[i];
        l.Wq = c.take(D * D);       l.bq = c.take(D);
        l.Wkv = c.take(2 * D * Dc); l.bkv = c.take(2 * D);
        l.Wcp = c.take(D * D);      l.bcp = c.take(D);
        l.Wqkv = c.take(3 * D * D); l.bqkv = c.take(3 * D);
        l.Wsp = c.take(D * D);      l.bsp = c.take(D);
        l.W1 = c.take(Dm * D);      l.b1 = c.take(Dm);
        l.W2 = c.take(D * Dm);      l.b2 = c.take(D);
    }
    h->Wada = c.take(static_cast<size_t>(h->Ltot) * D);
    h->bada = c.take(h->Ltot);
    h->null16 = c.take(Dc);
    h->Wx = c.take(D * h->Cin);  h->bx = c.take(D);
    h->Wt0 = c.take(D * 256);    h->bt0 = c.take(D);
    h->Wt2 = c.take(D * D);      h->bt2 = c.take(D);
    h->Wfl = c.take(static_cast<size_t>(h->Cout) * D);  h->bfl = c.take(h->Cout);
    h->uconst = c.take(static_cast<size_t>(h->L) * D);
    h->tmp_v = c.take(D);
    h->store_halves = c.off;
}

struct Slot {
    __half* ptr;
    int64_t d0, d1;  // expected shape (d1 == 0 -> 1-D)
    bool optional;
};

// Map a reference state_dict key to its slot.  Returns false for an unknown key.
bool find_slot(tpx_dit* h, const std::string& key, Slot* s) {
    const int64_t D = h->D, Dc = h->Dc, Dm = h->Dm;
    auto set = [&](__half* p, int64_t a, int64_t b, bool opt = false) { *s = Slot{p, a, b, opt}; return true; };
    if (key == "null_cond_embedding") return set(h->null16, Dc, 0);
    if (key == "x_embedder.weight") return set(h->Wx, D, h->Cin);
    if (key == "x_embedder.bias") return set(h->bx, D, 0);
    if (key == "t_embedder.mlp.0.weight") return set(h->Wt0, D, 256);
    if (key == "t_embedder.mlp.0.bias") return set(h->bt0, D, 0);
    if (key == "t_embedder.mlp.2.weight") return set(h->Wt2, D, D);
    if (key == "t_embedder.mlp.2.bias") return set(h->bt2, D, 0);
    if (key == "final_layer.linear.weight") return set(h->Wfl, h->Cout, D);
    if (key == "final_layer.linear.bias") return set(h->bfl, h->Cout, 0);
    if (key == "final_layer.adaLN_modulation.1.weight") return set(h->Wada + static_cast<size_t>(h->L) * 9 * D * D, 2 * D, D);
    if (key == "final_layer.adaLN_modulation.1.bias") return set(h->bada + static_cast<size_t>(h->L) * 9 * D, 2 * D, 0);
    if (key.rfind("blocks.", 0) != 0) return false;
    const size_t dot = key.find('.', 7);
    if (dot == std::string::npos) return false;
    char* endp = nullptr;
    const long i = strtol(key.c_str() + 7, &endp, 10);
    if (endp != key.c_str() + dot || i < 0 || i >= h->L) return false;
    const std::string r = key.substr(dot + 1);
    DitLayer& l = h->layers[i];
    if (r == "crossattn.to_q.weight") return set(l.Wq, D, D);
    if (r == "crossattn.to_q.bias") return set(l.bq, D, 0, true);
    if (r == "crossattn.to_k.weight") return set(l.Wkv, D, Dc);
    if (r == "crossattn.to_k.bias") return set(l.bkv, D, 0, true);
    if (r == "crossattn.to_v.weight") return set(l.Wkv + D * Dc, D, Dc);
    if (r == "crossattn.to_v.bias") return set(l.bkv + D, D, 0, true);
    if (r == "crossattn.proj.weight") return set(l.Wcp, D, D);
    if (r == "crossattn.proj.bias") return set(l.bcp, D, 0, true);
    if (r == "attn.qkv.weight") return set(l.Wqkv, 3 * D, D);
    if (r == "attn.qkv.bias") return set(l.bqkv, 3 * D, 0, true);
    if (r == "attn.proj.weight") return set(l.Wsp, D, D);
    if (r == "attn.proj.bias") return set(l.bsp, D, 0, true);
    if (r == "mlp.fc1.weight") return set(l.W1, Dm, D);
    if (r == "mlp.fc1.bias") return set(l.b1, Dm, 0);
    if (r == "mlp.fc2.weight") return set(l.W2, D, Dm);
    if (r == "mlp.fc2.bias") return set(l.b2, D, 0);
    if (r == "adaLN_modulation.1.weight") return set(h->Wada + static_cast<size_t>(i) * 9 * D * D, 9 * D, D);
    if (r == "adaLN_modulation.1.bias") return set(h->bada + static_cast<size_t>(i) * 9 * D, 9 * D, 0);
    return false;
}

void build_required(tpx_dit* h) {
    auto& r = h->required;
    // null_cond_embedding is optional: the reference only creates it when cond_drop_prob > 0 (dit_crossattn.py:143-146)
    for (const char* k : {"x_embedder.weight", "x_embedder.bias", "t_embedder.mlp.0.weight", "t_embedder.mlp.0.bias",
                          "t_embedder.mlp.2.weight", "t_embedder.mlp.2.bias", "final_layer.linear.weight", "final_layer.linear.bias",
                          "final_layer.adaLN_modulation.1.weight", "final_layer.adaLN_modulation.1.bias"})
        r.emplace_back(k);
    for (int i = 0; i < h->L; ++i)
        for (const char* k : {"crossattn.to_q.weight", "crossattn.to_k.weight", "crossattn.to_v.weight", "crossattn.proj.weight", "attn.qkv.weight",
                              "attn.proj.weight", "mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias", "adaLN_modulation.1.weight",
                              "adaLN_modulation.1.bias"})
            r.emplace_back("blocks." + std::to_string(i) + "." + k);
    h->seen.assign(r.size(), 0);
}

// tile width: minimise waves x BN over {128,192,256}; ties go to the wider tile (less operand re-read)
int pick_bn(int M, int N) {
    const int sms = gemm_num_sms();
    const int tm = (M + 127) / 128;
    int best = 128;
    long best_cost = -1;
    for (int bn : {128, 192, 256}) {
        const long tiles = static_cast<long>(tm) * ((N + bn - 1) / bn);
        const long cost = ((tiles + sms - 1) / sms) * bn;
        if (best_cost < 0 || cost < best_cost || (cost == best_cost && bn > best)) { best = bn; best_cost = cost; }
    }
    return best;
}

struct DitWs {
    float* xres;
    __half *h16, *q, *k, *v, *ao, *hid, *mod, *ts16, *fin;
    float *th1, *temb;
    size_t total;
};

DitWs carve_ws(const tpx_dit* h, int S, uint8_t* base) {
    DitWs w;
    size_t off = 0;
    auto take = [&](size_t bytes) { uint8_t* p = base == nullptr ? nullptr : base + off; off += (bytes + 255) & ~static_cast<size_t>(255); return p; };
    const size_t T = static_cast<size_t>(S) * h->N;
    w.xres = reinterpret_cast<float*>(take(T * h->D * 4));
    w.h16 = reinterpret_cast<__half*>(take(T * h->D * 2));
    const size_t hb = static_cast<size_t>(S) * h->H * h->N * h->DhP * 2;
    w.q = reinterpret_cast<__half*>(take(hb));
    w.k = reinterpret_cast<__half*>(take(hb));
    w.v = reinterpret_cast<__half*>(take(hb));
    w.ao = reinterpret_cast<__half*>(take(T * h->D * 2));
    w.hid = reinterpret_cast<__half*>(take(T * h->Dm * 2));
    w.mod = reinterpret_cast<__half*>(take(static_cast<size_t>(S) * h->Ltot * 2));
    w.ts16 = reinterpret_cast<__half*>(take(static_cast<size_t>(S) * h->D * 2));
    w.fin = reinterpret_cast<__half*>(take(T * h->Cout * 2));
    w.th1 = reinterpret_cast<float*>(take(static_cast<size_t>(S) * h->D * 4));
    w.temb = reinterpret_cast<float*>(take(static_cast<size_t>(S) * h->D * 4));
    w.total = off;
    return w;
}

// Timestep table: [K, Ltot] fp16 modulation rows | K int64 timesteps (device copy) | scratch of one chunk of TS_CHUNK timesteps
constexpr int TS_CHUNK = 8;    // = the batched GEMV's maximum batch: the adaLN weights are streamed once per 8 timesteps
constexpr int TS_MAX = 4096;
struct TsWs {
    __half* table;
    long long* t_dev;
    float *th1, *temb;
    __half* ts16;
    size_t total;
};

TsWs carve_ts(const tpx_dit* h, int K, uint8_t* base) {
    TsWs w;
    size_t off = 0;
    auto take = [&](size_t bytes) { uint8_t* p = base == nullptr ? nullptr : base + off; off += (bytes + 255) & ~static_cast<size_t>(255); return p; };
    w.table = reinterpret_cast<__half*>(take(static_cast<size_t>(K) * h->Ltot * 2));
    w.t_dev = reinterpret_cast<long long*>(take(static_cast<size_t>(K) * 8));
    w.th1 = reinterpret_cast<float*>(take(static_cast<size_t>(TS_CHUNK) * h->D * 4));
    w.temb = reinterpret_cast<float*>(take(static_cast<size_t>(TS_CHUNK) * h->D * 4));
    w.ts16 = reinterpret_cast<__half*>(take(static_cast<size_t>(TS_CHUNK) * h->D * 2));
    w.total = off;
    return w;
}

// tile width for the 2-CTA kernel (256-row pair tiles over gemm_num_sms()/2 clusters)
int pick_bn_2cta(int M, int N) {
    const int pairs = gemm_num_sms() / 2;
    const int tm = (M + 255) / 256;
    int best = 128;
    long best_cost = -1;
    for (int bn : {128, 192, 256}) {
        const long tiles = static_cast<long>(tm) * ((N + bn - 1) / bn);
        const long cost = ((tiles + pairs - 1) / pairs) * bn;
        if (best_cost < 0 || cost < best_cost || (cost == best_cost && bn > best)) { best = bn; best_cost = cost; }
    }
    return best;
}

// TPX_GEMM_2CTA: unset / "auto" = the cta_group::2 kernel where it measured faster (long-K or very wide linears: fc1, fc2);
// 0 = never, 1 = every linear with M >= 256.
int mode_2cta() {
    static const int v = getenv("TPX_GEMM_2CTA") ? (getenv("TPX_GEMM_2CTA")[0] == 'a' ? 2 : atoi(getenv("TPX_GEMM_2CTA"))) : 2;
    return v;
}
bool use_2cta() { return mode_2cta() == 1; }
bool want_2cta(int M, int N, int K, int epi) {
    const int m = mode_2cta();
    if (m == 0 || M < 256) return false;
    if (m == 1) return true;
    return M >= 1024 && (K >= 2304 || N >= 4096) && epi != EPI_HEADS;
}

// tile_n: 0 = automatic; > 0 = 1-CTA kernel with that tile width; < 0 = 2-CTA (cta_group::2) kernel with width -tile_n
int gemm_linear(const __half* A, int lda, const __half* W, int M, int N, int K, int epi, const GemmArgs& args, int tile_n, cudaStream_t st) {
    GemmProblem p{};
    p.A = A; p.a_mode = AMODE_LINEAR; p.lda = lda; p.W = W; p.M = M; p.N = N; p.K = K;
    p.epi = epi; p.args = args;
    if (tile_n < 0 || (tile_n == 0 && want_2cta(M, N, K, epi))) {
        p.BN = tile_n < 0 ? -tile_n : pick_bn_2cta(M, N);
        return launch_gemm_2cta(p, st);
    }
    p.BN = tile_n > 0 ? tile_n : pick_bn(M, N);
    return launch_gemm(p, st);
}
}  // namespace

extern "C" {

int tpx_version(void) { return TPX_VERSION; }
const char* tpx_last_error(void) { return tpx::last_error(); }

int64_t tpx_launch_count(void) { return tpx::launch_count(); }
int tpx_profile_begin(void) { tpx::prof_begin(); return TPX_OK; }
int tpx_profile_end(float* ms_by_class, int64_t* launches_by_class) {
    TPX_CHECK(ms_by_class != nullptr && launches_by_class != nullptr, TPX_ERR_ARG, "profile_end: null argument");
    long long n[PROF_NCLASS];
    int rc = tpx::prof_end(ms_by_class, n);
    for (int i = 0; i < PROF_NCLASS; ++i) launches_by_class[i] = n[i];
    return rc;
}

int tpx_device_check(void) {
    static int checked[64] = {0};   // per device ordinal: 0 unknown, 1 ok (attribute queries are cheap, but this sits on per-kernel entry points)
    int dev = 0;
    TPX_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && checked[dev] == 1) return TPX_OK;
    int major = 0, minor = 0;
    TPX_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    TPX_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    TPX_CHECK(major == 10, TPX_ERR_CUDA, "libtpx_b200 is sm_100a only; device %d is sm_%d%d (no fallback path exists)", dev, major, minor);
    if (dev >= 0 && dev < 64) checked[dev] = 1;
    return TPX_OK;
}

int tpx_dit_create(const tpx_dit_config* c, tpx_dit** out) {
    TPX_CHECK(c != nullptr && out != nullptr, TPX_ERR_ARG, "dit_create: null argument");
    TPX_CHECK(c->hidden_size % 128 == 0 && c->hidden_size <= 2048, TPX_ERR_SHAPE, "dit_create: hidden_size %d must be a multiple of 128 (<= 2048)", c->hidden_size);
    TPX_CHECK(c->num_heads > 0 && c->hidden_size % c->num_heads == 0, TPX_ERR_SHAPE, "dit_create: hidden %d not divisible by heads %d", c->hidden_size, c->num_heads);
    const int Dh = c->hidden_size / c->num_heads;
    TPX_CHECK(Dh % 8 == 0 && Dh <= 128, TPX_ERR_SHAPE, "dit_create: head dim %d must be a multiple of 8 and <= 128", Dh);
    TPX_CHECK(c->condition_channels % 8 == 0 && c->mlp_hidden % 8 == 0 && c->out_channels % 8 == 0, TPX_ERR_SHAPE,
              "dit_create: condition/mlp/out channels must be multiples of 8 (%d/%d/%d)", c->condition_channels, c->mlp_hidden, c->out_channels);
    TPX_CHECK(c->depth > 0 && c->seq_length > 0 && c->in_channels > 0, TPX_ERR_SHAPE, "dit_create: non-positive size");
    TPX_CHECK(c->in_channels % 4 == 0, TPX_ERR_SHAPE, "dit_create: in_channels %d must be a multiple of 4", c->in_channels);
    int rc = tpx_device_check();
    if (rc != TPX_OK) return rc;
    tpx_dit* h = new tpx_dit();
    h->N = c->seq_length; h->Cin = c->in_channels; h->Cout = c->out_channels; h->Dc = c->condition_channels;
    h->D = c->hidden_size; h->L = c->depth; h->H = c->num_heads; h->Dm = c->mlp_hidden; h->Dh = Dh;
    h->DhP = Dh <= 16 ? 16 : (Dh <= 32 ? 32 : (Dh <= 64 ? 64 : (Dh <= 80 ? 80 : 128)));
    h->Ltot = h->L * 9 * h->D + 2 * h->D;
    h->tc_attn = (h->Dh == 72 && h->N % 8 == 0);   // the tcgen05 kernel keeps the row sums in padding row 72 of the 80-wide tiles
    carve_store(h, nullptr);
    cudaError_t e = cudaMalloc(&h->store, h->store_halves * 2);
    if (e != cudaSuccess) { delete h; return cuda_fail(e, "cudaMalloc(parameter store)"); }
    e = cudaMemset(h->store, 0, h->store_halves * 2);
    if (e != cudaSuccess) { cudaFree(h->store); delete h; return cuda_fail(e, "cudaMemset(parameter store)"); }
    carve_store(h, h->store);
    build_required(h);
    *out = h;
    return TPX_OK;
}

void tpx_dit_destroy(tpx_dit* h) {
    if (h == nullptr) return;
    if (h->store != nullptr) cudaFree(h->store);
    delete h;
}

int tpx_dit_set_weight(tpx_dit* h, const char* ref_key, const void* dev_ptr, int dtype, const int64_t* shape, int ndim, void* stream) {
    TPX_CHECK(h != nullptr && ref_key != nullptr && dev_ptr != nullptr && shape != nullptr, TPX_ERR_ARG, "dit_set_weight: null argument");
    Slot s;
    const std::string key(ref_key);
    if (!find_slot(h, key, &s)) { set_error("dit_set_weight: unexpected key '%s'", ref_key); return TPX_ERR_KEY; }
    const bool ok = s.d1 == 0 ? (ndim == 1 && shape[0] == s.d0) : (ndim == 2 && shape[0] == s.d0 && shape[1] == s.d1);
    TPX_CHECK(ok, TPX_ERR_SHAPE, "dit_set_weight: size mismatch for %s: expected [%lld%s%lld], got %d-d [%lld,...]", ref_key, (long long)s.d0,
              s.d1 ? "," : "", (long long)s.d1, ndim, (long long)shape[0]);
    const long long n = s.d1 == 0 ? s.d0 : s.d0 * s.d1;
    int rc = launch_to_half(dev_ptr, dtype, s.ptr, n, static_cast<cudaStream_t>(stream));
    if (rc != TPX_OK) return rc;
    for (size_t i = 0; i < h->required.size(); ++i)
        if (h->required[i] == key) h->seen[i] = 1;
    if (key == "null_cond_embedding") h->has_null = true;
    h->finalized = false;
    h->ts_values.clear();       // a timestep table derived from the previous weights is stale
    h->ts_table = nullptr;
    return TPX_OK;
}

int tpx_dit_get_weight(tpx_dit* h, const char* ref_key, void* dst_dev, int dtype, void* stream) {
    TPX_CHECK(h != nullptr && ref_key != nullptr && dst_dev != nullptr, TPX_ERR_ARG, "dit_get_weight: null argument");
    Slot s;
    if (!find_slot(h, std::string(ref_key), &s)) { set_error("dit_get_weight: unexpected key '%s'", ref_key); return TPX_ERR_KEY; }
    return launch_from_half(s.ptr, dst_dev, dtype, s.d1 == 0 ? s.d0 : s.d0 * s.d1, static_cast<cudaStream_t>(stream));
}

int tpx_dit_finalize(tpx_dit* h, void* stream) {
    TPX_CHECK(h != nullptr, TPX_ERR_ARG, "dit_finalize: null handle");
    for (size_t i = 0; i < h->required.size(); ++i)
        TPX_CHECK(h->seen[i], TPX_ERR_STATE, "dit_finalize: missing key '%s' in state_dict", h->required[i].c_str());
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // per-block constant of cross-attention against an all-null context: h(Wp . h(Wv . null + bv) + bp)
    for (int i = 0; h->has_null && i < h->L; ++i) {
        const DitLayer& l = h->layers[i];
        int rc = launch_gemv(GEMV_IN_F16, GEMV_OUT_F16, l.Wkv + static_cast<size_t>(h->D) * h->Dc, l.bkv + h->D, h->null16, nullptr, 1, h->D, h->Dc,
                             h->tmp_v, nullptr, h->D, st);
        if (rc != TPX_OK) return rc;
        rc = launch_gemv(GEMV_IN_F16, GEMV_OUT_F16, l.Wcp, l.bcp, h->tmp_v, nullptr, 1, h->D, h->D, h->uconst + static_cast<size_t>(i) * h->D, nullptr,
                         h->D, st);
        if (rc != TPX_OK) return rc;
    }
    h->finalized = true;
    return TPX_OK;
}

size_t tpx_dit_cond_bytes(const tpx_dit* h, int n_cross, int M) {
    if (h == nullptr || n_cross <= 0 || M <= 0) return 0;
    const int MP = (M + 7) & ~7;   // key count padded for the transposed V (TMA row pitch must be a multiple of 16 B)
    const size_t kv = static_cast<size_t>(h->L) * n_cross * h->H * MP * h->DhP * 2;
    const size_t y = (static_cast<size_t>(n_cross) * M * h->Dc * 2 + 255) & ~static_cast<size_t>(255);
    return 2 * ((kv + 255) & ~static_cast<size_t>(255)) + y;
}

size_t tpx_dit_workspace_bytes(const tpx_dit* h, int n_seq) {
    if (h == nullptr || n_seq <= 0) return 0;
    return carve_ws(h, n_seq, nullptr).total;
}

int tpx_dit_set_cond(tpx_dit* h, const float* y, int n_cross, int M, void* cond_ws, size_t cond_bytes, void* stream) {
    TPX_CHECK(h != nullptr && y != nullptr && cond_ws != nullptr, TPX_ERR_ARG, "dit_set_cond: null argument");
    TPX_CHECK(h->finalized, TPX_ERR_STATE, "dit_set_cond: weights not finalized (load_state_dict first)");
    TPX_CHECK(n_cross > 0 && M > 0, TPX_ERR_SHAPE, "dit_set_cond: empty conditioning (%d x %d)", n_cross, M);
    TPX_CHECK(cond_bytes >= tpx_dit_cond_bytes(h, n_cross, M), TPX_ERR_ARG, "dit_set_cond: conditioning store too small (%zu < %zu)", cond_bytes,
              tpx_dit_cond_bytes(h, n_cross, M));
    TPX_CHECK((reinterpret_cast<uintptr_t>(cond_ws) & 255) == 0, TPX_ERR_ARG, "dit_set_cond: store must be 256-B aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int MP = (M + 7) & ~7;
    const size_t kv = (static_cast<size_t>(h->L) * n_cross * h->H * MP * h->DhP * 2 + 255) & ~static_cast<size_t>(255);
    uint8_t* base = static_cast<uint8_t*>(cond_ws);
    h->ck = reinterpret_cast<__half*>(base);
    h->cv = reinterpret_cast<__half*>(base + kv);
    h->y16 = reinterpret_cast<__half*>(base + 2 * kv);
    h->cond_n = n_cross;
    h->cond_M = M;
    h->cond_MP = MP;
    if (h->tc_attn) TPX_CUDA(cudaMemsetAsync(h->cv, 0, kv, st));   // key columns [M, MP) of the transposed V must be finite
    int rc = launch_to_half(y, TPX_DTYPE_F32, h->y16, static_cast<long long>(n_cross) * M * h->Dc, st);
    if (rc != TPX_OK) return rc;
    const size_t per_layer = static_cast<size_t>(n_cross) * h->H * M * h->DhP;
    const size_t per_layer_v = h->tc_attn ? static_cast<size_t>(n_cross) * h->H * MP * h->DhP : per_layer;
    for (int i = 0; i < h->L; ++i) {
        GemmArgs a{};
        a.bias = h->layers[i].bkv;
        a.post_scale = 1.0f;
        a.out0 = h->ck + i * per_layer;
        a.out1 = h->cv + i * per_layer_v;
        a.split_cols = h->D; a.Dh = h->Dh; a.DhP = h->DhP; a.H = h->H; a.Nseq = M;
        if (h->tc_attn) { a.vt_which_plus1 = 2; a.vt_ld = MP; }
        rc = gemm_linear(h->y16, h->Dc, h->layers[i].Wkv, n_cross * M, 2 * h->D, h->Dc, EPI_HEADS, a, 0, st);
        if (rc != TPX_OK) return rc;
    }
    return TPX_OK;
}

// One forward.  The modulation vectors of the step come either from `t` (device timesteps: timestep MLP + the adaLN GEMV pass run here and
// fill the workspace's [B, Ltot] table) or, when `t` is null, from `mod_rows` = one precomputed row shared by the whole batch (tpx_dit_forward_step).
static int dit_forward_impl(tpx_dit* h, const float* x, const int64_t* t, const __half* mod_rows, int B, int use_cfg, float cfg_scale, void* out, void* ws,
                            size_t ws_bytes, void* stream) {
    TPX_CHECK(h != nullptr && x != nullptr && (t != nullptr || mod_rows != nullptr) && out != nullptr && ws != nullptr, TPX_ERR_ARG, "dit_forward: null argument");
    TPX_CHECK(h->finalized, TPX_ERR_STATE, "dit_forward: weights not finalized");
    TPX_CHECK(B >= 1 && B <= 8, TPX_ERR_SHAPE, "dit_forward: batch %d must be in [1,8]", B);
    TPX_CHECK(use_cfg >= 0 && use_cfg <= 2, TPX_ERR_ARG, "dit_forward: use_cfg %d", use_cfg);
    TPX_CHECK(use_cfg == 0 || h->has_null, TPX_ERR_STATE, "dit_forward: classifier-free guidance needs null_cond_embedding (model built with cond_drop_prob = 0)");
    const int S = use_cfg ? 2 * B : B;       // sequences in the batch
    const int Sc = use_cfg == 1 ? B : S;     // leading sequences with real cross-attention
    TPX_CHECK(h->ck != nullptr && h->cond_n == Sc, TPX_ERR_STATE, "dit_forward: conditioning holds %d sequences, this call needs %d (call set_cond)",
              h->cond_n, Sc);
    TPX_CHECK(ws_bytes >= tpx_dit_workspace_bytes(h, S), TPX_ERR_ARG, "dit_forward: workspace too small (%zu < %zu)", ws_bytes,
              tpx_dit_workspace_bytes(h, S));
    TPX_CHECK((reinterpret_cast<uintptr_t>(ws) & 255) == 0, TPX_ERR_ARG, "dit_forward: workspace must be 256-B aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const DitWs w = carve_ws(h, S, static_cast<uint8_t*>(ws));
    const int N = h->N, D = h->D, Ltot = h->Ltot, M = h->cond_M;
    const float qscale = 1.0f / sqrtf(static_cast<float>(h->Dh));
    int rc;
    static const bool dbg_sync = getenv("TPX_DEBUG_SYNC") != nullptr && getenv("TPX_DEBUG_SYNC")[0] == '1';
    int op_index = 0;
#define TPX_RC(call)                                                                                              \
    do {                                                                                                          \
        rc = (call);                                                                                              \
        if (rc != TPX_OK) return rc;                                                                              \
        if (dbg_sync) {                                                                                           \
            fprintf(stderr, "[tpx] op %d: %s\n", op_index, #call);                                               \
            fflush(stderr);                                                                                       \
            TPX_CUDA(cudaStreamSynchronize(st));                                                                  \
        }                                                                                                         \
        ++op_index;                                                                                               \
    } while (0)

    // The head-split GEMM epilogues write only the Dh real columns of each 80-wide head row (24-column bulk stores); the padding
    // columns / rows of q, k, v(T) must be zero, so the three buffers (contiguous in the workspace) are cleared once per forward.
    TPX_CUDA(cudaMemsetAsync(w.q, 0, reinterpret_cast<uint8_t*>(w.ao) - reinterpret_cast<uint8_t*>(w.q), st));
    // timestep embedding (fp32) -> silu -> fp16, then every adaLN modulation of the network in one GEMV pass — unless the row of this
    // step was hoisted out of the sampling loop (tpx_dit_set_timesteps): then the whole batch reads that one row (batch stride 0)
    const __half* mod_base = w.mod;
    int mod_bs = Ltot;
    if (t != nullptr) {
        TPX_RC(launch_gemv(GEMV_IN_TIMESTEP, GEMV_OUT_F32_SILU, h->Wt0, h->bt0, nullptr, reinterpret_cast<const long long*>(t), B, D, 256, w.th1, nullptr, D, st));
        TPX_RC(launch_gemv(GEMV_IN_F32, GEMV_OUT_F32_AND_SILU16, h->Wt2, h->bt2, w.th1, nullptr, B, D, D, w.temb, w.ts16, D, st));
        TPX_RC(launch_gemv(GEMV_IN_F16, GEMV_OUT_F16, h->Wada, h->bada, w.ts16, nullptr, B, Ltot, D, w.mod, nullptr, Ltot, st));
    } else {
        mod_base = mod_rows;
        mod_bs = 0;
    }
    // token embedding, fp32, written to both CFG halves
    TPX_RC(launch_x_embed(x, h->Wx, h->bx, B * N, h->Cin, D, w.xres, use_cfg ? static_cast<long long>(B) * N * D : 0, st));

    const size_t per_layer_kv = static_cast<size_t>(h->cond_n) * h->H * M * h->DhP;
    const size_t per_layer_v = h->tc_attn ? static_cast<size_t>(h->cond_n) * h->H * h->cond_MP * h->DhP : per_layer_kv;
    for (int i = 0; i < h->L; ++i) {
        const DitLayer& l = h->layers[i];
        const __half* mod = mod_base + static_cast<size_t>(i) * 9 * D;   // (shift,scale,gate) x (mca,msa,mlp)
        // ---- cross-attention branch (sequences [0,Sc)) ----
        TPX_RC(launch_ln_modulate(w.xres, Sc * N, D, 1e-6f, mod + 0 * D, mod + 1 * D, mod_bs, N, B, w.h16, nullptr, nullptr, 0, st));
        {
            GemmArgs a{};
            a.bias = l.bq; a.post_scale = qscale; a.out0 = w.q;
            a.split_cols = D; a.Dh = h->Dh; a.DhP = h->DhP; a.H = h->H; a.Nseq = N; a.heads_tma = 1;
            TPX_RC(gemm_linear(w.h16, D, l.Wq, Sc * N, D, D, EPI_HEADS, a, D % 144 == 0 && h->Dh % 24 == 0 && N % 32 == 0 ? (use_2cta() ? -144 : 144) : 0, st));
        }
        if (h->tc_attn) TPX_RC(launch_attention_tc(w.q, h->ck + i * per_layer_kv, h->cv + i * per_layer_v, w.ao, Sc, h->H, N, M, h->cond_MP, h->Dh, qscale, st));
        else TPX_RC(launch_attention(w.q, h->ck + i * per_layer_kv, h->cv + i * per_layer_kv, w.ao, Sc, h->H, N, M, h->Dh, h->DhP, qscale, st));
        {
            GemmArgs a{};
            a.bias = l.bcp; a.post_scale = 1.0f; a.xres = w.xres; a.ldx = D;
            a.gate = mod + 2 * D; a.gate_bstride = mod_bs; a.rows_per_batch = N; a.gate_batches = B;
            TPX_RC(gemm_linear(w.ao, D, l.Wcp, Sc * N, D, D, EPI_GATED, a, 0, st));
        }
        // ---- self-attention branch (all sequences); the collapsed cross-attention of the null half is added here ----
        TPX_RC(launch_ln_modulate(w.xres, S * N, D, 1e-6f, mod + 3 * D, mod + 4 * D, mod_bs, N, B, w.h16, Sc < S ? mod + 2 * D : nullptr,
                                  h->uconst + static_cast<size_t>(i) * D, Sc * N, st));
        {
            GemmArgs a{};
            a.bias = l.bqkv; a.post_scale = 1.0f; a.out0 = w.q; a.out1 = w.k; a.out2 = w.v;
            a.split_cols = D; a.Dh = h->Dh; a.DhP = h->DhP; a.H = h->H; a.Nseq = N; a.heads_tma = 1;
            if (h->tc_attn) { a.vt_which_plus1 = 3; a.vt_ld = N; }
            TPX_RC(gemm_linear(w.h16, D, l.Wqkv, S * N, 3 * D, D, EPI_HEADS, a, 0, st));
        }
        if (h->tc_attn) TPX_RC(launch_attention_tc(w.q, w.k, w.v, w.ao, S, h->H, N, N, N, h->Dh, qscale, st));
        else TPX_RC(launch_attention(w.q, w.k, w.v, w.ao, S, h->H, N, N, h->Dh, h->DhP, qscale, st));
        {
            GemmArgs a{};
            a.bias = l.bsp; a.post_scale = 1.0f; a.xres = w.xres; a.ldx = D;
            a.gate = mod + 5 * D; a.gate_bstride = mod_bs; a.rows_per_batch = N; a.gate_batches = B;
            TPX_RC(gemm_linear(w.ao, D, l.Wsp, S * N, D, D, EPI_GATED, a, 0, st));
        }
        // ---- MLP branch ----
        TPX_RC(launch_ln_modulate(w.xres, S * N, D, 1e-6f, mod + 6 * D, mod + 7 * D, mod_bs, N, B, w.h16, nullptr, nullptr, 0, st));
        {
            GemmArgs a{};
            a.bias = l.b1; a.post_scale = 1.0f; a.out0 = w.hid; a.ldo = h->Dm;
            TPX_RC(gemm_linear(w.h16, D, l.W1, S * N, h->Dm, D, EPI_GELU, a, 0, st));
        }
        {
            GemmArgs a{};
            a.bias = l.b2; a.post_scale = 1.0f; a.xres = w.xres; a.ldx = D;
            a.gate = mod + 8 * D; a.gate_bstride = mod_bs; a.rows_per_batch = N; a.gate_batches = B;
            TPX_RC(gemm_linear(w.hid, h->Dm, l.W2, S * N, D, h->Dm, EPI_GATED, a, 0, st));
        }
    }
    // ---- final layer ----
    const __half* fmod = mod_base + static_cast<size_t>(h->L) * 9 * D;
    TPX_RC(launch_ln_modulate(w.xres, S * N, D, 1e-6f, fmod, fmod + D, mod_bs, N, B, w.h16, nullptr, nullptr, 0, st));
    {
        GemmArgs a{};
        a.bias = h->bfl; a.post_scale = 1.0f; a.out0 = use_cfg ? w.fin : static_cast<__half*>(out); a.ldo = h->Cout;
        TPX_RC(gemm_linear(w.h16, D, h->Wfl, S * N, h->Cout, D, EPI_STORE, a, 128, st));
    }
    if (use_cfg) TPX_RC(launch_cfg_combine(w.fin, static_cast<long long>(B) * N * h->Cout, cfg_scale, static_cast<__half*>(out), st));
#undef TPX_RC
    return TPX_OK;
}

int tpx_dit_forward(tpx_dit* h, const float* x, const int64_t* t, int B, int use_cfg, float cfg_scale, void* out, void* ws, size_t ws_bytes,
                    void* stream) {
    TPX_CHECK(t != nullptr, TPX_ERR_ARG, "dit_forward: null timesteps");
    return dit_forward_impl(h, x, t, nullptr, B, use_cfg, cfg_scale, out, ws, ws_bytes, stream);
}

size_t tpx_dit_timesteps_bytes(const tpx_dit* h, int K) {
    if (h == nullptr || K <= 0 || K > TS_MAX) return 0;
    return carve_ts(h, K, nullptr).total;
}

int tpx_dit_set_timesteps(tpx_dit* h, const int64_t* t_host, int K, void* ts_ws, size_t ts_bytes, void* stream) {
    TPX_CHECK(h != nullptr && t_host != nullptr && ts_ws != nullptr, TPX_ERR_ARG, "dit_set_timesteps: null argument");
    TPX_CHECK(h->finalized, TPX_ERR_STATE, "dit_set_timesteps: weights not finalized (load_state_dict first)");
    TPX_CHECK(K >= 1 && K <= TS_MAX, TPX_ERR_SHAPE, "dit_set_timesteps: %d timesteps (1..%d)", K, TS_MAX);
    TPX_CHECK(ts_bytes >= tpx_dit_timesteps_bytes(h, K), TPX_ERR_ARG, "dit_set_timesteps: table store too small (%zu < %zu)", ts_bytes,
              tpx_dit_timesteps_bytes(h, K));
    TPX_CHECK((reinterpret_cast<uintptr_t>(ts_ws) & 255) == 0, TPX_ERR_ARG, "dit_set_timesteps: store must be 256-B aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const TsWs w = carve_ts(h, K, static_cast<uint8_t*>(ts_ws));
    h->ts_values.clear();
    h->ts_table = nullptr;
    // pageable source: the driver stages the values before the call returns, so t_host need not outlive it
    TPX_CUDA(cudaMemcpyAsync(w.t_dev, t_host, static_cast<size_t>(K) * 8, cudaMemcpyHostToDevice, st));
    const int D = h->D, Ltot = h->Ltot;
    for (int c = 0; c < K; c += TS_CHUNK) {
        const int nb = K - c < TS_CHUNK ? K - c : TS_CHUNK;
        // the same three launches a forward issues for its batch (utils.py:27-64, dit_crossattn.py:54,75), over 8 timesteps at once:
        // per output element the arithmetic does not depend on the batch it runs in, so a row equals what tpx_dit_forward computes
        int rc = launch_gemv(GEMV_IN_TIMESTEP, GEMV_OUT_F32_SILU, h->Wt0, h->bt0, nullptr, w.t_dev + c, nb, D, 256, w.th1, nullptr, D, st);
        if (rc != TPX_OK) return rc;
        rc = launch_gemv(GEMV_IN_F32, GEMV_OUT_F32_AND_SILU16, h->Wt2, h->bt2, w.th1, nullptr, nb, D, D, w.temb, w.ts16, D, st);
        if (rc != TPX_OK) return rc;
        rc = launch_gemv(GEMV_IN_F16, GEMV_OUT_F16, h->Wada, h->bada, w.ts16, nullptr, nb, Ltot, D, w.table + static_cast<size_t>(c) * Ltot, nullptr, Ltot, st);
        if (rc != TPX_OK) return rc;
    }
    h->ts_values.assign(t_host, t_host + K);
    h->ts_table = w.table;
    return TPX_OK;
}

int tpx_dit_forward_step(tpx_dit* h, const float* x, int64_t t, int B, int use_cfg, float cfg_scale, void* out, void* ws, size_t ws_bytes, void* stream) {
    TPX_CHECK(h != nullptr, TPX_ERR_ARG, "dit_forward_step: null handle");
    TPX_CHECK(h->ts_table != nullptr && !h->ts_values.empty(), TPX_ERR_STATE, "dit_forward_step: no timestep table (call tpx_dit_set_timesteps after loading weights)");
    size_t row = h->ts_values.size();
    for (size_t i = 0; i < h->ts_values.size(); ++i)
        if (h->ts_values[i] == static_cast<long long>(t)) { row = i; break; }
    TPX_CHECK(row < h->ts_values.size(), TPX_ERR_STATE, "dit_forward_step: timestep %lld is not in the table given to tpx_dit_set_timesteps", (long long)t);
    return dit_forward_impl(h, x, nullptr, h->ts_table + row * static_cast<size_t>(h->Ltot), B, use_cfg, cfg_scale, out, ws, ws_bytes, stream);
}

int tpx_dit_debug_residual(const tpx_dit* h, const void* ws, int n_seq, float* out, void* stream) {
    TPX_CHECK(h != nullptr && ws != nullptr && out != nullptr, TPX_ERR_ARG, "dit_debug_residual: null argument");
    const DitWs w = carve_ws(h, n_seq, static_cast<uint8_t*>(const_cast<void*>(ws)));
    TPX_CUDA(cudaMemcpyAsync(out, w.xres, static_cast<size_t>(n_seq) * h->N * h->D * 4, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
    return TPX_OK;
}

// ---------------------------------------------------------------------------------------------------------
// sampler + per-kernel entry points
// ---------------------------------------------------------------------------------------------------------
int tpx_sampler_step(int ddim, const float* x, const void* mo, int mo_dtype, const float* noise, int64_t n, int C, const tpx_sampler_coefs* k,
                     float* x_prev, float* x0, void* stream) {
    TPX_CHECK(x != nullptr && mo != nullptr && k != nullptr && x_prev != nullptr && x0 != nullptr, TPX_ERR_ARG, "sampler_step: null argument");
    TPX_CHECK(C > 0 && n % C == 0, TPX_ERR_SHAPE, "sampler_step: n %lld not a multiple of C %d", (long long)n, C);
    SamplerCoefs c;
    static_assert(sizeof(SamplerCoefs) == sizeof(tpx_sampler_coefs), "coef struct mirror");
    memcpy(&c, k, sizeof(c));
    return launch_sampler_step(ddim, x, mo, mo_dtype == TPX_DTYPE_F16, noise, n, C, c, x_prev, x0, static_cast<cudaStream_t>(stream));
}

int tpx_latent_split(const float* sample, const float* mean, const float* stdv, float inv_nf, int64_t T, int C, float* srt, float* z, void* stream) {
    TPX_CHECK(sample != nullptr && srt != nullptr && z != nullptr, TPX_ERR_ARG, "latent_split: null argument");
    TPX_CHECK((mean == nullptr) == (stdv == nullptr), TPX_ERR_ARG, "latent_split: latent_mean and latent_std go together");
    TPX_CHECK(T >= 0 && C > 4, TPX_ERR_SHAPE, "latent_split: need more than 4 channels (got %d)", C);
    return launch_latent_split(sample, mean, stdv, inv_nf, T, C, srt, z, static_cast<cudaStream_t>(stream));
}

int tpx_primvolume_pack(const float* srt, const void* decoded, int decoded_dtype, int64_t T, int F, int vox, int srt_fix, float* out, void* stream) {
    TPX_CHECK(srt != nullptr && decoded != nullptr && out != nullptr, TPX_ERR_ARG, "primvolume_pack: null argument");
    TPX_CHECK(decoded_dtype == TPX_DTYPE_F32 || decoded_dtype == TPX_DTYPE_F16, TPX_ERR_ARG, "primvolume_pack: dtype %d", decoded_dtype);
    return launch_primvolume_pack(srt, decoded, decoded_dtype == TPX_DTYPE_F16, T, F, vox, srt_fix, out, static_cast<cudaStream_t>(stream));
}

int tpx_linear(const void* A, int lda, const void* W, const void* bias, void* out, int ldo, int M, int N, int K, int act, float post_scale, int tile_n,
               void* stream) {
    TPX_CHECK(A != nullptr && W != nullptr && out != nullptr, TPX_ERR_ARG, "linear: null argument");
    int rc = tpx_device_check();
    if (rc != TPX_OK) return rc;
    GemmArgs a{};
    a.bias = static_cast<const __half*>(bias); a.post_scale = post_scale; a.out0 = static_cast<__half*>(out); a.ldo = ldo;
    return gemm_linear(static_cast<const __half*>(A), lda, static_cast<const __half*>(W), M, N, K, act ? EPI_GELU : EPI_STORE, a, tile_n,
                       static_cast<cudaStream_t>(stream));
}

int tpx_linear_gated(const void* A, int lda, const void* W, const void* bias, const void* gate, int gate_bstride, int gate_batches, int rows_per_batch,
                     float* xres, int ldx, int M, int N, int K, int tile_n, void* stream) {
    TPX_CHECK(A != nullptr && W != nullptr && gate != nullptr && xres != nullptr, TPX_ERR_ARG, "linear_gated: null argument");
    TPX_CHECK(gate_batches > 0 && rows_per_batch > 0, TPX_ERR_ARG, "linear_gated: bad gate batching");
    int rc = tpx_device_check();
    if (rc != TPX_OK) return rc;
    GemmArgs a{};
    a.bias = static_cast<const __half*>(bias); a.post_scale = 1.0f; a.xres = xres; a.ldx = ldx;
    a.gate = static_cast<const __half*>(gate); a.gate_bstride = gate_bstride; a.rows_per_batch = rows_per_batch; a.gate_batches = gate_batches;
    return gemm_linear(static_cast<const __half*>(A), lda, static_cast<const __half*>(W), M, N, K, EPI_GATED, a, tile_n, static_cast<cudaStream_t>(stream));
}

int tpx_linear_heads(const void* A, int lda, const void* W, const void* bias, void* out0, void* out1, void* out2, int M, int N, int K, int split_cols,
                     int H, int Dh, int DhP, int n_seq_tokens, float post_scale, int tile_n, int transposed_which, int transposed_ld, void* stream) {
    TPX_CHECK(A != nullptr && W != nullptr && out0 != nullptr, TPX_ERR_ARG, "linear_heads: null argument");
    TPX_CHECK(Dh % 8 == 0 && DhP % 8 == 0 && DhP >= Dh && split_cols == H * Dh && N % split_cols == 0, TPX_ERR_SHAPE, "linear_heads: bad head geometry");
    int rc = tpx_device_check();
    if (rc != TPX_OK) return rc;
    GemmArgs a{};
    a.bias = static_cast<const __half*>(bias); a.post_scale = post_scale;
    a.out0 = static_cast<__half*>(out0); a.out1 = static_cast<__half*>(out1); a.out2 = static_cast<__half*>(out2);
    a.split_cols = split_cols; a.Dh = Dh; a.DhP = DhP; a.H = H; a.Nseq = n_seq_tokens;
    a.vt_which_plus1 = transposed_which >= 0 ? transposed_which + 1 : 0;
    a.vt_ld = transposed_ld;
    if (Dh % 24 == 0 && n_seq_tokens % 32 == 0 && n_seq_tokens > 0 && M % n_seq_tokens == 0) {
        // bulk-store epilogue: it writes the Dh real columns only, so the padding is cleared here (the DiT forward clears its
        // workspace once per call instead)
        void* outs[3] = {out0, out1, out2};
        const size_t per = static_cast<size_t>(M / n_seq_tokens) * H * n_seq_tokens * DhP * 2;
        for (int w = 0; w < N / split_cols && w < 3; ++w) {
            TPX_CHECK(outs[w] != nullptr, TPX_ERR_ARG, "linear_heads: output %d is null", w);
            const size_t bytes = w == transposed_which ? static_cast<size_t>(M / n_seq_tokens) * H * DhP * transposed_ld * 2 : per;
            TPX_CUDA(cudaMemsetAsync(outs[w], 0, bytes, static_cast<cudaStream_t>(stream)));
        }
        a.heads_tma = 1;
    }
    return gemm_linear(static_cast<const __half*>(A), lda, static_cast<const __half*>(W), M, N, K, EPI_HEADS, a, tile_n, static_cast<cudaStream_t>(stream));
}

int tpx_ln_modulate(float* x, int rows, int D, float eps, const void* shift, const void* scale, int mod_bstride, int rows_per_batch, int mod_batches,
                    void* out, const void* pre_gate, const void* pre_const, int pre_row0, void* stream) {
    TPX_CHECK(x != nullptr && shift != nullptr && scale != nullptr && out != nullptr, TPX_ERR_ARG, "ln_modulate: null argument");
    TPX_CHECK(rows_per_batch > 0 && mod_batches > 0, TPX_ERR_ARG, "ln_modulate: bad batching");
    return launch_ln_modulate(x, rows, D, eps, static_cast<const __half*>(shift), static_cast<const __half*>(scale), mod_bstride, rows_per_batch,
                              mod_batches, static_cast<__half*>(out), static_cast<const __half*>(pre_gate), static_cast<const __half*>(pre_const),
                              pre_row0, static_cast<cudaStream_t>(stream));
}

int tpx_attention(const void* q, const void* k, const void* v, void* out, int B, int H, int Nq, int Nk, int Dh, int DhP, float scale, void* stream) {
    TPX_CHECK(q != nullptr && k != nullptr && v != nullptr && out != nullptr, TPX_ERR_ARG, "attention: null argument");
    return launch_attention(static_cast<const __half*>(q), static_cast<const __half*>(k), static_cast<const __half*>(v), static_cast<__half*>(out), B, H,
                            Nq, Nk, Dh, DhP, scale, static_cast<cudaStream_t>(stream));
}

int tpx_attention_tc(const void* q, const void* k, const void* vT, void* out, int B, int H, int Nq, int Nk, int NkPad, int Dh, float scale, void* stream) {
    TPX_CHECK(q != nullptr && k != nullptr && vT != nullptr && out != nullptr, TPX_ERR_ARG, "attention_tc: null argument");
    int rc = tpx_device_check();
    if (rc != TPX_OK) return rc;
    return launch_attention_tc(static_cast<const __half*>(q), static_cast<const __half*>(k), static_cast<const __half*>(vT), static_cast<__half*>(out), B, H,
                               Nq, Nk, NkPad, Dh, scale, static_cast<cudaStream_t>(stream));
}

int tpx_attention_tc_debug(const void* q, const void* k, const void* vT, void* out, int B, int H, int Nq, int Nk, int NkPad, int Dh, float scale,
                            int64_t* timeline_dev, void* stream) {
    return launch_attention_tc(static_cast<const __half*>(q), static_cast<const __half*>(k), static_cast<const __half*>(vT), static_cast<__half*>(out), B, H,
                               Nq, Nk, NkPad, Dh, scale, static_cast<cudaStream_t>(stream), reinterpret_cast<long long*>(timeline_dev));
}

int tpx_debug_gemm_timeline(int64_t* timeline_dev) {
    tpx::set_gemm_timeline(reinterpret_cast<long long*>(timeline_dev));
    return TPX_OK;
}

int tpx_cfg_combine(const void* both, int64_t n_half, float s, void* out, void* stream) {
    TPX_CHECK(both != nullptr && out != nullptr, TPX_ERR_ARG, "cfg_combine: null argument");
    return launch_cfg_combine(static_cast<const __half*>(both), n_half, s, static_cast<__half*>(out), static_cast<cudaStream_t>(stream));
}

}  // extern "C"
