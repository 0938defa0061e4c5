// Engine + batch state behind the C ABI (include/vvb200.h): weight loading, device layouts, the preprocess ->
// (nfe-1) x DiT step -> decode schedule, CUDA-graph capture of the sampling loop.
//
// Host-side counterpart of ModelSessionManager's three sessions and TTSEngine's per-chunk loop
// (/root/reference/vietvoicetts/core/model.py:65-135, /root/reference/vietvoicetts/core/tts_engine.py:225-238).
#include "../../include/vvb200.h"
#include "frontend.h"
#include "kernels.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <vector>

using namespace vv;

// ------------------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CK(call)                                                                                     \
  do {                                                                                               \
    cudaError_t _e = (call);                                                                         \
    if (_e != cudaSuccess)                                                                           \
      return fail(VV_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)
#define CKL()                                                                                        \
  do {                                                                                               \
    cudaError_t _e = cudaGetLastError();                                                             \
    if (_e != cudaSuccess)                                                                           \
      return fail(VV_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// ------------------------------------------------------------------------------------------------ structs
struct WTensor {
  float* d = nullptr;
  int64_t shape[4] = {0, 0, 0, 0};
  int ndim = 0;
  size_t numel = 0;
};

struct ModTable {
  int nfe = 0, steps = 0;
  float* blocks = nullptr;  // [steps][depth][6*dim]
  float* fin = nullptr;     // [steps][2*dim]
  std::vector<float> dt;    // [steps]
  uint64_t last_use = 0;
};

// A reference prompt resident in HBM: PCM, its log-mel and ref_signal_len, keyed by a content
// hash.  All chunks of one long text — and every request for the same voice — share ONE prompt
// (/root/reference/vietvoicetts/core/tts_engine.py:225-238 feeds the same audio to every chunk;
// core/model.py:204-211 re-reads it from the tar per request), so it is uploaded and its mel computed once.
struct Prompt {
  uint64_t id = 0, check = 0;     // two independent 64-bit content hashes
  int64_t n_samples = 0;
  int ref_len = 0;                // n_samples / hop + 1 mel frames
  int16_t* pcm = nullptr;
  float* mel = nullptr;           // [ref_len, n_mel]
  float* scale = nullptr;         // RMS normalisation factor (device scalar)
  cudaStream_t st = nullptr;
  uint64_t last_use = 0;
  ~Prompt() {
    if (pcm) cudaFreeAsync(pcm, st);
    if (mel) cudaFreeAsync(mel, st);
    if (scale) cudaFreeAsync(scale, st);
  }
};

struct GemmOp {
  CUtensorMap tA, tB;
  CUtensorMap tBt;          // CTA-pair kernel: B with 32-row boxes (tail-wave column slices);
                            // conv op: the ACTIVATIONS with 32-row boxes (last piece of the resident halo, conv.cu)
  GemmShape s;
  int bn = 128;
};

struct LayerW {
  bf16 *qkv, *out, *ff1, *ff2;
  const float *qkv_b, *out_b, *ff1_b, *ff2_b;
};
struct ConvNextW {
  const float *dw_w, *dw_b, *ln_g, *ln_b, *pw1_b, *pw2_b, *grn_g, *grn_b, *gamma;
  bf16 *pw1, *pw2;
};

struct vv_engine {
  vv_arch a;
  int device = 0;
  cudaStream_t st = nullptr;
  bool own_stream = false;
  int num_sms = 148;
  bool finalized = false;
  std::map<std::string, WTensor> w;
  std::vector<void*> allocs;
  int64_t launches = 0;
  // tables
  float2* rope_cs = nullptr;
  int rope_max = 4096;
  float2* fft_tw = nullptr;
  float* hann = nullptr;
  float* pos_table = nullptr;
  // per-NFE time-embedding / AdaLN modulation tables: a small LRU (VVB200_MOD_TABLES, default 4) — a table costs
  // (nfe-1) * depth * 6 * dim * 4 bytes (0.54 MB per step at full size) and a request stream may sweep nfe
  std::map<int, ModTable> mod;
  uint64_t mod_tick = 0;
  std::set<vv_batch*> live;       // every batch of this engine (cached or caller-owned)
  // device-resident prompts, LRU of VVB200_PROMPT_CACHE entries (default 64; 0 disables caching)
  std::map<uint64_t, std::shared_ptr<Prompt>> prompts;
  uint64_t prompt_tick = 0;
  int64_t prompt_uploads = 0, prompt_hits = 0;
  // bf16 weights
  bf16 *in_n = nullptr, *in_c = nullptr, *c1 = nullptr, *c2 = nullptr, *out_w = nullptr;
  int Kn = 128, Kc = 0, Kemb = 0;
  std::vector<LayerW> layers;
  std::vector<ConvNextW> text_blocks, voc_blocks;
  bf16 *voc_embed = nullptr, *voc_head = nullptr;
  // cached batches for vv_synthesize_batch, keyed by the frame counts of the chunks.  A batch owns ~1 GB of
  // activations (B = 8, T = 1500) plus its CUDA graphs, and a request stream produces a new key for nearly every
  // micro-batch, so the cache is a small LRU (VVB200_BATCH_CACHE entries, default 6), not a map that only grows.
  std::map<std::vector<int64_t>, vv_batch*> batch_cache;
  std::map<vv_batch*, uint64_t> batch_last_use;
  uint64_t batch_tick = 0;
  // every public entry point that submits work or touches engine state takes this: callers may share one engine
  // between threads (the REST layer of the reference runs requests on worker threads, api/tts_engine.py:79-87)
  std::recursive_mutex mu;
  std::vector<std::pair<size_t, void*>> pinned_free;   // recycled pinned staging buffers (bytes, pointer)
  // ONE instantiated graph of the sampling loop per (nfe, launch mode), shared by all batches: a batch keeps only its
  // captured cudaGraph_t, and running a batch other than the one the executable was last set up for is a
  // cudaGraphExecUpdate (same topology, new pointers / grids: a few ms) instead of an instantiation (10-140 ms) plus
  // the first-launch upload.  `owner` is the batch the executable currently holds the parameters of.
  struct LoopExec {
    cudaGraphExec_t exec = nullptr;
    const vv_batch* owner = nullptr;
  };
  std::map<std::pair<int, int>, LoopExec> loop_exec;
};

struct vv_batch {
  vv_engine* e = nullptr;
  int B = 0;
  std::vector<int> T, ref_len;
  std::vector<int> seq_off, seq_len;  // 2B sequences (cond then uncond)
  int R = 0, M = 0, maxT = 0;
  std::vector<void*> allocs;
  std::vector<bool> prepped;
  bool committed = false, decoded = false;
  int steps_done = 0;
  // device index arrays
  int32_t *seq_off_d = nullptr, *seq_len_d = nullptr, *row_pos_d = nullptr, *row_len_d = nullptr, *row_seq_d = nullptr;
  uint8_t* row_mask_d = nullptr;
  int32_t *tile_seq_d = nullptr, *tile_q0_d = nullptr;
  int n_tiles = 0;
  int32_t* ids_d = nullptr;
  int32_t* ids_h = nullptr;             // pinned staging for the id upload
  size_t ids_h_bytes = 0;
  float* noise0 = nullptr;              // y0 as preprocessed (restored by vv_run_resident)
  std::vector<std::shared_ptr<Prompt>> prompt;   // per chunk: the resident prompt its mel rows come from
  std::vector<cudaEvent_t> ids_ev;               // per chunk: the H2D copy that last read this chunk's id staging
  std::vector<int> dec_ref_len;         // ref_len snapshot the cached decode layout was built for
  std::vector<GemmOp> dec_ops;
  // DiT buffers
  float *noise = nullptr, *mel = nullptr, *cond_proj = nullptr, *x = nullptr, *x0 = nullptr, *v = nullptr,
        *cat_f32 = nullptr;
  bf16 *noise_b = nullptr, *cat_b = nullptr, *x0b = nullptr, *h1b = nullptr, *hb = nullptr, *qkv = nullptr,
       *attn_o = nullptr, *ffb = nullptr;
  // text buffers
  float *tx = nullptr, *t_tmp = nullptr, *t_ff = nullptr, *gx2 = nullptr, *nx = nullptr;
  bf16 *t_hb = nullptr, *t_ffb = nullptr;
  // decode buffers
  int Rd_max = 0, Rd = 0;
  std::vector<int> dec_off, dec_len;
  std::vector<int64_t> pcm_off, pcm_len;
  int32_t *d_src_row = nullptr, *d_row_pos = nullptr, *d_row_len = nullptr;
  int32_t *d_dec_off = nullptr, *d_dec_len = nullptr;
  int64_t* d_pcm_off = nullptr;
  bf16 *v_emb = nullptr, *v_hb = nullptr, *v_ffb = nullptr;
  float *vx = nullptr, *v_tmp = nullptr, *v_head = nullptr;
  int ld_head = 0;
  int16_t* pcm_d = nullptr;
  int64_t pcm_total = 0;
  // ops
  GemmOp op_in, op_cond, op_c1, op_c2, op_fin;
  std::vector<GemmOp> op_qkv, op_out, op_ff1, op_ff2, op_tpw1, op_tpw2, op_vpw1, op_vpw2;
  GemmOp op_vemb, op_vhead;
  CUtensorMap tQKV;
  std::map<int, cudaGraph_t> graphs;       // captured sampling loop per nfe (the executable lives in the engine)
  std::map<int, int64_t> graph_launches;
};

// ------------------------------------------------------------------------------------------------ helpers
template <typename T>
static int dev_alloc(std::vector<void*>& list, T** out, size_t count, bool zero = true) {
  void* p = nullptr;
  size_t bytes = std::max<size_t>(count * sizeof(T), 256);
  CK(cudaMalloc(&p, bytes));
  if (zero) CK(cudaMemset(p, 0, bytes));
  list.push_back(p);
  *out = reinterpret_cast<T*>(p);
  return 0;
}
// Batch buffers come from the device's stream-ordered memory pool (its release threshold is raised at engine creation
// so that freed blocks stay cached): a request stream builds and drops a ~1 GB batch for nearly every micro-batch, and
// cudaMalloc / cudaFree of that (a device-wide synchronisation each) cost 0.1-2 s per new shape — more than the
// synthesis itself.
template <typename T>
static int pool_alloc(std::vector<void*>& list, T** out, size_t count, cudaStream_t st) {
  void* p = nullptr;
  size_t bytes = std::max<size_t>(count * sizeof(T), 256);
  CK(cudaMallocAsync(&p, bytes, st));
  list.push_back(p);
  *out = reinterpret_cast<T*>(p);
  CK(cudaMemsetAsync(p, 0, bytes, st));
  return 0;
}
#define ENG_LOCK(e) std::lock_guard<std::recursive_mutex> _eng_lock((e)->mu)
#define TRY(x)            \
  do {                    \
    int _r = (x);         \
    if (_r != 0) return _r; \
  } while (0)

static const WTensor* find_w(const vv_engine* e, const std::string& n) {
  auto it = e->w.find(n);
  return it == e->w.end() ? nullptr : &it->second;
}
static int need_w(const vv_engine* e, const std::string& n, const float** out, size_t numel) {
  const WTensor* t = find_w(e, n);
  if (!t) return fail(VV_ERR_FORMAT, "weight tensor '%s' missing from loaded blobs", n.c_str());
  if (numel && t->numel != numel)
    return fail(VV_ERR_FORMAT, "weight tensor '%s' has %zu elements, expected %zu", n.c_str(), t->numel, numel);
  *out = t->d;
  return 0;
}

static int pick_bn(int M, int N, int sms) {
  if (N <= 64) return 64;
  if (N <= 128) return 128;
  const long mt = (M + 127) / 128;
  const long t128 = mt * ((N + 127) / 128), t256 = mt * ((N + 255) / 256);
  const double c128 = (double)((t128 + sms - 1) / sms) * 1.0;
  const double c256 = (double)((t256 + sms - 1) / sms) * 2.0 * 0.85;
  return c256 <= c128 ? 256 : 128;
}

// VVB200_GEMM_PAIR=0 keeps every GEMM on the 1-CTA kernel (A/B experiments)
static bool pair_enabled() {
  static const bool on = [] {
    const char* v = getenv("VVB200_GEMM_PAIR");
    return !(v && v[0] == '0');
  }();
  return on;
}
// large GEMMs whose N is a multiple of 256 run on the CTA-pair kernel (bn code 512)
static int pick_tile(int M, int N, int K, int sms) {
  if (pair_enabled() && N % 256 == 0 && K % 64 == 0 && M >= 1024) return 512;
  return pick_bn(M, N, sms);
}

static GemmOp make_op(vv_engine* e, const bf16* A, int lda, int a_rows, int M, const bf16* Bw, int ldb, int N, int K) {
  GemmOp op;
  op.s.M = M; op.s.N = N; op.s.K = K;
  op.bn = pick_tile(M, N, K, e->num_sms);
  op.tA = make_tmap_bf16(A, a_rows, K, lda, 128);
  op.tB = make_tmap_bf16(Bw, N, K, ldb, op.bn == 512 ? 128 : op.bn);
  if (op.bn == 512) op.tBt = make_tmap_bf16(Bw, N, K, ldb, 32);
  return op;
}
static GemmOp make_conv_op(vv_engine* e, const bf16* X, int ldx, int a_rows, int M, const bf16* Wt, int groups, int taps) {
  GemmOp op;
  op.s.M = M; op.s.N = groups * 64; op.s.K = taps * 64;
  op.s.conv_taps = taps; op.s.conv_groups = groups;
  op.bn = 64;
  op.tA = make_tmap_bf16(X, a_rows, groups * 64, ldx, 128);
  op.tBt = make_tmap_bf16(X, a_rows, groups * 64, ldx, 32);
  op.tB = make_tmap_bf16(Wt, (uint64_t)groups * taps * 64, 64, 64, 64);
  return op;
}
static inline void run_gemm(vv_engine* e, const GemmOp& op, const GemmEpi& epi) {
  if (op.bn == 512 && !gemm_pair_supported(op.s, epi)) {
    fprintf(stderr, "vvb200: GEMM planned for the CTA-pair kernel has an unsupported epilogue/shape\n");
    abort();
  }
  if (op.s.conv_taps > 0 && conv_pos_supported(op.s, epi))
    launch_conv_pos(op.tA, op.tBt, op.tB, op.s, epi, e->num_sms, e->st);
  else
    launch_gemm(op.tA, op.tB, op.s, epi, op.bn, e->num_sms, e->st, op.bn == 512 ? &op.tBt : nullptr);
  e->launches++;
}

// ------------------------------------------------------------------------------------------------ C ABI: misc
extern "C" const char* vv_last_error(void) { return g_err.c_str(); }
extern "C" int vv_version(void) { return 100; }
extern "C" int vv_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

static int arch_check(const vv_arch& a) {
  if (a.heads * a.head_dim != a.dim || a.head_dim != 64) return fail(VV_ERR_ARG, "arch: need head_dim 64 and heads*64 == dim");
  if (a.dim / a.conv_pos_groups != 64) return fail(VV_ERR_ARG, "arch: conv_pos needs 64 channels per group");
  if (a.n_fft != 1024 || a.hop != 256) return fail(VV_ERR_ARG, "arch: mel/iSTFT kernels need n_fft 1024, hop 256");
  const int dims[] = {a.dim, a.ff_dim, a.text_dim, a.text_ff, a.voc_dim, a.voc_ff};
  for (int d : dims)
    if (d % 64) return fail(VV_ERR_ARG, "arch: layer widths must be multiples of 64");
  const int lnd[] = {a.dim, a.text_dim, a.voc_dim};
  for (int d : lnd)
    if (d % 128 || d > 2048) return fail(VV_ERR_ARG, "arch: LayerNorm widths must be multiples of 128, <= 2048");
  if (a.n_mel > 128 || a.rope_heads < 0 || a.rope_heads > a.heads || a.nfe < 2) return fail(VV_ERR_ARG, "arch: bad n_mel/rope_heads/nfe");
  return 0;
}

static int build_tables(vv_engine* e) {
  const vv_arch& a = e->a;
  // ---- tables (host double -> float, as the oracle does)
  {
    std::vector<float2> cs((size_t)e->rope_max * 32);
    for (int p = 0; p < e->rope_max; ++p)
      for (int k = 0; k < 32; ++k) {
        const double inv = 1.0 / pow((double)a.rope_theta, (double)(2 * k) / (double)a.head_dim);
        cs[(size_t)p * 32 + k] = make_float2((float)cos(p * inv), (float)sin(p * inv));
      }
    TRY(dev_alloc(e->allocs, &e->rope_cs, cs.size()));
    CK(cudaMemcpy(e->rope_cs, cs.data(), cs.size() * sizeof(float2), cudaMemcpyHostToDevice));
    std::vector<float2> tw(512);
    for (int k = 0; k < 512; ++k) tw[k] = make_float2((float)cos(-2.0 * M_PI * k / 1024.0), (float)sin(-2.0 * M_PI * k / 1024.0));
    TRY(dev_alloc(e->allocs, &e->fft_tw, tw.size()));
    CK(cudaMemcpy(e->fft_tw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
    std::vector<float> hn(1024);
    for (int n = 0; n < 1024; ++n) hn[n] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * n / 1024.0));
    TRY(dev_alloc(e->allocs, &e->hann, hn.size()));
    CK(cudaMemcpy(e->hann, hn.data(), hn.size() * 4, cudaMemcpyHostToDevice));
    const int td = a.text_dim, hd = td / 2;
    std::vector<float> pt((size_t)a.pos_table_len * td);
    for (int p = 0; p < a.pos_table_len; ++p)
      for (int k = 0; k < hd; ++k) {
        const double f = 1.0 / pow(10000.0, (double)(2 * k) / (double)td);
        pt[(size_t)p * td + k] = (float)cos(p * f);
        pt[(size_t)p * td + hd + k] = (float)sin(p * f);
      }
    TRY(dev_alloc(e->allocs, &e->pos_table, pt.size()));
    CK(cudaMemcpy(e->pos_table, pt.data(), pt.size() * 4, cudaMemcpyHostToDevice));
  }
  return 0;
}

extern "C" int vv_engine_create(const vv_arch* arch, int device, void* stream, vv_engine** out) {
  if (!arch || !out) return fail(VV_ERR_ARG, "vv_engine_create: null argument");
  TRY(arch_check(*arch));
  int n = vv_device_count();
  if (n <= 0) return fail(VV_ERR_CUDA, "no CUDA device available (this engine has no CPU fallback)");
  if (device < 0 || device >= n) return fail(VV_ERR_ARG, "device %d out of range (%d devices)", device, n);
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(VV_ERR_CUDA, "device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major, prop.minor);
  vv_engine* e = new vv_engine();
  e->a = *arch;
  e->device = device;
  e->num_sms = prop.multiProcessorCount;
  {  // keep freed batch memory cached in the stream-ordered pool instead of returning it to the driver at every sync
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      uint64_t keep = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    (void)cudaGetLastError();
  }
  if (stream) {
    e->st = reinterpret_cast<cudaStream_t>(stream);
  } else {
    cudaError_t r = cudaStreamCreateWithFlags(&e->st, cudaStreamNonBlocking);
    if (r != cudaSuccess) {
      delete e;
      return fail(VV_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(r));
    }
    e->own_stream = true;
  }
  {
    int r = build_tables(e);
    if (r) {
      vv_engine_destroy(e);
      return r;
    }
  }
  *out = e;
  return 0;
}

extern "C" void vv_engine_destroy(vv_engine* e);
extern "C" void* vv_engine_stream(const vv_engine* e) { return e ? (void*)e->st : nullptr; }
extern "C" int64_t vv_engine_launch_count(const vv_engine* e) { return e ? e->launches : 0; }
extern "C" int vv_sync(vv_engine* e) {
  if (!e) return fail(VV_ERR_ARG, "null engine");
  CK(cudaStreamSynchronize(e->st));
  return 0;
}

struct BlobEntry {
  char name[96];
  uint32_t dtype, ndim;
  int64_t shape[4];
  uint64_t offset, nbytes;
};
static_assert(sizeof(BlobEntry) == 152, "blob entry layout");

extern "C" int vv_engine_load_blob(vv_engine* e, const void* blob, size_t nbytes) {
  if (!e || !blob) return fail(VV_ERR_ARG, "vv_engine_load_blob: null argument");
  ENG_LOCK(e);
  if (e->finalized) return fail(VV_ERR_STATE, "engine already finalized");
  const uint8_t* p = static_cast<const uint8_t*>(blob);
  if (nbytes < 256 || memcmp(p, "VVB200W1", 8) != 0) return fail(VV_ERR_FORMAT, "not a VVB200 weight blob (bad magic)");
  uint32_t n;
  memcpy(&n, p + 8, 4);
  vv_arch ba;
  memcpy(&ba, p + 16, sizeof(vv_arch));
  if (memcmp(&ba, &e->a, sizeof(vv_arch)) != 0) return fail(VV_ERR_FORMAT, "blob architecture differs from the engine's");
  if (256 + (size_t)n * sizeof(BlobEntry) > nbytes) return fail(VV_ERR_FORMAT, "blob truncated (entry table)");
  CK(cudaSetDevice(e->device));
  for (uint32_t i = 0; i < n; ++i) {
    BlobEntry en;
    memcpy(&en, p + 256 + (size_t)i * sizeof(BlobEntry), sizeof(BlobEntry));
    en.name[95] = 0;
    // offset / size are untrusted: the comparison must not wrap in uint64
    if (en.dtype != 0 || en.ndim > 4 || en.offset > nbytes || en.nbytes > nbytes - en.offset || (en.nbytes & 3))
      return fail(VV_ERR_FORMAT, "blob entry '%s' malformed", en.name);
    uint64_t prod = 1;
    for (uint32_t k = 0; k < en.ndim; ++k) {
      if (en.shape[k] <= 0 || (uint64_t)en.shape[k] > (1ull << 32) || prod > (1ull << 40))
        return fail(VV_ERR_FORMAT, "blob entry '%s': bad shape", en.name);
      prod *= (uint64_t)en.shape[k];
    }
    if (prod * 4 != en.nbytes)
      return fail(VV_ERR_FORMAT, "blob entry '%s': shape holds %llu elements, payload %llu bytes", en.name,
                  (unsigned long long)prod, (unsigned long long)en.nbytes);
    if (e->w.count(en.name)) return fail(VV_ERR_FORMAT, "blob entry '%s' loaded twice", en.name);
    WTensor t;
    t.ndim = en.ndim;
    t.numel = en.nbytes / 4;
    for (int k = 0; k < 4; ++k) t.shape[k] = en.shape[k];
    TRY(dev_alloc(e->allocs, &t.d, t.numel, false));
    CK(cudaMemcpy(t.d, p + en.offset, en.nbytes, cudaMemcpyHostToDevice));
    e->w[en.name] = t;
  }
  return 0;
}

// fp32 [rows, cols] (row stride ld_src, starting at column col0) -> bf16 [rows, kpad] zero padded
static int to_bf16(vv_engine* e, const float* src, int rows, int cols, int ld_src, int col0, int kpad, bf16** out) {
  TRY(dev_alloc(e->allocs, out, (size_t)rows * kpad));
  launch_f32_to_bf16_2d(src + col0, rows, cols, ld_src, *out, kpad, kpad, e->st);
  e->launches++;
  return 0;
}

static int load_convnext(vv_engine* e, const std::string& p, int d, int ff, bool grn, ConvNextW* cw, int k) {
  const float *pw1 = nullptr, *pw2 = nullptr;
  TRY(need_w(e, p + ".dw.w", &cw->dw_w, (size_t)d * k));
  TRY(need_w(e, p + ".dw.b", &cw->dw_b, d));
  TRY(need_w(e, p + ".ln.g", &cw->ln_g, d));
  TRY(need_w(e, p + ".ln.b", &cw->ln_b, d));
  TRY(need_w(e, p + ".pw1.w", &pw1, (size_t)ff * d));
  TRY(need_w(e, p + ".pw1.b", &cw->pw1_b, ff));
  TRY(need_w(e, p + ".pw2.w", &pw2, (size_t)ff * d));
  TRY(need_w(e, p + ".pw2.b", &cw->pw2_b, d));
  cw->grn_g = cw->grn_b = cw->gamma = nullptr;
  if (grn) {
    TRY(need_w(e, p + ".grn.g", &cw->grn_g, ff));
    TRY(need_w(e, p + ".grn.b", &cw->grn_b, ff));
  } else {
    TRY(need_w(e, p + ".gamma", &cw->gamma, d));
  }
  TRY(to_bf16(e, pw1, ff, d, d, 0, d, &cw->pw1));
  TRY(to_bf16(e, pw2, d, ff, ff, 0, ff, &cw->pw2));
  return 0;
}

static constexpr int VV_MAX_NFE = 256;

static void destroy_batch_graph(vv_batch* b, int nfe);

// Drops the least recently used modulation table other than `keep_nfe`.  Captured sampling-loop graphs hold pointers
// into the table, so every batch's graph for that nfe and the shared executables go with it.
static bool evict_mod_table(vv_engine* e, int keep_nfe) {
  auto victim = e->mod.end();
  for (auto it = e->mod.begin(); it != e->mod.end(); ++it)
    if (it->first != keep_nfe && it->first != e->a.nfe &&
        (victim == e->mod.end() || it->second.last_use < victim->second.last_use))
      victim = it;
  if (victim == e->mod.end()) return false;
  const int nfe = victim->first;
  cudaStreamSynchronize(e->st);
  for (vv_batch* b : e->live) destroy_batch_graph(b, nfe);
  for (auto le = e->loop_exec.begin(); le != e->loop_exec.end();) {
    if (le->first.first == nfe) {
      if (le->second.exec) cudaGraphExecDestroy(le->second.exec);
      le = e->loop_exec.erase(le);
    } else {
      ++le;
    }
  }
  cudaFree(victim->second.blocks);
  cudaFree(victim->second.fin);
  e->mod.erase(victim);
  return true;
}

static int build_mod_table(vv_engine* e, int nfe, ModTable** out) {
  auto it = e->mod.find(nfe);
  if (it != e->mod.end()) {
    it->second.last_use = ++e->mod_tick;
    *out = &it->second;
    return 0;
  }
  if (nfe < 2 || nfe > VV_MAX_NFE) return fail(VV_ERR_ARG, "nfe %d out of range [2, %d]", nfe, VV_MAX_NFE);
  static const size_t cap = [] {
    const char* v = getenv("VVB200_MOD_TABLES");
    const int n = v ? atoi(v) : 4;
    return (size_t)(n < 1 ? 1 : n);
  }();
  while (e->mod.size() >= cap && evict_mod_table(e, nfe)) {}
  const vv_arch& a = e->a;
  ModTable mt;
  mt.nfe = nfe;
  mt.steps = nfe - 1;
  const int S = mt.steps, half = a.time_freq_dim / 2;
  std::vector<double> t(nfe);
  for (int i = 0; i < nfe; ++i) {
    double ti = nfe > 1 ? (double)i / (double)(nfe - 1) : 0.0;
    t[i] = ti + (double)a.sway * (cos(M_PI / 2.0 * ti) - 1.0 + ti);
  }
  mt.dt.resize(S);
  std::vector<float> emb((size_t)S * a.time_freq_dim);
  for (int i = 0; i < S; ++i) {
    mt.dt[i] = (float)(t[i + 1] - t[i]);
    for (int k = 0; k < half; ++k) {
      const double f = exp((double)k * (-log(10000.0) / (double)(half - 1)));
      const double ang = 1000.0 * t[i] * f;
      emb[(size_t)i * a.time_freq_dim + k] = (float)sin(ang);
      emb[(size_t)i * a.time_freq_dim + half + k] = (float)cos(ang);
    }
  }
  // the table owns `blocks` / `fin` (freed on eviction); the three scratch buffers are released below
  std::vector<void*> own, scratch;
  struct Guard {
    std::vector<void*>&own, &scratch;
    bool keep = false;
    ~Guard() {
      for (void* q : scratch) cudaFree(q);
      if (!keep)
        for (void* q : own) cudaFree(q);
    }
  } guard{own, scratch};
  float *emb_d, *h1, *st;
  TRY(dev_alloc(scratch, &emb_d, emb.size()));
  TRY(dev_alloc(scratch, &h1, (size_t)S * a.dim));
  TRY(dev_alloc(scratch, &st, (size_t)S * a.dim));
  TRY(dev_alloc(own, &mt.blocks, (size_t)S * a.depth * 6 * a.dim));
  TRY(dev_alloc(own, &mt.fin, (size_t)S * 2 * a.dim));
  CK(cudaMemcpyAsync(emb_d, emb.data(), emb.size() * 4, cudaMemcpyHostToDevice, e->st));
  const float *w1 = nullptr, *b1 = nullptr, *w2 = nullptr, *b2 = nullptr;
  TRY(need_w(e, "dit.time.l1.w", &w1, (size_t)a.dim * a.time_freq_dim));
  TRY(need_w(e, "dit.time.l1.b", &b1, a.dim));
  TRY(need_w(e, "dit.time.l2.w", &w2, (size_t)a.dim * a.dim));
  TRY(need_w(e, "dit.time.l2.b", &b2, a.dim));
  launch_linear_f32(emb_d, w1, b1, h1, S, a.time_freq_dim, a.dim, a.dim, 1, e->st);
  launch_linear_f32(h1, w2, b2, st, S, a.dim, a.dim, a.dim, 1, e->st);  // st = SiLU(temb)
  e->launches += 2;
  for (int l = 0; l < a.depth; ++l) {
    const float *aw = nullptr, *ab = nullptr;
    const std::string p = "dit.blocks." + std::to_string(l) + ".ada";
    TRY(need_w(e, p + ".w", &aw, (size_t)6 * a.dim * a.dim));
    TRY(need_w(e, p + ".b", &ab, (size_t)6 * a.dim));
    launch_linear_f32(st, aw, ab, mt.blocks + (size_t)l * 6 * a.dim, S, a.dim, 6 * a.dim, a.depth * 6 * a.dim, 0, e->st);
    e->launches++;
  }
  const float *fw = nullptr, *fb = nullptr;
  TRY(need_w(e, "dit.final.ada.w", &fw, (size_t)2 * a.dim * a.dim));
  TRY(need_w(e, "dit.final.ada.b", &fb, (size_t)2 * a.dim));
  launch_linear_f32(st, fw, fb, mt.fin, S, a.dim, 2 * a.dim, 2 * a.dim, 0, e->st);
  e->launches++;
  CK(cudaStreamSynchronize(e->st));
  CKL();
  guard.keep = true;
  mt.last_use = ++e->mod_tick;
  e->mod[nfe] = mt;
  *out = &e->mod[nfe];
  return 0;
}

extern "C" int vv_engine_finalize(vv_engine* e) {
  if (!e) return fail(VV_ERR_ARG, "null engine");
  ENG_LOCK(e);
  if (e->finalized) return 0;
  const vv_arch& a = e->a;
  CK(cudaSetDevice(e->device));
  // ---- bf16 GEMM operands
  const int d = a.dim;
  e->Kn = 128;
  e->Kc = round_up(a.n_mel + a.text_dim, 64);
  e->Kemb = round_up(a.voc_k * a.n_mel, 64);
  const int in_dim = 2 * a.n_mel + a.text_dim;
  const float* wp = nullptr;
  TRY(need_w(e, "dit.in.w", &wp, (size_t)d * in_dim));
  TRY(to_bf16(e, wp, d, a.n_mel, in_dim, 0, e->Kn, &e->in_n));
  TRY(to_bf16(e, wp, d, a.n_mel + a.text_dim, in_dim, a.n_mel, e->Kc, &e->in_c));
  const int cg = 64;
  for (int c = 0; c < 2; ++c) {
    TRY(need_w(e, c == 0 ? "dit.pos.c1.w" : "dit.pos.c2.w", &wp, (size_t)d * cg * a.conv_pos_k));
    bf16** dst = c == 0 ? &e->c1 : &e->c2;
    TRY(dev_alloc(e->allocs, dst, (size_t)d * cg * a.conv_pos_k));
    launch_permute_conv_w(wp, d, cg, a.conv_pos_k, *dst, e->st);
    e->launches++;
  }
  e->layers.resize(a.depth);
  for (int l = 0; l < a.depth; ++l) {
    const std::string p = "dit.blocks." + std::to_string(l);
    LayerW& L = e->layers[l];
    TRY(need_w(e, p + ".qkv.w", &wp, (size_t)3 * d * d));
    TRY(to_bf16(e, wp, 3 * d, d, d, 0, d, &L.qkv));
    TRY(need_w(e, p + ".out.w", &wp, (size_t)d * d));
    TRY(to_bf16(e, wp, d, d, d, 0, d, &L.out));
    TRY(need_w(e, p + ".ff1.w", &wp, (size_t)a.ff_dim * d));
    TRY(to_bf16(e, wp, a.ff_dim, d, d, 0, d, &L.ff1));
    TRY(need_w(e, p + ".ff2.w", &wp, (size_t)a.ff_dim * d));
    TRY(to_bf16(e, wp, d, a.ff_dim, a.ff_dim, 0, a.ff_dim, &L.ff2));
    TRY(need_w(e, p + ".qkv.b", &L.qkv_b, (size_t)3 * d));
    TRY(need_w(e, p + ".out.b", &L.out_b, d));
    TRY(need_w(e, p + ".ff1.b", &L.ff1_b, a.ff_dim));
    TRY(need_w(e, p + ".ff2.b", &L.ff2_b, d));
  }
  TRY(need_w(e, "dit.out.w", &wp, (size_t)a.n_mel * d));
  TRY(to_bf16(e, wp, a.n_mel, d, d, 0, d, &e->out_w));
  e->text_blocks.resize(a.text_layers);
  for (int i = 0; i < a.text_layers; ++i)
    TRY(load_convnext(e, "pre.text_blocks." + std::to_string(i), a.text_dim, a.text_ff, true, &e->text_blocks[i], 7));
  e->voc_blocks.resize(a.voc_layers);
  for (int i = 0; i < a.voc_layers; ++i)
    TRY(load_convnext(e, "voc.blocks." + std::to_string(i), a.voc_dim, a.voc_ff, false, &e->voc_blocks[i], a.voc_k));
  TRY(need_w(e, "voc.embed.w", &wp, (size_t)a.voc_dim * a.n_mel * a.voc_k));
  TRY(dev_alloc(e->allocs, &e->voc_embed, (size_t)a.voc_dim * e->Kemb));
  launch_permute_embed_w(wp, a.voc_dim, a.n_mel, a.voc_k, e->Kemb, e->voc_embed, e->st);
  e->launches++;
  TRY(need_w(e, "voc.head.w", &wp, (size_t)(a.n_fft + 2) * a.voc_dim));
  TRY(to_bf16(e, wp, a.n_fft + 2, a.voc_dim, a.voc_dim, 0, a.voc_dim, &e->voc_head));
  // presence checks for the remaining fp32 tensors
  const char* req[] = {"pre.mel_fb", "pre.text_embed", "dit.in.b", "dit.pos.c1.b", "dit.pos.c2.b", "dit.out.b",
                       "voc.embed.b", "voc.norm.g", "voc.norm.b", "voc.final.g", "voc.final.b", "voc.head.b"};
  for (const char* n : req) TRY(need_w(e, n, &wp, 0));
  CK(cudaStreamSynchronize(e->st));
  CKL();
  ModTable* mt;
  TRY(build_mod_table(e, a.nfe, &mt));
  e->finalized = true;
  return 0;
}

extern "C" void vv_engine_destroy(vv_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->st);
  {
    ENG_LOCK(e);
    {
      std::vector<vv_batch*> cached;
      for (auto& kv : e->batch_cache) cached.push_back(kv.second);
      for (vv_batch* cb : cached) vv_batch_destroy(cb);     // each removes itself from the cache
    }
    e->batch_cache.clear();
    for (vv_batch* b : e->live) b->e = nullptr;     // caller-owned batches outliving the engine: destroy is a no-op
    e->live.clear();
    e->prompts.clear();
    for (auto& m : e->mod) {
      cudaFree(m.second.blocks);
      cudaFree(m.second.fin);
    }
  }
  cudaStreamSynchronize(e->st);
  for (void* p : e->allocs) cudaFree(p);
  for (auto& pf : e->pinned_free) cudaFreeHost(pf.second);
  for (auto& le : e->loop_exec)
    if (le.second.exec) cudaGraphExecDestroy(le.second.exec);
  if (e->own_stream) cudaStreamDestroy(e->st);
  delete e;
}

// ------------------------------------------------------------------------------------------------ batch
extern "C" int vv_batch_create(vv_engine* e, int B, const int64_t* total_frames, vv_batch** out) {
  if (!e || !total_frames || !out || B <= 0) return fail(VV_ERR_ARG, "vv_batch_create: bad argument");
  ENG_LOCK(e);
  if (!e->finalized) return fail(VV_ERR_STATE, "engine not finalized");
  const vv_arch& a = e->a;
  CK(cudaSetDevice(e->device));
  vv_batch* b = new vv_batch();
  b->e = e;
  b->B = B;
  e->live.insert(b);
  const int gap = 16;  // >= conv_pos_k/2 zero rows between sequences
  if (a.conv_pos_k / 2 > gap) {
    vv_batch_destroy(b);
    return fail(VV_ERR_ARG, "conv_pos_k too large for the row layout");
  }
  b->T.resize(B);
  b->ref_len.assign(B, 0);
  b->prepped.assign(B, false);
  b->seq_off.resize(2 * B);
  b->seq_len.resize(2 * B);
  int R = 0;
  for (int i = 0; i < B; ++i) {
    if (total_frames[i] < 2 || total_frames[i] > e->rope_max) {
      vv_batch_destroy(b);
      return fail(VV_ERR_ARG, "total_frames[%d] = %lld out of range [2, %d]", i, (long long)total_frames[i], e->rope_max);
    }
    b->T[i] = (int)total_frames[i];
    b->maxT = std::max(b->maxT, b->T[i]);
    b->seq_off[i] = R;
    b->seq_len[i] = b->T[i];
    R += b->T[i] + gap;
  }
  R = round_up(R, 8);
  b->R = R;
  b->M = 2 * R;
  for (int i = 0; i < B; ++i) {
    b->seq_off[B + i] = R + b->seq_off[i];
    b->seq_len[B + i] = b->T[i];
  }
  const int M = b->M, d = a.dim;
  std::vector<int32_t> row_pos(M, 0), row_len(M, 0), row_seq(M, -1);
  std::vector<uint8_t> row_mask(M, 0);
  std::vector<int32_t> tile_seq, tile_q0;
  for (int s = 0; s < 2 * B; ++s) {
    for (int p = 0; p < b->seq_len[s]; ++p) {
      const int r = b->seq_off[s] + p;
      row_pos[r] = p; row_len[r] = b->seq_len[s]; row_seq[r] = s; row_mask[r] = 1;
    }
    for (int q0 = 0; q0 < b->seq_len[s]; q0 += attn_q_tile()) {
      tile_seq.push_back(s);
      tile_q0.push_back(q0);
    }
  }
  b->n_tiles = (int)tile_seq.size();
  auto& AL = b->allocs;
#define UP(dst, vec)                                                                              \
  do {                                                                                            \
    int _r = pool_alloc(AL, &dst, vec.size(), e->st);                                             \
    if (_r) { vv_batch_destroy(b); return _r; }                                                   \
    cudaMemcpyAsync(dst, vec.data(), vec.size() * sizeof(vec[0]), cudaMemcpyHostToDevice, e->st); \
  } while (0)
#define AB(ptr, count)                                                                            \
  do {                                                                                            \
    int _r = pool_alloc(AL, &ptr, (size_t)(count), e->st);                                        \
    if (_r) { vv_batch_destroy(b); return _r; }                                                   \
  } while (0)
  UP(b->seq_off_d, b->seq_off);
  UP(b->seq_len_d, b->seq_len);
  UP(b->row_pos_d, row_pos);
  UP(b->row_len_d, row_len);
  UP(b->row_seq_d, row_seq);
  UP(b->row_mask_d, row_mask);
  UP(b->tile_seq_d, tile_seq);
  UP(b->tile_q0_d, tile_q0);
  AB(b->ids_d, M);
  AB(b->noise0, (size_t)R * a.n_mel);
  {  // pinned staging buffer for the text ids: recycled through the engine (cudaMallocHost costs milliseconds)
    const size_t need = (size_t)R * 4;
    auto best = e->pinned_free.end();
    for (auto it = e->pinned_free.begin(); it != e->pinned_free.end(); ++it)
      if (it->first >= need && (best == e->pinned_free.end() || it->first < best->first)) best = it;
    if (best != e->pinned_free.end()) {
      b->ids_h = reinterpret_cast<int32_t*>(best->second);
      b->ids_h_bytes = best->first;
      e->pinned_free.erase(best);
    } else {
      const size_t bytes = std::max<size_t>(need * 2, 1 << 16);      // head room: the next shape is rarely smaller
      if (cudaMallocHost(&b->ids_h, bytes) != cudaSuccess) {
        vv_batch_destroy(b);
        return fail(VV_ERR_CUDA, "cudaMallocHost failed");
      }
      b->ids_h_bytes = bytes;
    }
  }
  b->prompt.assign(B, nullptr);
  b->ids_ev.assign(B, nullptr);
  AB(b->noise, (size_t)R * a.n_mel);
  AB(b->mel, (size_t)R * a.n_mel);
  AB(b->cond_proj, (size_t)M * d);
  AB(b->x, (size_t)M * d);
  AB(b->x0, (size_t)M * d);
  AB(b->v, (size_t)M * 128);
  AB(b->cat_f32, (size_t)M * e->Kc);
  AB(b->noise_b, (size_t)M * e->Kn);
  AB(b->cat_b, (size_t)M * e->Kc);
  AB(b->x0b, (size_t)M * d);
  AB(b->h1b, (size_t)M * d);
  AB(b->hb, (size_t)M * d);
  AB(b->qkv, (size_t)M * 3 * d);
  AB(b->attn_o, (size_t)M * d);
  AB(b->ffb, (size_t)M * a.ff_dim);
  AB(b->tx, (size_t)M * a.text_dim);
  AB(b->t_tmp, (size_t)M * a.text_dim);
  AB(b->t_ff, (size_t)M * a.text_ff);
  AB(b->gx2, (size_t)2 * B * ((b->maxT + 63) / 64) * a.text_ff);
  AB(b->nx, (size_t)2 * B * a.text_ff);
  AB(b->t_hb, (size_t)M * a.text_dim);
  AB(b->t_ffb, (size_t)M * a.text_ff);
  // decode
  int rd = 0;
  for (int i = 0; i < B; ++i) rd += b->T[i];
  b->Rd_max = round_up(rd, 8);
  b->ld_head = round_up(a.n_fft + 2, 4);
  AB(b->d_src_row, b->Rd_max);
  AB(b->d_row_pos, b->Rd_max);
  AB(b->d_row_len, b->Rd_max);
  AB(b->v_emb, (size_t)b->Rd_max * e->Kemb);
  AB(b->v_hb, (size_t)b->Rd_max * a.voc_dim);
  AB(b->v_ffb, (size_t)b->Rd_max * a.voc_ff);
  AB(b->vx, (size_t)b->Rd_max * a.voc_dim);
  AB(b->v_tmp, (size_t)b->Rd_max * a.voc_dim);
  AB(b->v_head, (size_t)b->Rd_max * b->ld_head);
  AB(b->d_dec_off, B);
  AB(b->d_dec_len, B);
  AB(b->d_pcm_off, B);
  AB(b->pcm_d, (size_t)b->Rd_max * a.hop);
  b->dec_off.assign(B, 0);
  b->dec_len.assign(B, 0);
  b->pcm_off.assign(B, 0);
  b->pcm_len.assign(B, 0);
#undef UP
#undef AB
  // ---- GEMM plans
  b->op_in = make_op(e, b->noise_b, e->Kn, M, M, e->in_n, e->Kn, d, e->Kn);
  b->op_cond = make_op(e, b->cat_b, e->Kc, M, M, e->in_c, e->Kc, d, e->Kc);
  b->op_c1 = make_conv_op(e, b->x0b, d, M, M, e->c1, a.conv_pos_groups, a.conv_pos_k);
  b->op_c2 = make_conv_op(e, b->h1b, d, M, M, e->c2, a.conv_pos_groups, a.conv_pos_k);
  b->op_fin = make_op(e, b->hb, d, M, M, e->out_w, d, a.n_mel, d);
  for (int l = 0; l < a.depth; ++l) {
    const LayerW& L = e->layers[l];
    b->op_qkv.push_back(make_op(e, b->hb, d, M, M, L.qkv, d, 3 * d, d));
    b->op_out.push_back(make_op(e, b->attn_o, d, M, M, L.out, d, d, d));
    b->op_ff1.push_back(make_op(e, b->hb, d, M, M, L.ff1, d, a.ff_dim, d));
    b->op_ff2.push_back(make_op(e, b->ffb, a.ff_dim, M, M, L.ff2, a.ff_dim, d, a.ff_dim));
  }
  for (int i = 0; i < a.text_layers; ++i) {
    b->op_tpw1.push_back(make_op(e, b->t_hb, a.text_dim, M, M, e->text_blocks[i].pw1, a.text_dim, a.text_ff, a.text_dim));
    b->op_tpw2.push_back(make_op(e, b->t_ffb, a.text_ff, M, M, e->text_blocks[i].pw2, a.text_ff, a.text_dim, a.text_ff));
  }
  b->tQKV = make_tmap_bf16(b->qkv, M, 3 * d, 3 * d, 128);
  *out = b;
  return 0;
}

static void destroy_batch_graph(vv_batch* b, int nfe) {
  auto it = b->graphs.find(nfe);
  if (it == b->graphs.end()) return;
  cudaGraphDestroy(it->second);
  b->graphs.erase(it);
  b->graph_launches.erase(nfe);
}

extern "C" void vv_batch_destroy(vv_batch* b) {
  if (!b) return;
  vv_engine* e = b->e;
  if (!e) {            // the engine went first (vv_engine_destroy): its stream and pool are gone, only host state is left
    delete b;
    return;
  }
  // Python reaches this from Batch.__del__ / cache eviction on any thread while another thread may be inside
  // vv_synthesize_batch: every entry point that touches engine state takes the (recursive) engine lock
  ENG_LOCK(e);
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->st);
  for (auto& g : b->graphs) cudaGraphDestroy(g.second);
  for (auto& le : e->loop_exec)
    if (le.second.owner == b) le.second.owner = nullptr;     // the executable must be updated before its next launch
  for (void* p : b->allocs) cudaFreeAsync(p, e->st);        // back to the pool, no device-wide synchronisation
  for (cudaEvent_t ev : b->ids_ev)
    if (ev) cudaEventDestroy(ev);
  b->prompt.clear();                                         // resident prompts are shared: last owner frees them
  if (b->ids_h) e->pinned_free.emplace_back(b->ids_h_bytes, b->ids_h);   // pinned staging buffers are recycled
  e->live.erase(b);
  e->batch_last_use.erase(b);
  for (auto it = e->batch_cache.begin(); it != e->batch_cache.end(); ++it)
    if (it->second == b) {
      e->batch_cache.erase(it);
      break;
    }
  delete b;
}

// ------------------------------------------------------------------------------------------------ resident prompts
// Two independent 64-bit content hashes of the PCM (4 interleaved multiply-xorshift lanes each; ~10 GB/s on the host).
static void hash_pcm(const int16_t* audio, int64_t n, uint64_t* id, uint64_t* check) {
  const uint8_t* p = reinterpret_cast<const uint8_t*>(audio);
  const size_t bytes = (size_t)n * 2;
  uint64_t h[4] = {0x9E3779B97F4A7C15ull, 0xC2B2AE3D27D4EB4Full, 0x165667B19E3779F9ull, 0x27D4EB2F165667C5ull};
  uint64_t g[4] = {0xD6E8FEB86659FD93ull, 0xA0761D6478BD642Full, 0xE7037ED1A0B428DBull, 0x8EBC6AF09C88C6E3ull};
  size_t i = 0;
  for (; i + 32 <= bytes; i += 32) {
    uint64_t w[4];
    memcpy(w, p + i, 32);
    for (int k = 0; k < 4; ++k) {
      h[k] = (h[k] ^ w[k]) * 0x9FB21C651E98DF25ull;
      h[k] ^= h[k] >> 29;
      g[k] = (g[k] + w[k]) * 0xFF51AFD7ED558CCDull;
      g[k] ^= g[k] >> 31;
    }
  }
  uint64_t tail[4] = {0, 0, 0, 0};
  memcpy(tail, p + i, bytes - i);
  for (int k = 0; k < 4; ++k) {
    h[k] = (h[k] ^ tail[k]) * 0x9FB21C651E98DF25ull;
    g[k] = (g[k] + tail[k]) * 0xFF51AFD7ED558CCDull;
  }
  uint64_t a = (uint64_t)bytes * 0x9E3779B97F4A7C15ull, c = ~(uint64_t)bytes;
  for (int k = 0; k < 4; ++k) {
    a = (a ^ h[k]) * 0xC4CEB9FE1A85EC53ull;
    a ^= a >> 32;
    c = (c + g[k]) * 0xBF58476D1CE4E5B9ull;
    c ^= c >> 30;
  }
  *id = a ? a : 1;      // 0 means "no prompt id" in vv_request
  *check = c;
}

static size_t prompt_cache_cap() {
  static const size_t cap = [] {
    const char* v = getenv("VVB200_PROMPT_CACHE");
    const int n = v ? atoi(v) : 64;
    return (size_t)(n < 0 ? 0 : n);
  }();
  return cap;
}

// log-mel of a resident prompt (all ref_len frames), recomputed in place
static void prompt_mel(vv_engine* e, Prompt* pr) {
  const vv_arch& a = e->a;
  const WTensor* fb = find_w(e, "pre.mel_fb");
  launch_mel(pr->pcm, pr->n_samples, a.target_rms, pr->scale, e->hann, e->fft_tw, fb->d, a.n_mel, a.mel_clamp,
             pr->ref_len, pr->mel, e->st);
  e->launches += 2;
}

// Returns the resident prompt for this PCM, uploading it and computing its mel if it is not cached.
static int prompt_acquire(vv_engine* e, const int16_t* audio, int64_t n_samples, std::shared_ptr<Prompt>* out) {
  const vv_arch& a = e->a;
  if (!audio) return fail(VV_ERR_ARG, "prompt audio is null");
  if (n_samples < a.n_fft / 2 + 1) return fail(VV_ERR_ARG, "prompt audio too short (%lld samples)", (long long)n_samples);
  if (n_samples / a.hop + 1 > e->rope_max) return fail(VV_ERR_ARG, "prompt audio too long (%lld samples)", (long long)n_samples);
  uint64_t id, check;
  hash_pcm(audio, n_samples, &id, &check);
  auto it = e->prompts.find(id);
  if (it != e->prompts.end() && it->second->check == check && it->second->n_samples == n_samples) {
    it->second->last_use = ++e->prompt_tick;
    e->prompt_hits++;
    *out = it->second;
    return 0;
  }
  auto pr = std::make_shared<Prompt>();
  pr->id = id;
  pr->check = check;
  pr->n_samples = n_samples;
  pr->ref_len = (int)(n_samples / a.hop) + 1;
  pr->st = e->st;
  CK(cudaMallocAsync(&pr->pcm, (size_t)n_samples * 2, e->st));
  CK(cudaMallocAsync(&pr->mel, (size_t)pr->ref_len * a.n_mel * 4, e->st));
  CK(cudaMallocAsync(&pr->scale, 16, e->st));
  CK(cudaMemcpyAsync(pr->pcm, audio, (size_t)n_samples * 2, cudaMemcpyHostToDevice, e->st));
  prompt_mel(e, pr.get());
  e->prompt_uploads++;
  pr->last_use = ++e->prompt_tick;
  const size_t cap = prompt_cache_cap();
  if (cap > 0) {
    while (e->prompts.size() >= cap) {           // LRU: batches that still use an evicted prompt keep it alive
      auto victim = e->prompts.begin();
      for (auto c = e->prompts.begin(); c != e->prompts.end(); ++c)
        if (c->second->last_use < victim->second->last_use) victim = c;
      e->prompts.erase(victim);
    }
    e->prompts[id] = pr;
  }
  *out = pr;
  return 0;
}

extern "C" int vv_prompt_put(vv_engine* e, const int16_t* audio, int64_t n_samples, uint64_t* prompt_id_out,
                             int64_t* ref_len_out) {
  if (!e || !prompt_id_out) return fail(VV_ERR_ARG, "vv_prompt_put: bad argument");
  ENG_LOCK(e);
  if (!e->finalized) return fail(VV_ERR_STATE, "engine not finalized");
  CK(cudaSetDevice(e->device));
  std::shared_ptr<Prompt> pr;
  TRY(prompt_acquire(e, audio, n_samples, &pr));
  if (prompt_cache_cap() == 0) return fail(VV_ERR_STATE, "prompt cache disabled (VVB200_PROMPT_CACHE=0)");
  CKL();
  *prompt_id_out = pr->id;
  if (ref_len_out) *ref_len_out = pr->ref_len;
  return 0;
}

extern "C" int vv_prompt_drop(vv_engine* e, uint64_t prompt_id) {
  if (!e) return fail(VV_ERR_ARG, "null engine");
  ENG_LOCK(e);
  CK(cudaSetDevice(e->device));
  return e->prompts.erase(prompt_id) ? 0 : fail(VV_ERR_ARG, "prompt %llu is not resident", (unsigned long long)prompt_id);
}

extern "C" int vv_prompt_cache_clear(vv_engine* e) {
  if (!e) return fail(VV_ERR_ARG, "null engine");
  ENG_LOCK(e);
  CK(cudaSetDevice(e->device));
  e->prompts.clear();
  return 0;
}

/* what[0] = prompts resident, what[1] = uploads (misses) so far, what[2] = hits so far */
extern "C" int vv_prompt_cache_stats(vv_engine* e, int64_t* what) {
  if (!e || !what) return fail(VV_ERR_ARG, "vv_prompt_cache_stats: bad argument");
  ENG_LOCK(e);
  what[0] = (int64_t)e->prompts.size();
  what[1] = e->prompt_uploads;
  what[2] = e->prompt_hits;
  return 0;
}

// device part of the preprocess graph for chunk idx: the prompt's resident log-mel -> rows of the batch, y0 -> bf16
static int preprocess_device(vv_batch* b, int idx) {
  vv_engine* e = b->e;
  const vv_arch& a = e->a;
  const int T = b->T[idx], off = b->seq_off[idx];
  const Prompt* pr = b->prompt[idx].get();
  const int rows = std::min(pr->ref_len, T);
  float* dst = b->mel + (size_t)off * a.n_mel;
  CK(cudaMemcpyAsync(dst, pr->mel, (size_t)rows * a.n_mel * 4, cudaMemcpyDeviceToDevice, e->st));
  if (T > rows) CK(cudaMemsetAsync(dst + (size_t)rows * a.n_mel, 0, (size_t)(T - rows) * a.n_mel * 4, e->st));
  float* nz = b->noise + (size_t)off * a.n_mel;
  launch_noise_to_bf16(nz, T, a.n_mel, b->noise_b + (size_t)off * e->Kn, b->noise_b + (size_t)(b->R + off) * e->Kn,
                       e->Kn, e->st);
  e->launches++;
  return 0;
}

// everything of vv_preprocess after the prompt has been resolved
static int preprocess_chunk(vv_batch* b, int idx, std::shared_ptr<Prompt> pr, const int32_t* text_ids, int64_t n_ids,
                            const float* noise_or_null, uint64_t seed, uint64_t chunk_key, int64_t* ref_len_out) {
  vv_engine* e = b->e;
  const vv_arch& a = e->a;
  const int T = b->T[idx], off = b->seq_off[idx];
  // text ids: +1, truncated / zero padded to T (uncond rows stay 0); validated before any state changes
  for (int i = 0; i < T && i < n_ids; ++i)
    if (text_ids[i] < 0 || text_ids[i] >= a.vocab) return fail(VV_ERR_ARG, "text id %d out of vocabulary range", text_ids[i]);
  b->prompt[idx] = pr;
  b->ref_len[idx] = pr->ref_len;
  if (ref_len_out) *ref_len_out = pr->ref_len;
  // the pinned staging slice of this chunk may still be read by the H2D copy of an earlier vv_preprocess of the same
  // chunk (no host sync in between): wait for that copy before overwriting it
  if (b->ids_ev[idx]) CK(cudaEventSynchronize(b->ids_ev[idx]));
  else CK(cudaEventCreateWithFlags(&b->ids_ev[idx], cudaEventDisableTiming));
  int32_t* stage = b->ids_h + off;
  for (int i = 0; i < T; ++i) stage[i] = i < n_ids ? text_ids[i] + 1 : 0;
  CK(cudaMemcpyAsync(b->ids_d + off, stage, (size_t)T * 4, cudaMemcpyHostToDevice, e->st));
  CK(cudaEventRecord(b->ids_ev[idx], e->st));
  float* nz = b->noise + (size_t)off * a.n_mel;
  if (noise_or_null) {
    CK(cudaMemcpyAsync(nz, noise_or_null, (size_t)T * a.n_mel * 4, cudaMemcpyHostToDevice, e->st));
  } else {
    launch_philox_normal(nz, (int64_t)T * a.n_mel, seed, chunk_key, e->st);
    e->launches++;
  }
  CK(cudaMemcpyAsync(b->noise0 + (size_t)off * a.n_mel, nz, (size_t)T * a.n_mel * 4, cudaMemcpyDeviceToDevice, e->st));
  TRY(preprocess_device(b, idx));
  CKL();
  b->prepped[idx] = true;
  b->committed = false;
  b->decoded = false;
  b->steps_done = 0;
  return 0;
}

extern "C" int vv_preprocess(vv_batch* b, int idx, const int16_t* audio, int64_t n_samples, const int32_t* text_ids,
                             int64_t n_ids, const float* noise_or_null, uint64_t seed, uint64_t chunk_key,
                             int64_t* ref_len_out) {
  if (!b || !b->e || !audio || idx < 0 || idx >= b->B || n_ids < 0 || (n_ids > 0 && !text_ids))
    return fail(VV_ERR_ARG, "vv_preprocess: bad argument");
  ENG_LOCK(b->e);
  vv_engine* e = b->e;
  CK(cudaSetDevice(e->device));
  std::shared_ptr<Prompt> pr;
  TRY(prompt_acquire(e, audio, n_samples, &pr));
  return preprocess_chunk(b, idx, pr, text_ids, n_ids, noise_or_null, seed, chunk_key, ref_len_out);
}

extern "C" int vv_preprocess_prompt(vv_batch* b, int idx, uint64_t prompt_id, const int32_t* text_ids, int64_t n_ids,
                                    const float* noise_or_null, uint64_t seed, uint64_t chunk_key,
                                    int64_t* ref_len_out) {
  if (!b || !b->e || idx < 0 || idx >= b->B || n_ids < 0 || (n_ids > 0 && !text_ids))
    return fail(VV_ERR_ARG, "vv_preprocess_prompt: bad argument");
  ENG_LOCK(b->e);
  vv_engine* e = b->e;
  CK(cudaSetDevice(e->device));
  auto it = e->prompts.find(prompt_id);
  if (it == e->prompts.end())
    return fail(VV_ERR_STATE, "prompt %llu is not resident (evicted or never put)", (unsigned long long)prompt_id);
  it->second->last_use = ++e->prompt_tick;
  e->prompt_hits++;
  return preprocess_chunk(b, idx, it->second, text_ids, n_ids, noise_or_null, seed, chunk_key, ref_len_out);
}

// text ConvNeXt-V2 over all rows, concat, conditioning projection
static int commit(vv_batch* b) {
  if (b->committed) return 0;
  vv_engine* e = b->e;
  const vv_arch& a = e->a;
  const int M = b->M;
  const WTensor* emb = find_w(e, "pre.text_embed");
  launch_text_gather(b->ids_d, b->row_pos_d, b->row_mask_d, emb->d, e->pos_table, a.pos_table_len, M, a.text_dim,
                     b->tx, e->st);
  e->launches++;
  for (int i = 0; i < a.text_layers; ++i) {
    const ConvNextW& w = e->text_blocks[i];
    launch_dwconv_rows(b->tx, b->row_pos_d, b->row_len_d, w.dw_w, w.dw_b, M, a.text_dim, 7, b->t_tmp, e->st);
    launch_ln_affine(b->t_tmp, M, a.text_dim, w.ln_g, w.ln_b, a.ln_eps, b->t_hb, nullptr, e->st);
    e->launches += 2;
    GemmEpi e1;
    e1.bias = w.pw1_b; e1.act = ACT_GELU_ERF; e1.out_f32 = b->t_ff; e1.ld_f32 = a.text_ff;
    run_gemm(e, b->op_tpw1[i], e1);
    launch_grn(b->t_ff, b->seq_off_d, b->seq_len_d, b->row_seq_d, 2 * b->B, b->maxT, M, a.text_ff, w.grn_g, w.grn_b,
               b->gx2, b->nx, b->t_ffb, e->st);
    e->launches += 3;
    GemmEpi e2;
    e2.bias = w.pw2_b; e2.resid = b->tx; e2.ld_resid = a.text_dim; e2.out_f32 = b->tx; e2.ld_f32 = a.text_dim;
    run_gemm(e, b->op_tpw2[i], e2);
  }
  launch_cat_cond(b->mel, b->tx, b->row_mask_d, M, b->R, a.n_mel, a.text_dim, e->Kc, b->cat_b, b->cat_f32, e->st);
  e->launches++;
  GemmEpi ec;
  const float* inb = nullptr;
  TRY(need_w(e, "dit.in.b", &inb, a.dim));
  ec.bias = inb; ec.row_mask = b->row_mask_d; ec.out_f32 = b->cond_proj; ec.ld_f32 = a.dim;
  run_gemm(e, b->op_cond, ec);
  CKL();
  b->committed = true;
  return 0;
}

static int decode_all(vv_batch* b);

static void run_attention(vv_batch* b) {
  vv_engine* e = b->e;
  AttnParams p;
  p.seq_off = b->seq_off_d; p.seq_len = b->seq_len_d; p.tile_seq = b->tile_seq_d; p.tile_q0 = b->tile_q0_d;
  p.n_tiles = b->n_tiles; p.heads = e->a.heads; p.dim = e->a.dim; p.out = b->attn_o;
  p.scale_log2 = (1.0f / sqrtf((float)e->a.head_dim)) * 1.4426950408889634f;
  launch_attention(b->tQKV, p, e->st);
  e->launches++;
}

// programmatic dependent launch for the kernels of a DiT evaluation: on for small batches (see kernels.h)
static bool want_pdl(const vv_batch* b) {
  static const int forced = [] {
    const char* v = getenv("VVB200_PDL");
    return v ? (v[0] == '0' ? 0 : 1) : -1;
  }();
  static const int max_rows = [] {
    const char* v = getenv("VVB200_PDL_MAX_ROWS");
    return v ? atoi(v) : 8192;
  }();
  return forced >= 0 ? forced == 1 : b->M <= max_rows;
}
struct PdlScope {
  explicit PdlScope(bool on) { pdl_set(on); }
  ~PdlScope() { pdl_set(false); }
};

// one DiT evaluation (both CFG branches) + Euler update.  n_layers < 0: all layers + final projection.
static int run_step(vv_batch* b, const ModTable& mt, int step, int n_layers) {
  vv_engine* e = b->e;
  PdlScope pdl(want_pdl(b));
  const vv_arch& a = e->a;
  const int M = b->M, d = a.dim;
  const float *c1b = nullptr, *c2b = nullptr, *outb = nullptr;
  TRY(need_w(e, "dit.pos.c1.b", &c1b, d));
  TRY(need_w(e, "dit.pos.c2.b", &c2b, d));
  TRY(need_w(e, "dit.out.b", &outb, a.n_mel));
  {  // input embedding: x0 = noise @ Wn^T + (cond @ Wc^T + b)
    GemmEpi ep;
    ep.resid = b->cond_proj; ep.ld_resid = d; ep.out_f32 = b->x0; ep.ld_f32 = d; ep.out_bf16 = b->x0b; ep.ld_bf16 = d;
    run_gemm(e, b->op_in, ep);
  }
  {  // conv_pos_embed: x = x0 + mish(conv2(mish(conv1(x0))))
    GemmEpi e1;
    e1.bias = c1b; e1.act = ACT_MISH; e1.row_mask = b->row_mask_d; e1.out_bf16 = b->h1b; e1.ld_bf16 = d;
    run_gemm(e, b->op_c1, e1);
    GemmEpi e2;
    e2.bias = c2b; e2.act = ACT_MISH; e2.resid = b->x0; e2.ld_resid = d; e2.row_mask = b->row_mask_d;
    e2.out_f32 = b->x; e2.ld_f32 = d;
    run_gemm(e, b->op_c2, e2);
  }
  const int L = n_layers < 0 ? a.depth : std::min(n_layers, a.depth);
  for (int l = 0; l < L; ++l) {
    const LayerW& W = e->layers[l];
    const float* m = mt.blocks + ((size_t)step * a.depth + l) * 6 * d;
    launch_ln_mod(b->x, M, d, m, m + d, a.ln_eps, b->hb, e->st);
    e->launches++;
    GemmEpi eq;
    eq.bias = W.qkv_b; eq.out_bf16 = b->qkv; eq.ld_bf16 = 3 * d;
    eq.rope_dim = a.rope_heads * a.head_dim; eq.rope_off2 = d; eq.row_pos = b->row_pos_d; eq.rope_cs = e->rope_cs;
    run_gemm(e, b->op_qkv[l], eq);
    run_attention(b);
    GemmEpi eo;
    eo.bias = W.out_b; eo.gate = m + 2 * d; eo.resid = b->x; eo.ld_resid = d; eo.out_f32 = b->x; eo.ld_f32 = d;
    run_gemm(e, b->op_out[l], eo);
    launch_ln_mod(b->x, M, d, m + 3 * d, m + 4 * d, a.ln_eps, b->hb, e->st);
    e->launches++;
    GemmEpi e1;
    e1.bias = W.ff1_b; e1.act = ACT_GELU_TANH; e1.out_bf16 = b->ffb; e1.ld_bf16 = a.ff_dim;
    run_gemm(e, b->op_ff1[l], e1);
    GemmEpi e2;
    e2.bias = W.ff2_b; e2.gate = m + 5 * d; e2.resid = b->x; e2.ld_resid = d; e2.out_f32 = b->x; e2.ld_f32 = d;
    run_gemm(e, b->op_ff2[l], e2);
  }
  if (n_layers < 0) {
    const float* f = mt.fin + (size_t)step * 2 * d;  // (scale, shift)
    launch_ln_mod(b->x, M, d, f + d, f, a.ln_eps, b->hb, e->st);
    e->launches++;
    GemmEpi ef;
    ef.bias = outb; ef.out_f32 = b->v; ef.ld_f32 = 128;
    run_gemm(e, b->op_fin, ef);
    launch_cfg_euler(b->noise, b->noise_b, e->Kn, b->v, 128, b->row_mask_d, b->R, a.n_mel, mt.dt[step],
                     a.cfg_strength, e->st);
    e->launches++;
  }
  return 0;
}

extern "C" int vv_sample(vv_batch* b, int nfe, int first_step, int n_steps) {
  if (!b) return fail(VV_ERR_ARG, "null batch");
  ENG_LOCK(b->e);
  vv_engine* e = b->e;
  CK(cudaSetDevice(e->device));
  for (int i = 0; i < b->B; ++i)
    if (!b->prepped[i]) return fail(VV_ERR_STATE, "vv_sample: chunk %d has not been preprocessed", i);
  if (nfe <= 0) nfe = e->a.nfe;
  ModTable* mt;
  TRY(build_mod_table(e, nfe, &mt));
  if (first_step < 0 || n_steps < 0 || first_step + n_steps > mt->steps)
    return fail(VV_ERR_ARG, "vv_sample: steps [%d, %d) outside the %d-step grid", first_step, first_step + n_steps, mt->steps);
  TRY(commit(b));
  b->decoded = false;
  if (first_step == 0 && n_steps == mt->steps && n_steps > 1) {
    auto it = b->graphs.find(nfe);
    if (it == b->graphs.end()) {
      cudaGraph_t g = nullptr;
      const int64_t before = e->launches;
      const auto tc0 = std::chrono::steady_clock::now();
      CK(cudaStreamBeginCapture(e->st, cudaStreamCaptureModeRelaxed));
      int rc = 0;
      for (int s = 0; s < n_steps && rc == 0; ++s) rc = run_step(b, *mt, s, -1);
      cudaError_t ce = cudaStreamEndCapture(e->st, &g);
      if (getenv("VVB200_VERBOSE"))
        fprintf(stderr, "vvb200: capture %.1f ms\n",
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tc0).count());
      if (rc) return rc;
      if (ce != cudaSuccess) return fail(VV_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
      b->graph_launches[nfe] = e->launches - before;
      e->launches = before;
      b->graphs[nfe] = g;
      it = b->graphs.find(nfe);
    }
    const auto tg0 = std::chrono::steady_clock::now();
    vv_engine::LoopExec& le = e->loop_exec[std::make_pair(nfe, want_pdl(b) ? 1 : 0)];
    const bool verbose_g = getenv("VVB200_VERBOSE") && le.owner != b;
    if (le.exec && le.owner != b) {
      cudaGraphExecUpdateResultInfo info;
      if (cudaGraphExecUpdate(le.exec, it->second, &info) != cudaSuccess) {   // different topology (tile plan): rebuild
        (void)cudaGetLastError();
        cudaGraphExecDestroy(le.exec);
        le.exec = nullptr;
      }
    }
    if (!le.exec) CK(cudaGraphInstantiate(&le.exec, it->second, 0));
    le.owner = b;
    const auto tg1 = std::chrono::steady_clock::now();
    CK(cudaGraphLaunch(le.exec, e->st));
    if (verbose_g)
      fprintf(stderr, "vvb200: graph exec update/instantiate %.1f ms, launch call %.1f ms\n",
              std::chrono::duration<double, std::milli>(tg1 - tg0).count(),
              std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tg1).count());
    e->launches += b->graph_launches[nfe];
  } else {
    for (int s = first_step; s < first_step + n_steps; ++s) TRY(run_step(b, *mt, s, -1));
    CKL();
  }
  b->steps_done = first_step + n_steps;
  return 0;
}

// debugging / parity aid: input embedding + the first n_layers blocks of step `step`, no Euler update
extern "C" int vv_debug_partial_step(vv_batch* b, int nfe, int step, int n_layers) {
  if (!b) return fail(VV_ERR_ARG, "null batch");
  ENG_LOCK(b->e);
  vv_engine* e = b->e;
  CK(cudaSetDevice(e->device));
  if (nfe <= 0) nfe = e->a.nfe;
  ModTable* mt;
  TRY(build_mod_table(e, nfe, &mt));
  if (step < 0 || step >= mt->steps || n_layers < 0) return fail(VV_ERR_ARG, "bad step / n_layers");
  TRY(commit(b));
  TRY(run_step(b, *mt, step, n_layers));
  CKL();
  return 0;
}


// Whole path with every input already resident in HBM (prompt PCM, text ids, y0): mel -> text embed -> cond ->
// (nfe-1) DiT steps -> Vocos/iSTFT -> int16 PCM left on the device.  No host<->device copies, no host sync.
// ev: optional 4 events recorded on the engine stream at the stage boundaries (start | preprocess done | loop done |
// decode done)
static int run_resident_impl(vv_batch* b, int nfe, cudaEvent_t* ev) {
  vv_engine* e = b->e;
  const vv_arch& a = e->a;
  for (int i = 0; i < b->B; ++i)
    if (!b->prepped[i]) return fail(VV_ERR_STATE, "vv_run_resident: chunk %d has not been preprocessed", i);
  if (ev) CK(cudaEventRecord(ev[0], e->st));
  CK(cudaMemcpyAsync(b->noise, b->noise0, (size_t)b->R * a.n_mel * 4, cudaMemcpyDeviceToDevice, e->st));
  // the mel front-end runs once per DISTINCT prompt of the batch (chunks of one text share theirs), from resident PCM
  std::set<const Prompt*> seen;
  for (int i = 0; i < b->B; ++i)
    if (seen.insert(b->prompt[i].get()).second) prompt_mel(e, b->prompt[i].get());
  for (int i = 0; i < b->B; ++i) TRY(preprocess_device(b, i));
  b->committed = false;
  b->decoded = false;
  if (nfe <= 0) nfe = a.nfe;
  TRY(commit(b));
  if (ev) CK(cudaEventRecord(ev[1], e->st));
  TRY(vv_sample(b, nfe, 0, nfe - 1));
  if (ev) CK(cudaEventRecord(ev[2], e->st));
  TRY(decode_all(b));
  if (ev) CK(cudaEventRecord(ev[3], e->st));
  return 0;
}

extern "C" int vv_run_resident(vv_batch* b, int nfe) {
  if (!b || !b->e) return fail(VV_ERR_ARG, "null batch");
  ENG_LOCK(b->e);
  CK(cudaSetDevice(b->e->device));
  return run_resident_impl(b, nfe, nullptr);
}

// The resident path once more with CUDA events at the stage boundaries: ms_out[0] preprocess (mel of the distinct
// prompts, text ConvNeXt, conditioning projection), [1] the (nfe-1)-step sampling loop, [2] decode (Vocos + iSTFT),
// [3] the whole call.  Synchronous.
extern "C" int vv_profile_stages(vv_batch* b, int nfe, float* ms_out /* [4] */) {
  if (!b || !b->e || !ms_out) return fail(VV_ERR_ARG, "vv_profile_stages: bad argument");
  ENG_LOCK(b->e);
  vv_engine* e = b->e;
  CK(cudaSetDevice(e->device));
  cudaEvent_t ev[4];
  for (int i = 0; i < 4; ++i) CK(cudaEventCreate(&ev[i]));
  int rc = run_resident_impl(b, nfe, ev);
  cudaError_t r = cudaStreamSynchronize(e->st);
  if (rc == 0 && r == cudaSuccess) {
    for (int i = 0; i < 3; ++i) cudaEventElapsedTime(&ms_out[i], ev[i], ev[i + 1]);
    cudaEventElapsedTime(&ms_out[3], ev[0], ev[3]);
  }
  for (int i = 0; i < 4; ++i) cudaEventDestroy(ev[i]);
  if (rc) return rc;
  if (r != cudaSuccess) return fail(VV_ERR_CUDA, "vv_profile_stages failed: %s", cudaGetErrorString(r));
  return 0;
}

// One eager DiT step with a CUDA-event pair around every launch; accumulates milliseconds per kernel class:
// 0 qkv GEMM, 1 out-proj GEMM, 2 ffn-up GEMM, 3 ffn-down GEMM, 4 attention, 5 LN-modulate, 6 conv_pos (2 launches),
// 7 input-embed GEMM + final projection GEMM + CFG/Euler.  The state (noise) advances by that one step.
extern "C" int vv_profile_step(vv_batch* b, int nfe, int step, float* ms_out /* [8] */) {
  if (!b || !ms_out) return fail(VV_ERR_ARG, "vv_profile_step: bad argument");
  ENG_LOCK(b->e);
  vv_engine* e = b->e;
  const vv_arch& a = e->a;
  CK(cudaSetDevice(e->device));
  if (nfe <= 0) nfe = a.nfe;
  ModTable* mt;
  TRY(build_mod_table(e, nfe, &mt));
  if (step < 0 || step >= mt->steps) return fail(VV_ERR_ARG, "bad step");
  TRY(commit(b));
  const int M = b->M, d = a.dim;
  std::vector<cudaEvent_t> evs;
  std::vector<int> cls;
  auto mark = [&](int c) {
    cudaEvent_t ev;
    cudaEventCreate(&ev);
    cudaEventRecord(ev, e->st);
    evs.push_back(ev);
    cls.push_back(c);
  };
  const float *c1b = nullptr, *c2b = nullptr, *outb = nullptr;
  TRY(need_w(e, "dit.pos.c1.b", &c1b, d));
  TRY(need_w(e, "dit.pos.c2.b", &c2b, d));
  TRY(need_w(e, "dit.out.b", &outb, a.n_mel));
  mark(-1);
  {
    GemmEpi ep;
    ep.resid = b->cond_proj; ep.ld_resid = d; ep.out_f32 = b->x0; ep.ld_f32 = d; ep.out_bf16 = b->x0b; ep.ld_bf16 = d;
    run_gemm(e, b->op_in, ep);
    mark(7);
    GemmEpi e1;
    e1.bias = c1b; e1.act = ACT_MISH; e1.row_mask = b->row_mask_d; e1.out_bf16 = b->h1b; e1.ld_bf16 = d;
    run_gemm(e, b->op_c1, e1);
    GemmEpi e2;
    e2.bias = c2b; e2.act = ACT_MISH; e2.resid = b->x0; e2.ld_resid = d; e2.row_mask = b->row_mask_d;
    e2.out_f32 = b->x; e2.ld_f32 = d;
    run_gemm(e, b->op_c2, e2);
    mark(6);
  }
  for (int l = 0; l < a.depth; ++l) {
    const LayerW& W = e->layers[l];
    const float* m = mt->blocks + ((size_t)step * a.depth + l) * 6 * d;
    launch_ln_mod(b->x, M, d, m, m + d, a.ln_eps, b->hb, e->st);
    e->launches++;
    mark(5);
    GemmEpi eq;
    eq.bias = W.qkv_b; eq.out_bf16 = b->qkv; eq.ld_bf16 = 3 * d;
    eq.rope_dim = a.rope_heads * a.head_dim; eq.rope_off2 = d; eq.row_pos = b->row_pos_d; eq.rope_cs = e->rope_cs;
    run_gemm(e, b->op_qkv[l], eq);
    mark(0);
    run_attention(b);
    mark(4);
    GemmEpi eo;
    eo.bias = W.out_b; eo.gate = m + 2 * d; eo.resid = b->x; eo.ld_resid = d; eo.out_f32 = b->x; eo.ld_f32 = d;
    run_gemm(e, b->op_out[l], eo);
    mark(1);
    launch_ln_mod(b->x, M, d, m + 3 * d, m + 4 * d, a.ln_eps, b->hb, e->st);
    e->launches++;
    mark(5);
    GemmEpi e1;
    e1.bias = W.ff1_b; e1.act = ACT_GELU_TANH; e1.out_bf16 = b->ffb; e1.ld_bf16 = a.ff_dim;
    run_gemm(e, b->op_ff1[l], e1);
    mark(2);
    GemmEpi e2;
    e2.bias = W.ff2_b; e2.gate = m + 5 * d; e2.resid = b->x; e2.ld_resid = d; e2.out_f32 = b->x; e2.ld_f32 = d;
    run_gemm(e, b->op_ff2[l], e2);
    mark(3);
  }
  {
    const float* f = mt->fin + (size_t)step * 2 * d;
    launch_ln_mod(b->x, M, d, f + d, f, a.ln_eps, b->hb, e->st);
    e->launches++;
    mark(5);
    GemmEpi ef;
    ef.bias = outb; ef.out_f32 = b->v; ef.ld_f32 = 128;
    run_gemm(e, b->op_fin, ef);
    launch_cfg_euler(b->noise, b->noise_b, e->Kn, b->v, 128, b->row_mask_d, b->R, a.n_mel, mt->dt[step],
                     a.cfg_strength, e->st);
    e->launches++;
    mark(7);
  }
  cudaError_t r = cudaStreamSynchronize(e->st);
  for (int i = 0; i < 8; ++i) ms_out[i] = 0.f;
  if (r == cudaSuccess)
    for (size_t i = 1; i < evs.size(); ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, evs[i - 1], evs[i]);
      ms_out[cls[i]] += ms;
    }
  for (cudaEvent_t ev : evs) cudaEventDestroy(ev);
  if (r != cudaSuccess) return fail(VV_ERR_CUDA, "profile step failed: %s", cudaGetErrorString(r));
  CKL();
  b->decoded = false;
  return 0;
}

// ------------------------------------------------------------------------------------------------ decode
static int decode_all(vv_batch* b) {
  if (b->decoded) return 0;
  vv_engine* e = b->e;
  const vv_arch& a = e->a;
  for (int i = 0; i < b->B; ++i)
    if (!b->prepped[i]) return fail(VV_ERR_STATE, "decode: chunk %d has not been preprocessed", i);
  if (b->dec_ref_len != b->ref_len) {  // (re)build the packed target-frame layout
    std::vector<int32_t> src(b->Rd_max, 0), pos(b->Rd_max, 0), len(b->Rd_max, 0);
    int rd0 = 0;
    int64_t po = 0;
    for (int i = 0; i < b->B; ++i) {
      const int tg = std::max(0, b->T[i] - b->ref_len[i]);
      b->dec_off[i] = rd0;
      b->dec_len[i] = tg;
      for (int p = 0; p < tg; ++p) {
        src[rd0 + p] = b->seq_off[i] + b->ref_len[i] + p;
        pos[rd0 + p] = p;
        len[rd0 + p] = tg;
      }
      rd0 += tg;
      b->pcm_off[i] = po;
      b->pcm_len[i] = tg > 1 ? (int64_t)(tg - 1) * a.hop : 0;
      po += b->pcm_len[i];
    }
    b->Rd = rd0;
    b->pcm_total = po;
    {
      std::vector<int32_t> doff(b->dec_off.begin(), b->dec_off.end()), dlen(b->dec_len.begin(), b->dec_len.end());
      CK(cudaMemcpyAsync(b->d_dec_off, doff.data(), (size_t)b->B * 4, cudaMemcpyHostToDevice, e->st));
      CK(cudaMemcpyAsync(b->d_dec_len, dlen.data(), (size_t)b->B * 4, cudaMemcpyHostToDevice, e->st));
      CK(cudaMemcpyAsync(b->d_pcm_off, b->pcm_off.data(), (size_t)b->B * 8, cudaMemcpyHostToDevice, e->st));
      CK(cudaStreamSynchronize(e->st));
    }
    if (rd0 > 0) {
      CK(cudaMemcpyAsync(b->d_src_row, src.data(), (size_t)rd0 * 4, cudaMemcpyHostToDevice, e->st));
      CK(cudaMemcpyAsync(b->d_row_pos, pos.data(), (size_t)rd0 * 4, cudaMemcpyHostToDevice, e->st));
      CK(cudaMemcpyAsync(b->d_row_len, len.data(), (size_t)rd0 * 4, cudaMemcpyHostToDevice, e->st));
      CK(cudaStreamSynchronize(e->st));
    }
    b->dec_ops.clear();
    const int vd0 = a.voc_dim;
    b->dec_ops.push_back(make_op(e, b->v_emb, e->Kemb, b->Rd_max, std::max(rd0, 1), e->voc_embed, e->Kemb, vd0, e->Kemb));
    for (int i = 0; i < a.voc_layers; ++i) {
      const ConvNextW& w = e->voc_blocks[i];
      b->dec_ops.push_back(make_op(e, b->v_hb, vd0, b->Rd_max, std::max(rd0, 1), w.pw1, vd0, a.voc_ff, vd0));
      b->dec_ops.push_back(make_op(e, b->v_ffb, a.voc_ff, b->Rd_max, std::max(rd0, 1), w.pw2, a.voc_ff, vd0, a.voc_ff));
    }
    b->dec_ops.push_back(make_op(e, b->v_hb, vd0, b->Rd_max, std::max(rd0, 1), e->voc_head, vd0, a.n_fft + 2, vd0));
    b->dec_ref_len = b->ref_len;
  }
  const int rd = b->Rd;
  if (rd == 0) {
    b->decoded = true;
    return 0;
  }
  const int vd = a.voc_dim;
  const float *eb = nullptr, *ng = nullptr, *nb = nullptr, *fg = nullptr, *fbb = nullptr, *hb = nullptr;
  TRY(need_w(e, "voc.embed.b", &eb, vd));
  TRY(need_w(e, "voc.norm.g", &ng, vd));
  TRY(need_w(e, "voc.norm.b", &nb, vd));
  TRY(need_w(e, "voc.final.g", &fg, vd));
  TRY(need_w(e, "voc.final.b", &fbb, vd));
  TRY(need_w(e, "voc.head.b", &hb, a.n_fft + 2));
  launch_voc_im2col(b->noise, b->d_src_row, b->d_row_pos, b->d_row_len, rd, a.n_mel, a.voc_k, e->Kemb, b->v_emb, e->st);
  e->launches++;
  const GemmOp& op = b->dec_ops[0];
  GemmEpi ee;
  ee.bias = eb; ee.out_f32 = b->vx; ee.ld_f32 = vd;
  run_gemm(e, op, ee);
  launch_ln_affine(b->vx, rd, vd, ng, nb, a.ln_eps, nullptr, b->vx, e->st);
  e->launches++;
  for (int i = 0; i < a.voc_layers; ++i) {
    const ConvNextW& w = e->voc_blocks[i];
    launch_dwconv_rows(b->vx, b->d_row_pos, b->d_row_len, w.dw_w, w.dw_b, rd, vd, a.voc_k, b->v_tmp, e->st);
    launch_ln_affine(b->v_tmp, rd, vd, w.ln_g, w.ln_b, a.ln_eps, b->v_hb, nullptr, e->st);
    e->launches += 2;
    const GemmOp& o1 = b->dec_ops[1 + 2 * i];
    GemmEpi e1;
    e1.bias = w.pw1_b; e1.act = ACT_GELU_ERF; e1.out_bf16 = b->v_ffb; e1.ld_bf16 = a.voc_ff;
    run_gemm(e, o1, e1);
    const GemmOp& o2 = b->dec_ops[2 + 2 * i];
    GemmEpi e2;
    e2.bias = w.pw2_b; e2.gate = w.gamma; e2.resid = b->vx; e2.ld_resid = vd; e2.out_f32 = b->vx; e2.ld_f32 = vd;
    run_gemm(e, o2, e2);
  }
  launch_ln_affine(b->vx, rd, vd, fg, fbb, a.ln_eps, b->v_hb, nullptr, e->st);
  e->launches++;
  const GemmOp& oh = b->dec_ops.back();
  GemmEpi eh;
  eh.bias = hb; eh.out_f32 = b->v_head; eh.ld_f32 = b->ld_head;
  run_gemm(e, oh, eh);
  {  // inverse FFT + window + overlap-add + envelope + int16 pack of ALL chunks in one launch; frames stay in smem
    int max_len = 0;
    for (int i = 0; i < b->B; ++i) max_len = std::max(max_len, b->dec_len[i]);
    launch_istft_ola(b->v_head, b->ld_head, e->hann, e->fft_tw, a.mag_clip, b->d_dec_off, b->d_dec_len, b->d_pcm_off,
                     b->B, max_len, a.pcm_scale, b->pcm_d, e->st);
    e->launches++;
  }
  CKL();
  b->decoded = true;
  return 0;
}

extern "C" int64_t vv_batch_pcm_len(const vv_batch* b, int idx) {
  if (!b || idx < 0 || idx >= b->B) return fail(VV_ERR_ARG, "bad argument");
  const int tg = std::max(0, b->T[idx] - b->ref_len[idx]);
  return tg > 1 ? (int64_t)(tg - 1) * b->e->a.hop : 0;
}

extern "C" int vv_decode(vv_batch* b, int idx, int16_t* pcm_out, int64_t capacity, int64_t* n_out) {
  if (!b || idx < 0 || idx >= b->B || !pcm_out) return fail(VV_ERR_ARG, "vv_decode: bad argument");
  ENG_LOCK(b->e);
  vv_engine* e = b->e;
  CK(cudaSetDevice(e->device));
  TRY(decode_all(b));
  const int64_t n = b->pcm_len[idx];
  if (capacity < n) return fail(VV_ERR_ARG, "vv_decode: capacity %lld < %lld samples", (long long)capacity, (long long)n);
  if (n > 0) CK(cudaMemcpyAsync(pcm_out, b->pcm_d + b->pcm_off[idx], (size_t)n * 2, cudaMemcpyDeviceToHost, e->st));
  CK(cudaStreamSynchronize(e->st));
  if (n_out) *n_out = n;
  return 0;
}

extern "C" int vv_decode_all(vv_batch* b, int16_t* const* pcm_out, int64_t* n_out) {
  if (!b || !pcm_out) return fail(VV_ERR_ARG, "vv_decode_all: bad argument");
  ENG_LOCK(b->e);
  vv_engine* e = b->e;
  CK(cudaSetDevice(e->device));
  TRY(decode_all(b));
  for (int i = 0; i < b->B; ++i) {
    if (b->pcm_len[i] > 0)
      CK(cudaMemcpyAsync(pcm_out[i], b->pcm_d + b->pcm_off[i], (size_t)b->pcm_len[i] * 2, cudaMemcpyDeviceToHost, e->st));
    if (n_out) n_out[i] = b->pcm_len[i];
  }
  CK(cudaStreamSynchronize(e->st));
  return 0;
}

// ------------------------------------------------------------------------------------------------ taps
static int copy_rows_f32(vv_engine* e, const float* src, int ld, int row0, int rows, int cols, float* out) {
  CK(cudaMemcpy2DAsync(out, (size_t)cols * 4, src + (size_t)row0 * ld, (size_t)ld * 4, (size_t)cols * 4, rows,
                       cudaMemcpyDeviceToHost, e->st));
  return 0;
}
static int copy_rows_bf16(vv_engine* e, const bf16* src, int ld, int row0, int rows, int cols, float* out) {
  std::vector<uint16_t> tmp((size_t)rows * cols);
  CK(cudaMemcpy2DAsync(tmp.data(), (size_t)cols * 2, src + (size_t)row0 * ld, (size_t)ld * 2, (size_t)cols * 2, rows,
                       cudaMemcpyDeviceToHost, e->st));
  CK(cudaStreamSynchronize(e->st));
  for (size_t i = 0; i < tmp.size(); ++i) {
    uint32_t u = (uint32_t)tmp[i] << 16;
    memcpy(&out[i], &u, 4);
  }
  return 0;
}

extern "C" int64_t vv_get_tensor(vv_batch* b, int idx, const char* name, float* out, int64_t capacity) {
  if (!b || !name || !out || idx < 0 || idx >= b->B) return fail(VV_ERR_ARG, "vv_get_tensor: bad argument");
  ENG_LOCK(b->e);
  vv_engine* e = b->e;
  const vv_arch& a = e->a;
  CK(cudaSetDevice(e->device));
  const int T = b->T[idx], off = b->seq_off[idx], offu = b->seq_off[b->B + idx];
  const std::string n(name);
  int64_t need = 0;
  auto two = [&](auto fn, int cols) -> int {  // [2, T, cols]: cond rows then uncond rows
    TRY(fn(off, out));
    TRY(fn(offu, out + (size_t)T * cols));
    return 0;
  };
  if (n == "noise") {
    need = (int64_t)T * a.n_mel;
    if (capacity < need) return fail(VV_ERR_ARG, "capacity too small");
    TRY(copy_rows_f32(e, b->noise, a.n_mel, off, T, a.n_mel, out));
  } else if (n == "mel") {
    const int rl = std::min(b->ref_len[idx], T);
    need = (int64_t)rl * a.n_mel;
    if (capacity < need) return fail(VV_ERR_ARG, "capacity too small");
    TRY(copy_rows_f32(e, b->mel, a.n_mel, off, rl, a.n_mel, out));
  } else if (n == "cat_mel_text" || n == "cat_mel_text_drop") {
    TRY(commit(b));
    const int cd = a.n_mel + a.text_dim;
    need = (int64_t)T * cd;
    if (capacity < need) return fail(VV_ERR_ARG, "capacity too small");
    TRY(copy_rows_f32(e, b->cat_f32, e->Kc, n == "cat_mel_text" ? off : offu, T, cd, out));
  } else if (n == "hidden" || n == "x0" || n == "cond_proj") {
    need = (int64_t)2 * T * a.dim;
    if (capacity < need) return fail(VV_ERR_ARG, "capacity too small");
    const float* src = n == "hidden" ? b->x : (n == "x0" ? b->x0 : b->cond_proj);
    TRY(two([&](int r0, float* o) { return copy_rows_f32(e, src, a.dim, r0, T, a.dim, o); }, a.dim));
  } else if (n == "v") {
    need = (int64_t)2 * T * a.n_mel;
    if (capacity < need) return fail(VV_ERR_ARG, "capacity too small");
    TRY(two([&](int r0, float* o) { return copy_rows_f32(e, b->v, 128, r0, T, a.n_mel, o); }, a.n_mel));
  } else if (n == "qkv" || n == "attn" || n == "hb" || n == "h1b" || n == "ffb") {
    const int cols = n == "qkv" ? 3 * a.dim : (n == "ffb" ? a.ff_dim : a.dim);
    const bf16* src = n == "qkv" ? b->qkv : (n == "attn" ? b->attn_o : (n == "hb" ? b->hb : (n == "h1b" ? b->h1b : b->ffb)));
    need = (int64_t)2 * T * cols;
    if (capacity < need) return fail(VV_ERR_ARG, "capacity too small");
    TRY(two([&](int r0, float* o) { return copy_rows_bf16(e, src, cols, r0, T, cols, o); }, cols));
  } else if (n == "voc_head") {
    TRY(decode_all(b));
    const int tg = b->dec_len[idx], cols = a.n_fft + 2;
    need = (int64_t)tg * cols;
    if (capacity < need) return fail(VV_ERR_ARG, "capacity too small");
    if (tg > 0) TRY(copy_rows_f32(e, b->v_head, b->ld_head, b->dec_off[idx], tg, cols, out));
  } else {
    return fail(VV_ERR_ARG, "vv_get_tensor: unknown tensor '%s'", name);
  }
  CK(cudaStreamSynchronize(e->st));
  return need;
}

extern "C" int vv_set_noise(vv_batch* b, int idx, const float* noise) {
  if (!b || !noise || idx < 0 || idx >= b->B) return fail(VV_ERR_ARG, "vv_set_noise: bad argument");
  ENG_LOCK(b->e);
  vv_engine* e = b->e;
  const vv_arch& a = e->a;
  CK(cudaSetDevice(e->device));
  const int T = b->T[idx], off = b->seq_off[idx];
  float* nz = b->noise + (size_t)off * a.n_mel;
  CK(cudaMemcpyAsync(nz, noise, (size_t)T * a.n_mel * 4, cudaMemcpyHostToDevice, e->st));
  launch_noise_to_bf16(nz, T, a.n_mel, b->noise_b + (size_t)off * e->Kn, b->noise_b + (size_t)(b->R + off) * e->Kn,
                       e->Kn, e->st);
  e->launches++;
  CK(cudaStreamSynchronize(e->st));
  b->decoded = false;
  return 0;
}

extern "C" int vv_set_ref_len(vv_batch* b, int idx, int64_t ref_len) {
  if (!b || idx < 0 || idx >= b->B || ref_len < 0 || ref_len > b->T[idx]) return fail(VV_ERR_ARG, "vv_set_ref_len: bad argument");
  ENG_LOCK(b->e);
  b->ref_len[idx] = (int)ref_len;
  b->decoded = false;
  return 0;
}

extern "C" int vv_set_cond(vv_batch* b, int idx, const float* cat_mel_text, const float* cat_mel_text_drop) {
  if (!b || !cat_mel_text || !cat_mel_text_drop || idx < 0 || idx >= b->B) return fail(VV_ERR_ARG, "vv_set_cond: bad argument");
  ENG_LOCK(b->e);
  vv_engine* e = b->e;
  const vv_arch& a = e->a;
  CK(cudaSetDevice(e->device));
  TRY(commit(b));
  const int T = b->T[idx], cd = a.n_mel + a.text_dim;
  const int offs[2] = {b->seq_off[idx], b->seq_off[b->B + idx]};
  const float* srcs[2] = {cat_mel_text, cat_mel_text_drop};
  for (int k = 0; k < 2; ++k) {
    float* dstf = b->cat_f32 + (size_t)offs[k] * e->Kc;
    CK(cudaMemcpy2DAsync(dstf, (size_t)e->Kc * 4, srcs[k], (size_t)cd * 4, (size_t)cd * 4, T, cudaMemcpyHostToDevice, e->st));
    launch_f32_to_bf16_2d(dstf, T, cd, e->Kc, b->cat_b + (size_t)offs[k] * e->Kc, e->Kc, e->Kc, e->st);
    e->launches++;
  }
  GemmEpi ec;
  const float* inb = nullptr;
  TRY(need_w(e, "dit.in.b", &inb, a.dim));
  ec.bias = inb; ec.row_mask = b->row_mask_d; ec.out_f32 = b->cond_proj; ec.ld_f32 = a.dim;
  run_gemm(e, b->op_cond, ec);
  CK(cudaStreamSynchronize(e->st));
  CKL();
  return 0;
}

// ------------------------------------------------------------------------------------------------ whole path
static int batch_crossfade(vv_batch* b, const int32_t* order, int n, const double* fo, const double* fi, int nf,
                           int16_t* pcm_out, int64_t capacity, int64_t* n_out);

// preprocess -> sampling loop -> decode for B requests on a cached batch; the PCM stays on the device
static int synthesize_on_device(vv_engine* e, vv_request* reqs, int B, int nfe, uint64_t seed, vv_batch** out_b) {
  if (!e->finalized) return fail(VV_ERR_STATE, "engine not finalized");
  CK(cudaSetDevice(e->device));
  std::vector<int64_t> key(B);
  for (int i = 0; i < B; ++i) key[i] = reqs[i].total_frames;
  vv_batch* b = nullptr;
  auto it = e->batch_cache.find(key);
  if (it != e->batch_cache.end()) {
    b = it->second;
  } else {
    static const size_t cap = [] {
      const char* v = getenv("VVB200_BATCH_CACHE");
      const int n = v ? atoi(v) : 6;
      return (size_t)(n < 1 ? 1 : n);
    }();
    auto evict_lru = [&]() {
      auto victim = e->batch_cache.end();
      for (auto c = e->batch_cache.begin(); c != e->batch_cache.end(); ++c)
        if (victim == e->batch_cache.end() || e->batch_last_use[c->second] < e->batch_last_use[victim->second]) victim = c;
      if (victim == e->batch_cache.end()) return false;
      vv_batch_destroy(victim->second);      // synchronises the stream, removes itself from the cache and the LRU
      return true;
    };
    while (e->batch_cache.size() >= cap && evict_lru()) {}
    const auto tb0 = std::chrono::steady_clock::now();
    int rc = vv_batch_create(e, B, key.data(), &b);
    if (getenv("VVB200_VERBOSE"))
      fprintf(stderr, "vvb200: batch create %.1f ms\n",
              std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tb0).count());
    while (rc == VV_ERR_CUDA && evict_lru()) {     // out of device memory: drop cached batches until it fits
      (void)cudaGetLastError();
      rc = vv_batch_create(e, B, key.data(), &b);
    }
    if (rc) return rc;
    e->batch_cache[key] = b;
  }
  e->batch_last_use[b] = ++e->batch_tick;
  for (int i = 0; i < B; ++i) {
    int64_t rl;
    if (reqs[i].prompt_id != 0 && e->prompts.count(reqs[i].prompt_id))
      TRY(vv_preprocess_prompt(b, i, reqs[i].prompt_id, reqs[i].text_ids, reqs[i].n_ids, reqs[i].noise, seed,
                               reqs[i].chunk_key, &rl));
    else if (reqs[i].audio)       // hashed: a prompt shared by several chunks / calls is uploaded once
      TRY(vv_preprocess(b, i, reqs[i].audio, reqs[i].n_samples, reqs[i].text_ids, reqs[i].n_ids, reqs[i].noise, seed,
                        reqs[i].chunk_key, &rl));
    else
      return fail(VV_ERR_STATE, "request %d: prompt %llu is not resident and no audio was given", i,
                  (unsigned long long)reqs[i].prompt_id);
  }
  if (nfe <= 0) nfe = e->a.nfe;
  TRY(vv_sample(b, nfe, 0, nfe - 1));
  TRY(decode_all(b));
  *out_b = b;
  return 0;
}

extern "C" int vv_synthesize_batch(vv_engine* e, vv_request* reqs, int B, int nfe, uint64_t seed) {
  if (!e || !reqs || B <= 0) return fail(VV_ERR_ARG, "vv_synthesize_batch: bad argument");
  ENG_LOCK(e);
  vv_batch* b = nullptr;
  TRY(synthesize_on_device(e, reqs, B, nfe, seed, &b));
  for (int i = 0; i < B; ++i) {
    if (reqs[i].pcm_capacity < b->pcm_len[i] || (!reqs[i].pcm_out && b->pcm_len[i] > 0))
      return fail(VV_ERR_ARG, "request %d: pcm_capacity %lld < %lld", i, (long long)reqs[i].pcm_capacity, (long long)b->pcm_len[i]);
    if (b->pcm_len[i] > 0)
      CK(cudaMemcpyAsync(reqs[i].pcm_out, b->pcm_d + b->pcm_off[i], (size_t)b->pcm_len[i] * 2, cudaMemcpyDeviceToHost, e->st));
    reqs[i].n_out = b->pcm_len[i];
  }
  CK(cudaStreamSynchronize(e->st));
  return 0;
}

// vv_synthesize_batch for the chunks of ONE text, joined on the device: the requests are synthesized as a batch, then
// clip-fixed and cross-faded in request order (vv_batch_crossfade); only the joined wave crosses PCIe.  The per-request
// pcm_out / pcm_capacity fields are ignored.
extern "C" int vv_synthesize_joined(vv_engine* e, vv_request* reqs, int B, int nfe, uint64_t seed, const double* fade_out,
                                    const double* fade_in, int n_fade, int16_t* pcm_out, int64_t capacity,
                                    int64_t* n_out) {
  if (!e || !reqs || B <= 0 || !pcm_out) return fail(VV_ERR_ARG, "vv_synthesize_joined: bad argument");
  ENG_LOCK(e);
  vv_batch* b = nullptr;
  TRY(synthesize_on_device(e, reqs, B, nfe, seed, &b));
  for (int i = 0; i < B; ++i) reqs[i].n_out = b->pcm_len[i];
  if (B == 1) {       // a single chunk is returned untouched (audio_processor.py:130-131)
    if (capacity < b->pcm_len[0]) return fail(VV_ERR_ARG, "vv_synthesize_joined: capacity too small");
    if (b->pcm_len[0] > 0)
      CK(cudaMemcpyAsync(pcm_out, b->pcm_d + b->pcm_off[0], (size_t)b->pcm_len[0] * 2, cudaMemcpyDeviceToHost, e->st));
    CK(cudaStreamSynchronize(e->st));
    if (n_out) *n_out = b->pcm_len[0];
    return 0;
  }
  return batch_crossfade(b, nullptr, B, fade_out, fade_in, n_fade, pcm_out, capacity, n_out);
}

// ------------------------------------------------------------------------------------------------ cross-fade
struct XfChunkH {          // = frontend.cu XfChunk
  const int16_t* src;
  int64_t len;
  int64_t out_off;
};
static_assert(sizeof(XfChunkH) == 24, "XfChunk layout");

// joins device-resident int16 chunks (src[i], lens[i]) into out_d; *total receives the joined length.
// Preconditions checked by the callers: n >= 2, nf >= 1, every chunk holds at least 2 * nf samples.
static int crossfade_device(vv_engine* e, const std::vector<const int16_t*>& src, const std::vector<int64_t>& lens, int nf,
                            const double* fade_out_h, const double* fade_in_h, int16_t* out_d, int64_t* total) {
  const int n = (int)src.size();
  std::vector<XfChunkH> ch(n);
  int64_t off = 0, max_len = 0;
  for (int i = 0; i < n; ++i) {
    ch[i].src = src[i];
    ch[i].len = lens[i];
    ch[i].out_off = off;
    off += lens[i] - (i + 1 < n ? nf : 0);
    max_len = std::max(max_len, lens[i]);
  }
  *total = off;
  void* scratch = nullptr;
  const size_t b_ch = (size_t)n * sizeof(XfChunkH), b_i = (size_t)n * 4, b_f = (size_t)nf * 8;
  const size_t o_clip = (b_ch + 255) / 256 * 256, o_has = o_clip + (b_i + 255) / 256 * 256,
               o_ratio = o_has + (b_i + 255) / 256 * 256, o_fo = o_ratio + (b_i + 255) / 256 * 256,
               o_fi = o_fo + (b_f + 255) / 256 * 256, bytes = o_fi + b_f;
  CK(cudaMallocAsync(&scratch, bytes, e->st));
  uint8_t* s8 = static_cast<uint8_t*>(scratch);
  CK(cudaMemcpyAsync(s8, ch.data(), b_ch, cudaMemcpyHostToDevice, e->st));
  CK(cudaMemcpyAsync(s8 + o_fo, fade_out_h, b_f, cudaMemcpyHostToDevice, e->st));
  CK(cudaMemcpyAsync(s8 + o_fi, fade_in_h, b_f, cudaMemcpyHostToDevice, e->st));
  CK(cudaStreamSynchronize(e->st));      // `ch` is a pageable host vector that dies with this frame
  launch_crossfade(s8, n, max_len, nf, reinterpret_cast<const double*>(s8 + o_fo),
                   reinterpret_cast<const double*>(s8 + o_fi), reinterpret_cast<int*>(s8 + o_clip),
                   reinterpret_cast<int*>(s8 + o_has), reinterpret_cast<float*>(s8 + o_ratio), out_d, e->st);
  e->launches += 3;
  CK(cudaFreeAsync(scratch, e->st));
  CKL();
  return 0;
}

static int crossfade_args_ok(int n, int nf, const double* fo, const double* fi) {
  if (n < 2 || nf < 1 || nf > 6144 || !fo || !fi)
    return fail(VV_ERR_ARG, "cross-fade needs >= 2 chunks, 1..6144 fade samples and both fade tables");
  return 0;
}

extern "C" int vv_crossfade_pcm(vv_engine* e, const int16_t* const* waves, const int64_t* lens, int n,
                                const double* fade_out, const double* fade_in, int n_fade, int16_t* pcm_out,
                                int64_t capacity, int64_t* n_out) {
  if (!e || !waves || !lens || !pcm_out) return fail(VV_ERR_ARG, "vv_crossfade_pcm: null argument");
  TRY(crossfade_args_ok(n, n_fade, fade_out, fade_in));
  ENG_LOCK(e);
  CK(cudaSetDevice(e->device));
  int64_t sum = 0;
  for (int i = 0; i < n; ++i) {
    if (lens[i] < 2 * (int64_t)n_fade) return fail(VV_ERR_ARG, "chunk %d holds %lld samples, fewer than two cross-fades", i, (long long)lens[i]);
    sum += lens[i];
  }
  const int64_t total = sum - (int64_t)n_fade * (n - 1);
  if (capacity < total) return fail(VV_ERR_ARG, "vv_crossfade_pcm: capacity %lld < %lld", (long long)capacity, (long long)total);
  int16_t *in_d = nullptr, *out_d = nullptr;
  CK(cudaMallocAsync(&in_d, (size_t)sum * 2, e->st));
  CK(cudaMallocAsync(&out_d, (size_t)total * 2, e->st));
  std::vector<const int16_t*> src(n);
  std::vector<int64_t> ln(lens, lens + n);
  int64_t o = 0;
  for (int i = 0; i < n; ++i) {
    CK(cudaMemcpyAsync(in_d + o, waves[i], (size_t)lens[i] * 2, cudaMemcpyHostToDevice, e->st));
    src[i] = in_d + o;
    o += lens[i];
  }
  int64_t got = 0;
  int rc = crossfade_device(e, src, ln, n_fade, fade_out, fade_in, out_d, &got);
  if (rc == 0) {
    cudaMemcpyAsync(pcm_out, out_d, (size_t)got * 2, cudaMemcpyDeviceToHost, e->st);
    if (n_out) *n_out = got;
  }
  cudaFreeAsync(in_d, e->st);
  cudaFreeAsync(out_d, e->st);
  CK(cudaStreamSynchronize(e->st));
  return rc;
}

// cross-fade of the decoded chunks order[0..n) of a batch, straight from its device PCM
static int batch_crossfade(vv_batch* b, const int32_t* order, int n, const double* fo, const double* fi, int nf,
                           int16_t* pcm_out, int64_t capacity, int64_t* n_out) {
  vv_engine* e = b->e;
  TRY(crossfade_args_ok(n, nf, fo, fi));
  TRY(decode_all(b));
  std::vector<const int16_t*> src(n);
  std::vector<int64_t> ln(n);
  int64_t sum = 0;
  for (int i = 0; i < n; ++i) {
    const int c = order ? order[i] : i;
    if (c < 0 || c >= b->B) return fail(VV_ERR_ARG, "cross-fade: chunk index %d out of range", c);
    if (b->pcm_len[c] < 2 * (int64_t)nf)
      return fail(VV_ERR_ARG, "chunk %d holds %lld samples, fewer than two cross-fades", c, (long long)b->pcm_len[c]);
    src[i] = b->pcm_d + b->pcm_off[c];
    ln[i] = b->pcm_len[c];
    sum += ln[i];
  }
  const int64_t total = sum - (int64_t)nf * (n - 1);
  if (capacity < total) return fail(VV_ERR_ARG, "cross-fade: capacity %lld < %lld", (long long)capacity, (long long)total);
  int16_t* out_d = nullptr;
  CK(cudaMallocAsync(&out_d, (size_t)total * 2, e->st));
  int64_t got = 0;
  int rc = crossfade_device(e, src, ln, nf, fo, fi, out_d, &got);
  if (rc == 0) {
    cudaMemcpyAsync(pcm_out, out_d, (size_t)got * 2, cudaMemcpyDeviceToHost, e->st);
    if (n_out) *n_out = got;
  }
  cudaFreeAsync(out_d, e->st);
  CK(cudaStreamSynchronize(e->st));
  return rc;
}

extern "C" int vv_batch_crossfade(vv_batch* b, const int32_t* order, int n, const double* fade_out,
                                  const double* fade_in, int n_fade, int16_t* pcm_out, int64_t capacity,
                                  int64_t* n_out) {
  if (!b || !b->e || !pcm_out) return fail(VV_ERR_ARG, "vv_batch_crossfade: null argument");
  ENG_LOCK(b->e);
  CK(cudaSetDevice(b->e->device));
  return batch_crossfade(b, order, n, fade_out, fade_in, n_fade, pcm_out, capacity, n_out);
}

// ------------------------------------------------------------------------------------------------ kernel-level ABI
static GemmEpi to_epi(vv_engine* e, const vv_gemm_epilogue* p) {
  GemmEpi g;
  if (!p) return g;
  g.bias = p->bias; g.gate = p->gate; g.resid = p->resid; g.ld_resid = p->ld_resid;
  g.out_f32 = p->out_f32; g.ld_f32 = p->ld_f32;
  g.out_bf16 = reinterpret_cast<bf16*>(p->out_bf16); g.ld_bf16 = p->ld_bf16;
  g.row_mask = p->row_mask; g.row_pos = p->row_pos; g.rope_cs = e->rope_cs;
  g.rope_dim = p->rope_dim; g.rope_off2 = p->rope_off2; g.act = p->act;
  return g;
}

extern "C" int vv_gemm_bf16(vv_engine* e, const void* A, int lda, const void* Bw, int ldb, int M, int N, int K,
                            const vv_gemm_epilogue* epi, int bn) {
  if (!e || !A || !Bw || M <= 0 || N <= 0 || K <= 0 || (K % 64) || (lda % 8) || (ldb % 8))
    return fail(VV_ERR_ARG, "vv_gemm_bf16: bad argument (K %% 64 == 0, ld %% 8 == 0 required)");
  ENG_LOCK(e);
  if (bn != 64 && bn != 128 && bn != 256 && bn != 512) bn = pick_tile(M, N, K, e->num_sms);
  CK(cudaSetDevice(e->device));
  GemmOp op;
  op.s.M = M; op.s.N = N; op.s.K = K;
  op.bn = bn;
  const GemmEpi ge = to_epi(e, epi);
  if (bn == 512 && !gemm_pair_supported(op.s, ge))
    return fail(VV_ERR_ARG, "vv_gemm_bf16: bn=512 (CTA pair) needs N %% 256 == 0 and 16-byte aligned bias/gate");
  op.tA = make_tmap_bf16(A, M, K, lda, 128);
  op.tB = make_tmap_bf16(Bw, N, K, ldb, bn == 512 ? 128 : bn);
  if (bn == 512) op.tBt = make_tmap_bf16(Bw, N, K, ldb, 32);
  run_gemm(e, op, ge);
  CKL();
  return 0;
}

extern "C" int vv_conv_rows_bf16(vv_engine* e, const void* X, int ldx, const void* Wt, int M, int groups, int taps,
                                 const vv_gemm_epilogue* epi) {
  if (!e || !X || !Wt || M <= 0 || groups <= 0 || taps <= 0 || (ldx % 8)) return fail(VV_ERR_ARG, "vv_conv_rows_bf16: bad argument");
  ENG_LOCK(e);
  CK(cudaSetDevice(e->device));
  GemmOp op = make_conv_op(e, reinterpret_cast<const bf16*>(X), ldx, M, M, reinterpret_cast<const bf16*>(Wt), groups, taps);
  run_gemm(e, op, to_epi(e, epi));
  CKL();
  return 0;
}

extern "C" int vv_attention_bf16(vv_engine* e, const void* qkv, void* out, int total_rows, const int32_t* seq_off,
                                 const int32_t* seq_len, int n_seq, int heads) {
  if (!e || !qkv || !out || !seq_off || !seq_len || n_seq <= 0 || heads <= 0) return fail(VV_ERR_ARG, "vv_attention_bf16: bad argument");
  ENG_LOCK(e);
  CK(cudaSetDevice(e->device));
  const int dim = heads * 64;
  std::vector<int32_t> ts, tq;
  for (int s = 0; s < n_seq; ++s)
    for (int q0 = 0; q0 < seq_len[s]; q0 += attn_q_tile()) {
      ts.push_back(s);
      tq.push_back(q0);
    }
  std::vector<void*> tmp;
  int32_t *so, *sl, *tsd, *tqd;
  TRY(dev_alloc(tmp, &so, n_seq));
  TRY(dev_alloc(tmp, &sl, n_seq));
  TRY(dev_alloc(tmp, &tsd, ts.size()));
  TRY(dev_alloc(tmp, &tqd, tq.size()));
  CK(cudaMemcpy(so, seq_off, n_seq * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(sl, seq_len, n_seq * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(tsd, ts.data(), ts.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(tqd, tq.data(), tq.size() * 4, cudaMemcpyHostToDevice));
  CUtensorMap tm = make_tmap_bf16(qkv, total_rows, 3 * dim, 3 * dim, 128);
  AttnParams p;
  p.seq_off = so; p.seq_len = sl; p.tile_seq = tsd; p.tile_q0 = tqd; p.n_tiles = (int)ts.size();
  p.heads = heads; p.dim = dim; p.out = reinterpret_cast<bf16*>(out);
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  launch_attention(tm, p, e->st);
  e->launches++;
  cudaError_t r = cudaStreamSynchronize(e->st);
  for (void* q : tmp) cudaFree(q);
  if (r != cudaSuccess) return fail(VV_ERR_CUDA, "attention kernel failed: %s", cudaGetErrorString(r));
  CKL();
  return 0;
}

extern "C" int vv_ln_modulate(vv_engine* e, const float* x, int rows, int dim, const float* shift, const float* scale,
                              float eps, void* out_bf16) {
  if (!e || !x || !shift || !scale || !out_bf16 || rows <= 0 || dim % 128 || dim > 2048) return fail(VV_ERR_ARG, "vv_ln_modulate: bad argument");
  ENG_LOCK(e);
  CK(cudaSetDevice(e->device));
  launch_ln_mod(x, rows, dim, shift, scale, eps, reinterpret_cast<bf16*>(out_bf16), e->st);
  e->launches++;
  CKL();
  return 0;
}
