// noahmp_b200 -- C ABI implementation (include/noahmp_b200.h): context, table upload, kernel dispatch,
// the host-buffer entry point and the synthetic forcing generator.
#include <cuda_runtime.h>

#include <sched.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "../../include/noahmp_b200.h"
#include "column_step.cuh"
#include "forcing_prep.cuh"

using namespace nmp;

static_assert(NSTATE == NOAHMP_NSTATE && NSTATE_CORE == NOAHMP_NSTATE_CORE && NFORCING == NOAHMP_NFORCING &&
                  NSTATIC_F == NOAHMP_NSTATIC_F && NDIAG == NOAHMP_NDIAG && NBUDGET == NOAHMP_NBUDGET &&
                  NGEN == NOAHMP_NGEN && NSOILP == NOAHMP_NSOILP && NCOLORP == NOAHMP_NCOLORP &&
                  NVEGF == NOAHMP_NVEGF && NVEGI == NOAHMP_NVEGI,
              "header and device layout disagree");

// A few host threads that run the chunks of one parallel loop (the record packer); the calling thread takes part.
class HostPool {
 public:
  explicit HostPool(int nthreads) {
    for (int k = 1; k < nthreads; ++k) th_.emplace_back([this] { loop(); });
  }
  ~HostPool() {
    {
      std::lock_guard<std::mutex> g(mu_);
      stop_ = true;
    }
    go_.notify_all();
    for (auto& t : th_) t.join();
  }
  int size() const { return (int)th_.size() + 1; }
  void run(int nchunks, const std::function<void(int)>& fn) {
    if (th_.empty() || nchunks <= 1) {
      for (int k = 0; k < nchunks; ++k) fn(k);
      return;
    }
    {
      std::lock_guard<std::mutex> g(mu_);
      fn_ = &fn; nchunks_ = nchunks; next_.store(0); pending_ = (int)th_.size(); ++gen_;
    }
    go_.notify_all();
    work();
    std::unique_lock<std::mutex> g(mu_);
    done_.wait(g, [this] { return pending_ == 0; });
  }

 private:
  void work() {
    for (int k; (k = next_.fetch_add(1)) < nchunks_;) (*fn_)(k);
  }
  void loop() {
    unsigned long long seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> g(mu_);
        go_.wait(g, [&] { return stop_ || gen_ != seen; });
        if (stop_) return;
        seen = gen_;
      }
      work();
      std::lock_guard<std::mutex> g(mu_);
      if (--pending_ == 0) done_.notify_one();
    }
  }
  std::vector<std::thread> th_;
  std::mutex mu_;
  std::condition_variable go_, done_;
  const std::function<void(int)>* fn_ = nullptr;
  int nchunks_ = 0, pending_ = 0;
  std::atomic<int> next_{0};
  unsigned long long gen_ = 0;
  bool stop_ = false;
};

struct noahmp_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool have_opts = false, have_tables = false;
  OptRuntime opt{};
  bool opt_is_default = false;
  Tables host_tables{};
  Tables* dev_tables = nullptr;
  double* dev_budget_scratch = nullptr;   // NBUDGET doubles
  double* dev_budget_rep = nullptr;       // NBUDGET_REP x NBUDGET: what the CTAs of a launch add into (column_step.cuh), zero between launches
  // grow-only DEVICE staging buffers of noahmp_run_host
  float* d_state = nullptr; float* d_forcing = nullptr; float* d_static_f = nullptr; int* d_static_i = nullptr;
  float* d_diag = nullptr;
  size_t cap_state = 0, cap_forcing = 0, cap_static_f = 0, cap_static_i = 0, cap_diag = 0;
  // noahmp_advance_host: copy stream, double-buffered forcing, hand-off events
  cudaStream_t copy_stream = nullptr;
  float* fbuf[2] = {nullptr, nullptr};
  size_t cap_fbuf = 0;
  cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  // noahmp_advance_records_host: ring of record slots on the device, one last-use event per slot
  float* rring = nullptr;
  size_t cap_rring = 0;            // bytes
  int rring_slots = 0;
  cudaEvent_t* ev_slot = nullptr;  // [rring_slots]
  cudaEvent_t ev_rcopied = nullptr;
  // zero suppression of record planes before they cross the host link (noahmp_set_record_packing)
  unsigned pack_planes = (1u << F_SOLDN) | (1u << F_PRCP);
  int pack_threads = 0;             // 0: default_pack_threads()
  HostPool* pool = nullptr;
  uint32_t* stage = nullptr;        // page-locked: per staging area and candidate plane, occupancy | before | values
  size_t cap_stage = 0;             // bytes
  cudaEvent_t ev_stage[4] = {nullptr, nullptr, nullptr, nullptr};   // the copies out of a staging area have completed
  unsigned stage_next = 0;          // staging areas are used in turn
  std::vector<uint32_t> pack_scratch;   // the packer's intermediate copy of one plane's non-zero words
  std::vector<int> slot_sparse;     // which planes of the record a ring slot holds are packed
  double pack_seconds = 0.0;        // wall time the calling thread has spent packing (its helpers work inside it)
  int64_t h2d_bytes = 0;            // bytes the advance_* entry points have copied host -> device
  int64_t launches = 0;
  std::string err;
};

namespace {

const char* kStateNames[NSTATE] = {
    "ALBOLD", "SNEQVO", "EAH", "FWET", "CANLIQ", "CANICE", "TV", "TG", "QSNOW", "ISNOW", "SNOWH", "SNEQV", "ZWT", "WA",
    "WT", "TAUSS", "STC-2", "STC-1", "STC+0", "STC+1", "STC+2", "STC+3", "STC+4", "SOILWAT1", "SOILWAT2", "SOILWAT3",
    "SOILWAT4", "SMC1", "SMC2", "SMC3", "SMC4", "ZSNSO-2", "ZSNSO-1", "ZSNSO+0", "ZSNSO+1", "ZSNSO+2", "ZSNSO+3",
    "ZSNSO+4", "SNICE-2", "SNICE-1", "SNICE+0", "SNLIQ-2", "SNLIQ-1", "SNLIQ+0", "QSFC", "TAH", "CM", "CH", "WSLAKE",
    "LFMASS", "RTMASS", "STMASS", "WOOD", "STBLCP", "FASTCP", "LAI", "SAI"};
const char* kForcingNames[NFORCING] = {"SFCTMP", "SFCPRS", "UU", "VV", "Q2", "SOLDN", "LWDN", "PRCP", "COSZ"};
const char* kDiagNames[NDIAG] = {
    "FSA", "FSR", "FIRA", "FSH", "SSOIL", "FCEV", "FGEV", "FCTR", "ECAN", "ETRAN", "EDIR", "TRAD", "TGB", "TGV", "T2MV",
    "T2MB", "Q2V", "Q2B", "RUNSRF", "RUNSUB", "APAR", "PSN", "SAV", "SAG", "FSNO", "NEE", "GPP", "NPP", "FVEG", "ALBEDO",
    "QSNBOT", "PONDING", "PONDING1", "PONDING2", "RSSUN", "RSSHA", "BGAP", "WGAP", "CHV", "CHB", "EMISSI", "SHG", "SHC",
    "SHB", "EVG", "EVB", "GHV", "GHB", "IRG", "IRC", "IRB", "TR", "EVC", "CHLEAF", "CHUC", "CHV2", "CHB2", "FPICE",
    "ERRWAT", "ERRSW", "ERRENG", "ITER", "FLAGS", "ACCPRCP", "ACCECAN", "ACCETRAN", "ACCEDIR", "SFCRNOFF", "UGDRNOFF"};
const char* kBudgetNames[NBUDGET] = {"NCELL", "NFLAGGED", "ERRWAT_SUM", "ERRWAT_ABS", "ERRWAT_MAXABS", "FSA", "FSR",
                                     "FIRA", "FSH", "LH", "SSOIL", "PRCP_DT", "ET_DT", "RUNOFF_DT", "SNEQV", "SOILWATER"};

int fail(noahmp_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  return code;
}
int cuda_fail(noahmp_ctx* c, cudaError_t e, const char* where) {
  return fail(c, -100 - (int)e, std::string(where) + ": " + cudaGetErrorString(e));
}
#define CUDA_TRY(ctx, call)                                  \
  do {                                                       \
    cudaError_t _e = (call);                                 \
    if (_e != cudaSuccess) return cuda_fail(ctx, _e, #call); \
  } while (0)

typedef void (*StepKernel)(const StepArgs);

StepKernel pick_kernel(bool is_default, bool diag, bool budget, bool packed = false) {
#ifdef NMP_EXPERIMENT_ONE_KERNEL   // tuning builds only (tools/tune_builds.sh): compile just the benchmarked variant
  (void)is_default; (void)diag; (void)budget; (void)packed;
  return column_step_kernel<OptDefault, false, false, false>;
#else
  // zero-suppressed record planes: only noahmp_advance_records_host produces them, and it launches without diagnostics
  if (packed) return is_default ? column_step_kernel<OptDefault, false, false, true, true> : column_step_kernel<OptRuntime, true, false, true, true>;
  if (is_default) {
    if (diag) return column_step_kernel<OptDefault, false, true, true>;
    if (budget) return column_step_kernel<OptDefault, false, false, true>;
    return column_step_kernel<OptDefault, false, false, false>;
  }
  if (diag) return column_step_kernel<OptRuntime, true, true, true>;
  if (budget) return column_step_kernel<OptRuntime, true, false, true>;
  return column_step_kernel<OptRuntime, true, false, false>;
#endif
}

// ---- synthetic forcing: integer hash + explicit round-to-nearest fp32 ops (no FMA contraction), so it
// is bit-identical to noahmp_b200/synth.py: forcing_step ----------------------------------------------
__device__ __forceinline__ uint32_t mix32(uint32_t h) {
  h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
  return h;
}
__device__ __forceinline__ float uniform24(uint32_t key, uint32_t col) {
  uint32_t h = mix32(col ^ key);
  h = mix32(h + 0x27D4EB2Fu);
  return __fmul_rn((float)(h >> 8), 1.0f / 16777216.0f);
}
struct SynthCoef {
  float t_mean, t_lat, t_noise, sw_max, lw_mean, lw_tcoef, lw_noise, p_mean, p_noise, wind_min, wind_range, q_mean,
      q_tcoef, q_noise, prcp_prob, inv_prob, prcp_scale;
  float sd, cdch, t_diur, t_seas;
  uint32_t seed;
  int step;
};
#define ADD(a, b) __fadd_rn((a), (b))
#define SUB(a, b) __fsub_rn((a), (b))
#define MUL(a, b) __fmul_rn((a), (b))
__global__ void __launch_bounds__(256) synth_forcing_kernel(int64_t ncol, int64_t col0, const int64_t* __restrict__ ids, SynthCoef p,
                                                            const float* __restrict__ sincos,
                                                            float* __restrict__ forcing) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncol) return;
  // the generator hashes the column's id in the GENERATOR's numbering: col0 + i, or ids[i] when the columns were re-ordered
  const uint32_t col = (uint32_t)(uint64_t)(ids ? ids[i] : col0 + i);
  float u[12];
#pragma unroll
  for (int s = 0; s < 12; ++s) {
    const uint32_t key = p.seed * 0x9E3779B1u + (uint32_t)p.step * 0x85EBCA77u + (uint32_t)s * 0xC2B2AE3Du;
    u[s] = uniform24(key, col);
  }
  const float sinlat = sincos[i], coslat = sincos[ncol + i];
  const float g1 = MUL(SUB(ADD(ADD(u[0], u[1]), ADD(u[2], u[3])), 2.0f), 1.7320508f);
  const float g2 = MUL(SUB(ADD(ADD(u[4], u[5]), ADD(u[6], u[7])), 2.0f), 1.7320508f);
  const float s2 = MUL(sinlat, sinlat);
  const float tcore = ADD(ADD(p.t_mean, MUL(ADD(p.t_lat, p.t_seas), s2)), p.t_diur);
  const float T = ADD(tcore, MUL(p.t_noise, g1));
  const float cosz = fminf(fmaxf(ADD(MUL(sinlat, p.sd), MUL(coslat, p.cdch)), -1.0f), 1.0f);
  const float sw = MUL(p.sw_max, fmaxf(cosz, 0.0f));
  const float lw = ADD(ADD(p.lw_mean, MUL(p.lw_tcoef, SUB(T, p.t_mean))), MUL(p.lw_noise, g2));
  const float ps = ADD(p.p_mean, MUL(p.p_noise, SUB(MUL(2.0f, u[8]), 1.0f)));
  const float wind = ADD(p.wind_min, MUL(p.wind_range, u[9]));
  float e = fmaxf(ADD(1.0f, MUL(MUL(p.q_tcoef, SUB(tcore, p.t_mean)), 0.0625f)), 0.05f);
  e = MUL(e, e); e = MUL(e, e); e = MUL(e, e); e = MUL(e, e);
  float q = MUL(MUL(p.q_mean, e), ADD(1.0f, MUL(p.q_noise, SUB(MUL(2.0f, u[10]), 1.0f))));
  q = fmaxf(q, 1.0e-4f);
  const float ur = MUL(u[11], p.inv_prob);
  const float pr = (u[11] < p.prcp_prob) ? MUL(p.prcp_scale, MUL(ur, ur)) : 0.0f;
  forcing[F_SFCTMP * ncol + i] = T;
  forcing[F_SFCPRS * ncol + i] = ps;
  forcing[F_UU * ncol + i] = MUL(wind, 0.8f);
  forcing[F_VV * ncol + i] = MUL(wind, 0.6f);
  forcing[F_Q2 * ncol + i] = q;
  forcing[F_SOLDN * ncol + i] = sw;
  forcing[F_LWDN * ncol + i] = lw;
  forcing[F_PRCP * ncol + i] = pr;
  forcing[F_COSZ * ncol + i] = cosz;
}

// ---- zero suppression of one record plane (host) --------------------------------------------------------------------
// occupancy[g] bit b = word 32 g + b of the plane is not the bit pattern 0; before[g] = non-zero words in groups < g;
// values = the non-zero words in order.  The plane is read ONCE: chunks of groups are compacted in parallel into `scratch`
// (each at the position of its first column), the chunk totals are summed, and a second parallel loop copies each chunk's
// words to their final place and adds the chunk's base to its `before` words.  A sample of the plane is looked at first:
// a plane that is clearly fuller than `max_fill` is not packed at all (the caller sends it as it is).  Returns the number
// of non-zero words, or -1 when the plane is (or is estimated to be) more than `max_fill` non-zero.
inline uint32_t occupancy32(const uint32_t* x) {
#if defined(__SSE2__)
  const __m128i z = _mm_setzero_si128();
  uint32_t zero = 0;
  for (int k = 0; k < 8; ++k) {
    const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(x + 4 * k));
    zero |= (uint32_t)_mm_movemask_ps(_mm_castsi128_ps(_mm_cmpeq_epi32(v, z))) << (4 * k);
  }
  return ~zero;
#else
  uint32_t m = 0;
  for (int b = 0; b < 32; ++b) m |= (uint32_t)(x[b] != 0u) << b;
  return m;
#endif
}
int64_t pack_plane(HostPool* pool, const float* plane, int64_t n, uint32_t* occupancy, uint32_t* before, float* values,
                   uint32_t* scratch, double max_fill) {
  const uint32_t* x = reinterpret_cast<const uint32_t*>(plane);
  const int64_t G = (n + 31) >> 5, Gfull = n >> 5;
  if (max_fill < 1.0 && n >= (1 << 16)) {   // 4096 evenly spaced words: is the plane worth reading?
    const int64_t stride = n / 4096;
    int64_t hit = 0;
    for (int64_t k = 0; k < 4096; ++k) hit += x[k * stride + (k & 31)] != 0u;
    if ((double)hit > (max_fill + 0.1) * 4096.0) return -1;
  }
  const int workers = pool ? pool->size() : 1;
  int nchunks = (int)std::min<int64_t>((int64_t)workers * 4, std::max<int64_t>(1, G / 2048));
  const int64_t per = (G + nchunks - 1) / nchunks;
  nchunks = (int)((G + per - 1) / per);
  std::vector<int64_t> count((size_t)nchunks + 1, 0);
  auto compact = [&](int c) {
    const int64_t g0 = c * per, g1 = std::min(G, g0 + per);
    uint32_t* out = scratch + 32 * g0;
    int64_t k = 0;
    for (int64_t g = g0; g < g1; ++g) {
      const uint32_t* src = x + 32 * g;
      uint32_t m;
      if (g < Gfull) m = occupancy32(src);
      else {
        m = 0;
        for (int64_t i = 32 * g; i < n; ++i) m |= (uint32_t)(x[i] != 0u) << (i - 32 * g);
      }
      occupancy[g] = m;
      before[g] = (uint32_t)k;
      if (m == 0xffffffffu) {
        std::memcpy(out + k, src, 128);
        k += 32;
      } else {
        for (; m; m &= m - 1) out[k++] = src[__builtin_ctz(m)];
      }
    }
    count[(size_t)c + 1] = k;
  };
  if (pool) pool->run(nchunks, compact);
  else for (int c = 0; c < nchunks; ++c) compact(c);
  for (int c = 0; c < nchunks; ++c) count[(size_t)c + 1] += count[(size_t)c];
  const int64_t nnz = count[(size_t)nchunks];
  if ((double)nnz > max_fill * (double)n) return -1;
  auto place = [&](int c) {
    const int64_t g0 = c * per, g1 = std::min(G, g0 + per), base = count[(size_t)c];
    std::memcpy(reinterpret_cast<uint32_t*>(values) + base, scratch + 32 * g0, sizeof(uint32_t) * (size_t)(count[(size_t)c + 1] - base));
    if (base)
      for (int64_t g = g0; g < g1; ++g) before[g] += (uint32_t)base;
  };
  if (pool) pool->run(nchunks, place);
  else for (int c = 0; c < nchunks; ++c) place(c);
  return nnz;
}
// a plane more than this full travels dense: packing it would save less than the occupancy words and the second pass cost
constexpr double kPackMaxFill = 0.80;
constexpr int kStageAreas = 4;     // page-locked staging areas of the packer (noahmp_ctx::ev_stage)
int default_pack_threads() {
  cpu_set_t set;
  int allowed = (int)std::thread::hardware_concurrency();
  if (sched_getaffinity(0, sizeof(set), &set) == 0) allowed = CPU_COUNT(&set);
  return std::max(1, std::min(8, allowed));
}

std::string g_init_error = "null context";   // text of the last failed noahmp_init (there is no context to hold it)

int ensure(noahmp_ctx* c, void** p, size_t* cap, size_t need) {
  if (need <= *cap) return 0;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  CUDA_TRY(c, cudaMalloc(p, need));
  *cap = need;
  return 0;
}

}  // namespace

extern "C" {

int noahmp_init(noahmp_ctx** out, int device) {
  if (!out) return -1;
  *out = nullptr;
  noahmp_ctx* c = new noahmp_ctx();
  c->device = device;
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaMalloc(&c->dev_tables, sizeof(Tables));
  if (e == cudaSuccess) e = cudaMalloc(&c->dev_budget_scratch, sizeof(double) * NBUDGET);
  if (e == cudaSuccess) e = cudaMalloc(&c->dev_budget_rep, sizeof(double) * NBUDGET * NBUDGET_REP);
  if (e == cudaSuccess) e = cudaMemset(c->dev_budget_rep, 0, sizeof(double) * NBUDGET * NBUDGET_REP);
  if (e != cudaSuccess) {   // no context to carry the text: noahmp_last_error(NULL) returns it
    g_init_error = std::string("noahmp_init(device ") + std::to_string(device) + "): " + cudaGetErrorString(e);
    cudaGetLastError();
    if (c->dev_tables) cudaFree(c->dev_tables);
    if (c->dev_budget_scratch) cudaFree(c->dev_budget_scratch);
    if (c->dev_budget_rep) cudaFree(c->dev_budget_rep);
    delete c;
    return -100 - (int)e;
  }
  *out = c;
  return 0;
}

int noahmp_finalize(noahmp_ctx* c) {
  if (!c) return -1;
  cudaSetDevice(c->device);
  cudaFree(c->dev_tables);
  cudaFree(c->dev_budget_scratch);
  cudaFree(c->dev_budget_rep);
  cudaFree(c->d_state); cudaFree(c->d_forcing); cudaFree(c->d_static_f); cudaFree(c->d_static_i); cudaFree(c->d_diag);
  cudaFree(c->fbuf[0]); cudaFree(c->fbuf[1]);
  for (int k = 0; k < 2; ++k) {
    if (c->ev_copied[k]) cudaEventDestroy(c->ev_copied[k]);
    if (c->ev_done[k]) cudaEventDestroy(c->ev_done[k]);
  }
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  cudaFree(c->rring);
  for (int k = 0; k < c->rring_slots; ++k) cudaEventDestroy(c->ev_slot[k]);
  delete[] c->ev_slot;
  if (c->ev_rcopied) cudaEventDestroy(c->ev_rcopied);
  for (int k = 0; k < 4; ++k)
    if (c->ev_stage[k]) cudaEventDestroy(c->ev_stage[k]);
  if (c->stage) cudaFreeHost(c->stage);
  delete c->pool;
  delete c;
  return 0;
}

int noahmp_set_stream(noahmp_ctx* c, void* s) {
  if (!c) return -1;
  c->stream = (cudaStream_t)s;
  return 0;
}

const char* noahmp_last_error(const noahmp_ctx* c) { return c ? c->err.c_str() : g_init_error.c_str(); }

int noahmp_set_options(noahmp_ctx* c, const int o[12]) {
  if (!c || !o) return -1;
  // ranges of core/module_noahmp_global.f90:17-74; opt_veg is checked per column like the reference does
  const int hi[12] = {1 << 30, 2, 3, 4, 2, 2, 2, 3, 2, 3, 2, 2};
  const char* nm[12] = {"veg", "crs", "btr", "run", "sfc", "frz", "inf", "rad", "alb", "snf", "tbot", "stc"};
  for (int k = 1; k < 12; ++k)
    if (o[k] < 1 || o[k] > hi[k]) return fail(c, -2, std::string("option opt_") + nm[k] + " out of range");
  c->opt = OptRuntime{o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7], o[8], o[9], o[10], o[11]};
  c->opt_is_default = o[0] == OptDefault::veg && o[1] == OptDefault::crs && o[2] == OptDefault::btr &&
                      o[3] == OptDefault::run && o[4] == OptDefault::sfc && o[5] == OptDefault::frz &&
                      o[6] == OptDefault::inf && o[7] == OptDefault::rad && o[8] == OptDefault::alb &&
                      o[9] == OptDefault::snf && o[10] == OptDefault::tbot && o[11] == OptDefault::stc;
  c->have_opts = true;
  return 0;
}

int noahmp_set_tables(noahmp_ctx* c, const float* gen, int ngen, const float* soil, int nsoiltyp, const float* color,
                      int ncolor, const float* veg_f, const int* veg_i, int nvegtyp, const int flags[5]) {
  if (!c || !gen || !soil || !color || !veg_f || !veg_i || !flags) return -1;
  if (ngen != NGEN) return fail(c, -3, "gen table must have NOAHMP_NGEN entries");
  if (nsoiltyp < 1 || nsoiltyp > MAXSOILTYP) return fail(c, -3, "Table PARM is too big to be read");      // soil_param:48
  if (ncolor < 1 || ncolor > MAXCOLOR) return fail(c, -3, "Table COLOR is too big to be read");           // soil_param:66
  if (nvegtyp < 1 || nvegtyp > MAXVEGTYP) return fail(c, -3, "Table RAD is too big to be read");          // veg_param:97
  Tables& t = c->host_tables;
  std::memset(&t, 0, sizeof(t));
  std::memcpy(t.gen, gen, sizeof(float) * NGEN);
  std::memcpy(t.soil, soil, sizeof(float) * NSOILP * nsoiltyp);
  std::memcpy(t.color, color, sizeof(float) * NCOLORP * ncolor);
  std::memcpy(t.veg, veg_f, sizeof(float) * NVEGF * nvegtyp);
  std::memcpy(t.vegi, veg_i, sizeof(int) * NVEGI * nvegtyp);
  std::memcpy(t.flags, flags, sizeof(int) * 5);
  t.nsoiltyp = nsoiltyp; t.ncolor = ncolor; t.nvegtyp = nvegtyp;
  CUDA_TRY(c, cudaSetDevice(c->device));
  CUDA_TRY(c, cudaMemcpy(c->dev_tables, &t, sizeof(Tables), cudaMemcpyHostToDevice));
  c->have_tables = true;
  return 0;
}

// where the forcing of a launch comes from: prepared planes [nsteps][9][ncol], or records + per-step table + geo planes
struct ForcingSource {
  const float* forcing = nullptr; int64_t forcing_stride = 0;
  const float* records = nullptr; int64_t record_stride = 0; const float* geo = nullptr; const noahmp_step_forcing* steps = nullptr;
  const int* sp0 = nullptr; const int* sp1 = nullptr;   // per step: packed planes of rec0 / rec1 (null: dense records)
};

static int run_impl(noahmp_ctx* c, const noahmp_run_params* rp, int64_t ncol, int nsteps, float* state_dev,
                    const ForcingSource& src, const float* static_f_dev,
                    const int32_t* static_i_dev, float* diag_dev, double* budget_dev, int final_launch) {
  if (!c) return -1;
  if (!c->have_opts) return fail(c, -4, "noahmp_set_options has not been called");
  if (!c->have_tables) return fail(c, -4, "noahmp_set_tables has not been called");
  if (ncol < 0 || nsteps < 0) return fail(c, -1, "negative size");
  if (ncol == 0 || nsteps == 0) return 0;   // empty batch: nothing to do (an empty tensor has a null pointer)
  if (!rp || !state_dev || !static_f_dev || !static_i_dev) return fail(c, -1, "null argument");
  if (!src.forcing && !(src.records && src.geo && src.steps)) return fail(c, -1, "null argument");
  if (src.records) {
    if (nsteps > MAX_RECORD_STEPS) return fail(c, -1, "records-driven launch longer than NOAHMP_MAX_RECORD_STEPS");
    for (int s = 0; s < nsteps; ++s) {
      if (src.steps[s].rec0 < 0 || src.steps[s].rec1 < 0) return fail(c, -5, "negative record index");
      if (!(src.steps[s].weight >= 0.0f && src.steps[s].weight <= 1.0f)) return fail(c, -5, "interpolation weight must be in [0,1]");
    }
  }
  if (!(rp->dt > 0.0f) || rp->yearlen <= 0) return fail(c, -5, "dt and yearlen must be positive");
  if (!std::isfinite(rp->julian0)) return fail(c, -5, "julian0 must be finite (it is wrapped into [0, yearlen) per step)");
  for (int k = 0; k < 4; ++k)
    if (!(rp->zsoil[k] < (k ? rp->zsoil[k - 1] : 0.0f))) return fail(c, -5, "zsoil must be negative and decreasing");
  CUDA_TRY(c, cudaSetDevice(c->device));
  StepArgs a;
  a.state = state_dev; a.forcing = src.forcing; a.forcing_stride = src.forcing_stride; a.static_f = static_f_dev;
  a.static_i = static_i_dev; a.diag = diag_dev;
  a.records = src.records; a.record_stride = src.record_stride; a.geo = src.geo;
  std::memset(a.tab, 0, sizeof(a.tab));
  bool packed = false;
  if (src.records)
    for (int s = 0; s < nsteps; ++s) {
      const noahmp_step_forcing& q = src.steps[s];
      a.tab[s] = StepForcing{q.rec0, q.rec1, q.weight, q.sd, q.cd, q.ch0, q.sh0, src.sp0 ? src.sp0[s] : 0, src.sp1 ? src.sp1[s] : 0};
      packed = packed || a.tab[s].sp0 != 0 || a.tab[s].sp1 != 0;
    }
  if (packed && (diag_dev != nullptr || budget_dev == nullptr)) return fail(c, -1, "packed records are launched with a budget and without diagnostics");
  // every instantiation with diagnostics also sums the budget: a launch with diag_dev but without budget_dev folds into the scratch vector
  const bool sums = budget_dev != nullptr || diag_dev != nullptr;
  a.budget = sums ? c->dev_budget_rep : nullptr;
  a.tables = c->dev_tables; a.ncol = ncol; a.nsteps = nsteps; a.final_launch = final_launch;
  a.rc.dt = rp->dt;
  for (int k = 0; k < 4; ++k) a.rc.zsoil[k] = rp->zsoil[k];
  a.rc.zref = rp->zref; a.rc.co2_frac = rp->co2_frac; a.rc.o2_frac = rp->o2_frac; a.rc.foln = rp->foln;
  a.rc.yearlen = rp->yearlen; a.rc.julian0 = rp->julian0;
  a.opt = c->opt;
  StepKernel k = pick_kernel(c->opt_is_default, diag_dev != nullptr, budget_dev != nullptr, packed);
  const int block = NMP_BLOCK;
  const int64_t grid = (ncol + block - 1) / block;
  if (grid > 0x7fffffffLL) return fail(c, -6, "ncol too large for one launch");
  k<<<(unsigned)grid, block, 0, c->stream>>>(a);
  c->launches += 1;
  CUDA_TRY(c, cudaGetLastError());
  if (sums) {
    budget_combine_kernel<<<1, 32, 0, c->stream>>>(c->dev_budget_rep, budget_dev ? budget_dev : c->dev_budget_scratch);
    CUDA_TRY(c, cudaGetLastError());
  }
  return 0;
}

int noahmp_run(noahmp_ctx* c, const noahmp_run_params* rp, int64_t ncol, int nsteps, float* state_dev,
               const float* forcing_dev, int64_t forcing_step_stride, const float* static_f_dev,
               const int32_t* static_i_dev, float* diag_dev, double* budget_dev) {
  ForcingSource src;
  src.forcing = forcing_dev; src.forcing_stride = forcing_step_stride;
  if (!forcing_dev && ncol > 0 && nsteps > 0) return fail(c, -1, "null argument");
  return run_impl(c, rp, ncol, nsteps, state_dev, src, static_f_dev, static_i_dev, diag_dev, budget_dev, 1);
}

static int run_records_impl(noahmp_ctx* c, const noahmp_run_params* rp, int64_t ncol, int nsteps, float* state_dev,
                            const float* records_dev, int64_t record_stride, const noahmp_step_forcing* steps,
                            const float* geo_dev, const float* static_f_dev, const int32_t* static_i_dev, float* diag_dev,
                            double* budget_dev, int final_launch) {
  if (!c) return -1;
  if (ncol < 0 || nsteps < 0) return fail(c, -1, "negative size");
  if (ncol == 0 || nsteps == 0) return 0;
  if (!rp || !records_dev || !steps || !geo_dev) return fail(c, -1, "null argument");
  for (int s0 = 0; s0 < nsteps; s0 += MAX_RECORD_STEPS) {   // the step table travels as kernel parameters: 32 steps a launch
    const int n = nsteps - s0 < MAX_RECORD_STEPS ? nsteps - s0 : MAX_RECORD_STEPS;
    noahmp_run_params r = *rp;
    r.julian0 = rp->julian0 + (double)s0 * (double)rp->dt / 86400.0;
    ForcingSource src;
    src.records = records_dev; src.record_stride = record_stride; src.geo = geo_dev; src.steps = steps + s0;
    const int rc = run_impl(c, &r, ncol, n, state_dev, src, static_f_dev, static_i_dev, diag_dev, budget_dev,
                            (final_launch && s0 + n >= nsteps) ? 1 : 0);
    if (rc) return rc;
  }
  return 0;
}

int noahmp_run_records(noahmp_ctx* c, const noahmp_run_params* rp, int64_t ncol, int nsteps, float* state_dev,
                       const float* records_dev, int64_t record_stride, const noahmp_step_forcing* steps, const float* geo_dev,
                       const float* static_f_dev, const int32_t* static_i_dev, float* diag_dev, double* budget_dev) {
  return run_records_impl(c, rp, ncol, nsteps, state_dev, records_dev, record_stride, steps, geo_dev, static_f_dev, static_i_dev,
                          diag_dev, budget_dev, 1);
}

int64_t noahmp_run_host(noahmp_ctx* c, const noahmp_run_params* rp, int64_t ncol, int nsteps, float* state,
                        const float* forcing, const float* static_f, const int32_t* static_i, float* diag,
                        double* budget) {
  if (!c) return -1;
  if (!state || !forcing || !static_f || !static_i) return fail(c, -1, "null argument");
  if (ncol <= 0 || nsteps <= 0) return 0;
  CUDA_TRY(c, cudaSetDevice(c->device));
  const size_t n = (size_t)ncol;
  const bool ext = !c->opt_is_default;
  const size_t nplanes = ext ? NSTATE : NSTATE_CORE;   // the 13 extended planes are dead for the shipped options
  int rc;
  if ((rc = ensure(c, (void**)&c->d_state, &c->cap_state, sizeof(float) * NSTATE * n))) return rc;
  if ((rc = ensure(c, (void**)&c->d_forcing, &c->cap_forcing, sizeof(float) * NFORCING * n * (size_t)nsteps))) return rc;
  if ((rc = ensure(c, (void**)&c->d_static_f, &c->cap_static_f, sizeof(float) * NSTATIC_F * n))) return rc;
  if ((rc = ensure(c, (void**)&c->d_static_i, &c->cap_static_i, sizeof(int) * n))) return rc;
  if (diag && (rc = ensure(c, (void**)&c->d_diag, &c->cap_diag, sizeof(float) * NDIAG * n))) return rc;
  cudaStream_t st = c->stream;
  CUDA_TRY(c, cudaMemcpyAsync(c->d_state, state, sizeof(float) * nplanes * n, cudaMemcpyHostToDevice, st));
  CUDA_TRY(c, cudaMemcpyAsync(c->d_forcing, forcing, sizeof(float) * NFORCING * n * (size_t)nsteps, cudaMemcpyHostToDevice, st));
  CUDA_TRY(c, cudaMemcpyAsync(c->d_static_f, static_f, sizeof(float) * NSTATIC_F * n, cudaMemcpyHostToDevice, st));
  CUDA_TRY(c, cudaMemcpyAsync(c->d_static_i, static_i, sizeof(int) * n, cudaMemcpyHostToDevice, st));
  CUDA_TRY(c, cudaMemsetAsync(c->dev_budget_scratch, 0, sizeof(double) * NBUDGET, st));
  // the accumulation planes of a host call cover this call: start them at zero
  if (diag) CUDA_TRY(c, cudaMemsetAsync(c->d_diag + (size_t)D_ACCPRCP * n, 0, sizeof(float) * (NDIAG - D_ACCPRCP) * n, st));
  rc = noahmp_run(c, rp, ncol, nsteps, c->d_state, c->d_forcing, (int64_t)NFORCING * ncol, c->d_static_f, c->d_static_i,
                  diag ? c->d_diag : nullptr, c->dev_budget_scratch);
  if (rc) return rc;
  CUDA_TRY(c, cudaMemcpyAsync(state, c->d_state, sizeof(float) * nplanes * n, cudaMemcpyDeviceToHost, st));
  if (diag) CUDA_TRY(c, cudaMemcpyAsync(diag, c->d_diag, sizeof(float) * NDIAG * n, cudaMemcpyDeviceToHost, st));
  double b[NBUDGET];
  CUDA_TRY(c, cudaMemcpyAsync(b, c->dev_budget_scratch, sizeof(double) * NBUDGET, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(c, cudaStreamSynchronize(st));
  if (budget) std::memcpy(budget, b, sizeof(b));
  return (int64_t)llround(b[B_NFLAGGED]);
}

int64_t noahmp_advance_host(noahmp_ctx* c, const noahmp_run_params* rp, int64_t ncol, int nsteps, float* state_dev,
                            const float* forcing_host, const float* static_f_dev, const int32_t* static_i_dev,
                            double* budget_host, int steps_per_launch) {
  if (!c) return -1;
  if (ncol <= 0 || nsteps <= 0) return 0;
  if (!rp || !state_dev || !forcing_host || !static_f_dev || !static_i_dev) return fail(c, -1, "null argument");
  CUDA_TRY(c, cudaSetDevice(c->device));
  const int g = steps_per_launch > 0 ? (steps_per_launch < nsteps ? steps_per_launch : nsteps) : 1;
  const size_t step_words = (size_t)NFORCING * (size_t)ncol;
  if (!c->copy_stream) {
    CUDA_TRY(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int k = 0; k < 2; ++k) {
      CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_copied[k], cudaEventDisableTiming));
      CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_done[k], cudaEventDisableTiming));
    }
  }
  const size_t need = sizeof(float) * step_words * (size_t)g;
  if (need > c->cap_fbuf) {
    CUDA_TRY(c, cudaStreamSynchronize(c->copy_stream));
    for (int k = 0; k < 2; ++k) {
      if (c->fbuf[k]) cudaFree(c->fbuf[k]);
      c->fbuf[k] = nullptr;
    }
    c->cap_fbuf = 0;
    CUDA_TRY(c, cudaMalloc((void**)&c->fbuf[0], need));
    CUDA_TRY(c, cudaMalloc((void**)&c->fbuf[1], need));
    c->cap_fbuf = need;
  }
  cudaStream_t cs = c->stream;   // compute stream
  CUDA_TRY(c, cudaMemsetAsync(c->dev_budget_scratch, 0, sizeof(double) * NBUDGET, cs));
  // the copy stream must not overwrite a buffer an earlier call on the compute stream may still read
  CUDA_TRY(c, cudaEventRecord(c->ev_done[0], cs));
  CUDA_TRY(c, cudaEventRecord(c->ev_done[1], cs));
  int chunk = 0;
  for (int s0 = 0; s0 < nsteps; s0 += g, ++chunk) {
    const int n = (nsteps - s0 < g) ? nsteps - s0 : g;
    const int b = chunk & 1;
    CUDA_TRY(c, cudaStreamWaitEvent(c->copy_stream, c->ev_done[b], 0));
    CUDA_TRY(c, cudaMemcpyAsync(c->fbuf[b], forcing_host + (size_t)s0 * step_words, sizeof(float) * step_words * (size_t)n,
                                cudaMemcpyHostToDevice, c->copy_stream));
    c->h2d_bytes += (int64_t)(sizeof(float) * step_words * (size_t)n);
    CUDA_TRY(c, cudaEventRecord(c->ev_copied[b], c->copy_stream));
    CUDA_TRY(c, cudaStreamWaitEvent(cs, c->ev_copied[b], 0));
    noahmp_run_params r = *rp;
    r.julian0 = rp->julian0 + (double)s0 * (double)rp->dt / 86400.0;
    ForcingSource src;
    src.forcing = c->fbuf[b]; src.forcing_stride = (int64_t)step_words;
    const int rc = run_impl(c, &r, ncol, n, state_dev, src, static_f_dev, static_i_dev, nullptr,
                            c->dev_budget_scratch, s0 + g >= nsteps ? 1 : 0);
    if (rc) return rc;
    CUDA_TRY(c, cudaEventRecord(c->ev_done[b], cs));
  }
  double bsum[NBUDGET];
  CUDA_TRY(c, cudaMemcpyAsync(bsum, c->dev_budget_scratch, sizeof(double) * NBUDGET, cudaMemcpyDeviceToHost, cs));
  CUDA_TRY(c, cudaStreamSynchronize(cs));
  if (budget_host) std::memcpy(budget_host, bsum, sizeof(bsum));
  return (int64_t)llround(bsum[B_NFLAGGED]);
}

int64_t noahmp_advance_records_host(noahmp_ctx* c, const noahmp_run_params* rp, int64_t ncol, int nsteps, float* state_dev,
                                    const float* records_host, int nrec, const noahmp_step_forcing* steps,
                                    const float* geo_dev, const float* static_f_dev, const int32_t* static_i_dev,
                                    double* budget_host, int steps_per_launch) {
  if (!c) return -1;
  if (ncol <= 0 || nsteps <= 0) return 0;
  if (!rp || !state_dev || !records_host || !steps || !geo_dev || !static_f_dev || !static_i_dev) return fail(c, -1, "null argument");
  for (int s = 0; s < nsteps; ++s) {
    const noahmp_step_forcing& q = steps[s];
    if (q.rec0 < 0 || q.rec0 >= nrec || q.rec1 < 0 || q.rec1 >= nrec) return fail(c, -5, "record index outside records_host");
    if (!(q.weight >= 0.0f && q.weight <= 1.0f)) return fail(c, -5, "interpolation weight must be in [0,1]");
  }
  CUDA_TRY(c, cudaSetDevice(c->device));
  int g = steps_per_launch > 0 ? steps_per_launch : 1;
  if (g > nsteps) g = nsteps;
  if (g > MAX_RECORD_STEPS) g = MAX_RECORD_STEPS;
  // records a launch can touch: at most two per step; the ring holds two launches' worth so that the copies of launch
  // k+1 run while launch k computes (a record shared by consecutive launches is copied once)
  const size_t n = (size_t)ncol, G = (n + 31) >> 5;
  const size_t host_rec_words = 8 * n;            // records_host is dense
  const size_t rec_words = 8 * n + 16 * G;        // a ring slot: the eight planes + occupancy / before words of packed planes
  // Ring depth.  A launch needs at most 2 g records, so 4 g slots keep the copies one launch ahead.  A deeper ring lets them
  // run further ahead: the cost of a model step varies over the day (at night there is no radiation and no photosynthesis to
  // solve), so a run alternates between stretches bound by the link and stretches bound by the kernel, and only copies
  // that were free to run ahead during the latter make up for the former.  Up to a day of hourly records (24), within a
  // quarter of the free device memory.
  int slots = 4 * g < 4 ? 4 : 4 * g;
  {
    int deep = nrec < 24 ? nrec : 24;
    if (deep > slots && deep > c->rring_slots) {
      size_t free_b = 0, total_b = 0;
      if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) free_b = 0;
      free_b += c->cap_rring;   // the present ring is given back before the deeper one is allocated
      while (deep > slots && sizeof(float) * rec_words * (size_t)deep > free_b / 4) --deep;
    }
    if (deep > slots) slots = deep;
  }
  if (!c->copy_stream) {
    CUDA_TRY(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int k = 0; k < 2; ++k) {
      CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_copied[k], cudaEventDisableTiming));
      CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_done[k], cudaEventDisableTiming));
    }
  }
  if (!c->ev_rcopied) CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_rcopied, cudaEventDisableTiming));
  if (slots > c->rring_slots) {
    cudaEvent_t* ev = new cudaEvent_t[slots];
    for (int k = 0; k < slots; ++k) {
      if (k < c->rring_slots) ev[k] = c->ev_slot[k];
      else CUDA_TRY(c, cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming));
    }
    delete[] c->ev_slot;
    c->ev_slot = ev;
    c->rring_slots = slots;
  }
  const size_t need = sizeof(float) * rec_words * (size_t)c->rring_slots;
  if (need > c->cap_rring) {
    CUDA_TRY(c, cudaStreamSynchronize(c->copy_stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (c->rring) cudaFree(c->rring);
    c->rring = nullptr;
    c->cap_rring = 0;
    CUDA_TRY(c, cudaMalloc((void**)&c->rring, need));
    c->cap_rring = need;
  }
  const int R = c->rring_slots;
  // packing: candidate planes, the page-locked staging areas (kStageAreas, used in turn), the host threads
  int cand[8], ncand = 0;
  for (int q = 0; q < 8; ++q)
    if ((c->pack_planes >> q) & 1u) cand[ncand++] = q;
  const size_t stage_words = 2 * G + n;           // occupancy | before | values of one plane
  if (ncand) {
    for (int k = 0; k < kStageAreas; ++k)
      if (!c->ev_stage[k]) CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_stage[k], cudaEventDisableTiming));
    const size_t need_stage = sizeof(uint32_t) * stage_words * (size_t)ncand * (size_t)kStageAreas;
    if (need_stage > c->cap_stage) {
      CUDA_TRY(c, cudaStreamSynchronize(c->copy_stream));
      if (c->stage) cudaFreeHost(c->stage);
      c->stage = nullptr;
      c->cap_stage = 0;
      CUDA_TRY(c, cudaHostAlloc((void**)&c->stage, need_stage, cudaHostAllocDefault));
      c->cap_stage = need_stage;
    }
    if (c->pack_scratch.size() < 32 * G) c->pack_scratch.resize(32 * G);
    const int want = c->pack_threads > 0 ? c->pack_threads : default_pack_threads();
    if (!c->pool || c->pool->size() != want) {
      delete c->pool;
      c->pool = new HostPool(want);
    }
  }
  cudaStream_t cs = c->stream;
  CUDA_TRY(c, cudaMemsetAsync(c->dev_budget_scratch, 0, sizeof(double) * NBUDGET, cs));
  // nothing of an earlier call may still read the ring when the first copy lands
  for (int k = 0; k < R; ++k) CUDA_TRY(c, cudaEventRecord(c->ev_slot[k], cs));
  std::vector<int> resident((size_t)R, -1);
  c->slot_sparse.assign((size_t)R, 0);
  noahmp_step_forcing tab[MAX_RECORD_STEPS];
  int sp0[MAX_RECORD_STEPS], sp1[MAX_RECORD_STEPS];
  for (int s0 = 0; s0 < nsteps; s0 += g) {
    const int nl = (nsteps - s0 < g) ? nsteps - s0 : g;
    bool copied = false;
    std::vector<int> claimed((size_t)R, -1);   // record each slot holds for THIS launch
    for (int s = 0; s < nl; ++s) {
      const noahmp_step_forcing& q = steps[s0 + s];
      const int need_rec[2] = {q.rec0, q.weight == 0.0f ? q.rec0 : q.rec1};
      for (int k = 0; k < 2; ++k) {
        const int r = need_rec[k], slot = r % R;
        if (claimed[(size_t)slot] >= 0 && claimed[(size_t)slot] != r)
          return fail(c, -5, "the records of one launch must be within 4 x steps_per_launch consecutive records");
        claimed[(size_t)slot] = r;
        if (resident[(size_t)slot] == r) continue;
        const float* hrec = records_host + (size_t)r * host_rec_words;
        float* drec = c->rring + (size_t)slot * rec_words;
        // pack the candidate planes of this record (host threads) while the device works on what is already queued;
        // a staging area is free once the copies of the record it held before have completed
        int sparse = 0;
        int64_t nnz[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const int area = (int)(c->stage_next % (unsigned)kStageAreas);   // copies leave in order: a few areas in turn suffice
        if (ncand) {
          c->stage_next += 1;
          CUDA_TRY(c, cudaEventSynchronize(c->ev_stage[area]));
          const auto t_pack = std::chrono::steady_clock::now();
          for (int j = 0; j < ncand; ++j) {
            uint32_t* st = c->stage + ((size_t)area * (size_t)ncand + (size_t)j) * stage_words;
            nnz[cand[j]] = pack_plane(c->pool, hrec + (size_t)cand[j] * n, ncol, st, st + G, reinterpret_cast<float*>(st + 2 * G),
                                      c->pack_scratch.data(), kPackMaxFill);
            if (nnz[cand[j]] >= 0) sparse |= 1 << cand[j];
          }
          c->pack_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_pack).count();
        }
        CUDA_TRY(c, cudaStreamWaitEvent(c->copy_stream, c->ev_slot[slot], 0));   // last launch that read this slot
        for (int qa = 0; qa < 8;) {
          if ((sparse >> qa) & 1) {
            int j = 0;
            while (cand[j] != qa) ++j;
            const uint32_t* st = c->stage + ((size_t)area * (size_t)ncand + (size_t)j) * stage_words;
            CUDA_TRY(c, cudaMemcpyAsync(reinterpret_cast<uint32_t*>(drec + 8 * n) + (size_t)(2 * qa) * G, st, sizeof(uint32_t) * 2 * G,
                                        cudaMemcpyHostToDevice, c->copy_stream));
            c->h2d_bytes += (int64_t)(sizeof(uint32_t) * 2 * G);
            if (nnz[qa] > 0) {
              CUDA_TRY(c, cudaMemcpyAsync(drec + (size_t)qa * n, st + 2 * G, sizeof(float) * (size_t)nnz[qa], cudaMemcpyHostToDevice,
                                          c->copy_stream));
              c->h2d_bytes += (int64_t)(sizeof(float) * (size_t)nnz[qa]);
            }
            ++qa;
          } else {   // a run of dense planes travels as one copy
            int qb = qa + 1;
            while (qb < 8 && !((sparse >> qb) & 1)) ++qb;
            CUDA_TRY(c, cudaMemcpyAsync(drec + (size_t)qa * n, hrec + (size_t)qa * n, sizeof(float) * n * (size_t)(qb - qa),
                                        cudaMemcpyHostToDevice, c->copy_stream));
            c->h2d_bytes += (int64_t)(sizeof(float) * n * (size_t)(qb - qa));
            qa = qb;
          }
        }
        if (ncand) CUDA_TRY(c, cudaEventRecord(c->ev_stage[area], c->copy_stream));
        resident[(size_t)slot] = r;
        c->slot_sparse[(size_t)slot] = sparse;
        copied = true;
      }
      tab[s] = q;
      tab[s].rec0 = need_rec[0] % R;
      tab[s].rec1 = need_rec[1] % R;
      sp0[s] = c->slot_sparse[(size_t)tab[s].rec0];
      sp1[s] = c->slot_sparse[(size_t)tab[s].rec1];
    }
    if (copied) {
      CUDA_TRY(c, cudaEventRecord(c->ev_rcopied, c->copy_stream));
      CUDA_TRY(c, cudaStreamWaitEvent(cs, c->ev_rcopied, 0));
    }
    noahmp_run_params r = *rp;
    r.julian0 = rp->julian0 + (double)s0 * (double)rp->dt / 86400.0;
    ForcingSource src;
    src.records = c->rring; src.record_stride = (int64_t)rec_words; src.geo = geo_dev; src.steps = tab;
    src.sp0 = sp0; src.sp1 = sp1;
    const int rc = run_impl(c, &r, ncol, nl, state_dev, src, static_f_dev, static_i_dev, nullptr, c->dev_budget_scratch,
                            s0 + g >= nsteps ? 1 : 0);
    if (rc) return rc;
    for (int s = 0; s < nl; ++s) {   // the slots this launch reads are busy until it ends
      CUDA_TRY(c, cudaEventRecord(c->ev_slot[tab[s].rec0], cs));
      if (tab[s].rec1 != tab[s].rec0) CUDA_TRY(c, cudaEventRecord(c->ev_slot[tab[s].rec1], cs));
    }
  }
  double bsum[NBUDGET];
  CUDA_TRY(c, cudaMemcpyAsync(bsum, c->dev_budget_scratch, sizeof(double) * NBUDGET, cudaMemcpyDeviceToHost, cs));
  CUDA_TRY(c, cudaStreamSynchronize(cs));
  if (budget_host) std::memcpy(budget_host, bsum, sizeof(bsum));
  return (int64_t)llround(bsum[B_NFLAGGED]);
}

static int synth_impl(noahmp_ctx* c, int64_t ncol, int64_t col0, const int64_t* ids_dev, uint32_t seed, int step, const float* coef17,
                      const float* sc4, const float* sincos_dev, float* forcing_dev) {
  if (!c || !coef17 || !sc4 || !sincos_dev || !forcing_dev) return -1;
  if (ncol <= 0) return 0;
  CUDA_TRY(c, cudaSetDevice(c->device));
  SynthCoef p;
  std::memcpy(&p, coef17, sizeof(float) * 17);
  p.sd = sc4[0]; p.cdch = sc4[1]; p.t_diur = sc4[2]; p.t_seas = sc4[3];
  p.seed = seed; p.step = step;
  const int block = 256;
  const int64_t grid = (ncol + block - 1) / block;
  synth_forcing_kernel<<<(unsigned)grid, block, 0, c->stream>>>(ncol, col0, ids_dev, p, sincos_dev, forcing_dev);
  c->launches += 1;
  CUDA_TRY(c, cudaGetLastError());
  return 0;
}

int noahmp_synth_forcing(noahmp_ctx* c, int64_t ncol, int64_t col0, uint32_t seed, int step, const float* coef17,
                         const float* sc4, const float* sincos_dev, float* forcing_dev) {
  return synth_impl(c, ncol, col0, nullptr, seed, step, coef17, sc4, sincos_dev, forcing_dev);
}

int noahmp_synth_forcing_ids(noahmp_ctx* c, int64_t ncol, const int64_t* ids_dev, uint32_t seed, int step, const float* coef17,
                             const float* sc4, const float* sincos_dev, float* forcing_dev) {
  if (!ids_dev) return -1;
  return synth_impl(c, ncol, 0, ids_dev, seed, step, coef17, sc4, sincos_dev, forcing_dev);
}

int noahmp_prepare_forcing(noahmp_ctx* c, int64_t ncol, const float* rec0_dev, const float* rec1_dev, float weight,
                           const float* geo_dev, const float scal4[4], float* forcing_dev) {
  if (!c) return -1;
  if (ncol <= 0) return 0;
  if (!rec0_dev || !rec1_dev || !geo_dev || !scal4 || !forcing_dev) return fail(c, -1, "null argument");
  if (!(weight >= 0.0f && weight <= 1.0f)) return fail(c, -5, "interpolation weight must be in [0,1]");
  CUDA_TRY(c, cudaSetDevice(c->device));
  PrepScalars p;
  p.w = weight; p.sd = scal4[0]; p.cd = scal4[1]; p.ch0 = scal4[2]; p.sh0 = scal4[3];
  const bool vec = (ncol % 4 == 0) && (((uintptr_t)rec0_dev | (uintptr_t)rec1_dev | (uintptr_t)geo_dev | (uintptr_t)forcing_dev) % 16 == 0);
  const int block = 256;
  const int64_t work = vec ? ncol / 4 : ncol;
  int64_t grid = (work + block - 1) / block;
  const int64_t cap = 148LL * 8 * 4;   // persistent-style: a few CTAs per SM, grid-stride loop
  if (grid > cap) grid = cap;
  if (vec) prepare_forcing_kernel<4><<<(unsigned)grid, block, 0, c->stream>>>(ncol, rec0_dev, rec1_dev, geo_dev, p, forcing_dev);
  else prepare_forcing_kernel<1><<<(unsigned)grid, block, 0, c->stream>>>(ncol, rec0_dev, rec1_dev, geo_dev, p, forcing_dev);
  c->launches += 1;
  CUDA_TRY(c, cudaGetLastError());
  return 0;
}

int noahmp_layout(int what) {
  switch (what) {
    case 0: return NSTATE;
    case 1: return NSTATE_CORE;
    case 2: return NFORCING;
    case 3: return NSTATIC_F;
    case 4: return NDIAG;
    case 5: return NGEN;
    case 6: return NSOILP;
    case 7: return NCOLORP;
    case 8: return NVEGF;
    case 9: return NVEGI;
    case 10: return (int)sizeof(noahmp_run_params);
    case 11: return NBUDGET;
  }
  return -1;
}
const char* noahmp_state_plane_name(int i) { return (i >= 0 && i < NSTATE) ? kStateNames[i] : nullptr; }
const char* noahmp_forcing_plane_name(int i) { return (i >= 0 && i < NFORCING) ? kForcingNames[i] : nullptr; }
const char* noahmp_diag_plane_name(int i) { return (i >= 0 && i < NDIAG) ? kDiagNames[i] : nullptr; }
const char* noahmp_budget_name(int i) { return (i >= 0 && i < NBUDGET) ? kBudgetNames[i] : nullptr; }
int64_t noahmp_launch_count(const noahmp_ctx* c) { return c ? c->launches : -1; }
int64_t noahmp_h2d_bytes(const noahmp_ctx* c) { return c ? c->h2d_bytes : -1; }
double noahmp_pack_seconds(const noahmp_ctx* c) { return c ? c->pack_seconds : -1.0; }

int noahmp_set_record_packing(noahmp_ctx* c, unsigned plane_mask, int host_threads) {
  if (!c) return -1;
  if (plane_mask > 0xffu) return fail(c, -5, "a record has eight planes: plane_mask must be below 256");
  if (host_threads > 256) return fail(c, -5, "host_threads above 256");
  c->pack_planes = plane_mask;
  c->pack_threads = host_threads > 0 ? host_threads : 0;
  return 0;
}

int64_t noahmp_pack_plane(const float* plane, int64_t ncol, uint32_t* occupancy, uint32_t* before, float* values,
                          int host_threads) {
  if (ncol < 0 || (ncol > 0 && (!plane || !occupancy || !before || !values))) return -1;
  if (ncol == 0) return 0;
  std::unique_ptr<uint32_t[]> scratch(new uint32_t[(size_t)(((ncol + 31) >> 5) * 32)]);   // not initialised on purpose
  if (host_threads > 1) {
    HostPool pool(host_threads > 256 ? 256 : host_threads);
    return pack_plane(&pool, plane, ncol, occupancy, before, values, scratch.get(), 1.0);
  }
  return pack_plane(nullptr, plane, ncol, occupancy, before, values, scratch.get(), 1.0);
}

int noahmp_kernel_info(noahmp_ctx* c, int info[4]) {
  if (!c || !info) return -1;
  CUDA_TRY(c, cudaSetDevice(c->device));
  cudaFuncAttributes at;
  StepKernel k = pick_kernel(c->opt_is_default, false, false);
  CUDA_TRY(c, cudaFuncGetAttributes(&at, (const void*)k));
  info[0] = at.numRegs; info[1] = (int)at.localSizeBytes; info[2] = (int)at.sharedSizeBytes; info[3] = at.maxThreadsPerBlock;
  return 0;
}

// ---- bounds-checked debug build (-DNMP_BOUNDS, common.cuh) -----------------------------------------------------------
#ifdef NMP_BOUNDS
namespace {
__global__ void bounds_selftest_kernel(float* out, int idx) {
  const float a[4] = {1.0f, 2.0f, 3.0f, 4.0f};
  out[0] = nmp::pick(a, idx);   // idx outside 0..3: counted at site 0 and clamped
}
}  // namespace
#endif

int noahmp_debug_bounds(noahmp_ctx* c, unsigned long long* hits, unsigned int* sites, int reset) {
  if (!c) return -1;
  if (hits) *hits = 0;
  if (sites) *sites = 0;
#ifdef NMP_BOUNDS
  CUDA_TRY(c, cudaSetDevice(c->device));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  unsigned long long h = 0;
  unsigned int m = 0;
  CUDA_TRY(c, cudaMemcpyFromSymbol(&h, nmp::g_bounds_hits, sizeof(h)));
  CUDA_TRY(c, cudaMemcpyFromSymbol(&m, nmp::g_bounds_sites, sizeof(m)));
  if (hits) *hits = h;
  if (sites) *sites = m;
  if (reset) {
    h = 0; m = 0;
    CUDA_TRY(c, cudaMemcpyToSymbol(nmp::g_bounds_hits, &h, sizeof(h)));
    CUDA_TRY(c, cudaMemcpyToSymbol(nmp::g_bounds_sites, &m, sizeof(m)));
  }
  return 1;
#else
  (void)reset;
  return 0;
#endif
}

int noahmp_debug_bounds_selftest(noahmp_ctx* c, int idx) {
  if (!c) return -1;
#ifdef NMP_BOUNDS
  CUDA_TRY(c, cudaSetDevice(c->device));
  float* d = nullptr;
  CUDA_TRY(c, cudaMalloc(&d, sizeof(float)));
  bounds_selftest_kernel<<<1, 1, 0, c->stream>>>(d, idx);
  CUDA_TRY(c, cudaGetLastError());
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  CUDA_TRY(c, cudaFree(d));
  return 1;
#else
  (void)idx;
  return 0;
#endif
}

}  // extern "C"
