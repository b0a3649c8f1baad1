// cmg_capi.cu -- host side of the C ABI declared in include/casm_monte_gpu.h.
// Owns device memory, builds the per-chain dE / acceptance tables with the
// reference's expression order, and launches the sm_100a kernels of
// cmg_device.cuh.  No CPU fallback: every entry point needs a CUDA device.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/casm_monte_gpu.h"
#include "../../include/casm_monte_b200/snf.hh"
#include "cmg_device.cuh"

using namespace cmg;

// ---------------------------------------------------------------------------
struct GraphKey {
  int variant;
  long long passes;
  long long sample_period;
  bool operator<(GraphKey const &o) const {
    if (variant != o.variant) return variant < o.variant;
    if (passes != o.passes) return passes < o.passes;
    return sample_period < o.sample_period;
  }
};

struct RunState {
  unsigned long long pass;
  long long n_samples;
};

enum Variant { V_AUTO = 0, V_GENERIC = 1, V_BULK2D = 2, V_BULK3D = 3, V_TILE2D = 4, V_RING2D = 5, V_TMA3D = 6 };

struct SeriesCheckBlock;

struct cmg_context {
  int device = 0;
  cudaStream_t stream = 0;
  int dim = 2;
  long long shape[3] = {1, 1, 1};
  long long n_sites = 0;
  int n_chains = 1;
  bool planar = false;  // all extents even -> colour planes
  // slab decomposition
  bool slab = false;
  long long col_begin = 0;
  long long global_shape[3] = {1, 1, 1};
  uint8_t *d_halo[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // [colour][side]
  uint8_t *push[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};    // [colour][side]
  unsigned long long *d_flags = nullptr;       // [2] wait flags, raised by the neighbours
  unsigned long long *peer_flag[2] = {nullptr, nullptr};  // the neighbours' wait flags
  unsigned int *d_done = nullptr;
  unsigned long long slab_epoch = 0;  // fused half-sweeps stepped since the peers were attached
  bool slab_exchange = true;          // false: halos neither pushed nor waited for (timing aid)
  std::vector<void *> ipc_opened;

  uint8_t *d_planes = nullptr;
  uint8_t *d_planes_alt = nullptr;  // second copy, written by k_tile2d launches with halos (then swapped with d_planes)
  long long plane_stride = 0, chain_stride = 0;
  uint8_t *d_nat = nullptr;  // [n_chains][n_sites]
  bool nat_is_current = false;     // natural copy mirrors the planes
  int32_t *d_stage = nullptr;      // [n_sites] int32 staging
  int *d_flag = nullptr;
  unsigned char *d_event = nullptr;  // event staging for cmg_event_delta

  ChainTables *d_tabs = nullptr;
  std::vector<ChainTables> h_tabs;
  double J = 1.0;
  int lattice_type = 1;
  bool model_set = false;

  unsigned long long philox_seed = 0;
  int chain_offset = 0;
  unsigned long long h_pass = 0;  // global pass index (Philox counter)
  RunState *d_run = nullptr;
  long long n_pass = 0;  // passes since reset_counters
  unsigned long long *d_n_accept = nullptr;

  MT64State *d_engines = nullptr;
  std::vector<char> engine_seeded;
  long long *d_cur_sb = nullptr;  // [n_chains][2]
  long long *d_scratch_sb = nullptr;  // [2]

  // sample series, sample-major: slot s at d_series + s*n_chains*2
  long long *d_series = nullptr;
  long long capacity = 0;
  long long n_samples = 0;
  double *d_dbl = nullptr;  // [n_chains][3][dbl_capacity]
  long long dbl_capacity = 0;
  long long dbl_valid = 0;

  // use_nlist = false energy form (2-d, checkerboard): per-sample line counts
  bool nonlist = false;
  int *d_lines = nullptr;  // [sample][chain][n0 + n1]
  long long lines_capacity = 0;
  int forced_variant = V_AUTO;
  int js = 0;  // 0 = auto
  int tile_passes = 3;   // passes per launch of the tiled kernel (halo = 2*P columns)
  int tile_threads = 512;
  bool tile_threads_forced = false;  // ":nt=" given
  int ring_passes = 128;  // passes per cooperative launch of the ring kernel
  uint8_t *d_ring_mailbox = nullptr;  // [n_chains][n_tiles][side][plane][h]
  size_t ring_mailbox_bytes = 0;
  // slab ring (k_ring2d continuing on the neighbour GPUs): the neighbours' mailboxes
  uint8_t *ring_peer_mb[2] = {nullptr, nullptr};
  int ring_peer_tiles[2] = {0, 0};
  int ring_tiles_cap = 0;              // "ring2d:rt=<n>": at most n tiles (several rings on one GPU)
  bool pdl = true;                     // half-sweep kernels launched as programmatic dependents (":pdl=0" turns it off)
  // chained half-sweeps (bulk2d / bulk3d): per-CTA completion flags, see hs_wait_for
  bool hs_chain = true;                // ":chain=0": every half-sweep waits for the whole previous grid
  unsigned int *d_hs_flags = nullptr;
  size_t hs_flags_count = 0;
  uint32_t hs_epoch = 1;
  bool hs_prev_chained = false;        // the kernel before this one on the stream is a flagged half-sweep of the same geometry
  unsigned long long ring_s0 = 0;      // half-sweeps the ring has stepped since the peers were attached
  // sticky device error word (kErr* bits, raised with atomicOr by kernels whose
  // waits are bounded) and its pinned host copy; zeroed at create and after a
  // failure has been reported
  unsigned int *d_error = nullptr;
  unsigned int *h_error = nullptr;
  int coop_launch = 0;
  int sm_count = 148;
  bool bulk_attr_set = false;
  bool tma3d_attr_set = false;
  int n_strips2d = 0;  // bulk2d: 0 = automatic, 1 = balanced strips whenever two to four waves of them exist, > 1 = that many
  int js_auto[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // cached strip length per kernel variant
  size_t smem_optin = 0;
  // cmg_mark / cmg_rollback: a restore point (planes, acceptance counters, host counters) and
  // a second, higher-priority stream on which the statistics of the samples taken up to
  // the mark are evaluated while the main stream already sweeps on
  cudaStream_t aux = nullptr;
  cudaEvent_t ev_mark = nullptr;
  uint8_t *d_shadow = nullptr;
  unsigned long long *d_shadow_accept = nullptr;
  bool mark_valid = false;
  int philox_rounds = 10;  // 10 (published default) or 7 (cmg_set_philox_rounds)
  bool reserve_sm = false;  // checks have run on the aux stream next to the sweep: k_ring2d leaves them an SM
  // a completion check enqueued ahead of its decision (cmg_series_check_prefetch)
  struct SeriesCheckBlock *d_check = nullptr, *h_check_in = nullptr, *h_check_out = nullptr;
  cudaEvent_t ev_check = nullptr;
  bool check_pending = false;
  int check_key_n = 0, check_key_chain = 0, check_key_q[3] = {0, 0, 0};
  double check_key_abs[3] = {0, 0, 0}, check_key_conf = 0;
  long long check_key_count = 0;
  unsigned long long mark_h_pass = 0;
  long long mark_n_pass = 0, mark_n_samples = 0;
  // k-state model; K == 0: the context runs the Ising path
  int ks_K = 0;
  double ks_V[kMaxSpecies * kMaxSpecies] = {0};
  KStateTables *d_ktabs = nullptr;  // [n_chains]
  std::vector<char> ks_valid;
  long long *d_kseries = nullptr;  // [sample][chain][K + K*K]
  long long ks_capacity = 0, ks_n_samples = 0;
  int *d_kloc = nullptr, *d_kloc_size = nullptr, *d_kmol_loc = nullptr;  // OccLocation of every chain
  bool ks_loc_valid = false;
  // N-fold way driver: class lists of every chain, weights of its samples, clock
  int *d_nf_members = nullptr, *d_nf_n = nullptr, *d_nf_pos = nullptr;
  uint8_t *d_nf_class = nullptr;
  double *d_nf_weight = nullptr, *d_nf_ratio = nullptr, *d_nf_time = nullptr;  // [sample][chain] x2, [chain]
  long long nf_capacity = 0, nf_first_sample = -1, nf_steps = 0;
  bool nf_lists_valid = false;
  long long launches = 0;
  std::string last_error;
  std::string variant_name = "auto";
};

static thread_local std::string g_last_error;

static int fail(cmg_context *ctx, int code, const std::string &msg) {
  if (ctx) ctx->last_error = msg;
  g_last_error = msg;
  return code;
}

#define CU(ctx, call)                                                          \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) {                                                  \
      char buf__[512];                                                         \
      snprintf(buf__, sizeof buf__, "%s failed: %s (%s:%d)", #call,            \
               cudaGetErrorString(e__), __FILE__, __LINE__);                   \
      return fail(ctx, (e__ == cudaErrorMemoryAllocation) ? CMG_ENOMEM : CMG_ECUDA, \
                  buf__);                                                      \
    }                                                                          \
  } while (0)

#define NEED(ctx)                                                 \
  do {                                                            \
    if (!(ctx)) return fail(nullptr, CMG_EINVAL, "null context"); \
    cudaSetDevice((ctx)->device);                                 \
  } while (0)

static inline unsigned int nblocks(long long n, int bs) {
  return (unsigned int)((n + bs - 1) / bs);
}

// ---------------------------------------------------------------------------
// Tables.  Expression order follows the reference exactly:
//   dE_f = (-J * (new_occ - s)) * sum_nbr      model.hh:312-314 (left to right)
//   Ndx  = 0.0 + (new_occ - s) / 2.0           model.hh:430-433
//   dE   = dE_f - mu * Ndx                     basic_semigrand_canonical.hh:185-191
//   prob = exp(-dE * beta), beta = 1/(KB*T)    metropolis.hh:33,
//                                              basic_occupation_metropolis.hh:361
// This translation unit is compiled with -fmad=false for device code and the
// host compiler's default (no contraction across statements on x86-64 without
// -march flags), and the statements are kept separate.
// ---------------------------------------------------------------------------
static void build_tables(ChainTables &t, int dim, double J, double T, double mu) {
  memset(&t, 0, sizeof t);
  t.J = J;
  t.mu = mu;
  t.temperature = T;
  volatile double kt = CMG_KB * T;
  t.beta = 1.0 / kt;
  const int z = 2 * dim;
  for (int b = 0; b < 2; ++b) {
    const int s = b ? 1 : -1;
    const int new_occ = -s;
    for (int nu = 0; nu <= z; ++nu) {
      const int nb_sum = 2 * nu - z;
      volatile double a = -J * (new_occ - s);
      volatile double dE_f = a * nb_sum;
      volatile double Ndx = 0.0;
      Ndx = Ndx + (new_occ - s) / 2.0;
      volatile double m = mu * Ndx;
      const double dE = dE_f - m;
      volatile double arg = -dE * t.beta;
      const double p = std::exp(arg);
      const int idx = 2 * nu + b;
      t.dE[idx] = dE;
      t.prob[idx] = p;
      uint32_t thr;
      if (dE < 0.0 || p >= 1.0) {
        thr = 0xFFFFFFFFu;
      } else if (!(p > 0.0)) {
        // exp underflowed to 0: the reference's `rand < prob` is never true
        thr = 0u;
        t.never_mask |= 1u << idx;
      } else {
        double scaled = std::ceil(p * 4294967296.0);
        if (scaled < 1.0) scaled = 1.0;
        if (scaled > 4294967296.0) scaled = 4294967296.0;
        thr = (uint32_t)((unsigned long long)scaled - 1ull);
      }
      t.thr_m1[idx] = thr;
    }
  }
  t.valid = 1;
}

// Winitzki's erf^-1 approximation as used for the confidence factor
// (include/casm/monte/misc/math.hh:62-72; BasicStatistics.cc:127)
static double z_confidence(double confidence) {
  const double a = 0.147, PI = 3.141592653589793238463;
  const double sgn = (confidence < 0.0) ? -1.0 : 1.0;
  const double b = std::log((1.0 - confidence) * (1.0 + confidence));
  const double c = 2.0 / (PI * a) + b * 0.5;
  const double d = b / a;
  return std::sqrt(2.0) * (sgn * std::sqrt(std::sqrt(c * c - d) - c));
}

static NaturalShape nat_shape(const cmg_context *c) {
  NaturalShape s;
  s.n0 = (int)c->shape[0];
  s.n1 = (int)c->shape[1];
  s.n2 = (int)c->shape[2];
  s.dim = c->dim;
  s.n_sites = c->n_sites;
  return s;
}

static LatticeView view(const cmg_context *c) {
  LatticeView L;
  memset(&L, 0, sizeof L);
  L.planes = c->d_planes;
  L.plane_stride = c->plane_stride;
  L.chain_stride = c->chain_stride;
  L.h = (int)(c->shape[0] / 2);
  L.n1 = (int)c->shape[1];
  L.n2 = (int)c->shape[2];
  L.dim = c->dim;
  L.col_offset = c->col_begin;
  L.error = c->d_error;
  if (c->slab) {
    // every CTA of k_halfsweep_bulk2d inside one strip: the boundary columns go first
    L.edge_mode = (c->shape[0] / 32) % 128 == 0 ? 1 : 0;
    for (int col = 0; col < 2; ++col) {
      L.halo_lo[col] = c->d_halo[col][0];
      L.halo_hi[col] = c->d_halo[col][1];
      if (c->slab_exchange) {
        L.push_lo[col] = c->push[col][0];
        L.push_hi[col] = c->push[col][1];
      }
    }
    if (c->slab_exchange) {
      for (int side = 0; side < 2; ++side) {
        // only meaningful once a peer is attached on that side
        L.wait_flag[side] = c->peer_flag[side] ? c->d_flags + side : nullptr;
        L.signal_flag[side] = c->peer_flag[side];
      }
      L.done_counter = (c->peer_flag[0] || c->peer_flag[1]) ? c->d_done : nullptr;
    }
  }
  return L;
}

// ---------------------------------------------------------------------------
extern "C" {

int cmg_abi_version(void) { return CMG_ABI_VERSION; }
const char *cmg_last_global_error(void) { return g_last_error.c_str(); }
const char *cmg_last_error(const cmg_context *ctx) {
  return ctx ? ctx->last_error.c_str() : g_last_error.c_str();
}

int cmg_device_count(int *count) {
  if (!count) return fail(nullptr, CMG_EINVAL, "count == null");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    *count = 0;
    cudaGetLastError();
    return fail(nullptr, CMG_ENODEVICE,
                std::string("no CUDA device: ") + cudaGetErrorString(e));
  }
  *count = n;
  return CMG_OK;
}

static int create_common(int dim, const int64_t *shape, int n_chains, int device,
                         bool slab, long long col_begin, const int64_t *gshape,
                         cmg_context **out) {
  if (!out) return fail(nullptr, CMG_EINVAL, "out == null");
  *out = nullptr;
  if (dim != 2 && dim != 3)
    return fail(nullptr, CMG_EINVAL, "IsingConfiguration only supports 2d (3d is the extension)");
  if (!shape) return fail(nullptr, CMG_EINVAL, "shape == null");
  if (n_chains < 1 || n_chains >= (1 << 24))
    return fail(nullptr, CMG_EINVAL, "n_chains must be in [1, 2^24)");
  for (int d = 0; d < dim; ++d)
    if (shape[d] < 2 || shape[d] > 0x7fffffffLL)
      return fail(nullptr, CMG_EINVAL, "extent out of range (need 2 <= n < 2^31)");
  int ndev = 0;
  int rc = cmg_device_count(&ndev);
  if (rc != CMG_OK) return rc;
  if (ndev == 0) return fail(nullptr, CMG_ENODEVICE, "no CUDA device");
  if (device < 0 || device >= ndev) return fail(nullptr, CMG_EINVAL, "bad device index");

  cmg_context *c = new cmg_context();
  c->device = device;
  c->dim = dim;
  c->n_chains = n_chains;
  c->n_sites = 1;
  for (int d = 0; d < 3; ++d) {
    c->shape[d] = d < dim ? shape[d] : 1;
    c->global_shape[d] = gshape ? (d < dim ? gshape[d] : 1) : c->shape[d];
    c->n_sites *= c->shape[d];
  }
  c->slab = slab;
  c->col_begin = col_begin;
  // colour planes need even extents; a slab needs only n0 (and the GLOBAL n1) even
  c->planar = (c->shape[0] % 2 == 0) && (slab || c->shape[1] % 2 == 0) &&
              (dim == 2 || c->shape[2] % 2 == 0);
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    delete c;
    return fail(nullptr, CMG_ECUDA, cudaGetErrorString(e));
  }
  {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && v > 0)
      c->sm_count = v;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) == cudaSuccess)
      c->smem_optin = (size_t)v;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, device) == cudaSuccess)
      c->coop_launch = v;
  }
#define CUC(call)                                                        \
  do {                                                                   \
    cudaError_t e__ = (call);                                            \
    if (e__ != cudaSuccess) {                                            \
      std::string m__ = std::string(#call) + ": " + cudaGetErrorString(e__); \
      cmg_destroy(c);                                                    \
      return fail(nullptr, e__ == cudaErrorMemoryAllocation ? CMG_ENOMEM : CMG_ECUDA, m__); \
    }                                                                    \
  } while (0)
  if (c->planar) {
    long long plane_bytes = c->n_sites / 2;
    c->plane_stride = (plane_bytes + 255) / 256 * 256;
    c->chain_stride = 2 * c->plane_stride;
    CUC(cudaMalloc(&c->d_planes, (size_t)c->chain_stride * n_chains));
    CUC(cudaMemset(c->d_planes, 1, (size_t)c->chain_stride * n_chains));
  } else {
    CUC(cudaMalloc(&c->d_nat, (size_t)c->n_sites * n_chains));
    CUC(cudaMemset(c->d_nat, 1, (size_t)c->n_sites * n_chains));
    c->nat_is_current = true;
  }
  CUC(cudaMalloc(&c->d_tabs, sizeof(ChainTables) * n_chains));
  CUC(cudaMemset(c->d_tabs, 0, sizeof(ChainTables) * n_chains));
  c->h_tabs.resize(n_chains);
  for (auto &t : c->h_tabs) memset(&t, 0, sizeof t);
  CUC(cudaMalloc(&c->d_n_accept, sizeof(unsigned long long) * n_chains));
  CUC(cudaMemset(c->d_n_accept, 0, sizeof(unsigned long long) * n_chains));
  CUC(cudaMalloc(&c->d_run, sizeof(RunState)));
  CUC(cudaMemset(c->d_run, 0, sizeof(RunState)));
  CUC(cudaMalloc(&c->d_flag, sizeof(int)));
  CUC(cudaMemset(c->d_flag, 0, sizeof(int)));
  CUC(cudaMalloc(&c->d_cur_sb, sizeof(long long) * 2 * n_chains));
  CUC(cudaMalloc(&c->d_scratch_sb, sizeof(long long) * 2));
  CUC(cudaMalloc(&c->d_error, sizeof(unsigned int)));
  CUC(cudaMemset(c->d_error, 0, sizeof(unsigned int)));
  CUC(cudaMallocHost(&c->h_error, sizeof(unsigned int)));
  *c->h_error = 0;
  c->engine_seeded.assign(n_chains, 0);
  if (slab) {
    const long long hb = c->shape[0] / 2;
    for (int col = 0; col < 2; ++col)
      for (int side = 0; side < 2; ++side) {
        CUC(cudaMalloc(&c->d_halo[col][side], (size_t)hb));
        CUC(cudaMemset(c->d_halo[col][side], 1, (size_t)hb));
      }
    CUC(cudaMalloc(&c->d_flags, 2 * sizeof(unsigned long long)));
    CUC(cudaMemset(c->d_flags, 0, 2 * sizeof(unsigned long long)));
    CUC(cudaMalloc(&c->d_done, 3 * sizeof(unsigned int)));
    CUC(cudaMemset(c->d_done, 0, 3 * sizeof(unsigned int)));
  }
#undef CUC
  *out = c;
  return CMG_OK;
}

int cmg_create(int dim, const int64_t *shape, int n_chains, int device,
               cmg_context **out) {
  return create_common(dim, shape, n_chains, device, false, 0, nullptr, out);
}

int cmg_create_slab(int dim, const int64_t *global_shape, int64_t col_begin,
                    int64_t n_cols_local, int device, cmg_context **out) {
  if (dim != 2) return fail(nullptr, CMG_EUNSUPPORTED, "slab decomposition is 2-d only");
  if (!global_shape) return fail(nullptr, CMG_EINVAL, "global_shape == null");
  if (global_shape[0] % 32 != 0)
    return fail(nullptr, CMG_EINVAL, "slab decomposition needs n0 % 32 == 0");
  if (global_shape[1] % 2 != 0)
    return fail(nullptr, CMG_EINVAL, "checkerboard needs even extents");
  if (col_begin < 0 || n_cols_local < 1 || col_begin + n_cols_local > global_shape[1])
    return fail(nullptr, CMG_EINVAL, "slab column range outside the lattice");
  if (col_begin % 2 != 0)
    return fail(nullptr, CMG_EINVAL, "slab col_begin must be even (local colour == global colour)");
  int64_t local[2] = {global_shape[0], n_cols_local};
  return create_common(2, local, 1, device, true, col_begin, global_shape, out);
}

int cmg_destroy(cmg_context *c) {
  if (!c) return CMG_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (void *p : c->ipc_opened) cudaIpcCloseMemHandle(p);
  cudaFree(c->d_planes);
  cudaFree(c->d_planes_alt);
  cudaFree(c->d_nat);
  cudaFree(c->d_stage);
  cudaFree(c->d_flag);
  cudaFree(c->d_event);
  cudaFree(c->d_tabs);
  cudaFree(c->d_run);
  cudaFree(c->d_n_accept);
  cudaFree(c->d_engines);
  cudaFree(c->d_cur_sb);
  cudaFree(c->d_scratch_sb);
  cudaFree(c->d_series);
  cudaFree(c->d_dbl);
  for (int col = 0; col < 2; ++col)
    for (int side = 0; side < 2; ++side) cudaFree(c->d_halo[col][side]);
  cudaFree(c->d_flags);
  cudaFree(c->d_done);
  cudaFree(c->d_ring_mailbox);
  cudaFree(c->d_hs_flags);
  cudaFree(c->d_check);
  if (c->h_check_in) cudaFreeHost(c->h_check_in);
  if (c->h_check_out) cudaFreeHost(c->h_check_out);
  if (c->ev_check) cudaEventDestroy(c->ev_check);
  cudaFree(c->d_lines);
  cudaFree(c->d_error);
  cudaFree(c->d_shadow);
  cudaFree(c->d_shadow_accept);
  if (c->aux) cudaStreamDestroy(c->aux);
  if (c->ev_mark) cudaEventDestroy(c->ev_mark);
  cudaFree(c->d_nf_members);
  cudaFree(c->d_nf_n);
  cudaFree(c->d_nf_pos);
  cudaFree(c->d_nf_class);
  cudaFree(c->d_nf_weight);
  cudaFree(c->d_nf_ratio);
  cudaFree(c->d_nf_time);
  cudaFree(c->d_ktabs);
  cudaFree(c->d_kseries);
  cudaFree(c->d_kloc);
  cudaFree(c->d_kloc_size);
  cudaFree(c->d_kmol_loc);
  if (c->h_error) cudaFreeHost(c->h_error);
  delete c;
  return CMG_OK;
}

int cmg_set_stream(cmg_context *c, void *cuda_stream) {
  NEED(c);
  CU(c, cudaStreamSynchronize(c->stream));
  c->stream = (cudaStream_t)cuda_stream;
  return CMG_OK;
}

// Kernels with waits (k_ring2d: neighbour edges and bulk-copy completion; slab
// half-sweeps: the neighbours' flags) bound them and raise bits of the context's
// STICKY device error word with atomicOr; no launch ever clears it.  Every entry
// point that synchronises the stream fetches it here, so a timeout in any launch
// of a multi-launch call surfaces at the next synchronising call -- and then the
// word is cleared, the failure having been reported (results up to here are
// invalid; re-upload the occupation to go on).
static int device_error_check(cmg_context *c) {
  if (!c->d_error) return CMG_OK;
  CU(c, cudaMemcpyAsync(c->h_error, c->d_error, sizeof(unsigned int), cudaMemcpyDeviceToHost,
                        c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  const unsigned int e = *c->h_error;
  if (!e) return CMG_OK;
  *c->h_error = 0;
  cudaMemsetAsync(c->d_error, 0, sizeof(unsigned int), c->stream);
  std::string msg;
  if (e & (kErrRingEdge | kErrRingCopy))
    msg += "ring2d / tile2d: a tile or warp waited too long for its neighbour or its bulk copy; ";
  if (e & kErrChain)
    msg += "chained half-sweeps: a CTA waited too long for its neighbours of the previous half-sweep; ";
  if (e & kErrSlabWait)
    msg += "slab half-sweep: a neighbour's flag did not arrive within 20 s (dead rank or "
           "half-sweep sequences that differ between ranks); ";
  return fail(c, CMG_ECUDA, msg + "results since the last successful synchronisation are invalid");
}

int cmg_sync(cmg_context *c) {
  NEED(c);
  CU(c, cudaStreamSynchronize(c->stream));
  return device_error_check(c);
}

int cmg_n_sites(const cmg_context *c, int64_t *n) {
  if (!c || !n) return fail(nullptr, CMG_EINVAL, "null argument");
  *n = c->n_sites;
  return CMG_OK;
}

// ---- model / conditions ------------------------------------------------------
// Samples are kept as integer sums and turned into doubles lazily with the
// chain's (J, mu): before either changes, the samples taken so far are converted
// with the values they were taken under (a temperature / mu sweep that re-uses a
// context keeps a consistent series).
static int ensure_doubles(cmg_context *c);

int cmg_set_model(cmg_context *c, double J, int lattice_type) {
  NEED(c);
  if (lattice_type != 1) return fail(c, CMG_EINVAL, "Unsupported lattice_type");
  {
    const int rc = ensure_doubles(c);
    if (rc) return rc;
  }
  c->J = J;
  c->lattice_type = lattice_type;
  c->model_set = true;
  // conditions already set keep their (T, mu) but need new tables
  for (int ch = 0; ch < c->n_chains; ++ch)
    if (c->h_tabs[ch].valid)
      build_tables(c->h_tabs[ch], c->dim, J, c->h_tabs[ch].temperature, c->h_tabs[ch].mu);
  CU(c, cudaMemcpyAsync(c->d_tabs, c->h_tabs.data(), sizeof(ChainTables) * c->n_chains,
                        cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return CMG_OK;
}

int cmg_set_conditions(cmg_context *c, int chain, double temperature, double mu) {
  NEED(c);
  if (chain < -1 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  if (!(temperature > 0.0)) return fail(c, CMG_EINVAL, "temperature must be > 0");
  {
    const int rc = ensure_doubles(c);
    if (rc) return rc;
  }
  const int lo = chain < 0 ? 0 : chain, hi = chain < 0 ? c->n_chains : chain + 1;
  for (int ch = lo; ch < hi; ++ch) build_tables(c->h_tabs[ch], c->dim, c->J, temperature, mu);
  CU(c, cudaMemcpyAsync(c->d_tabs + lo, c->h_tabs.data() + lo, sizeof(ChainTables) * (hi - lo),
                        cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return CMG_OK;
}

int cmg_get_tables(cmg_context *c, int chain, double *dE, double *prob, uint32_t *thr) {
  NEED(c);
  if (chain < 0 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  if (!c->h_tabs[chain].valid) return fail(c, CMG_ESTATE, "conditions not set");
  const int n = 2 * (2 * c->dim + 1);
  for (int i = 0; i < n; ++i) {
    if (dE) dE[i] = c->h_tabs[chain].dE[i];
    if (prob) prob[i] = c->h_tabs[chain].prob[i];
    if (thr) thr[i] = c->h_tabs[chain].thr_m1[i];
  }
  return CMG_OK;
}

// ---- occupation ----------------------------------------------------------------
static int ensure_stage(cmg_context *c) {
  if (!c->d_stage) CU(c, cudaMalloc(&c->d_stage, sizeof(int32_t) * (size_t)c->n_sites));
  return CMG_OK;
}
static int ensure_nat(cmg_context *c) {
  if (!c->d_nat) {
    CU(c, cudaMalloc(&c->d_nat, (size_t)c->n_sites * c->n_chains));
    c->nat_is_current = false;
  }
  return CMG_OK;
}
static NaturalShape slab_aware_shape(const cmg_context *c) { return nat_shape(c); }

// planes -> natural for all chains (no-op when the lattice is natural-only)
static int sync_nat_from_planes(cmg_context *c) {
  if (!c->planar) return CMG_OK;
  int rc = ensure_nat(c);
  if (rc) return rc;
  if (c->nat_is_current) return CMG_OK;
  NaturalShape s = slab_aware_shape(c);
  for (int ch = 0; ch < c->n_chains; ++ch) {
    k_planes_to_natural<<<nblocks(c->n_sites, 256), 256, 0, c->stream>>>(
        c->d_planes + ch * c->chain_stride, c->plane_stride,
        c->d_nat + ch * c->n_sites, s);
    ++c->launches;
  }
  CU(c, cudaGetLastError());
  c->nat_is_current = true;
  return CMG_OK;
}
static int sync_planes_from_nat(cmg_context *c) {
  if (!c->planar) return CMG_OK;
  NaturalShape s = slab_aware_shape(c);
  for (int ch = 0; ch < c->n_chains; ++ch) {
    k_natural_to_planes<<<nblocks(c->n_sites, 256), 256, 0, c->stream>>>(
        c->d_nat + ch * c->n_sites, c->d_planes + ch * c->chain_stride, c->plane_stride, s);
    ++c->launches;
  }
  CU(c, cudaGetLastError());
  return CMG_OK;
}

static int upload_from_stage(cmg_context *c, int chain, const int32_t *src_dev) {
  CU(c, cudaMemsetAsync(c->d_flag, 0, sizeof(int), c->stream));
  NaturalShape s = nat_shape(c);
  if (c->planar) {
    k_i32_to_planes<<<nblocks(c->n_sites, 256), 256, 0, c->stream>>>(
        src_dev, c->d_planes + chain * c->chain_stride, c->plane_stride, s, c->d_flag);
    c->nat_is_current = false;
  } else {
    k_i32_to_natural<<<nblocks(c->n_sites, 256), 256, 0, c->stream>>>(
        src_dev, c->d_nat + chain * c->n_sites, c->n_sites, c->d_flag);
  }
  ++c->launches;
  CU(c, cudaGetLastError());
  int bad = 0;
  CU(c, cudaMemcpyAsync(&bad, c->d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  if (bad) return fail(c, CMG_EINVAL, "occupation values must be +1 or -1");
  return CMG_OK;
}

int cmg_upload_occupation_i32(cmg_context *c, int chain, const int32_t *occ, int64_t n) {
  NEED(c);
  if (chain < 0 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  if (!occ || n != c->n_sites)
    return fail(c, CMG_EINVAL, "Error in set_occupation: size mismatch");
  if (c->slab && (c->col_begin & 1))
    return fail(c, CMG_EUNSUPPORTED, "slab col_begin must be even");
  int rc = ensure_stage(c);
  if (rc) return rc;
  CU(c, cudaMemcpyAsync(c->d_stage, occ, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice,
                        c->stream));
  return upload_from_stage(c, chain, c->d_stage);
}

int cmg_upload_occupation_i32_dev(cmg_context *c, int chain, const int32_t *occ_dev,
                                  int64_t n) {
  NEED(c);
  if (chain < 0 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  if (!occ_dev || n != c->n_sites)
    return fail(c, CMG_EINVAL, "Error in set_occupation: size mismatch");
  return upload_from_stage(c, chain, occ_dev);
}

static int download_to(cmg_context *c, int chain, int32_t *dst_dev) {
  NaturalShape s = nat_shape(c);
  if (c->planar) {
    k_planes_to_i32<<<nblocks(c->n_sites, 256), 256, 0, c->stream>>>(
        c->d_planes + chain * c->chain_stride, c->plane_stride, dst_dev, s);
  } else {
    k_natural_to_i32<<<nblocks(c->n_sites, 256), 256, 0, c->stream>>>(
        c->d_nat + chain * c->n_sites, dst_dev, c->n_sites);
  }
  ++c->launches;
  CU(c, cudaGetLastError());
  return CMG_OK;
}

int cmg_download_occupation_i32(cmg_context *c, int chain, int32_t *occ, int64_t n) {
  NEED(c);
  if (chain < 0 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  if (!occ || n != c->n_sites) return fail(c, CMG_EINVAL, "size mismatch");
  int rc = ensure_stage(c);
  if (rc) return rc;
  rc = download_to(c, chain, c->d_stage);
  if (rc) return rc;
  CU(c, cudaMemcpyAsync(occ, c->d_stage, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost,
                        c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return device_error_check(c);
}

int cmg_download_occupation_i32_dev(cmg_context *c, int chain, int32_t *occ_dev, int64_t n) {
  NEED(c);
  if (chain < 0 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  if (!occ_dev || n != c->n_sites) return fail(c, CMG_EINVAL, "size mismatch");
  return download_to(c, chain, occ_dev);
}

// ---- compact host formats (1 B or 1 bit per site instead of the reference's 4 B) ----
static uint8_t *chain_base(cmg_context *c, int chain);

int cmg_upload_occupation_i8(cmg_context *c, int chain, const int8_t *occ, int64_t n) {
  NEED(c);
  if (chain < 0 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  if (!occ || n != c->n_sites)
    return fail(c, CMG_EINVAL, "Error in set_occupation: size mismatch");
  int rc = ensure_stage(c);
  if (rc) return rc;
  CU(c, cudaMemcpyAsync(c->d_stage, occ, (size_t)n, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaMemsetAsync(c->d_flag, 0, sizeof(int), c->stream));
  k_i8_to_sites<<<nblocks(n, 256), 256, 0, c->stream>>>(
      reinterpret_cast<const int8_t *>(c->d_stage), chain_base(c, chain), c->plane_stride,
      nat_shape(c), c->planar ? 1 : 0, c->d_flag);
  ++c->launches;
  if (c->planar) c->nat_is_current = false;
  CU(c, cudaGetLastError());
  int bad = 0;
  CU(c, cudaMemcpyAsync(&bad, c->d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  if (bad) return fail(c, CMG_EINVAL, "occupation values must be +1 or -1");
  return CMG_OK;
}

int cmg_download_occupation_i8(cmg_context *c, int chain, int8_t *occ, int64_t n) {
  NEED(c);
  if (chain < 0 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  if (!occ || n != c->n_sites) return fail(c, CMG_EINVAL, "size mismatch");
  int rc = ensure_stage(c);
  if (rc) return rc;
  k_sites_to_i8<<<nblocks(n, 256), 256, 0, c->stream>>>(
      chain_base(c, chain), c->plane_stride, reinterpret_cast<int8_t *>(c->d_stage), nat_shape(c),
      c->planar ? 1 : 0);
  ++c->launches;
  CU(c, cudaGetLastError());
  CU(c, cudaMemcpyAsync(occ, c->d_stage, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return device_error_check(c);
}

int cmg_upload_occupation_bits(cmg_context *c, int chain, const uint8_t *bits, int64_t n_sites) {
  NEED(c);
  if (chain < 0 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  if (!bits || n_sites != c->n_sites)
    return fail(c, CMG_EINVAL, "Error in set_occupation: size mismatch");
  int rc = ensure_stage(c);
  if (rc) return rc;
  const size_t nbytes = (size_t)((n_sites + 7) / 8);
  CU(c, cudaMemcpyAsync(c->d_stage, bits, nbytes, cudaMemcpyHostToDevice, c->stream));
  k_bits_to_sites<<<nblocks(n_sites, 256), 256, 0, c->stream>>>(
      reinterpret_cast<const uint8_t *>(c->d_stage), chain_base(c, chain), c->plane_stride,
      nat_shape(c), c->planar ? 1 : 0);
  ++c->launches;
  if (c->planar) c->nat_is_current = false;
  CU(c, cudaGetLastError());
  return CMG_OK;
}

int cmg_download_occupation_bits(cmg_context *c, int chain, uint8_t *bits, int64_t n_sites) {
  NEED(c);
  if (chain < 0 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  if (!bits || n_sites != c->n_sites) return fail(c, CMG_EINVAL, "size mismatch");
  int rc = ensure_stage(c);
  if (rc) return rc;
  const long long nbytes = (n_sites + 7) / 8;
  k_sites_to_bits<<<nblocks(nbytes, 256), 256, 0, c->stream>>>(
      chain_base(c, chain), c->plane_stride, reinterpret_cast<uint8_t *>(c->d_stage), nat_shape(c),
      c->planar ? 1 : 0);
  ++c->launches;
  CU(c, cudaGetLastError());
  CU(c, cudaMemcpyAsync(bits, c->d_stage, (size_t)nbytes, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return device_error_check(c);
}

int cmg_fill_occupation(cmg_context *c, int chain, int value) {
  NEED(c);
  if (chain < -1 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  if (value != 1 && value != -1) return fail(c, CMG_EINVAL, "fill value must be +1 or -1");
  const int lo = chain < 0 ? 0 : chain, hi = chain < 0 ? c->n_chains : chain + 1;
  const int b = value > 0 ? 1 : 0;
  if (c->planar) {
    CU(c, cudaMemsetAsync(c->d_planes + lo * c->chain_stride, b,
                          (size_t)c->chain_stride * (hi - lo), c->stream));
    c->nat_is_current = false;
  } else {
    CU(c, cudaMemsetAsync(c->d_nat + lo * c->n_sites, b, (size_t)c->n_sites * (hi - lo),
                          c->stream));
  }
  return CMG_OK;
}

// base pointer of a chain for single-site kernels: planes when two-coloured
static uint8_t *chain_base(cmg_context *c, int chain) {
  return c->planar ? c->d_planes + chain * c->chain_stride : c->d_nat + chain * c->n_sites;
}

int cmg_get_occ(cmg_context *c, int chain, int64_t l, int32_t *value) {
  NEED(c);
  if (chain < 0 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  if (l < 0 || l >= c->n_sites || !value) return fail(c, CMG_EINVAL, "site index out of range");
  k_get_occ<<<1, 1, 0, c->stream>>>(chain_base(c, chain), nat_shape(c), c->planar ? 1 : 0,
                                    c->plane_stride, l, c->d_flag);
  ++c->launches;
  CU(c, cudaGetLastError());
  int v = 0;
  CU(c, cudaMemcpyAsync(&v, c->d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  *value = v;
  return CMG_OK;
}

int cmg_set_occ(cmg_context *c, int chain, int64_t l, int32_t value) {
  NEED(c);
  if (chain < 0 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  if (l < 0 || l >= c->n_sites) return fail(c, CMG_EINVAL, "site index out of range");
  if (value != 1 && value != -1) return fail(c, CMG_EINVAL, "occupation values must be +1 or -1");
  k_set_occ<<<1, 1, 0, c->stream>>>(chain_base(c, chain), nat_shape(c), c->planar ? 1 : 0,
                                    c->plane_stride, l, value);
  ++c->launches;
  if (c->planar) c->nat_is_current = false;
  CU(c, cudaGetLastError());
  return CMG_OK;
}

int cmg_event_delta(cmg_context *c, int chain, int n_event, const int64_t *ls,
                    const int32_t *new_occ, double *dE_formation, double *dNx) {
  NEED(c);
  if (chain < 0 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  if (n_event < 0 || n_event > kMaxEventSites || (n_event > 0 && (!ls || !new_occ)))
    return fail(c, CMG_EINVAL, "event size must be in [0, 64]");
  for (int e = 0; e < n_event; ++e)
    if (ls[e] < 0 || ls[e] >= c->n_sites) return fail(c, CMG_EINVAL, "site index out of range");
  if (n_event == 0) {
    if (dE_formation) *dE_formation = 0.0;
    if (dNx) *dNx = 0.0;
    return CMG_OK;
  }
  struct {
    long long l[kMaxEventSites];
    int v[kMaxEventSites];
  } h;
  for (int e = 0; e < n_event; ++e) {
    h.l[e] = ls[e];
    h.v[e] = new_occ[e];
  }
  if (!c->d_event) CU(c, cudaMalloc(&c->d_event, sizeof h + 2 * sizeof(double)));
  CU(c, cudaMemcpyAsync(c->d_event, &h, sizeof h, cudaMemcpyHostToDevice, c->stream));
  double *d_out = reinterpret_cast<double *>(c->d_event + sizeof h);
  k_event_delta<<<1, 1, 0, c->stream>>>(chain_base(c, chain), nat_shape(c), c->planar ? 1 : 0,
                                        c->plane_stride, c->J, n_event,
                                        reinterpret_cast<const long long *>(c->d_event),
                                        reinterpret_cast<const int *>(c->d_event + sizeof h.l),
                                        d_out);
  ++c->launches;
  CU(c, cudaGetLastError());
  double out[2];
  CU(c, cudaMemcpyAsync(out, d_out, sizeof out, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  if (dE_formation) *dE_formation = out[0];
  if (dNx) *dNx = out[1];
  return CMG_OK;
}

int cmg_randomize_occupation(cmg_context *c, int chain, uint64_t seed, double p_up) {
  NEED(c);
  if (chain < -1 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  if (!(p_up >= 0.0 && p_up <= 1.0)) return fail(c, CMG_EINVAL, "p_up outside [0,1]");
  int rc = ensure_nat(c);
  if (rc) return rc;
  const int lo = chain < 0 ? 0 : chain, hi = chain < 0 ? c->n_chains : chain + 1;
  double scaled = std::ceil(p_up * 4294967296.0);
  const int always = scaled >= 4294967296.0;
  const uint32_t thr = scaled < 1.0 ? 0u : (uint32_t)((unsigned long long)scaled - 1ull);
  if (c->planar && !(lo == 0 && hi == c->n_chains)) {
    rc = sync_nat_from_planes(c);
    if (rc) return rc;
  }
  // slabs key the draw on global site indices (n0 * col_begin is a multiple of 4:
  // n0 and col_begin are even)
  const long long g_offset = c->slab ? (long long)c->shape[0] * c->col_begin / 4 : 0;
  for (int ch = lo; ch < hi; ++ch) {
    if (scaled < 1.0) {
      CU(c, cudaMemsetAsync(c->d_nat + ch * c->n_sites, 0, (size_t)c->n_sites, c->stream));
    } else {
      k_randomize_natural<<<nblocks((c->n_sites + 3) / 4, 256), 256, 0, c->stream>>>(
          c->d_nat + ch * c->n_sites, c->n_sites, seed + 0x9E3779B97F4A7C15ull * (uint64_t)ch,
          thr, always, g_offset);
      ++c->launches;
    }
  }
  CU(c, cudaGetLastError());
  if (c->planar) {
    rc = sync_planes_from_nat(c);
    if (rc) return rc;
    c->nat_is_current = true;
  }
  return CMG_OK;
}

// ---- RNG -------------------------------------------------------------------------
int cmg_seed_philox(cmg_context *c, uint64_t seed) {
  NEED(c);
  c->philox_seed = seed;
  return CMG_OK;
}

int cmg_set_philox_rounds(cmg_context *c, int rounds) {
  NEED(c);
  if (rounds != 10 && rounds != 7)
    return fail(c, CMG_EINVAL, "philox rounds: 10 (default) or 7 (the fewest that pass BigCrush)");
  if (rounds != 10 && c->ks_K) return fail(c, CMG_EUNSUPPORTED, "the k-state model runs Philox4x32-10 only");
  c->philox_rounds = rounds;
  return CMG_OK;
}

static int push_run_state(cmg_context *c) {
  RunState rs;
  rs.pass = c->h_pass;
  rs.n_samples = c->n_samples;
  CU(c, cudaMemcpyAsync(c->d_run, &rs, sizeof rs, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return CMG_OK;
}

int cmg_set_chain_offset(cmg_context *c, int64_t global_index_of_chain_0) {
  NEED(c);
  if (global_index_of_chain_0 < 0 || global_index_of_chain_0 + c->n_chains > (1 << 24))
    return fail(c, CMG_EINVAL, "chain offset out of range");
  c->chain_offset = (int)global_index_of_chain_0;
  return CMG_OK;
}

int cmg_set_pass_counter(cmg_context *c, uint64_t pass_index) {
  NEED(c);
  c->h_pass = pass_index;
  return push_run_state(c);
}

static int ensure_engines(cmg_context *c) {
  if (!c->d_engines) {
    CU(c, cudaMalloc(&c->d_engines, sizeof(MT64State) * c->n_chains));
    CU(c, cudaMemset(c->d_engines, 0, sizeof(MT64State) * c->n_chains));
  }
  return CMG_OK;
}

// std::mt19937_64::seed(value): x[0] = value, x[i] = f*(x[i-1]^(x[i-1]>>62)) + i,
// position = 312 so the first draw regenerates (C++11 [rand.eng.mers])
int cmg_seed_mt19937_64(cmg_context *c, int chain, uint64_t seed) {
  NEED(c);
  if (chain < 0 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  int rc = ensure_engines(c);
  if (rc) return rc;
  MT64State st;
  memset(&st, 0, sizeof st);
  st.x[0] = seed;
  for (int i = 1; i < 312; ++i)
    st.x[i] = 6364136223846793005ull * (st.x[i - 1] ^ (st.x[i - 1] >> 62)) + (unsigned long long)i;
  st.pos = 312;
  CU(c, cudaMemcpyAsync(c->d_engines + chain, &st, sizeof st, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  c->engine_seeded[chain] = 1;
  return CMG_OK;
}

int cmg_set_mt19937_64_state(cmg_context *c, int chain, const uint64_t *state312, int position) {
  NEED(c);
  if (chain < 0 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  if (!state312 || position < 0 || position > 312)
    return fail(c, CMG_EINVAL, "bad engine state");
  int rc = ensure_engines(c);
  if (rc) return rc;
  MT64State st;
  memset(&st, 0, sizeof st);
  for (int i = 0; i < 312; ++i) st.x[i] = state312[i];
  st.pos = position;
  CU(c, cudaMemcpyAsync(c->d_engines + chain, &st, sizeof st, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  c->engine_seeded[chain] = 1;
  return CMG_OK;
}

int cmg_get_mt19937_64_state(cmg_context *c, int chain, uint64_t *state312, int *position) {
  NEED(c);
  if (chain < 0 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  if (!state312 || !position) return fail(c, CMG_EINVAL, "null argument");
  if (!c->d_engines || !c->engine_seeded[chain]) return fail(c, CMG_ESTATE, "engine not seeded");
  MT64State st;
  CU(c, cudaMemcpyAsync(&st, c->d_engines + chain, sizeof st, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < 312; ++i) state312[i] = st.x[i];
  *position = st.pos;
  return CMG_OK;
}

int cmg_rng_draw(cmg_context *c, int chain, int n, const int64_t *int_max, const double *real_max,
                 const uint8_t *is_real, int64_t *int_out, double *real_out) {
  NEED(c);
  if (chain < 0 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  if (n < 0 || !int_max || !real_max || !is_real || !int_out || !real_out)
    return fail(c, CMG_EINVAL, "null argument");
  if (!c->d_engines || !c->engine_seeded[chain]) return fail(c, CMG_ESTATE, "engine not seeded");
  if (n == 0) return CMG_OK;
  long long *d_imax = nullptr, *d_iout = nullptr;
  double *d_rmax = nullptr, *d_rout = nullptr;
  uint8_t *d_isr = nullptr;
  CU(c, cudaMalloc(&d_imax, 8 * n));
  CU(c, cudaMalloc(&d_iout, 8 * n));
  CU(c, cudaMalloc(&d_rmax, 8 * n));
  CU(c, cudaMalloc(&d_rout, 8 * n));
  CU(c, cudaMalloc(&d_isr, n));
  CU(c, cudaMemcpyAsync(d_imax, int_max, 8 * n, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaMemcpyAsync(d_rmax, real_max, 8 * n, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaMemcpyAsync(d_isr, is_real, n, cudaMemcpyHostToDevice, c->stream));
  k_rng_draw<<<1, 1, 0, c->stream>>>(c->d_engines + chain, n, d_imax, d_rmax, d_isr, d_iout, d_rout);
  ++c->launches;
  CU(c, cudaGetLastError());
  CU(c, cudaMemcpyAsync(int_out, d_iout, 8 * n, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(real_out, d_rout, 8 * n, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  cudaFree(d_imax);
  cudaFree(d_iout);
  cudaFree(d_rmax);
  cudaFree(d_rout);
  cudaFree(d_isr);
  return CMG_OK;
}

// ---- sample series storage -------------------------------------------------------
static int ensure_series(cmg_context *c, long long need) {
  if (need <= c->capacity) return CMG_OK;
  long long cap = c->capacity ? c->capacity : 1024;
  while (cap < need) cap *= 2;
  long long *nb = nullptr;
  const size_t slot_bytes = sizeof(long long) * 2 * (size_t)c->n_chains;
  if (c->aux) CU(c, cudaStreamSynchronize(c->aux));  // nobody reads the old series any more
  CU(c, cudaMalloc(&nb, slot_bytes * (size_t)cap));
  CU(c, cudaMemsetAsync(nb, 0, slot_bytes * (size_t)cap, c->stream));
  if (c->d_series && c->n_samples > 0)
    CU(c, cudaMemcpyAsync(nb, c->d_series, slot_bytes * (size_t)c->n_samples,
                          cudaMemcpyDeviceToDevice, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  cudaFree(c->d_series);
  c->d_series = nb;
  c->capacity = cap;
  return CMG_OK;
}

// ---- scratch memory of the statistics calls -----------------------------------------
// Stream-ordered allocations from the device's default pool, which is told to
// keep its memory: a completion check makes a dozen of these calls, and
// cudaMalloc / cudaFree (a device synchronisation each) dominated its cost.
#define scratch_alloc(pp, bytes, stream) scratch_alloc_raw(reinterpret_cast<void **>(pp), bytes, stream)
static cudaError_t scratch_alloc_raw(void **p, size_t bytes, cudaStream_t stream) {
  static thread_local bool pool_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !pool_set[dev]) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    pool_set[dev] = true;
  }
  return cudaMallocAsync(p, bytes, stream);
}
static void scratch_free(void *p, cudaStream_t stream) {
  if (p) cudaFreeAsync(p, stream);
}

// ---- the hot loop ------------------------------------------------------------------
// Geometry of the tiled kernel for this lattice: number of column tiles,
// passes per launch and dynamic shared memory.  ok = false if the lattice does
// not suit it (columns too tall for a useful tile, or n0 % 64 != 0).
struct TilePlan {
  bool ok = false;
  int n_tiles = 1, passes = 1, halo = 0, w_max = 0;
  size_t smem = 0;
};
static TilePlan plan_tiles(const cmg_context *c, long long passes_wanted) {
  TilePlan t;
  if (c->dim != 2 || c->slab || c->shape[0] % 64 != 0 || c->smem_optin < 64 * 1024) return t;
  const long long h = c->shape[0] / 2, n1 = c->shape[1];
  const long long budget = (long long)c->smem_optin - kSmemTile - 1024;  // tables + static smem
  if (2 * n1 * h <= budget) {  // whole lattice in one tile: periodic, no halo
    t.ok = true;
    t.n_tiles = 1;
    t.passes = (int)std::min<long long>(passes_wanted, 64);
    t.halo = 0;
    t.w_max = (int)n1;
    t.smem = (size_t)(2 * n1 * h) + kSmemTile;
    return t;
  }
  const int P = (int)std::min<long long>(passes_wanted, c->tile_passes);
  const long long H = 2 * P;
  const long long w_fit = budget / (2 * h);
  long long tw = w_fit - 2 * H;
  if (tw < 4 * H) return t;  // redundant halo work would exceed ~20 %
  long long n_tiles = (n1 + tw - 1) / tw;
  // use every SM: round the tile count up to a multiple of the SM count
  n_tiles = (n_tiles + c->sm_count - 1) / c->sm_count * c->sm_count;
  if (n_tiles > n1) n_tiles = n1;
  const long long tw_max = (n1 + n_tiles - 1) / n_tiles;
  if (tw_max < 4 * H && n_tiles > c->sm_count) {
    // too thin after rounding: fall back to the plain fit
    n_tiles = (n1 + tw - 1) / tw;
  }
  const long long tw_max2 = (n1 + n_tiles - 1) / n_tiles;
  if (tw_max2 + 2 * H > n1) return t;
  t.ok = true;
  t.n_tiles = (int)n_tiles;
  t.passes = P;
  t.halo = (int)H;
  t.w_max = (int)(tw_max2 + 2 * H);
  t.smem = (size_t)(2 * t.w_max * h) + kSmemTile;
  return t;
}

// Geometry of the ring kernel: the whole lattice resident in shared memory,
// one tile of whole columns per CTA, every CTA of the grid co-resident
// (cooperative launch), at least two columns per column group.
struct RingPlan {
  bool ok = false;
  int n_tiles = 0, w_max = 0;
  int nt = 512;  // threads per CTA: 512, or 128 for the short columns of n0 = 256 / 512
  size_t smem = 0;
};
static RingPlan plan_ring(const cmg_context *c) {
  RingPlan r;
  // (a slab runs resident only as part of a ring of slabs: cmg_slab_run_passes checks the peers)
  // n0 a multiple of 1024 (column groups of one to eight warps, 512 threads per CTA), or
  // n0 = 256 / 512 outside slabs (column groups of 8 / 16 threads, 128 threads per CTA: with
  // columns this short a 512-thread CTA would hold more groups than a mid-size lattice has
  // columns to give them)
  if (c->dim != 2 || !c->coop_launch) return r;
  const bool small = !c->slab && (c->shape[0] == 256 || c->shape[0] == 512);
  if (c->shape[0] % 1024 != 0 && !small) return r;
  const long long h = c->shape[0] / 2, n1 = c->shape[1], V = h / 16;
  r.nt = small ? 128 : 512;
  if (V > 256 || r.nt % V != 0) return r;  // at least two column groups per CTA
  const long long Q = r.nt / V;
  // A context whose completion checks run on the second stream NEXT TO the sweep (a check of
  // marked samples that was not prefetched) leaves one SM free when the widest tile stays the same (4096 / 147 and
  // 4096 / 148 both round up to 28 columns): the statistics kernels of cmg_series_check run there
  // while the cooperative kernel, whose CTAs fill the register file of their SMs, sweeps on.
  // Everybody else keeps all SMs (147 tiles measured 2.3 % slower than 148 on 4096^2: fewer
  // 27-column tiles to absorb the jitter of the edge exchange).
  long long n_tiles = c->sm_count / c->n_chains;
  if (c->reserve_sm && c->n_chains == 1 && n_tiles > 2 && (n1 + n_tiles - 2) / (n_tiles - 1) == (n1 + n_tiles - 1) / n_tiles) n_tiles -= 1;
  if (c->ring_tiles_cap > 0) n_tiles = std::min<long long>(n_tiles, c->ring_tiles_cap);
  n_tiles = std::min(n_tiles, n1 / (2 * Q));
  if (n_tiles < 2) return r;
  const long long w_max = (n1 + n_tiles - 1) / n_tiles;
  const long long smem = 2 * w_max * h + kSmemTile;
  if (smem + 6144 > (long long)c->smem_optin) return r;  // static: 4 KiB of per-pass sums
  r.ok = true;
  r.n_tiles = (int)n_tiles;
  r.w_max = (int)w_max;
  r.smem = (size_t)smem;
  return r;
}

static int pick_variant(cmg_context *c, long long n_passes) {
  if (c->forced_variant != V_AUTO) return c->forced_variant;
  if (c->philox_rounds != 10) {
    // the seven-round stream is compiled into the resident, the streaming 2-d and the generic kernel
    if (c->dim == 2 && n_passes >= 4 && plan_ring(c).ok && c->n_sites * c->n_chains <= (1ll << 25)) return V_RING2D;
    if (c->dim == 2 && c->shape[0] % 32 == 0) return V_BULK2D;
    return V_GENERIC;
  }
  // the tiled kernel wins while a half-sweep is short enough for launch ramp
  // and L2 latency to matter; very large batches stream better through bulk2d
  // (one lattice per CTA -- no halo work, many passes per launch -- wins at any number of
  // chains: the 1024 chains of 256^2 of BASELINE config 4 on one GPU 1.42e12 against 1.18e12
  // streamed)
  const TilePlan tp0 = plan_tiles(c, c->tile_passes);
  if (tp0.ok && (tp0.n_tiles == 1 || c->n_sites * c->n_chains <= (1ll << 25))) {
    // a lattice too large for one CTA but not for the GPU's shared memory stays
    // resident across the launch (no halo recomputation): 4096^2 1.4e12 vs 0.98e12
    // (a ring launch costs ~18 us of staging and set-up: worth it from 4 passes per call,
    // measured on one 4096^2 lattice; the trajectories are identical either way)
    if (n_passes >= 4 && plan_tiles(c, c->tile_passes).n_tiles > 1 && plan_ring(c).ok)
      return V_RING2D;
    // ONE lattice that fits a CTA but keeps it busy for more than the ~1.4 us an edge takes
    // from tile to tile through L2 (three or more vectors per thread and half-sweep) is spread
    // over several SMs too: 256^2 2.1e10 against 9.8e9 attempts/s in one CTA
    if (n_passes >= 4 && c->n_chains == 1 && c->n_sites >= 49152 && plan_ring(c).ok) return V_RING2D;
    return V_TILE2D;
  }
  if (c->dim == 2 && c->shape[0] % 32 == 0) return V_BULK2D;
  if (c->dim == 3 && c->shape[0] % 32 == 0) return V_BULK3D;
  return V_GENERIC;
}

static int pick_js(cmg_context *c, int variant) {
  if (c->js > 0) return c->js;
  if (c->js_auto[variant] > 0) return c->js_auto[variant];
  // CTAs of 128 threads that fit on the device at once (register-limited)
  int per_sm = 0;
  cudaError_t e = variant == V_BULK3D
                      ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_halfsweep_bulk3d<true>, 128, kSmemBulk3d)
                      : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_halfsweep_bulk2d<true>, 128, kSmemBulk2d);
  if (e != cudaSuccess || per_sm < 1) per_sm = 4;
  const double slots = (double)c->sm_count * per_sm;
  const long long V = c->shape[0] / 32;
  const long long layers = variant == V_BULK3D ? c->shape[2] : 1;
  // Strip length.  A launch takes ceil(CTAs / slots) waves and a wave lasts as
  // long as its longest strip plus the strip start-up (two extra column loads,
  // pipeline fill, table load: ~4 column-times), so minimise waves * (js + 4).
  // A ragged last strip does not shorten a wave, which is why strip lengths that
  // divide n1 win (512^3: js = 64 -> 512 CTAs in one wave, 1.17e12 against
  // 1.12e12 for js = 20 in 2.8 waves).  Multiples of four match the unrolled
  // column loop; 3-d strips must be even (two strips share a warp when n0 = 512
  // and the column parity has to be warp-uniform).
  static const int cand[] = {2,  4,  6,  8,  12, 16, 20, 24, 28, 32,  36,  40, 44,
                             48, 52, 56, 60, 64, 72, 80, 88, 96, 104, 112, 120, 128};
  int best = 16;
  double best_cost = 1e300;
  for (int js : cand) {
    if (js > c->shape[1] && js > 2) continue;
    const long long strips = (c->shape[1] + js - 1) / js;
    const double ctas = (double)nblocks(V * strips * layers, 128) * c->n_chains;
    // measured on 512^3: strips that leave a ragged remainder run ~10 % slower
    // than their CTA count predicts (js = 60: 1.06e12, js = 64: 1.19e12)
    const double ragged = (c->shape[1] % js) ? 1.1 : 1.0;
    const double cost = std::ceil(ctas / slots) * (js + 4.0) * ragged;
    if (cost <= best_cost) {  // ties: the longer strip
      best_cost = cost;
      best = js;
    }
  }
  c->js_auto[variant] = best;
  return best;
}

// 2-d: balanced strips (even starts, lengths within two columns of each other), their
// number chosen to minimise waves * (longest strip + ~4 column-times of start-up) with
// no upper limit on the strip length: a large lattice runs as ONE wave of long strips
// (65536 x 8192: 37 strips of 220-222 columns on 592 resident CTAs instead of 256
// strips of 32 in seven waves, whose start-up cost 14 %).
static int pick_strips2d(cmg_context *c) {
  const long long V = c->shape[0] / 32, n1 = c->shape[1];
  if (c->js > 0 || n1 % 2) return 0;
  if (c->n_strips2d > 1) return (int)std::min<long long>(c->n_strips2d, n1 / 2);
  if (c->js_auto[5] != 0) return std::max(c->js_auto[5], 0);
  // Measured on 65536 x {8192, 16384, 32768, 65536}, 16384^2 and 8 x 4096^2
  // (profiles/sweep2d_r2s.json): ONE wave of long strips is up to 2.5 % slower than
  // several waves (every CTA starts at once, so the start-up is exposed and nothing
  // rebalances a slow SM), and from two waves on the strip count matters by < 2 % --
  // except when wave quantisation pushes the uniform strips of pick_js below ~64
  // columns (65536 x 8192: js = 32 in 6.9 waves is 4 % slower than 74 balanced strips
  // of 110-112 columns in exactly two waves).  Only that case takes balanced strips.
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_halfsweep_bulk2d<true>, 128, kSmemBulk2d) != cudaSuccess || per_sm < 1)
    per_sm = 4;
  const double slots = (double)c->sm_count * per_sm;
  const int js = pick_js(c, V_BULK2D);
  int best = 0;
  if (js < 64 || c->n_strips2d == 1) {
    const double uniform_cost =
        std::ceil((double)nblocks(V * ((n1 + js - 1) / js), 128) * c->n_chains / slots) * (js + 2.0);
    double best_cost = c->n_strips2d == 1 ? 1e300 : 0.99 * uniform_cost;
    for (int waves = 2; waves <= 4; ++waves) {
      // the largest strip count that still fits `waves` waves
      long long S = (long long)(waves * slots / c->n_chains) * 128 / V;
      while (S > 1 && (double)nblocks(V * S, 128) * c->n_chains > waves * slots) --S;
      if (S < 2 || S > n1 / 2) continue;
      const double len = std::ceil((double)(n1 / 2) / S) * 2.0;
      if (len < 64) continue;
      const double cost = waves * (len + 2.0);
      if (cost < best_cost) {
        best_cost = cost;
        best = (int)S;
      }
    }
  }
  c->js_auto[5] = best > 0 ? best : -1;
  return best;
}

// 3-d: balanced strips, their number chosen so that the CTAs fill the resident
// places: the launch time is waves * (longest strip + ~4 column-times of start-up).
// With n0 < 1024 a warp holds several strips; they are made the same strip of
// layers k, k+2, ... (equal column parity), which needs n2 % (2 * strips per warp)
// == 0 -- otherwise the uniform strips of pick_js are used.
static int pick_strips3d(cmg_context *c, int *pair_layers) {
  const long long V = c->shape[0] / 32, n1 = c->shape[1], n2 = c->shape[2];
  const long long spw = V < 32 ? 32 / V : 1;
  *pair_layers = 0;
  if (c->js > 0 || n1 % 2 || V > 32 || (V < 32 && (32 % V || n2 % (2 * spw)))) return 0;
  *pair_layers = V < 32 ? 1 : 0;
  if (c->n_strips2d > 1) return (int)std::min<long long>(c->n_strips2d, n1 / 2);  // ":ns=<n>"
  if (c->js_auto[7] > 0) return c->js_auto[7];
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_halfsweep_bulk3d<true>, 128, kSmemBulk3d) != cudaSuccess || per_sm < 1)
    per_sm = 4;
  const double slots = (double)c->sm_count * per_sm;
  int best = 0;
  double best_cost = 1e300;
  for (long long S = 1; S <= n1 / 2; ++S) {
    const double len = std::ceil((double)(n1 / 2) / S) * 2.0;
    if (len > 128 && S < n1 / 2) continue;
    const double ctas = (double)nblocks(V * S * n2, 128) * c->n_chains;
    const double cost = std::ceil(ctas / slots) * (len + 4.0);
    if (cost < best_cost) {
      best_cost = cost;
      best = (int)S;
    }
  }
  c->js_auto[7] = best;
  return best;
}

// k_halfsweep_tma3d: K = 128 / V layers per CTA (V = n0 / 32 vectors per column), whole
// columns moved by bulk copies.  Strip count chosen like pick_strips3d's.
struct Tma3dPlan {
  bool ok;
  int n_strips, K;
  size_t smem;
};
static Tma3dPlan plan_tma3d(cmg_context *c) {
  Tma3dPlan r = {false, 0, 0, 0};
  if (c->dim != 3 || c->slab) return r;
  const long long n0 = c->shape[0], n1 = c->shape[1], n2 = c->shape[2];
  if ((n0 != 512 && n0 != 1024) || n1 % 2 || n1 < 2) return r;
  const long long V = n0 / 32, K = 128 / V;
  if (n2 % K) return r;
  r.K = (int)K;
  r.smem = (size_t)kSmemTile + (size_t)kTmaStages * (size_t)tma3d_stage_bytes((int)(n0 / 2));
  if (r.smem + 1024 > (size_t)c->smem_optin) return r;
  if (!c->tma3d_attr_set) {
    if (cudaFuncSetAttribute(k_halfsweep_tma3d<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)r.smem) != cudaSuccess ||
        cudaFuncSetAttribute(k_halfsweep_tma3d<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)r.smem) != cudaSuccess) {
      cudaGetLastError();
      return r;
    }
    c->tma3d_attr_set = true;
  }
  if (c->js_auto[6] > 0) {
    r.n_strips = c->js_auto[6];
    r.ok = true;
    return r;
  }
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_halfsweep_tma3d<true>, 128, r.smem) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    return r;
  }
  const double slots = (double)c->sm_count * per_sm;
  int best = 0;
  double best_cost = 1e300;
  const long long max_strips = c->js > 0 ? std::max<long long>(1, n1 / c->js) : n1 / 2;
  for (long long S = 1; S <= std::min(max_strips, n1 / 2); ++S) {
    const double len = std::ceil((double)(n1 / 2) / S) * 2.0;
    if (c->js > 0 && S != max_strips) continue;
    if (c->js <= 0 && len > 128 && S < n1 / 2) continue;
    const double ctas = (double)(S * (n2 / K)) * c->n_chains;
    const double cost = std::ceil(ctas / slots) * (len + 4.0);
    if (cost < best_cost) {
      best_cost = cost;
      best = (int)S;
    }
  }
  if (best < 1) return r;
  if (c->js <= 0) c->js_auto[6] = best;
  r.n_strips = best;
  r.ok = true;
  return r;
}

// A half-sweep kernel as a programmatic dependent of the kernel before it on the stream
// (pdl_launch_dependents / pdl_wait in the kernels): the ramp of one half-sweep overlaps
// the tail of the previous one.  Slab contexts (peer flags, pushes) keep plain launches.
static cudaError_t launch_dependent(const void *kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl,
                                    SweepArgs &A) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  void *args[] = {&A};
  return cudaLaunchKernelExC(&cfg, kernel, args);
}

// Per-CTA completion flags of the chained half-sweeps.  The first half-sweep of a run (the
// caller cleared hs_prev_chained) waits for the whole grid before it; the ones behind it
// wait for their neighbour CTAs only.
static int chain_half_sweep(cmg_context *c, SweepArgs &A, dim3 grid, bool pdl, long long cols_per_strip) {
  A.error = c->d_error;
  // Worth it when a CTA's strip is long against the wait (five to seven flags per thread and
  // a fence) and the grid fills the GPU: with a small grid the CTAs of the next half-sweep
  // would poll next to the ones that work (one 4096^2 lattice, 74 CTAs: 0.76e12 chained
  // against 0.88e12 waiting for the whole grid; 256^3 in 8-column strips: 0.70 against 0.76).
  const bool worth = cols_per_strip >= 24 && 2ll * grid.x * grid.y >= (long long)c->sm_count * CMG_BULK_CTAS;
  if (!pdl || !c->hs_chain || !worth) {
    c->hs_prev_chained = false;
    return CMG_OK;
  }
  const size_t need = (size_t)grid.x * grid.y;
  if (c->hs_flags_count < need) {
    CU(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->d_hs_flags);
    c->d_hs_flags = nullptr;
    CU(c, cudaMalloc(&c->d_hs_flags, need * sizeof(unsigned int)));
    CU(c, cudaMemset(c->d_hs_flags, 0, need * sizeof(unsigned int)));
    c->hs_flags_count = need;
    c->hs_prev_chained = false;
  }
  if (c->hs_epoch >= 0x7ffffff0u) {  // (2^31 half-sweeps: start the epochs over behind a full wait)
    CU(c, cudaMemsetAsync(c->d_hs_flags, 0, c->hs_flags_count * sizeof(unsigned int), c->stream));
    c->hs_epoch = 1;
    c->hs_prev_chained = false;
  }
  A.hs_flags = c->d_hs_flags;
  A.hs_epoch = ++c->hs_epoch;
  A.hs_wait = c->hs_prev_chained ? 1 : 0;
  c->hs_prev_chained = true;
  return CMG_OK;
}

// sample_kernel: run the sampling instantiation although this half-sweep is not sampled
// (sums computed and dropped).  Chained half-sweeps overlap in time, and two different
// kernels sharing an SM thrash its instruction cache (512^3 sampled every pass: 1.15e12
// with the two instantiations alternating, 1.33e12 unchained, 1.48e12 with one), so a 3-d
// run that samples often uses one instantiation throughout.  (The 2-d loops are half the
// size and share the cache: 8 x 4096^2 1.55e12 alternating, 1.50e12 with one.)
static int launch_half_sweep(cmg_context *c, int variant, int colour, unsigned long long pass,
                             bool sample, long long slot, bool sample_kernel = false) {
  SweepArgs A;
  memset(&A, 0, sizeof A);
  A.L = view(c);
  A.tabs = c->d_tabs;
  A.n_accept = c->d_n_accept;
  A.sb = sample ? c->d_series + slot * 2 * c->n_chains : nullptr;
  A.sb_chain_stride = 2;
  A.pass = pass;
  for (int r = 0; r < 10; ++r) {
    A.rk[2 * r] = (uint32_t)c->philox_seed + (uint32_t)r * kPhiloxW0;
    A.rk[2 * r + 1] = (uint32_t)(c->philox_seed >> 32) + (uint32_t)r * kPhiloxW1;
  }
  A.colour = colour;
  A.chain_offset = c->chain_offset;
  A.L.epoch = c->slab_epoch;
  if (!c->bulk_attr_set) {
    // the staging rings may exceed the 48 KiB default dynamic shared-memory limit
    cudaError_t e = cudaFuncSetAttribute(k_halfsweep_bulk2d<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBulk2d);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_halfsweep_bulk2d<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBulk2d);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_halfsweep_bulk2d<true, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBulk2d);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_halfsweep_bulk2d<false, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBulk2d);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_halfsweep_bulk3d<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBulk3d);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_halfsweep_bulk3d<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBulk3d);
    if (e != cudaSuccess) return fail(c, CMG_ECUDA, cudaGetErrorString(e));
    c->bulk_attr_set = true;
  }
  if (variant == V_TMA3D) {
    const Tma3dPlan tp = plan_tma3d(c);
    if (!tp.ok) return fail(c, CMG_EUNSUPPORTED, "tma3d does not fit this lattice");
    A.n_strips = tp.n_strips;
    dim3 grid((unsigned)(tp.n_strips * (c->shape[2] / tp.K)), c->n_chains);
    if (sample)
      k_halfsweep_tma3d<true><<<grid, dim3(128), tp.smem, c->stream>>>(A);
    else
      k_halfsweep_tma3d<false><<<grid, dim3(128), tp.smem, c->stream>>>(A);
    ++c->launches;
    return CMG_OK;
  }
  A.js = pick_js(c, variant);
  if (variant == V_BULK3D) A.n_strips = pick_strips3d(c, &A.pair_layers);
  if (variant == V_BULK2D) A.n_strips = pick_strips2d(c);
  const long long plane_size = c->n_sites / 2;
  dim3 block(128);
  if (variant == V_GENERIC) {
    dim3 grid(nblocks((plane_size + 7) / 8, 128), c->n_chains);
    if (sample)
      (c->philox_rounds == 7 ? k_halfsweep_generic<true, 7> : k_halfsweep_generic<true, 10>)<<<grid, block, kSmemSmall, c->stream>>>(A);
    else
      (c->philox_rounds == 7 ? k_halfsweep_generic<false, 7> : k_halfsweep_generic<false, 10>)<<<grid, block, kSmemSmall, c->stream>>>(A);
  } else if (variant == V_BULK2D) {
    const long long V = c->shape[0] / 32;
    const long long strips = A.n_strips > 0 ? A.n_strips : (c->shape[1] + A.js - 1) / A.js;
    dim3 grid(nblocks(V * strips, 128), c->n_chains);
    const bool pdl = c->pdl && !c->slab;
    int rc = chain_half_sweep(c, A, grid, pdl, c->shape[1] / strips);
    if (rc) return rc;
    cudaError_t e;
    if (sample)
      e = launch_dependent(c->philox_rounds == 7 ? (const void *)k_halfsweep_bulk2d<true, 7> : (const void *)k_halfsweep_bulk2d<true, 10>, grid, block, kSmemBulk2d, c->stream, pdl, A);
    else
      e = launch_dependent(c->philox_rounds == 7 ? (const void *)k_halfsweep_bulk2d<false, 7> : (const void *)k_halfsweep_bulk2d<false, 10>, grid, block, kSmemBulk2d, c->stream, pdl, A);
    if (e != cudaSuccess) return fail(c, CMG_ECUDA, cudaGetErrorString(e));
  } else if (variant == V_BULK3D) {
    const long long V = c->shape[0] / 32;
    const long long strips = A.n_strips > 0 ? A.n_strips : (c->shape[1] + A.js - 1) / A.js;
    dim3 grid(nblocks(V * strips * c->shape[2], 128), c->n_chains);
    const bool pdl = c->pdl && !c->slab;
    int rc = chain_half_sweep(c, A, grid, pdl, c->shape[1] / strips);
    if (rc) return rc;
    const cudaError_t e = (sample || (sample_kernel && A.hs_flags)) ? launch_dependent((const void *)k_halfsweep_bulk3d<true>, grid, block, kSmemBulk3d, c->stream, pdl, A)
                                 : launch_dependent((const void *)k_halfsweep_bulk3d<false>, grid, block, kSmemBulk3d, c->stream, pdl, A);
    if (e != cudaSuccess) return fail(c, CMG_ECUDA, cudaGetErrorString(e));
  } else {
    return fail(c, CMG_EUNSUPPORTED, "kernel variant not available");
  }
  ++c->launches;
  return CMG_OK;
}

static int launch_tile_passes_sp(cmg_context *c, const TilePlan &tp, int n_passes,
                                 long long sample_period) {
  TileArgs A;
  memset(&A, 0, sizeof A);
  A.L = view(c);
  A.tabs = c->d_tabs;
  A.n_accept = c->d_n_accept;
  A.sb = c->d_series ? c->d_series + c->n_samples * 2 * c->n_chains : nullptr;
  A.sb_chain_stride = 2;
  A.sb_slot_stride = 2 * c->n_chains;
  A.pass0 = c->h_pass;
  A.pass_phase = c->n_pass;
  A.sample_period = sample_period;
  for (int r = 0; r < 10; ++r) {
    A.rk[2 * r] = (uint32_t)c->philox_seed + (uint32_t)r * kPhiloxW0;
    A.rk[2 * r + 1] = (uint32_t)(c->philox_seed >> 32) + (uint32_t)r * kPhiloxW1;
  }
  A.n_passes = n_passes;
  A.chain_offset = c->chain_offset;
  A.n_tiles = tp.n_tiles;
  A.error = c->d_error;
  A.out_planes = c->d_planes;
  if (tp.n_tiles > 1) {  // tiles with halos never write back in place (TileArgs::out_planes)
    if (!c->d_planes_alt) CU(c, cudaMalloc(&c->d_planes_alt, (size_t)c->chain_stride * c->n_chains));
    A.out_planes = c->d_planes_alt;
  }
  A.halo = tp.n_tiles == 1 ? 0 : 2 * n_passes;
  A.w_max = tp.w_max;
  const unsigned long long V = (unsigned long long)(c->shape[0] / 32);
  A.v_magic = (uint32_t)((0x100000000ull + V - 1) / V);
  dim3 grid(tp.n_tiles, c->n_chains);
  cudaError_t e;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = grid;
  cfg.dynamicSmemBytes = tp.smem;
  cfg.stream = c->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  // (tiles with halos: launches of a few passes, 10-20 us each -- 512^2 +20 %, 768 x 2048 +15 %;
  // one lattice per CTA runs up to 64 passes per launch and gains nothing)
  cfg.numAttrs = (c->pdl && tp.n_tiles > 1) ? 1 : 0;
  void *args[] = {&A};
#define LAUNCH_TILE_S(NT, S)                                                                  \
  e = cudaFuncSetAttribute(k_tile2d<NT, S>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                           (int)tp.smem);                                                     \
  if (e != cudaSuccess) return fail(c, CMG_ECUDA, cudaGetErrorString(e));                     \
  cfg.blockDim = dim3(NT);                                                                    \
  e = cudaLaunchKernelExC(&cfg, (const void *)k_tile2d<NT, S>, args);                         \
  if (e != cudaSuccess) return fail(c, CMG_ECUDA, cudaGetErrorString(e));
#define LAUNCH_TILE(NT)                                                                       \
  if (tp.n_tiles == 1) {                                                                      \
    LAUNCH_TILE_S(NT, true)                                                                   \
  } else {                                                                                    \
    LAUNCH_TILE_S(NT, false)                                                                  \
  }
  // narrow tiles (mid-size lattices: a few owned columns plus halos) leave half the column
  // groups of a 512-thread CTA without a column; 256 threads do the same work with half the
  // per-half-sweep bookkeeping (512^2, 256 x 1024: +11 %)
  int nt = c->tile_threads;
  if (!c->tile_threads_forced && tp.n_tiles > 1 && (long long)tp.w_max * (c->shape[0] / 32) <= 256) nt = 256;
  if (nt == 1024) {
    LAUNCH_TILE(1024)
  } else if (nt == 256) {
    LAUNCH_TILE(256)
  } else if (nt == 640) {
    LAUNCH_TILE(640)
  } else if (nt == 768) {
    LAUNCH_TILE(768)
  } else {
    LAUNCH_TILE(512)
  }
#undef LAUNCH_TILE_S
#undef LAUNCH_TILE
  ++c->launches;
  if (tp.n_tiles > 1) std::swap(c->d_planes, c->d_planes_alt);  // (stream order: everything enqueued from here on sees the new planes)
  return CMG_OK;
}

static int launch_ring_passes(cmg_context *c, const RingPlan &rp, int n_passes,
                              long long sample_period, bool publish_first = false) {
  // edge mailbox: per tile and side one column of each plane; zeroed before every
  // launch because its bytes carry the half-sweep stamp that validates them
  const size_t mb_bytes = (size_t)c->n_chains * rp.n_tiles * 4 * (size_t)(c->shape[0] / 2);
  const bool peers = c->ring_peer_mb[0] && c->ring_peer_mb[1];
  if (c->ring_mailbox_bytes < mb_bytes) {
    if (peers) return fail(c, CMG_ESTATE, "ring2d: the mailbox the neighbours write into cannot grow");
    cudaFree(c->d_ring_mailbox);
    c->d_ring_mailbox = nullptr;
    CU(c, cudaMalloc(&c->d_ring_mailbox, mb_bytes));
    c->ring_mailbox_bytes = mb_bytes;
  }
  // (a mailbox the neighbour GPUs write into is never zeroed: their edges may arrive before
  // this launch starts; its stamps count the half-sweeps of the whole trajectory instead)
  if (!peers) CU(c, cudaMemsetAsync(c->d_ring_mailbox, 0, mb_bytes, c->stream));
  RingArgs A;
  memset(&A, 0, sizeof A);
  if (peers) {
    for (int side = 0; side < 2; ++side) {
      A.peer_mb[side] = c->ring_peer_mb[side];
      A.peer_tiles[side] = c->ring_peer_tiles[side];
    }
    A.stamp0 = (uint32_t)(c->ring_s0 % 127ull);
  }
  A.L = view(c);
  A.tabs = c->d_tabs;
  A.n_accept = c->d_n_accept;
  A.sb = c->d_series ? c->d_series + c->n_samples * 2 * c->n_chains : nullptr;
  A.sb_chain_stride = 2;
  A.sb_slot_stride = 2 * c->n_chains;
  A.pass0 = c->h_pass;
  A.pass_phase = c->n_pass;
  A.sample_period = sample_period;
  for (int r = 0; r < 10; ++r) {
    A.rk[2 * r] = (uint32_t)c->philox_seed + (uint32_t)r * kPhiloxW0;
    A.rk[2 * r + 1] = (uint32_t)(c->philox_seed >> 32) + (uint32_t)r * kPhiloxW1;
  }
  A.n_passes = n_passes;
  A.chain_offset = c->chain_offset;
  A.n_tiles = rp.n_tiles;
  A.w_max = rp.w_max;
  const unsigned long long V = (unsigned long long)(c->shape[0] / 32);
  A.v_magic = (uint32_t)((0x100000000ull + V - 1) / V);
  A.mailbox = c->d_ring_mailbox;
  A.error = c->d_error;
  const void *ring_kernel = rp.nt == 128 ? (c->philox_rounds == 7 ? (const void *)k_ring2d<128, 7> : (const void *)k_ring2d<128, 10>)
                                         : (c->philox_rounds == 7 ? (const void *)k_ring2d<512, 7> : (const void *)k_ring2d<512, 10>);
  cudaError_t e = cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rp.smem);
  if (e != cudaSuccess) return fail(c, CMG_ECUDA, cudaGetErrorString(e));
  dim3 grid(rp.n_tiles, c->n_chains);
  void *args[] = {&A};
  // cooperative: the grid starts only when every CTA can be resident, which
  // the edge waits between neighbouring tiles rely on
  if (peers && publish_first) {
    const int V = (int)(c->shape[0] / 32);
    k_ring_publish<<<dim3((unsigned)nblocks(V, 128), 2), 128, 0, c->stream>>>(A);
    ++c->launches;
  }
  e = cudaLaunchCooperativeKernel(ring_kernel, grid, dim3(rp.nt), args, rp.smem, c->stream);
  if (e != cudaSuccess) return fail(c, CMG_ECUDA, cudaGetErrorString(e));
  ++c->launches;
  if (peers) c->ring_s0 += 2ull * (unsigned long long)n_passes;
  return CMG_OK;
}

// the mailbox of a slab that may become part of a ring of slabs: allocated (and zeroed) once,
// before its address is handed to the neighbours
static int ensure_ring_mailbox(cmg_context *c, int *n_tiles) {
  *n_tiles = 0;
  const RingPlan rp = plan_ring(c);
  if (!rp.ok) return CMG_OK;
  const size_t mb_bytes = (size_t)c->n_chains * rp.n_tiles * 4 * (size_t)(c->shape[0] / 2);
  if (c->ring_mailbox_bytes < mb_bytes) {
    cudaFree(c->d_ring_mailbox);
    c->d_ring_mailbox = nullptr;
    CU(c, cudaMalloc(&c->d_ring_mailbox, mb_bytes));
    c->ring_mailbox_bytes = mb_bytes;
    CU(c, cudaMemset(c->d_ring_mailbox, 0, mb_bytes));
  }
  *n_tiles = rp.n_tiles;
  return CMG_OK;
}

static const char *variant_str(int v) {
  switch (v) {
    case V_RING2D: return "ring2d";
    case V_GENERIC: return "generic";
    case V_BULK2D: return "bulk2d";
    case V_BULK3D: return "bulk3d";
    case V_TMA3D: return "tma3d";
    case V_TILE2D: return "tile2d";
    default: return "auto";
  }
}

// line counts of sample `slot` for every chain (use_nlist = false energy form)
static int sample_lines(cmg_context *c, long long slot) {
  const long long per = c->shape[0] + c->shape[1];
  if (c->lines_capacity <= slot) {
    long long cap = c->lines_capacity ? c->lines_capacity : 256;
    while (cap <= slot) cap *= 2;
    int *nb = nullptr;
    CU(c, cudaMalloc(&nb, sizeof(int) * (size_t)cap * c->n_chains * per));
    if (c->d_lines && slot > 0)
      CU(c, cudaMemcpyAsync(nb, c->d_lines, sizeof(int) * (size_t)slot * c->n_chains * per,
                            cudaMemcpyDeviceToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->d_lines);
    c->d_lines = nb;
    c->lines_capacity = cap;
  }
  int *dst = c->d_lines + (size_t)slot * c->n_chains * per;
  CU(c, cudaMemsetAsync(dst, 0, sizeof(int) * (size_t)c->n_chains * per, c->stream));
  const int js = 64;
  const long long V = c->shape[0] / 32;
  const long long strips = (c->shape[1] + js - 1) / js;
  for (int ch = 0; ch < c->n_chains; ++ch) {
    int *row = dst + (size_t)ch * per;
    k_line_xor_planes<<<nblocks(V * strips, 128), 128, 0, c->stream>>>(view(c), ch, js, row,
                                                                       row + c->shape[0]);
    ++c->launches;
  }
  CU(c, cudaGetLastError());
  return CMG_OK;
}

static int check_ready(cmg_context *c) {
  for (int ch = 0; ch < c->n_chains; ++ch)
    if (!c->h_tabs[ch].valid) return fail(c, CMG_ESTATE, "conditions not set for every chain");
  return CMG_OK;
}

static int observables_now(cmg_context *c, int chain, long long *dst_dev /* {ones,B} */) {
  CU(c, cudaMemsetAsync(dst_dev, 0, 2 * sizeof(long long), c->stream));
  if (c->planar) {
    LatticeView L = view(c);
    k_observables_planes<<<(unsigned)std::min<long long>(nblocks(c->n_sites / 2, 256), 148 * 8),
                           256, 0, c->stream>>>(L, chain, dst_dev);
  } else {
    k_observables_natural<<<(unsigned)std::min<long long>(nblocks(c->n_sites, 256), 148 * 8), 256,
                            0, c->stream>>>(c->d_nat + chain * c->n_sites, nat_shape(c), dst_dev);
  }
  ++c->launches;
  CU(c, cudaGetLastError());
  return CMG_OK;
}

static int run_serial(cmg_context *c, long long n_passes, long long sample_period) {
  if (c->slab) return fail(c, CMG_EUNSUPPORTED, "serial reference mode is single-GPU only");
  for (int ch = 0; ch < c->n_chains; ++ch)
    if (!c->d_engines || !c->engine_seeded[ch])
      return fail(c, CMG_ESTATE, "mt19937_64 engine not seeded for every chain");
  int rc = sync_nat_from_planes(c);
  if (rc) return rc;
  for (int ch = 0; ch < c->n_chains; ++ch) {
    CU(c, cudaMemsetAsync(c->d_cur_sb + 2 * ch, 0, 2 * sizeof(long long), c->stream));
    k_observables_natural<<<(unsigned)std::min<long long>(nblocks(c->n_sites, 256), 148 * 8), 256,
                            0, c->stream>>>(c->d_nat + ch * c->n_sites, nat_shape(c),
                                            c->d_cur_sb + 2 * ch);
    ++c->launches;
  }
  long long n_new = 0;
  if (sample_period > 0)
    n_new = (c->n_pass + n_passes) / sample_period - c->n_pass / sample_period;
  rc = ensure_series(c, c->n_samples + n_new);
  if (rc) return rc;
  SerialArgs A;
  memset(&A, 0, sizeof A);
  A.nat = c->d_nat;
  A.shape = nat_shape(c);
  A.tabs = c->d_tabs;
  A.engines = c->d_engines;
  A.n_accept = c->d_n_accept;
  A.cur_sb = c->d_cur_sb;
  A.series = c->d_series + c->n_samples * 2 * c->n_chains;
  A.series_chain_stride = 2;
  A.n_passes = n_passes;
  A.sample_period = sample_period;
  A.pass_base = c->n_pass;
  // slots of consecutive samples are n_chains*2 apart: the kernel indexes
  // series + chain*2 + 2*slot, so give it a per-chain contiguous view only when
  // n_chains == 1; otherwise stage through a temporary
  long long *tmp = nullptr;
  if (c->n_chains > 1 && n_new > 0) {
    CU(c, scratch_alloc(&tmp, sizeof(long long) * 2 * (size_t)n_new * c->n_chains, c->stream));
    A.series = tmp;
    A.series_chain_stride = 2 * n_new;
  }
  // engine state + its tempered words + dE / prob tables
  size_t smem = 2 * 312 * sizeof(unsigned long long) + 32 * sizeof(double);
  A.use_smem = 0;
  if ((size_t)c->n_sites + smem <= 200 * 1024) {
    A.use_smem = 1;
    smem += (size_t)c->n_sites;
  }
  CU(c, cudaFuncSetAttribute(k_serial_reference, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(204 * 1024)));
  k_serial_reference<<<c->n_chains, kSerialThreads, smem, c->stream>>>(A);
  ++c->launches;
  CU(c, cudaGetLastError());
  if (tmp) {
    // tmp[chain][slot][2] -> series[slot][chain][2]
    for (int ch = 0; ch < c->n_chains; ++ch)
      CU(c, cudaMemcpy2DAsync(c->d_series + c->n_samples * 2 * c->n_chains + 2 * ch,
                              sizeof(long long) * 2 * c->n_chains, tmp + 2 * n_new * ch,
                              sizeof(long long) * 2, sizeof(long long) * 2, (size_t)n_new,
                              cudaMemcpyDeviceToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    scratch_free(tmp, c->stream);
  }
  c->n_samples += n_new;
  c->n_pass += n_passes;
  rc = sync_planes_from_nat(c);
  if (rc) return rc;
  c->nat_is_current = true;
  c->variant_name = "serial_reference";
  return CMG_OK;
}

int cmg_run_passes(cmg_context *c, int64_t n_passes, int mode, int64_t sample_period) {
  NEED(c);
  if (n_passes < 0 || sample_period < 0) return fail(c, CMG_EINVAL, "negative count");
  int rc = check_ready(c);
  if (rc) return rc;
  if (n_passes == 0) return CMG_OK;
  if (mode == CMG_MODE_SERIAL_REFERENCE) return run_serial(c, n_passes, sample_period);
  if (mode != CMG_MODE_CHECKERBOARD) return fail(c, CMG_EINVAL, "unknown mode");
  if (!c->planar)
    return fail(c, CMG_EINVAL,
                "checkerboard mode needs even extents; use CMG_MODE_SERIAL_REFERENCE");
  if (c->slab)
    return fail(c, CMG_ESTATE, "slab contexts are stepped with cmg_slab_half_sweep");
  if (c->nonlist && sample_period > 0) {
    // the row/column sums need the state at every sampled pass: advance to the next
    // sample, count the unequal pairs of every line, repeat
    if (c->dim != 2) return fail(c, CMG_EUNSUPPORTED, "the use_nlist=false energy form is 2-d only");
    c->nonlist = false;  // the inner calls take the ordinary path
    long long left = n_passes;
    rc = CMG_OK;
    while (left > 0 && rc == CMG_OK) {
      const long long chunk = std::min<long long>(left, sample_period - (c->n_pass % sample_period));
      rc = cmg_run_passes(c, chunk, mode, sample_period);
      if (rc == CMG_OK && (c->n_pass % sample_period) == 0) rc = sample_lines(c, c->n_samples - 1);
      left -= chunk;
    }
    c->nonlist = true;
    return rc;
  }
  const int variant = pick_variant(c, n_passes);
  if (variant == V_BULK2D && !(c->dim == 2 && c->shape[0] % 32 == 0))
    return fail(c, CMG_EINVAL, "bulk2d needs dim == 2 and n0 % 32 == 0");
  if (variant == V_BULK3D && !(c->dim == 3 && c->shape[0] % 32 == 0))
    return fail(c, CMG_EINVAL, "bulk3d needs dim == 3 and n0 % 32 == 0");
  if (variant == V_TMA3D && !plan_tma3d(c).ok)
    return fail(c, CMG_EINVAL, "tma3d needs dim == 3, n0 in {512, 1024}, n1 even and n2 a multiple of 4096 / n0");
  if (c->philox_rounds != 10 && (variant == V_TILE2D || variant == V_BULK3D || variant == V_TMA3D))
    return fail(c, CMG_EUNSUPPORTED, "seven Philox rounds: kernels ring2d, bulk2d and generic only");
  if (variant == V_TILE2D && !plan_tiles(c, 1).ok)
    return fail(c, CMG_EINVAL, "tile2d does not fit this lattice (need dim 2, n0 % 64 == 0, short columns)");
  if (variant == V_RING2D && !plan_ring(c).ok)
    return fail(c, CMG_EINVAL,
                "ring2d does not fit this lattice (need dim 2, n0 = 256, 512 or a power of two in {1024..8192}, "
                "the lattice within the GPU's shared memory, cooperative launch)");
  c->variant_name = variant_str(variant);
  long long n_new = 0;
  if (sample_period > 0)
    n_new = (c->n_pass + n_passes) / sample_period - c->n_pass / sample_period;
  rc = ensure_series(c, c->n_samples + n_new);
  if (rc) return rc;
  c->nat_is_current = false;
  if (variant == V_RING2D) {
    const RingPlan rp = plan_ring(c);
    long long left = n_passes;
    while (left > 0) {
      // at most kRingMaxPasses sampled passes per launch (shared per-pass accumulators)
      long long P = std::min<long long>(left, c->ring_passes);
      rc = launch_ring_passes(c, rp, (int)P, sample_period);
      if (rc) return rc;
      long long n_new_here = 0;
      if (sample_period > 0)
        n_new_here = (c->n_pass + P) / sample_period - c->n_pass / sample_period;
      c->h_pass += P;
      c->n_pass += P;
      c->n_samples += n_new_here;
      left -= P;
    }
    CU(c, cudaGetLastError());
    return CMG_OK;
  }
  if (variant == V_TILE2D) {
    long long left = n_passes;
    while (left > 0) {
      TilePlan tp = plan_tiles(c, left);
      if (!tp.ok) return fail(c, CMG_EINVAL, "tile2d does not fit this lattice (need dim 2, n0 % 64 == 0, short columns)");
      rc = launch_tile_passes_sp(c, tp, tp.passes, sample_period);
      if (rc) return rc;
      long long n_new_here = 0;
      if (sample_period > 0)
        n_new_here = (c->n_pass + tp.passes) / sample_period - c->n_pass / sample_period;
      c->h_pass += tp.passes;
      c->n_pass += tp.passes;
      c->n_samples += n_new_here;
      left -= tp.passes;
    }
    CU(c, cudaGetLastError());
    return CMG_OK;
  }
  c->hs_prev_chained = false;  // whatever is on the stream before this run is waited for as a whole
  for (long long t = 0; t < n_passes; ++t) {
    const bool sample = sample_period > 0 && ((c->n_pass + 1) % sample_period) == 0;
    const bool often = sample_period > 0 && sample_period <= 4;
    rc = launch_half_sweep(c, variant, 0, c->h_pass, false, 0, often);
    if (rc) return rc;
    rc = launch_half_sweep(c, variant, 1, c->h_pass, sample, c->n_samples, often);
    if (rc) return rc;
    ++c->h_pass;
    ++c->n_pass;
    if (sample) ++c->n_samples;
  }
  CU(c, cudaGetLastError());
  return CMG_OK;
}

int cmg_slab_half_sweep(cmg_context *c, int colour, uint64_t pass_index, int sample) {
  NEED(c);
  if (!c->slab) return fail(c, CMG_ESTATE, "not a slab context");
  if (colour != 0 && colour != 1) return fail(c, CMG_EINVAL, "colour must be 0 or 1");
  int rc = check_ready(c);
  if (rc) return rc;
  const bool do_sample = sample && colour == 1;
  if (do_sample) {
    rc = ensure_series(c, c->n_samples + 1);
    if (rc) return rc;
  }
  c->nat_is_current = false;
  c->variant_name = "bulk2d";
  c->hs_prev_chained = false;
  rc = launch_half_sweep(c, V_BULK2D, colour, pass_index, do_sample, c->n_samples);
  if (rc) return rc;
  // The neighbours' flags count fused half-sweeps since the peers were attached,
  // whatever pass indices they carry: restarting or re-running a trajectory
  // (cmg_set_pass_counter, a smaller pass_index) cannot wrap the epoch.  Every
  // rank of the ring must step the same sequence of half-sweeps.
  if (c->slab_exchange && (c->peer_flag[0] || c->peer_flag[1])) ++c->slab_epoch;
  if (colour == 1) {
    ++c->n_pass;
    c->h_pass = pass_index + 1;
    if (do_sample) ++c->n_samples;
  }
  CU(c, cudaGetLastError());
  return CMG_OK;
}

int cmg_slab_run_passes(cmg_context *c, int64_t n_passes, int64_t sample_period) {
  NEED(c);
  if (!c->slab) return fail(c, CMG_ESTATE, "not a slab context");
  if (n_passes < 0 || sample_period < 0) return fail(c, CMG_EINVAL, "negative count");
  if (c->slab_exchange && !(c->peer_flag[0] && c->peer_flag[1]))
    return fail(c, CMG_ESTATE,
                "cmg_slab_run_passes needs both neighbours attached (cmg_slab_ipc_attach); drivers "
                "with their own halo exchange step cmg_slab_half_sweep");
  if (c->forced_variant == V_RING2D) {
    // The slab stays resident in shared memory and its outer tiles exchange their edge
    // columns with the neighbour GPUs' outer tiles through the mailboxes (peer stores over
    // NVLink inside the cooperative kernel): one launch per GPU per block of passes.
    if (!c->slab_exchange || !(c->ring_peer_mb[0] && c->ring_peer_mb[1]))
      return fail(c, CMG_ESTATE, "ring2d on a slab needs both neighbours attached with a slab that can run resident too");
    int rc = check_ready(c);
    if (rc) return rc;
    const RingPlan rp = plan_ring(c);
    if (!rp.ok) return fail(c, CMG_EINVAL, "ring2d does not fit this slab");
    long long n_new = 0;
    if (sample_period > 0) n_new = (c->n_pass + n_passes) / sample_period - c->n_pass / sample_period;
    rc = ensure_series(c, c->n_samples + n_new);
    if (rc) return rc;
    c->nat_is_current = false;
    c->variant_name = "ring2d";
    // every call starts two stamps further on: what its first half-sweep consumes can only
    // come from the publish kernels of THIS call, whatever happened to the state in between
    // (upload, rollback); every rank of the ring makes the same calls
    c->ring_s0 += 2;
    long long left = n_passes;
    bool first = true;
    while (left > 0) {
      const long long P = std::min<long long>(left, c->ring_passes);
      rc = launch_ring_passes(c, rp, (int)P, sample_period, first);
      if (rc) return rc;
      first = false;
      long long n_new_here = 0;
      if (sample_period > 0) n_new_here = (c->n_pass + P) / sample_period - c->n_pass / sample_period;
      c->h_pass += P;
      c->n_pass += P;
      c->n_samples += n_new_here;
      left -= P;
    }
    if (n_passes > 0) {
      // leave the halos as the streaming kernel would: the neighbours' copies of our boundary
      // columns current, their flags one epoch further on
      LatticeView L = view(c);
      L.epoch = c->slab_epoch;
      k_slab_push_edges<<<dim3((unsigned)nblocks(c->shape[0] / 32, 128), 4), 128, 0, c->stream>>>(L);
      ++c->launches;
      ++c->slab_epoch;
    }
    CU(c, cudaGetLastError());
    return CMG_OK;
  }
  for (int64_t t = 0; t < n_passes; ++t) {
    const int sample = sample_period > 0 && ((c->n_pass + 1) % sample_period) == 0;
    const unsigned long long pass = c->h_pass;
    int rc = cmg_slab_half_sweep(c, 0, pass, 0);
    if (rc) return rc;
    rc = cmg_slab_half_sweep(c, 1, pass, sample);
    if (rc) return rc;
  }
  return CMG_OK;
}

int cmg_slab_set_halo_exchange(cmg_context *c, int enabled) {
  NEED(c);
  if (!c->slab) return fail(c, CMG_ESTATE, "not a slab context");
  c->slab_exchange = enabled != 0;
  return CMG_OK;
}

int cmg_slab_boundary_ptr(cmg_context *c, int colour, int side, void **dev_ptr, int64_t *n_bytes) {
  NEED(c);
  if (!c->slab) return fail(c, CMG_ESTATE, "not a slab context");
  if ((colour | side) & ~1) return fail(c, CMG_EINVAL, "colour/side must be 0 or 1");
  if (!dev_ptr || !n_bytes) return fail(c, CMG_EINVAL, "null argument");
  const long long hb = c->shape[0] / 2;
  *dev_ptr = c->d_planes + colour * c->plane_stride + (side ? hb * (c->shape[1] - 1) : 0);
  *n_bytes = hb;
  return CMG_OK;
}

int cmg_slab_halo_ptr(cmg_context *c, int colour, int side, void **dev_ptr, int64_t *n_bytes) {
  NEED(c);
  if (!c->slab) return fail(c, CMG_ESTATE, "not a slab context");
  if ((colour | side) & ~1) return fail(c, CMG_EINVAL, "colour/side must be 0 or 1");
  if (!dev_ptr || !n_bytes) return fail(c, CMG_EINVAL, "null argument");
  *dev_ptr = c->d_halo[colour][side];
  *n_bytes = c->shape[0] / 2;
  return CMG_OK;
}

struct SlabIpcBlob {
  cudaIpcMemHandle_t h[2][2];
  cudaIpcMemHandle_t flags;
  cudaIpcMemHandle_t mailbox;  // k_ring2d's edge mailbox (valid when ring_tiles > 0)
  int ring_tiles;              // tiles of the resident kernel on this slab, 0: it cannot run resident
  int pad;
};

int cmg_slab_ipc_export(cmg_context *c, void *handle_out, int64_t handle_bytes) {
  NEED(c);
  if (!c->slab) return fail(c, CMG_ESTATE, "not a slab context");
  if (!handle_out || handle_bytes < (int64_t)sizeof(SlabIpcBlob))
    return fail(c, CMG_EINVAL, "handle buffer too small (need 392 bytes)");
  SlabIpcBlob blob;
  memset(&blob, 0, sizeof blob);
  for (int col = 0; col < 2; ++col)
    for (int side = 0; side < 2; ++side)
      CU(c, cudaIpcGetMemHandle(&blob.h[col][side], c->d_halo[col][side]));
  CU(c, cudaIpcGetMemHandle(&blob.flags, c->d_flags));
  {
    const int rc = ensure_ring_mailbox(c, &blob.ring_tiles);
    if (rc) return rc;
    if (blob.ring_tiles > 0) CU(c, cudaIpcGetMemHandle(&blob.mailbox, c->d_ring_mailbox));
  }
  memcpy(handle_out, &blob, sizeof blob);
  return CMG_OK;
}

// side = which of OUR boundaries the peer sits on (0: peer is the low neighbour,
// so our column 0 is pushed into the peer's halo_hi; 1: high neighbour).
int cmg_slab_ipc_attach(cmg_context *c, int side, const void *handle, int64_t handle_bytes,
                        int same_process, cmg_context *peer) {
  NEED(c);
  if (!c->slab) return fail(c, CMG_ESTATE, "not a slab context");
  if (side & ~1) return fail(c, CMG_EINVAL, "side must be 0 or 1");
  if (same_process) {
    if (!peer || !peer->slab) return fail(c, CMG_EINVAL, "peer context missing");
    if (peer->device != c->device) {
      int can = 0;
      CU(c, cudaDeviceCanAccessPeer(&can, c->device, peer->device));
      if (!can) return fail(c, CMG_EUNSUPPORTED, "no peer access between the two devices");
      cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
        return fail(c, CMG_ECUDA, cudaGetErrorString(e));
      cudaGetLastError();
    }
    for (int col = 0; col < 2; ++col) c->push[col][side] = peer->d_halo[col][1 - side];
    c->peer_flag[side] = peer->d_flags + (1 - side);
    {
      int mine = 0, theirs = 0;
      int rc = ensure_ring_mailbox(c, &mine);
      if (rc == CMG_OK) rc = ensure_ring_mailbox(peer, &theirs);
      if (rc) return rc;
      c->ring_peer_mb[side] = (mine > 0 && theirs > 0) ? peer->d_ring_mailbox : nullptr;
      c->ring_peer_tiles[side] = theirs;
    }
    return CMG_OK;
  }
  if (!handle || handle_bytes < (int64_t)sizeof(SlabIpcBlob))
    return fail(c, CMG_EINVAL, "bad ipc handle");
  SlabIpcBlob blob;
  memcpy(&blob, handle, sizeof blob);
  for (int col = 0; col < 2; ++col) {
    void *p = nullptr;
    CU(c, cudaIpcOpenMemHandle(&p, blob.h[col][1 - side], cudaIpcMemLazyEnablePeerAccess));
    c->ipc_opened.push_back(p);
    c->push[col][side] = (uint8_t *)p;
  }
  {
    void *p = nullptr;
    CU(c, cudaIpcOpenMemHandle(&p, blob.flags, cudaIpcMemLazyEnablePeerAccess));
    c->ipc_opened.push_back(p);
    c->peer_flag[side] = (unsigned long long *)p + (1 - side);
  }
  {
    int mine = 0;
    const int rc = ensure_ring_mailbox(c, &mine);
    if (rc) return rc;
    c->ring_peer_mb[side] = nullptr;
    c->ring_peer_tiles[side] = blob.ring_tiles;
    if (mine > 0 && blob.ring_tiles > 0) {
      void *p = nullptr;
      CU(c, cudaIpcOpenMemHandle(&p, blob.mailbox, cudaIpcMemLazyEnablePeerAccess));
      c->ipc_opened.push_back(p);
      c->ring_peer_mb[side] = (uint8_t *)p;
    }
  }
  return CMG_OK;
}

int cmg_counters(cmg_context *c, int chain, int64_t *n_pass, int64_t *n_accept, int64_t *n_reject) {
  NEED(c);
  if (chain < 0 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  unsigned long long acc = 0;
  CU(c, cudaMemcpyAsync(&acc, c->d_n_accept + chain, sizeof acc, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  {
    const int rc = device_error_check(c);
    if (rc) return rc;
  }
  if (n_pass) *n_pass = c->n_pass;
  if (n_accept) *n_accept = (int64_t)acc;
  if (n_reject) *n_reject = c->n_pass * c->n_sites - (int64_t)acc;
  return CMG_OK;
}

int cmg_reset_counters(cmg_context *c) {
  NEED(c);
  CU(c, cudaMemsetAsync(c->d_n_accept, 0, sizeof(unsigned long long) * c->n_chains, c->stream));
  c->n_pass = 0;
  return CMG_OK;
}

// ---- sampling ---------------------------------------------------------------------
int cmg_sample_now(cmg_context *c, int chain, int64_t *S, int64_t *B) {
  NEED(c);
  if (chain < 0 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  int rc = observables_now(c, chain, c->d_scratch_sb);
  if (rc) return rc;
  long long sb[2];
  CU(c, cudaMemcpyAsync(sb, c->d_scratch_sb, sizeof sb, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  rc = device_error_check(c);
  if (rc) return rc;
  if (S) *S = 2 * sb[0] - c->n_sites;
  if (B) *B = sb[1];
  return CMG_OK;
}

int cmg_line_dots(cmg_context *c, int chain, int64_t *row_dots, int64_t *col_dots) {
  NEED(c);
  if (chain < 0 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  if (c->dim != 2) return fail(c, CMG_EUNSUPPORTED, "line dots are 2-d only");
  if (!row_dots || !col_dots) return fail(c, CMG_EINVAL, "null argument");
  int rc = sync_nat_from_planes(c);
  if (rc) return rc;
  long long *d = nullptr;
  const size_t n = (size_t)(c->shape[0] + c->shape[1]);
  CU(c, cudaMalloc(&d, sizeof(long long) * n));
  CU(c, cudaMemsetAsync(d, 0, sizeof(long long) * n, c->stream));
  k_line_dots<<<nblocks(c->n_sites, 256), 256, 0, c->stream>>>(c->d_nat + chain * c->n_sites,
                                                               nat_shape(c), d, d + c->shape[0]);
  ++c->launches;
  CU(c, cudaGetLastError());
  CU(c, cudaMemcpyAsync(row_dots, d, sizeof(long long) * c->shape[0], cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(col_dots, d + c->shape[0], sizeof(long long) * c->shape[1],
                        cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  cudaFree(d);
  return CMG_OK;
}

int cmg_n_samples(cmg_context *c, int64_t *n) {
  if (!c || !n) return fail(c, CMG_EINVAL, "null argument");
  *n = c->n_samples;
  return CMG_OK;
}

int cmg_clear_samples(cmg_context *c) {
  NEED(c);
  if (c->d_series)
    CU(c, cudaMemsetAsync(c->d_series, 0, sizeof(long long) * 2 * (size_t)c->n_chains * c->capacity,
                          c->stream));
  c->n_samples = 0;
  c->dbl_valid = 0;
  c->nf_first_sample = -1;
  return CMG_OK;
}

int cmg_read_samples_sb(cmg_context *c, int chain, int64_t first, int64_t count, int64_t *S,
                        int64_t *B) {
  NEED(c);
  if (chain < 0 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  if (first < 0 || count < 0 || first + count > c->n_samples)
    return fail(c, CMG_EINVAL, "sample range outside the series");
  if (count == 0) return CMG_OK;
  std::vector<long long> tmp(2 * (size_t)count);
  CU(c, cudaMemcpy2DAsync(tmp.data(), sizeof(long long) * 2,
                          c->d_series + first * 2 * c->n_chains + 2 * chain,
                          sizeof(long long) * 2 * c->n_chains, sizeof(long long) * 2, (size_t)count,
                          cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  {
    const int rc = device_error_check(c);
    if (rc) return rc;
  }
  for (long long i = 0; i < count; ++i) {
    if (S) S[i] = 2 * tmp[2 * i] - c->n_sites;
    if (B) B[i] = tmp[2 * i + 1];
  }
  return CMG_OK;
}

// convert new (ones,B) samples of every chain to the three double columns
static int ensure_doubles(cmg_context *c) {
  if (c->n_samples == 0) return CMG_OK;
  if (c->dbl_capacity < c->n_samples) {
    // grow, keeping what has been converted: those doubles were made with the
    // (J, mu) in force when they were converted and must never change afterwards
    long long cap = c->dbl_capacity ? c->dbl_capacity : 1024;
    while (cap < c->n_samples) cap *= 2;
    double *nb = nullptr;
    CU(c, cudaMalloc(&nb, sizeof(double) * 3 * (size_t)cap * c->n_chains));
    if (c->d_dbl && c->dbl_valid > 0)
      CU(c, cudaMemcpy2DAsync(nb, sizeof(double) * (size_t)cap, c->d_dbl,
                              sizeof(double) * (size_t)c->dbl_capacity,
                              sizeof(double) * (size_t)c->dbl_valid, (size_t)3 * c->n_chains,
                              cudaMemcpyDeviceToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->d_dbl);
    c->d_dbl = nullptr;
    c->d_dbl = nb;
    c->dbl_capacity = cap;
  }
  if (c->dbl_valid >= c->n_samples) return CMG_OK;
  const long long first = c->dbl_valid, count = c->n_samples - first;
  // per chain: gather the strided (ones,B) pairs into a contiguous temp, convert
  long long *tmp = nullptr;
  CU(c, scratch_alloc(&tmp, sizeof(long long) * 2 * (size_t)c->n_samples, c->stream));
  for (int ch = 0; ch < c->n_chains; ++ch) {
    CU(c, cudaMemcpy2DAsync(tmp + 2 * first, sizeof(long long) * 2,
                            c->d_series + first * 2 * c->n_chains + 2 * ch,
                            sizeof(long long) * 2 * c->n_chains, sizeof(long long) * 2,
                            (size_t)count, cudaMemcpyDeviceToDevice, c->stream));
    double *base = c->d_dbl + (size_t)3 * c->dbl_capacity * ch;
    k_series_to_doubles<<<nblocks(count, 256), 256, 0, c->stream>>>(
        tmp, first, count, c->n_sites, c->d_tabs + ch, base, base + c->dbl_capacity,
        base + 2 * c->dbl_capacity);
    ++c->launches;
    if (c->nonlist && c->d_lines) {
      // formation and potential energy in the row/column form (model.hh:273-285)
      const long long per = c->shape[0] + c->shape[1];
      k_nonlist_to_doubles<<<nblocks(count, 64), 64, 0, c->stream>>>(
          c->d_lines + (size_t)ch * per, per * c->n_chains, c->shape[0], c->shape[1], tmp, first,
          count, c->n_sites, c->d_tabs + ch, base + c->dbl_capacity, base + 2 * c->dbl_capacity);
      ++c->launches;
    }
  }
  CU(c, cudaGetLastError());
  CU(c, cudaStreamSynchronize(c->stream));
  scratch_free(tmp, c->stream);
  c->dbl_valid = c->n_samples;
  return CMG_OK;
}

int cmg_read_samples(cmg_context *c, int chain, int quantity, int64_t first, int64_t count,
                     double *out) {
  NEED(c);
  if (chain < 0 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  if (quantity < 0 || quantity > 2) return fail(c, CMG_EINVAL, "bad quantity");
  if (first < 0 || count < 0 || first + count > c->n_samples)
    return fail(c, CMG_EINVAL, "sample range outside the series");
  if (count == 0) return CMG_OK;
  if (!out) return fail(c, CMG_EINVAL, "null argument");
  int rc = ensure_doubles(c);
  if (rc) return rc;
  const double *src = c->d_dbl + (size_t)3 * c->dbl_capacity * chain +
                      (size_t)quantity * c->dbl_capacity + first;
  CU(c, cudaMemcpyAsync(out, src, sizeof(double) * (size_t)count, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return device_error_check(c);
}

// ---- probes -----------------------------------------------------------------------
int cmg_delta_e_probe(cmg_context *c, int chain, double *dE_per_site) {
  NEED(c);
  if (chain < 0 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  if (!dE_per_site) return fail(c, CMG_EINVAL, "null argument");
  if (!c->h_tabs[chain].valid) return fail(c, CMG_ESTATE, "conditions not set");
  if (c->slab) return fail(c, CMG_EUNSUPPORTED, "probe is single-GPU only");
  int rc = sync_nat_from_planes(c);
  if (rc) return rc;
  double *d = nullptr;
  CU(c, cudaMalloc(&d, sizeof(double) * (size_t)c->n_sites));
  k_delta_e_probe<<<nblocks(c->n_sites, 256), 256, 0, c->stream>>>(
      c->d_nat + chain * c->n_sites, nat_shape(c), c->d_tabs + chain, d);
  ++c->launches;
  CU(c, cudaGetLastError());
  CU(c, cudaMemcpyAsync(dE_per_site, d, sizeof(double) * (size_t)c->n_sites, cudaMemcpyDeviceToHost,
                        c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  cudaFree(d);
  return CMG_OK;
}

int cmg_accept_probe(cmg_context *c, int chain, const double *uniforms, uint8_t *accept) {
  NEED(c);
  if (chain < 0 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  if (!uniforms || !accept) return fail(c, CMG_EINVAL, "null argument");
  if (!c->h_tabs[chain].valid) return fail(c, CMG_ESTATE, "conditions not set");
  if (c->slab) return fail(c, CMG_EUNSUPPORTED, "probe is single-GPU only");
  int rc = sync_nat_from_planes(c);
  if (rc) return rc;
  double *du = nullptr;
  uint8_t *da = nullptr;
  CU(c, cudaMalloc(&du, sizeof(double) * (size_t)c->n_sites));
  CU(c, cudaMalloc(&da, (size_t)c->n_sites));
  CU(c, cudaMemcpyAsync(du, uniforms, sizeof(double) * (size_t)c->n_sites, cudaMemcpyHostToDevice,
                        c->stream));
  k_accept_probe<<<nblocks(c->n_sites, 256), 256, 0, c->stream>>>(
      c->d_nat + chain * c->n_sites, nat_shape(c), c->d_tabs + chain, du, da);
  ++c->launches;
  CU(c, cudaGetLastError());
  CU(c, cudaMemcpyAsync(accept, da, (size_t)c->n_sites, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  cudaFree(du);
  cudaFree(da);
  return CMG_OK;
}

// ---- series statistics ---------------------------------------------------------------
static int run_stats_jobs(cmg_context *c, cudaStream_t stream, const std::vector<SeriesJob> &jobs,
                          double confidence, double *mean, double *prec, double *var,
                          int64_t *k_star) {
  const int n = (int)jobs.size();
  if (n == 0) return CMG_OK;
  SeriesJob *dj = nullptr;
  double *dout = nullptr;
  long long *dk = nullptr;
  CU(c, scratch_alloc(&dj, sizeof(SeriesJob) * n, stream));
  CU(c, scratch_alloc(&dout, sizeof(double) * 4 * n, stream));
  CU(c, scratch_alloc(&dk, sizeof(long long) * n, stream));
  CU(c, cudaMemcpyAsync(dj, jobs.data(), sizeof(SeriesJob) * n, cudaMemcpyHostToDevice, stream));
  k_series_stats<<<n, 256, 0, stream>>>(dj, z_confidence(confidence), dout, dk);
  if (c) ++c->launches;
  CU(c, cudaGetLastError());
  std::vector<double> h(4 * (size_t)n);
  std::vector<long long> hk(n);
  CU(c, cudaMemcpyAsync(h.data(), dout, sizeof(double) * 4 * n, cudaMemcpyDeviceToHost, stream));
  CU(c, cudaMemcpyAsync(hk.data(), dk, sizeof(long long) * n, cudaMemcpyDeviceToHost, stream));
  CU(c, cudaStreamSynchronize(stream));
  for (int i = 0; i < n; ++i) {
    if (mean) mean[i] = h[4 * i];
    if (var) var[i] = h[4 * i + 1];
    if (prec) prec[i] = h[4 * i + 3];
    if (k_star) k_star[i] = hk[i];
  }
  scratch_free(dj, stream);
  scratch_free(dout, stream);
  scratch_free(dk, stream);
  return CMG_OK;
}

static int run_equil_jobs(cmg_context *c, cudaStream_t stream, const std::vector<SeriesJob> &jobs,
                          double prec, int *is_eq, int64_t *n_eq) {
  const int n = (int)jobs.size();
  if (n == 0) return CMG_OK;
  SeriesJob *dj = nullptr;
  int *de = nullptr;
  long long *dn = nullptr;
  CU(c, scratch_alloc(&dj, sizeof(SeriesJob) * n, stream));
  CU(c, scratch_alloc(&de, sizeof(int) * n, stream));
  CU(c, scratch_alloc(&dn, sizeof(long long) * n, stream));
  CU(c, cudaMemcpyAsync(dj, jobs.data(), sizeof(SeriesJob) * n, cudaMemcpyHostToDevice, stream));
  k_series_equilibration<<<n, kEquilThreads, 0, stream>>>(dj, n, prec, de, dn);
  if (c) ++c->launches;
  CU(c, cudaGetLastError());
  std::vector<int> he(n);
  std::vector<long long> hn(n);
  CU(c, cudaMemcpyAsync(he.data(), de, sizeof(int) * n, cudaMemcpyDeviceToHost, stream));
  CU(c, cudaMemcpyAsync(hn.data(), dn, sizeof(long long) * n, cudaMemcpyDeviceToHost, stream));
  CU(c, cudaStreamSynchronize(stream));
  for (int i = 0; i < n; ++i) {
    if (is_eq) is_eq[i] = he[i];
    if (n_eq) n_eq[i] = hn[i];
  }
  scratch_free(dj, stream);
  scratch_free(de, stream);
  scratch_free(dn, stream);
  return CMG_OK;
}

static const double *series_ptr(const cmg_context *c, int chain, int quantity) {
  return c->d_dbl + (size_t)3 * c->dbl_capacity * chain + (size_t)quantity * c->dbl_capacity;
}

int cmg_series_stats(cmg_context *c, int chain, int quantity, int64_t first, int64_t count,
                     double confidence, double *mean, double *calculated_precision,
                     double *variance, int64_t *k_star) {
  NEED(c);
  if (chain < 0 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  if (quantity < 0 || quantity > 2) return fail(c, CMG_EINVAL, "bad quantity");
  if (count <= 0)
    return fail(c, CMG_EINVAL, "Error in BasicStatisticsCalculator: observations.size()==0");
  if (first < 0 || first + count > c->n_samples)
    return fail(c, CMG_EINVAL, "sample range outside the series");
  int rc = ensure_doubles(c);
  if (rc) return rc;
  std::vector<SeriesJob> jobs(1);
  jobs[0].x = series_ptr(c, chain, quantity) + first;
  jobs[0].n = count;
  return run_stats_jobs(c, c->stream, jobs, confidence, mean, calculated_precision, variance, k_star);
}

int cmg_series_stats_all(cmg_context *c, int quantity, const int64_t *first, int64_t count_total,
                         double confidence, double *mean, double *calculated_precision,
                         double *variance, int64_t *k_star) {
  NEED(c);
  if (quantity < 0 || quantity > 2) return fail(c, CMG_EINVAL, "bad quantity");
  if (count_total <= 0 || count_total > c->n_samples)
    return fail(c, CMG_EINVAL, "sample range outside the series");
  int rc = ensure_doubles(c);
  if (rc) return rc;
  std::vector<SeriesJob> jobs(c->n_chains);
  for (int ch = 0; ch < c->n_chains; ++ch) {
    const long long f = first ? first[ch] : 0;
    if (f < 0 || f > count_total) return fail(c, CMG_EINVAL, "bad first index");
    jobs[ch].x = series_ptr(c, ch, quantity) + f;
    jobs[ch].n = count_total - f;
  }
  return run_stats_jobs(c, c->stream, jobs, confidence, mean, calculated_precision, variance, k_star);
}

int cmg_series_equilibration(cmg_context *c, int chain, int quantity, int64_t count,
                             double abs_precision, int *is_equilibrated, int64_t *n_equil) {
  NEED(c);
  if (chain < 0 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  if (quantity < 0 || quantity > 2) return fail(c, CMG_EINVAL, "bad quantity");
  if (count <= 0) return fail(c, CMG_EINVAL, "Error in equilibration_check: observations.size()==0");
  if (count > c->n_samples) return fail(c, CMG_EINVAL, "sample range outside the series");
  int rc = ensure_doubles(c);
  if (rc) return rc;
  std::vector<SeriesJob> jobs(1);
  jobs[0].x = series_ptr(c, chain, quantity);
  jobs[0].n = count;
  return run_equil_jobs(c, c->stream, jobs, abs_precision, is_equilibrated, n_equil);
}

int cmg_series_equilibration_all(cmg_context *c, int quantity, int64_t count, double abs_precision,
                                 int *is_equilibrated, int64_t *n_equil) {
  NEED(c);
  if (quantity < 0 || quantity > 2) return fail(c, CMG_EINVAL, "bad quantity");
  if (count <= 0 || count > c->n_samples)
    return fail(c, CMG_EINVAL, "sample range outside the series");
  int rc = ensure_doubles(c);
  if (rc) return rc;
  std::vector<SeriesJob> jobs(c->n_chains);
  for (int ch = 0; ch < c->n_chains; ++ch) {
    jobs[ch].x = series_ptr(c, ch, quantity);
    jobs[ch].n = count;
  }
  return run_equil_jobs(c, c->stream, jobs, abs_precision, is_equilibrated, n_equil);
}

int cmg_mark(cmg_context *c) {
  NEED(c);
  if (!c->planar || c->slab || c->ks_K) return fail(c, CMG_EUNSUPPORTED, "cmg_mark: checkerboard contexts only");
  if (!c->aux) {
    int lo = 0, hi = 0;
    CU(c, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CU(c, cudaStreamCreateWithPriority(&c->aux, cudaStreamNonBlocking, hi));
    CU(c, cudaEventCreateWithFlags(&c->ev_mark, cudaEventDisableTiming));
  }
  if (!c->d_shadow) {
    CU(c, cudaMalloc(&c->d_shadow, (size_t)c->chain_stride * c->n_chains));
    CU(c, cudaMalloc(&c->d_shadow_accept, sizeof(unsigned long long) * c->n_chains));
  }
  // everything enqueued so far is "before the mark": the aux stream waits for it, the copy
  // of the state follows it, whatever is enqueued next can be undone
  CU(c, cudaEventRecord(c->ev_mark, c->stream));
  CU(c, cudaStreamWaitEvent(c->aux, c->ev_mark, 0));
  CU(c, cudaMemcpyAsync(c->d_shadow, c->d_planes, (size_t)c->chain_stride * c->n_chains,
                        cudaMemcpyDeviceToDevice, c->stream));
  CU(c, cudaMemcpyAsync(c->d_shadow_accept, c->d_n_accept, sizeof(unsigned long long) * c->n_chains,
                        cudaMemcpyDeviceToDevice, c->stream));
  c->mark_h_pass = c->h_pass;
  c->mark_n_pass = c->n_pass;
  c->mark_n_samples = c->n_samples;
  c->mark_valid = true;
  return CMG_OK;
}

int cmg_rollback(cmg_context *c) {
  NEED(c);
  if (!c->mark_valid) return fail(c, CMG_ESTATE, "cmg_rollback: no mark");
  CU(c, cudaMemcpyAsync(c->d_planes, c->d_shadow, (size_t)c->chain_stride * c->n_chains,
                        cudaMemcpyDeviceToDevice, c->stream));
  CU(c, cudaMemcpyAsync(c->d_n_accept, c->d_shadow_accept, sizeof(unsigned long long) * c->n_chains,
                        cudaMemcpyDeviceToDevice, c->stream));
  // sample slots are accumulated into: the ones taken after the mark go back to zero
  if (c->n_samples > c->mark_n_samples && c->d_series)
    CU(c, cudaMemsetAsync(c->d_series + c->mark_n_samples * 2 * c->n_chains, 0,
                          sizeof(long long) * 2 * (size_t)c->n_chains * (size_t)(c->n_samples - c->mark_n_samples),
                          c->stream));
  c->h_pass = c->mark_h_pass;
  c->n_pass = c->mark_n_pass;
  c->n_samples = c->mark_n_samples;
  if (c->dbl_valid > c->n_samples) c->dbl_valid = c->n_samples;
  c->nat_is_current = false;
  c->mark_valid = false;
  return CMG_OK;
}

// one scratch block of a completion check: eq jobs | stat jobs | results | the error word
struct SeriesCheckBlock {
  SeriesJob eq[3], st[3];
  long long n_eq[3], n_stats, k_star[3];
  double out4[12];
  int is_eq[4];
  unsigned int error;
};

// The check of the first `count` samples, ENQUEUED on the context's stream and not waited
// for: a caller that knows a check is due enqueues it, then its next (speculative) block of
// passes, and collects the verdict with cmg_series_check (same arguments) while that block
// runs -- the device never idles between a block and the host's decision.  The series up to
// `count` is complete in stream order; later blocks only append.
int cmg_series_check_prefetch(cmg_context *c, int chain, int n_components, const int *quantity,
                              const double *abs_precision, int64_t count, double confidence) {
  NEED(c);
  if (chain < 0 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  if (n_components < 1 || n_components > 3 || !quantity || !abs_precision)
    return fail(c, CMG_EINVAL, "1 to 3 components");
  for (int i = 0; i < n_components; ++i)
    if (quantity[i] < 0 || quantity[i] > 2 || !(abs_precision[i] >= 0.0))
      return fail(c, CMG_EINVAL, "bad quantity or precision");
  if (count <= 0) return fail(c, CMG_EINVAL, "Error in equilibration_check: observations.size()==0");
  if (count > c->n_samples) return fail(c, CMG_EINVAL, "sample range outside the series");
  if (!c->d_check) {
    CU(c, cudaMalloc(&c->d_check, sizeof(SeriesCheckBlock)));
    CU(c, cudaMallocHost(&c->h_check_in, sizeof(SeriesCheckBlock)));
    CU(c, cudaMallocHost(&c->h_check_out, sizeof(SeriesCheckBlock)));
    CU(c, cudaEventCreateWithFlags(&c->ev_check, cudaEventDisableTiming));
  }
  if (c->check_pending) CU(c, cudaEventSynchronize(c->ev_check));  // the staging block is free again
  c->check_pending = false;
  int rc = ensure_doubles(c);
  if (rc) return rc;
  const int n = n_components;
  SeriesCheckBlock &h = *c->h_check_in;
  memset(&h, 0, sizeof h);
  for (int i = 0; i < n; ++i) {
    h.eq[i].x = series_ptr(c, chain, quantity[i]);
    h.eq[i].n = count;
  }
  SeriesCheckBlock *d = c->d_check;
  CU(c, cudaMemcpyAsync(d, &h, sizeof h, cudaMemcpyHostToDevice, c->stream));
  bool same = true;
  for (int i = 1; i < n; ++i) same = same && abs_precision[i] == abs_precision[0];
  if (same) {
    k_series_equilibration<<<n, kEquilThreads, 0, c->stream>>>(d->eq, n, abs_precision[0], d->is_eq, d->n_eq);
    ++c->launches;
  } else {
    for (int j = 0; j < n; ++j) {
      k_series_equilibration<<<1, kEquilThreads, 0, c->stream>>>(d->eq + j, 1, abs_precision[j], d->is_eq + j, d->n_eq + j);
      ++c->launches;
    }
  }
  k_make_tail_jobs<<<1, 32, 0, c->stream>>>(d->eq, n, d->is_eq, d->n_eq, d->st, &d->n_stats);
  k_series_stats<<<n, 256, 0, c->stream>>>(d->st, z_confidence(confidence), d->out4, d->k_star);
  c->launches += 2;
  CU(c, cudaGetLastError());
  if (c->d_error)
    CU(c, cudaMemcpyAsync(&d->error, c->d_error, sizeof(unsigned int), cudaMemcpyDeviceToDevice, c->stream));
  CU(c, cudaMemcpyAsync(c->h_check_out, d, sizeof(SeriesCheckBlock), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaEventRecord(c->ev_check, c->stream));
  c->check_pending = true;
  c->check_key_n = n;
  c->check_key_chain = chain;
  c->check_key_count = count;
  c->check_key_conf = confidence;
  for (int i = 0; i < n; ++i) {
    c->check_key_q[i] = quantity[i];
    c->check_key_abs[i] = abs_precision[i];
  }
  return CMG_OK;
}

int cmg_series_check(cmg_context *c, int chain, int n_components, const int *quantity,
                     const double *abs_precision, int64_t count, double confidence,
                     int *is_equilibrated, int64_t *n_equil, int64_t *n_stats, double *mean,
                     double *calculated_precision) {
  NEED(c);
  if (c->check_pending) {
    // the verdict of a prefetched check with these very arguments: wait for it alone
    bool match = n_components == c->check_key_n && chain == c->check_key_chain && count == c->check_key_count &&
                 confidence == c->check_key_conf && quantity && abs_precision;
    for (int i = 0; match && i < n_components; ++i)
      match = quantity[i] == c->check_key_q[i] && abs_precision[i] == c->check_key_abs[i];
    CU(c, cudaEventSynchronize(c->ev_check));
    c->check_pending = false;
    if (match) {
      const SeriesCheckBlock &h = *c->h_check_out;
      for (int i = 0; i < n_components; ++i) {
        if (is_equilibrated) is_equilibrated[i] = h.is_eq[i];
        if (n_equil) n_equil[i] = h.n_eq[i];
        if (mean) mean[i] = h.out4[4 * i];
        if (calculated_precision) calculated_precision[i] = h.out4[4 * i + 3];
      }
      if (n_stats) *n_stats = h.n_stats;
      if (h.error) return device_error_check(c);  // reports and clears the sticky word
      return CMG_OK;
    }
  }

  if (chain < 0 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  if (n_components < 1 || n_components > 3 || !quantity || !abs_precision)
    return fail(c, CMG_EINVAL, "1 to 3 components");
  for (int i = 0; i < n_components; ++i)
    if (quantity[i] < 0 || quantity[i] > 2 || !(abs_precision[i] >= 0.0))
      return fail(c, CMG_EINVAL, "bad quantity or precision");
  if (count <= 0) return fail(c, CMG_EINVAL, "Error in equilibration_check: observations.size()==0");
  if (count > c->n_samples) return fail(c, CMG_EINVAL, "sample range outside the series");
  // Samples up to a mark (cmg_mark) are complete once the work before the mark is: their
  // check runs on the second stream, next to whatever the main stream has been given since.
  struct StreamView {
    cmg_context *c;
    cudaStream_t stream;
    long long n_samples;
    bool on;
    StreamView(cmg_context *ctx, bool use_aux) : c(ctx), stream(ctx->stream), n_samples(ctx->n_samples), on(use_aux) {
      if (on) {
        c->stream = c->aux;
        c->n_samples = c->mark_n_samples;
      }
    }
    ~StreamView() {
      if (on) {
        c->stream = stream;
        c->n_samples = n_samples;
      }
    }
  } view_guard(c, c->mark_valid && c->aux && count <= c->mark_n_samples && c->n_samples > c->mark_n_samples);
  if (view_guard.on) c->reserve_sm = true;
  int rc = ensure_doubles(c);
  if (rc) return rc;
  const int n = n_components;
  // one scratch block: eq jobs | stat jobs | is_eq | n_eq | n_stats | out4 | k_star
  struct Host {
    SeriesJob eq[3], st[3];
    long long n_eq[3], n_stats, k_star[3];
    double out4[12];
    int is_eq[4];
  } h;
  memset(&h, 0, sizeof h);
  for (int i = 0; i < n; ++i) {
    h.eq[i].x = series_ptr(c, chain, quantity[i]);
    h.eq[i].n = count;
  }
  Host *d = nullptr;
  CU(c, scratch_alloc(&d, sizeof(Host), c->stream));
  CU(c, cudaMemcpyAsync(d, &h, sizeof(Host), cudaMemcpyHostToDevice, c->stream));
  // The kernel takes one precision: one launch (a CTA per series, side by side) when the
  // requested precisions are equal, as they usually are; one launch per series otherwise.
  // Next to a sweep (the view above is on) the series are taken one CTA after the other, so
  // that the check never holds more than one SM: a cooperative sweep kernel needs all the
  // others at once.
  bool same = true;
  for (int i = 1; i < n; ++i) same = same && abs_precision[i] == abs_precision[0];
  const bool one_at_a_time = view_guard.on;
  if (same && !one_at_a_time) {
    k_series_equilibration<<<n, kEquilThreads, 0, c->stream>>>(d->eq, n, abs_precision[0], d->is_eq, d->n_eq);
    ++c->launches;
  } else {
    for (int j = 0; j < n; ++j) {
      k_series_equilibration<<<1, kEquilThreads, 0, c->stream>>>(d->eq + j, 1, abs_precision[j], d->is_eq + j,
                                                                 d->n_eq + j);
      ++c->launches;
    }
  }
  k_make_tail_jobs<<<1, 32, 0, c->stream>>>(d->eq, n, d->is_eq, d->n_eq, d->st, &d->n_stats);
  if (!one_at_a_time) {
    k_series_stats<<<n, 256, 0, c->stream>>>(d->st, z_confidence(confidence), d->out4, d->k_star);
  } else {
    for (int j = 0; j < n; ++j)
      k_series_stats<<<1, 256, 0, c->stream>>>(d->st + j, z_confidence(confidence), d->out4 + 4 * j, d->k_star + j);
  }
  c->launches += 2;
  CU(c, cudaGetLastError());
  CU(c, cudaMemcpyAsync(&h, d, sizeof(Host), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  scratch_free(d, c->stream);
  for (int i = 0; i < n; ++i) {
    if (is_equilibrated) is_equilibrated[i] = h.is_eq[i];
    if (n_equil) n_equil[i] = h.n_eq[i];
    if (mean) mean[i] = h.out4[4 * i];
    if (calculated_precision) calculated_precision[i] = h.out4[4 * i + 3];
  }
  if (n_stats) *n_stats = h.n_stats;
  return device_error_check(c);
}

static int device_ok(int device) {
  int ndev = 0;
  int rc = cmg_device_count(&ndev);
  if (rc) return rc;
  if (ndev == 0) return fail(nullptr, CMG_ENODEVICE, "no CUDA device");
  if (device < 0 || device >= ndev) return fail(nullptr, CMG_EINVAL, "bad device index");
  cudaSetDevice(device);
  return CMG_OK;
}

int cmg_host_series_stats(int device, const double *x, int64_t n, double confidence, double *mean,
                          double *calculated_precision, double *variance, int64_t *k_star) {
  int rc = device_ok(device);
  if (rc) return rc;
  if (!x || n <= 0)
    return fail(nullptr, CMG_EINVAL, "Error in BasicStatisticsCalculator: observations.size()==0");
  double *d = nullptr;
  cmg_context *c = nullptr;
  CU(c, scratch_alloc(&d, sizeof(double) * (size_t)n, 0));
  CU(c, cudaMemcpy(d, x, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
  std::vector<SeriesJob> jobs(1);
  jobs[0].x = d;
  jobs[0].n = n;
  rc = run_stats_jobs(nullptr, 0, jobs, confidence, mean, calculated_precision, variance, k_star);
  scratch_free(d, 0);
  return rc;
}

int cmg_host_series_equilibration(int device, const double *x, int64_t n, double abs_precision,
                                  int *is_equilibrated, int64_t *n_equil) {
  int rc = device_ok(device);
  if (rc) return rc;
  if (!x || n <= 0)
    return fail(nullptr, CMG_EINVAL, "Error in equilibration_check: observations.size()==0");
  double *d = nullptr;
  cmg_context *c = nullptr;
  CU(c, scratch_alloc(&d, sizeof(double) * (size_t)n, 0));
  CU(c, cudaMemcpy(d, x, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
  std::vector<SeriesJob> jobs(1);
  jobs[0].x = d;
  jobs[0].n = n;
  rc = run_equil_jobs(nullptr, 0, jobs, abs_precision, is_equilibrated, n_equil);
  scratch_free(d, 0);
  return rc;
}

// ---- weighted observations --------------------------------------------------------------
static int run_weighted_job(const double *x, const double *w, int64_t n, double confidence,
                            int method, double weight_sum, int64_t n_resamples, double *out5,
                            int64_t *k_star, double *resampled_out) {
  cmg_context *c = nullptr;
  double *dx = nullptr, *dw = nullptr, *deq = nullptr, *dout = nullptr;
  long long *dk = nullptr;
  WeightedJob *dj = nullptr;
  CU(c, scratch_alloc(&dx, sizeof(double) * (size_t)n, 0));
  CU(c, scratch_alloc(&dw, sizeof(double) * (size_t)n, 0));
  CU(c, scratch_alloc(&deq, sizeof(double) * (size_t)n_resamples, 0));
  CU(c, scratch_alloc(&dout, sizeof(double) * 5, 0));
  CU(c, scratch_alloc(&dk, sizeof(long long), 0));
  CU(c, scratch_alloc(&dj, sizeof(WeightedJob), 0));
  CU(c, cudaMemcpy(dx, x, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
  CU(c, cudaMemcpy(dw, w, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
  WeightedJob job{dx, dw, (long long)n, deq, (long long)n_resamples, method, weight_sum};
  CU(c, cudaMemcpy(dj, &job, sizeof(job), cudaMemcpyHostToDevice));
  k_series_stats_weighted<<<1, 256>>>(dj, z_confidence(confidence), dout, dk);
  CU(c, cudaGetLastError());
  long long hk = 0;
  if (out5) CU(c, cudaMemcpy(out5, dout, sizeof(double) * 5, cudaMemcpyDeviceToHost));
  CU(c, cudaMemcpy(&hk, dk, sizeof(long long), cudaMemcpyDeviceToHost));
  if (resampled_out)
    CU(c, cudaMemcpy(resampled_out, deq, sizeof(double) * (size_t)n_resamples,
                     cudaMemcpyDeviceToHost));
  if (k_star) *k_star = hk;
  scratch_free(dx, 0);
  scratch_free(dw, 0);
  scratch_free(deq, 0);
  scratch_free(dout, 0);
  scratch_free(dk, 0);
  scratch_free(dj, 0);
  return CMG_OK;
}

int cmg_host_series_stats_weighted(int device, const double *x, const double *w, int64_t n,
                                   double confidence, int method, int64_t n_resamples,
                                   double *mean, double *calculated_precision, double *variance,
                                   double *weight_sum, int64_t *k_star) {
  int rc = device_ok(device);
  if (rc) return rc;
  if (!x || !w || n <= 0)
    return fail(nullptr, CMG_EINVAL, "Error in BasicStatisticsCalculator: observations.size()==0");
  if (method != 1 && method != 2)
    return fail(nullptr, CMG_EINVAL, "Error in BasicStatisticsCalculator: invalid method");
  if (n_resamples <= 0) return fail(nullptr, CMG_EINVAL, "n_resamples must be positive");
  double out[5];
  rc = run_weighted_job(x, w, n, confidence, method, 0.0, n_resamples, out, k_star, nullptr);
  if (rc) return rc;
  if (mean) *mean = out[0];
  if (variance) *variance = out[1];
  if (calculated_precision) *calculated_precision = out[3];
  if (weight_sum) *weight_sum = out[4];
  return CMG_OK;
}

int cmg_host_series_resample(int device, const double *x, const double *w, int64_t n,
                             double weight_sum, int64_t n_equally_spaced, double *out) {
  int rc = device_ok(device);
  if (rc) return rc;
  if (!x || !w || !out || n <= 0 || n_equally_spaced <= 0)
    return fail(nullptr, CMG_EINVAL, "bad argument");
  return run_weighted_job(x, w, n, 0.95, 0, weight_sum, n_equally_spaced, nullptr, nullptr, out);
}

int cmg_host_series_equilibration_weighted(int device, const double *x, const double *w,
                                           int64_t n, double abs_precision,
                                           int *is_equilibrated, int64_t *n_equil) {
  int rc = device_ok(device);
  if (rc) return rc;
  if (!x || !w || n <= 0)
    return fail(nullptr, CMG_EINVAL, "Error in equilibration_check: observations.size()==0");
  cmg_context *c = nullptr;
  double *dx = nullptr, *dw = nullptr, *dy = nullptr, *df = nullptr;
  CU(c, scratch_alloc(&dx, sizeof(double) * (size_t)n, 0));
  CU(c, scratch_alloc(&dw, sizeof(double) * (size_t)n, 0));
  CU(c, scratch_alloc(&dy, sizeof(double) * (size_t)n, 0));
  CU(c, scratch_alloc(&df, sizeof(double), 0));
  CU(c, cudaMemcpy(dx, x, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
  CU(c, cudaMemcpy(dw, w, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
  k_weight_factor<<<1, 32>>>(dw, n, df);
  k_apply_weight_factor<<<nblocks(n, 256), 256>>>(dx, dw, n, df, dy);
  CU(c, cudaGetLastError());
  std::vector<SeriesJob> jobs(1);
  jobs[0].x = dy;
  jobs[0].n = n;
  rc = run_equil_jobs(nullptr, 0, jobs, abs_precision, is_equilibrated, n_equil);
  scratch_free(dx, 0);
  scratch_free(dw, 0);
  scratch_free(dy, 0);
  scratch_free(df, 0);
  return rc;
}

// ---- conversions ----------------------------------------------------------------------
int cmg_conv_l_to_bijk(int device, const int64_t *n3, int64_t n_basis, const int64_t *l,
                       int64_t count, int64_t *bijk_out) {
  int rc = device_ok(device);
  if (rc) return rc;
  if (!n3 || !l || !bijk_out || count < 0 || n_basis < 1)
    return fail(nullptr, CMG_EINVAL, "bad argument");
  if (count == 0) return CMG_OK;
  const long long total = n3[0] * n3[1] * n3[2] * n_basis;
  for (int64_t i = 0; i < count; ++i)
    if (l[i] < 0 || l[i] >= total) return fail(nullptr, CMG_EINVAL, "linear index out of range");
  cmg_context *c = nullptr;
  long long *dl = nullptr, *db = nullptr;
  CU(c, cudaMalloc(&dl, 8 * (size_t)count));
  CU(c, cudaMalloc(&db, 32 * (size_t)count));
  CU(c, cudaMemcpy(dl, l, 8 * (size_t)count, cudaMemcpyHostToDevice));
  k_conv_l_to_bijk<<<nblocks(count, 256), 256>>>(n3[0], n3[1], n3[2], dl, count, db);
  CU(c, cudaGetLastError());
  CU(c, cudaMemcpy(bijk_out, db, 32 * (size_t)count, cudaMemcpyDeviceToHost));
  cudaFree(dl);
  cudaFree(db);
  return CMG_OK;
}

int cmg_conv_bijk_to_l(int device, const int64_t *n3, int64_t n_basis, const int64_t *bijk,
                       int64_t count, int64_t *l_out) {
  int rc = device_ok(device);
  if (rc) return rc;
  if (!n3 || !bijk || !l_out || count < 0 || n_basis < 1)
    return fail(nullptr, CMG_EINVAL, "bad argument");
  if (count == 0) return CMG_OK;
  for (int64_t i = 0; i < count; ++i)
    if (bijk[4 * i] < 0 || bijk[4 * i] >= n_basis)
      return fail(nullptr, CMG_EINVAL, "sublattice index out of range");
  cmg_context *c = nullptr;
  long long *dl = nullptr, *db = nullptr;
  CU(c, cudaMalloc(&dl, 8 * (size_t)count));
  CU(c, cudaMalloc(&db, 32 * (size_t)count));
  CU(c, cudaMemcpy(db, bijk, 32 * (size_t)count, cudaMemcpyHostToDevice));
  k_conv_bijk_to_l<<<nblocks(count, 256), 256>>>(n3[0], n3[1], n3[2], db, count, dl);
  CU(c, cudaGetLastError());
  CU(c, cudaMemcpy(l_out, dl, 8 * (size_t)count, cudaMemcpyDeviceToHost));
  cudaFree(dl);
  cudaFree(db);
  return CMG_OK;
}

static int make_conv_general(const int64_t *T9, int64_t n_basis, ConvGeneral *P) {
  if (!T9 || n_basis < 1) return fail(nullptr, CMG_EINVAL, "bad argument");
  try {
    casm_monte_b200::SiteIndexConverter f(casm_monte_b200::Mat3l::from_row_major(T9), n_basis);
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        P->T[3 * r + c] = f.T.a[r][c];
        P->adjT[3 * r + c] = f.adjT.a[r][c];
        P->U[3 * r + c] = f.U.a[r][c];
        P->Uinv[3 * r + c] = f.Uinv.a[r][c];
      }
    P->detT = f.detT;
    for (int d = 0; d < 3; ++d) P->s[d] = f.s[d];
    P->n_unitcells = f.n_unitcells;
  } catch (std::exception const &e) {
    return fail(nullptr, CMG_EINVAL, e.what());
  }
  return CMG_OK;
}

int cmg_conv_general_l_to_bijk(int device, const int64_t *T9, int64_t n_basis, const int64_t *l,
                               int64_t count, int64_t *bijk_out) {
  int rc = device_ok(device);
  if (rc) return rc;
  ConvGeneral P;
  rc = make_conv_general(T9, n_basis, &P);
  if (rc) return rc;
  if (!l || !bijk_out || count < 0) return fail(nullptr, CMG_EINVAL, "bad argument");
  if (count == 0) return CMG_OK;
  const long long total = P.n_unitcells * n_basis;
  for (int64_t i = 0; i < count; ++i)
    if (l[i] < 0 || l[i] >= total) return fail(nullptr, CMG_EINVAL, "linear index out of range");
  cmg_context *c = nullptr;
  long long *dl = nullptr, *db = nullptr;
  CU(c, scratch_alloc(&dl, 8 * (size_t)count, 0));
  CU(c, scratch_alloc(&db, 32 * (size_t)count, 0));
  CU(c, cudaMemcpy(dl, l, 8 * (size_t)count, cudaMemcpyHostToDevice));
  k_conv_general_l_to_bijk<<<nblocks(count, 256), 256>>>(P, dl, count, db);
  CU(c, cudaGetLastError());
  CU(c, cudaMemcpy(bijk_out, db, 32 * (size_t)count, cudaMemcpyDeviceToHost));
  scratch_free(dl, 0);
  scratch_free(db, 0);
  return CMG_OK;
}

int cmg_conv_general_bijk_to_l(int device, const int64_t *T9, int64_t n_basis, const int64_t *bijk,
                               int64_t count, int64_t *l_out) {
  int rc = device_ok(device);
  if (rc) return rc;
  ConvGeneral P;
  rc = make_conv_general(T9, n_basis, &P);
  if (rc) return rc;
  if (!bijk || !l_out || count < 0) return fail(nullptr, CMG_EINVAL, "bad argument");
  if (count == 0) return CMG_OK;
  for (int64_t i = 0; i < count; ++i)
    if (bijk[4 * i] < 0 || bijk[4 * i] >= n_basis)
      return fail(nullptr, CMG_EINVAL, "sublattice index out of range");
  cmg_context *c = nullptr;
  long long *dl = nullptr, *db = nullptr;
  CU(c, scratch_alloc(&dl, 8 * (size_t)count, 0));
  CU(c, scratch_alloc(&db, 32 * (size_t)count, 0));
  CU(c, cudaMemcpy(db, bijk, 32 * (size_t)count, cudaMemcpyHostToDevice));
  k_conv_general_bijk_to_l<<<nblocks(count, 256), 256>>>(P, db, count, dl);
  CU(c, cudaGetLastError());
  CU(c, cudaMemcpy(l_out, dl, 8 * (size_t)count, cudaMemcpyDeviceToHost));
  scratch_free(dl, 0);
  scratch_free(db, 0);
  return CMG_OK;
}

// ---- k-state model ------------------------------------------------------------------------
// Tables with the expression order of the restated model: dE = sum_s n_s * (V[to][s] -
// V[from][s]) accumulated in species order, dPhi = dE - (mu[to] - mu[from]),
// prob = exp(-dPhi * beta), beta = 1 / (KB * T); thresholds as for the Ising tables.
static void build_kstate_tables(KStateTables &t, int dim, int K, const double *V, double T, const double *mu) {
  memset(&t, 0, sizeof t);
  t.K = K;
  t.z = 2 * dim;
  t.n_cfg = 1;
  for (int s = 1; s < K; ++s) t.n_cfg *= (t.z + 1);
  volatile double kt = CMG_KB * T;
  const double beta = 1.0 / kt;
  int cnt[kMaxSpecies];
  for (int cfg = 0; cfg < t.n_cfg; ++cfg) {
    int rest = cfg, total = 0;
    for (int s = 1; s < K; ++s) {
      cnt[s] = rest % (t.z + 1);
      rest /= (t.z + 1);
      total += cnt[s];
    }
    if (total > t.z) continue;
    cnt[0] = t.z - total;
    for (int from = 0; from < K; ++from)
      for (int to = 0; to < K; ++to) {
        if (from == to) continue;
        volatile double dE = 0.0;
        for (int s = 0; s < K; ++s) {
          volatile double dv = V[to * K + s] - V[from * K + s];
          volatile double term = cnt[s] * dv;
          dE = dE + term;
        }
        volatile double dmu = mu[to] - mu[from];
        const double d = dE - dmu;
        volatile double arg = -d * beta;
        const double p = std::exp(arg);
        const int i = (from * K + to) * t.n_cfg + cfg;
        t.dPhi[i] = d;
        t.prob[i] = p;
        if (d < 0.0 || p >= 1.0) {
          t.thr_m1[i] = 0xFFFFFFFFu;
        } else if (!(p > 0.0)) {
          t.thr_m1[i] = 0u;
          t.never[i] = 1;
        } else {
          double scaled = std::ceil(p * 4294967296.0);
          if (scaled < 1.0) scaled = 1.0;
          if (scaled > 4294967296.0) scaled = 4294967296.0;
          t.thr_m1[i] = (uint32_t)((unsigned long long)scaled - 1ull);
        }
      }
  }
  t.valid = 1;
}

int cmg_kstate_set_model(cmg_context *c, int n_species, const double *V) {
  NEED(c);
  if (n_species < 2 || n_species > kMaxSpecies || !V) return fail(c, CMG_EINVAL, "k-state model: 2 <= n_species <= 4");
  if (c->slab) return fail(c, CMG_EUNSUPPORTED, "k-state model: single-GPU contexts only");
  for (int a = 0; a < n_species; ++a)
    for (int b = 0; b < n_species; ++b)
      if (V[a * n_species + b] != V[b * n_species + a]) return fail(c, CMG_EINVAL, "k-state model: V must be symmetric");
  c->ks_K = n_species;
  for (int i = 0; i < n_species * n_species; ++i) c->ks_V[i] = V[i];
  if (!c->d_ktabs) CU(c, cudaMalloc(&c->d_ktabs, sizeof(KStateTables) * (size_t)c->n_chains));
  c->ks_valid.assign(c->n_chains, 0);
  c->ks_n_samples = 0;
  c->ks_loc_valid = false;
  return CMG_OK;
}

int cmg_kstate_set_conditions(cmg_context *c, int chain, double temperature, const double *mu) {
  NEED(c);
  if (!c->ks_K) return fail(c, CMG_ESTATE, "k-state model not set");
  if (chain < -1 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  if (!(temperature > 0.0) || !mu) return fail(c, CMG_EINVAL, "temperature must be > 0");
  const int lo = chain < 0 ? 0 : chain, hi = chain < 0 ? c->n_chains : chain + 1;
  std::vector<KStateTables> h(1);
  build_kstate_tables(h[0], c->dim, c->ks_K, c->ks_V, temperature, mu);
  for (int ch = lo; ch < hi; ++ch) {
    CU(c, cudaMemcpyAsync(c->d_ktabs + ch, h.data(), sizeof(KStateTables), cudaMemcpyHostToDevice, c->stream));
    c->ks_valid[ch] = 1;
  }
  CU(c, cudaStreamSynchronize(c->stream));
  return CMG_OK;
}

int cmg_kstate_get_tables(cmg_context *c, int chain, double *dPhi, double *prob, uint32_t *thr_m1, uint8_t *never,
                          int64_t n_entries) {
  NEED(c);
  if (!c->ks_K) return fail(c, CMG_ESTATE, "k-state model not set");
  if (chain < 0 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  if (!c->ks_valid[chain]) return fail(c, CMG_ESTATE, "conditions not set");
  std::vector<KStateTables> h(1);
  CU(c, cudaMemcpyAsync(h.data(), c->d_ktabs + chain, sizeof(KStateTables), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  const int64_t n = (int64_t)h[0].K * h[0].K * h[0].n_cfg;
  if (n_entries != n) return fail(c, CMG_EINVAL, "n_entries must be K * K * (z + 1)^(K - 1)");
  for (int64_t i = 0; i < n; ++i) {
    if (dPhi) dPhi[i] = h[0].dPhi[i];
    if (prob) prob[i] = h[0].prob[i];
    if (thr_m1) thr_m1[i] = h[0].thr_m1[i];
    if (never) never[i] = h[0].never[i];
  }
  return CMG_OK;
}

int cmg_kstate_upload_occupation_i32(cmg_context *c, int chain, const int32_t *occ_index, int64_t n) {
  NEED(c);
  if (!c->ks_K) return fail(c, CMG_ESTATE, "k-state model not set");
  if (chain < 0 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  if (!occ_index || n != c->n_sites) return fail(c, CMG_EINVAL, "Error in set_occupation: size mismatch");
  int rc = ensure_stage(c);
  if (rc) return rc;
  CU(c, cudaMemcpyAsync(c->d_stage, occ_index, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaMemsetAsync(c->d_flag, 0, sizeof(int), c->stream));
  k_kstate_i32_to_sites<<<nblocks(n, 256), 256, 0, c->stream>>>(c->d_stage, chain_base(c, chain), c->plane_stride,
                                                                 nat_shape(c), c->planar ? 1 : 0, c->ks_K, c->d_flag);
  ++c->launches;
  if (c->planar) c->nat_is_current = false;
  c->ks_loc_valid = false;
  CU(c, cudaGetLastError());
  int bad = 0;
  CU(c, cudaMemcpyAsync(&bad, c->d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  if (bad) return fail(c, CMG_EINVAL, "occupation indices must be in [0, n_species)");
  return CMG_OK;
}

int cmg_kstate_download_occupation_i32(cmg_context *c, int chain, int32_t *occ_index, int64_t n) {
  NEED(c);
  if (!c->ks_K) return fail(c, CMG_ESTATE, "k-state model not set");
  if (chain < 0 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  if (!occ_index || n != c->n_sites) return fail(c, CMG_EINVAL, "size mismatch");
  int rc = ensure_stage(c);
  if (rc) return rc;
  k_kstate_sites_to_i32<<<nblocks(n, 256), 256, 0, c->stream>>>(chain_base(c, chain), c->plane_stride, c->d_stage,
                                                                 nat_shape(c), c->planar ? 1 : 0);
  ++c->launches;
  CU(c, cudaGetLastError());
  CU(c, cudaMemcpyAsync(occ_index, c->d_stage, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return CMG_OK;
}

static int kstate_sample(cmg_context *c) {
  const int per = c->ks_K + c->ks_K * c->ks_K;
  if (c->ks_n_samples >= c->ks_capacity) {
    long long cap = c->ks_capacity ? 2 * c->ks_capacity : 1024;
    long long *nb = nullptr;
    CU(c, cudaMalloc(&nb, sizeof(long long) * (size_t)cap * c->n_chains * per));
    CU(c, cudaMemsetAsync(nb, 0, sizeof(long long) * (size_t)cap * c->n_chains * per, c->stream));
    if (c->d_kseries && c->ks_n_samples > 0)
      CU(c, cudaMemcpyAsync(nb, c->d_kseries, sizeof(long long) * (size_t)c->ks_n_samples * c->n_chains * per,
                            cudaMemcpyDeviceToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->d_kseries);
    c->d_kseries = nb;
    c->ks_capacity = cap;
  }
  int rc = sync_nat_from_planes(c);
  if (rc) return rc;
  for (int ch = 0; ch < c->n_chains; ++ch) {
    long long *dst = c->d_kseries + ((size_t)c->ks_n_samples * c->n_chains + ch) * per;
    CU(c, cudaMemsetAsync(dst, 0, sizeof(long long) * per, c->stream));
    k_kstate_observables<<<(unsigned)std::min<long long>(nblocks(c->n_sites, 256), 148 * 4), 256, 0, c->stream>>>(
        c->d_nat + (size_t)ch * c->n_sites, nat_shape(c), c->ks_K, dst);
    ++c->launches;
  }
  CU(c, cudaGetLastError());
  ++c->ks_n_samples;
  return CMG_OK;
}

int cmg_kstate_run_passes(cmg_context *c, int64_t n_passes, int mode, int64_t sample_period) {
  NEED(c);
  if (!c->ks_K) return fail(c, CMG_ESTATE, "k-state model not set");
  if (n_passes < 0 || sample_period < 0) return fail(c, CMG_EINVAL, "negative count");
  for (int ch = 0; ch < c->n_chains; ++ch)
    if (!c->ks_valid[ch]) return fail(c, CMG_ESTATE, "conditions not set for every chain");
  if (mode == CMG_MODE_CHECKERBOARD) {
    if (!c->planar) return fail(c, CMG_EINVAL, "checkerboard mode needs even extents; use CMG_MODE_SERIAL_REFERENCE");
    KSweepArgs A;
    memset(&A, 0, sizeof A);
    A.L = view(c);
    A.tabs = c->d_ktabs;
    A.n_accept = c->d_n_accept;
    for (int r = 0; r < 10; ++r) {
      A.rk[2 * r] = (uint32_t)c->philox_seed + (uint32_t)r * kPhiloxW0;
      A.rk[2 * r + 1] = (uint32_t)(c->philox_seed >> 32) + (uint32_t)r * kPhiloxW1;
    }
    A.chain_offset = c->chain_offset;
    const dim3 grid((unsigned)nblocks((c->shape[0] / 2 + 1) / 2, 32), (unsigned)nblocks(c->shape[1], 8 * kKStateTrips), (unsigned)(c->shape[2] * c->n_chains));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = grid;
    cfg.blockDim = dim3(32, 8);
    cfg.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = c->pdl ? 1 : 0;
    for (int64_t t = 0; t < n_passes; ++t) {
      A.pass = c->h_pass;
      for (int colour = 0; colour < 2; ++colour) {
        A.colour = colour;
        void *args[] = {&A};
        CU(c, cudaLaunchKernelExC(&cfg, (const void *)k_kstate_halfsweep, args));
        ++c->launches;
      }
      c->nat_is_current = false;
      c->ks_loc_valid = false;
      ++c->h_pass;
      ++c->n_pass;
      if (sample_period > 0 && (c->n_pass % sample_period) == 0) {
        int rc = kstate_sample(c);
        if (rc) return rc;
      }
    }
    CU(c, cudaGetLastError());
    c->variant_name = "kstate_generic";
    return CMG_OK;
  }
  if (mode != CMG_MODE_SERIAL_REFERENCE) return fail(c, CMG_EINVAL, "unknown mode");
  for (int ch = 0; ch < c->n_chains; ++ch)
    if (!c->d_engines || !c->engine_seeded[ch]) return fail(c, CMG_ESTATE, "mt19937_64 engine not seeded for every chain");
  if (c->n_sites >= (1ll << 31)) return fail(c, CMG_EUNSUPPORTED, "k-state serial mode: n_sites < 2^31");
  int rc = sync_nat_from_planes(c);
  if (rc) return rc;
  if (!c->d_kloc) {
    CU(c, cudaMalloc(&c->d_kloc, sizeof(int) * (size_t)c->n_chains * kMaxSpecies * c->n_sites));
    CU(c, cudaMalloc(&c->d_kloc_size, sizeof(int) * (size_t)c->n_chains * kMaxSpecies));
    CU(c, cudaMalloc(&c->d_kmol_loc, sizeof(int) * (size_t)c->n_chains * c->n_sites));
    c->ks_loc_valid = false;
  }
  KLocation P;
  if (!c->ks_loc_valid) {
    // OccLocation::initialize from the current occupation
    for (int ch = 0; ch < c->n_chains; ++ch) {
      P.loc = c->d_kloc + (size_t)ch * c->ks_K * c->n_sites;
      P.loc_size = c->d_kloc_size + (size_t)ch * kMaxSpecies;
      P.mol_loc = c->d_kmol_loc + (size_t)ch * c->n_sites;
      k_kstate_location_init<<<1, 1, 0, c->stream>>>(c->d_nat + (size_t)ch * c->n_sites, c->n_sites, c->ks_K, P);
      ++c->launches;
    }
    c->ks_loc_valid = true;
  }
  KSerialArgs A;
  memset(&A, 0, sizeof A);
  A.nat = c->d_nat;
  A.shape = nat_shape(c);
  A.tabs = c->d_ktabs;
  A.engines = c->d_engines;
  A.n_accept = c->d_n_accept;
  A.loc.loc = c->d_kloc;
  A.loc.loc_size = c->d_kloc_size;
  A.loc.mol_loc = c->d_kmol_loc;
  int64_t left = n_passes;
  while (left > 0) {
    int64_t chunk = left;
    if (sample_period > 0) chunk = std::min<int64_t>(left, sample_period - (c->n_pass % sample_period));
    A.n_passes = chunk;
    k_kstate_serial<<<c->n_chains, kSerialThreads, 0, c->stream>>>(A);
    ++c->launches;
    CU(c, cudaGetLastError());
    c->n_pass += chunk;
    left -= chunk;
    // the planes follow the natural copy (which stays current: it is what the walk updates)
    rc = sync_planes_from_nat(c);
    if (rc) return rc;
    c->nat_is_current = true;
    if (sample_period > 0 && (c->n_pass % sample_period) == 0) {
      rc = kstate_sample(c);
      if (rc) return rc;
    }
  }
  c->variant_name = "kstate_serial_reference";
  return CMG_OK;
}

int cmg_kstate_n_samples(cmg_context *c, int64_t *n_samples) {
  if (!c || !n_samples) return fail(c, CMG_EINVAL, "null argument");
  *n_samples = c->ks_n_samples;
  return CMG_OK;
}

int cmg_kstate_clear_samples(cmg_context *c) {
  NEED(c);
  c->ks_n_samples = 0;
  return CMG_OK;
}

int cmg_kstate_read_samples(cmg_context *c, int chain, int64_t first, int64_t count, int64_t *counts,
                            int64_t *bonds) {
  NEED(c);
  if (!c->ks_K) return fail(c, CMG_ESTATE, "k-state model not set");
  if (chain < 0 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  if (first < 0 || count < 0 || first + count > c->ks_n_samples) return fail(c, CMG_EINVAL, "sample range outside the series");
  if (count == 0) return CMG_OK;
  const int K = c->ks_K, per = K + K * K;
  std::vector<long long> tmp((size_t)count * per);
  CU(c, cudaMemcpy2DAsync(tmp.data(), sizeof(long long) * per,
                          c->d_kseries + ((size_t)first * c->n_chains + chain) * per,
                          sizeof(long long) * per * c->n_chains, sizeof(long long) * per, (size_t)count,
                          cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  for (int64_t i = 0; i < count; ++i) {
    for (int a = 0; a < K; ++a)
      if (counts) counts[i * K + a] = tmp[(size_t)i * per + a];
    for (int e = 0; e < K * K; ++e)
      if (bonds) bonds[i * K * K + e] = tmp[(size_t)i * per + K + e];
  }
  return CMG_OK;
}

// ---- N-fold way driver ------------------------------------------------------------------
int cmg_nfold_run(cmg_context *c, int64_t n_steps, int64_t sample_period_steps) {
  NEED(c);
  if (n_steps < 0 || sample_period_steps < 0) return fail(c, CMG_EINVAL, "negative count");
  if (c->slab || c->ks_K) return fail(c, CMG_EUNSUPPORTED, "nfold: Ising contexts on one GPU");
  if (c->n_sites >= (1ll << 31)) return fail(c, CMG_EUNSUPPORTED, "nfold: n_sites < 2^31");
  int rc = check_ready(c);
  if (rc) return rc;
  for (int ch = 0; ch < c->n_chains; ++ch)
    if (!c->d_engines || !c->engine_seeded[ch]) return fail(c, CMG_ESTATE, "mt19937_64 engine not seeded for every chain");
  if (n_steps == 0) return CMG_OK;
  rc = sync_nat_from_planes(c);
  if (rc) return rc;
  if (!c->d_nf_members) {
    CU(c, cudaMalloc(&c->d_nf_members, sizeof(int) * (size_t)c->n_chains * 16 * c->n_sites));
    CU(c, cudaMalloc(&c->d_nf_n, sizeof(int) * (size_t)c->n_chains * 16));
    CU(c, cudaMalloc(&c->d_nf_pos, sizeof(int) * (size_t)c->n_chains * c->n_sites));
    CU(c, cudaMalloc(&c->d_nf_class, (size_t)c->n_chains * c->n_sites));
    CU(c, cudaMalloc(&c->d_nf_time, sizeof(double) * c->n_chains));
    CU(c, cudaMemsetAsync(c->d_nf_time, 0, sizeof(double) * c->n_chains, c->stream));
    c->nf_lists_valid = false;
  }
  NfoldArgs A;
  memset(&A, 0, sizeof A);
  A.lists.members = c->d_nf_members;
  A.lists.n_members = c->d_nf_n;
  A.lists.site_pos = c->d_nf_pos;
  A.lists.site_class = c->d_nf_class;
  // the lists follow the occupation as long as only this driver changes it
  if (!c->nf_lists_valid || !c->nat_is_current) c->nf_lists_valid = false;
  for (int ch = 0; ch < c->n_chains; ++ch) {
    CU(c, cudaMemsetAsync(c->d_cur_sb + 2 * ch, 0, 2 * sizeof(long long), c->stream));
    k_observables_natural<<<(unsigned)std::min<long long>(nblocks(c->n_sites, 256), 148 * 8), 256, 0, c->stream>>>(
        c->d_nat + (size_t)ch * c->n_sites, nat_shape(c), c->d_cur_sb + 2 * ch);
    ++c->launches;
    if (!c->nf_lists_valid) {
      NfoldLists P;
      P.members = c->d_nf_members + (size_t)ch * 16 * c->n_sites;
      P.n_members = c->d_nf_n + ch * 16;
      P.site_pos = c->d_nf_pos + (size_t)ch * c->n_sites;
      P.site_class = c->d_nf_class + (size_t)ch * c->n_sites;
      k_nfold_init<<<1, 1, 0, c->stream>>>(c->d_nat + (size_t)ch * c->n_sites, nat_shape(c), P);
      ++c->launches;
    }
  }
  c->nf_lists_valid = true;
  long long n_new = 0;
  if (sample_period_steps > 0) n_new = (c->nf_steps + n_steps) / sample_period_steps - c->nf_steps / sample_period_steps;
  rc = ensure_series(c, c->n_samples + n_new);
  if (rc) return rc;
  if (c->nf_first_sample < 0) c->nf_first_sample = c->n_samples;
  const long long w_need = c->n_samples - c->nf_first_sample + n_new;
  if (w_need > c->nf_capacity) {
    long long cap = c->nf_capacity ? c->nf_capacity : 1024;
    while (cap < w_need) cap *= 2;
    double *nw = nullptr, *nr = nullptr;
    CU(c, cudaMalloc(&nw, sizeof(double) * (size_t)cap * c->n_chains));
    CU(c, cudaMalloc(&nr, sizeof(double) * (size_t)cap * c->n_chains));
    const size_t used = sizeof(double) * (size_t)(c->n_samples - c->nf_first_sample) * c->n_chains;
    if (used) {
      CU(c, cudaMemcpyAsync(nw, c->d_nf_weight, used, cudaMemcpyDeviceToDevice, c->stream));
      CU(c, cudaMemcpyAsync(nr, c->d_nf_ratio, used, cudaMemcpyDeviceToDevice, c->stream));
    }
    CU(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->d_nf_weight);
    cudaFree(c->d_nf_ratio);
    c->d_nf_weight = nw;
    c->d_nf_ratio = nr;
    c->nf_capacity = cap;
  }
  A.nat = c->d_nat;
  A.shape = nat_shape(c);
  A.tabs = c->d_tabs;
  A.engines = c->d_engines;
  A.cur_sb = c->d_cur_sb;
  A.series = c->d_series + c->n_samples * 2 * c->n_chains;
  A.series_slot_stride = 2 * c->n_chains;
  A.weight = c->d_nf_weight + (size_t)(c->n_samples - c->nf_first_sample) * c->n_chains;
  A.rate_ratio = c->d_nf_ratio + (size_t)(c->n_samples - c->nf_first_sample) * c->n_chains;
  A.wstride = c->n_chains;
  A.time = c->d_nf_time;
  A.n_steps = n_steps;
  A.sample_period = sample_period_steps;
  A.step_base = c->nf_steps;
  k_nfold<<<c->n_chains, kSerialThreads, 0, c->stream>>>(A);
  ++c->launches;
  CU(c, cudaGetLastError());
  c->nf_steps += n_steps;
  c->n_samples += n_new;
  rc = sync_planes_from_nat(c);
  if (rc) return rc;
  c->nat_is_current = true;
  c->variant_name = "nfold";
  return CMG_OK;
}

int cmg_nfold_read_weights(cmg_context *c, int chain, int64_t first, int64_t count, double *weight,
                           double *expected_acceptance_rate) {
  NEED(c);
  if (chain < 0 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  if (c->nf_first_sample < 0) return fail(c, CMG_ESTATE, "no nfold samples");
  if (first < c->nf_first_sample || count < 0 || first + count > c->n_samples)
    return fail(c, CMG_EINVAL, "sample range outside the nfold samples");
  if (count == 0) return CMG_OK;
  const size_t off = (size_t)(first - c->nf_first_sample) * c->n_chains + chain;
  if (weight)
    CU(c, cudaMemcpy2DAsync(weight, sizeof(double), c->d_nf_weight + off, sizeof(double) * c->n_chains, sizeof(double),
                            (size_t)count, cudaMemcpyDeviceToHost, c->stream));
  if (expected_acceptance_rate)
    CU(c, cudaMemcpy2DAsync(expected_acceptance_rate, sizeof(double), c->d_nf_ratio + off,
                            sizeof(double) * c->n_chains, sizeof(double), (size_t)count, cudaMemcpyDeviceToHost,
                            c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return CMG_OK;
}

int cmg_nfold_time(cmg_context *c, int chain, double *time, int64_t *n_steps) {
  NEED(c);
  if (chain < 0 || chain >= c->n_chains) return fail(c, CMG_EINVAL, "bad chain index");
  if (time) {
    *time = 0.0;
    if (c->d_nf_time) {
      CU(c, cudaMemcpyAsync(time, c->d_nf_time + chain, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      CU(c, cudaStreamSynchronize(c->stream));
    }
  }
  if (n_steps) *n_steps = c->nf_steps;
  return CMG_OK;
}

// ---- introspection ----------------------------------------------------------------------
int cmg_launch_count(const cmg_context *c, int64_t *n) {
  if (!c || !n) return fail(nullptr, CMG_EINVAL, "null argument");
  *n = c->launches;
  return CMG_OK;
}

const char *cmg_kernel_variant(const cmg_context *c) { return c ? c->variant_name.c_str() : ""; }

int cmg_set_energy_form(cmg_context *c, int use_nlist) {
  NEED(c);
  if (!use_nlist) {
    if (c->dim != 2 || !c->planar || c->slab || c->shape[0] % 32 != 0)
      return fail(c, CMG_EUNSUPPORTED,
                  "the use_nlist=false energy form is sampled on the device for 2-d lattices with "
                  "even extents and n0 % 32 == 0");
  }
  if (c->nonlist != (use_nlist == 0) && c->n_samples > 0)
    return fail(c, CMG_ESTATE, "change the energy form on an empty sample series (cmg_clear_samples)");
  c->nonlist = (use_nlist == 0);
  return CMG_OK;
}

int cmg_set_kernel_variant(cmg_context *c, const char *name) {
  NEED(c);
  if (!name) return fail(c, CMG_EINVAL, "null name");
  std::string s(name);
  // "bulk2d:js=8" sets the strip length; "tile2d:p=2:nt=512" passes per launch / CTA size
  c->js = 0;
  c->tile_threads = 512;
  c->tile_threads_forced = false;
  size_t p = s.find(":js=");
  if (p != std::string::npos) {
    c->js = atoi(s.c_str() + p + 4);
    if (c->js < 1) return fail(c, CMG_EINVAL, "bad js");
  }
  p = s.find(":p=");
  if (p != std::string::npos) {
    c->tile_passes = atoi(s.c_str() + p + 3);
    if (c->tile_passes < 1 || c->tile_passes > 16) return fail(c, CMG_EINVAL, "bad p");
  }
  c->n_strips2d = 0;
  c->js_auto[5] = 0;
  p = s.find(":rt=");
  if (p != std::string::npos) {
    if (c->ring_peer_mb[0] || c->ring_peer_mb[1]) return fail(c, CMG_ESTATE, "rt cannot change once ring peers are attached");
    c->ring_tiles_cap = std::max(0, atoi(s.c_str() + p + 4));
  }
  c->pdl = s.find(":pdl=0") == std::string::npos;
  c->hs_chain = s.find(":chain=0") == std::string::npos;
  p = s.find(":ns=");
  if (p != std::string::npos) c->n_strips2d = std::max(0, atoi(s.c_str() + p + 4));
  p = s.find(":rp=");
  if (p != std::string::npos) {
    c->ring_passes = atoi(s.c_str() + p + 4);
    if (c->ring_passes < 1 || c->ring_passes > kRingMaxPasses)
      return fail(c, CMG_EINVAL, "rp must be in [1, 256]");
  }
  p = s.find(":nt=");
  if (p != std::string::npos) {
    c->tile_threads = atoi(s.c_str() + p + 4);
    c->tile_threads_forced = true;
    if (c->tile_threads != 256 && c->tile_threads != 512 && c->tile_threads != 640 &&
        c->tile_threads != 768 && c->tile_threads != 1024)
      return fail(c, CMG_EINVAL, "nt must be 256, 512, 640, 768 or 1024");
  }
  p = s.find(':');
  if (p != std::string::npos) s = s.substr(0, p);
  if (s == "auto") c->forced_variant = V_AUTO;
  else if (s == "generic") c->forced_variant = V_GENERIC;
  else if (s == "bulk2d") c->forced_variant = V_BULK2D;
  else if (s == "bulk3d") c->forced_variant = V_BULK3D;
  else if (s == "tma3d") c->forced_variant = V_TMA3D;
  else if (s == "tile2d") c->forced_variant = V_TILE2D;
  else if (s == "ring2d") c->forced_variant = V_RING2D;
  else return fail(c, CMG_EINVAL, "unknown kernel variant");
  return CMG_OK;
}

}  // extern "C"
