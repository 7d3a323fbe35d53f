// cabi.cpp -- the extern "C" boundary (include/eigkl.h).  No exception crosses it.
#include "internal.h"
#include <algorithm>
#include <cstdlib>
#include <mutex>

using namespace eigkl;

namespace eigkl {

void KernelProfiler::begin(int c, cudaStream_t s, int w) {
  if (!on || suppress) return;
  if (used + 2 > ev.size()) {
    const size_t old = ev.size();
    ev.resize(old + 4096);
    for (size_t i = old; i < ev.size(); ++i) cudaEventCreate(&ev[i]);
  }
  cls.push_back(c);
  weight.push_back(w);
  cudaEventRecord(ev[used], s);
}
void KernelProfiler::end(cudaStream_t s) {
  if (!on || suppress) return;
  cudaEventRecord(ev[used + 1], s);
  used += 2;
}
void KernelProfiler::resolve() {
  if (!on || used == 0) return;
  cudaEventSynchronize(ev[used - 1]);
  for (size_t i = 0; i < used; i += 2) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, ev[i], ev[i + 1]) == cudaSuccess) { ms[cls[i / 2]] += t; cnt[cls[i / 2]] += weight[i / 2]; }
  }
  used = 0;
  cls.clear();
  weight.clear();
}
void KernelProfiler::reset() {
  resolve();
  for (int i = 0; i < 8; ++i) { ms[i] = 0; cnt[i] = 0; }
}
KernelProfiler::~KernelProfiler() {
  for (auto e : ev) cudaEventDestroy(e);
}

}  // namespace eigkl

static std::string g_create_error;
static std::mutex g_create_mutex;

template <typename F>
static int guarded(eigkl_handle *h, F &&f) {
  try {
    if (!h) return EIGKL_E_ARG;
    f();
    return EIGKL_OK;
  } catch (const eigkl::Error &e) {
    h->err = e.what();
    cudaGetLastError();          // clear a sticky-less error state
    return e.code;
  } catch (const std::bad_alloc &) {
    h->err = "out of host memory";
    return EIGKL_E_NOMEM;
  } catch (const std::exception &e) {
    h->err = e.what();
    return EIGKL_E_ARG;
  } catch (...) {
    h->err = "unknown error";
    return EIGKL_E_ARG;
  }
}

static void sync_profile_into_stats(eigkl_handle *h) {
  auto &p = h->prof;
  p.resolve();
  auto &s = h->stats;
  s.ms_spmv = p.ms[KC_SPMV]; s.n_spmv = p.cnt[KC_SPMV];
  s.ms_multidot = p.ms[KC_MULTIDOT]; s.n_multidot = p.cnt[KC_MULTIDOT];
  s.ms_update = p.ms[KC_UPDATE]; s.n_update = p.cnt[KC_UPDATE];
  s.ms_restart = p.ms[KC_RESTART]; s.n_restart = p.cnt[KC_RESTART];
  s.ms_dvalues = p.ms[KC_DVALUES]; s.n_dvalues = p.cnt[KC_DVALUES];
  s.ms_comm = p.ms[KC_COMM]; s.n_comm = p.cnt[KC_COMM];
  s.ms_push = p.ms[KC_PUSH]; s.n_push = p.cnt[KC_PUSH];
  s.gpu_launches = h->launches;
}

extern "C" {

int eigkl_abi_version(void) { return EIGKL_ABI_VERSION; }

int eigkl_nccl_unique_id(void *id128) {
  try {
    if (!id128) return EIGKL_E_ARG;
    comm_unique_id(id128);
    return EIGKL_OK;
  } catch (const eigkl::Error &e) {
    std::lock_guard<std::mutex> lk(g_create_mutex);
    g_create_error = e.what();
    return e.code;
  } catch (...) {
    return EIGKL_E_NCCL;
  }
}

int eigkl_create(eigkl_handle **out, const eigkl_opts *opts) {
  if (!out) return EIGKL_E_ARG;
  *out = nullptr;
  eigkl_handle *h = nullptr;
  try {
    h = new eigkl_handle();
    if (opts) {
      if (opts->struct_size != sizeof(eigkl_opts)) throw Error(EIGKL_E_ARG, "eigkl_opts.struct_size mismatch (ABI)");
      h->opts = *opts;
    } else {
      h->opts.struct_size = sizeof(eigkl_opts);
    }
    if (h->opts.nranks <= 0) h->opts.nranks = 1;
    if (h->opts.rank < 0 || h->opts.rank >= h->opts.nranks) throw Error(EIGKL_E_ARG, "rank out of range");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev <= 0)
      throw Error(EIGKL_E_CUDA, std::string("no usable CUDA device (libeigkl has no CPU fallback): ") + cudaGetErrorString(ce));
    if (h->opts.device < 0 || h->opts.device >= ndev) throw Error(EIGKL_E_ARG, "device ordinal out of range");
    h->device = h->opts.device;
    EIGKL_CUDA(cudaSetDevice(h->device));
    cudaDeviceProp prop;
    EIGKL_CUDA(cudaGetDeviceProperties(&prop, h->device));
    if (prop.major < 10) throw Error(EIGKL_E_CUDA, std::string("device ") + prop.name + " is not sm_100 class; this library is built for sm_100a only");
    h->sm_count = prop.multiProcessorCount;
    EIGKL_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->timer.init();
    h->prof.on = (h->opts.flags & EIGKL_F_PROFILE) != 0;
    if (const char *m = getenv("EIGKL_SPMV_MODE")) h->spmv_mode = atoi(m);
    if (const char *m = getenv("EIGKL_SPMV_PDL")) h->spmv_pdl = atoi(m);
    if (const char *m = getenv("EIGKL_SPMV_RESIDENT")) h->spmv_resident = atoi(m);
    if (const char *m = getenv("EIGKL_GS_FUSED")) h->gs_fused = atoi(m);
    if (const char *m = getenv("EIGKL_KL_LOCAL")) h->kl_local = atoi(m);
    if (const char *m = getenv("EIGKL_KL_FLAT")) h->kl_flat = atoi(m);
    if (const char *m = getenv("EIGKL_DIST")) {       // rows | replicate | auto
      h->dist_mode = (strcmp(m, "rows") == 0 || strcmp(m, "1") == 0) ? 1 : (strcmp(m, "replicate") == 0 || strcmp(m, "2") == 0) ? 2 : 0;
    }
    if (const char *m = getenv("EIGKL_KL_DIST")) h->kl_dist = atoi(m);
    h->stats.struct_size = sizeof(eigkl_stats);
    if (h->opts.nranks > 1) comm_init(h);
    *out = h;
    return EIGKL_OK;
  } catch (const eigkl::Error &e) {
    std::lock_guard<std::mutex> lk(g_create_mutex);
    g_create_error = e.what();
    eigkl_destroy(h);            // releases whatever was already created (stream, timer events, communicator)
    return e.code;
  } catch (...) {
    std::lock_guard<std::mutex> lk(g_create_mutex);
    g_create_error = "eigkl_create: unknown error";
    eigkl_destroy(h);
    return EIGKL_E_NOMEM;
  }
}

void eigkl_destroy(eigkl_handle *h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  peer_arena_destroy(h);
  try { comm_destroy(h); } catch (...) {}
  if (h->l2_flush) cudaFree(h->l2_flush);
  h->timer.destroy();
  cudaStream_t s = h->stream;
  delete h;
  if (s) cudaStreamDestroy(s);
}

const char *eigkl_last_error(const eigkl_handle *h) {
  if (h) return h->err.c_str();
  return g_create_error.c_str();
}

int eigkl_get_stats(const eigkl_handle *hc, eigkl_stats *out) {
  eigkl_handle *h = const_cast<eigkl_handle *>(hc);
  return guarded(h, [&] {
    EIGKL_REQUIRE(out && out->struct_size == sizeof(eigkl_stats), EIGKL_E_ARG, "eigkl_stats.struct_size mismatch (ABI)");
    sync_profile_into_stats(h);
    *out = h->stats;
    out->struct_size = sizeof(eigkl_stats);
  });
}

int eigkl_set_profile(eigkl_handle *h, int on) {
  return guarded(h, [&] {
    EIGKL_CUDA(cudaSetDevice(h->device));
    EIGKL_CUDA(cudaStreamSynchronize(h->stream));
    h->prof.reset();
    h->prof.on = on != 0;
    h->stats.bytes_multidot_total = h->stats.bytes_update_total = 0.0;
  });
}

int eigkl_synchronize(eigkl_handle *h) {
  return guarded(h, [&] { EIGKL_CUDA(cudaStreamSynchronize(h->stream)); });
}

int eigkl_load_hgr(eigkl_handle *h, const char *path) {
  return guarded(h, [&] {
    EIGKL_REQUIRE(path, EIGKL_E_ARG, "path is NULL");
    EIGKL_CUDA(cudaSetDevice(h->device));
    HostHgr hg;
    parse_hgr(path, hg);
    upload_pins(h, hg.n_nodes, hg.n_nets, hg.net_off.data(), hg.pins.data());
  });
}

int eigkl_set_pins(eigkl_handle *h, int32_t n_nodes, int32_t n_nets, const int64_t *net_off, const int32_t *pins) {
  return guarded(h, [&] {
    EIGKL_CUDA(cudaSetDevice(h->device));
    upload_pins(h, n_nodes, n_nets, net_off, pins);
  });
}

int eigkl_get_sizes(const eigkl_handle *h, int32_t *n_nodes, int32_t *n_nets, int64_t *n_pins) {
  if (!h || !h->hg.loaded) return EIGKL_E_ARG;
  if (n_nodes) *n_nodes = h->hg.n_nodes;
  if (n_nets) *n_nets = h->hg.n_nets;
  if (n_pins) *n_pins = h->hg.n_pins;
  return EIGKL_OK;
}

int eigkl_invalidate(eigkl_handle *h) {
  return guarded(h, [&] {
    EIGKL_REQUIRE(h->hg.loaded, EIGKL_E_ARG, "eigkl_invalidate: no hypergraph loaded");
    EIGKL_CUDA(cudaSetDevice(h->device));
    EIGKL_CUDA(cudaStreamSynchronize(h->stream));
    h->ue.valid = false;
    h->ueL.valid = false;
    h->order.valid = false;
    h->L.valid = false;
    h->A.valid = false;
    h->kl.have_partition = false;
    h->eig.have_vector = h->eig.have_median = false;
  });
}

int eigkl_row_partition(int32_t n_rows, int32_t nranks, int32_t rank, int32_t *row_lo, int32_t *row_hi, int32_t *rows_padded) {
  if (n_rows <= 0 || nranks <= 0 || rank < 0 || rank >= nranks || !row_lo || !row_hi || !rows_padded) return EIGKL_E_ARG;
  row_partition(n_rows, nranks, rank, row_lo, row_hi, rows_padded);
  return EIGKL_OK;
}

int eigkl_get_stream(const eigkl_handle *h, void **stream) {
  if (!h || !stream) return EIGKL_E_ARG;
  *stream = (void *)h->stream;
  return EIGKL_OK;
}

int eigkl_assemble_laplacian(eigkl_handle *h) {
  return guarded(h, [&] {
    EIGKL_CUDA(cudaSetDevice(h->device));
    h->timer.start(h->stream);
    assemble_laplacian(h);
    h->timer.stop(h->stream);
    h->stats.ms_assemble_laplacian = h->timer.ms();
  });
}

int eigkl_fiedler(eigkl_handle *h, double *lambda2, double *vec) {
  return guarded(h, [&] {
    EIGKL_CUDA(cudaSetDevice(h->device));
    cudaEvent_t a, b;                       // own pair: the solver synchronises internally
    EIGKL_CUDA(cudaEventCreate(&a)); EIGKL_CUDA(cudaEventCreate(&b));
    EIGKL_CUDA(cudaEventRecord(a, h->stream));
    try {
      fiedler_solve(h);
    } catch (...) {
      cudaEventDestroy(a); cudaEventDestroy(b);
      throw;
    }
    EIGKL_CUDA(cudaEventRecord(b, h->stream));
    EIGKL_CUDA(cudaEventSynchronize(b));
    float ms = 0.f;
    EIGKL_CUDA(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a); cudaEventDestroy(b);
    h->stats.ms_fiedler = ms;
    spmv_resident_print_phases();
    if (lambda2) *lambda2 = h->eig.lambda2;
    if (vec) {
      EIGKL_CUDA(cudaMemcpyAsync(vec, h->eig.fiedler.p, (size_t)h->eig.n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
      EIGKL_CUDA(cudaStreamSynchronize(h->stream));
    }
  });
}

int eigkl_partition_from_fiedler(eigkl_handle *h, double *median, uint8_t *side) {
  return guarded(h, [&] {
    EIGKL_CUDA(cudaSetDevice(h->device));
    h->timer.start(h->stream);
    partition_from_fiedler(h);
    kl_set_partition_device(h, h->eig.side.p);
    h->timer.stop(h->stream);
    h->stats.ms_partition = h->timer.ms();
    if (median) *median = h->eig.median;
    if (side) {
      EIGKL_CUDA(cudaMemcpyAsync(side, h->eig.side.p, (size_t)h->eig.n, cudaMemcpyDeviceToHost, h->stream));
      EIGKL_CUDA(cudaStreamSynchronize(h->stream));
    }
  });
}

int eigkl_write_eig(eigkl_handle *h, const char *path) {
  return guarded(h, [&] {
    EIGKL_REQUIRE(path, EIGKL_E_ARG, "path is NULL");
    EIGKL_CUDA(cudaSetDevice(h->device));
    EIGKL_REQUIRE(h->eig.have_vector, EIGKL_E_ARG, "eigkl_write_eig: no Fiedler vector");
    if (!h->eig.have_median) partition_from_fiedler(h);
    std::vector<double> v((size_t)h->eig.n);
    EIGKL_CUDA(cudaMemcpyAsync(v.data(), h->eig.fiedler.p, v.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    EIGKL_CUDA(cudaStreamSynchronize(h->stream));
    write_eig_file(path, h->eig.lambda2, h->eig.median, v.data(), h->eig.n);
  });
}

int eigkl_assemble_kl_graph(eigkl_handle *h) {
  return guarded(h, [&] {
    EIGKL_CUDA(cudaSetDevice(h->device));
    h->timer.start(h->stream);
    assemble_kl_graph(h);
    h->timer.stop(h->stream);
    h->stats.ms_assemble_kl = h->timer.ms();
  });
}

int eigkl_set_partition(eigkl_handle *h, const uint8_t *side) {
  return guarded(h, [&] {
    EIGKL_CUDA(cudaSetDevice(h->device));
    kl_set_partition(h, side, nullptr, 0, nullptr, 0, true);
  });
}

int eigkl_set_partition_ordered(eigkl_handle *h, const int32_t *order0, int64_t n0, const int32_t *order1, int64_t n1) {
  return guarded(h, [&] {
    EIGKL_CUDA(cudaSetDevice(h->device));
    kl_set_partition(h, nullptr, order0, n0, order1, n1, false);
  });
}

int eigkl_load_eig(eigkl_handle *h, const char *path) {
  return guarded(h, [&] {
    EIGKL_REQUIRE(path, EIGKL_E_ARG, "path is NULL");
    EIGKL_REQUIRE(h->hg.loaded, EIGKL_E_ARG, "eigkl_load_eig: load the hypergraph first");
    EIGKL_CUDA(cudaSetDevice(h->device));
    std::vector<uint8_t> side;
    std::vector<int32_t> o0, o1;
    bool asc = true;
    read_eig_file(path, h->hg.n_nodes, side, o0, o1, asc);
    if (asc) kl_set_partition(h, side.data(), nullptr, 0, nullptr, 0, true);
    else kl_set_partition(h, nullptr, o0.data(), (int64_t)o0.size(), o1.data(), (int64_t)o1.size(), false);   // cKL.cpp:166-173
  });
}

int eigkl_kl_run(eigkl_handle *h, eigkl_trace *trace) {
  return guarded(h, [&] {
    EIGKL_CUDA(cudaSetDevice(h->device));
    // checked BEFORE the pass runs: a short trace must not cost a whole pass
    if (trace && h->kl.have_partition)
      EIGKL_REQUIRE(trace->capacity >= std::min(h->kl.n0, h->kl.n1) + 1, EIGKL_E_ARG,
                    "eigkl_trace.capacity too small (need min(|left|,|right|)+1)");
    kl_run(h);
    if (trace) {
      auto &k = h->kl;
      const int64_t rows = k.swaps + 1;
      EIGKL_REQUIRE(trace->capacity >= rows, EIGKL_E_ARG, "eigkl_trace.capacity too small (need min(|left|,|right|)+1)");
      trace->swaps = k.swaps;
      if (trace->cut) EIGKL_CUDA(cudaMemcpyAsync(trace->cut, k.t_cut.p, (size_t)rows * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
      if (trace->gain) EIGKL_CUDA(cudaMemcpyAsync(trace->gain, k.t_gain.p, (size_t)rows * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
      if (trace->node1) EIGKL_CUDA(cudaMemcpyAsync(trace->node1, k.t_n1.p, (size_t)rows * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
      if (trace->node2) EIGKL_CUDA(cudaMemcpyAsync(trace->node2, k.t_n2.p, (size_t)rows * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
      EIGKL_CUDA(cudaStreamSynchronize(h->stream));
    }
  });
}

int eigkl_write_trace(const char *path, const eigkl_trace *trace) {
  try {
    if (!path || !trace) return EIGKL_E_ARG;
    write_trace_file(path, trace);
    return EIGKL_OK;
  } catch (const eigkl::Error &e) {
    std::lock_guard<std::mutex> lk(g_create_mutex);
    g_create_error = e.what();
    return e.code;
  } catch (...) {
    return EIGKL_E_IO;
  }
}

int eigkl_get_partition(eigkl_handle *h, uint8_t *side) {
  return guarded(h, [&] {
    EIGKL_REQUIRE(side && h->kl.have_partition, EIGKL_E_ARG, "eigkl_get_partition: no partition");
    EIGKL_CUDA(cudaSetDevice(h->device));
    const size_t n = (size_t)h->hg.n_nodes;
    EIGKL_CUDA(cudaMemcpyAsync(side, h->kl.state.p, n, cudaMemcpyDeviceToHost, h->stream));
    EIGKL_CUDA(cudaStreamSynchronize(h->stream));
    for (size_t i = 0; i < n; ++i) side[i] &= 1u;
  });
}

int eigkl_kl_rollback(eigkl_handle *h, int64_t *best_row, float *best_cut) {
  return guarded(h, [&] {
    EIGKL_CUDA(cudaSetDevice(h->device));
    const int64_t b = kl_rollback(h, best_cut);
    if (best_row) *best_row = b;
  });
}

int eigkl_write_partition(eigkl_handle *h, const char *path) {
  return guarded(h, [&] {
    EIGKL_REQUIRE(path && h->kl.have_partition, EIGKL_E_ARG, "eigkl_write_partition: no partition");
    EIGKL_CUDA(cudaSetDevice(h->device));
    std::vector<uint8_t> side((size_t)h->hg.n_nodes);
    EIGKL_CUDA(cudaMemcpyAsync(side.data(), h->kl.state.p, side.size(), cudaMemcpyDeviceToHost, h->stream));
    EIGKL_CUDA(cudaStreamSynchronize(h->stream));
    write_partition_file(path, side.data(), h->hg.n_nodes);
  });
}

int eigkl_spmv(eigkl_handle *h, const double *x, double *y) {
  return guarded(h, [&] {
    EIGKL_REQUIRE(x && y && h->L.valid, EIGKL_E_ARG, "eigkl_spmv: Laplacian not assembled");
    EIGKL_CUDA(cudaSetDevice(h->device));
    const size_t n = (size_t)h->L.n;
    const bool dist = h->dist.valid;
    const int32_t n_pad = dist ? h->dist.n_pad : (int32_t)(ceil_div(h->L.n, 32) * 32);
    DBuf<double> dx, dy, dg; dx.alloc(n); dy.alloc((size_t)n_pad); dg.alloc(n);
    std::vector<int32_t> perm(n);
    EIGKL_CUDA(cudaMemcpyAsync(perm.data(), h->order.perm.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    EIGKL_CUDA(cudaStreamSynchronize(h->stream));
    std::vector<double> xp(n), yp(n);
    for (size_t i = 0; i < n; ++i) xp[i] = x[perm[i]];                // file ids -> the matrix's node order
    EIGKL_CUDA(cudaMemcpyAsync(dx.p, xp.data(), n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    EIGKL_CUDA(cudaMemsetAsync(dy.p, 0, (size_t)n_pad * sizeof(double), h->stream));
    const double *src = dy.p;
    if (dist) {
      // row-partitioned: this rank's rows of x into the stage buffer, halo pushed by the peers, rows of y gathered
      dist_stage_load(h, dx.p + h->L.row_lo);
      SpmvDist d{dist_push(h, 3), 0u, 0};
      spmv_launch_ex(h, dist_buf(h, 3), dist_own(h, 3), nullptr, dy.p, nullptr, nullptr, 1.0, 0.0, 0.0, &d);
      dist_gather_full(h, dy.p, dg.p);
      dist_check(h);
      src = dg.p;
    } else {
      spmv_launch(h, dx.p, dy.p, nullptr, nullptr);
    }
    EIGKL_CUDA(cudaMemcpyAsync(yp.data(), src, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    EIGKL_CUDA(cudaStreamSynchronize(h->stream));
    for (size_t i = 0; i < n; ++i) y[perm[i]] = yp[i];
    EIGKL_CUDA(cudaGetLastError());
  });
}

int eigkl_dvalues(eigkl_handle *h, float *val) {
  return guarded(h, [&] {
    EIGKL_REQUIRE(val, EIGKL_E_ARG, "val is NULL");
    EIGKL_CUDA(cudaSetDevice(h->device));
    kl_dvalues(h);
    EIGKL_CUDA(cudaMemcpyAsync(val, h->kl.val.p, (size_t)h->A.n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    EIGKL_CUDA(cudaStreamSynchronize(h->stream));
  });
}

int eigkl_get_kl_values(eigkl_handle *h, float *val) {
  return guarded(h, [&] {
    EIGKL_REQUIRE(val && h->A.valid && h->kl.have_partition, EIGKL_E_ARG, "eigkl_get_kl_values: no KL state");
    EIGKL_CUDA(cudaSetDevice(h->device));
    EIGKL_CUDA(cudaMemcpyAsync(val, h->kl.val.p, (size_t)h->A.n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    EIGKL_CUDA(cudaStreamSynchronize(h->stream));
  });
}

int eigkl_cut(eigkl_handle *h, float *cut) {
  return guarded(h, [&] {
    EIGKL_REQUIRE(cut, EIGKL_E_ARG, "cut is NULL");
    EIGKL_CUDA(cudaSetDevice(h->device));
    *cut = kl_cut0(h);
  });
}

int eigkl_get_laplacian(eigkl_handle *h, int32_t *rowptr, int32_t *col, double *val) {
  return guarded(h, [&] {
    EIGKL_REQUIRE(h->L.valid, EIGKL_E_ARG, "Laplacian not assembled");
    EIGKL_CUDA(cudaSetDevice(h->device));
    auto &L = h->L;
    if (rowptr) EIGKL_CUDA(cudaMemcpyAsync(rowptr, L.rowptr.p, ((size_t)L.n + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    if (col) EIGKL_CUDA(cudaMemcpyAsync(col, L.col.p, (size_t)L.nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    if (val) EIGKL_CUDA(cudaMemcpyAsync(val, L.val.p, (size_t)L.nnz * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    EIGKL_CUDA(cudaStreamSynchronize(h->stream));
  });
}

int eigkl_get_node_order(eigkl_handle *h, int32_t *perm) {
  return guarded(h, [&] {
    EIGKL_REQUIRE(perm && h->order.valid, EIGKL_E_ARG, "eigkl_get_node_order: assemble the Laplacian first");
    EIGKL_CUDA(cudaSetDevice(h->device));
    EIGKL_CUDA(cudaMemcpyAsync(perm, h->order.perm.p, (size_t)h->hg.n_nodes * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    EIGKL_CUDA(cudaStreamSynchronize(h->stream));
  });
}

int eigkl_get_kl_graph(eigkl_handle *h, int32_t *rowptr, int32_t *fwd_end, int32_t *col, float *w) {
  return guarded(h, [&] {
    EIGKL_REQUIRE(h->A.valid, EIGKL_E_ARG, "KL graph not assembled");
    EIGKL_CUDA(cudaSetDevice(h->device));
    auto &A = h->A;
    if (rowptr) EIGKL_CUDA(cudaMemcpyAsync(rowptr, A.rowptr.p, ((size_t)A.n + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    if (fwd_end) EIGKL_CUDA(cudaMemcpyAsync(fwd_end, A.fwd_end.p, (size_t)A.n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    if (col) EIGKL_CUDA(cudaMemcpyAsync(col, A.col.p, (size_t)A.nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    if (w) EIGKL_CUDA(cudaMemcpyAsync(w, A.w.p, (size_t)A.nnz * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    EIGKL_CUDA(cudaStreamSynchronize(h->stream));
  });
}

int eigkl_time_kernel(eigkl_handle *h, int what, int iters, int flush_l2, double *ms_avg) {
  return guarded(h, [&] {
    EIGKL_REQUIRE(ms_avg && iters > 0 && (what == 0 || what == 1), EIGKL_E_ARG, "eigkl_time_kernel: bad arguments");
    EIGKL_CUDA(cudaSetDevice(h->device));
    const size_t flush_bytes = (size_t)256 << 20;          // > 126 MB L2
    if (flush_l2 && !h->l2_flush) EIGKL_CUDA(cudaMalloc(&h->l2_flush, flush_bytes));
    DBuf<double> x, y;
    if (what == 0) {
      EIGKL_REQUIRE(h->L.valid, EIGKL_E_ARG, "Laplacian not assembled");
      x.alloc((size_t)h->L.n); y.alloc((size_t)h->L.n);
      EIGKL_CUDA(cudaMemsetAsync(x.p, 0, x.bytes(), h->stream));
    } else {
      EIGKL_REQUIRE(h->A.valid && h->kl.have_partition, EIGKL_E_ARG, "KL graph / partition missing");
    }
    const bool prof = h->prof.on;
    h->prof.on = false;
    cudaEvent_t a, b;
    EIGKL_CUDA(cudaEventCreate(&a)); EIGKL_CUDA(cudaEventCreate(&b));
    double total = 0.0;
    // row-partitioned: a chain of SpMVs through the three exchange buffers, each pushing its export rows to the peers
    // from its epilogue and waiting for the peers' halo of its input -- the chain a filter application runs
    // (collective: every rank calls with the same arguments)
    const bool dist = what == 0 && h->dist.valid;
    int cur = 0;
    uint32_t have = 0;
    if (dist) {
      for (int b = 0; b < 3; ++b) EIGKL_CUDA(cudaMemsetAsync(dist_own(h, b), 0, (size_t)h->dist.n_pad * sizeof(double), h->stream));
      have = dist_push(h, 0);
    }
    auto launch = [&] {
      if (dist) {
        const int out = (cur + 1) % 3;
        SpmvDist d{have, ++h->arena.seq, out};
        spmv_launch_ex(h, dist_buf(h, cur), dist_own(h, cur), nullptr, dist_own(h, out), nullptr, nullptr, 1.0, 0.0, 0.0, &d);
        have = d.push_seq;
        cur = out;
      } else if (what == 0) spmv_launch(h, x.p, y.p, nullptr, nullptr);
      else kl_dvalues(h);
    };
    launch(); launch();                                     // warm-up
    if (!flush_l2) {
      // back-to-back launches between ONE event pair: the queue stays full, so host launch latency is
      // not counted (this is how the kernel runs inside the Lanczos loop)
      EIGKL_CUDA(cudaEventRecord(a, h->stream));
      for (int i = 0; i < iters; ++i) launch();
      EIGKL_CUDA(cudaEventRecord(b, h->stream));
      EIGKL_CUDA(cudaEventSynchronize(b));
      float t = 0.f;
      EIGKL_CUDA(cudaEventElapsedTime(&t, a, b));
      total = t;
    } else {
      for (int i = 0; i < iters; ++i) {
        EIGKL_CUDA(cudaMemsetAsync(h->l2_flush, i & 0xff, flush_bytes, h->stream));
        EIGKL_CUDA(cudaEventRecord(a, h->stream));
        launch();
        EIGKL_CUDA(cudaEventRecord(b, h->stream));
        EIGKL_CUDA(cudaEventSynchronize(b));
        float t = 0.f;
        EIGKL_CUDA(cudaEventElapsedTime(&t, a, b));
        total += t;
      }
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    h->prof.on = prof;
    if (dist) dist_check(h);
    *ms_avg = total / iters;
  });
}

}  // extern "C"
