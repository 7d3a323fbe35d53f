/*
 * eigkl_oracle.h -- CPU restatement of the reference EIG+KL path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library.  The product (libeigkl.so, the CLIs, eig_kl_algorithm_b200/) never links or calls it.
 *
 * Parity status
 *   KL  : PINNED.  Checked against traces produced by running the reference itself
 *         (oracle/_ref/cKL and its instrumented twin, built by oracle/build_ref.sh from
 *         /root/reference/cKL.cpp): tests/golden/<c>.kl_trace_1core.txt (byte-exact) and
 *         <c>.kl_swaps.txt (swap node ids), circuits fract / ibm01 / industry2 / ibm10.
 *   EIG : PINNED to the reference's golden outputs pre_saved_EIG/{fract,ibm01,industry2}.hgr_out.txt
 *         (lambda2 rel. err <= 1e-8, sine <= 1e-6); ibm10's golden file is itself unconverged
 *          and is checked by residual only.  The reference solver is the
 *         third-party, un-pinned Spectra library (README.md:77; cEIG.cpp:194-198), absent from the
 *         image, so the Lanczos iteration below restates the published algorithm class
 *         (restarted Lanczos, nev=2, ncv=min(100,n/2), tol=1e-10, maxit=1000), not Spectra's code.
 */
#ifndef EIGKL_ORACLE_H
#define EIGKL_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- hypergraph -------------------------------------------------------------------------- */
typedef struct {
  int32_t n_nets, n_nodes;
  int64_t *net_off;   /* n_nets+1 */
  int32_t *pins;      /* 0-based node ids, file order */
} orc_hgr;
int  orc_hgr_load(const char *path, orc_hgr *out);          /* 0 ok, <0 error */
void orc_hgr_free(orc_hgr *h);

/* ---- libstdc++ hashtable iteration order ------------------------------ */
/* order[i] = index (into keys[]) of the i-th element visited when iterating a
 * std::unordered_{map,set}<uint32_t> into which keys[0..n) (distinct) were inserted one by one. */
void orc_stl_hash_order(const uint32_t *keys, int64_t n, int64_t *order);

/* ---- KL graph in the reference's traversal order (cKL.cpp:84-149, 225-251) ---------------- */
typedef struct {
  int32_t n;
  int64_t *rowptr;   /* n+1 */
  int64_t *fwd_end;  /* n : rowptr[v] <= fwd_end[v] <= rowptr[v+1]; [rowptr,fwd_end) = forward nbrs */
  int32_t *col;
  float   *w;
} orc_klgraph;
int  orc_kl_build(const orc_hgr *h, orc_klgraph *g);
void orc_kl_free(orc_klgraph *g);
/* val[v] = external - internal, cKL.cpp:225-251; side[v] in {0,1}, 0 = left */
void  orc_kl_dvalues(const orc_klgraph *g, const uint8_t *side, float *val);
/* cKL.cpp:199-223 evaluated on ONE thread; order0/order1 = remain[0]/remain[1] */
float orc_kl_cut0(const orc_klgraph *g, const uint8_t *side,
                  const int32_t *order0, int64_t n0, const int32_t *order1, int64_t n1);
/* One KL pass, cKL.cpp:288-390.  side is updated in place.  Returns number of swaps (rows 1..),
 * row 0 (initial cut) is cut[0]; arrays need capacity min(n0,n1)+1.                            */
int64_t orc_kl_run(const orc_klgraph *g, uint8_t *side,
                   const int32_t *order0, int64_t n0, const int32_t *order1, int64_t n1,
                   float *cut, float *gain, int32_t *node1, int32_t *node2, int64_t capacity);
/* the same pass with the literal O(|remain|) selection scans of cKL.cpp:341-355 (orc_kl_run caches the first
 * best element per block of 256 positions so that 2 M-node passes finish in seconds; results are identical) */
int64_t orc_kl_run_linear(const orc_klgraph *g, uint8_t *side,
                          const int32_t *order0, int64_t n0, const int32_t *order1, int64_t n1,
                          float *cut, float *gain, int32_t *node1, int32_t *node2, int64_t capacity);

/* ---- EIG (cEIG.cpp:86-133, 194-220) -------------------------------------------------------- */
typedef struct {
  int32_t n;
  int64_t *rowptr;  /* n+1 */
  int32_t *col;     /* ascending within a row, diagonal included */
  double  *val;
} orc_csr;
int  orc_laplacian(const orc_hgr *h, orc_csr *L);
void orc_csr_free(orc_csr *L);
void orc_spmv(const orc_csr *L, const double *x, double *y);   /* OpenMP over rows */
typedef struct {
  int32_t matvecs, restarts, converged, ncv;
  double  resid_est[2];
} orc_eig_stats;
/* two algebraically smallest eigenpairs; returns the larger one (lambda2) and its unit vector */
int    orc_fiedler(const orc_csr *L, double *lambda2, double *vec, orc_eig_stats *st);
int    orc_fiedler_bounded(const orc_csr *L, int max_restarts, double *lambda2, double *vec, orc_eig_stats *st);
double orc_median(const double *v, int32_t n);                 /* cEIG.cpp:55-65 */
/* writes the cEIG output format (cEIG.cpp:213-220) */
int    orc_write_eig(const char *path, double lambda2, const double *vec, int32_t n);
/* reads the side column of a cEIG output file the way cKL does (cKL.cpp:155-174) */
int    orc_read_eig_sides(const char *path, int32_t n, uint8_t *side, double *lambda2, double *median, double *vec);
/* writes a KL trace the way cKL does (cKL.cpp:315,380) */
int    orc_write_trace(const char *path, const float *cut, const float *gain, int64_t swaps);

/* dense symmetric eigen-solver used by orc_fiedler (Householder + implicit QL); exposed for tests */
void   orc_sym_eig(int n, double *a /* n*n row-major in, eigenvectors (columns) out */, double *evals /* ascending */);

#ifdef __cplusplus
}
#endif
#endif
