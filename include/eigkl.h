/*
 * eigkl.h -- C ABI of libeigkl.so, the B200-native (sm_100a) EIG+KL bipartitioner.
 *
 * This is the drop-in boundary for the hot path of yhinai/EIG-KL-Algorithm.  The reference has no
 * library or FFI surface: its operator surface is three executables coupled by text files
 * (SURVEY.md section 8b).  Each entry point below therefore names the reference code it replaces;
 * the executables cEIG / cKL / gKL shipped with this repo (eig_kl_algorithm_b200/cli/) are ~60-line
 * callers of this ABI that keep the reference's argv, file names, formats and exit codes.
 *
 * Conventions: plain C types only; every function returns an int status (EIGKL_OK == 0, negative
 * on error) and never throws; eigkl_last_error() gives the message.  The caller owns every host
 * buffer it passes; the handle owns all device memory, one CUDA stream and (optionally) one NCCL
 * communicator.  A handle is not thread-safe; use one handle per GPU / rank.
 * There is no CPU fallback: every compute entry point fails with EIGKL_E_CUDA without a GPU.
 */
#ifndef EIGKL_H
#define EIGKL_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EIGKL_ABI_VERSION 2

enum {
  EIGKL_OK        = 0,
  EIGKL_E_ARG     = -1,   /* bad argument / call order                               */
  EIGKL_E_IO      = -2,   /* cannot open / read / write a file                       */
  EIGKL_E_FORMAT  = -3,   /* malformed .hgr or EIG file, pin id out of range, duplicate pin in a net */
  EIGKL_E_CUDA    = -4,   /* CUDA runtime error or no usable device                  */
  EIGKL_E_NCCL    = -5,   /* NCCL error                                              */
  EIGKL_E_NOCONV  = -6,   /* Lanczos did not converge within max_restarts            */
  EIGKL_E_NOMEM   = -7
};

typedef struct eigkl_handle eigkl_handle;

/* options; zero-initialise and override.  sizeof is checked through struct_size. */
typedef struct {
  uint32_t struct_size;      /* = sizeof(eigkl_opts)                                              */
  int32_t  device;           /* CUDA device ordinal (default 0)                                   */
  int32_t  rank, nranks;     /* row-partition rank / world size (default 0 / 1)                   */
  const void *nccl_unique_id;/* 128-byte ncclUniqueId shared by all ranks, or NULL when nranks==1 */
  /* Fiedler solve -- defaults restate cEIG.cpp:195-198 (Spectra: nev=2, ncv=min(100,n/2),
   * tol=1e-10, maxit=1000)                                                                        */
  int32_t  ncv;              /* 0 => min(100, n/2)                                                */
  int32_t  max_restarts;     /* 0 => 1000                                                         */
  double   tol;              /* 0 => 1e-10                                                        */
  int32_t  keep;             /* Ritz vectors kept at a thick restart; 0 => library default        */
  uint64_t seed;             /* start-vector seed (the reference's is Spectra's fixed seed 0)     */
  /* KL */
  int32_t  kl_cluster;       /* CTAs in the KL cluster (1,2,4,8,16); 0 => chosen from the size    */
  uint32_t flags;            /* EIGKL_F_*                                                         */
} eigkl_opts;

#define EIGKL_F_PROFILE   0x1u  /* bracket every kernel class with CUDA events (see eigkl_stats); bit 0x2 is unused */
#define EIGKL_F_NATURAL_ORDER 0x8u /* EIG stage keeps the file's node numbering (default: nodes renumbered by first net, for gather locality) */
#define EIGKL_F_PLAIN_LANCZOS 0x4u /* no Chebyshev filter: Lanczos on L itself (degree-1 map), as Spectra does */

/* per-call statistics.  Times are device times from CUDA events on the handle's stream, in ms.   */
typedef struct {
  uint32_t struct_size;
  /* sizes */
  int64_t  n_nodes, n_nets, n_pins, n_pairs;  /* pairs = sum k(k-1)/2                              */
  int64_t  nnz_laplacian;                     /* symmetric off-diagonals + diagonals               */
  int64_t  nnz_kl;                            /* symmetric off-diagonals of the KL graph           */
  /* Fiedler solve */
  int32_t  ncv, matvecs, restarts, converged;
  double   resid_est[2];                      /* [0] Ritz estimate |beta y_last| (filtered space), [1] TRUE |L v - lambda2 v| */
  double   lambda[2];                         /* the two smallest eigenvalues found (ascending)    */
  int32_t  cheb_degree, lanczos_steps;        /* filter degree (matvecs ~ steps * degree)          */
  /* KL */
  int64_t  kl_swaps;
  int32_t  kl_cluster, kl_threads;
  /* kernel launches issued by this handle since creation (graph replays count their nodes)        */
  int64_t  gpu_launches;
  /* device time per stage of the last call of each kind (always measured, 2 events per stage)     */
  double   ms_assemble_laplacian, ms_assemble_kl, ms_fiedler, ms_partition, ms_kl_setup, ms_kl_loop;
  /* per kernel class, only with EIGKL_F_PROFILE: summed device time and launch count              */
  double   ms_spmv, ms_multidot, ms_update, ms_restart, ms_dvalues;
  int64_t  n_spmv, n_multidot, n_update, n_restart, n_dvalues;
  /* algorithmic bytes of ONE launch of the kernel class at this problem size          */
  double   bytes_spmv, bytes_dvalues;
  double   bytes_multidot_total, bytes_update_total;   /* summed over the profiled launches        */
  /* SpMVs carried by one launch of the SpMV kernel class: 1 (one spmv kernel per product), or the
   * Chebyshev degree when the resident filter runs a whole filter application as one launch       */
  int32_t  spmv_per_launch;
  int32_t  resident_k;           /* entries per thread of the resident filter kernel, 0 = not used  */
  /* 1 when both Gram-Schmidt passes of a Lanczos step run as ONE cooperative launch (counted in
   * n_multidot / ms_multidot; n_update stays 0), 0 when they are separate multidot/update launches */
  int32_t  gs_fused;
  int32_t  gs_cache_cols;        /* basis columns the fused kernel keeps in shared memory            */
  int32_t  kl_local;             /* swap loop as one CTA: 1 = tile keys and side bits in shared memory, 2 = tile keys
                                  * in shared memory and state bytes in global memory; 0 = global-memory cluster kernel */
  int32_t  kl_flat;              /* 1 = the flat (by-entry) form of the shared-memory swap loop, 0 = warp per row */
  /* multi-rank Lanczos (nranks > 1): 1 = every rank solved the whole problem (the matrix fits one chip), R = rows
   * partitioned over R ranks; then this rank's rows, the halo values it receives and the rows it pushes per SpMV */
  int32_t  dist_ranks, dist_rows;
  int64_t  dist_halo, dist_exports;
  /* only with profiling on, row-partitioned mode: the NCCL all-reduces of the Lanczos dot products / norms, and the
   * stand-alone halo pushes (one per filter application; the other pushes ride inside the SpMV kernel)            */
  double   ms_comm, ms_push;
  int64_t  n_comm, n_push;
} eigkl_stats;

/* KL trace, one row per swap plus row 0 (the initial cut) -- the rows cKL writes to
 * results/<base>_KL_CutSize[_EIG]_output.txt (cKL.cpp:315,380) plus the swapped node ids.
 * Arrays are caller-owned with `capacity` entries (need min(|left|,|right|)+1); any may be NULL.  */
typedef struct {
  int64_t  capacity;
  int64_t  swaps;            /* out: rows 1..swaps are swaps                                      */
  float   *cut;              /* cut[0] = initial cut; cut[i] = cut after swap i                   */
  float   *gain;             /* gain[0] = 0                                                       */
  int32_t *node1, *node2;    /* node1 leaves side 0, node2 leaves side 1 (0-based); -1 in row 0   */
} eigkl_trace;

/* ---- lifetime -------------------------------------------------------------------------------- */
int  eigkl_abi_version(void);
/* fills id[128] with a fresh ncclUniqueId (rank 0 calls it, then shares it with the other ranks) */
int  eigkl_nccl_unique_id(void *id128);
int  eigkl_create(eigkl_handle **out, const eigkl_opts *opts);
void eigkl_destroy(eigkl_handle *h);
const char *eigkl_last_error(const eigkl_handle *h);   /* h may be NULL: error of a failed create */
int  eigkl_get_stats(const eigkl_handle *h, eigkl_stats *out);
int  eigkl_synchronize(eigkl_handle *h);
/* switches the per-kernel-class event brackets (EIGKL_F_PROFILE) on or off for the calls that follow, and clears
 * the per-class sums.  With several ranks every rank must switch alike.                                          */
int  eigkl_set_profile(eigkl_handle *h, int on);

/* ---- input: the hypergraph ------------------------------------------------------------------- */
/* Parses a .hgr file: header "<nets> <nodes>", then <nets> lines of 1-based pin ids.
 * Replaces cEIG.cpp:178-182,94-101 ; cKL.cpp:92-115 ; gKL.cu:581-620.                            */
int  eigkl_load_hgr(eigkl_handle *h, const char *path);
/* Same, from host arrays: net e has pins[net_off[e] .. net_off[e+1]) (0-based node ids).         */
int  eigkl_set_pins(eigkl_handle *h, int32_t n_nodes, int32_t n_nets,
                    const int64_t *net_off, const int32_t *pins);
int  eigkl_get_sizes(const eigkl_handle *h, int32_t *n_nodes, int32_t *n_nets, int64_t *n_pins);
/* Drops everything derived from the pins (assembled matrices, Fiedler vector, partition) but keeps the
 * pins resident in HBM, so that the next eigkl_assemble_* call redoes the sort + segmented reduce.   */
int  eigkl_invalidate(eigkl_handle *h);
/* The 1-D row partition used when nranks > 1 (host-only, needs no GPU): rank owns rows [row_lo,row_hi)
 * of the Laplacian and of every Lanczos vector; rows_padded (a multiple of 32) is the per-rank slot of
 * the all-gathered vectors, so global row g sits at index g of a gathered buffer.                     */
int  eigkl_row_partition(int32_t n_rows, int32_t nranks, int32_t rank, int32_t *row_lo, int32_t *row_hi,
                         int32_t *rows_padded);
/* The CUDA stream (cudaStream_t) every kernel of this handle is launched on, for callers that want to
 * record their own events around calls.                                                              */
int  eigkl_get_stream(const eigkl_handle *h, void **stream);

/* ---- EIG stage (the cEIG executable) ---------------------------------------------------------- */
/* Clique-model Laplacian L = D - A, A_ij = sum over nets containing i and j of 2.0/|net| (fp64), as a
 * GPU sort + segmented reduce over net pins.  Replaces initializeMatrix, cEIG.cpp:86-133.        */
int  eigkl_assemble_laplacian(eigkl_handle *h);
/* Two algebraically smallest eigenpairs of L by restarted Lanczos; reports the larger one
 * (lambda2, Fiedler vector, unit norm).  Replaces the Spectra call cEIG.cpp:194-207.
 * lambda2 / vec (n_nodes doubles, host) may be NULL: the vector then stays on the device.        */
int  eigkl_fiedler(eigkl_handle *h, double *lambda2, double *vec);
/* median of the Fiedler vector (cEIG.cpp:55-65) and side_i = (median > v_i) (cEIG.cpp:218), on the
 * device; makes that partition the KL initial partition (the fused pipeline gKL2.cu:1018-1024
 * aimed at).  median / side (n_nodes bytes, host) may be NULL.                                   */
int  eigkl_partition_from_fiedler(eigkl_handle *h, double *median, uint8_t *side);
/* Writes pre_saved_EIG/<base>_out.txt: lambda2, median, then "i\tside\tv_i" with 12 significant
 * digits.  Replaces cEIG.cpp:213-220.                                                            */
int  eigkl_write_eig(eigkl_handle *h, const char *path);

/* ---- KL stage (the cKL / gKL executables) ------------------------------------------------------ */
/* KL graph: A[a][b] += 1.0f/(|net|-1) in file order (fp32), rows stored in the reference's traversal
 * order (forward neighbours in libstdc++ unordered_map order, then backward neighbours ascending).
 * Replaces InitializeSparsMatrix + initNodeConnections, cKL.cpp:84-149,53-72 ; gKL.cu:573-666.   */
int  eigkl_assemble_kl_graph(eigkl_handle *h);
/* Initial partition.  side[i] in {0,1}; remain[0]/remain[1] are the nodes of each side in ascending
 * id order (what the -EIG branch produces).  Replaces shuffleSparceMatrix, cKL.cpp:151-174.      */
int  eigkl_set_partition(eigkl_handle *h, const uint8_t *side);
/* Same with explicit remain[] orders (the random branch, cKL.cpp:175-193, with the shuffle done by
 * the caller): ties in the pair selection go to the earlier position.                            */
int  eigkl_set_partition_ordered(eigkl_handle *h, const int32_t *order0, int64_t n0,
                                 const int32_t *order1, int64_t n1);
/* Reads the side column of a pre_saved_EIG file the way cKL does (skip 2 lines, "node side w").
 * Replaces cKL.cpp:155-174 ; gKL.cu:270-300.                                                     */
int  eigkl_load_eig(eigkl_handle *h, const char *path);
/* One KL pass: D-values, pair selection (first max / first min, lowest position wins ties),
 * gain = D1 + D2 - 2w, lock-and-swap, recompute of the neighbours' D-values; stops after
 * floor(log2 N)+6 consecutive non-positive gains.  Replaces KL(), cKL.cpp:288-390 ; gKL.cu:417-549.
 * trace may be NULL (the trace then stays on the device; swaps is reported in eigkl_stats).      */
int  eigkl_kl_run(eigkl_handle *h, eigkl_trace *trace);
/* Writes results/<base>_KL_CutSize[_EIG]_output.txt exactly as cKL.cpp:315,380 does.             */
int  eigkl_write_trace(const char *path, const eigkl_trace *trace);
/* current side of every node (after eigkl_kl_run: the final partition, which the reference never
 * writes anywhere -- SURVEY.md section 8f.3)                                                     */
int  eigkl_get_partition(eigkl_handle *h, uint8_t *side);
/* Extensions the reference stops short of (cKL.cpp:363 tracks the best cut, cKL.cpp:395-405
 * neither rolls back to it nor saves the partition).  eigkl_kl_rollback undoes, on the device, the swaps after the
 * first minimum of the last pass' cut column and reports that row / cut; eigkl_write_partition writes "<node>\t<side>"
 * per line (0-based ids, ascending).  Calling eigkl_kl_run again without a new partition starts a NEW pass from the
 * current sides with fresh locks and ascending remain[] lists, as KL() does on every call (cKL.cpp:290-301).           */
int  eigkl_kl_rollback(eigkl_handle *h, int64_t *best_row, float *best_cut);
int  eigkl_write_partition(eigkl_handle *h, const char *path);

/* ---- test / measurement hooks ------------------------------------------------------------------ */
int  eigkl_spmv(eigkl_handle *h, const double *x, double *y);                 /* y = L x (host buffers) */
int  eigkl_dvalues(eigkl_handle *h, float *val);       /* connections() for every node, cKL.cpp:225-251 */
int  eigkl_cut(eigkl_handle *h, float *cut);           /* calCutSize() on one thread, cKL.cpp:199-223   */
/* the D-values as the swap loop left them (nodeGains[] after KL(), cKL.cpp:40,270) -- no recomputation  */
int  eigkl_get_kl_values(eigkl_handle *h, float *val);
/* device copies of the assembled matrices (any pointer may be NULL); sizes from eigkl_get_stats   */
int  eigkl_get_laplacian(eigkl_handle *h, int32_t *rowptr, int32_t *col, double *val);
/* the matrix above is stored in the EIG stage's node order: perm[new id] = file id (identity with
 * EIGKL_F_NATURAL_ORDER).  Every other entry point takes and returns vectors in the file's ids.          */
int  eigkl_get_node_order(eigkl_handle *h, int32_t *perm);
int  eigkl_get_kl_graph(eigkl_handle *h, int32_t *rowptr, int32_t *fwd_end, int32_t *col, float *w);
/* runs `iters` back-to-back launches of one kernel class on resident data and returns the average
 * device time per launch in ms (CUDA events on the handle's stream).  what: 0 = SpMV, 1 = D-values.
 * flush_l2 != 0 writes a >L2 buffer between launches (excluded from the time).                   */
int  eigkl_time_kernel(eigkl_handle *h, int what, int iters, int flush_l2, double *ms_avg);

#ifdef __cplusplus
}
#endif
#endif /* EIGKL_H */
