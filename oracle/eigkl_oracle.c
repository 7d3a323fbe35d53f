/*
 * eigkl_oracle.c -- CPU restatement of the reference EIG+KL path.  TEST INFRASTRUCTURE ONLY
 * (see eigkl_oracle.h for who may load it and for the parity status).
 *
 * Every function cites the reference lines it restates (paths relative to /root/reference).
 * The algorithmic results are the reference's; the data structures are not: the reference keeps
 * vector<unordered_map> and re-derives neighbour sets with O(N) hash probes per call
 * (cKL.cpp:53-72, 225-251), this file keeps one CSR whose rows are stored in exactly the order the
 * reference's loops visit them, which makes every float sum bit-identical at O(degree) cost.
 */
#define _GNU_SOURCE
#include "eigkl_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#include <ctype.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ================================================================================================
 * .hgr parser -- cEIG.cpp:178-182,94-101 ; cKL.cpp:92-115
 * line 1 = "<nets> <nodes>", then <nets> lines of 1-based whitespace separated pin ids.
 * ============================================================================================== */
int orc_hgr_load(const char *path, orc_hgr *out) {
  memset(out, 0, sizeof(*out));
  FILE *f = fopen(path, "rb");
  if (!f) return -1;
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  char *buf = (char *)malloc((size_t)sz + 1);
  if (fread(buf, 1, (size_t)sz, f) != (size_t)sz) { fclose(f); free(buf); return -1; }
  fclose(f);
  buf[sz] = 0;
  char *p = buf, *end = buf + sz;
  long nets = strtol(p, &p, 10), nodes = strtol(p, &p, 10);
  if (nets < 0 || nodes <= 0) { free(buf); return -2; }
  while (p < end && *p != '\n') ++p;      /* rest of header ignored (stringstream reads two tokens) */
  if (p < end) ++p;
  out->n_nets = (int32_t)nets;
  out->n_nodes = (int32_t)nodes;
  out->net_off = (int64_t *)calloc((size_t)nets + 1, sizeof(int64_t));
  size_t cap = (size_t)sz / 2 + 16, np = 0;
  out->pins = (int32_t *)malloc(cap * sizeof(int32_t));
  for (long i = 0; i < nets; ++i) {       /* exactly <nets> getline calls */
    while (p < end && *p != '\n') {
      if (isdigit((unsigned char)*p)) {
        long v = strtol(p, &p, 10);
        if (v < 1 || v > nodes) { orc_hgr_free(out); free(buf); return -2; }
        out->pins[np++] = (int32_t)(v - 1);   /* 0-based: cEIG.cpp:99, cKL.cpp:114 */
      } else if (*p == ' ' || *p == '\t' || *p == '\r') {
        ++p;
      } else {                                /* operator>> stops at a non-numeric token */
        while (p < end && *p != '\n') ++p;
      }
    }
    if (p < end) ++p;
    out->net_off[i + 1] = (int64_t)np;
  }
  free(buf);
  return 0;
}
void orc_hgr_free(orc_hgr *h) {
  free(h->net_off); free(h->pins);
  memset(h, 0, sizeof(*h));
}

/* ================================================================================================
 * libstdc++ (GCC 13) _Hashtable emulation -- SURVEY.md Appendix E.
 * std::hash<uint32_t> is the identity; bucket = key % bucket_count; load factor 1.0; growth chain
 * below is what _Prime_rehash_policy::_M_next_bkt yields for one-by-one inserts (measured with the
 * real container in tests/test_stl_order.py).  A node inserted into an empty bucket goes to the
 * FRONT of the global list; into a non-empty bucket it goes to the front of that bucket's chain.
 * A rehash walks the old list in order and re-inserts with the same two rules.
 * ============================================================================================== */
static const uint32_t ORC_BKT_CHAIN[] = {13u, 29u, 59u, 127u, 257u, 541u, 1109u, 2357u, 5087u, 10273u,
    20753u, 42043u, 85229u, 172933u, 351061u, 712697u, 1447153u, 2938679u, 5967347u, 12117689u,
    24607243u, 49969847u, 101473717u};
#define ORC_N_CHAIN ((int)(sizeof(ORC_BKT_CHAIN) / sizeof(ORC_BKT_CHAIN[0])))
#define HT_EMPTY (-2)   /* bucket has no nodes          */
#define HT_BB    (-1)   /* bucket's before-node is _M_before_begin */

typedef struct {
  uint32_t *key;      /* node -> key          */
  int64_t  *next;     /* node -> next node    */
  int64_t  *bkt;      /* bucket -> before node */
  int64_t  head;      /* _M_before_begin._M_nxt */
  int64_t  count;
  int      level;     /* -1: single bucket (empty table), else index in ORC_BKT_CHAIN */
  uint32_t nb;        /* bucket count */
  int64_t  cap_nodes, cap_bkt;
} ht_t;

static void ht_init(ht_t *t, int64_t max_nodes) {
  int lv = 0;
  while (lv < ORC_N_CHAIN - 1 && (int64_t)ORC_BKT_CHAIN[lv] < max_nodes) ++lv;
  t->cap_nodes = max_nodes > 0 ? max_nodes : 1;
  t->cap_bkt = ORC_BKT_CHAIN[lv];
  t->key = (uint32_t *)malloc((size_t)t->cap_nodes * sizeof(uint32_t));
  t->next = (int64_t *)malloc((size_t)t->cap_nodes * sizeof(int64_t));
  t->bkt = (int64_t *)malloc((size_t)t->cap_bkt * sizeof(int64_t));
  t->head = -1; t->count = 0; t->level = -1; t->nb = 1;
  t->bkt[0] = HT_EMPTY;
}
static void ht_reset(ht_t *t) {
  t->head = -1; t->count = 0; t->level = -1; t->nb = 1;
  t->bkt[0] = HT_EMPTY;
}
static void ht_free(ht_t *t) { free(t->key); free(t->next); free(t->bkt); }

static inline void ht_link(ht_t *t, int64_t node, uint32_t b, uint32_t *bbegin_bkt) {
  if (t->bkt[b] == HT_EMPTY) {
    t->next[node] = t->head;
    t->head = node;
    if (t->next[node] >= 0) {
      uint32_t ob = bbegin_bkt ? *bbegin_bkt : (t->key[t->next[node]] % t->nb);
      t->bkt[ob] = node;
    }
    t->bkt[b] = HT_BB;
    if (bbegin_bkt) *bbegin_bkt = b;
  } else {
    int64_t prev = t->bkt[b];
    if (prev == HT_BB) { t->next[node] = t->head; t->head = node; }
    else               { t->next[node] = t->next[prev]; t->next[prev] = node; }
  }
}
static void ht_rehash(ht_t *t, uint32_t nb) {        /* hashtable.h:_M_rehash_aux(unique) */
  for (uint32_t i = 0; i < nb; ++i) t->bkt[i] = HT_EMPTY;
  int64_t p = t->head;
  t->head = -1; t->nb = nb;
  uint32_t bbegin = 0;
  while (p >= 0) {
    int64_t nx = t->next[p];
    ht_link(t, p, t->key[p] % nb, &bbegin);
    p = nx;
  }
}
/* returns node index of key, inserting it (value semantics are the caller's) ; *fresh = 1 if new */
static inline int64_t ht_find_or_insert(ht_t *t, uint32_t k, int *fresh) {
  uint32_t b = k % t->nb;
  if (t->bkt[b] != HT_EMPTY) {
    int64_t p = (t->bkt[b] == HT_BB) ? t->head : t->next[t->bkt[b]];
    while (p >= 0 && t->key[p] % t->nb == b) {
      if (t->key[p] == k) { *fresh = 0; return p; }
      p = t->next[p];
    }
  }
  /* _M_insert_unique_node: rehash check first (count+1 > next_resize, next_resize == nb, or 0) */
  int64_t next_resize = (t->level < 0) ? 0 : (int64_t)t->nb;
  if (t->count + 1 > next_resize) {
    t->level += 1;
    ht_rehash(t, ORC_BKT_CHAIN[t->level]);
    b = k % t->nb;
  }
  int64_t node = t->count++;
  t->key[node] = k;
  ht_link(t, node, b, NULL);
  *fresh = 1;
  return node;
}

void orc_stl_hash_order(const uint32_t *keys, int64_t n, int64_t *order) {
  ht_t t;
  ht_init(&t, n);
  for (int64_t i = 0; i < n; ++i) { int fr; ht_find_or_insert(&t, keys[i], &fr); }
  int64_t i = 0;
  for (int64_t p = t.head; p >= 0; p = t.next[p]) order[i++] = p;
  ht_free(&t);
}

/* ================================================================================================
 * KL graph -- cKL.cpp:84-149 (InitializeSparsMatrix) in the row order of cKL.cpp:225-251.
 *   w_net = 1.0f/(k-1) (float)                                           cKL.cpp:117
 *   adjacencyList[min][max] += w_net, nets in file order, pairs (j<k)    cKL.cpp:119-131
 *   row(v) = [forward nbrs (b>v) in unordered_map iteration order] ++ [backward nbrs ascending]
 * ============================================================================================== */
int orc_kl_build(const orc_hgr *h, orc_klgraph *g) {
  memset(g, 0, sizeof(*g));
  const int32_t n = h->n_nodes;
  /* pass 1: pairs per min-node */
  int64_t *cnt = (int64_t *)calloc((size_t)n + 1, sizeof(int64_t));
  int64_t P = 0;
  for (int32_t e = 0; e < h->n_nets; ++e) {
    const int32_t *p = h->pins + h->net_off[e];
    int64_t k = h->net_off[e + 1] - h->net_off[e];
    for (int64_t j = 0; j < k; ++j)
      for (int64_t l = j + 1; l < k; ++l) {
        int32_t a = p[j] < p[l] ? p[j] : p[l];
        if (p[j] == p[l]) { free(cnt); return -2; }   /* duplicate pin: no pinned behaviour */
        cnt[a + 1]++; P++;
      }
  }
  for (int32_t v = 0; v < n; ++v) cnt[v + 1] += cnt[v];
  int32_t *pb = (int32_t *)malloc((size_t)(P ? P : 1) * sizeof(int32_t));
  float   *pw = (float *)malloc((size_t)(P ? P : 1) * sizeof(float));
  int64_t *cur = (int64_t *)malloc((size_t)n * sizeof(int64_t));
  memcpy(cur, cnt, (size_t)n * sizeof(int64_t));
  for (int32_t e = 0; e < h->n_nets; ++e) {           /* stable: file order within a row */
    const int32_t *p = h->pins + h->net_off[e];
    int64_t k = h->net_off[e + 1] - h->net_off[e];
    float w = 1.0f / (float)(k - 1);                  /* cKL.cpp:117 (size_t k-1 -> float) */
    for (int64_t j = 0; j < k; ++j)
      for (int64_t l = j + 1; l < k; ++l) {
        int32_t a = p[j], b = p[l];
        if (a > b) { int32_t t = a; a = b; b = t; }
        pb[cur[a]] = b; pw[cur[a]] = w; cur[a]++;
      }
  }
  /* pass 2: per row, replay the unordered_map: find-or-insert, += w ; then read iteration order */
  int64_t maxrow = 0;
  for (int32_t v = 0; v < n; ++v) if (cnt[v + 1] - cnt[v] > maxrow) maxrow = cnt[v + 1] - cnt[v];
  ht_t t;
  ht_init(&t, maxrow);
  float *acc = (float *)malloc((size_t)(maxrow ? maxrow : 1) * sizeof(float));
  int64_t *fdeg = (int64_t *)calloc((size_t)n + 1, sizeof(int64_t));
  int64_t *bdeg = (int64_t *)calloc((size_t)n + 1, sizeof(int64_t));
  /* forward lists packed in place over pb/pw (unique <= raw) */
  int64_t *fstart = (int64_t *)malloc(((size_t)n + 1) * sizeof(int64_t));
  int32_t *fcol = (int32_t *)malloc((size_t)(P ? P : 1) * sizeof(int32_t));
  float   *fw = (float *)malloc((size_t)(P ? P : 1) * sizeof(float));
  int64_t U = 0;
  for (int32_t v = 0; v < n; ++v) {
    fstart[v] = U;
    ht_reset(&t);
    for (int64_t i = cnt[v]; i < cnt[v + 1]; ++i) {
      int fr;
      int64_t node = ht_find_or_insert(&t, (uint32_t)pb[i], &fr);
      if (fr) acc[node] = 0.0f;
      acc[node] += pw[i];                             /* cKL.cpp:128 */
    }
    for (int64_t p = t.head; p >= 0; p = t.next[p]) {
      fcol[U] = (int32_t)t.key[p]; fw[U] = acc[p]; ++U;
      bdeg[t.key[p]]++;
    }
    fdeg[v] = U - fstart[v];
  }
  fstart[n] = U;
  ht_free(&t); free(acc); free(pb); free(pw); free(cur); free(cnt);
  /* pass 3: rows = forward (map order) ++ backward (ascending source) */
  g->n = n;
  g->rowptr = (int64_t *)malloc(((size_t)n + 1) * sizeof(int64_t));
  g->fwd_end = (int64_t *)malloc((size_t)(n ? n : 1) * sizeof(int64_t));
  g->rowptr[0] = 0;
  for (int32_t v = 0; v < n; ++v) g->rowptr[v + 1] = g->rowptr[v] + fdeg[v] + bdeg[v];
  int64_t nnz = g->rowptr[n];
  g->col = (int32_t *)malloc((size_t)(nnz ? nnz : 1) * sizeof(int32_t));
  g->w = (float *)malloc((size_t)(nnz ? nnz : 1) * sizeof(float));
  int64_t *bcur = (int64_t *)malloc((size_t)(n ? n : 1) * sizeof(int64_t));
  for (int32_t v = 0; v < n; ++v) {
    g->fwd_end[v] = g->rowptr[v] + fdeg[v];
    bcur[v] = g->fwd_end[v];
    memcpy(g->col + g->rowptr[v], fcol + fstart[v], (size_t)fdeg[v] * sizeof(int32_t));
    memcpy(g->w + g->rowptr[v], fw + fstart[v], (size_t)fdeg[v] * sizeof(float));
  }
  for (int32_t a = 0; a < n; ++a)                      /* ascending a => backward lists ascending */
    for (int64_t i = fstart[a]; i < fstart[a + 1]; ++i) {
      int32_t b = fcol[i];
      g->col[bcur[b]] = a; g->w[bcur[b]] = fw[i]; bcur[b]++;
    }
  free(bcur); free(fdeg); free(bdeg); free(fstart); free(fcol); free(fw);
  return 0;
}
void orc_kl_free(orc_klgraph *g) {
  free(g->rowptr); free(g->fwd_end); free(g->col); free(g->w);
  memset(g, 0, sizeof(*g));
}

/* connections(), cKL.cpp:225-251: two float accumulators, forward edges then backward edges */
static inline float orc_connections(const orc_klgraph *g, const uint8_t *side, int32_t v) {
  float external = 0.0f, internal = 0.0f;
  for (int64_t i = g->rowptr[v]; i < g->rowptr[v + 1]; ++i) {
    if (side[g->col[i]] == 0) internal += g->w[i];
    else                      external += g->w[i];
  }
  return external - internal;
}
void orc_kl_dvalues(const orc_klgraph *g, const uint8_t *side, float *val) {   /* cKL.cpp:318-321 */
#pragma omp parallel for schedule(dynamic, 256)
  for (int32_t v = 0; v < g->n; ++v) val[v] = orc_connections(g, side, v);
}

typedef struct { int64_t rank; float w; } rank_w;
static int cmp_rank(const void *a, const void *b) {
  int64_t x = ((const rank_w *)a)->rank, y = ((const rank_w *)b)->rank;
  return (x > y) - (x < y);
}
/* calCutSize(), cKL.cpp:199-223, on one thread: a single float accumulator over remain[0] in order;
 * per node: forward edges in map order, then backward edges in the iteration order of
 * unordered_set<uint32_t>(remain[1].begin(), remain[1].end()) (GCC 13: element-by-element insert) */
float orc_kl_cut0(const orc_klgraph *g, const uint8_t *side,
                  const int32_t *order0, int64_t n0, const int32_t *order1, int64_t n1) {
  (void)side;
  int64_t *rank = (int64_t *)malloc((size_t)(g->n ? g->n : 1) * sizeof(int64_t));
  for (int32_t v = 0; v < g->n; ++v) rank[v] = -1;               /* -1: not in rightNodes */
  {
    int64_t *ord = (int64_t *)malloc((size_t)(n1 ? n1 : 1) * sizeof(int64_t));
    orc_stl_hash_order((const uint32_t *)order1, n1, ord);
    for (int64_t i = 0; i < n1; ++i) rank[order1[ord[i]]] = i;
    free(ord);
  }
  int64_t maxdeg = 0;
  for (int32_t v = 0; v < g->n; ++v)
    if (g->rowptr[v + 1] - g->rowptr[v] > maxdeg) maxdeg = g->rowptr[v + 1] - g->rowptr[v];
  rank_w *tmp = (rank_w *)malloc((size_t)(maxdeg ? maxdeg : 1) * sizeof(rank_w));
  float cut = 0.0f;
  for (int64_t i = 0; i < n0; ++i) {
    int32_t v = order0[i];
    for (int64_t e = g->rowptr[v]; e < g->fwd_end[v]; ++e)
      if (rank[g->col[e]] >= 0) cut += g->w[e];                   /* cKL.cpp:207-211 */
    int64_t m = 0;
    for (int64_t e = g->fwd_end[v]; e < g->rowptr[v + 1]; ++e)
      if (rank[g->col[e]] >= 0) { tmp[m].rank = rank[g->col[e]]; tmp[m].w = g->w[e]; ++m; }
    qsort(tmp, (size_t)m, sizeof(rank_w), cmp_rank);
    for (int64_t j = 0; j < m; ++j) cut += tmp[j].w;              /* cKL.cpp:213-220 */
  }
  free(tmp); free(rank);
  return cut;
}

/* getEdgeWeight(), cKL.cpp:75-82 */
static float orc_edge_weight(const orc_klgraph *g, int32_t a, int32_t b) {
  if (a > b) { int32_t t = a; a = b; b = t; }
  for (int64_t e = g->rowptr[a]; e < g->fwd_end[a]; ++e)
    if (g->col[e] == b) return g->w[e];
  return 0.0f;
}

/* KL(), cKL.cpp:288-390: one pass, no rollback.
 *
 * Pair selection (cKL.cpp:337-355) is "first strictly greatest val over remain[0] in order" / "first
 * strictly smallest over remain[1]".  orc_kl_run_linear does literally that, O(|remain|) per swap -- fine up
 * to ibm10, hopeless at 2 M nodes (5e11 comparisons per pass).  orc_kl_run keeps, per block of ORC_SEL_BLOCK
 * consecutive POSITIONS of remain[s], the first best element of the block and rescans a block only when one of
 * its nodes changed value or was locked since the last selection; the winner is the first block (in position
 * order) holding the strictly best block value -- the same element the linear scan returns, by construction
 * (tests/test_oracle.py checks the two against each other and against the reference's traces).            */
#define ORC_SEL_BLOCK 256
typedef struct {
  const int32_t *order; int64_t n; int64_t nb;
  float *bval; int64_t *bidx; uint8_t *dirty;
} orc_sel;
static void sel_init(orc_sel *s, const int32_t *order, int64_t n) {
  s->order = order; s->n = n; s->nb = (n + ORC_SEL_BLOCK - 1) / ORC_SEL_BLOCK;
  s->bval = (float *)malloc((size_t)(s->nb ? s->nb : 1) * sizeof(float));
  s->bidx = (int64_t *)malloc((size_t)(s->nb ? s->nb : 1) * sizeof(int64_t));
  s->dirty = (uint8_t *)malloc((size_t)(s->nb ? s->nb : 1));
  memset(s->dirty, 1, (size_t)(s->nb ? s->nb : 1));
}
static void sel_free(orc_sel *s) { free(s->bval); free(s->bidx); free(s->dirty); }
/* want_max: first strictly greatest; else first strictly smallest.  Returns the position or -1. */
static int64_t sel_pick(orc_sel *s, const float *val, const uint8_t *locked, int want_max, float *best_out) {
  float best = want_max ? -FLT_MAX : FLT_MAX;
  int64_t best_i = -1;
  for (int64_t b = 0; b < s->nb; ++b) {
    if (s->dirty[b]) {
      float bv = want_max ? -FLT_MAX : FLT_MAX;
      int64_t bi = -1;
      const int64_t hi = (b + 1) * ORC_SEL_BLOCK < s->n ? (b + 1) * ORC_SEL_BLOCK : s->n;
      for (int64_t i = b * ORC_SEL_BLOCK; i < hi; ++i) {
        const int32_t v = s->order[i];
        if (locked[v]) continue;
        if (want_max ? (val[v] > bv) : (val[v] < bv)) { bv = val[v]; bi = i; }
      }
      s->bval[b] = bv; s->bidx[b] = bi; s->dirty[b] = 0;
    }
    if (s->bidx[b] >= 0 && (want_max ? (s->bval[b] > best) : (s->bval[b] < best))) { best = s->bval[b]; best_i = s->bidx[b]; }
  }
  *best_out = best;
  return best_i;
}

static int64_t orc_kl_run_impl(const orc_klgraph *g, uint8_t *side,
                               const int32_t *order0, int64_t n0, const int32_t *order1, int64_t n1,
                               float *cut, float *gain, int32_t *node1, int32_t *node2, int64_t capacity, int linear) {
  const int32_t n = g->n;
  float *val = (float *)malloc((size_t)(n ? n : 1) * sizeof(float));
  uint8_t *locked = (uint8_t *)calloc((size_t)(n ? n : 1), 1);
  int64_t *pos = NULL;                                             /* position of a node in its remain[] list */
  orc_sel s0, s1;
  if (!linear) {
    pos = (int64_t *)malloc((size_t)(n ? n : 1) * sizeof(int64_t));
    for (int64_t i = 0; i < n0; ++i) pos[order0[i]] = i;
    for (int64_t i = 0; i < n1; ++i) pos[order1[i]] = i;
    sel_init(&s0, order0, n0); sel_init(&s1, order1, n1);
  }
  uint32_t terminate = 0;
  uint32_t terminateLimit = (uint32_t)log2((double)n) + 5;        /* cKL.cpp:303 */
  float cutSize = orc_kl_cut0(g, side, order0, n0, order1, n1);   /* cKL.cpp:306 */
  int64_t it = 0;
  cut[0] = cutSize; gain[0] = 0.0f; node1[0] = -1; node2[0] = -1; /* row 0, cKL.cpp:315 */
  orc_kl_dvalues(g, side, val);                                   /* cKL.cpp:318-321 */
  int64_t rem0 = n0, rem1 = n1;
  int64_t lo0 = 0, lo1 = 0;                                        /* first possibly-unlocked index */
  while (rem0 > 0 && rem1 > 0) {                                   /* cKL.cpp:334 */
    float maxGain = -FLT_MAX, minGain = FLT_MAX;
    int64_t maxIdx = -1, minIdx = -1;
    if (linear) {
      while (lo0 < n0 && locked[order0[lo0]]) ++lo0;
      while (lo1 < n1 && locked[order1[lo1]]) ++lo1;
      for (int64_t i = lo0; i < n0; ++i) {                         /* cKL.cpp:341-347 */
        int32_t v = order0[i];
        if (!locked[v] && val[v] > maxGain) { maxGain = val[v]; maxIdx = i; }
      }
      for (int64_t i = lo1; i < n1; ++i) {                         /* cKL.cpp:349-355 */
        int32_t v = order1[i];
        if (!locked[v] && val[v] < minGain) { minGain = val[v]; minIdx = i; }
      }
    } else {
      maxIdx = sel_pick(&s0, val, locked, 1, &maxGain);
      minIdx = sel_pick(&s1, val, locked, 0, &minGain);
    }
    if (maxIdx < 0 || minIdx < 0) break;                           /* cKL.cpp:387-389 */
    int32_t a = order0[maxIdx], b = order1[minIdx];
    float gn = maxGain - minGain - 2.0f * orc_edge_weight(g, a, b);   /* cKL.cpp:360 */
    cutSize -= gn;                                                 /* cKL.cpp:362 */
    locked[a] = 1; locked[b] = 1; --rem0; --rem1;                  /* swip, cKL.cpp:274-286 */
    side[a] = 1; side[b] = 0;
    if (!linear) { s0.dirty[maxIdx / ORC_SEL_BLOCK] = 1; s1.dirty[minIdx / ORC_SEL_BLOCK] = 1; }
    /* updateAffectedNodeGains, cKL.cpp:253-272: recompute N(a) u N(b) from scratch (incl. locked).
     * A node keeps its place in the remain[] list of the side it STARTED on (a and b are erased, nobody moves). */
    for (int pass = 0; pass < 2; ++pass) {
      const int32_t u = pass ? b : a;
      for (int64_t e = g->rowptr[u]; e < g->rowptr[u + 1]; ++e) {
        const int32_t v = g->col[e];
        val[v] = orc_connections(g, side, v);
        if (!linear && v != a && v != b) {
          /* which list holds v: the side it had before this pass started = current side unless it was swapped
           * (swapped nodes are locked and never selected again, so their block need not be refreshed) */
          if (!locked[v]) { if (side[v] == 0) s0.dirty[pos[v] / ORC_SEL_BLOCK] = 1; else s1.dirty[pos[v] / ORC_SEL_BLOCK] = 1; }
        }
      }
    }
    ++it;
    if (it < capacity) { cut[it] = cutSize; gain[it] = gn; node1[it] = a; node2[it] = b; }
    if (gn <= 0.0f) { if (++terminate > terminateLimit) break; }  /* cKL.cpp:382-386 */
    else terminate = 0;
  }
  if (!linear) { sel_free(&s0); sel_free(&s1); free(pos); }
  free(val); free(locked);
  return it;
}
int64_t orc_kl_run(const orc_klgraph *g, uint8_t *side,
                   const int32_t *order0, int64_t n0, const int32_t *order1, int64_t n1,
                   float *cut, float *gain, int32_t *node1, int32_t *node2, int64_t capacity) {
  return orc_kl_run_impl(g, side, order0, n0, order1, n1, cut, gain, node1, node2, capacity, 0);
}
int64_t orc_kl_run_linear(const orc_klgraph *g, uint8_t *side,
                          const int32_t *order0, int64_t n0, const int32_t *order1, int64_t n1,
                          float *cut, float *gain, int32_t *node1, int32_t *node2, int64_t capacity) {
  return orc_kl_run_impl(g, side, order0, n0, order1, n1, cut, gain, node1, node2, capacity, 1);
}

/* ================================================================================================
 * EIG -- cEIG.cpp:86-133 (matrix), 194-207 (solve), 55-65 (median), 213-220 (file)
 * ============================================================================================== */
typedef struct { int32_t c; double v; } cv_t;
typedef struct { int32_t c; int64_t s; double v; } cs_t;
static int cmp_cs(const void *a, const void *b) {
  const cs_t *x = (const cs_t *)a, *y = (const cs_t *)b;
  if (x->c != y->c) return (x->c > y->c) - (x->c < y->c);
  return (x->s > y->s) - (x->s < y->s);
}
/* L = D - A, A_ij = sum over nets containing i,j of 2.0/|net| ; L_ii = -sum_j L_ij (cEIG.cpp:110-130) */
int orc_laplacian(const orc_hgr *h, orc_csr *L) {
  memset(L, 0, sizeof(*L));
  const int32_t n = h->n_nodes;
  int64_t *cnt = (int64_t *)calloc((size_t)n + 2, sizeof(int64_t));
  for (int32_t e = 0; e < h->n_nets; ++e) {
    int64_t k = h->net_off[e + 1] - h->net_off[e];
    if (k < 2) continue;
    for (int64_t j = 0; j < k; ++j) cnt[h->pins[h->net_off[e] + j] + 1] += k - 1;
  }
  for (int32_t v = 0; v < n; ++v) cnt[v + 1] += cnt[v];
  int64_t T = cnt[n];
  cv_t *tr = (cv_t *)malloc((size_t)(T ? T : 1) * sizeof(cv_t));
  int64_t *cur = (int64_t *)malloc((size_t)(n ? n : 1) * sizeof(int64_t));
  memcpy(cur, cnt, (size_t)n * sizeof(int64_t));
  for (int32_t e = 0; e < h->n_nets; ++e) {
    const int32_t *p = h->pins + h->net_off[e];
    int64_t k = h->net_off[e + 1] - h->net_off[e];
    if (k < 2) continue;
    double w = 2.0 / (double)k;                                  /* cEIG.cpp:110 */
    for (int64_t j = 0; j < k; ++j)
      for (int64_t l = j + 1; l < k; ++l) {
        if (p[j] == p[l]) { free(tr); free(cur); free(cnt); return -2; }
        tr[cur[p[j]]].c = p[l]; tr[cur[p[j]]].v = -w; cur[p[j]]++;   /* cEIG.cpp:114-115 */
        tr[cur[p[l]]].c = p[j]; tr[cur[p[l]]].v = -w; cur[p[l]]++;
      }
  }
  L->n = n;
  L->rowptr = (int64_t *)malloc(((size_t)n + 1) * sizeof(int64_t));
  L->col = (int32_t *)malloc((size_t)(T + n + 1) * sizeof(int32_t));
  L->val = (double *)malloc((size_t)(T + n + 1) * sizeof(double));
  int64_t nz = 0;
  L->rowptr[0] = 0;
  for (int32_t v = 0; v < n; ++v) {
    int64_t lo = cnt[v], hi = cnt[v + 1];
    /* stable by column so duplicates add in file order (setFromTriplets sums duplicates) */
    if (hi - lo <= 64) {
      for (int64_t i = lo + 1; i < hi; ++i) {                     /* insertion sort is stable */
        cv_t x = tr[i]; int64_t j = i;
        while (j > lo && tr[j - 1].c > x.c) { tr[j] = tr[j - 1]; --j; }
        tr[j] = x;
      }
    } else {                                                      /* long rows: qsort on (col,seq) */
      cs_t *t2 = (cs_t *)malloc((size_t)(hi - lo) * sizeof(cs_t));
      for (int64_t i = lo; i < hi; ++i) { t2[i - lo].c = tr[i].c; t2[i - lo].s = i; t2[i - lo].v = tr[i].v; }
      qsort(t2, (size_t)(hi - lo), sizeof(cs_t), cmp_cs);
      for (int64_t i = lo; i < hi; ++i) { tr[i].c = t2[i - lo].c; tr[i].v = t2[i - lo].v; }
      free(t2);
    }
    double rowsum = 0.0;
    int64_t row0 = nz, diag_at = -1;
    for (int64_t i = lo; i < hi;) {
      int32_t c = tr[i].c; double s = 0.0;
      while (i < hi && tr[i].c == c) { s += tr[i].v; ++i; }
      if (diag_at < 0 && c > v) { diag_at = nz; L->col[nz] = v; L->val[nz] = 0.0; ++nz; }
      L->col[nz] = c; L->val[nz] = s; ++nz;
      rowsum += s;
    }
    if (diag_at < 0) { diag_at = nz; L->col[nz] = v; L->val[nz] = 0.0; ++nz; }
    L->val[diag_at] = -rowsum;                                    /* cEIG.cpp:127-130 */
    (void)row0;
    L->rowptr[v + 1] = nz;
  }
  free(tr); free(cur); free(cnt);
  return 0;
}
void orc_csr_free(orc_csr *L) {
  free(L->rowptr); free(L->col); free(L->val);
  memset(L, 0, sizeof(*L));
}
void orc_spmv(const orc_csr *L, const double *x, double *y) {
#pragma omp parallel for schedule(static)
  for (int32_t v = 0; v < L->n; ++v) {
    double s = 0.0;
    for (int64_t e = L->rowptr[v]; e < L->rowptr[v + 1]; ++e) s += L->val[e] * x[L->col[e]];
    y[v] = s;
  }
}

/* ---- dense symmetric eigen-solver: Householder tridiagonalisation + implicit-shift QL ----------
 * (textbook EISPACK tred2/tql2 scheme; eigenvalues ascending, eigenvectors in the columns of a)  */
void orc_sym_eig(int n, double *a, double *d) {
  double *e = (double *)calloc((size_t)n + 1, sizeof(double));
#define A(i, j) a[(size_t)(i) * n + (j)]
  for (int j = 0; j < n; ++j) d[j] = A(n - 1, j);
  for (int i = n - 1; i > 0; --i) {
    double scale = 0.0, hh = 0.0;
    for (int k = 0; k < i; ++k) scale += fabs(d[k]);
    if (scale == 0.0) {
      e[i] = d[i - 1];
      for (int j = 0; j < i; ++j) { d[j] = A(i - 1, j); A(i, j) = 0.0; A(j, i) = 0.0; }
    } else {
      for (int k = 0; k < i; ++k) { d[k] /= scale; hh += d[k] * d[k]; }
      double f = d[i - 1], g = sqrt(hh);
      if (f > 0) g = -g;
      e[i] = scale * g; hh -= f * g; d[i - 1] = f - g;
      for (int j = 0; j < i; ++j) e[j] = 0.0;
      for (int j = 0; j < i; ++j) {
        f = d[j]; A(j, i) = f; g = e[j] + A(j, j) * f;
        for (int k = j + 1; k <= i - 1; ++k) { g += A(k, j) * d[k]; e[k] += A(k, j) * f; }
        e[j] = g;
      }
      f = 0.0;
      for (int j = 0; j < i; ++j) { e[j] /= hh; f += e[j] * d[j]; }
      double hk = f / (hh + hh);
      for (int j = 0; j < i; ++j) e[j] -= hk * d[j];
      for (int j = 0; j < i; ++j) {
        f = d[j]; g = e[j];
        for (int k = j; k <= i - 1; ++k) A(k, j) -= (f * e[k] + g * d[k]);
        d[j] = A(i - 1, j); A(i, j) = 0.0;
      }
    }
    d[i] = hh;
  }
  for (int i = 0; i < n - 1; ++i) {
    A(n - 1, i) = A(i, i); A(i, i) = 1.0;
    double hh = d[i + 1];
    if (hh != 0.0) {
      for (int k = 0; k <= i; ++k) d[k] = A(k, i + 1) / hh;
      for (int j = 0; j <= i; ++j) {
        double g = 0.0;
        for (int k = 0; k <= i; ++k) g += A(k, i + 1) * A(k, j);
        for (int k = 0; k <= i; ++k) A(k, j) -= g * d[k];
      }
    }
    for (int k = 0; k <= i; ++k) A(k, i + 1) = 0.0;
  }
  for (int j = 0; j < n; ++j) { d[j] = A(n - 1, j); A(n - 1, j) = 0.0; }
  A(n - 1, n - 1) = 1.0; e[0] = 0.0;
  /* QL */
  for (int i = 1; i < n; ++i) e[i - 1] = e[i];
  e[n - 1] = 0.0;
  double f = 0.0, tst1 = 0.0, eps = ldexp(1.0, -52);
  for (int l = 0; l < n; ++l) {
    double t = fabs(d[l]) + fabs(e[l]);
    if (t > tst1) tst1 = t;
    int m = l;
    while (m < n) { if (fabs(e[m]) <= eps * tst1) break; ++m; }
    if (m > l) {
      int iter = 0;
      do {
        ++iter;
        double g = d[l], p = (d[l + 1] - g) / (2.0 * e[l]), r = hypot(p, 1.0);
        if (p < 0) r = -r;
        d[l] = e[l] / (p + r); d[l + 1] = e[l] * (p + r);
        double dl1 = d[l + 1], h = g - d[l];
        for (int i = l + 2; i < n; ++i) d[i] -= h;
        f += h;
        p = d[m];
        double c = 1.0, c2 = c, c3 = c, el1 = e[l + 1], s = 0.0, s2 = 0.0;
        for (int i = m - 1; i >= l; --i) {
          c3 = c2; c2 = c; s2 = s;
          g = c * e[i]; h = c * p; r = hypot(p, e[i]);
          e[i + 1] = s * r; s = e[i] / r; c = p / r; p = c * d[i] - s * g;
          d[i + 1] = h + s * (c * g + s * d[i]);
          for (int k = 0; k < n; ++k) {
            h = A(k, i + 1);
            A(k, i + 1) = s * A(k, i) + c * h;
            A(k, i) = c * A(k, i) - s * h;
          }
        }
        p = -s * s2 * c3 * el1 * e[l] / dl1;
        e[l] = s * p; d[l] = c * p;
      } while (fabs(e[l]) > eps * tst1 && iter < 200);
    }
    d[l] += f; e[l] = 0.0;
  }
  for (int i = 0; i < n - 1; ++i) {       /* sort ascending */
    int k = i; double p = d[i];
    for (int j = i + 1; j < n; ++j) if (d[j] < p) { k = j; p = d[j]; }
    if (k != i) {
      d[k] = d[i]; d[i] = p;
      for (int j = 0; j < n; ++j) { double t = A(j, i); A(j, i) = A(j, k); A(j, k) = t; }
    }
  }
#undef A
  free(e);
}

/* ---- thick-restart Lanczos with full (twice) Gram-Schmidt re-orthogonalisation ------------------
 * Restates the solve cEIG.cpp:194-198 asks Spectra for: nev = 2 algebraically smallest eigenpairs,
 * ncv = min(100, n/2), tolerance 1e-10, at most 1000 restarts; converged when, for both wanted Ritz
 * pairs, |beta_m * y_last| < tol * max(eps^(2/3), |theta|).  Restart keeps
 * nev + min(nconv, (ncv-nev)/2) Ritz vectors (the same count an implicitly restarted Lanczos with
 * exact shifts retains).  cEIG.cpp:205-207 reports the LARGER of the two (lambda2).             */
static double orc_dot(const double *a, const double *b, int32_t n) {
  double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (int32_t i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}
static int orc_fiedler_impl(const orc_csr *L, double *lambda2, double *vec, orc_eig_stats *st, int maxit_override);
int orc_fiedler(const orc_csr *L, double *lambda2, double *vec, orc_eig_stats *st) {
  return orc_fiedler_impl(L, lambda2, vec, st, 0);
}
/* bounded sample for bench.py's cpu_baseline: stops after max_restarts restart cycles (st->converged
 * tells whether that was enough); st->matvecs is the work actually done */
int orc_fiedler_bounded(const orc_csr *L, int max_restarts, double *lambda2, double *vec, orc_eig_stats *st) {
  return orc_fiedler_impl(L, lambda2, vec, st, max_restarts);
}
static int orc_fiedler_impl(const orc_csr *L, double *lambda2, double *vec, orc_eig_stats *st, int maxit_override) {
  const int32_t n = L->n;
  const int nev = 2;
  int m = n / 2 < 100 ? n / 2 : 100;                               /* cEIG.cpp:195 */
  if (m <= nev || m > n) return -3;
  const double tol = 1e-10, eps23 = pow(DBL_EPSILON, 2.0 / 3.0);
  const int maxit = maxit_override > 0 ? maxit_override : 1000;
  double *V = (double *)malloc((size_t)n * (size_t)(m + 1) * sizeof(double));
  double *T = (double *)calloc((size_t)m * m, sizeof(double));
  double *Y = (double *)malloc((size_t)m * m * sizeof(double));
  double *th = (double *)malloc((size_t)m * sizeof(double));
  double *h = (double *)malloc((size_t)(m + 1) * sizeof(double));
  double *h2 = (double *)malloc((size_t)(m + 1) * sizeof(double));
  double *w = (double *)malloc((size_t)n * sizeof(double));
  double *Vk = (double *)malloc((size_t)n * (size_t)(m) * sizeof(double));
#ifdef _OPENMP
  const int max_threads = omp_get_max_threads();
#else
  const int max_threads = 1;
#endif
  double *part = (double *)malloc((size_t)max_threads * (size_t)(m + 1) * sizeof(double));
#define VC(j) (V + (size_t)(j) * n)
  /* deterministic start vector: splitmix64 uniform in [-0.5, 0.5) */
  uint64_t sm = 0x9E3779B97F4A7C15ull;
  for (int32_t i = 0; i < n; ++i) {
    sm += 0x9E3779B97F4A7C15ull;
    uint64_t z = sm;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    VC(0)[i] = (double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5;
  }
  double nrm = sqrt(orc_dot(VC(0), VC(0), n));
  for (int32_t i = 0; i < n; ++i) VC(0)[i] /= nrm;
  int k = 0, it, nmv = 0, done = 0;
  double beta = 0.0, res[2] = {0, 0};
  for (it = 0; it < maxit && !done; ++it) {
    for (int j = k; j < m; ++j) {
      orc_spmv(L, VC(j), w); ++nmv;
      for (int pass = 0; pass < 2; ++pass) {
        double *hh = pass ? h2 : h;
        const int nc = j + 1;
        /* hh = V^T w : every thread owns a contiguous row chunk (cache-blocked), partials folded in order */
#pragma omp parallel
        {
#ifdef _OPENMP
          const int nt = omp_get_num_threads(), tn = omp_get_thread_num();
#else
          const int nt = 1, tn = 0;
#endif
          const int64_t lo = (int64_t)n * tn / nt, hi = (int64_t)n * (tn + 1) / nt;
          double *pp = part + (size_t)tn * (m + 1);
          for (int c = 0; c < nc; ++c) pp[c] = 0.0;
          for (int64_t b0 = lo; b0 < hi; b0 += 2048) {
            const int64_t b1 = b0 + 2048 < hi ? b0 + 2048 : hi;
            for (int c = 0; c < nc; ++c) {
              const double *vc = VC(c);
              double s = 0.0;
              for (int64_t i = b0; i < b1; ++i) s += vc[i] * w[i];
              pp[c] += s;
            }
          }
#pragma omp barrier
#pragma omp for schedule(static)
          for (int c = 0; c < nc; ++c) {
            double s = 0.0;
            for (int t = 0; t < nt; ++t) s += part[(size_t)t * (m + 1) + c];
            hh[c] = s;
          }
          /* w -= V hh */
          for (int64_t b0 = lo; b0 < hi; b0 += 2048) {
            const int64_t b1 = b0 + 2048 < hi ? b0 + 2048 : hi;
            for (int c = 0; c < nc; ++c) {
              const double *vc = VC(c);
              const double hc = hh[c];
              for (int64_t i = b0; i < b1; ++i) w[i] -= vc[i] * hc;
            }
          }
        }
      }
      T[(size_t)j * m + j] = h[j] + h2[j];
      beta = sqrt(orc_dot(w, w, n));
      for (int32_t i = 0; i < n; ++i) VC(j + 1)[i] = w[i] / beta;
      if (j + 1 < m) { T[(size_t)j * m + j + 1] = beta; T[(size_t)(j + 1) * m + j] = beta; }
    }
    memcpy(Y, T, (size_t)m * m * sizeof(double));
    orc_sym_eig(m, Y, th);
    int nconv = 0;
    for (int i = 0; i < nev; ++i) {
      res[i] = fabs(beta * Y[(size_t)(m - 1) * m + i]);
      double thr = tol * (fabs(th[i]) > eps23 ? fabs(th[i]) : eps23);
      if (res[i] < thr) ++nconv;
    }
    if (nconv == nev || it == maxit - 1) { done = 1; break; }
    int kk = nev + (nconv < (m - nev) / 2 ? nconv : (m - nev) / 2);
    if (kk > m - 1) kk = m - 1;
    /* V[:, 0:kk] = V[:, 0:m] * Y[:, 0:kk] ; V[:, kk] = v_{m} */
#pragma omp parallel for schedule(static)
    for (int64_t b0 = 0; b0 < n; b0 += 1024) {
      const int64_t b1 = b0 + 1024 < n ? b0 + 1024 : n;
      for (int c = 0; c < kk; ++c) {
        double *out = Vk + (size_t)c * n;
        for (int64_t i = b0; i < b1; ++i) out[i] = 0.0;
        for (int j = 0; j < m; ++j) {
          const double y = Y[(size_t)j * m + c];
          const double *vj = VC(j);
          for (int64_t i = b0; i < b1; ++i) out[i] += vj[i] * y;
        }
      }
    }
    memcpy(V, Vk, (size_t)n * kk * sizeof(double));
    memcpy(VC(kk), VC(m), (size_t)n * sizeof(double));
    memset(T, 0, (size_t)m * m * sizeof(double));
    for (int c = 0; c < kk; ++c) {
      T[(size_t)c * m + c] = th[c];
      double s = beta * Y[(size_t)(m - 1) * m + c];
      T[(size_t)kk * m + c] = s; T[(size_t)c * m + kk] = s;
    }
    k = kk;
  }
  /* Ritz vector of the larger wanted value (index 1), normalised */
#pragma omp parallel for schedule(static)
  for (int32_t i = 0; i < n; ++i) {
    double s = 0.0;
    for (int j = 0; j < m; ++j) s += VC(j)[i] * Y[(size_t)j * m + 1];
    vec[i] = s;
  }
  nrm = sqrt(orc_dot(vec, vec, n));
  for (int32_t i = 0; i < n; ++i) vec[i] /= nrm;
  *lambda2 = th[1];
  if (st) {
    st->matvecs = nmv; st->restarts = it + 1; st->ncv = m;
    st->converged = (res[0] < tol * (fabs(th[0]) > eps23 ? fabs(th[0]) : eps23)) &&
                    (res[1] < tol * (fabs(th[1]) > eps23 ? fabs(th[1]) : eps23));
    st->resid_est[0] = res[0]; st->resid_est[1] = res[1];
  }
#undef VC
  free(V); free(T); free(Y); free(th); free(h); free(h2); free(w); free(Vk); free(part);
  return 0;
}

static int cmp_dbl(const void *a, const void *b) {
  double x = *(const double *)a, y = *(const double *)b;
  return (x > y) - (x < y);
}
double orc_median(const double *v, int32_t n) {                    /* cEIG.cpp:55-65 */
  if (n <= 0) return 0.0;
  double *s = (double *)malloc((size_t)n * sizeof(double));
  memcpy(s, v, (size_t)n * sizeof(double));
  qsort(s, (size_t)n, sizeof(double), cmp_dbl);
  double r = (n % 2 != 0) ? s[n / 2] : (s[(n - 1) / 2] + s[n / 2]) / 2.0;
  free(s);
  return r;
}
int orc_write_eig(const char *path, double lambda2, const double *vec, int32_t n) {  /* cEIG.cpp:213-220 */
  FILE *f = fopen(path, "w");
  if (!f) return -1;
  double med = orc_median(vec, n);
  fprintf(f, "%.12g\n%.12g\n", lambda2, med);
  for (int32_t i = 0; i < n; ++i) fprintf(f, "%d\t%d\t%.12g\n", i, (med > vec[i]) ? 1 : 0, vec[i]);
  fclose(f);
  return 0;
}
int orc_read_eig_sides(const char *path, int32_t n, uint8_t *side, double *lambda2, double *median, double *vec) {
  FILE *f = fopen(path, "r");                                       /* cKL.cpp:155-174 */
  if (!f) return -1;
  double l = 0, m = 0;
  if (fscanf(f, "%lf %lf", &l, &m) != 2) { fclose(f); return -2; }
  if (lambda2) *lambda2 = l;
  if (median) *median = m;
  int32_t cnt = 0;
  long node; int s; double wv;
  while (fscanf(f, "%ld %d %lf", &node, &s, &wv) == 3) {
    if (node < 0 || node >= n || (s != 0 && s != 1)) { fclose(f); return -2; }
    side[node] = (uint8_t)s;
    if (vec) vec[node] = wv;
    ++cnt;
  }
  fclose(f);
  return cnt == n ? 0 : -2;
}
int orc_write_trace(const char *path, const float *cut, const float *gain, int64_t swaps) {
  FILE *f = fopen(path, "w");                                       /* cKL.cpp:315,380 */
  if (!f) return -1;
  fprintf(f, "0\t%g\t0\n", (double)cut[0]);
  for (int64_t i = 1; i <= swaps; ++i) fprintf(f, "%lld\t%g\t%g\n", (long long)i, (double)cut[i], (double)gain[i]);
  fclose(f);
  return 0;
}
