// stl_order.h -- iteration order of libstdc++'s std::unordered_{map,set}<uint32_t> (GCC 13), as a
// pure function of the insertion sequence.  Host + device.
//
// Why the product needs this: cKL sums each node's forward edge weights in the iteration order of a
// std::unordered_map<uint32_t,float> (cKL.cpp:230-236, 207-211) and the backward cut edges in the
// iteration order of a std::unordered_set<uint32_t> (cKL.cpp:201,213).  fp32 addition is not
// associative, so the KL gains -- and from the 5th swap on, the swap sequence itself -- depend on
// that order.  Rules (bits/hashtable.h, bits/hashtable_policy.h):
//   * hash(k) = k, bucket = k % bucket_count, max load factor 1;
//   * bucket counts for one-by-one inserts follow the chain 1 -> 13 -> 29 -> 59 -> 127 -> ...
//     (next listed prime >= 2 * count); the rehash happens BEFORE inserting element count+1 when
//     count + 1 > bucket_count;
//   * a node entering an empty bucket is put at the FRONT of the global singly linked list; a node
//     entering a non-empty bucket is put at the front of that bucket's run;
//   * a rehash walks the old list front to back and re-inserts every node with the same two rules.
// Consequence used by the parallel version (kl.cu): after processing a sequence into B buckets the
// list is the buckets in DESCENDING order of the time they were first hit, each bucket's nodes in
// DESCENDING order of processing time.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define EIGKL_HD __host__ __device__ __forceinline__
#else
#define EIGKL_HD inline
#endif

namespace eigkl {

constexpr int STL_CHAIN_LEN = 23;

EIGKL_HD uint32_t stl_bucket_chain(int level) {
  constexpr uint32_t chain[STL_CHAIN_LEN] = {13u, 29u, 59u, 127u, 257u, 541u, 1109u, 2357u, 5087u, 10273u, 20753u,
                                             42043u, 85229u, 172933u, 351061u, 712697u, 1447153u, 2938679u,
                                             5967347u, 12117689u, 24607243u, 49969847u, 101473717u};
  return chain[level];
}

// number of buckets the table ends with after n one-by-one inserts (1 for the empty table)
EIGKL_HD uint32_t stl_final_buckets(uint32_t n) {
  if (n == 0) return 1u;
  int lv = 0;
  while (lv < STL_CHAIN_LEN - 1 && stl_bucket_chain(lv) < n) ++lv;
  return stl_bucket_chain(lv);
}

constexpr int32_t STL_EMPTY = -2;   // bucket has no node
constexpr int32_t STL_BB = -1;      // bucket's before-node is the list head sentinel

// Replays n inserts of distinct keys key(0..n-1).  next[] needs n entries, bkt[] stl_final_buckets(n).
// On return the iteration order is head, next[head], ... (-1 terminated).
template <typename KeyFn>
EIGKL_HD void stl_replay_inserts(int32_t n, KeyFn key, int32_t *next, int32_t *bkt, int32_t &head) {
  head = -1;
  uint32_t nb = 1;
  int level = -1;
  bkt[0] = STL_EMPTY;
  for (int32_t i = 0; i < n; ++i) {
    const uint32_t limit = (level < 0) ? 0u : nb;          // _M_next_resize
    if ((uint32_t)i + 1u > limit) {                        // rehash before the insert
      ++level;
      nb = stl_bucket_chain(level);
      for (uint32_t b = 0; b < nb; ++b) bkt[b] = STL_EMPTY;
      int32_t p = head;
      head = -1;
      uint32_t bbegin = 0;
      while (p >= 0) {
        const int32_t nx = next[p];
        const uint32_t b = key(p) % nb;
        if (bkt[b] == STL_EMPTY) {
          next[p] = head;
          head = p;
          bkt[b] = STL_BB;
          if (next[p] >= 0) bkt[bbegin] = p;
          bbegin = b;
        } else {
          const int32_t prev = bkt[b];
          if (prev == STL_BB) { next[p] = head; head = p; }
          else                { next[p] = next[prev]; next[prev] = p; }
        }
        p = nx;
      }
    }
    const uint32_t b = key(i) % nb;
    if (bkt[b] == STL_EMPTY) {
      next[i] = head;
      head = i;
      if (next[i] >= 0) bkt[key(next[i]) % nb] = i;
      bkt[b] = STL_BB;
    } else {
      const int32_t prev = bkt[b];
      if (prev == STL_BB) { next[i] = head; head = i; }
      else                { next[i] = next[prev]; next[prev] = i; }
    }
  }
}

}  // namespace eigkl
