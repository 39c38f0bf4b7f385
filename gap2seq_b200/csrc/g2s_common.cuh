// g2s_common.cuh - k-mer arithmetic, the global k-mer table and the portability shims that let the per-gap
// algorithm (g2s_fill.cuh) compile both as sm_100a device code and, for debugging in tests/ only
// (-DG2S_HOST_EMUL, group size 1), as plain host C++.  The product always runs the device build.
//
// Encoding follows GATB (reference paths relative to thirdparty/gatb-core/gatb-core/src/gatb):
//   ASCII -> 2 bit: (c >> 1) & 3, i.e. A=0 C=1 T=2 G=3              tools/misc/api/Data.hpp:176
//   a k-mer is the big-endian 2k-bit number of its bases              tools/math/LargeInt1.pri:119-133
//   canonical = numeric min(forward, reverse complement)              kmer/impl/Model.hpp:224,294
//   strand FORWARD (0) iff the oriented k-mer is the canonical one    debruijn/impl/Graph.cpp:1742
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__) && !defined(G2S_HOST_EMUL)
#define G2S_DEVICE_BUILD 1
#define G2S_HD __host__ __device__ __forceinline__
#define G2S_D __device__ __forceinline__
#else
#define G2S_HD inline
#define G2S_D inline
#endif

namespace g2s {

static const uint64_t KEY_EMPTY = ~0ULL;
static const uint32_t NONE32 = 0xFFFFFFFFu;
static const uint32_t MAX_PATHS = 2147483647u / 2u - 1u;   // src/Gap2Seq.cpp:38
static const int MAX_FUZ = 31;

// 16-byte slot of the global open-addressing table: one 32-byte sector holds two slots, so a probe that
// hits in the first or second position costs a single sector gather.
struct __attribute__((aligned(16))) Slot {
  uint64_t key;   // canonical k-mer, KEY_EMPTY when free
  uint64_t val;   // bits 0-7: adjacency byte (low nibble out A,C,T,G, high nibble in; canonical orientation)
                  // bits 8-35 / 36-63: skip hint of the canonical / reverse-complement orientation (see hint_of)
};

// Skip hints.  For an oriented k-mer x the hint lists the next j <= HINT_MAX bases of its UNIQUE forward extension:
// x and the j-1 nodes after it each have exactly one successor (the listed base); nothing is claimed about the node
// the last base leads to.  A walker that holds a hint can therefore emit j levels from ONE table gather - the
// per-level dependent gather of a chain of unitig-interior nodes disappears - while every node still becomes its
// own DP cell.  28 bits: length j in bits 24-27, the bases as a 2j-bit big-endian string in the low bits (first base
// highest, like a k-mer), so that the state i steps ahead is ((x << 2i) | (bases >> 2(j-i))) & kmask.
static const int HINT_MAX = 12;
G2S_HD uint32_t hint_of(uint64_t val, uint32_t strand) { return (uint32_t)(val >> (8 + 28 * strand)) & 0xFFFFFFFu; }

struct Table {
  const Slot *slots;
  uint64_t cap_mask;   // capacity - 1 (capacity is a power of two >= 2 * n)
  uint64_t kmask;      // (1 << 2k) - 1
  int k;
};

G2S_HD uint64_t mix64(uint64_t x) {
  x ^= x >> 31; x *= 0x7fb5d329728ea185ULL; x ^= x >> 27; x *= 0x81dadef4bc2dd44dULL; x ^= x >> 33;
  return x;
}

// reverse complement of a 2k-bit k-mer: reverse the 2-bit groups, complement = xor 2 (A<->T, C<->G)
G2S_HD uint64_t revcomp(uint64_t x, int k) {
#ifdef __CUDA_ARCH__
  uint64_t r = __brevll(x);
#else
  uint64_t r = x;
  r = ((r >> 1) & 0x5555555555555555ULL) | ((r & 0x5555555555555555ULL) << 1);
  r = ((r >> 2) & 0x3333333333333333ULL) | ((r & 0x3333333333333333ULL) << 2);
  r = ((r >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((r & 0x0F0F0F0F0F0F0F0FULL) << 4);
  r = __builtin_bswap64(r);
#endif
  // bits inside every 2-bit group are now swapped as well: undo that, then complement
  r = ((r >> 1) & 0x5555555555555555ULL) | ((r & 0x5555555555555555ULL) << 1);
  r ^= 0xAAAAAAAAAAAAAAAAULL;
  return r >> (64 - 2 * k);
}

struct Canon {
  uint64_t key;
  uint32_t strand;   // 0: oriented == canonical, 1: oriented == revcomp(canonical)
};

G2S_HD Canon canonical(uint64_t oriented, int k) {
  uint64_t rc = revcomp(oriented, k);
  Canon c;
  c.key = oriented < rc ? oriented : rc;
  c.strand = oriented < rc ? 0u : (oriented == rc ? 0u : 1u);
  return c;
}

// adjacency byte -> edge masks relative to the oriented k-mer (bit nt = base A,C,T,G appended / prepended).
// Reverse strand: appending nt to revcomp(c) is prepending (nt^2) to c, so nibbles swap and bits move by nt^2.
G2S_HD uint32_t swap2(uint32_t m) { return ((m & 3u) << 2) | ((m >> 2) & 3u); }
G2S_HD uint32_t out_mask(uint32_t adj, uint32_t strand) { return strand ? swap2(adj >> 4) : (adj & 15u); }
G2S_HD uint32_t in_mask(uint32_t adj, uint32_t strand) { return strand ? swap2(adj & 15u) : (adj >> 4); }

G2S_HD uint64_t succ_kmer(uint64_t oriented, uint32_t nt, const Table &t) { return ((oriented << 2) | nt) & t.kmask; }
G2S_HD uint64_t pred_kmer(uint64_t oriented, uint32_t nt, const Table &t) {
  return (oriented >> 2) | ((uint64_t)nt << (2 * (t.k - 1)));
}

G2S_D Slot load_slot(const Slot *p) {
#ifdef __CUDA_ARCH__
  Slot s;
  // read-only path, 16 bytes in one request
  asm volatile("ld.global.nc.v2.u64 {%0, %1}, [%2];" : "=l"(s.key), "=l"(s.val) : "l"(p));
  return s;
#else
  return *p;
#endif
}

// Graph::contains + adjacency fetch.  Returns the slot index or -1.
G2S_D int64_t table_find(const Table &t, uint64_t canon, uint32_t *adj) {
  uint64_t h = mix64(canon) & t.cap_mask;
  for (;;) {
    Slot s = load_slot(t.slots + h);
    if (s.key == canon) { *adj = (uint32_t)(s.val & 0xFF); return (int64_t)h; }
    if (s.key == KEY_EMPTY) return -1;
    h = (h + 1) & t.cap_mask;
  }
}

// same, returning the whole value word (adjacency byte + skip hints)
G2S_D int64_t table_find_val(const Table &t, uint64_t canon, uint64_t *val) {
  uint64_t h = mix64(canon) & t.cap_mask;
  for (;;) {
    Slot s = load_slot(t.slots + h);
    if (s.key == canon) { *val = s.val; return (int64_t)h; }
    if (s.key == KEY_EMPTY) return -1;
    h = (h + 1) & t.cap_mask;
  }
}

// Skip hints of one node, computed from the adjacency bytes by walking (used by the table build kernel and by the
// host emulation in tests/): returns the 56 hint bits to be stored at bit 8 of Slot::val.
G2S_D uint64_t compute_hints(const Table &t, uint64_t key, uint32_t adj) {
  uint64_t out = 0;
  for (uint32_t o = 0; o < 2; o++) {
    uint64_t x = o ? revcomp(key, t.k) : key;
    uint32_t mask = out_mask(adj, o), bits = 0, len = 0;
    while (len < (uint32_t)HINT_MAX && mask != 0 && !(mask & (mask - 1))) {
      const uint32_t nt = mask == 1 ? 0 : mask == 2 ? 1 : mask == 4 ? 2 : 3;
      bits = (bits << 2) | nt;
      len++;
      x = succ_kmer(x, nt, t);
      const Canon c = canonical(x, t.k);
      uint32_t a2 = 0;
      if (table_find(t, c.key, &a2) < 0) break;   // cannot happen: the adjacency byte said the neighbour exists
      mask = out_mask(a2, c.strand);
    }
    out |= (uint64_t)((bits & 0xFFFFFFu) | (len << 24)) << (28 * o);
  }
  return out;
}

// ASCII -> oriented k-mer value (no validity check, exactly like Graph::buildNode, Graph.cpp:1726-1744)
G2S_HD uint64_t kmer_from_ascii(const char *s, int k) {
  uint64_t v = 0;
  for (int i = 0; i < k; i++) v = (v << 2) | (uint64_t)(((unsigned char)s[i] >> 1) & 3);
  return v;
}

}  // namespace g2s
