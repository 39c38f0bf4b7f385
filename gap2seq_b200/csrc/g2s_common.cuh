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

// 32-byte slot of the global open-addressing table = one sector: a probe that hits in its first position costs a
// single sector gather and brings the node's neighbourhood AND its skip hints.
struct __attribute__((aligned(32))) Slot {
  uint64_t key;   // canonical k-mer, KEY_EMPTY when free
  uint64_t val;   // bits 0-7: adjacency byte (low nibble out A,C,T,G, high nibble in; canonical orientation)
                  // bits 8-13 / 14-19: hint length of the canonical / reverse-complement orientation
                  // bits 20-35 / 36-51: bits 64-79 of that orientation's hint bases
  uint64_t h[2];  // bits 0-63 of the hint bases, per orientation
};

// Skip hints.  For an oriented k-mer x the hint lists the next j <= HINT_MAX bases of its UNIQUE forward extension:
// x and the j-1 nodes after it each have exactly one successor (the listed base); nothing is claimed about the node
// the last base leads to.  A walker that holds a hint can therefore emit j levels from ONE table gather - the
// per-level dependent gather of a chain of unitig-interior nodes disappears - while every node still becomes its
// own DP cell.  The bases form a 2j-bit big-endian string (first base highest, like a k-mer; up to 80 bits), so that
// the state i steps ahead is ((x << 2i) | (bases >> 2(j-i))) & kmask.
static const int HINT_MAX = 40;
struct Hint {
  uint64_t lo;    // bits 0-63 of the bases
  uint32_t hi;    // bits 64-79
  uint32_t len;   // j
};
G2S_HD Hint hint_of(const Slot &s, uint32_t strand) {
  Hint h;
  h.lo = strand ? s.h[1] : s.h[0];   // a select, not an indexed load: the slot stays in registers
  h.hi = (uint32_t)(s.val >> (20 + 16 * strand)) & 0xFFFFu;
  h.len = (uint32_t)(s.val >> (8 + 6 * strand)) & 63u;
  return h;
}
G2S_HD uint64_t hint_shr(const Hint &h, uint32_t sh) {   // low 64 bits of (bases >> sh), sh < 160
  if (sh == 0) return h.lo;
  if (sh < 64) return (h.lo >> sh) | ((uint64_t)h.hi << (64 - sh));
  return sh < 96 ? (uint64_t)h.hi >> (sh - 64) : 0;
}
// the state j (1..len) steps past tip x, and the base of step j (0-based: the base that leads to state j+1)
G2S_HD uint64_t hint_state(uint64_t x, const Hint &h, uint32_t j, uint64_t kmask) {
  return ((2 * j < 64 ? x << (2 * j) : 0) | hint_shr(h, 2 * (h.len - j))) & kmask;
}
G2S_HD uint32_t hint_base(const Hint &h, uint32_t j) { return (uint32_t)hint_shr(h, 2 * (h.len - j - 1)) & 3u; }
G2S_HD void hint_push(Hint &h, uint32_t nt) {   // append one base
  h.hi = ((h.hi << 2) | (uint32_t)(h.lo >> 62)) & 0xFFFFu;
  h.lo = (h.lo << 2) | nt;
  h.len++;
}

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

// the whole 32-byte slot = one sector in ONE request (sm_100 256-bit load, LDG.E.256)
G2S_D Slot load_slot32(const Slot *p) {
#ifdef __CUDA_ARCH__
  Slot s;
  asm volatile("ld.global.nc.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(s.key), "=l"(s.val), "=l"(s.h[0]), "=l"(s.h[1]) : "l"(p));
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

// same, returning the whole slot (adjacency byte + skip hints): both halves of the sector
G2S_D int64_t table_find_slot(const Table &t, uint64_t canon, Slot *out) {
  uint64_t h = mix64(canon) & t.cap_mask;
  for (;;) {
#ifdef __CUDA_ARCH__
    Slot s;   // the whole 32-byte slot = one sector in ONE request (sm_100 256-bit load, LDG.E.256)
    asm volatile("ld.global.nc.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(s.key), "=l"(s.val), "=l"(s.h[0]), "=l"(s.h[1]) : "l"(t.slots + h));
#else
    const Slot s = t.slots[h];
#endif
    if (s.key == canon) { *out = s; return (int64_t)h; }
    if (s.key == KEY_EMPTY) return -1;
    h = (h + 1) & t.cap_mask;
  }
}

// Skip hints of one node, computed from the adjacency bytes by walking (used by the table build kernel and by the
// host emulation in tests/): fills the hint fields of `out` (val bits 8-51 and h[]); the adjacency byte is kept.
G2S_D void compute_hints(const Table &t, uint64_t key, uint32_t adj, Slot *out) {
  uint64_t val = adj & 0xFFu;
  for (uint32_t o = 0; o < 2; o++) {
    uint64_t x = o ? revcomp(key, t.k) : key;
    uint32_t mask = out_mask(adj, o);
    Hint h{0, 0, 0};
    while (h.len < (uint32_t)HINT_MAX && mask != 0 && !(mask & (mask - 1))) {
      const uint32_t nt = mask == 1 ? 0 : mask == 2 ? 1 : mask == 4 ? 2 : 3;
      hint_push(h, nt);
      x = succ_kmer(x, nt, t);
      const Canon c = canonical(x, t.k);
      uint32_t a2 = 0;
      if (table_find(t, c.key, &a2) < 0) break;   // cannot happen: the adjacency byte said the neighbour exists
      mask = out_mask(a2, c.strand);
    }
    out->h[o] = h.lo;
    val |= ((uint64_t)h.len << (8 + 6 * o)) | ((uint64_t)h.hi << (20 + 16 * o));
  }
  out->val = val;
}

// ASCII -> oriented k-mer value (no validity check, exactly like Graph::buildNode, Graph.cpp:1726-1744)
G2S_HD uint64_t kmer_from_ascii(const char *s, int k) {
  uint64_t v = 0;
  for (int i = 0; i < k; i++) v = (v << 2) | (uint64_t)(((unsigned char)s[i] >> 1) & 3);
  return v;
}

}  // namespace g2s
