// g2s_h5.h - reads the solid k-mers of a GATB graph file (<reads>.h5) without GATB or libhdf5.
//
// What the reference does with such a file: Gap2Seq::execute loads it instead of building the graph when
// <reads>.h5 is readable (src/Gap2Seq.cpp:171, 195-197); GATB then opens group "dsk/solid" and iterates its
// partitions (gatb-core/src/gatb/debruijn/impl/Graph.cpp:171-184, 2243-2362) - one dataset per partition, named
// "0" .. "N-1", of Kmer<span>::Count records { value (the canonical k-mer), abundance }.  Only k-mers at or above the
// solidity threshold are stored, so the union of the partitions IS the node set of the graph.
//
// The files GATB writes (HDF5 1.8, library defaults) use the original on-disk structures only: superblock version 0,
// version-1 object headers, symbol-table groups (version-1 B-tree of type 0 + local heap + SNOD nodes), and chunked
// datasets indexed by a version-1 B-tree of type 1.  That subset of the HDF5 file format specification is what this
// reader implements; anything else (superblock >= 2, "OHDR" headers, fractal heaps, filters other than shuffle /
// deflate) is reported as an error, never guessed at.  Chunks may carry the shuffle and deflate filters (GATB's
// optional compression); zlib does the inflation.
#pragma once
#include <stdint.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace g2s {
namespace h5 {

struct Error : std::runtime_error {
  explicit Error(const std::string &m) : std::runtime_error("g2s h5: " + m) {}
};

struct Message {
  uint16_t type;
  uint64_t off;    // of the message body in the file
  uint16_t size;
};

struct DatasetInfo {
  std::vector<uint64_t> dims;
  uint32_t elem_size = 0;
  int layout = -1;                 // 1 contiguous, 2 chunked
  uint64_t addr = 0, bytes = 0;    // contiguous: data address and size; chunked: address of the chunk B-tree
  std::vector<uint32_t> chunk;     // chunk extent per dimension (elements), the element size last
  std::vector<uint16_t> filters;   // filter ids in pipeline order (1 deflate, 2 shuffle)
  std::vector<std::pair<std::string, uint32_t>> members;   // compound type: member name, byte offset
};

class File {
 public:
  File(const uint8_t *data, size_t size) : d_(data), n_(size) {
    static const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    if (n_ < 96 || memcmp(d_, sig, 8) != 0) throw Error("not an HDF5 file");
    if (d_[8] != 0) throw Error("superblock version " + std::to_string(d_[8]) + " is not supported (GATB writes version 0)");
    if (d_[13] != 8 || d_[14] != 8) throw Error("offsets and lengths must be 8 bytes wide");
    base_ = u(24, 8);
    // root group symbol table entry: link name offset, object header address, cache type, reserved, scratch pad
    root_ = u(56 + 8, 8);
  }

  uint64_t root() const { return root_; }

  // object header (version 1) -> its messages, continuation blocks followed
  std::vector<Message> messages(uint64_t oh) const {
    need(oh, 16);
    if (memcmp(d_ + base_ + oh, "OHDR", 4) == 0) throw Error("version-2 object headers are not supported");
    if (b(oh) != 1) throw Error("object header version " + std::to_string(b(oh)));
    const uint32_t total = (uint32_t)u(oh + 2, 2);
    std::vector<std::pair<uint64_t, uint64_t>> blocks;
    blocks.push_back(std::make_pair(oh + 16, u(oh + 8, 4)));
    std::vector<Message> out;
    for (size_t bi = 0; bi < blocks.size() && out.size() < total; bi++) {
      uint64_t p = blocks[bi].first;
      const uint64_t end = p + blocks[bi].second;
      while (p + 8 <= end && out.size() < total) {
        Message m;
        m.type = (uint16_t)u(p, 2); m.size = (uint16_t)u(p + 2, 2); m.off = p + 8;
        need(m.off, m.size);
        out.push_back(m);
        if (m.type == 0x10) blocks.push_back(std::make_pair(u(m.off, 8), u(m.off + 8, 8)));   // continuation
        p += 8 + (uint64_t)m.size;
      }
    }
    return out;
  }

  // members of a symbol-table group: (name, object header address)
  std::vector<std::pair<std::string, uint64_t>> children(uint64_t group_oh) const {
    std::vector<std::pair<std::string, uint64_t>> out;
    for (const Message &m : messages(group_oh)) {
      if (m.type != 0x11) continue;
      const uint64_t btree = u(m.off, 8), heap = u(m.off + 8, 8);
      need(heap, 32);
      if (memcmp(d_ + base_ + heap, "HEAP", 4) != 0) throw Error("local heap signature");
      const uint64_t heap_data = u(heap + 24, 8);
      group_node(btree, heap_data, out, 0);
      return out;
    }
    throw Error("object is not a group");
  }

  bool child(uint64_t group_oh, const std::string &name, uint64_t *oh) const {
    for (const auto &c : children(group_oh))
      if (c.first == name) { *oh = c.second; return true; }
    return false;
  }

  // path like "dsk/solid"
  uint64_t open(const std::string &path) const {
    uint64_t cur = root_;
    size_t i = 0;
    while (i < path.size()) {
      size_t j = path.find('/', i);
      if (j == std::string::npos) j = path.size();
      if (j > i && !child(cur, path.substr(i, j - i), &cur)) throw Error("no object '" + path + "'");
      i = j + 1;
    }
    return cur;
  }

  // string attribute of an object (GATB keeps its properties that way, e.g. "kmer_size"); false when absent
  bool attribute(uint64_t oh, const std::string &name, std::string *value) const {
    for (const Message &m : messages(oh)) {
      if (m.type != 0x0C || m.size < 8 || b(m.off) != 1) continue;
      const uint32_t nsz = (uint32_t)u(m.off + 2, 2), tsz = (uint32_t)u(m.off + 4, 2), ssz = (uint32_t)u(m.off + 6, 2);
      const uint64_t pn = m.off + 8, pt = pn + pad8(nsz), ps = pt + pad8(tsz), pd = ps + pad8(ssz);
      if (pd > m.off + m.size) continue;
      std::string nm((const char *)d_ + base_ + pn, strnlen((const char *)d_ + base_ + pn, nsz));
      if (nm != name) continue;
      const uint32_t esz = (uint32_t)u(pt + 4, 4);
      if ((b(pt) & 15) == 9) {   // variable-length string: length, global heap collection, object index
        if (pd + 16 > m.off + m.size) continue;
        const uint64_t len = u(pd, 4), coll = u(pd + 4, 8), want = u(pd + 12, 4);
        need(coll, 16);
        if (memcmp(d_ + base_ + coll, "GCOL", 4) != 0) throw Error("global heap collection");
        const uint64_t cend = coll + u(coll + 8, 8);
        for (uint64_t p = coll + 16; p + 16 <= cend;) {
          const uint64_t idx = u(p, 2), osz = u(p + 8, 8);
          if (idx == 0) break;   // free space
          if (idx == want) {
            need(p + 16, osz);
            value->assign((const char *)d_ + base_ + p + 16, (size_t)(len < osz ? len : osz));
            while (!value->empty() && value->back() == 0) value->pop_back();
            return true;
          }
          p += 16 + pad8(osz);
        }
        continue;
      }
      const uint64_t avail = m.off + m.size - pd;
      const uint64_t len = esz < avail ? esz : avail;
      value->assign((const char *)d_ + base_ + pd, strnlen((const char *)d_ + base_ + pd, (size_t)len));
      return true;
    }
    return false;
  }

  DatasetInfo dataset(uint64_t oh) const {
    DatasetInfo di;
    for (const Message &m : messages(oh)) {
      if (m.type == 0x01) {          // dataspace
        const int ver = b(m.off), rank = b(m.off + 1);
        const uint64_t p = m.off + (ver == 1 ? 8 : 4);
        for (int i = 0; i < rank; i++) di.dims.push_back(u(p + 8 * (uint64_t)i, 8));
      } else if (m.type == 0x03) {   // datatype
        di.elem_size = (uint32_t)u(m.off + 4, 4);
        const int cls = b(m.off) & 15, ver = b(m.off) >> 4;
        if (cls == 6 && ver == 1) {  // compound, version 1 members: name (padded to 8), offset, 28 bytes of array info, member type
          const uint32_t nmem = (uint32_t)u(m.off + 1, 2);
          uint64_t p = m.off + 8;
          for (uint32_t i = 0; i < nmem && p < m.off + m.size; i++) {
            const size_t len = strnlen((const char *)d_ + base_ + p, (size_t)(m.off + m.size - p));
            std::string nm((const char *)d_ + base_ + p, len);
            p += pad8((uint32_t)len + 1);
            const uint32_t off = (uint32_t)u(p, 4);
            p += 4 + 1 + 3 + 4 + 4 + 16;
            const int mcls = b(p) & 15;
            uint32_t prop = 0;   // size of the member type's properties
            if (mcls == 0) prop = 4; else if (mcls == 1) prop = 12; else throw Error("compound member of class " + std::to_string(mcls));
            p += 8 + prop;
            di.members.push_back(std::make_pair(nm, off));
          }
        }
      } else if (m.type == 0x08) {   // data layout
        if (b(m.off) != 3) throw Error("data layout message version " + std::to_string(b(m.off)));
        di.layout = b(m.off + 1);
        if (di.layout == 1) { di.addr = u(m.off + 2, 8); di.bytes = u(m.off + 10, 8); }
        else if (di.layout == 2) {
          const int nd = b(m.off + 2);
          di.addr = u(m.off + 3, 8);
          for (int i = 0; i < nd; i++) di.chunk.push_back((uint32_t)u(m.off + 11 + 4 * (uint64_t)i, 4));
        } else throw Error("compact datasets are not supported");
      } else if (m.type == 0x0B) {   // filter pipeline
        const int ver = b(m.off), nf = b(m.off + 1);
        uint64_t p = m.off + (ver == 1 ? 8 : 2);
        for (int i = 0; i < nf; i++) {
          const uint16_t id = (uint16_t)u(p, 2);
          uint32_t nlen = (ver == 1 || id >= 256) ? (uint32_t)u(p + 2, 2) : 0;
          const uint64_t q = p + ((ver == 1 || id >= 256) ? 4 : 2);
          const uint32_t ncd = (uint32_t)u(q + 2, 2);
          p = q + 4 + (ver == 1 ? pad8(nlen) : nlen) + 4 * (uint64_t)ncd;
          if (ver == 1 && (ncd & 1)) p += 4;
          di.filters.push_back(id);
        }
      }
    }
    if (di.layout < 0 || di.elem_size == 0) throw Error("object is not a dataset");
    return di;
  }

  // whole one-dimensional dataset, raw records
  void read(uint64_t oh, std::vector<uint8_t> &out, DatasetInfo *info = nullptr) const {
    const DatasetInfo di = dataset(oh);
    if (info) *info = di;
    if (di.dims.size() != 1) throw Error("only one-dimensional datasets are read");
    const uint64_t total = di.dims[0] * (uint64_t)di.elem_size;
    out.assign((size_t)total, 0);
    if (total == 0) return;
    if (di.layout == 1) {
      if (di.addr == ~0ULL) return;   // never written
      need(di.addr, total);
      memcpy(out.data(), d_ + base_ + di.addr, (size_t)total);
      return;
    }
    if (di.chunk.size() != 2 || di.chunk[1] != di.elem_size) throw Error("unexpected chunk geometry");
    for (uint16_t f : di.filters) if (f != 1 && f != 2) throw Error("filter " + std::to_string(f) + " is not supported");
    if (di.addr != ~0ULL) chunk_node(di.addr, di, out, 0);
  }

 private:
  const uint8_t *d_;
  size_t n_;
  uint64_t base_ = 0, root_ = 0;

  static uint64_t pad8(uint64_t x) { return (x + 7) & ~7ULL; }
  void need(uint64_t off, uint64_t len) const {
    if (off > n_ || len > n_ || base_ + off + len > n_) throw Error("address beyond the end of the file");
  }
  int b(uint64_t off) const { need(off, 1); return d_[base_ + off]; }
  uint64_t u(uint64_t off, int nbytes) const {
    need(off, (uint64_t)nbytes);
    uint64_t v = 0;
    for (int i = nbytes - 1; i >= 0; i--) v = (v << 8) | d_[base_ + off + (uint64_t)i];
    return v;
  }

  void group_node(uint64_t node, uint64_t heap_data, std::vector<std::pair<std::string, uint64_t>> &out, int depth) const {
    if (depth > 32) throw Error("group B-tree too deep");
    need(node, 24);
    if (memcmp(d_ + base_ + node, "TREE", 4) != 0 || b(node + 4) != 0) throw Error("group B-tree node");
    const int level = b(node + 5);
    const uint32_t n = (uint32_t)u(node + 6, 2);
    uint64_t p = node + 24;                      // key 0 (8 bytes), child 0, key 1, ...
    for (uint32_t i = 0; i < n; i++, p += 16) {
      const uint64_t ch = u(p + 8, 8);
      if (level > 0) { group_node(ch, heap_data, out, depth + 1); continue; }
      need(ch, 8);
      if (memcmp(d_ + base_ + ch, "SNOD", 4) != 0) throw Error("symbol table node");
      const uint32_t ns = (uint32_t)u(ch + 6, 2);
      for (uint32_t s = 0; s < ns; s++) {
        const uint64_t e = ch + 8 + 40 * (uint64_t)s;
        const uint64_t name_off = heap_data + u(e, 8);
        need(name_off, 1);
        const size_t len = strnlen((const char *)d_ + base_ + name_off, (size_t)(n_ - base_ - name_off));
        out.push_back(std::make_pair(std::string((const char *)d_ + base_ + name_off, len), u(e + 8, 8)));
      }
    }
  }

  void chunk_node(uint64_t node, const DatasetInfo &di, std::vector<uint8_t> &out, int depth) const {
    if (depth > 32) throw Error("chunk B-tree too deep");
    need(node, 24);
    if (memcmp(d_ + base_ + node, "TREE", 4) != 0 || b(node + 4) != 1) throw Error("chunk B-tree node");
    const int level = b(node + 5);
    const uint32_t n = (uint32_t)u(node + 6, 2);
    const uint64_t key = 8 + 8 * (uint64_t)di.chunk.size();   // chunk size, filter mask, offset per dimension (+ the element one)
    uint64_t p = node + 24;
    for (uint32_t i = 0; i < n; i++, p += key + 8) {
      const uint64_t ch = u(p + key, 8);
      if (level > 0) { chunk_node(ch, di, out, depth + 1); continue; }
      const uint32_t stored = (uint32_t)u(p, 4), mask = (uint32_t)u(p + 4, 4);
      const uint64_t first = u(p + 8, 8);        // element offset of the chunk in the dataset
      if (first >= di.dims[0]) continue;
      const uint64_t want_el = std::min<uint64_t>(di.chunk[0], di.dims[0] - first);
      const uint64_t full = (uint64_t)di.chunk[0] * di.elem_size;
      need(ch, stored);
      std::vector<uint8_t> buf(d_ + base_ + ch, d_ + base_ + ch + stored);
      for (size_t fi = di.filters.size(); fi-- > 0;) {   // undo the pipeline back to front
        if (mask & (1u << fi)) continue;
        if (di.filters[fi] == 1) {
          std::vector<uint8_t> raw((size_t)full);
          uLongf len = (uLongf)full;
          if (uncompress(raw.data(), &len, buf.data(), (uLong)buf.size()) != Z_OK) throw Error("deflate stream of a chunk");
          raw.resize(len);
          buf.swap(raw);
        } else {                                          // shuffle: byte b of every element stored together
          const size_t es = di.elem_size, ne = buf.size() / es;
          std::vector<uint8_t> raw(buf.size());
          for (size_t bb = 0; bb < es; bb++)
            for (size_t e = 0; e < ne; e++) raw[e * es + bb] = buf[bb * ne + e];
          memcpy(raw.data() + ne * es, buf.data() + ne * es, buf.size() - ne * es);
          buf.swap(raw);
        }
      }
      if (buf.size() < want_el * di.elem_size) throw Error("short chunk");
      memcpy(out.data() + first * di.elem_size, buf.data(), (size_t)(want_el * di.elem_size));
    }
  }
};

// The node set of a GATB graph file: canonical k-mers (64-bit values, k <= 32) of every partition of dsk/solid, in
// file order, with their abundances when asked for.  *kmer_size = the "kmer_size" property, 0 when the file has none.
inline void read_solid_kmers(const uint8_t *data, size_t size, std::vector<uint64_t> &kmers, std::vector<uint32_t> *abundance,
                             int *kmer_size) {
  File f(data, size);
  const uint64_t solid = f.open("dsk/solid");
  if (kmer_size) {
    *kmer_size = 0;
    std::string v;
    for (const char *where : {"dsk", "dsk/solid", "bloom/bloom", ""}) {
      uint64_t oh;
      try { oh = f.open(where); } catch (const Error &) { continue; }
      if (f.attribute(oh, "kmer_size", &v)) { *kmer_size = atoi(v.c_str()); break; }
    }
  }
  std::vector<std::pair<std::string, uint64_t>> parts = f.children(solid);
  // partitions are named by their number; keep numeric order so that the output is reproducible
  std::vector<std::pair<long, uint64_t>> ordered;
  for (const auto &c : parts) {
    char *end = nullptr;
    const long id = strtol(c.first.c_str(), &end, 10);
    if (end == c.first.c_str() || *end) continue;
    ordered.push_back(std::make_pair(id, c.second));
  }
  std::sort(ordered.begin(), ordered.end());
  kmers.clear();
  if (abundance) abundance->clear();
  std::vector<uint8_t> raw;
  for (const auto &pt : ordered) {
    DatasetInfo di;
    f.read(pt.second, raw, &di);
    uint32_t off_value = 0, off_ab = 8;
    for (const auto &m : di.members) { if (m.first == "value") off_value = m.second; if (m.first == "abundance") off_ab = m.second; }
    if (di.elem_size < 12 || off_value + 8 > di.elem_size || off_ab + 4 > di.elem_size)
      throw Error("k-mer records of " + std::to_string(di.elem_size) + " bytes: only 64-bit k-mers (k <= 32) are read");
    if (di.elem_size != 16) throw Error("k-mer records of " + std::to_string(di.elem_size) + " bytes: k > 32 is not supported");
    const size_t n = raw.size() / di.elem_size;
    for (size_t i = 0; i < n; i++) {
      uint64_t v; uint32_t a;
      memcpy(&v, raw.data() + i * di.elem_size + off_value, 8);
      memcpy(&a, raw.data() + i * di.elem_size + off_ab, 4);
      kmers.push_back(v);
      if (abundance) abundance->push_back(a);
    }
  }
}

}  // namespace h5
}  // namespace g2s
