// TEST INFRASTRUCTURE ONLY (fixture generator, never linked into the product).
// Writes a GATB graph file whose datasets carry the shuffle + deflate filters: the same Graph::create call as the
// reference (src/Gap2Seq.cpp:213-214) plus GATB's own "-out-compress" option (debruijn/impl/Graph.cpp:295,
// tools/storage/impl/CollectionHDF5Patch.hpp:184-187).  The reference never passes that option, but graph files made by
// other GATB tools may - the reader of gap2seq_b200/csrc/g2s_h5.h must understand them.
//   usage: h5_compressed <reads.fa> <k> <solid> <level> <out prefix>      ->  <out prefix>.h5
#include <gatb/gatb_core.hpp>

#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

int main(int argc, char **argv) {
  if (argc < 6) { std::cerr << "usage: h5_compressed reads.fa k solid level out_prefix" << std::endl; return 2; }
  try {
    std::vector<std::string> files(1, argv[1]);
    BankAlbum *b = new BankAlbum(files);
    Graph graph = Graph::create(b, "-kmer-size %d -abundance-min %d -debloom original -debloom-impl basic -out-compress %d -out %s -nb-cores 1",
                                atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), argv[5]);
    std::cout << graph.getInfo() << std::endl;
  } catch (Exception &e) {
    std::cerr << "EXCEPTION: " << e.getMessage() << std::endl;
    return 1;
  }
  return 0;
}
