// TEST INFRASTRUCTURE ONLY - never linked into the product library.
//
// Pins the caller side of the hot path: the UNMODIFIED bodies of
//     PctgBuilder::alignMergeBlock / findBestAlignment / alignBlocks / is_good
//     (lib/src/pctg/PctgBuilder.cc:726-844, 1361-1614, 1617-1708, 1711-1730)
// are copied out of the reference source at build time by oracle/pctg_extract.py into
// oracle/_ref/pctg_extract.inc (generated, git-ignored) and compiled here against the reference's real
// Contig / MyAlignment / BandedSmithWaterman / ABlast / BestCtgAlignment / MergeBlock, plus stand-ins for what
// the whole file would drag in (Boost.Graph, sparsehash, BamTools):
//     Frame   - the five getters the bodies call, same types as lib/include/assembly/Frame.hpp:124-152,
//               getLength as Frame.cc:124-127
//     Block   - getMasterFrame / getSlaveFrame / getReadsNumber / getMasterId / getSlaveId, Block.hpp:115-149
//     CompactAssemblyGraph - Vertex + getBlocks(v), CompactAssemblyGraph.hpp:69,138
//     PctgBuilder - the declarations of the five member functions (PctgBuilder.hpp:162-199) and the two contig
//               loaders (:116-133), which here return the contigs handed to the C entry point
// The C entry point takes a merge block the way include/gamx.h describes it (gamx_merge_block / gamx_block).
#include <stdint.h>
#include <string.h>

#include <list>
#include <stdexcept>
#include <vector>

#include "alignment/ablast.hpp"
#include "alignment/banded_smith_waterman.hpp"
#include "alignment/my_alignment.hpp"
#include "assembly/contig.hpp"
#include "pctg/BestCtgAlignment.hpp"
#include "pctg/MergeDescriptor.hpp"

#define GAMX_PCTG_PART 1
#include "pctg_extract.inc"  // thresholds (MIN_HOMOLOGY ...)
#undef GAMX_PCTG_PART

typedef int64_t IntTypeStub;  // (Block::getReadsNumber returns IntType = a signed 64-bit integer, types.hpp)

class Frame {
 public:
  Frame() : _ctgId(0), _strand('?'), _begin(0), _end(0) {}
  Frame(int32_t ctg, char strand, int32_t begin, int32_t end) : _ctgId(ctg), _strand(strand), _begin(begin), _end(end) {}
  int32_t getContigId() const { return _ctgId; }
  char getStrand() const { return _strand; }
  int32_t getBegin() const { return _begin; }
  int32_t getEnd() const { return _end; }
  int32_t getLength() const { return (_end < _begin) ? 0 : _end - _begin + 1; }

 private:
  int32_t _ctgId;
  char _strand;
  int32_t _begin, _end;
};

class Block {
 public:
  Block(IntTypeStub nr, const Frame& m, const Frame& s) : _numReads(nr), _masterFrame(m), _slaveFrame(s) {}
  IntTypeStub getReadsNumber() const { return _numReads; }
  const Frame& getMasterFrame() const { return _masterFrame; }
  const Frame& getSlaveFrame() const { return _slaveFrame; }
  int32_t getMasterId() const { return _masterFrame.getContigId(); }
  int32_t getSlaveId() const { return _slaveFrame.getContigId(); }

 private:
  IntTypeStub _numReads;
  Frame _masterFrame, _slaveFrame;
};

class CompactAssemblyGraph {
 public:
  typedef uint64_t Vertex;
  std::vector<std::list<Block> > blocks;
  const std::list<Block>& getBlocks(const Vertex& pos) const { return blocks[pos]; }
};

class PctgBuilder {
 public:
  const Contig* master;
  const Contig* slave;
  const Contig& loadMasterContig(const int32_t) const { return *master; }
  const Contig& loadSlaveContig(const int32_t) const { return *slave; }
  void alignMergeBlock(const CompactAssemblyGraph& graph, MergeBlock& mb) const;
  void findBestAlignment(BestCtgAlignment& bestAlign, Contig& masterCtg, uint64_t masterStart, uint64_t masterEnd, Contig& slaveCtg,
                         uint64_t slaveStart, uint64_t slaveEnd, const std::list<Block>& blocks_list) const;
  void alignBlocks(const Contig& masterCtg, const uint64_t& masterStart, const Contig& slaveCtg, const uint64_t& slaveStart,
                   const std::list<Block>& blocks_list, std::vector<MyAlignment>& alignments) const;
  bool is_good(const std::vector<MyAlignment>& align, uint64_t min_align_len = MIN_ALIGNMENT_LEN) const;
  bool is_good(const MyAlignment& align, uint64_t min_align_len = MIN_ALIGNMENT_LEN) const;
};

#define GAMX_PCTG_PART 2
#include "pctg_extract.inc"  // the four function bodies, verbatim
#undef GAMX_PCTG_PART

extern "C" {

struct gamref_block {  // = gamx_block (include/gamx.h)
  int32_t num_reads;
  uint8_t m_strand, s_strand;
  uint8_t reserved_[2];
  int32_t m_begin, m_end, s_begin, s_end;
};

struct gamref_merge_result {  // = the fields of gamx_merge_result the reference defines
  int32_t status;      // 0 ok, 2: the reference threw
  int32_t align_ok, align_rev;
  int32_t coords_set;  // 0: alignMergeBlock returned before assigning m_start .. s_end
  int32_t m_start, m_end, s_start, s_end;
};

static Contig contig_of(const uint8_t* codes, uint64_t len) {
  Contig c("c", size_t(len));
  for (uint64_t i = 0; i < len; i++) c.at(i) = Nucleotide(BaseType(codes[i] > 4 ? 4 : codes[i]));
  return c;
}

// PctgBuilder::alignMergeBlock on one merge block: master / slave as base codes (0..4), the vertex's blocks,
// the four tail flags of the MergeBlock.
int gamref_align_merge_block(const uint8_t* m_codes, uint64_t m_len, const uint8_t* s_codes, uint64_t s_len,
                             const gamref_block* blocks, uint32_t n_blocks, int m_ltail, int m_rtail, int s_ltail, int s_rtail,
                             gamref_merge_result* out) {
  memset(out, 0, sizeof(*out));
  if (n_blocks == 0) return -1;  // (front() of an empty list: undefined in the reference)
  const Contig m = contig_of(m_codes, m_len), s = contig_of(s_codes, s_len);
  CompactAssemblyGraph graph;
  graph.blocks.resize(1);
  for (uint32_t k = 0; k < n_blocks; k++)
    graph.blocks[0].push_back(Block(blocks[k].num_reads, Frame(0, blocks[k].m_strand ? '-' : '+', blocks[k].m_begin, blocks[k].m_end),
                                    Frame(1, blocks[k].s_strand ? '-' : '+', blocks[k].s_begin, blocks[k].s_end)));
  PctgBuilder pb;
  pb.master = &m; pb.slave = &s;
  MergeBlock mb;
  memset(&mb, 0, sizeof(mb));
  const int32_t kUnset = -0x5eed;
  mb.vertex = 0; mb.m_id = 0; mb.s_id = 1;
  mb.m_start = mb.m_end = mb.s_start = mb.s_end = kUnset;
  mb.m_ltail = m_ltail != 0; mb.m_rtail = m_rtail != 0; mb.s_ltail = s_ltail != 0; mb.s_rtail = s_rtail != 0;
  try {
    pb.alignMergeBlock(graph, mb);
  } catch (const std::exception&) {
    out->status = 2;
    return 0;
  }
  out->align_ok = mb.align_ok ? 1 : 0;
  out->coords_set = !(mb.m_start == kUnset && mb.m_end == kUnset && mb.s_start == kUnset && mb.s_end == kUnset);
  if (out->coords_set) {
    out->align_rev = mb.align_rev ? 1 : 0;
    out->m_start = mb.m_start; out->m_end = mb.m_end; out->s_start = mb.s_start; out->s_end = mb.s_end;
  }
  return 0;
}

}  // extern "C"
