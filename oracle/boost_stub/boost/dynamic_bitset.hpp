// Build stub, test infrastructure only.
// The reference's lib/types.hpp:35 includes <boost/dynamic_bitset.hpp> for one
// typedef (lib/types.hpp:37) that the aligner translation units never use.
// Boost is not installed in this image, so the oracle/_ref build puts this
// empty template on the include path instead.
#pragma once
namespace boost {
template <class Block = unsigned long, class Alloc = void>
class dynamic_bitset {};
}  // namespace boost
