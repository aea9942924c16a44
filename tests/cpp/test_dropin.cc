// Compile-and-link check of the C++ drop-in (gam_ngs_b200/cpp/gamx_dropin.hpp) against libgamx.so,
// written like a reference call site (PctgBuilder.cc:1628,1669).  On a box with a GPU it also runs.
#include <cstdio>
#include <vector>

#include "../../gam_ngs_b200/cpp/gamx_dropin.hpp"

using namespace gamx_dropin;

int main() {
  std::vector<uint8_t> a(300), b;
  unsigned s = 12345;
  for (auto& c : a) { s = s * 1103515245u + 12345u; c = (s >> 16) & 3; }
  b = a;
  b[100] = (b[100] + 1) & 3;
  b.erase(b.begin() + 200);
  try {
    BandedSmithWaterman aligner;  // band 150, gap -8
    MyAlignment al = aligner.find_alignment(a, 0, a.size() - 1, b, 0, b.size() - 1);
    std::pair<MyAlignment::size_type, MyAlignment::size_type> p;
    bool ok = last_match_pos(al, p);
    std::printf("score=%ld begin=(%lu,%lu) len=%lu homology=%.3f last_match=(%lu,%lu,%d)\n", (long)al.score(),
                (unsigned long)al.begin_a(), (unsigned long)al.begin_b(), (unsigned long)al.length(), al.homology(),
                (unsigned long)p.first, (unsigned long)p.second, (int)ok);
    // expected from the oracle for this input: score=1478 begin=(0,0) len=300 homology=99.333
    return (al.score() == 1478 && al.length() == 300) ? 0 : 1;
  } catch (const std::runtime_error& e) {
    std::printf("no GPU: %s\n", e.what());
    return 77;
  }
}
