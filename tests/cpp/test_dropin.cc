// Compile-and-link check of the C++ drop-in (gam_ngs_b200/cpp/gamx_dropin.hpp) against libgamx.so,
// written like a reference call site (PctgBuilder.cc:1628,1669).  On a box with a GPU it also runs.
#include <cstdio>
#include <thread>
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
    if (!(al.score() == 1478 && al.length() == 300)) return 1;

    // The legacy call pattern (ThreadedBuildPctg.cc:159-169 -> PctgBuilder.cc:1628,1669): N host threads, each
    // with its own stack-local aligner, issuing synchronous calls at the same time.  Every thread must get
    // exactly what a single thread gets for the same pairs.
    const int kThreads = 8, kCalls = 12;
    std::vector<std::vector<uint8_t>> A(kThreads * kCalls), B(kThreads * kCalls);
    for (size_t k = 0; k < A.size(); k++) {
      A[k].resize(200 + 37 * (k % 11));
      for (auto& c : A[k]) { s = s * 1103515245u + 12345u; c = (s >> 16) & 3; }
      B[k] = A[k];
      B[k][50 + k % 40] = (B[k][50 + k % 40] + 1) & 3;
      if (k % 3 == 0) B[k].erase(B[k].begin() + 120);
      if (k % 5 == 0) B[k].insert(B[k].begin() + 30, (uint8_t)(k & 3));
    }
    struct Out { long score; unsigned long ba, bb, len; std::vector<AlignmentAlphabet> ops; };
    auto one = [&](size_t k) {
      BandedSmithWaterman al2(k % 2 ? 64 : 150);
      MyAlignment r = al2.find_alignment(A[k], 0, A[k].size() - 1, B[k], 0, B[k].size() - 1);
      Out o{(long)r.score(), (unsigned long)r.begin_a(), (unsigned long)r.begin_b(), (unsigned long)r.length(), {}};
      o.ops = r.sequence();
      return o;
    };
    std::vector<Out> serial(A.size()), par(A.size());
    for (size_t k = 0; k < A.size(); k++) serial[k] = one(k);
    std::vector<std::thread> th;
    for (int t = 0; t < kThreads; t++)
      th.emplace_back([&, t] { for (int c = 0; c < kCalls; c++) par[(size_t)t * kCalls + c] = one((size_t)t * kCalls + c); });
    for (auto& x : th) x.join();
    int bad = 0;
    for (size_t k = 0; k < A.size(); k++)
      bad += !(serial[k].score == par[k].score && serial[k].ba == par[k].ba && serial[k].bb == par[k].bb &&
               serial[k].len == par[k].len && serial[k].ops == par[k].ops);
    std::printf("concurrent callers: %d threads x %d calls, %d mismatches\n", kThreads, kCalls, bad);
    if (bad) return 2;

    // ... and concurrent batches on ONE shared context through the C ABI (the context serialises them)
    gamx_ctx* shared = default_context();
    std::vector<uint32_t> ida(A.size()), idb(A.size());
    for (size_t k = 0; k < A.size(); k++) {
      ida[k] = (uint32_t)gamx_add_contig(shared, A[k].data(), A[k].size());
      idb[k] = (uint32_t)gamx_add_contig(shared, B[k].data(), B[k].size());
    }
    std::vector<gamx_result> res(A.size());
    std::vector<int> rcs(kThreads, 0);
    th.clear();
    for (int t = 0; t < kThreads; t++)
      th.emplace_back([&, t] {
        std::vector<gamx_job> jobs(kCalls);
        for (int c = 0; c < kCalls; c++) {
          const size_t k = (size_t)t * kCalls + c;
          gamx_job j = {};
          j.a_id = ida[k]; j.b_id = idb[k]; j.a_len = j.b_len = UINT64_MAX;
          j.end_a = A[k].size() - 1; j.end_b = B[k].size() - 1;
          j.band = k % 2 ? 64 : 150; j.gap = GAMX_DEFAULT_GAP; j.mode = GAMX_MODE_ENDPOINTS;
          jobs[c] = j;
        }
        rcs[t] = gamx_align_batch(shared, jobs.data(), jobs.size(), res.data() + (size_t)t * kCalls, nullptr, 0);
        (void)gamx_contig_length(shared, ida[0]); (void)gamx_last_error(shared);
      });
    for (auto& x : th) x.join();
    bad = 0;
    for (int t = 0; t < kThreads; t++) bad += rcs[t] != GAMX_OK;
    for (size_t k = 0; k < A.size(); k++)
      bad += !(res[k].status == GAMX_JOB_OK && res[k].score == serial[k].score && res[k].begin_a == serial[k].ba &&
               res[k].begin_b == serial[k].bb && res[k].n_ops == serial[k].len);
    std::printf("concurrent batches on one context: %d mismatches\n", bad);
    return bad ? 3 : 0;
  } catch (const std::runtime_error& e) {
    std::printf("no GPU: %s\n", e.what());
    return 77;
  }
}
