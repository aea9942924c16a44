// TEST INFRASTRUCTURE ONLY: CPU lane simulator for the kernel bodies.
//
// There is no GPU in the development container, so the warp-level kernel body
// (gam_ngs_b200/csrc/bsw_warp.h) is written against a small warp policy and executed here by
// 32 cooperative fibers (ucontext) in SPMD fashion: every collective (shuffle / warp sync) is
// a rendezvous of all 32 lanes, exactly like the *_sync intrinsics.  Lanes run either in
// ascending or descending order between rendezvous points, so a missing warp sync between a
// shared-memory write and a cross-lane read shows up as a mismatch in one of the two orders.
//
// The simulator shares the host-side preparation, packing and finalisation code with the
// product (bsw_host.h), so those are exercised on the CPU too.  Nothing here is used by the
// product path; the -m gpu tests exercise the real kernels through the C ABI.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <algorithm>
#include <vector>

#include "../../gam_ngs_b200/csrc/bsw_generic.h"
#include "../../gam_ngs_b200/csrc/bsw_host.h"
#include "../../gam_ngs_b200/csrc/bsw_warp.h"
#include "../../gam_ngs_b200/csrc/bsw_warp16.h"

using namespace gamx;

// (a named namespace: the build may split the template instantiations over several translation units,
// SIM_PART below)
namespace simns {

struct Sched;
struct SimWarp {
  Sched* s;
  int lane_;
  int lane() const { return lane_; }
  int exchange(int v, int src);
  int shfl_up(int v, int d, int width) { return exchange(v, (lane_ % width) - d >= 0 ? lane_ - d : lane_); }
  int shfl_down(int v, int d, int width) { return exchange(v, (lane_ % width) + d < width ? lane_ + d : lane_); }
  int shfl_xor(int v, int m, int) { return exchange(v, lane_ ^ m); }
  void sync() { exchange(0, lane_); }
};

constexpr int kMaxLanes = 256;  // K2 runs up to 256 lanes (one CTA) per pair
struct Sched {
  ucontext_t main_ctx;
  ucontext_t ctx[kMaxLanes];
  std::vector<char> stacks[kMaxLanes];
  bool done[kMaxLanes];
  int slot[kMaxLanes];
  int n = 32;  // lanes in this run
  int arrived = 0;
  unsigned gen = 0;
  int cur = 0;
  bool descending = false;
  void (*body)(SimWarp&, void*) = nullptr;
  void* arg = nullptr;
  SimWarp warps[kMaxLanes];

  void yield(int lane) { swapcontext(&ctx[lane], &main_ctx); }
  // two-phase rendezvous: everybody publishes, then everybody reads
  void barrier(int lane) {
    const unsigned g = gen;
    if (++arrived == n) { arrived = 0; gen++; }
    while (gen == g) yield(lane);
  }
  static void tramp(int lane_lo, int sp_lo, int sp_hi) {
    Sched* s = (Sched*)(((uintptr_t)(uint32_t)sp_hi << 32) | (uint32_t)sp_lo);
    s->body(s->warps[lane_lo], s->arg);
    s->done[lane_lo] = true;
    swapcontext(&s->ctx[lane_lo], &s->main_ctx);
  }
  void run(void (*b)(SimWarp&, void*), void* a, bool desc, int lanes = 32) {
    body = b; arg = a; descending = desc; arrived = 0; gen = 0; n = lanes;
    for (int l = 0; l < n; l++) {
      warps[l].s = this; warps[l].lane_ = l; done[l] = false;
      stacks[l].assign(1 << 18, 0);
      getcontext(&ctx[l]);
      ctx[l].uc_stack.ss_sp = stacks[l].data();
      ctx[l].uc_stack.ss_size = stacks[l].size();
      ctx[l].uc_link = &main_ctx;
      const uintptr_t p = (uintptr_t)this;
      makecontext(&ctx[l], (void (*)())tramp, 3, l, (int)(uint32_t)p, (int)(uint32_t)(p >> 32));
    }
    for (;;) {
      bool any = false;
      for (int q = 0; q < n; q++) {
        const int l = descending ? n - 1 - q : q;
        if (done[l]) continue;
        any = true;
        swapcontext(&main_ctx, &ctx[l]);
      }
      if (!any) break;
    }
  }
};

inline int SimWarp::exchange(int v, int src) {
  s->slot[lane_] = v;
  s->barrier(lane_);
  const int r = s->slot[src];
  s->barrier(lane_);
  return r;
}

struct WarpArgs {
  const DevJob* job[8];   // per group (null: idle)
  DevResult* out[8];
  SeqStore store;
  void* smem;
  uint32_t* dirs;
  uint64_t group_stride;
  uint32_t* ops;
  int c, lg;
  bool dirs_on;
};

template <int C, int LG>
void body_c(SimWarp& w, void* p) {
  WarpArgs* a = (WarpArgs*)p;
  WarpSmem<C, LG>& sm = *(WarpSmem<C, LG>*)a->smem;
  const int grp = LG >= 32 ? 0 : w.lane() / LG;
  if (a->dirs_on) warp_align<C, LG, true>(w, a->job[grp], a->store, sm, a->dirs, a->group_stride, a->ops, a->out[grp]);
  else warp_align<C, LG, false>(w, a->job[grp], a->store, sm, a->dirs, a->group_stride, a->ops, a->out[grp]);
}

template <int C, int LG>
void dispatch(WarpArgs& a, bool desc) {
  std::vector<uint64_t> smem((sizeof(WarpSmem<C, LG>) + 7) / 8, 0x9e3779b97f4a7c15ull);  // (garbage, like a reused block)
  a.smem = smem.data();
  Sched* s = new Sched();
  s->run(body_c<C, LG>, &a, desc, LG > 32 ? LG : 32);
  delete s;
}

template <int LG>
void run_warp_lg(WarpArgs& a, bool desc) {
  switch (a.c) {
#define CASE(N) case N: dispatch<N, LG>(a, desc); break;
    CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13)
    CASE(14) CASE(15) CASE(16) CASE(17) CASE(18)
#undef CASE
    default: break;
  }
}
// Parallel build (tests/simlib.py): part 0 holds the glue and only DECLARES the per-LG instantiations, parts
// 1..11 define one of them each.  Without SIM_PART everything is one translation unit.
#if defined(SIM_PART) && SIM_PART == 0
extern template void run_warp_lg<4>(WarpArgs&, bool);
extern template void run_warp_lg<8>(WarpArgs&, bool);
extern template void run_warp_lg<16>(WarpArgs&, bool);
extern template void run_warp_lg<32>(WarpArgs&, bool);
extern template void run_warp_lg<64>(WarpArgs&, bool);
extern template void run_warp_lg<128>(WarpArgs&, bool);
extern template void run_warp_lg<256>(WarpArgs&, bool);
#elif defined(SIM_PART) && SIM_PART == 1
template void run_warp_lg<4>(WarpArgs&, bool);
#elif defined(SIM_PART) && SIM_PART == 2
template void run_warp_lg<8>(WarpArgs&, bool);
#elif defined(SIM_PART) && SIM_PART == 3
template void run_warp_lg<16>(WarpArgs&, bool);
#elif defined(SIM_PART) && SIM_PART == 4
template void run_warp_lg<32>(WarpArgs&, bool);
#elif defined(SIM_PART) && SIM_PART == 5
template void run_warp_lg<64>(WarpArgs&, bool);
#elif defined(SIM_PART) && SIM_PART == 6
template void run_warp_lg<128>(WarpArgs&, bool);
#elif defined(SIM_PART) && SIM_PART == 7
template void run_warp_lg<256>(WarpArgs&, bool);
#endif
#if !defined(SIM_PART) || SIM_PART == 0
void run_warp(WarpArgs& a, bool desc) {
  if (a.lg == 32) run_warp_lg<32>(a, desc);
  else if (a.lg == 16) run_warp_lg<16>(a, desc);
  else if (a.lg == 8) run_warp_lg<8>(a, desc);
  else if (a.lg == 4) run_warp_lg<4>(a, desc);
  else if (a.lg == 64) run_warp_lg<64>(a, desc);
  else if (a.lg == 128) run_warp_lg<128>(a, desc);
  else run_warp_lg<256>(a, desc);
}
#endif


// ---- 16x2 pairs (bsw_warp16.h) -------------------------------------------------------------------
struct PairArgs {
  const DevJob* ja[8];    // per group (null: idle half / group)
  const DevJob* jb[8];
  DevResult* oa[8];
  DevResult* ob[8];
  uint32_t* pdirs[8];
  SeqStore store;
  void* smem;
  int c, lg;
  bool dirs_on;
  uint32_t nbits[8];      // OR over the group's lanes of job_n_bits (checked by the harness)
};

template <int C, int LG>
void body_pair(SimWarp& w, void* p) {
  PairArgs* a = (PairArgs*)p;
  WarpSmem16<C, LG>& sm = *(WarpSmem16<C, LG>*)a->smem;
  const int grp = w.lane() / LG, gl = w.lane() % LG;
  // the N pre-scan the kernel does (every lane a share of the mask words, OR-reduced over the group)
  uint32_t nb = job_n_bits(a->store, a->ja[grp], gl, LG) | job_n_bits(a->store, a->jb[grp], gl, LG);
  for (int d = 1; d < LG; d <<= 1) nb |= (uint32_t)w.shfl_xor((int)nb, d, 32);
  if (gl == 0) a->nbits[grp] = nb;
  int any = nb != 0;
  for (int d = 1; d < 32; d <<= 1) any |= w.shfl_xor(any, d, 32);
  if (any) return;
  if (a->dirs_on) warp_align16<C, LG, true>(w, a->ja[grp], a->jb[grp], a->store, sm, a->pdirs[grp], a->oa[grp], a->ob[grp]);
  else warp_align16<C, LG, false>(w, a->ja[grp], a->jb[grp], a->store, sm, a->pdirs[grp], a->oa[grp], a->ob[grp]);
}

template <int C, int LG>
void dispatch_pair(PairArgs& a, bool desc) {
  std::vector<uint64_t> smem((sizeof(WarpSmem16<C, LG>) + 7) / 8);
  {  // garbage, like a reused block (SIM_SMEM_SEED varies it)
    const char* e = getenv("SIM_SMEM_SEED");
    uint64_t x = 0x9e3779b97f4a7c15ull * (uint64_t)(e ? atoll(e) + 1 : 1);
    for (auto& v : smem) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; v = x; }
  }
  a.smem = smem.data();
  Sched* s = new Sched();
  s->run(body_pair<C, LG>, &a, desc, 32);
  delete s;
}
template <int LG>
void run_pair_lg(PairArgs& a, bool desc) {
  switch (a.c) {
#define CASE(N) case N: dispatch_pair<N, LG>(a, desc); break;
    CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13)
    CASE(14) CASE(15) CASE(16) CASE(17) CASE(18)
#undef CASE
    default: break;
  }
}
#if defined(SIM_PART) && SIM_PART == 0
extern template void run_pair_lg<4>(PairArgs&, bool);
extern template void run_pair_lg<8>(PairArgs&, bool);
extern template void run_pair_lg<16>(PairArgs&, bool);
extern template void run_pair_lg<32>(PairArgs&, bool);
#elif defined(SIM_PART) && SIM_PART == 8
template void run_pair_lg<4>(PairArgs&, bool);
#elif defined(SIM_PART) && SIM_PART == 9
template void run_pair_lg<8>(PairArgs&, bool);
#elif defined(SIM_PART) && SIM_PART == 10
template void run_pair_lg<16>(PairArgs&, bool);
#elif defined(SIM_PART) && SIM_PART == 11
template void run_pair_lg<32>(PairArgs&, bool);
#endif
#if !defined(SIM_PART) || SIM_PART == 0
void run_pair(PairArgs& a, bool desc) {
  if (a.lg == 32) run_pair_lg<32>(a, desc);
  else if (a.lg == 16) run_pair_lg<16>(a, desc);
  else if (a.lg == 8) run_pair_lg<8>(a, desc);
  else run_pair_lg<4>(a, desc);
}
#endif

}  // namespace simns
using namespace simns;

#if !defined(SIM_PART) || SIM_PART == 0
extern "C" {

// One job through packing -> prepare_job -> kernel body (simulated warp or generic thread)
// -> finalize_result.  a/b are the CONTIGS (codes 0..4); the job's views are taken from them.
// force_class: 0 = as classified, 2 = force the generic body.  lane_order: 0 ascending, 1 descending.
// ops_out receives one byte per op (FULL mode).  Returns the class that ran.
int sim_align(const uint8_t* a, uint64_t a_clen, int a_rc, uint64_t a_off, uint64_t a_len,
              const uint8_t* b, uint64_t b_clen, int b_rc, uint64_t b_off, uint64_t b_len,
              uint64_t begin_a, uint64_t end_a, uint64_t begin_b, uint64_t end_b, uint64_t band,
              int64_t gap, int fs, int fe, int mode, int force_class, int lane_order,
              gamx_result* result, uint8_t* ops_out, uint64_t ops_out_cap) {
  HostStore hs;
  const int64_t ia = hs.add(a, a_clen), ib = hs.add(b, b_clen);
  if (a_len == UINT64_MAX) a_len = a_clen - a_off;
  if (b_len == UINT64_MAX) b_len = b_clen - b_off;
  const SeqView va = make_view(hs.start[ia], a_clen, a_rc != 0, a_off);
  const SeqView vb = make_view(hs.start[ib], b_clen, b_rc != 0, b_off);
  Prepared P;
  GenJob GJ;
  memset(&GJ, 0, sizeof(GJ));
  prepare_job(P, &GJ, va, a_len, vb, b_len, begin_a, end_a, begin_b, end_b, band, gap, fs != 0, fe != 0, mode);
  if (force_class == kClassGeneric && (P.cls == kClassWarp || P.cls == kClassCta)) {
    // re-prepare as generic
    P.cls = kClassGeneric;
    GenJob& g = GJ;
    memset(&g, 0, sizeof(g));
    g.a = va; g.b = vb; g.la = a_len; g.lb = b_len;
    g.begin_a = begin_a; g.end_a = end_a; g.begin_b = begin_b; g.end_b = end_b;
    g.band = band; g.gap = gap; g.force_start = fs; g.force_end = fe; g.mode = mode;
    g.ops_cap = P.ops_cap; g.x_size = P.x_size;
    P.dir_words = (P.x_size * (2 * band + 1) + 15) / 16;
    P.gen_rows = 2 * (2 * band + 1);
  }
  if (force_class == kClassCta && P.cls == kClassWarp) {
    // a warp-class job on the CTA kernel (what the dispatcher does for small batches of long pairs)
    P.cls = kClassCta;
    cta_geometry_for_latency(band, &P.c, &P.lg);
    P.dir_words = (mode == kModeScore) ? 0 : k1_dir_words((int)P.x_size, (int)band, P.c, P.lg);
  }
  SeqStore st{hs.packed.data(), hs.nmask.data()};
  DevResult dr;
  memset(&dr, 0, sizeof(dr));
  std::vector<uint32_t> ops(P.ops_cap / 16 + 1, 0u);
  if (P.cls == kClassWarp || P.cls == kClassCta) {
    const int G = P.lg >= 32 ? 1 : 32 / P.lg;
    std::vector<uint32_t> dirs((P.dir_words + 1) * G, 0xdeadbeefu);
    P.dj.ops_word = 0;
    WarpArgs wa;
    memset(&wa, 0, sizeof(wa));
    const int g = (lane_order >> 1) % G;  // which group of the warp runs the job; the others idle
    wa.job[g] = &P.dj; wa.out[g] = &dr;
    wa.store = st; wa.dirs = dirs.data(); wa.group_stride = P.dir_words + 1; wa.ops = ops.data();
    wa.c = P.c; wa.lg = P.lg; wa.dirs_on = mode != kModeScore;
    run_warp(wa, (lane_order & 1) != 0);
  } else if (P.cls == kClassGeneric) {
    std::vector<int64_t> rows(P.gen_rows + 1, 0x5a5a5a5a5a5a5a5aLL);
    std::vector<uint32_t> dirs(P.dir_words + 1, 0xdeadbeefu);
    GJ.ops_word = 0;
    generic_align(GJ, st, rows.data(), dirs.data(), ops.data(), dr);
  }
  finalize_result(P, &dr, mode, result);
  if (result->status == GAMX_JOB_OK && mode == kModeFull && ops_out) {
    for (uint64_t k = 0; k < result->n_ops && k < ops_out_cap; k++) {
      const uint64_t g = result->ops_offset + k;
      ops_out[k] = (uint8_t)((ops[g >> 4] >> (2 * (g & 15))) & 3u);
    }
  }
  return P.cls;
}


// Up to 4 regular jobs with the same band in ONE simulated warp (one per lane group), to exercise
// groups of different length sharing a step loop.  Jobs are full-contig views.  Returns the number
// of jobs that ran on the warp kernel (others are skipped and get status -1).
int sim_align_multi(int n_jobs, const uint8_t* const* a, const uint64_t* la, const uint8_t* const* b,
                    const uint64_t* lb, const uint64_t* begin_a, const uint64_t* end_a, const uint64_t* begin_b,
                    const uint64_t* end_b, uint64_t band, int64_t gap, const int* fs, const int* fe, int mode,
                    int lane_order, gamx_result* results, uint8_t* const* ops_out, uint64_t ops_out_cap) {
  HostStore hs;
  std::vector<Prepared> P(n_jobs);
  std::vector<DevResult> dr(n_jobs);
  int c = 0, lg = 0, ran = 0;
  uint64_t stride = 0, ops_words = 0;
  for (int k = 0; k < n_jobs; k++) {
    const int64_t ia = hs.add(a[k], la[k]), ib = hs.add(b[k], lb[k]);
    prepare_job(P[k], nullptr, make_view(hs.start[ia], la[k], false, 0), la[k], make_view(hs.start[ib], lb[k], false, 0),
                lb[k], begin_a[k], end_a[k], begin_b[k], end_b[k], band, gap, fs[k] != 0, fe[k] != 0, mode);
    memset(&dr[k], 0, sizeof(DevResult));
    if (P[k].cls == kClassWarp) {
      c = P[k].c; lg = P[k].lg;
      stride = std::max<uint64_t>(stride, P[k].dir_words + 1);
      P[k].dj.ops_word = ops_words;
      ops_words += P[k].ops_cap / 16 + 1;
    }
  }
  std::vector<uint32_t> ops(ops_words + 1, 0u);
  SeqStore st{hs.packed.data(), hs.nmask.data()};
  if (c) {
    const int G = 32 / lg;
    std::vector<uint32_t> dirs(stride * G + 1, 0xdeadbeefu);
    WarpArgs wa;
    memset(&wa, 0, sizeof(wa));
    int g = 0;
    for (int k = 0; k < n_jobs && g < G; k++)
      if (P[k].cls == kClassWarp) { wa.job[g] = &P[k].dj; wa.out[g] = &dr[k]; g++; ran++; }
    wa.store = st; wa.dirs = dirs.data(); wa.group_stride = stride; wa.ops = ops.data();
    wa.c = c; wa.lg = lg; wa.dirs_on = mode != kModeScore;
    run_warp(wa, (lane_order & 1) != 0);
  }
  int placed = 0;
  for (int k = 0; k < n_jobs; k++) {
    if (P[k].cls == kClassWarp && placed < 32 / (lg ? lg : 32)) {
      placed++;
      finalize_result(P[k], &dr[k], mode, &results[k]);
      if (results[k].status == GAMX_JOB_OK && mode == kModeFull && ops_out && ops_out[k])
        for (uint64_t q = 0; q < results[k].n_ops && q < ops_out_cap; q++) {
          const uint64_t gpos = results[k].ops_offset + q;
          ops_out[k][q] = (uint8_t)((ops[gpos >> 4] >> (2 * (gpos & 15))) & 3u);
        }
    } else {
      memset(&results[k], 0, sizeof(gamx_result));
      results[k].status = -1;
    }
  }
  return ran;
}

// Jobs 2g and 2g+1 run as the 16x2 PAIR of lane group g of ONE simulated warp (bsw_warp16.h), then the
// traceback walks each job's half of the pair region (PairFetch), like k1s_kernel + tb_kernel do.  All
// jobs take the same band and gap and full-contig views; an odd job count leaves the last high half idle.
// first_group: lane group of pair 0 (groups before it idle).  Returns the number of jobs that ran, or
// -1 when a job is not a warp-kernel job, -2 when a window holds an N (the kernel would fall back).
int sim_align_pairs(int n_jobs, const uint8_t* const* a, const uint64_t* la, const uint8_t* const* b,
                    const uint64_t* lb, const uint64_t* begin_a, const uint64_t* end_a, const uint64_t* begin_b,
                    const uint64_t* end_b, uint64_t band, int64_t gap, const int* fs, const int* fe, int mode,
                    int lane_order, int first_group, int rc, gamx_result* results, uint8_t* const* ops_out, uint64_t ops_out_cap) {
  // rc: bit 0 / bit 1 = the jobs see the reverse complement of the stored a / b contigs
  HostStore hs;
  std::vector<Prepared> P(n_jobs);
  std::vector<DevResult> dr(n_jobs);
  int c = 0, lg = 0;
  uint64_t xmax = 0, ops_words = 0;
  for (int k = 0; k < n_jobs; k++) {
    const int64_t ia = hs.add(a[k], la[k]), ib = hs.add(b[k], lb[k]);
    prepare_job(P[k], nullptr, make_view(hs.start[ia], la[k], (rc & 1) != 0, 0), la[k], make_view(hs.start[ib], lb[k], (rc & 2) != 0, 0),
                lb[k], begin_a[k], end_a[k], begin_b[k], end_b[k], band, gap, fs[k] != 0, fe[k] != 0, mode);
    memset(&dr[k], 0, sizeof(DevResult));
    if (P[k].cls != kClassWarp) return -1;
    c = P[k].c; lg = P[k].lg;
    xmax = std::max(xmax, P[k].x_size);
    P[k].dj.ops_word = ops_words;
    ops_words += P[k].ops_cap / 16 + 1;
  }
  const int G = 32 / lg;
  if (first_group + (n_jobs + 1) / 2 > G) return -1;
  const uint64_t stride = k1_dir_words16((int)xmax, c, lg);  // per pair
  std::vector<uint32_t> ops(ops_words + 1, 0u);
  std::vector<uint32_t> dirs(stride * G + 1, 0xdeadbeefu);
  SeqStore st{hs.packed.data(), hs.nmask.data()};
  PairArgs pa;
  memset(&pa, 0, sizeof(pa));
  for (int k = 0; k < n_jobs; k++) {
    const int g = first_group + k / 2;
    if (k & 1) { pa.jb[g] = &P[k].dj; pa.ob[g] = &dr[k]; }
    else { pa.ja[g] = &P[k].dj; pa.oa[g] = &dr[k]; }
  }
  for (int g = 0; g < G; g++) pa.pdirs[g] = dirs.data() + (uint64_t)g * stride;  // (idle groups: their sink)
  pa.store = st; pa.c = c; pa.lg = lg; pa.dirs_on = mode != kModeScore;
  run_pair(pa, (lane_order & 1) != 0);
  for (int g = 0; g < G; g++) if (pa.nbits[g]) return -2;
  for (int k = 0; k < n_jobs; k++) {
    DevResult& R = dr[k];
    if (mode != kModeScore && R.status == kStatusOk) {
      if (R.has_match != 1 + (k & 1)) return -3;  // layout tag
      PairFetch f{dirs.data() + (uint64_t)(first_group + k / 2) * stride, c, lg, k & 1};
      k1_traceback_t(f, c, R.end_i, R.end_j, P[k].dj.p0, mode == kModeFull, ops.data() + P[k].dj.ops_word, P[k].dj.ops_cap, R);
      R.ops_start = P[k].dj.ops_word * 16 + P[k].dj.ops_cap - R.n_ops;
    }
    finalize_result(P[k], &R, mode, &results[k]);
    if (results[k].status == GAMX_JOB_OK && mode == kModeFull && ops_out && ops_out[k])
      for (uint64_t q = 0; q < results[k].n_ops && q < ops_out_cap; q++) {
        const uint64_t gpos = results[k].ops_offset + q;
        ops_out[k][q] = (uint8_t)((ops[gpos >> 4] >> (2 * (gpos & 15))) & 3u);
      }
  }
  return n_jobs;
}

void sim_band_geometry(uint64_t band, int* c, int* lg) { geometry_for_band(band, true, c, lg); }

}  // extern "C"
#endif  // glue part
