// Host build of the range-space solver core that prints the working-set changes of ONE QP (which row enters or leaves at every
// trip of the loop, and the pairs the set-up pass commits): what the long QPs of a batch do.  Driven by tools/trace_qp.py.
// A developer tool; nothing in the product library links or loads it.
#include <cstdio>
#include <cstdint>
#include <cstring>
#include "../quadruped_control_b200/csrc/qpb_tpq_core.h"
using namespace qpb::tpq;
// trace the loop of one QP at LPQ=1
extern "C" int trace_one(const qpb_params* P, const qpb_state_rec* in, int verbose) {
  FastParams K; make_fast_params(*P, K);
  const double* rec = reinterpret_cast<const double*>(in);
  uint32_t cbytes, hint; memcpy(&cbytes, in->contact, 4); hint = 0;
  State st, keep; double b6[6], G[21], Gkeep[21]; uint32_t key = 0;
  auto commit = [&](const State& s, const double (&g)[21], uint32_t k) { keep = s; memcpy(Gkeep, g, sizeof(Gkeep)); key = k; if (verbose) printf("commit iters=%d word=%06x nact=%d key=%08x\n", s.iters, s.word, __builtin_popcount(s.word), k); };
  setup(*P, K, rec, cbytes, hint, st, b6, G, commit);
  st = keep; memcpy(G, Gkeep, sizeof(G));
  if (!(st.status == QPB_OK && key)) return st.iters;
  double side[kSideSize] = {}; memcpy(side + kSideG, G, sizeof(G)); memcpy(side + kSideR, st.r, sizeof(st.r));
  Lane<4> ln; lane_init(ln, 0, st.f, st.r, st.u, st.word, st.stance, st.status, st.iters, key);
  ln.sp = row_slack_share(K, ln, 0);
  while (!ln.done) {
    StepTmp<4> T; double ub, rb; int kb;
    uint32_t w0 = ln.word; int p0 = ln.p; uint32_t pc0 = ln.pc;
    direction(K, ln, 0, side, T, ub, rb, kb);
    advance(K, ln, 0, side, T, ub, rb, kb, true);
    if (verbose) {
      if (ln.word & ~w0) printf("  it %2d ADD  row %2d%c \n", ln.iters, p0, pc0 == 1 ? 'A' : 'B');
      else printf("  it %2d DROP row %2d (pending %2d%c)\n", ln.iters, kb, p0, pc0 == 1 ? 'A' : 'B');
    }
    uint32_t best = select_local(K, ln, 0); bool fresh;
    double s = select_commit(K, ln, 0, best, fresh); if (fresh) ln.sp = s;
  }
  if (verbose) printf("final word=%06x nact=%d iters=%d\n", ln.word, __builtin_popcount(ln.word), ln.iters);
  return ln.iters;
}

// working-set changes spent by the set-up pass (block rounds) and in total, for n records, without printing
extern "C" void trace_counts(const qpb_params* P, const qpb_state_rec* in, int64_t n, int* setup_rounds, int* total) {
  FastParams K;
  make_fast_params(*P, K);
  for (int64_t q = 0; q < n; q++) {
    const double* rec = reinterpret_cast<const double*>(in + q);
    uint32_t cbytes;
    memcpy(&cbytes, in[q].contact, 4);
    State st, keep;
    double b6[6], G[21];
    uint32_t key = 0;
    auto commit = [&](const State& s, const double (&)[21], uint32_t k) { keep = s; key = k; };
    setup(*P, K, rec, cbytes, 0u, st, b6, G, commit);
    setup_rounds[q] = keep.iters;
    total[q] = key ? trace_one(P, in + q, 0) : keep.iters;
  }
}
