// Host build of the thread-per-QP solver core (quadruped_control_b200/csrc/qpb_tpq_core.h) for the CPU tests:
// the same source the CUDA kernel compiles, run one QP at a time, so the algorithm can be checked against the oracle
// without a GPU.  The LPQ lanes that share a QP on the device are stepped here in lock step, and the three warp
// exchanges of the kernel (integer max, sum, fraction min over the lanes of a QP) are plain loops.
// TEST INFRASTRUCTURE ONLY: nothing in the product library links or loads this.
#include <cstdint>
#include <cstring>

#include "../../quadruped_control_b200/csrc/qpb_tpq_core.h"

using namespace qpb::tpq;

template <int LPQ>
static void solve_loop(const FastParams& K, State& st, const double* G0, uint32_t key) {
  constexpr int LPL = 4 / LPQ;
  double side[kSideSize] = {};
  std::memcpy(side + kSideG, G0, 21 * sizeof(double));
  std::memcpy(side + kSideR, st.r, 12 * sizeof(double));
  double* G = side + kSideG;
  Lane<LPL> ln[LPQ];
  double slack = 0.0;
  for (int j = 0; j < LPQ; j++) {
    lane_init(ln[j], j, st.f, st.r, st.u, st.word, st.stance, st.status, st.iters, key);
    slack += row_slack_share(K, ln[j], j);
  }
  for (int j = 0; j < LPQ; j++) ln[j].sp = slack;
  // a trip: direction -> (exchange: blocking row) -> step + working-set change -> selection of the next row
  // (exchanges: most violated row, its slack).  The loop is entered with the first row already chosen.
  while (!ln[0].done) {
    StepTmp<LPL> T[LPQ];
    double ub = 1.0, rb = 0.0;
    int kb = -1;
    for (int j = 0; j < LPQ; j++) {
      double u1, r1;
      int k1;
      direction(K, ln[j], j, side, T[j], u1, r1, k1);
      if (j == 0) { ub = u1; rb = r1; kb = k1; }
      else better_ratio(ub, rb, kb, u1, r1, k1);
    }
    double Gnew[21];
    std::memcpy(Gnew, G, sizeof(Gnew));
    for (int j = 0; j < LPQ; j++) {  // every lane reads the old G; one writes the new one
      double sj[kSideSize];
      std::memcpy(sj, side, sizeof(sj));
      advance(K, ln[j], j, sj, T[j], ub, rb, kb, true);
      if (j == 0) std::memcpy(Gnew, sj + kSideG, sizeof(Gnew));
    }
    std::memcpy(G, Gnew, sizeof(Gnew));
    uint32_t best = 0;
    for (int j = 0; j < LPQ; j++) {
      const uint32_t k = select_local(K, ln[j], j);
      best = k > best ? k : best;
    }
    double s2 = 0.0;
    bool fresh[LPQ];
    for (int j = 0; j < LPQ; j++) s2 += select_commit(K, ln[j], j, best, fresh[j]);
    for (int j = 0; j < LPQ; j++)
      if (fresh[j]) ln[j].sp = s2;
  }
  st.word = ln[0].word;
  st.status = ln[0].status;
  st.iters = ln[0].iters;
}

extern "C" int tpq_host_control_batch(const qpb_params* P, const qpb_state_rec* in, int64_t n, qpb_out_rec* out,
                                      int lpq /* lanes per QP: 1, 2 or 4 */, int do_polish) {
  FastParams K;
  if (!make_fast_params(*P, K)) return -1;
  for (int64_t i = 0; i < n; i++) {
    const double* rec = reinterpret_cast<const double*>(in + i);
    uint32_t cbytes, hint;
    std::memcpy(&cbytes, in[i].contact, 4);
    std::memcpy(&hint, in[i].pad, 4);
    State st, keep;
    double b6[6], G[21], Gkeep[21];
    uint32_t key = 0;
    auto commit = [&](const State& s, const double (&g)[21], uint32_t k) {
      keep = s;
      std::memcpy(Gkeep, g, sizeof(Gkeep));
      key = k;
    };
    setup(*P, K, rec, cbytes, hint, st, b6, G, commit);
    st = keep;  // the loop starts from the last dual-feasible pair the set-up committed
    std::memcpy(G, Gkeep, sizeof(G));
    if (st.status == QPB_OK && key != 0u) {  // the set-up's pair is not optimal yet: the loop (the kernel's worklist)
      if (lpq == 4) solve_loop<4>(K, st, G, key);
      else if (lpq == 2) solve_loop<2>(K, st, G, key);
      else solve_loop<1>(K, st, G, key);
    }
    if (do_polish) polish(K, st, b6);
    double grf[12], tau[12];
    finish(*P, rec, rec + qpb::kQ, st, grf, tau);
    std::memset(&out[i], 0, sizeof(qpb_out_rec));
    std::memcpy(out[i].grf_body, grf, sizeof(grf));
    std::memcpy(out[i].tau, tau, sizeof(tau));
    out[i].status = st.status;
    out[i].iters = st.iters;
    const uint32_t word = st.word | 0x80000000u;
    std::memcpy(out[i].pad, &word, 4);
  }
  return 0;
}
