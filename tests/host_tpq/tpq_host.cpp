// Host build of the thread-per-QP solver core (quadruped_control_b200/csrc/qpb_tpq_core.h) for the CPU tests:
// the same source the CUDA kernel compiles, run one QP at a time, so the algorithm can be checked against the oracle
// without a GPU.  TEST INFRASTRUCTURE ONLY: nothing in the product library links or loads this.
#include <cstdint>
#include <cstring>

#include "../../quadruped_control_b200/csrc/qpb_tpq_core.h"



extern "C" int tpq_host_control_batch(const qpb_params* P, const qpb_state_rec* in, int64_t n, qpb_out_rec* out,
                                      double* fw /* n x 12 world-frame or null */, int do_polish) {
  using namespace qpb::tpq;
  FastParams K;
  if (!make_fast_params(*P, K)) return -1;
  for (int64_t i = 0; i < n; i++) {
    const double* rec = reinterpret_cast<const double*>(in + i);
    uint32_t cbytes, hint;
    std::memcpy(&cbytes, in[i].contact, 4);
    std::memcpy(&hint, in[i].pad, 4);
    State st;
    double b6[6];
    setup(*P, K, rec, cbytes, hint, st, b6);
    while (!st.done) iterate(K, st);
    if (do_polish) polish(K, st, b6);
    double grf[12], tau[12];
    finish(*P, rec, rec + qpb::kQ, st, grf, tau);
    std::memset(&out[i], 0, sizeof(qpb_out_rec));
    std::memcpy(out[i].grf_body, grf, sizeof(grf));
    std::memcpy(out[i].tau, tau, sizeof(tau));
    out[i].status = st.status;
    out[i].iters = st.iters;
    const uint32_t word = wset_encode(st.sg) | 0x80000000u;
    std::memcpy(out[i].pad, &word, 4);
    if (fw) std::memcpy(fw + 12 * i, st.f, sizeof(st.f));
  }
  return 0;
}
