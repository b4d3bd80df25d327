/* ASan / UBSan driver for the CPU oracle: reads parameter blocks and records the test wrote
 * to a file, runs every batched oracle entry point on them, and writes the results back so the test can compare them
 * with the optimised build.  Built by tests/test_oracle_sanitizers.py with
 *   gcc -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=all
 * TEST INFRASTRUCTURE ONLY. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../oracle/mpc_oracle.h"
#include "../../oracle/plan_oracle.h"
#include "../../oracle/qpb_oracle.h"

static void* slurp(FILE* f, size_t bytes) {
  void* p = malloc(bytes ? bytes : 1);
  if (!p || fread(p, 1, bytes, f) != bytes) {
    fprintf(stderr, "short read\n");
    exit(3);
  }
  return p;
}

int main(int argc, char** argv) {
  if (argc != 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  int64_t hdr[2];  /* balance records, MPC records */
  if (fread(hdr, sizeof(hdr), 1, f) != 1) return 3;
  const int64_t n = hdr[0], nm = hdr[1];
  orc_params* P = slurp(f, sizeof(orc_params));
  orc_state* S = slurp(f, (size_t)n * sizeof(orc_state));
  orc_swing* SW = slurp(f, (size_t)n * sizeof(orc_swing));
  orc_mpc_params* MP = slurp(f, sizeof(orc_mpc_params));
  orc_mpc_rec* MR = slurp(f, (size_t)nm * sizeof(orc_mpc_rec));
  fclose(f);
  orc_out* out = calloc((size_t)n, sizeof(orc_out));
  orc_out* tick = calloc((size_t)n, sizeof(orc_out));
  orc_mpc_out* mout = calloc((size_t)nm, sizeof(orc_mpc_out));
  orc_control_batch(P, S, n, out, 2);
  orc_joint_gains g;
  orc_default_joint_gains(&g);
  orc_tick_batch(P, &g, S, SW, n, tick, 2);
  orc_mpc_batch(MP, MR, nm, mout, 2);
  FILE* o = fopen(argv[2], "wb");
  if (!o) return 2;
  fwrite(out, sizeof(orc_out), (size_t)n, o);
  fwrite(tick, sizeof(orc_out), (size_t)n, o);
  fwrite(mout, sizeof(orc_mpc_out), (size_t)nm, o);
  fclose(o);
  free(P); free(S); free(SW); free(MP); free(MR); free(out); free(tick); free(mout);
  return 0;
}
