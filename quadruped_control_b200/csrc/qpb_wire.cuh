// qpb_wire.cuh -- widen / narrow the wire records of the host-buffer calls (include/qpb200.h: qpb_wire_state 488 B,
// qpb_wire_out 200 B) to and from the device records (512 B / 256 B).  One thread per 8-byte word, fully coalesced; the
// records are moved as 64-bit integers so that every bit pattern (NaNs of a bad input included) arrives unchanged.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/qpb200.h"

namespace qpb {

static_assert(sizeof(qpb_wire_state) == 488 && sizeof(qpb_wire_out) == 200, "wire record layout");
static_assert(sizeof(qpb_state_rec) == 512 && sizeof(qpb_out_rec) == 256, "device record layout");
constexpr int kWireInWords = 61, kWireOutWords = 25;

// words 0..59: the sixty doubles; word 60: contact[4] + warm-start word = contact[4] + pad[0..3] of the device record
__global__ void wire_unpack_kernel(const uint64_t* __restrict__ in, uint64_t* __restrict__ out, int64_t n) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n * 64) return;
  const int64_t rec = g >> 6;
  const int k = (int)(g & 63);
  out[g] = k < kWireInWords ? __ldg(in + rec * kWireInWords + k) : 0ULL;
}

// words 0..23: forces and torques; word 24: status (16) | iters (16, saturated) | working-set word (32)
__global__ void wire_pack_kernel(const uint64_t* __restrict__ in, uint64_t* __restrict__ out, int64_t n) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n * kWireOutWords) return;
  const int64_t rec = g / kWireOutWords;
  const int k = (int)(g - rec * kWireOutWords);
  uint64_t v;
  if (k < 24) {
    v = __ldg(in + rec * 32 + k);
  } else {
    const uint64_t si = __ldg(in + rec * 32 + 24), ws = __ldg(in + rec * 32 + 25);
    const int32_t status = (int32_t)(uint32_t)si, iters = (int32_t)(uint32_t)(si >> 32);
    const uint32_t it16 = (uint32_t)(iters < 0 ? 0 : (iters > 32767 ? 32767 : iters));
    v = (uint64_t)((uint32_t)status & 0xffffu) | ((uint64_t)it16 << 16) | ((ws & 0xffffffffULL) << 32);
  }
  out[g] = v;
}

}  // namespace qpb
