// qpb_internal.h -- shared between the translation units of libqpb200.so (not part of the ABI).
#pragma once
#include <stdint.h>

#include <string>

// records msg as the calling thread's qpb_last_error() text and returns code
__attribute__((visibility("hidden"))) int qpb_internal_fail(int code, const std::string& msg);

// one shard of a host batch of whole_n records whose first record does (whole_warm = 1) or does not (0) carry a
// warm-start word: the kernels are chosen as for the whole batch, so a record's result does not depend on the sharding
struct qpb_handle;
struct qpb_state_rec;
struct qpb_swing_rec;
struct qpb_out_rec;
extern "C" __attribute__((visibility("hidden"))) int qpb_internal_host_shard(qpb_handle* h, int64_t n, const qpb_state_rec* h_states,
                                                                             const qpb_swing_rec* h_swing, qpb_out_rec* h_out,
                                                                             int64_t whole_n, int whole_warm);
