// qpb_internal.h -- shared between the translation units of libqpb200.so (not part of the ABI).
#pragma once
#include <string>

// records msg as the calling thread's qpb_last_error() text and returns code
__attribute__((visibility("hidden"))) int qpb_internal_fail(int code, const std::string& msg);
