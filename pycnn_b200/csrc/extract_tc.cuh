// extract_tc.cuh -- tensor-core form of the fused extractor (stub until the tcgen05 kernel lands).
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace pb {
static inline bool tc_extract_supported(const pycnn_b200_ctx*, int) { return false; }
static inline int tc_extract(pycnn_b200_ctx*, const uint8_t*, int64_t, int, float*, int64_t) {
  return fail("tensor-core extractor not built");
}
static inline int tc_extract_set_filters(pycnn_b200_ctx*) { return 0; }
}  // namespace pb
