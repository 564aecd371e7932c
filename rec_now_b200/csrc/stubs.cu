// TEMPORARY: entry points not implemented yet return RN_ERR_UNSUPPORTED (replaced file by file).
#include "common.cuh"
extern "C" size_t rn_pair_indices_scratch_bytes(int64_t, int32_t) { return 0; }
extern "C" int rn_pair_indices_count(const rn_pairwise_args*, int32_t, void*, size_t, int64_t*, void*) { return RN_ERR_UNSUPPORTED; }
extern "C" int rn_pair_indices_fill(const rn_pairwise_args*, int32_t, void*, size_t, int32_t*, int32_t*, float*, int64_t, void*) { return RN_ERR_UNSUPPORTED; }
extern "C" size_t rn_occurrence_scratch_bytes(int64_t) { return 0; }
extern "C" int rn_occurrence_power_weight(const int64_t*, int64_t, float, float*, void*, size_t, void*) { return RN_ERR_UNSUPPORTED; }
extern "C" size_t rn_listwise_scratch_bytes(int64_t) { return 0; }
extern "C" int rn_listwise_fwd_bwd(const rn_listwise_args*, void*, size_t, void*) { return RN_ERR_UNSUPPORTED; }
extern "C" int rn_listwise_dense(const rn_listwise_args*, void*, size_t, int64_t, uint8_t*, float*, float*, int32_t, float, void*) { return RN_ERR_UNSUPPORTED; }
extern "C" int rn_listwise_launch_count(int64_t) { return 0; }
