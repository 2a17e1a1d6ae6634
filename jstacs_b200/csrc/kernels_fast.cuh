// kernels_fast.cuh — scaled linear-space kernels (placeholder: tables only; kernels follow).
#pragma once
#include "common.cuh"

struct jh_dataset;
template <typename T> struct DevBuf;

namespace jhf {

struct FastTables {
  bool usable = false;
};

inline int build(FastTables &F, int, int, int, bool, int, const std::vector<int> &, const std::vector<int> &, const std::vector<int> &,
                 const std::vector<int> &, const std::vector<int> &, const std::vector<int> &, const std::vector<int> &,
                 const std::vector<int> &, const std::vector<int> &, const std::vector<uint8_t> &, int, int, cudaStream_t) {
  F.usable = false;
  return JH_OK;
}
inline int prepare(FastTables &, const DevModel &, cudaStream_t) { return JH_OK; }
inline void destroy(FastTables &) {}
inline int run_loglik(FastTables &, const DevModel &, const DevData &, int64_t, int64_t, double *, unsigned long long *, int, cudaStream_t) {
  return JH_ERR_UNSUPPORTED;
}
inline int run_estep(FastTables &, const DevModel &, const jh_dataset &, double *, unsigned long long *, int32_t *, DevBuf<double> &, int,
                     int64_t, int, cudaStream_t) {
  return JH_ERR_UNSUPPORTED;
}

}  // namespace jhf
