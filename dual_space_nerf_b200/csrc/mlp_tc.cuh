// placeholder until the tcgen05 kernel lands (next commit)
#pragma once
#include <cuda_runtime.h>
#include <vector>
namespace dsn {
struct TcWeights {
  template <typename... A> int stage(A&&...) { return 0; }
  void release() {}
};
inline void tc_configure() {}
inline int tc_launch(TcWeights&, const float*, const float4*, const unsigned long long*, int64_t, float4*, float4*, int, int, cudaStream_t) {
  return (int)cudaErrorNotSupported;
}
}  // namespace dsn
