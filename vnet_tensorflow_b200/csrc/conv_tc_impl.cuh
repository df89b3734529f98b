// tcgen05 implicit-GEMM convolution path: engine glue (filled in by the tensor-core kernel commit).
#pragma once
#include "engine.cuh"
namespace vnb {
inline void Engine::tc_setup() {}
inline void Engine::tc_prepare_weights() {}
inline void Engine::tc_run_fprop(Unit&, int) {}
inline void Engine::tc_run_dgrad(Unit&, int) {}
inline void Engine::tc_run_wgrad(Unit&, int) {}
inline void tc_op_conv5(int, const float*, const float*, const float*, const float*, float*, int, Dims, int, int, bool) {
  throw std::invalid_argument("tensor-core convolution path not built yet");
}
inline void tc_op_wgrad5(int, const float*, const float*, float*, int, Dims, int, int) {
  throw std::invalid_argument("tensor-core wgrad path not built yet");
}
}  // namespace vnb
