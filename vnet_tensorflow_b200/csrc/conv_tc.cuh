// tcgen05 implicit-GEMM convolution path (placeholder plan types; kernels land in a later commit).
#pragma once
#include "vnb_cuda.h"
namespace vnb {
struct TcKernelPlan {
  bool valid = false;
};
struct TcConvPlan {
  TcKernelPlan fprop, dgrad, wgrad;
};
}  // namespace vnb
