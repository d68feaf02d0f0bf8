// Stub standing in for cusp/detail/config.h when compiling the reference's host CSR kernels
// from /root/reference with the CUDA 12.9 Thrust (bundled cusp does not build against it).
// Written for this repo; contains no reference code.
#pragma once
#include <thrust/detail/config.h>
#include <thrust/system/cpp/execution_policy.h>
#include <thrust/functional.h>
#include <cstddef>
