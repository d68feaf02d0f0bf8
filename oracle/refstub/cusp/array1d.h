// Stub: only what cusp/system/omp/detail/multiply/csr_spgemm.h needs -- the omp policy names.
#pragma once
#include <thrust/system/omp/execution_policy.h>
namespace cusp { namespace system { namespace omp {
using thrust::system::omp::execution_policy;
using thrust::system::omp::tag;
}}}
