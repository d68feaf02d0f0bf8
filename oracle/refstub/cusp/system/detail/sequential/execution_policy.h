// Stub: the sequential back-end is thrust::cpp.
#pragma once
#include <thrust/system/cpp/execution_policy.h>
