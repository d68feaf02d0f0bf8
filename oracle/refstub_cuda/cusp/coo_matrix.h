// Minimal stand-in for the cusp headers the reference's bmSparse sources include, so that those sources
// compile from /root/reference against the CUDA 12.9 Thrust (bundled cusp does not).  Written for this repo
// following SURVEY.md Appendix C; contains no reference code.  Only the members the reference touches exist.
#pragma once
#include <thrust/host_vector.h>
#include <thrust/device_vector.h>
#include <thrust/sort.h>
#include <thrust/remove.h>
#include <thrust/reduce.h>
#include <thrust/transform.h>
#include <thrust/scatter.h>
#include <thrust/gather.h>
#include <thrust/scan.h>
#include <thrust/execution_policy.h>
#include <thrust/functional.h>
#include <thrust/tuple.h>
#include <thrust/iterator/constant_iterator.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/discard_iterator.h>
#include <thrust/iterator/transform_iterator.h>
#include <thrust/iterator/zip_iterator.h>
#include <iostream>
#include <string>
#include <vector>

namespace cusp {
struct host_memory {};
struct device_memory {};
template <typename T, typename M> struct vec_of { typedef thrust::host_vector<T> type; };
template <typename T> struct vec_of<T, device_memory> { typedef thrust::device_vector<T> type; };

template <typename I, typename V, typename M>
struct coo_matrix {
    size_t num_rows = 0, num_cols = 0, num_entries = 0;
    typename vec_of<I, M>::type row_indices, column_indices;
    typename vec_of<V, M>::type values;
    coo_matrix() {}
    template <typename M2>
    coo_matrix(const coo_matrix<I, V, M2>& o)
        : num_rows(o.num_rows), num_cols(o.num_cols), num_entries(o.num_entries), row_indices(o.row_indices),
          column_indices(o.column_indices), values(o.values) {}
};
template <typename I, typename V, typename M>
struct csr_matrix {
    size_t num_rows = 0, num_cols = 0, num_entries = 0;
    typename vec_of<I, M>::type row_offsets, column_indices;
    typename vec_of<V, M>::type values;
};
}  // namespace cusp
