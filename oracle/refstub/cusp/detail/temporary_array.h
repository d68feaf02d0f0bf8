// Stub: cusp::detail::temporary_array<T,Policy>(exec, n, init) as a std::vector.
#pragma once
#include <vector>
#include <cstddef>
namespace cusp { namespace detail {
template <typename T, typename Policy>
struct temporary_array : std::vector<T> {
    template <typename Exec> temporary_array(Exec&, size_t n, T init = T()) : std::vector<T>(n, init) {}
};
}}
