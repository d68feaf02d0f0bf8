// Stub: the format tags the reference's multiply overloads dispatch on.
#pragma once
namespace cusp {
struct known_format {};
struct csr_format : known_format {};
struct coo_format : known_format {};
struct array1d_format : known_format {};
}
