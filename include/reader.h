/*
 * reader.h -- shim for the reference's include/reader.h: mmread_bmSparse (src/reader.cu:49-110) over the C ABI.
 * The reference fills four thrust::device_vectors through O(nnz * 64) single-element device inserts; here the file is
 * parsed on the host once and converted on the device; the result is returned as a bmSpMatrix whose keys/bmps/offsets/
 * values members are the four vectors.  Like the reference function it builds the fp16, non-transposed form and it
 * does NOT mirror `symmetric` files (reader.cu ignores the banner) -- use bmSpMatrix(path, transpose) for that.
 */
#ifndef READER_HPP_
#define READER_HPP_
#include <tuple>
#include <string>
#include <fstream>
#include <sstream>
#include <vector>
#include "bmSpMatrix.h"

/* returns (rows, cols, entry lines) like the reference's thrust::tuple<int,int,int> */
inline std::tuple<int, int, int> mmread_bmSparse(const std::string& path, bmSpMatrix<bmsp::half_t>& out, void* stream = nullptr) {
    std::ifstream f(path.c_str());
    if (!f) throw std::runtime_error("mmread_bmSparse: cannot open " + path);
    std::string line;
    while (f.peek() == '%') std::getline(f, line);
    int nr = 0, nc = 0, nl = 0;
    f >> nr >> nc >> nl;
    std::vector<int32_t> r(nl), c(nl); std::vector<double> v(nl);
    for (int i = 0; i < nl; i++) { float t; f >> r[i] >> c[i] >> t; r[i]--; c[i]--; v[i] = t; }   /* text -> float, reader.cu:70-75 */
    bmsp_matrix_t h = nullptr;
    bmsp::check(bmsp_create_from_coo(nr, nc, nl, r.data(), c.data(), v.data(), 0, BMSP_F16, stream, &h));
    out.adopt(h);
    return std::make_tuple(nr, nc, nl);
}

#ifdef BMSP_HAVE_THRUST
#include <thrust/tuple.h>
#include <cuda_fp16.h>
/* The reference's signature, include/reader.h:14-15 (defined in src/reader.cu:49-110): fills the caller's four device vectors. */
typedef thrust::device_vector<uint64_t> uint64_vec;
typedef thrust::device_vector<__half> half_vec;
inline thrust::tuple<int, int, int> mmread_bmSparse(std::string path, uint64_vec& keys, uint64_vec& bmps, uint64_vec& offsets, half_vec& values) {
    bmSpMatrix<bmsp::half_t> m;
    const std::tuple<int, int, int> dims = mmread_bmSparse(path, m);
    keys.resize(m.keys.size()); bmps.resize(m.bmps.size()); offsets.resize(m.offsets.size()); values.resize(m.values.size());
    cudaMemcpy(thrust::raw_pointer_cast(keys.data()), m.keys.data(), 8 * m.keys.size(), cudaMemcpyDeviceToDevice);
    cudaMemcpy(thrust::raw_pointer_cast(bmps.data()), m.bmps.data(), 8 * m.bmps.size(), cudaMemcpyDeviceToDevice);
    cudaMemcpy(thrust::raw_pointer_cast(offsets.data()), m.offsets.data(), 8 * m.offsets.size(), cudaMemcpyDeviceToDevice);
    cudaMemcpy(thrust::raw_pointer_cast(values.data()), m.values.data(), 2 * m.values.size(), cudaMemcpyDeviceToDevice);
    return thrust::make_tuple(std::get<0>(dims), std::get<1>(dims), std::get<2>(dims));
}
#endif
#endif /* READER_HPP_ */
