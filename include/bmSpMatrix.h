/*
 * bmSpMatrix.h -- header-only C++ shim that re-creates the reference's class and operator names on top of the
 * C ABI in bmsparse_b200.h, so that code written against GonzaBerger/bmSparse-SPGEMM-SPMV's include/bmSpMatrix.h
 * (class bmSpMatrix<valueType>, public keys/bmps/offsets/values, num_rows/num_cols/nnz/block_num) and against the
 * operator templates bmSparse_SpMV (src/bmSparse_SPMV.cu:191) and bmSparse_mult (src/bmSparse_SPGEMM.cu:827)
 * compiles against libbmsparse_b200.so.  Differences, all deliberate:
 *   - keys/bmps/offsets/values are non-owning device views (data(), size(), raw pointer) instead of
 *     thrust::device_vector: no Thrust dependency; thrust::raw_pointer_cast(v.data()) becomes v.data();
 *   - errors throw std::runtime_error(bmsp_last_error()) instead of printing and calling exit();
 *   - operators run on an explicit stream (default 0) and do not call cudaDeviceSynchronize().
 * Link with -lbmsparse_b200.  No CUDA headers are needed to include this file.
 */
#ifndef BMSPMATRIX_H_
#define BMSPMATRIX_H_

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>
#include "bmsparse_b200.h"

/* Translation units that are compiled by nvcc and can see Thrust also get the reference's Thrust-typed signatures (the adopting
 * constructor of include/bmSpMatrix.h:33-34 and the five-argument mmread_bmSparse of include/reader.h:14-15), so that reference
 * call sites compile unchanged.  Define BMSP_NO_THRUST to keep Thrust out. */
#if !defined(BMSP_NO_THRUST) && defined(__CUDACC__) && defined(__has_include)
#if __has_include(<thrust/device_vector.h>)
#include <thrust/device_vector.h>
#define BMSP_HAVE_THRUST 1
#endif
#endif

#define BLOCK_WIDTH 8
#define BLOCK_HEIGHT 8
#define BMSP_BLOCK_SIZE (BLOCK_WIDTH * BLOCK_HEIGHT)

namespace bmsp {
struct half_t { uint16_t bits; };                     // stands in for CUDA's half when cuda_fp16.h is not included
template <class T> struct dtype_of;
template <> struct dtype_of<float> { static const int value = BMSP_F32; };
template <> struct dtype_of<half_t> { static const int value = BMSP_F16; };
#ifdef __CUDA_FP16_H__
template <> struct dtype_of<__half> { static const int value = BMSP_F16; };
#endif
inline void check(int status) { if (status != BMSP_OK) throw std::runtime_error(bmsp_last_error()); }

template <class T>
class device_view {                                    // what the reference's public device_vector members expose
public:
    device_view() : p_(nullptr), n_(0) {}
    device_view(const T* p, size_t n) : p_(p), n_(n) {}
    const T* data() const { return p_; }
    size_t size() const { return n_; }
private:
    const T* p_; size_t n_;
};
}  // namespace bmsp

template <class valueType>
class bmSpMatrix {
public:
    bmsp::device_view<uint64_t> keys, bmps, offsets;
    bmsp::device_view<valueType> values;
    int num_rows, num_cols, nnz, block_num;

    bmSpMatrix() : num_rows(0), num_cols(0), nnz(0), block_num(0), h_(nullptr) {}
    /* bmSpMatrix(std::string, bool transpose): MatrixMarket ingest, src/bmSpMatrix.cu:111-219 */
    bmSpMatrix(std::string path, bool transpose, void* stream = nullptr) : h_(nullptr) {
        bmsp::check(bmsp_create_from_mtx(path.c_str(), transpose ? 1 : 0, bmsp::dtype_of<valueType>::value, stream, &h_));
        refresh();
    }
    /* adopting constructor, src/bmSpMatrix.cu:30-43 (device pointers; the arrays are copied, not swapped) */
    bmSpMatrix(int rows, int cols, int blocks, const uint64_t* d_keys, const uint64_t* d_bmps, const uint64_t* d_offsets,
               size_t offsets_len, const valueType* d_values, size_t n_values, bool transposed = false, void* stream = nullptr) : h_(nullptr) {
        bmsp::check(bmsp_create_from_arrays(rows, cols, blocks, (int64_t)n_values, d_keys, d_bmps, d_offsets, (int64_t)offsets_len, d_values,
                                            bmsp::dtype_of<valueType>::value, BMSP_DEVICE, transposed ? 1 : 0, stream, &h_));
        refresh();
    }
#ifdef BMSP_HAVE_THRUST
    /* the reference's own signature (include/bmSpMatrix.h:33-34): adopts the four vectors.  The reference swaps them into the
     * object (src/bmSpMatrix.cu:38-42), leaving the caller's vectors empty; here the arrays are copied into the handle and the
     * caller's vectors are released, which is the same observable state. */
    bmSpMatrix(int rows, int cols, int blocks, thrust::device_vector<uint64_t>& k, thrust::device_vector<uint64_t>& b,
               thrust::device_vector<uint64_t>& o, thrust::device_vector<valueType>& v) : h_(nullptr) {
        bmsp::check(bmsp_create_from_arrays(rows, cols, blocks, (int64_t)v.size(), thrust::raw_pointer_cast(k.data()),
                                            thrust::raw_pointer_cast(b.data()), thrust::raw_pointer_cast(o.data()), (int64_t)o.size(),
                                            thrust::raw_pointer_cast(v.data()), bmsp::dtype_of<valueType>::value, BMSP_DEVICE, 0, nullptr, &h_));
        cudaStreamSynchronize(0);
        thrust::device_vector<uint64_t>().swap(k); thrust::device_vector<uint64_t>().swap(b); thrust::device_vector<uint64_t>().swap(o);
        thrust::device_vector<valueType>().swap(v);
        refresh();
    }
#endif
    /* the CSR entry point the north star adds (CSRMatrix.h role): host or device CSR */
    static bmSpMatrix from_csr(int rows, int cols, int64_t n, const int32_t* row_ptr, const int32_t* col_idx, const float* vals,
                               bool on_device, bool transpose, void* stream = nullptr) {
        bmSpMatrix m;
        bmsp::check(bmsp_create_from_csr(rows, cols, n, row_ptr, col_idx, vals, BMSP_F32, on_device ? BMSP_DEVICE : BMSP_HOST,
                                         transpose ? 1 : 0, bmsp::dtype_of<valueType>::value, stream, &m.h_));
        m.refresh();
        return m;
    }
    bmSpMatrix(const bmSpMatrix&) = delete;
    bmSpMatrix& operator=(const bmSpMatrix&) = delete;
    bmSpMatrix(bmSpMatrix&& o) noexcept : h_(nullptr) { *this = static_cast<bmSpMatrix&&>(o); }
    bmSpMatrix& operator=(bmSpMatrix&& o) noexcept {
        if (this != &o) { if (h_) bmsp_destroy(h_); h_ = o.h_; o.h_ = nullptr; refresh(); o.refresh(); }
        return *this;
    }
    ~bmSpMatrix() { if (h_) bmsp_destroy(h_); }

    /* generate_coo, src/bmSpMatrix.cu:320-363: host COO in block order then bit order */
    void generate_coo() {
        coo_rows.resize(nnz); coo_cols.resize(nnz); coo_vals.resize(nnz);
        if (nnz) bmsp::check(bmsp_to_coo(h_, coo_rows.data(), coo_cols.data(), coo_vals.data()));
    }
    /* compare, src/bmSpMatrix.cu:381-432 -- returns a real verdict: same pattern and max relative error <= tol */
    bool compare(const std::vector<int32_t>& rows, const std::vector<int32_t>& cols, const std::vector<float>& vals, double tol = 1e-3,
                 double* mean_rel_err = nullptr) const {
        int64_t a = 0, b = 0; double mean = 0, mx = 0;
        bmsp::check(bmsp_compare(h_, (int64_t)rows.size(), rows.data(), cols.data(), vals.data(), &a, &b, &mean, &mx));
        if (mean_rel_err) *mean_rel_err = mean;
        return a == 0 && b == 0 && mx <= tol;
    }
    bmsp_matrix_t handle() const { return h_; }
    void adopt(bmsp_matrix_t h) { if (h_) bmsp_destroy(h_); h_ = h; refresh(); }

    std::vector<int32_t> coo_rows, coo_cols;
    std::vector<float> coo_vals;

private:
    void refresh() {
        if (!h_) { keys = bmps = offsets = bmsp::device_view<uint64_t>(); values = bmsp::device_view<valueType>(); num_rows = num_cols = nnz = block_num = 0; return; }
        bmsp_view v;
        bmsp::check(bmsp_get(h_, &v));
        keys = bmsp::device_view<uint64_t>(v.keys, (size_t)v.block_num);
        bmps = bmsp::device_view<uint64_t>(v.bmps, (size_t)v.block_num);
        offsets = bmsp::device_view<uint64_t>(v.offsets, (size_t)v.offsets_len);
        values = bmsp::device_view<valueType>(static_cast<const valueType*>(v.values), (size_t)v.nnz);
        num_rows = v.num_rows; num_cols = v.num_cols; nnz = (int)v.nnz; block_num = (int)v.block_num;
    }
    bmsp_matrix_t h_;
};

/* bmSparse_SpMV<ValueIn,ValueOut>(A, v, u, batched), src/bmSparse_SPMV.cu:191-230.  v has the matrix' value type (as in the
 * reference); u is fp32.  `batched` picked the reference's second kernel (broken as shipped): accepted, ignored. */
template <class ValueIn, class ValueOut>
inline void bmSparse_SpMV(bmSpMatrix<ValueIn>& A, ValueIn* v, ValueOut* u, bool batched = false, void* stream = nullptr) {
    static_assert(sizeof(ValueOut) == 4, "the output vector is fp32");
    (void)batched;
    bmsp::check(bmsp_spmv(A.handle(), v, bmsp::dtype_of<ValueIn>::value, reinterpret_cast<float*>(u), stream));
}
/* fp16 matrix with an fp32 x vector: the intended instantiation (BASELINE config 2) */
template <class ValueIn>
inline void bmSparse_SpMV_f32x(bmSpMatrix<ValueIn>& A, const float* v, float* u, void* stream = nullptr) {
    bmsp::check(bmsp_spmv(A.handle(), v, BMSP_F32, u, stream));
}

/* The driver's copy-in / multiply / copy-out sequence (SPMV.cu:276-285, :299, :308-309) as one pipelined call: v and u are HOST
 * pointers (pinned for full speed); u is complete once `stream` is synchronised. */
template <class ValueIn>
inline void bmSparse_SpMV_host(bmSpMatrix<ValueIn>& A, const float* v_host, float* u_host, void* stream = nullptr) {
    bmsp::check(bmsp_spmv_host(A.handle(), v_host, BMSP_F32, u_host, stream));
}

/* bmSparse_mult<valueIn,valueOut>(A, B, C, mode, VERBOSE, tc_version), src/bmSparse_SPGEMM.cu:827-1223.  B must have been built
 * with transpose = true (SPGEMM.cu:1262).  mode / tc_version are accepted and ignored. */
template <class valueIn, class valueOut>
inline void bmSparse_mult(bmSpMatrix<valueIn>& A, bmSpMatrix<valueIn>& B, bmSpMatrix<valueOut>& C, bool mode = false, bool VERBOSE = false,
                          long tc_version = 5, bmsp_spgemm_info* info = nullptr, void* stream = nullptr) {
    bmsp_spgemm_opts o = {mode ? 1 : 0, (int32_t)tc_version, VERBOSE ? 1 : 0, -1, 0, 0};
    bmsp_matrix_t c = nullptr;
    bmsp::check(bmsp_spgemm(A.handle(), B.handle(), &o, stream, &c, info));
    C.adopt(c);
}

#endif /* BMSPMATRIX_H_ */
