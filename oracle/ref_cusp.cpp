// ref_cusp.cpp -- C-ABI shim around the REFERENCE's own host CSR kernels.  TEST/BASELINE ONLY.
//
// Compiles cusp's sequential and OpenMP CSR SpMV/SpGEMM straight from the reference tree
// (paths injected by oracle/Makefile as REF_SEQ_SPGEMM_H etc.; nothing is copied into this
// repo).  Bundled cusp does not build against Thrust 2.8, so the four cusp support headers
// those files include are replaced by the stubs in oracle/refstub/.  The output library
// oracle/_ref/libcusp_ref.so is what bench.py --impl reference and cpu_baseline time
// ("kind": "reference"), and what tests use to validate oracle/bmsp_oracle.c's restatement.
//
// This is what cusp::multiply(A_csr_host, B_csr_host, C) runs for host_memory
// (cusp/cusp/detail/multiply.inl:44-59 -> system/detail/generic/multiply.inl:93-132 ->
//  system/detail/sequential/multiply/csr_spgemm.h:165-197, or system/omp/... under OMP).
#include REF_SEQ_SPGEMM_H
#include REF_SEQ_SPMV_H
#include REF_OMP_SPGEMM_H
#include REF_OMP_SPMV_H
#include <vector>
#include <cstdint>
#include <cstring>

namespace {
struct Csr {
    typedef int index_type;
    typedef float value_type;
    size_t num_rows = 0, num_cols = 0, num_entries = 0;
    std::vector<int> row_offsets, column_indices;
    std::vector<float> values;
    void resize(size_t r, size_t c, size_t n) {
        num_rows = r; num_cols = c; num_entries = n;
        row_offsets.resize(r + 1); column_indices.resize(n); values.resize(n);
    }
};
struct View {   // non-owning array with the vector interface the kernels use
    const int* p; size_t n; typedef int value_type;
    const int& operator[](size_t i) const { return p[i]; }
};
struct ViewF {
    const float* p; size_t n; typedef float value_type;
    const float& operator[](size_t i) const { return p[i]; }
};
struct CsrView {
    typedef int index_type; typedef float value_type;
    size_t num_rows, num_cols, num_entries; View row_offsets, column_indices; ViewF values;
};
struct VecOut { float* p; typedef float value_type; float& operator[](size_t i) { return p[i]; } const float& operator[](size_t i) const { return p[i]; } };
struct zero_init { float operator()(const float&) const { return 0.0f; } };
}

extern "C" {

// y = A x through sequential/multiply/csr_spmv.h:42-74
void ref_csr_spmv_seq(int rows, int cols, const int* rp, const int* ci, const float* v, const float* x, float* y) {
    CsrView A{(size_t)rows, (size_t)cols, (size_t)rp[rows], {rp, (size_t)rows + 1}, {ci, (size_t)rp[rows]}, {v, (size_t)rp[rows]}};
    ViewF X{x, (size_t)cols}; VecOut Y{y};
    thrust::cpp::tag exec;
    cusp::system::detail::sequential::multiply(exec, A, X, Y, zero_init(), thrust::multiplies<float>(), thrust::plus<float>(),
                                               cusp::csr_format(), cusp::array1d_format(), cusp::array1d_format());
}

// y = A x through omp/detail/multiply/csr_spmv.h:43-86 (#pragma omp parallel for over rows)
void ref_csr_spmv_omp(int rows, int cols, const int* rp, const int* ci, const float* v, const float* x, float* y) {
    CsrView A{(size_t)rows, (size_t)cols, (size_t)rp[rows], {rp, (size_t)rows + 1}, {ci, (size_t)rp[rows]}, {v, (size_t)rp[rows]}};
    ViewF X{x, (size_t)cols}; VecOut Y{y};
    thrust::omp::tag exec;
    cusp::system::omp::detail::multiply(exec, A, X, Y, zero_init(), thrust::multiplies<float>(), thrust::plus<float>(),
                                        cusp::csr_format(), cusp::array1d_format(), cusp::array1d_format());
}

// C = A B; omp == 0: sequential/multiply/csr_spgemm.h:165-197, omp != 0: omp/detail/multiply/csr_spgemm.h:166-198
void* ref_csr_spgemm(int a_rows, int a_cols, const int* a_rp, const int* a_ci, const float* a_v,
                     int b_rows, int b_cols, const int* b_rp, const int* b_ci, const float* b_v, int omp) {
    CsrView A{(size_t)a_rows, (size_t)a_cols, (size_t)a_rp[a_rows], {a_rp, (size_t)a_rows + 1}, {a_ci, (size_t)a_rp[a_rows]}, {a_v, (size_t)a_rp[a_rows]}};
    CsrView B{(size_t)b_rows, (size_t)b_cols, (size_t)b_rp[b_rows], {b_rp, (size_t)b_rows + 1}, {b_ci, (size_t)b_rp[b_rows]}, {b_v, (size_t)b_rp[b_rows]}};
    Csr* C = new Csr();
    if (omp) {
        thrust::omp::tag exec;
        cusp::system::omp::detail::multiply(exec, A, B, *C, zero_init(), thrust::multiplies<float>(), thrust::plus<float>(),
                                            cusp::csr_format(), cusp::csr_format(), cusp::csr_format());
    } else {
        thrust::cpp::tag exec;
        cusp::system::detail::sequential::multiply(exec, A, B, *C, zero_init(), thrust::multiplies<float>(), thrust::plus<float>(),
                                                   cusp::csr_format(), cusp::csr_format(), cusp::csr_format());
    }
    return C;
}
long long ref_result_nnz(void* h) { return (long long)static_cast<Csr*>(h)->num_entries; }
void ref_result_copy(void* h, int* rp, int* ci, float* v) {
    Csr* C = static_cast<Csr*>(h);
    std::memcpy(rp, C->row_offsets.data(), sizeof(int) * (C->num_rows + 1));
    std::memcpy(ci, C->column_indices.data(), sizeof(int) * C->num_entries);
    std::memcpy(v, C->values.data(), sizeof(float) * C->num_entries);
}
void ref_result_free(void* h) { delete static_cast<Csr*>(h); }
}
