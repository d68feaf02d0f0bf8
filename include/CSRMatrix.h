/*
 * CSRMatrix.h -- shim for the reference's include/CSRMatrix.h (class CSRMatrix: constructor from a file, constructor from
 * a host CSR, multiply()).  The reference only declares the class (no definition anywhere in its tree); its role -- CSR
 * ingest and A*B -- is implemented here on the B200 path: both operands are converted to bmSparse on the device (the right
 * one in transposed-operand form), multiplied with bmsp_spgemm and brought back as host CSR.  No CPU multiply is involved.
 */
#ifndef CUSPARSE_H_
#define CUSPARSE_H_
#include <algorithm>
#include <numeric>
#include "bmSpMatrix.h"

class CSRMatrix {
public:
    int num_rows, num_cols;
    std::vector<int32_t> row_offsets, column_indices;
    std::vector<float> values;

    CSRMatrix() : num_rows(0), num_cols(0) {}
    /* CSRMatrix(std::string): MatrixMarket file */
    explicit CSRMatrix(const std::string& path) {
        bmSpMatrix<float> m(path, false);
        from_bm(m);
    }
    /* CSRMatrix(cusp::csr_matrix<float,float,host>*): any host CSR triple (columns ascending inside each row) */
    CSRMatrix(int rows, int cols, const int32_t* rp, const int32_t* ci, const float* v) : num_rows(rows), num_cols(cols),
        row_offsets(rp, rp + rows + 1), column_indices(ci, ci + rp[rows]), values(v, v + rp[rows]) {}

    /* C = this * other: fp16 operands, fp32 accumulate (bmSparse_mult<half,float>) */
    CSRMatrix multiply(const CSRMatrix& other) const {
        bmSpMatrix<bmsp::half_t> A = bmSpMatrix<bmsp::half_t>::from_csr(num_rows, num_cols, (int64_t)values.size(), row_offsets.data(),
                                                                        column_indices.data(), values.data(), false, false);
        bmSpMatrix<bmsp::half_t> Bt = bmSpMatrix<bmsp::half_t>::from_csr(other.num_rows, other.num_cols, (int64_t)other.values.size(),
                                                                         other.row_offsets.data(), other.column_indices.data(),
                                                                         other.values.data(), false, true);
        bmSpMatrix<float> C;
        bmSparse_mult<bmsp::half_t, float>(A, Bt, C);
        CSRMatrix out;
        out.from_bm(C);
        return out;
    }

private:
    template <class T>
    void from_bm(bmSpMatrix<T>& m) {
        num_rows = m.num_rows; num_cols = m.num_cols;
        m.generate_coo();
        const size_t n = m.coo_rows.size();
        std::vector<size_t> perm(n);
        std::iota(perm.begin(), perm.end(), (size_t)0);
        std::sort(perm.begin(), perm.end(), [&](size_t a, size_t b) {
            return m.coo_rows[a] != m.coo_rows[b] ? m.coo_rows[a] < m.coo_rows[b] : m.coo_cols[a] < m.coo_cols[b]; });
        row_offsets.assign(num_rows + 1, 0); column_indices.resize(n); values.resize(n);
        for (size_t i = 0; i < n; i++) {
            row_offsets[m.coo_rows[perm[i]] + 1]++; column_indices[i] = m.coo_cols[perm[i]]; values[i] = m.coo_vals[perm[i]];
        }
        for (int r = 0; r < num_rows; r++) row_offsets[r + 1] += row_offsets[r];
    }
};
#endif /* CUSPARSE_H_ */
