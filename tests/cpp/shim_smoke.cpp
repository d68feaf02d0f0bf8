// Reads like the reference's mains (src/bmSparse_SPGEMM.cu:1226-1288, src/bmSparse_SPMV.cu:232-312) but through the shim
// headers: build A and B from MatrixMarket files, C = A*B, y = A*1, print the reference's summary lines.
#include <cstdio>
#include <cuda_runtime.h>
#include "bmSpMatrix.h"
#include "reader.h"
#include "CSRMatrix.h"

int main(int argc, char** argv) {
    if (argc < 3) { std::printf("usage: shim_smoke A.mtx B.mtx\n"); return 2; }
    try {
        bmSpMatrix<bmsp::half_t> A(argv[1], false), B(argv[2], true);
        bmSpMatrix<float> C;
        bmSparse_mult<bmsp::half_t, float>(A, B, C, false, false, 5);
        cudaDeviceSynchronize();
        std::printf("C blocks: %zu\nC nnz: %d\n", C.keys.size(), C.nnz);
        bmSpMatrix<float> A32(argv[1], false);
        float *v, *u;
        cudaMalloc(&v, sizeof(float) * A32.num_cols); cudaMalloc(&u, sizeof(float) * A32.num_rows);
        std::vector<float> ones(A32.num_cols, 1.0f), y(A32.num_rows);
        cudaMemcpy(v, ones.data(), sizeof(float) * ones.size(), cudaMemcpyHostToDevice);
        bmSparse_SpMV<float, float>(A32, v, u, false);
        cudaMemcpy(y.data(), u, sizeof(float) * y.size(), cudaMemcpyDeviceToHost);
        double s = 0; for (float t : y) s += t;
        std::printf("SpMV sum: %.1f\n", s);
        bmSpMatrix<bmsp::half_t> R; auto dims = mmread_bmSparse(argv[1], R);
        std::printf("mmread: %d %d %d blocks %d\n", std::get<0>(dims), std::get<1>(dims), std::get<2>(dims), R.block_num);
        CSRMatrix a(argv[1]), b(argv[2]);
        CSRMatrix c = a.multiply(b);
        std::printf("CSR C nnz: %zu\n", c.values.size());
    } catch (const std::exception& e) { std::printf("error: %s\n", e.what()); return 1; }
    return 0;
}
