// The reference's Thrust-typed call sites against the shim headers: the adopting constructor (include/bmSpMatrix.h:33-34, used by
// callers that build the four vectors themselves) and the five-argument mmread_bmSparse (include/reader.h:14-15).
#include <cstdio>
#include <cuda_fp16.h>
#include <thrust/device_vector.h>
#include "bmSpMatrix.h"
#include "reader.h"

int main(int argc, char** argv) {
    if (argc < 2) { std::printf("usage: thrust_shim A.mtx\n"); return 2; }
    try {
        uint64_vec keys, bmps, offsets; half_vec values;
        thrust::tuple<int, int, int> dims = mmread_bmSparse(std::string(argv[1]), keys, bmps, offsets, values);
        std::printf("mmread5: %d %d %d blocks %zu values %zu\n", thrust::get<0>(dims), thrust::get<1>(dims), thrust::get<2>(dims), keys.size(), values.size());
        const int blocks = (int)keys.size();
        bmSpMatrix<half> A(thrust::get<0>(dims), thrust::get<1>(dims), blocks, keys, bmps, offsets, values);
        std::printf("adopted: blocks %d nnz %d caller vectors now %zu %zu %zu %zu\n", A.block_num, A.nnz, keys.size(), bmps.size(), offsets.size(), values.size());
        float *v, *u;
        cudaMalloc(&v, sizeof(float) * A.num_cols); cudaMalloc(&u, sizeof(float) * A.num_rows);
        std::vector<float> ones(A.num_cols, 1.0f), y(A.num_rows);
        cudaMemcpy(v, ones.data(), sizeof(float) * ones.size(), cudaMemcpyHostToDevice);
        bmSparse_SpMV_f32x(A, v, u);
        cudaMemcpy(y.data(), u, sizeof(float) * y.size(), cudaMemcpyDeviceToHost);
        double s = 0; for (float t : y) s += t;
        std::printf("SpMV sum: %.1f\n", s);
    } catch (const std::exception& e) { std::printf("error: %s\n", e.what()); return 1; }
    return 0;
}
