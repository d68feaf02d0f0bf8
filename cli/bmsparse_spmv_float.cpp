// bmsparse_spmv_float -- command-line driver with the reference's calling convention and output lines, so that
// spmv_run_batch.sh runs unchanged against libbmsparse_b200.so:
//     bmsparse_spmv_float <MatrixFolder> <A_Matrix> [<A_Matrix>] [batched]
// (reference main: src/bmSparse_SPMV.cu:232-312: x = ones, y = A x, prints "bmSparse SpMV execution").  Differences, all
// deliberate: the fp16 matrix the reference builds and then ignores IS the one multiplied here (fp16 values, fp32 x / y /
// accumulate -- BASELINE config 2; pass BMSP_CLI_FP32=1 for the fp32 matrix the reference actually ran); the second path
// gets its ".mtx" suffix; `batched` is accepted and ignored; the checksum of y is printed so runs can be compared.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "bmSpMatrix.h"

template <class V>
static int run(const std::string& path, int reps) {
    using clk = std::chrono::steady_clock;
    auto us = [](clk::time_point a, clk::time_point b) { return (long long)std::chrono::duration_cast<std::chrono::microseconds>(b - a).count(); };
    auto t0 = clk::now();
    bmSpMatrix<V> A(path + ".mtx", false);
    cudaDeviceSynchronize();
    std::cout << "Parsing mtx files / Loading matrices from disk BMSP: " << us(t0, clk::now()) << " μs" << std::endl;
    std::cout << "Running SpMV \n";
    float *v = nullptr, *u = nullptr;
    cudaMalloc((void**)&v, sizeof(float) * (size_t)A.num_cols);
    cudaMalloc((void**)&u, sizeof(float) * (size_t)A.num_rows);
    std::vector<float> ones((size_t)A.num_cols, 1.0f), y((size_t)A.num_rows);
    t0 = clk::now();
    cudaMemcpy(v, ones.data(), sizeof(float) * ones.size(), cudaMemcpyHostToDevice);
    std::cout << "Parsing mtx files / Loading matrix and vectors: " << us(t0, clk::now()) << " μs" << std::endl;
    bmSparse_SpMV_f32x(A, v, u);        // first call builds the SpMV plan (the reference rebuilds its row pointers on every call)
    cudaDeviceSynchronize();
    t0 = clk::now();
    for (int i = 0; i < reps; i++) bmSparse_SpMV_f32x(A, v, u);
    cudaDeviceSynchronize();
    std::cout << "bmSparse SpMV execution: " << us(t0, clk::now()) / reps << " μs" << std::endl;
    cudaMemcpy(y.data(), u, sizeof(float) * y.size(), cudaMemcpyDeviceToHost);
    double s = 0;
    for (float t : y) s += t;
    std::cout << "y checksum: " << s << std::endl;
    cudaFree(v); cudaFree(u);
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 3) {
        std::cout << "./main MatrixFolder A_Matrix [A_Matrix] [batched]" << std::endl;
        return 1;
    }
    const std::string A_path = std::string(argv[1]) + "/" + argv[2];
    std::cout << "A matrix: " << A_path << std::endl;
    const char* e = getenv("BMSP_CLI_REPS");
    const int reps = e ? std::max(1, atoi(e)) : 1;
    const char* f = getenv("BMSP_CLI_FP32");
    try {
        cudaFree(0);
        return (f && *f == '1') ? run<float>(A_path, reps) : run<bmsp::half_t>(A_path, reps);
    } catch (const std::exception& ex) {
        std::cerr << "error: " << ex.what() << std::endl;
        return 2;
    }
}
