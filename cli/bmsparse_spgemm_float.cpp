// bmsparse_spgemm_float -- command-line driver with the reference's calling convention and output lines, so that
// spgemm_run_batch.sh runs unchanged against libbmsparse_b200.so:
//     bmsparse_spgemm_float <MatrixFolder> <A_Matrix> <B_Matrix> [segmented tc_version verbose]
// (reference main: src/bmSparse_SPGEMM.cu:1226-1288; it reads <folder>/<name>.mtx, A plain and B in transposed-operand form,
// times bmSparse_mult with a host clock and prints "bmSparse execution", "C blocks", "C nnz").  Differences, all deliberate:
// arguments are read from the positions the batch script passes them in (the reference indexes one past);
// `segmented` / `tc_version` are accepted and ignored; verbose = 1 prints the per-phase device times under the reference's
// T_n labels (symbolic = T_1..T_6,T_9, numeric = T_7); errors are reported instead of exit() inside the library.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>
#include <cuda_runtime.h>
#include "bmSpMatrix.h"

int main(int argc, char** argv) {
    if (argc < 4) {
        std::cout << "./main MatrixFolder A_Matrix B_Matrix [segmented tc_version verbose]" << std::endl;
        return 1;
    }
    const long segmented = argc > 4 ? strtol(argv[4], NULL, 10) : 0;
    const long tc_version = argc > 5 ? strtol(argv[5], NULL, 10) : 5;
    const bool verbose = argc > 6 && *argv[6] == '1';
    const std::string A_path = std::string(argv[1]) + "/" + argv[2], B_path = std::string(argv[1]) + "/" + argv[3];
    std::cout << "A matrix: " << A_path << std::endl << "B matrix: " << B_path << std::endl;
    using clk = std::chrono::steady_clock;
    auto us = [](clk::time_point a, clk::time_point b) { return (long long)std::chrono::duration_cast<std::chrono::microseconds>(b - a).count(); };
    try {
        cudaFree(0);
        auto t0 = clk::now();
        bmSpMatrix<bmsp::half_t> A(A_path + ".mtx", false), B(B_path + ".mtx", true);
        cudaDeviceSynchronize();
        std::cout << "Parsing mtx files / Loading matrices from disk BMSP: " << us(t0, clk::now()) << " μs" << std::endl;
        bmSpMatrix<float> C;
        bmsp_spgemm_info info;
        {   // first call pays one-time work the reference also leaves outside its numbers (context, module load, B's packed metadata)
            bmSpMatrix<float> warm;
            bmSparse_mult<bmsp::half_t, float>(A, B, warm, segmented != 0, false, tc_version);
            cudaDeviceSynchronize();
        }
        t0 = clk::now();
        bmSparse_mult<bmsp::half_t, float>(A, B, C, segmented != 0, verbose, tc_version, &info);
        cudaDeviceSynchronize();
        const long long t = us(t0, clk::now());
        if (verbose) {
            std::cout << "T_1-T_6,T_9 (symbolic): " << (long long)(info.symbolic_ms * 1e3) << " μs" << std::endl;
            std::cout << "T_7 (numeric): " << (long long)(info.numeric_ms * 1e3) << " μs" << std::endl;
            std::cout << "Task list size: " << info.surviving_pairs << " (of " << info.candidate_pairs << " candidate pairs)" << std::endl;
        }
        std::cout << "bmSparse execution: " << t << " μs" << std::endl;
        std::cout << "C blocks: " << C.keys.size() << std::endl;
        std::cout << "C nnz: " << C.nnz << std::endl;
    } catch (const std::exception& e) {
        std::cerr << "error: " << e.what() << std::endl;
        return 2;
    }
    return 0;
}
