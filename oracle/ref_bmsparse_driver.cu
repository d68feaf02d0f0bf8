// ref_bmsparse_driver.cu -- runs the REFERENCE's own bmSparse CUDA operators on the B200.  TEST/BASELINE ONLY.
//
// The reference sources are compiled from where they lie under /root/reference (paths injected by
// oracle/Makefile; nothing is copied; src/bmSpMatrix.cu is compiled as its own object) with main() renamed; this file adds a driver that feeds them matrices
// through the reference's own adopting constructor (src/bmSpMatrix.cu:30-43) from raw array files written by
// the tests, times exactly the region the reference's main() times (SPGEMM.cu:1274-1280, SPMV.cu:297-304) and
// dumps the result arrays so that tests can compare them bit for bit with the oracle and the B200 kernels.
//
//   ref_spgemm <A.bin> <Bt.bin> <out.bin> [tc_version=5] [mode=0] [reps=1]
//   ref_spmv   <A.bin> <out.bin> [reps=1]           (fp32 matrix, x = ones, as the shipped main does)
//
// .bin layout: int64 rows, cols, nblk, nnz, value_bytes; keys[nblk] bmps[nblk] offsets[nblk] u64; values[nnz].
#define main reference_main
#ifdef REF_BUILD_SPGEMM
#include REF_SPGEMM_CU
#else
#include REF_SPMV_CU
#endif
#undef main
#include <cstdio>
#include <vector>

struct HostMat {
    long long rows, cols, nblk, nnz, vbytes;
    std::vector<uint64_t> keys, bmps, offsets;
    std::vector<char> values;
};
static bool read_mat(const char* path, HostMat& m) {
    FILE* f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); return false; }
    long long h[5];
    if (fread(h, 8, 5, f) != 5) return false;
    m.rows = h[0]; m.cols = h[1]; m.nblk = h[2]; m.nnz = h[3]; m.vbytes = h[4];
    m.keys.resize(m.nblk); m.bmps.resize(m.nblk); m.offsets.resize(m.nblk); m.values.resize(m.nnz * m.vbytes);
    bool ok = fread(m.keys.data(), 8, m.nblk, f) == (size_t)m.nblk && fread(m.bmps.data(), 8, m.nblk, f) == (size_t)m.nblk &&
              fread(m.offsets.data(), 8, m.nblk, f) == (size_t)m.nblk && fread(m.values.data(), m.vbytes, m.nnz, f) == (size_t)m.nnz;
    fclose(f);
    return ok;
}
template <typename T>
static bmSpMatrix<T>* adopt(const HostMat& m) {
    thrust::device_vector<uint64_t> k(m.keys.begin(), m.keys.end()), b(m.bmps.begin(), m.bmps.end()), o(m.offsets.begin(), m.offsets.end());
    thrust::device_vector<T> v(m.nnz);
    cudaMemcpy(thrust::raw_pointer_cast(v.data()), m.values.data(), m.nnz * sizeof(T), cudaMemcpyHostToDevice);
    return new bmSpMatrix<T>((int)m.rows, (int)m.cols, (int)m.nblk, k, b, o, v);
}

int main(int argc, char** argv) {
    cudaFree(0);
#ifdef REF_BUILD_SPGEMM
    if (argc < 4) { fprintf(stderr, "usage: ref_spgemm A.bin Bt.bin out.bin [tc_version] [mode] [reps]\n"); return 2; }
    long tc = argc > 4 ? atol(argv[4]) : 5; long mode = argc > 5 ? atol(argv[5]) : 0; int reps = argc > 6 ? atoi(argv[6]) : 1;
    HostMat ha, hb;
    if (!read_mat(argv[1], ha) || !read_mat(argv[2], hb)) return 3;
    bmSpMatrix<half>* A = adopt<half>(ha);
    bmSpMatrix<half>* B = adopt<half>(hb);
    double best = 1e30;
    bmSpMatrix<float>* C = nullptr;
    for (int r = 0; r < reps; r++) {
        delete C; C = new bmSpMatrix<float>();
        cudaDeviceSynchronize();
        auto t0 = std::chrono::steady_clock::now();
        bmSparse_mult<half, float>(*A, *B, *C, (bool)mode, false, tc);
        cudaDeviceSynchronize();
        double us = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
        if (us < best) best = us;
    }
    cudaError_t e = cudaGetLastError();
    printf("REF_SPGEMM_US %.1f C_blocks %zu C_nnz %d cuda_status %d\n", best, C->keys.size(), C->nnz, (int)e);
    thrust::host_vector<uint64_t> k = C->keys, b = C->bmps, o = C->offsets;
    thrust::host_vector<float> v = C->values;
    FILE* f = fopen(argv[3], "wb");
    long long h[5] = {C->num_rows, C->num_cols, (long long)k.size(), (long long)v.size(), (long long)o.size()};
    fwrite(h, 8, 5, f);
    fwrite(k.data(), 8, k.size(), f); fwrite(b.data(), 8, k.size(), f); fwrite(o.data(), 8, o.size(), f); fwrite(v.data(), 4, v.size(), f);
    fclose(f);
#else
    if (argc < 3) { fprintf(stderr, "usage: ref_spmv A.bin out.bin [reps]\n"); return 2; }
    int reps = argc > 3 ? atoi(argv[3]) : 1;
    HostMat ha;
    if (!read_mat(argv[1], ha)) return 3;
    bmSpMatrix<float>* A = adopt<float>(ha);
    thrust::device_vector<float> x(ha.cols, 1.0f), y(ha.rows + 8, 0.0f);
    double best = 1e30;
    for (int r = 0; r < reps; r++) {
        cudaDeviceSynchronize();
        auto t0 = std::chrono::steady_clock::now();
        bmSparse_SpMV<float, float>(*A, thrust::raw_pointer_cast(x.data()), thrust::raw_pointer_cast(y.data()), false);
        cudaDeviceSynchronize();
        double us = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
        if (us < best) best = us;
    }
    printf("REF_SPMV_US %.1f cuda_status %d\n", best, (int)cudaGetLastError());
    thrust::host_vector<float> hy = y;
    FILE* f = fopen(argv[2], "wb");
    long long n = ha.rows;
    fwrite(&n, 8, 1, f); fwrite(hy.data(), 4, n, f); fclose(f);
#endif
    return 0;
}
