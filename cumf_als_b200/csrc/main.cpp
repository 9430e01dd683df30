// main.cpp -- the cuMF ALS command line on top of libcumf_als_b200.so.
//
// Same contract as the reference CLI (main.cpp:19-172): nine positional
// arguments, the ten .bin files of DATA_DIR, theta0 = 0.2*rand()/RAND_MAX after
// srand(0), X0 = 0, ITERS = 10 on device 0, and the stdout lines the reference's
// log scrapers read (print-test-result.sh:8-12).  The reference's own main.cpp
// also links unmodified against the library; this file
// exists so the repository builds a CLI without the reference tree.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <string>
#include <sys/time.h>

#include "../../include/cumf_als.h"

static const int kDevice = 0;   // DEVICEID, main.cpp:16
static const int kIters = 10;   // ITERS,    main.cpp:17

static double now_seconds() {
    struct timeval tv;
    gettimeofday(&tv, nullptr);
    return tv.tv_sec + 1e-6 * tv.tv_usec;
}

template <typename T>
static T* pinned(size_t count) {
    T* p = nullptr;
    if (cudaMallocHost((void**)&p, count * sizeof(T)) != cudaSuccess) {
        fprintf(stderr, "cudaMallocHost of %zu bytes failed\n", count * sizeof(T));
        exit(EXIT_FAILURE);
    }
    return p;
}

int main(int argc, char** argv) {
    if (argc != 10) {
        printf("Usage: give M, N, F, NNZ, NNZ_TEST, lambda, X_BATCH, THETA_BATCH and DATA_DIR.\n");
        printf("E.g., for netflix data set, use: \n");
        printf("./main 17770 480189 100 99072112 1408395 0.048 1 3 ./data/netflix/ \n");
        return 0;   // the reference also exits 0 on usage errors (main.cpp:29)
    }
    const int f = atoi(argv[3]);
    if (f % 10 != 0) {
        printf("F has to be a multiple of %d \n", 10);
        return 0;
    }
    const int m = atoi(argv[1]);
    const int n = atoi(argv[2]);
    const long nnz = atoi(argv[4]);        // parsed with atoi like main.cpp:39-40
    const long nnz_test = atoi(argv[5]);
    const float lambda = (float)atof(argv[6]);
    const int x_batch = atoi(argv[7]);
    const int theta_batch = atoi(argv[8]);
    const std::string dir(argv[9]);
    printf("M = %d, N = %d, F = %d, NNZ = %ld, NNZ_TEST = %ld, lambda = %f\nX_BATCH = %d, THETA_BATCH = %d\nDATA_DIR = %s \n",
           m, n, f, nnz, nnz_test, lambda, x_batch, theta_batch, dir.c_str());

    cudaSetDevice(kDevice);
    int* csr_ptr = pinned<int>((size_t)m + 1);
    int* csr_col = pinned<int>((size_t)nnz);
    float* csr_val = pinned<float>((size_t)nnz);
    float* csc_val = pinned<float>((size_t)nnz);
    int* csc_row = pinned<int>((size_t)nnz);
    int* csc_ptr = pinned<int>((size_t)n + 1);
    int* coo_row = pinned<int>((size_t)nnz);
    float* theta = pinned<float>((size_t)n * f);
    float* x = pinned<float>((size_t)m * f);

    srand(0);
    for (long k = 0; k < (long)n * f; ++k) theta[k] = 0.2 * ((float)rand() / (float)RAND_MAX);
    for (long k = 0; k < (long)m * f; ++k) x[k] = 0;   // CG warm-starts from X (main.cpp:76-78)

    printf("*******start loading training and testing sets to host.\n");
    int* test_row = (int*)malloc(sizeof(int) * (size_t)nnz_test);
    int* test_col = (int*)malloc(sizeof(int) * (size_t)nnz_test);
    float* test_val = (float*)malloc(sizeof(float) * (size_t)nnz_test);
    int bad = 0;
    bad |= cumf_load_coo_bin((dir + "/R_test_coo.data.bin").c_str(), (dir + "/R_test_coo.row.bin").c_str(),
                             (dir + "/R_test_coo.col.bin").c_str(), test_val, test_row, test_col, nnz_test);
    bad |= cumf_load_csr_bin((dir + "/R_train_csr.data.bin").c_str(), (dir + "/R_train_csr.indptr.bin").c_str(),
                             (dir + "/R_train_csr.indices.bin").c_str(), csr_val, csr_ptr, csr_col, m, nnz);
    bad |= cumf_load_csc_bin((dir + "/R_train_csc.data.bin").c_str(), (dir + "/R_train_csc.indices.bin").c_str(),
                             (dir + "/R_train_csc.indptr.bin").c_str(), csc_val, csc_row, csc_ptr, n, nnz);
    bad |= cumf_load_coo_row_bin((dir + "/R_train_coo.row.bin").c_str(), coo_row, nnz);
    if (bad) printf("Unable to open file!");

    const double t0 = now_seconds();
    cumf_doALS(csr_ptr, csr_col, csr_val, csc_row, csc_ptr, csc_val, coo_row, theta, x, test_row, test_col, test_val, m,
               n, f, nnz, nnz_test, lambda, kIters, x_batch, theta_batch, kDevice);
    printf("\ndoALS takes seconds: %.3f for F = %d\n", now_seconds() - t0, f);

    const char* dump = getenv("CUMF_DUMP_FACTORS");   // optional: write XT / thetaT for offline diffing
    if (dump && *dump) {
        FILE* fx = fopen((std::string(dump) + "/XT.bin").c_str(), "wb");
        FILE* ft = fopen((std::string(dump) + "/thetaT.bin").c_str(), "wb");
        if (fx && ft) {
            fwrite(x, sizeof(float), (size_t)m * f, fx);
            fwrite(theta, sizeof(float), (size_t)n * f, ft);
        }
        if (fx) fclose(fx);
        if (ft) fclose(ft);
    }
    cudaFreeHost(csr_ptr); cudaFreeHost(csr_col); cudaFreeHost(csr_val); cudaFreeHost(csc_val);
    cudaFreeHost(csc_row); cudaFreeHost(csc_ptr); cudaFreeHost(coo_row); cudaFreeHost(x); cudaFreeHost(theta);
    free(test_row); free(test_col); free(test_val);
    cudaDeviceReset();
    printf("\nALS Done.\n");
    return 0;
}
