// host_io.cpp -- the CLI's raw-binary loaders (drop-in for host_utilities.cpp:19-98).
//
// File format: headerless little-endian arrays, int32 indices and
// float32 values.  The C-ABI functions report failures; the C++-linkage functions
// keep the reference's names and void signatures (host_utilities.h:31-40) so the
// reference's unmodified main.cpp links against this library.
#include <cstdio>
#include <cstdlib>

#include "../../include/cumf_als.h"

namespace {
// read `count` 4-byte items; 0 on success
int read_items(const char* path, void* dst, size_t count) {
    FILE* fp = fopen(path, "rb");
    if (!fp) return -1;
    const size_t got = fread(dst, 4, count, fp);
    fclose(fp);
    return got == count ? 0 : -1;
}
}  // namespace

extern "C" int cumf_load_csr_bin(const char* dataFile, const char* rowFile, const char* colFile, float* data,
                                 int* row, int* col, int m, long nnz) {
    // indptr has m+1 entries, indices/data nnz (host_utilities.cpp:33-35)
    int rc = read_items(rowFile, row, (size_t)m + 1);
    rc |= read_items(colFile, col, (size_t)nnz);
    rc |= read_items(dataFile, data, (size_t)nnz);
    return rc ? -1 : 0;
}
extern "C" int cumf_load_csc_bin(const char* dataFile, const char* rowFile, const char* colFile, float* data,
                                 int* row, int* col, int n, long nnz) {
    // row ids nnz, column pointers n+1 (host_utilities.cpp:57-59)
    int rc = read_items(rowFile, row, (size_t)nnz);
    rc |= read_items(colFile, col, (size_t)n + 1);
    rc |= read_items(dataFile, data, (size_t)nnz);
    return rc ? -1 : 0;
}
extern "C" int cumf_load_coo_row_bin(const char* rowFile, int* row, long nnz) {
    return read_items(rowFile, row, (size_t)nnz) ? -1 : 0;   // host_utilities.cpp:71
}
extern "C" int cumf_load_coo_bin(const char* dataFile, const char* rowFile, const char* colFile, float* data,
                                 int* row, int* col, long nnz) {
    int rc = read_items(rowFile, row, (size_t)nnz);            // host_utilities.cpp:90-92
    rc |= read_items(colFile, col, (size_t)nnz);
    rc |= read_items(dataFile, data, (size_t)nnz);
    return rc ? -1 : 0;
}

// Factor initialisation loops of the reference's two front ends, on glibc rand() like they are.
// `scale` is a double like the literals 0.2 (main.cpp:75) and 0.1 (als_tf.cc:121): the product is formed in double and
// rounded once, so the values are bit-identical to the front ends' own loops.
extern "C" void cumf_init_factors(float* thetaTHost, float* XTHost, int m, int n, int f, double scale, long seed) {
    if (seed >= 0) srand((unsigned)seed);                                        // main.cpp:73; als_tf.cc never seeds
    if (thetaTHost)
        for (long k = 0; k < (long)n * f; ++k) thetaTHost[k] = (float)(scale * (double)((float)rand() / (float)RAND_MAX));   // main.cpp:75, als_tf.cc:121
    if (XTHost)
        for (long k = 0; k < (long)m * f; ++k) XTHost[k] = 0.f;                  // CG warm-starts from X: main.cpp:78, als_tf.cc:124
}

// Reference-named symbols (C++ linkage, same mangled names as host_utilities.cpp).
void loadCSRSparseMatrixBin(const char* dataFile, const char* rowFile, const char* colFile, float* data, int* row,
                            int* col, const int m, const long nnz) {
    if (cumf_load_csr_bin(dataFile, rowFile, colFile, data, row, col, m, nnz)) printf("Unable to open file!");
}
void loadCSCSparseMatrixBin(const char* dataFile, const char* rowFile, const char* colFile, float* data, int* row,
                            int* col, const int n, const long nnz) {
    if (cumf_load_csc_bin(dataFile, rowFile, colFile, data, row, col, n, nnz)) printf("Unable to open file!");
}
void loadCooSparseMatrixRowPtrBin(const char* rowFile, int* row, const long nnz) {
    if (cumf_load_coo_row_bin(rowFile, row, nnz)) printf("Unable to open file!");
}
void loadCooSparseMatrixBin(const char* dataFile, const char* rowFile, const char* colFile, float* data, int* row,
                            int* col, const long nnz) {
    if (cumf_load_coo_bin(dataFile, rowFile, colFile, data, row, col, nnz)) printf("Unable to open file!");
}
