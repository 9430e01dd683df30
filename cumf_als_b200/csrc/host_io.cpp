// host_io.cpp -- the CLI's raw-binary loaders (drop-in for host_utilities.cpp:19-98).
//
// File format: headerless little-endian arrays, int32 indices and
// float32 values.  The C-ABI functions report failures; the C++-linkage functions
// keep the reference's names and void signatures (host_utilities.h:31-40) so the
// reference's unmodified main.cpp links against this library.
#include <cstdio>
#include <vector>
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

// ---- sharded loading: one rank reads only ITS rows of the CLI's .bin files (hugewiki.cu:2332-2340 keeps one file set per GPU
// batch; here the reference's own single file set is sliced with seeks, so no process ever holds the whole matrix) ----------
extern "C" int cumf_bin_shard_extent(const char* indptrFile, int rows, int row_begin, int row_end, long long* first, long long* count) {
    if (!indptrFile || !first || !count || row_begin < 0 || row_begin > row_end || row_end > rows) return -1;
    FILE* f = fopen(indptrFile, "rb");
    if (!f) return -1;
    int a = 0, b = 0;
    int ok = fseek(f, (long)sizeof(int) * row_begin, SEEK_SET) == 0 && fread(&a, sizeof(int), 1, f) == 1 &&
             fseek(f, (long)sizeof(int) * row_end, SEEK_SET) == 0 && fread(&b, sizeof(int), 1, f) == 1;
    fclose(f);
    if (!ok || b < a) return -1;
    *first = a;
    *count = (long long)b - a;
    return 0;
}
// `count` items of `elem_size` bytes starting at item `first` of a headerless .bin file
extern "C" int cumf_load_bin_slice(const char* file, int elem_size, long long first, long long count, void* dst) {
    if (!file || !dst || elem_size <= 0 || first < 0 || count < 0) return -1;
    FILE* f = fopen(file, "rb");
    if (!f) return -1;
    int ok = fseek(f, (long)(first * elem_size), SEEK_SET) == 0 && fread(dst, (size_t)elem_size, (size_t)count, f) == (size_t)count;
    fclose(f);
    return ok ? 0 : -1;
}
// rows [row_begin, row_end) of a CSR (or columns of a CSC) file set: ptr_out gets row_end - row_begin + 1 pointers rebased to
// 0 (int64, what cumf_als_create_device / cumf_plan_create64 take), idx_out / val_out the slice (sized by cumf_bin_shard_extent)
extern "C" int cumf_load_csr_shard_bin(const char* dataFile, const char* indptrFile, const char* indicesFile, int rows, int row_begin,
                                       int row_end, long long* ptr_out, int* idx_out, float* val_out) {
    long long first = 0, count = 0;
    if (!ptr_out || cumf_bin_shard_extent(indptrFile, rows, row_begin, row_end, &first, &count)) return -1;
    std::vector<int> ptr((size_t)(row_end - row_begin) + 1);
    if (cumf_load_bin_slice(indptrFile, (int)sizeof(int), row_begin, (long long)ptr.size(), ptr.data())) return -1;
    for (size_t i = 0; i < ptr.size(); ++i) ptr_out[i] = (long long)ptr[i] - first;
    if (count == 0) return 0;
    if (!idx_out || !val_out) return -1;
    if (cumf_load_bin_slice(indicesFile, (int)sizeof(int), first, count, idx_out)) return -1;
    if (cumf_load_bin_slice(dataFile, (int)sizeof(float), first, count, val_out)) return -1;
    return 0;
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
