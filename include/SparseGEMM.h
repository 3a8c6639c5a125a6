// SparseGEMM.h -- C++ header: drop-in for the reference's SparseGEMM.h entry points on the hot path.
//
//   class SparseFormat(int* matrix, int K, int N)   reference SparseGEMM.h:13-40  (int matrix, >=1 / <=-1)
//   sparseGEMM<T>(X, csp, csn, rip, rin, b, Y, M, N, K)          reference SparseGEMM.h:104-119
//   sparseGEMM_PReLU<T>(X, csp, csn, rip, rin, b, Y, M, N, K, a) reference SparseGEMM.h:151-168
//
// Same names, argument order and meaning.  Work runs on the GPU through libtsgemm_b200.so's C-ABI
// (tsg_sparse_gemm_f32, tsg_sparse_format_*); there is no CPU fallback, so only T = float -- the only instantiation the
// reference itself uses (SparseGEMM.cpp:109,128) -- is provided.
//
// So that the reference's own driver (SparseGEMM.cpp) compiles unchanged against this header, the CPU-side benchmark
// support it takes from the reference header is provided too, re-written here: the dense checkers GEMM / GEMM_PReLU
// (SparseGEMM.h:121-149), the input generators initX / generateSparseMatrix (SparseGEMM.h:42-102) and compare_results
// (SparseGEMM.h:171-184).  They run on the CPU by design -- they are the driver's reference and its inputs, not the hot
// path -- and, like the reference header, this one ends up injecting `using namespace std` because the driver relies on it.
#pragma once

#include <cmath>
#include <cstdlib>
#include <ctime>
#include <iostream>
#include <random>
#include <stdexcept>
#include <type_traits>
#include <vector>

using namespace std;  // SparseGEMM.cpp:95-101 writes vector<...>, cout, ... unqualified (reference SparseGEMM.h:10)

extern "C" {
int tsg_sparse_gemm_f32(const float *X, const int *col_start_pos, const int *col_start_neg, const int *row_index_pos,
                        const int *row_index_neg, const float *b, float *Y, int M, int N, int K, float a, int use_prelu);
int tsg_sparse_format_build_i32(const int *matrix, int K, int N, void **handle, int *n_pos, int *n_neg);
int tsg_sparse_format_fetch(void *handle, int *csp, int *csn, int *rip, int *rin);
const char *sparse_last_error(void);
}

class SparseFormat {
public:
    std::vector<int> col_start_pos;
    std::vector<int> col_start_neg;
    std::vector<int> row_index_pos;
    std::vector<int> row_index_neg;

    SparseFormat(int *matrix, int K, int N) {
        void *h = nullptr;
        int np = 0, nn = 0;
        if (tsg_sparse_format_build_i32(matrix, K, N, &h, &np, &nn) != 0) throw std::runtime_error(sparse_last_error());
        col_start_pos.resize((size_t)N + 1);
        col_start_neg.resize((size_t)N + 1);
        row_index_pos.resize((size_t)np);
        row_index_neg.resize((size_t)nn);
        // vectors of size 0 may hand out data()==nullptr; the library never dereferences a pointer for a zero count
        if (tsg_sparse_format_fetch(h, col_start_pos.data(), col_start_neg.data(), row_index_pos.data(), row_index_neg.data()) != 0)
            throw std::runtime_error(sparse_last_error());
    }
};

template <typename T>
void sparseGEMM(T *X, int *col_start_pos, int *col_start_neg, int *row_index_pos, int *row_index_neg, T *b, T *Y, int M, int N, int K) {
    static_assert(std::is_same<T, float>::value, "libtsgemm_b200 computes in fp32 only (no CPU fallback for other T)");
    tsg_sparse_gemm_f32(X, col_start_pos, col_start_neg, row_index_pos, row_index_neg, b, Y, M, N, K, 0.0f, 0);
}

template <typename T>
void sparseGEMM_PReLU(T *X, int *col_start_pos, int *col_start_neg, int *row_index_pos, int *row_index_neg, T *b, T *Y, int M, int N, int K, T a) {
    static_assert(std::is_same<T, float>::value, "libtsgemm_b200 computes in fp32 only (no CPU fallback for other T)");
    tsg_sparse_gemm_f32(X, col_start_pos, col_start_neg, row_index_pos, row_index_neg, b, Y, M, N, K, a, 1);
}

// ---- CPU-side benchmark support for the reference's driver -------------------------------------------------------------

// integer-valued inputs in [-Range, Range] (reference SparseGEMM.h:42-51: mt19937 seeded with time(0))
template <typename T>
vector<T> initX(int LEN, int Range) {
    mt19937 rng(static_cast<unsigned int>(time(nullptr)));
    uniform_int_distribution<int> pick(-Range, Range);
    vector<T> x(static_cast<size_t>(LEN));
    for (auto &v : x) v = static_cast<T>(pick(rng));
    return x;
}

// H x W ternary matrix, about W/nonZero non-zeros per row (reference SparseGEMM.h:53-102).
//   uniformDistribution: every window of 2*nonZero columns receives one +1 and one -1 on even offsets;
//   otherwise: per row, (W/nonZero)/2 + d entries are +1 and (W/nonZero)/2 - d are -1 at random free positions, with
//   d drawn from [0, W/nonZero/20 + 1] so that the rows (and hence the columns' +1/-1 lists) are unbalanced.
template <typename T>
vector<T> generateSparseMatrix(int H, int W, int nonZero, bool uniformDistribution) {
    vector<T> mat(static_cast<size_t>(H) * W, T(0));
    if (uniformDistribution) {
        const int window = 2 * nonZero;
        for (int h = 0; h < H; ++h)
            for (int w0 = 0; w0 < W; w0 += window) {
                const int plus = (rand() % nonZero) * 2;
                int minus = (rand() % nonZero) * 2;
                mat[static_cast<size_t>(h) * W + w0 + plus] = T(1);
                while (minus == plus) minus = (rand() % nonZero) * 2;
                mat[static_cast<size_t>(h) * W + w0 + minus] = T(-1);
            }
        return mat;
    }
    mt19937 rng(static_cast<unsigned int>(time(nullptr)));
    uniform_int_distribution<int> column(0, W - 1);
    uniform_int_distribution<int> skew(0, int(W / nonZero / 20 + 1));
    auto scatter = [&](T *row, int howMany, T value) {
        for (int placed = 0; placed < howMany;) {
            const int c = column(rng);
            if (row[c] == T(0)) {
                row[c] = value;
                ++placed;
            }
        }
    };
    for (int h = 0; h < H; ++h) {
        const int d = skew(rng), half = (W / nonZero) / 2;
        scatter(&mat[static_cast<size_t>(h) * W], half + d, T(1));
        scatter(&mat[static_cast<size_t>(h) * W], half - d, T(-1));
    }
    return mat;
}

// dense checkers: y = sum_k X[m,k]*W[k,n] accumulated in T over ascending k, then + b[n] (reference SparseGEMM.h:121-149)
template <typename T>
void GEMM(T *X, T *W, T *b, T *Y, int M, int N, int K) {
#ifdef _OPENMP
#pragma omp parallel for
#endif
    for (int m = 0; m < M; ++m) {
        const T *x = X + static_cast<size_t>(m) * K;
        for (int n = 0; n < N; ++n) {
            T acc = 0;
            for (int k = 0; k < K; ++k) acc += x[k] * W[static_cast<size_t>(k) * N + n];
            Y[static_cast<size_t>(m) * N + n] = acc + b[n];
        }
    }
}

template <typename T>
void GEMM_PReLU(T *X, T *W, T *b, T *Y, int M, int N, int K, T a) {
#ifdef _OPENMP
#pragma omp parallel for
#endif
    for (int m = 0; m < M; ++m) {
        const T *x = X + static_cast<size_t>(m) * K;
        for (int n = 0; n < N; ++n) {
            T acc = 0;
            for (int k = 0; k < K; ++k) acc += x[k] * W[static_cast<size_t>(k) * N + n];
            acc += b[n];
            Y[static_cast<size_t>(m) * N + n] = (acc < 0) ? a * acc : acc;
        }
    }
}

// absolute tolerance 10e-6 = 1e-5, first offender reported (reference SparseGEMM.h:171-184)
template <typename T>
bool compare_results(T *result, T *groundTruth, int H, int W) {
    for (int h = 0; h < H; ++h)
        for (int w = 0; w < W; ++w) {
            const size_t i = static_cast<size_t>(h) * W + w;
            if (abs(result[i] - groundTruth[i]) > 10e-6) {
                cout << "Error at: H=" << h << ", W=" << w << ", result=" << result[i] << ", groundTruth=" << groundTruth[i] << endl;
                return false;
            }
        }
    return true;
}
