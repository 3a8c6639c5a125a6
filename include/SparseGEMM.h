// SparseGEMM.h -- C++ header: drop-in for the reference's SparseGEMM.h entry points on the hot path.
//
//   class SparseFormat(int* matrix, int K, int N)   reference SparseGEMM.h:13-40  (int matrix, >=1 / <=-1)
//   sparseGEMM<T>(X, csp, csn, rip, rin, b, Y, M, N, K)          reference SparseGEMM.h:104-119
//   sparseGEMM_PReLU<T>(X, csp, csn, rip, rin, b, Y, M, N, K, a) reference SparseGEMM.h:151-168
//
// Same names, argument order and meaning.  Work runs on the GPU through libtsgemm_b200.so's C-ABI
// (tsg_sparse_gemm_f32, tsg_sparse_format_*); there is no CPU fallback, so only T = float -- the only instantiation the
// reference itself uses (SparseGEMM.cpp:109,128) -- is provided.  The reference's dense helpers and generators
// (GEMM, GEMM_PReLU, initX, generateSparseMatrix, compare_results: SparseGEMM.h:42-102,121-149,171-184) are benchmark
// support that stays on the CPU and is not part of this header.  Unlike the reference this header does not inject
// `using namespace std`.
#pragma once

#include <stdexcept>
#include <type_traits>
#include <vector>

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
