// eval.cuh -- streaming evaluate_kernel! for the single-component kernels (see eval.cu).
#pragma once
#include "common.cuh"

namespace mct {
size_t sc3d_eval_scratch_doubles(int Nk, int n);
// n evaluations with the same tables: F2 is [n][Nk]; F1 is [n][Nk] with stride f1_stride (0 = the same F1 for
// every evaluation; pass F1 == F2, f1_stride == Nk for the collective kernel).
int sc3d_eval_launch(int Nk, int n, const double *dk, const double *dV1, const double *dV2, const double *dV3, const double *dF1,
                     long long f1_stride, const double *dF2, double *d_scratch, double *d_out, cudaStream_t s);
}  // namespace mct
