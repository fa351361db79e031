// score_launch.h -- host interface of the score-pass translation unit (score_launch.cu).  The score kernels
// (tcgen05 variants + the fp64 cross-check kernel) are compiled separately from the per-SNP test kernels so that the
// two halves of the library build in parallel.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <string>
#include "types.cuh"

namespace spagxe {

struct ScoreGeom {
    int64_t n;          // samples
    int64_t row_bytes;  // ceil(n / 4)
    int64_t dpitch;     // device pitch of the packed rows
    int mc;             // SNP rows in this chunk
};

// one-time per-process setup: opt-in shared memory sizes, driver entry point for tensor maps
cudaError_t score_init(std::string *err);
int64_t score_bmat_bytes(int64_t n);
int score_acc_rows(int mc);   // rows of the int64 accumulator block (32 x int64 each)
// build the int8 digit matrix B of (R, V); `bmat` must hold score_bmat_bytes(n) bytes
cudaError_t score_build_b(cudaStream_t s, const double *R, const double *V, int64_t n, BmatInfo *d_info, int8_t *d_bmat, int64_t *launches);
// tensor-core score pass (+ digit recombination) into `st`; returns cudaSuccess or sets *err
cudaError_t score_launch_tc(cudaStream_t s, int num_sms, int tc_cfg, const uint8_t *d_rows, const ScoreGeom &g, const int8_t *d_bmat,
                            const BmatInfo *d_info, long long *d_acc, SnpStats st, int64_t *launches, std::string *err, bool hom_pass = false);
// fp64 CUDA-core cross-check kernel (SPAGXE_SCORE=v1)
int64_t score_v1_partial_elems(const ScoreGeom &g, int chunk_m);
cudaError_t score_launch_v1(cudaStream_t s, const uint8_t *d_rows, const ScoreGeom &g, const double *R, const double *V,
                            double *d_partial, int *d_ipartial, SnpStats st, int64_t *launches);
constexpr int kScoreRowsPad = 128;

}  // namespace spagxe
