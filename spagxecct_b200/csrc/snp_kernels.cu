// snp_kernels.cu -- translation unit of the per-SNP test kernels (kernels.cuh): SNP-invariant vector statistics,
// R-matrix packing, decode / counts, the per-SNP epilogues with routing, the SPA kernels (quadrature and exact), the
// SPAGxE_Plus sparse-GRM branch B and the per-SNP SPAGxEmix kernel.  Compiled on its own (the template-heavy cluster
// kernels take minutes) so that the host pipeline in spagxe_cuda.cu rebuilds in seconds.
#define SPAGXE_TU 1   // the small kernels and k_mix_snp; snp_kernels_tu{2,3,4}.cu hold the other heavy kernels
#include <cuda_runtime.h>
#include <climits>
#include "snp_kernels.h"
#include "kernels.cuh"

namespace spagxe {

cudaError_t launch_pack_dense(cudaStream_t s, int f64, const void *G, int64_t ld, int64_t n, int m, uint32_t *bed,
                              int64_t pitch_words, Counters *ctr) {
    dim3 grid((unsigned)((pitch_words + 255) / 256), (unsigned)m);
    if (f64) k_pack_dense<double><<<grid, 256, 0, s>>>((const double *)G, ld, n, m, bed, pitch_words, ctr);
    else k_pack_dense<int><<<grid, 256, 0, s>>>((const int *)G, ld, n, m, bed, pitch_words, ctr);
    return cudaGetLastError();
}

}  // namespace spagxe
