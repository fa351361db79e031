// snp_kernels_tu4.cu -- translation unit 4 of the per-SNP test kernels (kernels.cuh): k_cct_dosage_snp, k_mixplus_snp.
// The heavy cluster kernels take a minute or more of ptxas each; one group per .cu file lets `make -j` compile them side by side.
#define SPAGXE_TU 4
#include <cuda_runtime.h>
#include <climits>
#include "snp_kernels.h"
#include "kernels.cuh"
