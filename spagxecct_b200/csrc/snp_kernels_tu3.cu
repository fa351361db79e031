// snp_kernels_tu3.cu -- translation unit 3 of the per-SNP test kernels (kernels.cuh): k_mix_dosage_snp, k_mixplus_dosage_snp, k_plus_dosage_snp.
// The heavy cluster kernels take a minute or more of ptxas each; one group per .cu file lets `make -j` compile them side by side.
#define SPAGXE_TU 3
#include <cuda_runtime.h>
#include <climits>
#include "snp_kernels.h"
#include "kernels.cuh"
