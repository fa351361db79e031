// snp_kernels_tu2.cu -- translation unit 2 of the per-SNP test kernels (kernels.cuh): k_localance_snp.
// The heavy cluster kernels take a minute or more of ptxas each; one group per .cu file lets `make -j` compile them side by side.
#define SPAGXE_TU 2
#include <cuda_runtime.h>
#include <climits>
#include "snp_kernels.h"
#include "kernels.cuh"
