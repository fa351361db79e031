// snp_kernels.h -- declarations of the per-SNP test kernels (defined in kernels.cuh, compiled once in snp_kernels.cu).
// The host pipeline (spagxe_cuda.cu) launches them through these declarations, so that an edit of the host code does
// not recompile the kernels (and vice versa).  DESIGN.md section 4 lists what each kernel does.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "types.cuh"

namespace spagxe {

__global__ void __launch_bounds__(1024) k_vec_prep1(const double *__restrict__ R, const double *__restrict__ E, int64_t n, VecConst *vc);
__global__ void __launch_bounds__(1024) k_vec_prep2(const double *__restrict__ R, const double *__restrict__ E, const double *__restrict__ Rn_in, int64_t n, int mode, double *__restrict__ Rn, double *__restrict__ V, VecConst *vc);
// R matrix (REALSXP: f64 != 0, INTSXP: f64 == 0) -> packed PLINK rows; host launcher of the template kernel k_pack_dense
cudaError_t launch_pack_dense(cudaStream_t s, int f64, const void *G, int64_t ld, int64_t n, int m, uint32_t *bed, int64_t pitch_words, Counters *ctr);
__global__ void k_bed_prepare(const uint8_t *__restrict__ src, int64_t spitch, uint32_t *__restrict__ dst, int64_t dpitch_words, int64_t n, int m, const int64_t *__restrict__ idx, int count_a2);
__global__ void k_bed_counts(const uint8_t *__restrict__ bed, int64_t pitch, int64_t n, int m, long long *counts);
__global__ void k_bed_decode(const uint8_t *__restrict__ bed, int64_t pitch, int64_t n, int m, double *__restrict__ out);
__global__ void k_finalize_cct(SnpStats st, int m_chunk, int64_t snp0, int64_t m_total, int64_t n, const VecConst *__restrict__ vcp, double epsilon, double min_maf, double *__restrict__ out, SpaItem *spa_list, BItem *b_list, const uint8_t *bed, int64_t pitch, Counters *ctr, int g_model);
__global__ void __launch_bounds__(1024) k_quad_range(const double *__restrict__ rv, int64_t n, Quad q);
__global__ void k_quad_accumulate(const double *__restrict__ rv, int64_t n, Quad q);
__global__ void __launch_bounds__(1024) k_quad_finish(Quad q);
__global__ void __launch_bounds__(kSpaQThreads) k_spa_quad(const SpaItem *__restrict__ list, Counters *ctr, Quad q, SpaItem *__restrict__ exact_list, int64_t m_total, double *__restrict__ out);
__global__ void __launch_bounds__(kSpaThreads) k_spa_homo(const SpaItem *__restrict__ list, const int *count_ptr, Counters *ctr, const double *__restrict__ rv, int64_t n, int64_t m_total, double *__restrict__ out);
__global__ void __launch_bounds__(1024) k_plus_nullmodel(const double *__restrict__ R, const double *__restrict__ E, int64_t n, GrmCoo grm, double *__restrict__ Rnew, double *__restrict__ out3);
__global__ void k_finalize_plus(SnpStats st, int m_chunk, int64_t snp0, int64_t m_total, int64_t n, const VecConst *__restrict__ vcp, PlusConst pc, double epsilon, double cutoff, double min_maf, double *__restrict__ out, SpaItem *spa_list, BItem *b_list, const uint8_t *bed, int64_t pitch, Counters *ctr, int g_model);
__global__ void __launch_bounds__(kBThreads) k_branch_b_plus(const BItem *__restrict__ b_list, const int *__restrict__ count_ptr, Counters *ctr, int64_t m_total, int64_t n, const double *__restrict__ R, const double *__restrict__ E, const VecConst *__restrict__ vcp, GrmCoo grm, double cutoff, double *__restrict__ out, int g_model);
__global__ void k_finalize_mix(SnpStats st, int m_chunk, int64_t snp0, int64_t m_total, int64_t n, double min_maf, double *__restrict__ out, int *b_list, Counters *ctr);
__global__ void __launch_bounds__(kBThreads) k_mix_snp(const int *__restrict__ list, Counters *ctr, const uint8_t *__restrict__ bed, int64_t pitch, SnpStats st, int64_t snp0, int64_t m_total, int64_t n, const double *__restrict__ R, const double *__restrict__ E, const VecConst *__restrict__ vcp, WaldData wd, MixData md, double epsilon, double cutoff, double *__restrict__ out, double *__restrict__ mcache_all, const double *__restrict__ hbeta, int hbeta_ld, const MixHand *__restrict__ hand);
__global__ void __launch_bounds__(kBThreads) k_mixplus_snp(const int *__restrict__ list, Counters *ctr, const uint8_t *__restrict__ bed, int64_t pitch, SnpStats st, int64_t snp0, int64_t m_total, int64_t n, const double *__restrict__ R, const double *__restrict__ E, const VecConst *__restrict__ vcp, MixData md, GrmCoo grm, double epsilon, double cutoff, double *__restrict__ out, double *__restrict__ mcache_all);

__global__ void __launch_bounds__(kBThreads) k_localance_snp(LocalArgs A);
__global__ void __launch_bounds__(kBThreads) k_cct_dosage_snp(DosageArgs A);
__global__ void __launch_bounds__(kBThreads) k_plus_dosage_snp(PlusDosageArgs A);
__global__ void __launch_bounds__(kBThreads) k_mix_dosage_snp(MixDosageArgs A);
__global__ void __launch_bounds__(kBThreads) k_mixplus_dosage_snp(MixDosageArgs A);

}  // namespace spagxe
