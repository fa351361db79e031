// score_launch.cu -- translation unit of the score pass (K1+K2): tcgen05 kernels, the fp64 cross-check kernel and
// their launch logic.  See score_tc.cuh for the math; score_launch.h for the interface used by spagxe_cuda.cu.
#include <cuda_runtime.h>
#include <cuda.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "score_launch.h"
#include "score.cuh"
#include "score_tc.cuh"
#include "score_tc3.cuh"
#include "score_tc4.cuh"

namespace spagxe {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static void *g_encode_tiled = nullptr;

static int64_t round_up64(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

cudaError_t score_init(std::string *err) {
    cudaError_t e;
#define SMEM_OPT_IN(k)                                                                                          \
    if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes)) != cudaSuccess) { \
        *err = "cudaFuncSetAttribute(" #k ")"; return e;                                                        \
    }
    SMEM_OPT_IN(k_score_unpack_mma);
#ifdef SPAGXE_PROBES
    SMEM_OPT_IN((k_score_tc<8, 3>));
    SMEM_OPT_IN(k_score_tc3_probe<0>);
    SMEM_OPT_IN(k_score_tc3_probe<1>);
    SMEM_OPT_IN(k_score_tc3_probe<2>);
    SMEM_OPT_IN(k_score_tc3_probe<3>);
    SMEM_OPT_IN(k_score_tc3_probe<4>);
    SMEM_OPT_IN(k_score_tc3_probe<5>);
    SMEM_OPT_IN(k_score_tc3_probe<6>);
    SMEM_OPT_IN(k_score_tc3_probe<7>);
    SMEM_OPT_IN(k_score_tc3_probe<8>);
    SMEM_OPT_IN(k_score_tc3_probe<9>);
    SMEM_OPT_IN(k_score_tc3_probe<10>);
    SMEM_OPT_IN(k_score_tc3_probe<12>);
    SMEM_OPT_IN(k_score_tc3_probe<11>);
    SMEM_OPT_IN((k_score_tc4_probe<4, 1>));
    SMEM_OPT_IN((k_score_tc4_probe<6, 1>));
    SMEM_OPT_IN((k_score_tc4_probe<3, 1>));
    SMEM_OPT_IN((k_score_tc4_probe<2, 2>));
    SMEM_OPT_IN((k_score_tc4_probe<2, 3>));
    SMEM_OPT_IN((k_score_tc4_probe<3, 2, 1>));
    SMEM_OPT_IN((k_score_tc4_probe<3, 2, 2>));
    SMEM_OPT_IN((k_score_tc4_probe<3, 2, 3>));
    SMEM_OPT_IN((k_score_tc4_probe<3, 2, 4>));
    SMEM_OPT_IN((k_score_tc4_probe<3, 2, 6>));
    SMEM_OPT_IN((k_score_tc4_probe<3, 2, 7>));
    SMEM_OPT_IN((k_score_tc4_probe<3, 2, 8>));
    SMEM_OPT_IN((k_score_tc4_probe<3, 2, 9>));
    SMEM_OPT_IN((k_score_tc4_probe<3, 2, 10>));
    SMEM_OPT_IN((k_score_tc4_probe<3, 2, 11>));
#endif
#undef SMEM_OPT_IN
    if (!g_encode_tiled) {
        cudaDriverEntryPointQueryResult qres;
        e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &g_encode_tiled, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !g_encode_tiled) {
            *err = "cudaGetDriverEntryPoint(cuTensorMapEncodeTiled)";
            return e == cudaSuccess ? cudaErrorUnknown : e;
        }
    }
    return cudaSuccess;
}

int64_t score_bmat_bytes(int64_t n) { return ((n + kTcTileSamples - 1) / kTcTileSamples) * (int64_t)kTcBTileBytes; }
int score_acc_rows(int mc) { return (int)round_up64(mc, kTcRows); }

cudaError_t score_build_b(cudaStream_t s, const double *R, const double *V, int64_t n, BmatInfo *d_info, int8_t *d_bmat,
                          int64_t *launches) {
    cudaError_t e = cudaMemsetAsync(d_bmat, 0, score_bmat_bytes(n), s);
    if (e != cudaSuccess) return e;
    k_tc_scale<<<1, 1024, 0, s>>>(R, V, n, d_info);
    k_tc_build_b<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(R, V, n, d_info, d_bmat);
    *launches += 2;
    return cudaGetLastError();
}

#ifdef SPAGXE_PROBES
template <int NBOX>
static float run_probe(cudaStream_t s, int num_sms, const CUtensorMap &map, int n_tiles, int tps, int n_splits, int n_row_tiles,
                       long long *d_acc) {
    const size_t smem = 192 * 1024 + 1024 + 256;
    cudaFuncSetAttribute(k_tma_probe<NBOX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int grid = std::min(n_row_tiles * n_splits, num_sms);
    k_tma_probe<NBOX><<<grid, 64, smem, s>>>(map, n_tiles, tps, n_splits, n_row_tiles, (unsigned long long *)d_acc);
    cudaEventRecord(a, s);
    k_tma_probe<NBOX><<<grid, 64, smem, s>>>(map, n_tiles, tps, n_splits, n_row_tiles, (unsigned long long *)d_acc);
    cudaEventRecord(b, s); cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a); cudaEventDestroy(b);
    return ms;
}

#endif

cudaError_t score_launch_tc(cudaStream_t s, int num_sms, int tc_cfg, const uint8_t *d_rows, const ScoreGeom &g, const int8_t *d_bmat,
                            const BmatInfo *d_info, long long *d_acc, SnpStats st, int64_t *launches, std::string *err, bool hom_pass) {
    const int mc = g.mc;
    const int m_pad = score_acc_rows(mc);
    cudaError_t e = cudaMemsetAsync(d_acc, 0, sizeof(long long) * (size_t)m_pad * 32, s);
    if (e != cudaSuccess) { *err = "cudaMemsetAsync(acc)"; return e; }
    CUtensorMap map;
    const cuuint64_t gdim[2] = {(cuuint64_t)g.dpitch, (cuuint64_t)mc};
    const cuuint64_t gstr[1] = {(cuuint64_t)g.dpitch};
    const cuuint32_t box[2] = {kTcTileBytes, kTcRows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult cr = ((EncodeTiledFn)g_encode_tiled)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void *)d_rows, gdim, gstr, box, estr,
                                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        char buf[160];
        snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d): pitch=%lld rows=%d", (int)cr, (long long)g.dpitch, mc);
        *err = buf;
        return cudaErrorInvalidValue;
    }
    const int n_tiles = (int)((g.n + kTcTileSamples - 1) / kTcTileSamples);
    const int n_row_tiles = m_pad / kTcRows;
    int want_splits = std::max(1, (4 * num_sms + n_row_tiles - 1) / n_row_tiles);
    want_splits = std::min(want_splits, std::max(1, n_tiles / 16));
    // int32 TMEM accumulators are exact while samples per split * 128 * 128 < 2^31 (unsigned view of the missing code 0x80
    // against a digit of -128): at most 255 tiles of 512 samples per split
    want_splits = std::max(want_splits, (n_tiles + kT4MaxTilesPerSplit - 1) / kT4MaxTilesPerSplit);
    const int tiles_per_split = (n_tiles + want_splits - 1) / want_splits;
    const int n_splits = (n_tiles + tiles_per_split - 1) / tiles_per_split;
    const int n_items = n_row_tiles * n_splits;
    const int grid = std::min(n_items, num_sms);   // round-1 kernels of the probe build: whole items, round-robin
    (void)grid;
    // k_score_unpack_mma: one CTA per SM, each takes an equal contiguous range of the (row tile, sample tile) units
    const long long n_units = (long long)n_row_tiles * n_tiles;
    const int grid4 = (int)std::min<long long>(num_sms, std::max<long long>(1, n_units / 8));
#define TC_ARGS map, d_bmat, n_tiles, tiles_per_split, n_splits, n_row_tiles, mc, d_acc
    // A1-dosage byte per PLINK code: 00 -> 2, 01 (missing) -> 0x80, 10 -> 1, 11 -> 0.  The second pass of G.model Dom / Rec
    // counts hom-A1 samples only (00 -> 1), which separates the dosage sums into their het and hom parts
    const uint32_t lut = hom_pass ? 0x00008001u : 0x00018002u;
#ifdef SPAGXE_PROBES
    if (tc_cfg == 99) {   // one-off diagnostic of the access pattern
        const double gb = (double)mc * g.row_bytes / 1e9;
        float m1 = run_probe<1>(s, num_sms, map, n_tiles, tiles_per_split, n_splits, n_row_tiles, d_acc);
        float m2 = run_probe<2>(s, num_sms, map, n_tiles, tiles_per_split, n_splits, n_row_tiles, d_acc);
        float m4 = run_probe<4>(s, num_sms, map, n_tiles, tiles_per_split, n_splits, n_row_tiles, d_acc);
        fprintf(stderr, "[tma probe] rows=%d %.3f GB: 1 box %.3f ms (%.0f GB/s), 2 boxes %.3f ms (%.0f GB/s), 4 boxes %.3f ms (%.0f GB/s)\n", mc, gb,
                m1, gb / m1 * 1e3, m2, gb / m2 * 1e3, m4, gb / m4 * 1e3);
    }
    switch (tc_cfg) {   // probe builds only: schedule variants and timing probes (several give WRONG sums by design)
        case 1: k_score_tc<8, 3><<<grid, 12 * 32, kTcSmemBytes, s>>>(TC_ARGS); break;
        case 9: k_score_tc3_probe<0><<<grid, kT3Threads, kTcSmemBytes, s>>>(TC_ARGS); break;    // whole-round hand-over
        case 11: k_score_tc3_probe<1><<<grid, kT3Threads, kTcSmemBytes, s>>>(TC_ARGS); break;   // no MMAs
        case 12: k_score_tc3_probe<2><<<grid, kT3Threads, kTcSmemBytes, s>>>(TC_ARGS); break;   // no expansion
        case 32: k_score_tc3_probe<12><<<grid, kT3Threads, kTcSmemBytes, s>>>(TC_ARGS); break;  // first chunk expanded early
        case 27: k_score_tc3_probe<7><<<grid, kT3Threads, kTcSmemBytes, s>>>(TC_ARGS); break;   // half-round hand-over only
        case 29: k_score_tc3_probe<9><<<grid, kT3Threads, kTcSmemBytes, s>>>(TC_ARGS); break;   // hand-over after 3 of 4 chunks
        case 30: k_score_tc3_probe<10><<<grid, kT3Threads, kTcSmemBytes, s>>>(TC_ARGS); break;  // hand-over after 1 of 4 chunks
        case 28: k_score_tc3_probe<8><<<grid, kT3Threads, kTcSmemBytes, s>>>(TC_ARGS); break;   // per-chunk hand-over
        case 19: k_score_tc3_probe<6><<<grid, kT3Threads, kTcSmemBytes, s>>>(TC_ARGS); break;   // expansion token ring
        case 18: k_score_tc3_probe<5><<<grid, kT3Threads, kTcSmemBytes, s>>>(TC_ARGS); break;   // only group 0 expands
        case 17: k_score_tc3_probe<4><<<grid, kT3Threads, kTcSmemBytes, s>>>(TC_ARGS); break;   // B tiles not loaded
        case 16: k_score_tc3_probe<3><<<grid, kT3Threads, kTcSmemBytes, s>>>(TC_ARGS); break;   // hand-over timing trace
        case 10: k_score_tc3_probe<11><<<grid, kT3Threads, kTcSmemBytes, s>>>(TC_ARGS); break;   // round 1's product kernel
        case 41: k_score_tc4_probe<4, 1><<<grid4, t4_threads(4), kTcSmemBytes, s>>>(TC_ARGS, lut); break;   // 4 groups, single buffers
        case 61: k_score_tc4_probe<6, 1><<<grid4, t4_threads(6), kTcSmemBytes, s>>>(TC_ARGS, lut); break;   // 6 groups, single buffers
        case 31: k_score_tc4_probe<3, 1><<<grid4, t4_threads(3), kTcSmemBytes, s>>>(TC_ARGS, lut); break;   // 3 groups, single buffers
        case 22: k_score_tc4_probe<2, 2><<<grid4, t4_threads(2), kTcSmemBytes, s>>>(TC_ARGS, lut); break;   // 2 groups, ping-pong
        case 23: k_score_tc4_probe<2, 3><<<grid4, t4_threads(2), kTcSmemBytes, s>>>(TC_ARGS, lut); break;   // 2 groups, 3 buffers
        case 51: k_score_tc4_probe<3, 2, 1><<<grid4, t4_threads(3), kTcSmemBytes, s>>>(TC_ARGS, lut); break;   // no UTCIMMA (wrong sums)
        case 52: k_score_tc4_probe<3, 2, 2><<<grid4, t4_threads(3), kTcSmemBytes, s>>>(TC_ARGS, lut); break;   // signed view only (wrong sums)
        case 53: k_score_tc4_probe<3, 2, 3><<<grid4, t4_threads(3), kTcSmemBytes, s>>>(TC_ARGS, lut); break;   // no expansion ALU (wrong sums)
        case 54: k_score_tc4_probe<3, 2, 4><<<grid4, t4_threads(3), kTcSmemBytes, s>>>(TC_ARGS, lut); break;   // no tcgen05.st (wrong sums)
        case 58: k_score_tc4_probe<3, 2, 8><<<grid4, t4_threads(3), kTcSmemBytes, s>>>(TC_ARGS, lut); break;   // software-pipelined barrier probes (exact)
        case 59: k_score_tc4_probe<3, 2, 9><<<grid4, t4_threads(3), kTcSmemBytes, s>>>(TC_ARGS, lut); break;   // 58 with the timing trace
        case 60: k_score_tc4_probe<3, 2, 10><<<grid4, t4_threads(3), kTcSmemBytes, s>>>(TC_ARGS, lut); break;   // in-body timing trace (exact)
        case 62: k_score_tc4_probe<3, 2, 11><<<grid4, t4_threads(3), kTcSmemBytes, s>>>(TC_ARGS, lut); break;   // incremental loop state (exact)
        case 57: k_score_tc4_probe<3, 2, 7><<<grid4, t4_threads(3), kTcSmemBytes, s>>>(TC_ARGS, lut); break;   // timing trace (exact)
        case 56: k_score_tc4_probe<3, 2, 6><<<grid4, t4_threads(3), kTcSmemBytes, s>>>(TC_ARGS, lut); break;   // packed words read one round ahead (exact)
        default: k_score_unpack_mma<<<grid4, kT4Threads, kTcSmemBytes, s>>>(TC_ARGS, lut); break;
    }
    const bool tc3 = (tc_cfg != 0 && tc_cfg != 41 && tc_cfg != 61 && tc_cfg != 31 && tc_cfg != 22 && tc_cfg != 23 && (tc_cfg < 51 || tc_cfg > 62));
#else
    (void)tc_cfg;   // the product library has exactly one tensor-core score kernel
    k_score_unpack_mma<<<grid4, kT4Threads, kTcSmemBytes, s>>>(TC_ARGS, lut);
    const bool tc3 = false;
#endif
#undef TC_ARGS
    if ((e = cudaGetLastError()) != cudaSuccess) { *err = "k_score_unpack_mma launch"; return e; }
#ifdef SPAGXE_PROBES
    if (tc_cfg == 16) {   // dump the timing trace of CTA 0
        static int dumped = 0;
        cudaStreamSynchronize(s);
        if (dumped++ == 3) {
            static unsigned long long h[kT3Warps * kT3TraceRounds * kT3TraceEv];
            cudaMemcpyFromSymbol(h, g_tc_trace, sizeof h);
            unsigned long long t0 = ~0ull;
            for (size_t i = 0; i < sizeof h / sizeof h[0]; i++) if (h[i] && h[i] < t0) t0 = h[i];
            for (int w = 0; w < kT3Warps; w++) {
                if (w == kT3ExpWarps || w == kT3ExpWarps + 1) continue;
                for (int r = 0; r < kT3TraceRounds; r++) {
                    fprintf(stderr, "[trace] warp %2d round %2d:", w, r);
                    for (int k = 0; k < 6; k++) {
                        unsigned long long v = h[((size_t)w * kT3TraceRounds + r) * kT3TraceEv + k];
                        fprintf(stderr, " %8lld", v ? (long long)(v - t0) : -1ll);
                    }
                    fprintf(stderr, "\n");
                }
            }
        }
    }
#endif
#ifdef SPAGXE_PROBES
    if (tc_cfg == 57 || tc_cfg == 59 || tc_cfg == 60) {   // dump the tc4 timing trace of CTA 0 (fourth call)
        static int dumped4 = 0;
        cudaStreamSynchronize(s);
        if (dumped4++ == 3) {
            static unsigned long long h[kT4TraceWarps * kT4TraceRounds * kT4TraceEv];
            cudaMemcpyFromSymbol(h, g_t4_trace, sizeof h);
            unsigned long long t0 = ~0ull;
            for (size_t i = 0; i < sizeof h / sizeof h[0]; i++) if (i % kT4TraceEv < 6 && h[i] && h[i] < t0) t0 = h[i];
            for (int w = 0; w < t4_threads(3) / 32; w++) {
                for (int r = 0; r < kT4TraceRounds; r++) {
                    bool any = false;
                    for (int k = 0; k < kT4TraceEv; k++) any |= h[((size_t)w * kT4TraceRounds + r) * kT4TraceEv + k] != 0;
                    if (!any) continue;
                    fprintf(stderr, "[trace4] warp %2d use %2d:", w, r + 16);
                    for (int k = 0; k < 7; k++) {
                        unsigned long long v = h[((size_t)w * kT4TraceRounds + r) * kT4TraceEv + k];
                        fprintf(stderr, " %8lld", v ? (k < 6 ? (long long)(v - t0) : (long long)v) : -1ll);
                    }
                    fprintf(stderr, "\n");
                }
            }
        }
    }
#endif
    if (hom_pass) k_score_tc4_reduce_hom<<<(mc + 127) / 128, 128, 0, s>>>(d_acc, mc, d_info, st);
    else if (tc3) k_score_tc_reduce<<<(mc + 127) / 128, 128, 0, s>>>(d_acc, mc, d_info, st);
    else k_score_tc4_reduce<<<(mc + 127) / 128, 128, 0, s>>>(d_acc, mc, d_info, st);
    *launches += 2;
    if ((e = cudaGetLastError()) != cudaSuccess) { *err = "k_score_tc_reduce launch"; return e; }
    return cudaSuccess;
}

int64_t score_v1_partial_elems(const ScoreGeom &g, int chunk_m) {
    const int tiles_s = (int)((g.dpitch / 4 + kScoreThreads - 1) / kScoreThreads);
    return (int64_t)tiles_s * 4 * round_up64(chunk_m, kScoreRows);
}

cudaError_t score_launch_v1(cudaStream_t s, const uint8_t *d_rows, const ScoreGeom &g, const double *R, const double *V,
                            double *d_partial, int *d_ipartial, SnpStats st, int64_t *launches) {
    const int64_t pw = g.dpitch / 4;
    const int tiles_s = (int)((pw + kScoreThreads - 1) / kScoreThreads);
    const int m_pad = (int)round_up64(g.mc, kScoreRows);
    dim3 grid((unsigned)tiles_s, (unsigned)(m_pad / kScoreRows));
    k_score_v1<<<grid, kScoreThreads, 0, s>>>((const uint32_t *)d_rows, pw, g.n, g.mc, m_pad, R, V, d_partial, d_ipartial);
    k_score_reduce<<<(g.mc + 127) / 128, 128, 0, s>>>(d_partial, d_ipartial, tiles_s, g.mc, m_pad, st);
    *launches += 2;
    return cudaGetLastError();
}

}  // namespace spagxe
