// score_tc3.cuh -- round-partitioned variant of the tensor-core score pass (see score_tc.cuh for the math).
//
// ncu + timing probes on k_score_tc (profiles/r01_score_tc_analysis.md) showed that neither HBM (the same TMA
// stream reaches 6.5 TB/s), nor the ALU pipe (50 %), nor the tensor pipe (18 %) bounds it: every expander warp
// and the MMA issuer take part in EVERY 256-sample round, so the per-round fixed latency of each warp
// (barrier wake-ups, fences, tcgen05.wait::st, operand set-up at CPI ~5) is on the critical path.
// Here rounds are dealt round-robin to three independent groups instead:
//     group q = 4 expander warps (128 TMEM lanes = 128 SNP rows) + 1 MMA warp + A buffer q (128 TMEM columns)
//     round G (256 samples of the 128-row tile) belongs to group G mod 3
// so a warp sees only every third round and the three groups overlap each other's latencies.  All MMAs
// accumulate (integer adds commute); the epilogue warps zero the accumulators after reading them.
#pragma once
#include "score_tc.cuh"
#ifdef SPAGXE_PROBES

namespace spagxe {

constexpr int kT3Groups = 3;
constexpr int kT3ExpWarps = 4 * kT3Groups;       // 12
constexpr int kT3Warps = kT3ExpWarps + 2 + kT3Groups;  // + geno producer, B producer, 3 MMA warps = 17
constexpr int kT3Threads = kT3Warps * 32;

// kProbe == 3: CTA 0 records clock64() at the hand-over points of rounds [16, 64) of every warp (timing trace)
constexpr int kT3TraceRounds = 48, kT3TraceEv = 8;
#ifdef SPAGXE_PROBES
__device__ unsigned long long g_tc_trace[kT3Warps * kT3TraceRounds * kT3TraceEv];
#define T3_TRACE(k)                                                                                        \
    if (kProbe == 3 && blockIdx.x == 0 && lane == 0 && use >= 16 && use < 16 + kT3TraceRounds)               \
        g_tc_trace[((size_t)warp * kT3TraceRounds + (use - 16)) * kT3TraceEv + (k)] = (unsigned long long)clock64();
#else
#define T3_TRACE(k)
#endif

// The kernel body is a template over the schedule variant.  The library ships exactly one instantiation
// kT3Product was round 1's product kernel; round 2's is score_tc4.cuh (one A matrix read as signed and unsigned bytes).
// Everything in this file only exists in builds with -DSPAGXE_PROBES (tools/, never the product .so).
constexpr int kT3Product = 11;   // half-round hand-over + packed bytes read before the A-buffer wait

template <int kProbe>
__device__ __forceinline__ void
score_tc3_body(const CUtensorMap &bedmap, const int8_t *__restrict__ Bmat, int n_tiles, int tiles_per_split,
               int n_splits, int n_row_tiles, int m_chunk, long long *__restrict__ acc) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bbase = sbase + kTcGStages * kTcGStageBytes;
    const uint32_t bars = bbase + kTcBStages * kTcBTileBytes;
    auto bar_gfull = [&](int s) { return bars + 8u * s; };
    auto bar_gempty = [&](int s) { return bars + 8u * (kTcGStages + s); };
    auto bar_bfull = [&](int s) { return bars + 8u * (2 * kTcGStages + s); };
    auto bar_bempty = [&](int s) { return bars + 8u * (2 * kTcGStages + kTcBStages + s); };
    constexpr int kB0 = 2 * kTcGStages + 2 * kTcBStages;
    auto bar_afull = [&](int b) { return bars + 8u * (kB0 + b); };
    auto bar_aempty = [&](int b) { return bars + 8u * (kB0 + kT3Groups + b); };
    const uint32_t bar_accfull = bars + 8u * (kB0 + 2 * kT3Groups), bar_accempty = bars + 8u * (kB0 + 2 * kT3Groups + 1);
    const uint32_t tmem_slot = bars + 8u * (kB0 + 2 * kT3Groups + 2);
    // kProbe == 6: expansion token ring.  Round G may start expanding only when round G - 1 has issued its last STTM, so
    // the three groups take the ALU / MIO in turn instead of all at once (the trace shows them phase-locked).
    auto bar_tok = [&](uint32_t q) { return bars + 8u * (kB0 + 2 * kT3Groups + 3 + q); };
    // kProbe == 7: the first 128 samples of a round are handed to the MMA warp while the second 128 are still expanding
    // kProbe == 8: the same per 64-sample chunk (three early hand-overs per round)
    // kProbe == 9 / 10: two parts split after chunk 3 / chunk 1
    // kProbe == 11: as 7, and the packed bytes are read before waiting for the A buffer (their latency hides in that wait)
    constexpr int kParts = (kProbe == 7 || kProbe == 9 || kProbe == 10 || kProbe == 11 || kProbe == 12) ? 2 : (kProbe == 8 ? 4 : 1);
    // kProbe == 12: as 11, and the first chunk is expanded into registers before that wait as well
    constexpr bool kLoadEarly = (kProbe == 11 || kProbe == 12);
    constexpr bool kExpandEarly = (kProbe == 12);
    constexpr int kFirst = (kProbe == 9) ? 3 : (kProbe == 10 ? 1 : 4 / kParts);   // chunks in every early part (2-part modes: in the first)
    auto bar_apart = [&](uint32_t q, uint32_t part) { return bars + 8u * (kB0 + 3 * kT3Groups + 3 + q * 3u + part); };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kTcGStages; s++) { mbar_init(bar_gfull(s), 1); mbar_init(bar_gempty(s), 8); }  // 2 rounds x 4 warps
        for (int s = 0; s < kTcBStages; s++) { mbar_init(bar_bfull(s), 1); mbar_init(bar_bempty(s), 2); }  // 1 commit per round
        for (int b = 0; b < kT3Groups; b++) { mbar_init(bar_afull(b), 4); mbar_init(bar_aempty(b), 1); }
        mbar_init(bar_accfull, kT3Groups); mbar_init(bar_accempty, 4);
        for (uint32_t b = 0; b < (uint32_t)kT3Groups; b++) { mbar_init(bar_tok(b), 4); for (uint32_t c = 0; c < 3; c++) mbar_init(bar_apart(b, c), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kT3ExpWarps) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);
    constexpr uint32_t kColAcc = kT3Groups * 128;
    if (warp < 4) {  // zero the accumulators once; afterwards the epilogue re-zeroes them
        const uint32_t z[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        const uint32_t la = tmem_base + ((uint32_t)(warp * 32) << 16);
        tmem_st16(la + kColAcc, z); tmem_st16(la + kColAcc + 16, z);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    const int n_items = n_row_tiles * n_splits;
    if (warp == kT3ExpWarps) {
        if (lane == 0) {   // ---- TMA producer: packed genotype tiles (HBM)
            uint32_t f = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int rt = item / n_splits, ks = item % n_splits;
                const int t0 = ks * tiles_per_split, t1 = min(n_tiles, t0 + tiles_per_split);
                for (int t = t0; t < t1; t++, f++) {
                    const int s = f % kTcGStages;
                    mbar_wait(bar_gempty(s), ((f / kTcGStages) & 1) ^ 1);
                    mbar_expect_tx(bar_gfull(s), kTcGStageBytes);
                    tma_load_2d(sbase + s * kTcGStageBytes, &bedmap, t * kTcTileBytes, rt * kTcRows, bar_gfull(s));
                }
            }
        }
    } else if (warp == kT3ExpWarps + 1) {
        if (lane == 0) {   // ---- bulk-copy producer: B tiles (L2)
            uint32_t f = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int ks = item % n_splits;
                const int t0 = ks * tiles_per_split, t1 = min(n_tiles, t0 + tiles_per_split);
                for (int t = t0; t < t1; t++, f++) {
                    const int s = f % kTcBStages;
                    mbar_wait(bar_bempty(s), ((f / kTcBStages) & 1) ^ 1);
                    if (kProbe == 4) { mbar_arrive(bar_bfull(s)); continue; }   // probe: no B traffic (stale tiles, wrong sums)
                    mbar_expect_tx(bar_bfull(s), kTcBTileBytes);
                    bulk_load_1d(bbase + s * kTcBTileBytes, Bmat + (int64_t)t * kTcBTileBytes, kTcBTileBytes, bar_bfull(s));
                }
            }
        }
    } else if (warp > kT3ExpWarps + 1) {
        // ---- MMA issuer of group q: both A matrices of buffer q, 16 UTCIMMA per round, everything accumulates
        const uint32_t q = (uint32_t)(warp - (kT3ExpWarps + 2));
        const uint32_t abase = tmem_base + q * 128u, dd = tmem_base + kColAcc, dm = tmem_base + kColAcc + 16;
        uint32_t gbase = 0, use = 0, itc = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int ks = item % n_splits;
            const int t0 = ks * tiles_per_split, t1 = min(n_tiles, t0 + tiles_per_split);
            const uint32_t nr = 2u * (uint32_t)max(t1 - t0, 0);
            if (nr == 0) continue;
            bool waited = false;
            for (uint32_t G = gbase + ((q + 3u - gbase % 3u) % 3u); G < gbase + nr; G += 3u, use++) {
                if (!waited) { mbar_wait(bar_accempty, (itc & 1) ^ 1); tc_fence_after(); waited = true; }
                const uint32_t F = G >> 1, hs = G & 1u, sb = F % kTcBStages;
                mbar_wait(bar_bfull(sb), (F / kTcBStages) & 1);
                T3_TRACE(0)
                if (kParts > 1) {   // early parts as soon as they are in TMEM
#pragma unroll
                    for (int part = 0; part < kParts - 1; part++) {
                        mbar_wait(bar_apart(q, (uint32_t)part), use & 1);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint64_t bd0 = tc_bdesc(bbase + sb * kTcBTileBytes + hs * (kTcBTileBytes / 2));
#pragma unroll
                            for (int c4 = part * kFirst; c4 < (part + 1) * kFirst; c4++) {
#pragma unroll
                                for (int j = 0; j < 2; j++) {
                                    const uint64_t bd = bd0 + (uint64_t)(((c4 * 4 + 2 * j) * 256) >> 4);
                                    const uint32_t acol = (uint32_t)(c4 * 16 + j * 8);
                                    tc_mma_i8_ts(dd, abase + acol, bd, kTcIdesc, 1u);
                                    tc_mma_i8_ts(dm, abase + acol + 64, bd, kTcIdesc, 1u);
                                }
                            }
                        }
                        __syncwarp();
                    }
                }
                mbar_wait(bar_afull(q), use & 1);
                T3_TRACE(1)
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t bd0 = tc_bdesc(bbase + sb * kTcBTileBytes + hs * (kTcBTileBytes / 2));
#pragma unroll
                    for (int c4 = (kParts - 1) * kFirst; c4 < 4; c4++) {
#pragma unroll
                        for (int j = 0; j < 2; j++) {
                            const uint64_t bd = bd0 + (uint64_t)(((c4 * 4 + 2 * j) * 256) >> 4);
                            const uint32_t acol = (uint32_t)(c4 * 16 + j * 8);
                            if (kProbe != 1) {
                                tc_mma_i8_ts(dd, abase + acol, bd, kTcIdesc, 1u);
                                tc_mma_i8_ts(dm, abase + acol + 64, bd, kTcIdesc, 1u);
                            }
                        }
                    }
                    tc_commit(bar_aempty(q));
                    tc_commit(bar_bempty(sb));
                }
                __syncwarp();
                T3_TRACE(2)
            }
            if (elect_one()) tc_commit(bar_accfull);
            __syncwarp();
            gbase += nr; itc++;
        }
    } else {
        // ---- expander warps: group q = warp / 4 fills A buffer q; thread = SNP row = TMEM lane
        const uint32_t q = (uint32_t)(warp >> 2), quarter = (uint32_t)(warp & 3);
        const int row = (int)quarter * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((quarter * 32u) << 16);
        const uint32_t swz = (uint32_t)(row & 7);
        uint32_t gbase = 0, use = 0, itc = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int rt = item / n_splits, ks = item % n_splits;
            const int t0 = ks * tiles_per_split, t1 = min(n_tiles, t0 + tiles_per_split);
            const uint32_t nr = 2u * (uint32_t)max(t1 - t0, 0);
            if (nr == 0) continue;
            for (uint32_t G = gbase + ((q + 3u - gbase % 3u) % 3u); G < gbase + nr; G += 3u, use++) {
                const uint32_t F = G >> 1, hs = G & 1u, s = F % kTcGStages;
                T3_TRACE(0)
                mbar_wait(bar_gfull(s), (F / kTcGStages) & 1);
                T3_TRACE(1)
                if (!kLoadEarly) {
                    mbar_wait(bar_aempty(q), (use & 1) ^ 1);
                    if (kProbe == 6) mbar_wait(bar_tok(q), q == 0 ? ((use & 1) ^ 1) : (use & 1));   // predecessor round has expanded
                    T3_TRACE(2)
                    tc_fence_after();
                }
                const uint32_t rowaddr = sbase + s * kTcGStageBytes + (uint32_t)row * kTcTileBytes;
                // all four 16-byte chunks of this row first (loads in flight together), then expand
                uint4 xv[4];
                if (kProbe != 2 && !(kProbe == 5 && q != 0)) {   // probe 5: only group 0 expands (contention vs. chain latency)
#pragma unroll
                    for (int c4 = 0; c4 < 4; c4++) {
                        const uint32_t chunk = hs * 4u + (uint32_t)c4;
                        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                                     : "=r"(xv[c4].x), "=r"(xv[c4].y), "=r"(xv[c4].z), "=r"(xv[c4].w)
                                     : "r"(rowaddr + ((chunk ^ swz) << 4)));
                    }
                }
                auto expand = [&](const uint4 &v, uint32_t (&d)[16], uint32_t (&m)[16]) {
                    const uint32_t x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint32_t so_lo = x[k], se_lo = x[k] * 4u, so_hi = __umulhi(x[k], 1u << 16), se_hi = __umulhi(x[k], 1u << 18);
                        d[4 * k + 0] = prmt(0xFEFEFEFEu, 0u, so_lo); m[4 * k + 0] = prmt(0u, 0x01010101u, so_lo);
                        d[4 * k + 1] = prmt(0xFEFEFEFEu, 0u, se_lo); m[4 * k + 1] = prmt(0u, 0x01010101u, se_lo);
                        d[4 * k + 2] = prmt(0xFEFEFEFEu, 0u, so_hi); m[4 * k + 2] = prmt(0u, 0x01010101u, so_hi);
                        d[4 * k + 3] = prmt(0xFEFEFEFEu, 0u, se_hi); m[4 * k + 3] = prmt(0u, 0x01010101u, se_hi);
                    }
                };
                uint32_t d0[16], m0[16];
                if (kExpandEarly) expand(xv[0], d0, m0);
                if (kLoadEarly) { mbar_wait(bar_aempty(q), (use & 1) ^ 1); tc_fence_after(); }
#pragma unroll
                for (int c4 = 0; c4 < ((kProbe == 2 || (kProbe == 5 && q != 0)) ? 0 : 4); c4++) {
                    uint32_t d[16], m[16];
                    if (kExpandEarly && c4 == 0) {
#pragma unroll
                        for (int k = 0; k < 16; k++) { d[k] = d0[k]; m[k] = m0[k]; }
                    } else expand(xv[c4], d, m);
                    tmem_st16(lane_addr + q * 128u + (uint32_t)(c4 * 16), d);
                    tmem_st16(lane_addr + q * 128u + (uint32_t)(64 + c4 * 16), m);
                    if (kParts > 1 && (c4 + 1) % kFirst == 0 && (c4 + 1) / kFirst <= kParts - 1) {
                        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_apart(q, (uint32_t)((c4 + 1) / kFirst - 1)));
                    }
                }
                T3_TRACE(3)
                if (kProbe == 6) { __syncwarp(); if (lane == 0) mbar_arrive(bar_tok((q + 1u) % 3u)); }   // pass the token on
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                T3_TRACE(4)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { mbar_arrive(bar_afull(q)); mbar_arrive(bar_gempty(s)); }
                T3_TRACE(5)
            }
            gbase += nr;
            if (q == 0) {
                // epilogue: read D (2 x 16 int32 per row), re-zero it, add to the per-row int64 totals
                mbar_wait(bar_accfull, itc & 1);
                tc_fence_after();
                uint32_t dd[16], dm[16];
                tmem_ld16(lane_addr + kColAcc, dd);
                tmem_ld16(lane_addr + kColAcc + 16, dm);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const uint32_t z[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
                tmem_st16(lane_addr + kColAcc, z); tmem_st16(lane_addr + kColAcc + 16, z);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_accempty);
                const int grow = rt * kTcRows + row;
                if (grow < m_chunk) {
                    unsigned long long *dst = (unsigned long long *)(acc + (int64_t)grow * 32);
#pragma unroll
                    for (int k = 0; k < 15; k++) {
                        atomicAdd(dst + k, (unsigned long long)(long long)(int)dd[k]);
                        atomicAdd(dst + 16 + k, (unsigned long long)(long long)(int)dm[k]);
                    }
                }
            }
            itc++;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kT3ExpWarps) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

#ifdef SPAGXE_PROBES
template <int kProbe>
__global__ void __launch_bounds__(kT3Threads, 1)
k_score_tc3_probe(const __grid_constant__ CUtensorMap bedmap, const int8_t *__restrict__ Bmat, int n_tiles, int tiles_per_split,
                  int n_splits, int n_row_tiles, int m_chunk, long long *__restrict__ acc) {
    score_tc3_body<kProbe>(bedmap, Bmat, n_tiles, tiles_per_split, n_splits, n_row_tiles, m_chunk, acc);
}
#endif

}  // namespace spagxe
#endif  // SPAGXE_PROBES
