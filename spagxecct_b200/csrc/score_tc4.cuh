// score_tc4.cuh -- K1+K2, round-2 kernel: ONE expanded A matrix, read twice by the tensor cores.
//
// The round-1 kernel (score_tc3.cuh, now a probe build) expands every packed row into two int8 matrices (-dosage and
// missing) and is paced by that expansion (LDS + 8 PRMT + 2 STTM.x16 per 16 samples and row; profiles/
// r01_score_tc_analysis.md).  Here a row is expanded ONCE, with the byte code
//        00 (hom A1) -> 0x02     10 (het) -> 0x01     11 (hom A2) -> 0x00     01 (missing) -> 0x80
// and the tensor cores read that matrix twice per K chunk: once as SIGNED bytes and once as UNSIGNED bytes (the
// a-format field of the tcgen05 instruction descriptor).  With class sums T_c of a digit column,
//        signed view    E1 = 2 T00 + T10 - 128 T01
//        unsigned view  E2 = 2 T00 + T10 + 128 T01          =>  T01 = (E2 - E1) / 256,   2 T00 + T10 = E1 + 128 T01
// both exact in integers.  So: half the PRMT / STTM work and half the TMEM per sample, the same number of UTCIMMA.
// The freed TMEM gives every expander group TWO A buffers (ping-pong), so a group expands round k+1 while the tensor
// cores still read round k -- the serial "expand, then wait for the MMA round trip" chain of the round-1 kernel is gone.
//
// A full 4-entry byte LUT needs the selector nibble's mode bit (bit 3) cleared: the packed word is masked with
// 0x77777777 (even samples) / shifted by 2 and masked (odd samples), then PRMT picks LUT[code] with the code in the low
// two bits of each nibble; bit 2 (the neighbour's low bit) selects between two identical LUT registers.
// The B matrix (digits of R and V, a column of ones), its tiling and the sample order inside a 64-sample chunk are
// those of score_tc.cuh; int32 accumulation is exact while samples per split * 128 * 128 < 2^31 (<= 255 tiles of 512).
#pragma once
#include "score_tc.cuh"

namespace spagxe {

constexpr int kT4Groups = 3;                               // product geometry: 3 groups x 2 A buffers
constexpr int kT4Bufs = 2;
constexpr uint32_t kT4ACols = 64;                          // TMEM columns of one A buffer: 256 samples x int8
constexpr int kT4MaxTilesPerSplit = 255;                   // int32 exactness of the unsigned view
constexpr int t4_threads(int groups) { return (4 * groups + 2 + groups) * 32; }   // expanders + 2 producers + MMA warps
constexpr int kT4Threads = t4_threads(kT4Groups);          // 17 warps
constexpr uint32_t kT4IdescS = kTcIdesc;                                   // A = s8, B = s8
constexpr uint32_t kT4IdescU = kTcIdesc & ~(7u << 7);                      // A = u8, B = s8

#ifdef SPAGXE_PROBES
// PROBE 7: CTA 0 records clock64() at the hand-over points of uses [16, 48) of every warp (timing trace; exact sums)
constexpr int kT4TraceRounds = 32, kT4TraceEv = 8, kT4TraceWarps = 32;
__device__ unsigned long long g_t4_trace[kT4TraceWarps * kT4TraceRounds * kT4TraceEv];
#define T4_TRACE(k)                                                                                        \
    if ((PROBE == 7 || PROBE == 9) && blockIdx.x == 0 && lane == 0 && use >= 16 && use < 16 + kT4TraceRounds)              \
        g_t4_trace[((size_t)warp * kT4TraceRounds + (use - 16)) * kT4TraceEv + (k)] = (unsigned long long)clock64();
#define T4_TRACEB(k)                                                                                       \
    if (PROBE == 10 && blockIdx.x == 0 && lane == 0 && use >= 16 && use < 16 + kT4TraceRounds)             \
        g_t4_trace[((size_t)warp * kT4TraceRounds + (use - 16)) * kT4TraceEv + (k)] = (unsigned long long)clock64();
#else
#define T4_TRACE(k)
#define T4_TRACEB(k)
#endif

// NG expander groups (4 warps + 1 MMA warp each), NB A buffers per group; the product instantiation is <3, 2>
// PROBE (probe build only, timing experiments that give WRONG sums): 1 no UTCIMMA, 2 no unsigned-view UTCIMMA,
// 3 no expansion ALU work (raw words stored), 4 no tcgen05.st; 6 (exact sums) next round's packed words read one round ahead
template <int NG, int NB, int PROBE = 0>
__device__ __forceinline__ void
score_tc4_body(const CUtensorMap &bedmap, const int8_t *__restrict__ Bmat, int n_tiles, int tiles_per_split,
               int n_splits, int n_row_tiles, int m_chunk, long long *__restrict__ acc, uint32_t lut) {
    constexpr int kT4Groups = NG;
    constexpr int kT4ExpWarps = 4 * NG;
    constexpr uint32_t kT4ColAcc = NG * NB * kT4ACols;
    static_assert(NG * NB * kT4ACols + 32 <= 512, "TMEM columns");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bbase = sbase + kTcGStages * kTcGStageBytes;
    const uint32_t bars = bbase + kTcBStages * kTcBTileBytes;
    auto bar_gfull = [&](int s) { return bars + 8u * s; };
    auto bar_gempty = [&](int s) { return bars + 8u * (kTcGStages + s); };
    auto bar_bfull = [&](int s) { return bars + 8u * (2 * kTcGStages + s); };
    auto bar_bempty = [&](int s) { return bars + 8u * (2 * kTcGStages + kTcBStages + s); };
    constexpr int kB0 = 2 * kTcGStages + 2 * kTcBStages;
    auto bar_afull = [&](uint32_t q, uint32_t pp) { return bars + 8u * (kB0 + q * NB + pp); };
    auto bar_aempty = [&](uint32_t q, uint32_t pp) { return bars + 8u * (kB0 + NB * kT4Groups + q * NB + pp); };
    const uint32_t bar_accfull = bars + 8u * (kB0 + 2 * NB * kT4Groups), bar_accempty = bars + 8u * (kB0 + 2 * NB * kT4Groups + 1);
    const uint32_t tmem_slot = bars + 8u * (kB0 + 2 * NB * kT4Groups + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kTcGStages; s++) { mbar_init(bar_gfull(s), 1); mbar_init(bar_gempty(s), 8); }   // 2 rounds x 4 warps
        for (int s = 0; s < kTcBStages; s++) { mbar_init(bar_bfull(s), 1); mbar_init(bar_bempty(s), 2); }   // 1 commit per round
        for (uint32_t q = 0; q < (uint32_t)kT4Groups; q++)
            for (uint32_t pp = 0; pp < (uint32_t)NB; pp++) { mbar_init(bar_afull(q, pp), 4); mbar_init(bar_aempty(q, pp), 1); }
        mbar_init(bar_accfull, kT4Groups); mbar_init(bar_accempty, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kT4ExpWarps) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);
    if (warp < 4) {  // zero the accumulators once; afterwards the epilogue re-zeroes them
        const uint32_t z[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        const uint32_t la = tmem_base + ((uint32_t)(warp * 32) << 16);
        tmem_st16(la + kT4ColAcc, z); tmem_st16(la + kT4ColAcc + 16, z);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // Work partition: the (row tile, sample tile) units of the launch form one linear sequence, CTA c takes the c-th of
    // gridDim.x equal contiguous ranges (they differ by at most one unit) and walks it as items (row tile, [t0, t1)) cut at
    // row-tile boundaries and at the int32 exactness limit.  Any partition gives the same bits (int64 atomics), so
    // the split only has to balance the SMs: with whole (row tile x sample split) items dealt round-robin, 640 items on
    // 148 SMs left 100 SMs idle for the last fifth of the kernel.
    const long long n_units = (long long)n_row_tiles * n_tiles;
    const long long u_begin = n_units * blockIdx.x / gridDim.x, u_end = n_units * (blockIdx.x + 1) / gridDim.x;
    (void)tiles_per_split; (void)n_splits;
#define T4_FOR_ITEMS                                                                                             \
    for (long long u_ = u_begin; u_ < u_end;) {                                                                  \
        const int rt = (int)(u_ / n_tiles), t0 = (int)(u_ % n_tiles);                                            \
        const int t1 = (int)min((long long)min(n_tiles, t0 + kT4MaxTilesPerSplit), (long long)t0 + (u_end - u_)); \
        u_ += t1 - t0;                                                                                           \
        (void)rt;
    if (warp == kT4ExpWarps) {
        if (lane == 0) {   // ---- TMA producer: packed genotype tiles (HBM)
            uint32_t f = 0;
            T4_FOR_ITEMS
                for (int t = t0; t < t1; t++, f++) {
                    const int s = f % kTcGStages;
                    mbar_wait(bar_gempty(s), ((f / kTcGStages) & 1) ^ 1);
                    mbar_expect_tx(bar_gfull(s), kTcGStageBytes);
                    tma_load_2d(sbase + s * kTcGStageBytes, &bedmap, t * kTcTileBytes, rt * kTcRows, bar_gfull(s));
                }
            }
        }
    } else if (warp == kT4ExpWarps + 1) {
        if (lane == 0) {   // ---- bulk-copy producer: B tiles (L2)
            uint32_t f = 0;
            T4_FOR_ITEMS
                for (int t = t0; t < t1; t++, f++) {
                    const int s = f % kTcBStages;
                    mbar_wait(bar_bempty(s), ((f / kTcBStages) & 1) ^ 1);
                    mbar_expect_tx(bar_bfull(s), kTcBTileBytes);
                    bulk_load_1d(bbase + s * kTcBTileBytes, Bmat + (int64_t)t * kTcBTileBytes, kTcBTileBytes, bar_bfull(s));
                }
            }
        }
    } else if (warp > kT4ExpWarps + 1) {
        // ---- MMA issuer of group q: 8 K-chunks of a round, each read as signed and as unsigned bytes
        const uint32_t q = (uint32_t)(warp - (kT4ExpWarps + 2));
        const uint32_t ds = tmem_base + kT4ColAcc, du = tmem_base + kT4ColAcc + 16;
        uint32_t gbase = 0, use = 0, itc = 0;
        T4_FOR_ITEMS
            const uint32_t nr = 2u * (uint32_t)max(t1 - t0, 0);
            if (nr == 0) continue;
            bool waited = false;
            for (uint32_t G = gbase + ((q + (uint32_t)NG - gbase % (uint32_t)NG) % (uint32_t)NG); G < gbase + nr; G += (uint32_t)NG, use++) {
                if (!waited) { mbar_wait(bar_accempty, (itc & 1) ^ 1); tc_fence_after(); waited = true; }
                const uint32_t F = G >> 1, hs = G & 1u, sb = F % kTcBStages, pp = use % (uint32_t)NB;
                T4_TRACE(0)
                uint32_t b_ready = 0;
                if (PROBE == 8 || PROBE == 9) b_ready = mbar_test(bar_bfull(sb), (F / kTcBStages) & 1);   // answered while the A buffer is awaited
                else mbar_wait(bar_bfull(sb), (F / kTcBStages) & 1);
                T4_TRACE(1)
                mbar_wait(bar_afull(q, pp), (use / (uint32_t)NB) & 1);
                if ((PROBE == 8 || PROBE == 9) && !b_ready) mbar_wait(bar_bfull(sb), (F / kTcBStages) & 1);
                tc_fence_after();
                T4_TRACE(2)
                if (elect_one()) {
                    const uint32_t abase = tmem_base + (q * (uint32_t)NB + pp) * kT4ACols;
                    const uint64_t bd0 = tc_bdesc(bbase + sb * kTcBTileBytes + hs * (kTcBTileBytes / 2));
#pragma unroll
                    for (int c4 = 0; c4 < 4; c4++) {
#pragma unroll
                        for (int j = 0; j < 2; j++) {
                            const uint64_t bd = bd0 + (uint64_t)(((c4 * 4 + 2 * j) * 256) >> 4);
                            const uint32_t acol = (uint32_t)(c4 * 16 + j * 8);
                            if (PROBE != 1) tc_mma_i8_ts(ds, abase + acol, bd, kT4IdescS, 1u);
                            if (PROBE != 1 && PROBE != 2) tc_mma_i8_ts(du, abase + acol, bd, kT4IdescU, 1u);
                        }
                    }
                    tc_commit(bar_aempty(q, pp));
                    tc_commit(bar_bempty(sb));
                }
                __syncwarp();
                T4_TRACE(3)
            }
            if (elect_one()) tc_commit(bar_accfull);
            __syncwarp();
            gbase += nr; itc++;
        }
    } else {
        // ---- expander warps: group q = warp / 4 fills its two A buffers in turn; thread = SNP row = TMEM lane
        const uint32_t q = (uint32_t)(warp >> 2), quarter = (uint32_t)(warp & 3);
        const int row = (int)quarter * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((quarter * 32u) << 16);
        const uint32_t swz = (uint32_t)(row & 7);
        if constexpr ((PROBE == 0 && NB == 2) || PROBE == 11) {
        // ---- same protocol as the product path, loop state kept incrementally (stage / phase / addresses advance by
        // additions; no division by the ring depth per round)
        static_assert(NB == 2, "streamlined path: ping-pong buffers");
        uint32_t gbase = 0, use = 0, itc = 0;
        uint32_t o0[4];
#pragma unroll
        for (int c4 = 0; c4 < 4; c4++) o0[c4] = (((uint32_t)c4 ^ swz) << 4) + (uint32_t)row * kTcTileBytes;
        const uint32_t a_bar0 = bar_afull(q, 0), abuf0 = lane_addr + (q * (uint32_t)NB) * kT4ACols;
        constexpr uint32_t kAEmptyOff = 8u * NB * NG;
        T4_FOR_ITEMS
            const uint32_t nr = 2u * (uint32_t)max(t1 - t0, 0);
            if (nr == 0) continue;
            uint32_t G = gbase + ((q + (uint32_t)NG - gbase % (uint32_t)NG) % (uint32_t)NG);
            const uint32_t Gend = gbase + nr;
            // ring state of this group's first round of the item
            uint32_t F = G >> 1, hs = G & 1u, s = F % kTcGStages, gph = (F / kTcGStages) & 1;
            uint32_t sdata = sbase + s * kTcGStageBytes, gbar = bar_gfull(0) + 8u * s;
            uint32_t pp = use & 1u, aph = ((use >> 1) & 1u) ^ 1u;
            for (; G < Gend; G += (uint32_t)NG) {
                uint4 xv[4];
                mbar_wait(gbar, gph);
                const uint32_t hx = hs << 6;
#pragma unroll
                for (int c4 = 0; c4 < 4; c4++)
                    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(xv[c4].x), "=r"(xv[c4].y), "=r"(xv[c4].z), "=r"(xv[c4].w)
                                 : "r"(sdata + (o0[c4] ^ hx)));
                const uint32_t abar = a_bar0 + 8u * pp;
                mbar_wait(abar + kAEmptyOff, aph);
                tc_fence_after();
                const uint32_t abuf = abuf0 + pp * kT4ACols;
#pragma unroll
                for (int c4 = 0; c4 < 4; c4++) {
                    const uint32_t x[4] = {xv[c4].x, xv[c4].y, xv[c4].z, xv[c4].w};
                    uint32_t d[16];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint32_t xe = x[k] & 0x77777777u;
                        const uint32_t xo = __umulhi(x[k], 1u << 30) & 0x77777777u;
                        d[4 * k + 0] = prmt(lut, lut, xo);
                        d[4 * k + 1] = prmt(lut, lut, xe);
                        d[4 * k + 2] = prmt(lut, lut, __umulhi(xo, 1u << 16));
                        d[4 * k + 3] = prmt(lut, lut, __umulhi(xe, 1u << 16));
                    }
                    tmem_st16(abuf + (uint32_t)(c4 * 16), d);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { mbar_arrive(abar); mbar_arrive(gbar + 8u * kTcGStages); }
                // advance by NG = 3 rounds: one or two stages
                use++;
                aph ^= pp; pp ^= 1u;                       // the phase flips when the ping-pong wraps
                const uint32_t dF = 1u + hs;
                hs ^= 1u;
                s += dF; sdata += dF * kTcGStageBytes; gbar += 8u * dF;
                if (s >= (uint32_t)kTcGStages) { s -= kTcGStages; sdata -= kTcGStages * kTcGStageBytes; gbar -= 8u * kTcGStages; gph ^= 1u; }
            }
            gbase += nr;
            if (q == 0) {
                mbar_wait(bar_accfull, itc & 1);
                tc_fence_after();
                uint32_t es[16], eu[16];
                tmem_ld16(lane_addr + kT4ColAcc, es);
                tmem_ld16(lane_addr + kT4ColAcc + 16, eu);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const uint32_t z[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
                tmem_st16(lane_addr + kT4ColAcc, z); tmem_st16(lane_addr + kT4ColAcc + 16, z);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_accempty);
                const int grow = rt * kTcRows + row;
                if (grow < m_chunk) {
                    unsigned long long *dst = (unsigned long long *)(acc + (int64_t)grow * 32);
#pragma unroll
                    for (int k = 0; k < 15; k++) {
                        atomicAdd(dst + k, (unsigned long long)(long long)(int)es[k]);
                        atomicAdd(dst + 16 + k, (unsigned long long)(long long)(int)eu[k]);
                    }
                }
            }
            itc++;
        }
        } else if constexpr (PROBE == 8 || PROBE == 9) {
        // ---- software-pipelined rounds: a completed mbarrier still costs ~200 clk to query (the predicate comes back
        // through the MIO queue, behind the tcgen05.st traffic), so the state of the NEXT round's barriers is probed
        // with non-blocking test_wait while this round expands, and the next round's packed words are read while the
        // last tcgen05.st of this round drains.  The blocking wait is only the fallback.
        uint32_t gbase = 0, use = 0, itc = 0;
        auto expand_chunk = [&](const uint4 &xq, uint32_t taddr) {
            const uint32_t x[4] = {xq.x, xq.y, xq.z, xq.w};
            uint32_t d[16];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t xe = x[k] & 0x77777777u;                          // even samples: code in nibble bits 0-1
                const uint32_t xo = __umulhi(x[k], 1u << 30) & 0x77777777u;      // odd samples (x >> 2)
                d[4 * k + 0] = prmt(lut, lut, xo);
                d[4 * k + 1] = prmt(lut, lut, xe);
                d[4 * k + 2] = prmt(lut, lut, __umulhi(xo, 1u << 16));
                d[4 * k + 3] = prmt(lut, lut, __umulhi(xe, 1u << 16));
            }
            tmem_st16(taddr, d);
        };
        auto read_round = [&](uint32_t G, uint4 (&xr)[4]) {
            const uint32_t F = G >> 1, hs = G & 1u, s = F % kTcGStages;
            const uint32_t rowaddr = sbase + s * kTcGStageBytes + (uint32_t)row * kTcTileBytes;
#pragma unroll
            for (int c4 = 0; c4 < 4; c4++) {
                const uint32_t chunk = hs * 4u + (uint32_t)c4;
                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(xr[c4].x), "=r"(xr[c4].y), "=r"(xr[c4].z), "=r"(xr[c4].w)
                             : "r"(rowaddr + ((chunk ^ swz) << 4)));
            }
        };
        T4_FOR_ITEMS
            const uint32_t nr = 2u * (uint32_t)max(t1 - t0, 0);
            if (nr == 0) continue;
            uint32_t G = gbase + ((q + (uint32_t)NG - gbase % (uint32_t)NG) % (uint32_t)NG);
            const uint32_t Gend = gbase + nr;
            if (G < Gend) {
                uint4 xv[4];
                mbar_wait(bar_gfull((G >> 1) % kTcGStages), ((G >> 1) / kTcGStages) & 1);
                read_round(G, xv);
                uint32_t a_ready = 0;
                for (;;) {
                    const uint32_t s = (G >> 1) % kTcGStages, pp = use % (uint32_t)NB;
                    const uint32_t Gn = G + (uint32_t)NG, Fn = Gn >> 1, sn = Fn % kTcGStages, phn = (Fn / kTcGStages) & 1;
                    const uint32_t ppn = (use + 1) % (uint32_t)NB, aphn = (((use + 1) / (uint32_t)NB) & 1) ^ 1;
                    const bool have_next = Gn < Gend;
                    T4_TRACE(0)
                    if (!a_ready) mbar_wait(bar_aempty(q, pp), ((use / (uint32_t)NB) & 1) ^ 1);
                    tc_fence_after();
                    T4_TRACE(1)
                    const uint32_t g_ready = have_next ? mbar_test(bar_gfull(sn), phn) : 0u;   // consumed after the expansion
                    const uint32_t abuf = lane_addr + (q * (uint32_t)NB + pp) * kT4ACols;
#pragma unroll
                    for (int c4 = 0; c4 < 4; c4++) expand_chunk(xv[c4], abuf + (uint32_t)(c4 * 16));
                    T4_TRACE(2)
                    uint32_t a_next = 0;
                    if (have_next) {
                        a_next = mbar_test(bar_aempty(q, ppn), aphn);
                        if (!g_ready) mbar_wait(bar_gfull(sn), phn);
                        read_round(Gn, xv);                 // xv is dead: its words went through the expansion
                    }
                    T4_TRACE(3)
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    T4_TRACE(4)
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) { mbar_arrive(bar_afull(q, pp)); mbar_arrive(bar_gempty(s)); }
                    T4_TRACE(5)
#ifdef SPAGXE_PROBES
                    if (PROBE == 9 && blockIdx.x == 0 && lane == 0 && use >= 16 && use < 16 + kT4TraceRounds)
                        g_t4_trace[((size_t)warp * kT4TraceRounds + (use - 16)) * kT4TraceEv + 6] = 1000ull + a_ready * 2 + (g_ready ? 1 : 0);
#endif
                    use++;
                    if (!have_next) break;
                    G = Gn; a_ready = a_next;
                }
            }
            gbase += nr;
            if (q == 0) {
                mbar_wait(bar_accfull, itc & 1);
                tc_fence_after();
                uint32_t es[16], eu[16];
                tmem_ld16(lane_addr + kT4ColAcc, es);
                tmem_ld16(lane_addr + kT4ColAcc + 16, eu);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const uint32_t z[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
                tmem_st16(lane_addr + kT4ColAcc, z); tmem_st16(lane_addr + kT4ColAcc + 16, z);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_accempty);
                const int grow = rt * kTcRows + row;
                if (grow < m_chunk) {
                    unsigned long long *dst = (unsigned long long *)(acc + (int64_t)grow * 32);
#pragma unroll
                    for (int k = 0; k < 15; k++) {
                        atomicAdd(dst + k, (unsigned long long)(long long)(int)es[k]);
                        atomicAdd(dst + 16 + k, (unsigned long long)(long long)(int)eu[k]);
                    }
                }
            }
            itc++;
        }
        } else {
        uint32_t gbase = 0, use = 0, itc = 0;
        bool have_next = false;
        uint4 xn[4];
        T4_FOR_ITEMS
            const uint32_t nr = 2u * (uint32_t)max(t1 - t0, 0);
            if (nr == 0) continue;
            for (uint32_t G = gbase + ((q + (uint32_t)NG - gbase % (uint32_t)NG) % (uint32_t)NG); G < gbase + nr; G += (uint32_t)NG, use++) {
                const uint32_t F = G >> 1, hs = G & 1u, s = F % kTcGStages, pp = use % (uint32_t)NB;
                uint4 xv[4];   // the four 16-byte chunks of this row's half stage (loads in flight together)
                T4_TRACE(0)
                if (PROBE == 6 && have_next) {
#pragma unroll
                    for (int c4 = 0; c4 < 4; c4++) xv[c4] = xn[c4];
                } else {
                    mbar_wait(bar_gfull(s), (F / kTcGStages) & 1);
                    T4_TRACE(1)
                    const uint32_t rowaddr = sbase + s * kTcGStageBytes + (uint32_t)row * kTcTileBytes;
#pragma unroll
                    for (int c4 = 0; c4 < 4; c4++) {
                        const uint32_t chunk = hs * 4u + (uint32_t)c4;
                        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                                     : "=r"(xv[c4].x), "=r"(xv[c4].y), "=r"(xv[c4].z), "=r"(xv[c4].w)
                                     : "r"(rowaddr + ((chunk ^ swz) << 4)));
                    }
                }
                mbar_wait(bar_aempty(q, pp), ((use / (uint32_t)NB) & 1) ^ 1);
                tc_fence_after();
                T4_TRACE(2)
                T4_TRACEB(0)
                const uint32_t abuf = lane_addr + (q * (uint32_t)NB + pp) * kT4ACols;
#pragma unroll
                for (int c4 = 0; c4 < 4; c4++) {
                    const uint32_t x[4] = {xv[c4].x, xv[c4].y, xv[c4].z, xv[c4].w};
                    uint32_t d[16];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        if (PROBE == 3) { d[4 * k + 0] = x[k]; d[4 * k + 1] = x[k] + 1u; d[4 * k + 2] = x[k] + 2u; d[4 * k + 3] = x[k] + 3u; continue; }
                        // right shifts through IMAD.HI (FMA pipe): the ALU pipe is left to LOP3 / PRMT
                        const uint32_t xe = x[k] & 0x77777777u;                          // even samples: code in nibble bits 0-1
                        const uint32_t xo = __umulhi(x[k], 1u << 30) & 0x77777777u;      // odd samples (x >> 2)
                        d[4 * k + 0] = prmt(lut, lut, xo);
                        d[4 * k + 1] = prmt(lut, lut, xe);
                        d[4 * k + 2] = prmt(lut, lut, __umulhi(xo, 1u << 16));
                        d[4 * k + 3] = prmt(lut, lut, __umulhi(xe, 1u << 16));
                    }
                    if (PROBE != 4) { tmem_st16(abuf + (uint32_t)(c4 * 16), d); T4_TRACEB(1 + c4) }
                    else {
                        uint32_t v = 0;
#pragma unroll
                        for (int k = 0; k < 16; k++) v ^= d[k];
                        if (v == 0x12345679u) tmem_st16(abuf + (uint32_t)(c4 * 16), d);   // keeps the expansion alive
                    }
                }
                if (PROBE == 6) {   // the next round's packed words: their latency hides behind the store wait and the hand-over
                    have_next = (G + (uint32_t)NG < gbase + nr);
                    if (have_next) {
                        const uint32_t Gn = G + (uint32_t)NG, Fn = Gn >> 1, hn = Gn & 1u, sn = Fn % kTcGStages;
                        mbar_wait(bar_gfull(sn), (Fn / kTcGStages) & 1);
                        const uint32_t rowaddr = sbase + sn * kTcGStageBytes + (uint32_t)row * kTcTileBytes;
#pragma unroll
                        for (int c4 = 0; c4 < 4; c4++) {
                            const uint32_t chunk = hn * 4u + (uint32_t)c4;
                            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                                         : "=r"(xn[c4].x), "=r"(xn[c4].y), "=r"(xn[c4].z), "=r"(xn[c4].w)
                                         : "r"(rowaddr + ((chunk ^ swz) << 4)));
                        }
                    }
                }
                T4_TRACE(3)
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                T4_TRACE(4)
                T4_TRACEB(5)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { mbar_arrive(bar_afull(q, pp)); mbar_arrive(bar_gempty(s)); }
                T4_TRACE(5)
            }
            gbase += nr;
            if (q == 0) {
                // epilogue: read the two accumulator blocks (2 x 16 int32 per row), re-zero them, add to the int64 totals
                mbar_wait(bar_accfull, itc & 1);
                tc_fence_after();
                uint32_t es[16], eu[16];
                tmem_ld16(lane_addr + kT4ColAcc, es);
                tmem_ld16(lane_addr + kT4ColAcc + 16, eu);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const uint32_t z[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
                tmem_st16(lane_addr + kT4ColAcc, z); tmem_st16(lane_addr + kT4ColAcc + 16, z);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_accempty);
                const int grow = rt * kTcRows + row;
                if (grow < m_chunk) {
                    unsigned long long *dst = (unsigned long long *)(acc + (int64_t)grow * 32);
#pragma unroll
                    for (int k = 0; k < 15; k++) {
                        atomicAdd(dst + k, (unsigned long long)(long long)(int)es[k]);
                        atomicAdd(dst + 16 + k, (unsigned long long)(long long)(int)eu[k]);
                    }
                }
            }
            itc++;
        }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kT4ExpWarps) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

#undef T4_FOR_ITEMS

// K1+K2 product kernel (the only tensor-core score kernel in the product .so)
__global__ void __launch_bounds__(kT4Threads, 1)
k_score_unpack_mma(const __grid_constant__ CUtensorMap bedmap, const int8_t *__restrict__ Bmat, int n_tiles, int tiles_per_split,
                   int n_splits, int n_row_tiles, int m_chunk, long long *__restrict__ acc, uint32_t lut) {
    score_tc4_body<kT4Groups, kT4Bufs>(bedmap, Bmat, n_tiles, tiles_per_split, n_splits, n_row_tiles, m_chunk, acc, lut);
}
#ifdef SPAGXE_PROBES
template <int NG, int NB, int PROBE = 0>
__global__ void __launch_bounds__(t4_threads(NG), 1)
k_score_tc4_probe(const __grid_constant__ CUtensorMap bedmap, const int8_t *__restrict__ Bmat, int n_tiles, int tiles_per_split,
                  int n_splits, int n_row_tiles, int m_chunk, long long *__restrict__ acc, uint32_t lut) {
    score_tc4_body<NG, NB, PROBE>(bedmap, Bmat, n_tiles, tiles_per_split, n_splits, n_row_tiles, m_chunk, acc, lut);
}
#endif

// digit recombination for the signed / unsigned pair: per digit column k, missing sum M_k = (E2_k - E1_k) / 256 and
// dosage sum D_k = E1_k + 128 M_k, both exact; then value = sum_k 256^k (.)_k as in tc_combine
__global__ void k_score_tc4_reduce(const long long *__restrict__ acc, int m_chunk, const BmatInfo *__restrict__ info, SnpStats st) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m_chunk) return;
    const long long *a = acc + (int64_t)j * 32;
    long long d[16], m[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const long long e1 = a[k], e2 = a[16 + k];
        m[k] = (e2 - e1) >> 8;            // exact: the difference is 256 * T01
        d[k] = e1 + 128 * m[k];
    }
    st.D0[j] = tc_combine(d, info->sR);
    st.D1[j] = tc_combine(d + 8, info->sV);
    st.asum[j] = (int)d[7];
    st.T0[j] = tc_combine(m, info->sR);
    st.T1[j] = tc_combine(m + 8, info->sV);
    st.nmiss[j] = (int)m[7];
}

// second pass of G.model Dom / Rec (LUT 00 -> 1, missing -> 0x80, others 0): the dosage part is the hom-A1 class sum
__global__ void k_score_tc4_reduce_hom(const long long *__restrict__ acc, int m_chunk, const BmatInfo *__restrict__ info, SnpStats st) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m_chunk) return;
    const long long *a = acc + (int64_t)j * 32;
    long long d[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const long long e1 = a[k], e2 = a[16 + k];
        d[k] = e1 + 128 * ((e2 - e1) >> 8);
    }
    st.H0[j] = tc_combine(d, info->sR);
    st.H1[j] = tc_combine(d + 8, info->sV);
    st.nhom[j] = (int)d[7];
}

}  // namespace spagxe
