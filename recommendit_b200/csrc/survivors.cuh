// Per-warp survivor sink of the pruned exhaustive scans (flat_scan_tc.cu, flat_stream_tc.cu).
//
// A score that beats its query's threshold has to be appended to that query's survivor list in HBM: a slot from
// atomicAdd(count[q], 1), then (score, row).  Done straight from the epilogue, every survivor costs the warp one full atomic round
// trip (≈ 1 us under load) ON the critical path of its accumulator: a pruned round leaves ≈ 3·k survivors per query whatever its
// size, i.e. ≈ 100 k (64 queries) … 6 M (4096 queries) serial round trips per round over the chip — measured ≈ 40 us per round of
// the 64-query scan and about a third of the 4096-query scan.
//
// Here the epilogue only DROPS the survivor into a small shared-memory buffer owned by its warp (no atomics: the warp is converged,
// positions come from a ballot), and the buffer is flushed when it cannot take the next block and at the end of the kernel: 128
// entries per step, four independent atomics per lane in flight, so a round trip is paid per 128 survivors instead of per one.
// The order of a query's survivor list was already arbitrary (exact-score ties inside a round).
#pragma once
#include "common.cuh"
#include "umma.cuh"

// VAL = false: the filter kernels — their accumulator holds an approximate (and threshold-shifted) score that flat_rescore_kernel replaces
// anyway, so only (query, row) are kept: no score has to be re-read from TMEM and the epilogue can release its accumulator as soon as the
// scores are in registers.
template <int CAPW, bool VAL = true>
struct WarpSurvivors {
    float* sc;                 // [CAPW] scores
    int* rw;                   // [CAPW] rows (relative to the round's first row)
    unsigned short* qq;        // [CAPW] queries
    int fill;                  // warp-uniform
    int lane;

    static constexpr int BYTES = CAPW * 10;

    // `mem`: BYTES of shared memory owned by this warp (4-byte aligned)
    __device__ __forceinline__ void init(unsigned char* mem, int lane_) {
        sc = reinterpret_cast<float*>(mem);
        rw = reinterpret_cast<int*>(mem + CAPW * 4);
        qq = reinterpret_cast<unsigned short*>(mem + CAPW * 8);
        fill = 0;
        lane = lane_;
    }

    __device__ __forceinline__ void flush(int* __restrict__ count, float* __restrict__ cand_s, long long stride, int kprev,
                                          int* __restrict__ cand_r, int cap, int* __restrict__ flags) {
        __syncwarp();
        for (int e0 = 0; e0 < fill; e0 += 128) {
            int qv[4], pos[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int e = e0 + i * 32 + lane;
                qv[i] = e < fill ? (int)qq[e] : -1;
                pos[i] = qv[i] >= 0 ? atomicAdd(count + qv[i], 1) : 0;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int e = e0 + i * 32 + lane;
                if (qv[i] >= 0) {
                    if (pos[i] < cap) {
                        if (VAL) cand_s[(long long)qv[i] * stride + kprev + pos[i]] = sc[e];
                        cand_r[(long long)qv[i] * cap + pos[i]] = rw[e];
                    } else {
                        flags[0] = 1;      // survivor list full: the caller redoes the search on the chunked path
                    }
                }
            }
        }
        fill = 0;
        __syncwarp();
    }

    // One 32-row × 32-query block of scores.  m: this lane's (= row's) survivor mask over the block's 32 queries; acc: TMEM address of
    // the block's first column in this warp's lane quadrant (the scores are re-read from there, one column for the 32 lanes, instead
    // of indexing registers dynamically); q0: the block's first query; row: this lane's row.  All 32 lanes must call (converged).
    __device__ __forceinline__ void add_block(uint32_t m, uint32_t acc, int q0, int row, int* __restrict__ count, float* __restrict__ cand_s,
                                              long long stride, int kprev, int* __restrict__ cand_r, int cap, int* __restrict__ flags) {
        uint32_t u = __reduce_or_sync(0xffffffffu, m);
        if (!u) return;
        const int total = (int)__reduce_add_sync(0xffffffffu, (unsigned)__popc(m));
        if (total > CAPW) {
            // more survivors in one block than the buffer holds (thresholds far too low: an adversarial row order) — append directly
            while (u) {
                const int j = __ffs(u) - 1;
                u &= u - 1;
                const float v = VAL ? umma::tmem_ld1(acc + j) : 0.f;
                if ((m >> j) & 1u) {
                    const int pos = atomicAdd(count + q0 + j, 1);
                    if (pos < cap) {
                        if (VAL) cand_s[(long long)(q0 + j) * stride + kprev + pos] = v;
                        cand_r[(long long)(q0 + j) * cap + pos] = row;
                    } else {
                        flags[0] = 1;
                    }
                }
            }
            return;
        }
        if (fill + total > CAPW) flush(count, cand_s, stride, kprev, cand_r, cap, flags);
        const uint32_t lt = (1u << lane) - 1u;
        while (u) {
            const int j = __ffs(u) - 1;
            u &= u - 1;
            const float v = VAL ? umma::tmem_ld1(acc + j) : 0.f;
            const bool mine = (m >> j) & 1u;
            const uint32_t b = __ballot_sync(0xffffffffu, mine);
            if (mine) {
                const int idx = fill + __popc(b & lt);
                if (VAL) sc[idx] = v;
                rw[idx] = row;
                qq[idx] = (unsigned short)(q0 + j);
            }
            fill += __popc(b);
        }
    }
};
