// vl_scan.cuh — single-pass prefix scan (decoupled look-back, warp-shuffle tiles).
//
// Replaces the serial prefix loops of the reference (HostMapBuffer::build, src/core/hosts.rs:178-181) and
// backs the compaction of the haplotype table.  One launch: every CTA takes tiles from an atomic ticket,
// scans 4096 items held as 4 x uint4 per thread, publishes (status | value) in one 64-bit descriptor and
// resolves its exclusive prefix by looking back over predecessor descriptors with one warp.
#pragma once
#include <stdint.h>

namespace vl {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_VEC = 4;
constexpr int SCAN_ROUNDS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_VEC * SCAN_ROUNDS;  // 4096

constexpr uint64_t SCAN_ST_AGG = 1ull << 62, SCAN_ST_INC = 2ull << 62, SCAN_VAL_MASK = (1ull << 62) - 1;

__device__ __forceinline__ uint64_t scan_ld_desc(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void scan_st_desc(uint64_t *p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// work[0] = ticket counter, work[1..] = tile descriptors; all zero before the launch.
// n is read from device memory (*n_ptr + n_add).  out[i] = sum_{j<i} in[j] (EXCL) or sum_{j<=i} (inclusive).
// If total_out != nullptr the grand total is written there.  If EXCL and write_end, out[n] = total too.
template <typename OutT, bool EXCL>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_u32(const uint32_t *__restrict__ in, OutT *__restrict__ out,
                                                            const uint64_t *n_ptr, uint64_t n_add, uint64_t *work,
                                                            uint64_t *total_out, int write_end, const uint32_t *active) {
    __shared__ uint64_t s_warp[SCAN_THREADS / 32];
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_excl;
    if (active && !*active) return;
    const uint64_t n = (n_ptr ? *n_ptr : 0) + n_add;
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint64_t n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    uint64_t *desc = work + 1;
    if (n == 0) {
        if (blockIdx.x == 0 && tid == 0) {
            if (total_out) *total_out = 0;
            if (EXCL && write_end) out[0] = (OutT)0;
        }
        return;
    }
    for (;;) {
        if (tid == 0) s_tile = (uint32_t)atomicAdd((unsigned long long *)work, 1ull);
        __syncthreads();
        const uint64_t tile = s_tile;
        if (tile >= n_tiles) return;
        const uint64_t base = tile * SCAN_TILE;
        // ---- load 4 rounds of uint4 (coalesced 16 B per lane)
        uint32_t v[SCAN_ROUNDS][SCAN_VEC];
        uint64_t tsum[SCAN_ROUNDS];
#pragma unroll
        for (int r = 0; r < SCAN_ROUNDS; r++) {
            uint64_t i0 = base + ((uint64_t)r * SCAN_THREADS + tid) * SCAN_VEC;
            if (i0 + SCAN_VEC <= n) {
                uint4 q = *reinterpret_cast<const uint4 *>(in + i0);
                v[r][0] = q.x; v[r][1] = q.y; v[r][2] = q.z; v[r][3] = q.w;
            } else {
#pragma unroll
                for (int k = 0; k < SCAN_VEC; k++) v[r][k] = (i0 + k < n) ? in[i0 + k] : 0u;
            }
            tsum[r] = (uint64_t)v[r][0] + v[r][1] + v[r][2] + v[r][3];
        }
        // ---- block scan of the per-thread group sums, round by round, carrying the running total
        uint64_t excl_in_tile[SCAN_ROUNDS];
        uint64_t carry = 0;
#pragma unroll
        for (int r = 0; r < SCAN_ROUNDS; r++) {
            uint64_t x = tsum[r];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint64_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
                if (lane >= (uint32_t)d) x += y;
            }
            if (lane == 31) s_warp[wid] = x;
            __syncthreads();
            uint64_t woff = 0, rtotal = 0;
#pragma unroll
            for (int w = 0; w < SCAN_THREADS / 32; w++) {
                uint64_t t = s_warp[w];
                if (w < (int)wid) woff += t;
                rtotal += t;
            }
            excl_in_tile[r] = carry + woff + x - tsum[r];
            carry += rtotal;
            __syncthreads();
        }
        // ---- publish and look back
        if (wid == 0) {
            uint64_t excl = 0;
            if (tile == 0) {
                if (lane == 0) scan_st_desc(&desc[0], SCAN_ST_INC | carry);
            } else {
                if (lane == 0) scan_st_desc(&desc[tile], SCAN_ST_AGG | carry);
                int64_t idx = (int64_t)tile - 1;
                for (;;) {
                    int64_t my = idx - lane;
                    uint64_t d = SCAN_ST_INC;  // virtual tiles before 0: inclusive prefix 0
                    if (my >= 0) {
                        do { d = scan_ld_desc(&desc[my]); } while ((d >> 62) == 0);
                    }
                    uint32_t inc_mask = __ballot_sync(0xFFFFFFFFu, (d >> 62) == 2);
                    uint64_t val = d & SCAN_VAL_MASK;
                    if (inc_mask) {
                        int first = __ffs(inc_mask) - 1;
                        if ((int)lane > first) val = 0;
                    }
#pragma unroll
                    for (int s = 16; s > 0; s >>= 1) val += __shfl_xor_sync(0xFFFFFFFFu, val, s);
                    excl += val;
                    if (inc_mask) break;
                    idx -= 32;
                }
                if (lane == 0) scan_st_desc(&desc[tile], SCAN_ST_INC | ((excl + carry) & SCAN_VAL_MASK));
            }
            if (lane == 0) s_excl = excl;
        }
        __syncthreads();
        const uint64_t tile_excl = s_excl;
        if (tile == n_tiles - 1 && tid == 0) {
            if (total_out) *total_out = tile_excl + carry;
            if (EXCL && write_end) out[n] = (OutT)(tile_excl + carry);
        }
        // ---- write
#pragma unroll
        for (int r = 0; r < SCAN_ROUNDS; r++) {
            uint64_t i0 = base + ((uint64_t)r * SCAN_THREADS + tid) * SCAN_VEC;
            uint64_t run = tile_excl + excl_in_tile[r];
            OutT o[SCAN_VEC];
#pragma unroll
            for (int k = 0; k < SCAN_VEC; k++) {
                if (EXCL) { o[k] = (OutT)run; run += v[r][k]; }
                else { run += v[r][k]; o[k] = (OutT)run; }
            }
            if (i0 + SCAN_VEC <= n) {
                if (sizeof(OutT) == 4) {
                    *reinterpret_cast<uint4 *>(out + i0) = make_uint4((uint32_t)o[0], (uint32_t)o[1], (uint32_t)o[2], (uint32_t)o[3]);
                } else {
                    *reinterpret_cast<ulonglong2 *>(out + i0) = make_ulonglong2((unsigned long long)o[0], (unsigned long long)o[1]);
                    *reinterpret_cast<ulonglong2 *>(out + i0 + 2) = make_ulonglong2((unsigned long long)o[2], (unsigned long long)o[3]);
                }
            } else {
#pragma unroll
                for (int k = 0; k < SCAN_VEC; k++) if (i0 + k < n) out[i0 + k] = o[k];
            }
        }
        __syncthreads();  // s_tile / s_excl reuse
    }
}

}  // namespace vl
