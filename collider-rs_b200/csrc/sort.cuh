// Event-queue materialisation: build 128-bit order keys for every queued event of the last step, sort
// them (stable LSD radix sort, 4-bit digits, only the digits that differ), decode to cb200_event.
// This is the lazy path behind cb200_fetch_events — EventManager's BTreeMap order (events.rs:28-68) with
// the canonical tie-break of SURVEY.md 8c — not part of the per-step hot path.
#pragma once
#include "common.cuh"

namespace cb200 {

// Reiterate events live in the state columns (reit flag + internal end_time); append their keys.
// Both key kernels also reduce the range of the time keys they produce into *range (min, max; null: not wanted): one
// atomic pair per block instead of a separate pass over the keys.
__device__ __forceinline__ void block_range_reduce(unsigned long long lo, unsigned long long hi, BsRange* range) {
    __shared__ unsigned long long s_lo[32], s_hi[32];
    __syncthreads();  // (called twice per kernel: time range, label range)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) {
        s_lo[warp] = lo;
        s_hi[warp] = hi;
    }
    __syncthreads();
    if (warp == 0) {
        lo = lane < nw ? s_lo[lane] : ~0ull;
        hi = lane < nw ? s_hi[lane] : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (lane == 0 && lo <= hi) {
            atomicMin(&range->lo, lo);
            atomicMax(&range->hi, hi);
        }
    }
}
// (bid / nblk: the block's index and the number of blocks that share the loop — a kernel may run both key loops side by side)
__device__ __forceinline__ void keys_solitaire_body(uint32_t bid, uint32_t nblk, uint32_t n, const uint8_t* __restrict__ reit, const double* __restrict__ end,
                                                    const uint32_t* __restrict__ gid, Key128* keys, Counters* ctr, BsRange* range, BsRange* lab_range) {
    unsigned long long lo = ~0ull, hi = 0ull, llo = ~0ull, lhi = 0ull;
    for (uint32_t i = bid * blockDim.x + threadIdx.x; i < n; i += nblk * blockDim.x) {
        if (!reit[i]) continue;
        const unsigned long long pos = warp_agg_claim(&ctr->n_sort_keys);
        const unsigned long long t = time_key(end[i]);
        const unsigned long long label = (unsigned long long)(gid ? gid[i] : i);
        keys[pos] = Key128{label, t};
        lo = min(lo, t);
        hi = max(hi, t);
        llo = min(llo, label);
        lhi = max(lhi, label);
    }
    if (range) block_range_reduce(lo, hi, range);
    if (lab_range) block_range_reduce(llo, lhi, lab_range);
}
__global__ void k_keys_solitaire(uint32_t n, const uint8_t* __restrict__ reit, const double* __restrict__ end, const uint32_t* __restrict__ gid,
                                 Key128* keys, Counters* ctr, BsRange* range, BsRange* lab_range) {
    keys_solitaire_body(blockIdx.x, gridDim.x, n, reit, end, gid, keys, ctr, range, lab_range);
}
__device__ __forceinline__ void keys_pairs_body(uint32_t bid, uint32_t nblk, uint32_t n_ev, const PairEvent* __restrict__ ev, Key128* keys, unsigned long long base,
                                                BsRange* range, BsRange* lab_range) {
    unsigned long long lo = ~0ull, hi = 0ull, llo = ~0ull, lhi = 0ull;
    for (uint32_t i = bid * blockDim.x + threadIdx.x; i < n_ev; i += nblk * blockDim.x) {
        const PairEvent e = ev[i];
        const unsigned long long t = time_key(e.time);
        keys[base + i] = Key128{kPairClass | ((unsigned long long)e.a << 32) | e.b_kind, t};
        lo = min(lo, t);
        hi = max(hi, t);
        llo = min(llo, (unsigned long long)e.a);
        lhi = max(lhi, (unsigned long long)e.a);
    }
    if (range) block_range_reduce(lo, hi, range);
    if (lab_range) block_range_reduce(llo, lhi, lab_range);
}
__global__ void k_keys_pairs(uint32_t n_ev, const PairEvent* __restrict__ ev, Key128* keys, unsigned long long base, BsRange* range, BsRange* lab_range) {
    keys_pairs_body(blockIdx.x, gridDim.x, n_ev, ev, keys, base, range, lab_range);
}
__global__ void k_key_diff(uint32_t n, const Key128* __restrict__ keys, Counters* ctr) {
    const Key128 k0 = keys[0];
    unsigned long long dl = 0, dh = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const Key128 k = keys[i];
        dl |= k.lo ^ k0.lo;
        dh |= k.hi ^ k0.hi;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        dl |= __shfl_xor_sync(0xffffffffu, dl, o);
        dh |= __shfl_xor_sync(0xffffffffu, dh, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (dl) atomicOr(&ctr->diff_lo, dl);
        if (dh) atomicOr(&ctr->diff_hi, dh);
    }
}

constexpr int kRsThreads = 256, kRsItems = 8, kRsTile = kRsThreads * kRsItems;

__device__ __forceinline__ uint32_t key_digit(const Key128& k, int shift) {
    return (uint32_t)((shift < 64 ? (k.lo >> shift) : (k.hi >> (shift - 64))) & 15ull);
}

// hist[d * n_blocks + b] = number of keys of block b with digit d
__global__ void __launch_bounds__(kRsThreads) k_rs_hist(const Key128* __restrict__ in, uint32_t n, int shift, uint32_t* hist, uint32_t n_blocks) {
    __shared__ uint32_t s_h[16];
    if (threadIdx.x < 16) s_h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * kRsTile + threadIdx.x * kRsItems;
#pragma unroll
    for (int k = 0; k < kRsItems; ++k)
        if (base + k < n) atomicAdd(&s_h[key_digit(in[base + k], shift)], 1u);
    __syncthreads();
    if (threadIdx.x < 16) hist[threadIdx.x * n_blocks + blockIdx.x] = s_h[threadIdx.x];
}
// exclusive scan of hist (length m) by one block
__global__ void __launch_bounds__(1024) k_rs_scan(uint32_t* hist, uint32_t m) {
    __shared__ uint32_t s_w[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < m; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < m ? hist[i] : 0;
        uint32_t inc = v;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint32_t ws = s_w[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, ws, o);
                if (lane >= o) ws += t;
            }
            s_w[lane] = ws;
        }
        __syncthreads();
        const uint32_t excl = s_carry + inc - v + (warp ? s_w[warp - 1] : 0);
        if (i < m) hist[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
}
// stable scatter: thread t owns items [t*8, t*8+8) of the tile, so (thread, item) order == input order
__global__ void __launch_bounds__(kRsThreads) k_rs_scatter(const Key128* __restrict__ in, Key128* __restrict__ out, uint32_t n, int shift,
                                                           const uint32_t* __restrict__ hist, uint32_t n_blocks) {
    __shared__ uint32_t s_cnt[16 * kRsThreads];  // [digit][thread]
    __shared__ uint32_t s_w[kRsThreads / 32];
    const uint32_t base = blockIdx.x * kRsTile + threadIdx.x * kRsItems;
    Key128 k[kRsItems];
    uint32_t d[kRsItems];
    uint32_t cnt[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) cnt[q] = 0;
#pragma unroll
    for (int q = 0; q < kRsItems; ++q) {
        d[q] = 16;
        if (base + q < n) {
            k[q] = in[base + q];
            d[q] = key_digit(k[q], shift);
#pragma unroll
            for (int r = 0; r < 16; ++r) cnt[r] += (d[q] == (uint32_t)r);
        }
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) s_cnt[q * kRsThreads + threadIdx.x] = cnt[q];
    __syncthreads();
    // exclusive scan of the 4096-long flattened [digit][thread] array: thread t scans elements [16t, 16t+16)
    uint32_t loc[16];
    uint32_t sum = 0;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        loc[q] = s_cnt[threadIdx.x * 16 + q];
        sum += loc[q];
    }
    uint32_t inc = sum;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t ws = lane < kRsThreads / 32 ? s_w[lane] : 0;
#pragma unroll
        for (int o = 1; o < kRsThreads / 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, ws, o);
            if (lane >= o) ws += t;
        }
        if (lane < kRsThreads / 32) s_w[lane] = ws;
    }
    __syncthreads();
    uint32_t run = inc - sum + (warp ? s_w[warp - 1] : 0);
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        s_cnt[threadIdx.x * 16 + q] = run;
        run += loc[q];
    }
    __syncthreads();
    // position = global base of (digit, block) + rank inside the block among equal digits
    uint32_t seen[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) seen[q] = 0;
#pragma unroll
    for (int q = 0; q < kRsItems; ++q) {
        if (d[q] < 16) {
            const uint32_t dg = d[q];
            uint32_t before = 0;
#pragma unroll
            for (int r = 0; r < 16; ++r)
                if (dg == (uint32_t)r) {
                    before = seen[r];
                    seen[r] += 1;
                }
            const uint32_t in_block = s_cnt[dg * kRsThreads + threadIdx.x] - s_cnt[dg * kRsThreads] + before;
            out[hist[dg * n_blocks + blockIdx.x] + in_block] = k[q];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Fast path: bucket sort.  Event times of a step are spread over [T, end_time], so the order-preserving time key is cut
// into 2^b equal ranges (~12 events each), the keys are counting-sorted into those buckets, and every bucket is finished by
// one warp with a rank sort on the full 128-bit key in shared memory.  Six launches instead of ~70 for the LSD radix sort
// above, which remains the fallback when some bucket is too large (many events at exactly the same time).
constexpr int kBsMaxBucket = 256;  // keys a warp sorts in shared memory (8 warps x 256 x 16 B = 32 KB per block)

// The buckets are cut on ONE 64-bit field of the key: the time key (field 1), or — for a set of keys that all carry the
// same time, e.g. everything that happens exactly at T — the tie word (field 0).  Buckets that come out too large for a warp
// are sorted again on their own range (the host recurses, sort_range in engine.cu).
__global__ void k_bs_minmax(uint32_t n, const Key128* __restrict__ keys, int field, BsRange* r) {
    unsigned long long lo = ~0ull, hi = 0ull;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned long long t = bs_field(keys[i], field);
        lo = min(lo, t);
        hi = max(hi, t);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&r->lo, lo);
        atomicMax(&r->hi, hi);
    }
}
__global__ void k_bs_count(uint32_t n, const Key128* __restrict__ keys, int field, const BsRange* r, int log2_buckets, uint32_t* count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&count[r->lo == r->hi ? 0u : bs_bucket(bs_field(keys[i], field), *r, log2_buckets, field)], 1u);
}
// one block: offsets[b] = exclusive prefix of count (offsets[n_buckets] = total), cursor = a copy; big[0] = number of buckets
// with more than kBsMaxBucket keys, big[1 + 2k], big[2 + 2k] = offset and size of the k-th of them (k < kBsMaxBig)
constexpr int kBsMaxBig = 15;
__global__ void __launch_bounds__(1024) k_bs_scan(const uint32_t* count, uint32_t n_buckets, uint32_t* offsets, uint32_t* cursor, uint32_t* big) {
    __shared__ uint32_t s_w[32];
    __shared__ uint32_t s_carry, s_big;
    if (threadIdx.x == 0) {
        s_carry = 0;
        s_big = 0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < n_buckets; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n_buckets ? count[i] : 0;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint32_t ws = s_w[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, ws, o);
                if (lane >= o) ws += t;
            }
            s_w[lane] = ws;
        }
        __syncthreads();
        const uint32_t excl = s_carry + inc - v + (warp ? s_w[warp - 1] : 0);
        if (i < n_buckets) {
            offsets[i] = excl;
            cursor[i] = excl;
            if (v > (uint32_t)kBsMaxBucket) {
                const uint32_t k = atomicAdd(&s_big, 1u);
                if (k < (uint32_t)kBsMaxBig) {
                    big[1 + 2 * k] = excl;
                    big[2 + 2 * k] = v;
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        offsets[n_buckets] = s_carry;
        big[0] = s_big;
    }
}
__global__ void k_bs_scatter(uint32_t n, const Key128* __restrict__ keys, int field, const BsRange* r, int log2_buckets, uint32_t* cursor, Key128* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Key128 k = keys[i];
    out[atomicAdd(&cursor[r->lo == r->hi ? 0u : bs_bucket(bs_field(k, field), *r, log2_buckets, field)], 1u)] = k;
}
__device__ __forceinline__ bool key128_less(const Key128& a, const Key128& b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); }
// one warp per bucket: rank sort in shared memory; larger buckets are left in `in` (the host sorts them there next)
__device__ __forceinline__ void bs_local_body(const Key128* __restrict__ in, const uint32_t* __restrict__ offsets, uint32_t n_buckets, Key128* out) {
    __shared__ Key128 s_keys[8][kBsMaxBucket];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t b = blockIdx.x * 8 + warp; b < n_buckets; b += gridDim.x * 8) {
        const uint32_t o0 = offsets[b], m = offsets[b + 1] - o0;
        if (m == 0) continue;
        if (m > (uint32_t)kBsMaxBucket) continue;
        if (m == 1) {
            if (lane == 0) out[o0] = in[o0];
            continue;
        }
        for (uint32_t i = lane; i < m; i += 32) s_keys[warp][i] = in[o0 + i];
        __syncwarp();
        for (uint32_t i = lane; i < m; i += 32) {
            const Key128 k = s_keys[warp][i];
            uint32_t rank = 0;
            for (uint32_t j = 0; j < m; ++j) {
                const Key128 q = s_keys[warp][j];
                rank += (key128_less(q, k) || (q.hi == k.hi && q.lo == k.lo && j < i)) ? 1u : 0u;
            }
            out[o0 + rank] = k;
        }
        __syncwarp();
    }
}
__global__ void __launch_bounds__(256) k_bs_local(const Key128* __restrict__ in, const uint32_t* __restrict__ offsets, uint32_t n_buckets, Key128* out) {
    bs_local_body(in, offsets, n_buckets, out);
}

// ---------------------------------------------------------------------------------------------------------
// One-shot variant (no host round trip between the launches): two bucket regions cut in one pass.
//   region A  keys whose time equals the smallest time of the set — in a bulk step that is everything that happens exactly
//             at T (delay 0.0): there the tie word decides, so they are cut linearly on (class, first label)
//   region B  all later keys, cut linearly in time as above
// Both cuts are monotone in the (time, tie) order, so bucket order == key order and the warp rank sort finishes the job.
// The host looks at `big` only after everything (decode and the copy to the host included) has been queued; a bucket that
// came out too large for a warp sends it to the recursive path above.
// The labels are cut over the range [lab.lo, lab.hi] the step's keys actually carry (reduced by the key kernels next to the
// time range): a sharded rank's labels (= ids) cover about 1 / world of the id space, and cutting over the whole id space
// left most of region A's buckets empty and the rest oversized (every fetch then fell back to the recursive sort).
__global__ void k_bs2_count(uint32_t n, const Key128* __restrict__ keys, const BsRange* r /* [2]: time range, label range */, int log2_buckets, uint32_t* count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&count[bs2_bucket(keys[i], r[0], log2_buckets, r[1])], 1u);
}
__global__ void k_bs2_scatter(uint32_t n, const Key128* __restrict__ keys, const BsRange* r, int log2_buckets, uint32_t* cursor, Key128* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Key128 k = keys[i];
    out[atomicAdd(&cursor[bs2_bucket(k, r[0], log2_buckets, r[1])], 1u)] = k;
}

// The bucket scan of the one-shot path, over all SMs (one block walking 2 x 32768 buckets took 81 us — its strided 4-byte
// stores serialise in one SM's L1): k_bs2_scan_local scans kBs2Chunk buckets per block (eight per thread, 32-byte
// accesses) and leaves the block totals; k_bs2_scan_add gives every block the sum of the totals before it (there are at
// most a few thousand), writes offsets and cursors, and reports the buckets that are too large for a warp.
constexpr int kBs2Chunk = 256 * 8;
__device__ __forceinline__ void bs2_scan_local_body(const uint32_t* __restrict__ count, uint32_t n_buckets, uint32_t* __restrict__ local,
                                                    uint32_t* __restrict__ block_total) {
    __shared__ uint32_t s_w[8];
    const uint32_t i0 = blockIdx.x * kBs2Chunk + threadIdx.x * 8u;
    uint32_t v[8], sum = 0;
    if (i0 + 8 <= n_buckets) {
        const uint4 q0 = *reinterpret_cast<const uint4*>(count + i0), q1 = *reinterpret_cast<const uint4*>(count + i0 + 4);
        v[0] = q0.x; v[1] = q0.y; v[2] = q0.z; v[3] = q0.w; v[4] = q1.x; v[5] = q1.y; v[6] = q1.z; v[7] = q1.w;
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = i0 + j < n_buckets ? count[i0 + j] : 0u;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) sum += v[j];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    uint32_t before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        before += w < warp ? s_w[w] : 0u;
        total += s_w[w];
    }
    uint32_t excl = before + inc - sum;
    if (i0 + 8 <= n_buckets) {
        uint4 q0, q1;
        q0.x = excl; excl += v[0]; q0.y = excl; excl += v[1]; q0.z = excl; excl += v[2]; q0.w = excl; excl += v[3];
        q1.x = excl; excl += v[4]; q1.y = excl; excl += v[5]; q1.z = excl; excl += v[6]; q1.w = excl;
        *reinterpret_cast<uint4*>(local + i0) = q0;
        *reinterpret_cast<uint4*>(local + i0 + 4) = q1;
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (i0 + j < n_buckets) local[i0 + j] = excl;
            excl += v[j];
        }
    }
    if (threadIdx.x == 0) block_total[blockIdx.x] = total;
}
__global__ void __launch_bounds__(256) k_bs2_scan_local(const uint32_t* __restrict__ count, uint32_t n_buckets, uint32_t* __restrict__ local,
                                                        uint32_t* __restrict__ block_total) {
    bs2_scan_local_body(count, n_buckets, local, block_total);
}
// count -> offsets in place (offsets[n_buckets] = total); cursor = the block-local scan on entry, a copy of offsets on exit
__device__ __forceinline__ void bs2_scan_add_body(uint32_t* count_offsets, uint32_t n_buckets, uint32_t* cursor, const uint32_t* __restrict__ block_total,
                                                  uint32_t* big /* mapped host memory; big[0] zeroed by the host */) {
    __shared__ uint32_t s_w[8];
    uint32_t part = 0;
    for (uint32_t b = threadIdx.x; b < blockIdx.x; b += blockDim.x) part += block_total[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = part;
    __syncthreads();
    uint32_t prefix = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) prefix += s_w[w];
    const uint32_t i0 = blockIdx.x * kBs2Chunk + threadIdx.x * 8u;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t i = i0 + j;
        if (i < n_buckets) {
            const uint32_t v = count_offsets[i], off = prefix + cursor[i];
            count_offsets[i] = off;
            cursor[i] = off;
            if (v > (uint32_t)kBsMaxBucket) {
                const uint32_t k = atomicAdd(big, 1u);
                if (k < (uint32_t)kBsMaxBig) {
                    big[1 + 2 * k] = off;
                    big[2 + 2 * k] = v;
                }
            }
            if (i == n_buckets - 1) count_offsets[n_buckets] = off + v;
        }
    }
}
__global__ void __launch_bounds__(256) k_bs2_scan_add(uint32_t* count_offsets, uint32_t n_buckets, uint32_t* cursor, const uint32_t* __restrict__ block_total,
                                                      uint32_t* big) {
    bs2_scan_add_body(count_offsets, n_buckets, cursor, block_total, big);
}

// Decode sorted keys to the C-ABI event records (layout of cb200_event: f64 time, u64 a, u64 b, u32 kind, u32 pad).
struct EventOut {
    double time;
    uint64_t a, b;
    uint32_t kind, pad;
};
__global__ void k_decode_events(uint32_t n, const Key128* __restrict__ keys, const uint64_t* __restrict__ ids, EventOut* out, int add_mode) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Key128 k = keys[i];
    EventOut e;
    e.time = key_time(k.hi);
    e.pad = 0;
    if (k.lo & kPairClass) {
        const uint32_t hi = (uint32_t)((k.lo >> 32) & 0x7fffffffu);
        const uint32_t lo = add_mode ? (uint32_t)((k.lo & 0xffffffffu) >> 1) : (uint32_t)(k.lo & 0x7fffffffu);
        e.a = ids ? ids[hi] : hi;
        e.b = ids ? ids[lo] : lo;
        e.kind = (add_mode ? (k.lo & 1ull) : (k.lo & kCollideBit)) ? 1u : 2u;  // CB200_EV_COLLIDE / CB200_EV_SEPARATE
    } else {
        e.a = ids ? ids[(uint32_t)k.lo] : (uint32_t)k.lo;
        e.b = 0;
        e.kind = 0;  // CB200_EV_REITERATE
    }
    out[i] = e;
}

__global__ void k_decode_events16(uint32_t n, const Key128* __restrict__ keys, Event16* out, int add_mode) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = decode_event16(keys[i], add_mode);
}

// ---------------------------------------------------------------------------------------------------------
// The one-shot sort with its sizes in DEVICE memory: the same kernel bodies, launched with grids sized for the buffers'
// capacity, so that the whole sequence is one pair of CUDA graphs replayed every step (engine.cu cb200_fetch_events_begin:
// sixteen driver calls per step became a handful).  k_sortdyn_init reads the step's counters where the step left them (the
// device copy of Counters) and turns them into the sort's parameters.  Only two launches sit on the step's stream — the init
// and ONE kernel that builds both kinds of keys — everything else, the zeroing of the bucket counters included, runs on the
// copy stream beside the next step.
struct SortDyn {
    uint32_t nt, n_pairs, n_buckets;
    int log2_buckets;
    unsigned long long n_sol;
    uint32_t add_mode, pad;
};
__global__ void k_sortdyn_init(Counters* ctr, uint32_t cap_events, uint32_t add_mode, SortDyn* d, BsRange* range /* [2] */, int log2_cap) {
    SortDyn v;
    v.n_pairs = (uint32_t)min(ctr->n_pair_events, (unsigned long long)cap_events);
    v.n_sol = ctr->n_solitaire;
    v.nt = v.n_pairs + (uint32_t)v.n_sol;
    v.log2_buckets = bs2_log2_buckets(v.nt, log2_cap);
    v.n_buckets = 2u << v.log2_buckets;  // region A + region B
    v.add_mode = add_mode;
    v.pad = 0;
    *d = v;
    range[0] = BsRange{~0ull, 0ull};
    range[1] = BsRange{~0ull, 0ull};
    ctr->n_sort_keys = 0ull;
}
__global__ void k_sortdyn_zero(const SortDyn* d, uint32_t* count) {
    const uint32_t m = d->n_buckets + 1u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) count[i] = 0u;
}
// blocks [0, g_rows): the Reiterate keys of the table rows; blocks [g_rows, gridDim.x): the pair-event keys
__global__ void k_keys_all_dyn(const SortDyn* d, uint32_t g_rows, uint32_t n, const uint8_t* __restrict__ reit, const double* __restrict__ end,
                               const uint32_t* __restrict__ gid, const PairEvent* __restrict__ ev, Key128* keys, Counters* ctr, BsRange* range, BsRange* lab_range) {
    if (blockIdx.x < g_rows) {
        if (d->n_sol) keys_solitaire_body(blockIdx.x, g_rows, n, reit, end, gid, keys, ctr, range, lab_range);
    } else {
        keys_pairs_body(blockIdx.x - g_rows, gridDim.x - g_rows, d->n_pairs, ev, keys, d->n_sol, range, lab_range);
    }
}
__global__ void k_bs2_count_dyn(const SortDyn* d, const Key128* __restrict__ keys, const BsRange* r, uint32_t* count) {
    const uint32_t n = d->nt;
    const int l = d->log2_buckets;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) atomicAdd(&count[bs2_bucket(keys[i], r[0], l, r[1])], 1u);
}
__global__ void k_bs2_scatter_dyn(const SortDyn* d, const Key128* __restrict__ keys, const BsRange* r, uint32_t* cursor, Key128* out) {
    const uint32_t n = d->nt;
    const int l = d->log2_buckets;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const Key128 k = keys[i];
        out[atomicAdd(&cursor[bs2_bucket(k, r[0], l, r[1])], 1u)] = k;
    }
}
__global__ void __launch_bounds__(256) k_bs2_scan_local_dyn(const uint32_t* __restrict__ count, const SortDyn* d, uint32_t* __restrict__ local,
                                                            uint32_t* __restrict__ block_total) {
    const uint32_t n_buckets = d->n_buckets;
    if (blockIdx.x * (uint32_t)kBs2Chunk >= n_buckets) return;  // (the grid is sized for the largest table)
    bs2_scan_local_body(count, n_buckets, local, block_total);
}
__global__ void __launch_bounds__(256) k_bs2_scan_add_dyn(uint32_t* count_offsets, const SortDyn* d, uint32_t* cursor, const uint32_t* __restrict__ block_total,
                                                          uint32_t* big) {
    const uint32_t n_buckets = d->n_buckets;
    if (blockIdx.x * (uint32_t)kBs2Chunk >= n_buckets) return;
    bs2_scan_add_body(count_offsets, n_buckets, cursor, block_total, big);
}
__global__ void __launch_bounds__(256) k_bs_local_dyn(const Key128* __restrict__ in, const uint32_t* __restrict__ offsets, const SortDyn* d, Key128* out) {
    bs_local_body(in, offsets, d->n_buckets, out);
}
__global__ void k_decode_events16_dyn(const SortDyn* d, const Key128* __restrict__ keys, Event16* out) {
    const uint32_t n = d->nt;
    const int add_mode = (int)d->add_mode;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = decode_event16(keys[i], add_mode);
}

}  // namespace cb200
