// Event-queue materialisation: build 128-bit order keys for every queued event of the last step, sort
// them (stable LSD radix sort, 4-bit digits, only the digits that differ), decode to cb200_event.
// This is the lazy path behind cb200_fetch_events — EventManager's BTreeMap order (events.rs:28-68) with
// the canonical tie-break of SURVEY.md 8c — not part of the per-step hot path.
#pragma once
#include "common.cuh"

namespace cb200 {

struct alignas(16) Key128 {
    unsigned long long lo;  // tie: class | hi << 32 | kind << 31 | lo     (solitaire: slot)
    unsigned long long hi;  // time_key(time)
};

// Reiterate events live in the state columns (reit flag + internal end_time); append their keys.
__global__ void k_keys_solitaire(uint32_t n, const uint8_t* __restrict__ reit, const double* __restrict__ end, const uint32_t* __restrict__ gid,
                                 Key128* keys, Counters* ctr) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !reit[i]) return;
    const unsigned long long pos = warp_agg_claim(&ctr->n_sort_keys);
    keys[pos] = Key128{(unsigned long long)(gid ? gid[i] : i), time_key(end[i])};
}
__global__ void k_keys_pairs(uint32_t n_ev, const PairEvent* __restrict__ ev, Key128* keys, unsigned long long base) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_ev) return;
    const PairEvent e = ev[i];
    keys[base + i] = Key128{kPairClass | ((unsigned long long)e.a << 32) | e.b_kind, time_key(e.time)};
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

// Decode sorted keys to the C-ABI event records (layout of cb200_event: f64 time, u64 a, u64 b, u32 kind, u32 pad).
struct EventOut {
    double time;
    uint64_t a, b;
    uint32_t kind, pad;
};
__global__ void k_decode_events(uint32_t n, const Key128* __restrict__ keys, const uint64_t* __restrict__ ids, EventOut* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Key128 k = keys[i];
    EventOut e;
    e.time = key_time(k.hi);
    e.pad = 0;
    if (k.lo & kPairClass) {
        const uint32_t hi = (uint32_t)((k.lo >> 32) & 0x7fffffffu), lo = (uint32_t)(k.lo & 0x7fffffffu);
        e.a = ids ? ids[hi] : hi;
        e.b = ids ? ids[lo] : lo;
        e.kind = (k.lo & kCollideBit) ? 1u : 2u;  // CB200_EV_COLLIDE / CB200_EV_SEPARATE
    } else {
        e.a = ids ? ids[(uint32_t)k.lo] : (uint32_t)k.lo;
        e.b = 0;
        e.kind = 0;  // CB200_EV_REITERATE
    }
    out[i] = e;
}

}  // namespace cb200
