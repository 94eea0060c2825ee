// Shared device-side types and helpers of the B200 engine (sm_100a only).
//
// Vocabulary follows the reference (SergiusIW/collider-rs 0.3.1): hitbox, DurHitbox, IndexRect, grid
// cell, HbEvent.  A *slot* is one element of the device cell table: cell (x, y) folded modulo the grid
// period (2^lw x 2^lh).  Aliasing between cells that share a slot costs time, never correctness (see
// pair ownership in kernels.cuh).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "narrowphase.cuh"
#include "sort_cut.cuh"

namespace cb200 {

// ---- error flags (mirrors include/collider_b200.h CB200_F_*) ----
constexpr uint32_t kFInvalidTime = 1u << 0, kFEndTime = 1u << 1, kFCircleResize = 1u << 2, kFTooSmall = 1u << 3,
                   kFEmptyRect = 1u << 4, kFRectSpan = 1u << 5, kFNan = 1u << 6, kFHighTime = 1u << 7,
                   kFIdNotFound = 1u << 8, kFPeerLost = 1u << 9;

constexpr uint32_t kGroupNone = 0xFFFFFFFFu;

// The DurHitbox a hitbox presents to the pair stage at the step time (core/dur_hitbox/mod.rs:27-70)
// plus its IndexRect (index_rect.rs:17-31) and profile bits: 96 B = three 32-B sectors.  Sector 0 is
// all the binning needs; sectors 1-2 are fetched only for pairs that are solved.  A record is
// self-contained (it carries the hitbox's global id rank), so it is also the unit of the multi-GPU halo.
struct alignas(32) HbRec {
    int32_t sx, sy;  // IndexRect start
    uint32_t span;   // (ex - sx) | (ey - sy) << 16     (spans are limited to 65535 cells)
    uint32_t gid;    // rank of the HbId in ascending-id order (one device: the table slot; sharded: the id itself)
    double dur;      // internal end_time - T
    uint32_t kgb;    // bit0 kind (0 circle, 1 rect), bit1 binned (has a group and a valid rect), bits 8..15 group
    uint32_t mask;   // interact_groups as a bit mask
    double px, py, w, h;
    double vx, vy, rw, rh;
};
static_assert(sizeof(HbRec) == 96, "HbRec must be three sectors");

struct RecHead {  // sector 0 of HbRec, decoded
    int32_t sx, sy, ex, ey;
    uint32_t gid;
    double dur;
    uint32_t kgb, mask;
};

__device__ __forceinline__ RecHead load_head(const HbRec* __restrict__ rec, uint32_t i) {
    const int4* p = reinterpret_cast<const int4*>(rec + i);
    const int4 a = __ldg(p), b = __ldg(p + 1);
    RecHead h;
    h.sx = a.x; h.sy = a.y;
    h.ex = a.x + (int32_t)((uint32_t)a.z & 0xffffu);
    h.ey = a.y + (int32_t)((uint32_t)a.z >> 16);
    h.gid = (uint32_t)a.w;
    h.dur = __hiloint2double(b.y, b.x);
    h.kgb = (uint32_t)b.z;
    h.mask = (uint32_t)b.w;
    return h;
}

__device__ __forceinline__ DurHb load_body(const HbRec* __restrict__ rec, uint32_t i, const RecHead& h) {
    const double2* p = reinterpret_cast<const double2*>(rec + i);
    const double2 a = __ldg(p + 2), b = __ldg(p + 3), c = __ldg(p + 4), d = __ldg(p + 5);
    DurHb r;
    r.px = a.x; r.py = a.y; r.w = b.x; r.h = b.y;
    r.vx = c.x; r.vy = c.y; r.rw = d.x; r.rh = d.y;
    r.dur = h.dur;
    r.kind = (int)(h.kgb & 1u);
    return r;
}

// L2 residency hints.  The 96-byte records are written by stage A and gathered by the solve stage four kernels later, with
// ~110 MB of streaming traffic (slot table, bin keys, entries, work items) in between on a 126 MB L2: records are stored
// and loaded evict_last so that the gathers mostly hit L2 instead of DRAM.  CB200_L2_HINTS=0 at build time turns this off.
#ifndef CB200_L2_HINTS
#define CB200_L2_HINTS 0  /* measured: no effect on the solve stage (bench_p16 vs bench_p16_nohint) */
#endif
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void st_int4_keep(int4* ptr, const int4& v, uint64_t policy) {
#if CB200_L2_HINTS
    asm volatile("st.global.L2::cache_hint.v4.s32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(ptr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(policy) : "memory");
#else
    *ptr = v;
#endif
}
__device__ __forceinline__ int4 ld_int4_keep(const int4* ptr, uint64_t policy) {
#if CB200_L2_HINTS
    int4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.s32 {%0, %1, %2, %3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr), "l"(policy));
    return v;
#else
    return __ldg(ptr);
#endif
}
// Record loads of the solve stage (both halves of a record, evict_last).
__device__ __forceinline__ void load_rec_keep(const HbRec* __restrict__ rec, uint32_t i, uint64_t policy, RecHead* h, DurHb* d) {
    const int4* p = reinterpret_cast<const int4*>(rec + i);
    const int4 a = ld_int4_keep(p, policy), b = ld_int4_keep(p + 1, policy), c = ld_int4_keep(p + 2, policy), e = ld_int4_keep(p + 3, policy),
               f = ld_int4_keep(p + 4, policy), g = ld_int4_keep(p + 5, policy);
    h->sx = a.x; h->sy = a.y;
    h->ex = a.x + (int32_t)((uint32_t)a.z & 0xffffu);
    h->ey = a.y + (int32_t)((uint32_t)a.z >> 16);
    h->gid = (uint32_t)a.w;
    h->dur = __hiloint2double(b.y, b.x);
    h->kgb = (uint32_t)b.z;
    h->mask = (uint32_t)b.w;
    d->px = __hiloint2double(c.y, c.x); d->py = __hiloint2double(c.w, c.z);
    d->w = __hiloint2double(e.y, e.x); d->h = __hiloint2double(e.w, e.z);
    d->vx = __hiloint2double(f.y, f.x); d->vy = __hiloint2double(f.w, f.z);
    d->rw = __hiloint2double(g.y, g.x); d->rh = __hiloint2double(g.w, g.z);
    d->dur = h->dur;
    d->kind = (int)(h->kgb & 1u);
}

// The same record through three 256-bit loads (LDG.E.256 on sm_100a): one request per 32-B sector instead of two.
__device__ __forceinline__ void ld256_nc(const void* p, unsigned long long (&v)[4]) {
    asm("ld.global.nc.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]) : "l"(p));
}
__device__ __forceinline__ void load_rec_wide(const HbRec* __restrict__ rec, uint32_t i, RecHead* h, DurHb* d) {
    const char* p = reinterpret_cast<const char*>(rec + i);
    unsigned long long s0[4], s1[4], s2[4];
    ld256_nc(p, s0);
    ld256_nc(p + 32, s1);
    ld256_nc(p + 64, s2);
    h->sx = (int32_t)(uint32_t)s0[0]; h->sy = (int32_t)(uint32_t)(s0[0] >> 32);
    const uint32_t span = (uint32_t)s0[1];
    h->ex = h->sx + (int32_t)(span & 0xffffu);
    h->ey = h->sy + (int32_t)(span >> 16);
    h->gid = (uint32_t)(s0[1] >> 32);
    h->dur = __longlong_as_double((long long)s0[2]);
    h->kgb = (uint32_t)s0[3];
    h->mask = (uint32_t)(s0[3] >> 32);
    d->px = __longlong_as_double((long long)s1[0]); d->py = __longlong_as_double((long long)s1[1]);
    d->w = __longlong_as_double((long long)s1[2]); d->h = __longlong_as_double((long long)s1[3]);
    d->vx = __longlong_as_double((long long)s2[0]); d->vy = __longlong_as_double((long long)s2[1]);
    d->rw = __longlong_as_double((long long)s2[2]); d->rh = __longlong_as_double((long long)s2[3]);
    d->dur = h->dur;
    d->kind = (int)(h->kgb & 1u);
}

// Coherent (plain ld.global) variants for kernels that read records written earlier in the same launch.
__device__ __forceinline__ RecHead load_head_nc(const HbRec* rec, uint32_t i) {
    const int4* p = reinterpret_cast<const int4*>(rec + i);
    const int4 a = p[0], b = p[1];
    RecHead h;
    h.sx = a.x; h.sy = a.y;
    h.ex = a.x + (int32_t)((uint32_t)a.z & 0xffffu);
    h.ey = a.y + (int32_t)((uint32_t)a.z >> 16);
    h.gid = (uint32_t)a.w;
    h.dur = __hiloint2double(b.y, b.x);
    h.kgb = (uint32_t)b.z;
    h.mask = (uint32_t)b.w;
    return h;
}
__device__ __forceinline__ DurHb load_body_nc(const HbRec* rec, uint32_t i, const RecHead& h) {
    const double2* p = reinterpret_cast<const double2*>(rec + i);
    const double2 a = p[2], b = p[3], c = p[4], d = p[5];
    DurHb r;
    r.px = a.x; r.py = a.y; r.w = b.x; r.h = b.y;
    r.vx = c.x; r.vy = c.y; r.rw = d.x; r.rh = d.y;
    r.dur = h.dur;
    r.kind = (int)(h.kgb & 1u);
    return r;
}

// One pair event queued by a step: time, creator (higher slot), partner (lower slot) | kind bit.
struct alignas(16) PairEvent {
    double time;
    uint32_t a;
    uint32_t b_kind;
};

// Total order of the event queue (events.rs:28-68 + SURVEY.md 8c canonical tie-break):
// t = order-preserving bits of the time; tie = class | hi << 32 | kind << 31 | lo  (solitaire: slot).
struct MinKey {
    uint64_t t, tie;
};
__device__ __forceinline__ bool key_less(const MinKey& a, const MinKey& b) { return a.t < b.t || (a.t == b.t && a.tie < b.tie); }
constexpr uint64_t kKeyMax = 0xFFFFFFFFFFFFFFFFull;

__device__ __forceinline__ MinKey warp_min(MinKey k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        MinKey other;
        other.t = __shfl_xor_sync(0xffffffffu, k.t, o);
        other.tie = __shfl_xor_sync(0xffffffffu, k.tie, o);
        if (key_less(other, k)) k = other;
    }
    return k;
}

// Block-wide min of a MinKey; the result is valid in thread 0.  blockDim.x must be a multiple of 32, <= 1024.
__device__ __forceinline__ MinKey block_min(MinKey k) {
    __shared__ MinKey s_part[32];
    k = warp_min(k);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_part[warp] = k;
    __syncthreads();
    if (warp == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        k = lane < nw ? s_part[lane] : MinKey{kKeyMax, kKeyMax};
        k = warp_min(k);
    }
    return k;
}

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Block-wide sum; valid in thread 0.
__device__ __forceinline__ unsigned long long block_sum(unsigned long long v) {
    __shared__ unsigned long long s_sum[32];
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_sum[warp] = v;
    __syncthreads();
    if (warp == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        v = lane < nw ? s_sum[lane] : 0ull;
        v = warp_sum(v);
    }
    return v;
}

// Device-resident step counters (zeroed at the start of every step).
struct Counters {
    unsigned long long n_entries;      // E, written by the scan
    unsigned long long n_cells;        // non-empty slots
    unsigned long long n_collide;      // collide_time evaluations = work items that passed the candidate test
    unsigned int work_overflow;        // step result: item count of the fullest work sub-list (vs the segment capacity)
    unsigned int pad0;
    unsigned long long n_work;         // step result: (record, record, slot) work items enumerated
    unsigned long long n_separate;     // separate_time evaluations
    unsigned long long n_pair_events;  // pair events with time < 1e50 (may exceed the buffer capacity -> overflow)
    unsigned long long n_solitaire;    // Reiterate events
    unsigned long long err_slot;       // lowest slot that reported an error
    unsigned int err_flags;
    unsigned int scan_ticket;
    unsigned int overflow_entries;
    unsigned int done_blocks;          // solve stage: blocks finished (last-block finalize)
    unsigned int n_local;              // sharded: records this rank keeps for its own strip
    unsigned int overflow_send;        // sharded: a send buffer overflowed
    unsigned long long diff_hi, diff_lo;  // sort: OR of key differences
    unsigned long long n_sort_keys;
    unsigned long long n_tile_pairs;   // tile path: (entry, predecessor) pairs walked in shared memory (the step result's work items)
    unsigned long long n_first_events; // pair events when the solve stage finished (the window drain appends its follow-ups behind them)
    unsigned long long n_surv;         // tile path: survivors of the swept-bounds filter queued for k_tile_finish (may exceed the buffer: overflow)
    unsigned int tile_fail;            // tile path: a visitor slab or a single cell exceeded its capacity -> the host repeats the step on the global path
    unsigned int pad1;
};

__device__ __forceinline__ void report_error(Counters* ctr, uint32_t flags, uint64_t slot) {
    atomicOr(&ctr->err_flags, flags);
    atomicMin(&ctr->err_slot, (unsigned long long)slot);
}

// Block-wide exclusive scan of one uint32 per thread (blockDim.x == 256 or any multiple of 32 <= 1024);
// returns the exclusive prefix, *total receives the block sum in every thread.
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t s_scan[33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();  // protects s_scan against the previous call's readers
    if (lane == 31) s_scan[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        uint32_t ws = lane < nw ? s_scan[lane] : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, ws, o);
            if (lane >= o) ws += t;
        }
        if (lane < nw) s_scan[lane] = ws;  // inclusive over warps
        if (lane == nw - 1) s_scan[32] = ws;
    }
    __syncthreads();
    *total = s_scan[32];
    return inc - v + (warp ? s_scan[warp - 1] : 0u);
}

// Grid fold: cell (x, y) -> slot.  Cells are folded modulo the grid period and laid out in 8 x 8-cell tiles
// (slot bits: y_hi | x_hi | y_lo3 | x_lo3), so that the hitboxes of one neighbourhood touch a few consecutive cache
// lines of the slot table and of the slot-sorted entry array (the hitbox table itself is kept in the same order,
// see k_resort_*).
struct GridGeom {
    int lw, lh;
    __host__ __device__ __forceinline__ int tx() const { return lw < 3 ? lw : 3; }
    __host__ __device__ __forceinline__ int ty() const { return lh < 3 ? lh : 3; }
    __host__ __device__ __forceinline__ uint32_t slot(int32_t x, int32_t y) const {
        const uint32_t ux = (uint32_t)x & ((1u << lw) - 1u), uy = (uint32_t)y & ((1u << lh) - 1u);
        const int a = tx(), b = ty();
        return (ux & ((1u << a) - 1u)) | ((uy & ((1u << b) - 1u)) << a) | ((ux >> a) << (a + b)) | ((uy >> b) << (lw + b));
    }
    // inverse: the folded (x, y) of a slot
    __host__ __device__ __forceinline__ void unslot(uint32_t s, uint32_t* ux, uint32_t* uy) const {
        const int a = tx(), b = ty();
        const uint32_t xl = s & ((1u << a) - 1u), yl = (s >> a) & ((1u << b) - 1u);
        const uint32_t xh = (s >> (a + b)) & ((1u << (lw - a)) - 1u), yh = s >> (lw + b);
        *ux = xl | (xh << a);
        *uy = yl | (yh << b);
    }
    __host__ __device__ __forceinline__ uint32_t n_slots() const { return 1u << (lw + lh); }
};

// Warp-aggregated append: the lanes that call this together claim consecutive positions with one atomic.
__device__ __forceinline__ unsigned long long warp_agg_claim(unsigned long long* counter) {
    const unsigned m = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(m));
    base = __shfl_sync(m, base, leader);
    return base + (unsigned long long)__popc(m & ((1u << lane) - 1u));
}
__device__ __forceinline__ unsigned int warp_agg_claim32(unsigned int* counter) {
    const unsigned m = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    unsigned int base = 0;
    if (lane == leader) base = atomicAdd(counter, (unsigned int)__popc(m));
    base = __shfl_sync(m, base, leader);
    return base + (unsigned int)__popc(m & ((1u << lane) - 1u));
}

}  // namespace cb200
