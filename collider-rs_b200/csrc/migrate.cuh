// Owner migration of a sharded table ("hitboxes migrate owners when `start` crosses a strip edge").
//
// A rank is the RESIDENT of the rows in its table: it advances them, writes their state and emits their Reiterate events.
// Which rank solves a pair is decided by the pair's owner cell, not by residence, so any partition of the rows is correct —
// but a row that has drifted into another strip is pure overhead where it lives (a halo record every step, no local cell).
// Between two steps the host moves such rows to the rank whose strip holds their centre column:
//   k_mig_mark     lists the rows whose centre column left [col_lo - margin, col_hi + margin) with their destination strip
//   k_mig_pack     copies the listed rows (every column) into a staging block the host downloads; clears their vmap entries
//   k_mig_fill     fills the holes among the first n - m rows with the surviving rows of the tail
//   k_mig_ord_out  ord = rank of the row's id among the residents (the caller's array order): minus the movers below it
//   k_mig_ord_in   plus the arrivals below it; counts, for every arrival, the residents below it
// The arrivals are then appended by plain copies (engine.cu cb200_shard_add_rows).  Row order only affects speed.
#pragma once
#include "kernels.cuh"

namespace cb200 {

struct MigCols {
    double* col[11];
    uint8_t *kind, *reit;
    uint32_t *group, *mask, *label, *ord;
};

__global__ void __launch_bounds__(kThreads) k_mig_mark(uint32_t n, const double* __restrict__ px, double cell_width, const ShardGeom sh, int32_t margin,
                                                       uint32_t* rows, int32_t* dest, uint32_t* count, uint32_t cap) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t c = (int64_t)f64_as_i32(floor(px[i] / cell_width));
    if (c >= (int64_t)sh.col_lo - margin && c < (int64_t)sh.col_hi + margin) return;
    int d = 0;
    while (d + 1 < sh.world && c >= (int64_t)sh.bounds[d + 1]) ++d;  // the strip that holds column c
    if (d == sh.rank) return;
    const uint32_t pos = atomicAdd(count, 1u);
    if (pos < cap) {
        rows[pos] = i;
        dest[pos] = d;
    }
}

// Staging block: m doubles per column (11 columns), then label, ord, group, mask (m u32 each), then kind (m bytes).
__host__ __device__ inline size_t mig_stage_bytes(size_t m) { return m * (11 * 8 + 4 * 4 + 1) + 64; }

__global__ void __launch_bounds__(kThreads) k_mig_pack(uint32_t m, const uint32_t* __restrict__ rows, const MigCols c, unsigned char* stage, uint32_t* vmap,
                                                       uint32_t id_space) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const uint32_t r = rows[j];
    double* d = reinterpret_cast<double*>(stage);
#pragma unroll
    for (int k = 0; k < 11; ++k) d[(size_t)k * m + j] = c.col[k][r];
    uint32_t* u = reinterpret_cast<uint32_t*>(d + (size_t)11 * m);
    const uint32_t gid = c.label[r];
    u[j] = gid;
    u[(size_t)m + j] = c.ord[r];
    u[(size_t)2 * m + j] = c.group[r];
    u[(size_t)3 * m + j] = c.mask[r];
    reinterpret_cast<unsigned char*>(u + (size_t)4 * m)[j] = c.kind[r];
    if (vmap && gid < id_space) vmap[gid] = 0xffffffffu;
}

// moves[2k] = destination row (a hole), moves[2k + 1] = source row (a survivor of the tail)
__global__ void __launch_bounds__(kThreads) k_mig_fill(uint32_t n_moves, const uint32_t* __restrict__ moves, const MigCols c) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_moves) return;
    const uint32_t dst = moves[2 * j], src = moves[2 * j + 1];
#pragma unroll
    for (int k = 0; k < 11; ++k) c.col[k][dst] = c.col[k][src];
    c.kind[dst] = c.kind[src];
    c.reit[dst] = c.reit[src];
    c.group[dst] = c.group[src];
    c.mask[dst] = c.mask[src];
    c.label[dst] = c.label[src];
    c.ord[dst] = c.ord[src];
}

// number of elements of the ascending array a[0, m) that are < x
__device__ __forceinline__ uint32_t count_below(const uint32_t* __restrict__ a, uint32_t m, uint32_t x) {
    uint32_t lo = 0, hi = m;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < x) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(kThreads) k_mig_ord_out(uint32_t n, uint32_t* ord, const uint32_t* __restrict__ gone_ord_sorted, uint32_t m) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ord[i] -= count_below(gone_ord_sorted, m, ord[i]);
}

// below[p] += 1 for every resident with exactly p arrivals below it: the residents below arrival j are sum(below[0..j])
__global__ void __launch_bounds__(kThreads) k_mig_ord_in(uint32_t n, uint32_t* ord, const uint32_t* __restrict__ label, const uint32_t* __restrict__ new_ids_sorted,
                                                         uint32_t m, uint32_t* below) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t p = count_below(new_ids_sorted, m, label[i]);
    ord[i] += p;
    atomicAdd(below + p, 1u);
}

}  // namespace cb200
