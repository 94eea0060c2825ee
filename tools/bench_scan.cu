// Micro-benchmark of slot-scan variants on a synthetic slot table (4M slots, ~0.47 entries/slot).  Diagnostic only.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cub/cub.cuh>
#include "../collider-rs_b200/csrc/kernels.cuh"
using namespace cb200;

// V1: one tile per block held in registers, decoupled via epoch flags (all predecessors), single global read.
template <int ITEMS>
__global__ void __launch_bounds__(256) k_scan_v1(uint32_t* data, unsigned long long* state, uint32_t epoch, uint32_t* ticket, uint32_t ticket_base,
                                                 unsigned long long* totals) {
    __shared__ uint32_t s_chunk, s_prefix;
    __shared__ uint32_t s_warp[8];
    if (threadIdx.x == 0) s_chunk = atomicAdd(ticket, 1u) - ticket_base;
    __syncthreads();
    const uint32_t chunk = s_chunk;
    uint4* p = reinterpret_cast<uint4*>(data + (size_t)chunk * 256 * ITEMS + (size_t)threadIdx.x * ITEMS);
    uint32_t v[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS / 4; ++k) {
        const uint4 q = p[k];
        v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
    }
    uint32_t sum = 0, nz = 0;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) { sum += v[k]; nz += v[k] != 0; }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { const uint32_t x = s_warp[w]; before += w < warp ? x : 0u; total += x; }
    if (threadIdx.x == 0) atomicExch(&state[chunk], ((unsigned long long)epoch << 32) | total);
    if (warp == 0) {
        uint32_t prefix = 0;
        for (uint32_t j = lane; j < chunk; j += 32) {
            unsigned long long st;
            do { st = *reinterpret_cast<volatile unsigned long long*>(&state[j]); } while ((uint32_t)(st >> 32) != epoch);
            prefix += (uint32_t)st;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) prefix += __shfl_xor_sync(0xffffffffu, prefix, o);
        if (lane == 0) { s_prefix = prefix; if (chunk == gridDim.x - 1) totals[0] = prefix + total; }
    }
    nz = (uint32_t)block_sum(nz);
    if (threadIdx.x == 0 && nz) atomicAdd(&totals[1], (unsigned long long)nz);
    uint32_t run = s_prefix + before + inc - sum;
#pragma unroll
    for (int k = 0; k < ITEMS / 4; ++k) {
        uint4 q;
        q.x = run; run += v[4 * k]; q.y = run; run += v[4 * k + 1]; q.z = run; run += v[4 * k + 2]; q.w = run; run += v[4 * k + 3];
        p[k] = q;
    }
}

// V2: like V1 but the base comes from an atomicAdd on a global cursor (no waiting; chunk order arbitrary)
template <int ITEMS>
__global__ void __launch_bounds__(256) k_scan_v2(uint32_t* data, uint32_t* cursor, uint32_t* chunk_base, unsigned long long* totals) {
    __shared__ uint32_t s_prefix;
    __shared__ uint32_t s_warp[8];
    const uint32_t chunk = blockIdx.x;
    uint4* p = reinterpret_cast<uint4*>(data + (size_t)chunk * 256 * ITEMS + (size_t)threadIdx.x * ITEMS);
    uint32_t v[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS / 4; ++k) {
        const uint4 q = p[k];
        v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
    }
    uint32_t sum = 0, nz = 0;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) { sum += v[k]; nz += v[k] != 0; }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { const uint32_t x = s_warp[w]; before += w < warp ? x : 0u; total += x; }
    if (threadIdx.x == 0) { s_prefix = atomicAdd(cursor, total); chunk_base[chunk] = s_prefix; }
    nz = (uint32_t)block_sum(nz);
    if (threadIdx.x == 0 && nz) atomicAdd(&totals[1], (unsigned long long)nz);
    uint32_t run = s_prefix + before + inc - sum;
#pragma unroll
    for (int k = 0; k < ITEMS / 4; ++k) {
        uint4 q;
        q.x = run; run += v[4 * k]; q.y = run; run += v[4 * k + 1]; q.z = run; run += v[4 * k + 2]; q.w = run; run += v[4 * k + 3];
        p[k] = q;
    }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

int main() {
    const uint32_t n = 1u << 22;
    std::vector<uint32_t> h(n);
    uint64_t s = 12345;
    for (uint32_t i = 0; i < n; ++i) { s = s * 6364136223846793005ull + 1442695040888963407ull; uint32_t r = (uint32_t)(s >> 33) % 100; h[i] = r < 63 ? 0 : r < 90 ? 1 : r < 98 ? 2 : 3; }
    uint32_t *d, *src, *ticket, *cursor, *chunk_base, *flush;
    unsigned long long *state, *totals;
    Counters* ctr;
    CK(cudaMalloc(&d, n * 4)); CK(cudaMalloc(&src, n * 4)); CK(cudaMalloc(&ticket, 4)); CK(cudaMalloc(&cursor, 4)); CK(cudaMalloc(&chunk_base, 65536 * 4));
    CK(cudaMalloc(&state, 65536 * 8)); CK(cudaMalloc(&totals, 16)); CK(cudaMalloc(&ctr, sizeof(Counters)));
    const size_t flush_bytes = 512u << 20;
    CK(cudaMalloc(&flush, flush_bytes));
    CK(cudaMemcpy(src, h.data(), n * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(state, 0, 65536 * 8)); CK(cudaMemset(ticket, 0, 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    void* cub_tmp = nullptr; size_t cub_bytes = 0;
    cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, d, d, n);
    CK(cudaMalloc(&cub_tmp, cub_bytes));
    uint32_t epoch = 0, tbase = 0;
    for (int cold = 0; cold < 2; ++cold)
        for (int variant = 0; variant < 8; ++variant) {
            float best = 1e9, sumt = 0;
            const int reps = 20;
            for (int r = 0; r < reps; ++r) {
                CK(cudaMemcpy(d, src, n * 4, cudaMemcpyDeviceToDevice));
                if (cold) CK(cudaMemset(flush, r, flush_bytes));
                CK(cudaMemset(cursor, 0, 4)); CK(cudaMemset(totals, 0, 16)); CK(cudaMemset(ctr, 0, sizeof(Counters)));
                CK(cudaDeviceSynchronize());
                cudaEventRecord(e0);
                switch (variant) {
                    case 0: k_scan_slots<<<n / kScanChunk, 256>>>(d, totals, chunk_base, totals + 1); break;  // the engine's kernel
                    case 1: k_scan_v1<16><<<n / 4096, 256>>>(d, state, ++epoch, ticket, tbase, totals); tbase += n / 4096; break;
                    case 2: k_scan_v1<32><<<n / 8192, 256>>>(d, state, ++epoch, ticket, tbase, totals); tbase += n / 8192; break;
                    case 3: k_scan_v2<16><<<n / 4096, 256>>>(d, cursor, chunk_base, totals); break;
                    case 4: k_scan_v2<8><<<n / 2048, 256>>>(d, cursor, chunk_base, totals); break;
                    case 5: k_scan_v2<32><<<n / 8192, 256>>>(d, cursor, chunk_base, totals); break;
                    case 6: cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, d, d, n); break;
                    case 7: CK(cudaMemcpyAsync(flush, d, n * 4, cudaMemcpyDeviceToDevice)); break;
                }
                cudaEventRecord(e1);
                CK(cudaDeviceSynchronize());
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                best = fminf(best, ms); sumt += ms;
            }
            const char* names[] = {"engine k_scan_slots", "v1 regs 1024x4096 wait-all", "v1 regs 512x8192 wait-all", "v2 atomic 1024x4096", "v2 atomic 2048x2048", "v2 atomic 512x8192", "cub ExclusiveSum", "memcpy d2d 16MB"};
            printf("%s %-32s best %.2f us  mean %.2f us\n", cold ? "cold" : "warm", names[variant], best * 1e3, sumt / reps * 1e3);
        }
    return 0;
}
