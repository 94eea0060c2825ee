// The tile path of the bulk step: binning, candidate enumeration and the pair solve of one block of grid cells done
// entirely in shared memory — no global slot table, no global cell entries, no work items, no record gathers.
//
// Rows of the hitbox table are kept in grid-slot order (kernels.cuh "Table order"), so the hitboxes whose centre cell lies
// in a *block tile* — 2^tj consecutive slots, an (8 * 2^k) x 8 or 8 x 2^k block of cells — are one contiguous run of rows.
// Per step:
//   k_stage_a_tile   stage A as before (rebase, solitaire check, swept bounds, IndexRect; coalesced SoA reads and writes) —
//                    but no cell counting and no 96-byte record per row.  A hitbox whose IndexRect reaches into a tile other
//                    than its home tile (the tile its row belongs to) is appended, as a self-contained 96-byte record, to that
//                    tile's *visitor slab*: the single-GPU twin of the multi-GPU halo.
//   k_tile_solve     persistent blocks, one tile at a time: the tile's rows (coalesced SoA loads) and its visitor slab become
//                    records in shared memory; Grid::update_area (grid.rs:137-175) is a shared-memory counting sort over the
//                    tile's cells (histogram, block scan, scatter of 4-byte local entries); Grid::overlapping_ids
//                    (grid.rs:115-135) walks every cell run there; the first exit of collide_time (solvers.rs:27-30, swept
//                    bounds do not touch) is tested from edges computed once per record; the survivors queue up in shared
//                    memory and are solved with full warps; events leave through a shared-memory buffer.
//   k_solve          (kernels.cuh) over the overlapping pairs only — separate_time — and the last-block min-reduction.
// Exactness never depends on the row order: a tile's block considers its rows plus its visitors = every hitbox with a cell in
// the tile, and clips every rect to the cells it is working on; a pair is taken in the cell (max sx, max sy) only, as on the
// global path.  A tile (or part of a tile) whose records do not fit the shared-memory capacity is split by cell range and
// processed in several passes; a single cell that does not fit, or a visitor slab that overflows, raises `tile_fail` and the
// host repeats the step on the global path (engine.cu).
#pragma once
#include "kernels.cuh"

namespace cb200 {

struct TileGeom {
    int tj;       // log2 of the slots in a block tile (4 .. 11: a tile never straddles a scan chunk of 2048 slots)
    int bx, by;   // log2 of the tile's extent in cells
    int nx, ny;   // log2 of the number of tiles per grid period along x / y
    uint32_t n_tiles;
};
__host__ __device__ inline TileGeom make_tile_geom(const GridGeom& g, int tj) {
    TileGeom t;
    const int a = g.tx(), b = g.ty();
    t.tj = tj;
    if (tj <= a) {
        t.bx = tj; t.by = 0;
    } else if (tj <= a + b) {
        t.bx = a; t.by = tj - a;
    } else if (tj <= g.lw + b) {
        t.bx = tj - b; t.by = b;
    } else {
        t.bx = g.lw; t.by = tj - g.lw;
    }
    t.nx = g.lw - t.bx;
    t.ny = g.lh - t.by;
    t.n_tiles = 1u << (g.lw + g.lh - tj);
    return t;
}

struct TileTables {
    const uint2* rows;      // [n_tiles]: (first row, end row) of the tile's home rows (written at the re-sort)
    const uint32_t* home;   // [n]: home tile of each row
    uint32_t* vis_count;    // [2][n_tiles] by step parity: this step's counts are appended by stage A, the other half is zeroed by k_tile_solve
    HbRec* vis;             // [n_tiles * vcap]
    uint32_t vcap;
    uint32_t n_tiles;
};

// ---------------------------------------------------------------------------------------------------------
// Stage A of the tile path.  Same per-hitbox arithmetic as k_stage_a (hb_update_core); differs in what it writes.
struct TileStageArgs {
    TileGeom tg;
    TileTables tt;
    uint32_t has_overlaps;  // the overlap set is not empty: rows whose hitbox has an overlap also store their record (k_solve gathers those)
};

__global__ void __launch_bounds__(kThreads, CB200_LB_A) k_stage_a_tile(const StepArgs a, const TileStageArgs ta) {
    MinKey best{kKeyMax, kKeyMax};
    unsigned long long n_sol = 0;
    const StepDyn dyn = *a.dyn;
    uint32_t* const vis_count = ta.tt.vis_count + (size_t)(dyn.tag & 1ull) * ta.tt.n_tiles;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
        HbCore h;
        h.px = a.s.px[i]; h.py = a.s.py[i]; h.w = a.s.w[i]; h.h = a.s.h[i];
        h.vx = a.s.vx[i]; h.vy = a.s.vy[i]; h.rw = a.s.rw[i]; h.rh = a.s.rh[i];
        h.start = a.s.start[i]; h.pub_end = a.s.pub_end[i];
        h.kind = a.s.kind[i]; h.group = a.s.group[i]; h.mask = a.s.mask[i];
        const uint32_t gid = a.label[i], o = a.ord[i];
        NewVel nv;
        nv.has_vx = a.nvx != nullptr; nv.has_vy = a.nvy != nullptr; nv.has_rw = a.nrw != nullptr; nv.has_rh = a.nrh != nullptr;
        nv.vx = nv.has_vx ? __ldg(a.nvx + o) : 0.0;
        nv.vy = nv.has_vy ? __ldg(a.nvy + o) : 0.0;
        nv.rw = nv.has_rw ? __ldg(a.nrw + o) : 0.0;
        nv.rh = nv.has_rh ? __ldg(a.nrh + o) : 0.0;
        nv.end_time = a.nend ? __ldg(a.nend + (a.rebase ? o : i)) : dyn.end_scalar;
        const uint32_t err = hb_update_core(h, dyn.T, a.rebase != 0, nv, a.cell_width, a.padding, a.geom);
        if (err) report_error(a.ctr, err, gid);
        a.s.px[i] = h.px; a.s.py[i] = h.py; a.s.w[i] = h.w; a.s.h[i] = h.h;
        if (a.nvx) a.s.vx[i] = h.vx;
        if (a.nvy) a.s.vy[i] = h.vy;
        if (a.nrw) a.s.rw[i] = h.rw;
        if (a.nrh) a.s.rh[i] = h.rh;
        a.s.start[i] = h.start;
        a.s.end[i] = h.end;
        a.s.pub_end[i] = h.pub_end;
        a.s.reit[i] = h.reit ? 1 : 0;
        const HbRec rec = make_rec(h, gid);
        a.bink[i] = make_int4(rec.sx, rec.sy, (int)rec.span, (int)rec.kgb);
        if (ta.has_overlaps && ((__ldg(a.ov_bits + (gid >> 5)) >> (gid & 31u)) & 1u)) store_rec(a.rec + i, rec);
        if (h.binned) {
            // every tile the rect reaches into, other than the home tile of the row
            const uint32_t home = __ldg(ta.tt.home + i);
            const int32_t X0 = h.sx >> ta.tg.bx, X1 = (h.ex - 1) >> ta.tg.bx, Y0 = h.sy >> ta.tg.by, Y1 = (h.ey - 1) >> ta.tg.by;
            // (a rect spans at most one grid period — kFRectSpan — so at most 2^nx + 1 tile columns: the last one would alias the first)
            const int32_t Xe = min(X1, X0 + (int32_t)((1u << ta.tg.nx) - 1u)), Ye = min(Y1, Y0 + (int32_t)((1u << ta.tg.ny) - 1u));
            for (int32_t X = X0; X <= Xe; ++X)
                for (int32_t Y = Y0; Y <= Ye; ++Y) {
                    const uint32_t t = a.geom.slot(X << ta.tg.bx, Y << ta.tg.by) >> ta.tg.tj;
                    if (t == home) continue;
                    const uint32_t pos = atomicAdd(&vis_count[t], 1u);
                    if (pos < ta.tt.vcap) store_rec(ta.tt.vis + (size_t)t * ta.tt.vcap + pos, rec);
                    else a.ctr->tile_fail = 1u;
                }
        }
        if (h.reit) {
            ++n_sol;
            const MinKey k{time_key(h.end), (uint64_t)gid};
            if (key_less(k, best)) best = k;
        }
    }
    best = block_min(best);
    const unsigned long long tot = block_sum(n_sol);
    if (threadIdx.x == 0) {
        a.partial[a.partial_base + blockIdx.x] = best;
        if (tot) atomicAdd(&a.ctr->n_solitaire, tot);
    }
}

// At the re-sort: the row range of every tile (the runs of a tile's slots are contiguous: a tile lies inside one scan chunk).
// slot_end = the slot table after k_resort_rank (end of each slot's run).
__global__ void __launch_bounds__(kThreads) k_tile_rows(uint32_t n_tiles, int tj, const uint32_t* __restrict__ slot_end, const uint32_t* __restrict__ chunk_base,
                                                        uint2* rows) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const uint32_t s0 = t << tj, s1 = ((t + 1u) << tj) - 1u;
    rows[t] = make_uint2(slot_run_start(slot_end, chunk_base, s0), slot_end[s1]);
}
__global__ void __launch_bounds__(kThreads) k_tile_home(uint32_t n, int tj, const uint32_t* __restrict__ src_row, const uint32_t* __restrict__ key, uint32_t* home) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) home[i] = key[src_row[i]] >> tj;
}

// ---------------------------------------------------------------------------------------------------------
struct TileSolveArgs {
    SolveArgs sv;           // cd, dyn, padding, events, ctr, partial, add_mode ... (rec / work unused here)
    StateCols s;            // own rows: rebased state
    const uint32_t* label;
    const int4* bink;       // (sx, sy, span, kgb) per row, written by stage A
    TileGeom tg;
    TileTables tt;
    uint32_t cap_rec, cap_ent, cap_q;  // shared-memory capacities: records, cell entries, queued survivors
    uint32_t partial_at;    // first slot of this kernel's block minima in sv.partial
    uint2* list_out;        // LIST: accepted candidate pairs (gid hi, gid lo) instead of solving them
    unsigned long long* n_list;
    unsigned long long cap_list;
};

constexpr uint32_t kTileThreads = 256;
constexpr uint32_t kTileEvBuf = 2 * kTileThreads;  // events a block collects before it appends them with one atomic
constexpr uint32_t kTileVisFlag = 0x80000000u;
constexpr int kTileStack = 12;

__host__ __device__ inline size_t tile_smem_bytes(uint32_t cap_rec, uint32_t cap_ent, uint32_t cap_q, int tj) {
    const size_t nc = (size_t)1 << tj;
    return (size_t)cap_rec * (96 + 40 + 4) + (nc + 1) * 4 * 2 + (size_t)cap_ent * 4 + (size_t)cap_q * 4 + (size_t)kTileEvBuf * 16 + 64;
}

// Does any cell of the rect fold into local cells [lo, hi) of the tile whose first slot is `base`?
__device__ __forceinline__ bool rect_touches(const GridGeom& g, int32_t sx, int32_t sy, uint32_t span, uint32_t base, uint32_t lo, uint32_t hi) {
    const int32_t ex = sx + (int32_t)(span & 0xffffu), ey = sy + (int32_t)(span >> 16);
    for (int32_t x = sx; x < ex; ++x)
        for (int32_t y = sy; y != ey; ++y) {
            const uint32_t c = g.slot(x, y) - base;  // (unsigned: slots below the tile wrap to huge values)
            if (c >= lo && c < hi) return true;
        }
    return false;
}

#ifndef CB200_LB_TILE
#define CB200_LB_TILE 2
#endif
template <bool LIST>
__global__ void __launch_bounds__(kTileThreads, CB200_LB_TILE) k_tile_solve(const TileSolveArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    // ---- shared-memory carve-up ----
    int4* const s_rec = reinterpret_cast<int4*>(smem);                                         // [cap_rec * 6]
    double* const s_edge = reinterpret_cast<double*>(smem + (size_t)a.cap_rec * 96);           // [cap_rec * 5]: l, r, b, t, dur_key
    uint32_t* const s_src = reinterpret_cast<uint32_t*>(smem + (size_t)a.cap_rec * 136);       // [cap_rec]: row | visitor index
    const uint32_t NC = 1u << a.tg.tj;
    uint32_t* const s_off = s_src + a.cap_rec;   // [NC + 1] run starts
    uint32_t* const s_cur = s_off + NC + 1;      // [NC + 1] counts, then cursors
    uint32_t* const s_ent = s_cur + NC + 1;      // [cap_ent]: record index | local cell << 16
    uint32_t* const s_q = s_ent + a.cap_ent;     // [cap_q]: record of the entry | record of its predecessor << 16
    PairEvent* const s_ev = reinterpret_cast<PairEvent*>((reinterpret_cast<uintptr_t>(s_q + a.cap_q) + 15) & ~(uintptr_t)15);  // [kTileEvBuf]
    __shared__ uint32_t s_nsel, s_nq, s_nev, s_sp, s_any;
    __shared__ unsigned long long s_evbase;
    __shared__ uint2 s_stack[kTileStack];

    const int lane = threadIdx.x & 31;
    const StepDyn dyn = *a.sv.dyn;
    const double T = dyn.T;
    const uint32_t par = (uint32_t)(dyn.tag & 1ull);
    uint32_t* const cnt_now = a.tt.vis_count + (size_t)par * a.tt.n_tiles;
    uint32_t* const cnt_next = a.tt.vis_count + (size_t)(par ^ 1u) * a.tt.n_tiles;
    const GridGeom geom = a.sv.cd.geom;

    MinKey best{kKeyMax, kKeyMax};
    unsigned long long n_col = 0, n_entries = 0, n_cells = 0, n_pairs = 0;

    auto flush_events = [&]() {  // all threads
        __syncthreads();
        const uint32_t n_ev = s_nev;
        if (n_ev) {
            if (threadIdx.x == 0) s_evbase = atomicAdd(&a.sv.ctr->n_pair_events, (unsigned long long)n_ev);
            __syncthreads();
            const unsigned long long pos = s_evbase;
            for (uint32_t k = threadIdx.x; k < n_ev; k += blockDim.x)
                if (pos + k < a.sv.cap_events) a.sv.events[pos + k] = s_ev[k];
            __syncthreads();
            if (threadIdx.x == 0) s_nev = 0;
            __syncthreads();
        }
    };
    if (threadIdx.x == 0) {
        s_nev = 0;
        s_nq = 0;
    }
    __syncthreads();

    for (uint32_t t = blockIdx.x; t < a.tt.n_tiles; t += gridDim.x) {
        const uint2 rows = __ldg(a.tt.rows + t);
        const uint32_t n_vis = min(cnt_now[t], a.tt.vcap);
        if (!LIST && threadIdx.x == 0) cnt_next[t] = 0u;  // the slab counter the NEXT step appends to
        if (rows.y == rows.x && n_vis == 0) continue;
        const uint32_t base = t << a.tg.tj;
        const HbRec* const vis = a.tt.vis + (size_t)t * a.tt.vcap;
        if (threadIdx.x == 0) {
            s_stack[0] = make_uint2(0u, NC);
            s_sp = 1;
        }
        __syncthreads();
        while (s_sp) {  // (block-uniform: s_sp changes only between barriers)
            const uint2 range = s_stack[s_sp - 1];
            const uint32_t lo = range.x, hi = range.y;
            __syncthreads();
            if (threadIdx.x == 0) {
                s_sp -= 1;
                s_nsel = 0;
            }
            for (uint32_t c = threadIdx.x; c <= hi - lo; c += blockDim.x) s_cur[c] = 0u;
            __syncthreads();
            // ---- select: the rows / visitors with a cell in [lo, hi), in order (a warp's 32 consecutive rows stay together) ----
            const uint32_t n_own = rows.y - rows.x, n_all = n_own + n_vis;
            for (uint32_t b0 = (threadIdx.x & ~31u); b0 < n_all; b0 += blockDim.x) {
                const uint32_t k = b0 + (uint32_t)lane;
                bool take = false;
                uint32_t src = 0;
                if (k < n_own) {
                    const int4 bk = __ldg(a.bink + rows.x + k);
                    take = ((uint32_t)bk.w & 2u) && rect_touches(geom, bk.x, bk.y, (uint32_t)bk.z, base, lo, hi);
                    src = rows.x + k;
                } else if (k < n_all) {
                    const int4 hd = __ldg(reinterpret_cast<const int4*>(vis + (k - n_own)));
                    take = rect_touches(geom, hd.x, hd.y, (uint32_t)hd.z, base, lo, hi);
                    src = kTileVisFlag | (k - n_own);
                }
                const unsigned m = __ballot_sync(0xffffffffu, take);
                if (m) {
                    uint32_t at = 0;
                    if (lane == 0) at = atomicAdd(&s_nsel, (uint32_t)__popc(m));
                    at = __shfl_sync(0xffffffffu, at, 0) + (uint32_t)__popc(m & ((1u << lane) - 1u));
                    if (take && at < a.cap_rec) s_src[at] = src;
                }
            }
            __syncthreads();
            const uint32_t n_sel = s_nsel;
            if (n_sel > a.cap_rec) {
                // too many records for one pass: halve the cell range (a single cell that does not fit -> global path)
                if (threadIdx.x == 0) {
                    if (hi - lo < 2u || s_sp + 2 > kTileStack) {
                        a.sv.ctr->tile_fail = 1u;
                    } else {
                        const uint32_t mid = lo + ((hi - lo) >> 1);
                        s_stack[s_sp] = make_uint2(mid, hi);
                        s_stack[s_sp + 1] = make_uint2(lo, mid);
                        s_sp += 2;
                    }
                }
                __syncthreads();
                continue;
            }
            if (n_sel == 0) continue;
            // ---- load the records (own rows: from the SoA columns; visitors: their 96 bytes), swept edges, cell counts ----
            for (uint32_t k = threadIdx.x; k < n_sel; k += blockDim.x) {
                const uint32_t src = s_src[k];
                HbRec r;
                if (src & kTileVisFlag) {
                    const int4* p = reinterpret_cast<const int4*>(vis + (src & ~kTileVisFlag));
                    int4* q = reinterpret_cast<int4*>(&r);
#pragma unroll
                    for (int j = 0; j < 6; ++j) q[j] = __ldg(p + j);
                } else {
                    const int4 bk = __ldg(a.bink + src);
                    r.sx = bk.x; r.sy = bk.y; r.span = (uint32_t)bk.z; r.kgb = (uint32_t)bk.w;
                    r.gid = __ldg(a.label + src);
                    r.mask = a.s.mask[src];
                    r.dur = a.s.end[src] - T;  // (= r_time - T of stage A: the same operation on the same operands)
                    r.px = a.s.px[src]; r.py = a.s.py[src]; r.w = a.s.w[src]; r.h = a.s.h[src];
                    r.vx = a.s.vx[src]; r.vy = a.s.vy[src]; r.rw = a.s.rw[src]; r.rh = a.s.rh[src];
                }
                const int4* q = reinterpret_cast<const int4*>(&r);
#pragma unroll
                for (int j = 0; j < 6; ++j) s_rec[k * 6 + j] = q[j];
                DurHb d;
                d.px = r.px; d.py = r.py; d.w = r.w; d.h = r.h; d.vx = r.vx; d.vy = r.vy; d.rw = r.rw; d.rh = r.rh; d.dur = r.dur; d.kind = (int)(r.kgb & 1u);
                const SweptEdges e = swept_edges(d);
                s_edge[k * 5 + 0] = e.l; s_edge[k * 5 + 1] = e.r; s_edge[k * 5 + 2] = e.b; s_edge[k * 5 + 3] = e.t; s_edge[k * 5 + 4] = e.dur_key;
                const int32_t ex = r.sx + (int32_t)(r.span & 0xffffu), ey = r.sy + (int32_t)(r.span >> 16);
                for (int32_t x = r.sx; x < ex; ++x)
                    for (int32_t y = r.sy; y != ey; ++y) {
                        const uint32_t c = geom.slot(x, y) - base;
                        if (c >= lo && c < hi) atomicAdd(&s_cur[c - lo], 1u);
                    }
            }
            __syncthreads();
            // ---- scan the cell counts: run starts ----
            const uint32_t n_c = hi - lo;
            uint32_t n_ent, nz = 0;
            {
                const uint32_t per = (n_c + blockDim.x - 1) / blockDim.x;
                const uint32_t c0 = min(threadIdx.x * per, n_c), c1 = min(c0 + per, n_c);
                uint32_t sum = 0;
                for (uint32_t c = c0; c < c1; ++c) {
                    sum += s_cur[c];
                    nz += s_cur[c] != 0u;
                }
                uint32_t run = block_excl_scan(sum, &n_ent);
                for (uint32_t c = c0; c < c1; ++c) {
                    const uint32_t v = s_cur[c];
                    s_off[c] = run;
                    s_cur[c] = run;
                    run += v;
                }
                if (c1 == n_c) s_off[n_c] = n_ent;
            }
            __syncthreads();
            if (n_ent > a.cap_ent) {
                // too many cell entries for one pass: halve the cell range as above
                if (threadIdx.x == 0) {
                    if (hi - lo < 2u || s_sp + 2 > kTileStack) {
                        a.sv.ctr->tile_fail = 1u;
                    } else {
                        const uint32_t mid = lo + ((hi - lo) >> 1);
                        s_stack[s_sp] = make_uint2(mid, hi);
                        s_stack[s_sp + 1] = make_uint2(lo, mid);
                        s_sp += 2;
                    }
                }
                __syncthreads();
                continue;
            }
            n_cells += nz;
            if (threadIdx.x == 0) n_entries += n_ent;
            // ---- scatter: 4-byte local entries (record index | local cell << 16) ----
            for (uint32_t k = threadIdx.x; k < n_sel; k += blockDim.x) {
                const int4 hd = s_rec[k * 6];
                const int32_t ex = hd.x + (int32_t)((uint32_t)hd.z & 0xffffu), ey = hd.y + (int32_t)((uint32_t)hd.z >> 16);
                for (int32_t x = hd.x; x < ex; ++x)
                    for (int32_t y = hd.y; y != ey; ++y) {
                        const uint32_t c = geom.slot(x, y) - base;
                        if (c >= lo && c < hi) s_ent[atomicAdd(&s_cur[c - lo], 1u)] = k | ((c - lo) << 16);
                    }
            }
            __syncthreads();
            // ---- enumerate + filter (each thread walks the predecessors of its entries), solve whenever the queue fills ----
            uint32_t e = threadIdx.x;  // this thread's current entry
            uint32_t p = 0;            // next predecessor position of that entry
            bool fresh = true, done = e >= n_ent;
            for (;;) {
                while (!done) {
                    const uint32_t ent = s_ent[e];
                    const uint32_t ka = ent & 0xffffu, c = ent >> 16;
                    if (fresh) {
                        p = s_off[c];
                        fresh = false;
                    }
                    if (p == e) {
                        e += blockDim.x;
                        fresh = true;
                        done = e >= n_ent;
                        continue;
                    }
                    const uint32_t kb = s_ent[p] & 0xffffu;
                    // candidate test from the two record heads (candidate_geom: owner cell, column strip, interact groups)
                    const int4 a0 = s_rec[ka * 6], a1 = s_rec[ka * 6 + 1], b0 = s_rec[kb * 6], b1 = s_rec[kb * 6 + 1];
                    RecHead ha, hb;
                    ha.sx = a0.x; ha.sy = a0.y; ha.ex = a0.x + (int32_t)((uint32_t)a0.z & 0xffffu); ha.ey = a0.y + (int32_t)((uint32_t)a0.z >> 16);
                    ha.gid = (uint32_t)a0.w; ha.kgb = (uint32_t)a1.z; ha.mask = (uint32_t)a1.w;
                    hb.sx = b0.x; hb.sy = b0.y; hb.ex = b0.x + (int32_t)((uint32_t)b0.z & 0xffffu); hb.ey = b0.y + (int32_t)((uint32_t)b0.z >> 16);
                    hb.gid = (uint32_t)b0.w; hb.kgb = (uint32_t)b1.z; hb.mask = (uint32_t)b1.w;
                    bool a_is_hi;
                    bool pass = candidate_geom(a.sv.cd, ha, hb, base + lo + c, &a_is_hi);
                    if (pass) {
                        if (!LIST) {
                            SweptEdges ea;
                            ea.l = s_edge[ka * 5]; ea.r = s_edge[ka * 5 + 1]; ea.b = s_edge[ka * 5 + 2]; ea.t = s_edge[ka * 5 + 3]; ea.dur_key = s_edge[ka * 5 + 4];
                            pass = swept_filter_pass(ea, s_edge[kb * 5], s_edge[kb * 5 + 1], s_edge[kb * 5 + 2], s_edge[kb * 5 + 3], s_edge[kb * 5 + 4]);
                        }
                        if (pass) {
                            const uint32_t at = atomicAdd(&s_nq, 1u);
                            if (at >= a.cap_q) break;  // queue full: this pair is retried after the solve round (p not advanced)
                            s_q[at] = ka | (kb << 16);
                        }
                        n_col += 1;  // (counted when the pair leaves the filter, exactly once: a retried pair breaks before this)
                    }
                    n_pairs += 1;
                    p += 1;
                }
                // ---- solve round ----
                __syncthreads();
                const uint32_t n_q = min(s_nq, a.cap_q);
                for (uint32_t q0 = 0; q0 < n_q; q0 += blockDim.x) {
                    if (s_nev + blockDim.x > kTileEvBuf) flush_events();  // (block-uniform: s_nev is stable between barriers)
                    const uint32_t q = q0 + threadIdx.x;
                    if (q < n_q) {
                        const uint32_t pr = s_q[q];
                        const int4* pa = s_rec + (pr & 0xffffu) * 6;
                        const int4* pb = s_rec + (pr >> 16) * 6;
                        RecHead h1, h2;
                        DurHb d1, d2;
                        decode_rec(pa[0], pa[1], pa[2], pa[3], pa[4], pa[5], &h1, &d1);
                        decode_rec(pb[0], pb[1], pb[2], pb[3], pb[4], pb[5], &h2, &d2);
                        const bool first_hi = h1.gid > h2.gid;
                        if (!is_overlapping(a.sv.cd, h1, h2, first_hi)) {
                            if (LIST) {
                                const unsigned long long pos = atomicAdd(a.n_list, 1ull);
                                if (pos < a.cap_list) a.list_out[pos] = first_hi ? make_uint2(h1.gid, h2.gid) : make_uint2(h2.gid, h1.gid);
                            } else {
                                PairEvent ev;
                                if (solve_pair(a.sv, T, h1, d1, h2, d2, first_hi, &ev, &best)) s_ev[atomicAdd(&s_nev, 1u)] = ev;
                            }
                        }
                    }
                    __syncthreads();
                }
                __syncthreads();
                if (threadIdx.x == 0) {
                    s_nq = 0;
                    s_any = 0;
                }
                __syncthreads();
                if (!done) s_any = 1u;
                __syncthreads();
                if (!s_any) break;
            }
        }
        __syncthreads();
    }
    flush_events();
    best = block_min(best);
    const unsigned long long col_tot = block_sum(n_col), ent_tot = block_sum(n_entries), cell_tot = block_sum(n_cells), pair_tot = block_sum(n_pairs);
    if (threadIdx.x == 0 && !LIST) {
        a.sv.partial[a.partial_at + blockIdx.x] = best;
        if (col_tot) atomicAdd(&a.sv.ctr->n_collide, col_tot);
        if (ent_tot) atomicAdd(&a.sv.ctr->n_entries, ent_tot);
        if (cell_tot) atomicAdd(&a.sv.ctr->n_cells, cell_tot);
        if (pair_tot) atomicAdd(&a.sv.ctr->n_tile_pairs, pair_tot);
    }
}

}  // namespace cb200
