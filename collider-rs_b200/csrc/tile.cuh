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
                    else atomicOr(&a.ctr->tile_fail, 1u);  // (1: a visitor slab is full)
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
// Survivors: the candidate pairs whose swept bounds touch, as (source of the higher id, source of the lower id).  A source
// names where the hitbox's data lives: a table row, or kTileVisFlag | index into the visitor slabs.  8 bytes per pair; the
// finish kernel gathers the two hitboxes (only the few per cent of the candidate pairs that survive the filter are ever
// gathered).  (Measured alternative: the tile kernel assembling both 96-byte records per survivor so that the finish kernel
// streams them — the lane-serial assembly cost the tile kernel 35 us for 40k survivors and saved the finish kernel 3.)
struct SurvList {
    uint2* items;
    uint32_t cap;
};

struct TileSolveArgs {
    SolveArgs sv;           // cd, dyn, padding, events, ctr, partial, add_mode ... (work lists unused here)
    StateCols s;            // own rows: rebased state
    const uint32_t* label;
    const int4* bink;       // (sx, sy, span, kgb) per row, written by stage A
    TileGeom tg;
    TileTables tt;
    SurvList surv;
    uint32_t cap_rec, cap_ent;  // shared-memory capacities per warp: records, cell entries
};

constexpr uint32_t kTileWarps = 4;                  // warps per block: every warp works on its own tile
constexpr uint32_t kTileThreads = 32 * kTileWarps;
constexpr uint32_t kTileVisFlag = 0x80000000u;
constexpr int kTileStack = 24;
constexpr uint32_t kTileIter = 3;                   // (entry, predecessor) pairs a lane tests per round

// One record of a tile in shared memory: all the candidate test and the swept-bounds filter need (80 bytes).
struct alignas(16) TileRec {
    int32_t sx, sy;
    uint32_t span, gid;
    uint32_t kgb, mask, src, pad;
    double l, r, b, t;  // swept edges over the record's own duration (narrowphase.cuh swept_edges)
    double dur_key;
    unsigned long long cells;  // tile_cell_list (valid when kgb bit 4 is set; bits 4.. of kgb are free in a record)
};
static_assert(sizeof(TileRec) == 80, "TileRec is five 16-byte chunks");

__host__ __device__ inline size_t tile_warp_smem_bytes(uint32_t cap_rec, uint32_t cap_ent, int tj) {
    const size_t nc = (size_t)1 << tj;
    return ((size_t)cap_rec * sizeof(TileRec) + (nc + 1) * 4 * 2 + (size_t)cap_ent * 4 + 8 * kTileStack + 15) & ~(size_t)15;
}

// GridGeom::slot is separable: slot(x, y) = slot_x(x) | slot_y(y).
__device__ __forceinline__ uint32_t slot_x(const GridGeom& g, int32_t x) {
    const uint32_t ux = (uint32_t)x & ((1u << g.lw) - 1u);
    const int a = g.tx(), b = g.ty();
    return (ux & ((1u << a) - 1u)) | ((ux >> a) << (a + b));
}
__device__ __forceinline__ uint32_t slot_y(const GridGeom& g, int32_t y) {
    const uint32_t uy = (uint32_t)y & ((1u << g.lh) - 1u);
    const int a = g.tx(), b = g.ty();
    return ((uy & ((1u << b) - 1u)) << a) | ((uy >> b) << (g.lw + b));
}
// The cells of a rect that fold into the tile whose first slot is `base` (NC slots), as up to four 16-bit local cell ids
// (0xffff = none) — rects of at most 2 x 2 cells, i.e. nearly all of them; *complete = false for a larger rect (its cells are
// then walked by for_tile_cells every time).
__device__ __forceinline__ unsigned long long tile_cell_list(const GridGeom& g, int32_t sx, int32_t sy, uint32_t span, uint32_t base, uint32_t NC, bool* complete) {
    const uint32_t nx = span & 0xffffu, ny = span >> 16;
    *complete = nx <= 2u && ny <= 2u;
    const uint32_t fx0 = slot_x(g, sx), fx1 = slot_x(g, sx + 1), fy0 = slot_y(g, sy), fy1 = slot_y(g, sy + 1);
    uint32_t c[4] = {(fx0 | fy0) - base, (fx0 | fy1) - base, (fx1 | fy0) - base, (fx1 | fy1) - base};
    const bool ok[4] = {nx >= 1u && ny >= 1u, nx >= 1u && ny >= 2u, nx >= 2u && ny >= 1u, nx >= 2u && ny >= 2u};
    unsigned long long list = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) list |= (unsigned long long)((ok[k] && c[k] < NC) ? c[k] : 0xffffu) << (16 * k);
    return list;
}
// The local cells of a rect inside a tile.  f(c) is called for every cell of the rect that folds into local cells
// [lo, hi) of the tile whose first slot is `base`.
template <class F>
__device__ __forceinline__ void for_tile_cells(const GridGeom& g, int32_t sx, int32_t sy, uint32_t span, uint32_t base, uint32_t lo, uint32_t hi, F&& f) {
    const int32_t ex = sx + (int32_t)(span & 0xffffu), ey = sy + (int32_t)(span >> 16);
    for (int32_t x = sx; x < ex; ++x) {
        const uint32_t fx = slot_x(g, x);
        for (int32_t y = sy; y != ey; ++y) {
            const uint32_t c = (fx | slot_y(g, y)) - base;  // (unsigned: slots below the tile wrap to huge values)
            if (c >= lo && c < hi) f(c - lo);
        }
    }
}
// The same from a record's stored cell list when it is complete (kgb bit 4).
template <class F>
__device__ __forceinline__ void for_rec_cells(const GridGeom& g, int32_t sx, int32_t sy, uint32_t span, uint32_t kgb, unsigned long long list, uint32_t base, uint32_t lo,
                                              uint32_t hi, F&& f) {
    if (kgb & 16u) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t c = (uint32_t)(list >> (16 * k)) & 0xffffu;
            if (c >= lo && c < hi) f(c - lo);  // (0xffff >= hi always: NC <= 2048)
        }
    } else {
        for_tile_cells(g, sx, sy, span, base, lo, hi, f);
    }
}

// The hitbox a source names, as the solve stage needs it.
__device__ __forceinline__ void load_source(const TileSolveArgs& a, uint32_t src, double T, RecHead* h, DurHb* d) {
    if (src & kTileVisFlag) {
        const int4* p = reinterpret_cast<const int4*>(a.tt.vis + (src & ~kTileVisFlag));
        decode_rec(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3), __ldg(p + 4), __ldg(p + 5), h, d);
        return;
    }
    const int4 bk = __ldg(a.bink + src);
    h->sx = bk.x; h->sy = bk.y;
    h->ex = bk.x + (int32_t)((uint32_t)bk.z & 0xffffu);
    h->ey = bk.y + (int32_t)((uint32_t)bk.z >> 16);
    h->gid = __ldg(a.label + src);
    h->kgb = (uint32_t)bk.w;
    h->mask = __ldg(a.s.mask + src);
    h->dur = __ldg(a.s.end + src) - T;  // (= r_time - T of stage A: the same operation on the same operands)
    d->px = __ldg(a.s.px + src); d->py = __ldg(a.s.py + src); d->w = __ldg(a.s.w + src); d->h = __ldg(a.s.h + src);
    d->vx = __ldg(a.s.vx + src); d->vy = __ldg(a.s.vy + src); d->rw = __ldg(a.s.rw + src); d->rh = __ldg(a.s.rh + src);
    d->dur = h->dur;
    d->kind = (int)(h->kgb & 1u);
}

#ifndef CB200_LB_TILE
#define CB200_LB_TILE 5
#endif
// Binning + candidate enumeration + swept-bounds filter, one WARP per tile (no block-wide barrier anywhere); the survivors
// leave as source pairs (SurvList).  LIST: the filter is skipped, i.e. the list holds every pair that passes candidate_geom
// (the parity view).
template <bool LIST>
__global__ void __launch_bounds__(kTileThreads, CB200_LB_TILE) k_tile_bin(const TileSolveArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t NC = 1u << a.tg.tj;
    // ---- this warp's shared memory ----
    unsigned char* const mine = smem + (size_t)warp * tile_warp_smem_bytes(a.cap_rec, a.cap_ent, a.tg.tj);
    TileRec* const s_rec = reinterpret_cast<TileRec*>(mine);                 // [cap_rec]
    uint2* const s_stack = reinterpret_cast<uint2*>(s_rec + a.cap_rec);      // [kTileStack]: cell ranges still to do
    uint32_t* const s_off = reinterpret_cast<uint32_t*>(s_stack + kTileStack);  // [NC + 1] run starts
    uint32_t* const s_cur = s_off + NC + 1;                                  // [NC + 1] counts, then cursors
    uint32_t* const s_ent = s_cur + NC + 1;                                  // [cap_ent]: record index | local cell << 16

    const StepDyn dyn = *a.sv.dyn;
    const double T = dyn.T;
    const uint32_t par = (uint32_t)(dyn.tag & 1ull);
    const uint32_t* const cnt_now = a.tt.vis_count + (size_t)par * a.tt.n_tiles;
    uint32_t* const cnt_next = a.tt.vis_count + (size_t)(par ^ 1u) * a.tt.n_tiles;
    const GridGeom geom = a.sv.cd.geom;
    const uint32_t fold_x = (1u << geom.lw) - 1u, fold_y = (1u << geom.lh) - 1u;
    const unsigned lt_mask = (1u << lane) - 1u;

    unsigned long long n_col = 0, n_pairs = 0;
    uint32_t n_entries = 0, n_cells = 0;  // (lane 0 / per lane)

    const uint32_t n_warps = gridDim.x * kTileWarps;
    for (uint32_t t = blockIdx.x * kTileWarps + warp; t < a.tt.n_tiles; t += n_warps) {
        const uint2 rows = __ldg(a.tt.rows + t);
        const uint32_t n_vis = min(cnt_now[t], a.tt.vcap);
        if (!LIST && lane == 0) cnt_next[t] = 0u;  // the slab counter the NEXT step appends to
        if (rows.y == rows.x && n_vis == 0) continue;
        const uint32_t base = t << a.tg.tj;
        const size_t vis0 = (size_t)t * a.tt.vcap;
        const uint32_t n_own = rows.y - rows.x, n_all = n_own + n_vis;
        uint32_t sp = 1;  // (warp-uniform)
        if (lane == 0) s_stack[0] = make_uint2(0u, NC);
        __syncwarp();
        while (sp) {
            const uint2 range = s_stack[sp - 1];
            sp -= 1;
            const uint32_t lo = range.x, hi = range.y, n_c = hi - lo;
            for (uint32_t c = lane; c <= n_c; c += 32) s_cur[c] = 0u;
            __syncwarp();
            // ---- load: every row / visitor with a cell in [lo, hi) becomes a record in shared memory, in order; its cells are
            //      counted.  One round trip to global memory: all the columns of a row are requested before anything is known about it
            uint32_t n_sel = 0;
            for (uint32_t b0 = 0; b0 < n_all; b0 += 32) {
                const uint32_t k = b0 + (uint32_t)lane;
                bool take = false;
                TileRec r;
                DurHb d;
                if (k < n_own) {
                    const uint32_t src = rows.x + k;
                    const int4 bk = __ldg(a.bink + src);
                    r.gid = __ldg(a.label + src);
                    r.mask = a.s.mask[src];
                    const double end = a.s.end[src];
                    d.px = a.s.px[src]; d.py = a.s.py[src]; d.w = a.s.w[src]; d.h = a.s.h[src];
                    d.vx = a.s.vx[src]; d.vy = a.s.vy[src]; d.rw = a.s.rw[src]; d.rh = a.s.rh[src];
                    r.sx = bk.x; r.sy = bk.y; r.span = (uint32_t)bk.z; r.kgb = (uint32_t)bk.w;
                    r.src = src;
                    d.dur = end - T;  // (= r_time - T of stage A: the same operation on the same operands)
                    take = (r.kgb & 2u) != 0u;
                } else if (k < n_all) {
                    const int4* p = reinterpret_cast<const int4*>(a.tt.vis + vis0 + (k - n_own));
                    const int4 q0 = __ldg(p), q1 = __ldg(p + 1), q2 = __ldg(p + 2), q3 = __ldg(p + 3), q4 = __ldg(p + 4), q5 = __ldg(p + 5);
                    RecHead h;
                    decode_rec(q0, q1, q2, q3, q4, q5, &h, &d);
                    r.sx = q0.x; r.sy = q0.y; r.span = (uint32_t)q0.z; r.gid = h.gid; r.kgb = h.kgb; r.mask = h.mask;
                    r.src = kTileVisFlag | (uint32_t)(vis0 + (k - n_own));
                    take = true;
                }
                if (take) {
                    bool complete;
                    r.cells = tile_cell_list(geom, r.sx, r.sy, r.span, base, NC, &complete);
                    r.kgb = (r.kgb & ~16u) | (complete ? 16u : 0u);
                    bool any = false;
                    for_rec_cells(geom, r.sx, r.sy, r.span, r.kgb, r.cells, base, lo, hi, [&](uint32_t) { any = true; });
                    take = any;
                }
                const unsigned m = __ballot_sync(0xffffffffu, take);
                const uint32_t at = n_sel + (uint32_t)__popc(m & lt_mask);
                n_sel += (uint32_t)__popc(m);
                if (take && at < a.cap_rec) {
                    d.kind = (int)(r.kgb & 1u);
                    const SweptEdges e = swept_edges(d);
                    r.l = e.l; r.r = e.r; r.b = e.b; r.t = e.t; r.dur_key = e.dur_key;
                    r.pad = 0u;
                    const int4* q = reinterpret_cast<const int4*>(&r);
                    int4* o = reinterpret_cast<int4*>(s_rec + at);
#pragma unroll
                    for (int j = 0; j < 5; ++j) o[j] = q[j];
                    for_rec_cells(geom, r.sx, r.sy, r.span, r.kgb, r.cells, base, lo, hi, [&](uint32_t c) { atomicAdd(&s_cur[c], 1u); });
                }
            }
            __syncwarp();
            // ---- scan the cell counts: run starts ----
            uint32_t n_ent = 0, nz = 0;
            if (n_sel <= a.cap_rec) {  // (warp-uniform)
                const uint32_t per = (n_c + 31u) / 32u;
                const uint32_t c0 = min((uint32_t)lane * per, n_c), c1 = min(c0 + per, n_c);
                uint32_t sum = 0;
                for (uint32_t c = c0; c < c1; ++c) {
                    sum += s_cur[c];
                    nz += s_cur[c] != 0u;
                }
                uint32_t inc = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += u;
                }
                n_ent = __shfl_sync(0xffffffffu, inc, 31);
                uint32_t run = inc - sum;
                for (uint32_t c = c0; c < c1; ++c) {
                    const uint32_t v = s_cur[c];
                    s_off[c] = run;
                    s_cur[c] = run;
                    run += v;
                }
            }
            __syncwarp();
            if (n_sel > a.cap_rec || n_ent > a.cap_ent) {
                // too many records / cell entries for one pass: halve the cell range (a single cell that does not fit -> global path)
                if (n_c < 2u || sp + 2 > (uint32_t)kTileStack) {
                    if (lane == 0) atomicOr(&a.sv.ctr->tile_fail, n_c < 2u ? (n_sel > a.cap_rec ? 2u : 4u) : 8u);  // (2 / 4: one cell's records / entries do not fit; 8: stack)
                } else {
                    const uint32_t mid = lo + (n_c >> 1);
                    if (lane == 0) {
                        s_stack[sp] = make_uint2(mid, hi);
                        s_stack[sp + 1] = make_uint2(lo, mid);
                    }
                    sp += 2;
                }
                __syncwarp();
                continue;
            }
            if (n_sel == 0) continue;
            n_cells += nz;
            if (lane == 0) n_entries += n_ent;
            // ---- scatter: 4-byte local entries (record index | local cell << 16) ----
            for (uint32_t k = lane; k < n_sel; k += 32) {
                const int4 h0 = *reinterpret_cast<const int4*>(s_rec + k);
                const uint32_t kgb = s_rec[k].kgb;
                const unsigned long long cells = s_rec[k].cells;
                for_rec_cells(geom, h0.x, h0.y, (uint32_t)h0.z, kgb, cells, base, lo, hi, [&](uint32_t c) { s_ent[atomicAdd(&s_cur[c], 1u)] = k | (c << 16); });
            }
            __syncwarp();
            // ---- enumerate + filter: each lane walks the predecessors of its entries, at most kTileIter pairs per round; then the
            //      round's survivors are appended to the global list with one atomic ----
            uint32_t e = lane, p = 0, p_end = 0;
            bool fresh = true;
            for (;;) {
                uint32_t budget = kTileIter, my_q = 0;
                uint2 mine_q[kTileIter];
                while (e < n_ent && budget) {
                    const uint32_t ent = s_ent[e];
                    const uint32_t ka = ent & 0xffffu, c = ent >> 16;
                    if (fresh) {
                        p = s_off[c];
                        p_end = e;
                        fresh = false;
                    }
                    if (p < p_end) {
                        const TileRec ra = s_rec[ka];
                        const int32_t aex = ra.sx + (int32_t)(ra.span & 0xffffu), aey = ra.sy + (int32_t)(ra.span >> 16);
                        uint32_t cell_x, cell_y;  // the folded cell of the run: a pair belongs to it iff its owner cell folds to the same
                        geom.unslot(base + lo + c, &cell_x, &cell_y);
                        for (; p < p_end && budget; ++p, --budget) {
                            const uint32_t kb = s_ent[p] & 0xffffu;
                            const int4 b0 = *reinterpret_cast<const int4*>(s_rec + kb), b1 = *(reinterpret_cast<const int4*>(s_rec + kb) + 1);
                            const int32_t bex = b0.x + (int32_t)((uint32_t)b0.z & 0xffffu), bey = b0.y + (int32_t)((uint32_t)b0.z >> 16);
                            const uint32_t bgid = (uint32_t)b0.w, bkgb = (uint32_t)b1.x, bmask = (uint32_t)b1.y, bsrc = (uint32_t)b1.z;
                            // candidate_geom with the slot test written as "the owner cell folds to the run's cell"
                            const int32_t ox = max(ra.sx, b0.x), oy = max(ra.sy, b0.y);
                            const bool first_hi = ra.gid > bgid;
                            const uint32_t mask_hi = first_hi ? ra.mask : bmask, group_lo = ((first_hi ? bkgb : ra.kgb) >> 8) & 31u;
                            const bool geom_ok = (ox < min(aex, bex)) & (oy < min(aey, bey)) & (((uint32_t)ox & fold_x) == cell_x) & (((uint32_t)oy & fold_y) == cell_y) &
                                                 (ox >= a.sv.cd.col_lo) & (ox < a.sv.cd.col_hi) & (((mask_hi >> group_lo) & 1u) != 0u);
                            n_pairs += 1;
                            if (!geom_ok) continue;
                            n_col += 1;
                            bool pass = true;
                            if (!LIST) {
                                const TileRec& rb = s_rec[kb];
                                const bool touch = (ra.r - rb.l >= 0.0) & (ra.t - rb.b >= 0.0) & (rb.r - ra.l >= 0.0) & (rb.t - ra.b >= 0.0);
                                pass = touch | !(ra.dur_key == rb.dur_key);  // swept_filter_pass (narrowphase.cuh)
                            }
                            if (pass) mine_q[my_q++] = first_hi ? make_uint2(ra.src, bsrc) : make_uint2(bsrc, ra.src);
                        }
                    }
                    if (p == p_end) {
                        e += 32;
                        fresh = true;
                    }
                }
                // ---- survivors of the round out: one reservation per warp, consecutive 8-byte items ----
                uint32_t inc = my_q;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += u;
                }
                const uint32_t n_q = __shfl_sync(0xffffffffu, inc, 31);
                if (n_q) {
                    unsigned long long at = 0;
                    if (lane == 0) at = atomicAdd(&a.sv.ctr->n_surv, (unsigned long long)n_q);
                    at = __shfl_sync(0xffffffffu, at, 0) + (inc - my_q);
#pragma unroll
                    for (uint32_t j = 0; j < kTileIter; ++j)
                        if (j < my_q && at + j < a.surv.cap) a.surv.items[at + j] = mine_q[j];
                }
                if (!__any_sync(0xffffffffu, e < n_ent)) break;
            }
            __syncwarp();
        }
        __syncwarp();
    }
    const unsigned long long col_tot = block_sum(n_col), ent_tot = block_sum((unsigned long long)n_entries), cell_tot = block_sum((unsigned long long)n_cells),
                             pair_tot = block_sum(n_pairs);
    if (threadIdx.x == 0 && !LIST) {
        if (col_tot) atomicAdd(&a.sv.ctr->n_collide, col_tot);
        if (ent_tot) atomicAdd(&a.sv.ctr->n_entries, ent_tot);
        if (cell_tot) atomicAdd(&a.sv.ctr->n_cells, cell_tot);
        if (pair_tot) atomicAdd(&a.sv.ctr->n_tile_pairs, pair_tot);
    }
}

// The solve stage of the tile path: the survivors (collide_time), then the overlapping pairs (separate_time; their records
// are gathered like in k_solve), then k_solve's epilogue: block minima, counters, and the last block reduces every partial
// minimum of the step and publishes the result.
// LIST: the survivors that are not overlapping already are listed as (gid hi, gid lo); nothing else happens.
template <bool LIST>
__global__ void __launch_bounds__(kThreads, CB200_LB_SOLVE) k_tile_finish(const TileSolveArgs ta, uint2* list_out, unsigned long long* n_list, unsigned long long cap_list) {
    const SolveArgs& a = ta.sv;
    __shared__ PairEvent s_ev[kSolveBuf];
    __shared__ uint32_t s_n;
    __shared__ unsigned long long s_base;
    MinKey best{kKeyMax, kKeyMax};
    uint32_t n_sep = 0;
    int32_t n_col = 0;  // net: minus the overlapping pairs that are candidates (the tile kernel counted every pair passing candidate_geom)
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    const uint32_t n_surv = (uint32_t)min(a.ctr->n_surv, (unsigned long long)ta.surv.cap);
    const uint32_t total = n_surv + (LIST ? 0u : (a.n_ov_dev ? min(*a.n_ov_dev, a.n_ov) : a.n_ov));
    const uint32_t n_local = a.ctr->n_local;
    const uint32_t stride = gridDim.x * blockDim.x;
    auto flush = [&]() {  // all threads
        __syncthreads();
        const uint32_t n_ev = s_n;
        if (n_ev) {
            if (threadIdx.x == 0) s_base = atomicAdd(&a.ctr->n_pair_events, (unsigned long long)n_ev);
            __syncthreads();
            const unsigned long long pos = s_base;
            for (uint32_t k = threadIdx.x; k < n_ev; k += blockDim.x)
                if (pos + k < a.cap_events) a.events[pos + k] = s_ev[k];
            __syncthreads();
            if (threadIdx.x == 0) s_n = 0;
            __syncthreads();
        }
    };
    uint32_t pending_rounds = 0;
    const double T = a.dyn->T;
    const uint64_t keep = l2_policy_evict_last();
    for (uint32_t base = blockIdx.x * blockDim.x; base < total; base += stride) {
        const uint32_t p = base + threadIdx.x;
        if (p < n_surv) {
            const uint2 it = __ldg(ta.surv.items + p);
            RecHead hh, hl;
            DurHb dh, dl;
            load_source(ta, it.x, T, &hh, &dh);
            load_source(ta, it.y, T, &hl, &dl);
            if (!is_overlapping(a.cd, hh, hl, true)) {
                if (LIST) {
                    const unsigned long long pos = warp_agg_claim(n_list);
                    if (pos < cap_list) list_out[pos] = make_uint2(hh.gid, hl.gid);
                } else {
                    PairEvent ev;
                    if (solve_pair(a, T, hh, dh, hl, dl, true, &ev, &best)) s_ev[atomicAdd(&s_n, 1u)] = ev;
                }
            }
        } else if (p < total) {
            // an overlapping pair (gid hi, gid lo): separate_time (collider.rs:329-339)
            const uint2 pr = a.ov_pairs[p - n_surv];
            uint32_t ia, ib;
            if ((!a.ov_live_check || overlap_contains(a.cd.ov, pr.x, pr.y)) && gid_lookup(a, n_local, pr.x, &ia) && gid_lookup(a, n_local, pr.y, &ib)) {
                RecHead ha, hb;
                DurHb da, db;
                load_rec_keep(a.rec, ia, keep, &ha, &da);
                load_rec_keep(a.rec, ib, keep, &hb, &db);
                const int32_t owner_col = max(ha.sx, hb.sx);
                if (owner_col >= a.col_lo && owner_col < a.col_hi) {
                    n_col -= overlap_pair_is_candidate(a.cd, ha, hb);
                    const DurHb B = advanced(db, 0.0);  // the partner is `other.hitbox_at_time(T)` (collider.rs:478-486)
                    const double delay = separate_time(da, B, a.padding);
                    if (delay != delay) report_error(a.ctr, kFNan, ha.gid);
                    n_sep += 1;
                    const double t = T + delay;
                    if (t < kHighTime) {
                        const uint32_t low = hb.gid;  // Separate: no collide bit
                        s_ev[atomicAdd(&s_n, 1u)] = PairEvent{t, ha.gid, low};
                        const MinKey k{time_key(t), kPairClass | ((uint64_t)ha.gid << 32) | low};
                        if (key_less(k, best)) best = k;
                    }
                }
            }
        }
        if (++pending_rounds == kSolveBuf / kThreads) {
            flush();
            pending_rounds = 0;
        }
    }
    if (LIST) return;
    flush();
    solve_epilogue(a, best, n_sep, n_col);
}

}  // namespace cb200
