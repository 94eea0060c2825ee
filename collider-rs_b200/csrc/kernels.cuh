// The bulk-step kernels: stage A (rebase + solitaire check + swept bounds + IndexRect + cell count),
// cell scatter, pair stage (candidate generation + collide_time + event emit + min-reduce), separate
// stage (overlapping pairs), finalize.  sm_100a, FMA contraction off (-fmad=false).
//
// Reference-equivalent definition of one bulk step at time T:
//     for id in ascending HbId: Collider::internal_update_hitbox(id, Some(vel_id))   src/core/collider.rs:231-250
// which leaves, for every pair whose NEW index rects share a cell, one event created by the higher id
// against the lower id already rebased to T.  Each kernel cites the reference lines whose results it reproduces.
#pragma once
#include <cuda/barrier>

#include "common.cuh"

namespace cb200 {

constexpr int kThreads = 256;

struct StateCols {  // SoA hitbox table; rows are kept in grid-slot order (k_resort_*), label[row] carries the HbId rank
    double *px, *py, *w, *h, *vx, *vy, *rw, *rh, *start, *end, *pub_end;
    uint8_t* kind;
    uint32_t *group, *mask;
    uint8_t* reit;  // 1 = a Reiterate event is queued at `end` (collider.rs:441-447)
};

// Spatial sharding: cell columns are split into contiguous strips, one per rank.  bounds[d] <=
// column < bounds[d+1] belongs to rank d (bounds[0] = INT32_MIN, bounds[world] = INT32_MAX).  A pair is solved by
// the rank whose strip holds the pair's owner cell column max(sx_i, sx_j), which is never left of either rect's
// start: records only ever travel to ranks at or right of their start column.
constexpr int kMaxWorld = 64;
constexpr int kSubLists = 8;               // work sub-lists per kind pair (k_enum_pairs)
constexpr int kWorkLists = 3 * kSubLists;
struct ShardGeom {
    int world, rank;
    int32_t col_lo, col_hi;          // own strip [col_lo, col_hi); one device: the whole i32 range
    int32_t bounds[kMaxWorld + 1];
};

// The two scalars that change from step to step live in device memory, so that the kernel arguments of a step stay the same
// and the whole step can be replayed as one CUDA graph (engine.cu).
struct StepDyn {
    double T, end_scalar;
    unsigned long long tag;  // direct exchange: the step number the flags of this step carry
};

struct StepArgs {
    uint32_t n;
    int rebase;  // 0 = re-run of binning on an already-rebased table (after a buffer grew): skip the advance
    const StepDyn* dyn;
    double cell_width, padding;
    StateCols s;
    const double *nvx, *nvy, *nrw, *nrh, *nend;  // optional new velocity columns (device), null = keep
    const uint32_t* label; // label[row]: rank of the row's HbId in ascending-id order (sharded: the id itself)
    const uint32_t* ord;   // ord[row]: index of the row's hitbox in the caller's (upload-order) arrays; one device: == label
    HbRec* rec;            // one device: rec[i].  Sharded: the visitor table, own records at [0, n), received ones behind
    int4* bink;            // bink[i] = (sx, sy, span, kgb) of rec[i]: all the scatter needs, 16 B instead of a 96-B record
    uint32_t* slot_count;  // one device only (the count is fused here)
    GridGeom geom;
    Counters* ctr;
    MinKey* partial;  // solitaire minima, one per block of the stage-A launches
    uint32_t partial_base;  // first slot of this launch
    uint32_t dbg;           // timing experiments only (CB200_DBG_A): 1 skip the cell count, 2 skip the record store, 4 skip the state write-back
    // sharded only: one outgoing slab of `slab` records PER DESTINATION rank; record 0 of a slab is its header (first word =
    // count).  A record goes into the slab of every other strip its columns touch (usually one neighbour), so a rank
    // receives only what it has to bin.
    HbRec* send;           // [world][slab]
    uint32_t* send_count;  // [world]
    uint32_t slab;
    uint32_t* vmap;        // gid -> record index of this step (direct-addressed; null when there are no overlapping pairs)
    const uint32_t* ov_bits;  // one bit per gid: the hitbox has an overlap — only those are looked up through vmap
};

// Sharded visitor index space: the own records sit at their rows rec[0, n_local = n); the records
// received from rank d sit in rec[recv0 + d * slab + 1 ...] behind that slab's header.
struct VisSpace {
    uint32_t n_own;  // capacity of the own region (= residents)
    uint32_t recv0;  // first record of this step's received slabs (n_own; direct exchange: + parity * world * slab)
    uint32_t slab;   // records per slab including the header
    int world, rank; // slab `rank` (this rank's own slot) is never filled: ignored
};
__device__ __forceinline__ uint32_t slab_count(const HbRec* rec, const VisSpace& vs, int d) {
    return __ldg(reinterpret_cast<const unsigned int*>(rec + vs.recv0 + (size_t)d * vs.slab));
}
// i in [0, n_local + world * slab) -> record index, or false for headers / unused slab tail
__device__ __forceinline__ bool vis_record(const HbRec* rec, const VisSpace& vs, uint32_t n_local, uint32_t i, uint32_t* v) {
    if (i < n_local) {
        *v = i;
        return true;
    }
    const uint32_t j = i - n_local, d = j / vs.slab, k = j % vs.slab;
    if ((int)d == vs.rank || k == 0 || k - 1 >= slab_count(rec, vs, (int)d)) return false;
    *v = vs.recv0 + j;
    return true;
}

__device__ __forceinline__ bool is_finite(double x) { return fabs(x) < cb_inf(); }

// Rust `f64 as i32`: saturating, NaN -> 0 — exactly cvt.rzi.s32.f64.
__device__ __forceinline__ int32_t f64_as_i32(double v) { return __double2int_rz(v); }

__device__ __forceinline__ void store_rec(HbRec* dst, const HbRec& rec) {
    int4* out = reinterpret_cast<int4*>(dst);
    const int4* in = reinterpret_cast<const int4*>(&rec);
#pragma unroll
    for (int k = 0; k < 6; ++k) out[k] = in[k];
}

// ---------------------------------------------------------------------------------------------------------
// Stage A: per hitbox, the part of internal_update_hitbox that touches only that hitbox.
//   rebase        HitboxInfo::pub_hitbox_at_time (collider.rs:488-497), PlacedShape::advance (shape/mod.rs:260-265)
//   install vel   collider.rs:237-240, Hitbox::validate (core/mod.rs:146-162)
//   solitaire     Collider::solitaire_event_check, release build (collider.rs:420-448); Grid::cell_period
//                 (grid.rs:61-72); PlacedBounds::max_edge over the velocity bounds (shape/mod.rs:294-300,
//                 core/mod.rs:109-116); Hitbox::time_until_too_small (core/mod.rs:164-175)
//   DurHitbox     Hitbox::to_dur_hitbox (core/mod.rs:177-187)
//   swept bounds  DurHitbox::bounding_box (dur_hitbox/mod.rs:88-99)
//   IndexRect     Grid::index_bounds (grid.rs:101-113)
//   cell count    first half of the counting sort that replaces Grid::update_area (grid.rs:137-175)   [one device]
//   halo pack     the record goes to every rank whose strip its rect (+1 column) touches                [sharded]
// Reads the SoA columns coalesced, writes the rebased columns back and the 96-B HbRec for the pair stage.
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

// The per-hitbox part of Collider::internal_update_hitbox (collider.rs:231-250), shared by the bulk stage A and by the
// single-hitbox update (single.cuh).  `h` enters as the stored hitbox (shape at its start_time) and leaves rebased to T
// with the new velocity installed; returns the error flags (h is then not to be stored).
struct HbCore {
    double px, py, w, h, vx, vy, rw, rh;  // Hitbox.value / Hitbox.vel
    double start, pub_end;                 // HitboxInfo.start_time / pub_end_time (in: stored; out: T / user end_time)
    double end;                            // out: internal end_time = min(cell period, user end, too-small)  (collider.rs:420-448)
    double dur;                            // out: end - T
    uint32_t kind, group, mask;
    int32_t sx, sy, ex, ey;                // out: IndexRect (valid when binned)
    bool reit, binned;                     // out: Reiterate queued at `end`; entered the grid
};
struct NewVel {
    bool has_vx, has_vy, has_rw, has_rh;
    double vx, vy, rw, rh, end_time;
};
__device__ __forceinline__ uint32_t hb_update_core(HbCore& h, double T, bool rebase, const NewVel& nv, double cell_width, double padding,
                                                   const GridGeom& geom) {
    uint32_t err = 0;
    // pub_hitbox_at_time(T): assert start <= T <= pub_end, then advance by T - start.
    if (!(T >= h.start && T <= h.pub_end)) err |= kFInvalidTime;
    const double dt = T - h.start;
    if (!(dt < kHighTime)) err |= kFHighTime;
    if (rebase) {
        h.px = h.px + h.vx * dt;
        h.py = h.py + h.vy * dt;
        h.w = h.w + h.rw * dt;
        h.h = h.h + h.rh * dt;
    }
    // install the new velocity, validate (core/mod.rs:146-162)
    if (nv.has_vx) h.vx = nv.vx;
    if (nv.has_vy) h.vy = nv.vy;
    if (nv.has_rw) h.rw = nv.rw;
    if (nv.has_rh) h.rh = nv.rh;
    const double user_end = nv.end_time;
    if (!(user_end >= T)) err |= kFEndTime;  // also catches NaN
    if (h.kind == (uint32_t)kCircle && !(h.rw == h.rh)) err |= kFCircleResize;
    if (!(h.w >= padding && h.h >= padding)) err |= kFTooSmall;
    if (!(is_finite(h.px) && is_finite(h.py) && is_finite(h.w) && is_finite(h.h) && is_finite(h.vx) && is_finite(h.vy) && is_finite(h.rw) &&
          is_finite(h.rh)))
        err |= kFNan;

    // solitaire_event_check (release build)
    const bool has_group = h.group != kGroupNone;
    double period = cb_inf();
    if (has_group) {
        const double speed = fmax(fmax(fabs(h.vx - h.rw * 0.5), fabs(h.vy - h.rh * 0.5)), fmax(fabs(h.vx + h.rw * 0.5), fabs(h.vy + h.rh * 0.5)));
        if (!(speed <= 0.0)) period = cell_width / speed;
    }
    double r_time = T + period;
    bool reit = true;
    if (user_end < r_time) {
        r_time = user_end;
        reit = false;
    }
    {
        const double min_size = padding * 0.9;
        if (!(h.w > min_size && h.h > min_size)) err |= kFTooSmall;
        double tus = cb_inf();
        if (h.rw < 0.0) tus = fmin(tus, (min_size - h.w) / h.rw);
        if (h.rh < 0.0) tus = fmin(tus, (min_size - h.h) / h.rh);
        const double small_end = T + tus;
        if (small_end < r_time) {
            r_time = small_end;
            reit = false;
        }
    }
    reit = reit && (r_time < kHighTime);  // EventManager::new_event_key drops time >= 1e50 (events.rs:156-159)
    const double dur = r_time - T;

    // swept bounds + IndexRect (only hitboxes with a group enter the grid, collider.rs:328)
    h.sx = h.sy = h.ex = h.ey = 0;
    h.binned = false;
    if (has_group && err == 0) {
        DurHb d;
        d.px = h.px; d.py = h.py; d.w = h.w; d.h = h.h; d.vx = h.vx; d.vy = h.vy; d.rw = h.rw; d.rh = h.rh; d.dur = dur; d.kind = (int)h.kind;
        const Box4 bb = swept_bounds(d, dur);
        if (!bb.ok) {
            err |= (dur < kHighTime) ? kFNan : kFHighTime;
        } else {
            const double min_x = bb.cx - bb.w * 0.5, max_x = bb.cx + bb.w * 0.5;
            const double min_y = bb.cy - bb.h * 0.5, max_y = bb.cy + bb.h * 0.5;
            // Grid::index_bounds divides by cell_width (grid.rs:101-113).  When cell_width is a power of two the quotient is an
            // exact scaling, so multiplying by the (exact) reciprocal gives the same bits for every input — including subnormal
            // results, which both operations round from the same real value — at a fraction of the cost of four f64 divisions.
            const long long cw_bits = __double_as_longlong(cell_width);
            const bool cw_pow2 = (cw_bits & 0x000fffffffffffffLL) == 0 && cw_bits > (2LL << 52) && cw_bits < (2044LL << 52);
            const double cw_inv = __longlong_as_double((2046LL << 52) - cw_bits);
            const double qx0 = cw_pow2 ? min_x * cw_inv : min_x / cell_width, qy0 = cw_pow2 ? min_y * cw_inv : min_y / cell_width;
            const double qx1 = cw_pow2 ? max_x * cw_inv : max_x / cell_width, qy1 = cw_pow2 ? max_y * cw_inv : max_y / cell_width;
            const int32_t sx = f64_as_i32(floor(qx0));
            const int32_t sy = f64_as_i32(floor(qy0));
            const int32_t ex = max(f64_as_i32(ceil(qx1)), (int32_t)((uint32_t)sx + 1u));
            const int32_t ey = max(f64_as_i32(ceil(qy1)), (int32_t)((uint32_t)sy + 1u));
            const int64_t nx = (int64_t)ex - sx, ny = (int64_t)ey - sy;
            if (nx <= 0 || ny <= 0) {
                err |= kFEmptyRect;  // IndexRect::new (index_rect.rs:26-29)
            } else if (nx > (1ll << geom.lw) || ny > (1ll << geom.lh) || nx > 0xffff || ny > 0xffff) {
                err |= kFRectSpan;
            } else {
                h.sx = sx; h.sy = sy; h.ex = ex; h.ey = ey;
                h.binned = true;
            }
        }
    }
    h.start = T;
    h.pub_end = user_end;
    h.end = r_time;
    h.dur = dur;
    h.reit = reit;
    return err;
}
__device__ __forceinline__ HbRec make_rec(const HbCore& h, uint32_t gid) {
    HbRec rec;
    rec.sx = h.sx; rec.sy = h.sy;
    rec.span = h.binned ? ((uint32_t)(h.ex - h.sx) | ((uint32_t)(h.ey - h.sy) << 16)) : 0u;
    rec.gid = gid;
    rec.dur = h.dur;
    rec.kgb = (h.kind & 1u) | (h.binned ? 2u : 0u) | ((h.group != kGroupNone ? (h.group & 0xffu) : 0xffu) << 8);
    rec.mask = h.mask;
    rec.px = h.px; rec.py = h.py; rec.w = h.w; rec.h = h.h;
    rec.vx = h.vx; rec.vy = h.vy; rec.rw = h.rw; rec.rh = h.rh;
    return rec;
}

// Everything of stage A after the row's stored state has been read into `h`.
template <bool SHARDED>
__device__ __forceinline__ HbRec stage_a_row(const StepArgs& a, const ShardGeom& sh, uint32_t i, HbCore& h, uint32_t gid, uint32_t o, MinKey& best,
                                             unsigned long long& n_sol) {
    // (the caller's arrays are in upload order: gathered through ord; a re-run reads the stored pub_end_time by row)
    NewVel nv;
    nv.has_vx = a.nvx != nullptr; nv.has_vy = a.nvy != nullptr; nv.has_rw = a.nrw != nullptr; nv.has_rh = a.nrh != nullptr;
    nv.vx = nv.has_vx ? __ldg(a.nvx + o) : 0.0;
    nv.vy = nv.has_vy ? __ldg(a.nvy + o) : 0.0;
    nv.rw = nv.has_rw ? __ldg(a.nrw + o) : 0.0;
    nv.rh = nv.has_rh ? __ldg(a.nrh + o) : 0.0;
    const StepDyn dyn = *a.dyn;
    nv.end_time = a.nend ? __ldg(a.nend + (a.rebase ? o : i)) : dyn.end_scalar;
    const uint32_t err = hb_update_core(h, dyn.T, a.rebase != 0, nv, a.cell_width, a.padding, a.geom);
    if (h.binned && !(a.dbg & 1u)) {
        // cell count (sharded: only the cells inside the own strip)
        const int32_t x0 = SHARDED ? max(h.sx, sh.col_lo) : h.sx, x1 = SHARDED ? min(h.ex, sh.col_hi) : h.ex;
        for (int32_t x = x0; x < x1; ++x)
            for (int32_t y = h.sy; y != h.ey; ++y) atomicAdd(&a.slot_count[a.geom.slot(x, y)], 1u);
    }
    if (err) report_error(a.ctr, err, gid);

    // write back: state columns + pair-stage record
    if (!(a.dbg & 4u)) {
    a.s.px[i] = h.px; a.s.py[i] = h.py; a.s.w[i] = h.w; a.s.h[i] = h.h;
    if (a.nvx) a.s.vx[i] = h.vx;
    if (a.nvy) a.s.vy[i] = h.vy;
    if (a.nrw) a.s.rw[i] = h.rw;
    if (a.nrh) a.s.rh[i] = h.rh;
    a.s.start[i] = h.start;
    a.s.end[i] = h.end;
    a.s.pub_end[i] = h.pub_end;
    a.s.reit[i] = h.reit ? 1 : 0;
    }

    // the record stays at its row; a sharded rank bins only the cells inside its own strip (above and in k_scatter)
    const HbRec rec = make_rec(h, gid);
    if (!(a.dbg & 2u)) a.bink[i] = make_int4(rec.sx, rec.sy, (int)rec.span, (int)rec.kgb);
    if (SHARDED && a.vmap && ((__ldg(a.ov_bits + (gid >> 5)) >> (gid & 31u)) & 1u)) a.vmap[gid] = i;
    if (SHARDED && h.binned) {
        // The record is needed by every rank whose strip intersects columns [sx, ex + 1): the extra column keeps
        // a partner that is within `padding` across a cell line (an overlapping pair whose rects do not
        // intersect) visible.  Every strip other than the own one -> that rank's halo slab (sharding.py strips_touched).
        const int64_t hi_col = (int64_t)h.ex + 1;
        const bool leaves = (int64_t)rec.sx < (int64_t)sh.col_lo || hi_col > (int64_t)sh.col_hi;
        if (leaves) {
            for (int d = 0; d < sh.world; ++d) {
                if (d == sh.rank || !((int64_t)sh.bounds[d + 1] > (int64_t)rec.sx && (int64_t)sh.bounds[d] < hi_col)) continue;
                const uint32_t pos = atomicAdd(a.send_count + d, 1u);
                if (pos + 1 < a.slab) store_rec(a.send + (size_t)d * a.slab + 1 + pos, rec);
                else a.ctr->overflow_send = 1u;
            }
        }
    }
    if (h.reit) {
        ++n_sol;
        const MinKey k{time_key(h.end), (uint64_t)gid};
        if (key_less(k, best)) best = k;
    }
    return rec;
}

// The 96-byte records of a warp's 32 consecutive rows, written as 3 KB of contiguous 16-byte chunks: each lane parks its
// record in shared memory (stride 112 B: conflict-free 128-bit accesses) and then stores chunks lane, lane + 32, ... of the
// warp's block — full sectors, half the store transactions of 32 lanes writing 16 bytes each at a 96-byte stride (the
// record store was 25 of stage A's 62 us: profiles/stage_a_ablation.py).
constexpr int kRecStageStride = 7;  // int4 slots per record in shared memory (6 used)
__device__ __forceinline__ void store_recs_warp(HbRec* rec_base /* rec + first row of the warp */, uint32_t n_valid, const HbRec& mine, int4* s_warp /* [32 * 7] */) {
    const int lane = threadIdx.x & 31;
    const int4* in = reinterpret_cast<const int4*>(&mine);
#pragma unroll
    for (int k = 0; k < 6; ++k) s_warp[lane * kRecStageStride + k] = in[k];
    __syncwarp();
    int4* out = reinterpret_cast<int4*>(rec_base);
    const uint32_t n_chunks = n_valid * 6u;
    const uint64_t keep = l2_policy_evict_last();
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        const uint32_t c = (uint32_t)lane + 32u * j;
        if (c < n_chunks) st_int4_keep(out + c, s_warp[(c / 6u) * kRecStageStride + (c % 6u)], keep);
    }
    __syncwarp();
}

#ifndef CB200_LB_A
#define CB200_LB_A 4
#endif
// Plain variant: every thread loads its row's columns itself (also the tail of the staged variant).
template <bool SHARDED>
__global__ void __launch_bounds__(kThreads, CB200_LB_A) k_stage_a(const StepArgs a, const ShardGeom sh, uint32_t first_row) {
    __shared__ int4 s_rec[kThreads / 32][32 * kRecStageStride];
    MinKey best{kKeyMax, kKeyMax};
    unsigned long long n_sol = 0;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t base = first_row + blockIdx.x * blockDim.x; base < a.n; base += stride) {
        const uint32_t i = base + threadIdx.x;
        HbRec rec;
        if (i < a.n) {
            HbCore h;
            h.px = a.s.px[i]; h.py = a.s.py[i]; h.w = a.s.w[i]; h.h = a.s.h[i];
            h.vx = a.s.vx[i]; h.vy = a.s.vy[i]; h.rw = a.s.rw[i]; h.rh = a.s.rh[i];
            h.start = a.s.start[i]; h.pub_end = a.s.pub_end[i];
            h.kind = a.s.kind[i]; h.group = a.s.group[i]; h.mask = a.s.mask[i];
            rec = stage_a_row<SHARDED>(a, sh, i, h, a.label[i], a.ord[i], best, n_sol);
        }
        const uint32_t warp_row = base + (threadIdx.x & ~31u);
        if (warp_row < a.n && !(a.dbg & 2u)) store_recs_warp(a.rec + warp_row, min(32u, a.n - warp_row), rec, s_rec[threadIdx.x >> 5]);
    }
    best = block_min(best);
    const unsigned long long tot = block_sum(n_sol);
    if (threadIdx.x == 0) {
        a.partial[a.partial_base + blockIdx.x] = best;
        if (tot) atomicAdd(&a.ctr->n_solitaire, tot);
    }
}

// Staged variant: the 14 input columns of a 256-row tile are brought into shared memory by bulk async copies (TMA,
// cp.async.bulk, one elected thread, completion on an mbarrier), two tiles deep, so the bytes in flight per SM do not
// depend on how many threads happen to be in their load phase (the plain variant sat at 55 % of the HBM peak with every
// warp stalled on its first use of a loaded value, profiles/r1c_step_kernels.txt).  Persistent blocks; full tiles only.
struct StageATile {
    double col[10][kThreads];  // px py w h vx vy rw rh start pub_end
    uint32_t group[kThreads], mask[kThreads], label[kThreads], ord[kThreads];
    uint8_t kind[kThreads];
};
constexpr uint32_t kStageATileBytes = 10 * kThreads * 8 + 4 * kThreads * 4 + kThreads;

template <bool SHARDED>
__global__ void __launch_bounds__(kThreads, CB200_LB_A) k_stage_a_tma(const StepArgs a, const ShardGeom sh, uint32_t n_tiles) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    StageATile* tile = reinterpret_cast<StageATile*>(smem_raw);  // [2]
    __shared__ int4 s_rec[kThreads / 32][32 * kRecStageStride];
    __shared__ cuda::barrier<cuda::thread_scope_block> bar[2];
    if (threadIdx.x == 0) {
        init(&bar[0], blockDim.x);
        init(&bar[1], blockDim.x);
        cuda::device::experimental::fence_proxy_async_shared_cta();
    }
    __syncthreads();
    const double* src[10] = {a.s.px, a.s.py, a.s.w, a.s.h, a.s.vx, a.s.vy, a.s.rw, a.s.rh, a.s.start, a.s.pub_end};
    auto issue = [&](uint32_t t, int stage) {  // thread 0 only
        const size_t row0 = (size_t)t * kThreads;
        StageATile& dst = tile[stage];
#pragma unroll
        for (int k = 0; k < 10; ++k) cuda::device::memcpy_async_tx(dst.col[k], src[k] + row0, cuda::aligned_size_t<16>(kThreads * 8), bar[stage]);
        cuda::device::memcpy_async_tx(dst.group, a.s.group + row0, cuda::aligned_size_t<16>(kThreads * 4), bar[stage]);
        cuda::device::memcpy_async_tx(dst.mask, a.s.mask + row0, cuda::aligned_size_t<16>(kThreads * 4), bar[stage]);
        cuda::device::memcpy_async_tx(dst.label, a.label + row0, cuda::aligned_size_t<16>(kThreads * 4), bar[stage]);
        cuda::device::memcpy_async_tx(dst.ord, a.ord + row0, cuda::aligned_size_t<16>(kThreads * 4), bar[stage]);
        cuda::device::memcpy_async_tx(dst.kind, a.s.kind + row0, cuda::aligned_size_t<16>(kThreads), bar[stage]);
    };
    MinKey best{kKeyMax, kKeyMax};
    unsigned long long n_sol = 0;
    cuda::barrier<cuda::thread_scope_block>::arrival_token token[2];
    auto arm = [&](uint32_t t, int stage) {  // all threads: arrive on the stage's barrier; thread 0 also posts the copies
        if (threadIdx.x == 0) {
            issue(t, stage);
            token[stage] = cuda::device::barrier_arrive_tx(bar[stage], 1, kStageATileBytes);
        } else {
            token[stage] = bar[stage].arrive();
        }
    };
    uint32_t t = blockIdx.x;
    int stage = 0;
    if (t < n_tiles) arm(t, 0);
    for (; t < n_tiles; t += gridDim.x, stage ^= 1) {
        const uint32_t t_next = t + gridDim.x;
        if (t_next < n_tiles) arm(t_next, stage ^ 1);  // (that stage was released by the barrier at the end of the previous iteration)
        bar[stage].wait(std::move(token[stage]));
        const StageATile& src_tile = tile[stage];
        const uint32_t i = t * kThreads + threadIdx.x;
        HbCore h;
        h.px = src_tile.col[0][threadIdx.x]; h.py = src_tile.col[1][threadIdx.x]; h.w = src_tile.col[2][threadIdx.x]; h.h = src_tile.col[3][threadIdx.x];
        h.vx = src_tile.col[4][threadIdx.x]; h.vy = src_tile.col[5][threadIdx.x]; h.rw = src_tile.col[6][threadIdx.x]; h.rh = src_tile.col[7][threadIdx.x];
        h.start = src_tile.col[8][threadIdx.x]; h.pub_end = src_tile.col[9][threadIdx.x];
        h.kind = src_tile.kind[threadIdx.x]; h.group = src_tile.group[threadIdx.x]; h.mask = src_tile.mask[threadIdx.x];
        const uint32_t gid = src_tile.label[threadIdx.x], o = src_tile.ord[threadIdx.x];
        __syncthreads();  // every thread has its row in registers: the stage may be refilled
        const HbRec rec = stage_a_row<SHARDED>(a, sh, i, h, gid, o, best, n_sol);
        if (!(a.dbg & 2u)) store_recs_warp(a.rec + (i & ~31u), 32u, rec, s_rec[threadIdx.x >> 5]);
    }
    best = block_min(best);
    const unsigned long long tot = block_sum(n_sol);
    if (threadIdx.x == 0) {
        a.partial[a.partial_base + blockIdx.x] = best;
        if (tot) atomicAdd(&a.ctr->n_solitaire, tot);
    }
}

// Sharded: publish the record counts in the headers of the halo slabs before the exchange (one thread per destination).
__global__ void k_slab_header(HbRec* send, const uint32_t* __restrict__ send_count, uint32_t slab, int world) {
    const int d = threadIdx.x;
    if (d < world) *reinterpret_cast<unsigned int*>(send + (size_t)d * slab) = min(send_count[d], slab - 1u);
}

// Sharded, after the exchange: count the strip cells of the RECEIVED records (the own ones were counted in stage A;
// COUNT_OWN re-counts them too — used when a step is re-run after a buffer grew), and index them by gid for the
// overlapping-pair lookups.
template <bool COUNT_OWN>
__global__ void __launch_bounds__(kThreads) k_index_visitors(const HbRec* __restrict__ rec, int4* bink, VisSpace vs, const Counters* ctr, uint32_t* slot_count,
                                                             GridGeom geom, int32_t col_lo, int32_t col_hi, uint32_t* vmap, const uint32_t* __restrict__ ov_bits) {
    const uint32_t n_local = ctr->n_local;
    const uint32_t total = n_local + (uint32_t)vs.world * vs.slab;
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t first = COUNT_OWN ? 0u : n_local;
    for (uint32_t i = first + blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        uint32_t v;
        if (!vis_record(rec, vs, n_local, i, &v)) continue;
        const RecHead h = load_head(rec, v);
        if (i >= n_local) bink[v] = make_int4(h.sx, h.sy, (int)((uint32_t)(h.ex - h.sx) | ((uint32_t)(h.ey - h.sy) << 16)), (int)h.kgb);
        if (!(h.kgb & 2u)) continue;
        const int32_t x0 = max(h.sx, col_lo), x1 = min(h.ex, col_hi);
        for (int32_t x = x0; x < x1; ++x)
            for (int32_t y = h.sy; y != h.ey; ++y) atomicAdd(&slot_count[geom.slot(x, y)], 1u);
        if (vmap && i >= n_local && ((__ldg(ov_bits + (h.gid >> 5)) >> (h.gid & 31u)) & 1u)) vmap[h.gid] = v;  // (own records: stage A)
    }
}

// ---------------------------------------------------------------------------------------------------------
// Direct halo exchange over peer memory (NVLink / NVSwitch): every rank's visitor table and mailbox are mapped into its
// peers (cudaIpc or same-process pointers), so the exchange is two tiny kernels instead of two NCCL collectives —
//   k_push_halo   copies the rank's halo slab for peer d (header + the records that are actually there) straight into slot
//                 `rank` of that peer's receive area for this step's parity, then raises flag_halo[rank] = tag on the peer
//   k_wait_halo   spins (on local memory) until every peer has raised its flag for this step
// and the same pattern for the 24-byte next-event keys after the solve stage (k_push_key / k_reduce_keys).  Slab areas and
// key rows are double-buffered by step parity; a rank can run at most one step ahead of a peer, because its next back half
// needs that peer's next flag.  tag = the step number.  A rank whose buffers overflowed in the step marks its key incomplete;
// every rank then sees complete == 0 in the reduced key and the host repeats the key exchange after the re-run.
struct Mailbox {
    unsigned long long flag_halo[kMaxWorld];
    unsigned long long flag_key[kMaxWorld];
    unsigned long long keys[2][kMaxWorld][4];  // [parity][rank] = {time key, tie, complete, -}
    unsigned long long abort;                  // raised by a peer that gave up on its step (k_raise_abort): waiters stop spinning
};
struct PeerTable {
    HbRec* recv[kMaxWorld];     // peer d's receive area for parity 0, slot 0 (parity 1 follows world * slab records later)
    Mailbox* mail[kMaxWorld];
};

__global__ void __launch_bounds__(1024) k_push_halo(const HbRec* __restrict__ send, PeerTable peers, int world, int rank, uint32_t slab, int parity,
                                                    const StepDyn* dyn) {
    const unsigned long long tag = dyn->tag;
    const int d = blockIdx.x + (blockIdx.x >= (unsigned)rank ? 1 : 0);  // blocks enumerate the peers
    if (d >= world) return;
    send += (size_t)d * slab;  // the slab addressed to peer d
    const uint32_t count = min(*reinterpret_cast<const unsigned int*>(send), slab - 1u);
    const int4* src = reinterpret_cast<const int4*>(send);
    int4* dst = reinterpret_cast<int4*>(peers.recv[d] + ((size_t)parity * world + rank) * slab);
    const uint32_t n16 = (count + 1u) * 6u;  // header + records, 96 B = 6 x 16 B each
    for (uint32_t i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = src[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        *reinterpret_cast<volatile unsigned long long*>(&peers.mail[d]->flag_halo[rank]) = tag;
        __threadfence_system();
    }
}
// Waiting for a peer is bounded: a rank that stopped stepping (an error on its side, a Python exception, a crashed process)
// must not leave its peers inside a kernel for ever.  The wait ends when the flag arrives, when any peer raised this rank's
// abort word, or after timeout_ns of %globaltimer — the last two set kFPeerLost in the step's error flags, which the host
// reports as a failed step.
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ bool wait_flag(const unsigned long long* flag, unsigned long long tag, const Mailbox* own, unsigned long long timeout_ns) {
    const unsigned long long t0 = global_timer_ns();
    unsigned int spins = 0;
    while (*reinterpret_cast<const volatile unsigned long long*>(flag) < tag) {
        __nanosleep(100);
        if ((++spins & 63u) == 0u) {
            if (*reinterpret_cast<const volatile unsigned long long*>(&own->abort) != 0ull) return false;
            if (global_timer_ns() - t0 > timeout_ns) return false;
        }
    }
    return true;
}
__global__ void k_wait_flags(const Mailbox* own, int world, int rank, const StepDyn* dyn, Counters* ctr, unsigned long long timeout_ns) {
    const unsigned long long tag = dyn->tag;
    const int d = threadIdx.x;
    if (d >= world || d == rank) return;
    if (!wait_flag(&own->flag_halo[d], tag, own, timeout_ns)) report_error(ctr, kFPeerLost, (uint64_t)d);
    __threadfence_system();
}
__global__ void k_push_key(const MinKey* __restrict__ local_key, const Counters* ctr, const uint32_t* work_count, uint32_t seg_cap, uint32_t cap_entries,
                           uint32_t cap_events, PeerTable peers, Mailbox* own, int world, int rank, int parity, const StepDyn* dyn) {
    const unsigned long long tag = dyn->tag;
    const int d = threadIdx.x;
    if (d >= world) return;
    // complete = no buffer of this rank overflowed in this step (else the step is re-run and the keys exchanged again)
    unsigned long long complete = 1;
    if (ctr->overflow_entries || ctr->overflow_send || ctr->n_entries > cap_entries || ctr->n_pair_events > cap_events) complete = 0;
    for (int l = 0; l < kWorkLists; ++l)
        if (work_count[l] > seg_cap) complete = 0;
    Mailbox* m = d == rank ? own : peers.mail[d];
    volatile unsigned long long* row = m->keys[parity][rank];
    row[0] = local_key->t;
    row[1] = local_key->tie;
    row[2] = complete;
    __threadfence_system();
    if (d != rank) {
        *reinterpret_cast<volatile unsigned long long*>(&m->flag_key[rank]) = tag;
        __threadfence_system();
    }
}
struct GlobalKey {
    unsigned long long t, tie, complete, tag, lost;  // lost: a peer never delivered its key (timeout / abort)
};
__global__ void k_reduce_keys(const Mailbox* own, int world, int rank, int parity, const StepDyn* dyn, GlobalKey* out /* mapped host */,
                              unsigned long long timeout_ns) {
    __shared__ unsigned long long s_t[kMaxWorld], s_tie[kMaxWorld], s_ok[kMaxWorld];
    __shared__ unsigned int s_lost;
    const unsigned long long tag = dyn->tag;
    const int d = threadIdx.x;
    if (d == 0) s_lost = 0u;
    __syncthreads();
    if (d < world) {
        bool arrived = true;
        if (d != rank) arrived = wait_flag(&own->flag_key[d], tag, own, timeout_ns);
        __threadfence_system();
        const volatile unsigned long long* row = own->keys[parity][d];
        s_t[d] = arrived ? row[0] : kKeyMax;
        s_tie[d] = arrived ? row[1] : kKeyMax;
        s_ok[d] = arrived ? row[2] : 0ull;
        if (!arrived) atomicOr(&s_lost, 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        MinKey best{kKeyMax, kKeyMax};
        unsigned long long ok = 1;
        for (int k = 0; k < world; ++k) {
            const MinKey c{s_t[k], s_tie[k]};
            if (key_less(c, best)) best = c;
            ok &= s_ok[k];
        }
        out->t = best.t;
        out->tie = best.tie;
        out->complete = ok;
        out->lost = s_lost;
        out->tag = tag;
        __threadfence_system();
    }
}
// A rank that fails its step tells every peer: their waits end with kFPeerLost instead of spinning until the timeout.
__global__ void k_raise_abort(PeerTable peers, int world, int rank) {
    const int d = threadIdx.x;
    if (d >= world || d == rank) return;
    *reinterpret_cast<volatile unsigned long long*>(&peers.mail[d]->abort) = 1ull;
    __threadfence_system();
}

// ---------------------------------------------------------------------------------------------------------
// Step prologue: one kernel zeroes everything a step accumulates into (slot table, counters, halo count).
__global__ void __launch_bounds__(kThreads) k_step_zero(uint4* slots4, uint32_t n4, Counters* ctr, uint32_t* send_count, uint32_t* work_count,
                                                        uint32_t n_work_lists, uint32_t n_local, uint32_t n_send) {
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) slots4[i] = z;
    if (blockIdx.x == 0) {
        uint32_t* w = reinterpret_cast<uint32_t*>(ctr);
        static_assert(sizeof(Counters) / 4 <= kThreads, "one thread per counter word");
        if (threadIdx.x < sizeof(Counters) / 4) w[threadIdx.x] = 0u;
        if (work_count)
            for (uint32_t l = threadIdx.x; l < n_work_lists; l += blockDim.x) work_count[l] = 0u;
        __syncthreads();
        if (threadIdx.x == 0) {
            ctr->err_slot = ~0ull;
            ctr->n_local = n_local;  // sharded: the own records occupy rows [0, n) of the visitor table
        }
        if (send_count)
            for (uint32_t d = threadIdx.x; d < n_send; d += blockDim.x) send_count[d] = 0u;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Slot offsets: the counting sort needs, for every slot, the start of its run in the entry array — but not a
// particular ORDER of the runs.  So instead of a global prefix sum (whose look-back / wait chain measured 2x the
// cost of the data movement, profiles/r1c_scan_variants.txt) every block scans its own chunk of kScanChunk slots in
// registers and takes the chunk's base from one atomicAdd on a global cursor.  One read and one write of the table,
// no inter-block waiting.  Runs of one chunk are contiguous and in slot order; chunks land in arrival order, and
// chunk_base[] remembers where each one starts (the run of slot s starts at chunk_base[s / kScanChunk] when s is the
// first slot of its chunk, else where the run of slot s - 1 ends).
constexpr int kScanItems = 8;
constexpr int kScanChunk = kThreads * kScanItems;  // 2048 slots
static_assert(kScanChunk == 2048, "chunk size is part of the grid view (slot_run_start)");

__global__ void __launch_bounds__(kThreads) k_scan_slots(uint32_t* __restrict__ data, unsigned long long* cursor, uint32_t* __restrict__ chunk_base,
                                                         unsigned long long* n_cells) {
    __shared__ uint32_t s_prefix;
    __shared__ uint32_t s_warp[kThreads / 32];
    const uint32_t chunk = blockIdx.x;
    uint4* p = reinterpret_cast<uint4*>(data + (size_t)chunk * kScanChunk + (size_t)threadIdx.x * kScanItems);
    uint32_t v[kScanItems];
#pragma unroll
    for (int k = 0; k < kScanItems / 4; ++k) {
        const uint4 q = p[k];
        v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
    }
    uint32_t sum = 0, nz = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        sum += v[k];
        nz += v[k] != 0;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
        const uint32_t x = s_warp[w];
        before += w < warp ? x : 0u;
        total += x;
    }
    if (threadIdx.x == 0) {
        const uint32_t base = (uint32_t)atomicAdd(cursor, (unsigned long long)total);
        s_prefix = base;
        chunk_base[chunk] = base;
    }
    const unsigned long long nz_block = block_sum(nz);  // (its barriers also publish s_prefix)
    if (threadIdx.x == 0 && nz_block && n_cells) atomicAdd(n_cells, nz_block);
    uint32_t run = s_prefix + before + inc - sum;
#pragma unroll
    for (int k = 0; k < kScanItems / 4; ++k) {
        uint4 q;
        q.x = run; run += v[4 * k];
        q.y = run; run += v[4 * k + 1];
        q.z = run; run += v[4 * k + 2];
        q.w = run; run += v[4 * k + 3];
        p[k] = q;
    }
}

// After the scatter, slot_end[s] = end of slot s's run.  Start of the run (see k_scan_slots).
__device__ __forceinline__ uint32_t slot_run_start(const uint32_t* __restrict__ slot_end, const uint32_t* __restrict__ chunk_base, uint32_t s) {
    return (s % kScanChunk) ? slot_end[s - 1] : chunk_base[s / kScanChunk];
}

// ---------------------------------------------------------------------------------------------------------
// Table order.  Rows of the hitbox table are kept sorted by the grid slot of their centre cell, so that the
// atomics of the binning, the entry writes of the scatter and the record gathers of the solve stage all land
// in nearby cache lines.  The order only affects speed: ids travel in label[row], every result is keyed by label.
// A re-sort is a counting sort over the slot table (count -> k_scan_slots -> rank) followed by one gather of all
// columns into the alternate buffers; the engine runs it after upload and every few steps (drift is slow).
__global__ void __launch_bounds__(kThreads) k_resort_count(uint32_t n, const double* __restrict__ px, const double* __restrict__ py, double cell_width,
                                                           GridGeom geom, uint32_t* slot_count, uint32_t* key) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = geom.slot(f64_as_i32(floor(px[i] / cell_width)), f64_as_i32(floor(py[i] / cell_width)));
    key[i] = s;
    atomicAdd(&slot_count[s], 1u);
}
__global__ void __launch_bounds__(kThreads) k_resort_rank(uint32_t n, const uint32_t* __restrict__ key, uint32_t* cursor, uint32_t* src_row) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    src_row[atomicAdd(&cursor[key[i]], 1u)] = i;
}
struct ResortCols {
    const double* in[11];
    double* out[11];
    const uint8_t *kind_in, *reit_in;
    uint8_t *kind_out, *reit_out;
    const uint32_t *group_in, *mask_in, *label_in, *ord_in;  // ord_in null: ord == label (one device)
    uint32_t *group_out, *mask_out, *label_out, *ord_out;
    uint32_t* row_of;  // label -> new row (null: not maintained)
};
__global__ void __launch_bounds__(kThreads) k_resort_gather(uint32_t n, const uint32_t* __restrict__ src_row, const ResortCols c) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = src_row[i];
#pragma unroll
    for (int k = 0; k < 11; ++k) c.out[k][i] = c.in[k][s];
    c.kind_out[i] = c.kind_in[s];
    c.reit_out[i] = c.reit_in[s];
    c.group_out[i] = c.group_in[s];
    c.mask_out[i] = c.mask_in[s];
    const uint32_t l = c.label_in[s];
    c.label_out[i] = l;
    if (c.ord_in) c.ord_out[i] = c.ord_in[s];
    if (c.row_of) c.row_of[l] = i;
}
// Column in the caller's order for the host (download_state): out[ord[row]] = col[row].
__global__ void __launch_bounds__(kThreads) k_unpermute(uint32_t n, const uint32_t* __restrict__ ord, const double* __restrict__ col, double* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[ord[i]] = col[i];
}
__global__ void __launch_bounds__(kThreads) k_iota_rows(uint32_t n, uint32_t* label, uint32_t* row_of) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (label) label[i] = i;
    if (row_of) row_of[i] = i;
}

// ---------------------------------------------------------------------------------------------------------
// Scatter: second half of the counting sort.  offsets[s] enters as the start of slot s and is used as the
// slot's cursor.  A cell entry is 8 bytes — the key (slot, group, kind) the enumeration stage streams, and the index
// of the hitbox's record, where everything else about it lives.
struct alignas(8) CellEntry {
    uint32_t key;   // grid slot (26 bits) | group << 26 (5 bits) | rect << 31
    uint32_t vidx;  // index of the hitbox's record in the visitor table (one device: the table row)
};
static_assert(sizeof(CellEntry) == 8, "CellEntry is one 8-byte word");
constexpr uint32_t kSlotMask = (1u << 26) - 1u;
constexpr uint32_t kEntryKindBit = 0x80000000u;  // key bit 31: the hitbox is a rect

template <bool SHARDED>
__global__ void __launch_bounds__(kThreads) k_scatter(uint32_t n, const HbRec* __restrict__ rec, const int4* __restrict__ bink, VisSpace vs, uint32_t* offsets,
                                                      CellEntry* entries, uint32_t cap_entries, GridGeom geom, int32_t col_lo, int32_t col_hi, Counters* ctr) {
    const uint32_t n_local = SHARDED ? ctr->n_local : n;
    const uint32_t total = SHARDED ? n_local + (uint32_t)vs.world * vs.slab : n;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < total; j += stride) {
        uint32_t i = j;
        if (SHARDED && !vis_record(rec, vs, n_local, j, &i)) continue;
        const int4 b = __ldg(bink + i);  // (sx, sy, span, kgb)
        const uint32_t kgb = (uint32_t)b.w;
        if (!(kgb & 2u)) continue;
        const uint32_t group = ((kgb >> 8) & 0x1fu) | ((kgb & 1u) << 5);  // bit 5 -> key bit 31: the hitbox is a rect
        const int32_t sy = b.y, ex = b.x + (int32_t)((uint32_t)b.z & 0xffffu), ey = b.y + (int32_t)((uint32_t)b.z >> 16);
        const int32_t x0 = max(b.x, col_lo), x1 = min(ex, col_hi);
        for (int32_t x = x0; x < x1; ++x)
            for (int32_t y = sy; y != ey; ++y) {
                const uint32_t s = geom.slot(x, y);
                const uint32_t pos = atomicAdd(&offsets[s], 1u);
                if (pos < cap_entries) *reinterpret_cast<uint2*>(entries + pos) = make_uint2(s | (group << 26), i);
                else ctr->overflow_entries = 1u;
            }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Overlap set (HitboxInfo.overlaps, collider.rs:463): "is (hi, lo) overlapping already?" is asked for every candidate
// pair, so the common answer — no — must not cost a dependent trip to L2 (it was 22 % of the candidate stage's stall
// samples, profiles/r1f_candidates.txt).  Two levels:
//   bits    one bit per gid: the hitbox has at least one overlap.  ~n/8 bytes, cache-resident; a pair is looked up only
//           if both bits are set
//   pair    one bit per hash(hi, lo) in a table of >= 16 bits per pair (a one-hash Bloom filter): where most hitboxes overlap
//           something (1 shape per unit^2: all of them) the gid bits pass everything and every candidate paid a 32-byte DRAM
//           sector in the table below (18 % of k_solve_dense's stall samples, profiles/r1n); this 4-byte load hits L2 and
//           answers "no" for ~15 of 16 pairs that do not overlap
//   table   open addressing over 32-byte buckets of four (hi << 32 | lo) keys: one sector per probe step
struct OverlapSet {
    const ulonglong2* table;  // [2 * n_buckets]; kKeyMax = empty
    const uint32_t* bits;
    uint32_t bucket_mask;     // n_buckets - 1 (power of two); unused when n_pairs == 0
    uint32_t n_pairs;
    const uint32_t* pair_bits;  // [pair_mask + 1] words
    uint32_t pair_mask;         // words - 1 (power of two)
};
__device__ __forceinline__ uint32_t pair_bit_index(unsigned long long h) { return (uint32_t)(h >> 32); }  // (the bucket uses the low half)
__device__ __forceinline__ bool overlap_contains(const OverlapSet& ov, uint32_t hi, uint32_t lo) {
    if (!((__ldg(ov.bits + (hi >> 5)) >> (hi & 31u)) & 1u) || !((__ldg(ov.bits + (lo >> 5)) >> (lo & 31u)) & 1u)) return false;
    const unsigned long long key = ((unsigned long long)hi << 32) | lo;
    const unsigned long long h = mix64(key);
    const uint32_t pb = pair_bit_index(h);
    if (!((__ldg(ov.pair_bits + ((pb >> 5) & ov.pair_mask)) >> (pb & 31u)) & 1u)) return false;
    uint32_t b = (uint32_t)h & ov.bucket_mask;
    for (;;) {
        const ulonglong2 q0 = __ldg(ov.table + 2 * (size_t)b), q1 = __ldg(ov.table + 2 * (size_t)b + 1);
        if (q0.x == key || q0.y == key || q1.x == key || q1.y == key) return true;
        if (q1.y == kKeyMax) return false;  // buckets fill front to back: a free last slot ends the probe sequence
        b = (b + 1) & ov.bucket_mask;
    }
}
__global__ void k_overlap_insert(const uint2* __restrict__ pairs, uint32_t n_pairs, unsigned long long* table, uint32_t bucket_mask, uint32_t* bits,
                                 uint32_t* pair_bits, uint32_t pair_mask) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    const uint2 pr = pairs[i];
    atomicOr(bits + (pr.x >> 5), 1u << (pr.x & 31u));
    atomicOr(bits + (pr.y >> 5), 1u << (pr.y & 31u));
    const unsigned long long key = ((unsigned long long)pr.x << 32) | pr.y;
    const unsigned long long h = mix64(key);
    const uint32_t pb = pair_bit_index(h);
    atomicOr(pair_bits + ((pb >> 5) & pair_mask), 1u << (pb & 31u));
    uint32_t b = (uint32_t)h & bucket_mask;
    for (;;) {
        // slots of a bucket are claimed front to back, so a bucket never has a hole before a key
        for (int k = 0; k < 4; ++k) {
            const unsigned long long prev = atomicCAS(&table[4 * (size_t)b + k], kKeyMax, key);
            if (prev == kKeyMax || prev == key) return;
        }
        b = (b + 1) & bucket_mask;
    }
}
// Order of the overlapping-pair list.  The solve stage gathers both records of every listed pair; rows are kept in grid-slot
// order, and two overlapping hitboxes are neighbours, so a list ordered by the ROW of its first hitbox turns those gathers into
// near-sequential reads (config-4 density: 3.4 M pairs per million hitboxes, 650 MB of gathers per step).  A counting sort on
// row >> shift (buckets of 2^shift rows; the order inside a bucket does not matter), redone after every re-sort of the table.
__global__ void __launch_bounds__(kThreads) k_ovsort_count(const uint2* __restrict__ pairs, uint32_t n_pairs, const uint32_t* __restrict__ row_of, uint32_t n_labels,
                                                           int shift, uint32_t n_buckets, uint32_t* hist) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += gridDim.x * blockDim.x) {
        const uint32_t hi = pairs[i].x;
        const uint32_t row = hi < n_labels ? row_of[hi] : 0xffffffffu;
        atomicAdd(&hist[min(row >> shift, n_buckets - 1u)], 1u);
    }
}
__global__ void __launch_bounds__(1024) k_ovsort_scan(uint32_t* hist, uint32_t n_buckets) {  // one block: hist -> exclusive prefix (cursors)
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_buckets; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n_buckets ? hist[i] : 0u;
        uint32_t total;
        const uint32_t excl = block_excl_scan(v, &total);
        const uint32_t carry = s_carry;
        if (i < n_buckets) hist[i] = carry + excl;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + total;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(kThreads) k_ovsort_scatter(const uint2* __restrict__ pairs, uint32_t n_pairs, const uint32_t* __restrict__ row_of, uint32_t n_labels,
                                                             int shift, uint32_t n_buckets, uint32_t* cursor, uint2* out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += gridDim.x * blockDim.x) {
        const uint2 pr = pairs[i];
        const uint32_t row = pr.x < n_labels ? row_of[pr.x] : 0xffffffffu;
        out[atomicAdd(&cursor[min(row >> shift, n_buckets - 1u)], 1u)] = pr;
    }
}

// ids -> slots (binary search in the ascending id column); pairs out as (hi slot, lo slot).
__global__ void k_overlap_map_ids(const uint64_t* __restrict__ ids, uint32_t n, const uint64_t* __restrict__ id_a,
                                  const uint64_t* __restrict__ id_b, uint32_t n_pairs, uint2* pairs, Counters* ctr) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    uint32_t s[2];
    const uint64_t want[2] = {id_a[i], id_b[i]};
    if (ids == nullptr) {  // sharded: the id is its own rank
        if (want[0] >= 0x7fffffffull || want[1] >= 0x7fffffffull || want[0] == want[1]) report_error(ctr, kFIdNotFound, i);
        pairs[i] = make_uint2((uint32_t)max(want[0], want[1]), (uint32_t)min(want[0], want[1]));
        return;
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        uint32_t lo = 0, hi = n;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (ids[mid] < want[k]) lo = mid + 1;
            else hi = mid;
        }
        if (lo >= n || ids[lo] != want[k]) {
            report_error(ctr, kFIdNotFound, i);
            lo = 0;
        }
        s[k] = lo;
    }
    if (s[0] == s[1]) report_error(ctr, kFIdNotFound, i);
    pairs[i] = make_uint2(max(s[0], s[1]), min(s[0], s[1]));
}

// ---------------------------------------------------------------------------------------------------------
// Candidate stage, part 1: enumeration.  The entry array is sorted by slot, so the possible partners of an entry are the
// entries that precede it in its slot run.  This kernel only ENUMERATES those (entry, predecessor) pairs from the 8-byte
// entries and writes them as dense work lists of (record, record, slot) items — grouped by kind pair (rect/rect,
// circle/circle, mixed) — for the solve stage, which tests and solves them from the records it has to load anyway.
// (Testing inside a per-entry walk, a block-wide flattening and a warp-shuffle walk over 32-byte entries were all measured
// first: each spent 22-33 M warp instructions per step with most lanes idle or waiting — profiles/r1c…r1g.)
// A block takes kEnumTile consecutive entries per round, one chunk of kEnumChunk per warp, 32 entries per step:
//   rank in run   one ballot of the run heads; a run that began before the step continues through `last_rank`
//   item counts   the kind of the k-th predecessor is a bit of the kind ballots, so a lane's items per kind pair are two
//                 popcounts per entry; one warp scan per kind pair gives every lane its own output range
//   output        the block reserves its ranges with ONE atomic per kind pair and round (on one of kSubLists sub-lists:
//                 same-address atomics serialise at ~2 ns each in L2, and a warp that appended through its own atomics
//                 spent 45 % of its time waiting for them, profiles/r1g_enum_solve.txt); each lane then writes its items
//                 in place, reading the predecessors' 8-byte entries back from L1 — no staging, no per-warp atomics
constexpr int kEnumChunk = 128;                  // entries per warp and round
constexpr int kEnumSteps = kEnumChunk / 32;
constexpr int kEnumTile = kEnumChunk * (kThreads / 32);  // entries per block and round

struct WorkItem {
    uint32_t a, b;  // record indices of the entry and of its predecessor
    uint32_t slot;  // the slot of the run they were found in (pair ownership)
    uint32_t pad;
};
struct WorkLists {  // work buffer = kWorkLists equal segments; list l holds kind pair l / kSubLists
    WorkItem* items;
    uint32_t seg_cap;
    uint32_t* count;  // [kWorkLists]
};

// Long slot runs (dense cells) are not expanded into work items: an entry whose rank in its run is >= kDenseRank leaves its
// pairs to k_solve_dense, which walks the run from shared memory.  The entries of rank kDenseRank, kDenseRank + 32, ... each
// append one descriptor (first entry of the run, own rank): the 32 entries from that rank on are the lanes of one warp there.
constexpr uint32_t kDenseRank = 4;  // measured at 1 shape per unit^2: 1.09 ms per step with 2 or 4, 1.16 with 8; clustered config 5: 0.54 with 4, 0.57 with 8, 0.61 with 2
constexpr uint32_t kDenseOff = 0xffffffffu;  // DenseList.rank0: every pair becomes a work item (A/B: CB200_DENSE_RANK=0)
struct DenseList {
    uint2* desc;      // (entry index of the run's first entry, rank j0)
    uint32_t* count;
    uint32_t cap;     // >= entries / (rank0 + 1) + entries / 32 (one per long run + one per 32 entries): cannot overflow
    uint32_t* cursor; // next descriptor to hand out (k_solve_dense); zeroed with the counts
    uint32_t rank0;   // first dense rank (kDenseRank; 1 ..: tests force everything through the dense kernel)
};
struct EnumArgs {
    const CellEntry* entries;
    uint32_t cap_entries;  // the scatter wrote only the entries below it (an overflowing step is re-run)
    WorkLists work;
    DenseList dense;
    Counters* ctr;
};

// Work items of one lane: for each of its entries, the `rank` entries before it.  The kind of a predecessor is a bit of the
// window (previous step's kind ballot, this step's), so the per-list COUNTS are two popcounts per entry; the items
// themselves are written by a lane-private loop that reads the predecessor's 8-byte entry back from L1.
__device__ __forceinline__ uint32_t rect_preds(unsigned prev_rects, unsigned rects, int lane, uint32_t rank) {
    // predecessors k = 1..min(rank, 32 + lane) sit at window bits [32 + lane - k]; window = prev_rects | rects << 32
    const unsigned long long window = (unsigned long long)prev_rects | ((unsigned long long)rects << 32);
    const uint32_t in_window = min(rank, 32u + (uint32_t)lane);
    const unsigned long long below_me = (1ull << (32 + lane)) - 1ull;                         // bits under this lane's
    const unsigned long long from = in_window >= 64u ? 0ull : ~((1ull << (32 + lane - in_window)) - 1ull);
    return (uint32_t)__popcll(window & below_me & from);
}

__global__ void __launch_bounds__(kThreads) k_enum_pairs(const EnumArgs a) {
    __shared__ uint32_t s_cnt[kThreads / 32][3], s_off[kThreads / 32][3], s_base[3];
    const uint32_t n_entries = (uint32_t)min(a.ctr->n_entries, (unsigned long long)a.cap_entries);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr uint32_t kNoSlot = 0xffffffffu;  // (slots have 26 bits)
    const uint2* ent = reinterpret_cast<const uint2*>(a.entries);
    const uint32_t sub = blockIdx.x & (kSubLists - 1);
    for (uint32_t t0 = blockIdx.x * kEnumTile; t0 < n_entries; t0 += gridDim.x * kEnumTile) {
        const uint32_t c0 = t0 + (uint32_t)warp * kEnumChunk;
        const uint32_t c1 = min(c0 + (uint32_t)kEnumChunk, n_entries);  // (c1 <= c0: this warp has nothing)
        // all loads of the round first: the chunk and the 32 entries before it
        uint2 ent_s[kEnumSteps];
#pragma unroll
        for (int s = 0; s < kEnumSteps; ++s) ent_s[s] = c0 + s * 32 + lane < c1 ? __ldg(ent + c0 + s * 32 + lane) : make_uint2(0u, 0u);
        const bool halo_live = c0 < c1 && c0 + lane >= 32;
        const uint2 halo = halo_live ? __ldg(ent + (c0 - 32 + lane)) : make_uint2(0u, 0u);
        const uint32_t halo_slot = halo_live ? (halo.x & kSlotMask) : kNoSlot;
        unsigned prev_rects = __ballot_sync(0xffffffffu, halo_live && (halo.x & kEntryKindBit));
        // rank of the last halo entry in its run
        uint32_t last_slot = __shfl_sync(0xffffffffu, halo_slot, 31), last_rank = 0;
        if (last_slot != kNoSlot) {
            const unsigned same = __ballot_sync(0xffffffffu, halo_slot == last_slot);
            const uint32_t trail = __clz(~same);  // consecutive halo lanes 31, 30, ... in last's run (32: maybe more before the halo)
            last_rank = trail - 1;
            if (trail == 32) {
                uint32_t f = c0 - 32;
                while (f > 0 && (__ldg(ent + (f - 1)).x & kSlotMask) == last_slot) {
                    --f;
                    ++last_rank;
                }
            }
        }
        // ranks, and per-lane item counts per kind pair: [0] rect/rect, [1] circle/circle, [2] mixed
        uint32_t rank[kEnumSteps], cnt[3] = {0u, 0u, 0u};
#pragma unroll
        for (int s = 0; s < kEnumSteps; ++s) {
            const bool live = c0 + s * 32 + lane < c1;
            const uint32_t slot = live ? (ent_s[s].x & kSlotMask) : kNoSlot;
            uint32_t slot_up = __shfl_up_sync(0xffffffffu, slot, 1);
            if (lane == 0) slot_up = last_slot;
            const unsigned heads = __ballot_sync(0xffffffffu, live && slot != slot_up);
            const unsigned rects = __ballot_sync(0xffffffffu, live && (ent_s[s].x & kEntryKindBit));
            const unsigned below = heads & (0xffffffffu >> (31 - lane));  // heads at or below this lane
            const uint32_t rk = !live ? 0u : (below ? (uint32_t)(lane - (31 - __clz(below))) : (uint32_t)lane + last_rank + 1u);  // rank in the run
            // dense part of a run: one descriptor per 32 ranks from kDenseRank on, no work items
            const uint32_t rank0 = a.dense.rank0;
            const bool desc = live && rk >= rank0 && ((rk - rank0) & 31u) == 0u;
            const unsigned dm = __ballot_sync(0xffffffffu, desc);
            if (dm) {
                uint32_t at = 0;
                if (lane == __ffs(dm) - 1) at = atomicAdd(a.dense.count, (uint32_t)__popc(dm));
                at = __shfl_sync(0xffffffffu, at, __ffs(dm) - 1) + (uint32_t)__popc(dm & ((1u << lane) - 1u));
                if (desc && at < a.dense.cap) a.dense.desc[at] = make_uint2(c0 + s * 32 + lane - rk, rk);
            }
            rank[s] = rk < rank0 ? rk : 0u;  // predecessors that become work items
            uint32_t n_rect = rect_preds(prev_rects, rects, lane, rank[s]);
            for (uint32_t k = 33u + (uint32_t)lane; k <= rank[s]; ++k)  // a run longer than the window (dense path off): global memory
                n_rect += __ldg(ent + (c0 + s * 32 + lane - k)).x >> 31;
            if (ent_s[s].x >> 31) {
                cnt[0] += n_rect;
                cnt[2] += rank[s] - n_rect;
            } else {
                cnt[2] += n_rect;
                cnt[1] += rank[s] - n_rect;
            }
            prev_rects = rects;
            last_slot = __shfl_sync(0xffffffffu, slot, 31);
            last_rank = __shfl_sync(0xffffffffu, rk, 31);
        }
        // the block reserves its ranges with one atomic per kind pair; lanes get disjoint sub-ranges by a warp scan
        uint32_t lane_off[3], warp_tot[3];
#pragma unroll
        for (int l = 0; l < 3; ++l) {
            uint32_t inc = cnt[l];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += u;
            }
            lane_off[l] = inc - cnt[l];
            warp_tot[l] = __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane < 3) s_cnt[warp][lane] = lane == 0 ? warp_tot[0] : (lane == 1 ? warp_tot[1] : warp_tot[2]);
        __syncthreads();
        if (threadIdx.x < 3) {
            uint32_t run = 0;
            for (int w = 0; w < kThreads / 32; ++w) {
                s_off[w][threadIdx.x] = run;
                run += s_cnt[w][threadIdx.x];
            }
            s_base[threadIdx.x] = run ? atomicAdd(&a.work.count[threadIdx.x * kSubLists + sub], run) : 0u;
        }
        __syncthreads();
        // Items past the end of a segment are dropped (the step is re-run with a larger buffer) — one by one, so that every
        // position below seg_cap IS written: the solve stage reads min(count, seg_cap) items of each list.
        uint32_t pos[3];
#pragma unroll
        for (int l = 0; l < 3; ++l) pos[l] = s_base[l] + s_off[warp][l] + lane_off[l];
        {
            const uint32_t cap = a.work.seg_cap;
            uint4* const out0 = reinterpret_cast<uint4*>(a.work.items + (size_t)(0 * kSubLists + sub) * a.work.seg_cap);
            uint4* const out1 = reinterpret_cast<uint4*>(a.work.items + (size_t)(1 * kSubLists + sub) * a.work.seg_cap);
            uint4* const out2 = reinterpret_cast<uint4*>(a.work.items + (size_t)(2 * kSubLists + sub) * a.work.seg_cap);
#pragma unroll
            for (int s = 0; s < kEnumSteps; ++s) {
                const uint32_t e = c0 + s * 32 + lane, my_rect = ent_s[s].x >> 31, slot = ent_s[s].x & kSlotMask;
                for (uint32_t k = 1; k <= rank[s]; ++k) {
                    const uint2 pred = __ldg(ent + (e - k));
                    const uint4 item = make_uint4(ent_s[s].y, pred.y, slot, 0u);
                    const uint32_t n_rect = my_rect + (pred.x >> 31);
                    if (n_rect == 2) {
                        if (pos[0] < cap) out0[pos[0]] = item;
                        ++pos[0];
                    } else if (n_rect == 0) {
                        if (pos[1] < cap) out1[pos[1]] = item;
                        ++pos[1];
                    } else {
                        if (pos[2] < cap) out2[pos[2]] = item;
                        ++pos[2];
                    }
                }
            }
        }
        __syncthreads();  // s_cnt / s_off / s_base are rewritten by the next round
    }
}

// Candidate stage, part 2: the test of one work item, from the two record heads — run by the solve stage (and by the
// debug listing kernel).
//   candidates   Grid::overlapping_ids (grid.rs:115-135) restricted to the surviving (hi, lo) pairs; the
//                hash-set de-duplication is replaced by pair ownership: a pair is taken only in the slot of
//                cell (max(sx_i, sx_j), max(sy_i, sy_j)) — the min corner of the rect intersection — so every
//                pair sharing >= 1 cell is produced exactly once, and aliased slot-mates are rejected.  Sharded:
//                only the rank whose strip holds the owner cell's column takes the pair.
//   filters      lo.group in hi.interact_groups (grid.rs:122, key group); not already overlapping (collider.rs:351)
struct CandArgs {
    GridGeom geom;
    int32_t col_lo, col_hi;
    OverlapSet ov;
};
// *a_is_hi: the item's first record carries the higher id.
// candidate_geom: everything but the overlap lookup.  The step's solve count is (pairs passing candidate_geom) - (overlapping
// pairs passing candidate_geom): every pair that shares a cell is enumerated exactly once, so the solve stage counts the
// first term over its items and the second over the overlap list it walks anyway — the dense kernel then needs the hash
// lookup only for the pairs that survive its swept-bounds filter.
__device__ __forceinline__ bool candidate_geom(const CandArgs& a, const RecHead& ha, const RecHead& hb, uint32_t slot, bool* a_is_hi) {
    const int32_t ox = max(ha.sx, hb.sx), oy = max(ha.sy, hb.sy);
    if (ox >= min(ha.ex, hb.ex) || oy >= min(ha.ey, hb.ey)) return false;
    if (a.geom.slot(ox, oy) != slot) return false;
    if (ox < a.col_lo || ox >= a.col_hi) return false;
    const bool first_hi = ha.gid > hb.gid;
    *a_is_hi = first_hi;
    const uint32_t mask_hi = first_hi ? ha.mask : hb.mask, group_lo = ((first_hi ? hb.kgb : ha.kgb) >> 8) & 31u;
    return ((mask_hi >> group_lo) & 1u) != 0u;
}
__device__ __forceinline__ bool is_overlapping(const CandArgs& a, const RecHead& ha, const RecHead& hb, bool a_is_hi) {
    return a.ov.n_pairs && overlap_contains(a.ov, a_is_hi ? ha.gid : hb.gid, a_is_hi ? hb.gid : ha.gid);
}
__device__ __forceinline__ bool candidate_test(const CandArgs& a, const RecHead& ha, const RecHead& hb, uint32_t slot, bool* a_is_hi) {
    return candidate_geom(a, ha, hb, slot, a_is_hi) && !is_overlapping(a, ha, hb, *a_is_hi);
}
// An overlapping pair (both record heads at hand): would it have been enumerated as a candidate but for the overlap lookup?
__device__ __forceinline__ bool overlap_pair_is_candidate(const CandArgs& a, const RecHead& ha, const RecHead& hb) {
    bool hi;
    return candidate_geom(a, ha, hb, a.geom.slot(max(ha.sx, hb.sx), max(ha.sy, hb.sy)), &hi);
}

// Index space of the work lists of one step: list l holds min(count[l], seg_cap) items; first[l] = items before it.
struct WorkIndex {
    uint32_t first[kWorkLists + 1];
};
__device__ __forceinline__ void work_index_build(const WorkLists& w, WorkIndex* s_idx) {  // all threads of the block; barrier inside
    __shared__ uint32_t s_tmp[kWorkLists];
    for (int l = threadIdx.x; l < kWorkLists; l += blockDim.x) s_tmp[l] = min(w.count[l], w.seg_cap);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int l = 0; l < kWorkLists; ++l) {
            s_idx->first[l] = run;
            run += s_tmp[l];
        }
        s_idx->first[kWorkLists] = run;
    }
    __syncthreads();
}
__device__ __forceinline__ WorkItem work_item_at(const WorkLists& w, const WorkIndex* s_idx, uint32_t p) {
    int lo = 0, hi = kWorkLists;  // the list l with first[l] <= p < first[l + 1]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (s_idx->first[mid] <= p) lo = mid;
        else hi = mid;
    }
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(w.items + (size_t)lo * w.seg_cap + (p - s_idx->first[lo])));
    return WorkItem{q.x, q.y, q.z, q.w};
}

// Debug / parity view: the accepted candidate pairs as (gid hi, gid lo), unordered (cb200_fetch_candidate_pairs).
__global__ void __launch_bounds__(kThreads) k_list_candidates(const CandArgs a, const WorkLists w, const HbRec* __restrict__ rec, uint2* out,
                                                              unsigned long long* n_out, uint64_t cap) {
    __shared__ WorkIndex s_idx;
    work_index_build(w, &s_idx);
    const uint32_t total = s_idx.first[kWorkLists];
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
        const WorkItem it = work_item_at(w, &s_idx, p);
        const RecHead ha = load_head(rec, it.a), hb = load_head(rec, it.b);
        bool a_is_hi;
        if (!candidate_test(a, ha, hb, it.slot, &a_is_hi)) continue;
        const unsigned long long pos = warp_agg_claim(n_out);
        if (pos < cap) out[pos] = a_is_hi ? make_uint2(ha.gid, hb.gid) : make_uint2(hb.gid, ha.gid);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Solve stage: one thread per pair.  Pairs [0, n_collide) are the candidates (record indices), pairs
// [n_collide, n_collide + n_ov) the currently-overlapping pairs (HitboxInfo.overlaps) as (gid hi, gid lo).
//   collide      hi.collide_time(lo rebased to T) (collider.rs:354 -> solvers.rs:25-34)
//                -> add_pair_event(T + delay, Collide(hi, lo)), dropped when >= 1e50 (collider.rs:367-372, events.rs:156-159)
//   separate     hi.separate_time(lo rebased to T, padding) (collider.rs:329-339 -> solvers.rs:36-44)
//                -> add_pair_event(T + delay, Separate(hi, lo))
//   min          EventManager::peek_time under the canonical order (events.rs:171-173): warp shuffle + block reduce,
//                the last block to finish reduces all per-block minima and publishes the step result
// The step result, written by the last block straight into mapped pinned host memory.
struct StepResultDev {
    double next_time;
    uint64_t next_tie;  // MinKey.tie of the minimum (kKeyMax if none)
    uint64_t next_id_a, next_id_b;  // its HbIds (label -> id), so the host needs no further copy
    Counters ctr;
    // window drain (drain.cuh), written by k_drain_commit
    unsigned long long n_followups;
    uint32_t ov_pairs_now, drain_flags, n_removed, n_added, ov_live_now, pad;
};

struct SolveArgs {
    CandArgs cd;             // the candidate test of a work item
    WorkLists work;          // k_enum_pairs' lists
    const uint2* ov_pairs;
    uint32_t n_ov;             // overlapping pairs (with n_ov_dev: the capacity of the list)
    const uint32_t* n_ov_dev;  // window drain: the pair count lives on the device (drain.cuh OvDyn.n_pairs); null = n_ov
    int ov_live_check;         // window drain: the list also holds pairs that separated since the last rebuild — skip those (table lookup)
    const HbRec* rec;
    const StepDyn* dyn;
    double padding;
    PairEvent* events;
    uint32_t cap_events;
    Counters* ctr;
    MinKey* partial;         // [n_partial_a + gridDim.x + n_partial_dense]: stage A's minima, this kernel's, k_solve_dense's
    uint32_t n_partial_a;
    uint32_t n_partial_dense;
    StepResultDev* result;   // mapped host memory
    MinKey* local_key;       // device copy of the local minimum (sharded: input of the all-gather min)
    // overlapping pairs are looked up by gid; sharded: only the owner-column rank solves them
    const uint32_t* vmap;    // gid -> record index (entries may be stale: validated)
    uint32_t id_space;
    int sharded;
    VisSpace vs;
    int32_t col_lo, col_hi;
    const uint64_t* ids;     // label -> HbId (null: the label is the id)
    uint32_t dbg;            // timing experiments only (CB200_DBG_S): 1 skip the narrowphase, 2 skip the event append, 4 skip the candidate test
    uint32_t prefetch;       // k_solve: request the next work item's records into L2 one iteration ahead (CB200_SOLVE_PREFETCH)
    // add mode (batched Collider::add_hitbox, collider.rs:214-222 + 355-365): a candidate whose collide delay is exactly 0.0
    // overlaps NOW — its pair is reported in imm_pairs and its event is the Separate; simultaneous events of one hitbox are
    // ordered by partner id alone (they were created in one ascending loop), so the kind bit sits below the partner label
    int add_mode;
    uint2* imm_pairs;        // (gid hi, gid lo)
    unsigned long long* n_imm;
    uint32_t cap_imm;
};

// The three 32-byte sectors of a 96-byte record, requested into L2 (no register, no scoreboard entry).
__device__ __forceinline__ void prefetch_rec_l2(const HbRec* rec, uint32_t v) {
    const char* q = reinterpret_cast<const char*>(rec + v);
    asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(q + 32));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(q + 64));
}

// A map entry is trusted only if it names a record slot that is live in THIS step and that record carries the gid.
// One device: vmap = row_of (label -> row; record index == row).  Sharded: vmap = gid -> visitor index.
__device__ __forceinline__ bool gid_lookup(const SolveArgs& a, uint32_t n_local, uint32_t gid, uint32_t* vidx) {
    if (gid >= a.id_space) return false;
    const uint32_t v = __ldg(a.vmap + gid);
    if (!a.sharded) {
        if (v >= a.vs.n_own) return false;
    } else if (v >= n_local) {
        if (v < a.vs.recv0) return false;
        const uint32_t j = v - a.vs.recv0;
        if (j >= (uint32_t)a.vs.world * a.vs.slab) return false;
        const uint32_t d = j / a.vs.slab, k = j % a.vs.slab;
        if ((int)d == a.vs.rank || k == 0 || k - 1 >= slab_count(a.rec, a.vs, (int)d)) return false;
    }
    if (__ldg(reinterpret_cast<const unsigned int*>(a.rec + v) + 3) != gid) return false;
    *vidx = v;
    return true;
}

// End of the solve stage (all threads of the block): block minimum and counters out; the last block to finish reduces every
// per-block minimum of stage A and of this stage and publishes the step result.
// n_col: items that passed candidate_geom MINUS overlapping pairs that pass it (see candidate_geom) — per thread, may be negative.
__device__ __forceinline__ void solve_epilogue(const SolveArgs& a, MinKey best, uint32_t n_sep, int32_t n_col) {
    __shared__ uint32_t s_last;
    best = block_min(best);
    const unsigned long long sep_tot = block_sum((unsigned long long)n_sep), col_tot = block_sum((unsigned long long)(long long)n_col);
    if (threadIdx.x == 0) {
        a.partial[a.n_partial_a + blockIdx.x] = best;
        if (sep_tot) atomicAdd(&a.ctr->n_separate, sep_tot);
        if (col_tot) atomicAdd(&a.ctr->n_collide, col_tot);  // (mod 2^64: the sum over all blocks of both solve kernels is >= 0)
        __threadfence();
        s_last = atomicAdd(&a.ctr->done_blocks, 1u) == gridDim.x - 1 ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    // last block: reduce every per-block minimum of stage A and of this stage, publish the step result
    __threadfence();
    best = MinKey{kKeyMax, kKeyMax};
    const uint32_t n_partial = a.n_partial_a + gridDim.x + a.n_partial_dense;
    for (uint32_t i = threadIdx.x; i < n_partial; i += blockDim.x) {
        const volatile unsigned long long* q = reinterpret_cast<const volatile unsigned long long*>(&a.partial[i]);
        const MinKey k{q[0], q[1]};
        if (key_less(k, best)) best = k;
    }
    best = block_min(best);
    if (threadIdx.x == 0) {
        a.ctr->n_first_events = a.ctr->n_pair_events;  // (every block of the solve stage has flushed its events)
        a.result->next_time = (best.t == kKeyMax && best.tie == kKeyMax) ? cb_inf() : key_time(best.t);
        a.result->next_tie = best.tie;
        {
            const bool pair = (best.tie & kPairClass) != 0;
            const uint32_t la = pair ? (uint32_t)((best.tie >> 32) & 0x7fffffffu) : (uint32_t)best.tie;
            const uint32_t lb = a.add_mode ? (uint32_t)((best.tie & 0xffffffffu) >> 1) : (uint32_t)(best.tie & 0x7fffffffu);
            const bool any = !(best.t == kKeyMax && best.tie == kKeyMax);
            a.result->next_id_a = !any ? 0ull : (a.ids ? a.ids[la] : (uint64_t)la);
            a.result->next_id_b = !any || !pair ? 0ull : (a.ids ? a.ids[lb] : (uint64_t)lb);
        }
        if (a.local_key) *a.local_key = best;
        const volatile Counters* c = a.ctr;
        Counters out;
        out.n_entries = c->n_entries; out.n_cells = c->n_cells; out.n_separate = c->n_separate;
        out.work_overflow = 0;
        out.n_work = 0;
        for (int l = 0; l < kWorkLists; ++l) {
            out.work_overflow = max(out.work_overflow, a.work.count[l]);  // fullest sub-list
            out.n_work += a.work.count[l];
        }
        out.pad0 = 0;
        out.n_collide = c->n_collide;
        out.n_pair_events = c->n_pair_events; out.n_solitaire = c->n_solitaire; out.err_slot = c->err_slot;
        out.err_flags = c->err_flags; out.scan_ticket = c->scan_ticket; out.overflow_entries = c->overflow_entries;
        out.done_blocks = c->done_blocks; out.n_local = c->n_local; out.overflow_send = c->overflow_send;
        out.diff_hi = 0; out.diff_lo = 0; out.n_sort_keys = 0;
        out.n_tile_pairs = c->n_tile_pairs; out.n_surv = c->n_surv; out.tile_fail = c->tile_fail; out.pad1 = 0; out.n_first_events = c->n_pair_events;
        out.n_work += out.n_tile_pairs;
        a.result->ctr = out;
        __threadfence_system();
    }
}

// Pair events are collected in shared memory and flushed with one global atomic per kSolveBuf events (instead of a
// block scan + atomic + two barriers per iteration).
constexpr uint32_t kSolveBuf = 4 * kThreads;

#ifndef CB200_LB_SOLVE
#define CB200_LB_SOLVE 4
#endif
// WIDE: records through 256-bit loads.  NET: the step also runs k_solve_dense, whose filter sees candidates before the overlap
// lookup — the solve count is then kept as (pairs passing candidate_geom) - (overlapping pairs passing it), see candidate_geom.
// Without the dense kernel the count is simply the items that pass candidate_test; that instantiation is the one a sparse
// scene runs, and it stays free of the extra bookkeeping (measured: 77 vs 94 us at config 3).
template <bool WIDE, bool NET>
__global__ void __launch_bounds__(kThreads, CB200_LB_SOLVE) k_solve(const SolveArgs a) {
    __shared__ PairEvent s_ev[kSolveBuf];
    __shared__ WorkIndex s_idx;
    __shared__ uint32_t s_n;
    __shared__ unsigned long long s_base;
    MinKey best{kKeyMax, kKeyMax};
    uint32_t n_sep = 0;
    int32_t n_col = 0;  // net: candidates - overlapping pairs that are candidates (solve_epilogue)
    if (threadIdx.x == 0) s_n = 0;
    work_index_build(a.work, &s_idx);
    const uint32_t n_cand = s_idx.first[kWorkLists];
    const uint32_t total = n_cand + (a.n_ov_dev ? min(*a.n_ov_dev, a.n_ov) : a.n_ov);
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
    // Software prefetch (SolveArgs.prefetch): the work item of the thread's NEXT iteration is loaded one iteration early and its
    // two records are requested into L2 while this iteration's records are in flight and its narrowphase runs.
    const bool pf = a.prefetch != 0u;
    WorkItem nxt{0u, 0u, 0u, 0u};
    {
        const uint32_t p0 = blockIdx.x * blockDim.x + threadIdx.x;
        if (pf && p0 < n_cand) nxt = work_item_at(a.work, &s_idx, p0);
    }
    for (uint32_t base = blockIdx.x * blockDim.x; base < total; base += stride) {
        const uint32_t p = base + threadIdx.x;
        if (p < total) {
            const bool sep = p >= n_cand;
            uint32_t ia, ib, slot = 0;  // record indices
            bool present = true;
            if (sep) {
                const uint2 pr = a.ov_pairs[p - n_cand];
                present = (!a.ov_live_check || overlap_contains(a.cd.ov, pr.x, pr.y)) && gid_lookup(a, n_local, pr.x, &ia) && gid_lookup(a, n_local, pr.y, &ib);
            } else {
                const WorkItem it = pf ? nxt : work_item_at(a.work, &s_idx, p);
                ia = it.a; ib = it.b; slot = it.slot;
                if (pf && p + stride < n_cand) {
                    nxt = work_item_at(a.work, &s_idx, p + stride);
                    prefetch_rec_l2(a.rec, nxt.a);
                    prefetch_rec_l2(a.rec, nxt.b);
                }
            }
            if (present) {
                // all twelve 16-byte loads of the two records are issued before anything depends on them
                RecHead ha, hb;
                DurHb da, db;
                if (WIDE) {
                    load_rec_wide(a.rec, ia, &ha, &da);
                    load_rec_wide(a.rec, ib, &hb, &db);
                } else {
                    load_rec_keep(a.rec, ia, keep, &ha, &da);
                    load_rec_keep(a.rec, ib, keep, &hb, &db);
                }
                bool a_is_hi = true;  // overlap pairs arrive as (hi, lo)
                if (!sep) {
                    if (NET) {
                        present = (a.dbg & 4u) ? true : candidate_geom(a.cd, ha, hb, slot, &a_is_hi);
                        n_col += present;
                        present = present && ((a.dbg & 4u) || !is_overlapping(a.cd, ha, hb, a_is_hi));
                    } else {
                        present = (a.dbg & 4u) ? true : candidate_test(a.cd, ha, hb, slot, &a_is_hi);
                        n_col += present;
                    }
                } else {
                    const int32_t owner_col = max(ha.sx, hb.sx);
                    present = owner_col >= a.col_lo && owner_col < a.col_hi;
                    if (NET) n_col -= present && overlap_pair_is_candidate(a.cd, ha, hb);
                }
                if (present) {
                    const DurHb A = pick(a_is_hi, da, db);
                    // the partner is `other.hitbox_at_time(T)` (collider.rs:478-486): advanced by T - start = 0
                    const DurHb B = advanced(pick(a_is_hi, db, da), 0.0);
                    const uint32_t gid_hi = a_is_hi ? ha.gid : hb.gid, gid_lo = a_is_hi ? hb.gid : ha.gid;
                    double delay = (a.dbg & 1u) ? (A.px + B.px) : (sep ? separate_time(A, B, a.padding) : collide_time(A, B));
                    bool is_sep = sep;
                    if (a.add_mode && !sep && delay == 0.0) {
                        // immediate overlap: process_collision (collider.rs:177-197) queues the Separate
                        const unsigned long long at = atomicAdd(a.n_imm, 1ull);
                        if (at < a.cap_imm) a.imm_pairs[at] = make_uint2(gid_hi, gid_lo);
                        delay = separate_time(A, B, a.padding);
                        is_sep = true;
                    }
                    if (delay != delay) report_error(a.ctr, kFNan, gid_hi);
                    n_sep += sep;
                    const double t = T + delay;
                    if (t < kHighTime && !(a.dbg & 2u)) {
                        const uint32_t low = a.add_mode ? ((gid_lo << 1) | (is_sep ? 0u : 1u)) : (gid_lo | (is_sep ? 0u : kCollideBit));
                        s_ev[atomicAdd(&s_n, 1u)] = PairEvent{t, gid_hi, low};
                        const MinKey k{time_key(t), kPairClass | ((uint64_t)gid_hi << 32) | low};
                        if (key_less(k, best)) best = k;
                    }
                }
            }
        }
        // the buffer holds kSolveBuf events and an iteration adds at most blockDim.x
        if (++pending_rounds == kSolveBuf / kThreads) {
            flush();
            pending_rounds = 0;
        }
    }
    flush();
    solve_epilogue(a, best, n_sep, n_col);
}

// Two-phase solve stage (CB200_SOLVE=2; measured SLOWER than k_solve — 89 vs 79 us at config 3, 2.12 vs 1.92 ms at 1 shape
// per unit^2 — and kept as an A/B variant and as a second implementation the parity tests run).  In k_solve above a thread carries its item through everything, and most items stop
// early — the candidate test, and above all collide_time's swept-bounds test (solvers.rs:27-30) — so the sweep math ran with
// 17 of 32 lanes on average (ncu, profiles/r1k_step_kernels.txt).  Here a block alternates between
//   filter   one thread per work item: records, candidate test, duration, swept bounds of both shapes, touch test.  The
//            survivors (record of the higher id, record of the lower id) go to a queue in shared memory; the overlapping
//            pairs of the separate list go there directly
//   solve    whenever 256 survivors are queued, one thread each: sweep_unpadded / separate_time with full warps, event
//            append, running minimum.  The records are read again (they are in L1 / L2 from the filter)
// The arithmetic is the same device functions in the same order as in k_solve (collide_time = bounds + touch + sweep), so
// the results are bit-identical; tests/test_bulk_gpu.py runs both kernels against the oracle.  Why it loses: the second read
// of the records and the extra barriers cost more than the idle lanes did — the stage is bound by the record gathers
// (l1tex data-pipe wavefronts 46 %, the busiest unit), not by issue slots (31 %).
constexpr uint32_t kQueueCap = 2 * kThreads;
constexpr uint32_t kSepBit = 0x80000000u;  // in Survivor.y

__global__ void __launch_bounds__(kThreads, CB200_LB_SOLVE) k_solve2(const SolveArgs a) {
    __shared__ PairEvent s_ev[kSolveBuf];
    __shared__ uint2 s_q[kQueueCap];  // (record of the hi id, record of the lo id | kSepBit)
    __shared__ WorkIndex s_idx;
    __shared__ uint32_t s_n, s_qn;
    __shared__ unsigned long long s_base;
    MinKey best{kKeyMax, kKeyMax};
    uint32_t n_sep = 0;
    int32_t n_col = 0;  // net: candidates - overlapping pairs that are candidates (solve_epilogue)
    if (threadIdx.x == 0) {
        s_n = 0;
        s_qn = 0;
    }
    work_index_build(a.work, &s_idx);  // (barriers inside)
    const uint32_t n_cand = s_idx.first[kWorkLists];
    const uint32_t total = n_cand + (a.n_ov_dev ? min(*a.n_ov_dev, a.n_ov) : a.n_ov);
    const uint32_t n_local = a.ctr->n_local;
    const uint32_t stride = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    const double T = a.dyn->T;
    const uint64_t keep = l2_policy_evict_last();
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
    // solve phase: survivors s_q[first .. first + count), count <= blockDim.x (all threads call)
    auto solve_batch = [&](uint32_t first, uint32_t count) {
        if (threadIdx.x < count) {
            const uint2 sv = s_q[first + threadIdx.x];
            const bool sep = (sv.y & kSepBit) != 0;
            RecHead hh, hl;
            DurHb A, Bl;
            load_rec_keep(a.rec, sv.x, keep, &hh, &A);
            load_rec_keep(a.rec, sv.y & ~kSepBit, keep, &hl, &Bl);
            // the partner is `other.hitbox_at_time(T)` (collider.rs:478-486): advanced by T - start = 0
            const DurHb B = advanced(Bl, 0.0);
            double delay = sep ? separate_time(A, B, a.padding) : sweep_unpadded(A, B, true, fmin(A.dur, B.dur));
            bool is_sep = sep;
            if (a.add_mode && !sep && delay == 0.0) {
                // immediate overlap: process_collision (collider.rs:177-197) queues the Separate
                const unsigned long long at = atomicAdd(a.n_imm, 1ull);
                if (at < a.cap_imm) a.imm_pairs[at] = make_uint2(hh.gid, hl.gid);
                delay = separate_time(A, B, a.padding);
                is_sep = true;
            }
            if (delay != delay) report_error(a.ctr, kFNan, hh.gid);
            const double t = T + delay;
            if (t < kHighTime) {
                const uint32_t low = a.add_mode ? ((hl.gid << 1) | (is_sep ? 0u : 1u)) : (hl.gid | (is_sep ? 0u : kCollideBit));
                s_ev[atomicAdd(&s_n, 1u)] = PairEvent{t, hh.gid, low};
                const MinKey k{time_key(t), kPairClass | ((uint64_t)hh.gid << 32) | low};
                if (key_less(k, best)) best = k;
            }
        }
        // the event buffer holds kSolveBuf events and a batch adds at most blockDim.x
        if (++pending_rounds == kSolveBuf / kThreads) {
            flush();
            pending_rounds = 0;
        }
    };
    for (uint32_t base = blockIdx.x * blockDim.x; base < total; base += stride) {
        // ---- filter ----
        const uint32_t p = base + threadIdx.x;
        bool push = false;
        uint2 sv = make_uint2(0u, 0u);
        if (p < total) {
            if (p >= n_cand) {
                // an overlapping pair (hi, lo): solved by the rank that owns its cell column
                const uint2 pr = a.ov_pairs[p - n_cand];
                uint32_t ia, ib;
                if ((!a.ov_live_check || overlap_contains(a.cd.ov, pr.x, pr.y)) && gid_lookup(a, n_local, pr.x, &ia) && gid_lookup(a, n_local, pr.y, &ib)) {
                    const RecHead ha = load_head(a.rec, ia), hb = load_head(a.rec, ib);
                    const int32_t owner_col = max(ha.sx, hb.sx);
                    if (owner_col >= a.col_lo && owner_col < a.col_hi) {
                        push = true;
                        sv = make_uint2(ia, ib | kSepBit);
                        n_sep += 1;
                        n_col -= overlap_pair_is_candidate(a.cd, ha, hb);
                    }
                }
            } else {
                const WorkItem it = work_item_at(a.work, &s_idx, p);
                RecHead ha, hb;
                DurHb da, db;
                load_rec_keep(a.rec, it.a, keep, &ha, &da);
                load_rec_keep(a.rec, it.b, keep, &hb, &db);
                bool a_is_hi = true;
                const bool geom = candidate_geom(a.cd, ha, hb, it.slot, &a_is_hi);
                n_col += geom;
                if (geom && !is_overlapping(a.cd, ha, hb, a_is_hi)) {
                    // collide_time (solvers.rs:25-34) up to the touch test, on (hi, lo advanced by zero): only the four
                    // advanced fields of the lower id's shape differ from its record
                    const DurHb da0 = advanced(da, 0.0), db0 = advanced(db, 0.0);
                    da.px = a_is_hi ? da.px : da0.px; da.py = a_is_hi ? da.py : da0.py; da.w = a_is_hi ? da.w : da0.w; da.h = a_is_hi ? da.h : da0.h;
                    db.px = a_is_hi ? db0.px : db.px; db.py = a_is_hi ? db0.py : db.py; db.w = a_is_hi ? db0.w : db.w; db.h = a_is_hi ? db0.h : db.h;
                    const double duration = a_is_hi ? fmin(da.dur, db.dur) : fmin(db.dur, da.dur);
                    const Box4 ba = swept_bounds(da, duration), bb = swept_bounds(db, duration);
                    if (!(ba.ok && bb.ok)) {
                        report_error(a.ctr, kFNan, a_is_hi ? ha.gid : hb.gid);
                    } else if (boxes_touch(ba, bb)) {
                        push = true;
                        sv = a_is_hi ? make_uint2(it.a, it.b) : make_uint2(it.b, it.a);
                    }
                }
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, push);
        if (m) {
            uint32_t at = 0;
            if (lane == 0) at = atomicAdd(&s_qn, (uint32_t)__popc(m));
            at = __shfl_sync(0xffffffffu, at, 0);
            if (push) s_q[at + __popc(m & ((1u << lane) - 1u))] = sv;
        }
        __syncthreads();
        // ---- solve a full batch ----
        const uint32_t q = s_qn;  // < 2 * blockDim.x
        if (q >= blockDim.x) {
            solve_batch(q - blockDim.x, blockDim.x);
            __syncthreads();
            if (threadIdx.x == 0) s_qn = q - blockDim.x;
            __syncthreads();
        }
    }
    __syncthreads();
    {
        const uint32_t q = s_qn;
        if (q) solve_batch(0, q);
    }
    flush();
    solve_epilogue(a, best, n_sep, n_col);
}

// ---------------------------------------------------------------------------------------------------------
// Dense cells.  A slot run of L entries holds L (L - 1) / 2 (entry, predecessor) pairs; as work items each costs 16 bytes
// written, 16 read and two 96-byte gathers.  At 1 shape per unit^2 (config 4's density, ~30 entries per run) that was 30 M
// items per million hitboxes: 0.53 ms to enumerate and 1.94 ms to solve.  The dense kernels instead give 32 consecutive
// entries of a run (ranks j0 .. j0 + 31, one per lane) to one warp, which walks ALL their predecessors in chunks of 32
// records staged in shared memory: no work items, no gathers.  Ranks below DenseList.rank0 stay work items (short runs
// are the common case at low density).  Two kernels: k_solve_dense_lockstep (first version, A/B only) and k_solve_dense
// (filter + solve; the one in use).  Measurements and the reasoning: DESIGN.md section 4, "Dense cells".

#ifndef CB200_DENSE_MERGED
#define CB200_DENSE_MERGED 1  /* collide_time_merged in the dense kernels' solve (narrowphase.cuh) */
#endif
// The pair has passed the candidate test: narrowphase + event (the body of k_solve for one candidate).
__device__ __forceinline__ bool solve_pair(const SolveArgs& a, double T, const RecHead& ha, const DurHb& da, const RecHead& hb, const DurHb& db, bool a_is_hi,
                                           PairEvent* ev, MinKey* best) {
    const DurHb A = pick(a_is_hi, da, db);
    // the partner is `other.hitbox_at_time(T)` (collider.rs:478-486): advanced by T - start = 0
    const DurHb B = advanced(pick(a_is_hi, db, da), 0.0);
    const uint32_t gid_hi = a_is_hi ? ha.gid : hb.gid, gid_lo = a_is_hi ? hb.gid : ha.gid;
#if CB200_DENSE_MERGED
    double delay = collide_time_merged(A, B);
#else
    double delay = collide_time(A, B);
#endif
    bool is_sep = false;
    if (a.add_mode && delay == 0.0) {
        // immediate overlap: process_collision (collider.rs:177-197) queues the Separate
        const unsigned long long at = atomicAdd(a.n_imm, 1ull);
        if (at < a.cap_imm) a.imm_pairs[at] = make_uint2(gid_hi, gid_lo);
        delay = separate_time(A, B, a.padding);
        is_sep = true;
    }
    if (delay != delay) report_error(a.ctr, kFNan, gid_hi);
    const double t = T + delay;
    if (!(t < kHighTime)) return false;
    const uint32_t low = a.add_mode ? ((gid_lo << 1) | (is_sep ? 0u : 1u)) : (gid_lo | (is_sep ? 0u : kCollideBit));
    *ev = PairEvent{t, gid_hi, low};
    const MinKey k{time_key(t), kPairClass | ((uint64_t)gid_hi << 32) | low};
    if (key_less(k, *best)) *best = k;
    return true;
}

// One candidate = (record of an entry, record of one of its predecessors in the slot run, the slot): candidate test,
// collide_time, event.  *n_col counts the candidates that passed candidate_geom.  Returns true if *ev was produced.
__device__ __forceinline__ bool solve_candidate(const SolveArgs& a, double T, const RecHead& ha, const DurHb& da, const RecHead& hb, const DurHb& db,
                                                uint32_t slot, unsigned long long* n_col, PairEvent* ev, MinKey* best) {
    bool a_is_hi = true;
    if (!candidate_geom(a.cd, ha, hb, slot, &a_is_hi)) return false;
    *n_col += 1;
    if (is_overlapping(a.cd, ha, hb, a_is_hi)) return false;
    return solve_pair(a, T, ha, da, hb, db, a_is_hi, ev, best);
}

__device__ __forceinline__ void decode_rec(const int4& a, const int4& b, const int4& c, const int4& e, const int4& f, const int4& g, RecHead* h, DurHb* d) {
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

// The walk of one descriptor by one warp.  f(active, ha, da, hb, db, slot) is called with the whole warp converged for every
// predecessor step; `active` lanes hold a pair (their entry, the predecessor of this step).
constexpr int kRecInt4 = sizeof(HbRec) / 16;
template <class F>
__device__ __forceinline__ void dense_walk(const uint2* __restrict__ entries, uint32_t n_entries, const HbRec* __restrict__ rec, uint2 ds,
                                           int4* s_chunk /* this warp's [32 * kRecInt4] */, F&& f) {
    const int lane = threadIdx.x & 31;
    const uint32_t run0 = ds.x, j0 = ds.y;
    const uint32_t slot = __ldg(entries + run0).x & kSlotMask;
    const uint32_t ej = run0 + j0 + (uint32_t)lane;
    uint2 entj = make_uint2(0u, 0u);
    if (ej < n_entries) entj = __ldg(entries + ej);
    const bool valid_j = ej < n_entries && (entj.x & kSlotMask) == slot;  // (the run may end inside the warp's 32 ranks)
    RecHead ha{};
    DurHb da{};
    if (valid_j) {
        const int4* p = reinterpret_cast<const int4*>(rec + entj.y);
        decode_rec(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3), __ldg(p + 4), __ldg(p + 5), &ha, &da);
    }
    const uint32_t n_valid = (uint32_t)__popc(__ballot_sync(0xffffffffu, valid_j));  // lanes 0 .. n_valid - 1 (a run is contiguous)
    const uint32_t i_end = j0 + n_valid - 1u;  // predecessors 0 .. i_end - 1 are needed (n_valid >= 1: the descriptor's own entry)
    for (uint32_t i0 = 0; i0 < i_end; i0 += 32) {
        // stage predecessors i0 .. i0 + 31 (they lie inside the run: i < j for a valid j)
        const uint32_t i = i0 + (uint32_t)lane;
        if (i < i_end) {
            const uint2 enti = __ldg(entries + run0 + i);
            const int4* p = reinterpret_cast<const int4*>(rec + enti.y);
            int4* q = s_chunk + lane * kRecInt4;
#pragma unroll
            for (int k = 0; k < kRecInt4; ++k) q[k] = __ldg(p + k);
        }
        __syncwarp();
        const uint32_t n_i = min(32u, i_end - i0);
        for (uint32_t k = 0; k < n_i; ++k) {
            const bool active = valid_j && (i0 + k) < (j0 + (uint32_t)lane);
            const int4* q = s_chunk + k * kRecInt4;
            RecHead hb;
            DurHb db;
            decode_rec(q[0], q[1], q[2], q[3], q[4], q[5], &hb, &db);
            f(active, ha, da, hb, db, slot);
        }
        __syncwarp();
    }
}

struct DenseArgs {
    DenseList list;
    const uint2* entries;
    uint32_t cap_entries;
    uint32_t partial_at;  // this kernel's first slot in SolveArgs.partial
};
constexpr int kDenseEvBuf = 64;  // pair events a warp collects before it appends 32 of them with one atomic

#ifndef CB200_LB_DENSE
#define CB200_LB_DENSE 2
#endif
// First version (CB200_DENSE_KERNEL=1; measured SLOWER than work items: 3.77 vs 1.94 + 0.54 ms at 1 shape per unit^2): every
// (entry, predecessor) step runs the whole candidate test + narrowphase in lockstep.  The records come from shared memory, but
// the stage is issue-bound there, not gather-bound: ~1000 warp instructions per step with 44 % of the lanes holding a pair
// (lane j only pairs with predecessors i < j), and the hash lookup of the overlap set on every pair.
__global__ void __launch_bounds__(kThreads, CB200_LB_DENSE) k_solve_dense_lockstep(const SolveArgs a, const DenseArgs d) {
    __shared__ int4 s_rec[kThreads / 32][32 * kRecInt4];
    __shared__ PairEvent s_ev[kThreads / 32][kDenseEvBuf];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n_desc = min(*d.list.count, d.list.cap);
    const uint32_t n_entries = (uint32_t)min(a.ctr->n_entries, (unsigned long long)d.cap_entries);
    const double T = a.dyn->T;
    MinKey best{kKeyMax, kKeyMax};
    unsigned long long n_col = 0;
    uint32_t n_buf = 0;  // events in this warp's buffer (the same value in every lane)
    auto append = [&](uint32_t n) {  // the first n buffered events -> the global event list (whole warp)
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&a.ctr->n_pair_events, (unsigned long long)n);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (uint32_t k = lane; k < n; k += 32)
            if (base + k < a.cap_events) a.events[base + k] = s_ev[warp][k];
        __syncwarp();
    };
    const uint32_t n_warps = gridDim.x * (kThreads / 32);
    for (uint32_t q = blockIdx.x * (kThreads / 32) + warp; q < n_desc; q += n_warps) {
        dense_walk(d.entries, n_entries, a.rec, d.list.desc[q], s_rec[warp], [&](bool active, const RecHead& ha, const DurHb& da, const RecHead& hb, const DurHb& db, uint32_t slot) {
            PairEvent ev;
            bool made = false;
            if (active) made = solve_candidate(a, T, ha, da, hb, db, slot, &n_col, &ev, &best);
            const unsigned m = __ballot_sync(0xffffffffu, made);
            if (m) {
                if (made) s_ev[warp][n_buf + (uint32_t)__popc(m & ((1u << lane) - 1u))] = ev;
                n_buf += (uint32_t)__popc(m);
                __syncwarp();
                if (n_buf >= 32u) {
                    append(32u);
                    const uint32_t rest = n_buf - 32u;  // < 32
                    PairEvent keep_ev;
                    if ((uint32_t)lane < rest) keep_ev = s_ev[warp][32 + lane];
                    __syncwarp();
                    if ((uint32_t)lane < rest) s_ev[warp][lane] = keep_ev;
                    __syncwarp();
                    n_buf = rest;
                }
            }
        });
    }
    if (n_buf) append(n_buf);
    best = block_min(best);
    const unsigned long long col_tot = block_sum(n_col);
    if (threadIdx.x == 0) {
        a.partial[d.partial_at + blockIdx.x] = best;
        if (col_tot) atomicAdd(&a.ctr->n_collide, col_tot);
    }
}

// The dense kernel in use: filter, then solve.
//   filter   (lockstep, one predecessor per step, cheap) candidate_geom from the record heads, then the first exit of
//            collide_time (solvers.rs:27-30: the swept bounds of the two hitboxes over min(duration) do not touch -> no
//            event) from four edges per record that are computed ONCE per record when it is staged: the swept bounds over the
//            record's OWN duration.  They are the bounds collide_time would compute whenever the two durations are equal
//            bit for bit (every hitbox whose cell period exceeds the step has duration = end_time - T): then the test here is
//            that very test, same operations on the same operands.  If the durations differ, or either record's bounds are
//            not `ok` (the reference would panic), the pair simply passes.  The partner of the pair is `advanced by 0.0` in
//            collide_time (x + v * 0.0): for finite velocities that only turns -0.0 into +0.0, which no comparison or later
//            sum can see; records with a non-finite velocity pass unfiltered.  Survivors go to a queue in shared memory as
//            (lane of the entry, index of the predecessor in the staged chunk).
//   solve    whenever 32 survivors are queued (and when a chunk of predecessors ends) each lane takes one: both records from
//            shared memory, the overlap-set lookup, the full collide_time — all 32 lanes busy, and only for the ~1/3 of the
//            pairs whose swept bounds touch.
constexpr uint32_t kDenseThreads = 128;
constexpr uint32_t kDenseWarps = kDenseThreads / 32;
#ifndef CB200_LB_DENSE2
#define CB200_LB_DENSE2 4
#endif
__global__ void __launch_bounds__(kDenseThreads, CB200_LB_DENSE2) k_solve_dense(const SolveArgs a, const DenseArgs d) {
    __shared__ int4 s_own[kDenseWarps][32 * kRecInt4];   // the records of the warp's 32 entries (lane j's at j)
    __shared__ int4 s_pred[kDenseWarps][32 * kRecInt4];  // the staged chunk of predecessors
    __shared__ double s_edge[kDenseWarps][32][4];        // their swept edges l, r, b, t
    __shared__ double s_dur[kDenseWarps][32];            // and dur_key
    __shared__ uint16_t s_q[kDenseWarps][1024];          // survivors of a chunk: predecessor index << 5 | lane of the entry
    __shared__ PairEvent s_ev[kDenseWarps][kDenseEvBuf];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n_desc = min(*d.list.count, d.list.cap);
    const uint32_t n_entries = (uint32_t)min(a.ctr->n_entries, (unsigned long long)d.cap_entries);
    const double T = a.dyn->T;
    MinKey best{kKeyMax, kKeyMax};
    unsigned long long n_col = 0;
    uint32_t n_buf = 0;  // events in this warp's buffer (the same value in every lane)
    auto append = [&](uint32_t n) {  // the first n buffered events -> the global event list (whole warp)
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&a.ctr->n_pair_events, (unsigned long long)n);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (uint32_t k = lane; k < n; k += 32)
            if (base + k < a.cap_events) a.events[base + k] = s_ev[warp][k];
        __syncwarp();
    };
    for (;;) {
        // descriptors are handed out one by one (their cost grows with the rank they start at)
        uint32_t q = 0;
        if (lane == 0) q = atomicAdd(d.list.cursor, 1u);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= n_desc) break;
        const uint2 ds = d.list.desc[q];
        const uint32_t run0 = ds.x, j0 = ds.y;
        const uint32_t slot = __ldg(d.entries + run0).x & kSlotMask;
        uint32_t cell_x, cell_y;  // the folded cell of the run: a pair belongs to it iff its owner cell folds to the same
        a.cd.geom.unslot(slot, &cell_x, &cell_y);
        const uint32_t fold_x = (1u << a.cd.geom.lw) - 1u, fold_y = (1u << a.cd.geom.lh) - 1u;
        // ---- this warp's entries: ranks j0 .. j0 + 31 of the run, one per lane ----
        const uint32_t ej = run0 + j0 + (uint32_t)lane;
        uint2 entj = make_uint2(0u, 0u);
        if (ej < n_entries) entj = __ldg(d.entries + ej);
        const bool valid_j = ej < n_entries && (entj.x & kSlotMask) == slot;  // (the run may end inside the warp's 32 ranks)
        RecHead ha{};
        SweptEdges ea{};
        if (valid_j) {
            const int4* p = reinterpret_cast<const int4*>(a.rec + entj.y);
            int4 r[kRecInt4];
#pragma unroll
            for (int k = 0; k < kRecInt4; ++k) r[k] = __ldg(p + k);
            int4* o = s_own[warp] + lane * kRecInt4;
#pragma unroll
            for (int k = 0; k < kRecInt4; ++k) o[k] = r[k];
            DurHb da;
            decode_rec(r[0], r[1], r[2], r[3], r[4], r[5], &ha, &da);
            ea = swept_edges(da);
        }
        const uint32_t group_a = (ha.kgb >> 8) & 31u;
        const uint32_t n_valid = (uint32_t)__popc(__ballot_sync(0xffffffffu, valid_j));  // lanes 0 .. n_valid - 1 (a run is contiguous)
        const uint32_t i_end = j0 + n_valid - 1u;  // predecessors 0 .. i_end - 1 are needed (n_valid >= 1: the descriptor's own entry)
        for (uint32_t i0 = 0; i0 < i_end; i0 += 32) {
            // ---- stage predecessors i0 .. i0 + 31 (they lie inside the run: i < j for a valid j) ----
            const uint32_t i = i0 + (uint32_t)lane;
            if (i < i_end) {
                const uint2 enti = __ldg(d.entries + run0 + i);
                const int4* p = reinterpret_cast<const int4*>(a.rec + enti.y);
                int4 r[kRecInt4];
#pragma unroll
                for (int k = 0; k < kRecInt4; ++k) r[k] = __ldg(p + k);
                int4* o = s_pred[warp] + lane * kRecInt4;
#pragma unroll
                for (int k = 0; k < kRecInt4; ++k) o[k] = r[k];
                RecHead hp;
                DurHb dp;
                decode_rec(r[0], r[1], r[2], r[3], r[4], r[5], &hp, &dp);
                const SweptEdges ep = swept_edges(dp);
                s_edge[warp][lane][0] = ep.l; s_edge[warp][lane][1] = ep.r; s_edge[warp][lane][2] = ep.b; s_edge[warp][lane][3] = ep.t;
                s_dur[warp][lane] = ep.dur_key;
            }
            __syncwarp();
            const uint32_t n_i = min(32u, i_end - i0);
            // ---- filter: (lane's entry, predecessor i0 + k) for every k; the survivors of this lane are bits of `keep` ----
            // candidate_geom with the slot test written as "the owner cell folds to the run's cell" (GridGeom::slot is a
            // bijection on the folded coordinates), then the swept-edge test
            uint32_t keep = 0u;
            const uint32_t rank_j = j0 + (uint32_t)lane;
            const uint32_t k_end = (!valid_j || rank_j <= i0) ? 0u : min(n_i, rank_j - i0);  // predecessors of this lane in the chunk
            for (uint32_t k = 0; k < n_i; ++k) {
                const int4 h0 = s_pred[warp][k * kRecInt4], h1 = s_pred[warp][k * kRecInt4 + 1];
                const double2 e0 = *reinterpret_cast<const double2*>(&s_edge[warp][k][0]), e1 = *reinterpret_cast<const double2*>(&s_edge[warp][k][2]);
                const double dk = s_dur[warp][k];
                const int32_t bsx = h0.x, bsy = h0.y;
                const int32_t bex = bsx + (int32_t)((uint32_t)h0.z & 0xffffu), bey = bsy + (int32_t)((uint32_t)h0.z >> 16);
                const uint32_t bgid = (uint32_t)h0.w, bkgb = (uint32_t)h1.z, bmask = (uint32_t)h1.w;
                const int32_t ox = max(ha.sx, bsx), oy = max(ha.sy, bsy);
                const bool first_hi = ha.gid > bgid;
                const uint32_t mask_hi = first_hi ? ha.mask : bmask, group_lo = first_hi ? ((bkgb >> 8) & 31u) : group_a;
                const bool geom = (k < k_end) & (ox < min(ha.ex, bex)) & (oy < min(ha.ey, bey)) & (((uint32_t)ox & fold_x) == cell_x) &
                                  (((uint32_t)oy & fold_y) == cell_y) & (ox >= a.cd.col_lo) & (ox < a.cd.col_hi) & (((mask_hi >> group_lo) & 1u) != 0u);
                const bool touch = (ea.r - e0.x >= 0.0) & (ea.t - e1.x >= 0.0) & (e0.y - ea.l >= 0.0) & (e1.y - ea.b >= 0.0);
                const bool pass = geom & (touch | !(ea.dur_key == dk));
                n_col += geom;
                keep |= (pass ? 1u : 0u) << k;
            }
            // ---- queue: every lane writes its survivors behind those of the lanes below it ----
            uint32_t incl = (uint32_t)__popc(keep);
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += u;
            }
            const uint32_t n_q = __shfl_sync(0xffffffffu, incl, 31);
            for (uint32_t at = incl - (uint32_t)__popc(keep); keep; keep &= keep - 1u) s_q[warp][at++] = (uint16_t)(((uint32_t)(__ffs(keep) - 1) << 5) | (uint32_t)lane);
            __syncwarp();
            // ---- solve: 32 survivors at a time ----
            for (uint32_t first = 0; first < n_q; first += 32) {
                PairEvent ev;
                bool made = false;
                if (first + (uint32_t)lane < n_q) {
                    const uint32_t e = s_q[warp][first + lane];
                    const int4* pa = s_own[warp] + (e & 31u) * kRecInt4;
                    const int4* pb = s_pred[warp] + (e >> 5) * kRecInt4;
                    RecHead h1, h2;
                    DurHb d1, d2;
                    decode_rec(pa[0], pa[1], pa[2], pa[3], pa[4], pa[5], &h1, &d1);
                    decode_rec(pb[0], pb[1], pb[2], pb[3], pb[4], pb[5], &h2, &d2);
                    const bool first_hi = h1.gid > h2.gid;
                    if (!is_overlapping(a.cd, h1, h2, first_hi)) made = solve_pair(a, T, h1, d1, h2, d2, first_hi, &ev, &best);
                }
                const unsigned me = __ballot_sync(0xffffffffu, made);
                if (made) s_ev[warp][n_buf + (uint32_t)__popc(me & ((1u << lane) - 1u))] = ev;
                n_buf += (uint32_t)__popc(me);
                __syncwarp();
                if (n_buf >= 32u) {
                    append(32u);
                    const uint32_t rest = n_buf - 32u;  // < 32
                    PairEvent keep_ev;
                    if ((uint32_t)lane < rest) keep_ev = s_ev[warp][32 + lane];
                    __syncwarp();
                    if ((uint32_t)lane < rest) s_ev[warp][lane] = keep_ev;
                    __syncwarp();
                    n_buf = rest;
                }
            }
            __syncwarp();  // (s_pred and s_q are rewritten by the next chunk)
        }
        __syncwarp();  // (s_own is rewritten by the next descriptor)
    }
    if (n_buf) append(n_buf);
    best = block_min(best);
    const unsigned long long col_tot = block_sum(n_col);
    if (threadIdx.x == 0) {
        a.partial[d.partial_at + blockIdx.x] = best;
        if (col_tot) atomicAdd(&a.ctr->n_collide, col_tot);
    }
}

// Debug / parity view of the dense part (see k_list_candidates).
__global__ void __launch_bounds__(kThreads) k_list_candidates_dense(const CandArgs a, const DenseArgs d, const Counters* ctr, const HbRec* __restrict__ rec, uint2* out,
                                                                    unsigned long long* n_out, uint64_t cap) {
    __shared__ int4 s_rec[kThreads / 32][32 * kRecInt4];
    const int warp = threadIdx.x >> 5;
    const uint32_t n_desc = min(*d.list.count, d.list.cap);
    const uint32_t n_entries = (uint32_t)min(ctr->n_entries, (unsigned long long)d.cap_entries);
    const uint32_t n_warps = gridDim.x * (kThreads / 32);
    for (uint32_t q = blockIdx.x * (kThreads / 32) + warp; q < n_desc; q += n_warps) {
        dense_walk(d.entries, n_entries, rec, d.list.desc[q], s_rec[warp], [&](bool active, const RecHead& ha, const DurHb&, const RecHead& hb, const DurHb&, uint32_t slot) {
            bool a_is_hi = true;
            if (active && candidate_test(a, ha, hb, slot, &a_is_hi)) {
                const unsigned long long pos = warp_agg_claim(n_out);
                if (pos < cap) out[pos] = a_is_hi ? make_uint2(ha.gid, hb.gid) : make_uint2(hb.gid, ha.gid);
            }
        });
    }
}

// ---------------------------------------------------------------------------------------------------------
// Batched narrowphase on explicit DurHitbox pairs (n x 9 doubles: px,py,dw,dh,vx,vy,rw,rh,duration).
__device__ __forceinline__ DurHb load_dur9(const double* __restrict__ r, uint8_t kind) {
    DurHb d;
    d.px = r[0]; d.py = r[1]; d.w = r[2]; d.h = r[3]; d.vx = r[4]; d.vy = r[5]; d.rw = r[6]; d.rh = r[7]; d.dur = r[8];
    d.kind = kind;
    return d;
}
__global__ void k_collide_batch(uint64_t n, const double* __restrict__ a, const uint8_t* __restrict__ ka, const double* __restrict__ b,
                                const uint8_t* __restrict__ kb, double* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = collide_time(load_dur9(a + 9 * i, ka[i]), load_dur9(b + 9 * i, kb[i]));
}
__global__ void k_separate_batch(uint64_t n, const double* __restrict__ a, const uint8_t* __restrict__ ka, const double* __restrict__ b,
                                 const uint8_t* __restrict__ kb, double padding, double* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = separate_time(load_dur9(a + 9 * i, ka[i]), load_dur9(b + 9 * i, kb[i]), padding);
}

}  // namespace cb200
