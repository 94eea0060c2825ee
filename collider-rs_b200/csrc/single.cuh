// Single-hitbox operations against the device-resident table and grid: the device half of the reference's
// incremental API —
//   Collider::add_hitbox / set_hitbox_vel / Reiterate  -> k_update_one   (collider.rs:214-250, 320-381, 420-448)
//   Collider::process_event (Collide / Separate)       -> k_pair_solve   (collider.rs:126-197)
//   Collider::get_hitbox                               -> k_get_hitbox   (collider.rs:200-202, 488-497)
//   Collider::query_overlaps                           -> k_query        (collider.rs:309-318, grid.rs:74-77)
//   Collider::remove_hitbox                            -> k_remove_row   (collider.rs:256-275)
//
// The grid these operations see is the one the last bulk step (or k_rebin_count + scan + scatter) built — the
// slot-sorted entry array — plus the *dirty list*: rows whose hitbox changed since that build.  A dirty row's
// entries in the sorted array are stale and are skipped; dirty rows are tested directly through their current
// IndexRect in rec[] (two rects share a cell iff they intersect).  The host re-bins when the list grows long.
// Every pair found in the sorted array is taken exactly once by the same ownership rule as the bulk candidate
// stage (the min corner of the rect intersection must be the cell being walked).
#pragma once
#include "kernels.cuh"

namespace cb200 {

constexpr uint32_t kRowNone = 0xFFFFFFFFu;

// Item kinds reported by k_update_one
constexpr uint32_t kItemCollide = 1;    // Collide(self, other) at `time`
constexpr uint32_t kItemSeparate = 2;   // Separate(self, other) at `time`  (other is a current overlap)
constexpr uint32_t kItemImmediate = 3;  // add only: collide delay == 0.0 -> overlap now; `time` = T + separate_time (collider.rs:355-365)

struct OneItem {
    uint32_t label, kind;
    double time;  // T + delay (may be >= 1e50 / inf: the host drops those like EventManager::new_event_key)
};
struct OneHeader {
    uint32_t n_items;    // items produced (may exceed the capacity: overflow)
    uint32_t err_flags;  // kF*
    uint32_t noop;       // set_hitbox_vel with an identical HbVel (collider.rs:226)
    uint32_t reit;       // a Reiterate event is due at end_time
    double end_time;     // internal end_time
    uint32_t n_dirty;    // length of the dirty list after this operation
    uint32_t n_rows;
    double hb[9];        // k_get_hitbox: px py w h vx vy rw rh pub_end
    uint32_t kind, pad;
};

struct GridView {
    const CellEntry* entries;
    const uint32_t* slot_end;    // after the scatter: end of each slot's run
    const uint32_t* chunk_base;  // start of each scan chunk's runs (slot_run_start)
    GridGeom geom;
    int have_grid;
};

struct TableView {
    StateCols s;
    uint32_t* label;
    uint32_t* row_of;  // label -> row
    HbRec* rec;
    int4* bink;             // (sx, sy, span, kgb) of every record (the scatter's input)
    uint8_t* dirty;         // row is in dirty_list
    uint32_t* dirty_list;
    uint32_t* dirty_count;
    uint32_t n;             // live rows
};

enum { kOpUpdate = 0, kOpAdd = 1, kOpEnumerate = 2 };  // enumerate: phase 2 only (re-run after an item overflow)

struct OneArgs {
    TableView t;
    GridView g;
    int op;
    uint32_t label;       // self
    uint32_t row_new;     // add: the row to create (== t.n before the call)
    int has_vel;          // update: install `vel` (0 = Reiterate: keep the velocity, end_time = stored pub_end_time)
    int check_noop;       // set_hitbox_vel: compare with the stored HbVel first
    double px, py, w, h;  // add: Hitbox.value
    uint32_t kind, group, mask;
    double vx, vy, rw, rh, end_time;
    double T, cell_width, padding;
    const uint32_t* ov_labels;  // labels of the hitboxes overlapping self now
    uint32_t n_ov;
    OneHeader* head;  // mapped pinned
    OneItem* items;   // mapped pinned
    uint32_t cap_items;
};

// HitboxInfo::hitbox_at_time(T) of row j (collider.rs:478-486): shape advanced by T - start, duration end - T.
__device__ __forceinline__ DurHb row_at_time(const StateCols& s, uint32_t j, double T, uint32_t* err) {
    const double start = s.start[j], end = s.end[j];
    if (!(T >= start && T <= end)) *err |= kFInvalidTime;
    const double dt = T - start;
    if (!(dt < kHighTime)) *err |= kFHighTime;
    DurHb d;
    d.px = s.px[j]; d.py = s.py[j]; d.w = s.w[j]; d.h = s.h[j];
    d.vx = s.vx[j]; d.vy = s.vy[j]; d.rw = s.rw[j]; d.rh = s.rh[j];
    d.kind = (int)(s.kind[j] & 1u);
    d = advanced(d, dt);
    d.dur = end - T;
    return d;
}

__device__ __forceinline__ void mark_dirty(const TableView& t, uint32_t row) {
    if (!t.dirty[row]) {
        t.dirty[row] = 1;
        t.dirty_list[atomicAdd(t.dirty_count, 1u)] = row;
    }
}

struct SelfInfo {
    DurHb d;
    int32_t sx, sy, ex, ey;
    uint32_t row, mask, binned, label;
};

// Calls f(row_j, head_j) once for every live, binned hitbox j != self whose CURRENT IndexRect shares a cell with
// [sx,ex) x [sy,ey) — Grid::overlapping_ids (grid.rs:115-135) without the group filter.  All threads of the block.
template <class F>
__device__ __forceinline__ void for_each_cellmate(const TableView& t, const GridView& g, int32_t sx, int32_t sy, int32_t ex, int32_t ey,
                                                  uint32_t self_row, F&& f) {
    if (g.have_grid) {
        const uint32_t ny = (uint32_t)(ey - sy), ncells = (uint32_t)(ex - sx) * ny;
        for (uint32_t c = threadIdx.x; c < ncells; c += blockDim.x) {
            const int32_t x = sx + (int32_t)(c / ny), y = sy + (int32_t)(c % ny);
            const uint32_t s = g.geom.slot(x, y);
            const uint32_t e0 = slot_run_start(g.slot_end, g.chunk_base, s), e1 = g.slot_end[s];
            for (uint32_t e = e0; e < e1; ++e) {
                const uint32_t v = g.entries[e].vidx;
                // rows that moved, appeared or changed since the entry array was built are dirty; every other row's record
                // still holds the IndexRect it was binned with
                if (v >= t.n || v == self_row || t.dirty[v]) continue;
                const RecHead hj = load_head_nc(t.rec, v);
                if (!(hj.kgb & 2u)) continue;
                // ownership: (x, y) must be the min corner of the rect intersection (also rejects aliased slot-mates)
                if (max(sx, hj.sx) != x || max(sy, hj.sy) != y) continue;
                if (x >= min(ex, hj.ex) || y >= min(ey, hj.ey)) continue;
                f(v, hj);
            }
        }
    }
    const uint32_t nd = *t.dirty_count;
    for (uint32_t k = threadIdx.x; k < nd; k += blockDim.x) {
        const uint32_t v = t.dirty_list[k];
        if (v >= t.n || v == self_row) continue;
        const RecHead hj = load_head_nc(t.rec, v);
        if (!(hj.kgb & 2u)) continue;
        if (max(sx, hj.sx) >= min(ex, hj.ex) || max(sy, hj.sy) >= min(ey, hj.ey)) continue;
        f(v, hj);
    }
}

__global__ void __launch_bounds__(kThreads) k_update_one(const OneArgs a) {
    __shared__ SelfInfo self;
    __shared__ uint32_t s_count, s_err, s_stop;
    const TableView& t = a.t;
    if (threadIdx.x == 0) {
        s_count = 0;
        s_err = 0;
        s_stop = 0;
        OneHeader hd;
        hd.n_items = 0; hd.err_flags = 0; hd.noop = 0; hd.reit = 0; hd.end_time = 0.0; hd.kind = 0; hd.pad = 0;
        const uint32_t row = a.op == kOpAdd ? a.row_new : t.row_of[a.label];
        uint32_t err = 0;
        if (a.op != kOpEnumerate) {
            HbCore h;
            NewVel nv;
            if (a.op == kOpAdd) {
                h.px = a.px; h.py = a.py; h.w = a.w; h.h = a.h;
                h.vx = a.vx; h.vy = a.vy; h.rw = a.rw; h.rh = a.rh;
                h.start = a.T; h.pub_end = a.end_time;  // HitboxInfo::new (collider.rs:467-476)
                h.kind = a.kind; h.group = a.group; h.mask = a.mask;
                nv.has_vx = nv.has_vy = nv.has_rw = nv.has_rh = false;
                nv.vx = nv.vy = nv.rw = nv.rh = 0.0;
                nv.end_time = a.end_time;
                if (!(a.w >= 0.0 && a.h >= 0.0) || (a.kind == (uint32_t)kCircle && !(a.w == a.h))) err |= kFTooSmall;  // Shape::new
            } else {
                h.px = t.s.px[row]; h.py = t.s.py[row]; h.w = t.s.w[row]; h.h = t.s.h[row];
                h.vx = t.s.vx[row]; h.vy = t.s.vy[row]; h.rw = t.s.rw[row]; h.rh = t.s.rh[row];
                h.start = t.s.start[row]; h.pub_end = t.s.pub_end[row];
                h.kind = t.s.kind[row]; h.group = t.s.group[row]; h.mask = t.s.mask[row];
                if (a.check_noop && h.vx == a.vx && h.vy == a.vy && h.rw == a.rw && h.rh == a.rh && t.s.end[row] == a.end_time) {
                    hd.noop = 1;  // Collider::set_hitbox_vel: `if self.hitboxes[&id].hitbox.vel != vel` (collider.rs:226)
                    s_stop = 1;
                }
                nv.has_vx = nv.has_vy = nv.has_rw = nv.has_rh = a.has_vel != 0;
                nv.vx = a.vx; nv.vy = a.vy; nv.rw = a.rw; nv.rh = a.rh;
                nv.end_time = a.has_vel ? a.end_time : h.pub_end;  // Reiterate re-installs the public end_time (collider.rs:236)
            }
            if (!s_stop) {
                err |= hb_update_core(h, a.T, a.op != kOpAdd, nv, a.cell_width, a.padding, a.g.geom);
                if (err) {
                    s_stop = 1;
                } else {
                    t.s.px[row] = h.px; t.s.py[row] = h.py; t.s.w[row] = h.w; t.s.h[row] = h.h;
                    t.s.vx[row] = h.vx; t.s.vy[row] = h.vy; t.s.rw[row] = h.rw; t.s.rh[row] = h.rh;
                    t.s.start[row] = h.start; t.s.end[row] = h.end; t.s.pub_end[row] = h.pub_end;
                    t.s.reit[row] = h.reit ? 1 : 0;
                    if (a.op == kOpAdd) {
                        t.s.kind[row] = (uint8_t)a.kind; t.s.group[row] = a.group; t.s.mask[row] = a.mask;
                        t.label[row] = a.label;
                        t.row_of[a.label] = row;
                    }
                    const HbRec nr = make_rec(h, a.label);
                    store_rec(t.rec + row, nr);
                    t.bink[row] = make_int4(nr.sx, nr.sy, (int)nr.span, (int)nr.kgb);
                    mark_dirty(t, row);
                }
            }
        }
        hd.err_flags = err;
        s_err = err;
        if (!s_stop) {
            // phase 2 works from the stored row (also the path of a re-run)
            const RecHead hs = load_head_nc(t.rec, row);
            self.d = load_body_nc(t.rec, row, hs);
            self.sx = hs.sx; self.sy = hs.sy; self.ex = hs.ex; self.ey = hs.ey;
            self.row = row; self.mask = hs.mask; self.binned = (hs.kgb & 2u) ? 1u : 0u; self.label = hs.gid;
            hd.reit = t.s.reit[row];
            hd.end_time = t.s.end[row];
        }
        *a.head = hd;
    }
    __syncthreads();
    if (s_stop) {
        if (threadIdx.x == 0) {
            a.head->n_dirty = *t.dirty_count;
            a.head->n_rows = t.n + (a.op == kOpAdd && !s_err ? 1u : 0u);
            __threadfence_system();
        }
        return;
    }
    const TableView tv{t.s, t.label, t.row_of, t.rec, t.bink, t.dirty, t.dirty_list, t.dirty_count, t.n + (a.op == kOpAdd ? 1u : 0u)};
    if (self.binned) {  // collider.rs:328: only hitboxes with a group are tracked
        const DurHb me = self.d;
        // Separate events for the current overlaps (collider.rs:329-339)
        for (uint32_t k = threadIdx.x; k < a.n_ov; k += blockDim.x) {
            const uint32_t lj = a.ov_labels[k], j = t.row_of[lj];
            uint32_t err = 0;
            const DurHb other = row_at_time(t.s, j, a.T, &err);
            const double delay = separate_time(me, other, a.padding);
            if (delay != delay) err |= kFNan;
            if (err) atomicOr(&s_err, err);
            const uint32_t pos = atomicAdd(&s_count, 1u);
            if (pos < a.cap_items) a.items[pos] = OneItem{lj, kItemSeparate, a.T + delay};
        }
        // Collide events for the cell-mates (collider.rs:340-372)
        for_each_cellmate(tv, a.g, self.sx, self.sy, self.ex, self.ey, self.row, [&](uint32_t j, const RecHead& hj) {
            const uint32_t group_j = (hj.kgb >> 8) & 0xffu;
            if (group_j > 31u || !((self.mask >> group_j) & 1u)) return;  // grid key group must be one of self's interact_groups
            for (uint32_t k = 0; k < a.n_ov; ++k)
                if (a.ov_labels[k] == hj.gid) return;  // already overlapping (collider.rs:351)
            uint32_t err = 0;
            const DurHb other = row_at_time(t.s, j, a.T, &err);
            const double delay = collide_time(me, other);
            if (delay != delay) err |= kFNan;
            OneItem it{hj.gid, kItemCollide, a.T + delay};
            if (a.op == kOpAdd && delay == 0.0) {
                const double sep = separate_time(me, other, a.padding);
                if (sep != sep) err |= kFNan;
                it.kind = kItemImmediate;
                it.time = a.T + sep;
            }
            if (err) atomicOr(&s_err, err);
            const uint32_t pos = atomicAdd(&s_count, 1u);
            if (pos < a.cap_items) a.items[pos] = it;
        });
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        a.head->n_items = s_count;
        a.head->err_flags = s_err;
        a.head->n_dirty = *t.dirty_count;
        a.head->n_rows = tv.n;
        __threadfence_system();
    }
}

// One narrowphase solve between two table rows at time T: Collider::process_collision's separate_time
// (collider.rs:177-197) or the collide_time after a Separate event (collider.rs:148-158).  out[0] = T + delay.
__global__ void k_pair_solve(TableView t, uint32_t label_a, uint32_t label_b, int separate, double T, double padding, OneHeader* head) {
    uint32_t err = 0;
    const DurHb A = row_at_time(t.s, t.row_of[label_a], T, &err), B = row_at_time(t.s, t.row_of[label_b], T, &err);
    const double delay = separate ? separate_time(A, B, padding) : collide_time(A, B);
    if (delay != delay) err |= kFNan;
    head->end_time = T + delay;
    head->err_flags = err;
    __threadfence_system();
}

// The same for a batch of queued events, each at its own event time: the host prefetches the solves of the next events in
// one launch instead of one launch per event (collider.inl prefetch_solves); a result is used only if neither hitbox
// changed in between.
struct PairJob {
    uint32_t label_a, label_b;
    uint32_t separate, pad;
    double t;
};
struct PairOut {
    double time;  // t + delay
    uint32_t err, pad;
    // Collider::get_hitbox of both hitboxes at the event time (what the user's handler asks for next): pos, dims, vel, resize,
    // public end_time; kind; contract flags
    double hb_a[9], hb_b[9];
    uint32_t kind_a, kind_b, err_a, err_b;
};
// HitboxInfo::pub_hitbox_at_time(T) (collider.rs:488-497) of row j; returns the contract flags.
__device__ __forceinline__ uint32_t pub_hitbox_at(const StateCols& s, uint32_t j, double T, double* hb, uint32_t* kind) {
    uint32_t err = 0;
    const double start = s.start[j], pub_end = s.pub_end[j];
    if (!(T >= start && T <= pub_end)) err |= kFInvalidTime;
    const double dt = T - start;
    if (!(dt < kHighTime)) err |= kFHighTime;
    const double vx = s.vx[j], vy = s.vy[j], rw = s.rw[j], rh = s.rh[j];
    hb[0] = s.px[j] + vx * dt;
    hb[1] = s.py[j] + vy * dt;
    hb[2] = s.w[j] + rw * dt;
    hb[3] = s.h[j] + rh * dt;
    hb[4] = vx; hb[5] = vy; hb[6] = rw; hb[7] = rh;
    hb[8] = pub_end;
    *kind = s.kind[j] & 1u;
    return err;
}
__global__ void __launch_bounds__(128) k_pair_solve_batch(TableView t, const PairJob* __restrict__ jobs, uint32_t n, double padding, PairOut* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const PairJob j = jobs[i];
    uint32_t err = 0;
    const uint32_t ra = t.row_of[j.label_a], rb = t.row_of[j.label_b];
    if (ra >= t.n || rb >= t.n) {
        out[i].time = 0.0;
        out[i].err = kFIdNotFound;
        out[i].err_a = out[i].err_b = kFIdNotFound;
        return;
    }
    const DurHb A = row_at_time(t.s, ra, j.t, &err), B = row_at_time(t.s, rb, j.t, &err);
    const double delay = j.separate ? separate_time(A, B, padding) : collide_time(A, B);
    if (delay != delay) err |= kFNan;
    PairOut o;
    o.time = j.t + delay;
    o.err = err;
    o.pad = 0;
    o.err_a = pub_hitbox_at(t.s, ra, j.t, o.hb_a, &o.kind_a);
    o.err_b = pub_hitbox_at(t.s, rb, j.t, o.hb_b, &o.kind_b);
    out[i] = o;
}

// Collider::get_hitbox = HitboxInfo::pub_hitbox_at_time(T) (collider.rs:200-202, 488-497).
__global__ void k_get_hitbox(TableView t, uint32_t label, double T, OneHeader* head) {
    uint32_t kind;
    head->err_flags = pub_hitbox_at(t.s, t.row_of[label], T, head->hb, &kind);
    head->kind = kind;
    __threadfence_system();
}

// PlacedShape::overlaps (geom/shape/mod.rs:154-177 -> normals.rs:22-51): normal_from(...).len() >= 0.
struct PShape {
    double px, py, w, h;
    int kind;
};
__device__ __forceinline__ double box_box_normal_len(const PShape& d, const PShape& s) {
    // min over Card::values (MinusX, MinusY, PlusX, PlusY) of dst.card_overlap(src, card) = src.edge(card) + dst.edge(flip card)
    const double d_l = d.px - d.w * 0.5, d_r = d.px + d.w * 0.5, d_b = d.py - d.h * 0.5, d_t = d.py + d.h * 0.5;
    const double s_l = s.px - s.w * 0.5, s_r = s.px + s.w * 0.5, s_b = s.py - s.h * 0.5, s_t = s.py + s.h * 0.5;
    return fmin(fmin(d_r - s_l, d_t - s_b), fmin(s_r - d_l, s_t - d_b));
}
__device__ __forceinline__ double disc_disc_normal_len(double dpx, double dpy, double dd, double spx, double spy, double sd) {
    const double x = dpx - spx, y = dpy - spy;
    return (sd + dd) * 0.5 - sqrt(x * x + y * y);
}
__device__ __forceinline__ double box_disc_normal_len(const PShape& box, const PShape& disc) {
    const double l = box.px - box.w * 0.5, r = box.px + box.w * 0.5, b = box.py - box.h * 0.5, t = box.py + box.h * 0.5;
    const int sx = disc.px < l ? -1 : (disc.px > r ? 1 : 0), sy = disc.py < b ? -1 : (disc.py > t ? 1 : 0);
    if (sx != 0 && sy != 0) return disc_disc_normal_len(sx < 0 ? l : r, sy < 0 ? b : t, 0.0, disc.px, disc.py, disc.w);
    return box_box_normal_len(box, disc);
}
__device__ __forceinline__ bool shapes_overlap(const PShape& a, const PShape& b) {
    double len;
    if (a.kind == kRect && b.kind == kRect) len = box_box_normal_len(a, b);
    else if (a.kind == kRect) len = box_disc_normal_len(a, b);
    else if (b.kind == kRect) len = box_disc_normal_len(b, a);
    else len = disc_disc_normal_len(a.px, a.py, a.w, b.px, b.py, b.w);
    return len >= 0.0;
}

// Collider::query_overlaps (collider.rs:309-318): cell-mates of the shape's IndexRect (Grid::shape_cellmates,
// grid.rs:74-77) in the profile's interact_groups whose shape at T overlaps `q`.  Labels out (unordered).
struct QueryArgs {
    TableView t;
    GridView g;
    PShape q;
    uint32_t mask;
    double T, cell_width;
    OneHeader* head;
    uint32_t* out;  // mapped pinned
    uint32_t cap;
};
// One query, run by one block (all threads call; barriers inside).  Results: labels (unordered) in out[0 .. min(count, cap)).
__device__ __forceinline__ void query_block(const TableView& t, const GridView& g, const PShape& q, uint32_t mask, double T, double cell_width, uint32_t* out,
                                            uint32_t cap, OneHeader* head) {
    __shared__ uint32_t s_count, s_err;
    __shared__ int32_t s_rect[4];
    if (threadIdx.x == 0) {
        s_count = 0;
        s_err = 0;
        // Grid::index_bounds (grid.rs:101-113)
        const double min_x = q.px - q.w * 0.5, max_x = q.px + q.w * 0.5, min_y = q.py - q.h * 0.5, max_y = q.py + q.h * 0.5;
        const int32_t sx = f64_as_i32(floor(min_x / cell_width)), sy = f64_as_i32(floor(min_y / cell_width));
        const int32_t ex = max(f64_as_i32(ceil(max_x / cell_width)), (int32_t)((uint32_t)sx + 1u));
        const int32_t ey = max(f64_as_i32(ceil(max_y / cell_width)), (int32_t)((uint32_t)sy + 1u));
        s_rect[0] = sx; s_rect[1] = sy; s_rect[2] = ex; s_rect[3] = ey;
        const int64_t nx = (int64_t)ex - sx, ny = (int64_t)ey - sy;
        if (nx <= 0 || ny <= 0) s_err = kFEmptyRect;
        else if (nx > (1ll << g.geom.lw) || ny > (1ll << g.geom.lh) || nx > 0xffff || ny > 0xffff) s_err = kFRectSpan;
    }
    __syncthreads();
    if (!s_err) {
        for_each_cellmate(t, g, s_rect[0], s_rect[1], s_rect[2], s_rect[3], kRowNone, [&](uint32_t j, const RecHead& hj) {
            const uint32_t group_j = (hj.kgb >> 8) & 0xffu;
            if (group_j > 31u || !((mask >> group_j) & 1u)) return;
            uint32_t err = 0;
            const double start = t.s.start[j], pub_end = t.s.pub_end[j];
            if (!(T >= start && T <= pub_end)) err |= kFInvalidTime;
            const double dt = T - start;
            if (!(dt < kHighTime)) err |= kFHighTime;
            PShape p;
            p.px = t.s.px[j] + t.s.vx[j] * dt;
            p.py = t.s.py[j] + t.s.vy[j] * dt;
            p.w = t.s.w[j] + t.s.rw[j] * dt;
            p.h = t.s.h[j] + t.s.rh[j] * dt;
            p.kind = (int)(t.s.kind[j] & 1u);
            if (err) atomicOr(&s_err, err);
            if (!shapes_overlap(p, q)) return;
            const uint32_t pos = atomicAdd(&s_count, 1u);
            if (pos < cap) out[pos] = hj.gid;
        });
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        head->n_items = s_count;
        head->err_flags = s_err;
        __threadfence_system();
    }
}
__global__ void __launch_bounds__(kThreads) k_query(const QueryArgs a) { query_block(a.t, a.g, a.q, a.mask, a.T, a.cell_width, a.out, a.cap, a.head); }

// Many queries in one launch, one block each (Collider::query_overlaps for a batch of shapes): query b reads items[b] and
// writes its labels to out[b * cap_each ...] and its count / flags to heads[b].
struct QueryItem {
    PShape q;
    uint32_t mask;
};
struct QueryBatchArgs {
    TableView t;
    GridView g;
    const QueryItem* items;
    double T, cell_width;
    OneHeader* heads;  // mapped pinned
    uint32_t* out;     // mapped pinned, [n * cap_each]
    uint32_t cap_each;
};
__global__ void __launch_bounds__(kThreads) k_query_batch(const QueryBatchArgs a) {
    const QueryItem it = a.items[blockIdx.x];
    query_block(a.t, a.g, it.q, it.mask, a.T, a.cell_width, a.out + (size_t)blockIdx.x * a.cap_each, a.cap_each, a.heads + blockIdx.x);
}

// Collider::remove_hitbox, table part: the last row moves into the hole (its sorted-array entries go stale, so it
// becomes dirty); the removed label no longer maps to a row.
__global__ void k_remove_row(TableView t, uint32_t label, OneHeader* head) {
    const uint32_t row = t.row_of[label], last = t.n - 1;
    t.row_of[label] = kRowNone;
    if (row != last) {
        double* cols[11] = {t.s.px, t.s.py, t.s.w, t.s.h, t.s.vx, t.s.vy, t.s.rw, t.s.rh, t.s.start, t.s.end, t.s.pub_end};
        for (int k = 0; k < 11; ++k) cols[k][row] = cols[k][last];
        t.s.kind[row] = t.s.kind[last]; t.s.group[row] = t.s.group[last]; t.s.mask[row] = t.s.mask[last]; t.s.reit[row] = t.s.reit[last];
        const uint32_t moved = t.label[last];
        t.label[row] = moved;
        t.row_of[moved] = row;
        const int4* src = reinterpret_cast<const int4*>(t.rec + last);
        int4* dst = reinterpret_cast<int4*>(t.rec + row);
        for (int k = 0; k < 6; ++k) dst[k] = src[k];
        t.bink[row] = t.bink[last];
        mark_dirty(t, row);
    }
    head->n_dirty = *t.dirty_count;
    head->n_rows = last;
    head->err_flags = 0;
    __threadfence_system();
}

// Dense relabel (the id -> label ranks changed): label column, records and row_of follow.
__global__ void __launch_bounds__(kThreads) k_relabel(uint32_t n, const uint32_t* __restrict__ new_of_old, uint32_t* label, HbRec* rec, uint32_t* row_of) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t l = new_of_old[label[i]];
    label[i] = l;
    rec[i].gid = l;
    row_of[l] = i;
}

// Re-bin: count the cells of every live binned row from its current record (then k_scan_slots + k_scatter); clears
// the dirty flags.  The counting half of Grid::update_area for the whole table (grid.rs:137-175).
__global__ void __launch_bounds__(kThreads) k_rebin_count(uint32_t n, const HbRec* __restrict__ rec, uint32_t* slot_count, GridGeom geom, uint8_t* dirty,
                                                          uint32_t* dirty_count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *dirty_count = 0u;
    if (i >= n) return;
    dirty[i] = 0;
    const RecHead h = load_head_nc(rec, i);
    if (!(h.kgb & 2u)) return;
    for (int32_t x = h.sx; x < h.ex; ++x)
        for (int32_t y = h.sy; y != h.ey; ++y) atomicAdd(&slot_count[geom.slot(x, y)], 1u);
}

}  // namespace cb200
