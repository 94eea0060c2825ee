// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_geom.hpp).  extern "C" surface over the C++
// restatement so that tests/ and bench.py's cpu_baseline leg can drive it through ctypes.
#include <cstring>
#include <memory>

#include "oracle_collider.hpp"

using namespace oracle;

namespace {
struct Handle {
    std::unique_ptr<Collider> c;
    std::string err;
};
thread_local std::string g_err;

Profile make_profile(uint64_t id, int64_t group, uint32_t interact_mask) {
    Profile p;
    p.id = id;
    p.has_group = group >= 0;
    p.group = group >= 0 ? (HbGroup)group : 0;
    p.interact_groups.clear();
    for (uint32_t g = 0; g < 32; ++g)
        if (interact_mask & (1u << g)) p.interact_groups.push_back(g);
    return p;
}

DurHitbox make_dur(const double* r, int kind) {
    // r = px, py, dw, dh, vx, vy, rw, rh, duration
    DurHitbox d;
    d.value = PlacedShape{v2(r[0], r[1]), (ShapeKind)kind, v2(r[2], r[3])};
    d.vel = DurHbVel{v2(r[4], r[5]), v2(r[6], r[7]), r[8]};
    return d;
}

template <class F>
int guard(Handle* h, F&& f) {
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        if (h) h->err = e.what();
        g_err = e.what();
        return -1;
    }
}
}  // namespace

extern "C" {

const char* orc_last_error(void* hp) { return hp ? ((Handle*)hp)->err.c_str() : g_err.c_str(); }

void* orc_new(double cell_width, double padding) {
    auto* h = new Handle;
    if (guard(h, [&] { h->c = std::make_unique<Collider>(cell_width, padding); }) != 0) {
        g_err = h->err;
        delete h;
        return nullptr;
    }
    return h;
}
void orc_free(void* hp) { delete (Handle*)hp; }

typedef int (*orc_can_interact_fn)(uint64_t id_a, uint64_t id_b);
void orc_set_can_interact(void* hp, orc_can_interact_fn fn) {
    auto* h = (Handle*)hp;
    if (fn) h->c->set_can_interact([fn](const Profile& a, const Profile& b) { return fn(a.id, b.id) != 0; });
    else h->c->set_can_interact(nullptr);
}

double orc_time(void* hp) { return ((Handle*)hp)->c->time(); }
double orc_next_time(void* hp) { return ((Handle*)hp)->c->next_time(); }
int orc_set_time(void* hp, double t) {
    auto* h = (Handle*)hp;
    return guard(h, [&] { h->c->set_time(t); });
}
// returns 1 and fills out (kind 1 Collide / 2 Separate, ids ascending), 0 for None, -1 on panic
int orc_next(void* hp, uint32_t* kind, uint64_t* id_1, uint64_t* id_2) {
    auto* h = (Handle*)hp;
    int got = 0;
    int rc = guard(h, [&] {
        if (auto ev = h->c->next()) {
            *kind = (uint32_t)ev->kind;
            *id_1 = ev->id_1;
            *id_2 = ev->id_2;
            got = 1;
        }
    });
    return rc != 0 ? -1 : got;
}

// hb = px, py, dw, dh, vx, vy, rw, rh, end_time ; kind 0 circle / 1 rect ; group < 0 => None
int64_t orc_add_hitbox(void* hp, uint64_t id, int64_t group, uint32_t interact_mask, int kind, const double* hb, uint64_t* out, uint64_t cap) {
    auto* h = (Handle*)hp;
    int64_t n = 0;
    int rc = guard(h, [&] {
        Hitbox hitbox;
        hitbox.value = placed(v2(hb[0], hb[1]), (ShapeKind)kind, v2(hb[2], hb[3]));
        hitbox.vel = HbVel{v2(hb[4], hb[5]), v2(hb[6], hb[7]), hb[8]};
        auto res = h->c->add_hitbox(make_profile(id, group, interact_mask), hitbox);
        n = (int64_t)res.size();
        for (size_t i = 0; i < res.size() && i < cap; ++i) out[i] = res[i];
    });
    return rc != 0 ? -1 : n;
}
int orc_set_hitbox_vel(void* hp, uint64_t id, const double* vel) {  // vx, vy, rw, rh, end_time
    auto* h = (Handle*)hp;
    return guard(h, [&] { h->c->set_hitbox_vel(id, HbVel{v2(vel[0], vel[1]), v2(vel[2], vel[3]), vel[4]}); });
}
int64_t orc_remove_hitbox(void* hp, uint64_t id, uint64_t* out, uint64_t cap) {
    auto* h = (Handle*)hp;
    int64_t n = 0;
    int rc = guard(h, [&] {
        auto res = h->c->remove_hitbox(id);
        n = (int64_t)res.size();
        for (size_t i = 0; i < res.size() && i < cap; ++i) out[i] = res[i];
    });
    return rc != 0 ? -1 : n;
}
// out = px, py, dw, dh, vx, vy, rw, rh, end_time ; *kind
int orc_get_hitbox(void* hp, uint64_t id, double* out, int* kind) {
    auto* h = (Handle*)hp;
    return guard(h, [&] {
        Hitbox hb = h->c->get_hitbox(id);
        out[0] = hb.value.pos.x; out[1] = hb.value.pos.y; out[2] = hb.value.dims.x; out[3] = hb.value.dims.y;
        out[4] = hb.vel.value.x; out[5] = hb.vel.value.y; out[6] = hb.vel.resize.x; out[7] = hb.vel.resize.y;
        out[8] = hb.vel.end_time;
        *kind = (int)hb.value.kind;
    });
}
// raw internal state: px,py,dw,dh,vx,vy,rw,rh,start_time,internal end_time,pub_end_time
int orc_get_raw(void* hp, uint64_t id, double* out, int* kind) {
    auto* h = (Handle*)hp;
    return guard(h, [&] {
        const HitboxInfo& i = h->c->info(id);
        out[0] = i.hitbox.value.pos.x; out[1] = i.hitbox.value.pos.y; out[2] = i.hitbox.value.dims.x; out[3] = i.hitbox.value.dims.y;
        out[4] = i.hitbox.vel.value.x; out[5] = i.hitbox.vel.value.y; out[6] = i.hitbox.vel.resize.x; out[7] = i.hitbox.vel.resize.y;
        out[8] = i.start_time; out[9] = i.hitbox.vel.end_time; out[10] = i.pub_end_time;
        *kind = (int)i.hitbox.value.kind;
    });
}
int64_t orc_get_overlaps(void* hp, uint64_t id, uint64_t* out, uint64_t cap) {
    auto* h = (Handle*)hp;
    int64_t n = 0;
    int rc = guard(h, [&] {
        auto res = h->c->get_overlaps(id);
        n = (int64_t)res.size();
        for (size_t i = 0; i < res.size() && i < cap; ++i) out[i] = res[i];
    });
    return rc != 0 ? -1 : n;
}
int orc_is_overlapping(void* hp, uint64_t a, uint64_t b) { return ((Handle*)hp)->c->is_overlapping(a, b) ? 1 : 0; }
int64_t orc_query_overlaps(void* hp, int kind, const double* shape /*px,py,dw,dh*/, uint64_t id, int64_t group, uint32_t mask, uint64_t* out,
                           uint64_t cap) {
    auto* h = (Handle*)hp;
    int64_t n = 0;
    int rc = guard(h, [&] {
        PlacedShape s = placed(v2(shape[0], shape[1]), (ShapeKind)kind, v2(shape[2], shape[3]));
        auto res = h->c->query_overlaps(s, make_profile(id, group, mask));
        n = (int64_t)res.size();
        for (size_t i = 0; i < res.size() && i < cap; ++i) out[i] = res[i];
    });
    return rc != 0 ? -1 : n;
}
uint64_t orc_len(void* hp) { return ((Handle*)hp)->c->len(); }

// Bulk step at the current time.  Arrays are indexed by the rank of the id in ascending-id order; any of
// vx/vy/rw/rh may be null (keep the current component); end_time array may be null => scalar end_time.
int orc_bulk_update(void* hp, const double* vx, const double* vy, const double* rw, const double* rh, const double* end_time_arr,
                    double end_time, int record_candidates) {
    auto* h = (Handle*)hp;
    return guard(h, [&] {
        std::vector<HbId> ids = h->c->ids();
        std::unordered_map<HbId, size_t> rank;
        if (vx || vy || rw || rh || end_time_arr)
            for (size_t i = 0; i < ids.size(); ++i) rank[ids[i]] = i;
        h->c->record_candidates = record_candidates != 0;
        h->c->bulk_update([&](HbId id, HbVel cur) {
            size_t r = rank.empty() ? 0 : rank[id];
            if (vx) cur.value.x = vx[r];
            if (vy) cur.value.y = vy[r];
            if (rw) cur.resize.x = rw[r];
            if (rh) cur.resize.y = rh[r];
            cur.end_time = end_time_arr ? end_time_arr[r] : end_time;
            return cur;
        });
    });
}
// stats: updates, collide_solves, separate_solves, grid_cells ; reset != 0 clears them afterwards
void orc_stats(void* hp, uint64_t* out, int reset) {
    auto* h = (Handle*)hp;
    out[0] = h->c->stats.updates;
    out[1] = h->c->stats.collide_solves;
    out[2] = h->c->stats.separate_solves;
    out[3] = h->c->stats.grid_cells;
    if (reset) h->c->stats = BulkStats{};
}
int64_t orc_bulk_candidates(void* hp, uint64_t* hi, uint64_t* lo, uint64_t cap) {
    auto* h = (Handle*)hp;
    auto& v = h->c->bulk_candidates();
    for (size_t i = 0; i < v.size() && i < cap; ++i) {
        hi[i] = v[i].first;
        lo[i] = v[i].second;
    }
    return (int64_t)v.size();
}
// queue in (time, index) order; kind 0 Reiterate / 1 Collide / 2 Separate
int64_t orc_dump_events(void* hp, double* time, uint64_t* index, uint32_t* kind, uint64_t* a, uint64_t* b, uint64_t cap) {
    auto* h = (Handle*)hp;
    auto v = h->c->dump_events();
    for (size_t i = 0; i < v.size() && i < cap; ++i) {
        time[i] = v[i].time;
        index[i] = v[i].index;
        kind[i] = (uint32_t)v[i].kind;
        a[i] = v[i].a;
        b[i] = v[i].b;
    }
    return (int64_t)v.size();
}
int64_t orc_dump_grid(void* hp, int32_t* x, int32_t* y, uint32_t* group, uint64_t* id, uint64_t cap) {
    auto* h = (Handle*)hp;
    auto v = h->c->dump_grid();
    for (size_t i = 0; i < v.size() && i < cap; ++i) {
        x[i] = v[i].x;
        y[i] = v[i].y;
        group[i] = v[i].group;
        id[i] = v[i].id;
    }
    return (int64_t)v.size();
}

// ---- bulk helpers for large parity scenes (avoid one ctypes call per hitbox) ----
// cols: 9 columns of n doubles, column k at cols + k*n: px,py,dw,dh,vx,vy,rw,rh,end_time.  group 0xFFFFFFFF = None.
int orc_add_hitboxes(void* hp, uint64_t n, const uint64_t* ids, const double* cols, const uint8_t* kind, const uint32_t* group,
                     const uint32_t* mask) {
    auto* h = (Handle*)hp;
    return guard(h, [&] {
        for (uint64_t i = 0; i < n; ++i) {
            Hitbox hitbox;
            hitbox.value = placed(v2(cols[0 * n + i], cols[1 * n + i]), (ShapeKind)kind[i], v2(cols[2 * n + i], cols[3 * n + i]));
            hitbox.vel = HbVel{v2(cols[4 * n + i], cols[5 * n + i]), v2(cols[6 * n + i], cols[7 * n + i]), cols[8 * n + i]};
            int64_t g = group[i] == 0xFFFFFFFFu ? -1 : (int64_t)group[i];
            h->c->add_hitbox(make_profile(ids[i], g, mask[i]), hitbox);
        }
    });
}
// ids ascending; cols: 11 columns of cap doubles (px,py,dw,dh,vx,vy,rw,rh,start_time,internal end_time,pub_end_time)
int64_t orc_dump_state(void* hp, uint64_t* ids, double* cols, uint8_t* kind, uint32_t* group, uint32_t* mask, uint64_t cap) {
    auto* h = (Handle*)hp;
    std::vector<HbId> all = h->c->ids();
    for (size_t r = 0; r < all.size() && r < cap; ++r) {
        const HitboxInfo& i = h->c->info(all[r]);
        ids[r] = all[r];
        const double v[11] = {i.hitbox.value.pos.x, i.hitbox.value.pos.y, i.hitbox.value.dims.x, i.hitbox.value.dims.y,
                              i.hitbox.vel.value.x, i.hitbox.vel.value.y, i.hitbox.vel.resize.x, i.hitbox.vel.resize.y,
                              i.start_time, i.hitbox.vel.end_time, i.pub_end_time};
        for (int k = 0; k < 11; ++k) cols[(size_t)k * cap + r] = v[k];
        kind[r] = (uint8_t)i.hitbox.value.kind;
        group[r] = i.profile.has_group ? i.profile.group : 0xFFFFFFFFu;
        uint32_t m = 0;
        for (HbGroup g : i.profile.interact_groups) m |= 1u << g;
        mask[r] = m;
    }
    return (int64_t)all.size();
}
// every overlapping pair once, as (a < b)
int64_t orc_dump_overlaps(void* hp, uint64_t* a, uint64_t* b, uint64_t cap) {
    auto* h = (Handle*)hp;
    int64_t n = 0;
    for (HbId id : h->c->ids())
        for (HbId other : h->c->info(id).overlaps.v)
            if (id < other) {
                if ((uint64_t)n < cap) {
                    a[n] = id;
                    b[n] = other;
                }
                ++n;
            }
    return n;
}

// ---- scalar / batched primitives for KATs and narrowphase parity ----
// a, b: n x 9 doubles (px,py,dw,dh,vx,vy,rw,rh,duration); kinds: n bytes each.  NaN-free inputs assumed;
// a panic inside one pair writes NaN for that pair.
void orc_collide_time_batch(uint64_t n, const double* a, const uint8_t* ka, const double* b, const uint8_t* kb, double* out) {
    for (uint64_t i = 0; i < n; ++i) {
        try {
            out[i] = collide_time(make_dur(a + 9 * i, ka[i]), make_dur(b + 9 * i, kb[i]));
        } catch (const std::exception&) {
            out[i] = std::numeric_limits<double>::quiet_NaN();
        }
    }
}
void orc_separate_time_batch(uint64_t n, const double* a, const uint8_t* ka, const double* b, const uint8_t* kb, double padding, double* out) {
    for (uint64_t i = 0; i < n; ++i) {
        try {
            out[i] = separate_time(make_dur(a + 9 * i, ka[i]), make_dur(b + 9 * i, kb[i]), padding);
        } catch (const std::exception&) {
            out[i] = std::numeric_limits<double>::quiet_NaN();
        }
    }
}
// returns 1 if Some(root)
// Benchmark driver (bench.py cpu_baseline rows for configs 1 and 2): the user loop of README.md:66-80 run natively — advance
// to the next event, pop it, and on Collide halve both hitboxes' velocities (get_hitbox + set_hitbox_vel, end_time kept).
// out = {user-visible events, Collide events among them}.  Returns -1 on a panic.
int orc_halving_loop(void* hp, double t_end, uint64_t* out) {
    auto* h = (Handle*)hp;
    uint64_t n_ev = 0, n_col = 0;
    int rc = guard(h, [&] {
        Collider& c = *h->c;
        while (c.time() < t_end) {
            const double nt = c.next_time();
            c.set_time(nt < t_end ? nt : t_end);
            if (auto ev = c.next()) {
                ++n_ev;
                if ((int)ev->kind == 1) {
                    ++n_col;
                    for (HbId id : {ev->id_1, ev->id_2}) {
                        Hitbox hb = c.get_hitbox(id);
                        hb.vel.value.x *= 0.5;
                        hb.vel.value.y *= 0.5;
                        c.set_hitbox_vel(id, hb.vel);
                    }
                }
            }
        }
    });
    out[0] = n_ev;
    out[1] = n_col;
    return rc;
}

int orc_quad_root_ascending(double a, double b, double c, double* root) { return quad_root_ascending(a, b, c, root) ? 1 : 0; }
// IndexRect iteration: fills xs/ys (cap entries), returns count or -1 on panic
int64_t orc_index_rect_iter(int32_t sx, int32_t sy, int32_t ex, int32_t ey, int32_t* xs, int32_t* ys, uint64_t cap) {
    int64_t n = 0;
    int rc = guard(nullptr, [&] {
        IndexRect r = IndexRect::make(sx, sy, ex, ey);
        r.for_each([&](int32_t x, int32_t y) {
            if ((uint64_t)n < cap) {
                xs[n] = x;
                ys[n] = y;
            }
            ++n;
        });
    });
    return rc != 0 ? -1 : n;
}
int orc_index_rect_contains(int32_t sx, int32_t sy, int32_t ex, int32_t ey, int32_t x, int32_t y) {
    return IndexRect{sx, sy, ex, ey}.contains(x, y) ? 1 : 0;
}
// Grid::index_bounds of a rect-shaped bounds (px,py,dw,dh) -> sx,sy,ex,ey
int orc_index_bounds(double cell_width, const double* s, int32_t* out) {
    return guard(nullptr, [&] {
        Grid g(cell_width);
        IndexRect r = g.index_bounds(PlacedShape{v2(s[0], s[1]), ShapeKind::Rect, v2(s[2], s[3])});
        out[0] = r.sx; out[1] = r.sy; out[2] = r.ex; out[3] = r.ey;
    });
}
// PlacedShape::advance; in/out = px,py,dw,dh
int orc_advance(int kind, const double* s, double vx, double vy, double rw, double rh, double elapsed, double* out) {
    return guard(nullptr, [&] {
        PlacedShape r = advance(placed(v2(s[0], s[1]), (ShapeKind)kind, v2(s[2], s[3])), v2(vx, vy), v2(rw, rh), elapsed);
        out[0] = r.pos.x; out[1] = r.pos.y; out[2] = r.dims.x; out[3] = r.dims.y;
    });
}
// min_x, min_y, max_x, max_y
void orc_edges(const double* s, double* out) {
    PlacedShape p{v2(s[0], s[1]), ShapeKind::Rect, v2(s[2], s[3])};
    out[0] = p.min_x(); out[1] = p.min_y(); out[2] = p.max_x(); out[3] = p.max_y();
}
int orc_shape_overlaps(int ka, const double* a, int kb, const double* b) {
    return overlaps(PlacedShape{v2(a[0], a[1]), (ShapeKind)ka, v2(a[2], a[3])}, PlacedShape{v2(b[0], b[1]), (ShapeKind)kb, v2(b[2], b[3])}) ? 1 : 0;
}
// swept bounding box + index rect of a DurHitbox (r: 9 doubles) -> bbox px,py,dw,dh and rect
int orc_swept_rect(double cell_width, const double* r, int kind, double* bbox, int32_t* rect) {
    return guard(nullptr, [&] {
        DurHitbox d = make_dur(r, kind);
        PlacedShape bb = d.bounding_box();
        bbox[0] = bb.pos.x; bbox[1] = bb.pos.y; bbox[2] = bb.dims.x; bbox[3] = bb.dims.y;
        Grid g(cell_width);
        IndexRect ir = g.index_bounds(bb);
        rect[0] = ir.sx; rect[1] = ir.sy; rect[2] = ir.ex; rect[3] = ir.ey;
    });
}

}  // extern "C"
