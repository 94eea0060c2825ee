// Host side of the reference's `Collider` API over the device engine (included by engine.cu — same translation unit).
//
// What lives where (SURVEY.md 8a-16 … a-19, 8f-1 … f-3):
//   device  every geometric computation: rebase, solitaire check, swept bounds, IndexRect, grid membership and
//           cell-mate enumeration, collide_time / separate_time, get_hitbox's advance, query_overlaps' shape test
//   host    bookkeeping with no arithmetic in it: id <-> label maps, the overlap sets (HitboxInfo.overlaps,
//           collider.rs:463), the event queue order (EventManager, events.rs:102-194) and the clock
//
// Event queue = the device-sorted run of the last bulk step (read lazily, in chunks) + an ordered overlay of the events
// created since.  An event of the run is dead once one of its hitboxes had clear_related_events (events.rs:141-154)
// run on it after the bulk step — tracked per label.  Overlay events carry the reference's insertion index
// (events.rs:156-169), so simultaneous events come out in the canonical order of SURVEY.md 8c.

namespace {

struct EvKey {
    double time;
    uint64_t index;
};
struct EvKeyLess {
    bool operator()(const EvKey& a, const EvKey& b) const { return a.time < b.time || (a.time == b.time && a.index < b.index); }
};
struct QEvent {
    uint32_t kind;  // CB200_EV_*
    uint64_t a, b;
};
constexpr uint64_t kPairBase = 0x8000000000000000ull;  // events.rs:26

template <class T>
struct Mapped {  // page-locked host memory mapped into the device address space: small per-call arguments and results
    T* h = nullptr;
    T* d = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        release();
        size_t ncap = std::max<size_t>(n + n / 2, 64);
        cudaError_t e = cudaHostAlloc((void**)&h, ncap * sizeof(T), cudaHostAllocMapped);
        if (e != cudaSuccess) return e;
        e = cudaHostGetDevicePointer((void**)&d, h, 0);
        if (e == cudaSuccess) cap = ncap;
        return e;
    }
    void release() {
        if (h) cudaFreeHost(h);
        h = d = nullptr;
        cap = 0;
    }
};

}  // namespace

struct cb200_collider {
    cb200_ctx* ctx = nullptr;
    double time = 0.0;
    std::string err;

    struct Info {
        uint32_t label, group, mask;
        std::vector<uint64_t> overlaps;  // ascending ids
        std::vector<EvKey> keys;         // overlay events this hitbox is part of
        uint64_t version = 0;            // changes whenever the hitbox's device state changes (solve cache validity)
    };
    // Prefetched narrowphase results for queued Collide / Separate events (uid -> result): see prefetch_solves.
    struct CachedSolve {
        double time;
        uint32_t err;
        uint64_t ver_a, ver_b;
    };
    std::unordered_map<uint64_t, CachedSolve> solve_cache;
    // get_hitbox answers that came back with the prefetched solves: (id, bits of the event time) -> hitbox at that time.
    // Served only while the collider stands at that time and the hitbox still carries the version seen by the prefetch.
    struct CachedHitbox {
        uint64_t version;
        double hb[9];
        uint32_t kind, err;
    };
    struct HbKey {
        uint64_t id, tbits;
        bool operator==(const HbKey& o) const { return id == o.id && tbits == o.tbits; }
    };
    struct HbKeyHash {
        size_t operator()(const HbKey& k) const { return (size_t)((k.id * 0x9E3779B97F4A7C15ull) ^ (k.tbits * 0xC2B2AE3D27D4EB4Full)); }
    };
    std::unordered_map<HbKey, CachedHitbox, HbKeyHash> hitbox_cache;
    uint64_t hitbox_cache_hits = 0;
    uint64_t version_counter = 0;
    Mapped<PairJob> jobs;
    Mapped<PairOut> job_out;
    uint64_t prefetches = 0, prefetched = 0;
    std::unordered_map<uint64_t, Info> infos;
    std::map<uint64_t, uint32_t> by_id;  // ordered id -> label: labels are order-isomorphic to the ids
    std::vector<uint64_t> id_of_label;   // UINT64_MAX = free
    bool ids_dirty = true;               // the device label -> id table is stale
    bool overlaps_dirty = true;          // the device overlap set is stale
    std::set<std::pair<uint64_t, uint64_t>> overlap_pairs;  // every overlapping pair once, (lower id, higher id)
    std::vector<uint64_t> with_keys;     // ids whose Info.keys may be non-empty (cleared wholesale by a bulk step)

    std::map<EvKey, QEvent, EvKeyLess> overlay;
    uint64_t next_index = 0;
    std::vector<cb200_event> base;  // window [base_off, base_off + base.size()) of the device run
    uint64_t base_off = 0, base_cur = 0, base_total = 0;
    bool base_on_host = false;      // the whole run was fetched (can_interact filtering)
    std::vector<uint8_t> cleared;   // by label: events of the run involving it are dead

    Mapped<OneHeader> head;
    Mapped<OneItem> items;
    Mapped<uint32_t> ov, qout;
    Mapped<QueryItem> bq_items;   // batched queries
    Mapped<OneHeader> bq_heads;
    Mapped<uint32_t> bq_out;
    std::unordered_map<uint32_t, uint32_t> group_index;  // HbGroup value -> device group index (0..31), in order of first use
    DevBuf<uint8_t> dirty;
    DevBuf<uint32_t> dirty_list;
    uint32_t* d_dirty_count = nullptr;
    uint32_t n_dirty = 0;
    bool have_grid = false;
    uint32_t rebin_threshold = 4096;
    uint64_t rebins = 0;

    cb200_can_interact_fn can_interact = nullptr;
    void* can_interact_user = nullptr;
};

namespace {

int cfail(cb200_collider* c, int code, const std::string& msg) {
    c->err = msg;
    return code;
}
#define CCU(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return cfail(c, CB200_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

int not_found(cb200_collider* c, uint64_t id) { return cfail(c, CB200_E_CONTRACT, "hitbox id " + std::to_string(id) + " not found"); }

bool interacts(cb200_collider* c, uint64_t a, uint64_t b) { return !c->can_interact || c->can_interact(c->can_interact_user, a, b) != 0; }

TableView table_of(cb200_collider* c) {
    cb200_ctx* x = c->ctx;
    TableView t;
    t.s = cols_of(x);
    t.label = x->label.p;
    t.row_of = x->row_of.p;
    t.rec = x->rec.p;
    t.bink = x->bink.p;
    t.dirty = c->dirty.p;
    t.dirty_list = c->dirty_list.p;
    t.dirty_count = c->d_dirty_count;
    t.n = x->n;
    return t;
}
GridView grid_of(cb200_collider* c) {
    cb200_ctx* x = c->ctx;
    return GridView{x->entries.p, x->slots.p, x->chunk_base.p, x->geom, c->have_grid ? 1 : 0};
}

// Room for `rows` table rows and `labels` labels; existing contents are kept.
int ensure_capacity(cb200_collider* c, size_t rows, size_t labels) {
    cb200_ctx* x = c->ctx;
    cudaStream_t st = x->stream;
    const size_t old_rows = c->dirty.cap, old_labels = x->row_of.cap;
    for (auto& b : x->col) CCU(b.reserve(rows, true, st));
    CCU(x->kind.reserve(rows, true, st)); CCU(x->reit.reserve(rows, true, st)); CCU(x->group.reserve(rows, true, st));
    CCU(x->mask.reserve(rows, true, st)); CCU(x->label.reserve(rows, true, st)); CCU(x->rec.reserve(rows, true, st)); CCU(x->bink.reserve(rows, true, st));
    CCU(c->dirty.reserve(rows, true, st)); CCU(c->dirty_list.reserve(rows, true, st));
    if (c->dirty.cap > old_rows) CCU(cudaMemsetAsync(c->dirty.p + old_rows, 0, c->dirty.cap - old_rows, st));
    CCU(x->row_of.reserve(labels, true, st));
    if (x->row_of.cap > old_labels) CCU(cudaMemsetAsync(x->row_of.p + old_labels, 0xff, (x->row_of.cap - old_labels) * 4, st));
    if (c->id_of_label.size() < labels) c->id_of_label.resize(std::max(labels, c->id_of_label.size() * 2), UINT64_MAX);
    if (c->cleared.size() < c->id_of_label.size()) c->cleared.resize(c->id_of_label.size(), 1);
    return CB200_OK;
}

// Grid period for the incremental API: ~4 slots per hitbox, square, at least 256 x 256 cells.
void collider_geom(cb200_collider* c) {
    cb200_ctx* x = c->ctx;
    if (x->geom_forced) return;
    const int bits = std::min(26, std::max(16, ilog2_ceil(std::max<uint64_t>(x->n, 1) * 4)));
    x->geom = GridGeom{(bits + 1) / 2, bits / 2};
}

int wait_head(cb200_collider* c) {
    CCU(cudaStreamSynchronize(c->ctx->stream));
    CCU(cudaGetLastError());
    return CB200_OK;
}

// A bulk step on the tile path leaves no global grid (slot table, entry array, per-row records) behind; the single-hitbox
// operations read it, so the first of them after such a step has the engine rebuild it from the rebased table.
int sync_grid(cb200_collider* c) {
    cb200_ctx* x = c->ctx;
    if (!x->have_step || !x->last_step_tile || x->global_grid_valid) return CB200_OK;
    int rc = rebuild_global_grid(x);
    if (rc != CB200_OK) return cfail(c, rc, x->err);
    return CB200_OK;
}

// Rebuild the slot-sorted entry array from the current records (no rebase, no solve); empties the dirty list.
int rebin(cb200_collider* c) {
    cb200_ctx* x = c->ctx;
    {
        int rc = sync_grid(c);
        if (rc != CB200_OK) return rc;
    }
    const uint32_t n = x->n;
    collider_geom(c);
    const uint32_t n_slots = x->geom.n_slots();
    CCU(x->slots.reserve(n_slots));
    CCU(x->chunk_base.reserve(n_slots / kScanChunk));
    if (x->entries.cap == 0) CCU(x->entries.reserve((size_t)n * 4 + 1024));
    for (int attempt = 0; attempt < 3; ++attempt) {
        k_step_zero<<<x->sm_count * 4, kThreads, 0, x->stream>>>(reinterpret_cast<uint4*>(x->slots.p), n_slots / 4, x->d_ctr, nullptr, nullptr, 0, 0, 0);
        if (n) k_rebin_count<<<(n + kThreads - 1) / kThreads, kThreads, 0, x->stream>>>(n, x->rec.p, x->slots.p, x->geom, c->dirty.p, c->d_dirty_count);
        int rc = launch_scan(x, true);
        if (rc != CB200_OK) return cfail(c, rc, x->err);
        const VisSpace vs{n, 0, 1, 0};
        const uint32_t cap_e = (uint32_t)std::min<size_t>(x->entries.cap, 0xffffffffu);
        const uint32_t grid_v = std::max<uint32_t>(1, std::min<uint32_t>((n + kThreads - 1) / kThreads, x->sm_count * 8));
        if (n) k_scatter<false><<<grid_v, kThreads, 0, x->stream>>>(n, x->rec.p, x->bink.p, vs, x->slots.p, x->entries.p, cap_e, x->geom, INT32_MIN, INT32_MAX, x->d_ctr);
        x->launches += 3;
        Counters h;
        CCU(cudaMemcpyAsync(&h, x->d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, x->stream));
        int rcw = wait_head(c);
        if (rcw != CB200_OK) return rcw;
        if (h.overflow_entries || h.n_entries > x->entries.cap) {
            CCU(x->entries.reserve((size_t)h.n_entries + h.n_entries / 4 + 1024));
                    continue;
        }
        c->have_grid = true;
        c->n_dirty = 0;
        c->rebins += 1;
        return CB200_OK;
    }
    return cfail(c, CB200_E_STATE, "re-bin buffers failed to converge");
}

int maybe_rebin(cb200_collider* c) {
    int rc = sync_grid(c);
    if (rc != CB200_OK) return rc;
    if (c->n_dirty > c->rebin_threshold) return rebin(c);
    return CB200_OK;
}

// ---- event queue ------------------------------------------------------------------------------------------
bool base_dead(cb200_collider* c, const cb200_event& e) {
    auto ia = c->infos.find(e.a);
    if (ia == c->infos.end() || c->cleared[ia->second.label]) return true;
    if (e.kind != CB200_EV_REITERATE) {
        auto ib = c->infos.find(e.b);
        if (ib == c->infos.end() || c->cleared[ib->second.label]) return true;
    }
    return false;
}

// Positions base_cur on the first live event of the run (fetching windows as needed); *ev = it, or null at the end.
int base_peek(cb200_collider* c, const cb200_event** ev) {
    *ev = nullptr;
    for (;;) {
        if (c->base_cur >= c->base_total) return CB200_OK;
        if (c->base_cur < c->base_off || c->base_cur >= c->base_off + c->base.size()) {
            if (c->base_on_host) return CB200_OK;  // the whole run is in `base`
            const uint64_t want = std::min<uint64_t>(4096, c->base_total - c->base_cur);
            c->base.resize(want);
            uint64_t got = 0;
            int rc = cb200_fetch_events(c->ctx, c->base.data(), want, c->base_cur, &got);
            if (rc != CB200_OK) return cfail(c, rc, c->ctx->err);
            c->base.resize(got);
            c->base_off = c->base_cur;
            if (got == 0) return CB200_OK;
        }
        const cb200_event& e = c->base[c->base_cur - c->base_off];
        if (!base_dead(c, e)) {
            *ev = &e;
            return CB200_OK;
        }
        c->base_cur += 1;
    }
}

struct Head {
    bool any = false, from_base = false;
    double time = INFINITY;
    QEvent ev{};
    EvKey key{};
};
int queue_head(cb200_collider* c, Head* out) {
    const cb200_event* be = nullptr;
    int rc = base_peek(c, &be);
    if (rc != CB200_OK) return rc;
    Head h;
    if (be) {
        h.any = true;
        h.from_base = true;
        h.time = be->time;
        h.ev = QEvent{be->kind, be->a, be->b};
    }
    if (!c->overlay.empty()) {
        const auto& o = *c->overlay.begin();
        bool take = !h.any || o.first.time < h.time;
        if (!take && o.first.time == h.time) {
            // equal times: solitaire before pair (index < PAIR_BASE); within a class the run's events are older
            const bool o_sol = o.first.index < kPairBase, b_sol = h.ev.kind == CB200_EV_REITERATE;
            take = o_sol && !b_sol;
        }
        if (take) {
            h.any = true;
            h.from_base = false;
            h.time = o.first.time;
            h.ev = o.second;
            h.key = o.first;
        }
    }
    *out = h;
    return CB200_OK;
}

void drop_key(std::vector<EvKey>& keys, const EvKey& k) {
    for (size_t i = 0; i < keys.size(); ++i)
        if (keys[i].index == k.index) {
            keys[i] = keys.back();
            keys.pop_back();
            return;
        }
}

// EventManager::new_event_key + add_*_event (events.rs:115-169)
int add_event(cb200_collider* c, double time, const QEvent& ev) {
    if (time != time) return cfail(c, CB200_E_CONTRACT, "unexpected NaN");  // float.rs:31
    if (time >= 1e50) return CB200_OK;
    uint64_t index = c->next_index++;
    if (ev.kind != CB200_EV_REITERATE) index += kPairBase;
    const EvKey key{time, index};
    c->overlay.emplace(key, ev);
    c->infos[ev.a].keys.push_back(key);
    if (c->with_keys.size() > (1u << 22)) {  // (no bulk step for a long time: keep the list a set)
        std::sort(c->with_keys.begin(), c->with_keys.end());
        c->with_keys.erase(std::unique(c->with_keys.begin(), c->with_keys.end()), c->with_keys.end());
    }
    c->with_keys.push_back(ev.a);
    if (ev.kind != CB200_EV_REITERATE) {
        c->infos[ev.b].keys.push_back(key);
        c->with_keys.push_back(ev.b);
    }
    return CB200_OK;
}

// EventManager::clear_related_events (events.rs:141-154)
void clear_related(cb200_collider* c, uint64_t id, cb200_collider::Info& inf) {
    for (const EvKey& k : inf.keys) {
        auto it = c->overlay.find(k);
        if (it == c->overlay.end()) continue;
        const QEvent ev = it->second;
        c->overlay.erase(it);
        if (ev.kind != CB200_EV_REITERATE) {
            const uint64_t other = ev.a == id ? ev.b : ev.a;
            auto io = c->infos.find(other);
            if (io != c->infos.end()) drop_key(io->second.keys, k);
        }
    }
    inf.keys.clear();
    c->cleared[inf.label] = 1;
}

void overlap_insert(std::vector<uint64_t>& v, uint64_t id) { v.insert(std::lower_bound(v.begin(), v.end(), id), id); }
void pair_insert(cb200_collider* c, uint64_t a, uint64_t b) {
    overlap_insert(c->infos[a].overlaps, b);
    overlap_insert(c->infos[b].overlaps, a);
    c->overlap_pairs.emplace(std::min(a, b), std::max(a, b));
    c->overlaps_dirty = true;
}
bool overlap_remove(std::vector<uint64_t>& v, uint64_t id) {
    auto it = std::lower_bound(v.begin(), v.end(), id);
    if (it == v.end() || *it != id) return false;
    v.erase(it);
    return true;
}

// ---- labels -----------------------------------------------------------------------------------------------
// Dense relabel: label = rank of the id.  Device: label column, records and row_of follow; the sorted entry array
// carries the old labels, so the grid is rebuilt.
int relabel(cb200_collider* c) {
    cb200_ctx* x = c->ctx;
    {
        int rc = sync_grid(c);
        if (rc != CB200_OK) return rc;
    }
    std::vector<uint32_t> new_of_old(c->id_of_label.size(), kRowNone);
    std::vector<uint64_t> ids;
    ids.reserve(c->by_id.size());
    uint32_t r = 0;
    std::vector<uint8_t> cleared(c->cleared.size(), 1);
    for (auto& kv : c->by_id) {
        new_of_old[kv.second] = r;
        cleared[r] = c->cleared[kv.second];
        kv.second = r;
        c->infos[kv.first].label = r;
        ids.push_back(kv.first);
        ++r;
    }
    c->cleared.swap(cleared);
    std::fill(c->id_of_label.begin(), c->id_of_label.end(), UINT64_MAX);
    std::copy(ids.begin(), ids.end(), c->id_of_label.begin());
    c->ids_dirty = true;
    c->overlaps_dirty = true;
    if (x->n) {
        DevBuf<uint32_t> map;
        CCU(map.reserve(new_of_old.size()));
        CCU(cudaMemcpyAsync(map.p, new_of_old.data(), new_of_old.size() * 4, cudaMemcpyHostToDevice, x->stream));
        CCU(cudaMemsetAsync(x->row_of.p, 0xff, x->row_of.cap * 4, x->stream));
        k_relabel<<<(x->n + kThreads - 1) / kThreads, kThreads, 0, x->stream>>>(x->n, map.p, x->label.p, x->rec.p, x->row_of.p);
        x->launches += 1;
        int rc = wait_head(c);
        map.release();
        if (rc != CB200_OK) return rc;
        return rebin(c);
    }
    return CB200_OK;
}

// A label for a new id that keeps labels order-isomorphic to ids; relabels everything when no gap is left.
int assign_label(cb200_collider* c, uint64_t id, uint32_t* out) {
    auto next = c->by_id.lower_bound(id);
    const int64_t hi = next == c->by_id.end() ? (int64_t)0x7ffffff0 : (int64_t)next->second;
    const int64_t lo = next == c->by_id.begin() ? -1 : (int64_t)std::prev(next)->second;
    int64_t label;
    if (next == c->by_id.end()) {
        label = lo + 1;  // appending ids (the common case) keeps labels dense
    } else if (hi - lo >= 2) {
        label = lo + (hi - lo) / 2;  // a hole left by a removal
    } else {
        // no room between the neighbours: park the id on a fresh label past every label in use, then rank everything
        const uint32_t spare = (uint32_t)c->id_of_label.size();
        c->id_of_label.resize(spare + 1, UINT64_MAX);
        c->cleared.resize(c->id_of_label.size(), 1);
        c->by_id[id] = spare;
        c->infos[id].label = spare;
        int rc = ensure_capacity(c, (size_t)c->ctx->n + 1, c->id_of_label.size());
        if (rc != CB200_OK) return rc;
        if ((rc = relabel(c)) != CB200_OK) return rc;
        *out = c->by_id[id];
        return CB200_OK;
    }
    if (label >= 0x7ffffff0) return cfail(c, CB200_E_ARG, "too many hitboxes (label space exhausted)");
    c->by_id[id] = (uint32_t)label;
    *out = (uint32_t)label;
    return CB200_OK;
}

int contract_error(cb200_collider* c, uint32_t flags, uint64_t id) {
    return cfail(c, CB200_E_CONTRACT, flags_message(flags) + " (hitbox id " + std::to_string(id) + ")");
}

// Runs k_update_one for `id`, re-running the enumeration if the item buffer was too small.
int run_update(cb200_collider* c, OneArgs& a) {
    cb200_ctx* x = c->ctx;
    for (int attempt = 0; attempt < 4; ++attempt) {
        a.t = table_of(c);
        a.g = grid_of(c);
        a.head = c->head.d;
        a.items = c->items.d;
        a.cap_items = (uint32_t)c->items.cap;
        k_update_one<<<1, kThreads, 0, x->stream>>>(a);
        x->launches += 1;
        int rc = wait_head(c);
        if (rc != CB200_OK) return rc;
        const OneHeader& h = *c->head.h;
        if (h.err_flags || h.noop || h.n_items <= c->items.cap) return CB200_OK;
        CCU(c->items.reserve(h.n_items));
        if (a.op == kOpAdd) x->n += 1;  // the row exists now; only the enumeration is repeated
        a.op = kOpEnumerate;
    }
    return cfail(c, CB200_E_STATE, "item buffer failed to converge");
}

// Collider::internal_update_hitbox (collider.rs:231-250) for one hitbox; vel == null: Reiterate.
int internal_update(cb200_collider* c, uint64_t id, const double* vel5, bool check_noop) {
    auto it = c->infos.find(id);
    if (it == c->infos.end()) return not_found(c, id);
    cb200_collider::Info& inf = it->second;
    int rc = maybe_rebin(c);
    if (rc != CB200_OK) return rc;
    const size_t n_ov = inf.group == CB200_GROUP_NONE ? 0 : inf.overlaps.size();
    CCU(c->ov.reserve(std::max<size_t>(n_ov, 1)));
    for (size_t k = 0; k < n_ov; ++k) c->ov.h[k] = c->infos[inf.overlaps[k]].label;
    CCU(c->items.reserve(n_ov + 64));
    OneArgs a{};
    a.op = kOpUpdate;
    a.label = inf.label;
    a.has_vel = vel5 ? 1 : 0;
    a.check_noop = check_noop ? 1 : 0;
    if (vel5) {
        a.vx = vel5[0]; a.vy = vel5[1]; a.rw = vel5[2]; a.rh = vel5[3]; a.end_time = vel5[4];
    }
    a.T = c->time;
    a.cell_width = c->ctx->cell_width;
    a.padding = c->ctx->padding;
    a.ov_labels = c->ov.d;
    a.n_ov = (uint32_t)n_ov;
    rc = run_update(c, a);
    if (rc != CB200_OK) return rc;
    const OneHeader h = *c->head.h;
    if (h.err_flags) return contract_error(c, h.err_flags, id);
    if (h.noop) return CB200_OK;
    c->n_dirty = h.n_dirty;
    inf.version = ++c->version_counter;
    clear_related(c, id, inf);
    // solitaire_event_check (collider.rs:441-447), then Separate events in ascending partner id, then Collide events
    if (h.reit && (rc = add_event(c, h.end_time, QEvent{CB200_EV_REITERATE, id, 0})) != CB200_OK) return rc;
    std::vector<std::pair<uint64_t, OneItem>> sep, col;
    for (uint32_t k = 0; k < h.n_items; ++k) {
        const OneItem& itv = c->items.h[k];
        const uint64_t other = c->id_of_label[itv.label];
        (itv.kind == kItemSeparate ? sep : col).emplace_back(other, itv);
    }
    auto by_id = [](const std::pair<uint64_t, OneItem>& p, const std::pair<uint64_t, OneItem>& q) { return p.first < q.first; };
    std::sort(sep.begin(), sep.end(), by_id);
    std::sort(col.begin(), col.end(), by_id);
    for (auto& p : sep)
        if ((rc = add_event(c, p.second.time, QEvent{CB200_EV_SEPARATE, id, p.first})) != CB200_OK) return rc;
    for (auto& p : col)
        if (interacts(c, id, p.first))
            if ((rc = add_event(c, p.second.time, QEvent{CB200_EV_COLLIDE, id, p.first})) != CB200_OK) return rc;
    return CB200_OK;
}

// uid of a queued event for the solve cache: overlay events by their insertion index (pair events carry bit 63), events of
// the device-sorted run by their position (bit 62).
inline uint64_t uid_of(bool from_base, uint64_t base_pos, const EvKey& key) { return from_base ? (0x4000000000000000ull | base_pos) : key.index; }

// One launch for the solves of the next queued events (the one being processed first): what process_event would compute
// for each of them at its own event time — separate_time after a Collide (collider.rs:177-197), collide_time after a
// Separate (collider.rs:148-158).  Speculative: a result is used only if both hitboxes still carry the versions seen here.
constexpr size_t kPrefetchMax = 2048;
int prefetch_solves(cb200_collider* c, uint64_t need_uid, const QEvent& need_ev, double need_time) {
    cb200_ctx* x = c->ctx;
    CCU(c->jobs.reserve(kPrefetchMax));
    CCU(c->job_out.reserve(kPrefetchMax));
    std::vector<uint64_t> uids;
    std::vector<std::pair<uint64_t, uint64_t>> ab, ids;
    uids.reserve(kPrefetchMax);
    auto push = [&](uint64_t uid, const QEvent& ev, double t) {
        auto ia = c->infos.find(ev.a), ib = c->infos.find(ev.b);
        if (ia == c->infos.end() || ib == c->infos.end()) return;
        c->jobs.h[uids.size()] = PairJob{ia->second.label, ib->second.label, ev.kind == CB200_EV_COLLIDE ? 1u : 0u, 0u, t};
        uids.push_back(uid);
        ab.emplace_back(ia->second.version, ib->second.version);
        ids.emplace_back(ev.a, ev.b);
    };
    auto cached_valid = [&](uint64_t uid, const QEvent& ev) {
        auto it = c->solve_cache.find(uid);
        if (it == c->solve_cache.end()) return false;
        auto ia = c->infos.find(ev.a), ib = c->infos.find(ev.b);
        return ia != c->infos.end() && ib != c->infos.end() && it->second.ver_a == ia->second.version && it->second.ver_b == ib->second.version;
    };
    push(need_uid, need_ev, need_time);
    // upcoming events of the run (the fetched window) and of the overlay, half the budget each
    for (uint64_t p = c->base_cur; p < c->base_off + c->base.size() && p < c->base_total && uids.size() < kPrefetchMax / 2; ++p) {
        if (p < c->base_off) continue;
        const cb200_event& e = c->base[p - c->base_off];
        if (e.kind == CB200_EV_REITERATE || base_dead(c, e)) continue;
        const uint64_t uid = uid_of(true, p, EvKey{});
        const QEvent ev{e.kind, e.a, e.b};
        if (uid == need_uid || cached_valid(uid, ev)) continue;
        push(uid, ev, e.time);
    }
    for (auto it = c->overlay.begin(); it != c->overlay.end() && uids.size() < kPrefetchMax; ++it) {
        if (it->second.kind == CB200_EV_REITERATE) continue;
        const uint64_t uid = uid_of(false, 0, it->first);
        if (uid == need_uid || cached_valid(uid, it->second)) continue;
        push(uid, it->second, it->first.time);
    }
    const uint32_t n = (uint32_t)uids.size();
    k_pair_solve_batch<<<(n + 127) / 128, 128, 0, x->stream>>>(table_of(c), c->jobs.d, n, x->padding, c->job_out.d);
    x->launches += 1;
    int rc = wait_head(c);
    if (rc != CB200_OK) return rc;
    if (c->solve_cache.size() > 8 * kPrefetchMax) c->solve_cache.clear();  // (stale entries of events that were cancelled)
    if (c->hitbox_cache.size() > 16 * kPrefetchMax) c->hitbox_cache.clear();
    for (uint32_t k = 0; k < n; ++k) {
        const PairOut& o = c->job_out.h[k];
        c->solve_cache[uids[k]] = cb200_collider::CachedSolve{o.time, o.err, ab[k].first, ab[k].second};
        if (o.err_a == kFIdNotFound) continue;
        uint64_t tbits;
        const double t = c->jobs.h[k].t;
        memcpy(&tbits, &t, 8);
        cb200_collider::CachedHitbox ha{ab[k].first, {}, o.kind_a, o.err_a}, hb{ab[k].second, {}, o.kind_b, o.err_b};
        memcpy(ha.hb, o.hb_a, sizeof(ha.hb));
        memcpy(hb.hb, o.hb_b, sizeof(hb.hb));
        c->hitbox_cache[cb200_collider::HbKey{ids[k].first, tbits}] = ha;
        c->hitbox_cache[cb200_collider::HbKey{ids[k].second, tbits}] = hb;
    }
    c->prefetches += 1;
    c->prefetched += n;
    return CB200_OK;
}

// The narrowphase result process_event needs for the event `uid` = (ev, at the current time).
int pair_solve(cb200_collider* c, uint64_t uid, const QEvent& ev, double* time_out) {
    for (int attempt = 0; attempt < 2; ++attempt) {
        auto it = c->solve_cache.find(uid);
        if (it != c->solve_cache.end()) {
            const cb200_collider::CachedSolve cs = it->second;
            c->solve_cache.erase(it);
            if (cs.ver_a == c->infos[ev.a].version && cs.ver_b == c->infos[ev.b].version) {
                if (cs.err) return contract_error(c, cs.err, ev.a);
                *time_out = cs.time;
                return CB200_OK;
            }
        }
        int rc = prefetch_solves(c, uid, ev, c->time);
        if (rc != CB200_OK) return rc;
    }
    return cfail(c, CB200_E_STATE, "solve prefetch failed to produce the requested event");
}

// Collider::process_collision (collider.rs:177-197)
int process_collision(cb200_collider* c, uint64_t a, uint64_t b, const double* sep_time, uint64_t uid) {
    pair_insert(c, a, b);
    double t;
    if (sep_time) t = *sep_time;
    else {
        int rc = pair_solve(c, uid, QEvent{CB200_EV_COLLIDE, a, b}, &t);
        if (rc != CB200_OK) return rc;
    }
    return add_event(c, t, QEvent{CB200_EV_SEPARATE, a, b});
}

}  // namespace

extern "C" {

int cb200_collider_new(double cell_width, double padding, int device, cb200_collider** out) {
    if (!out) return CB200_E_ARG;
    *out = nullptr;
    cb200_ctx* ctx = nullptr;
    int rc = cb200_create(cell_width, padding, device, &ctx);
    if (rc != CB200_OK) return rc;
    cb200_collider* c = new cb200_collider;
    c->ctx = ctx;
    auto bail = [&](const char* what, cudaError_t e) {
        g_create_err = std::string(what) + ": " + cudaGetErrorString(e);
        cb200_collider_free(c);
        return CB200_E_CUDA;
    };
    cudaError_t e;
    if ((e = c->head.reserve(1)) != cudaSuccess) return bail("cudaHostAlloc", e);
    if ((e = c->items.reserve(256)) != cudaSuccess) return bail("cudaHostAlloc", e);
    if ((e = c->ov.reserve(64)) != cudaSuccess) return bail("cudaHostAlloc", e);
    if ((e = c->qout.reserve(256)) != cudaSuccess) return bail("cudaHostAlloc", e);
    if ((e = cudaMalloc(&c->d_dirty_count, 4)) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMemset(c->d_dirty_count, 0, 4)) != cudaSuccess) return bail("cudaMemset", e);
    ctx->collider_dirty = nullptr;
    *out = c;
    return CB200_OK;
}

void cb200_collider_free(cb200_collider* c) {
    if (!c) return;
    if (c->ctx) {
        cudaSetDevice(c->ctx->device);
        cudaStreamSynchronize(c->ctx->stream);
    }
    c->head.release(); c->items.release(); c->ov.release(); c->qout.release(); c->jobs.release(); c->job_out.release();
    c->bq_items.release(); c->bq_heads.release(); c->bq_out.release();
    c->dirty.release(); c->dirty_list.release();
    if (c->d_dirty_count) cudaFree(c->d_dirty_count);
    cb200_destroy(c->ctx);
    delete c;
}

const char* cb200_collider_last_error(const cb200_collider* c) { return c ? c->err.c_str() : g_create_err.c_str(); }
cb200_ctx* cb200_collider_engine(cb200_collider* c) { return c ? c->ctx : nullptr; }
uint64_t cb200_collider_len(const cb200_collider* c) { return c ? c->infos.size() : 0; }
uint64_t cb200_collider_rebin_count(const cb200_collider* c) { return c ? c->rebins : 0; }
uint64_t cb200_collider_prefetch_count(const cb200_collider* c) { return c ? c->prefetches : 0; }
uint64_t cb200_collider_hitbox_cache_hits(const cb200_collider* c) { return c ? c->hitbox_cache_hits : 0; }

int cb200_collider_set_can_interact(cb200_collider* c, cb200_can_interact_fn fn, void* user) {
    if (!c) return CB200_E_ARG;
    c->can_interact = fn;
    c->can_interact_user = user;
    return CB200_OK;
}
int cb200_collider_set_rebin_threshold(cb200_collider* c, uint32_t dirty_rows) {
    if (!c) return CB200_E_ARG;
    c->rebin_threshold = dirty_rows;
    return CB200_OK;
}

double cb200_collider_time(const cb200_collider* c) { return c ? c->time : 0.0; }  // collider.rs:72-74

int cb200_collider_next_time(cb200_collider* c, double* out) {  // collider.rs:85-87
    if (!c || !out) return CB200_E_ARG;
    cudaSetDevice(c->ctx->device);
    Head h;
    int rc = queue_head(c, &h);
    if (rc != CB200_OK) return rc;
    *out = h.any ? h.time : INFINITY;
    return CB200_OK;
}

int cb200_collider_set_time(cb200_collider* c, double time) {  // collider.rs:97-102
    if (!c) return CB200_E_ARG;
    double nt;
    int rc = cb200_collider_next_time(c, &nt);
    if (rc != CB200_OK) return rc;
    if (!(time >= c->time)) return cfail(c, CB200_E_CONTRACT, "cannot rewind time");
    if (!(time <= nt)) return cfail(c, CB200_E_CONTRACT, "time must not exceed next_time()");
    if (!(time < 1e50)) return cfail(c, CB200_E_CONTRACT, "time must not exceed 1e50");
    c->time = time;
    return CB200_OK;
}

int cb200_collider_next(cb200_collider* c, int* has_event, uint32_t* kind, uint64_t* id_1, uint64_t* id_2) {  // collider.rs:113-175
    if (!c || !has_event) return CB200_E_ARG;
    cudaSetDevice(c->ctx->device);
    *has_event = 0;
    for (;;) {
        Head h;
        int rc = queue_head(c, &h);
        if (rc != CB200_OK) return rc;
        if (!h.any || !(h.time == c->time)) return CB200_OK;  // EventManager::next (events.rs:175-189)
        const uint64_t uid = uid_of(h.from_base, c->base_cur, h.key);
        if (h.from_base) {
            c->base_cur += 1;
        } else {
            c->overlay.erase(h.key);
            drop_key(c->infos[h.ev.a].keys, h.key);
            if (h.ev.kind != CB200_EV_REITERATE) drop_key(c->infos[h.ev.b].keys, h.key);
        }
        const QEvent ev = h.ev;
        if (ev.kind == CB200_EV_REITERATE) {
            if ((rc = internal_update(c, ev.a, nullptr, false)) != CB200_OK) return rc;
            continue;
        }
        if (ev.kind == CB200_EV_COLLIDE) {
            if ((rc = process_collision(c, ev.a, ev.b, nullptr, uid)) != CB200_OK) return rc;
        } else {
            if (!overlap_remove(c->infos[ev.a].overlaps, ev.b) || !overlap_remove(c->infos[ev.b].overlaps, ev.a))
                return cfail(c, CB200_E_STATE, "Separate event for a pair that is not overlapping");
            c->overlap_pairs.erase({std::min(ev.a, ev.b), std::max(ev.a, ev.b)});
            c->overlaps_dirty = true;
            double t;
            if ((rc = pair_solve(c, uid, ev, &t)) != CB200_OK) return rc;
            if ((rc = add_event(c, t, QEvent{CB200_EV_COLLIDE, ev.a, ev.b})) != CB200_OK) return rc;
        }
        *has_event = 1;
        if (kind) *kind = ev.kind;
        if (id_1) *id_1 = std::min(ev.a, ev.b);  // Collider::new_event (collider.rs:513-519)
        if (id_2) *id_2 = std::max(ev.a, ev.b);
        return CB200_OK;
    }
}

int cb200_collider_get_hitbox(cb200_collider* c, uint64_t id, cb200_hitbox* out) {  // collider.rs:200-202
    if (!c || !out) return CB200_E_ARG;
    auto it = c->infos.find(id);
    if (it == c->infos.end()) return not_found(c, id);
    cb200_ctx* x = c->ctx;
    {
        // the answer may have come back with the prefetched solve of the event the caller is handling (prefetch_solves)
        uint64_t tbits;
        memcpy(&tbits, &c->time, 8);
        auto hc = c->hitbox_cache.find(cb200_collider::HbKey{id, tbits});
        if (hc != c->hitbox_cache.end() && hc->second.version == it->second.version) {
            const cb200_collider::CachedHitbox& h = hc->second;
            if (h.err) return contract_error(c, h.err, id);
            out->pos_x = h.hb[0]; out->pos_y = h.hb[1]; out->dim_x = h.hb[2]; out->dim_y = h.hb[3];
            out->vel_x = h.hb[4]; out->vel_y = h.hb[5]; out->rsz_x = h.hb[6]; out->rsz_y = h.hb[7];
            out->end_time = h.hb[8];
            out->kind = h.kind;
            out->_pad = 0;
            c->hitbox_cache_hits += 1;
            return CB200_OK;
        }
    }
    cudaSetDevice(x->device);
    k_get_hitbox<<<1, 1, 0, x->stream>>>(table_of(c), it->second.label, c->time, c->head.d);
    x->launches += 1;
    int rc = wait_head(c);
    if (rc != CB200_OK) return rc;
    const OneHeader& h = *c->head.h;
    if (h.err_flags) return contract_error(c, h.err_flags, id);
    out->pos_x = h.hb[0]; out->pos_y = h.hb[1]; out->dim_x = h.hb[2]; out->dim_y = h.hb[3];
    out->vel_x = h.hb[4]; out->vel_y = h.hb[5]; out->rsz_x = h.hb[6]; out->rsz_y = h.hb[7];
    out->end_time = h.hb[8];
    out->kind = h.kind;
    out->_pad = 0;
    return CB200_OK;
}

int cb200_collider_add_hitbox(cb200_collider* c, const cb200_profile* profile, const cb200_hitbox* hb, uint64_t* overlaps, uint64_t cap,
                              uint64_t* n_out) {  // collider.rs:214-222
    if (!c || !profile || !hb) return CB200_E_ARG;
    if (n_out) *n_out = 0;
    cb200_ctx* x = c->ctx;
    cudaSetDevice(x->device);
    const uint64_t id = profile->id;
    if (c->infos.count(id)) return cfail(c, CB200_E_CONTRACT, "assertion failed: self.hitboxes.insert(id, info).is_none()");
    if (profile->group != CB200_GROUP_NONE && profile->group > 31) return cfail(c, CB200_E_ARG, "group must be 0..31 or CB200_GROUP_NONE");
    if (x->sharded) return cfail(c, CB200_E_STATE, "the incremental API runs on one device");
    if (hb->kind > 1) return cfail(c, CB200_E_ARG, "bad shape kind");
    if (!(hb->dim_x >= 0.0 && hb->dim_y >= 0.0)) return cfail(c, CB200_E_CONTRACT, "dims must be non-negative");          // shape/mod.rs:44-47
    if (hb->kind == CB200_KIND_CIRCLE && !(hb->dim_x == hb->dim_y)) return cfail(c, CB200_E_CONTRACT, "circle width must equal height");
    int rc = maybe_rebin(c);
    if (rc != CB200_OK) return rc;
    if (x->n == 0 && !c->have_grid) collider_geom(c);
    uint32_t label = 0;
    c->infos[id] = cb200_collider::Info{0, profile->group, profile->interact_mask, {}, {}, ++c->version_counter};
    if ((rc = ensure_capacity(c, (size_t)x->n + 1, c->by_id.size() + 2)) != CB200_OK || (rc = assign_label(c, id, &label)) != CB200_OK) {
        c->infos.erase(id);
        c->by_id.erase(id);
        return rc;
    }
    if ((rc = ensure_capacity(c, (size_t)x->n + 1, (size_t)label + 1)) != CB200_OK) return rc;
    c->infos[id].label = label;
    c->id_of_label[label] = id;
    c->cleared[label] = 1;  // nothing of the run belongs to a hitbox added after it
    c->ids_dirty = true;
    CCU(c->items.reserve(64));
    OneArgs a{};
    a.op = kOpAdd;
    a.label = label;
    a.row_new = x->n;
    a.px = hb->pos_x; a.py = hb->pos_y; a.w = hb->dim_x; a.h = hb->dim_y;
    a.kind = hb->kind; a.group = profile->group; a.mask = profile->interact_mask;
    a.vx = hb->vel_x; a.vy = hb->vel_y; a.rw = hb->rsz_x; a.rh = hb->rsz_y; a.end_time = hb->end_time;
    a.T = c->time;
    a.cell_width = x->cell_width;
    a.padding = x->padding;
    a.ov_labels = c->ov.d;
    a.n_ov = 0;
    const uint32_t n_before = x->n;
    rc = run_update(c, a);
    const OneHeader h = *c->head.h;
    if (rc != CB200_OK || h.err_flags) {
        x->n = n_before;
        c->infos.erase(id);
        c->by_id.erase(id);
        c->id_of_label[label] = UINT64_MAX;
        if (rc != CB200_OK) return rc;
        return contract_error(c, h.err_flags, id);
    }
    x->n = n_before + 1;
    x->have_state = true;
    x->tile_ok = false;  // a row was appended: the tile tables are rebuilt by the re-sort the next bulk step starts with
    x->need_resort = true;
    c->n_dirty = h.n_dirty;
    if (h.reit && (rc = add_event(c, h.end_time, QEvent{CB200_EV_REITERATE, id, 0})) != CB200_OK) return rc;
    std::vector<std::pair<uint64_t, OneItem>> col;
    for (uint32_t k = 0; k < h.n_items; ++k) col.emplace_back(c->id_of_label[c->items.h[k].label], c->items.h[k]);
    std::sort(col.begin(), col.end(), [](const std::pair<uint64_t, OneItem>& p, const std::pair<uint64_t, OneItem>& q) { return p.first < q.first; });
    uint64_t n_res = 0;
    for (auto& p : col) {
        if (!interacts(c, id, p.first)) continue;
        if (p.second.kind == kItemImmediate) {
            if (overlaps && n_res < cap) overlaps[n_res] = p.first;
            ++n_res;
            if ((rc = process_collision(c, id, p.first, &p.second.time, 0)) != CB200_OK) return rc;
        } else if ((rc = add_event(c, p.second.time, QEvent{CB200_EV_COLLIDE, id, p.first})) != CB200_OK) {
            return rc;
        }
    }
    if (n_out) *n_out = n_res;
    return CB200_OK;
}

// Batched add_hitbox: exactly `for k in ascending id: add_hitbox(profiles[k], hitboxes[k])` at the current time
// (collider.rs:214-222), as ONE device pass when the collider is empty — the table is uploaded and an add-mode step
// (SolveArgs.add_mode) bins it, finds the candidate pairs, reports the pairs that overlap immediately and queues Separate /
// Collide / Reiterate events.  A non-empty collider takes the hitboxes one by one.  pair_a[k] was added after pair_b[k].
int cb200_collider_add_hitboxes(cb200_collider* c, uint64_t n, const cb200_profile* profiles, const cb200_hitbox* hitboxes, uint64_t* pair_a,
                                uint64_t* pair_b, uint64_t cap_pairs, uint64_t* n_pairs) {
    if (!c || (n && (!profiles || !hitboxes))) return CB200_E_ARG;
    if (n_pairs) *n_pairs = 0;
    if (n == 0) return CB200_OK;
    cb200_ctx* x = c->ctx;
    cudaSetDevice(x->device);
    std::vector<uint32_t> order(n);
    for (uint64_t k = 0; k < n; ++k) order[k] = (uint32_t)k;
    std::sort(order.begin(), order.end(), [&](uint32_t p, uint32_t q) { return profiles[p].id < profiles[q].id; });
    for (uint64_t k = 1; k < n; ++k)
        if (profiles[order[k - 1]].id == profiles[order[k]].id) return cfail(c, CB200_E_CONTRACT, "assertion failed: self.hitboxes.insert(id, info).is_none()");
    uint64_t n_out = 0;
    if (!c->infos.empty() || c->can_interact || x->sharded || n >= 0x7ffffff0ull) {
        // one by one (also when a can_interact callback must see every candidate before an event exists)
        std::vector<uint64_t> ov(64);
        for (uint64_t k = 0; k < n; ++k) {
            uint64_t m = 0;
            int rc = cb200_collider_add_hitbox(c, &profiles[order[k]], &hitboxes[order[k]], ov.data(), ov.size(), &m);
            if (rc != CB200_OK) return rc;
            if (m > ov.size()) {
                ov.resize(m);
                if ((rc = cb200_collider_get_overlaps(c, profiles[order[k]].id, ov.data(), ov.size(), &m)) != CB200_OK) return rc;
            }
            for (uint64_t j = 0; j < m; ++j, ++n_out)
                if (n_out < cap_pairs) {
                    if (pair_a) pair_a[n_out] = profiles[order[k]].id;
                    if (pair_b) pair_b[n_out] = ov[j];
                }
        }
        if (n_pairs) *n_pairs = n_out;
        return CB200_OK;
    }
    // ---- empty collider: upload + one add-mode step ----
    for (uint64_t k = 0; k < n; ++k) {
        const cb200_profile& p = profiles[k];
        const cb200_hitbox& h = hitboxes[k];
        if (p.group != CB200_GROUP_NONE && p.group > 31) return cfail(c, CB200_E_ARG, "group must be 0..31 or CB200_GROUP_NONE");
        if (h.kind > 1) return cfail(c, CB200_E_ARG, "bad shape kind");
        if (!(h.dim_x >= 0.0 && h.dim_y >= 0.0)) return cfail(c, CB200_E_CONTRACT, "dims must be non-negative");
        if (h.kind == CB200_KIND_CIRCLE && !(h.dim_x == h.dim_y)) return cfail(c, CB200_E_CONTRACT, "circle width must equal height");
    }
    int rc = ensure_capacity(c, n, n);
    if (rc != CB200_OK) return rc;
    {
        std::vector<double> col(n);
        auto put = [&](int k, auto get) -> int {
            for (uint64_t i = 0; i < n; ++i) col[i] = get(hitboxes[order[i]]);
            CCU(cudaMemcpy(x->col[k].p, col.data(), n * 8, cudaMemcpyHostToDevice));
            return CB200_OK;
        };
        const double T = c->time;
        if ((rc = put(0, [](const cb200_hitbox& h) { return h.pos_x; })) != CB200_OK || (rc = put(1, [](const cb200_hitbox& h) { return h.pos_y; })) != CB200_OK ||
            (rc = put(2, [](const cb200_hitbox& h) { return h.dim_x; })) != CB200_OK || (rc = put(3, [](const cb200_hitbox& h) { return h.dim_y; })) != CB200_OK ||
            (rc = put(4, [](const cb200_hitbox& h) { return h.vel_x; })) != CB200_OK || (rc = put(5, [](const cb200_hitbox& h) { return h.vel_y; })) != CB200_OK ||
            (rc = put(6, [](const cb200_hitbox& h) { return h.rsz_x; })) != CB200_OK || (rc = put(7, [](const cb200_hitbox& h) { return h.rsz_y; })) != CB200_OK ||
            (rc = put(8, [T](const cb200_hitbox&) { return T; })) != CB200_OK || (rc = put(9, [](const cb200_hitbox& h) { return h.end_time; })) != CB200_OK ||
            (rc = put(10, [](const cb200_hitbox& h) { return h.end_time; })) != CB200_OK)
            return rc;
        std::vector<uint8_t> kind(n);
        std::vector<uint32_t> group(n), mask(n);
        for (uint64_t i = 0; i < n; ++i) {
            kind[i] = (uint8_t)hitboxes[order[i]].kind;
            group[i] = profiles[order[i]].group;
            mask[i] = profiles[order[i]].interact_mask;
        }
        CCU(cudaMemcpy(x->kind.p, kind.data(), n, cudaMemcpyHostToDevice));
        CCU(cudaMemcpy(x->group.p, group.data(), n * 4, cudaMemcpyHostToDevice));
        CCU(cudaMemcpy(x->mask.p, mask.data(), n * 4, cudaMemcpyHostToDevice));
        CCU(cudaMemset(x->reit.p, 0, n));
        k_iota_rows<<<((uint32_t)n + kThreads - 1) / kThreads, kThreads, 0, x->stream>>>((uint32_t)n, x->label.p, x->row_of.p);
    }
    x->n = (uint32_t)n;
    x->have_state = true;
    x->need_resort = true;
    x->steps_since_resort = 0;
    x->n_ov = 0;
    x->ov_bucket_mask = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const cb200_profile& p = profiles[order[i]];
        c->infos[p.id] = cb200_collider::Info{(uint32_t)i, p.group, p.interact_mask, {}, {}, ++c->version_counter};
        c->by_id[p.id] = (uint32_t)i;
        c->id_of_label[i] = p.id;
    }
    CCU(x->ids.reserve(n));
    CCU(cudaMemcpy(x->ids.p, c->id_of_label.data(), n * 8, cudaMemcpyHostToDevice));
    c->ids_dirty = false;
    collider_geom(c);
    x->collider_dirty = c->dirty.p;
    x->collider_dirty_count = c->d_dirty_count;
    // the add-mode step: like a re-run of a bulk step (no rebase; the user end_time is read from the stored column)
    if (!x->d_n_imm) CCU(cudaMalloc(&x->d_n_imm, 8));
    if (x->imm_pairs.cap == 0) CCU(x->imm_pairs.reserve(n / 4 + 1024));
    if ((rc = prepare_buffers(x)) != CB200_OK) return cfail(c, rc, x->err);
    if ((rc = set_step_dyn(x, c->time, INFINITY)) != CB200_OK) return cfail(c, rc, x->err);
    const double* nv[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    x->add_mode = true;
    x->step_is_tile = false;  // (the add-mode step runs on the global path)
    bool again = true;
    unsigned long long n_imm = 0;
    for (int attempt = 0; attempt < 5 && again; ++attempt) {
        CCU(cudaMemsetAsync(x->d_n_imm, 0, 8, x->stream));
        CCU(cudaEventRecord(x->ev_t[ST_PRE], x->stream));
        if ((rc = enqueue_front(x, c->time, false, nv, INFINITY)) != CB200_OK || (rc = enqueue_back(x, c->time, false)) != CB200_OK ||
            (rc = complete_step(x, c->time, &again)) != CB200_OK) {
            x->add_mode = false;
            return cfail(c, rc, x->err);
        }
        CCU(cudaMemcpy(&n_imm, x->d_n_imm, 8, cudaMemcpyDeviceToHost));
        if (n_imm > x->imm_pairs.cap) {
            CCU(x->imm_pairs.reserve(n_imm + n_imm / 4));
            again = true;
        }
        if (x->last.ctr.err_flags) break;
    }
    x->add_mode = false;
    cb200_step_result res;
    if ((rc = fill_result(x, &res)) != CB200_OK) {
        // a contract failure (e.g. a shape smaller than the padding): nothing was added
        const std::string msg = x->err;
        c->infos.clear();
        c->by_id.clear();
        std::fill(c->id_of_label.begin(), c->id_of_label.end(), UINT64_MAX);
        x->n = 0;
        x->have_step = false;
        return cfail(c, rc, msg);
    }
    if (again) return cfail(c, CB200_E_STATE, "step buffers failed to converge");
    if ((rc = sort_events(x)) != CB200_OK) return cfail(c, rc, x->err);
    // immediate overlaps (collider.rs:355-365)
    std::vector<uint2> imm(n_imm);
    if (n_imm) CCU(cudaMemcpy(imm.data(), x->imm_pairs.p, n_imm * sizeof(uint2), cudaMemcpyDeviceToHost));
    std::sort(imm.begin(), imm.end(), [](const uint2& p, const uint2& q) { return p.x < q.x || (p.x == q.x && p.y < q.y); });
    for (const uint2& pr : imm) {
        const uint64_t a_id = c->id_of_label[pr.x], b_id = c->id_of_label[pr.y];
        pair_insert(c, a_id, b_id);
        if (n_out < cap_pairs) {
            if (pair_a) pair_a[n_out] = a_id;
            if (pair_b) pair_b[n_out] = b_id;
        }
        ++n_out;
    }
    if (n_pairs) *n_pairs = n_out;
    c->overlaps_dirty = true;
    // the queue is the device-sorted run
    c->hitbox_cache.clear();
    c->solve_cache.clear();
    c->overlay.clear();
    std::fill(c->cleared.begin(), c->cleared.end(), 0);
    c->base.clear();
    c->base_off = c->base_cur = 0;
    c->base_total = res.n_events;
    c->base_on_host = false;
    c->next_index += res.n_events;
    c->have_grid = true;
    c->n_dirty = 0;
    return CB200_OK;
}

int cb200_collider_set_hitbox_vel(cb200_collider* c, uint64_t id, double vel_x, double vel_y, double rsz_x, double rsz_y, double end_time) {
    if (!c) return CB200_E_ARG;  // collider.rs:225-229
    cudaSetDevice(c->ctx->device);
    const double v[5] = {vel_x, vel_y, rsz_x, rsz_y, end_time};
    return internal_update(c, id, v, true);
}

int cb200_collider_remove_hitbox(cb200_collider* c, uint64_t id, uint64_t* overlaps, uint64_t cap, uint64_t* n_out) {  // collider.rs:256-275
    if (!c) return CB200_E_ARG;
    if (n_out) *n_out = 0;
    auto it = c->infos.find(id);
    if (it == c->infos.end()) return not_found(c, id);
    cb200_ctx* x = c->ctx;
    cudaSetDevice(x->device);
    {
        int rc = sync_grid(c);
        if (rc != CB200_OK) return rc;
    }
    x->tile_ok = false;  // rows move: the tile tables are rebuilt by the re-sort the next bulk step starts with
    x->need_resort = true;
    clear_related(c, id, it->second);
    k_remove_row<<<1, 1, 0, x->stream>>>(table_of(c), it->second.label, c->head.d);
    x->launches += 1;
    int rc = wait_head(c);
    if (rc != CB200_OK) return rc;
    x->n -= 1;
    c->n_dirty = c->head.h->n_dirty;
    // clear_overlaps (collider.rs:383-393)
    uint64_t n_res = 0;
    for (uint64_t other : it->second.overlaps) {
        overlap_remove(c->infos[other].overlaps, id);
        c->overlap_pairs.erase({std::min(other, id), std::max(other, id)});
        if (overlaps && n_res < cap) overlaps[n_res] = other;
        ++n_res;
    }
    if (n_out) *n_out = n_res;
    c->id_of_label[it->second.label] = UINT64_MAX;
    c->by_id.erase(id);
    c->infos.erase(it);
    c->ids_dirty = true;
    c->overlaps_dirty = true;
    return CB200_OK;
}

int cb200_collider_get_overlaps(cb200_collider* c, uint64_t id, uint64_t* out, uint64_t cap, uint64_t* n_out) {  // collider.rs:279-288
    if (!c) return CB200_E_ARG;
    auto it = c->infos.find(id);
    if (it == c->infos.end()) return not_found(c, id);
    const auto& v = it->second.overlaps;
    for (size_t k = 0; k < v.size() && k < cap; ++k)
        if (out) out[k] = v[k];
    if (n_out) *n_out = v.size();
    return CB200_OK;
}

int cb200_collider_is_overlapping(cb200_collider* c, uint64_t id_1, uint64_t id_2, int* out) {  // collider.rs:300-305
    if (!c || !out) return CB200_E_ARG;
    auto it = c->infos.find(id_1);
    *out = it != c->infos.end() && std::binary_search(it->second.overlaps.begin(), it->second.overlaps.end(), id_2) ? 1 : 0;
    return CB200_OK;
}

int cb200_collider_query_overlaps(cb200_collider* c, uint32_t kind, double pos_x, double pos_y, double dim_x, double dim_y,
                                  const cb200_profile* profile, uint64_t* out, uint64_t cap, uint64_t* n_out) {  // collider.rs:309-318
    if (!c || !profile) return CB200_E_ARG;
    if (n_out) *n_out = 0;
    if (kind > 1) return cfail(c, CB200_E_ARG, "bad shape kind");
    if (!(dim_x >= 0.0 && dim_y >= 0.0)) return cfail(c, CB200_E_CONTRACT, "dims must be non-negative");
    if (kind == CB200_KIND_CIRCLE && !(dim_x == dim_y)) return cfail(c, CB200_E_CONTRACT, "circle width must equal height");
    cb200_ctx* x = c->ctx;
    cudaSetDevice(x->device);
    if (x->n == 0) return CB200_OK;
    int rc = maybe_rebin(c);
    if (rc != CB200_OK) return rc;
    for (int attempt = 0; attempt < 3; ++attempt) {
        QueryArgs q{};
        q.t = table_of(c);
        q.g = grid_of(c);
        q.q = PShape{pos_x, pos_y, dim_x, dim_y, (int)kind};
        q.mask = profile->interact_mask;
        q.T = c->time;
        q.cell_width = x->cell_width;
        q.head = c->head.d;
        q.out = c->qout.d;
        q.cap = (uint32_t)c->qout.cap;
        k_query<<<1, kThreads, 0, x->stream>>>(q);
        x->launches += 1;
        if ((rc = wait_head(c)) != CB200_OK) return rc;
        const OneHeader& h = *c->head.h;
        if (h.err_flags) return cfail(c, CB200_E_CONTRACT, flags_message(h.err_flags));
        if (h.n_items > c->qout.cap) {
            CCU(c->qout.reserve(h.n_items));
            continue;
        }
        std::vector<uint64_t> ids;
        for (uint32_t k = 0; k < h.n_items; ++k) {
            const uint64_t other = c->id_of_label[c->qout.h[k]];
            if (interacts(c, other, profile->id)) ids.push_back(other);
        }
        std::sort(ids.begin(), ids.end());
        for (size_t k = 0; k < ids.size() && k < cap; ++k)
            if (out) out[k] = ids[k];
        if (n_out) *n_out = ids.size();
        return CB200_OK;
    }
    return cfail(c, CB200_E_STATE, "query buffer failed to converge");
}

// Collider::query_overlaps (collider.rs:309-318) for a batch of shapes: one device launch, one block per query.
// ids of query q: out_ids[offsets[q] .. offsets[q + 1]), ascending.  offsets[n] = total; if it exceeds cap_ids only the
// queries that fit completely were written (call again with a larger buffer).
int cb200_collider_query_overlaps_batch(cb200_collider* c, uint64_t n, const cb200_query* queries, uint64_t* out_ids, uint64_t cap_ids, uint64_t* offsets) {
    if (!c || !offsets || (n && !queries)) return CB200_E_ARG;
    for (uint64_t q = 0; q <= n; ++q) offsets[q] = 0;
    if (n == 0) return CB200_OK;
    if (n > 0x3fffffffull) return cfail(c, CB200_E_ARG, "too many queries in one batch");
    cb200_ctx* x = c->ctx;
    cudaSetDevice(x->device);
    for (uint64_t q = 0; q < n; ++q) {
        const cb200_query& qu = queries[q];
        if (qu.kind > 1) return cfail(c, CB200_E_ARG, "bad shape kind");
        if (!(qu.dim_x >= 0.0 && qu.dim_y >= 0.0)) return cfail(c, CB200_E_CONTRACT, "dims must be non-negative");
        if (qu.kind == CB200_KIND_CIRCLE && !(qu.dim_x == qu.dim_y)) return cfail(c, CB200_E_CONTRACT, "circle width must equal height");
    }
    if (x->n == 0) return CB200_OK;
    int rc = maybe_rebin(c);
    if (rc != CB200_OK) return rc;
    CCU(c->bq_items.reserve(n));
    CCU(c->bq_heads.reserve(n));
    for (uint64_t q = 0; q < n; ++q)
        c->bq_items.h[q] = QueryItem{PShape{queries[q].pos_x, queries[q].pos_y, queries[q].dim_x, queries[q].dim_y, (int)queries[q].kind}, queries[q].profile.interact_mask};
    uint32_t cap_each = 32;
    for (int attempt = 0; attempt < 4; ++attempt) {
        CCU(c->bq_out.reserve((size_t)n * cap_each));
        QueryBatchArgs a{};
        a.t = table_of(c);
        a.g = grid_of(c);
        a.items = c->bq_items.d;
        a.T = c->time;
        a.cell_width = x->cell_width;
        a.heads = c->bq_heads.d;
        a.out = c->bq_out.d;
        a.cap_each = cap_each;
        k_query_batch<<<(uint32_t)n, kThreads, 0, x->stream>>>(a);
        x->launches += 1;
        if ((rc = wait_head(c)) != CB200_OK) return rc;
        uint32_t need = 0;
        for (uint64_t q = 0; q < n; ++q) {
            if (c->bq_heads.h[q].err_flags) return cfail(c, CB200_E_CONTRACT, flags_message(c->bq_heads.h[q].err_flags) + " (query " + std::to_string(q) + ")");
            need = std::max(need, c->bq_heads.h[q].n_items);
        }
        if (need <= cap_each) break;
        cap_each = need + need / 4;
        if (attempt == 3) return cfail(c, CB200_E_STATE, "query buffers failed to converge");
    }
    uint64_t total = 0;
    std::vector<uint64_t> ids;
    for (uint64_t q = 0; q < n; ++q) {
        ids.clear();
        const uint32_t m = c->bq_heads.h[q].n_items;
        for (uint32_t k = 0; k < m; ++k) {
            const uint64_t other = c->id_of_label[c->bq_out.h[(size_t)q * cap_each + k]];
            if (interacts(c, other, queries[q].profile.id)) ids.push_back(other);
        }
        std::sort(ids.begin(), ids.end());
        offsets[q] = total;
        if (out_ids && total + ids.size() <= cap_ids) std::copy(ids.begin(), ids.end(), out_ids + total);
        total += ids.size();
    }
    offsets[n] = total;
    return CB200_OK;
}

// HbGroup is an arbitrary u32 in the reference (core/mod.rs:199); the device keys its grid by a group INDEX 0..31 and takes
// interact_groups as a bit mask over those indices.  This maps a group value to its index, assigning indices in order of
// first use — a wrapper builds cb200_profile.group / .interact_mask from it (at most 32 distinct groups per collider).
int cb200_collider_group_index(cb200_collider* c, uint32_t group_value, uint32_t* index_out) {
    if (!c || !index_out) return CB200_E_ARG;
    auto it = c->group_index.find(group_value);
    if (it == c->group_index.end()) {
        if (c->group_index.size() >= 32) return cfail(c, CB200_E_ARG, "more than 32 distinct HbGroup values (the device grid key holds 32 group indices)");
        it = c->group_index.emplace(group_value, (uint32_t)c->group_index.size()).first;
    }
    *index_out = it->second;
    return CB200_OK;
}

// Benchmark driver: the user loop of the reference's README (README.md:66-80) run natively over this API — advance to the
// next event, pop it, and on Collide halve the velocity of both hitboxes (get_hitbox + set_hitbox_vel, end_time kept) —
// so that bench.py times the Collider API itself rather than Python's call overhead.  The oracle has the same loop
// (orc_halving_loop).  *n_events / *n_collides: user-visible events and the Collide events among them.
int cb200_collider_run_halving_loop(cb200_collider* c, double t_end, uint64_t* n_events, uint64_t* n_collides) {
    if (!c) return CB200_E_ARG;
    uint64_t n_ev = 0, n_col = 0;
    int rc = CB200_OK;
    while (c->time < t_end) {
        double nt;
        if ((rc = cb200_collider_next_time(c, &nt)) != CB200_OK) break;
        if ((rc = cb200_collider_set_time(c, nt < t_end ? nt : t_end)) != CB200_OK) break;
        int has = 0;
        uint32_t kind = 0;
        uint64_t ids[2] = {0, 0};
        if ((rc = cb200_collider_next(c, &has, &kind, &ids[0], &ids[1])) != CB200_OK) break;
        if (!has) continue;
        ++n_ev;
        if (kind != CB200_EV_COLLIDE) continue;
        ++n_col;
        for (int k = 0; k < 2 && rc == CB200_OK; ++k) {
            cb200_hitbox hb;
            if ((rc = cb200_collider_get_hitbox(c, ids[k], &hb)) != CB200_OK) break;
            rc = cb200_collider_set_hitbox_vel(c, ids[k], hb.vel_x * 0.5, hb.vel_y * 0.5, hb.rsz_x, hb.rsz_y, hb.end_time);
        }
        if (rc != CB200_OK) break;
    }
    if (n_events) *n_events = n_ev;
    if (n_collides) *n_collides = n_col;
    return rc;
}

// The bulk entry point on the Collider: for id ascending, internal_update_hitbox(id, Some(vel_id)) at the current time
// (SURVEY.md 8b).  `vel` arrays are in ascending-id order.
int cb200_collider_bulk_update(cb200_collider* c, const cb200_vel_view* vel, double end_time, cb200_step_result* out) {
    if (!c) return CB200_E_ARG;
    cb200_ctx* x = c->ctx;
    cudaSetDevice(x->device);
    if (x->n == 0) return cfail(c, CB200_E_STATE, "bulk_update on an empty collider");
    int rc;
    // dense labels: the caller's arrays are indexed by id rank
    bool dense = !c->by_id.empty() && std::prev(c->by_id.end())->second + 1 == c->by_id.size();
    if (!dense && (rc = relabel(c)) != CB200_OK) return rc;
    if (c->ids_dirty) {
        CCU(x->ids.reserve(c->by_id.size()));
        CCU(cudaMemcpyAsync(x->ids.p, c->id_of_label.data(), c->by_id.size() * 8, cudaMemcpyHostToDevice, x->stream));
        CCU(cudaStreamSynchronize(x->stream));
        c->ids_dirty = false;
    }
    if (c->overlaps_dirty) {
        std::vector<uint64_t> pa, pb;
        pa.reserve(c->overlap_pairs.size());
        pb.reserve(c->overlap_pairs.size());
        for (const auto& pr : c->overlap_pairs) {
            pa.push_back(pr.first);
            pb.push_back(pr.second);
        }
        rc = cb200_set_overlaps(x, pa.data(), pb.data(), pa.size());
        if (rc != CB200_OK) return cfail(c, rc, x->err);
        c->overlaps_dirty = false;
    }
    collider_geom(c);
    x->collider_dirty = c->dirty.p;
    x->collider_dirty_count = c->d_dirty_count;
    cb200_step_result res;
    rc = cb200_bulk_update(x, c->time, vel, end_time, &res);
    if (out) *out = res;
    if (rc != CB200_OK) return cfail(c, rc, x->err);
    // materialise the sorted run now: its solitaire keys are read from the state columns, which later single-hitbox
    // updates overwrite
    if ((rc = sort_events(x)) != CB200_OK) return cfail(c, rc, x->err);
    // every hitbox was updated: every older event is gone (clear_related_events per hitbox); the run is the queue
    c->solve_cache.clear();  // (so the hitbox versions need no bump)
    c->hitbox_cache.clear();
    c->overlay.clear();
    for (uint64_t id : c->with_keys) {
        auto it = c->infos.find(id);
        if (it != c->infos.end()) it->second.keys.clear();
    }
    c->with_keys.clear();
    std::fill(c->cleared.begin(), c->cleared.end(), 0);
    c->base.clear();
    c->base_off = c->base_cur = 0;
    c->base_total = res.n_events;
    c->base_on_host = false;
    c->next_index += res.n_events;
    c->have_grid = true;
    c->n_dirty = 0;
    if (c->can_interact) {
        // can_interact is host code: filter the Collide events of the run (HbProfile::can_interact, core/mod.rs:233-237)
        std::vector<cb200_event> all(res.n_events);
        uint64_t got = 0;
        if (res.n_events && (rc = cb200_fetch_events(x, all.data(), all.size(), 0, &got)) != CB200_OK) return cfail(c, rc, x->err);
        all.resize(got);
        c->base.clear();
        for (const cb200_event& e : all)
            if (e.kind != CB200_EV_COLLIDE || interacts(c, e.a, e.b)) c->base.push_back(e);
        c->base_total = c->base.size();
        c->base_on_host = true;
    }
    return CB200_OK;
}

}  // extern "C"
