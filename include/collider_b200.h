/* collider_b200.h — C ABI of the B200 device engine behind collider-rs's `Collider<P: HbProfile>`.
 *
 * This is the drop-in boundary.  The reference (SergiusIW/collider-rs 0.3.1, pure Rust)
 * has no FFI of its own; the cut is made at the three internal seams its orchestrator calls:
 *   (1) Grid::update_hitbox -> candidate ids           src/core/grid.rs:79-92   (called collider.rs:266-272,340-349)
 *   (2) DurHitbox::{collide_time, separate_time}       src/core/dur_hitbox/mod.rs:101-107 (called collider.rs:151-153,188-190,331-332,354)
 *   (3) EventManager::{peek_time, add_*_event}         src/core/events.rs:115-173 (called collider.rs:86,154,191,333,367,442)
 * and the per-hitbox driver around them, Collider::internal_update_hitbox / update_hitbox_tracking /
 * solitaire_event_check (src/core/collider.rs:231-250, 320-381, 420-448), executed for ALL hitboxes in one
 * device pass by cb200_bulk_update.
 *
 * Plain pointers and sizes only.  Every function returns 0 on success, a negative CB200_E_* code on
 * failure; cb200_last_error() gives the message (the Rust wrapper re-raises it as panic!, preserving the
 * reference's contract-panic behaviour — see INTEGRATION.md).  A context is single-threaded like the
 * reference's `&mut self` API: one caller at a time, one CUDA stream per context.
 *
 * There is no CPU fallback: every entry point that computes runs CUDA kernels on `device`; if no CUDA
 * device is usable, cb200_create fails with CB200_E_CUDA.
 */
#ifndef COLLIDER_B200_H
#define COLLIDER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CB200_ABI_VERSION 1

enum {
    CB200_OK = 0,
    CB200_E_ARG = -1,      /* bad argument (null pointer, ids not ascending, ...) */
    CB200_E_CUDA = -2,     /* CUDA runtime failure */
    CB200_E_CONTRACT = -3, /* a reference contract assertion failed on the device (see error_flags) */
    CB200_E_STATE = -4     /* call sequence error (e.g. bulk_update before upload_state) */
};

/* Contract violations detected on the device; names follow the reference's panics. */
enum {
    CB200_F_INVALID_TIME = 1u << 0,   /* "invalid time": T outside [start_time, pub_end_time]   collider.rs:489-492 */
    CB200_F_END_TIME = 1u << 1,       /* "end time must exceed present time"                      core/mod.rs:147-150 */
    CB200_F_CIRCLE_RESIZE = 1u << 2,  /* "circle resize velocity must maintain aspect ratio"      core/mod.rs:151-156 */
    CB200_F_TOO_SMALL = 1u << 3,      /* "shape width/height must be at least {padding}"          core/mod.rs:157-161,166 */
    CB200_F_EMPTY_RECT = 1u << 4,     /* "IndexRect contains no elements"                         index_rect.rs:26-29 */
    CB200_F_RECT_SPAN = 1u << 5,      /* index rect wider than the device grid period (engine limit, not a reference panic) */
    CB200_F_NAN = 1u << 6,            /* "unexpected NaN"                                         float.rs:31 */
    CB200_F_HIGH_TIME = 1u << 7,      /* "requires time < 1e50"                                   core/mod.rs:142 */
    CB200_F_ID_NOT_FOUND = 1u << 8,   /* "hitbox id {} not found"                                 collider.rs:235,260,283,294 */
    CB200_F_PEER_LOST = 1u << 9       /* sharded, direct exchange: a peer rank aborted or did not arrive within the timeout (engine condition) */
};

#define CB200_KIND_CIRCLE 0 /* ShapeKind::Circle, geom/shape/mod.rs:27-31 */
#define CB200_KIND_RECT 1   /* ShapeKind::Rect */
#define CB200_GROUP_NONE 0xFFFFFFFFu /* HbProfile::group() == None: never binned, no events (collider.rs:328) */

/* InternalEvent kinds (events.rs:74-82, release build) */
#define CB200_EV_REITERATE 0
#define CB200_EV_COLLIDE 1
#define CB200_EV_SEPARATE 2

typedef struct cb200_ctx cb200_ctx;

/* Host SoA view of the hitbox table: HitboxInfo fields (collider.rs:457-464) column by column.
 * `id` must be strictly ascending (array order == HbId order; the canonical tie-break of SURVEY.md §8c
 * relies on the rank of an id).  group: HbProfile::group() value in 0..31 or CB200_GROUP_NONE; interact_mask: bit g set
 * iff g is in HbProfile::interact_groups() (core/mod.rs:218-231). */
typedef struct cb200_state_view {
    uint64_t n;
    const uint64_t* id;
    const double* pos_x;        /* Hitbox.value.pos          core/mod.rs:119-130 */
    const double* pos_y;
    const double* dim_x;        /* Hitbox.value.shape.dims */
    const double* dim_y;
    const double* vel_x;        /* Hitbox.vel.value          core/mod.rs:34-59 */
    const double* vel_y;
    const double* rsz_x;        /* Hitbox.vel.resize */
    const double* rsz_y;
    const double* start_time;   /* HitboxInfo.start_time */
    const double* end_time;     /* internal Hitbox.vel.end_time (collider.rs:440) */
    const double* pub_end_time; /* HitboxInfo.pub_end_time */
    const uint8_t* kind;        /* CB200_KIND_* */
    const uint32_t* group;
    const uint32_t* interact_mask;
} cb200_state_view;

typedef struct cb200_state_out {
    uint64_t n; /* capacity of every array below, in hitboxes */
    double *pos_x, *pos_y, *dim_x, *dim_y, *vel_x, *vel_y, *rsz_x, *rsz_y, *start_time, *end_time, *pub_end_time;
} cb200_state_out;

/* New velocities for a bulk step; any pointer may be NULL (= keep the current component). */
typedef struct cb200_vel_view {
    const double* vel_x;
    const double* vel_y;
    const double* rsz_x;
    const double* rsz_y;
    const double* end_time; /* per-hitbox HbVel.end_time; NULL = use the scalar argument */
} cb200_vel_view;

/* One queued event as the host's EventManager would hold it (events.rs:28-100): `a` is the hitbox that
 * was being updated when the event was created (always the higher HbId after a bulk step),
 * `b` the partner (0 for Reiterate). */
typedef struct cb200_event {
    double time;
    uint64_t a;
    uint64_t b;
    uint32_t kind; /* CB200_EV_* */
    uint32_t _pad;
} cb200_event;

typedef struct cb200_step_result {
    double next_time;       /* EventManager::peek_time (events.rs:171-173): min event time or +inf */
    cb200_event next_event; /* the minimum under the canonical order (time, class, hi, kind, lo); kind/ids 0 if none */
    uint64_t n_hitboxes;    /* N: hitboxes updated (a-2 … a-7 each) */
    uint64_t n_entries;     /* E: (hitbox, cell) entries binned */
    uint64_t n_cells;       /* C: grid slots touched (non-empty) */
    uint64_t n_collide_solves;  /* unique candidate pairs, one collide_time each (a-10) */
    uint64_t n_separate_solves; /* overlapping pairs, one separate_time each (a-11) */
    uint64_t n_events;      /* V: events emitted (time < 1e50), solitaire included */
    uint64_t n_solitaire;   /* Reiterate events among them */
    uint32_t error_flags;   /* CB200_F_* */
    uint32_t _pad;
    uint64_t error_slot;    /* lowest-numbered reporting slot's id, when error_flags != 0 */
    uint64_t n_work_items;  /* W: (hitbox, hitbox) pairs of one slot run enumerated for the candidate test (>= n_collide_solves) */
} cb200_step_result;

/* Collider::new (collider.rs:59-69): requires cell_width > padding > 0 (else CB200_E_ARG with the
 * reference's message).  `device` is the CUDA ordinal.  *out receives the context. */
int cb200_create(double cell_width, double padding, int device, cb200_ctx** out);
void cb200_destroy(cb200_ctx* ctx);
/* Message of the last failure on `ctx` (or of the last failed cb200_create when ctx == NULL). */
const char* cb200_last_error(const cb200_ctx* ctx);
int cb200_abi_version(void);

/* Device grid period (slots = 2^log2_w x 2^log2_h, cells folded modulo the period; aliasing costs time,
 * never correctness).  0,0 = choose from the uploaded extent (default). */
int cb200_set_grid(cb200_ctx* ctx, int log2_w, int log2_h);
int cb200_get_grid(cb200_ctx* ctx, int* log2_w, int* log2_h);

/* Table order.  The device keeps the rows of the hitbox table sorted by the grid slot of their centre cell (so that
 * binning atomics, entry writes and record gathers stay in nearby cache lines) and re-sorts them after an upload and
 * then every `steps` bulk steps; 0 = only after an upload.  Default 16 (environment: CB200_RESORT_PERIOD).  The order
 * affects speed only: every input and output of this API stays in the caller's (ascending-id) order. */
int cb200_set_resort_period(cb200_ctx* ctx, int steps);
uint64_t cb200_resort_count(const cb200_ctx* ctx);

/* Event-queue order (EventKey's Ord, events.rs:28-68, under the canonical tie order): cb200_fetch_events sorts the
 * step's events on the device with a one-shot bucket sort (ties at the step time cut on the tie word, later events cut
 * by time) and copies them out behind it — one synchronisation per fetch.  A step with a heavy tie at a LATER time takes
 * the recursive path instead; this counts how often that happened (environment: CB200_SORT=slow forces it). */
uint64_t cb200_event_sort_fallbacks(const cb200_ctx* ctx);

/* Pinned (page-locked) host memory for the arrays handed to upload_state / bulk_update / fetch_events, so that
 * the H2D / D2H copies inside those calls run at full PCIe speed.  Plain malloc'd memory is accepted too. */
void* cb200_host_alloc(uint64_t bytes);
void cb200_host_free(void* p);

/* Replace the device-resident hitbox table (H2D).  Replaces FnvHashMap<HbId, HitboxInfo> (collider.rs:31). */
int cb200_upload_state(cb200_ctx* ctx, const cb200_state_view* view);
/* Read the table back (D2H) — Collider::get_hitbox's data source (collider.rs:200-202). */
int cb200_download_state(cb200_ctx* ctx, const cb200_state_out* out);

/* Replace the set of currently-overlapping pairs (HitboxInfo.overlaps, collider.rs:463) as id pairs. */
int cb200_set_overlaps(cb200_ctx* ctx, const uint64_t* id_a, const uint64_t* id_b, uint64_t n_pairs);

/* Window drain: process, on the device, the pair events of every bulk step that fall inside the step's own window
 * [time, end_time] — exactly as Collider::next / process_event / process_collision (collider.rs:113-197) would for a caller
 * that pops every event up to end_time WITHOUT reacting to it (no set_hitbox_vel in between): Collide -> the pair joins the
 * overlap set and its Separate is computed at the event time, Separate -> it leaves and its next Collide is computed; every
 * follow-up event with time <= end_time is appended to the step's event list (cb200_fetch_events then holds every event the
 * reference delivers up to end_time), and the overlap set — kept on the device from step to step, no cb200_set_overlaps
 * between steps — is the set at end_time.  Exact as long as no Reiterate event falls inside a window (a hitbox crossing a
 * whole cell within one step); cb200_last_drain reports how many did.  One device only; the step takes one scalar end_time.
 * cb200_fetch_overlaps reads the set back (id pairs, higher id first). */
int cb200_set_window_drain(cb200_ctx* ctx, int on);
int cb200_fetch_overlaps(cb200_ctx* ctx, uint64_t* id_a, uint64_t* id_b, uint64_t cap, uint64_t* n_out);
int cb200_last_drain(const cb200_ctx* ctx, uint64_t* followup_events, uint64_t* pairs_added, uint64_t* pairs_removed, uint64_t* pairs_now,
                     uint64_t* reiterates_in_windows);

/* The bulk advance/rebuild entry point.  Reference-equivalent definition: at `time`,
 *   for id in ascending HbId: Collider::internal_update_hitbox(id, Some(vel_id))   (collider.rs:231-250)
 * i.e. rebase, clear every queued event, solitaire_event_check, re-bin into the grid, candidate pairs,
 * collide_time / separate_time per pair, queue the events; then EventManager::peek_time.
 * `vel` may be NULL (keep velocities); `end_time` is the HbVel.end_time installed when vel->end_time is NULL. */
int cb200_bulk_update(cb200_ctx* ctx, double time, const cb200_vel_view* vel, double end_time, cb200_step_result* out);

/* ---- Spatial sharding across the GPUs of one box — one context per GPU / process ----------------
 * Cell columns are split into `world` contiguous strips: rank d owns columns [col_bounds[d], col_bounds[d+1])
 * (the first and last strips are open-ended).  Each rank keeps a set of resident hitboxes (upload_state; ids must be
 * < 2^31 - 1 and globally unique; residents change only through the owner migration below) and, every step, hands the
 * 96-byte record of each resident to every other rank whose strip its index rect touches: one halo slab per destination,
 * so a rank receives only what it has to bin.  A pair is solved by exactly one rank: the one whose strip holds the pair's
 * owner cell (max of the two rect starts).  With the collective exchange the caller moves the slabs between the two halves
 * (NCCL over NVLink: torch.distributed in this repo, ncclSend/ncclRecv from a Rust host); the global next event is the
 * minimum of the per-rank keys.
 * All three calls below are asynchronous on the context's stream except cb200_shard_result:
 *   cb200_set_stream  : run the context on the caller's CUDA stream (e.g. torch's current stream) so that the
 *                       exchange collectives are stream-ordered between the two halves — no host sync mid-step.
 *   cb200_shard_begin : stage A on the residents + halo pack.  Every record whose rect (+1 column) touches another strip is
 *                       appended to the halo slab of that strip's rank: *send_base holds `world` slabs of slab_bytes bytes,
 *                       slab d addressed to rank d (a 96-byte header whose first word is the record count, then the
 *                       records).  The caller exchanges them all-to-all: slab d of rank r -> slot r of rank d's
 *                       *recv_base (ncclSend/ncclRecv pairs; torch.distributed.all_to_all_single) — ONE collective.
 *   cb200_shard_finish: bins own + received records restricted to the own strip, candidate pairs, solves, min.
 *                       *local_key = device address of this rank's (time key, tie) minimum, 2 x u64: all-gather it and
 *                       take the lexicographic minimum for the global next event.
 *   cb200_shard_result: waits for the step and returns this rank's result. */
#define CB200_RECORD_BYTES 96
int cb200_shard_configure(cb200_ctx* ctx, int rank, int world, const int32_t* col_bounds /* world + 1 */,
                          uint64_t id_space /* every HbId of every rank is < id_space < 2^31 - 1 */);
int cb200_set_stream(cb200_ctx* ctx, void* cuda_stream);
int cb200_shard_set_slab(cb200_ctx* ctx, uint64_t records); /* capacity of the halo slab; default max(4095, n/64): the direct exchange only moves the records that are there; before upload_state */

/* Owner migration ("hitboxes migrate owners when `start` crosses a strip edge"; the reference keeps one grid,
 * grid.rs:37-70, so there is nothing to cite beyond the cell rule of index_bounds, dur_hitbox/mod.rs:83-102).  Between two
 * steps of a sharded context:
 *   cb200_shard_take_movers  every row whose centre column (floor(pos_x / cell_width)) lies more than margin_cols columns
 *                            outside this rank's strip is copied into `out` — every state column, its profile and dest[i] = the
 *                            rank whose strip holds it — and removed from the table.  *n_out = rows moved.  If out is NULL or
 *                            out->cap is too small, nothing is removed, *n_out = the rows that qualify and CB200_E_ARG is
 *                            returned (out == NULL, cap 0: a pure count, CB200_OK when nothing qualifies)
 *   cb200_shard_add_rows     appends rows (ids strictly ascending, none resident here) — what the peers' take_movers addressed
 *                            to this rank.  The table has room for cb200_shard_set_row_capacity rows (default: 25 % more than
 *                            were uploaded); the receive area of the halo exchange does not move, the peers' mappings stay valid
 * Afterwards the caller's array order (new velocities of cb200_vel_view, cb200_download_state) is the ascending-id order of
 * the rank's residents, as after cb200_upload_state.  Which rank solves a pair never depended on residence (owner-cell rule):
 * results are unchanged; the last step's events can no longer be fetched. */
typedef struct cb200_rows_out {
    uint64_t cap; /* capacity of every array below, in rows */
    uint64_t* id;
    double *pos_x, *pos_y, *dim_x, *dim_y, *vel_x, *vel_y, *rsz_x, *rsz_y, *start_time, *end_time, *pub_end_time;
    uint8_t* kind;
    uint32_t* group;
    uint32_t* interact_mask;
    int32_t* dest;
} cb200_rows_out;
int cb200_shard_set_row_capacity(cb200_ctx* ctx, uint64_t rows); /* 0 = default; before upload_state */
int cb200_shard_rows(const cb200_ctx* ctx, uint64_t* rows, uint64_t* capacity, uint64_t* migrated_out, uint64_t* migrated_in);
int cb200_shard_take_movers(cb200_ctx* ctx, int32_t margin_cols, const cb200_rows_out* out, uint64_t* n_out);
int cb200_shard_add_rows(cb200_ctx* ctx, const cb200_state_view* rows);
int cb200_shard_begin(cb200_ctx* ctx, double time, const cb200_vel_view* vel, double end_time, void** send_base, void** recv_base,
                      uint64_t* slab_bytes);
int cb200_shard_finish(cb200_ctx* ctx, void** local_key);
int cb200_shard_result(cb200_ctx* ctx, cb200_step_result* out);

/* Direct halo exchange over peer memory (NVLink 5 / NVSwitch) — replaces the two collectives of a sharded step by device
 * kernels that write straight into the peers' memory: after stage A each rank copies its halo slab (only the records that
 * are there) into slot `rank` of every peer's visitor table and raises a flag in the peer's mailbox; the binning kernels
 * start once all flags of the step are up; the 24-byte next-event keys travel the same way after the solve stage.  One
 * host synchronisation per step, no NCCL call on the data path.
 *   cb200_shard_p2p_export : after cb200_upload_state; fills `info` (cudaIpc handles of the visitor table and the mailbox,
 *                            the receive offset).  The host exchanges the infos of all ranks (any transport: they are
 *                            sent once) and hands the table, indexed by rank, to
 *   cb200_shard_p2p_import : opens the peers (cudaIpcOpenMemHandle; contexts of the same process are addressed directly).
 *   cb200_shard_step       : the whole step, asynchronous on the context's stream.  Then cb200_shard_result (waits; local
 *                            result) and cb200_shard_global_next (the minimum over all ranks as (time key, tie):
 *                            cb200_step_result.next_event of the winning rank in the order of the device keys). */
#define CB200_IPC_HANDLE_BYTES 64
typedef struct cb200_p2p_info {
    uint8_t rec_handle[CB200_IPC_HANDLE_BYTES];
    uint8_t mail_handle[CB200_IPC_HANDLE_BYTES];
    uint64_t rec_ptr, mail_ptr;   /* valid in the exporting process only */
    uint64_t recv_offset_bytes;   /* of the receive area inside the visitor table allocation */
    uint64_t pid;
    int32_t device;
    uint32_t slab_records;
} cb200_p2p_info;
int cb200_shard_p2p_export(cb200_ctx* ctx, cb200_p2p_info* info);
int cb200_shard_p2p_import(cb200_ctx* ctx, const cb200_p2p_info* infos /* world entries, indexed by rank */);
int cb200_shard_step(cb200_ctx* ctx, double time, const cb200_vel_view* vel, double end_time);
int cb200_shard_global_next(cb200_ctx* ctx, uint64_t* time_key, uint64_t* tie, int* complete);
void* cb200_shard_local_key(cb200_ctx* ctx); /* device address of this rank's (time key, tie): input of a repeated key all-gather */
/* Every wait of a step kernel for a peer is bounded (environment CB200_PEER_TIMEOUT_MS, default 20000) and ends early when a
 * peer raises this rank's abort word; either way the step fails with CB200_F_PEER_LOST instead of hanging.  The library
 * raises the abort word of all peers when a sharded step of this rank fails; a host that stops stepping for a reason of
 * its own (an exception between steps) calls cb200_shard_abort so that its peers return at once. */
int cb200_shard_abort(cb200_ctx* ctx);

/* Events queued by the last bulk step, sorted by the canonical order (time, class, hi id, kind, lo id)
 * (events.rs:60-68 with SURVEY.md §8c tie-break).  Copies events [offset, offset+cap) into buf. */
int cb200_fetch_events(cb200_ctx* ctx, cb200_event* buf, uint64_t cap, uint64_t offset, uint64_t* n_out);

/* ---- Pipelined host I/O ----------------------------------------------------------------------------------------------
 * A bulk step is bound by PCIe when its inputs and outputs cross the host boundary every step.  These three calls let the
 * transfers run beside the compute (each direction on its own CUDA stream of the context, double-buffered):
 *   cb200_stage_velocities   starts the host -> device copy of a coming step's velocities and returns; up to two sets may
 *                            be staged, and each cb200_bulk_update called with vel == NULL consumes the oldest (the host
 *                            arrays must stay valid and unchanged until that step has returned; pinned memory —
 *                            cb200_host_alloc — for a truly asynchronous copy).  So the upload of step k+1 runs beside step k.
 *   cb200_fetch_events_begin sorts the last step's events on the device (as cb200_fetch_events) and starts their download
 *                            in the 16-byte compact form; returns without waiting, so the next step can be started
 *   cb200_fetch_events_end   waits for that download
 * cb200_event16: `a` = label of the first hitbox (bit 31 set: the event is a Reiterate), `b` = label of the second hitbox
 * (bit 31 set: Collide, clear: Separate).  A label is the rank of the HbId in the uploaded (ascending) id array — the host
 * holds that array — or the id itself in sharded mode. */
typedef struct cb200_event16 {
    double time;
    uint32_t a;
    uint32_t b;
} cb200_event16;
int cb200_stage_velocities(cb200_ctx* ctx, const cb200_vel_view* vel);
int cb200_fetch_events_begin(cb200_ctx* ctx, cb200_event16* buf, uint64_t cap, uint64_t* n_out);
int cb200_fetch_events_end(cb200_ctx* ctx);

/* Parity/debug views of the last bulk step.
 * Candidate pairs: (hi id, lo id) for every collide_time evaluated (unsorted).  Grid::overlapping_ids, grid.rs:115-135. */
int cb200_fetch_candidate_pairs(cb200_ctx* ctx, uint64_t* id_hi, uint64_t* id_lo, uint64_t cap, uint64_t* n_out);
/* Cell entries: (cell x, cell y, id) for every (hitbox, cell) binned (unsorted).  Grid.map, grid.rs:48-51. */
int cb200_fetch_cell_entries(cb200_ctx* ctx, int32_t* cell_x, int32_t* cell_y, uint64_t* id, uint64_t cap, uint64_t* n_out);
/* Per-hitbox IndexRect (sx, sy, ex, ey) x n.  Grid::index_bounds, grid.rs:101-113. */
int cb200_fetch_index_rects(cb200_ctx* ctx, int32_t* rects, uint64_t cap_hitboxes);

/* Batched narrowphase on explicit DurHitbox pairs (host arrays, n x 9 doubles each:
 * px,py,dw,dh,vx,vy,rw,rh,duration; kinds n bytes each).  DurHitbox::collide_time / separate_time,
 * dur_hitbox/mod.rs:101-107.  NaN is written where the reference would panic. */
int cb200_collide_time_batch(cb200_ctx* ctx, uint64_t n, const double* a, const uint8_t* kind_a, const double* b, const uint8_t* kind_b,
                             double* out);
int cb200_separate_time_batch(cb200_ctx* ctx, uint64_t n, const double* a, const uint8_t* kind_a, const double* b, const uint8_t* kind_b,
                              double padding, double* out);

/* ---- The reference's `Collider<P: HbProfile>` API on the device (src/core/collider.rs:59-318) ---------------------------
 * A cb200_collider owns one device context and mirrors the reference's public methods one to one; a Rust
 * `Collider<P>` wrapper keeps a `HashMap<HbId, P>` and forwards every call.  All geometry — rebase,
 * solitaire check, swept bounds, IndexRect, cell-mate enumeration, collide_time / separate_time, get_hitbox's advance,
 * query_overlaps' shape test — runs in CUDA kernels; the host keeps only bookkeeping without arithmetic (id maps,
 * HitboxInfo.overlaps, the order of the event queue, the clock).  Hitboxes may be added, re-aimed and removed one at a
 * time (k_update_one against the device grid + a dirty list) or all at once (cb200_collider_bulk_update = the bulk step).
 * Contract panics of the reference come back as CB200_E_CONTRACT with the reference's message.
 * Simultaneous events are delivered in the canonical order of SURVEY.md 8c. */
typedef struct cb200_collider cb200_collider;

typedef struct cb200_hitbox { /* Hitbox (core/mod.rs:119-130): PlacedShape + HbVel */
    double pos_x, pos_y, dim_x, dim_y; /* circle: dim_x == dim_y == diameter */
    double vel_x, vel_y, rsz_x, rsz_y;
    double end_time;                   /* HbVel.end_time (core/mod.rs:34-59) */
    uint32_t kind;                     /* CB200_KIND_* */
    uint32_t _pad;
} cb200_hitbox;

typedef struct cb200_profile { /* HbProfile (core/mod.rs:210-238) */
    uint64_t id;            /* id() */
    uint32_t group;         /* group(): 0..31 or CB200_GROUP_NONE */
    uint32_t interact_mask; /* interact_groups() as a bit mask */
} cb200_profile;

/* HbProfile::can_interact (core/mod.rs:233-237), evaluated on the host for every candidate pair; NULL = always true. */
typedef int (*cb200_can_interact_fn)(void* user, uint64_t id_a, uint64_t id_b);

int cb200_collider_new(double cell_width, double padding, int device, cb200_collider** out); /* Collider::new      :59-69  */
void cb200_collider_free(cb200_collider* c);
const char* cb200_collider_last_error(const cb200_collider* c);
int cb200_collider_set_can_interact(cb200_collider* c, cb200_can_interact_fn fn, void* user);
double cb200_collider_time(const cb200_collider* c);                                          /* time               :72-74  */
int cb200_collider_next_time(cb200_collider* c, double* out);                                 /* next_time          :85-87  */
int cb200_collider_set_time(cb200_collider* c, double time);                                  /* set_time           :97-102 */
/* next :113-124 — *has_event = 0 when no event is due at the current time; kind is CB200_EV_COLLIDE / CB200_EV_SEPARATE,
 * id_1 < id_2.  Reiterate events are processed internally (collider.rs:164-174). */
int cb200_collider_next(cb200_collider* c, int* has_event, uint32_t* kind, uint64_t* id_1, uint64_t* id_2);
int cb200_collider_get_hitbox(cb200_collider* c, uint64_t id, cb200_hitbox* out);             /* get_hitbox         :200-202 */
/* add_hitbox :214-222 — ids of the hitboxes the new one overlaps immediately (ascending), at most `cap` written, *n_out = count */
int cb200_collider_add_hitbox(cb200_collider* c, const cb200_profile* profile, const cb200_hitbox* hitbox, uint64_t* overlaps, uint64_t cap,
                              uint64_t* n_out);
/* Batched add_hitbox: the same as calling add_hitbox for every entry in ascending id order at the current time; on an
 * empty collider it is one device pass (upload + an add-mode step).  (pair_a[k], pair_b[k]): pair_a[k] overlapped
 * pair_b[k] the moment it was added (pair_a[k] > pair_b[k]); at most cap_pairs written, *n_pairs = count. */
int cb200_collider_add_hitboxes(cb200_collider* c, uint64_t n, const cb200_profile* profiles, const cb200_hitbox* hitboxes, uint64_t* pair_a,
                                uint64_t* pair_b, uint64_t cap_pairs, uint64_t* n_pairs);
int cb200_collider_set_hitbox_vel(cb200_collider* c, uint64_t id, double vel_x, double vel_y, double rsz_x, double rsz_y,
                                  double end_time);                                           /* set_hitbox_vel     :225-229 */
int cb200_collider_remove_hitbox(cb200_collider* c, uint64_t id, uint64_t* overlaps, uint64_t cap, uint64_t* n_out); /* :256-275 */
int cb200_collider_get_overlaps(cb200_collider* c, uint64_t id, uint64_t* out, uint64_t cap, uint64_t* n_out);       /* :279-288 */
int cb200_collider_is_overlapping(cb200_collider* c, uint64_t id_1, uint64_t id_2, int* out);                        /* :300-305 */
int cb200_collider_query_overlaps(cb200_collider* c, uint32_t kind, double pos_x, double pos_y, double dim_x, double dim_y,
                                  const cb200_profile* profile, uint64_t* out, uint64_t cap, uint64_t* n_out);       /* :309-318 */
/* Collider::query_overlaps for a batch of shapes in ONE device launch (one thread block per query).  ids of query q:
 * out_ids[offsets[q] .. offsets[q + 1]), ascending; offsets has n + 1 entries and offsets[n] is the total — if it exceeds
 * cap_ids, only the queries that fit completely were written. */
typedef struct cb200_query {
    uint32_t kind; /* CB200_KIND_* */
    uint32_t _pad;
    double pos_x, pos_y, dim_x, dim_y;
    cb200_profile profile; /* id (for can_interact), interact_mask */
} cb200_query;
int cb200_collider_query_overlaps_batch(cb200_collider* c, uint64_t n, const cb200_query* queries, uint64_t* out_ids, uint64_t cap_ids, uint64_t* offsets);
/* HbGroup is an arbitrary u32 in the reference (core/mod.rs:199); the device takes group INDICES 0..31 (cb200_profile.group)
 * and interact_groups as a bit mask over them.  Maps a group value to its index (assigned in order of first use; at most 32
 * distinct values per collider) so that a wrapper can accept any HbGroup. */
int cb200_collider_group_index(cb200_collider* c, uint32_t group_value, uint32_t* index_out);
/* The bulk advance/rebuild entry point on a Collider: at the current time, for id ascending,
 * internal_update_hitbox(id, Some(vel_id)) — one device pass (cb200_bulk_update).  `vel` arrays are in ascending-id order. */
int cb200_collider_bulk_update(cb200_collider* c, const cb200_vel_view* vel, double end_time, cb200_step_result* out);
/* Benchmark driver: the user loop of the reference's README (README.md:66-80) up to t_end, run natively over the calls above
 * with the Collide handler halving both hitboxes' velocities (get_hitbox + set_hitbox_vel) — BASELINE config 1/2's loop. */
int cb200_collider_run_halving_loop(cb200_collider* c, double t_end, uint64_t* n_events, uint64_t* n_collides);
uint64_t cb200_collider_len(const cb200_collider* c);
/* Rows changed since the grid was last rebuilt are kept on a dirty list that every single-hitbox operation scans; the
 * grid is rebuilt on the device (count / scan / scatter, no solve) when the list exceeds this many rows.  Default 4096. */
int cb200_collider_set_rebin_threshold(cb200_collider* c, uint32_t dirty_rows);
uint64_t cb200_collider_rebin_count(const cb200_collider* c);
/* next() needs one narrowphase solve per Collide / Separate event; the library computes them for up to 2048 queued events
 * per launch ahead of time and uses a prefetched result only if neither hitbox changed since.  Number of such launches. */
uint64_t cb200_collider_prefetch_count(const cb200_collider* c);
/* get_hitbox calls answered from the prefetch: the launch that solves the next queued events also evaluates
 * pub_hitbox_at_time (collider.rs:488-497) of their two hitboxes at the event time, which is what a Collide handler
 * asks for next; an answer is used only while the collider stands at that time and the hitbox is unchanged. */
uint64_t cb200_collider_hitbox_cache_hits(const cb200_collider* c);
cb200_ctx* cb200_collider_engine(cb200_collider* c); /* the underlying device context (timing, launch counts, grid period) */

/* Kernel launches issued by this context since creation (bench.py's gpu_launches). */
uint64_t cb200_launch_count(const cb200_ctx* ctx);
/* Which device path the last bulk step took: 1 = tile path (binning, candidate enumeration and pair solve per block of grid
 * cells in shared memory: k_stage_a_tile + k_tile_solve + k_solve for the overlapping pairs), 0 = global path (slot table,
 * entry array, work lists: k_stage_a, k_scan_slots, k_scatter, k_enum_pairs, k_solve[_dense]); -1 = no step yet.  Both are
 * exact; the tile path is taken by single-device bulk steps (environment CB200_TILE=0 disables it), and a step whose tiles
 * exceed the shared-memory capacities is repeated on the global path (cb200_tile_fallbacks counts those). */
int cb200_last_step_path(const cb200_ctx* ctx);
uint64_t cb200_tile_fallbacks(const cb200_ctx* ctx);
int cb200_get_tile(const cb200_ctx* ctx, int* log2_slots, int* cells_x, int* cells_y); /* tile size chosen at the last re-sort (0: none) */
/* Device-side duration of cb200_bulk_update and of each of its stages in milliseconds (CUDA events on the
 * context's stream).  For bench.py's roofline. */
/* Averages since the last cb200_reset_timing.  Stages: 0 pre (H2D velocities + table zeroing), 1 stage A (+ halo
 * pack), 2 halo exchange (sharded; the caller's collective on the context's stream), 3 slot scan (+ indexing of
 * received records), 4 scatter, 5 candidate stage, 6 solve stage + min-reduce + result, 7 tail. */
#define CB200_N_STAGES 8
int cb200_last_step_timing(cb200_ctx* ctx, float* total_ms, float* stage_ms, int n_stages);
int cb200_reset_timing(cb200_ctx* ctx);
/* on (default): CUDA events between the stages of a step (each costs ~2 us of stream time); off: only around the whole
 * step, stage_ms reads 0.  Resets the averages. */
int cb200_set_stage_timing(cb200_ctx* ctx, int on);

#ifdef __cplusplus
}
#endif
#endif /* COLLIDER_B200_H */
