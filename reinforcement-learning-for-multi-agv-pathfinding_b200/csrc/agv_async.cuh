// agv_async.cuh -- lane-per-scene fused tick with decoupled lanes.
//
// Same scene-per-lane layout as agv_lanes.cuh, but every lane keeps its OWN turn cursor: a
// scene whose AGV needs the level-synchronous BFS (no Manhattan-length path) posts the request
// to a BFS worker warp of the CTA and simply sits out the following iterations until the
// answer is there, while the other 31 scenes of the warp keep taking turns.  Turn order only
// exists inside a scene (Scene.py:124-162), so lanes never have to wait for each other: the
// tick ends when the slowest SCENE is done, not after 64 x (slowest lane of each iteration).
//
// Warp roles in a CTA:  warp 0           scene warp of the light scenes (one scene per lane)
//                       warp 1           scene warp of the heavy scenes (same lane <-> scene map;
//                                        a scene is heavy when its last tick needed many BFS)
//                       warps 2..NB+1    BFS workers  (claim posted requests, warp_bfs)
//                       warp NB+2        record writer (OBS_CLIP: expands 49-bit windows
//                                        to [3,7,7] bf16 records and stores them)
// All hand-offs are flags in shared memory (fence + volatile), polled with __nanosleep.
// Covers OBS_NONE and OBS_CLIP with K = 7 without the shaped reward; the other variants
// run on tick_lanes (agv_lanes.cuh).  Results are bit-identical to k_tick.
#pragma once
#include "agv_lanes.cuh"

namespace agv {

struct AsyncShared {
    unsigned bfs_state[32];  // per scene slot: 0 idle, 1 posted, 2 taken by a worker, 3 answered
    unsigned bfs_req[32];    // coop_bfs request word
    unsigned bfs_res[32];    // action | len << 8
    int exit_flag;
    unsigned done_warps;     // scene warps that have finished their lanes
    int pad[2];
    // record batches, one queue per scene warp (q = 0 light scenes, 1 heavy scenes)
    unsigned enc_seq[2];     // batches posted
    unsigned enc_done[2];    // batches written by the record writer
    int enc_n[2][2];
    uint4 enc_req[2][2][64]; // {window lo, window hi, tpos | which << 8, record index}
};

// A scene whose previous tick needed at least this many BFS searches is "heavy": it is carried by
// the CTA's second scene warp, whose iterations are short (few lanes), instead of dragging along
// (and being slowed down by) the 30-odd light scenes of the first.  The count lives in the scene
// header's reserved word.
#ifndef AGV_HEAVY_BFS
#define AGV_HEAVY_BFS 16
#endif

constexpr int LS_START = 0, LS_WAIT = 1, LS_HAVE = 2, LS_DONE = 3;
#ifndef AGV_POLL_NS
#define AGV_POLL_NS 400  // sleep between two polls of an idle worker warp (40 ns cost 10 %: the polls compete for issue slots)
#endif

template <int NW, int RPL, int SPW>
__device__ __forceinline__ void async_bfs_worker(const DevCfg &c, const LaneArgs &t, uint8_t *smem) {
    const int lane = threadIdx.x & 31, HN = c.H * NW;
    AsyncShared *sh = reinterpret_cast<AsyncShared *>(smem + t.off_queue);
    volatile unsigned *vs = sh->bfs_state;
    volatile int *vexit = &sh->exit_flag;
    const uint64_t *tab = reinterpret_cast<const uint64_t *>(smem + t.off_tab);
    const uint64_t *occ0 = reinterpret_cast<const uint64_t *>(smem + t.off_occ);
    for (;;) {
        unsigned m = __ballot_sync(FULL, vs[lane] == 1u);
        if (!m) {
            if (*vexit) return;
            __nanosleep(AGV_POLL_NS);
            continue;
        }
        int j = -1;
        while (m) {
            const int k = __ffs(m) - 1;
            m &= m - 1;
            int ok = 0;
            if (lane == 0) ok = atomicCAS(&sh->bfs_state[k], 1u, 2u) == 1u;
            ok = __shfl_sync(FULL, ok, 0);
            if (ok) { j = k; break; }
        }
        if (j < 0) continue;
        __threadfence_block();
        const unsigned pk = *reinterpret_cast<volatile unsigned *>(&sh->bfs_req[j]);
        int bl = 0;
#ifdef AGV_LANES_PROF
        const long long b0_ = clock64();
#endif
        const int ba = coop_bfs<NW, RPL>(pk, tab, HN, occ0 + (size_t)j * t.occ_stride, c.H, c.W, lane, bl);
#ifdef AGV_LANES_PROF
        if (lane == 0) {  // [11] BFS cycles, [12] searches, [14] path lengths; [6..10]: searches without a path
            const long long d_ = clock64() - b0_;
            atomicAdd(c.counters + 11, (unsigned long long)d_);
            atomicAdd(c.counters + 12, 1ull);
            atomicAdd(c.counters + 14, (unsigned long long)bl);
            if (bl == 0) {
                atomicAdd(c.counters + 6, (unsigned long long)d_);
                atomicAdd(c.counters + 7, 1ull);
                atomicAdd(c.counters + (d_ < 1000 ? 8 : (d_ < 4000 ? 9 : 10)), 1ull);
            }
        }
#endif
        if (lane == 0) {
            *reinterpret_cast<volatile unsigned *>(&sh->bfs_res[j]) = (unsigned)ba | ((unsigned)bl << 8);
            __threadfence_block();
            vs[j] = 3u;
        }
        __syncwarp();
    }
}

__device__ __forceinline__ void async_record_writer(const LaneArgs &t, uint8_t *smem) {
    const int lane = threadIdx.x & 31;
    AsyncShared *sh = reinterpret_cast<AsyncShared *>(smem + t.off_queue);
    volatile unsigned *vseq = sh->enc_seq, *vdone = sh->enc_done;
    volatile int *vexit = &sh->exit_flag;
    unsigned mine[2] = {0, 0};
    for (;;) {
        bool any = false;
#pragma unroll
        for (int q = 0; q < 2; q++) {
            if (vseq[q] == mine[q]) continue;
            any = true;
            __threadfence_block();
            const int buf = mine[q] & 1;
            const int n = *reinterpret_cast<volatile int *>(&sh->enc_n[q][buf]);
            for (int r = 0; r < n; r++) {
                const uint4 rq = sh->enc_req[q][buf][r];
                uint16_t *base = (rq.z >> 8) ? t.next_obs : t.obs;
                expand_clip7(base + (size_t)rq.w * 147, (uint64_t)rq.x | ((uint64_t)rq.y << 32), (int)(rq.z & 255u), lane);
            }
            __threadfence_block();
            __syncwarp();
            mine[q]++;
            if (lane == 0) vdone[q] = mine[q];
        }
        if (any) continue;
        if (*vexit) {  // the last batches are posted before the exit flag
            if (vseq[0] == mine[0] && vseq[1] == mine[1]) return;
            continue;
        }
        __nanosleep(AGV_POLL_NS);
    }
}

// scene warp `which` (0: the light scenes of the CTA, 1: its heavy scenes)
template <int NW, int RPL, int OBS, int SPW>
__device__ __forceinline__ void tick_async(const DevCfg &c, const LaneArgs &t, uint8_t *smem, const int which) {
    const int lane = threadIdx.x & 31;
    const int e0 = blockIdx.x * SPW;
    LaneScene<NW, RPL, SPW, 0> ws(c, t, smem, lane, e0);
    AsyncShared *sh = reinterpret_cast<AsyncShared *>(smem + t.off_queue);
    volatile unsigned *vs = sh->bfs_state;
    volatile unsigned *vseq = &sh->enc_seq[which], *vdone = &sh->enc_done[which];
    const int N = c.N;
    if (which == 0) {
        ws.load_tables();
        ws.load_blocks(e0);
        // per-AGV outputs default to "no turn"
        const int nsc = min(SPW, c.E - e0);
        const size_t r0 = (size_t)e0 * N;
        for (int q = lane; q < nsc * N; q += 32) {
            if (t.act_out) t.act_out[r0 + q] = -1;
            if (t.done) t.done[r0 + q] = 0;
            if (t.reward) t.reward[r0 + q] = 0.f;
            if (t.plen) t.plen[r0 + q] = 0;
            if (t.slen) t.slen[r0 + q] = 0xFFFF;
        }
        __threadfence_block();
    }
    named_bar_sync(3, 64);  // both scene warps: tables, scene blocks and default outputs are in place
    if (ws.has) ws.has = (((ws.hdr_w[6] & 0xFFFFu) >= (unsigned)AGV_HEAVY_BFS) ? 1 : 0) == which;  // (high half: turns, agv_rollout.cuh)
    unsigned n_bfs = 0;
    bool live = ws.has;
    if (ws.has) {
        ws.hdr_to_regs();
        if (ws.over) {
            if (c.auto_reset) ws.reset();
            else live = false;
        }
        if (live) { ws.build_occ(); ws.tick++; ws.acted_mask = 0; }
    } else {
        ws.over = 1; ws.finished = 0; ws.n_created = 0; ws.tick = 0; ws.cursor = 0; ws.n_done = 0; ws.acted_mask = 0;
    }
    __syncwarp();
    const int n_loop = live ? min(N, ws.n_created) : 0;
    const uint64_t om_a = N > 1 ? ~0ull : 0ull;
    const bool want_obs = OBS == OBS_CLIP && t.obs != nullptr, want_next = OBS == OBS_CLIP && t.next_obs != nullptr;
    const uint64_t *zr = reinterpret_cast<const uint64_t *>(smem + t.off_zero);

#ifdef AGV_LANES_PROF
    const long long p0_ = clock64();
    long long iters_ = 0, idle_ = 0;
#endif
    int i = 0, st = n_loop > 0 ? LS_START : LS_DONE;
    Agv a;  // the AGV of the turn in progress (kept in registers from phase C to phase B)
    a.x = a.y = a.tx = a.ty = 1; a.stage = a.loaded = a.idle = a.order = 0;
    int action = -1, plen = 0, given = -1;
    int pf_idx = -1, pf_val = -1;  // prefetched POLICY_GIVEN action of AGV index pf_idx
    bool guard = false;
    unsigned seq = 0;
#ifdef AGV_LANES_PROF
    long long q_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long qc_ = clock64();
#define QP(k) { const long long n_ = clock64(); q_[k] += n_ - qc_; qc_ = n_; }
#else
#define QP(k)
#endif
    while (__any_sync(FULL, st != LS_DONE)) {
        bool progress = false;
        // nothing to do while every unfinished scene of this warp waits for its BFS: sleep instead of
        // running the (predicated-off) body, whose instructions would compete with the BFS workers
        if (!__any_sync(FULL, st == LS_START || st == LS_HAVE || (st == LS_WAIT && vs[lane] == 3u))) {
#ifdef AGV_LANES_PROF
            iters_++; idle_++;
#endif
            __nanosleep(100);
            continue;
        }
        // ---- A: answers of the BFS workers
        if (st == LS_WAIT && vs[lane] == 3u) {
            __threadfence_block();
            const unsigned r = *reinterpret_cast<volatile unsigned *>(&sh->bfs_res[lane]);
            vs[lane] = 0u;
            action = (int)(r & 0xFFu); plen = (int)(r >> 8);
            if (guard) { ws.occ_clear(0, 0); guard = false; }
            st = LS_HAVE;
        }
        // record batch of this iteration goes to buffer seq & 1: the batch before the previous
        // one must be out
        QP(0)
        if (want_obs || want_next) {
            while ((int)(seq - *vdone) > 1) __nanosleep(20);
        }
        QP(1)
        const int buf = seq & 1;
        int nrec = 0;
        // ---- B: finish the turn of the lanes that have their action
        // (Explorer.check_action, Explorer.py:151-208; Scene.py:155-158)
        const bool fin = st == LS_HAVE;
        if (fin) {  // `a` is still the AGV loaded when this turn started
            progress = true;
            const bool guided = (t.policy == POLICY_GUIDED) || (t.policy == POLICY_GIVEN && given < 0);
            const bool pol_a = !guided && t.policy == POLICY_ASTAR;
            if (!guided && !pol_a && action > A_STOP) { ws.cnt_bad++; action = A_STOP; }
            const int dx = (action == A_RIGHT) - (action == A_LEFT);
            const int dy = (action == A_DOWN) - (action == A_UP);
            const int nx = a.x + dx, ny = a.y + dy;
            int r = 0, end = 0;
            const bool oob = nx < 1 || ny < 1 || nx > c.W || ny > c.H;
            if (oob) { r = -1; end = 1; }
            else {
                const int code = ws.cell_code(nx - 1, ny - 1);
                const bool at_target = (nx == a.tx) && (ny == a.ty);
                if (code == 1 && a.loaded && !at_target) { r = -1; end = 1; }
                else if (code == 2 && !at_target) { r = -1; end = 1; }
                else if (t.proto == PROTO_INTELLIGENT && ws.other_at_from(i, nx | (ny << 8), a.x | (a.y << 8))) { r = -1; end = 1; }
                else if (at_target) { r = 1; ws.continue_working(a); }
            }
            // Explorer.py:145-148: the move is applied iff not is_end
            if (t.proto != PROTO_INTELLIGENT && !guided && !pol_a) ws.overlap = true;
            if (!end && (dx | dy)) {
                const bool still = ws.other_at_from(i, a.x | (a.y << 8), a.x | (a.y << 8));
                ws.occ_move(a.x - 1, a.y - 1, nx - 1, ny - 1, !still);
                a.x = nx; a.y = ny;
            }
            ws.store_agv(i, a);
            if (t.proto == PROTO_INTELLIGENT) {
                if (ws.finished || (long long)ws.tick >= c.max_steps) end = 1;  // Scene.py:155
                if (end) ws.over = 1;
            } else {
                end = 0;  // is_end is ignored in A_star / manual mode (Scene.py:159-162)
            }
            ws.acted_mask |= (1ull << i);
            ws.cnt_turns++;
            ws.cnt_ended += end;
            const size_t rec = (size_t)ws.e * N + i;
            if (t.act_out) t.act_out[rec] = (int8_t)action;
            if (t.done) t.done[rec] = (uint8_t)end;
            if (t.reward) t.reward[rec] = (float)r;
            if (t.plen) t.plen[rec] = (uint16_t)plen;
        }
        QP(2)
        if (want_next) {  // store_info's create_state on the snapshot after the move (Scene.py:156)
            const unsigned pend = __ballot_sync(FULL, fin);
            if (fin) {
                bool shared = false;
                if (ws.overlap) shared = ws.other_at(i, a.x | (a.y << 8));
                const uint64_t wm = lane_window7<NW>(ws.tab + a.loaded * ws.HN, ws.occ, c.H, a.x - 1, a.y - 1, a.tx - 1,
                                                     a.ty - 1, shared);
                const int tpos = (max(-3, min(3, a.ty - a.y)) + 3) * 7 + max(-3, min(3, a.tx - a.x)) + 3;
                sh->enc_req[which][buf][nrec + __popc(pend & ((1u << lane) - 1u))] =
                    make_uint4((unsigned)wm, (unsigned)(wm >> 32), (unsigned)tpos | 256u, (unsigned)(ws.e * N + i));
            }
            nrec += __popc(pend);
        }
        if (fin) { i++; st = LS_START; }
        // ---- C: start the next turn (Scene.py:124-134 skip rules first)
        if (st == LS_START) {
            for (;;) {
                if (i >= n_loop || ws.over) { st = LS_DONE; break; }
                if (ws.finished) { ws.over = 1; ws.cnt_ended++; st = LS_DONE; break; }  // Scene.py:129-130
                if (!((ws.meta_s[i] >> 3) & 1)) break;                                   // idle AGVs are skipped
                i++;
            }
        }
        const bool act = st == LS_START;
        QP(3)
        if (act) {
            progress = true;
            a = ws.load_agv(i);
            given = -1;
            if (t.policy == POLICY_GIVEN && t.action) {
                // the action of the next AGV index was requested a turn ago (a global load here would
                // put ~500 cycles of latency on every turn of the scene's chain)
                given = (pf_idx == i) ? pf_val : (int)t.action[(size_t)ws.e * N + i];
                pf_idx = i + 1;
                pf_val = (i + 1 < N) ? (int)t.action[(size_t)ws.e * N + i + 1] : -1;
            }
        }
        if (want_obs) {
            const unsigned pend = __ballot_sync(FULL, act);
            if (act) {
                bool shared = false;
                if (ws.overlap) shared = ws.other_at(i, a.x | (a.y << 8));
                const uint64_t wm = lane_window7<NW>(ws.tab + a.loaded * ws.HN, ws.occ, c.H, a.x - 1, a.y - 1, a.tx - 1,
                                                     a.ty - 1, shared);
                const int tpos = (max(-3, min(3, a.ty - a.y)) + 3) * 7 + max(-3, min(3, a.tx - a.x)) + 3;
                sh->enc_req[which][buf][nrec + __popc(pend & ((1u << lane) - 1u))] =
                    make_uint4((unsigned)wm, (unsigned)(wm >> 32), (unsigned)tpos, (unsigned)(ws.e * N + i));
            }
            nrec += __popc(pend);
        }
        QP(4)
        if (__any_sync(FULL, act)) {
            const bool guided = (t.policy == POLICY_GUIDED) || (t.policy == POLICY_GIVEN && given < 0);
            const bool pol_a = !guided && t.policy == POLICY_ASTAR;
            bool need = act && (guided || pol_a);
            if (act && !need) { action = given; plen = 0; st = LS_HAVE; }
            const int sx = a.x - 1, sy = a.y - 1, tx = a.tx - 1, ty = a.ty - 1;
            if (need && sx == tx && sy == ty) { action = A_STOP; plen = 0; st = LS_HAVE; need = false; }
            // matrix (a) also blocks (1,1) while uncreated AGVs park there (Explorer.py:274-283)
            if (need && pol_a && N > 1 && ws.n_created < N && !ws.occ_test(0, 0)) { ws.occ_set(0, 0); guard = true; }
            const uint64_t om = guided ? ~0ull : om_a;
            int l = 0;
            const int ma = lane_monotone<NW>(need, ws.tab, ws.HN, om ? ws.occ : zr, om ? ws.occm : zr, a.loaded, sx, sy, tx,
                                             ty, l);
            if (need) {
                if (ma >= 0) {
                    action = ma; plen = l; st = LS_HAVE;
                    if (guard) { ws.occ_clear(0, 0); guard = false; }
                } else {  // level-synchronous BFS by a worker warp; this scene waits, the others go on
                    *reinterpret_cast<volatile unsigned *>(&sh->bfs_req[lane]) =
                        (unsigned)sx | ((unsigned)sy << 7) | ((unsigned)tx << 14) | ((unsigned)ty << 21) |
                        ((unsigned)a.loaded << 28) | (om ? (1u << 29) : 0u);
                    __threadfence_block();
                    vs[lane] = 1u;
                    st = LS_WAIT;
                    n_bfs++;
                }
            }
        }
        QP(5)
        // ---- hand the iteration's records to the record writer
        if (nrec > 0) {
            if (lane == 0) *reinterpret_cast<volatile int *>(&sh->enc_n[which][buf]) = nrec;
            __threadfence_block();
            __syncwarp();
            seq++;
            if (lane == 0) *vseq = seq;
        }
        QP(6)
#ifdef AGV_LANES_PROF
        iters_++;
        if (!__any_sync(FULL, progress)) idle_++;
#endif
    }
#ifdef AGV_LANES_PROF
#ifndef AGV_PROF_WHICH
#define AGV_PROF_WHICH 1
#endif
    const unsigned mb_ = __reduce_max_sync(FULL, n_bfs);
    if (lane == 0 && which == AGV_PROF_WHICH) {
        atomicMax(c.counters + 19, (unsigned long long)mb_);
        const long long d_ = clock64() - p0_;
        atomicAdd(c.counters + 3 + 0, (unsigned long long)d_); atomicAdd(c.counters + 3 + 1, (unsigned long long)iters_);
        atomicAdd(c.counters + 3 + 2, (unsigned long long)idle_); atomicAdd(c.counters + 15, 1ull);
        atomicMax(c.counters + 16, (unsigned long long)d_); atomicMax(c.counters + 17, (unsigned long long)iters_);
        atomicMax(c.counters + 18, (unsigned long long)idle_);
        for (int k_ = 0; k_ < 7; k_++) atomicAdd(c.counters + 20 + k_, (unsigned long long)q_[k_]);
        { int b_ = (int)(d_ / 100000); if (b_ > 11) b_ = 11; atomicAdd(&c.counters[32 + b_], 1ull); }  // histogram of loop cycles, 100 k buckets
    }
#endif
    __threadfence_block();
    __syncwarp();
    if (lane == 0 && atomicAdd(&sh->done_warps, 1u) == 1u) *reinterpret_cast<volatile int *>(&sh->exit_flag) = 1;
    if (live && !ws.over) ws.spawn();
    if (ws.has) { ws.regs_to_hdr(); ws.hdr_w[6] = n_bfs; }
    {
        const unsigned ct = __reduce_add_sync(FULL, ws.cnt_turns), ce = __reduce_add_sync(FULL, ws.cnt_ended),
                       cb = __reduce_add_sync(FULL, ws.cnt_bad);
        if (lane == 0) {
            if (ct) atomicAdd(c.counters + 0, (unsigned long long)ct);
            if (ce) atomicAdd(c.counters + 1, (unsigned long long)ce);
            if (cb) atomicAdd(c.counters + 2, (unsigned long long)cb);
        }
    }
    __threadfence_block();
    named_bar_sync(3, 64);  // the headers of both scene warps are back in the blocks
    if (which == 0) ws.store_blocks(e0);
}

// kernel body: role by warp index
template <int NW, int RPL, int OBS, int SPW, int NB>
__device__ __forceinline__ void async_cta(const DevCfg &c, const LaneArgs &t, uint8_t *smem) {
    AsyncShared *sh = reinterpret_cast<AsyncShared *>(smem + t.off_queue);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        sh->bfs_state[lane] = 0u;
        if (lane == 0) {
            sh->exit_flag = 0; sh->done_warps = 0u;
            sh->enc_seq[0] = sh->enc_seq[1] = 0u; sh->enc_done[0] = sh->enc_done[1] = 0u;
        }
    }
    __syncthreads();
    if (warp < 2) tick_async<NW, RPL, OBS, SPW>(c, t, smem, warp);
    else if (warp < 2 + NB) async_bfs_worker<NW, RPL, SPW>(c, t, smem);
    else async_record_writer(t, smem);
}

}  // namespace agv
