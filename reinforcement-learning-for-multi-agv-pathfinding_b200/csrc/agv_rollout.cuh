// agv_rollout.cuh -- persistent multi-tick rollout: every scene runs n_ticks passes of
// Scene.run_mode's loop body (Scene.py:111-164) inside ONE kernel launch.
//
// In A_star / Expert mode and in the A*-guided branch of AG-DQN the scenes are independent across
// ticks as well as inside one, so the per-tick launch of agv_tick is only a barrier: it makes every
// tick as long as its most congested scene.  Here a lane keeps its scene for all n_ticks (its own
// tick counter next to its own turn cursor, same decoupled-lane state machine as agv_async.cuh),
// outputs go to [n_ticks, E, N, ...] rings, and CTAs are persistent: when a scene has finished its
// n_ticks the slot is handed to a BFS worker warp, which writes the scene block back, takes the next
// scene from a global queue (longest first -- ordered by the turns + BFS searches of the scene's last tick)
// and loads it.  Long scenes start first and short ones fill in behind them.
//
// Warp roles in a CTA:  warps 0..NB-1    BFS workers: level-synchronous BFS requests + slot swaps
//                       warp NB          record writer: unpacks the per-turn records of a finished
//                                        scene-tick -- act / reward / done / plen with coalesced stores,
//                                        the packed 49-bit windows into [3,7,7] bf16 records with
//                                        16-byte stores (8 records = 147 uint4, always aligned)
//                       warps NB+1 / +2  scene warps: the slots whose last tick was heavy / light
//                                        (a slot migrates between them at tick boundaries)
// Results are bit-identical to n_ticks calls of agv_tick (tests/test_gpu_rollout.py).
#pragma once
#include "agv_async.cuh"
#include <cstddef>

namespace agv {

constexpr int ROLL_RING = 32;  // tick-end record requests in flight per scene warp

// Poll intervals of the idle warps.  Every poll is ~25 issued instructions on the half-rate ALU pipe the
// scene warps' turn chains live on: with 400 / 100 ns everywhere (agv_async.cuh) the polls were half of
// the 5.2e9 instructions of a 16-tick c3 rollout (ncu: smsp__issue_active 50 %, ALU pipe 53 %).  The
// record writer is never on a scene's critical path; a BFS takes >= 3 us, so 1 us of wake-up is cheap.
#ifndef AGV_ROLL_WORKER_NS
#define AGV_ROLL_WORKER_NS 1000
#endif
#ifndef AGV_ROLL_WORKER_FAST_NS
#define AGV_ROLL_WORKER_FAST_NS 250
#endif
#ifndef AGV_ROLL_WRITER_NS
#define AGV_ROLL_WRITER_NS 4000
#endif
#ifndef AGV_ROLL_IDLE_NS
#define AGV_ROLL_IDLE_NS 2000   // scene warp without any scene (waiting for a hand-over)
#endif
#ifndef AGV_ROLL_WAIT_NS
#define AGV_ROLL_WAIT_NS 200    // scene warp whose scenes all wait for BFS answers
#endif

struct RollArgs {
    LaneArgs la;              // proto / policy / obs kind, tick-0 base pointers, shared-memory plan
    int K;                    // ticks per scene
    int vec_ok;               // N % 8 == 0 and 16-byte aligned rings: vector-store path
    int heavy_bfs, light_bfs; // searches per tick from which a scene runs on the heavy warp / below which it returns
    // per-turn records of the scene lanes, unpacked by the record writer when the scene-tick is over:
    // scr_turn [K*E*N] {wm | tpos << 49, reward f32 | plen << 32 | action << 48 | done << 56}
    // (wm = 49-bit window of matrix (b) before the move), scr_next [K*E*N] wm | tpos << 49 after the move
    ulonglong2 *scr_turn;
    unsigned long long *scr_next;
    const unsigned *order;    // [E] scene indices, longest first
    unsigned *qctr;           // next entry of `order` to hand out
    unsigned long long *stat; // pinned host memory, written by the order kernel: {BFS searches, turns} of every
                              // scene's last tick -- the load figure the NEXT launch picks its shape by
    unsigned long long *prof; // -DAGV_ROLL_PROF builds: [E][8] per-scene timeline (agv_rollout_profile)
};

#ifdef AGV_ROLL_PROF
__device__ __forceinline__ unsigned long long roll_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define RPROF(...) __VA_ARGS__
#else
#define RPROF(...)
#endif

struct RollShared {
    unsigned bfs_state[32];   // per slot: 0 idle, 1 BFS posted, 2 taken, 3 answered, 4 swap posted, 5 swap taken
    unsigned bfs_req[32];
    unsigned bfs_res[32];
    unsigned ctl[32];         // 0 / 1: running on scene warp 0 / 1; 2 / 3: handed to scene warp 0 / 1; 4: exhausted
    int slot_e[32];           // scene held by the slot (-1: none)
    int slot_k[32];           // ticks of the rollout this scene has done
    unsigned slot_fl[32];     // bit 0: two created AGVs may share a cell (LaneScene::overlap)
    int exit_flag;
    unsigned done_warps;
    unsigned done_slots;      // slots that found the scene queue empty
    unsigned heavy_live;      // scenes the heavy warp is running: the workers poll faster while there are any
    unsigned enc_head[2], enc_tail[2];
    uint4 enc_ring[2][ROLL_RING];          // {first record of the scene-tick, acted lo, acted hi, -}
    unsigned long long recbits[64][3];     // writer staging: the 147-bit strings of one scene-tick's records
    uint4 lut[256];                        // 8 bits -> 8 bf16 (0.0 / 1.0): the writer's expansion is one 16-byte load
};
static_assert(offsetof(RollShared, enc_ring) % 16 == 0 && offsetof(RollShared, lut) % 16 == 0, "RollShared layout");

constexpr int LS_IDLE = 4, LS_NEWTICK = 5, LS_TICKEND = 6;
// A scene runs on the heavy warp from AGV_ROLL_HEAVY_BFS searches in its last tick and returns to the light warp
// below AGV_ROLL_LIGHT_BFS.  The heavy warp buys its few scenes short BFS round trips, but an iteration costs a
// warp the same for one active lane as for 32, so every scene on it is paid for by the light warps -- measured
// at c3 (guided, AGV-steps/s; 2 workers x 3 CTAs per SM): 16 / 8: 1.19e9, 24 / 12: 1.24e9, 32 / 16: 1.27e9, 40 / 30:
// 1.29e9, 64 / 32: 1.24e9, none: 1.24e9; with the final shape (3 workers x 2 CTAs, 96 registers): 32 / 24: 1.415e9
// (e2e 1.29e9), 40 / 30: 1.40e9 (1.25e9), 64 / 48 and none: 1.33e9.  In the congested modes (heavy launch shape) the
// searches bind and 16 / 8 is better: c3 A_star control mode 6.6e8 against 6.3e8 with 40 / 30.  On 128 x 128 maps the
// heavy warp never pays (launch_shape in agv_rollout.cu sets RollArgs::heavy_bfs / light_bfs per class and shape).
#ifndef AGV_ROLL_HEAVY_BFS
#define AGV_ROLL_HEAVY_BFS 32
#endif
#ifndef AGV_ROLL_LIGHT_BFS
#define AGV_ROLL_LIGHT_BFS 24
#endif
#ifndef AGV_ROLL_SLOW_EVERY
#define AGV_ROLL_SLOW_EVERY 1   // power of two: the light scene warp runs its slow section every n-th iteration
                                // (measured at c3: 1 -> 1.125e9, 2 -> 1.12e9, 4 -> 1.10e9, 8 -> 1.0e9 AGV-steps/s: what the
                                // fast iterations save, the scenes waiting at their tick boundary lose)
#endif

// lane_monotone (agv_lanes.cuh) reshaped for the latency of ONE warp (nothing else runs on the scene warp's
// SM sub-partition, so every instruction and every exposed load latency of the row loop is on the scene's
// turn chain; the plain loop measured ~100 cycles per row, 6 k of the 14 k cycles of a warp iteration):
//  * the loop covers only the rows strictly between the target's and the start's; a lane that is through
//    re-reads its last row, which is idempotent (fill(fill(x)) == fill(x)), so no per-row snapshot selects;
//    the start row is applied once after the loop (its loads are issued before the loop);
//  * the row loads of the next pair of rows are issued while the current pair runs through the fill chain.
// 13 instead of 23 instructions per row.  Same result as lane_monotone for every input.
template <int NW>
__device__ __forceinline__ int lane_monotone_pf(bool need, const uint64_t *tab, int HN, const uint64_t *occ,
                                                const uint64_t *occm, int loaded, int sx, int sy, int tx, int ty,
                                                int &len) {
    constexpr int WB = 64 * NW;
    const int dxs = (tx > sx) - (tx < sx), dys = (ty > sy) - (ty < sy);
    const bool mirror = tx > sx;
    const int sxm = mirror ? WB - 1 - sx : sx, txm = mirror ? WB - 1 - tx : tx;  // txm <= sxm
    const bool flip = mirror && !mirror_copy<NW>();
    const uint64_t *bp = tab + ((mirror && mirror_copy<NW>() ? 2 : 0) + loaded) * HN + ty * NW;
    const uint64_t *op = (mirror && mirror_copy<NW>() ? occm : occ) + ty * NW;
    const int step = -dys * NW;
    const Row<NW> box = rrange<NW>(txm, sxm);
    Row<NW> reach = rbit<NW>(txm);
    const int n = need ? abs(ty - sy) : 0;  // rows after the target's; row n is the start's
    const int nm = max(n - 1, 0);           // rows strictly between
    if (!need) { bp = tab; op = tab; }
    Row<NW> b0 = ldrow<NW>(bp), o0 = ldrow<NW>(op);
    Row<NW> bs = ldrow<NW>(bp + n * step), os = ldrow<NW>(op + n * step);  // the start's row
    Row<NW> nb[2], no[2];
#pragma unroll
    for (int q = 0; q < 2; q++) { const int k = min(1 + q, nm) * step; nb[q] = ldrow<NW>(bp + k); no[q] = ldrow<NW>(op + k); }
    {   // the target's row: the target cell itself is valid unless occupied
        if (!mirror_copy<NW>() && flip) { b0 = rmirror(b0); o0 = rmirror(o0); }
        b0 = ror(b0, reach);
#pragma unroll
        for (int i = 0; i < NW; i++) b0.w[i] = b0.w[i] & ~o0.w[i] & box.w[i];
        reach = rfill_up(rand_(reach, b0), b0);
    }
    const Row<NW> r0 = reach;
    const int nmax = __reduce_max_sync(FULL, nm);
#pragma unroll 1
    for (int it = 1; it <= nmax; it += 2) {
        Row<NW> cb[2], co[2];
#pragma unroll
        for (int q = 0; q < 2; q++) { cb[q] = nb[q]; co[q] = no[q]; }
#pragma unroll
        for (int q = 0; q < 2; q++) { const int k = min(it + 2 + q, nm) * step; nb[q] = ldrow<NW>(bp + k); no[q] = ldrow<NW>(op + k); }
#pragma unroll
        for (int q = 0; q < 2; q++) {
            Row<NW> b = cb[q], o = co[q];
            if (!mirror_copy<NW>() && flip) { b = rmirror(b); o = rmirror(o); }
#pragma unroll
            for (int i = 0; i < NW; i++) b.w[i] = b.w[i] & ~o.w[i] & box.w[i];
            reach = rfill_up(rand_(reach, b), b);
        }
    }
    if (nm == 0) reach = r0;  // (the loop re-read the target's row without its target bit: discard)
    // reach = row before the start's (n >= 1) or the target's own row (n == 0)
    Row<NW> rfin = reach, pfin = rzero<NW>();
    if (n >= 1) {
        pfin = reach;
        if (!mirror_copy<NW>() && flip) { bs = rmirror(bs); os = rmirror(os); }
#pragma unroll
        for (int i = 0; i < NW; i++) bs.w[i] = bs.w[i] & ~os.w[i] & box.w[i];
        rfin = rfill_up(rand_(reach, bs), bs);
    }
    len = 0;
    int act = -1;
    if (need) {
        // the start's own row may be empty (the start cell is occupied by the AGV itself, hence
        // invalid) while the vertical neighbour in the row before it is still on a path
        const bool okH = dxs != 0 && rtest(rfin, sxm - 1);
        const bool okV = dys != 0 && rtest(pfin, sxm);
        if (okH && dxs > 0) act = A_RIGHT;
        else if (okV && dys > 0) act = A_DOWN;
        else if (okV && dys < 0) act = A_UP;
        else if (okH && dxs < 0) act = A_LEFT;
        if (act >= 0) len = abs(tx - sx) + abs(ty - sy);
    }
    return act;
}

// (A bidirectional variant -- target-side chain and two start-side chains on the mirrored copies, meeting
// in the middle row: half the dependent depth, 3 independent chains -- was parity-green but slower, 6.8 k
// vs 4.9 k cycles per warp iteration: 45 instead of 17 instructions per row step, and the warp was not
// dependency-bound but losing the issue arbitration, see rollout_cta.)

// ------------------------------------------------------------------ slot <-> HBM (one warp, coalesced)
template <int NW>
__device__ __forceinline__ void roll_store_slot(const DevCfg &c, const LaneArgs &t, uint8_t *smem, int slot, int e,
                                                int lane) {
    const int wps = c.env_stride >> 2;
    uint32_t *dst = reinterpret_cast<uint32_t *>(c.state + (size_t)e * c.env_stride);
    const uint32_t *src = reinterpret_cast<const uint32_t *>(smem + t.off_blk + (size_t)slot * t.blk_stride);
    for (int w = lane; w < wps; w += 32) dst[w] = src[w];
}

// loads scene e into `slot`, builds its occupancy rows and hands it to the scene warp that fits its weight
template <int NW>
__device__ __forceinline__ void roll_load_slot(const DevCfg &c, const LaneArgs &t, uint8_t *smem, RollShared *sh,
                                               int slot, int e, int lane, int heavy_bfs) {
    constexpr int WB = 64 * NW;
    const int wps = c.env_stride >> 2, HN = c.H * NW;
    const uint32_t *src = reinterpret_cast<const uint32_t *>(c.state + (size_t)e * c.env_stride);
    uint32_t *blk = reinterpret_cast<uint32_t *>(smem + t.off_blk + (size_t)slot * t.blk_stride);
    {
        uint32_t v[4];
#pragma unroll
        for (int q = 0; q < 4; q++) { const int w = lane + 32 * q; v[q] = w < wps ? __ldg(src + w) : 0u; }
#pragma unroll
        for (int q = 0; q < 4; q++) { const int w = lane + 32 * q; if (w < wps) blk[w] = v[q]; }
    }
    uint64_t *occ = reinterpret_cast<uint64_t *>(smem + t.off_occ) + (size_t)slot * t.occ_stride;
    uint64_t *occm = reinterpret_cast<uint64_t *>(smem + t.off_occm) + (size_t)slot * t.occ_stride;
    for (int q = lane; q < HN; q += 32) { occ[q] = 0ull; if (mirror_copy<NW>()) occm[q] = 0ull; }
    __syncwarp();
    const int n_created = (int)(blk[3] & 0xFFFFu);
    const uint16_t *pos = reinterpret_cast<const uint16_t *>(reinterpret_cast<const uint8_t *>(blk) + c.off_pos);
    for (int j = lane; j < n_created; j += 32) {
        const int p = pos[j], x0 = (p & 255) - 1, y0 = (p >> 8) - 1;
        atomicOr(reinterpret_cast<unsigned *>(occ) + y0 * NW * 2 + (x0 >> 5), 1u << (x0 & 31));
        if (mirror_copy<NW>()) {
            const int xm = WB - 1 - x0;
            atomicOr(reinterpret_cast<unsigned *>(occm) + y0 * NW * 2 + (xm >> 5), 1u << (xm & 31));
        }
    }
    __syncwarp();
    int bits = 0;
    for (int q = lane; q < HN; q += 32) bits += __popcll(occ[q]);
    bits = __reduce_add_sync(FULL, bits);
    if (lane == 0) {
        sh->slot_e[slot] = e;
        sh->slot_k[slot] = 0;
        sh->slot_fl[slot] = bits != n_created ? 1u : 0u;
        const unsigned heavy = (blk[6] & 0xFFFFu) >= (unsigned)heavy_bfs ? 1u : 0u;
        __threadfence_block();
        *reinterpret_cast<volatile unsigned *>(&sh->ctl[slot]) = 2u + heavy;
    }
    __syncwarp();
}

// ------------------------------------------------------------------ BFS worker (+ slot swaps)
template <int NW, int RPL, int SPW>
__device__ __forceinline__ void roll_worker(const DevCfg &c, const RollArgs &ra, uint8_t *smem) {
    const LaneArgs &t = ra.la;
    const int lane = threadIdx.x & 31, HN = c.H * NW;
    RollShared *sh = reinterpret_cast<RollShared *>(smem + t.off_queue);
    volatile unsigned *vs = sh->bfs_state;
    volatile int *vexit = &sh->exit_flag;
    const uint64_t *tab = reinterpret_cast<const uint64_t *>(smem + t.off_tab);
    const uint64_t *occ0 = reinterpret_cast<const uint64_t *>(smem + t.off_occ);
    for (;;) {
        const unsigned s = vs[lane];
        unsigned m = __ballot_sync(FULL, s == 1u || s == 4u);
        if (!m) {
            if (*vexit) return;
            // BFS round trips are the critical path of the heavy scenes only: quick wake-ups while the CTA has any.
            // (These polls are 12 % of the kernel's instructions.  A one-word poll -- a counter of posted requests kept
            // by the scene warps -- halves them and was measured 1.5 % SLOWER: the ballot + atomic it adds to every
            // scene-warp iteration costs more than the idle warps' instructions, which nothing is waiting for.)
            __nanosleep(*reinterpret_cast<volatile unsigned *>(&sh->heavy_live) ? AGV_ROLL_WORKER_FAST_NS : AGV_ROLL_WORKER_NS);
            continue;
        }
        int j = -1;
        unsigned kind = 0;
        while (m) {
            const int k = __ffs(m) - 1;
            m &= m - 1;
            const unsigned sk = __shfl_sync(FULL, s, k);
            int ok = 0;
            if (lane == 0) ok = atomicCAS(&sh->bfs_state[k], sk, sk + 1u) == sk;
            ok = __shfl_sync(FULL, ok, 0);
            if (ok) { j = k; kind = sk; break; }
        }
        if (j < 0) continue;
        __threadfence_block();
        if (kind == 1u) {
            const unsigned pk = *reinterpret_cast<volatile unsigned *>(&sh->bfs_req[j]);
            int bl = 0;
            const int ba = coop_bfs<NW, RPL>(pk, tab, HN, occ0 + (size_t)j * t.occ_stride, c.H, c.W, lane, bl);
            if (lane == 0) {
                *reinterpret_cast<volatile unsigned *>(&sh->bfs_res[j]) = (unsigned)ba | ((unsigned)bl << 8);
                __threadfence_block();
                vs[j] = 3u;
            }
        } else {
            // the scene of slot j has done its n_ticks: write it back, take the next one from the queue
            const int old = *reinterpret_cast<volatile int *>(&sh->slot_e[j]);
            if (old >= 0) roll_store_slot<NW>(c, t, smem, j, old, lane);
            unsigned idx = 0;
            if (lane == 0) idx = atomicAdd(ra.qctr, 1u);
            idx = __shfl_sync(FULL, idx, 0);
            __syncwarp();
            if (lane == 0) vs[j] = 0u;
            if (idx < (unsigned)c.E) {
                roll_load_slot<NW>(c, t, smem, sh, j, (int)__ldg(ra.order + idx), lane, ra.heavy_bfs);
            } else if (lane == 0) {
                sh->slot_e[j] = -1;
                *reinterpret_cast<volatile unsigned *>(&sh->ctl[j]) = 4u;
                __threadfence_block();
                atomicAdd(&sh->done_slots, 1u);
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------ record writer
// One finished scene-tick: the packed windows of its AGVs (scr, written by the scene lane turn by
// turn) become [3,7,7] bf16 records.  A record is a 147-bit string (49 bits per channel); 8 records
// are 1176 bits = 147 x 8 bits, and 8 bits expand to one uint4 of bf16 -- so with N % 8 == 0 every
// store is a full, aligned 16-byte vector.  Records of AGVs without a turn inside a written group
// are zero-filled (the consumer masks with act_out >= 0, as for agv_tick).
// (scr: packed windows, `stride` 64-bit words apart)
static __device__ __noinline__ void roll_expand(int vec_ok, RollShared *sh, uint16_t *dst, const unsigned long long *scr,
                                         int stride, unsigned rec_base, uint64_t acted, int lane) {
    if (!acted) return;
    const int n_hi = 64 - __clzll((long long)acted);
    const int G = (n_hi + 7) >> 3;
    if (!vec_ok) {
        for (int r = 0; r < n_hi; r++) {
            if (!((acted >> r) & 1ull)) continue;
            const unsigned long long m = __ldcg(scr + ((size_t)rec_base + r) * stride);
            expand_clip7(dst + ((size_t)rec_base + r) * 147, m & ((1ull << 49) - 1ull), (int)((m >> 49) & 63ull), lane);
        }
        return;
    }
#pragma unroll
    for (int q = 0; q < 2; q++) {
        const int r = lane + 32 * q;
        if (r < 8 * G) {
            const bool valid = (acted >> r) & 1ull;
            const unsigned long long m = valid ? __ldcg(scr + ((size_t)rec_base + r) * stride) : 0ull;
            const unsigned long long wm = m & ((1ull << 49) - 1ull);
            const int tp = (int)((m >> 49) & 63ull);
            unsigned long long w0 = 0ull, w1 = 0ull, w2 = 0ull;
            if (valid) {
                w0 = (1ull << 24) | (tp < 15 ? (1ull << (49 + tp)) : 0ull);
                w1 = (tp >= 15 ? (1ull << (tp - 15)) : 0ull) | (wm << 34);
                w2 = wm >> 30;
            }
            sh->recbits[r][0] = w0; sh->recbits[r][1] = w1; sh->recbits[r][2] = w2;
        }
    }
    __syncwarp();
    uint4 *d4 = reinterpret_cast<uint4 *>(dst + (size_t)rec_base * 147);
    // up to the end of the last record with a turn; the last vector may reach into the first <= 7 elements of
    // the record after it, which are zero in every record (its first-channel cells 0..6; the centre is 24)
    const int nv = (147 * n_hi + 7) >> 3;
    for (int v = lane; v < nv; v += 32) {
        const unsigned bit = 8u * (unsigned)v;
        const unsigned r = bit / 147u, off = bit - r * 147u;
        const unsigned wi = off >> 6, sft = off & 63u;
        unsigned long long lo = sh->recbits[r][wi] >> sft;
        if (sft && wi < 2u) lo |= sh->recbits[r][wi + 1] << (64u - sft);
        // (bits past the end of record r are bits 0..6 of record r+1's first channel: always zero)
        d4[v] = sh->lut[(unsigned)lo & 0xFFu];
    }
    __syncwarp();
}

// act / done / reward / plen of a finished scene-tick: the scene lane left them packed in the second word
// of its per-turn records (one 16-byte store per turn instead of five scattered ones on the turn chain);
// here they go out to the [n_ticks, E, N] rings, one coalesced store per array
__device__ __forceinline__ void roll_unpack(const RollArgs &ra, unsigned rec_base, uint64_t acted, int lane) {
    const LaneArgs &t = ra.la;
    const unsigned long long *scr = reinterpret_cast<const unsigned long long *>(ra.scr_turn);
#pragma unroll
    for (int q = 0; q < 2; q++) {
        const int r = lane + 32 * q;
        if ((acted >> r) & 1ull) {
            const size_t rec = (size_t)rec_base + r;
            const unsigned long long hi = __ldcg(scr + rec * 2 + 1);
            if (t.reward) t.reward[rec] = __uint_as_float((unsigned)hi);
            if (t.plen) t.plen[rec] = (uint16_t)(hi >> 32);
            if (t.act_out) t.act_out[rec] = (int8_t)(hi >> 48);
            if (t.done) t.done[rec] = (uint8_t)(hi >> 56);
        }
    }
}

__device__ __forceinline__ void roll_writer(const RollArgs &ra, uint8_t *smem) {
    const LaneArgs &t = ra.la;
    const int lane = threadIdx.x & 31;
    RollShared *sh = reinterpret_cast<RollShared *>(smem + t.off_queue);
    volatile unsigned *vhead = sh->enc_head, *vtail = sh->enc_tail;
    volatile int *vexit = &sh->exit_flag;
    unsigned mine[2] = {0u, 0u};
    for (;;) {
        bool any = false;
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const unsigned h = vhead[q];
            if (h == mine[q]) continue;
            any = true;
            __threadfence_block();
            for (; mine[q] != h; mine[q]++) {
                const uint4 rq = sh->enc_ring[q][mine[q] % ROLL_RING];
                const uint64_t acted = (uint64_t)rq.y | ((uint64_t)rq.z << 32);
                roll_unpack(ra, rq.x, acted, lane);
                if (t.obs) roll_expand(ra.vec_ok, sh, t.obs, reinterpret_cast<const unsigned long long *>(ra.scr_turn), 2, rq.x, acted, lane);
                if (t.next_obs) roll_expand(ra.vec_ok, sh, t.next_obs, ra.scr_next, 1, rq.x, acted, lane);
            }
            __threadfence_block();
            __syncwarp();
            if (lane == 0) vtail[q] = mine[q];
        }
        if (any) continue;
        if (*vexit) {  // the last requests are posted before the exit flag
            if (vhead[0] == mine[0] && vhead[1] == mine[1]) return;
            continue;
        }
        __nanosleep(AGV_ROLL_WRITER_NS);
    }
}

// ------------------------------------------------------------------ scene warp
// MODE fixes protocol and policy at compile time: 1 = "intelligent" + A* guidance (AG-DQN's guided branch,
// the headline), 2 = A_star control mode (Expert rollouts), 3 = "intelligent" + per-AGV actions with the
// guidance fallback (the epsilon-gate: what the host-buffer entry points run), 0 = whatever the call says.
// The specialised kernels carry a third less code: a scene-warp iteration streams its whole body through
// the 6 KB L0 / 32 KB L1.5 instruction caches once per turn, and instruction fetch was the top stall of the
// generic body (ncu source view: stall_no_inst 37 % of the samples outside the row-DP loop).
//
// One loop iteration = a FAST section (finish the turns that have their action, start the next ones: window,
// row DP, BFS request) and, every 4th iteration only, a SLOW section (hand-overs, tick boundaries, spawn,
// record requests, BFS answers, exit test).  The 32 scenes of a warp are at different points of their ticks,
// so tick-boundary code would otherwise run -- for one or two lanes -- in nearly every iteration.  The heavy
// warp (few scenes, all waiting for BFS answers most of the time) runs the slow section every iteration.
template <int NW, int RPL, int OBS, int SPW, int MODE>
__device__ __forceinline__ void roll_scene_warp(const DevCfg &c, const RollArgs &ra, uint8_t *smem, const int which) {
    const LaneArgs &t = ra.la;
    const int lane = threadIdx.x & 31;
    LaneScene<NW, RPL, SPW, 0> ws(c, t, smem, lane, 0);
    RollShared *sh = reinterpret_cast<RollShared *>(smem + t.off_queue);
    volatile unsigned *vs = sh->bfs_state, *vctl = sh->ctl;
    volatile unsigned *vdone_slots = &sh->done_slots;
    volatile unsigned *vhead = &sh->enc_head[which], *vtail = &sh->enc_tail[which];
    const int N = c.N, K = ra.K;
    constexpr bool GIVEN = MODE == 0 || MODE == 3;   // per-AGV actions from the caller (negative = guidance)
    constexpr bool MAT_A = MODE == 0 || MODE == 2;   // searches on matrix (a) can occur ((1,1) guard while AGVs are uncreated)
    const int proto = (MODE == 1 || MODE == 3) ? PROTO_INTELLIGENT : (MODE == 2 ? PROTO_ASTAR : t.proto);
    const int policy = MODE == 1 ? POLICY_GUIDED : (MODE == 2 ? POLICY_ASTAR : (MODE == 3 ? POLICY_GIVEN : t.policy));
    const uint64_t om_a = N > 1 ? ~0ull : 0ull;
    const bool want_obs = OBS == OBS_CLIP && t.obs != nullptr, want_next = OBS == OBS_CLIP && t.next_obs != nullptr;
    const uint64_t *zr = reinterpret_cast<const uint64_t *>(smem + t.off_zero);
    ws.has = false; ws.e = 0; ws.tasks_e = c.tasks; ws.pf = true;
    ws.over = 1; ws.finished = 0; ws.n_created = 0; ws.tick = 0; ws.cursor = 0; ws.n_done = 0; ws.acted_mask = 0;

    int st = LS_IDLE, i = 0, n_loop = 0, k = 0;
    unsigned n_bfs = 0;     // BFS searches of the tick in progress ...
    unsigned n_turns = 0;   // ... and turns of the last finished tick: the scene's weight in the next launch's order
    unsigned rec0 = 0;      // (k * E + e) * N: first output record of the scene's tick in progress
    unsigned head = 0, it = 0;
    Agv a;
    a.x = a.y = a.tx = a.ty = 1; a.stage = a.loaded = a.idle = a.order = 0;
    int action = -1, plen = 0, given = -1;
    int pf_idx = -1, pf_val = -1;
    bool guard = false;
    unsigned long long wm_turn = 0ull;  // packed window of the turn in progress (obs before the move)
    RPROF(long long t_post = 0; long long q_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long qc_ = clock64(); long long iters_ = 0; long long rows_ = 0;)
#ifdef AGV_ROLL_PROF
#define RQ(k) { const long long n_ = clock64(); q_[k] += n_ - qc_; qc_ = n_; }
#else
#define RQ(k)
#endif

    for (;;) {
        const bool any_run = __any_sync(FULL, st == LS_START || st == LS_HAVE);
        if (which == 1 || (it & (unsigned)(AGV_ROLL_SLOW_EVERY - 1)) == 0u || !any_run) {
            // ================================================================ slow section
            if (*vdone_slots >= (unsigned)SPW) break;  // every slot of the CTA has found the queue empty
            const bool pick = st == LS_IDLE && vctl[lane] == 2u + (unsigned)which;
            const bool ans = st == LS_WAIT && vs[lane] == 3u;
            if (!any_run && !__any_sync(FULL, pick || ans || st == LS_NEWTICK || st == LS_TICKEND)) {
                if (__any_sync(FULL, st == LS_WAIT)) __nanosleep(AGV_ROLL_WAIT_NS);
                else __nanosleep(AGV_ROLL_IDLE_NS);
                RQ(0)
                continue;
            }
            if (which == 1) {
                const unsigned live = __ballot_sync(FULL, st != LS_IDLE);
                if (lane == 0) *reinterpret_cast<volatile unsigned *>(&sh->heavy_live) = live;
            }
            // ---- slots handed to this warp (freshly loaded by a worker, or migrated from the other scene warp)
            if (pick) {
                __threadfence_block();
                ws.hdr_to_regs();
                ws.e = *reinterpret_cast<volatile int *>(&sh->slot_e[lane]);
                ws.tasks_e = c.tasks + (size_t)ws.e * c.T;
                ws.prefetch_cursor_task();
                k = *reinterpret_cast<volatile int *>(&sh->slot_k[lane]);
                ws.overlap = (*reinterpret_cast<volatile unsigned *>(&sh->slot_fl[lane]) & 1u) != 0u;
                n_bfs = ws.hdr_w[6] & 0xFFFFu;
                n_turns = ws.hdr_w[6] >> 16;
                vctl[lane] = (unsigned)which;
                st = LS_NEWTICK;
                RPROF(if (ra.prof) { if (k == 0) ra.prof[(size_t)ws.e * 8 + 4] = roll_now(); atomicAdd(ra.prof + (size_t)ws.e * 8 + 5, 1ull); })
            }
            // ---- end of a tick: Scene.check_new_veh (Scene.py:164), then the finished scene-tick goes to the
            // record writer: {first record, who took a turn}
            const bool tick_end = st == LS_TICKEND;
            if (tick_end) {
                n_turns = (unsigned)__popcll(ws.acted_mask);
                if (!ws.over) ws.spawn();
                k++;
                st = LS_NEWTICK;
                RPROF(if (ra.prof) { atomicAdd(ra.prof + (size_t)ws.e * 8 + 1, (unsigned long long)n_bfs); atomicAdd(ra.prof + (size_t)ws.e * 8 + 2, (unsigned long long)__popcll(ws.acted_mask)); })
            }
            {
                const bool post = tick_end && ws.acted_mask != 0ull;
                const unsigned pend = __ballot_sync(FULL, post);
                if (pend) {
                    const int n = __popc(pend);
                    while ((int)(head + (unsigned)n - *vtail) > ROLL_RING) __nanosleep(40);
                    if (post)
                        sh->enc_ring[which][(head + (unsigned)__popc(pend & ((1u << lane) - 1u))) % ROLL_RING] =
                            make_uint4(rec0, (unsigned)ws.acted_mask, (unsigned)(ws.acted_mask >> 32), 0u);
                    __threadfence_block();  // also orders the lane's per-turn records before the request
                    __syncwarp();
                    head += (unsigned)n;
                    if (lane == 0) *vhead = head;
                }
            }
            // ---- start of a tick
            if (st == LS_NEWTICK) {
                // hysteresis: a scene goes to the heavy warp at AGV_ROLL_HEAVY_BFS searches per tick and returns below
                // AGV_ROLL_LIGHT_BFS (scenes around the threshold otherwise change warps every other tick, and in the
                // light warp every BFS costs them a whole -- long -- iteration).  (Measured against 16 / 8: 12 / 6 and
                // 10 / 5 -1.2 %; also sending scenes with turns + 1.5 x searches >= 70 / 80 / 90 -- the ones a launch ends
                // on: 64 turns and 15-20 searches a tick in the light warp -- to the heavy warp: -2 / -2 / -1 %.)
                const int w_new = n_bfs >= (unsigned)(which ? ra.light_bfs : ra.heavy_bfs) ? 1 : 0;
                if (k >= K) {          // this scene is through: a worker stores it and fetches the next one
                    ws.regs_to_hdr(); ws.hdr_w[6] = n_bfs | (n_turns << 16);
                    __threadfence_block();
                    vs[lane] = 4u;
                    st = LS_IDLE;
                    RPROF(if (ra.prof) { ra.prof[(size_t)ws.e * 8 + 0] = roll_now(); ra.prof[(size_t)ws.e * 8 + 6] = blockIdx.x; })
                } else if (w_new != which) {  // its last tick was heavy (light): continue on the other scene warp
                    ws.regs_to_hdr(); ws.hdr_w[6] = n_bfs | (n_turns << 16);
                    sh->slot_k[lane] = k; sh->slot_fl[lane] = ws.overlap ? 1u : 0u;
                    __threadfence_block();
                    vctl[lane] = 2u + (unsigned)w_new;
                    st = LS_IDLE;
                } else {
                    bool live = true;
                    if (ws.over) {
                        if (c.auto_reset) {  // Scene.init: every AGV back at (1,1), AGV 0 created
                            ws.reset();
                            for (int q = 0; q < ws.HN; q++) { ws.occ[q] = 0ull; if (mirror_copy<NW>()) ws.occm[q] = 0ull; }
                            ws.occ_set(0, 0);
                            ws.overlap = false;
                        } else live = false;
                    }
                    if (!live) k = K;  // a finished episode without auto-reset: the remaining ticks do nothing
                    else {
                        ws.tick++; ws.acted_mask = 0;
                        i = 0; n_loop = min(N, ws.n_created); n_bfs = 0; pf_idx = -1;
                        rec0 = ((unsigned)k * (unsigned)c.E + (unsigned)ws.e) * (unsigned)N;
                        st = LS_START;
                    }
                }
            }
            // ---- answers of the BFS workers
            if (ans) {
                __threadfence_block();
                const unsigned r = *reinterpret_cast<volatile unsigned *>(&sh->bfs_res[lane]);
                vs[lane] = 0u;
                action = (int)(r & 0xFFu); plen = (int)(r >> 8);
                if (MAT_A && guard) { ws.occ_clear(0, 0); guard = false; }
                st = LS_HAVE;
                RPROF(if (ra.prof) { atomicAdd(ra.prof + (size_t)ws.e * 8 + 3, (unsigned long long)(clock64() - t_post)); atomicAdd(ra.prof + (size_t)ws.e * 8 + 7, (unsigned long long)plen); })
            }
            RQ(1)
        }
        it++;
        RPROF(iters_++;)
        // ==================================================================== fast section
        // ---- B: finish the turn of the lanes that have their action
        // (Explorer.check_action, Explorer.py:151-208; Scene.py:155-158)
        if (st == LS_HAVE) {
            const bool guided = (policy == POLICY_GUIDED) || (policy == POLICY_GIVEN && given < 0);
            const bool pol_a = !guided && policy == POLICY_ASTAR;
            if (GIVEN && !guided && !pol_a && action > A_STOP) { ws.cnt_bad++; action = A_STOP; }
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
                else if (proto == PROTO_INTELLIGENT && ws.other_at_from(i, nx | (ny << 8), a.x | (a.y << 8))) { r = -1; end = 1; }
                else if (at_target) { r = 1; ws.continue_working(a); }
            }
            // Explorer.py:145-148: the move is applied iff not is_end
            if (MODE == 0 && proto != PROTO_INTELLIGENT && !guided && !pol_a) ws.overlap = true;
            if (!end && (dx | dy)) {
                const bool still = ws.other_at_from(i, a.x | (a.y << 8), a.x | (a.y << 8));
                ws.occ_move(a.x - 1, a.y - 1, nx - 1, ny - 1, !still);
                a.x = nx; a.y = ny;
            }
            ws.store_agv(i, a);
            if (proto == PROTO_INTELLIGENT) {
                if (ws.finished || (long long)ws.tick >= c.max_steps) end = 1;  // Scene.py:155
                if (end) ws.over = 1;
            } else {
                end = 0;  // is_end is ignored in A_star / manual mode (Scene.py:159-162)
            }
            ws.acted_mask |= (1ull << i);
            ws.cnt_turns++;
            ws.cnt_ended += end;
            const size_t rec = (size_t)rec0 + i;
            // one 16-byte record per turn: {window before the move, reward | plen | action | done}
            ra.scr_turn[rec] = make_ulonglong2(wm_turn, (unsigned long long)__float_as_uint((float)r) |
                                                            ((unsigned long long)(plen & 0xFFFF) << 32) |
                                                            ((unsigned long long)(action & 0xFF) << 48) |
                                                            ((unsigned long long)end << 56));
            if (want_next) {  // store_info's create_state on the snapshot after the move (Scene.py:156)
                bool shared = false;
                if (ws.overlap) shared = ws.other_at(i, a.x | (a.y << 8));
                const uint64_t wm = lane_window7<NW>(ws.tab + a.loaded * ws.HN, ws.occ, c.H, a.x - 1, a.y - 1, a.tx - 1,
                                                     a.ty - 1, shared);
                const int tpos = (max(-3, min(3, a.ty - a.y)) + 3) * 7 + max(-3, min(3, a.tx - a.x)) + 3;
                ra.scr_next[rec] = wm | ((unsigned long long)tpos << 49);
            }
            i++;
            st = LS_START;
        }
        RQ(2)
        // ---- C: start the next turn (Scene.py:124-134 skip rules first) -- or leave the tick's end to the
        // slow section
        if (st == LS_START) {
            for (;;) {
                if (i >= n_loop || ws.over) { st = LS_TICKEND; break; }
                if (ws.finished) { ws.over = 1; ws.cnt_ended++; st = LS_TICKEND; break; }  // Scene.py:129-130
                if (!((ws.meta_s[i] >> 3) & 1)) break;                                    // idle AGVs are skipped
                i++;
            }
        }
        RQ(3)
        const bool act = st == LS_START;
        uint64_t wm_obs = ~0ull;
        if (act) {
            a = ws.load_agv(i);
            given = -1;
            if (GIVEN && policy == POLICY_GIVEN && t.action) {
                // the action of the next AGV index was requested a turn ago (a global load here would
                // put ~500 cycles of latency on every turn of the scene's chain)
                given = (pf_idx == i) ? pf_val : (int)t.action[(size_t)rec0 + i];
                pf_idx = i + 1;
                pf_val = (i + 1 < N) ? (int)t.action[(size_t)rec0 + i + 1] : -1;
            }
            if (want_obs) {
                bool shared = false;
                if (ws.overlap) shared = ws.other_at(i, a.x | (a.y << 8));
                wm_obs = lane_window7<NW>(ws.tab + a.loaded * ws.HN, ws.occ, c.H, a.x - 1, a.y - 1, a.tx - 1, a.ty - 1,
                                          shared);
                const int tpos = (max(-3, min(3, a.ty - a.y)) + 3) * 7 + max(-3, min(3, a.tx - a.x)) + 3;
                wm_turn = wm_obs | ((unsigned long long)tpos << 49);
            }
        }
        RQ(5)
        if (__any_sync(FULL, act)) {
            const bool guided = (policy == POLICY_GUIDED) || (policy == POLICY_GIVEN && given < 0);
            const bool pol_a = !guided && policy == POLICY_ASTAR;
            bool need = act && (guided || pol_a);
            if (GIVEN && act && !need) { action = given; plen = 0; st = LS_HAVE; }
            const int sx = a.x - 1, sy = a.y - 1, tx = a.tx - 1, ty = a.ty - 1;
            if (need && sx == tx && sy == ty) { action = A_STOP; plen = 0; st = LS_HAVE; need = false; }
            // boxed in: none of the four neighbour cells is valid in matrix (b) -- they are cells 17, 23,
            // 25, 31 of the window just extracted -- so FindPathAstar cannot leave the start: STOP, length 0
            if (need && guided && want_obs && !ws.overlap &&
                !(wm_obs & ((1ull << 17) | (1ull << 23) | (1ull << 25) | (1ull << 31)))) {
                action = A_STOP; plen = 0; st = LS_HAVE; need = false;
            }
            // matrix (a) also blocks (1,1) while uncreated AGVs park there (Explorer.py:274-283)
            if (MAT_A && need && pol_a && N > 1 && ws.n_created < N && !ws.occ_test(0, 0)) { ws.occ_set(0, 0); guard = true; }
            const uint64_t om = guided ? ~0ull : om_a;
            int l = 0;
            RPROF(rows_ += __reduce_max_sync(FULL, need ? abs(ty - sy) : 0);)
            RQ(6)
            // (Remembering per AGV that its last search needed the BFS, and skipping the row DP for it next time,
            // was measured: -5 % at c3, -5 % in A_star mode -- the jams are too short-lived, and every stale bit
            // turns a 2k-cycle DP into a 10k-cycle BFS round trip.)
            const int ma = lane_monotone_pf<NW>(need, ws.tab, ws.HN, om ? ws.occ : zr, om ? ws.occm : zr, a.loaded, sx, sy,
                                                tx, ty, l);
            RQ(7)
            bool want_bfs = false;
            if (need) {
                if (ma >= 0) {
                    action = ma; plen = l; st = LS_HAVE;
                    if (MAT_A && guard) { ws.occ_clear(0, 0); guard = false; }
                    // about to reach the target: the task word of the next stage will be needed on the chain
                    if (l <= 2) asm volatile("prefetch.global.L1 [%0];" ::"l"(ws.tasks_e + a.order));
                } else want_bfs = true;   // no Manhattan-length path: the level-synchronous BFS decides
            }
            const unsigned pk = (unsigned)sx | ((unsigned)sy << 7) | ((unsigned)tx << 14) | ((unsigned)ty << 21) |
                                ((unsigned)a.loaded << 28) | (om ? (1u << 29) : 0u);
            // (Posting the heavy warp's request BEFORE its row DP, and cancelling or orphaning it when the DP answers,
            // was measured: -3 % guided, -11 % in A_star mode -- the orphans take the workers from real searches.)
            // (Letting the heavy warp run one pending search itself -- no hand-off -- was measured: -15 %; its other
            // scenes, whose answers are ready, then wait for that search.)
            if (want_bfs) {  // by a worker warp; this scene waits, the others go on
                *reinterpret_cast<volatile unsigned *>(&sh->bfs_req[lane]) = pk;
                __threadfence_block();
                vs[lane] = 1u;
                st = LS_WAIT;
                n_bfs++;
                RPROF(t_post = clock64();)
            }
        }
        RQ(6)
    }
#ifdef AGV_ROLL_PROF
    if (lane == 0) {  // counters[8 + 10 * which + k]: cycles of phase k, [.. + 8]: iterations, [.. + 9]: warps
        for (int k_ = 0; k_ < 8; k_++) atomicAdd(c.counters + 8 + 10 * which + k_, (unsigned long long)q_[k_]);
        atomicAdd(c.counters + 8 + 10 * which + 8, (unsigned long long)iters_);
        atomicAdd(c.counters + 8 + 10 * which + 9, 1ull);
        atomicAdd(c.counters + 28 + which, (unsigned long long)rows_);
    }
#endif
    __threadfence_block();
    __syncwarp();
    if (lane == 0 && atomicAdd(&sh->done_warps, 1u) == 1u) *reinterpret_cast<volatile int *>(&sh->exit_flag) = 1;
    {
        const unsigned ct = __reduce_add_sync(FULL, ws.cnt_turns), ce = __reduce_add_sync(FULL, ws.cnt_ended),
                       cb = __reduce_add_sync(FULL, ws.cnt_bad);
        if (lane == 0) {
            if (ct) atomicAdd(c.counters + 0, (unsigned long long)ct);
            if (ce) atomicAdd(c.counters + 1, (unsigned long long)ce);
            if (cb) atomicAdd(c.counters + 2, (unsigned long long)cb);
        }
    }
}

// kernel body: role by warp index
template <int NW, int RPL, int OBS, int SPW, int NB, int MODE>
__device__ __forceinline__ void rollout_cta(const DevCfg &c, const RollArgs &ra, uint8_t *smem) {
    const LaneArgs &t = ra.la;
    RollShared *sh = reinterpret_cast<RollShared *>(smem + t.off_queue);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    if (warp == 0) {
        sh->bfs_state[lane] = 0u; sh->ctl[lane] = 4u; sh->slot_e[lane] = -1; sh->slot_k[lane] = 0; sh->slot_fl[lane] = 0u;
        if (lane == 0) {
            sh->exit_flag = 0; sh->done_warps = 0u; sh->done_slots = 0u; sh->heavy_live = 0u;
            sh->enc_head[0] = sh->enc_head[1] = 0u; sh->enc_tail[0] = sh->enc_tail[1] = 0u;
        }
    } else if (warp == 1) {
        LaneScene<NW, RPL, SPW, 0> ws(c, t, smem, lane, 0);
        ws.load_tables();
    } else if (warp == 2) {
        for (int b = lane; b < 256; b += 32) {
            uint4 o;
            unsigned u, w;
            u = b & 3u;        w = (u | (u << 15)) & 0x10001u; o.x = w * (unsigned)BF16_ONE;
            u = (b >> 2) & 3u; w = (u | (u << 15)) & 0x10001u; o.y = w * (unsigned)BF16_ONE;
            u = (b >> 4) & 3u; w = (u | (u << 15)) & 0x10001u; o.z = w * (unsigned)BF16_ONE;
            u = (b >> 6) & 3u; w = (u | (u << 15)) & 0x10001u; o.w = w * (unsigned)BF16_ONE;
            sh->lut[b] = o;
        }
    }
    __syncthreads();
    // first fill: slot j of CTA b takes entry b + j * gridDim of the longest-first order, so the long
    // scenes are spread over all CTAs; later scenes come from the queue counter (starts at SPW * gridDim)
    for (int j = warp; j < SPW; j += n_warps) {
        const unsigned idx = blockIdx.x + (unsigned)j * gridDim.x;
        if (idx < (unsigned)c.E) roll_load_slot<NW>(c, t, smem, sh, j, (int)__ldg(ra.order + idx), lane, ra.heavy_bfs);
        else if (lane == 0) atomicAdd(&sh->done_slots, 1u);
    }
    __syncthreads();
    // (More heavy scene warps per CTA -- heavy warp 1 + slot % n, so that the congested scenes a launch ends on share
    // their warp with fewer others -- were measured: 2 -> -18 %, 3 -> -21 % guided, -4 / -15 % in A_star mode.  An
    // iteration costs a warp the same for one active lane as for 32; every additional scene warp that runs takes
    // that much issue / ALU-pipe time from the light warp, which carries most of the work.)
    // Role by warp index.  The SM sub-partition's arbiter is reported to pick the eligible warp with the highest
    // warp id first, so the scene warps -- whose dependent instruction chains are the critical path of every
    // scene -- get the highest ids of the CTA, the BFS workers and the record writer (throughput work with
    // slack) the lowest.  (Measured against scene warps at ids 0 / 1: no difference, 1.06e9 both ways.)
    if (warp < NB) roll_worker<NW, RPL, SPW>(c, ra, smem);
    else if (warp == NB) roll_writer(ra, smem);  // (also without observations: it unpacks act / reward / done / plen)
    else roll_scene_warp<NW, RPL, OBS, SPW, MODE>(c, ra, smem, warp == NB + 2 ? 0 : 1);
}

}  // namespace agv
