// agv_kernels.cu -- kernels + C ABI (include/agv_b200.h) of the batched scene.
// sm_100a only.  See agv_core.cuh for the warp-per-scene device core.
#include "agv_rollout.cuh"
#include "../../include/agv_b200.h"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

namespace agv {

int lanes_tick(const DevCfg &dc, int NW, int RPL, const LaneArgs &t, cudaStream_t st, int allow_async);  // agv_lanes.cu
int lanes_rollout(const DevCfg &dc, int NW, int RPL, const RollArgs &ra, unsigned *order, cudaStream_t st);  // agv_rollout.cu

static std::atomic<uint64_t> g_launches{0};
static const char *g_tick_kernel = "none";
static int g_last_cuda = 0;

struct TickArgs {
    int proto, policy, obs_kind;
    const int8_t *action;
    int8_t *act_out;
    float *reward;
    uint8_t *done;
    uint16_t *plen, *slen;
    uint16_t *obs, *next_obs;
    int rec_elems;  // bf16 elements per obs record
    int group;      // records per 16-byte aligned flush group (OBS_CLIP)
    // per-warp shared memory plan (bytes)
    int per_warp, off_blk, off_vrow, off_act, off_done, off_plen, off_slen, off_rew, off_stage, off_stage2;
};

struct StepArgs {
    int k, proto, obs_kind, post, matrix;
    const int8_t *action;
    int8_t *action_out;
    float *reward;
    uint8_t *done, *acted;
    uint16_t *slen, *plen;
    uint16_t *obs;
    int rec_elems;
    int per_warp, off_blk, off_vrow, off_stage;
};

#define WARP_PROLOGUE(ARGS)                                                        \
    extern __shared__ __align__(16) uint8_t smem[];                                \
    const int wic = threadIdx.x >> 5, lane = threadIdx.x & 31;                     \
    const int e = blockIdx.x * (blockDim.x >> 5) + wic;                            \
    if (e >= c.E) return;                                                          \
    uint8_t *wbase = smem + (size_t)wic * ARGS.per_warp;                           \
    WarpScene<NW, RPL> ws(c, lane);                                                \
    ws.bind(wbase + ARGS.off_blk, reinterpret_cast<uint64_t *>(wbase + ARGS.off_vrow), e); \
    ws.load_block(wbase + ARGS.off_blk, e);                                        \
    ws.load_static();                                                              \
    ws.hdr_to_regs();

// ---------------------------------------------------------------- fused tick
// resident CTAs per SM the register budget is tuned for (ptxas -v: no spills at these bounds)
template <int NW, int RPL>
constexpr int tick_min_blocks() {
#ifndef AGV_TICK_MINB
#define AGV_TICK_MINB 6   // rows fit 2 u64 per lane: 80 registers, no spills
#endif
#ifndef AGV_TICK_MINB4
#define AGV_TICK_MINB4 4  // 4 u64 per lane
#endif
#ifndef AGV_TICK_MINB8
#define AGV_TICK_MINB8 2  // 8 u64 per lane (128 x 128)
#endif
#ifndef AGV_TICK_MINB1
#define AGV_TICK_MINB1 7  // 1 u64 per lane (maps up to 32 rows x 64 columns): 72 registers, no spills -- 28 one-warp CTAs per SM,
                          // 4096 scenes are one wave on 148 SMs (README 9x9, 4096 scenes: 5.08e8 -> 5.65e8 AGV-steps/s)
#endif
    return NW * RPL <= 1 ? AGV_TICK_MINB1 : NW * RPL <= 2 ? AGV_TICK_MINB : (NW * RPL <= 4 ? AGV_TICK_MINB4 : AGV_TICK_MINB8);
}

template <int NW, int RPL, int OBS, bool SHAPED>
__global__ void __launch_bounds__(128, tick_min_blocks<NW, RPL>()) k_tick(const DevCfg c, const TickArgs t) {
    WARP_PROLOGUE(t)
    int8_t *o_act = reinterpret_cast<int8_t *>(wbase + t.off_act);
    uint8_t *o_done = wbase + t.off_done;
    uint16_t *o_plen = reinterpret_cast<uint16_t *>(wbase + t.off_plen);
    uint16_t *o_slen = reinterpret_cast<uint16_t *>(wbase + t.off_slen);
    float *o_rew = reinterpret_cast<float *>(wbase + t.off_rew);
    uint16_t *stage = reinterpret_cast<uint16_t *>(wbase + t.off_stage);
    uint16_t *stage2 = reinterpret_cast<uint16_t *>(wbase + t.off_stage2);
    const int N = c.N, R = t.rec_elems, G = t.group;
    constexpr bool clip = OBS == OBS_CLIP, full = OBS == OBS_FULL;

    bool live = true;
    if (ws.over) {
        if (c.auto_reset) ws.reset();
        else live = false;
    }
    if (live) { ws.build_occ(); ws.tick++; ws.acted_mask = 0; }
    // per-AGV outputs default to "no turn"; the turn loop only visits the created prefix and
    // stops at the first uncreated AGV (Scene.py:127-128) or when the episode is over
    for (int j = lane; j < N; j += 32) { o_act[j] = -1; o_done[j] = 0; o_rew[j] = 0.f; o_plen[j] = 0; o_slen[j] = 0xFFFF; }
    __syncwarp();
    const int n_loop = live ? min(N, ws.n_created) : 0;  // spawn happens after the loop
    const int8_t *given_row = (t.policy == POLICY_GIVEN && t.action) ? t.action + (size_t)e * N : nullptr;
    bool group_dirty = false;
    int i = 0;
    for (; i < n_loop && !ws.over; i++) {
        bool act = false;
        if (ws.finished) {                                   // Scene.py:129-130 return
            ws.over = 1;
            ws.cnt_ended++;
        } else act = !((ws.meta_s[i] >> 3) & 1);             // Scene.py:132-134 idle AGVs are skipped
        const size_t rec = (size_t)e * N + i;
        uint16_t *od = nullptr, *nd = nullptr;
        if (clip) {
            if (t.obs) od = stage + (size_t)(i & (G - 1)) * R;
            if (t.next_obs) nd = stage2 + (size_t)(i & (G - 1)) * R;
        } else if (full) {
            if (t.obs) od = t.obs + rec * R;
            if (t.next_obs) nd = t.next_obs + rec * R;
        }
        if (act) {
            TurnOut o;
            const int given = given_row ? (int)given_row[i] : -1;
            do_turn<NW, RPL, OBS, SHAPED>(ws, i, t.proto, t.policy, given, od, nd, o);
            // warp-uniform values, stored by every lane (no divergent lane-0 block)
            o_act[i] = (int8_t)o.action; o_done[i] = (uint8_t)o.done; o_rew[i] = o.reward;
            o_plen[i] = (uint16_t)o.plen; o_slen[i] = (uint16_t)o.slen;
        } else if (clip) {
            // no turn: observation records of this AGV are left untouched in HBM (the consumer
            // masks with act_out >= 0); inside a flush group the slot is zero-filled so that the
            // group still goes out as one aligned vector store
            if (od) for (int q = lane; q < R; q += 32) od[q] = 0;
            if (nd) for (int q = lane; q < R; q += 32) nd[q] = 0;
        }
        if (clip) {
            group_dirty |= act;
            if (((i + 1) & (G - 1)) == 0) {  // G is a power of two: group complete
                if (group_dirty) {
                    const size_t off = ((size_t)e * N + (i & ~(G - 1))) * R;
                    if (t.obs) flush_bytes(reinterpret_cast<uint8_t *>(t.obs + off),
                                           reinterpret_cast<uint8_t *>(stage), G * R * 2, lane);
                    if (t.next_obs) flush_bytes(reinterpret_cast<uint8_t *>(t.next_obs + off),
                                                reinterpret_cast<uint8_t *>(stage2), G * R * 2, lane);
                }
                group_dirty = false;
            }
        }
    }
    if (clip && group_dirty && (i & (G - 1)) != 0) {  // partially filled last group
        const int g0 = i & ~(G - 1), nrec = i - g0;
        const size_t off = ((size_t)e * N + g0) * R;
        if (t.obs) flush_bytes(reinterpret_cast<uint8_t *>(t.obs + off), reinterpret_cast<uint8_t *>(stage),
                               nrec * R * 2, lane);
        if (t.next_obs) flush_bytes(reinterpret_cast<uint8_t *>(t.next_obs + off),
                                    reinterpret_cast<uint8_t *>(stage2), nrec * R * 2, lane);
    }
    if (live && !ws.over) ws.spawn();
    ws.regs_to_hdr();
    ws.flush_counters();
    ws.store_block(wbase + t.off_blk, e);
    for (int j = lane; j < N; j += 32) {
        const size_t rec = (size_t)e * N + j;
        if (t.act_out) t.act_out[rec] = o_act[j];
        if (t.done) t.done[rec] = o_done[j];
        if (t.reward) t.reward[rec] = o_rew[j];
        if (t.plen) t.plen[rec] = o_plen[j];
        if (t.slen) t.slen[rec] = o_slen[j];
    }
}

// Small maps (up to 32 rows x 64 columns, one u64 per lane): the same tick with the loop over n_ticks INSIDE the
// launch -- a tick of the README 9x9 scene is launch-sized (24 us for 4096 scenes), so the scene stays in its warp
// and the outputs advance by E * N records per tick.  A separate kernel with its own n_ticks parameter on purpose:
// four more bytes in TickArgs changed the register allocation of every k_tick instance (8 -> 40 bytes of spills in
// the HBM-bound full-map encode at 64 x 64, -3 %).
template <int NW, int RPL, int OBS, bool SHAPED>
__global__ void __launch_bounds__(128, tick_min_blocks<NW, RPL>()) k_tick_small(const DevCfg c, const TickArgs t, const int n_ticks) {
    WARP_PROLOGUE(t)
    int8_t *o_act = reinterpret_cast<int8_t *>(wbase + t.off_act);
    uint8_t *o_done = wbase + t.off_done;
    uint16_t *o_plen = reinterpret_cast<uint16_t *>(wbase + t.off_plen);
    uint16_t *o_slen = reinterpret_cast<uint16_t *>(wbase + t.off_slen);
    float *o_rew = reinterpret_cast<float *>(wbase + t.off_rew);
    uint16_t *stage = reinterpret_cast<uint16_t *>(wbase + t.off_stage);
    uint16_t *stage2 = reinterpret_cast<uint16_t *>(wbase + t.off_stage2);
    const int N = c.N, R = t.rec_elems, G = t.group;
    constexpr bool clip = OBS == OBS_CLIP, full = OBS == OBS_FULL;

  const size_t tick_recs = (size_t)c.E * N;
  for (int tk = 0; tk < n_ticks; tk++) {
    const size_t rec0 = (size_t)tk * tick_recs + (size_t)e * N;   // first output record of this scene-tick
    bool live = true;
    if (ws.over) {
        if (c.auto_reset) ws.reset();
        else live = false;
    }
    if (live) { ws.build_occ(); ws.tick++; ws.acted_mask = 0; }
    // per-AGV outputs default to "no turn"; the turn loop only visits the created prefix and
    // stops at the first uncreated AGV (Scene.py:127-128) or when the episode is over
    for (int j = lane; j < N; j += 32) { o_act[j] = -1; o_done[j] = 0; o_rew[j] = 0.f; o_plen[j] = 0; o_slen[j] = 0xFFFF; }
    __syncwarp();
    const int n_loop = live ? min(N, ws.n_created) : 0;  // spawn happens after the loop
    const int8_t *given_row = (t.policy == POLICY_GIVEN && t.action) ? t.action + rec0 : nullptr;
    bool group_dirty = false;
    int i = 0;
    for (; i < n_loop && !ws.over; i++) {
        bool act = false;
        if (ws.finished) {                                   // Scene.py:129-130 return
            ws.over = 1;
            ws.cnt_ended++;
        } else act = !((ws.meta_s[i] >> 3) & 1);             // Scene.py:132-134 idle AGVs are skipped
        const size_t rec = rec0 + i;
        uint16_t *od = nullptr, *nd = nullptr;
        if (clip) {
            if (t.obs) od = stage + (size_t)(i & (G - 1)) * R;
            if (t.next_obs) nd = stage2 + (size_t)(i & (G - 1)) * R;
        } else if (full) {
            if (t.obs) od = t.obs + rec * R;
            if (t.next_obs) nd = t.next_obs + rec * R;
        }
        if (act) {
            TurnOut o;
            const int given = given_row ? (int)given_row[i] : -1;
            do_turn<NW, RPL, OBS, SHAPED>(ws, i, t.proto, t.policy, given, od, nd, o);
            // warp-uniform values, stored by every lane (no divergent lane-0 block)
            o_act[i] = (int8_t)o.action; o_done[i] = (uint8_t)o.done; o_rew[i] = o.reward;
            o_plen[i] = (uint16_t)o.plen; o_slen[i] = (uint16_t)o.slen;
        } else if (clip) {
            // no turn: observation records of this AGV are left untouched in HBM (the consumer
            // masks with act_out >= 0); inside a flush group the slot is zero-filled so that the
            // group still goes out as one aligned vector store
            if (od) for (int q = lane; q < R; q += 32) od[q] = 0;
            if (nd) for (int q = lane; q < R; q += 32) nd[q] = 0;
        }
        if (clip) {
            group_dirty |= act;
            if (((i + 1) & (G - 1)) == 0) {  // G is a power of two: group complete
                if (group_dirty) {
                    const size_t off = (rec0 + (i & ~(G - 1))) * R;
                    if (t.obs) flush_bytes(reinterpret_cast<uint8_t *>(t.obs + off),
                                           reinterpret_cast<uint8_t *>(stage), G * R * 2, lane);
                    if (t.next_obs) flush_bytes(reinterpret_cast<uint8_t *>(t.next_obs + off),
                                                reinterpret_cast<uint8_t *>(stage2), G * R * 2, lane);
                }
                group_dirty = false;
            }
        }
    }
    if (clip && group_dirty && (i & (G - 1)) != 0) {  // partially filled last group
        const int g0 = i & ~(G - 1), nrec = i - g0;
        const size_t off = (rec0 + g0) * R;
        if (t.obs) flush_bytes(reinterpret_cast<uint8_t *>(t.obs + off), reinterpret_cast<uint8_t *>(stage),
                               nrec * R * 2, lane);
        if (t.next_obs) flush_bytes(reinterpret_cast<uint8_t *>(t.next_obs + off),
                                    reinterpret_cast<uint8_t *>(stage2), nrec * R * 2, lane);
    }
    if (live && !ws.over) ws.spawn();
    if (tk + 1 == n_ticks) {  // (before the output stores, as in the one-tick kernel: the scene's registers die here)
        ws.regs_to_hdr();
        ws.flush_counters();
        ws.store_block(wbase + t.off_blk, e);
    }
    __syncwarp();
    for (int j = lane; j < N; j += 32) {
        const size_t rec = rec0 + j;
        if (t.act_out) t.act_out[rec] = o_act[j];
        if (t.done) t.done[rec] = o_done[j];
        if (t.reward) t.reward[rec] = o_rew[j];
        if (t.plen) t.plen[rec] = o_plen[j];
        if (t.slen) t.slen[rec] = o_slen[j];
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------- sub-step kernels
template <int NW, int RPL>
__global__ void __launch_bounds__(128) k_reset(const DevCfg c, const StepArgs t, const uint8_t *mask) {
    WARP_PROLOGUE(t)
    if (mask && !mask[e]) return;
    ws.reset();
    ws.regs_to_hdr();
    ws.store_block(wbase + t.off_blk, e);
}

template <int NW, int RPL>
__global__ void __launch_bounds__(128) k_begin_tick(const DevCfg c, const StepArgs t) {
    WARP_PROLOGUE(t)
    if (ws.over && c.auto_reset) ws.reset();
    if (!ws.over) { ws.tick++; ws.acted_mask = 0; }
    ws.regs_to_hdr();
    ws.store_block(wbase + t.off_blk, e);
}

template <int NW, int RPL>
__global__ void __launch_bounds__(128) k_end_tick(const DevCfg c, const StepArgs t) {
    WARP_PROLOGUE(t)
    if (ws.over) return;
    ws.build_occ();
    ws.spawn();
    ws.regs_to_hdr();
    ws.store_block(wbase + t.off_blk, e);
}

template <int NW, int RPL>
__global__ void __launch_bounds__(128) k_encode(const DevCfg c, const StepArgs t) {
    WARP_PROLOGUE(t)
    const int k = t.k, R = t.rec_elems;
    bool go;
    if (t.post) go = (ws.acted_mask >> k) & 1ull;
    else go = (!ws.over) && ws.turn_state(k) == 1;
    if (t.acted && lane == 0) t.acted[e] = go ? 1 : 0;
    if (!t.obs) return;
    uint16_t *dst = t.obs + (size_t)e * R;
    if (!go) { zero_global(dst, R, lane); return; }
    ws.build_occ();
    Agv a = ws.load_agv(k);
    Row<NW> v[RPL];
    ws.valid_b(a, k, v);
    ws.rows_to_smem(v);
    if (t.obs_kind == OBS_CLIP) {
        uint16_t *stage = reinterpret_cast<uint16_t *>(wbase + t.off_stage);
        ws.encode_clip(stage, a);
        flush_bytes(reinterpret_cast<uint8_t *>(dst), reinterpret_cast<uint8_t *>(stage), R * 2, lane);
    } else {
        ws.encode_full(dst, a);
    }
}

template <int NW, int RPL>
__global__ void __launch_bounds__(128) k_bfs_action(const DevCfg c, const StepArgs t) {
    WARP_PROLOGUE(t)
    const int k = t.k;
    const bool go = (!ws.over) && ws.turn_state(k) == 1;
    int action = -1, plen = 0;
    if (go) {
        ws.build_occ();
        Agv a = ws.load_agv(k);
        Row<NW> v[RPL];
        if (t.matrix == 0) ws.valid_a(a, v);
        else ws.valid_b(a, k, v);
        action = warp_path<NW, RPL>(v, a.x - 1, a.y - 1, a.tx - 1, a.ty - 1, c.H, c.W, lane, plen);
    }
    if (lane == 0) {
        if (t.action_out) t.action_out[e] = (int8_t)action;
        if (t.plen) t.plen[e] = (uint16_t)plen;
    }
}

template <int NW, int RPL>
__global__ void __launch_bounds__(128) k_step_turn(const DevCfg c, const StepArgs t) {
    WARP_PROLOGUE(t)
    const int k = t.k;
    TurnOut o;
    o.action = -1; o.reward = 0.f; o.done = 0; o.plen = 0; o.slen = 0xFFFF;
    bool acted = false;
    if (!ws.over) {
        const int st = ws.turn_state(k);
        if (st == -2) {
            ws.over = 1;
            ws.cnt_ended++;
        } else if (st == 1) {
            ws.build_occ();
            const int given = t.action ? (int)t.action[e] : -1;
            if (c.shaped) do_turn<NW, RPL, OBS_NONE, true>(ws, k, t.proto, POLICY_GIVEN, given, nullptr, nullptr, o);
            else do_turn<NW, RPL, OBS_NONE, false>(ws, k, t.proto, POLICY_GIVEN, given, nullptr, nullptr, o);
            acted = true;
        }
        ws.regs_to_hdr();
        ws.flush_counters();
        ws.store_block(wbase + t.off_blk, e);
    }
    if (lane == 0) {
        if (t.reward) t.reward[e] = o.reward;
        if (t.done) t.done[e] = (uint8_t)o.done;
        if (t.acted) t.acted[e] = acted ? 1 : 0;
        if (t.slen) t.slen[e] = (uint16_t)o.slen;
        if (t.action_out) t.action_out[e] = (int8_t)o.action;
    }
}

// ---------------------------------------------------------------- standalone BFS
template <int NW, int RPL>
__device__ __forceinline__ void rows_from_bytes(const uint8_t *m, int H, int W, int lane, Row<NW> (&v)[RPL]) {
#pragma unroll
    for (int s = 0; s < RPL; s++) {
        v[s] = rzero<NW>();
        const int r = lane * RPL + s;
        if (r < H) {
            for (int cc = 0; cc < W; cc++)
                if (m[r * W + cc] == 1) {
#pragma unroll
                    for (int i = 0; i < NW; i++)
                        if (i == (cc >> 6)) v[s].w[i] |= 1ull << (cc & 63);
                }
        }
    }
}

template <int NW, int RPL>
__global__ void __launch_bounds__(128) k_bfs_batch(int B, int H, int W, const uint8_t *valid, const int32_t *start,
                                                   const int32_t *target, int8_t *action, int32_t *pathlen,
                                                   uint8_t *found) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    Row<NW> v[RPL];
    rows_from_bytes<NW, RPL>(valid + (size_t)b * H * W, H, W, lane, v);
    const int sx = start[2 * b], sy = start[2 * b + 1], tx = target[2 * b], ty = target[2 * b + 1];
    int len = 0;
    const int a = warp_path<NW, RPL>(v, sx, sy, tx, ty, H, W, lane, len);
    if (lane == 0) {
        if (action) action[b] = (int8_t)a;
        if (pathlen) pathlen[b] = len;
        if (found) found[b] = (uint8_t)((sx == tx && sy == ty) || len > 0);
    }
}

template <int NW, int RPL>
__global__ void __launch_bounds__(128) k_bfs_field(int B, int H, int W, const uint8_t *valid, const int32_t *target,
                                                   uint16_t *dist) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    Row<NW> v[RPL];
    rows_from_bytes<NW, RPL>(valid + (size_t)b * H * W, H, W, lane, v);
    warp_bfs_field<NW, RPL>(v, target[2 * b], target[2 * b + 1], H, W, lane, dist + (size_t)b * H * W);
}

__global__ void k_init_state(const DevCfg c) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= c.E) return;
    uint8_t *blk = c.state + (size_t)e * c.env_stride;
    for (int i = 0; i < c.env_stride; i++) blk[i] = 0;
    EnvHdr *h = reinterpret_cast<EnvHdr *>(blk);
    h->over = 1;
    uint16_t *pos = reinterpret_cast<uint16_t *>(blk + c.off_pos);
    uint16_t *tgt = reinterpret_cast<uint16_t *>(blk + c.off_tgt);
    for (int j = 0; j < c.N; j++) { pos[j] = 0x0101; tgt[j] = 0x0101; }  // Explorer.py:27-28
}

}  // namespace agv

// =================================================================== host side
using namespace agv;

struct AgvHandle {
    AgvConfig cfg;
    DevCfg dc;
    int NW, RPL;
    uint8_t *d_state = nullptr;
    uint32_t *d_tasks = nullptr;
    uint8_t *d_grid = nullptr;
    uint64_t *d_rows = nullptr;  // storage | picking | inb
    unsigned long long *d_counters = nullptr;
    // staging for agv_tick_host: results are double-buffered on the device and copied back on an
    // internal stream, so the D2H of call k overlaps the tick of call k+1
    int8_t *d_action[2] = {nullptr, nullptr}, *d_act_out[2] = {nullptr, nullptr};
    float *d_reward[2] = {nullptr, nullptr};
    uint8_t *d_done[2] = {nullptr, nullptr};
    cudaStream_t copy_stream = nullptr, h2d_stream = nullptr;
    cudaEvent_t ev_tick[2] = {nullptr, nullptr}, ev_copy[2] = {nullptr, nullptr}, ev_h2d[2] = {nullptr, nullptr};
    bool tick_pending[2] = {false, false};
    bool copy_pending[2] = {false, false};
    unsigned host_calls = 0;
    std::vector<uint8_t> h_state;
    // persistent rollout (agv_rollout): longest-first scene order + queue counter, packed-window scratch
    // rings of the record writer, and double-buffered result rings for agv_rollout_host
    unsigned *d_order = nullptr, *d_qctr = nullptr;
    unsigned long long *h_rstat = nullptr;  // pinned: load figure written by k_rollout_order (agv_rollout.cu)
    unsigned long long *d_scr[2] = {nullptr, nullptr};
    size_t scr_cap[2] = {0, 0};
    int8_t *r_action[2] = {nullptr, nullptr}, *r_act_out[2] = {nullptr, nullptr};
    float *r_reward[2] = {nullptr, nullptr};
    uint8_t *r_done[2] = {nullptr, nullptr};
    size_t r_cap = 0;
    unsigned roll_calls = 0;
    unsigned long long *d_prof = nullptr;  // -DAGV_ROLL_PROF builds with AGV_ROLL_PROF=1 in the environment
};

static inline int align16(int x) { return (x + 15) & ~15; }

#define CUDA_TRY(expr)                              \
    do {                                            \
        cudaError_t _e = (expr);                    \
        if (_e != cudaSuccess) {                    \
            g_last_cuda = (int)_e;                  \
            return AGV_E_CUDA;                      \
        }                                           \
    } while (0)

#define DISPATCH(h, CALL)                                                  \
    do {                                                                   \
        if (h->NW == 1 && h->RPL == 1) { CALL(1, 1); }                     \
        else if (h->NW == 1 && h->RPL == 2) { CALL(1, 2); }                \
        else if (h->NW == 1 && h->RPL == 4) { CALL(1, 4); }                \
        else if (h->NW == 2 && h->RPL == 1) { CALL(2, 1); }                \
        else if (h->NW == 2 && h->RPL == 2) { CALL(2, 2); }                \
        else { CALL(2, 4); }                                               \
    } while (0)

static int pick_variant(int H, int W, int *NW, int *RPL) {
    if (H < 1 || W < 1 || H > 128 || W > 128) return AGV_E_UNSUPPORTED;
    *NW = W <= 64 ? 1 : 2;
    *RPL = H <= 32 ? 1 : (H <= 64 ? 2 : 4);
    return AGV_OK;
}

static int gcd_i(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }

static int warps_per_cta() {
    static int v = 0;
    if (v == 0) {
        const char *e = getenv("AGV_WARPS_PER_CTA");  // tuning knob; 1..4
        v = e ? atoi(e) : 1;  // 1 warp per CTA: a scene frees its SM slot as soon as it is done
        if (v < 1 || v > 4) v = 1;
    }
    return v;
}

// fused-tick implementation: "async" (lane-per-scene with decoupled lanes, agv_async.cuh),
// "lanes" (lane-per-scene in lock step, agv_lanes.cuh) or "warp" (warp-per-scene k_tick).  All are
// bit-identical (tests/test_gpu_parity.py runs every fused-tick test on each).  AGV_TICK_IMPL
// forces one for A/B runs and tests; otherwise the measured-fastest one for the shape is taken:
//   * full-map observations are HBM-write-bound and best written by a whole warp per scene;
//   * maps up to 32 x 32 have short searches and few scenes per SM: warp-per-scene;
//   * everything else runs lane-per-scene (decoupled lanes where agv_async.cuh covers it).
static int tick_impl(const AgvHandle *h, int obs_kind) {  // 0 warp, 1 lanes, 2 async, 3 rollout
    const char *e = getenv("AGV_TICK_IMPL");
    if (e && strcmp(e, "warp") == 0) return 0;
    if (e && strcmp(e, "lanes") == 0) return 1;
    if (e && strcmp(e, "async") == 0) return 2;
    if (e && strcmp(e, "rollout") == 0) return 3;  // the persistent rollout kernel with n_ticks = 1
    if (obs_kind == AGV_OBS_FULL) return 0;
    if (h->NW == 1 && h->RPL == 1) return 0;
    return 2;
}

template <typename K, typename... Args>
static int launch_warps(K kernel, int n_warps, size_t smem_per_warp, cudaStream_t st, Args... args) {
    const int WARPS_PER_CTA = warps_per_cta();
    const int ctas = (n_warps + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    const size_t smem = smem_per_warp * WARPS_PER_CTA;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { g_last_cuda = (int)e; return AGV_E_CUDA; }
    }
    if (smem > 0) {
        // the scenes live in shared memory and nothing is re-read through L1: give smem the carveout
        cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
    kernel<<<ctas, WARPS_PER_CTA * 32, smem, st>>>(args...);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_last_cuda = (int)e; return AGV_E_CUDA; }
    return AGV_OK;
}

static void plan_step(const AgvHandle *h, StepArgs &s, int rec_elems) {
    int off = 0;
    s.off_blk = off; off += align16(h->dc.env_stride);
    s.off_vrow = off; off += align16(h->cfg.H * h->NW * 8);
    s.off_stage = off; off += align16(rec_elems * 2);
    s.per_warp = off;
    s.rec_elems = rec_elems;
}

static int obs_elems(const AgvHandle *h, int obs_kind) {
    if (obs_kind == AGV_OBS_CLIP) return 3 * h->dc.K * h->dc.K;
    if (obs_kind == AGV_OBS_FULL) return 3 * h->cfg.H * h->cfg.W;
    return 0;
}

// library-owned copy streams + events of the *_host entry points (created on first use)
static cudaError_t host_streams(AgvHandle *h) {
    if (h->copy_stream) return cudaSuccess;
    cudaError_t e;
    for (int b = 0; b < 2; b++) {
        if ((e = cudaEventCreateWithFlags(&h->ev_tick[b], cudaEventDisableTiming)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&h->ev_copy[b], cudaEventDisableTiming)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&h->ev_h2d[b], cudaEventDisableTiming)) != cudaSuccess) return e;
    }
    if ((e = cudaStreamCreateWithFlags(&h->h2d_stream, cudaStreamNonBlocking)) != cudaSuccess) return e;
    return cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
}

// Launch of the persistent rollout kernel (agv_rollout.cuh) for n_ticks ticks; AGV_E_UNSUPPORTED when
// the configuration is outside what that kernel covers (the caller then loops over agv_tick).
static int rollout_launch(AgvHandle *h, int n_ticks, int proto, int policy, int obs_kind, const int8_t *action_dev,
                          int8_t *act_out_dev, float *reward_dev, uint8_t *done_dev, uint16_t *plen_dev,
                          uint16_t *slen_dev, void *obs_dev, void *next_obs_dev, cudaStream_t st) {
    const DevCfg &d = h->dc;
    const size_t per_tick = (size_t)d.E * d.N, n = per_tick * (size_t)n_ticks;
    if (n >= (1ull << 31)) return AGV_E_UNSUPPORTED;
    if (d.shaped || !(obs_kind == AGV_OBS_NONE || (obs_kind == AGV_OBS_CLIP && d.K == 7))) return AGV_E_UNSUPPORTED;
    // maps up to 32 rows x 64 columns: the warp-per-scene fused tick (grid in registers, BFS in the scene's own
    // warp) is faster than lane-per-scene with BFS hand-offs -- README 9x9, 4 AGVs, 4096 scenes: 5.1e8 AGV-steps/s
    // as 32 fused ticks per call against 3.0e8 in the persistent kernel.  AGV_ROLLOUT_IMPL=persistent forces the
    // persistent kernel (tests run it on every shape).
    {
        const char *e = getenv("AGV_ROLLOUT_IMPL");
        if (h->NW == 1 && h->RPL == 1 && !(e && strcmp(e, "persistent") == 0)) return AGV_E_UNSUPPORTED;
    }
    if (!h->d_order) {
        CUDA_TRY(cudaMalloc(&h->d_order, sizeof(unsigned) * (size_t)d.E));
        CUDA_TRY(cudaMalloc(&h->d_qctr, sizeof(unsigned)));
        CUDA_TRY(cudaHostAlloc(&h->h_rstat, 2 * sizeof(unsigned long long), cudaHostAllocMapped));
        h->h_rstat[0] = h->h_rstat[1] = 0ull;
    }
    // per-turn scratch records of the scene lanes (agv_rollout.cuh): 16 B per turn, + 8 B with next_obs
    const size_t need[2] = {n * 2, next_obs_dev ? n : 0};
    for (int b = 0; b < 2; b++)
        if (h->scr_cap[b] < need[b]) {
            if (h->d_scr[b]) { CUDA_TRY(cudaStreamSynchronize(st)); CUDA_TRY(cudaFree(h->d_scr[b])); h->d_scr[b] = nullptr; h->scr_cap[b] = 0; }
            CUDA_TRY(cudaMalloc(&h->d_scr[b], need[b] * sizeof(unsigned long long)));
            h->scr_cap[b] = need[b];
        }
    RollArgs ra; memset(&ra, 0, sizeof(ra));
    LaneArgs &la = ra.la;
    la.proto = proto; la.policy = policy; la.obs_kind = obs_kind;
    la.action = action_dev; la.act_out = act_out_dev; la.reward = reward_dev; la.done = done_dev;
    la.plen = plen_dev; la.slen = slen_dev; la.obs = (uint16_t *)obs_dev; la.next_obs = (uint16_t *)next_obs_dev;
    la.rec_elems = obs_elems(h, obs_kind);
    ra.K = n_ticks;
    ra.heavy_bfs = ra.light_bfs = -1;   // per map class and launch shape (agv_rollout.cu: launch_shape)
    ra.scr_turn = reinterpret_cast<ulonglong2 *>(h->d_scr[0]); ra.scr_next = h->d_scr[1];
    ra.qctr = h->d_qctr;
    ra.stat = h->h_rstat;  // (unified addressing: the pinned host pointer is valid on the device)
#ifdef AGV_ROLL_PROF
    if (getenv("AGV_ROLL_PROF")) {
        if (!h->d_prof) CUDA_TRY(cudaMalloc(&h->d_prof, sizeof(unsigned long long) * 8 * (size_t)d.E));
        CUDA_TRY(cudaMemsetAsync(h->d_prof, 0, sizeof(unsigned long long) * 8 * (size_t)d.E, st));
        ra.prof = h->d_prof;
    }
#endif
    ra.vec_ok = (d.N % 8 == 0) && (((uintptr_t)obs_dev | (uintptr_t)next_obs_dev) & 15) == 0;
    // per-AGV outputs default to "no turn" (the kernel only writes the turns that happen)
    if (act_out_dev) CUDA_TRY(cudaMemsetAsync(act_out_dev, 0xFF, n, st));
    if (done_dev) CUDA_TRY(cudaMemsetAsync(done_dev, 0, n, st));
    if (reward_dev) CUDA_TRY(cudaMemsetAsync(reward_dev, 0, n * sizeof(float), st));
    if (plen_dev) CUDA_TRY(cudaMemsetAsync(plen_dev, 0, n * sizeof(uint16_t), st));
    if (slen_dev) CUDA_TRY(cudaMemsetAsync(slen_dev, 0xFF, n * sizeof(uint16_t), st));
    const int rc = lanes_rollout(d, h->NW, h->RPL, ra, h->d_order, st);
    if (rc == 0) g_tick_kernel = "k_rollout (persistent lane-per-scene rollout, longest-first scene queue)";
    return rc;
}

extern "C" {

int agv_create(const AgvConfig *cfg, const uint8_t *grid_host, AgvHandle **out) {
    if (!cfg || !grid_host || !out) return AGV_E_INVALID;
    if (cfg->E < 1 || cfg->N < 1 || cfg->N > 64 || cfg->T < 0 || cfg->T > 65535 || cfg->P < 0 || cfg->P > 15)
        return AGV_E_INVALID;
    int NW, RPL;
    int rc = pick_variant(cfg->H, cfg->W, &NW, &RPL);
    if (rc != AGV_OK) return rc;
    for (int i = 0; i < cfg->H * cfg->W; i++)
        if (grid_host[i] > 2) return AGV_E_BADGRID;
    CUDA_TRY(cudaSetDevice(cfg->device));
    AgvHandle *h = new (std::nothrow) AgvHandle();
    if (!h) return AGV_E_NOMEM;
    h->cfg = *cfg; h->NW = NW; h->RPL = RPL;
    DevCfg &d = h->dc;
    memset(&d, 0, sizeof(d));
    d.H = cfg->H; d.W = cfg->W; d.E = cfg->E; d.N = cfg->N; d.T = cfg->T; d.P = cfg->P; d.K = 2 * cfg->P + 1;
    d.always_loaded = cfg->always_loaded; d.always_empty = cfg->always_empty; d.shaped = cfg->shaped_reward;
    d.auto_reset = cfg->auto_reset; d.max_steps = cfg->max_steps;
    const int N = cfg->N;
    d.off_pos = HDR_BYTES; d.off_tgt = d.off_pos + 2 * N; d.off_ord = d.off_tgt + 2 * N; d.off_meta = d.off_ord + 2 * N;
    d.env_stride = align16(d.off_meta + N);
    const size_t HW = (size_t)cfg->H * cfg->W;
    std::vector<uint64_t> rows(3 * (size_t)cfg->H * NW, 0ull);
    for (int r = 0; r < cfg->H; r++)
        for (int cc = 0; cc < cfg->W; cc++) {
            const int code = grid_host[r * cfg->W + cc];
            const uint64_t bit = 1ull << (cc & 63);
            const size_t idx = (size_t)r * NW + (cc >> 6);
            if (code == 1) rows[idx] |= bit;
            if (code == 2) rows[(size_t)cfg->H * NW + idx] |= bit;
            rows[2 * (size_t)cfg->H * NW + idx] |= bit;
        }
#define CREATE_TRY(expr)                                                   \
    do {                                                                   \
        cudaError_t _e = (expr);                                           \
        if (_e != cudaSuccess) { g_last_cuda = (int)_e; agv_destroy(h); return AGV_E_CUDA; } \
    } while (0)
    CREATE_TRY(cudaMalloc(&h->d_state, (size_t)cfg->E * d.env_stride));
    CREATE_TRY(cudaMalloc(&h->d_tasks, sizeof(uint32_t) * (size_t)cfg->E * (cfg->T > 0 ? cfg->T : 1)));
    CREATE_TRY(cudaMemset(h->d_tasks, 0, sizeof(uint32_t) * (size_t)cfg->E * (cfg->T > 0 ? cfg->T : 1)));
    CREATE_TRY(cudaMalloc(&h->d_grid, HW));
    CREATE_TRY(cudaMalloc(&h->d_rows, rows.size() * 8));
    CREATE_TRY(cudaMalloc(&h->d_counters, 48 * sizeof(unsigned long long)));
    CREATE_TRY(cudaMemset(h->d_counters, 0, 48 * sizeof(unsigned long long)));
    CREATE_TRY(cudaMemcpy(h->d_grid, grid_host, HW, cudaMemcpyHostToDevice));
    CREATE_TRY(cudaMemcpy(h->d_rows, rows.data(), rows.size() * 8, cudaMemcpyHostToDevice));
    d.state = h->d_state; d.tasks = h->d_tasks; d.grid = h->d_grid;
    d.storage_rows = h->d_rows; d.picking_rows = h->d_rows + (size_t)cfg->H * NW;
    d.inb_rows = h->d_rows + 2 * (size_t)cfg->H * NW;
    d.counters = h->d_counters;
    k_init_state<<<(cfg->E + 127) / 128, 128>>>(d);
    g_launches.fetch_add(1);
    CREATE_TRY(cudaGetLastError());
    CREATE_TRY(cudaDeviceSynchronize());
    *out = h;
    return AGV_OK;
}

int agv_destroy(AgvHandle *h) {
    if (!h) return AGV_OK;
    cudaSetDevice(h->cfg.device);
    cudaFree(h->d_state); cudaFree(h->d_tasks); cudaFree(h->d_grid); cudaFree(h->d_rows); cudaFree(h->d_counters);
    for (int b = 0; b < 2; b++) {
        cudaFree(h->d_action[b]); cudaFree(h->d_act_out[b]); cudaFree(h->d_reward[b]); cudaFree(h->d_done[b]);
        if (h->ev_tick[b]) cudaEventDestroy(h->ev_tick[b]);
        if (h->ev_copy[b]) cudaEventDestroy(h->ev_copy[b]);
        if (h->ev_h2d[b]) cudaEventDestroy(h->ev_h2d[b]);
    }
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->h2d_stream) cudaStreamDestroy(h->h2d_stream);
    cudaFree(h->d_order); cudaFree(h->d_qctr); cudaFree(h->d_prof);
    if (h->h_rstat) cudaFreeHost(h->h_rstat);
    for (int b = 0; b < 2; b++) {
        cudaFree(h->d_scr[b]);
        cudaFree(h->r_action[b]); cudaFree(h->r_act_out[b]); cudaFree(h->r_reward[b]); cudaFree(h->r_done[b]);
    }
    delete h;
    return AGV_OK;
}

int agv_set_tasks(AgvHandle *h, const uint8_t *tasks_host, int32_t env_begin, int32_t env_count, void *stream) {
    if (!h || !tasks_host || env_begin < 0 || env_count < 0 || env_begin + env_count > h->cfg.E) return AGV_E_INVALID;
    if (h->cfg.T == 0 || env_count == 0) return AGV_OK;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    const size_t n = (size_t)env_count * h->cfg.T;
    for (size_t i = 0; i < n; i++) {
        const uint8_t *t = tasks_host + 4 * i;
        if (t[0] < 1 || t[0] > h->cfg.W || t[2] < 1 || t[2] > h->cfg.W || t[1] < 1 || t[1] > h->cfg.H || t[3] < 1 ||
            t[3] > h->cfg.H)
            return AGV_E_INVALID;
    }
    // [.,4] uint8 (sx,sy,px,py) is bit-identical to the packed little-endian uint32
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemcpyAsync(h->d_tasks + (size_t)env_begin * h->cfg.T, tasks_host, n * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return AGV_OK;
}

int agv_reset(AgvHandle *h, const uint8_t *env_mask_dev, void *stream) {
    if (!h) return AGV_E_INVALID;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    StepArgs s; memset(&s, 0, sizeof(s));
    plan_step(h, s, 0);
#define CALL(NW_, RPL_) return launch_warps(k_reset<NW_, RPL_>, h->cfg.E, s.per_warp, (cudaStream_t)stream, h->dc, s, env_mask_dev)
    DISPATCH(h, CALL);
#undef CALL
}

// n_ticks fused ticks in one launch: only the warp-per-scene kernel loops over ticks itself (outputs advance by
// E * N records per tick); the lane-per-scene kernels return AGV_E_UNSUPPORTED for n_ticks > 1
static int tick_n(AgvHandle *h, int n_ticks, int32_t proto, int32_t policy, int32_t obs_kind, const int8_t *action_dev,
                  int8_t *act_out_dev, float *reward_dev, uint8_t *done_dev, uint16_t *plen_dev, uint16_t *slen_dev,
                  void *obs_dev, void *next_obs_dev, void *stream) {
    if (!h || proto < 0 || proto > 1 || policy < 0 || policy > 2 || obs_kind < 0 || obs_kind > 2) return AGV_E_INVALID;
    if (obs_kind == AGV_OBS_NONE && (obs_dev || next_obs_dev)) return AGV_E_INVALID;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    TickArgs t; memset(&t, 0, sizeof(t));
    t.proto = proto; t.policy = policy; t.obs_kind = (obs_dev || next_obs_dev) ? obs_kind : AGV_OBS_NONE;
    t.action = action_dev; t.act_out = act_out_dev; t.reward = reward_dev; t.done = done_dev;
    t.plen = plen_dev; t.slen = slen_dev;
    t.obs = (uint16_t *)obs_dev; t.next_obs = (uint16_t *)next_obs_dev;
    const int N = h->cfg.N;
    t.rec_elems = obs_elems(h, t.obs_kind);
    const int impl = tick_impl(h, t.obs_kind);
    // (several ticks per launch pay on small maps, where a tick is launch-sized: README 9x9, 4096 scenes, 32 ticks
    // per call 5.6e8 -> 6.7e8 AGV-steps/s; the HBM-bound full-map encode at 64 x 64 loses 2 % to the longer waves)
    if (n_ticks != 1 && (impl > 0 || !(h->NW == 1 && h->RPL == 1))) return AGV_E_UNSUPPORTED;
    if (impl == 3) {
        const int rc = rollout_launch(h, 1, proto, policy, t.obs_kind, action_dev, act_out_dev, reward_dev, done_dev,
                                      plen_dev, slen_dev, obs_dev, next_obs_dev, (cudaStream_t)stream);
        if (rc != AGV_E_UNSUPPORTED) return rc;
    }
    if (impl > 0) {
        LaneArgs la; memset(&la, 0, sizeof(la));
        la.proto = t.proto; la.policy = t.policy; la.obs_kind = t.obs_kind;
        la.action = t.action; la.act_out = t.act_out; la.reward = t.reward; la.done = t.done;
        la.plen = t.plen; la.slen = t.slen; la.obs = t.obs; la.next_obs = t.next_obs; la.rec_elems = t.rec_elems;
        const int rc = lanes_tick(h->dc, h->NW, h->RPL, la, (cudaStream_t)stream, impl >= 2);
        if (rc != AGV_E_UNSUPPORTED) {
            const bool as = impl >= 2 && !h->dc.shaped && (t.obs_kind == AGV_OBS_NONE || (t.obs_kind == AGV_OBS_CLIP && h->dc.K == 7));
            g_tick_kernel = as ? "k_tick_async (lane-per-scene, decoupled lanes + BFS worker warps)"
                               : "k_tick_lanes (lane-per-scene, lock step)";
            return rc;
        }  // scenes too large for a lane-per-scene warp: warp-per-scene path
    }
    g_tick_kernel = (h->NW == 1 && h->RPL == 1) ? "k_tick_small (warp-per-scene, tick loop inside the launch)" : "k_tick (warp-per-scene)";
    t.group = 1;
    if (t.obs_kind == AGV_OBS_CLIP) {
        const int rb = t.rec_elems * 2;
        int g = 16 / gcd_i(rb, 16);
        // vector flush needs every group start 16-byte aligned: N % g == 0
        t.group = (N % g == 0) ? g : 1;
    }
    int off = 0;
    t.off_blk = off; off += align16(h->dc.env_stride);
    t.off_vrow = off; off += align16(h->cfg.H * h->NW * 8);
    t.off_act = off; off += align16(N);
    t.off_done = off; off += align16(N);
    t.off_plen = off; off += align16(2 * N);
    t.off_slen = off; off += align16(2 * N);
    t.off_rew = off; off += align16(4 * N);
    const int stage_bytes = (t.obs_kind == AGV_OBS_CLIP) ? align16(t.group * t.rec_elems * 2) : 0;
    t.off_stage = off; off += t.obs ? stage_bytes : 0;
    t.off_stage2 = off; off += t.next_obs ? stage_bytes : 0;
    t.per_warp = off;
#define CALL_OS(NW_, RPL_, OBS_, SH_)                                                                                   \
    do {                                                                                                                \
        if (NW_ == 1 && RPL_ == 1)                                                                                      \
            return launch_warps(k_tick_small<1, 1, OBS_, SH_>, h->cfg.E, t.per_warp, (cudaStream_t)stream, h->dc, t, n_ticks); \
        return launch_warps(k_tick<NW_, RPL_, OBS_, SH_>, h->cfg.E, t.per_warp, (cudaStream_t)stream, h->dc, t);        \
    } while (0)
#define CALL(NW_, RPL_)                                                              \
    do {                                                                             \
        const bool sh = h->cfg.shaped_reward != 0;                                   \
        if (t.obs_kind == AGV_OBS_CLIP) { if (sh) CALL_OS(NW_, RPL_, OBS_CLIP, true); else CALL_OS(NW_, RPL_, OBS_CLIP, false); } \
        else if (t.obs_kind == AGV_OBS_FULL) { if (sh) CALL_OS(NW_, RPL_, OBS_FULL, true); else CALL_OS(NW_, RPL_, OBS_FULL, false); } \
        else { if (sh) CALL_OS(NW_, RPL_, OBS_NONE, true); else CALL_OS(NW_, RPL_, OBS_NONE, false); } \
    } while (0)
    DISPATCH(h, CALL);
#undef CALL
#undef CALL_OS
}

int agv_tick(AgvHandle *h, int32_t proto, int32_t policy, int32_t obs_kind, const int8_t *action_dev,
             int8_t *act_out_dev, float *reward_dev, uint8_t *done_dev, uint16_t *plen_dev, uint16_t *slen_dev,
             void *obs_dev, void *next_obs_dev, void *stream) {
    return tick_n(h, 1, proto, policy, obs_kind, action_dev, act_out_dev, reward_dev, done_dev, plen_dev, slen_dev,
                  obs_dev, next_obs_dev, stream);
}

int agv_tick_host(AgvHandle *h, int32_t proto, int32_t policy, int32_t obs_kind, const int8_t *action_host,
                  int8_t *act_out_host, float *reward_host, uint8_t *done_host, void *obs_dev, void *next_obs_dev,
                  void *stream) {
    if (!h) return AGV_E_INVALID;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    const size_t n = (size_t)h->cfg.E * h->cfg.N;
    if (!h->d_action[0]) {
        for (int b = 0; b < 2; b++) {
            CUDA_TRY(cudaMalloc(&h->d_action[b], n));
            CUDA_TRY(cudaMalloc(&h->d_act_out[b], n));
            CUDA_TRY(cudaMalloc(&h->d_reward[b], n * sizeof(float)));
            CUDA_TRY(cudaMalloc(&h->d_done[b], n));
        }
    }
    CUDA_TRY(host_streams(h));
    cudaStream_t st = (cudaStream_t)stream;
    const int b = h->host_calls & 1;
    // the tick overwrites result buffer b: the copy-back of the call before last must be out
    if (h->copy_pending[b]) CUDA_TRY(cudaStreamWaitEvent(st, h->ev_copy[b], 0));
    if (policy == AGV_POLICY_GIVEN) {
        if (!action_host) return AGV_E_INVALID;
        // upload on its own stream so that it overlaps the previous call's tick; only ordered after
        // the tick that last read buffer b (action_host must hold its final contents at call time)
        if (h->tick_pending[b]) CUDA_TRY(cudaStreamWaitEvent(h->h2d_stream, h->ev_tick[b], 0));
        CUDA_TRY(cudaMemcpyAsync(h->d_action[b], action_host, n, cudaMemcpyHostToDevice, h->h2d_stream));
        CUDA_TRY(cudaEventRecord(h->ev_h2d[b], h->h2d_stream));
        CUDA_TRY(cudaStreamWaitEvent(st, h->ev_h2d[b], 0));
    }
    int rc = agv_tick(h, proto, policy, obs_kind, policy == AGV_POLICY_GIVEN ? h->d_action[b] : nullptr, h->d_act_out[b],
                      h->d_reward[b], h->d_done[b], nullptr, nullptr, obs_dev, next_obs_dev, stream);
    if (rc != AGV_OK) return rc;
    CUDA_TRY(cudaEventRecord(h->ev_tick[b], st));
    h->tick_pending[b] = true;
    CUDA_TRY(cudaStreamWaitEvent(h->copy_stream, h->ev_tick[b], 0));
    if (act_out_host) CUDA_TRY(cudaMemcpyAsync(act_out_host, h->d_act_out[b], n, cudaMemcpyDeviceToHost, h->copy_stream));
    if (reward_host) CUDA_TRY(cudaMemcpyAsync(reward_host, h->d_reward[b], n * sizeof(float), cudaMemcpyDeviceToHost, h->copy_stream));
    if (done_host) CUDA_TRY(cudaMemcpyAsync(done_host, h->d_done[b], n, cudaMemcpyDeviceToHost, h->copy_stream));
    CUDA_TRY(cudaEventRecord(h->ev_copy[b], h->copy_stream));
    h->copy_pending[b] = true;
    h->host_calls++;
    return AGV_OK;
}

int agv_tick_host_join(AgvHandle *h, void *stream) {
    if (!h) return AGV_E_INVALID;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    for (int b = 0; b < 2; b++)
        if (h->copy_pending[b]) CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)stream, h->ev_copy[b], 0));
    return AGV_OK;
}

int agv_rollout(AgvHandle *h, int32_t n_ticks, int32_t proto, int32_t policy, int32_t obs_kind,
                const int8_t *action_dev, int8_t *act_out_dev, float *reward_dev, uint8_t *done_dev,
                uint16_t *plen_dev, uint16_t *slen_dev, void *obs_dev, void *next_obs_dev, void *stream) {
    if (!h || n_ticks < 0 || proto < 0 || proto > 1 || policy < 0 || policy > 2 || obs_kind < 0 || obs_kind > 2)
        return AGV_E_INVALID;
    if (obs_kind == AGV_OBS_NONE && (obs_dev || next_obs_dev)) return AGV_E_INVALID;
    if (n_ticks == 0) return AGV_OK;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    const int ok = (obs_dev || next_obs_dev) ? obs_kind : AGV_OBS_NONE;
    const size_t per = (size_t)h->cfg.E * h->cfg.N, R = (size_t)obs_elems(h, ok);
    const char *e = getenv("AGV_ROLLOUT_IMPL");  // "ticks" forces the per-tick fallback (A/B runs, tests)
    if (!(e && strcmp(e, "ticks") == 0)) {
        // Long rollouts are taken in chunks of at most 16 ticks per launch (AGV_ROLLOUT_CHUNK): it bounds the
        // per-turn scratch, and every launch re-sorts the scenes longest-first and re-deals them over the CTAs.
        // Measured at c3 in steady state, 16 ticks: 16 per launch 1.175e9, 4: 1.167e9, 2: 1.12e9, 1: 1.05e9
        // AGV-steps/s.  Same results for every chunk size.
        static const int chunk_env = getenv("AGV_ROLLOUT_CHUNK") ? atoi(getenv("AGV_ROLLOUT_CHUNK")) : 0;
        const int chunk = chunk_env > 0 ? chunk_env : 16;
        int rc = AGV_OK;
        for (int k0 = 0; k0 < n_ticks && rc == AGV_OK; k0 += chunk) {
            const size_t o = per * (size_t)k0;
            rc = rollout_launch(h, n_ticks - k0 < chunk ? n_ticks - k0 : chunk, proto, policy, ok,
                                action_dev ? action_dev + o : nullptr, act_out_dev ? act_out_dev + o : nullptr,
                                reward_dev ? reward_dev + o : nullptr, done_dev ? done_dev + o : nullptr,
                                plen_dev ? plen_dev + o : nullptr, slen_dev ? slen_dev + o : nullptr,
                                obs_dev ? (void *)((uint16_t *)obs_dev + o * R) : nullptr,
                                next_obs_dev ? (void *)((uint16_t *)next_obs_dev + o * R) : nullptr, (cudaStream_t)stream);
            if (rc == AGV_E_UNSUPPORTED && k0 > 0) return AGV_E_CUDA;  // (cannot happen: support depends on the configuration only)
        }
        if (rc != AGV_E_UNSUPPORTED) return rc;
    }
    // configurations the persistent kernel does not cover (shaped reward, full-map observations, K != 7) and
    // small maps: the warp-per-scene kernel keeps every scene in its warp for all n_ticks (one launch) ...
    {
        static const bool per_tick = getenv("AGV_ROLLOUT_PER_TICK") != nullptr;   // (A/B runs)
        const int rc = per_tick ? AGV_E_UNSUPPORTED
                                : tick_n(h, n_ticks, proto, policy, obs_kind, action_dev, act_out_dev, reward_dev, done_dev,
                                         plen_dev, slen_dev, obs_dev, next_obs_dev, stream);
        if (rc != AGV_E_UNSUPPORTED) return rc;
    }
    // ... the lane-per-scene ticks run as n_ticks launches, each writing its slice of the rings
    for (int k = 0; k < n_ticks; k++) {
        const size_t o = per * (size_t)k;
        const int rc = agv_tick(h, proto, policy, obs_kind, action_dev ? action_dev + o : nullptr,
                                act_out_dev ? act_out_dev + o : nullptr, reward_dev ? reward_dev + o : nullptr,
                                done_dev ? done_dev + o : nullptr, plen_dev ? plen_dev + o : nullptr,
                                slen_dev ? slen_dev + o : nullptr, obs_dev ? (uint16_t *)obs_dev + o * R : nullptr,
                                next_obs_dev ? (uint16_t *)next_obs_dev + o * R : nullptr, stream);
        if (rc != AGV_OK) return rc;
    }
    return AGV_OK;
}

int agv_rollout_host(AgvHandle *h, int32_t n_ticks, int32_t proto, int32_t policy, int32_t obs_kind,
                     const int8_t *action_host, int8_t *act_out_host, float *reward_host, uint8_t *done_host,
                     void *obs_dev, void *next_obs_dev, void *stream) {
    if (!h || n_ticks < 1) return AGV_E_INVALID;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    const size_t n = (size_t)h->cfg.E * h->cfg.N * (size_t)n_ticks;
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(host_streams(h));
    if (h->r_cap < n) {
        CUDA_TRY(cudaDeviceSynchronize());
        for (int b = 0; b < 2; b++) {
            cudaFree(h->r_action[b]); cudaFree(h->r_act_out[b]); cudaFree(h->r_reward[b]); cudaFree(h->r_done[b]);
            h->r_action[b] = h->r_act_out[b] = nullptr; h->r_reward[b] = nullptr; h->r_done[b] = nullptr;
            CUDA_TRY(cudaMalloc(&h->r_action[b], n));
            CUDA_TRY(cudaMalloc(&h->r_act_out[b], n));
            CUDA_TRY(cudaMalloc(&h->r_reward[b], n * sizeof(float)));
            CUDA_TRY(cudaMalloc(&h->r_done[b], n));
        }
        h->r_cap = n;
        h->copy_pending[0] = h->copy_pending[1] = false;
        h->tick_pending[0] = h->tick_pending[1] = false;
    }
    // Same double-buffered pipeline as agv_tick_host -- the upload of call k+1 and the copy-back of call k-1
    // overlap the rollout of call k -- and, inside one call, the rollout runs in chunks of a few ticks
    // (AGV_ROLLOUT_HOST_CHUNK, default 4) whose actions are uploaded and whose results are copied back while
    // the neighbouring chunks compute: a caller that waits for every call (closed loop) sees the last chunk's
    // copy-back only, not the whole ring's.
    static const int hc_env = getenv("AGV_ROLLOUT_HOST_CHUNK") ? atoi(getenv("AGV_ROLLOUT_HOST_CHUNK")) : 0;
    const int hc = hc_env > 0 ? hc_env : 4;
    const size_t per = (size_t)h->cfg.E * h->cfg.N;
    const int ok = (obs_dev || next_obs_dev) ? obs_kind : AGV_OBS_NONE;
    const size_t R = (size_t)obs_elems(h, ok);
    const int b = h->roll_calls & 1;
    if (h->copy_pending[b]) CUDA_TRY(cudaStreamWaitEvent(st, h->ev_copy[b], 0));
    if (policy == AGV_POLICY_GIVEN) {
        if (!action_host) return AGV_E_INVALID;
        if (h->tick_pending[b]) CUDA_TRY(cudaStreamWaitEvent(h->h2d_stream, h->ev_tick[b], 0));
    }
    for (int k0 = 0; k0 < n_ticks; k0 += hc) {
        const int kc = n_ticks - k0 < hc ? n_ticks - k0 : hc;
        const size_t o = per * (size_t)k0, nc = per * (size_t)kc;
        if (policy == AGV_POLICY_GIVEN) {
            CUDA_TRY(cudaMemcpyAsync(h->r_action[b] + o, action_host + o, nc, cudaMemcpyHostToDevice, h->h2d_stream));
            CUDA_TRY(cudaEventRecord(h->ev_h2d[b], h->h2d_stream));
            CUDA_TRY(cudaStreamWaitEvent(st, h->ev_h2d[b], 0));
        }
        const int rc = agv_rollout(h, kc, proto, policy, obs_kind, policy == AGV_POLICY_GIVEN ? h->r_action[b] + o : nullptr,
                                   h->r_act_out[b] + o, h->r_reward[b] + o, h->r_done[b] + o, nullptr, nullptr,
                                   obs_dev ? (void *)((uint16_t *)obs_dev + o * R) : nullptr,
                                   next_obs_dev ? (void *)((uint16_t *)next_obs_dev + o * R) : nullptr, stream);
        if (rc != AGV_OK) return rc;
        CUDA_TRY(cudaEventRecord(h->ev_tick[b], st));
        CUDA_TRY(cudaStreamWaitEvent(h->copy_stream, h->ev_tick[b], 0));
        if (act_out_host) CUDA_TRY(cudaMemcpyAsync(act_out_host + o, h->r_act_out[b] + o, nc, cudaMemcpyDeviceToHost, h->copy_stream));
        if (reward_host) CUDA_TRY(cudaMemcpyAsync(reward_host + o, h->r_reward[b] + o, nc * sizeof(float), cudaMemcpyDeviceToHost, h->copy_stream));
        if (done_host) CUDA_TRY(cudaMemcpyAsync(done_host + o, h->r_done[b] + o, nc, cudaMemcpyDeviceToHost, h->copy_stream));
    }
    h->tick_pending[b] = true;
    CUDA_TRY(cudaEventRecord(h->ev_copy[b], h->copy_stream));
    h->copy_pending[b] = true;
    h->roll_calls++;
    return AGV_OK;
}

int agv_begin_tick(AgvHandle *h, void *stream) {
    if (!h) return AGV_E_INVALID;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    StepArgs s; memset(&s, 0, sizeof(s));
    plan_step(h, s, 0);
#define CALL(NW_, RPL_) return launch_warps(k_begin_tick<NW_, RPL_>, h->cfg.E, s.per_warp, (cudaStream_t)stream, h->dc, s)
    DISPATCH(h, CALL);
#undef CALL
}

int agv_end_tick(AgvHandle *h, void *stream) {
    if (!h) return AGV_E_INVALID;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    StepArgs s; memset(&s, 0, sizeof(s));
    plan_step(h, s, 0);
#define CALL(NW_, RPL_) return launch_warps(k_end_tick<NW_, RPL_>, h->cfg.E, s.per_warp, (cudaStream_t)stream, h->dc, s)
    DISPATCH(h, CALL);
#undef CALL
}

int agv_encode(AgvHandle *h, int32_t k, int32_t obs_kind, int32_t post, void *out_dev, uint8_t *will_act_dev,
               void *stream) {
    if (!h || k < 0 || k >= h->cfg.N || obs_kind < 1 || obs_kind > 2) return AGV_E_INVALID;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    StepArgs s; memset(&s, 0, sizeof(s));
    plan_step(h, s, obs_elems(h, obs_kind));
    s.k = k; s.obs_kind = obs_kind; s.post = post; s.obs = (uint16_t *)out_dev; s.acted = will_act_dev;
#define CALL(NW_, RPL_) return launch_warps(k_encode<NW_, RPL_>, h->cfg.E, s.per_warp, (cudaStream_t)stream, h->dc, s)
    DISPATCH(h, CALL);
#undef CALL
}

int agv_bfs_action(AgvHandle *h, int32_t k, int32_t matrix, int8_t *action_dev, uint16_t *pathlen_dev, void *stream) {
    if (!h || k < 0 || k >= h->cfg.N || matrix < 0 || matrix > 1) return AGV_E_INVALID;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    StepArgs s; memset(&s, 0, sizeof(s));
    plan_step(h, s, 0);
    s.k = k; s.matrix = matrix; s.action_out = action_dev; s.plen = pathlen_dev;
#define CALL(NW_, RPL_) return launch_warps(k_bfs_action<NW_, RPL_>, h->cfg.E, s.per_warp, (cudaStream_t)stream, h->dc, s)
    DISPATCH(h, CALL);
#undef CALL
}

int agv_step_turn(AgvHandle *h, int32_t k, int32_t proto, const int8_t *action_dev, float *reward_dev,
                  uint8_t *done_dev, uint8_t *acted_dev, uint16_t *slen_dev, int8_t *act_out_dev, void *stream) {
    if (!h || k < 0 || k >= h->cfg.N || proto < 0 || proto > 1) return AGV_E_INVALID;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    StepArgs s; memset(&s, 0, sizeof(s));
    plan_step(h, s, 0);
    s.k = k; s.proto = proto; s.action = action_dev; s.reward = reward_dev; s.done = done_dev; s.acted = acted_dev;
    s.slen = slen_dev; s.action_out = act_out_dev;
#define CALL(NW_, RPL_) return launch_warps(k_step_turn<NW_, RPL_>, h->cfg.E, s.per_warp, (cudaStream_t)stream, h->dc, s)
    DISPATCH(h, CALL);
#undef CALL
}

int agv_get_state(AgvHandle *h, int64_t *hdr_host, int32_t *agv_host, void *stream) {
    if (!h) return AGV_E_INVALID;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    const DevCfg &d = h->dc;
    h->h_state.resize((size_t)d.E * d.env_stride);
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemcpyAsync(h->h_state.data(), h->d_state, h->h_state.size(), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (int e = 0; e < d.E; e++) {
        const uint8_t *blk = h->h_state.data() + (size_t)e * d.env_stride;
        EnvHdr hd; memcpy(&hd, blk, sizeof(hd));
        if (hdr_host) {
            int64_t *o = hdr_host + 6 * (size_t)e;
            o[0] = hd.tick; o[1] = hd.cursor; o[2] = hd.n_done; o[3] = hd.finished; o[4] = hd.n_created; o[5] = hd.over;
        }
        if (agv_host) {
            const uint16_t *pos = (const uint16_t *)(blk + d.off_pos), *tgt = (const uint16_t *)(blk + d.off_tgt),
                           *ord = (const uint16_t *)(blk + d.off_ord);
            const uint8_t *meta = blk + d.off_meta;
            for (int j = 0; j < d.N; j++) {
                int32_t *o = agv_host + 8 * ((size_t)e * d.N + j);
                o[0] = pos[j] & 255; o[1] = pos[j] >> 8; o[2] = tgt[j] & 255; o[3] = tgt[j] >> 8;
                o[4] = meta[j] & 3; o[5] = (meta[j] >> 2) & 1; o[6] = (meta[j] >> 3) & 1; o[7] = ord[j];
            }
        }
    }
    return AGV_OK;
}

int agv_set_state(AgvHandle *h, const int64_t *hdr_host, const int32_t *agv_host, void *stream) {
    if (!h || !hdr_host || !agv_host) return AGV_E_INVALID;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    const DevCfg &d = h->dc;
    h->h_state.assign((size_t)d.E * d.env_stride, 0);
    for (int e = 0; e < d.E; e++) {
        uint8_t *blk = h->h_state.data() + (size_t)e * d.env_stride;
        const int64_t *i6 = hdr_host + 6 * (size_t)e;
        // the kernels index pos[j] for j < n_created and tasks[order]: a snapshot outside these ranges would
        // read out of bounds instead of failing
        if (i6[0] < 0 || i6[1] < 0 || i6[1] > d.T || i6[2] < 0 || i6[2] > d.T || i6[4] < 0 || i6[4] > d.N)
            return AGV_E_INVALID;
        EnvHdr hd; memset(&hd, 0, sizeof(hd));
        hd.tick = (uint32_t)i6[0]; hd.cursor = (uint32_t)i6[1]; hd.n_done = (uint32_t)i6[2];
        hd.finished = (uint8_t)i6[3]; hd.n_created = (uint16_t)i6[4]; hd.over = (uint8_t)i6[5];
        memcpy(blk, &hd, sizeof(hd));
        uint16_t *pos = (uint16_t *)(blk + d.off_pos), *tgt = (uint16_t *)(blk + d.off_tgt),
                 *ord = (uint16_t *)(blk + d.off_ord);
        uint8_t *meta = blk + d.off_meta;
        for (int j = 0; j < d.N; j++) {
            const int32_t *a = agv_host + 8 * ((size_t)e * d.N + j);
            if (a[0] < 1 || a[0] > d.W || a[1] < 1 || a[1] > d.H || a[2] < 1 || a[2] > d.W || a[3] < 1 || a[3] > d.H)
                return AGV_E_INVALID;
            if (a[4] < 0 || a[4] > 2 || a[7] < 0 || a[7] >= (d.T > 0 ? d.T : 1)) return AGV_E_INVALID;
            pos[j] = (uint16_t)(a[0] | (a[1] << 8)); tgt[j] = (uint16_t)(a[2] | (a[3] << 8));
            meta[j] = (uint8_t)((a[4] & 3) | (a[5] ? 4 : 0) | (a[6] ? 8 : 0)); ord[j] = (uint16_t)a[7];
        }
    }
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemcpyAsync(h->d_state, h->h_state.data(), h->h_state.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return AGV_OK;
}

int agv_read_counters(AgvHandle *h, uint64_t out[3], void *stream) {
    if (!h || !out) return AGV_E_INVALID;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long tmp[48];  // [3..31]: phase cycle counters of profiling builds (-DAGV_LANES_PROF)
    CUDA_TRY(cudaMemcpyAsync(tmp, h->d_counters, sizeof(tmp), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemsetAsync(h->d_counters, 0, sizeof(tmp), st));
    CUDA_TRY(cudaStreamSynchronize(st));
    static const bool dump = getenv("AGV_PROF_DUMP") != nullptr;  // profiling builds only; read once
    if (dump) {
        fprintf(stderr, "agv prof:");
        for (int i = 3; i < 48; i++) fprintf(stderr, " %llu", tmp[i]);
        fprintf(stderr, "\n");
    }
    out[0] = tmp[0]; out[1] = tmp[1]; out[2] = tmp[2];
    return AGV_OK;
}

int agv_bfs_batch(int32_t B, int32_t H, int32_t W, const uint8_t *valid_dev, const int32_t *start_dev,
                  const int32_t *target_dev, int8_t *action_dev, int32_t *pathlen_dev, uint8_t *found_dev,
                  void *stream) {
    if (B < 0 || !valid_dev || !start_dev || !target_dev) return AGV_E_INVALID;
    if (B == 0) return AGV_OK;
    AgvHandle tmp; AgvHandle *h = &tmp;
    int rc = pick_variant(H, W, &h->NW, &h->RPL);
    if (rc != AGV_OK) return rc;
#define CALL(NW_, RPL_) return launch_warps(k_bfs_batch<NW_, RPL_>, B, 0, (cudaStream_t)stream, B, H, W, valid_dev, start_dev, target_dev, action_dev, pathlen_dev, found_dev)
    DISPATCH(h, CALL);
#undef CALL
}

int agv_bfs_field(int32_t B, int32_t H, int32_t W, const uint8_t *valid_dev, const int32_t *target_dev,
                  uint16_t *dist_dev, void *stream) {
    if (B < 0 || !valid_dev || !target_dev || !dist_dev) return AGV_E_INVALID;
    if (B == 0) return AGV_OK;
    AgvHandle tmp; AgvHandle *h = &tmp;
    int rc = pick_variant(H, W, &h->NW, &h->RPL);
    if (rc != AGV_OK) return rc;
#define CALL(NW_, RPL_) return launch_warps(k_bfs_field<NW_, RPL_>, B, 0, (cudaStream_t)stream, B, H, W, valid_dev, target_dev, dist_dev)
    DISPATCH(h, CALL);
#undef CALL
}

int agv_rollout_profile(AgvHandle *h, uint64_t *out_host, void *stream) {
    if (!h || !out_host) return AGV_E_INVALID;
    if (!h->d_prof) return AGV_E_UNSUPPORTED;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    CUDA_TRY(cudaMemcpyAsync(out_host, h->d_prof, sizeof(unsigned long long) * 8 * (size_t)h->cfg.E,
                             cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return AGV_OK;
}

const char *agv_strerror(int status) {
    switch (status) {
        case AGV_OK: return "ok";
        case AGV_E_INVALID: return "invalid argument";
        case AGV_E_CUDA: return "CUDA runtime error";
        case AGV_E_NOMEM: return "out of memory";
        case AGV_E_UNSUPPORTED: return "size outside the compiled kernel range (H, W <= 128)";
        case AGV_E_BADGRID: return "grid code outside {0,1,2}";
        default: return "unknown status";
    }
}

int agv_last_cuda_error(void) { return g_last_cuda; }
uint64_t agv_launch_count(void) { return g_launches.load(); }
const char *agv_version(void) { return "agv_b200 0.2 (sm_100a)"; }
const char *agv_last_tick_kernel(void) { return g_tick_kernel; }

}  // extern "C"

namespace agv {
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
void note_cuda_error(int e) { g_last_cuda = e; }
}  // namespace agv
