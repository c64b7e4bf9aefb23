// agv_core.cuh -- warp-cooperative device core of the batched multi-AGV scene.
//
// One warp owns one scene (environment) for the duration of a kernel: the scene's
// state block is staged in shared memory, the grid lives in registers as one bit
// per cell (row r is owned by lane r / RPL, slot r % RPL, NW 64-bit words per
// row), and the reference's turn order (Scene.py:124-162) is a sequential loop
// inside the warp.  BFS ("A*", astar.py:73-153) is a bit-parallel, level-
// synchronous frontier expansion: one level = 2 row shuffles + shifts/ORs, no
// shared memory and no block barrier.
//
// Reference citations (file:line) are relative to
// stoneMan349/Reinforcement-learning-for-multi-AGV-pathfinding.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace agv {

constexpr unsigned FULL = 0xffffffffu;
constexpr int A_UP = 0, A_RIGHT = 1, A_DOWN = 2, A_LEFT = 3, A_STOP = 4;
constexpr int PROTO_ASTAR = 0, PROTO_INTELLIGENT = 1;
constexpr int POLICY_GIVEN = 0, POLICY_ASTAR = 1, POLICY_GUIDED = 2;
constexpr int OBS_NONE = 0, OBS_CLIP = 1, OBS_FULL = 2;
constexpr uint16_t BF16_ONE = 0x3F80;
constexpr int HDR_BYTES = 32;
#ifndef AGV_BFS_UNROLL
#define AGV_BFS_UNROLL 4  // BFS levels per loop body (= period of the empty-frontier test)
#endif

// Scene state header, first HDR_BYTES of a scene block.
struct EnvHdr {
    uint32_t tick;       // Scene.running_time
    uint32_t cursor;     // len(layout.task_arrangement[0])
    uint32_t n_done;     // sum(layout.task_arrangement[2])
    uint16_t n_created;  // created AGVs are the prefix [0, n_created)
    uint8_t finished;    // layout.task_finished
    uint8_t over;        // run_mode has returned
    uint64_t acted_mask; // AGVs that took a turn in the current tick (sub-step protocol)
    uint64_t reserved;
};
static_assert(sizeof(EnvHdr) == HDR_BYTES, "header layout");

struct DevCfg {
    int H, W, E, N, T, P, K;
    int always_loaded, always_empty, shaped, auto_reset;
    long long max_steps;
    int env_stride;                 // bytes per scene block (multiple of 16)
    int off_pos, off_tgt, off_ord, off_meta;  // byte offsets of the per-AGV arrays
    uint8_t *state;                 // [E * env_stride]
    const uint32_t *tasks;          // [E * T] sx | sy<<8 | px<<16 | py<<24
    const uint8_t *grid;            // [H * W] codes 0/1/2
    const uint64_t *storage_rows;   // [H * NW] bit c of row r = storage station
    const uint64_t *picking_rows;   // [H * NW]
    const uint64_t *inb_rows;       // [H * NW] columns < W
    unsigned long long *counters;   // [3] turns, episodes ended, invalid actions
};

// per-AGV meta byte: bits 0-1 task_stage, bit 2 loaded, bit 3 all_assigned (idle)
struct Agv {
    int x, y, tx, ty, stage, loaded, idle, order;
};

__device__ __forceinline__ int meta_pack(const Agv &a) {
    return (a.stage & 3) | (a.loaded ? 4 : 0) | (a.idle ? 8 : 0);
}

// ------------------------------------------------------------------ bit rows
template <int NW>
struct Row {
    uint64_t w[NW];
};

template <int NW>
__device__ __forceinline__ Row<NW> rzero() {
    Row<NW> r;
#pragma unroll
    for (int i = 0; i < NW; i++) r.w[i] = 0ull;
    return r;
}
template <int NW>
__device__ __forceinline__ Row<NW> ror(const Row<NW> &a, const Row<NW> &b) {
    Row<NW> r;
#pragma unroll
    for (int i = 0; i < NW; i++) r.w[i] = a.w[i] | b.w[i];
    return r;
}
template <int NW>
__device__ __forceinline__ Row<NW> rand_(const Row<NW> &a, const Row<NW> &b) {
    Row<NW> r;
#pragma unroll
    for (int i = 0; i < NW; i++) r.w[i] = a.w[i] & b.w[i];
    return r;
}
template <int NW>
__device__ __forceinline__ Row<NW> randn(const Row<NW> &a, const Row<NW> &b) {  // a & ~b
    Row<NW> r;
#pragma unroll
    for (int i = 0; i < NW; i++) r.w[i] = a.w[i] & ~b.w[i];
    return r;
}
template <int NW>
__device__ __forceinline__ bool rany(const Row<NW> &a) {
    uint64_t v = 0;
#pragma unroll
    for (int i = 0; i < NW; i++) v |= a.w[i];
    return v != 0ull;
}
// column c+1 <- column c
template <int NW>
__device__ __forceinline__ Row<NW> rshl1(const Row<NW> &a) {
    Row<NW> r;
#pragma unroll
    for (int i = NW - 1; i >= 0; i--) r.w[i] = (a.w[i] << 1) | (i > 0 ? (a.w[i - 1] >> 63) : 0ull);
    return r;
}
template <int NW>
__device__ __forceinline__ Row<NW> rshr1(const Row<NW> &a) {
    Row<NW> r;
#pragma unroll
    for (int i = 0; i < NW; i++) r.w[i] = (a.w[i] >> 1) | (i + 1 < NW ? (a.w[i + 1] << 63) : 0ull);
    return r;
}
template <int NW>
__device__ __forceinline__ Row<NW> rshfl_up(const Row<NW> &a) {
    Row<NW> r;
#pragma unroll
    for (int i = 0; i < NW; i++) r.w[i] = __shfl_up_sync(FULL, (unsigned long long)a.w[i], 1);
    return r;
}
template <int NW>
__device__ __forceinline__ Row<NW> rshfl_down(const Row<NW> &a) {
    Row<NW> r;
#pragma unroll
    for (int i = 0; i < NW; i++) r.w[i] = __shfl_down_sync(FULL, (unsigned long long)a.w[i], 1);
    return r;
}

// set bit (row r, col c) on the lane that owns row r
template <int NW, int RPL>
__device__ __forceinline__ void set_bit(Row<NW> (&rows)[RPL], int r, int c, int lane) {
    const int ol = r / RPL, os = r % RPL;
    const uint64_t bit = 1ull << (c & 63);
    const int wi = c >> 6;
#pragma unroll
    for (int s = 0; s < RPL; s++)
#pragma unroll
        for (int i = 0; i < NW; i++)
            if (lane == ol && s == os && i == wi) rows[s].w[i] |= bit;
}
template <int NW, int RPL>
__device__ __forceinline__ void clear_bit(Row<NW> (&rows)[RPL], int r, int c, int lane) {
    const int ol = r / RPL, os = r % RPL;
    const uint64_t bit = 1ull << (c & 63);
    const int wi = c >> 6;
#pragma unroll
    for (int s = 0; s < RPL; s++)
#pragma unroll
        for (int i = 0; i < NW; i++)
            if (lane == ol && s == os && i == wi) rows[s].w[i] &= ~bit;
}
// true only on the owner lane, when the bit is set
template <int NW, int RPL>
__device__ __forceinline__ bool test_bit_local(const Row<NW> (&rows)[RPL], int r, int c, int lane) {
    const int ol = r / RPL, os = r % RPL;
    const int wi = c >> 6;
    uint64_t v = 0;
#pragma unroll
    for (int s = 0; s < RPL; s++)
#pragma unroll
        for (int i = 0; i < NW; i++)
            if (s == os && i == wi) v = rows[s].w[i];
    return (lane == ol) && ((v >> (c & 63)) & 1ull);
}

// ------------------------------------------------------------------ BFS
// multiply both 32-bit halves by m in {0,1}: boundary masking on the FMA pipe (IMAD)
// instead of SEL on the (saturated) ALU pipe
template <int NW>
__device__ __forceinline__ Row<NW> rmask01(const Row<NW> &a, uint32_t m) {
    Row<NW> r;
#pragma unroll
    for (int i = 0; i < NW; i++) {
        const uint32_t lo = (uint32_t)a.w[i] * m, hi = (uint32_t)(a.w[i] >> 32) * m;
        r.w[i] = ((uint64_t)hi << 32) | lo;
    }
    return r;
}

// a & ~b for b a subset of a, as per-half a + b * 0xffffffff: mad.lo.u32 issues on the FMA
// pipe (IMAD), off-loading the ALU pipe that bounds the BFS level loop
template <int NW>
__device__ __forceinline__ Row<NW> rsub_subset(const Row<NW> &a, const Row<NW> &b) {
    Row<NW> r;
#pragma unroll
    for (int i = 0; i < NW; i++) {
        uint32_t lo, hi;
        asm("mad.lo.u32 %0, %1, 0xffffffff, %2;" : "=r"(lo) : "r"((uint32_t)b.w[i]), "r"((uint32_t)a.w[i]));
        asm("mad.lo.u32 %0, %1, 0xffffffff, %2;" : "=r"(hi) : "r"((uint32_t)(b.w[i] >> 32)), "r"((uint32_t)(a.w[i] >> 32)));
        r.w[i] = ((uint64_t)hi << 32) | lo;
    }
    return r;
}

// one BFS level: f <- neighbours(f) & av;  av <- av & ~f   (av = valid & ~visited)
// (Measured and dropped: an even/odd column layout of the 64-column row -- even' = odd | odd << 1, odd' = even |
// even >> 1, two 32-bit shifts instead of four funnel shifts -- takes the level from 28 to 22 instructions and
// the ALU-pipe share from 14.5 to 10, bit-exact, and changes the rollout by +0.5 % / -0.5 % (guided / A_star):
// a search is bound by the level-to-level chain SHFL -> LOP3 -> LOP3, not by the pipe.)
template <int NW, int RPL>
__device__ __forceinline__ void bfs_expand(Row<NW> (&f)[RPL], Row<NW> (&av)[RPL], uint32_t m_up, uint32_t m_dn) {
    // Lane 0 (31) gets its OWN register back from shfl_up (shfl_down).  For RPL <= 2 that
    // row is either the frontier row itself (f & av == 0) or a true vertical neighbour that
    // is OR-ed in anyway, so no boundary masking is needed; RPL > 2 masks explicitly.
    Row<NW> up_in = rshfl_up(f[RPL - 1]);
    Row<NW> dn_in = rshfl_down(f[0]);
    if (RPL > 2) { up_in = rmask01(up_in, m_up); dn_in = rmask01(dn_in, m_dn); }
    Row<NW> n[RPL];
#pragma unroll
    for (int s = 0; s < RPL; s++) {
        Row<NW> a = ror(rshl1(f[s]), rshr1(f[s]));
        const Row<NW> u = (s > 0) ? f[s > 0 ? s - 1 : 0] : up_in;
        const Row<NW> d = (s < RPL - 1) ? f[s < RPL - 1 ? s + 1 : 0] : dn_in;
        a = ror(a, u);
        n[s] = rand_(ror(a, d), av[s]);
    }
#pragma unroll
    for (int s = 0; s < RPL; s++) { av[s] = rsub_subset(av[s], n[s]); f[s] = n[s]; }
}

// Target-rooted bit-parallel BFS over `valid`; stops at the first level whose
// frontier touches a neighbour of the start cell.  Equivalent to
// FindPathAstar(valid, start, target).run_astar_method() (astar.py:66-153) as far
// as its callers use it (Explorer.py:286-302, MADQN.py:192-200): returns
// action_list[0] (STOP when not found or empty) and len(action_list) in `len`.
// The reference's open list is FIFO (all_cost is never assigned, astar.py:14,86)
// and re-parents on ties (astar.py:108-117), which makes the path the
// lexicographically largest shortest path under LEFT < UP < DOWN < RIGHT
// (neighbour order astar.py:33-38); hence the priority RIGHT, DOWN, UP, LEFT
// among start neighbours one step closer to the target.
// Coordinates are 0-based.  All lanes must call; the result is warp-uniform.
template <int NW, int RPL>
__device__ __forceinline__ int warp_bfs(const Row<NW> (&valid)[RPL], int sx, int sy, int tx, int ty, int H, int W,
                                        int lane, int &len) {
    len = 0;
    if (sx == tx && sy == ty) return A_STOP;  // found, empty action list (astar.py:90-92)
    Row<NW> f[RPL], av[RPL], nb[RPL];
#pragma unroll
    for (int s = 0; s < RPL; s++) f[s] = rzero<NW>();
    set_bit<NW, RPL>(f, ty, tx, lane);
#pragma unroll
    for (int s = 0; s < RPL; s++) { f[s] = rand_(f[s], valid[s]); av[s] = randn(valid[s], f[s]); }
    {   // the (up to) 4 neighbour cells of the start; columns outside the map are never valid
        Row<NW> ctr = rzero<NW>();
#pragma unroll
        for (int i = 0; i < NW; i++)
            if (i == (sx >> 6)) ctr.w[i] = 1ull << (sx & 63);
        const Row<NW> side = ror(rshl1(ctr), rshr1(ctr));
#pragma unroll
        for (int s = 0; s < RPL; s++) {
            const int dr = lane * RPL + s - sy;
            nb[s] = dr == 0 ? side : ((dr == 1 || dr == -1) ? ctr : rzero<NW>());
        }
    }
    const uint32_t m_up = lane != 0 ? 1u : 0u, m_dn = lane != 31 ? 1u : 0u;
    int level = 0;
    // Phase 1: a frontier cell at level L is exactly L steps from the target, and a neighbour
    // of the start is at least manhattan(start, target) - 1 steps away, so no hit is possible
    // before that level: expand without the per-level hit test / vote.
    {
        const int lmin = abs(sx - tx) + abs(sy - ty) - 1;
        while (level + AGV_BFS_UNROLL <= lmin) {
#pragma unroll
            for (int u = 0; u < AGV_BFS_UNROLL; u++) bfs_expand<NW, RPL>(f, av, m_up, m_dn);
            level += AGV_BFS_UNROLL;
            bool ne = false;
#pragma unroll
            for (int s = 0; s < RPL; s++) ne |= rany(f[s]);
            if (!__any_sync(FULL, ne)) return A_STOP;
        }
        // (the up to 3 remaining hit-free levels run in phase 2: one loop body less to fetch)
    }
    // Phase 2: blocks of AGV_BFS_UNROLL levels with a lane-local hit flag and ONE vote per block
    // (a vote + branch per level made every level wait for the vote's latency); the block in which
    // the frontier first touches a start neighbour is replayed level by level from its saved start.
    for (;;) {
        Row<NW> f0[RPL], av0[RPL];
#pragma unroll
        for (int s = 0; s < RPL; s++) { f0[s] = f[s]; av0[s] = av[s]; }
        bool h = false;
#pragma unroll
        for (int u = 0; u < AGV_BFS_UNROLL; u++) {
#pragma unroll
            for (int s = 0; s < RPL; s++) h |= rany(rand_(f[s], nb[s]));
            bfs_expand<NW, RPL>(f, av, m_up, m_dn);
        }
        if (__any_sync(FULL, h)) {
#pragma unroll
            for (int s = 0; s < RPL; s++) { f[s] = f0[s]; av[s] = av0[s]; }
#pragma unroll 1
            for (;;) {
                bool hit = false;
#pragma unroll
                for (int s = 0; s < RPL; s++) hit |= rany(rand_(f[s], nb[s]));
                if (__any_sync(FULL, hit)) {
                    unsigned code = 0;
                    if (sx + 1 < W && test_bit_local<NW, RPL>(f, sy, sx + 1, lane)) code |= 1u;
                    if (sy + 1 < H && test_bit_local<NW, RPL>(f, sy + 1, sx, lane)) code |= 2u;
                    if (sy - 1 >= 0 && test_bit_local<NW, RPL>(f, sy - 1, sx, lane)) code |= 4u;
                    if (sx - 1 >= 0 && test_bit_local<NW, RPL>(f, sy, sx - 1, lane)) code |= 8u;
                    code = __reduce_or_sync(FULL, code);
                    len = level + 1;
                    if (code & 1u) return A_RIGHT;
                    if (code & 2u) return A_DOWN;
                    if (code & 4u) return A_UP;
                    return A_LEFT;
                }
                bfs_expand<NW, RPL>(f, av, m_up, m_dn);
                ++level;
            }
        }
        level += AGV_BFS_UNROLL;
        // frontier-empty test (target unreachable) once per block: an empty frontier stays empty
        bool ne = false;
#pragma unroll
        for (int s = 0; s < RPL; s++) ne |= rany(f[s]);
        if (!__any_sync(FULL, ne)) return A_STOP;
    }
}

// ------------------------------------------------------------------ monotone fast path
// If a path of Manhattan length exists, every shortest path is a monotone staircase inside
// the bounding box of start and target, so only the (at most two) start neighbours that
// point towards the target can be one step closer, and a neighbour n is on a shortest path
// iff the target reaches n by a monotone path through valid cells.  That reachability is a
// row-by-row DP from the target's row to the start's row:
//     reach(r) = hfill( (reach(r -/+ 1) | seed) & valid(r) & box , valid(r) & box )
// where hfill spreads set bits towards the start column through runs of valid cells -- one
// 64-bit add: filled = (m & ~(m + x)) | x  (the carry ripples through the run).  Rows are
// mirrored first when the spread has to go towards lower columns.  ~15 instructions per row of
// the box instead of ~27 per BFS level over the whole grid; in the 64x64 benchmark 92.6 % of
// all searches are decided here.  Returns the action, or -1 when no Manhattan-length path
// exists (the caller then runs warp_bfs).  Exactly equivalent to the reference's result:
// with D = manhattan the neighbours at distance D-1 are precisely the forward neighbours
// with a monotone path, and the priority RIGHT, DOWN, UP, LEFT picks among them.
__device__ __forceinline__ uint64_t brev64(uint64_t x) {
    return ((uint64_t)__brev((uint32_t)x) << 32) | (uint64_t)__brev((uint32_t)(x >> 32));
}
template <int NW>
__device__ __forceinline__ Row<NW> rmirror(const Row<NW> &a) {
    Row<NW> r;
#pragma unroll
    for (int i = 0; i < NW; i++) r.w[i] = brev64(a.w[NW - 1 - i]);
    return r;
}
// bits [lo, hi] set
template <int NW>
__device__ __forceinline__ Row<NW> rrange(int lo, int hi) {
    Row<NW> r;
#pragma unroll
    for (int i = 0; i < NW; i++) {
        const int l = max(lo - 64 * i, 0), h = min(hi - 64 * i, 63);
        uint64_t m = 0ull;
        if (l <= h) m = (h == 63 ? ~0ull : ((1ull << (h + 1)) - 1ull)) & ~((1ull << l) - 1ull);
        r.w[i] = m;
    }
    return r;
}
// spread the bits of x (a subset of m) towards higher bit positions through the runs of m
template <int NW>
__device__ __forceinline__ Row<NW> rfill_up(const Row<NW> &x, const Row<NW> &m) {
    Row<NW> r;
    uint64_t carry = 0ull;
#pragma unroll
    for (int i = 0; i < NW; i++) {
        const uint64_t s1 = m.w[i] + x.w[i];
        const uint64_t c1 = s1 < m.w[i] ? 1ull : 0ull;
        const uint64_t s2 = s1 + carry;
        const uint64_t c2 = s2 < s1 ? 1ull : 0ull;
        carry = c1 | c2;
        r.w[i] = (m.w[i] & ~s2) | x.w[i];
    }
    return r;
}
template <int NW>
__device__ __forceinline__ bool rtest(const Row<NW> &a, int c) {
    uint64_t v = 0ull;
#pragma unroll
    for (int i = 0; i < NW; i++)
        if (i == (c >> 6)) v = a.w[i];
    return (v >> (c & 63)) & 1ull;
}

template <int NW, int RPL>
__device__ __forceinline__ int warp_monotone(const Row<NW> (&valid)[RPL], int sx, int sy, int tx, int ty, int lane,
                                             int &len) {
    const int dxs = (tx > sx) - (tx < sx), dys = (ty > sy) - (ty < sy);
    const bool mirror = tx > sx;  // the spread runs from the target's column towards the start's
    constexpr int WB = 64 * NW;
    const int sxm = mirror ? WB - 1 - sx : sx, txm = mirror ? WB - 1 - tx : tx;  // txm <= sxm
    Row<NW> vm[RPL];
#pragma unroll
    for (int s = 0; s < RPL; s++) vm[s] = mirror ? rmirror(valid[s]) : valid[s];
    const Row<NW> box = rrange<NW>(txm, sxm);
    Row<NW> reach = rzero<NW>(), prev = rzero<NW>();
#pragma unroll
    for (int i = 0; i < NW; i++)
        if (i == (txm >> 6)) reach.w[i] = 1ull << (txm & 63);  // seed: the target cell
    int r = ty;
    for (;;) {
        const int ol = r / RPL, os = r % RPL;
        Row<NW> mine = vm[0];
#pragma unroll
        for (int s = 1; s < RPL; s++)
            if (os == s) mine = vm[s];
        Row<NW> vr;
#pragma unroll
        for (int i = 0; i < NW; i++) vr.w[i] = __shfl_sync(FULL, (unsigned long long)mine.w[i], ol) & box.w[i];
        const Row<NW> now = rfill_up(rand_(reach, vr), vr);  // vertical step, then horizontal spread
        // the start's own row may come out empty when the start cell itself is invalid (matrix (a)
        // blocks every explorer, Explorer.py:274-283) while `prev` still holds the vertical neighbour
        if (r == sy) { reach = now; break; }
        if (!rany(now)) return -1;                            // (warp-uniform: every lane holds the same rows)
        prev = now;
        reach = now;
        r -= dys;
    }
    // reach = row sy, prev = row sy + dys (zero when dys == 0)
    const bool okH = dxs != 0 && rtest(reach, sxm - 1);
    const bool okV = dys != 0 && rtest(prev, sxm);
    len = abs(tx - sx) + abs(ty - sy);
    if (okH && dxs > 0) return A_RIGHT;
    if (okV && dys > 0) return A_DOWN;
    if (okV && dys < 0) return A_UP;
    if (okH && dxs < 0) return A_LEFT;
    len = 0;
    return -1;
}

// FindPathAstar's first action + length: monotone fast path, else the level-synchronous BFS
template <int NW, int RPL>
__device__ __forceinline__ int warp_path(const Row<NW> (&valid)[RPL], int sx, int sy, int tx, int ty, int H, int W,
                                         int lane, int &len) {
    len = 0;
    if (sx == tx && sy == ty) return A_STOP;
    const int a = warp_monotone<NW, RPL>(valid, sx, sy, tx, ty, lane, len);
    if (a >= 0) return a;
    return warp_bfs<NW, RPL>(valid, sx, sy, tx, ty, H, W, lane, len);
}

// Full distance field from the target (uint16, 0xFFFF unreachable) into dist[H*W]
// (global or shared).  Used by the host-side FindPathAstar mirror to rebuild
// path_list / action_list.
template <int NW, int RPL>
__device__ __forceinline__ void warp_bfs_field(const Row<NW> (&valid)[RPL], int tx, int ty, int H, int W, int lane,
                                               uint16_t *dist) {
    for (int p = lane; p < H * W; p += 32) dist[p] = 0xFFFF;
    __syncwarp();
    Row<NW> f[RPL], vis[RPL];
#pragma unroll
    for (int s = 0; s < RPL; s++) f[s] = rzero<NW>();
    set_bit<NW, RPL>(f, ty, tx, lane);
#pragma unroll
    for (int s = 0; s < RPL; s++) { f[s] = rand_(f[s], valid[s]); vis[s] = f[s]; }
    for (int level = 0; level <= H * W; ++level) {
        unsigned fl = 0;
#pragma unroll
        for (int s = 0; s < RPL; s++) {
            fl |= rany(f[s]) ? 1u : 0u;
            const int r = lane * RPL + s;
#pragma unroll
            for (int i = 0; i < NW; i++) {
                uint64_t m = f[s].w[i];
                while (m) {
                    const int b = __ffsll((long long)m) - 1;
                    m &= m - 1;
                    dist[r * W + i * 64 + b] = (uint16_t)level;
                }
            }
        }
        if (!__any_sync(FULL, fl != 0)) break;
        Row<NW> up_in = rshfl_up(f[RPL - 1]);
        Row<NW> dn_in = rshfl_down(f[0]);
        if (lane == 0) up_in = rzero<NW>();
        if (lane == 31) dn_in = rzero<NW>();
        Row<NW> n[RPL];
#pragma unroll
        for (int s = 0; s < RPL; s++) {
            Row<NW> a = ror(rshl1(f[s]), rshr1(f[s]));
            Row<NW> u = (s > 0) ? f[s > 0 ? s - 1 : 0] : up_in;
            Row<NW> d = (s < RPL - 1) ? f[s < RPL - 1 ? s + 1 : 0] : dn_in;
            a = ror(a, ror(u, d));
            n[s] = rand_(a, randn(valid[s], vis[s]));
        }
#pragma unroll
        for (int s = 0; s < RPL; s++) { vis[s] = ror(vis[s], n[s]); f[s] = n[s]; }
    }
}

// ------------------------------------------------------------------ scene in a warp
template <int NW, int RPL>
struct WarpScene {
    const DevCfg &c;
    int lane;
    // scene block staged in shared memory
    EnvHdr *hdr_s;
    uint16_t *pos_s, *tgt_s, *ord_s;
    uint8_t *meta_s;
    uint64_t *vrow_s;  // [H * NW] scratch rows of the current valid matrix (for the encoder)
    const uint32_t *tasks_e;
    // warp-uniform copy of the header
    uint32_t tick, cursor, n_done;
    int n_created, finished, over;
    uint64_t acted_mask;
    bool overlap;  // two created AGVs may share a cell (see build_occ)
    uint32_t cnt_turns, cnt_ended, cnt_bad;  // statistics, flushed once per kernel (flush_counters)
    // static layout rows + occupancy of created AGVs
    Row<NW> stor[RPL], pick[RPL], inb[RPL], occ[RPL];

    __device__ __forceinline__ WarpScene(const DevCfg &cfg, int lane_) : c(cfg), lane(lane_), overlap(true), cnt_turns(0), cnt_ended(0), cnt_bad(0) {}

    __device__ __forceinline__ void flush_counters() {
        if (lane == 0) {
            if (cnt_turns) atomicAdd(c.counters + 0, (unsigned long long)cnt_turns);
            if (cnt_ended) atomicAdd(c.counters + 1, (unsigned long long)cnt_ended);
            if (cnt_bad) atomicAdd(c.counters + 2, (unsigned long long)cnt_bad);
        }
    }

    __device__ __forceinline__ void bind(uint8_t *blk_s, uint64_t *vrow, int e) {
        hdr_s = reinterpret_cast<EnvHdr *>(blk_s);
        pos_s = reinterpret_cast<uint16_t *>(blk_s + c.off_pos);
        tgt_s = reinterpret_cast<uint16_t *>(blk_s + c.off_tgt);
        ord_s = reinterpret_cast<uint16_t *>(blk_s + c.off_ord);
        meta_s = blk_s + c.off_meta;
        vrow_s = vrow;
        tasks_e = c.tasks + (size_t)e * c.T;
    }

    // coalesced copy of the scene block global -> shared (uint4 granules)
    __device__ __forceinline__ void load_block(uint8_t *blk_s, int e) {
        const uint4 *src = reinterpret_cast<const uint4 *>(c.state + (size_t)e * c.env_stride);
        uint4 *dst = reinterpret_cast<uint4 *>(blk_s);
        for (int i = lane; i < c.env_stride / 16; i += 32) dst[i] = src[i];
        __syncwarp();
    }
    __device__ __forceinline__ void store_block(const uint8_t *blk_s, int e) {
        __syncwarp();
        uint4 *dst = reinterpret_cast<uint4 *>(c.state + (size_t)e * c.env_stride);
        const uint4 *src = reinterpret_cast<const uint4 *>(blk_s);
        for (int i = lane; i < c.env_stride / 16; i += 32) dst[i] = src[i];
    }

    __device__ __forceinline__ void load_static() {
#pragma unroll
        for (int s = 0; s < RPL; s++) {
            const int r = lane * RPL + s;
#pragma unroll
            for (int i = 0; i < NW; i++) {
                const bool ok = r < c.H;
                stor[s].w[i] = ok ? __ldg(c.storage_rows + r * NW + i) : 0ull;
                pick[s].w[i] = ok ? __ldg(c.picking_rows + r * NW + i) : 0ull;
                inb[s].w[i] = ok ? __ldg(c.inb_rows + r * NW + i) : 0ull;
            }
        }
    }

    __device__ __forceinline__ void hdr_to_regs() {
        tick = hdr_s->tick; cursor = hdr_s->cursor; n_done = hdr_s->n_done;
        n_created = hdr_s->n_created; finished = hdr_s->finished; over = hdr_s->over;
        acted_mask = hdr_s->acted_mask;
    }
    __device__ __forceinline__ void regs_to_hdr() {
        hdr_s->tick = tick; hdr_s->cursor = cursor; hdr_s->n_done = n_done;
        hdr_s->n_created = (uint16_t)n_created; hdr_s->finished = (uint8_t)finished;
        hdr_s->over = (uint8_t)over; hdr_s->acted_mask = acted_mask;
        __syncwarp();
    }

    // occupancy rows of the created AGVs.  `overlap` is raised when two created AGVs share a
    // cell: only possible after moves under the A_star protocol with caller-supplied actions
    // (no AGV-AGV test, Scene.py:159-162) or agv_set_state; while it is false the per-turn
    // "is another AGV on this cell" ballots reduce to one occupancy-bit test.
    __device__ __forceinline__ void build_occ() {
        // parallel over AGVs: every lane ORs the cells of its AGVs into the shared scratch rows
        // (vrow_s), then picks up the rows it owns; the value returned by atomicOr tells whether
        // the cell was already taken (two created AGVs on one cell)
        for (int q = lane; q < c.H * NW; q += 32) vrow_s[q] = 0ull;
        __syncwarp();
        bool dup = false;
        for (int j = lane; j < n_created; j += 32) {
            const int p = pos_s[j], cc = (p & 255) - 1, r = (p >> 8) - 1;
            const unsigned long long bit = 1ull << (cc & 63);
            const unsigned long long old = atomicOr(reinterpret_cast<unsigned long long *>(vrow_s) + r * NW + (cc >> 6), bit);
            dup |= (old & bit) != 0ull;
        }
        __syncwarp();
#pragma unroll
        for (int s = 0; s < RPL; s++) {
            const int r = lane * RPL + s;
#pragma unroll
            for (int i = 0; i < NW; i++) occ[s].w[i] = r < c.H ? vrow_s[r * NW + i] : 0ull;
        }
        overlap = __any_sync(FULL, dup);
    }

    __device__ __forceinline__ Agv load_agv(int i) const {
        Agv a;
        const int p = pos_s[i], t = tgt_s[i], m = meta_s[i];
        a.x = p & 255; a.y = p >> 8; a.tx = t & 255; a.ty = t >> 8;
        a.stage = m & 3; a.loaded = (m >> 2) & 1; a.idle = (m >> 3) & 1; a.order = ord_s[i];
        return a;
    }
    // every lane stores the same (warp-uniform) values: one wavefront per store and no
    // divergent `if (lane == 0)` block (BSSY/BSYNC pairs were top stall_no_instruction sites)
    __device__ __forceinline__ void store_agv(int i, const Agv &a) {
        pos_s[i] = (uint16_t)(a.x | (a.y << 8));
        tgt_s[i] = (uint16_t)(a.tx | (a.ty << 8));
        ord_s[i] = (uint16_t)a.order;
        meta_s[i] = (uint8_t)meta_pack(a);
        __syncwarp();
    }

    // any created AGV j != i standing on cell code (x | y<<8)?   (warp ballot)
    __device__ __forceinline__ bool other_at(int i, int cell) const {
        if (!overlap) {  // occupancy bit, minus AGV i itself
            if (pos_s[i] == cell) return false;
            return __any_sync(FULL, test_bit_local<NW, RPL>(occ, (cell >> 8) - 1, (cell & 255) - 1, lane));
        }
        bool m = false;
        for (int j = lane; j < n_created; j += 32) m |= (j != i) && (pos_s[j] == cell);
        return __any_sync(FULL, m);
    }

    // Explorer.load_condition, Explorer.py:102-116
    __device__ __forceinline__ void load_condition(Agv &a) const {
        if (c.always_loaded) a.loaded = 1;
        else if (c.always_empty) a.loaded = 0;
        else a.loaded = (a.stage == 0) ? 0 : 1;
    }
    // Explorer.get_task, Explorer.py:68-100 (AGV is created)
    __device__ __forceinline__ void get_task(Agv &a) {
        if ((uint32_t)c.T == cursor && a.stage == 0) {
            a.idle = 1;
            if ((uint32_t)c.T == n_done) finished = 1;
            return;  // target_position / loaded keep their values
        }
        if (a.stage == 0) a.order = (int)(cursor++);
        const uint32_t t = __ldg(tasks_e + a.order);
        if (a.stage == 1) { a.tx = (t >> 16) & 255; a.ty = (t >> 24) & 255; }
        else { a.tx = t & 255; a.ty = (t >> 8) & 255; }
        load_condition(a);
    }
    // Explorer.continue_working, Explorer.py:251-259
    __device__ __forceinline__ void continue_working(Agv &a) {
        a.stage++;
        if (a.stage == 3) { a.stage = 0; n_done++; }
        get_task(a);
    }

    // Scene.init (Scene.py:61-65) + create_explorer of AGV 0 (Scene.py:86)
    __device__ __forceinline__ void reset() {
        tick = 0; cursor = 0; n_done = 0; finished = 0; over = 0; acted_mask = 0;
        for (int j = lane; j < c.N; j += 32) {
            pos_s[j] = (uint16_t)(1 | (1 << 8));
            ord_s[j] = 0; meta_s[j] = 0;  // tgt_s is NOT reset (Explorer.py:45-58)
        }
        __syncwarp();
        n_created = 0;
        if (c.N > 0) {
            n_created = 1;
            Agv a = load_agv(0);
            get_task(a);
            store_agv(0, a);
        }
    }

    // Scene.py:124-134: 1 act, 0 skip (idle), -1 break (uncreated), -2 return (finished)
    __device__ __forceinline__ int turn_state(int i) const {
        if (i >= n_created) return -1;
        if (finished) return -2;
        if ((meta_s[i] >> 3) & 1) return 0;
        return 1;
    }

    // base passability by cell code and load state (Explorer.py:266-272, StateManager.py:92-120)
    __device__ __forceinline__ Row<NW> base_row(int s, int loaded) const {
        Row<NW> b = randn(inb[s], pick[s]);
        if (loaded) b = randn(b, stor[s]);
        return b;
    }

    // matrix (b): StateManager.create_path_matrix, StateManager.py:43-61
    __device__ __forceinline__ void valid_b(const Agv &a, int i, Row<NW> (&v)[RPL]) const {
        const bool shared = other_at(i, a.x | (a.y << 8));
        Row<NW> own[RPL], oth[RPL];
#pragma unroll
        for (int s = 0; s < RPL; s++) { own[s] = rzero<NW>(); oth[s] = occ[s]; }
        set_bit<NW, RPL>(own, a.y - 1, a.x - 1, lane);
        if (!shared) clear_bit<NW, RPL>(oth, a.y - 1, a.x - 1, lane);
        set_bit<NW, RPL>(own, a.ty - 1, a.tx - 1, lane);
#pragma unroll
        for (int s = 0; s < RPL; s++) v[s] = randn(ror(base_row(s, a.loaded), own[s]), oth[s]);
    }
    // matrix (a): Explorer.create_valid_matrix, Explorer.py:261-284
    __device__ __forceinline__ void valid_a(const Agv &a, Row<NW> (&v)[RPL]) const {
        Row<NW> tg[RPL], blk[RPL];
#pragma unroll
        for (int s = 0; s < RPL; s++) { tg[s] = rzero<NW>(); blk[s] = rzero<NW>(); }
        set_bit<NW, RPL>(tg, a.ty - 1, a.tx - 1, lane);
        if (c.N > 1) {
#pragma unroll
            for (int s = 0; s < RPL; s++) blk[s] = occ[s];
            if (n_created < c.N) set_bit<NW, RPL>(blk, 0, 0, lane);  // uncreated AGVs park at (1,1)
        }
#pragma unroll
        for (int s = 0; s < RPL; s++) v[s] = randn(ror(base_row(s, a.loaded), tg[s]), blk[s]);
    }

    __device__ __forceinline__ void rows_to_smem(const Row<NW> (&v)[RPL]) {
        __syncwarp();
#pragma unroll
        for (int s = 0; s < RPL; s++) {
            const int r = lane * RPL + s;
            if (r < c.H) {
#pragma unroll
                for (int i = 0; i < NW; i++) vrow_s[r * NW + i] = v[s].w[i];
            }
        }
        __syncwarp();
    }

    // StateManager.create_state(obs_clip=True) (StateManager.py:14-41,145-192) into a
    // [3,K,K] bf16 record in shared memory; vrow_s holds matrix (b).  KT > 0 fixes K at
    // compile time (the index divisions become multiplies).
    template <int KT>
    __device__ __forceinline__ void encode_clip_impl(uint16_t *dst, const Agv &a) const {
        const int K = KT > 0 ? KT : c.K, P = (K - 1) >> 1, KK = K * K, n = 3 * KK;
        const int cx = a.x - 1, cy = a.y - 1;
        int tdx = a.tx - a.x, tdy = a.ty - a.y;
        tdx = max(-P, min(P, tdx)) + P;
        tdy = max(-P, min(P, tdy)) + P;
        for (int idx = lane; idx < n; idx += 32) {
            const int ch = idx / KK, rem = idx - ch * KK, wy = rem / K, wx = rem - wy * K;
            bool v;
            if (ch == 0) v = (wy == P) && (wx == P);
            else if (ch == 1) v = (wy == tdy) && (wx == tdx);
            else {
                const int r = cy - P + wy, cc = cx - P + wx;
                v = false;
                if (r >= 0 && r < c.H && cc >= 0 && cc < c.W) v = (vrow_s[r * NW + (cc >> 6)] >> (cc & 63)) & 1ull;
            }
            dst[idx] = v ? BF16_ONE : (uint16_t)0;
        }
    }
    // K = 7 (the reference's default padding 3): 49 window cells, at most two per lane, whose
    // window coordinates are compile-time functions of the lane; three stores per cell.
    __device__ __forceinline__ void encode_clip7(uint16_t *dst, const Agv &a) const {
        constexpr int K = 7, P = 3, KK = 49;
        const int cx = a.x - 1, cy = a.y - 1;
        const int tdx = max(-P, min(P, a.tx - a.x)) + P, tdy = max(-P, min(P, a.ty - a.y)) + P;
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int cell = lane + 32 * q;
            if (cell < KK) {
                const int wy = cell / K, wx = cell - wy * K;
                const int r = cy - P + wy, cc = cx - P + wx;
                bool v = false;
                if (r >= 0 && r < c.H && cc >= 0 && cc < c.W) v = (vrow_s[r * NW + (cc >> 6)] >> (cc & 63)) & 1ull;
                dst[cell] = (cell == P * K + P) ? BF16_ONE : (uint16_t)0;
                dst[KK + cell] = (wy == tdy && wx == tdx) ? BF16_ONE : (uint16_t)0;
                dst[2 * KK + cell] = v ? BF16_ONE : (uint16_t)0;
            }
        }
    }
    __device__ __forceinline__ void encode_clip(uint16_t *dst, const Agv &a) const {
        if (c.K == 7) encode_clip7(dst, a);
        else encode_clip_impl<0>(dst, a);
    }
    // StateManager.create_state(obs_clip=False) straight to a [3,H,W] bf16 record in HBM.
    // W % 8 == 0: one 16-byte store = 8 cells of one row; the two one-hot planes are zero
    // vectors except one; the valid plane expands 8 mask bits with two-bits-per-IMAD
    // (t = b0 | b1 << 16; t * 0x3F80 = {bf16(b0), bf16(b1)}).
    __device__ __forceinline__ void encode_full(uint16_t *dst, const Agv &a) const {
        const int H = c.H, W = c.W, HW = H * W, n = 3 * HW;
        const int cpos = (a.y - 1) * W + (a.x - 1), tpos = (a.ty - 1) * W + (a.tx - 1);
        if ((W & 7) == 0 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
            const int nvr = W >> 3, nvc = H * nvr;
            uint4 *d4 = reinterpret_cast<uint4 *>(dst);
            const int cv = cpos >> 3, tv = tpos >> 3;
            for (int v = lane; v < nvc; v += 32) {
                uint4 o0 = make_uint4(0, 0, 0, 0), o1 = make_uint4(0, 0, 0, 0);
                if (v == cv) {
                    const unsigned val = (unsigned)BF16_ONE << ((cpos & 1) * 16), w = (cpos & 7) >> 1;
                    o0.x = w == 0 ? val : 0u; o0.y = w == 1 ? val : 0u; o0.z = w == 2 ? val : 0u; o0.w = w == 3 ? val : 0u;
                }
                if (v == tv) {
                    const unsigned val = (unsigned)BF16_ONE << ((tpos & 1) * 16), w = (tpos & 7) >> 1;
                    o1.x = w == 0 ? val : 0u; o1.y = w == 1 ? val : 0u; o1.z = w == 2 ? val : 0u; o1.w = w == 3 ? val : 0u;
                }
                d4[v] = o0;
                d4[nvc + v] = o1;
            }
            int r = lane / nvr, c8 = lane - r * nvr;
            const int dr = 32 / nvr, dc = 32 - dr * nvr;
            uint4 *d2 = d4 + 2 * nvc;
            for (int v = lane; v < nvc; v += 32) {
                const unsigned bits = (unsigned)(vrow_s[r * NW + (c8 >> 3)] >> ((c8 & 7) * 8)) & 0xffu;
                uint4 o;
                unsigned u, t;
                u = bits & 3u;        t = (u | (u << 15)) & 0x10001u; o.x = t * (unsigned)BF16_ONE;
                u = (bits >> 2) & 3u; t = (u | (u << 15)) & 0x10001u; o.y = t * (unsigned)BF16_ONE;
                u = (bits >> 4) & 3u; t = (u | (u << 15)) & 0x10001u; o.z = t * (unsigned)BF16_ONE;
                u = (bits >> 6) & 3u; t = (u | (u << 15)) & 0x10001u; o.w = t * (unsigned)BF16_ONE;
                d2[v] = o;
                c8 += dc; r += dr;
                if (c8 >= nvr) { c8 -= nvr; r++; }
            }
        } else {
            for (int idx = lane; idx < n; idx += 32) {
                const int ch = idx / HW, p = idx - ch * HW;
                bool v;
                if (ch == 0) v = (p == cpos);
                else if (ch == 1) v = (p == tpos);
                else {
                    const int r = p / W, cc = p - r * W;
                    v = (vrow_s[r * NW + (cc >> 6)] >> (cc & 63)) & 1ull;
                }
                dst[idx] = v ? BF16_ONE : (uint16_t)0;
            }
        }
    }

    // Scene.check_new_veh, Scene.py:354-366
    __device__ __forceinline__ void spawn() {
        if (n_created >= c.N) return;
        const bool occupied = test_any(occ, 0, 0);
        if (occupied) return;
        const int k = n_created++;
        Agv a = load_agv(k);
        get_task(a);
        store_agv(k, a);
        set_bit<NW, RPL>(occ, a.y - 1, a.x - 1, lane);
    }
    __device__ __forceinline__ bool test_any(const Row<NW> (&rows)[RPL], int r, int cc) const {
        return __any_sync(FULL, test_bit_local<NW, RPL>(rows, r, cc, lane));
    }
};

struct TurnOut {
    int action;
    float reward;
    int done;
    int plen;
    int slen;
};

// One AGV turn = [encode obs] -> action -> Explorer.execute_action (Explorer.py:131-208)
// -> Scene.py:155 correction -> [encode next_obs].  Warp-uniform control flow.
// obs_dst / next_dst: destination record (shared staging for OBS_CLIP, global for
// OBS_FULL) or nullptr.
// OBS (observation kind) and SHAPED (settings.Sparse_Reward) are compile-time so that each
// kernel carries only the code of its own path: the turn body is re-fetched by every warp
// once per turn and instruction-fetch stalls were the top stall reason with one fat kernel.
template <int NW, int RPL, int OBS, bool SHAPED>
__device__ __forceinline__ void do_turn(WarpScene<NW, RPL> &ws, int i, int proto, int policy, int given,
                                        uint16_t *obs_dst, uint16_t *next_dst, TurnOut &out) {
    const DevCfg &c = ws.c;
    const int lane = ws.lane;
    Agv a = ws.load_agv(i);
    Row<NW> v[RPL];
    const bool want_obs = (OBS != OBS_NONE) && (obs_dst != nullptr);
    const bool guided = (policy == POLICY_GUIDED) || (policy == POLICY_GIVEN && given < 0);
    if (want_obs || guided) ws.valid_b(a, i, v);
    if (OBS != OBS_NONE && want_obs) {
        ws.rows_to_smem(v);
        if (OBS == OBS_CLIP) ws.encode_clip(obs_dst, a);
        else ws.encode_full(obs_dst, a);
    }
    int action = given, plen = 0;
    const bool policy_bfs = guided || policy == POLICY_ASTAR;
    if (!guided && policy == POLICY_ASTAR) ws.valid_a(a, v);
    if (policy_bfs) {  // the one policy-search call site (matrix a or b is already in v)
        action = warp_path<NW, RPL>(v, a.x - 1, a.y - 1, a.tx - 1, a.ty - 1, c.H, c.W, lane, plen);
    } else if (action > A_STOP) {  // the reference raises KeyError (utils.py:21); counted, treated as STOP
        ws.cnt_bad++;
        action = A_STOP;
    }
    // ---- Explorer.check_action, Explorer.py:151-208
    const int dx = (action == A_RIGHT) - (action == A_LEFT);
    const int dy = (action == A_DOWN) - (action == A_UP);
    const int nx = a.x + dx, ny = a.y + dy;
    int r = 0, end = 0, slen = 0xFFFF;
    double rshaped = 0.0;
    bool use_shaped = false;
    const bool oob = nx < 1 || ny < 1 || nx > c.W || ny > c.H;
    if (oob) { r = -1; end = 1; }
    else {
        const int code = __ldg(c.grid + (ny - 1) * c.W + (nx - 1));
        const bool at_target = (nx == a.tx) && (ny == a.ty);
        if (code == 1 && a.loaded && !at_target) { r = -1; end = 1; }
        else if (code == 2 && !at_target) { r = -1; end = 1; }
        else if (proto == PROTO_INTELLIGENT && ws.other_at(i, nx | (ny << 8))) { r = -1; end = 1; }
        else if (at_target) { r = 1; ws.continue_working(a); }
        else if (SHAPED) {
            // Explorer.rectify_reward (Explorer.py:210-249): only path_length_new matters
            ws.valid_a(a, v);
            int l = 0;
            warp_path<NW, RPL>(v, nx - 1, ny - 1, a.tx - 1, a.ty - 1, c.H, c.W, lane, l);
            slen = l;
            const double M = 0.5 * (double)(c.H + c.W);
            rshaped = 0.001 * (M - (double)l) / M;
            use_shaped = true;
        }
    }
    // ---- Explorer.py:145-148: the move is applied iff not is_end
    if (proto != PROTO_INTELLIGENT && !guided && policy != POLICY_ASTAR) ws.overlap = true;
    if (!end && (dx | dy)) {
        const int oldcell = a.x | (a.y << 8);
        const bool still = ws.other_at(i, oldcell);
        if (!still) clear_bit<NW, RPL>(ws.occ, a.y - 1, a.x - 1, lane);
        a.x = nx; a.y = ny;
        set_bit<NW, RPL>(ws.occ, ny - 1, nx - 1, lane);
    }
    ws.store_agv(i, a);
    if (proto == PROTO_INTELLIGENT) {
        if (ws.finished || (long long)ws.tick >= c.max_steps) end = 1;  // Scene.py:155
        if (end) ws.over = 1;
    } else {
        end = 0;  // is_end is ignored by the loop in A_star / manual mode (Scene.py:159-162)
    }
    ws.acted_mask |= (1ull << i);
    ws.cnt_turns++;
    ws.cnt_ended += end;
    if (OBS != OBS_NONE && next_dst != nullptr) {
        // store_info's create_state on the snapshot after the move (Scene.py:156)
        ws.valid_b(a, i, v);
        ws.rows_to_smem(v);
        if (OBS == OBS_CLIP) ws.encode_clip(next_dst, a);
        else ws.encode_full(next_dst, a);
    }
    out.action = action;
    out.reward = use_shaped ? (float)rshaped : (float)r;
    out.done = end;
    out.plen = plen;
    out.slen = slen;
}

// copy nbytes shared -> global with the widest aligned stores available
__device__ __forceinline__ void flush_bytes(uint8_t *dst, const uint8_t *src, int nbytes, int lane) {
    __syncwarp();
    const uintptr_t d = reinterpret_cast<uintptr_t>(dst), s = reinterpret_cast<uintptr_t>(src);
    if (((d | s) & 15) == 0 && (nbytes & 15) == 0) {
        for (int i = lane; i < (nbytes >> 4); i += 32)
            reinterpret_cast<uint4 *>(dst)[i] = reinterpret_cast<const uint4 *>(src)[i];
    } else if (((d | s) & 3) == 0 && (nbytes & 3) == 0) {
        for (int i = lane; i < (nbytes >> 2); i += 32)
            reinterpret_cast<uint32_t *>(dst)[i] = reinterpret_cast<const uint32_t *>(src)[i];
    } else {
        for (int i = lane; i < (nbytes >> 1); i += 32)
            reinterpret_cast<uint16_t *>(dst)[i] = reinterpret_cast<const uint16_t *>(src)[i];
    }
    __syncwarp();
}

__device__ __forceinline__ void zero_global(uint16_t *dst, int n_elems, int lane) {
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (n_elems & 7) == 0) {
        for (int i = lane; i < (n_elems >> 3); i += 32) reinterpret_cast<uint4 *>(dst)[i] = make_uint4(0, 0, 0, 0);
    } else {
        for (int i = lane; i < n_elems; i += 32) dst[i] = 0;
    }
}

}  // namespace agv
