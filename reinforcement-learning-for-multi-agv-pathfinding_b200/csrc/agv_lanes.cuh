// agv_lanes.cuh -- lane-per-scene fused tick.
//
// The turn logic of a scene (Scene.py:124-162, Explorer.py:131-208) is scalar and strictly
// sequential, so a warp that owns ONE scene (agv_core.cuh, k_tick) executes it 32 times
// redundantly.  Here one LANE owns one scene: a warp carries SPW scenes whose state blocks and
// occupancy bit rows live in shared memory, and the scalar parts (task FSM, move test, the
// Manhattan-path row DP) run lane-parallel.  Only the wide operations are warp-cooperative, one
// scene at a time over a ballot of the lanes that need them:
//   * the observation record (3 x K x K or 3 x H x W bf16) -- 32 lanes write one record;
//   * the level-synchronous bit-parallel BFS (warp_bfs) for the searches the row DP cannot
//     decide (no Manhattan-length path).
// Results are bit-identical to k_tick; tests/test_gpu_parity.py runs both.
#pragma once
#include "agv_core.cuh"

namespace agv {

struct LaneArgs {
    int proto, policy, obs_kind;
    const int8_t *action;
    int8_t *act_out;
    float *reward;
    uint8_t *done;
    uint16_t *plen, *slen;
    uint16_t *obs, *next_obs;
    int rec_elems;
    // shared memory plan (bytes from the start of dynamic shared memory; one warp per CTA)
    int off_tab;     // u64 [4][H*NW]: base(empty), base(loaded), and both mirrored
    int off_blk;     // SPW scene blocks, blk_stride bytes apart (odd number of words)
    int blk_stride;
    int off_occ;     // SPW x occ_stride u64: occupancy rows of the created AGVs
    int off_occm;    // same, columns mirrored (c -> 64*NW-1-c)
    int occ_stride;  // in u64 (odd)
    int off_vrow;    // u64 [H*NW]: matrix (b) of the scene being encoded
    int off_queue;   // LaneQueue (worker-warp variant)
    int off_zero;    // u64 [H*NW] of zeros: "occupancy" of a matrix that does not block the explorers
};

// Column-mirrored copies of the occupancy rows / base tables are kept only for maps up to 64 columns
// (NW = 1): at 128 x 128 they would double the 2 KB of occupancy per scene and halve the number of
// resident CTAs, so there the row DP mirrors the rows it loads (BREV, off its dependent chain).
template <int NW>
__host__ __device__ constexpr bool mirror_copy() { return NW == 1; }

template <int NW>
__device__ __forceinline__ Row<NW> ldrow(const uint64_t *p) {
    Row<NW> r;
#pragma unroll
    for (int i = 0; i < NW; i++) r.w[i] = p[i];
    return r;
}
template <int NW>
__device__ __forceinline__ Row<NW> rbit(int c) {
    Row<NW> r;
#pragma unroll
    for (int i = 0; i < NW; i++) r.w[i] = (i == (c >> 6)) ? (1ull << (c & 63)) : 0ull;
    return r;
}

// ------------------------------------------------------------------ Manhattan-path row DP
// Lane-parallel version of warp_monotone (agv_core.cuh): every lane runs its own search on
// its own scene; rows come from shared memory (static base rows + the scene's occupancy
// rows, both also kept column-mirrored so that the carry fill always runs upwards).
// matrix row r = (base[loaded][r] | target bit) & ~(occ[r] & om).  Returns the action or -1.
template <int NW>
__device__ __forceinline__ int lane_monotone(bool need, const uint64_t *tab, int HN, const uint64_t *occ,
                                             const uint64_t *occm, int loaded, int sx, int sy, int tx, int ty,
                                             int &len) {
    constexpr int WB = 64 * NW;
    const int dxs = (tx > sx) - (tx < sx), dys = (ty > sy) - (ty < sy);
    const bool mirror = tx > sx;
    const int sxm = mirror ? WB - 1 - sx : sx, txm = mirror ? WB - 1 - tx : tx;  // txm <= sxm
    const bool flip = mirror && !mirror_copy<NW>();  // mirror the loaded rows on the fly
    const uint64_t *bp = tab + ((mirror && mirror_copy<NW>() ? 2 : 0) + loaded) * HN + ty * NW;
    const uint64_t *op = (mirror && mirror_copy<NW>() ? occm : occ) + ty * NW;
    const int step = -dys * NW;
    const Row<NW> box = rrange<NW>(txm, sxm);
    Row<NW> reach = rbit<NW>(txm);
    Row<NW> rfin = rzero<NW>(), pfin = rzero<NW>();  // reach of the start's row / of the row before it
    const int n = abs(ty - sy);                       // rows after the target's
    const int nl = need ? n : 0;
    if (!need) { bp = tab; op = tab; }
    {   // the target's row: the target cell itself is valid unless occupied
        Row<NW> b = ldrow<NW>(bp), o = ldrow<NW>(op);
        if (!mirror_copy<NW>() && flip) { b = rmirror(b); o = rmirror(o); }
        b = ror(b, reach);
#pragma unroll
        for (int i = 0; i < NW; i++) b.w[i] = b.w[i] & ~o.w[i] & box.w[i];
        reach = rfill_up(rand_(reach, b), b);
        if (nl == 0) rfin = reach;
    }
    const int nmax = __reduce_max_sync(FULL, nl);
    // The loop-carried chain is only reach -> and -> add -> and/or -> reach (4 dependent ops per
    // row): the row loads do not depend on the DP state, a dead reach set stays empty by itself
    // (no liveness flag), and a lane that has passed its last row keeps re-reading that row and
    // computing into the void -- its result was snapshotted when it == nl, off the chain.
#pragma unroll 4
    for (int it = 1; it <= nmax; it++) {
        const int k = min(it, nl) * step;
        Row<NW> b = ldrow<NW>(bp + k), o = ldrow<NW>(op + k);
        if (!mirror_copy<NW>() && flip) { b = rmirror(b); o = rmirror(o); }
#pragma unroll
        for (int i = 0; i < NW; i++) b.w[i] = b.w[i] & ~o.w[i] & box.w[i];
        const Row<NW> now = rfill_up(rand_(reach, b), b);
        if (it == nl) { rfin = now; pfin = reach; }
        reach = now;
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

// 7 x 7 window of matrix (b) around the AGV as a 49-bit mask (bit wy*7+wx), lane-local:
// StateManager.py:145-192 with padding_size 3; cells outside the map read 0.
template <int NW>
__device__ __forceinline__ uint64_t lane_window7(const uint64_t *tabl, const uint64_t *occ, int H, int cx, int cy,
                                                 int tx0, int ty0, bool shared) {
    // Branch-free: rows outside the map are clamped for the load and masked afterwards, the own /
    // target bits go in by selects -- so all 14 row loads issue up front instead of one
    // load-latency per row behind a branch (the window was 1 k of the ~5 k cycles of a turn).
    uint64_t wm = 0ull;
    const int lo = cx - 3;
    const Row<NW> ownb = rbit<NW>(cx), tgtb = rbit<NW>(tx0);
#pragma unroll
    for (int wy = 0; wy < 7; wy++) {
        const int r = cy - 3 + wy, rc = min(max(r, 0), H - 1);
        Row<NW> b = ldrow<NW>(tabl + rc * NW), o = ldrow<NW>(occ + rc * NW);
        const bool is_t = r == ty0, is_o = r == cy;
#pragma unroll
        for (int i = 0; i < NW; i++) {
            const uint64_t ob = is_o ? ownb.w[i] : 0ull;
            b.w[i] = (b.w[i] | (is_t ? tgtb.w[i] : 0ull) | ob) & ~(o.w[i] & ~(shared ? 0ull : ob));
        }
        uint64_t bits;
        if (NW == 1) {
            bits = lo >= 0 ? (b.w[0] >> lo) : (b.w[0] << (-lo));
        } else {
            if (lo < 0) bits = b.w[0] << (-lo);
            else {
                const int wi = lo >> 6, off = lo & 63;
                uint64_t a0 = 0ull, a1 = 0ull;
#pragma unroll
                for (int i = 0; i < NW; i++) { if (i == wi) a0 = b.w[i]; if (i == wi + 1) a1 = b.w[i]; }
                bits = (a0 >> off) | (off ? (a1 << (64 - off)) : 0ull);
            }
        }
        wm |= (r == rc) ? ((bits & 0x7Full) << (7 * wy)) : 0ull;
    }
    return wm;
}

// [3,7,7] bf16 record from the window mask; tpos = clamped target cell index (tdy*7+tdx)
__device__ __forceinline__ void expand_clip7(uint16_t *dst, uint64_t wm, int tpos, int lane) {
#pragma unroll
    for (int q = 0; q < 2; q++) {
        const int cell = lane + 32 * q;
        if (cell < 49) {
            dst[cell] = (cell == 24) ? BF16_ONE : (uint16_t)0;
            dst[49 + cell] = (cell == tpos) ? BF16_ONE : (uint16_t)0;
            dst[98 + cell] = ((wm >> cell) & 1ull) ? BF16_ONE : (uint16_t)0;
        }
    }
}

// ------------------------------------------------------------------ observation encoders
// (warp-cooperative, one record, straight to HBM; vrow = matrix (b) rows in shared memory)
// StateManager.create_state(obs_clip=True), StateManager.py:14-41,145-192
template <int NW>
__device__ __forceinline__ void coop_encode_clip(uint16_t *dst, const uint64_t *vrow, int H, int W, int K, int x, int y,
                                                 int tx, int ty, int lane) {
    const int P = (K - 1) >> 1, KK = K * K;
    const int cx = x - 1, cy = y - 1;
    const int tdx = max(-P, min(P, tx - x)) + P, tdy = max(-P, min(P, ty - y)) + P;
    if (K == 7) {
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int cell = lane + 32 * q;
            if (cell < 49) {
                const int wy = cell / 7, wx = cell - wy * 7;
                const int r = cy - 3 + wy, cc = cx - 3 + wx;
                bool v = false;
                if (r >= 0 && r < H && cc >= 0 && cc < W) v = (vrow[r * NW + (cc >> 6)] >> (cc & 63)) & 1ull;
                dst[cell] = (cell == 24) ? BF16_ONE : (uint16_t)0;
                dst[49 + cell] = (wy == tdy && wx == tdx) ? BF16_ONE : (uint16_t)0;
                dst[98 + cell] = v ? BF16_ONE : (uint16_t)0;
            }
        }
        return;
    }
    for (int cell = lane; cell < KK; cell += 32) {
        const int wy = cell / K, wx = cell - wy * K;
        const int r = cy - P + wy, cc = cx - P + wx;
        bool v = false;
        if (r >= 0 && r < H && cc >= 0 && cc < W) v = (vrow[r * NW + (cc >> 6)] >> (cc & 63)) & 1ull;
        dst[cell] = (wy == P && wx == P) ? BF16_ONE : (uint16_t)0;
        dst[KK + cell] = (wy == tdy && wx == tdx) ? BF16_ONE : (uint16_t)0;
        dst[2 * KK + cell] = v ? BF16_ONE : (uint16_t)0;
    }
}

// StateManager.create_state(obs_clip=False): [3,H,W] bf16 (same scheme as WarpScene::encode_full)
template <int NW>
__device__ __forceinline__ void coop_encode_full(uint16_t *dst, const uint64_t *vrow, int H, int W, int x, int y, int tx,
                                                 int ty, int lane) {
    const int HW = H * W, n = 3 * HW;
    const int cpos = (y - 1) * W + (x - 1), tpos = (ty - 1) * W + (tx - 1);
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
            const unsigned bits = (unsigned)(vrow[r * NW + (c8 >> 3)] >> ((c8 & 7) * 8)) & 0xffu;
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
                v = (vrow[r * NW + (cc >> 6)] >> (cc & 63)) & 1ull;
            }
            dst[idx] = v ? BF16_ONE : (uint16_t)0;
        }
    }
}

// ------------------------------------------------------------------ warp-cooperative pieces
// (run either by the scene warp itself or by a worker warp of the same CTA)
constexpr int Q_EXIT = 0, Q_BFS = 1, Q_ENC7 = 2, Q_ENCG = 3;
constexpr int BAR_GO = 1, BAR_DONE = 2;
struct LaneQueue {
    int kind, n, i, which;
    uint4 req[32];
    unsigned res[32];
};
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// BFS of one scene.  pk = sx | sy<<7 | tx<<14 | ty<<21 | loaded<<28 | blocks<<29 (0-based cells)
template <int NW, int RPL>
__device__ __forceinline__ int coop_bfs(unsigned pk, const uint64_t *tab, int HN, const uint64_t *oj, int H, int W,
                                        int lane, int &len) {
    const int jsx = pk & 127, jsy = (pk >> 7) & 127, jtx = (pk >> 14) & 127, jty = (pk >> 21) & 127;
    const int jl = (pk >> 28) & 1;
    const uint64_t jom = (pk >> 29) & 1 ? ~0ull : 0ull;
    const uint64_t *bj = tab + jl * HN;
    Row<NW> v[RPL];
#pragma unroll
    for (int s = 0; s < RPL; s++) {
        const int r = lane * RPL + s;
        v[s] = rzero<NW>();
        if (r < H) {
#pragma unroll
            for (int i = 0; i < NW; i++) {
                uint64_t b = bj[r * NW + i];
                if (r == jty && i == (jtx >> 6)) b |= 1ull << (jtx & 63);
                v[s].w[i] = b & ~(oj[r * NW + i] & jom);
            }
        }
    }
    return warp_bfs<NW, RPL>(v, jsx, jsy, jtx, jty, H, W, lane, len);
}

// generic observation record of one scene: materialise matrix (b) in `vrow`, then encode.
// pk = x | y<<8 | tx<<16 | ty<<24 (1-based), fl = loaded | shared<<1
template <int NW, int OBS>
__device__ __forceinline__ void coop_encode(unsigned pk, unsigned fl, const uint64_t *tab, int HN, const uint64_t *oj,
                                            uint64_t *vrow, int H, int W, int K, uint16_t *dst, int lane) {
    const int jx = pk & 255, jy = (pk >> 8) & 255, jtx = (pk >> 16) & 255, jty = pk >> 24;
    const int jl = fl & 1;
    const bool jshared = fl & 2;
    const uint64_t *bj = tab + jl * HN;
    for (int r = lane; r < H; r += 32) {
#pragma unroll
        for (int w = 0; w < NW; w++) {
            uint64_t b = bj[r * NW + w], o = oj[r * NW + w];
            if (r == jy - 1 && w == ((jx - 1) >> 6)) {
                b |= 1ull << ((jx - 1) & 63);
                if (!jshared) o &= ~(1ull << ((jx - 1) & 63));
            }
            if (r == jty - 1 && w == ((jtx - 1) >> 6)) b |= 1ull << ((jtx - 1) & 63);
            vrow[r * NW + w] = b & ~o;
        }
    }
    __syncwarp();
    if (OBS == OBS_CLIP) coop_encode_clip<NW>(dst, vrow, H, W, K, jx, jy, jtx, jty, lane);
    else coop_encode_full<NW>(dst, vrow, H, W, jx, jy, jtx, jty, lane);
    __syncwarp();
}

// ------------------------------------------------------------------ one scene per lane
template <int NW, int RPL, int SPW, int NWK>
struct LaneScene {
    static constexpr int WB = 64 * NW;
    const DevCfg &c;
    const LaneArgs &t;
    int lane, e;
    bool has;
    uint8_t *smem;
    const uint64_t *tab;
    uint64_t *occ, *occm, *vrow;
    uint16_t *pos_s, *tgt_s, *ord_s;
    uint8_t *meta_s;
    uint32_t *hdr_w;
    const uint32_t *tasks_e;
    int HN;
    // header in registers
    uint32_t tick, cursor, n_done;
    int n_created, finished, over;
    uint64_t acted_mask;
    bool overlap;
    bool inflight;  // a request batch is with the worker warps (warp-uniform)
    long long pbfs = 0, pnb = 0;  // profiling builds: cycles / count of BFS sections
    uint32_t cnt_turns, cnt_ended, cnt_bad;
    // rollout kernel: the task word at the scene's cursor is requested as soon as the cursor moves, so
    // that the next stage-0 get_task finds it in a register instead of waiting for HBM on the turn chain
    bool pf = false;
    uint32_t tk_pf = 0;

    __device__ __forceinline__ LaneScene(const DevCfg &cfg, const LaneArgs &args, uint8_t *sm, int lane_, int e0)
        : c(cfg), t(args), lane(lane_), smem(sm), overlap(true), inflight(false), cnt_turns(0), cnt_ended(0), cnt_bad(0) {
        e = e0 + lane;
        has = lane < SPW && e < c.E;
        HN = c.H * NW;
        tab = reinterpret_cast<const uint64_t *>(smem + t.off_tab);
        const int slot = lane < SPW ? lane : 0;
        uint8_t *blk = smem + t.off_blk + (size_t)slot * t.blk_stride;
        hdr_w = reinterpret_cast<uint32_t *>(blk);
        pos_s = reinterpret_cast<uint16_t *>(blk + c.off_pos);
        tgt_s = reinterpret_cast<uint16_t *>(blk + c.off_tgt);
        ord_s = reinterpret_cast<uint16_t *>(blk + c.off_ord);
        meta_s = blk + c.off_meta;
        occ = reinterpret_cast<uint64_t *>(smem + t.off_occ) + (size_t)slot * t.occ_stride;
        occm = reinterpret_cast<uint64_t *>(smem + t.off_occm) + (size_t)slot * t.occ_stride;
        vrow = reinterpret_cast<uint64_t *>(smem + t.off_vrow);
        tasks_e = c.tasks + (size_t)(has ? e : 0) * c.T;
    }
    __device__ __forceinline__ const uint64_t *occ_of(int slot) const {
        return reinterpret_cast<const uint64_t *>(smem + t.off_occ) + (size_t)slot * t.occ_stride;
    }

    // ---- cooperative prologue / epilogue
    __device__ __forceinline__ void load_tables() {
        uint64_t *tb = reinterpret_cast<uint64_t *>(smem + t.off_tab);
        for (int q = lane; q < HN; q += 32) reinterpret_cast<uint64_t *>(smem + t.off_zero)[q] = 0ull;
        for (int r = lane; r < c.H; r += 32) {
            Row<NW> b0, b1;
#pragma unroll
            for (int i = 0; i < NW; i++) {
                const uint64_t inb = __ldg(c.inb_rows + r * NW + i), pk = __ldg(c.picking_rows + r * NW + i),
                               st = __ldg(c.storage_rows + r * NW + i);
                b0.w[i] = inb & ~pk;
                b1.w[i] = b0.w[i] & ~st;
            }
            const Row<NW> m0 = rmirror(b0), m1 = rmirror(b1);
#pragma unroll
            for (int i = 0; i < NW; i++) {
                tb[0 * HN + r * NW + i] = b0.w[i];
                tb[1 * HN + r * NW + i] = b1.w[i];
                if (mirror_copy<NW>()) {
                    tb[2 * HN + r * NW + i] = m0.w[i];
                    tb[3 * HN + r * NW + i] = m1.w[i];
                }
            }
        }
    }
    __device__ __forceinline__ void load_blocks(int e0) {
        // 8 scene blocks (<= 32 words per lane) in flight at a time: the copy is latency-bound
        // (one DRAM round trip per batch), so batches of independent loads, not a load-store loop
        const int nsc = min(SPW, c.E - e0), wps = c.env_stride >> 2;  // wps <= 120 (N <= 64)
        const uint32_t *src = reinterpret_cast<const uint32_t *>(c.state + (size_t)e0 * c.env_stride);
        for (int j0 = 0; j0 < nsc; j0 += 8) {
            uint32_t v[8][4];
#pragma unroll
            for (int jj = 0; jj < 8; jj++)
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int w = lane + 32 * q;
                    v[jj][q] = (j0 + jj < nsc && w < wps) ? __ldg(src + (j0 + jj) * wps + w) : 0u;
                }
#pragma unroll
            for (int jj = 0; jj < 8; jj++) {
                uint32_t *dst = reinterpret_cast<uint32_t *>(smem + t.off_blk + (size_t)(j0 + jj) * t.blk_stride);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int w = lane + 32 * q;
                    if (j0 + jj < nsc && w < wps) dst[w] = v[jj][q];
                }
            }
        }
        __syncwarp();
    }
    __device__ __forceinline__ void store_blocks(int e0) {
        __syncwarp();
        const int nsc = min(SPW, c.E - e0), wps = c.env_stride >> 2;
        uint32_t *dst = reinterpret_cast<uint32_t *>(c.state + (size_t)e0 * c.env_stride);
        for (int j = 0; j < nsc; j++) {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(smem + t.off_blk + (size_t)j * t.blk_stride);
#pragma unroll 4
            for (int w = lane; w < wps; w += 32) dst[j * wps + w] = src[w];
        }
    }
    // blocks are only 4-byte aligned in shared memory (odd word stride): the header is accessed as
    // 32-bit words, never through EnvHdr (whose 8-byte alignment the compiler would rely on)
    __device__ __forceinline__ void hdr_to_regs() {
        const uint32_t *h = hdr_w;
        tick = h[0]; cursor = h[1]; n_done = h[2];
        const uint32_t w3 = h[3];
        n_created = w3 & 0xFFFF; finished = (w3 >> 16) & 0xFF; over = (w3 >> 24) & 0xFF;
        acted_mask = (uint64_t)h[4] | ((uint64_t)h[5] << 32);
    }
    __device__ __forceinline__ void regs_to_hdr() {
        uint32_t *h = hdr_w;
        h[0] = tick; h[1] = cursor; h[2] = n_done;
        h[3] = (uint32_t)(n_created & 0xFFFF) | ((uint32_t)(finished & 0xFF) << 16) | ((uint32_t)(over & 0xFF) << 24);
        h[4] = (uint32_t)acted_mask; h[5] = (uint32_t)(acted_mask >> 32);
    }

    // ---- occupancy (cells are 0-based here)
    __device__ __forceinline__ void occ_set(int x0, int y0) {
        occ[y0 * NW + (x0 >> 6)] |= 1ull << (x0 & 63);
        if (!mirror_copy<NW>()) return;
        const int xm = WB - 1 - x0;
        occm[y0 * NW + (xm >> 6)] |= 1ull << (xm & 63);
    }
    __device__ __forceinline__ void occ_clear(int x0, int y0) {
        occ[y0 * NW + (x0 >> 6)] &= ~(1ull << (x0 & 63));
        if (!mirror_copy<NW>()) return;
        const int xm = WB - 1 - x0;
        occm[y0 * NW + (xm >> 6)] &= ~(1ull << (xm & 63));
    }
    // grid code of an in-bounds cell (0 track, 1 storage, 2 picking) from the base rows in shared
    // memory: with four 54 KB CTAs on an SM there is next to no L1 left for a global grid lookup
    __device__ __forceinline__ int cell_code(int x0, int y0) const {
        const int q = y0 * NW + (x0 >> 6), sh = x0 & 63;
        const bool b0 = (tab[q] >> sh) & 1ull, b1 = (tab[HN + q] >> sh) & 1ull;
        return b0 ? (b1 ? 0 : 1) : 2;
    }
    // AGV moves (x0,y0) -> (x1,y1): the four row loads (plain + mirrored copy) issue together
    __device__ __forceinline__ void occ_move(int x0, int y0, int x1, int y1, bool clear_old) {
        const int q0 = y0 * NW + (x0 >> 6), q1 = y1 * NW + (x1 >> 6);
        const int xm0 = WB - 1 - x0, xm1 = WB - 1 - x1;
        const int p0 = y0 * NW + (xm0 >> 6), p1 = y1 * NW + (xm1 >> 6);
        const uint64_t b0 = clear_old ? (1ull << (x0 & 63)) : 0ull, b1 = 1ull << (x1 & 63);
        const uint64_t c0 = clear_old ? (1ull << (xm0 & 63)) : 0ull, c1 = 1ull << (xm1 & 63);
        uint64_t r0 = occ[q0], r1 = occ[q1], m0 = 0ull, m1 = 0ull;
        if (mirror_copy<NW>()) { m0 = occm[p0]; m1 = occm[p1]; }
        if (q0 == q1) { r1 = (r1 & ~b0) | b1; occ[q1] = r1; }
        else { occ[q0] = r0 & ~b0; occ[q1] = r1 | b1; }
        if (mirror_copy<NW>()) {
            if (p0 == p1) { m1 = (m1 & ~c0) | c1; occm[p1] = m1; }
            else { occm[p0] = m0 & ~c0; occm[p1] = m1 | c1; }
        }
    }
    // other_at for the AGV whose (unmoved) position `mypos` is already in registers
    __device__ __forceinline__ bool other_at_from(int i, int cell, int mypos) const {
        if (!overlap) {
            if (mypos == cell) return false;
            return occ_test((cell & 255) - 1, (cell >> 8) - 1);
        }
        return other_at(i, cell);
    }
    __device__ __forceinline__ bool occ_test(int x0, int y0) const {
        return (occ[y0 * NW + (x0 >> 6)] >> (x0 & 63)) & 1ull;
    }
    __device__ __forceinline__ void build_occ() {
        // ORs without a returned value (no load -> test -> store chain per AGV); two created AGVs on
        // one cell show up as fewer set bits than AGVs
        for (int q = 0; q < HN; q++) { occ[q] = 0ull; if (mirror_copy<NW>()) occm[q] = 0ull; }
        for (int j = 0; j < n_created; j++) {
            const int p = pos_s[j], x0 = (p & 255) - 1, y0 = (p >> 8) - 1;
            // (32-bit halves: the 64-bit shared-memory OR is a compare-and-swap loop)
            atomicOr(reinterpret_cast<unsigned *>(occ) + y0 * NW * 2 + (x0 >> 5), 1u << (x0 & 31));
            if (mirror_copy<NW>()) {
                const int xm = WB - 1 - x0;
                atomicOr(reinterpret_cast<unsigned *>(occm) + y0 * NW * 2 + (xm >> 5), 1u << (xm & 31));
            }
        }
        int bits = 0;
        for (int q = 0; q < HN; q++) bits += __popcll(occ[q]);
        overlap = bits != n_created;
    }

    __device__ __forceinline__ Agv load_agv(int i) const {
        Agv a;
        const int p = pos_s[i], tg = tgt_s[i], m = meta_s[i];
        a.x = p & 255; a.y = p >> 8; a.tx = tg & 255; a.ty = tg >> 8;
        a.stage = m & 3; a.loaded = (m >> 2) & 1; a.idle = (m >> 3) & 1; a.order = ord_s[i];
        return a;
    }
    __device__ __forceinline__ void store_agv(int i, const Agv &a) {
        pos_s[i] = (uint16_t)(a.x | (a.y << 8));
        tgt_s[i] = (uint16_t)(a.tx | (a.ty << 8));
        ord_s[i] = (uint16_t)a.order;
        meta_s[i] = (uint8_t)meta_pack(a);
    }
    // any created AGV j != i standing on cell (x | y<<8), 1-based
    __device__ __forceinline__ bool other_at(int i, int cell) const {
        if (!overlap) {
            if (pos_s[i] == cell) return false;
            return occ_test((cell & 255) - 1, (cell >> 8) - 1);
        }
        bool m = false;
        for (int j = 0; j < n_created; j++) m |= (j != i) && (pos_s[j] == cell);
        return m;
    }

    // Explorer.load_condition / get_task / continue_working (Explorer.py:68-116,251-259)
    __device__ __forceinline__ void load_condition(Agv &a) const {
        if (c.always_loaded) a.loaded = 1;
        else if (c.always_empty) a.loaded = 0;
        else a.loaded = (a.stage == 0) ? 0 : 1;
    }
    __device__ __forceinline__ void get_task(Agv &a) {
        if ((uint32_t)c.T == cursor && a.stage == 0) {
            a.idle = 1;
            if ((uint32_t)c.T == n_done) finished = 1;
            return;
        }
        uint32_t tk;
        if (a.stage == 0) {
            a.order = (int)(cursor++);
            if (pf) { tk = tk_pf; tk_pf = __ldg(tasks_e + (cursor < (uint32_t)c.T ? cursor : 0u)); }
            else tk = __ldg(tasks_e + a.order);
        } else tk = __ldg(tasks_e + a.order);
        if (a.stage == 1) { a.tx = (tk >> 16) & 255; a.ty = (tk >> 24) & 255; }
        else { a.tx = tk & 255; a.ty = (tk >> 8) & 255; }
        load_condition(a);
    }
    __device__ __forceinline__ void prefetch_cursor_task() {
        if (pf) tk_pf = __ldg(tasks_e + (cursor < (uint32_t)c.T ? cursor : 0u));
    }
    __device__ __forceinline__ void continue_working(Agv &a) {
        a.stage++;
        if (a.stage == 3) { a.stage = 0; n_done++; }
        get_task(a);
    }
    // Scene.init (Scene.py:61-65) + create_explorer of AGV 0 (Scene.py:86)
    __device__ __forceinline__ void reset() {
        tick = 0; cursor = 0; n_done = 0; finished = 0; over = 0; acted_mask = 0;
        prefetch_cursor_task();
        for (int j = 0; j < c.N; j++) { pos_s[j] = (uint16_t)(1 | (1 << 8)); ord_s[j] = 0; meta_s[j] = 0; }
        n_created = 0;
        if (c.N > 0) {
            n_created = 1;
            Agv a = load_agv(0);
            get_task(a);
            store_agv(0, a);
        }
    }
    // Scene.check_new_veh, Scene.py:354-366
    __device__ __forceinline__ void spawn() {
        if (n_created >= c.N) return;
        if (occ_test(0, 0)) return;
        const int k = n_created++;
        Agv a = load_agv(k);
        get_task(a);
        store_agv(k, a);
        occ_set(a.x - 1, a.y - 1);
    }

    // ---- FindPathAstar's first action + length for the lanes with `need` (all lanes call).
    // matrix row r = (base[loaded][r] | target) & ~(occ[r] & om); coordinates 0-based.
    __device__ __forceinline__ void search(bool need, int loaded, uint64_t om, int sx, int sy, int tx, int ty,
                                           int &action, int &len) {
        if (need && sx == tx && sy == ty) { action = A_STOP; len = 0; need = false; }
        int l = 0;
        const uint64_t *zr = reinterpret_cast<const uint64_t *>(smem + t.off_zero);
        const int a = lane_monotone<NW>(need, tab, HN, om ? occ : zr, om ? occm : zr, loaded, sx, sy, tx, ty, l);
        if (need && a >= 0) { action = a; len = l; need = false; }
        // the rest needs the level-synchronous BFS: warp-cooperative, one scene at a time
        unsigned pend = __ballot_sync(FULL, need);
        if (!pend) return;
#ifdef AGV_LANES_PROF
        const long long pb0_ = clock64();
        struct Acc { long long &a, &n; long long t0; __device__ ~Acc() { a += clock64() - t0; n++; } } acc_{pbfs, pnb, pb0_};
#endif
        const unsigned pk = (unsigned)sx | ((unsigned)sy << 7) | ((unsigned)tx << 14) | ((unsigned)ty << 21) |
                            ((unsigned)loaded << 28) | (om ? (1u << 29) : 0u);
        if (NWK > 0) {  // handed to the worker warps of the CTA
            LaneQueue *q = queue();
            wait_done();
            const int slot = __popc(pend & ((1u << lane) - 1u));
            if (need) q->req[slot] = make_uint4(pk, (unsigned)lane, 0u, 0u);
            post(Q_BFS, __popc(pend), 0, 0);
            wait_done();
            if (need) { const unsigned r = q->res[slot]; action = (int)(r & 0xFFu); len = (int)(r >> 8); }
            return;
        }
        while (pend) {
            const int j = __ffs(pend) - 1;
            pend &= pend - 1;
            int bl = 0;
            const int ba = coop_bfs<NW, RPL>(__shfl_sync(FULL, pk, j), tab, HN, occ_of(j), c.H, c.W, lane, bl);
            if (lane == j) { action = ba; len = bl; }
        }
    }

    // ---- observation of AGV `a` (index i) of the lanes with `want`, written to base + rec * R
    // matrix (b): StateManager.create_path_matrix, StateManager.py:43-61
    template <int OBS>
    __device__ __forceinline__ void encode(bool want, const Agv &a, int i, int which) {
        bool shared = false;
        if (want && overlap) shared = other_at(i, a.x | (a.y << 8));
        unsigned pend = __ballot_sync(FULL, want);
        if (!pend) return;
        uint16_t *base = which ? t.next_obs : t.obs;
        const int slot = __popc(pend & ((1u << lane) - 1u));
        if (OBS == OBS_CLIP && c.K == 7) {
            // lane-local window extraction; the expansion to 147 bf16 and the stores are warp-wide
            uint64_t wm = 0ull;
            if (want) wm = lane_window7<NW>(tab + a.loaded * HN, occ, c.H, a.x - 1, a.y - 1, a.tx - 1, a.ty - 1, shared);
            const int tpos = (max(-3, min(3, a.ty - a.y)) + 3) * 7 + max(-3, min(3, a.tx - a.x)) + 3;
            if (NWK > 0) {  // fire and forget: the request carries everything the record needs
                LaneQueue *q = queue();
                wait_done();
                if (want) q->req[slot] = make_uint4((unsigned)wm, (unsigned)(wm >> 32), (unsigned)tpos, (unsigned)lane);
                post(Q_ENC7, __popc(pend), i, which);
                return;
            }
            while (pend) {
                const int j = __ffs(pend) - 1;
                pend &= pend - 1;
                const uint64_t jwm = __shfl_sync(FULL, wm, j);
                const int jt = __shfl_sync(FULL, tpos, j);
                expand_clip7(base + ((size_t)(e - lane + j) * c.N + i) * 147, jwm, jt, lane);
            }
            return;
        }
        const unsigned pk = (unsigned)a.x | ((unsigned)a.y << 8) | ((unsigned)a.tx << 16) | ((unsigned)a.ty << 24);
        const unsigned fl = (unsigned)a.loaded | (shared ? 2u : 0u);
        if (NWK > 0) {  // the workers read this scene's occupancy rows: wait before the move is applied
            LaneQueue *q = queue();
            wait_done();
            if (want) q->req[slot] = make_uint4(pk, fl, 0u, (unsigned)lane);
            post(Q_ENCG, __popc(pend), i, which);
            wait_done();
            return;
        }
        while (pend) {
            const int j = __ffs(pend) - 1;
            pend &= pend - 1;
            const unsigned q = __shfl_sync(FULL, pk, j), f = __shfl_sync(FULL, fl, j);
            uint16_t *dst = base + ((size_t)(e - lane + j) * c.N + i) * t.rec_elems;
            coop_encode<NW, OBS>(q, f, tab, HN, occ_of(j), vrow, c.H, c.W, c.K, dst, lane);
        }
    }

    // ---- request queue to the worker warps (NWK > 0)
    __device__ __forceinline__ LaneQueue *queue() const { return reinterpret_cast<LaneQueue *>(smem + t.off_queue); }
    __device__ __forceinline__ void wait_done() {
        if (inflight) { named_bar_sync(BAR_DONE, 32 * (1 + NWK)); inflight = false; }
    }
    __device__ __forceinline__ void post(int kind, int n, int i, int which) {
        if (lane == 0) {
            LaneQueue *q = queue();
            q->kind = kind; q->n = n; q->i = i; q->which = which;
        }
        __threadfence_block();
        named_bar_arrive(BAR_GO, 32 * (1 + NWK));
        inflight = true;
    }
};

// worker warp of a lane-per-scene CTA: serves BFS and record-expansion requests
template <int NW, int RPL, int OBS, int SPW, int NWK>
__device__ __forceinline__ void lane_worker(const DevCfg &c, const LaneArgs &t, uint8_t *smem, int wid) {
    const int lane = threadIdx.x & 31;
    const int e0 = blockIdx.x * SPW, HN = c.H * NW;
    LaneQueue *q = reinterpret_cast<LaneQueue *>(smem + t.off_queue);
    const uint64_t *tab = reinterpret_cast<const uint64_t *>(smem + t.off_tab);
    const uint64_t *occ0 = reinterpret_cast<const uint64_t *>(smem + t.off_occ);
    uint64_t *vrow = reinterpret_cast<uint64_t *>(smem + t.off_vrow) + (size_t)(wid + 1) * HN;
    for (;;) {
        named_bar_sync(BAR_GO, 32 * (1 + NWK));
        const int kind = q->kind, n = q->n;
        if (kind == Q_EXIT) return;
        if (kind == Q_BFS) {
            for (int r = wid; r < n; r += NWK) {
                const uint4 rq = q->req[r];
                int bl = 0;
                const int ba = coop_bfs<NW, RPL>(rq.x, tab, HN, occ0 + (size_t)rq.y * t.occ_stride, c.H, c.W, lane, bl);
                if (lane == 0) q->res[r] = (unsigned)ba | ((unsigned)bl << 8);
            }
        } else {
            uint16_t *base = q->which ? t.next_obs : t.obs;
            const int i = q->i;
            for (int r = wid; r < n; r += NWK) {
                const uint4 rq = q->req[r];
                uint16_t *dst = base + ((size_t)(e0 + (int)rq.w) * c.N + i) * t.rec_elems;
                if (kind == Q_ENC7) expand_clip7(dst, (uint64_t)rq.x | ((uint64_t)rq.y << 32), (int)rq.z, lane);
                else coop_encode<NW, OBS>(rq.x, rq.y, tab, HN, occ0 + (size_t)rq.w * t.occ_stride, vrow, c.H, c.W, c.K, dst, lane);
            }
        }
        __threadfence_block();
        named_bar_arrive(BAR_DONE, 32 * (1 + NWK));
    }
}

// Profiling builds (-DAGV_LANES_PROF): clock64 phase counters of the scene warp, summed over CTAs into
// counters[3..] (dumped by agv_read_counters when AGV_PROF_DUMP is set): [3+k] cycles of phase k
// (0 prologue, 1 loop head, 2 obs encode, 3 search, 4 check/move/outputs, 5 next_obs encode,
// 7 epilogue), [11] cycles inside BFS sections, [12] their count, [13] iterations, [15] CTAs,
// [16..19] maxima over CTAs of total / search / BFS cycles / BFS sections.
#ifdef AGV_LANES_PROF
#define PROF_DECL long long pt_[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long pc_ = clock64();
#define PROF(k) { const long long n_ = clock64(); pt_[k] += n_ - pc_; pc_ = n_; }
#define PROF_FLUSH                                                                                   \
    pt_[8] = ws.pbfs; pt_[9] = ws.pnb; pt_[10] = n_max;                                              \
    if (lane == 0) {                                                                                 \
        long long tot_ = 0;                                                                          \
        for (int k_ = 0; k_ < 8; k_++) tot_ += pt_[k_];                                              \
        atomicMax(c.counters + 16, (unsigned long long)tot_);                                        \
        atomicMax(c.counters + 17, (unsigned long long)pt_[3]);                                      \
        atomicMax(c.counters + 18, (unsigned long long)ws.pbfs);                                     \
        atomicMax(c.counters + 19, (unsigned long long)ws.pnb);                                      \
        for (int k_ = 0; k_ < 12; k_++) atomicAdd(c.counters + 3 + k_, (unsigned long long)pt_[k_]); \
        atomicAdd(c.counters + 15, 1ull);                                                            \
    }
#else
#define PROF_DECL
#define PROF(k)
#define PROF_FLUSH
#endif

// One pass of Scene.run_mode's loop body (Scene.py:111-164) for SPW scenes per warp.
template <int NW, int RPL, int OBS, bool SHAPED, int SPW, int NWK>
__device__ __forceinline__ void tick_lanes(const DevCfg &c, const LaneArgs &t, uint8_t *smem) {
    const int lane = threadIdx.x & 31;
    const int e0 = blockIdx.x * SPW;
    LaneScene<NW, RPL, SPW, NWK> ws(c, t, smem, lane, e0);
    PROF_DECL
    ws.load_tables();
    ws.load_blocks(e0);
    const int N = c.N;
    {   // per-AGV outputs default to "no turn"
        const int nsc = min(SPW, c.E - e0);
        const size_t r0 = (size_t)e0 * N;
        for (int q = lane; q < nsc * N; q += 32) {
            if (t.act_out) t.act_out[r0 + q] = -1;
            if (t.done) t.done[r0 + q] = 0;
            if (t.reward) t.reward[r0 + q] = 0.f;
            if (t.plen) t.plen[r0 + q] = 0;
            if (t.slen) t.slen[r0 + q] = 0xFFFF;
        }
        __syncwarp();
    }
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
    const int n_max = __reduce_max_sync(FULL, n_loop);
    const uint64_t om_a = N > 1 ? ~0ull : 0ull;  // matrix (a) blocks the explorers only when group size > 1
    const bool want_obs = OBS != OBS_NONE && t.obs != nullptr, want_next = OBS != OBS_NONE && t.next_obs != nullptr;

    PROF(0)
    for (int i = 0; i < n_max; i++) {
        bool act = false;
        if (i < n_loop && !ws.over) {
            if (ws.finished) { ws.over = 1; ws.cnt_ended++; }        // Scene.py:129-130
            else act = !((ws.meta_s[i] >> 3) & 1);                   // Scene.py:132-134
        }
        if (!__any_sync(FULL, act)) continue;
        Agv a;
        a.x = a.y = a.tx = a.ty = 1; a.stage = a.loaded = a.idle = a.order = 0;
        int given = -1;
        if (act) {
            a = ws.load_agv(i);
            if (t.policy == POLICY_GIVEN && t.action) given = (int)t.action[(size_t)ws.e * N + i];
        }
        PROF(1)
        if (want_obs) ws.template encode<OBS>(act, a, i, 0);
        PROF(2)
        // ---- action
        const bool guided = (t.policy == POLICY_GUIDED) || (t.policy == POLICY_GIVEN && given < 0);
        const bool pol_a = !guided && t.policy == POLICY_ASTAR;
        int action = given, plen = 0;
        {
            // matrix (a) also blocks (1,1) while uncreated AGVs park there (Explorer.py:274-283)
            const bool guard = act && pol_a && N > 1 && ws.n_created < N && !ws.occ_test(0, 0);
            if (guard) ws.occ_set(0, 0);
            __syncwarp();
            ws.search(act && (guided || pol_a), a.loaded, guided ? ~0ull : om_a, a.x - 1, a.y - 1, a.tx - 1, a.ty - 1,
                      action, plen);
            __syncwarp();
            if (guard) ws.occ_clear(0, 0);
        }
        PROF(3)
        if (act && !guided && !pol_a && action > A_STOP) { ws.cnt_bad++; action = A_STOP; }
        // ---- Explorer.check_action, Explorer.py:151-208
        const int dx = (action == A_RIGHT) - (action == A_LEFT);
        const int dy = (action == A_DOWN) - (action == A_UP);
        const int nx = a.x + dx, ny = a.y + dy;
        int r = 0, end = 0, slen = 0xFFFF;
        bool want_shaped = false;
        if (act) {
            const bool oob = nx < 1 || ny < 1 || nx > c.W || ny > c.H;
            if (oob) { r = -1; end = 1; }
            else {
                const int code = ws.cell_code(nx - 1, ny - 1);
                const bool at_target = (nx == a.tx) && (ny == a.ty);
                if (code == 1 && a.loaded && !at_target) { r = -1; end = 1; }
                else if (code == 2 && !at_target) { r = -1; end = 1; }
                else if (t.proto == PROTO_INTELLIGENT && ws.other_at(i, nx | (ny << 8))) { r = -1; end = 1; }
                else if (at_target) { r = 1; ws.continue_working(a); }
                else want_shaped = SHAPED;
            }
        }
        double rshaped = 0.0;
        if (SHAPED) {
            // Explorer.rectify_reward (Explorer.py:210-249): only path_length_new matters
            const bool guard = want_shaped && N > 1 && ws.n_created < N && !ws.occ_test(0, 0);
            if (guard) ws.occ_set(0, 0);
            __syncwarp();
            int sa = 0, l = 0;
            ws.search(want_shaped, a.loaded, om_a, nx - 1, ny - 1, a.tx - 1, a.ty - 1, sa, l);
            __syncwarp();
            if (guard) ws.occ_clear(0, 0);
            if (want_shaped) {
                slen = l;
                const double M = 0.5 * (double)(c.H + c.W);
                rshaped = 0.001 * (M - (double)l) / M;
            }
        }
        if (act) {
            // ---- Explorer.py:145-148: the move is applied iff not is_end
            if (t.proto != PROTO_INTELLIGENT && !guided && !pol_a) ws.overlap = true;
            if (!end && (dx | dy)) {
                const bool still = ws.other_at(i, a.x | (a.y << 8));
                if (!still) ws.occ_clear(a.x - 1, a.y - 1);
                a.x = nx; a.y = ny;
                ws.occ_set(nx - 1, ny - 1);
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
            if (t.reward) t.reward[rec] = want_shaped ? (float)rshaped : (float)r;
            if (t.plen) t.plen[rec] = (uint16_t)plen;
            if (t.slen) t.slen[rec] = (uint16_t)slen;
        }
        __syncwarp();
        PROF(4)
        if (want_next) ws.template encode<OBS>(act, a, i, 1);
        PROF(5)  // store_info's create_state (Scene.py:156)
    }
    PROF(6)
    if (NWK > 0) { ws.wait_done(); ws.post(Q_EXIT, 0, 0, 0); }
    if (live && !ws.over) ws.spawn();
    if (ws.has) ws.regs_to_hdr();
    {
        const unsigned ct = __reduce_add_sync(FULL, ws.cnt_turns), ce = __reduce_add_sync(FULL, ws.cnt_ended),
                       cb = __reduce_add_sync(FULL, ws.cnt_bad);
        if (lane == 0) {
            if (ct) atomicAdd(c.counters + 0, (unsigned long long)ct);
            if (ce) atomicAdd(c.counters + 1, (unsigned long long)ce);
            if (cb) atomicAdd(c.counters + 2, (unsigned long long)cb);
        }
    }
    ws.store_blocks(e0);
    PROF(7)
    PROF_FLUSH
}

}  // namespace agv
