// agv_rollout.cu -- kernels + launcher of the persistent multi-tick rollout (agv_rollout.cuh).
// sm_100a only.  Called from agv_rollout / agv_tick (agv_kernels.cu).
#include "agv_rollout.cuh"

#include <cstdio>
#include <cstdlib>

namespace agv {
void note_launch();
void note_cuda_error(int e);

// Resident CTAs the register allocation has to leave room for.  The <= 64 x 64 kernels run 2 CTAs per SM (launch
// shapes below); compiled for 4 they got 80 registers with 3 workers -- and 56 with 6, spilling.  Measured at c3
// with 3 workers: room for 4 CTAs (80 registers) 1.34e9 guided / 1.24e9 with per-AGV actions, for 3 (96 registers)
// 1.41e9 / 1.33e9, for 2 (108 registers, the compiler's own choice) 1.38e9 / 1.35e9.
#ifndef AGV_ROLL_MINCTAS
#define AGV_ROLL_MINCTAS 0
#endif
template <int NW, int RPL, int NB, int MODE>
constexpr int rollout_min_ctas() {
    return (NW == 1 && RPL <= 2) ? (AGV_ROLL_MINCTAS ? AGV_ROLL_MINCTAS : (NB >= 4 || MODE == 3 || MODE == 0 ? 2 : 3))
                         : (NW * RPL <= 4 ? 2 : 1);
}

template <int NW, int RPL, int OBS, int SPW, int NB, int MODE>
__global__ void __launch_bounds__(32 * (3 + NB), rollout_min_ctas<NW, RPL, NB, MODE>())
k_rollout(const DevCfg c, const RollArgs ra) {
    extern __shared__ __align__(16) uint8_t lane_smem[];
    rollout_cta<NW, RPL, OBS, SPW, NB, MODE>(c, ra, lane_smem);
}

// Longest-first scene order: counting sort of the scenes by the length of their last tick -- turns taken
// plus BFS searches (a BFS round trip costs about as much as a turn), scene header word 6 = BFS count |
// turns << 16 -- in 64 buckets, longest first.  One CTA; the order inside a bucket does not matter (scenes
// are independent: the schedule never changes a result).
__device__ __forceinline__ unsigned order_bucket(unsigned w) {
    return 63u - min(((w & 0xFFFFu) + (w >> 16)) >> 1, 63u);
}

__global__ void __launch_bounds__(1024) k_rollout_order(const DevCfg c, unsigned *order, unsigned *qctr, unsigned q0,
                                                        unsigned long long *stat) {
    __shared__ unsigned hist[64], base[64];
    __shared__ unsigned long long load[2];
    if (threadIdx.x < 64) hist[threadIdx.x] = 0u;
    if (threadIdx.x < 2) load[threadIdx.x] = 0ull;
    __syncthreads();
    unsigned n_bfs = 0, n_turns = 0;
    for (int e = threadIdx.x; e < c.E; e += blockDim.x) {
        const unsigned w = *reinterpret_cast<const unsigned *>(c.state + (size_t)e * c.env_stride + 24);
        atomicAdd(&hist[order_bucket(w)], 1u);
        n_bfs += w & 0xFFFFu; n_turns += w >> 16;
    }
    n_bfs = __reduce_add_sync(FULL, n_bfs); n_turns = __reduce_add_sync(FULL, n_turns);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&load[0], (unsigned long long)n_bfs); atomicAdd(&load[1], (unsigned long long)n_turns); }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned s = 0;
        for (int b = 0; b < 64; b++) { base[b] = s; s += hist[b]; }
        *qctr = q0;
        if (stat) {  // read by the host before the next launch, without synchronisation: a hint, never a result
            reinterpret_cast<volatile unsigned long long *>(stat)[0] = load[0];
            reinterpret_cast<volatile unsigned long long *>(stat)[1] = load[1];
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < c.E; e += blockDim.x) {
        const unsigned w = *reinterpret_cast<const unsigned *>(c.state + (size_t)e * c.env_stride + 24);
        order[atomicAdd(&base[order_bucket(w)], 1u)] = (unsigned)e;
    }
}

static inline int align8r(int x) { return (x + 7) & ~7; }

template <int NW, int SPW>
static int plan_rollout(const DevCfg &dc, LaneArgs &t) {
    const int HN = dc.H * NW;
    int off = 0;
    t.off_tab = off; off += (mirror_copy<NW>() ? 4 : 2) * HN * 8;
    t.blk_stride = dc.env_stride + 4;  // odd number of words: same-field accesses of the 32 lanes hit 32 banks
    t.off_blk = off; off = align8r(off + SPW * t.blk_stride);
    t.occ_stride = HN | 1;
    t.off_occ = off; off += SPW * t.occ_stride * 8;
    t.off_occm = off; off += mirror_copy<NW>() ? SPW * t.occ_stride * 8 : 0;
    t.off_vrow = off;
    t.off_zero = off; off += HN * 8;
    t.off_queue = (off + 15) & ~15; off = t.off_queue + (int)sizeof(RollShared);
    return off;
}

static int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

template <int NW, int RPL, int OBS, int SPW, int NB, int MODE>
static int launch_rollout(const DevCfg &dc, RollArgs ra, unsigned *order, cudaStream_t st, int cps_want = 0) {
    const int smem = plan_rollout<NW, SPW>(dc, ra.la);
    auto kern = k_rollout<NW, RPL, OBS, SPW, NB, MODE>;
    const int threads = 32 * (3 + NB);  // 2 scene warps, NB BFS workers, the record writer
    if (smem > 227 * 1024) return -4;  // AGV_E_UNSUPPORTED
    static int cps_cached = 0, sms = 0, smem_cached = -1;  // per template instance (one device per process)
    if (smem != smem_cached) {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) { note_cuda_error((int)e); return -2; }
        }
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        int dev = 0, n = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, threads, smem);
        if (e != cudaSuccess || n < 1) { note_cuda_error((int)e); return -2; }
        cps_cached = n;
        smem_cached = smem;
    }
    // Persistent CTAs, never more than the scenes need.  One CTA slot per SM is left free when there are three
    // or more: co-resident CTAs slow each other's dependent turn chains more than the scene queue costs
    // (c3, 16384 scenes on 148 SMs: 4 per SM = all scenes resident, 1.10e9 AGV-steps/s; 3 per SM = 13 % of the
    // scenes -- the shortest ones -- wait in the queue, 1.15e9; 2 per SM 0.97e9).  AGV_ROLLOUT_CPS overrides.
    int cps = env_int("AGV_ROLLOUT_CPS", cps_want);
    if (cps < 1 || cps > cps_cached) cps = cps_cached >= 3 ? cps_cached - 1 : cps_cached;
    int ctas = (dc.E + SPW - 1) / SPW;
    if (ctas > sms * cps) ctas = sms * cps;
    k_rollout_order<<<1, 1024, 0, st>>>(dc, order, ra.qctr, (unsigned)ctas * SPW, ra.stat);
    note_launch();
    ra.order = order;
    kern<<<ctas, threads, smem, st>>>(dc, ra);
    note_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { note_cuda_error((int)e); return -2; }
    return 0;
}

// protocol / policy specialisations (agv_rollout.cuh, MODE): the two modes the rollout exists for get their
// own kernels; open-loop actions and mixed settings run on the generic one.
//
// Launch shape = BFS worker warps per CTA x CTAs per SM.  The best shape depends on how many turns need the
// level-synchronous BFS (the row DP decides the others inside the scene warp): with few searches the scene
// warps' turn chains bind and idle workers only compete for issue slots; with many, the searches bind.  Maps
// up to 64 x 64 have two shapes, picked per launch by the load of every scene's last tick -- searches per 100 turns,
// summed by the previous launch's order kernel into pinned host memory (a hint: the schedule never changes a
// result).  Measured, AGV-steps/s (light / heavy shape):
//   64 x 64 (c3, 16 ticks), light = 3 workers x 2 CTAs, heavy = 8 x 2:
//       guided rollout (10 searches per 100 turns)       1.41e9 / 0.95e9     (2 x 3: 1.28e9, 3 x 3: 1.17e9, 2 x 2:
//                                                         1.05e9, 4 x 2: 1.23e9, 4 x 1: 1.07e9)
//       A_star control mode (54 per 100)                 4.4e8 / 6.6e8       (heavy with 6 workers: 6.3e8, 10: 6.7e8)
//   128 x 128 (c5, 8 ticks; 4 rows x 2 words per lane, 144 registers, 95 KB shared memory per CTA): one shape, 3
//   workers x 2 CTAs and no heavy scenes: guided 3.30e8, A_star control mode 2.27e8, Expert collection 3.46e8.
//   (2 workers x 2: 2.86e8 / 1.88e8; 4 or 5 workers fit only one CTA per SM: 2.2e8 - 2.5e8 / 1.9e8.  With scenes
//   moved to the heavy warp from 16 searches per tick the same shape gave 2.3e8 / 1.6e8 -- that, not the worker
//   count, was what an earlier two-shape scheme for this class worked around.)
// WHERE the warps land matters as much as how many there are: the hardware gives a CTA consecutive warp slots and
// slot % 4 is the SM sub-partition (tuning/micro/placement.cu prints %smid / %warpid: CTAs b and b + 148 share an
// SM, the second one starts at the next free slot).  Two 6-warp CTAs put {worker, heavy, worker}, {worker, LIGHT,
// writer}, {worker, worker, heavy}, {writer, worker, LIGHT} on the four sub-partitions.  The same twelve warps
// padded to 8-warp CTAs (both light warps meet two workers each): 1.10e9; light warps kept clear of every worker,
// the six workers on two sub-partitions: 1.21e9; the natural 6-warp layout: 1.34e9.  (With the final register
// budget: natural 1.41e9; each light warp with exactly one worker and nothing else, roles taken from %warpid:
// 1.06e9 -- every 8-warp arrangement tried lost 15-25 %, so slot % 4 is at best part of the story.)
// Map classes in between (64 x 128, 128 x 64) run 3 workers and no heavy scenes either; AGV_ROLL_NB fixes the
// worker count (tuning builds),
// AGV_ROLLOUT_SHAPE=1|2 forces light | heavy (tests), AGV_ROLLOUT_CPS the CTAs per SM.
#ifndef AGV_ROLL_SPW
#define AGV_ROLL_SPW 32   // scene slots per CTA = lanes of a scene warp in use (tuning builds: 16)
#endif
#ifndef AGV_ROLL_NB_HEAVY
#define AGV_ROLL_NB_HEAVY 8   // workers per CTA of the <= 64 x 64 heavy shape
#endif
template <int NW, int RPL>
struct RollShape {
    static constexpr bool two = NW == 1 && RPL <= 2;
    static constexpr int nb_light = 3, cps_light = NW * RPL > 4 || (NW == 1 && RPL <= 2) ? 2 : 0;   // cps 0: by occupancy
    static constexpr int nb_heavy = two ? AGV_ROLL_NB_HEAVY : 3, cps_heavy = 0;
    static constexpr int heavy_pct = 15;   // searches per 100 turns from which the heavy shape runs
};

template <int NW, int RPL, int OBS, int MODE>
static int launch_shape(const DevCfg &dc, const RollArgs &ra_in, unsigned *order, cudaStream_t st) {
    using S = RollShape<NW, RPL>;
    RollArgs ra = ra_in;
    // searches per tick from which a scene runs on the heavy warp / below which it returns (agv_rollout.cuh):
    // 32 / 24 for maps up to 64 x 64, never on larger maps (c5: 3.3e8 AGV-steps/s without heavy scenes, 2.7e8 with
    // 24 / 16, 2.3e8 with 16 / 8); AGV_ROLL_HEAVY_BFS / AGV_ROLL_LIGHT_BFS override (tuning)
    ra.heavy_bfs = env_int("AGV_ROLL_HEAVY_BFS", S::two ? AGV_ROLL_HEAVY_BFS : (1 << 20));
    ra.light_bfs = env_int("AGV_ROLL_LIGHT_BFS", S::two ? AGV_ROLL_LIGHT_BFS : (1 << 20));
#ifdef AGV_ROLL_NB
    return launch_rollout<NW, RPL, OBS, AGV_ROLL_SPW, AGV_ROLL_NB, MODE>(dc, ra, order, st);
#else
    if (S::two) {
        static bool heavy = false;  // (kept while the figure is unknown: right after agv_set_state / agv_reset)
        if (ra.stat && reinterpret_cast<volatile unsigned long long *>(ra.stat)[1] > 0) {
            const unsigned long long n_bfs = reinterpret_cast<volatile unsigned long long *>(ra.stat)[0],
                                     n_turns = reinterpret_cast<volatile unsigned long long *>(ra.stat)[1];
            heavy = n_bfs * 100ull >= n_turns * (unsigned long long)S::heavy_pct;
            static const bool dump = getenv("AGV_PROF_DUMP") != nullptr;
            if (dump) fprintf(stderr, "agv rollout load: %llu searches / %llu turns in the scenes' last ticks -> %s\n", n_bfs, n_turns, heavy ? "heavy" : "light");
        }
        const int shape_env = env_int("AGV_ROLLOUT_SHAPE", 0);
        if (shape_env ? shape_env == 2 : heavy) {
            // congested launches: the searches bind, quick round trips for the scenes with many of them pay
            ra.heavy_bfs = env_int("AGV_ROLL_HEAVY_BFS", 16); ra.light_bfs = env_int("AGV_ROLL_LIGHT_BFS", 8);
            return launch_rollout<NW, RPL, OBS, AGV_ROLL_SPW, S::nb_heavy, MODE>(dc, ra, order, st, S::cps_heavy);
        }
    }
    return launch_rollout<NW, RPL, OBS, AGV_ROLL_SPW, S::nb_light, MODE>(dc, ra, order, st, S::cps_light);
#endif
}

template <int NW, int RPL, int OBS>
static int dispatch_mode(const DevCfg &dc, const RollArgs &ra, unsigned *order, cudaStream_t st) {
    const LaneArgs &t = ra.la;
#ifndef AGV_ROLL_GENERIC_ONLY
    if (t.proto == PROTO_INTELLIGENT && t.policy == POLICY_GUIDED) return launch_shape<NW, RPL, OBS, 1>(dc, ra, order, st);
    if (t.proto == PROTO_ASTAR && t.policy == POLICY_ASTAR) return launch_shape<NW, RPL, OBS, 2>(dc, ra, order, st);
    if (t.proto == PROTO_INTELLIGENT && t.policy == POLICY_GIVEN) return launch_shape<NW, RPL, OBS, 3>(dc, ra, order, st);
#endif
    return launch_shape<NW, RPL, OBS, 0>(dc, ra, order, st);
}

template <int NW, int RPL>
static int dispatch_rollout(const DevCfg &dc, const RollArgs &ra, unsigned *order, cudaStream_t st) {
    if (ra.la.obs_kind == OBS_CLIP) return dispatch_mode<NW, RPL, OBS_CLIP>(dc, ra, order, st);
    return dispatch_mode<NW, RPL, OBS_NONE>(dc, ra, order, st);
}

// covers no-observation and 3x7x7 limited-visual rollouts without the shaped reward; everything else
// returns AGV_E_UNSUPPORTED and the caller falls back to n_ticks launches of agv_tick
int lanes_rollout(const DevCfg &dc, int NW, int RPL, const RollArgs &ra, unsigned *order, cudaStream_t st) {
    if (dc.shaped || !(ra.la.obs_kind == OBS_NONE || (ra.la.obs_kind == OBS_CLIP && dc.K == 7))) return -4;
#ifdef AGV_ROLL_ONLY12   // tuning builds: only the 64 x 64 shape (compile time)
    if (NW == 1 && RPL == 2) return dispatch_rollout<1, 2>(dc, ra, order, st);
    return -4;
#endif
    if (NW == 1 && RPL == 1) return dispatch_rollout<1, 1>(dc, ra, order, st);
    if (NW == 1 && RPL == 2) return dispatch_rollout<1, 2>(dc, ra, order, st);
    if (NW == 1 && RPL == 4) return dispatch_rollout<1, 4>(dc, ra, order, st);
    if (NW == 2 && RPL == 1) return dispatch_rollout<2, 1>(dc, ra, order, st);
    if (NW == 2 && RPL == 2) return dispatch_rollout<2, 2>(dc, ra, order, st);
    return dispatch_rollout<2, 4>(dc, ra, order, st);
}

}  // namespace agv
