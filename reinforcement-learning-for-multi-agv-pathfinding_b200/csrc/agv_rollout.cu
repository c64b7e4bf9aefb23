// agv_rollout.cu -- kernels + launcher of the persistent multi-tick rollout (agv_rollout.cuh).
// sm_100a only.  Called from agv_rollout / agv_tick (agv_kernels.cu).
#include "agv_rollout.cuh"

#include <cstdlib>

namespace agv {
void note_launch();
void note_cuda_error(int e);

template <int NW, int RPL>
constexpr int rollout_min_ctas() { return NW * RPL <= 2 ? 4 : (NW * RPL <= 4 ? 2 : 1); }

template <int NW, int RPL, int OBS, int SPW, int NB, int MODE>
__global__ void __launch_bounds__(32 * (3 + NB), rollout_min_ctas<NW, RPL>())
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

__global__ void __launch_bounds__(1024) k_rollout_order(const DevCfg c, unsigned *order, unsigned *qctr, unsigned q0) {
    __shared__ unsigned hist[64], base[64];
    if (threadIdx.x < 64) hist[threadIdx.x] = 0u;
    __syncthreads();
    for (int e = threadIdx.x; e < c.E; e += blockDim.x) {
        const unsigned w = *reinterpret_cast<const unsigned *>(c.state + (size_t)e * c.env_stride + 24);
        atomicAdd(&hist[order_bucket(w)], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned s = 0;
        for (int b = 0; b < 64; b++) { base[b] = s; s += hist[b]; }
        *qctr = q0;
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
static int launch_rollout(const DevCfg &dc, RollArgs ra, unsigned *order, cudaStream_t st) {
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
    int cps = env_int("AGV_ROLLOUT_CPS", 0);
    if (cps < 1 || cps > cps_cached) cps = cps_cached >= 3 ? cps_cached - 1 : cps_cached;
    int ctas = (dc.E + SPW - 1) / SPW;
    if (ctas > sms * cps) ctas = sms * cps;
    k_rollout_order<<<1, 1024, 0, st>>>(dc, order, ra.qctr, (unsigned)ctas * SPW);
    note_launch();
    ra.order = order;
    kern<<<ctas, threads, smem, st>>>(dc, ra);
    note_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { note_cuda_error((int)e); return -2; }
    return 0;
}

// protocol / policy specialisations (agv_rollout.cuh, MODE): the two modes the rollout exists for get their
// own kernels; open-loop actions and mixed settings run on the generic one
// BFS worker warps per CTA.  64 x 64 (c3, 16 ticks): 2 -> 1.10e9, 3 -> 1.06e9, 4 -> 0.99e9 AGV-steps/s (the scene warps'
// turn chains bind, more workers only compete for issue slots).  128 x 128 (c5, 8 ticks): searches are 4x longer
// (4 rows x 2 words per lane, paths up to ~250 levels) and bind: 2 -> 2.08e8, 3 -> 2.31e8, 4 -> 2.44e8 (A_star
// control mode 1.42e8 -> 1.89e8); 5 would not leave registers for 2 CTAs per SM.  AGV_ROLL_NB overrides (tuning).
template <int NW, int RPL>
constexpr int rollout_nb() {
#ifdef AGV_ROLL_NB
    return AGV_ROLL_NB;
#else
    return NW * RPL <= 2 ? 2 : (NW * RPL <= 4 ? 3 : 4);
#endif
}

template <int NW, int RPL, int OBS>
static int dispatch_mode(const DevCfg &dc, const RollArgs &ra, unsigned *order, cudaStream_t st) {
    const LaneArgs &t = ra.la;
#ifndef AGV_ROLL_GENERIC_ONLY
    if (t.proto == PROTO_INTELLIGENT && t.policy == POLICY_GUIDED) return launch_rollout<NW, RPL, OBS, 32, rollout_nb<NW, RPL>(), 1>(dc, ra, order, st);
    if (t.proto == PROTO_ASTAR && t.policy == POLICY_ASTAR) return launch_rollout<NW, RPL, OBS, 32, rollout_nb<NW, RPL>(), 2>(dc, ra, order, st);
    if (t.proto == PROTO_INTELLIGENT && t.policy == POLICY_GIVEN) return launch_rollout<NW, RPL, OBS, 32, rollout_nb<NW, RPL>(), 3>(dc, ra, order, st);
#endif
    return launch_rollout<NW, RPL, OBS, 32, rollout_nb<NW, RPL>(), 0>(dc, ra, order, st);
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
