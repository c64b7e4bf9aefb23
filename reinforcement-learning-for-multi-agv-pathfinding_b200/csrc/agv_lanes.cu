// agv_lanes.cu -- kernels + launcher of the lane-per-scene fused tick (agv_lanes.cuh).
// sm_100a only.  Called from agv_tick (agv_kernels.cu).
#include "agv_async.cuh"

#include <cstdlib>

namespace agv {
void note_launch();
void note_cuda_error(int e);

// warp 0 carries the SPW scenes; warps 1..NWK (if any) serve its BFS / record-expansion requests
template <int NW, int RPL, int OBS, bool SHAPED, int SPW, int NWK>
__global__ void __launch_bounds__(32 * (1 + NWK)) k_tick_lanes(const DevCfg c, const LaneArgs t) {
    extern __shared__ __align__(16) uint8_t lane_smem[];
    // the scene warp's slot rotates with the CTA's wave position so that co-resident CTAs put their
    // (latency-bound) scene warps on different SM sub-partitions (warp slot % 4)
    const int role = NWK > 0 ? (int)((blockIdx.x / 148u) % (unsigned)(1 + NWK)) : 0;
    const int warp = threadIdx.x >> 5;
    if (NWK > 0 && warp != role) lane_worker<NW, RPL, OBS, SPW, NWK>(c, t, lane_smem, warp < role ? warp : warp - 1);
    else tick_lanes<NW, RPL, OBS, SHAPED, SPW, NWK>(c, t, lane_smem);
}

// decoupled lanes: two scene warps (light / heavy scenes) + NB BFS workers (+ the record writer for OBS_CLIP)
// 4 CTAs per SM hold the 512 CTAs of 16384 scenes in one wave: the 64-column variants are compiled
// for that register budget (one register too many means 3 CTAs per SM and a second wave, +20 %)
template <int NW, int RPL>
constexpr int async_min_ctas() { return NW * RPL <= 2 ? 4 : (NW * RPL <= 4 ? 2 : 1); }

template <int NW, int RPL, int OBS, int SPW, int NB>
__global__ void __launch_bounds__(32 * (2 + NB + (OBS == OBS_CLIP ? 1 : 0)), async_min_ctas<NW, RPL>())
k_tick_async(const DevCfg c, const LaneArgs t) {
    extern __shared__ __align__(16) uint8_t lane_smem[];
    async_cta<NW, RPL, OBS, SPW, NB>(c, t, lane_smem);
}

static inline int align8(int x) { return (x + 7) & ~7; }

template <int NW, int SPW>
static int plan_lanes(const DevCfg &dc, LaneArgs &t, int n_vrow, int queue_bytes) {
    const int HN = dc.H * NW;
    int off = 0;
    t.off_tab = off; off += (mirror_copy<NW>() ? 4 : 2) * HN * 8;
    t.blk_stride = dc.env_stride + 4;  // odd number of words: same-field accesses of the 32 lanes hit 32 banks
    t.off_blk = off; off = align8(off + SPW * t.blk_stride);
    t.occ_stride = HN | 1;
    t.off_occ = off; off += SPW * t.occ_stride * 8;
    t.off_occm = off; off += mirror_copy<NW>() ? SPW * t.occ_stride * 8 : 0;
    t.off_vrow = off; off += n_vrow * HN * 8;
    t.off_zero = off; off += HN * 8;
    t.off_queue = (off + 15) & ~15; off = t.off_queue + queue_bytes;
    return off;
}

template <int NW, int RPL, int OBS, int SPW, int NB>
static int launch_async(const DevCfg &dc, LaneArgs t, cudaStream_t st) {
    const int smem = plan_lanes<NW, SPW>(dc, t, 1, (int)sizeof(AsyncShared));
    auto kern = k_tick_async<NW, RPL, OBS, SPW, NB>;
    if (smem > 227 * 1024) return -4;  // AGV_E_UNSUPPORTED
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) { note_cuda_error((int)e); return -2; }
    }
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    const int ctas = (dc.E + SPW - 1) / SPW;
    kern<<<ctas, 32 * (2 + NB + (OBS == OBS_CLIP ? 1 : 0)), smem, st>>>(dc, t);
    note_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { note_cuda_error((int)e); return -2; }
    return 0;
}

template <int NW, int RPL, int SPW, int NB>
static int dispatch_async(const DevCfg &dc, const LaneArgs &t, cudaStream_t st) {
    if (t.obs_kind == OBS_CLIP) return launch_async<NW, RPL, OBS_CLIP, SPW, NB>(dc, t, st);
    return launch_async<NW, RPL, OBS_NONE, SPW, NB>(dc, t, st);
}

template <int NW, int RPL, int OBS, bool SHAPED, int SPW, int NWK>
static int launch_lanes(const DevCfg &dc, LaneArgs t, cudaStream_t st) {
    const int off = plan_lanes<NW, SPW>(dc, t, 1 + NWK, (int)sizeof(LaneQueue));
    auto kern = k_tick_lanes<NW, RPL, OBS, SHAPED, SPW, NWK>;
    if (off > 227 * 1024) return -4;  // AGV_E_UNSUPPORTED
    if (off > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, off);
        if (e != cudaSuccess) { note_cuda_error((int)e); return -2; }
    }
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    const int ctas = (dc.E + SPW - 1) / SPW;
    kern<<<ctas, 32 * (1 + NWK), off, st>>>(dc, t);
    note_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { note_cuda_error((int)e); return -2; }
    return 0;
}

template <int NW, int RPL, int SPW, int NWK>
static int dispatch_obs(const DevCfg &dc, const LaneArgs &t, cudaStream_t st) {
    const bool sh = dc.shaped != 0;
    if (t.obs_kind == OBS_CLIP) return sh ? launch_lanes<NW, RPL, OBS_CLIP, true, SPW, NWK>(dc, t, st) : launch_lanes<NW, RPL, OBS_CLIP, false, SPW, NWK>(dc, t, st);
    if (t.obs_kind == OBS_FULL) return sh ? launch_lanes<NW, RPL, OBS_FULL, true, SPW, NWK>(dc, t, st) : launch_lanes<NW, RPL, OBS_FULL, false, SPW, NWK>(dc, t, st);
    return sh ? launch_lanes<NW, RPL, OBS_NONE, true, SPW, NWK>(dc, t, st) : launch_lanes<NW, RPL, OBS_NONE, false, SPW, NWK>(dc, t, st);
}

// scenes per warp (SPW) and worker warps per CTA (NWK); AGV_LANES_SPW / AGV_LANES_NWK select among
// the variants compiled into tuning builds (-DAGV_LANES_TUNE)
int lanes_tick(const DevCfg &dc, int NW, int RPL, const LaneArgs &t, cudaStream_t st, int allow_async) {
    // decoupled lanes cover no-observation and 3x7x7 limited-visual ticks without the shaped reward
    if (allow_async && !dc.shaped && (t.obs_kind == OBS_NONE || (t.obs_kind == OBS_CLIP && dc.K == 7))) {
        if (NW == 1 && RPL == 1) return dispatch_async<1, 1, 32, 3>(dc, t, st);
        if (NW == 1 && RPL == 2) {
#ifdef AGV_LANES_TUNE
            const char *e = getenv("AGV_LANES_SPW"), *w = getenv("AGV_LANES_NWK");
            const int spw = e ? atoi(e) : 32, nb = w ? atoi(w) : 3;
            if (spw == 32 && nb == 1) return dispatch_async<1, 2, 32, 1>(dc, t, st);
            if (spw == 32 && nb == 2) return dispatch_async<1, 2, 32, 2>(dc, t, st);
            if (spw == 16 && nb == 1) return dispatch_async<1, 2, 16, 1>(dc, t, st);
            if (spw == 16 && nb == 2) return dispatch_async<1, 2, 16, 2>(dc, t, st);
            if (spw == 16 && nb == 3) return dispatch_async<1, 2, 16, 3>(dc, t, st);
            if (spw == 32 && nb == 4) return dispatch_async<1, 2, 32, 4>(dc, t, st);
            if (spw == 8 && nb == 1) return dispatch_async<1, 2, 8, 1>(dc, t, st);
#endif
            return dispatch_async<1, 2, 32, 3>(dc, t, st);  // measured best at 64 x 64, 16384 scenes
        }
        if (NW == 1 && RPL == 4) return dispatch_async<1, 4, 32, 3>(dc, t, st);
        if (NW == 2 && RPL == 1) return dispatch_async<2, 1, 32, 3>(dc, t, st);
        if (NW == 2 && RPL == 2) return dispatch_async<2, 2, 32, 3>(dc, t, st);
        return dispatch_async<2, 4, 32, 3>(dc, t, st);
    }
    if (NW == 1 && RPL == 1) return dispatch_obs<1, 1, 32, 0>(dc, t, st);
    if (NW == 1 && RPL == 2) {
#ifdef AGV_LANES_TUNE
        const char *e = getenv("AGV_LANES_SPW"), *w = getenv("AGV_LANES_NWK");
        const int spw = e ? atoi(e) : 32, nwk = w ? atoi(w) : 3;
        if (spw == 8 && nwk == 0) return dispatch_obs<1, 2, 8, 0>(dc, t, st);
        if (spw == 32 && nwk == 2) return dispatch_obs<1, 2, 32, 2>(dc, t, st);
        if (spw == 16 && nwk == 3) return dispatch_obs<1, 2, 16, 3>(dc, t, st);
        if (spw == 16 && nwk == 2) return dispatch_obs<1, 2, 16, 2>(dc, t, st);
        if (spw == 16 && nwk == 1) return dispatch_obs<1, 2, 16, 1>(dc, t, st);
        if (spw == 8 && nwk == 1) return dispatch_obs<1, 2, 8, 1>(dc, t, st);
        if (spw == 16) return dispatch_obs<1, 2, 16, 0>(dc, t, st);
        if (spw == 32) return dispatch_obs<1, 2, 32, 0>(dc, t, st);
#endif
        return dispatch_obs<1, 2, 32, 3>(dc, t, st);  // shaped reward at 64 x 64: 1.25 ms vs 1.74 ms with <8, 0>
    }
    if (NW == 1 && RPL == 4) return dispatch_obs<1, 4, 32, 0>(dc, t, st);
    if (NW == 2 && RPL == 1) return dispatch_obs<2, 1, 32, 0>(dc, t, st);
    if (NW == 2 && RPL == 2) return dispatch_obs<2, 2, 32, 0>(dc, t, st);
    return dispatch_obs<2, 4, 32, 0>(dc, t, st);
}

}  // namespace agv
