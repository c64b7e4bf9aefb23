// agv_lanes.cu -- kernels + launcher of the lane-per-scene fused tick (agv_lanes.cuh).
// sm_100a only.  Called from agv_tick (agv_kernels.cu).
#include "agv_lanes.cuh"

#include <cstdlib>

namespace agv {
void note_launch();
void note_cuda_error(int e);

template <int NW, int RPL, int OBS, bool SHAPED, int SPW>
__global__ void __launch_bounds__(32) k_tick_lanes(const DevCfg c, const LaneArgs t) {
    extern __shared__ __align__(16) uint8_t lane_smem[];
    tick_lanes<NW, RPL, OBS, SHAPED, SPW>(c, t, lane_smem);
}

static inline int align8(int x) { return (x + 7) & ~7; }

template <int NW, int RPL, int OBS, bool SHAPED, int SPW>
static int launch_lanes(const DevCfg &dc, LaneArgs t, cudaStream_t st) {
    const int HN = dc.H * NW;
    int off = 0;
    t.off_tab = off; off += 4 * HN * 8;
    t.blk_stride = dc.env_stride + 4;  // odd number of words: same-field accesses of the 32 lanes hit 32 banks
    t.off_blk = off; off = align8(off + SPW * t.blk_stride);
    t.occ_stride = HN | 1;
    t.off_occ = off; off += SPW * t.occ_stride * 8;
    t.off_occm = off; off += SPW * t.occ_stride * 8;
    t.off_vrow = off; off += HN * 8;
    t.off_zero = off; off += HN * 8;
    auto kern = k_tick_lanes<NW, RPL, OBS, SHAPED, SPW>;
    if (off > 227 * 1024) return -4;  // AGV_E_UNSUPPORTED
    if (off > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, off);
        if (e != cudaSuccess) { note_cuda_error((int)e); return -2; }
    }
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    const int ctas = (dc.E + SPW - 1) / SPW;
    kern<<<ctas, 32, off, st>>>(dc, t);
    note_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { note_cuda_error((int)e); return -2; }
    return 0;
}

template <int NW, int RPL, int SPW>
static int dispatch_obs(const DevCfg &dc, const LaneArgs &t, cudaStream_t st) {
    const bool sh = dc.shaped != 0;
    if (t.obs_kind == OBS_CLIP) return sh ? launch_lanes<NW, RPL, OBS_CLIP, true, SPW>(dc, t, st) : launch_lanes<NW, RPL, OBS_CLIP, false, SPW>(dc, t, st);
    if (t.obs_kind == OBS_FULL) return sh ? launch_lanes<NW, RPL, OBS_FULL, true, SPW>(dc, t, st) : launch_lanes<NW, RPL, OBS_FULL, false, SPW>(dc, t, st);
    return sh ? launch_lanes<NW, RPL, OBS_NONE, true, SPW>(dc, t, st) : launch_lanes<NW, RPL, OBS_NONE, false, SPW>(dc, t, st);
}

// scenes per warp (SPW): fewer scenes per warp = more warps to hide the latency of the scalar
// chains, at the price of executing them once per warp; 8 measured best at 64 x 64 / 16384 scenes
// (AGV_LANES_SPW = 8|16|32 selects in tuning builds, -DAGV_LANES_TUNE)
int lanes_tick(const DevCfg &dc, int NW, int RPL, const LaneArgs &t, cudaStream_t st) {
    if (NW == 1 && RPL == 1) return dispatch_obs<1, 1, 32>(dc, t, st);
    if (NW == 1 && RPL == 2) {
#ifdef AGV_LANES_TUNE
        const char *e = getenv("AGV_LANES_SPW");
        const int spw = e ? atoi(e) : 8;
        if (spw == 16) return dispatch_obs<1, 2, 16>(dc, t, st);
        if (spw == 32) return dispatch_obs<1, 2, 32>(dc, t, st);
#endif
        return dispatch_obs<1, 2, 8>(dc, t, st);
    }
    if (NW == 1 && RPL == 4) return dispatch_obs<1, 4, 32>(dc, t, st);
    if (NW == 2 && RPL == 1) return dispatch_obs<2, 1, 32>(dc, t, st);
    if (NW == 2 && RPL == 2) return dispatch_obs<2, 2, 32>(dc, t, st);
    return dispatch_obs<2, 4, 32>(dc, t, st);
}

}  // namespace agv
