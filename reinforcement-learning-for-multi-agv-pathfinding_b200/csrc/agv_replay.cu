// agv_replay.cu -- device-resident replay buffer with a PER sum-tree
// (reference: src/algorithm/MADQN_structure/PER.py:6-112, MADQN.py:113-171).
//
// Storage (SoA, HBM): obs bf16[cap, R], next_obs bf16[cap, R], action int8[cap],
// reward f32[cap], done u8[cap], tree f64[2*cap-1] in the reference's heap layout
// (leaf of data slot s is node s + cap - 1, PER.py:42).
// Push compacts the rows that took a turn (act >= 0) in row order and appends them
// at the ring cursor (SumTree.add, PER.py:41-52).  Sampling is the reference's
// stratified prefix-sum descent (PER.py:25-35,86-108) with injected uniforms, fused
// with the gather of the batch.
#include <cuda_runtime.h>
#include <stdint.h>

#include <new>

#include "../../include/agv_b200.h"

namespace agv {
void note_launch();
void note_cuda_error(int e);

constexpr int PUSH_ROWS = 256;        // rows per CTA in the push kernels
constexpr int EXACT_TREE_LIMIT = 2048; // pushes up to this many rows replay the reference's
                                      // sequential `tree[parent] += change` order bit-exactly

struct ReplayDev {
    long long cap;
    int R;
    uint16_t *obs, *next_obs;
    int8_t *act;
    float *rew;
    uint8_t *done;
    double *tree;
    long long *meta;      // [0] write cursor, [1] n_entries, [2] base cursor of the push in flight,
                          // [3] rows accepted by the push in flight, [4] rows skipped (overflowing cap)
    long long *blk_count; // per-CTA valid-row counts / exclusive offsets
};

__global__ void k_push_count(long long n_rows, const int8_t *act, long long *blk_count) {
    __shared__ int warp_cnt[PUSH_ROWS / 32];
    const long long row = (long long)blockIdx.x * PUSH_ROWS + threadIdx.x;
    const bool v = row < n_rows && act[row] >= 0;
    const unsigned b = __ballot_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = __popc(b);
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int i = 0; i < PUSH_ROWS / 32; i++) s += warp_cnt[i];
        blk_count[blockIdx.x] = s;
    }
}

// single CTA: exclusive scan of the per-CTA counts + ring bookkeeping
__global__ void k_push_scan(int n_blocks, ReplayDev r) {
    __shared__ long long carry;
    __shared__ long long wsum[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n_blocks; base += blockDim.x) {
        const int i = base + threadIdx.x;
        long long v = i < n_blocks ? r.blk_count[i] : 0;
        // inclusive warp scan
        long long x = v;
        for (int o = 1; o < 32; o <<= 1) {
            long long y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            long long w = threadIdx.x < (blockDim.x >> 5) ? wsum[threadIdx.x] : 0;
            long long xs = w;
            for (int o = 1; o < 32; o <<= 1) {
                long long y = __shfl_up_sync(0xffffffffu, xs, o);
                if (threadIdx.x >= o) xs += y;
            }
            wsum[threadIdx.x] = xs - w;  // exclusive
        }
        __syncthreads();
        const long long excl = carry + wsum[threadIdx.x >> 5] + (x - v);
        if (i < n_blocks) r.blk_count[i] = excl;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const long long total = carry;
        const long long skip = total > r.cap ? total - r.cap : 0;  // older rows of this push would be overwritten
        const long long w0 = r.meta[0];
        r.meta[2] = w0;
        r.meta[3] = total;
        r.meta[4] = skip;
        r.meta[0] = (w0 + total) % r.cap;
        long long ne = r.meta[1] + total;
        r.meta[1] = ne > r.cap ? r.cap : ne;
    }
}

__global__ void k_push_scatter(long long n_rows, ReplayDev r, const int8_t *act, const float *rew,
                               const uint8_t *done, const uint16_t *obs, const uint16_t *next_obs,
                               const double *prio, double *leaf_new, long long *leaf_slot) {
    __shared__ int warp_off[PUSH_ROWS / 32];
    __shared__ long long slot_s[PUSH_ROWS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long row = (long long)blockIdx.x * PUSH_ROWS + threadIdx.x;
    const bool v = row < n_rows && act[row] >= 0;
    const unsigned b = __ballot_sync(0xffffffffu, v);
    if (lane == 0) warp_off[w] = __popc(b);
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int i = 0; i < PUSH_ROWS / 32; i++) { int t = warp_off[i]; warp_off[i] = s; s += t; }
    }
    __syncthreads();
    long long slot = -1;
    if (v) {
        const long long rank = r.blk_count[blockIdx.x] + warp_off[w] + __popc(b & ((1u << lane) - 1));
        if (rank >= r.meta[4]) {
            slot = (r.meta[2] + rank) % r.cap;
            r.act[slot] = act[row];
            r.rew[slot] = rew ? rew[row] : 0.f;
            r.done[slot] = done ? done[row] : 0;
            // insertion-ordered list of (slot, priority) for the tree update
            leaf_slot[rank] = slot;
            leaf_new[rank] = prio ? prio[row] : 1.0;
        } else {
            leaf_slot[rank] = -1;
        }
    }
    slot_s[threadIdx.x] = slot;
    __syncthreads();
    // cooperative record copy: one warp per row
    for (int q = w; q < PUSH_ROWS; q += PUSH_ROWS / 32) {
        const long long s = slot_s[q];
        if (s < 0) continue;
        const long long src_row = (long long)blockIdx.x * PUSH_ROWS + q;
        if (obs) for (int i = lane; i < r.R; i += 32) r.obs[s * r.R + i] = obs[src_row * r.R + i];
        if (next_obs) for (int i = lane; i < r.R; i += 32) r.next_obs[s * r.R + i] = next_obs[src_row * r.R + i];
    }
}

// SumTree.update for every pushed row in insertion order (PER.py:16-22,55-59):
// change = p - tree[leaf]; tree[leaf] = p; every ancestor += change.  One thread, so
// the float64 additions happen in exactly the reference's order.
__global__ void k_tree_update_exact(ReplayDev r, const double *leaf_new, const long long *leaf_slot) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const long long total = r.meta[3];
    for (long long t = 0; t < total; t++) {
        const long long s = leaf_slot[t];
        if (s < 0) continue;
        long long idx = s + r.cap - 1;
        const double change = leaf_new[t] - r.tree[idx];
        r.tree[idx] = leaf_new[t];
        while (idx != 0) {
            idx = (idx - 1) / 2;
            r.tree[idx] += change;
        }
    }
}

// large pushes: write the leaves, then rebuild the internal nodes level by level as
// tree[v] = tree[2v+1] + tree[2v+2] (same sums up to float64 rounding order).
__global__ void k_tree_set_leaves(ReplayDev r, const double *leaf_new, const long long *leaf_slot) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= r.meta[3]) return;
    const long long s = leaf_slot[t];
    if (s >= 0) r.tree[s + r.cap - 1] = leaf_new[t];
}
// Bottom-up rebuild of all internal nodes in two launches.  k_tree_rebuild_sub: CTA b owns the
// subtree rooted at heap node (2^d0 - 1) + b and recomputes its levels deepest-first (a node is
// internal iff its left child exists: the reference's `left >= len(tree)` test, PER.py:29);
// k_tree_rebuild_top: one CTA finishes the levels above d0.
__global__ void k_tree_rebuild_sub(ReplayDev r, int d0, int lmax) {
    const long long n = 2 * r.cap - 1;
    const long long root = ((1ll << d0) - 1) + blockIdx.x;
    if (root >= n) return;
    for (int d = lmax; d >= d0; d--) {
        const long long first = ((root + 1) << (d - d0)) - 1, count = 1ll << (d - d0);
        for (long long q = threadIdx.x; q < count; q += blockDim.x) {
            const long long v = first + q, l = 2 * v + 1;
            if (v < n && l < n) r.tree[v] = r.tree[l] + (l + 1 < n ? r.tree[l + 1] : 0.0);
        }
        __syncthreads();
    }
}
__global__ void k_tree_rebuild_top(ReplayDev r, int d0) {
    const long long n = 2 * r.cap - 1;
    for (int d = d0 - 1; d >= 0; d--) {
        const long long first = (1ll << d) - 1, count = 1ll << d;
        for (long long q = threadIdx.x; q < count; q += blockDim.x) {
            const long long v = first + q, l = 2 * v + 1;
            if (v < n && l < n) r.tree[v] = r.tree[l] + (l + 1 < n ? r.tree[l + 1] : 0.0);
        }
        __syncthreads();
    }
}

__device__ __forceinline__ void gather_row(const ReplayDev &r, long long s, int q, uint16_t *obs, int8_t *act,
                                           float *rew, uint16_t *next_obs, uint8_t *done, int lane) {
    if (obs) for (int i = lane; i < r.R; i += 32) obs[(long long)q * r.R + i] = r.obs[s * r.R + i];
    if (next_obs) for (int i = lane; i < r.R; i += 32) next_obs[(long long)q * r.R + i] = r.next_obs[s * r.R + i];
    if (lane == 0) {
        if (act) act[q] = r.act[s];
        if (rew) rew[q] = r.rew[s];
        if (done) done[q] = r.done[s];
    }
}

// Memory.sample (PER.py:86-108): one warp per sample; lane 0 walks the tree.
__global__ void k_sample(ReplayDev r, int n, const double *u01, long long *slot_out, double *prio_out, uint16_t *obs,
                         int8_t *act, float *rew, uint16_t *next_obs, uint8_t *done) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= n) return;
    long long slot = 0;
    if (lane == 0) {
        const double segment = r.tree[0] / (double)n;
        const double a = segment * (double)q, b = segment * (double)(q + 1);
        double s = a + (b - a) * u01[q];  // random.uniform(a, b)
        const long long nn = 2 * r.cap - 1;
        long long idx = 0;
        while (true) {
            const long long left = 2 * idx + 1;
            if (left >= nn) break;
            const double tl = r.tree[left];
            if (s <= tl) idx = left;
            else { s -= tl; idx = left + 1; }
        }
        slot = idx - r.cap + 1;
        if (slot_out) slot_out[q] = slot;
        if (prio_out) prio_out[q] = r.tree[idx];
    }
    slot = __shfl_sync(0xffffffffu, slot, 0);
    gather_row(r, slot, q, obs, act, rew, next_obs, done, lane);
}

// SumTree.update (PER.py:55-59) for a list of tree indices, sequentially (reference order)
__global__ void k_tree_update_list(ReplayDev r, int n, const long long *tree_idx, const double *prio) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const long long nn = 2 * r.cap - 1;
    for (int t = 0; t < n; t++) {
        long long idx = tree_idx[t];
        if (idx < 0 || idx >= nn) continue;
        const double change = prio[t] - r.tree[idx];
        r.tree[idx] = prio[t];
        while (idx != 0) {
            idx = (idx - 1) / 2;
            r.tree[idx] += change;
        }
    }
}

__global__ void k_gather(ReplayDev r, int n, const long long *slots, uint16_t *obs, int8_t *act, float *rew,
                         uint16_t *next_obs, uint8_t *done) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= n) return;
    long long s = slots[q];
    if (s < 0 || s >= r.cap) s = 0;
    gather_row(r, s, q, obs, act, rew, next_obs, done, lane);
}

}  // namespace agv

using namespace agv;

struct AgvReplay {
    int device;
    ReplayDev d;
    double *leaf_new = nullptr;
    long long *leaf_slot = nullptr;
    long long scratch_rows = 0;
    long long scratch_blocks = 0;
};

#define R_TRY(expr)                                                  \
    do {                                                             \
        cudaError_t _e = (expr);                                     \
        if (_e != cudaSuccess) { note_cuda_error((int)_e); return AGV_E_CUDA; } \
    } while (0)
#define R_LAUNCH_CHECK()                                             \
    do {                                                             \
        note_launch();                                               \
        cudaError_t _e = cudaGetLastError();                         \
        if (_e != cudaSuccess) { note_cuda_error((int)_e); return AGV_E_CUDA; } \
    } while (0)

extern "C" {

int agv_replay_create(int32_t device, int64_t capacity, int32_t obs_elems, AgvReplay **out) {
    if (!out || capacity < 1 || obs_elems < 0) return AGV_E_INVALID;
    R_TRY(cudaSetDevice(device));
    AgvReplay *r = new (std::nothrow) AgvReplay();
    if (!r) return AGV_E_NOMEM;
    r->device = device;
    ReplayDev &d = r->d;
    d.cap = capacity; d.R = obs_elems;
    const size_t ob = (size_t)capacity * (obs_elems > 0 ? obs_elems : 1) * 2;
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = cudaMalloc(&d.obs, ob);
    if (e == cudaSuccess) e = cudaMalloc(&d.next_obs, ob);
    if (e == cudaSuccess) e = cudaMalloc(&d.act, capacity);
    if (e == cudaSuccess) e = cudaMalloc(&d.rew, capacity * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d.done, capacity);
    if (e == cudaSuccess) e = cudaMalloc(&d.tree, (2 * capacity - 1) * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d.meta, 8 * sizeof(long long));
    if (e == cudaSuccess) e = cudaMemset(d.tree, 0, (2 * capacity - 1) * sizeof(double));
    if (e == cudaSuccess) e = cudaMemset(d.meta, 0, 8 * sizeof(long long));
    if (e == cudaSuccess) e = cudaMemset(d.obs, 0, ob);
    if (e == cudaSuccess) e = cudaMemset(d.next_obs, 0, ob);
    if (e == cudaSuccess) e = cudaMemset(d.act, 0, capacity);
    if (e == cudaSuccess) e = cudaMemset(d.rew, 0, capacity * sizeof(float));
    if (e == cudaSuccess) e = cudaMemset(d.done, 0, capacity);
    if (e != cudaSuccess) { note_cuda_error((int)e); agv_replay_destroy(r); return AGV_E_CUDA; }
    *out = r;
    return AGV_OK;
}

int agv_replay_destroy(AgvReplay *r) {
    if (!r) return AGV_OK;
    cudaSetDevice(r->device);
    cudaFree(r->d.obs); cudaFree(r->d.next_obs); cudaFree(r->d.act); cudaFree(r->d.rew); cudaFree(r->d.done);
    cudaFree(r->d.tree); cudaFree(r->d.meta); cudaFree(r->d.blk_count); cudaFree(r->leaf_new);
    cudaFree(r->leaf_slot);
    delete r;
    return AGV_OK;
}

int agv_replay_push(AgvReplay *r, int64_t n_rows, const int8_t *act_dev, const float *reward_dev,
                    const uint8_t *done_dev, const void *obs_dev, const void *next_obs_dev,
                    const double *priority_dev, void *stream) {
    if (!r || n_rows < 0 || !act_dev) return AGV_E_INVALID;
    if (n_rows == 0) return AGV_OK;
    R_TRY(cudaSetDevice(r->device));
    cudaStream_t st = (cudaStream_t)stream;
    const long long nb = (n_rows + PUSH_ROWS - 1) / PUSH_ROWS;
    if (n_rows > r->scratch_rows || nb > r->scratch_blocks) {
        R_TRY(cudaStreamSynchronize(st));
        cudaFree(r->leaf_new); cudaFree(r->leaf_slot); cudaFree(r->d.blk_count);
        r->leaf_new = nullptr; r->leaf_slot = nullptr; r->d.blk_count = nullptr;
        R_TRY(cudaMalloc(&r->leaf_new, n_rows * sizeof(double)));
        R_TRY(cudaMalloc(&r->leaf_slot, n_rows * sizeof(long long)));
        R_TRY(cudaMalloc(&r->d.blk_count, nb * sizeof(long long)));
        r->scratch_rows = n_rows; r->scratch_blocks = nb;
    }
    k_push_count<<<(unsigned)nb, PUSH_ROWS, 0, st>>>(n_rows, act_dev, r->d.blk_count);
    R_LAUNCH_CHECK();
    k_push_scan<<<1, 1024, 0, st>>>((int)nb, r->d);
    R_LAUNCH_CHECK();
    k_push_scatter<<<(unsigned)nb, PUSH_ROWS, 0, st>>>(n_rows, r->d, act_dev, reward_dev, done_dev,
                                                       (const uint16_t *)obs_dev, (const uint16_t *)next_obs_dev,
                                                       priority_dev, r->leaf_new, r->leaf_slot);
    R_LAUNCH_CHECK();
    if (n_rows <= EXACT_TREE_LIMIT) {
        k_tree_update_exact<<<1, 32, 0, st>>>(r->d, r->leaf_new, r->leaf_slot);
        R_LAUNCH_CHECK();
    } else {
        k_tree_set_leaves<<<(unsigned)((n_rows + 255) / 256), 256, 0, st>>>(r->d, r->leaf_new, r->leaf_slot);
        R_LAUNCH_CHECK();
        // heap levels: level L holds nodes [2^L - 1, 2^(L+1) - 1)
        const long long n = 2 * r->d.cap - 1;
        int lmax = 0;
        while (((1ll << (lmax + 1)) - 1) < n) lmax++;
        const int d0 = lmax > 9 ? lmax - 9 : 0;
        k_tree_rebuild_sub<<<(unsigned)(1ll << d0), 1024, 0, st>>>(r->d, d0, lmax);
        R_LAUNCH_CHECK();
        if (d0 > 0) {
            k_tree_rebuild_top<<<1, 1024, 0, st>>>(r->d, d0);
            R_LAUNCH_CHECK();
        }
    }
    return AGV_OK;
}

int agv_replay_sample(AgvReplay *r, int32_t n, const double *u01_dev, int64_t *slot_dev, double *prio_dev,
                      void *obs_dev, int8_t *act_dev, float *reward_dev, void *next_obs_dev, uint8_t *done_dev,
                      void *stream) {
    if (!r || n < 1 || !u01_dev) return AGV_E_INVALID;
    R_TRY(cudaSetDevice(r->device));
    k_sample<<<(n + 3) / 4, 128, 0, (cudaStream_t)stream>>>(r->d, n, u01_dev, (long long *)slot_dev, prio_dev,
                                                           (uint16_t *)obs_dev, act_dev, reward_dev,
                                                           (uint16_t *)next_obs_dev, done_dev);
    R_LAUNCH_CHECK();
    return AGV_OK;
}

int agv_replay_update(AgvReplay *r, int32_t n, const int64_t *tree_idx_dev, const double *priority_dev,
                      void *stream) {
    if (!r || n < 1 || !tree_idx_dev || !priority_dev) return AGV_E_INVALID;
    R_TRY(cudaSetDevice(r->device));
    k_tree_update_list<<<1, 32, 0, (cudaStream_t)stream>>>(r->d, n, (const long long *)tree_idx_dev, priority_dev);
    R_LAUNCH_CHECK();
    return AGV_OK;
}

int agv_replay_gather(AgvReplay *r, int32_t n, const int64_t *slot_dev, void *obs_dev, int8_t *act_dev,
                      float *reward_dev, void *next_obs_dev, uint8_t *done_dev, void *stream) {
    if (!r || n < 1 || !slot_dev) return AGV_E_INVALID;
    R_TRY(cudaSetDevice(r->device));
    k_gather<<<(n + 3) / 4, 128, 0, (cudaStream_t)stream>>>(r->d, n, (const long long *)slot_dev, (uint16_t *)obs_dev,
                                                           act_dev, reward_dev, (uint16_t *)next_obs_dev, done_dev);
    R_LAUNCH_CHECK();
    return AGV_OK;
}

int agv_replay_info(AgvReplay *r, int64_t out[2], double *total_out, void *stream) {
    if (!r) return AGV_E_INVALID;
    R_TRY(cudaSetDevice(r->device));
    cudaStream_t st = (cudaStream_t)stream;
    long long m[2];
    double tot = 0;
    R_TRY(cudaMemcpyAsync(m, r->d.meta, sizeof(m), cudaMemcpyDeviceToHost, st));
    R_TRY(cudaMemcpyAsync(&tot, r->d.tree, sizeof(double), cudaMemcpyDeviceToHost, st));
    R_TRY(cudaStreamSynchronize(st));
    if (out) { out[0] = m[1]; out[1] = m[0]; }
    if (total_out) *total_out = tot;
    return AGV_OK;
}

}  // extern "C"
