/*
 * agv_b200.h -- C ABI of the B200-native batched multi-AGV (RMFS) environment.
 *
 * This is the drop-in boundary for the reference's hot path
 * (stoneMan349/Reinforcement-learning-for-multi-AGV-pathfinding): env step +
 * state encode + BFS "A*" guidance + replay.  The reference is pure Python and has
 * no FFI of its own; each entry point below names the reference interface it
 * replaces (file:line relative to the reference checkout).  INTEGRATION.md shows
 * the ctypes binding a maintainer adds to src/multiAGVscene/Scene.py.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / C++ types.
 *  - every function returns an int status: 0 = AGV_OK, <0 = AGV_E_*; nothing throws.
 *    Illegal moves are data (reward -1, done 1), not errors (Explorer.py:151-187).
 *  - `*_dev` pointers are device memory owned by the caller (e.g. torch tensors);
 *    `*_host` pointers are host memory (pinned for the async copies to overlap).
 *  - all work is enqueued on the caller's `stream` (a cudaStream_t passed as void*;
 *    NULL = legacy default stream) and is asynchronous unless stated otherwise.
 *  - a handle belongs to one GPU and one stream at a time; it is not thread-safe.
 *  - positions are 1-based (x = column, y = row) like the reference (Explorer.py:27);
 *    actions: 0 UP, 1 RIGHT, 2 DOWN, 3 LEFT, 4 STOP (src/utils/utils.py:9-10).
 */
#ifndef AGV_B200_H
#define AGV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AGV_OK 0
#define AGV_E_INVALID (-1)   /* bad argument / config */
#define AGV_E_CUDA (-2)      /* CUDA runtime error (agv_last_cuda_error for the code) */
#define AGV_E_NOMEM (-3)
#define AGV_E_UNSUPPORTED (-4) /* size outside the compiled kernel range */
#define AGV_E_BADGRID (-5)   /* grid code outside {0,1,2} */

#define AGV_UP 0
#define AGV_RIGHT 1
#define AGV_DOWN 2
#define AGV_LEFT 3
#define AGV_STOP 4
#define AGV_NO_TURN (-1)     /* action-log value for an AGV that took no turn */

/* control protocol of a turn: Scene.run_mode, src/multiAGVscene/Scene.py:136-162 */
#define AGV_PROTO_ASTAR 0        /* "A_star"/"auto"/"manual": no AGV-AGV test, is_end ignored */
#define AGV_PROTO_INTELLIGENT 1  /* "intelligent": collision test, episode ends on is_end */

/* action source of a fused tick */
#define AGV_POLICY_GIVEN 0   /* action_dev[E,N]; a negative entry falls back to GUIDED */
#define AGV_POLICY_ASTAR 1   /* Explorer.find_path_astar, Explorer.py:261-302 (matrix a) */
#define AGV_POLICY_GUIDED 2  /* Agent.find_action_astar on StateManager's valid matrix,
                                MADQN_structure/MADQN.py:192-200 (matrix b) */

/* passability matrix for agv_bfs_action */
#define AGV_MATRIX_ASTAR 0   /* Explorer.create_valid_matrix, Explorer.py:261-284 */
#define AGV_MATRIX_STATE 1   /* StateManager.create_path_matrix, StateManager.py:43-61 */

/* observation kinds: StateManager.create_state(obs_clip=...), StateManager.py:14-41 */
#define AGV_OBS_NONE 0
#define AGV_OBS_CLIP 1       /* [3, K, K], K = 2P+1, limited visual */
#define AGV_OBS_FULL 2       /* [3, H, W] */

typedef struct AgvHandle AgvHandle;

typedef struct AgvConfig {
    int32_t H, W;            /* Layout.scene_y_width / scene_x_width (<= 128 each) */
    int32_t E;               /* scenes (environments) resident on this GPU */
    int32_t N;               /* AGVs (Explorers) per scene, <= 64 */
    int32_t T;               /* len(Layout.task_list) per scene, <= 65535 */
    int32_t P;               /* StateManager padding_size; K = 2P+1 */
    int32_t always_loaded;   /* SuperParas.Explorer_Always_Loaded, settings.py:11 */
    int32_t always_empty;    /* SuperParas.Explorer_Always_Empty, settings.py:12 */
    int32_t shaped_reward;   /* SuperParas.Sparse_Reward, settings.py:18 */
    int32_t auto_reset;      /* 1: a scene whose episode is over is re-initialised
                                (Scene.init with its fixed task list) at its next tick */
    int64_t max_steps;       /* Scene.max_training_steps, Scene.py:24 */
    int32_t device;          /* CUDA device ordinal */
    int32_t reserved;
} AgvConfig;

/* ---- lifetime --------------------------------------------------------------
 * Replaces Layout.__init__/__create_layout (Layout.py:15-64,118-132) +
 * Explorer.__init__ (Explorer.py:13-43) + Scene.__init__ (Scene.py:15-59) for E
 * scenes sharing one layout.  grid_host: H*W codes 0 track / 1 storage / 2 picking.
 * Synchronous.  The scenes start "over" until agv_reset. */
int agv_create(const AgvConfig *cfg, const uint8_t *grid_host, AgvHandle **out);
int agv_destroy(AgvHandle *h);

/* Layout.task_list (Layout.py:43-49,104-116): tasks_host is [env_count, T, 4] uint8
 * (sx, sy, px, py), 1-based, for scenes [env_begin, env_begin+env_count).
 * Synchronous w.r.t. the host buffer. */
int agv_set_tasks(AgvHandle *h, const uint8_t *tasks_host, int32_t env_begin, int32_t env_count, void *stream);

/* Scene.init() (Scene.py:61-65) + the run_game prologue create_explorer of AGV 0
 * (Scene.py:86, Explorer.py:64-66).  env_mask_dev: uint8[E], NULL = all scenes. */
int agv_reset(AgvHandle *h, const uint8_t *env_mask_dev, void *stream);

/* ---- fused tick ------------------------------------------------------------
 * One pass of the while-loop body of Scene.run_mode (Scene.py:111-164) for every
 * scene: running_time += 1, AGVs act in list order (each sees the moves of the AGVs
 * before it), optional state encode of the acting AGV before (obs) and after
 * (next_obs) its move, then check_new_veh (Scene.py:354-366).
 * All outputs are optional (NULL) and indexed [scene, agv]:
 *   action_dev   int8  [E,N]  in (POLICY_GIVEN) / unused otherwise
 *   act_out_dev  int8  [E,N]  action taken, AGV_NO_TURN if the AGV took no turn
 *   reward_dev   float [E,N]  Explorer.check_action reward (-1/0/1 or shaped)
 *   done_dev     uint8 [E,N]  is_end after the Scene.py:155 correction
 *   plen_dev     uint16[E,N]  len(action_list) of the policy BFS (0 = STOP/none)
 *   slen_dev     uint16[E,N]  rectify_reward's path_length_new (Explorer.py:220),
 *                             0xFFFF when the shaped branch was not taken
 *   obs_dev / next_obs_dev  bf16 [E,N,3,K,K] or [E,N,3,H,W]; records of AGVs that took no
 *                             turn are NOT written (mask with act_out >= 0) */
int agv_tick(AgvHandle *h, int32_t proto, int32_t policy, int32_t obs_kind,
             const int8_t *action_dev, int8_t *act_out_dev, float *reward_dev, uint8_t *done_dev,
             uint16_t *plen_dev, uint16_t *slen_dev, void *obs_dev, void *next_obs_dev, void *stream);

/* Same, with HOST buffers for the per-step inputs and results (pinned memory):
 * copies action_host -> device, runs the tick, copies act/reward/done back.
 * obs stay on the device (they feed the CNN).  The copy-back runs on an internal
 * stream from double-buffered device results, so it overlaps the next call's tick;
 * agv_tick_host_join makes `stream` wait for every outstanding copy-back -- after it,
 * anything ordered on `stream` (an event, a synchronize) covers the host results.
 * Callers that need step k's results before issuing step k+1 join after each call.
 * The upload of action_host also runs on an internal stream (overlapping the previous
 * tick): the buffer must hold its final contents when the call is made and stay
 * untouched until the tick has run. */
int agv_tick_host(AgvHandle *h, int32_t proto, int32_t policy, int32_t obs_kind,
                  const int8_t *action_host, int8_t *act_out_host, float *reward_host, uint8_t *done_host,
                  void *obs_dev, void *next_obs_dev, void *stream);
int agv_tick_host_join(AgvHandle *h, void *stream);

/* ---- multi-tick rollout ------------------------------------------------------
 * n_ticks passes of the while-loop body of Scene.run_mode (Scene.py:111-164) for every scene in ONE
 * launch: exactly the results of n_ticks consecutive agv_tick calls with the same proto / policy /
 * obs_kind, written to rings indexed [tick, scene, agv]:
 *   action_dev [n_ticks,E,N] (POLICY_GIVEN: open-loop actions, negative = A* guidance),
 *   act_out / reward / done / plen / slen [n_ticks,E,N], obs / next_obs bf16 [n_ticks,E,N,3,K,K].
 * This is Scene.run_game("A_star") / Expert.create_data_by_self (ExpertManager.py:35-52) and the
 * A*-guided branch of AG-DQN (MADQN.py:96-111 with u >= epsilon) for E scenes at once: no network is
 * in the loop, so a scene never has to wait for the other scenes between its ticks.  Scenes are
 * scheduled longest-first (by the turns + BFS searches of their last tick) onto persistent CTAs; long
 * rollouts are taken 16 ticks per launch.  Inside a written group of 8 observation records, records of AGVs
 * without a turn are zero-filled; mask with act_out >= 0 as for agv_tick.  Configurations the persistent
 * kernel does not cover (shaped reward, full-map observations, K != 7) run as n_ticks fused ticks inside this
 * call; maps up to 32 rows x 64 columns run on the warp-per-scene kernel, which keeps every scene in its warp
 * for all n_ticks (one launch) -- same results. */
int agv_rollout(AgvHandle *h, int32_t n_ticks, int32_t proto, int32_t policy, int32_t obs_kind,
                const int8_t *action_dev, int8_t *act_out_dev, float *reward_dev, uint8_t *done_dev,
                uint16_t *plen_dev, uint16_t *slen_dev, void *obs_dev, void *next_obs_dev, void *stream);
/* Same with HOST (pinned) buffers for the [n_ticks,E,N] actions in and act / reward / done out;
 * double-buffered and overlapped like agv_tick_host, and pipelined inside the call: the rollout runs in
 * chunks of 4 ticks whose actions are uploaded and whose results are copied back while the neighbouring
 * chunks compute.  agv_tick_host_join waits for the copy-backs. */
int agv_rollout_host(AgvHandle *h, int32_t n_ticks, int32_t proto, int32_t policy, int32_t obs_kind,
                     const int8_t *action_host, int8_t *act_out_host, float *reward_host, uint8_t *done_host,
                     void *obs_dev, void *next_obs_dev, void *stream);

/* ---- sub-step protocol (NN in the loop) -------------------------------------
 * The callback protocol of Scene.py:144-158 split into tensor calls per AGV index k:
 *   agv_begin_tick; for k in 0..N-1: agv_encode(k) -> policy -> agv_step_turn(k)
 *   [-> agv_encode(k) again for next_obs]; agv_end_tick.                          */
int agv_begin_tick(AgvHandle *h, void *stream);  /* running_time += 1 (Scene.py:112) */
/* smart_controller.choose_action's StateManager.create_state for AGV k of every
 * scene: out_dev bf16 [E,3,K,K] / [E,3,H,W]; will_act_dev uint8[E] (optional) = 1
 * where AGV k takes a turn in this tick (Scene.py:124-134).  `post` = 1 encodes the
 * scenes where AGV k HAS just acted (store_info's next_obs, Scene.py:156). */
int agv_encode(AgvHandle *h, int32_t k, int32_t obs_kind, int32_t post, void *out_dev, uint8_t *will_act_dev,
               void *stream);
/* find_path_astar / find_action_astar for AGV k: action int8[E], pathlen uint16[E] */
int agv_bfs_action(AgvHandle *h, int32_t k, int32_t matrix, int8_t *action_dev, uint16_t *pathlen_dev,
                   void *stream);
/* Explorer.execute_action + Scene.py:155-158 for AGV k with action_dev int8[E]
 * (negative = GUIDED fallback).  acted_dev uint8[E] = 1 where the turn happened;
 * act_out_dev int8[E] (optional) = the action executed (the BFS action for a fallback),
 * AGV_NO_TURN where no turn happened. */
int agv_step_turn(AgvHandle *h, int32_t k, int32_t proto, const int8_t *action_dev, float *reward_dev,
                  uint8_t *done_dev, uint8_t *acted_dev, uint16_t *slen_dev, int8_t *act_out_dev, void *stream);
int agv_end_tick(AgvHandle *h, void *stream);    /* Scene.check_new_veh */

/* ---- state access (parity dumps, Scene.create_info Scene.py:341-352) ---------
 * hdr_host  int64[E,6]  : running_time, cursor, n_done, task_finished, n_created, over
 * agv_host  int32[E,N,8]: x, y, tx, ty, task_stage, loaded, all_assigned, task_order
 * Synchronous. */
int agv_get_state(AgvHandle *h, int64_t *hdr_host, int32_t *agv_host, void *stream);
int agv_set_state(AgvHandle *h, const int64_t *hdr_host, const int32_t *agv_host, void *stream);

/* counters accumulated on the device since the last call (then cleared):
 * out[0] = AGV turns taken, out[1] = episodes ended, out[2] = invalid action codes
 * seen (treated as STOP; the reference raises KeyError, utils.py:21). Synchronous. */
int agv_read_counters(AgvHandle *h, uint64_t out[3], void *stream);

/* ---- standalone BFS ("A*", src/utils/astar.py:53-153) -------------------------
 * B independent problems on one H x W 0/1 matrix each: valid_dev uint8[B,H,W],
 * start/target int32[B,2] 0-based (x,y).  Outputs: action int8[B] (first action or
 * STOP), pathlen int32[B] (len(action_list)), found uint8[B]. */
int agv_bfs_batch(int32_t B, int32_t H, int32_t W, const uint8_t *valid_dev, const int32_t *start_dev,
                  const int32_t *target_dev, int8_t *action_dev, int32_t *pathlen_dev, uint8_t *found_dev,
                  void *stream);
/* full distance field from the target over valid cells (uint16, 0xFFFF = unreachable);
 * used to rebuild FindPathAstar's path_list/action_list on the host. */
int agv_bfs_field(int32_t B, int32_t H, int32_t W, const uint8_t *valid_dev, const int32_t *target_dev,
                  uint16_t *dist_dev, void *stream);

/* ---- replay buffer (MADQN_structure/PER.py:6-112, MADQN.py:113-171) ----------- */
typedef struct AgvReplay AgvReplay;
/* capacity rows of (obs[obs_elems] bf16, action int8, reward f32, next_obs bf16, done u8)
 * + a float64 sum-tree of 2*capacity-1 nodes (PER.SumTree). */
int agv_replay_create(int32_t device, int64_t capacity, int32_t obs_elems, AgvReplay **out);
int agv_replay_destroy(AgvReplay *r);
/* Memory.add for every (scene, agv) row with act_dev[row] >= 0, in row order
 * (ring write, SumTree.add PER.py:41-52).  priority_dev: float64[rows] =
 * (|td error|+0.01)^0.6 (PER.py:79-80) or NULL for priority 1.  n_rows = E*N. */
int agv_replay_push(AgvReplay *r, int64_t n_rows, const int8_t *act_dev, const float *reward_dev,
                    const uint8_t *done_dev, const void *obs_dev, const void *next_obs_dev,
                    const double *priority_dev, void *stream);
/* Memory.sample(n) (PER.py:86-108) with the n uniforms injected: u01_dev float64[n]
 * in [0,1); s_i = seg*i + seg*u_i, sum-tree descent.  Writes the chosen data slots
 * (int64[n]), their priorities tree[idx] (float64[n], optional) and gathers the batch. */
int agv_replay_sample(AgvReplay *r, int32_t n, const double *u01_dev, int64_t *slot_dev, double *prio_dev,
                      void *obs_dev, int8_t *act_dev, float *reward_dev, void *next_obs_dev, uint8_t *done_dev,
                      void *stream);
/* Memory.update (PER.py:110-112) / SumTree.update (55-59) for n tree indices
 * (tree_idx = slot + capacity - 1, as returned by Memory.sample), applied in order. */
int agv_replay_update(AgvReplay *r, int32_t n, const int64_t *tree_idx_dev, const double *priority_dev,
                      void *stream);
/* uniform gather by explicit slots (int64[n]) */
int agv_replay_gather(AgvReplay *r, int32_t n, const int64_t *slot_dev, void *obs_dev, int8_t *act_dev,
                      float *reward_dev, void *next_obs_dev, uint8_t *done_dev, void *stream);
/* out[0] = n_entries, out[1] = write cursor; total_out = SumTree.total(). Synchronous. */
int agv_replay_info(AgvReplay *r, int64_t out[2], double *total_out, void *stream);

/* ---- diagnostics --------------------------------------------------------- */
const char *agv_strerror(int status);
int agv_last_cuda_error(void);
/* number of kernel launches issued through this library since load (for bench.py) */
uint64_t agv_launch_count(void);
const char *agv_version(void);
/* name of the kernel the last agv_tick / agv_tick_host call launched (three bit-identical
 * implementations exist; the library picks by shape, AGV_TICK_IMPL=async|lanes|warp forces one) */
const char *agv_last_tick_kernel(void);
/* profiling builds only (-DAGV_ROLL_PROF, AGV_ROLL_PROF=1): per-scene timeline of the last agv_rollout,
 * uint64[E,8] = {finish ns, BFS searches, turns, cycles waiting for BFS answers, start ns, pick-ups,
 * CTA, sum of BFS path lengths}; AGV_E_UNSUPPORTED otherwise.  Synchronous. */
int agv_rollout_profile(AgvHandle *h, uint64_t *out_host, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* AGV_B200_H */
