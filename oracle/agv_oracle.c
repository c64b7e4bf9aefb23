/*
 * oracle/agv_oracle.c -- CPU restatement of the reference's multi-AGV hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import, link or
 * execute this file.  It is used by tests/ (as the checker for the CUDA path),
 * by __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference
 * legs.  It is pinned against the live Python reference (tests/golden/, made by
 * tests/golden/make_golden.py in the build container) -- see DESIGN.md "Oracle".
 *
 * All file:line citations are relative to the reference checkout
 * (stoneMan349/Reinforcement-learning-for-multi-AGV-pathfinding).
 *
 * Conventions: positions are 1-based (x = column,
 * y = row); grid[(y-1)*W + (x-1)] in {0 track, 1 storage, 2 picking}
 * (src/multiAGVscene/Layout.py:118-132); actions 0 UP (0,-1), 1 RIGHT (+1,0),
 * 2 DOWN (0,+1), 3 LEFT (-1,0), 4 STOP (src/utils/utils.py:9-10,
 * src/multiAGVscene/Explorer.py:127-129).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_UP 0
#define ORC_RIGHT 1
#define ORC_DOWN 2
#define ORC_LEFT 3
#define ORC_STOP 4

/* control protocol of a turn (src/multiAGVscene/Scene.py:136-162) */
#define ORC_PROTO_ASTAR 0       /* "A_star"/"auto": no AGV-AGV test, is_end ignored */
#define ORC_PROTO_INTELLIGENT 1 /* "intelligent": collision test, episode ends on is_end */

/* action source of a fused tick */
#define ORC_POLICY_GIVEN 0  /* caller supplies action[N] */
#define ORC_POLICY_ASTAR 1  /* Explorer.find_path_astar on matrix (a), Explorer.py:261-302 */
#define ORC_POLICY_GUIDED 2 /* Agent.find_action_astar on matrix (b), MADQN.py:192-200 */

typedef struct orc_env {
    int H, W, N, T;
    int always_loaded, always_empty, shaped;
    long max_steps;
    uint8_t *grid; /* H*W */
    int *tasks;    /* T*4: sx, sy, px, py (1-based) */
    /* per-episode layout state (Layout.py:50-51,56-61) */
    long tick;     /* Scene.running_time */
    int cursor;    /* len(task_arrangement[0]) */
    int n_done;    /* sum(task_arrangement[2]) */
    uint8_t *done; /* task_arrangement[2] */
    int finished;  /* layout.task_finished */
    int n_created; /* created AGVs are always the prefix [0, n_created) (Scene.py:354-366) */
    int over;      /* run_mode returned */
    /* per-AGV state (Explorer.py:13-58) */
    int *x, *y, *tx, *ty, *stage, *loaded, *idle, *order;
    /* scratch */
    uint8_t *valid;
    int *dist;
    int *queue;
} orc_env;

static const int DX[5] = {0, 1, 0, -1, 0};
static const int DY[5] = {-1, 0, 1, 0, 0};

orc_env *orc_create(int H, int W, int N, int T, const uint8_t *grid, const int *tasks,
                    int always_loaded, int always_empty, int shaped, long max_steps) {
    orc_env *e = (orc_env *)calloc(1, sizeof(orc_env));
    e->H = H; e->W = W; e->N = N; e->T = T;
    e->always_loaded = always_loaded; e->always_empty = always_empty; e->shaped = shaped;
    e->max_steps = max_steps;
    e->grid = (uint8_t *)malloc((size_t)H * W);
    memcpy(e->grid, grid, (size_t)H * W);
    e->tasks = (int *)malloc(sizeof(int) * 4 * (T > 0 ? T : 1));
    if (T > 0) memcpy(e->tasks, tasks, sizeof(int) * 4 * T);
    e->done = (uint8_t *)calloc(T > 0 ? T : 1, 1);
    e->x = (int *)calloc(N, sizeof(int)); e->y = (int *)calloc(N, sizeof(int));
    e->tx = (int *)calloc(N, sizeof(int)); e->ty = (int *)calloc(N, sizeof(int));
    e->stage = (int *)calloc(N, sizeof(int)); e->loaded = (int *)calloc(N, sizeof(int));
    e->idle = (int *)calloc(N, sizeof(int)); e->order = (int *)calloc(N, sizeof(int));
    e->valid = (uint8_t *)malloc((size_t)H * W);
    e->dist = (int *)malloc(sizeof(int) * H * W);
    e->queue = (int *)malloc(sizeof(int) * H * W);
    for (int i = 0; i < N; i++) { /* Explorer.__init__, Explorer.py:27-28 */
        e->x[i] = e->y[i] = 1; e->tx[i] = e->ty[i] = 1;
    }
    return e;
}

void orc_destroy(orc_env *e) {
    if (!e) return;
    free(e->grid); free(e->tasks); free(e->done);
    free(e->x); free(e->y); free(e->tx); free(e->ty);
    free(e->stage); free(e->loaded); free(e->idle); free(e->order);
    free(e->valid); free(e->dist); free(e->queue);
    free(e);
}

void orc_set_tasks(orc_env *e, const int *tasks) { memcpy(e->tasks, tasks, sizeof(int) * 4 * e->T); }

/* Explorer.load_condition, Explorer.py:102-116 */
static void load_condition(orc_env *e, int i, int stage) {
    if (e->always_loaded) { e->loaded[i] = 1; return; }
    if (e->always_empty) { e->loaded[i] = 0; return; }
    e->loaded[i] = (stage == 0) ? 0 : 1;
}

/* Explorer.get_task, Explorer.py:68-100 (caller guarantees has_created) */
static void get_task(orc_env *e, int i) {
    if (e->T == e->cursor && e->stage[i] == 0) {
        e->idle[i] = 1; /* all_assigned */
        if (e->T == e->n_done) e->finished = 1;
        return; /* target_position and loaded keep their old values */
    }
    if (e->stage[i] == 0) {
        e->order[i] = e->cursor++;
        e->tx[i] = e->tasks[4 * e->order[i] + 0];
        e->ty[i] = e->tasks[4 * e->order[i] + 1];
    } else if (e->stage[i] == 1) {
        e->tx[i] = e->tasks[4 * e->order[i] + 2];
        e->ty[i] = e->tasks[4 * e->order[i] + 3];
    } else {
        e->tx[i] = e->tasks[4 * e->order[i] + 0];
        e->ty[i] = e->tasks[4 * e->order[i] + 1];
    }
    load_condition(e, i, e->stage[i]);
}

/* Explorer.continue_working, Explorer.py:251-259 */
static void continue_working(orc_env *e, int i) {
    e->stage[i]++;
    if (e->stage[i] == 3) {
        e->stage[i] = 0;
        if (!e->done[e->order[i]]) { e->done[e->order[i]] = 1; e->n_done++; }
    }
    get_task(e, i);
}

/* Scene.init (Scene.py:61-65) = Layout.init (Layout.py:56-64, fixed task list)
 * + Explorer.init for every AGV (Explorer.py:45-58; target_position is NOT reset)
 * followed by the run_game prologue explorer_group[0].create_explorer()
 * (Scene.py:86, Explorer.py:64-66) and running_time = 0 (Scene.py:101). */
void orc_reset(orc_env *e) {
    e->tick = 0; e->cursor = 0; e->n_done = 0; e->finished = 0; e->over = 0;
    memset(e->done, 0, e->T > 0 ? e->T : 1);
    for (int i = 0; i < e->N; i++) {
        e->x[i] = e->y[i] = 1;
        e->stage[i] = 0; e->loaded[i] = 0; e->idle[i] = 0; e->order[i] = 0;
    }
    e->n_created = 0;
    if (e->N > 0) { e->n_created = 1; get_task(e, 0); }
}

/* Passability matrix (a): Explorer.create_valid_matrix, Explorer.py:261-284.
 * Uses AGV i's *current* position list of every explorer (created or not). */
void orc_valid_a(const orc_env *e, int i, uint8_t *out) {
    for (int r = 0; r < e->H; r++)
        for (int c = 0; c < e->W; c++) {
            int code = e->grid[r * e->W + c], v = 0;
            if (code == 0) v = 1;
            else if (code == 1) v = e->loaded[i] ? 0 : 1;
            else v = 0;
            if (e->tx[i] - 1 == c && e->ty[i] - 1 == r) v = 1;
            out[r * e->W + c] = (uint8_t)v;
        }
    if (e->N > 1)
        for (int j = 0; j < e->N; j++) out[(e->y[j] - 1) * e->W + (e->x[j] - 1)] = 0;
}

/* Passability matrix (b): StateManager.create_path_matrix + layout_to_matrix,
 * src/algorithm/Manager/StateManager.py:43-61,92-120, with occupied_place = the
 * other *created* AGVs (all_info_analysis 122-133 over Scene.create_info 341-352). */
void orc_valid_b(const orc_env *e, int i, uint8_t *out) {
    for (int r = 0; r < e->H; r++)
        for (int c = 0; c < e->W; c++) {
            int code = e->grid[r * e->W + c];
            out[r * e->W + c] = (uint8_t)((code == 0) || (code == 1 && !e->loaded[i]));
        }
    out[(e->y[i] - 1) * e->W + (e->x[i] - 1)] = 1;
    out[(e->ty[i] - 1) * e->W + (e->tx[i] - 1)] = 1;
    for (int j = 0; j < e->n_created; j++)
        if (j != i) out[(e->y[j] - 1) * e->W + (e->x[j] - 1)] = 0;
}

/* ---------------------------------------------------------------------------
 * "A*" (src/utils/astar.py:73-153).  Two restatements:
 *  orc_bfs_literal: follows astar_method step by step (FIFO open list because
 *    all_cost is never assigned, astar.py:14,86; replace-and-move-to-back when
 *    c.f >= n.f, astar.py:108-117; closed test 94-101), then derives the action
 *    list as astar_plot_action_route does (139-153).  O(V^2); for pinning.
 *  orc_bfs_field: target-rooted distance field + priority rule RIGHT, DOWN, UP,
 *    LEFT among neighbours one step closer.  O(V).
 * Coordinates here are 0-based (x, y) as in the reference's callers.
 * Both return the first action (ORC_STOP when not found or empty) and write
 * *found and *len (= len(action_list)).
 * ------------------------------------------------------------------------- */
typedef struct { int pos; int parent; int g; long f; } lit_cell;

int orc_bfs_literal(const uint8_t *valid, int H, int W, int sx, int sy, int tx, int ty,
                    int *found, int *len, int8_t *actions_out, int actions_cap) {
    /* neighbour order astar.py:33-38: (-1,0) (0,-1) (0,1) (1,0) */
    static const int NX[4] = {-1, 0, 0, 1};
    static const int NY[4] = {0, -1, 1, 0};
    int V = H * W;
    lit_cell *cells = (lit_cell *)malloc(sizeof(lit_cell) * (size_t)(4 * V + 8));
    int ncell = 0;
    int *open = (int *)malloc(sizeof(int) * (size_t)(V + 8));
    int nopen = 0;
    uint8_t *closed = (uint8_t *)calloc(V, 1);
    cells[ncell] = (lit_cell){sy * W + sx, -1, 0, 0};
    open[nopen++] = ncell++;
    int target = ty * W + tx, cur = -1;
    *found = 0; *len = 0;
    while (nopen > 0) {
        cur = open[0]; /* argmin of all-zero costs = index 0 */
        memmove(open, open + 1, sizeof(int) * (size_t)(nopen - 1));
        nopen--;
        closed[cells[cur].pos] = 1;
        if (cells[cur].pos == target) { *found = 1; break; }
        int cx = cells[cur].pos % W, cy = cells[cur].pos / W;
        for (int k = 0; k < 4; k++) {
            int x = cx + NX[k], y = cy + NY[k];
            if (!(0 <= x && x < W && 0 <= y && y < H && valid[y * W + x] == 1)) continue;
            int p = y * W + x;
            if (closed[p]) continue;
            long h = (long)(ty - y) * (ty - y) + (long)(tx - x) * (tx - x);
            lit_cell n = {p, cur, cells[cur].g + 1, 0};
            n.f = h + n.g;
            int already = 0;
            for (int q = 0; q < nopen; q++)
                if (cells[open[q]].pos == p) {
                    if (cells[open[q]].f < n.f) {
                    } else {
                        memmove(open + q, open + q + 1, sizeof(int) * (size_t)(nopen - q - 1));
                        nopen--;
                        cells[ncell] = n;
                        open[nopen++] = ncell++;
                    }
                    already = 1;
                    break;
                }
            if (!already) { cells[ncell] = n; open[nopen++] = ncell++; }
        }
    }
    int first = ORC_STOP;
    if (*found) {
        /* path from target back to start; actions from start to target */
        int n = 0;
        for (int c = cur; cells[c].parent >= 0; c = cells[c].parent) n++;
        *len = n;
        int idx = n - 1;
        for (int c = cur; cells[c].parent >= 0; c = cells[c].parent, idx--) {
            int p = cells[c].pos, pp = cells[cells[c].parent].pos;
            int a = ORC_STOP;
            /* astar.py:145-152: the y test overrides the x test */
            if (pp % W < p % W) a = ORC_RIGHT; else if (pp % W > p % W) a = ORC_LEFT;
            if (pp / W < p / W) a = ORC_DOWN; else if (pp / W > p / W) a = ORC_UP;
            if (actions_out && idx < actions_cap) actions_out[idx] = (int8_t)a;
            if (idx == 0) first = a;
        }
    }
    free(cells); free(open); free(closed);
    return first;
}

int orc_bfs_field(const uint8_t *valid, int H, int W, int sx, int sy, int tx, int ty,
                  int *found, int *len, int *dist, int *queue) {
    *found = 0; *len = 0;
    if (sx == tx && sy == ty) { *found = 1; return ORC_STOP; } /* empty action list */
    if (valid[ty * W + tx] != 1) return ORC_STOP;              /* target never enters open */
    int V = H * W;
    for (int i = 0; i < V; i++) dist[i] = -1;
    int head = 0, tail = 0;
    dist[ty * W + tx] = 0; queue[tail++] = ty * W + tx;
    while (head < tail) {
        int p = queue[head++], px = p % W, py = p / W;
        for (int k = 0; k < 4; k++) {
            int x = px + DX[k], y = py + DY[k];
            if (x < 0 || x >= W || y < 0 || y >= H) continue;
            int q = y * W + x;
            if (valid[q] != 1 || dist[q] >= 0) continue;
            dist[q] = dist[p] + 1; queue[tail++] = q;
        }
    }
    static const int PRIO[4] = {ORC_RIGHT, ORC_DOWN, ORC_UP, ORC_LEFT};
    int best = -1;
    for (int k = 0; k < 4; k++) {
        int x = sx + DX[PRIO[k]], y = sy + DY[PRIO[k]];
        if (x < 0 || x >= W || y < 0 || y >= H) continue;
        int q = y * W + x;
        if (valid[q] != 1 || dist[q] < 0) continue;
        if (best < 0 || dist[q] < best) best = dist[q];
    }
    if (best < 0) return ORC_STOP;
    *found = 1; *len = best + 1;
    for (int k = 0; k < 4; k++) {
        int x = sx + DX[PRIO[k]], y = sy + DY[PRIO[k]];
        if (x < 0 || x >= W || y < 0 || y >= H) continue;
        int q = y * W + x;
        if (valid[q] == 1 && dist[q] == best) return PRIO[k];
    }
    return ORC_STOP;
}

/* Explorer.find_path_astar (Explorer.py:286-302): matrix (a), STOP if not found/empty */
int orc_action_astar(orc_env *e, int i, int *len) {
    int found, l;
    orc_valid_a(e, i, e->valid);
    int a = orc_bfs_field(e->valid, e->H, e->W, e->x[i] - 1, e->y[i] - 1, e->tx[i] - 1, e->ty[i] - 1,
                          &found, &l, e->dist, e->queue);
    if (len) *len = l;
    return (!found || l == 0) ? ORC_STOP : a;
}

/* Agent.find_action_astar on StateManager's valid_path_matrix
 * (MADQN_structure/MADQN.py:192-200, Controller.py:116-121) */
int orc_action_guided(orc_env *e, int i, int *len) {
    int found, l;
    orc_valid_b(e, i, e->valid);
    int a = orc_bfs_field(e->valid, e->H, e->W, e->x[i] - 1, e->y[i] - 1, e->tx[i] - 1, e->ty[i] - 1,
                          &found, &l, e->dist, e->queue);
    if (len) *len = l;
    return (!found || l == 0) ? ORC_STOP : a;
}

/* Does AGV i take a turn this tick? (Scene.py:124-134).  Returns 1 act, 0 skip
 * (idle -> continue), -1 stop the loop (uncreated -> break), -2 episode returns
 * (task_finished seen at turn start). */
int orc_turn_state(const orc_env *e, int i) {
    if (i >= e->n_created) return -1;
    if (e->finished) return -2;
    if (e->idle[i]) return 0;
    return 1;
}

/* Explorer.execute_action + check_action (Explorer.py:131-208) and, for
 * PROTO_INTELLIGENT, the is_end correction of Scene.py:155.
 * reward_int in {-1,0,1}; shaped reward (settings.Sparse_Reward, Explorer.py:206,
 * 210-249) is returned as path_len_new (>=0) with *shaped_used=1 and the float
 * value 0.001*(M-L)/M, M=0.5*(H+W) in *reward_f64.  */
void orc_execute(orc_env *e, int i, int action, int proto,
                 int *reward_int, double *reward_f64, int *is_end, int *shaped_used, int *path_len_new) {
    int nx = e->x[i] + DX[action], ny = e->y[i] + DY[action];
    int r = 0, end = 0;
    *shaped_used = 0; *path_len_new = 0;
    do {
        if (nx < 1 || ny < 1 || nx > e->W || ny > e->H) { r = -1; end = 1; break; }
        int code = e->grid[(ny - 1) * e->W + (nx - 1)];
        int at_target = (nx == e->tx[i] && ny == e->ty[i]);
        if (code == 1 && e->loaded[i] && !at_target) { r = -1; end = 1; break; }
        if (code == 2 && !at_target) { r = -1; end = 1; break; }
        if (proto == ORC_PROTO_INTELLIGENT) { /* all_info = created AGVs, Scene.py:145,341-352 */
            int hit = 0;
            for (int j = 0; j < e->n_created; j++)
                if (j != i && e->x[j] == nx && e->y[j] == ny) { hit = 1; break; }
            if (hit) { r = -1; end = 1; break; }
        }
        if (at_target) { r = 1; end = 0; continue_working(e, i); break; }
        if (e->shaped) { /* rectify_reward: BFS from the proposed cell on matrix (a) */
            int found, l;
            orc_valid_a(e, i, e->valid);
            orc_bfs_field(e->valid, e->H, e->W, nx - 1, ny - 1, e->tx[i] - 1, e->ty[i] - 1,
                          &found, &l, e->dist, e->queue);
            *shaped_used = 1; *path_len_new = l;
        }
    } while (0);
    if (!end) { e->x[i] = nx; e->y[i] = ny; } /* Explorer.py:145-148 */
    if (proto == ORC_PROTO_INTELLIGENT && (e->finished || e->tick >= e->max_steps)) end = 1;
    *reward_int = r;
    if (*shaped_used) {
        double M = 0.5 * (double)(e->H + e->W);
        *reward_f64 = 0.001 * (M - (double)*path_len_new) / M;
    } else {
        *reward_f64 = (double)r;
    }
    *is_end = end;
}

/* Scene.check_new_veh (Scene.py:354-366) */
int orc_spawn(orc_env *e) {
    if (e->n_created >= e->N) return 0;
    for (int j = 0; j < e->n_created; j++)
        if (e->x[j] == 1 && e->y[j] == 1) return 0;
    int k = e->n_created++;
    get_task(e, k);
    return k;
}

void orc_begin_tick(orc_env *e) { e->tick++; }

void orc_encode(const orc_env *e, int i, int clip, int P, float *out);

/* One pass of the while-loop body of Scene.run_mode (Scene.py:111-164) with a
 * built-in action source.  Per-AGV logs (length N): act[i] = action taken or -1
 * if AGV i took no turn, rew[i] (double), end[i], plen[i] = len(action_list) of
 * the policy BFS, slen[i] = rectify_reward's path_length_new or 0xFFFF.
 * obs_kind 0 none / 1 clip [3,K,K] / 2 full [3,H,W]: obs (before the move, what
 * choose_action sees) and next_obs (after, what store_info sees) per AGV, zeros
 * for AGVs without a turn; either may be NULL.  A given action < 0 falls back to
 * the guided policy.  Returns the number of turns taken.  Sets e->over when
 * run_mode would have returned. */
int orc_tick_obs(orc_env *e, int proto, int policy, const int8_t *given,
                 int8_t *act, double *rew, uint8_t *end, int *plen, int *slen,
                 int obs_kind, int P, float *obs, float *next_obs) {
    int K = 2 * P + 1;
    size_t R = obs_kind == 1 ? (size_t)3 * K * K : (obs_kind == 2 ? (size_t)3 * e->H * e->W : 0);
    for (int i = 0; i < e->N; i++) { act[i] = -1; rew[i] = 0; end[i] = 0; plen[i] = 0; if (slen) slen[i] = 0xFFFF; }
    if (obs && R) memset(obs, 0, sizeof(float) * R * e->N);
    if (next_obs && R) memset(next_obs, 0, sizeof(float) * R * e->N);
    if (e->over) return 0;
    int turns = 0;
    e->tick++;
    for (int i = 0; i < e->N; i++) {
        int st = orc_turn_state(e, i);
        if (st == -1) break;
        if (st == -2) { e->over = 1; return turns; }
        if (st == 0) continue;
        if (obs && R) orc_encode(e, i, obs_kind == 1, P, obs + R * i);
        int a, l = 0;
        if (policy == ORC_POLICY_GIVEN && given[i] >= 0) a = given[i];
        else if (policy == ORC_POLICY_ASTAR) a = orc_action_astar(e, i, &l);
        else a = orc_action_guided(e, i, &l);
        int ri, ie, su, pl; double rf;
        orc_execute(e, i, a, proto, &ri, &rf, &ie, &su, &pl);
        if (proto == ORC_PROTO_ASTAR) ie = 0; /* ignored by the loop, Scene.py:159-162 */
        act[i] = (int8_t)a; rew[i] = rf; end[i] = (uint8_t)ie; plen[i] = l;
        if (slen && su) slen[i] = pl;
        if (next_obs && R) orc_encode(e, i, obs_kind == 1, P, next_obs + R * i);
        turns++;
        if (proto == ORC_PROTO_INTELLIGENT && ie) { e->over = 1; return turns; }
    }
    orc_spawn(e);
    return turns;
}

int orc_tick(orc_env *e, int proto, int policy, const int8_t *given,
             int8_t *act, double *rew, uint8_t *end, int *plen) {
    return orc_tick_obs(e, proto, policy, given, act, rew, end, plen, 0, 0, 0, 0, 0);
}

/* StateManager.create_state (StateManager.py:14-41) for AGV i of the current
 * snapshot: channels [current one-hot, target one-hot, valid (b)].
 * clip=0: out is 3*H*W; clip=1: out is 3*K*K, K=2P+1, window centred on the AGV,
 * zero padding, off-window target clamped to the border (clip_matrix 145-192). */
void orc_encode(const orc_env *e, int i, int clip, int P, float *out) {
    int H = e->H, W = e->W;
    uint8_t *vb = (uint8_t *)malloc((size_t)H * W);
    orc_valid_b(e, i, vb);
    int cx = e->x[i] - 1, cy = e->y[i] - 1, tx = e->tx[i] - 1, ty = e->ty[i] - 1;
    if (!clip) {
        memset(out, 0, sizeof(float) * 3 * H * W);
        out[cy * W + cx] = 1.f;
        out[H * W + ty * W + tx] = 1.f;
        for (int p = 0; p < H * W; p++) out[2 * H * W + p] = vb[p] ? 1.f : 0.f;
    } else {
        int K = 2 * P + 1;
        memset(out, 0, sizeof(float) * 3 * K * K);
        out[P * K + P] = 1.f;
        int dx = tx - cx, dy = ty - cy;
        if (dx < -P) dx = -P; if (dx > P) dx = P;
        if (dy < -P) dy = -P; if (dy > P) dy = P;
        out[K * K + (dy + P) * K + (dx + P)] = 1.f;
        for (int wy = 0; wy < K; wy++)
            for (int wx = 0; wx < K; wx++) {
                int r = cy - P + wy, c = cx - P + wx;
                float v = 0.f;
                if (r >= 0 && r < H && c >= 0 && c < W) v = vb[r * W + c] ? 1.f : 0.f;
                out[2 * K * K + wy * K + wx] = v;
            }
    }
    free(vb);
}

/* Scene.create_info snapshot (Scene.py:341-352) flattened: for each created AGV
 * [x, y, tx, ty, loaded].  Returns n_created. */
int orc_snapshot(const orc_env *e, int *out5) {
    for (int j = 0; j < e->n_created; j++) {
        out5[5 * j + 0] = e->x[j]; out5[5 * j + 1] = e->y[j];
        out5[5 * j + 2] = e->tx[j]; out5[5 * j + 3] = e->ty[j];
        out5[5 * j + 4] = e->loaded[j];
    }
    return e->n_created;
}

/* full state dump for parity: hdr[6] = tick,cursor,n_done,finished,n_created,over;
 * agv[N*8] = x,y,tx,ty,stage,loaded,idle,order */
void orc_get_state(const orc_env *e, long *hdr, int *agv) {
    hdr[0] = e->tick; hdr[1] = e->cursor; hdr[2] = e->n_done; hdr[3] = e->finished;
    hdr[4] = e->n_created; hdr[5] = e->over;
    for (int j = 0; j < e->N; j++) {
        agv[8 * j + 0] = e->x[j]; agv[8 * j + 1] = e->y[j]; agv[8 * j + 2] = e->tx[j]; agv[8 * j + 3] = e->ty[j];
        agv[8 * j + 4] = e->stage[j]; agv[8 * j + 5] = e->loaded[j]; agv[8 * j + 6] = e->idle[j];
        agv[8 * j + 7] = e->order[j];
    }
}

/* Bulk driver for CPU-baseline timing and long parity runs: runs `nticks` ticks
 * with the given protocol/policy, auto-resetting (Scene.init with the fixed task
 * list) whenever the episode is over.  When obs != NULL every turn also runs the
 * limited-visual encode of the acting AGV (one create_state per turn) into obs
 * (3*K*K floats, overwritten).  Returns turns taken; *checksum accumulates a
 * position hash so that work cannot be optimised away. */
long orc_run(orc_env *e, int proto, int policy, long nticks, int encode_P, float *obs, unsigned long *checksum) {
    long turns = 0;
    unsigned long cs = checksum ? *checksum : 0;
    for (long t = 0; t < nticks; t++) {
        if (e->over) orc_reset(e);
        e->tick++;
        int returned = 0;
        for (int i = 0; i < e->N; i++) {
            int st = orc_turn_state(e, i);
            if (st == -1) break;
            if (st == -2) { e->over = 1; returned = 1; break; }
            if (st == 0) continue;
            if (obs) orc_encode(e, i, 1, encode_P, obs);
            int a, l = 0;
            if (policy == ORC_POLICY_ASTAR) a = orc_action_astar(e, i, &l);
            else a = orc_action_guided(e, i, &l);
            int ri, ie, su, pl; double rf;
            orc_execute(e, i, a, proto, &ri, &rf, &ie, &su, &pl);
            turns++;
            cs = cs * 1000003UL + (unsigned long)(e->x[i] * 131 + e->y[i] + a * 7 + ri + 2);
            if (proto == ORC_PROTO_INTELLIGENT && ie) { e->over = 1; returned = 1; break; }
        }
        if (!returned) orc_spawn(e);
    }
    if (checksum) *checksum = cs;
    return turns;
}
