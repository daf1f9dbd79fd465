/*
 * heist_oracle.c -- CPU restatement of the Heist Architect environment hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (the CUDA library behind include/heist_b200.h) never calls into this file.
 *
 * Parity status: PINNED.  Every function below is checked bit-for-bit against the
 * unmodified Python reference (imported from /root/reference in the build container)
 * through the fixtures in tests/golden/ (generator: tests/golden/make_golden.py).
 *
 * All file:line citations are relative to the reference tree
 * (Shanmuk4622/RL-Project-Heist-Architect-...-CSE4019).  The reference is pure Python;
 * its floating point is CPython double arithmetic + the platform libm (glibc), which is
 * exactly what this file uses (compile with -ffp-contract=off, no -ffast-math).
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off -pthread -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

/* TileType, heist_architect/utils.py:31-37 */
enum { T_EMPTY = 0, T_WALL = 1, T_START = 2, T_VAULT = 3, T_CAMERA = 4, T_GUARD = 5 };
/* status codes shared with include/heist_b200.h */
enum { ST_RUNNING = 0, ST_DETECTED = 1, ST_VAULT = 2, ST_TIMEOUT = 3, ST_ALREADY_DONE = 4 };

/* BUDGET_COSTS, heist_architect/components/budget.py:13-17 */
#define COST_WALL 1
#define COST_CAMERA 3
#define COST_GUARD 5

typedef struct {
    int row, col;
    double fov, heading, speed;
    int range;
} OCam;

typedef struct {
    int len, speed, idx, range;
    double fov, heading;
    int *path; /* len x 2 (row, col) */
} OGuard;

typedef struct {
    int R, C, max_steps;
    int start_r, start_c, vault_r, vault_c;
    int budget_total, budget_spent;
    double reward_vault, reward_detection, reward_step;
    int32_t *grid; /* R*C tile codes */
    float *vis;    /* R*C, 0/1 */
    uint8_t *wallmask;
    int n_walls;
    int n_cams, cap_cams;
    OCam *cams;
    int n_guards, cap_guards;
    OGuard *guards;
    int solver_r, solver_c, tick;
    int done, detected, vault_reached;
    int prev_dist, init_dist;
} OEnv;

/* ------------------------------------------------------------------------------------ */
/* CPython float semantics                                                               */
/* ------------------------------------------------------------------------------------ */

/* Python float %  (Objects/floatobject.c float_rem): result takes the divisor's sign. */
static double py_fmod(double x, double y) {
    double m = fmod(x, y);
    if (m != 0.0) {
        if ((y < 0) != (m < 0)) m += y;
    } else {
        m = copysign(0.0, y);
    }
    return m;
}
/* math.radians / math.degrees (Modules/mathmodule.c): x * (pi/180), x * (180/pi) */
static const double PY_PI = 3.14159265358979323846;
static double py_radians(double x) { return x * (PY_PI / 180.0); }
static double py_degrees(double x) { return x * (180.0 / PY_PI); }
/* round(float) / round(np.float64) with ndigits=None: half-to-even -> rint in RN mode */
static long py_round(double x) { return (long)rint(x); }
static int iabs(int a) { return a < 0 ? -a : a; }

/* ------------------------------------------------------------------------------------ */
/* Env lifetime                                                                          */
/* ------------------------------------------------------------------------------------ */

/* create_empty_grid, utils.py:131-139 ; START/VAULT marks, environment.py:171-173 */
static void fresh_grid(OEnv *e) {
    int R = e->R, C = e->C;
    for (int i = 0; i < R * C; ++i) e->grid[i] = T_EMPTY;
    for (int c = 0; c < C; ++c) { e->grid[c] = T_WALL; e->grid[(R - 1) * C + c] = T_WALL; }
    for (int r = 0; r < R; ++r) { e->grid[r * C] = T_WALL; e->grid[r * C + C - 1] = T_WALL; }
    e->grid[e->start_r * C + e->start_c] = T_START;
    e->grid[e->vault_r * C + e->vault_c] = T_VAULT;
}

static void free_assets(OEnv *e) {
    for (int g = 0; g < e->n_guards; ++g) free(e->guards[g].path);
    e->n_guards = 0;
    e->n_cams = 0;
    e->n_walls = 0;
}

/* HeistEnvironment.__init__, environment.py:62-96 ; EnvironmentConfig :18-37 */
OEnv *oenv_create(int R, int C, int max_steps, int start_r, int start_c, int vault_r, int vault_c,
                  int budget, double reward_vault, double reward_detection, double reward_step) {
    OEnv *e = (OEnv *)calloc(1, sizeof(OEnv));
    e->R = R; e->C = C; e->max_steps = max_steps;
    e->start_r = start_r; e->start_c = start_c;
    e->vault_r = vault_r; e->vault_c = vault_c;
    e->budget_total = budget; e->budget_spent = 0;
    e->reward_vault = reward_vault; e->reward_detection = reward_detection; e->reward_step = reward_step;
    e->grid = (int32_t *)malloc(sizeof(int32_t) * R * C);
    e->vis = (float *)calloc(R * C, sizeof(float));
    e->wallmask = (uint8_t *)calloc(R * C, 1);
    e->cap_cams = 0; e->cams = NULL;
    e->cap_guards = 0; e->guards = NULL;
    fresh_grid(e);
    e->solver_r = start_r; e->solver_c = start_c;
    e->prev_dist = iabs(start_r - vault_r) + iabs(start_c - vault_c);
    e->init_dist = e->prev_dist;
    return e;
}

void oenv_free(OEnv *e) {
    if (!e) return;
    free_assets(e);
    free(e->cams); free(e->guards);
    free(e->grid); free(e->vis); free(e->wallmask);
    free(e);
}

/* BudgetManager.scale_budget, budget.py:64-67 */
void oenv_scale_budget(OEnv *e, int b) { e->budget_total = b; e->budget_spent = 0; }

/* BudgetManager.purchase, budget.py:48-58 */
static int purchase(OEnv *e, int cost) {
    if (e->budget_total - e->budget_spent >= cost) { e->budget_spent += cost; return 1; }
    return 0;
}

/* _is_valid_placement, environment.py:160-167 */
static int valid_placement(const OEnv *e, int r, int c) {
    if (r <= 0 || r >= e->R - 1) return 0;
    if (c <= 0 || c >= e->C - 1) return 0;
    return e->grid[r * e->C + c] == T_EMPTY;
}

/* bfs_path_exists, utils.py:52-85 : 4-connected, passable <=> tile != WALL */
int oracle_bfs(const int32_t *grid, int R, int C, int sr, int sc, int gr, int gc) {
    if (sr == gr && sc == gc) return 1;
    uint8_t *seen = (uint8_t *)calloc(R * C, 1);
    int *q = (int *)malloc(sizeof(int) * R * C);
    int head = 0, tail = 0, found = 0;
    q[tail++] = sr * C + sc; seen[sr * C + sc] = 1;
    static const int DR[4] = {-1, 1, 0, 0}, DC[4] = {0, 0, -1, 1};
    while (head < tail && !found) {
        int cur = q[head++], r = cur / C, c = cur % C;
        for (int k = 0; k < 4; ++k) {
            int nr = r + DR[k], nc = c + DC[k];
            if (nr < 0 || nr >= R || nc < 0 || nc >= C) continue;
            if (seen[nr * C + nc]) continue;
            if (grid[nr * C + nc] == T_WALL) continue;
            if (nr == gr && nc == gc) { found = 1; break; }
            seen[nr * C + nc] = 1; q[tail++] = nr * C + nc;
        }
    }
    free(seen); free(q);
    return found;
}

/* is_level_valid, environment.py:154-158 */
int oenv_is_valid(const OEnv *e) {
    return oracle_bfs(e->grid, e->R, e->C, e->start_r, e->start_c, e->vault_r, e->vault_c);
}

/*
 * set_layout, environment.py:102-152 (+ _reset_layout :169-177).
 *   walls:  wall_rc[2*i] = row, [2*i+1] = col
 *   cams:   cam_rc like walls; cam_f[3*i] = fov, [3*i+1] = heading, [3*i+2] = rotation_speed
 *   guards: guard_path is [n_guards][path_stride][2]; guard_len[i] waypoints used
 * Returns BFS validity.  Waypoints must be inside the grid (the reference would wrap
 * negative indices / raise IndexError; both are out of scope) -> returns -1.
 */
int oenv_set_layout(OEnv *e, int n_walls, const int *wall_rc, int n_cams, const int *cam_rc,
                    const double *cam_f, const int *cam_range, int n_guards, const int *guard_len,
                    const int *guard_path, int path_stride, const int *guard_speed,
                    const int *guard_range, const double *guard_fov) {
    free_assets(e);
    fresh_grid(e);
    e->budget_spent = 0; /* budget.reset(), budget.py:60-62 */
    if (n_cams > e->cap_cams) { e->cams = (OCam *)realloc(e->cams, sizeof(OCam) * n_cams); e->cap_cams = n_cams; }
    if (n_guards > e->cap_guards) { e->guards = (OGuard *)realloc(e->guards, sizeof(OGuard) * n_guards); e->cap_guards = n_guards; }
    for (int g = 0; g < n_guards; ++g)
        for (int k = 0; k < guard_len[g]; ++k) {
            int r = guard_path[(g * path_stride + k) * 2], c = guard_path[(g * path_stride + k) * 2 + 1];
            if (r < 0 || r >= e->R || c < 0 || c >= e->C) return -1;
        }
    /* walls :118-121 -- validity first (short-circuit), then purchase */
    for (int i = 0; i < n_walls; ++i) {
        int r = wall_rc[2 * i], c = wall_rc[2 * i + 1];
        if (valid_placement(e, r, c) && purchase(e, COST_WALL)) {
            e->grid[r * e->C + c] = T_WALL;
            e->n_walls++;
        }
    }
    /* cameras :124-135 */
    for (int i = 0; i < n_cams; ++i) {
        int r = cam_rc[2 * i], c = cam_rc[2 * i + 1];
        if (valid_placement(e, r, c) && purchase(e, COST_CAMERA)) {
            OCam *cam = &e->cams[e->n_cams++];
            cam->row = r; cam->col = c;
            cam->fov = cam_f[3 * i]; cam->heading = cam_f[3 * i + 1]; cam->speed = cam_f[3 * i + 2];
            cam->range = cam_range[i];
            e->grid[r * e->C + c] = T_CAMERA;
        }
    }
    /* guards :138-149 -- no placement check; start tile overwritten unconditionally */
    for (int i = 0; i < n_guards; ++i) {
        if (guard_len[i] > 0 && purchase(e, COST_GUARD)) {
            OGuard *g = &e->guards[e->n_guards++];
            g->len = guard_len[i]; g->speed = guard_speed[i]; g->idx = 0;
            g->range = guard_range[i]; g->fov = guard_fov[i]; g->heading = 0.0;
            g->path = (int *)malloc(sizeof(int) * 2 * g->len);
            memcpy(g->path, guard_path + (size_t)i * path_stride * 2, sizeof(int) * 2 * g->len);
            e->grid[g->path[0] * e->C + g->path[1]] = T_GUARD;
        }
    }
    return oenv_is_valid(e);
}

/* ------------------------------------------------------------------------------------ */
/* Visibility                                                                            */
/* ------------------------------------------------------------------------------------ */

/*
 * Camera.get_vision_cone_tiles, security.py:53-101 (sub=1) and
 * Guard.get_visible_tiles, security.py:161-192 (sub=0): ray-march, marks tiles in `out`.
 * The reference's `visible` list + later vis[r,c]=1 loop is a set union, so marking
 * directly is equivalent.
 */
static void cone(const OEnv *e, int row, int col, double fov, double heading, int range, int substeps,
                 float *out) {
    int R = e->R, C = e->C;
    double half_fov = fov / 2.0;
    double two_fov = fov * 2;
    int num_rays = (int)two_fov; /* int(): truncation toward zero */
    if (num_rays < 30) num_rays = 30;
    for (int i = 0; i <= num_rays; ++i) {
        double angle_deg = heading - half_fov + (fov * i / num_rays);
        double angle_rad = py_radians(angle_deg);
        double dx = cos(angle_rad);
        double dy = -sin(angle_rad);
        int blocked = 0;
        for (int step = 1; step <= range && !blocked; ++step) {
            int nsub = substeps ? 3 : 1;
            for (int s = 0; s < nsub; ++s) {
                double dist;
                if (substeps) {
                    /* np.linspace(0, 1, 3) == [0.0, 0.5, 1.0] ; dist = step - 1 + sub * 1.0 */
                    double sub = 0.5 * s;
                    dist = (double)(step - 1) + sub * 1.0;
                    if (dist == 0) continue;
                } else {
                    dist = (double)step;
                }
                double fx = col + dx * dist;
                double fy = row + dy * dist;
                long c = py_round(fx), r = py_round(fy);
                if (0 <= r && r < R && 0 <= c && c < C) {
                    if (e->wallmask[r * C + c]) { blocked = 1; break; }
                    if (!(r == row && c == col)) out[r * C + c] = 1.0f;
                } else { blocked = 1; break; }
            }
        }
    }
}

/* DynamicVisibilityMap.update, visibility.py:31-65 with wall_mask = (grid == WALL),
 * environment.py:211,257 */
static void update_visibility(OEnv *e) {
    int n = e->R * e->C;
    for (int i = 0; i < n; ++i) { e->wallmask[i] = (e->grid[i] == T_WALL); e->vis[i] = 0.0f; }
    for (int k = 0; k < e->n_cams; ++k) {
        OCam *cam = &e->cams[k];
        cone(e, cam->row, cam->col, cam->fov, cam->heading, cam->range, 1, e->vis);
    }
    for (int k = 0; k < e->n_guards; ++k) {
        OGuard *g = &e->guards[k];
        int gr = g->path[2 * g->idx], gc = g->path[2 * g->idx + 1];
        cone(e, gr, gc, g->fov, g->heading, g->range, 0, e->vis);
        e->vis[gr * e->C + gc] = 1.0f; /* guard's own tile, visibility.py:59 */
    }
}

/* ------------------------------------------------------------------------------------ */
/* reset / step                                                                          */
/* ------------------------------------------------------------------------------------ */

/* HeistEnvironment.reset, environment.py:183-214: headings persist, guard idx -> 0 */
void oenv_reset(OEnv *e) {
    e->solver_r = e->start_r; e->solver_c = e->start_c;
    e->tick = 0; e->done = 0; e->detected = 0; e->vault_reached = 0;
    e->prev_dist = iabs(e->solver_r - e->vault_r) + iabs(e->solver_c - e->vault_c);
    e->init_dist = e->prev_dist;
    for (int k = 0; k < e->n_guards; ++k) e->guards[k].idx = 0;
    update_visibility(e);
}

/* Guard.update, security.py:145-159 */
static void guard_update(OGuard *g) {
    if (g->len < 2) return;
    int old = g->idx;
    int ni = (g->idx + g->speed) % g->len;
    if (ni < 0) ni += g->len; /* Python int % */
    g->idx = ni;
    int dr = g->path[2 * ni] - g->path[2 * old];
    int dc = g->path[2 * ni + 1] - g->path[2 * old + 1];
    if (dr != 0 || dc != 0) {
        int ndr = -dr; /* integer negation: -0 stays +0 before the float conversion */
        g->heading = py_fmod(py_degrees(atan2((double)ndr, (double)dc)), 360.0);
    }
}

/* HeistEnvironment.step, environment.py:216-299.  Returns status code. */
int oenv_step(OEnv *e, int action, double *reward_out, int *done_out) {
    static const int AR[5] = {0, -1, 1, 0, 0}, AC[5] = {0, 0, 0, -1, 1}; /* :52-58 */
    if (e->done) { *reward_out = 0.0; *done_out = 1; return ST_ALREADY_DONE; } /* :232-233 */
    double reward = e->reward_step; /* :235 */
    int status = ST_RUNNING;
    int nr = e->solver_r + AR[action], nc = e->solver_c + AC[action]; /* :239-246 */
    if (0 <= nr && nr < e->R && 0 <= nc && nc < e->C && e->grid[nr * e->C + nc] != T_WALL) {
        e->solver_r = nr; e->solver_c = nc;
    }
    for (int k = 0; k < e->n_cams; ++k) /* Camera.update, security.py:49-51 */
        e->cams[k].heading = py_fmod(e->cams[k].heading + e->cams[k].speed * 1, 360.0);
    for (int k = 0; k < e->n_guards; ++k) guard_update(&e->guards[k]);
    update_visibility(e); /* :257-258 */
    int curr = iabs(e->solver_r - e->vault_r) + iabs(e->solver_c - e->vault_c); /* :261 */
    reward += (double)(e->prev_dist - curr) * 0.1; /* :263-264 */
    e->prev_dist = curr;
    if (curr <= 3 && e->init_dist > 3) reward += 0.05 * (double)(3 - curr); /* :268-269 */
    if (e->vis[e->solver_r * e->C + e->solver_c] > 0.5f) { /* :273-281 */
        e->detected = 1; reward += e->reward_detection; e->done = 1; status = ST_DETECTED;
    }
    if (e->solver_r == e->vault_r && e->solver_c == e->vault_c) { /* :284-288 */
        e->vault_reached = 1; reward += e->reward_vault; e->done = 1; status = ST_VAULT;
    }
    e->tick += 1; /* :291-297 */
    if (e->tick >= e->max_steps) {
        e->done = 1; status = ST_TIMEOUT;
        int denom = e->init_dist > 1 ? e->init_dist : 1;
        double cf = 1.0 - (double)curr / (double)denom;
        if (!(cf > 0)) cf = 0.0; /* max(0, x) */
        reward += cf * 2.0;
    }
    *reward_out = reward; *done_out = e->done;
    return status;
}

/* ------------------------------------------------------------------------------------ */
/* Observations                                                                          */
/* ------------------------------------------------------------------------------------ */

/* get_state_tensor, environment.py:347-374 -> out[3][R][C] float32 (numpy>=2 scalar rules) */
void oenv_state_tensor(const OEnv *e, float *out) {
    int R = e->R, C = e->C, n = R * C;
    for (int i = 0; i < n; ++i) out[i] = (float)e->grid[i] / 5.0f; /* :319 */
    for (int i = 0; i < n; ++i) out[n + i] = e->vis[i];            /* :322 */
    float *pos = out + 2 * n;
    for (int i = 0; i < n; ++i) pos[i] = 0.0f;
    pos[e->solver_r * C + e->solver_c] = 1.0f;  /* :357 */
    pos[e->vault_r * C + e->vault_c] = -1.0f;   /* :358 */
    int max_d = R + C;
    for (int r = 0; r < R; ++r)
        for (int c = 0; c < C; ++c) { /* :361-365 */
            int d = iabs(r - e->vault_r) + iabs(c - e->vault_c);
            double v = -0.3 * ((double)d / (double)max_d);
            pos[r * C + c] = pos[r * C + c] + (float)v; /* np.float32 + python float -> float32 */
        }
}

/* _get_observation, environment.py:305-345: small vectors (grids come from state tensor ch0/ch1) */
void oenv_obs_vectors(const OEnv *e, float *solver_pos2, float *vault_dir2, float *time1) {
    solver_pos2[0] = (float)((double)e->solver_r / (double)e->R);
    solver_pos2[1] = (float)((double)e->solver_c / (double)e->C);
    vault_dir2[0] = (float)((double)(e->vault_r - e->solver_r) / (double)e->R);
    vault_dir2[1] = (float)((double)(e->vault_c - e->solver_c) / (double)e->C);
    time1[0] = (float)((double)e->tick / (double)e->max_steps);
}

/* ------------------------------------------------------------------------------------ */
/* Getters                                                                               */
/* ------------------------------------------------------------------------------------ */
void oenv_get_grid(const OEnv *e, int32_t *out) { memcpy(out, e->grid, sizeof(int32_t) * e->R * e->C); }
void oenv_get_vis(const OEnv *e, float *out) { memcpy(out, e->vis, sizeof(float) * e->R * e->C); }
/* info[0..9]: solver_r, solver_c, tick, done, detected, vault_reached, n_walls, n_cams, n_guards, spent */
void oenv_get_info(const OEnv *e, int *info) {
    info[0] = e->solver_r; info[1] = e->solver_c; info[2] = e->tick; info[3] = e->done;
    info[4] = e->detected; info[5] = e->vault_reached; info[6] = e->n_walls; info[7] = e->n_cams;
    info[8] = e->n_guards; info[9] = e->budget_spent;
}
void oenv_get_cam_headings(const OEnv *e, double *out) { for (int k = 0; k < e->n_cams; ++k) out[k] = e->cams[k].heading; }
/* per guard: out[3k] = row, out[3k+1] = col, out[3k+2] = idx ; head[k] = heading */
void oenv_get_guards(const OEnv *e, int *out, double *head) {
    for (int k = 0; k < e->n_guards; ++k) {
        const OGuard *g = &e->guards[k];
        out[3 * k] = g->path[2 * g->idx]; out[3 * k + 1] = g->path[2 * g->idx + 1]; out[3 * k + 2] = g->idx;
        head[k] = g->heading;
    }
}

/* ------------------------------------------------------------------------------------ */
/* Architect-side layout decode                                                          */
/* ------------------------------------------------------------------------------------ */

/* ArchitectNetwork._generate_patrol, networks.py:324-335 */
void oracle_generate_patrol(int row, int col, int H, int W, int *path16) {
    static const int OFF[8][2] = {{0, 0}, {0, 1}, {0, 2}, {1, 2}, {2, 2}, {2, 1}, {2, 0}, {1, 0}};
    for (int k = 0; k < 8; ++k) {
        int r = row + OFF[k][0] - 1, c = col + OFF[k][1] - 1;
        if (r > H - 2) r = H - 2;
        if (r < 1) r = 1;
        if (c > W - 2) c = W - 2;
        if (c < 1) c = 1;
        path16[2 * k] = r; path16[2 * k + 1] = c;
    }
}

/*
 * Decode loop of ArchitectNetwork.generate_layout, networks.py:273-322.
 * asset_map[H][W] in {0 none, 1 wall, 2 camera, 3 guard}.  Outputs (caller-sized H*W):
 *   wall_rc [2*n], cam_rc [2*n], guard_path [n][8][2].  counts[0..2] = n_walls, n_cams, n_guards,
 *   counts[3] = remaining budget.  All cameras share (fov, speed, heading), range 6;
 *   guards: speed 1, range 4, fov 90 (filled in by the caller).
 */
void oracle_decode_layout(const int8_t *asset_map, int H, int W, int budget, int *wall_rc, int *cam_rc,
                          int *guard_path, int *counts) {
    int nw = 0, ncam = 0, ng = 0, remaining = budget, stop = 0;
    for (int r = 1; r < H - 1 && !stop; ++r) {
        for (int c = 1; c < W - 1; ++c) {
            int t = asset_map[r * W + c];
            if (t == 0) continue;
            else if (t == 1 && remaining >= COST_WALL) { wall_rc[2 * nw] = r; wall_rc[2 * nw + 1] = c; nw++; remaining -= COST_WALL; }
            else if (t == 2 && remaining >= COST_CAMERA) { cam_rc[2 * ncam] = r; cam_rc[2 * ncam + 1] = c; ncam++; remaining -= COST_CAMERA; }
            else if (t == 3 && remaining >= COST_GUARD) { oracle_generate_patrol(r, c, H, W, guard_path + 16 * ng); ng++; remaining -= COST_GUARD; }
            if (remaining <= 0) { stop = 1; break; }
        }
    }
    counts[0] = nw; counts[1] = ncam; counts[2] = ng; counts[3] = remaining;
}

/* ------------------------------------------------------------------------------------ */
/* Solver rollout: GAE + returns (fp32, torch scalar-tensor semantics)                   */
/* ------------------------------------------------------------------------------------ */

/*
 * SolverAgent._compute_gae + returns, agents/solver.py:141-143, 228-244, applied to each of
 * n_cols independent time-major columns x[t*n_cols + j] (n_cols = 1 is the reference's
 * single flat buffer).  Python scalars gamma, gamma*lambda are cast to fp32 by torch; every
 * op rounds to fp32 in the order written; the last element's next value is 0.
 */
void oracle_gae(const float *rew, const float *val, const float *done, int T, int n_cols, double gamma,
                double gae_lambda, float *adv, float *ret) {
    volatile float g = (float)gamma;
    volatile float gl = (float)(gamma * gae_lambda);
    for (int j = 0; j < n_cols; ++j) {
        float last = 0.0f;
        for (int t = T - 1; t >= 0; --t) {
            size_t i = (size_t)t * n_cols + j;
            float nv = (t == T - 1) ? 0.0f : val[i + n_cols];
            volatile float om = 1.0f - done[i];
            volatile float a = g * nv;
            volatile float b = a * om;
            volatile float c = rew[i] + b;
            volatile float delta = c - val[i];
            volatile float d1 = gl * om;
            volatile float d2 = d1 * last;
            volatile float A = delta + d2;
            last = A;
            adv[i] = A;
            volatile float rr = A + val[i];
            ret[i] = rr;
        }
    }
}

/* advantage normalisation, solver.py:146-147: (A - mean) / (std_unbiased + 1e-8), len > 1.
 * torch's reduction order is not sequential, so this is compared at 1e-5 relative, not bitwise. */
void oracle_normalize(const float *adv, int n, float *out) {
    if (n <= 1) { for (int i = 0; i < n; ++i) out[i] = adv[i]; return; }
    double s = 0; for (int i = 0; i < n; ++i) s += adv[i];
    double mean = s / n, ss = 0;
    for (int i = 0; i < n; ++i) { double d = adv[i] - mean; ss += d * d; }
    float fm = (float)mean, fs = (float)sqrt(ss / (n - 1));
    for (int i = 0; i < n; ++i) out[i] = (adv[i] - fm) / (fs + 1e-8f);
}

/* RewardCalculator.calculate_architect_reward, rewards.py:43-73 */
double oracle_architect_reward(int valid, double solve_rate) {
    if (!valid) return -1.0;
    double reward = 0.0;
    double detection_rate = 1.0 - solve_rate;
    reward += detection_rate * 1.0;
    if (solve_rate > 0.8) reward += -0.5;
    if (0.2 <= solve_rate && solve_rate <= 0.6) reward += 0.2;
    return reward;
}

/* ------------------------------------------------------------------------------------ */
/* Batch drivers (bench cpu_baseline / --impl reference; parity tests at size)           */
/* ------------------------------------------------------------------------------------ */

int oracle_num_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

typedef struct {
    OEnv **envs; int n; const int8_t *actions; int T; int autoreset;
    float *reward32; double *reward64; uint8_t *done; uint8_t *status; uint32_t *vis_bits;
    int *next; /* shared work counter (dynamic schedule, one env at a time) */
    long live; int mode; /* mode 0 = rollout, 1 = reset */
} Job;

static void rollout_one(Job *jb, int j, long *live) {
    OEnv *e = jb->envs[j];
    int n = jb->n, W = (e->C + 31) / 32;
    for (int t = 0; t < jb->T; ++t) {
        double rw; int dn;
        size_t i = (size_t)t * n + j;
        int st = oenv_step(e, jb->actions[i], &rw, &dn);
        if (st != ST_ALREADY_DONE) (*live)++;
        if (jb->reward32) jb->reward32[i] = (float)rw;
        if (jb->reward64) jb->reward64[i] = rw;
        if (jb->done) jb->done[i] = (uint8_t)dn;
        if (jb->status) jb->status[i] = (uint8_t)st;
        if (jb->autoreset && dn) oenv_reset(e);
        if (jb->vis_bits) {
            uint32_t *vb = jb->vis_bits + i * (size_t)e->R * W;
            for (int r = 0; r < e->R; ++r)
                for (int w = 0; w < W; ++w) {
                    uint32_t m = 0;
                    for (int b = 0; b < 32 && w * 32 + b < e->C; ++b)
                        if (e->vis[r * e->C + w * 32 + b] > 0.5f) m |= (1u << b);
                    vb[r * W + w] = m;
                }
        }
    }
}

static void *worker(void *arg) {
    Job *jb = (Job *)arg;
    long live = 0;
    for (;;) {
        int j = __atomic_fetch_add(jb->next, 1, __ATOMIC_RELAXED);
        if (j >= jb->n) break;
        if (jb->mode == 0) rollout_one(jb, j, &live);
        else oenv_reset(jb->envs[j]);
    }
    __atomic_fetch_add(&jb->live, live, __ATOMIC_RELAXED);
    return NULL;
}

static long run_job(Job *jb, int n_threads) {
    int next = 0;
    jb->next = &next; jb->live = 0;
    if (n_threads <= 0) n_threads = oracle_num_threads();
    if (n_threads > jb->n) n_threads = jb->n > 0 ? jb->n : 1;
    if (n_threads <= 1) { worker(jb); return jb->live; }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
    for (int k = 0; k < n_threads; ++k) pthread_create(&th[k], NULL, worker, jb);
    for (int k = 0; k < n_threads; ++k) pthread_join(th[k], NULL);
    free(th);
    return jb->live;
}

/*
 * Roll n envs for T steps on actions[t*n + j] (time-major) with optional auto-reset
 * (step; if done: reset -- the trainer's pattern, training.py:515-533).  Outputs are
 * time-major and may be NULL.  vis_bits: [T][n][R][W] uint32 row bitmaps (bit c = col c)
 * of the visibility map *after* each step (after the auto-reset when one happened).
 * n_threads <= 0: all online cores.  Returns the number of env steps executed on live envs.
 */
long oracle_rollout(OEnv **envs, int n, const int8_t *actions, int T, int autoreset, float *reward32,
                    double *reward64, uint8_t *done, uint8_t *status, uint32_t *vis_bits, int n_threads) {
    Job jb;
    memset(&jb, 0, sizeof(jb));
    jb.envs = envs; jb.n = n; jb.actions = actions; jb.T = T; jb.autoreset = autoreset;
    jb.reward32 = reward32; jb.reward64 = reward64; jb.done = done; jb.status = status;
    jb.vis_bits = vis_bits; jb.mode = 0;
    return run_job(&jb, n_threads);
}

void oracle_reset_all(OEnv **envs, int n, int n_threads) {
    Job jb;
    memset(&jb, 0, sizeof(jb));
    jb.envs = envs; jb.n = n; jb.mode = 1;
    run_job(&jb, n_threads);
}
