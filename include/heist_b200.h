/*
 * heist_b200.h -- C ABI of the B200-native batched Heist Architect environment hot path.
 *
 * The reference (Shanmuk4622/RL-Project-Heist-Architect-...-CSE4019) is pure Python and has no
 * FFI layer; its boundary for this path is the Python class HeistEnvironment
 * (heist_architect/environment.py:40-426) plus two pieces the trainer runs next to it: the
 * Architect's asset-map decode (heist_architect/networks.py:273-335) and the Solver's GAE scan
 * (heist_architect/agents/solver.py:228-244).  Each entry point below names the reference
 * function it replaces.  INTEGRATION.md shows the ctypes stub a maintainer of the reference
 * would add to bind them.
 *
 * Conventions
 *   - every function returns int: 0 = OK, <0 = argument/capacity error, >0 = cudaError_t;
 *     heist_last_error() returns a thread-local message for the last non-zero return.
 *   - all array arguments are CALLER-OWNED DEVICE pointers on the handle's device, valid
 *     until the work enqueued on `stream` has finished.  The library allocates only the
 *     per-env state owned by the handle.  `stream` is a cudaStream_t passed as void*
 *     (NULL = legacy default stream).  No call synchronises the host except
 *     heist_create, heist_destroy and heist_check_errors.
 *   - a handle is not thread-safe; use one handle per (process, device).
 *   - N = num_envs, R = grid_rows, C = grid_cols (R, C <= HEIST_MAX_DIM),
 *     W = (C + 31) / 32 words per bitmap row; bit (c % 32) of word (c / 32) is column c.
 */
#ifndef HEIST_B200_H
#define HEIST_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HEIST_MAX_DIM 64
#define HEIST_ABI_VERSION 2

/* info["status"] of HeistEnvironment.step, environment.py:233,236,281,288,294 */
enum HeistStatus {
    HEIST_RUNNING = 0,
    HEIST_DETECTED = 1,
    HEIST_VAULT_REACHED = 2,
    HEIST_TIMEOUT = 3,
    HEIST_ALREADY_DONE = 4
};

/* TileType, heist_architect/utils.py:31-37 */
enum HeistTile { HEIST_EMPTY = 0, HEIST_WALL = 1, HEIST_START = 2, HEIST_VAULT = 3, HEIST_CAMERA = 4, HEIST_GUARD = 5 };

/* EnvironmentConfig, environment.py:18-37 (+ per-env capacities of the batched state). */
typedef struct HeistParams {
    int32_t grid_rows, grid_cols;
    int32_t max_steps;
    int32_t start_row, start_col;
    int32_t vault_row, vault_col;
    int32_t architect_budget;      /* BudgetManager.total_budget default, budget.py:36 */
    int32_t max_walls;             /* capacity of the explicit layout lists, per env */
    int32_t max_cams;              /* K_c: cameras kept per env */
    int32_t max_guards;            /* K_g: guards kept per env */
    int32_t max_path;              /* waypoints per guard patrol (>= 8 for the Architect decode) */
    double reward_vault;           /* +10.0 */
    double reward_detection;       /* -1.0  */
    double reward_step;            /* -0.01 */
} HeistParams;

/*
 * Explicit layouts, the arguments of HeistEnvironment.set_layout (environment.py:102-113) for N
 * envs.  Lists longer than the capacities in HeistParams are a caller error.
 */
typedef struct HeistLayoutArrays {
    const int32_t *n_walls;      /* [N]                                             */
    const int16_t *wall_rc;      /* [N][max_walls][2]   (row, col)                  */
    const int32_t *n_cams;       /* [N]                                             */
    const int16_t *cam_rc;       /* [N][max_cams][2]                                */
    const double *cam_f;         /* [N][max_cams][3]    fov_angle, heading, rotation_speed */
    const int32_t *cam_range;    /* [N][max_cams]       vision_range                */
    const int32_t *n_guards;     /* [N]                                             */
    const int32_t *guard_len;    /* [N][max_guards]     waypoints in patrol_path (0 = skipped) */
    const int16_t *guard_path;   /* [N][max_guards][max_path][2]                    */
    const double *guard_head;    /* [N][max_guards][max_path] heading after leaving waypoint i:
                                    degrees(atan2(-dr, dc)) % 360 from the platform libm
                                    (security.py:159), NaN when the move is (0,0)   */
    const int32_t *guard_speed;  /* [N][max_guards]                                 */
    const int32_t *guard_range;  /* [N][max_guards]                                 */
    const double *guard_fov;     /* [N][max_guards]                                 */
} HeistLayoutArrays;

/* Device views of the handle's state (read-only for callers). */
typedef struct HeistStateView {
    const uint8_t *tile;          /* [N][R*C]   TileType codes (HeistEnvironment.grid)           */
    const uint32_t *wall_bits;    /* [N][R*W]   grid == WALL                                     */
    const uint32_t *vis_bits;     /* [N][R*W]   DynamicVisibilityMap.visibility > 0.5            */
    const int32_t *env_static;    /* [N][4]     n_cams, n_guards, valid, budget spent            */
    const int32_t *env_dyn;       /* [N][8]     row | col<<16, tick, prev_dist, init_dist,
                                                flags (bit0 done, bit1 detected, bit2 vault_reached,
                                                bits 8..15 status of the last step),
                                                episodes ended by vault / detection / timeout
                                                since the last set_layout                        */
    const double *cam_f;          /* [N][max_cams][2]   fov_angle, rotation_speed                */
    const int16_t *cam_i;         /* [N][max_cams][4]   row, col, vision_range, num_rays         */
    const double *cam_heading;    /* [N][max_cams]                                               */
    const double *guard_fov;      /* [N][max_guards]                                             */
    const int32_t *guard_i;       /* [N][max_guards][4] path length, speed, vision_range, num_rays */
    const uint8_t *guard_path;    /* [N][max_guards][max_path][2]                                */
    const double *guard_heading;  /* [N][max_guards]                                             */
    const int32_t *guard_idx;     /* [N][max_guards]    Guard.current_idx                        */
    const uint8_t *wall_accepted; /* [N][max_walls]     1: the i-th wall of the last explicit set_layout request was
                                                        placed and paid for, i.e. is in HeistEnvironment.walls
                                                        (environment.py:118-121)                                  */
} HeistStateView;

typedef struct HeistHandle HeistHandle;

int heist_abi_version(void);
const char *heist_last_error(void);
/* Message of the last condition that is not an error but costs performance (thread-local; "" if none): set by
 * heist_create when the angular visibility cache could not be allocated and every env falls back to the ray-march
 * kernels.  HEIST_REQUIRE_VIS_CACHE=1 in the environment turns that condition into an error (-11). */
const char *heist_last_warning(void);

/* HeistEnvironment.__init__ (environment.py:62-96) for num_envs independent envs on `device`. */
int heist_create(const HeistParams *params, int num_envs, int device, HeistHandle **out);
int heist_destroy(HeistHandle *h);

/*
 * ArchitectNetwork.generate_layout decode loop + _generate_patrol (networks.py:273-335), the
 * curriculum filter (training.py:464-467), then HeistEnvironment.set_layout incl. BFS validity
 * (environment.py:102-158; utils.py:52-85), for every env.
 *   asset_map   [N][R][C] int8 in {0 none, 1 wall, 2 camera, 3 guard}
 *   cam_params  [N][3] float32: fov, speed, heading (the order of the network's dict)
 *   budget      [N] int32 or NULL (NULL: HeistParams.architect_budget); used for both the decode
 *               (architect.budget) and the env (env.budget.scale_budget), training.py:444-445
 *   valid_out   [N] uint8 (may be NULL)
 */
int heist_decode_validate(HeistHandle *h, const int8_t *asset_map, const float *cam_params, const int32_t *budget,
                          int allow_cameras, int allow_guards, uint8_t *valid_out, void *stream);

/* HeistEnvironment.set_layout (environment.py:102-152) on explicit lists. budget as above. */
int heist_set_layout_explicit(HeistHandle *h, const HeistLayoutArrays *layout, const int32_t *budget,
                              uint8_t *valid_out, void *stream);

/* HeistEnvironment.reset (environment.py:183-214) for envs with mask[i] != 0 (NULL: all). */
int heist_reset(HeistHandle *h, const uint8_t *mask, void *stream);

/*
 * HeistEnvironment.step (environment.py:216-299) for all envs.
 *   actions [N] int8 (0 WAIT, 1 UP, 2 DOWN, 3 LEFT, 4 RIGHT; anything else acts as WAIT)
 *   reward [N] float32, reward64 [N] float64 (either may be NULL), done [N] uint8, status [N] uint8
 */
int heist_step(HeistHandle *h, const int8_t *actions, float *reward, double *reward64, uint8_t *done,
               uint8_t *status, void *stream);

/*
 * T consecutive steps in one launch (the trainer's inner loop, training.py:515-533, without the
 * policy): time-major actions [T][N]; outputs [T][N] (any may be NULL).  autoreset != 0 applies
 * reset() to an env right after a step that ended its episode.  vis_traj, if not NULL, receives
 * the visibility bitmap after every step (after the auto-reset when one happened): [T][N][R*W].
 */
int heist_step_many(HeistHandle *h, const int8_t *actions, int T, int autoreset, float *reward, uint8_t *done,
                    uint8_t *status, uint32_t *vis_traj, void *stream);

/*
 * heist_step_many for a driver whose rollout buffers live on the host (pinned memory): actions_host [T][N] int8
 * in, reward_host [T][N] float32 / done_host / status_host [T][N] uint8 out (any output may be NULL); vis_traj stays
 * a device pointer (may be NULL).  With auto-reset and T > 32 the copies are issued chunk by chunk on internal
 * streams next to the kernels of the pipelined launch and joined onto `stream`; otherwise they bracket the launch
 * on `stream`.  The host buffers are valid once `stream` has passed the call.
 */
int heist_step_many_host(HeistHandle *h, const int8_t *actions_host, int T, int autoreset, float *reward_host,
                         uint8_t *done_host, uint8_t *status_host, uint32_t *vis_traj, void *stream);

/* HeistEnvironment.get_state_tensor (environment.py:347-374): state [N][3][R][C] float32. */
int heist_observe(HeistHandle *h, float *state, void *stream);

/*
 * One tick plus the dense state the policy needs next, in one call: HeistEnvironment.step followed by
 * get_state_tensor (the trainer's inner loop, training.py:523-529), with the optional `if done: reset()`
 * applied before the state is built.  state [N][3][R][C] float32.
 */
int heist_step_observe(HeistHandle *h, const int8_t *actions, int autoreset, float *reward, uint8_t *done,
                       uint8_t *status, float *state, void *stream);

/*
 * Dense states for a PPO minibatch from a PACKED rollout buffer (agents/solver.py:134,165 keep and gather dense
 * (3,R,C) float32 states; here a transition is its visibility bitmap + solver position).  For m < M:
 * vis_bits [M][R*W], pos [M] = row | col << 16, env_idx [M] = env whose layout (tile codes) applies
 * -> state [M][3][R][C], identical to what heist_observe produced when the transition was recorded.
 */
int heist_expand_states(HeistHandle *h, const uint32_t *vis_bits, const int32_t *pos, const int32_t *env_idx, int M,
                        float *state, void *stream);

/*
 * HeistEnvironment._get_observation small vectors (environment.py:324-337):
 * obs_vec [N][5] float32 = solver_position(2), vault_direction(2), time_feature(1).
 * occupancy_grid and visibility_map are channels 0 and 1 of heist_observe.
 */
int heist_observation_vectors(HeistHandle *h, float *obs_vec, void *stream);

int heist_get_state(HeistHandle *h, HeistStateView *view);

/*
 * SolverAgent._compute_gae + returns (agents/solver.py:141-143, 228-244) on n_cols independent
 * time-major columns: rew, val [T][n_cols] float32, done [T][n_cols] uint8 -> adv, ret [T][n_cols].
 * gamma and gamma*lambda are rounded to float32 first, as torch does with Python scalars.
 */
int heist_gae(const float *rew, const float *val, const uint8_t *done, int T, int n_cols, double gamma,
              double gae_lambda, float *adv, float *ret, int device, void *stream);

/*
 * RewardCalculator.calculate_architect_reward (rewards.py:43-73) with the trainer's outcome
 * counting (training.py:535-550) per env: solve_rate = vault episodes / finished episodes since
 * the last set_layout; invalid layouts get -1.  reward_out [N] float64, solve_rate_out [N] float64
 * (may be NULL).
 */
int heist_architect_reward(HeistHandle *h, double *reward_out, double *solve_rate_out, void *stream);

/*
 * Verification / deployment knob.  All modes produce bit-identical results; tests compare them.
 *   HEIST_MODE_DEFAULT  visibility from the per-layout angular cache (table-driven kernel); envs whose assets the
 *                       cache does not cover (vision_range > 7, fov outside (0, 180], ...) are ray-marched
 *   HEIST_MODE_EXACT    every ray sample through the fp64 reference arithmetic (security.py:69-99)
 *   HEIST_MODE_MARCH    filtered fixed-point ray-march with exact fallback for every env (no cache)
 * Setting HEIST_NO_VIS_CACHE=1 in the environment before heist_create makes DEFAULT behave like MARCH.
 *   HEIST_MODE_TABLES   like DEFAULT, but the caller GUARANTEES that every env's assets are covered by the cache (true for
 *                       every layout the Architect decode produces: vision_range 6, fov 30..120, <= 4 guards): launches
 *                       then never carry the ray-march kernels -- a single tick is exactly one kernel, which is what a
 *                       CUDA graph of a policy loop wants.  A set_layout that breaks the guarantee raises the sticky
 *                       device error "layout not covered by the visibility cache" (heist_check_errors); the uncovered
 *                       envs do not advance.
 */
#define HEIST_MODE_DEFAULT 0
#define HEIST_MODE_EXACT 1
#define HEIST_MODE_MARCH 2
#define HEIST_MODE_TABLES 3
int heist_set_mode(HeistHandle *h, int mode);

/*
 * Coverage of the angular visibility cache after the last set_layout: how many envs every asset of which is
 * served from the cache (the rest are ray-marched), and the cache's device memory in bytes (0: disabled).
 * Synchronises `stream`.  Any of the outputs may be NULL.
 */
int heist_cache_stats(HeistHandle *h, int32_t *envs_cached, int64_t *cache_bytes, void *stream);

/* Kernels this handle has launched for reset / step / step_many / step_observe since heist_create (bench accounting). */
int heist_launch_count(HeistHandle *h, int64_t *count);

/*
 * Verification hook: the ray directions the kernels use, dx = cos(radians(a)), dy = -sin(radians(a))
 * (security.py:71-75), for n angles in degrees (device pointers; a handle must exist on `device`).  Angles within
 * 1e-11 degree of a multiple of 30 (|a| <= 1440) come from a table evaluated by the host libm at heist_create.
 */
int heist_debug_ray_dirs(int device, const double *angles_deg, int n, double *dx, double *dy, void *stream);

/* Synchronises `stream` and reports sticky device-side errors (capacity overflow, bad waypoint). */
int heist_check_errors(HeistHandle *h, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* HEIST_B200_H */
