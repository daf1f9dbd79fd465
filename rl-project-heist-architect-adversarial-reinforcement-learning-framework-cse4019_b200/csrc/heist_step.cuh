// heist_step.cuh -- per-step dynamics: reset, step, step_many (one warp per env).
//
// Reference: HeistEnvironment.reset / step (environment.py:183-299), Camera.update and
// get_vision_cone_tiles (security.py:49-101), Guard.update and get_visible_tiles
// (security.py:145-192), DynamicVisibilityMap.update (visibility.py:31-65).
#pragma once
#include "heist_common.cuh"

// Per-warp shared-memory working set of one env.
struct WarpEnv {
    uint32_t *wall;      // [RW] grid == WALL row bitmaps
    uint32_t *vis;       // [RW] visibility row bitmaps (rebuilt every tick)
    double *cam_fov;     // [Kc]
    double *cam_speed;   // [Kc]
    double *cam_head;    // [Kc]
    int4 *cam_i;         // [Kc] row, col, range, num_rays
    double *g_fov;       // [Kg]
    double *g_head;      // [Kg]
    int4 *g_i;           // [Kg] len, speed, range, num_rays
    int *g_idx;          // [Kg]
    int2 *g_pos;         // [Kg] current (row, col)
};

__host__ __device__ inline size_t warp_env_bytes(int RW, int Kc, int Kg) {
    size_t b = 0;
    b += (size_t)Kc * (3 * sizeof(double) + sizeof(int4));
    b += (size_t)Kg * (2 * sizeof(double) + sizeof(int4) + sizeof(int2) + sizeof(int));
    b += (size_t)2 * RW * sizeof(uint32_t);
    return (b + 15) & ~(size_t)15;
}

__device__ __forceinline__ WarpEnv carve_warp_env(unsigned char *base, int RW, int Kc, int Kg) {
    WarpEnv S;
    unsigned char *p = base;  // 16-byte aligned
    S.cam_i = (int4 *)p;      p += (size_t)Kc * sizeof(int4);
    S.g_i = (int4 *)p;        p += (size_t)Kg * sizeof(int4);
    S.cam_fov = (double *)p;  p += (size_t)Kc * sizeof(double);
    S.cam_speed = (double *)p;p += (size_t)Kc * sizeof(double);
    S.cam_head = (double *)p; p += (size_t)Kc * sizeof(double);
    S.g_fov = (double *)p;    p += (size_t)Kg * sizeof(double);
    S.g_head = (double *)p;   p += (size_t)Kg * sizeof(double);
    S.g_pos = (int2 *)p;      p += (size_t)Kg * sizeof(int2);
    S.g_idx = (int *)p;       p += (size_t)Kg * sizeof(int);
    S.wall = (uint32_t *)p;   p += (size_t)RW * sizeof(uint32_t);
    S.vis = (uint32_t *)p;
    return S;
}

struct EnvRegs {
    int r, c, tick, prev, init, flags, n_vault, n_detect, n_timeout;
};

// One vision cone: rays i = lane, lane+32, ... <= num_rays; samples dist = unit*j, j = 1..nsamp.
// Cameras: unit 0.5, nsamp 2*range (the reference's sub-steps 0/.5/1 repeat integer distances,
// which is idempotent).  Guards: unit 1, nsamp = range.  First out-of-bounds or WALL sample ends a ray.
__device__ __forceinline__ void cone_march(const Dev &D, const WarpEnv &S, int lane, int row, int col, double fov,
                                           double heading, int num_rays, int nsamp, double unit) {
    const double half_fov = __ddiv_rn(fov, 2.0);
    const double base = __dsub_rn(heading, half_fov);
    const double drow = (double)row, dcol = (double)col, dn = (double)num_rays;
    for (int i = lane; i <= num_rays; i += 32) {
        double angle_deg = __dadd_rn(base, __ddiv_rn(__dmul_rn(fov, (double)i), dn));
        double dx, dy;
        ray_dir(angle_deg, D.deg2rad, dx, dy);
        for (int j = 1; j <= nsamp; ++j) {
            double dist = unit * (double)j;  // exact
            double fx = __dadd_rn(dcol, __dmul_rn(dx, dist));
            double fy = __dadd_rn(drow, __dmul_rn(dy, dist));
            int c = __double2int_rn(fx);  // round half to even, like round()
            int r = __double2int_rn(fy);
            if ((unsigned)r >= (unsigned)D.R || (unsigned)c >= (unsigned)D.C) break;
            int word = r * D.W + (c >> 5);
            uint32_t bit = 1u << (c & 31);
            if (S.wall[word] & bit) break;
            if (r != row || c != col) atomicOr(&S.vis[word], bit);
        }
    }
}

// DynamicVisibilityMap.update (visibility.py:31-65)
__device__ __forceinline__ void compute_visibility(const Dev &D, const WarpEnv &S, int lane, int n_cams, int n_guards) {
    for (int i = lane; i < D.RW; i += 32) S.vis[i] = 0u;
    __syncwarp();
    for (int k = 0; k < n_cams; ++k) {
        int4 ci = S.cam_i[k];
        cone_march(D, S, lane, ci.x, ci.y, S.cam_fov[k], S.cam_head[k], ci.w, 2 * ci.z, 0.5);
    }
    for (int k = 0; k < n_guards; ++k) {
        int4 gi = S.g_i[k];
        int2 gp = S.g_pos[k];
        cone_march(D, S, lane, gp.x, gp.y, S.g_fov[k], S.g_head[k], gi.w, gi.z, 1.0);
        if (lane == 0) atomicOr(&S.vis[gp.x * D.W + (gp.y >> 5)], 1u << (gp.y & 31));  // guard's own tile
    }
    __syncwarp();
}

__device__ __forceinline__ void load_env(const Dev &D, const WarpEnv &S, int env, int lane, EnvRegs &E, int &n_cams,
                                         int &n_guards) {
    const int4 es = *reinterpret_cast<const int4 *>(D.env_s + (size_t)env * 4);
    n_cams = es.x;
    n_guards = es.y;
    const int4 d0 = *reinterpret_cast<const int4 *>(D.env_d + (size_t)env * 8);
    const int4 d1 = *reinterpret_cast<const int4 *>(D.env_d + (size_t)env * 8 + 4);
    E.r = d0.x & 0xffff; E.c = d0.x >> 16; E.tick = d0.y; E.prev = d0.z; E.init = d0.w;
    E.flags = d1.x & 0xff; E.n_vault = d1.y; E.n_detect = d1.z; E.n_timeout = d1.w;
    for (int i = lane; i < D.RW; i += 32) S.wall[i] = D.wall[(size_t)env * D.RW + i];
    if (lane < n_cams) {
        size_t o = (size_t)env * D.Kc + lane;
        S.cam_fov[lane] = D.cam_f[o * 2];
        S.cam_speed[lane] = D.cam_f[o * 2 + 1];
        S.cam_head[lane] = D.cam_heading[o];
        const int16_t *ci = D.cam_i + o * 4;
        S.cam_i[lane] = make_int4(ci[0], ci[1], ci[2], ci[3]);
    }
    if (lane < n_guards) {
        size_t o = (size_t)env * D.Kg + lane;
        S.g_fov[lane] = D.guard_fov[o];
        S.g_head[lane] = D.guard_heading[o];
        S.g_i[lane] = *reinterpret_cast<const int4 *>(D.guard_i + o * 4);
        int idx = D.guard_idx[o];
        S.g_idx[lane] = idx;
        const uint8_t *p = D.guard_path + (o * D.L + idx) * 2;
        S.g_pos[lane] = make_int2(p[0], p[1]);
    }
    __syncwarp();
}

__device__ __forceinline__ void store_env(const Dev &D, const WarpEnv &S, int env, int lane, const EnvRegs &E, int status,
                                          int n_cams, int n_guards) {
    if (lane == 0) {
        int4 d0 = make_int4(E.r | (E.c << 16), E.tick, E.prev, E.init);
        int4 d1 = make_int4(E.flags | (status << 8), E.n_vault, E.n_detect, E.n_timeout);
        *reinterpret_cast<int4 *>(D.env_d + (size_t)env * 8) = d0;
        *reinterpret_cast<int4 *>(D.env_d + (size_t)env * 8 + 4) = d1;
    }
    if (lane < n_cams) D.cam_heading[(size_t)env * D.Kc + lane] = S.cam_head[lane];
    if (lane < n_guards) {
        size_t o = (size_t)env * D.Kg + lane;
        D.guard_heading[o] = S.g_head[lane];
        D.guard_idx[o] = S.g_idx[lane];
    }
    for (int i = lane; i < D.RW; i += 32) D.vis[(size_t)env * D.RW + i] = S.vis[i];
}

// HeistEnvironment.reset (environment.py:183-214): camera and guard headings persist.
__device__ __forceinline__ void reset_env(const Dev &D, const WarpEnv &S, int env, int lane, EnvRegs &E, int n_cams,
                                          int n_guards) {
    E.r = D.start_r; E.c = D.start_c; E.tick = 0;
    E.flags = 0;
    E.prev = abs(E.r - D.vault_r) + abs(E.c - D.vault_c);
    E.init = E.prev;
    if (lane < n_guards) {
        S.g_idx[lane] = 0;
        const uint8_t *p = D.guard_path + ((size_t)env * D.Kg + lane) * D.L * 2;
        S.g_pos[lane] = make_int2(p[0], p[1]);
    }
    __syncwarp();
    compute_visibility(D, S, lane, n_cams, n_guards);
}

// HeistEnvironment.step (environment.py:216-299).  All lanes carry the scalar env state redundantly.
__device__ __forceinline__ int step_env(const Dev &D, const WarpEnv &S, int env, int lane, EnvRegs &E, int n_cams,
                                        int n_guards, int action, double &reward_out) {
    if (E.flags & F_DONE) { reward_out = 0.0; return HEIST_ALREADY_DONE; }  // :232-233
    double reward = D.reward_step;                                           // :235
    int status = HEIST_RUNNING;
    // 1. move (:239-246) -- blocked only by out-of-bounds or WALL
    int nr = E.r + (action == 2) - (action == 1);
    int nc = E.c + (action == 4) - (action == 3);
    if ((unsigned)nr < (unsigned)D.R && (unsigned)nc < (unsigned)D.C &&
        !((S.wall[nr * D.W + (nc >> 5)] >> (nc & 31)) & 1u)) { E.r = nr; E.c = nc; }
    // 2. cameras rotate (security.py:49-51), guards advance (security.py:145-159)
    if (lane < n_cams) S.cam_head[lane] = py_mod360(__dadd_rn(S.cam_head[lane], S.cam_speed[lane]));
    if (lane < n_guards) {
        int4 gi = S.g_i[lane];
        if (gi.x >= 2) {
            int old = S.g_idx[lane];
            int ni = py_imod(old + gi.y, gi.x);
            size_t o = ((size_t)env * D.Kg + lane) * D.L;
            double h = D.guard_head[o + old];
            if (h == h) S.g_head[lane] = h;  // NaN <=> the move is (0,0): heading unchanged
            S.g_idx[lane] = ni;
            const uint8_t *p = D.guard_path + (o + ni) * 2;
            S.g_pos[lane] = make_int2(p[0], p[1]);
        }
    }
    __syncwarp();
    // 3. visibility (:257-258)
    compute_visibility(D, S, lane, n_cams, n_guards);
    // 4. shaping (:261-269)
    int curr = abs(E.r - D.vault_r) + abs(E.c - D.vault_c);
    reward = __dadd_rn(reward, __dmul_rn((double)(E.prev - curr), 0.1));
    E.prev = curr;
    if (curr <= 3 && E.init > 3) reward = __dadd_rn(reward, __dmul_rn(0.05, (double)(3 - curr)));
    // 5. detection (:273-281), vault (:284-288), timeout (:291-297)
    if ((S.vis[E.r * D.W + (E.c >> 5)] >> (E.c & 31)) & 1u) {
        E.flags |= F_DETECTED | F_DONE;
        reward = __dadd_rn(reward, D.reward_detection);
        status = HEIST_DETECTED;
    }
    if (E.r == D.vault_r && E.c == D.vault_c) {
        E.flags |= F_VAULT | F_DONE;
        reward = __dadd_rn(reward, D.reward_vault);
        status = HEIST_VAULT_REACHED;
    }
    E.tick += 1;
    if (E.tick >= D.max_steps) {
        E.flags |= F_DONE;
        status = HEIST_TIMEOUT;
        double cf = __dsub_rn(1.0, __ddiv_rn((double)curr, (double)max(E.init, 1)));
        if (!(cf > 0.0)) cf = 0.0;
        reward = __dadd_rn(reward, __dmul_rn(cf, 2.0));
    }
    if (status == HEIST_VAULT_REACHED) E.n_vault++;          // trainer's outcome counting,
    else if (status == HEIST_DETECTED) E.n_detect++;         // training.py:535-540
    else if (status == HEIST_TIMEOUT) E.n_timeout++;
    reward_out = reward;
    return status;
}

// T steps per launch; T = 1 with vis_traj = NULL is HeistEnvironment.step for the batch.
__global__ void __launch_bounds__(HEIST_WARPS_PER_CTA * 32)
k_step_many(Dev D, const int8_t *__restrict__ actions, int T, int autoreset, float *__restrict__ reward,
            double *__restrict__ reward64, uint8_t *__restrict__ done, uint8_t *__restrict__ status_out,
            uint32_t *__restrict__ vis_traj) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int env = blockIdx.x * HEIST_WARPS_PER_CTA + warp;
    if (env >= D.N) return;
    WarpEnv S = carve_warp_env(smem + (size_t)warp * warp_env_bytes(D.RW, D.Kc, D.Kg), D.RW, D.Kc, D.Kg);
    EnvRegs E;
    int n_cams, n_guards;
    load_env(D, S, env, lane, E, n_cams, n_guards);
    // the visibility map is state too (a done env keeps it): start from the stored one
    for (int i = lane; i < D.RW; i += 32) S.vis[i] = D.vis[(size_t)env * D.RW + i];
    __syncwarp();
    int status = HEIST_RUNNING;
    for (int t = 0; t < T; ++t) {
        const size_t o = (size_t)t * D.N + env;
        int action = actions[o];
        double rw;
        status = step_env(D, S, env, lane, E, n_cams, n_guards, action, rw);
        if (lane == 0) {
            if (reward) reward[o] = (float)rw;
            if (reward64) reward64[o] = rw;
            if (done) done[o] = (E.flags & F_DONE) ? 1 : 0;
            if (status_out) status_out[o] = (uint8_t)status;
        }
        if (autoreset && (E.flags & F_DONE)) reset_env(D, S, env, lane, E, n_cams, n_guards);
        if (vis_traj) {
            uint32_t *vt = vis_traj + o * D.RW;
            for (int i = lane; i < D.RW; i += 32) vt[i] = S.vis[i];
        }
    }
    store_env(D, S, env, lane, E, status, n_cams, n_guards);
}

__global__ void __launch_bounds__(HEIST_WARPS_PER_CTA * 32)
k_reset(Dev D, const uint8_t *__restrict__ mask) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int env = blockIdx.x * HEIST_WARPS_PER_CTA + warp;
    if (env >= D.N) return;
    if (mask && !mask[env]) return;
    WarpEnv S = carve_warp_env(smem + (size_t)warp * warp_env_bytes(D.RW, D.Kc, D.Kg), D.RW, D.Kc, D.Kg);
    EnvRegs E;
    int n_cams, n_guards;
    load_env(D, S, env, lane, E, n_cams, n_guards);
    reset_env(D, S, env, lane, E, n_cams, n_guards);
    store_env(D, S, env, lane, E, HEIST_RUNNING, n_cams, n_guards);
}
