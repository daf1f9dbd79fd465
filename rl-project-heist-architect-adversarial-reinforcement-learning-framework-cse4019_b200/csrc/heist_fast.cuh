// heist_fast.cuh -- table-driven reset / step / step_many for the envs the angular visibility cache covers.
//
// Reference: HeistEnvironment.reset / step (environment.py:183-299), Camera.update (security.py:49-51),
// Guard.update (security.py:145-159), DynamicVisibilityMap.update (visibility.py:31-65).
//
// Same semantics and the same persistent state arrays as heist_step.cuh (the two paths can be mixed freely on
// one handle).  The observation that shapes this file: cameras never react to the Solver -- a camera's heading at
// tick t is `heading0` advanced a known number of times (environment.py:251-252), whatever the agent does and
// whenever episodes end (reset() keeps headings, environment.py:205-208).  So a launch of T ticks splits into
//   k_heads    thread per camera: the heading at the start of every block of FAST_TB ticks (sequential fp64
//              Python-float modulo, T steps, negligible);
//   k_cam_vis  warp per (env, tick block): union of the camera cones of every tick of the block from the cache
//              (heist_cache.cuh) -> cam_vis[t][env] row bitmaps.  Embarrassingly parallel over env x time: no
//              sequential dependence, no load imbalance between envs, latency hidden by occupancy;
//   k_dyn      warp per env, lane = grid row, sequential in t: move, guard patrol + guard cone (one cached mask
//              per (waypoint, heading)), cam_vis[t] OR guards -> visibility map, detection / vault / timeout,
//              rewards, auto-reset.  HBM-streaming: reads cam_vis + actions, writes the visibility trajectory
//              and reward / done / status.
// Rays that fall inside a tie band (or outside the cached angle domain) are marched exactly like the reference
// does, so results are bit-identical to heist_step.cuh's.
#pragma once
#include "heist_cache.cuh"
#include "heist_step.cuh"

#define FAST_WARPS 4
#define FAST_TB 8      // ticks per k_cam_vis warp

struct FastCam {
    double heading, speed, fov, inv_step, dom_lo;
    const double *P;        // boundary points
    const uint16_t *MK;     // gap masks
    int row, col, range, num_rays, n_points, s0, carry, pad;
};
struct FastGuard {
    double heading, fov;
    int len, speed, range, num_rays, idx, hslot, nh, pad;
};

__host__ __device__ inline size_t camvis_warp_bytes(int RW, int Kc) {
    return (size_t)Kc * sizeof(FastCam) + (((size_t)RW * 4 + 15) & ~(size_t)15) + 32;
}
__host__ __device__ inline size_t dyn_warp_bytes(int RW, int Kg) {
    return (size_t)Kg * sizeof(FastGuard) + (((size_t)RW * 4 + 15) & ~(size_t)15);
}

// bit (r, c) of a lane-per-row bitmap (all lanes get the answer; r, c warp-uniform)
template <int RPL, int W>
__device__ __forceinline__ unsigned fast_bit(const uint32_t (&m)[RPL][W], int r, int c) {
    uint32_t w = m[0][0];
    if (W == 2 && (c >> 5)) w = m[0][W - 1];
    if (RPL == 2) {
        uint32_t w1 = m[RPL - 1][0];
        if (W == 2 && (c >> 5)) w1 = m[RPL - 1][W - 1];
        if (r >> 5) w = w1;
    }
    return (__shfl_sync(0xffffffffu, w, r & 31) >> (c & 31)) & 1u;
}

// OR a 16-bit window row (bit i = column col0 + i) into lane-row words
template <int W>
__device__ __forceinline__ void fast_or_row(uint32_t (&v)[W], unsigned bits, int col0) {
    const unsigned long long b = col0 >= 0 ? ((unsigned long long)bits << col0) : ((unsigned long long)bits >> (-col0));
    v[0] |= (uint32_t)b;
    if (W == 2) v[W - 1] |= (uint32_t)(b >> 32);
}

// adv0: camera updates that precede tick 0 of a launch.  A step launch updates the cameras once per tick
// (environment.py:251-252) -- except that an env which was already done when the launch began spends its first
// tick on the "already done" early-out (:232-233); a reset launch keeps the headings (:205-208).
__device__ __forceinline__ int fast_adv0(const Dev &D, int env, int do_reset) {
    if (do_reset) return 0;
    return (D.env_d[(size_t)env * 8 + 4] & F_DONE) ? 0 : 1;
}

// heads[b][env][k] = heading of camera k at tick b * FAST_TB of this launch
__global__ void __launch_bounds__(128) k_heads(Dev D, int T, int do_reset, double *__restrict__ heads) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= D.N * D.Kc) return;
    const int env = i / D.Kc, k = i - env * D.Kc;
    if (!D.env_cached[env] || k >= D.env_s[(size_t)env * 4]) return;
    double h = D.cam_heading[i];
    const double speed = D.cam_f[(size_t)i * 2 + 1];
    const int adv0 = fast_adv0(D, env, do_reset);
    for (int a = 0; a < adv0; ++a) h = py_mod360(__dadd_rn(h, speed));
    for (int t = 0; t < T; ++t) {
        if (t % FAST_TB == 0) heads[(size_t)(t / FAST_TB) * D.N * D.Kc + i] = h;
        h = py_mod360(__dadd_rn(h, speed));
    }
}

// rays below boundary point p: clamp(ceil((p - base) / step), 0, NR)
__device__ __forceinline__ int fast_nrays(double p, double base, double inv_step, int NR) {
    return max(0, min(NR, __double2int_ru((p - base) * inv_step)));
}

// Union of the camera cones of one env for FAST_TB consecutive ticks -> out[t][env][RW].
template <int RPL, int W>
__global__ void __launch_bounds__(FAST_WARPS * 32)
k_cam_vis(Dev D, int T, int nblk, const double *__restrict__ heads, uint32_t *__restrict__ out, const uint8_t *__restrict__ mask) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long wid = (long long)blockIdx.x * FAST_WARPS + warp;
    const int env = (int)(wid / nblk), b = (int)(wid - (long long)env * nblk);
    if (env >= D.N || !D.env_cached[env]) return;
    if (mask && !mask[env]) return;
    unsigned char *sp = smem + (size_t)warp * camvis_warp_bytes(D.RW, D.Kc);
    FastCam *cams = reinterpret_cast<FastCam *>(sp);        sp += (size_t)D.Kc * sizeof(FastCam);
    uint32_t *xvis = reinterpret_cast<uint32_t *>(sp);      sp += ((size_t)D.RW * 4 + 15) & ~(size_t)15;
    uint32_t *stage = reinterpret_cast<uint32_t *>(sp);
    const uint32_t *wall_g = D.wall + (size_t)env * D.RW;
    const int n_cams = D.env_s[(size_t)env * 4];
    const uint16_t *IX = nullptr;
    if (lane < n_cams) {
        const size_t o = (size_t)env * D.Kc + lane;
        FastCam &Cm = cams[lane];
        const int16_t *ci = D.cam_i + o * 4;
        Cm.fov = D.cam_f[o * 2]; Cm.speed = D.cam_f[o * 2 + 1];
        Cm.heading = heads[(size_t)b * D.N * D.Kc + o];
        Cm.row = ci[0]; Cm.col = ci[1]; Cm.range = ci[2]; Cm.num_rays = ci[3];
        Cm.inv_step = 1.0 / (Cm.fov / (double)Cm.num_rays);
        Cm.dom_lo = D.vc_lo[o];
        Cm.n_points = D.vc_meta[o * 2];
        Cm.P = D.vc_p + o * VC_POINTS;
        Cm.MK = D.vc_mask + o * (size_t)(VC_POINTS / 2) * VC_ROWS;
        IX = D.vc_idx + o * VC_IDX;
    }
    for (int i = lane; i < D.RW; i += 32) xvis[i] = 0;
    const int t_end = min(T, (b + 1) * FAST_TB);
    for (int t = b * FAST_TB; t < t_end; ++t) {
        // every camera's window start (coarse index) and the ray count below it: one load level for all cameras
        if (lane < n_cams) {
            FastCam &Cm = cams[lane];
            if (t > b * FAST_TB) Cm.heading = py_mod360(__dadd_rn(Cm.heading, Cm.speed));
            const double base = Cm.heading - Cm.fov * 0.5;
            const int q = max(0, min(VC_IDX - 1, (int)floor(base - Cm.dom_lo)));
            const int s0 = IX[q];
            Cm.s0 = s0;
            Cm.carry = s0 > 0 ? fast_nrays(Cm.P[s0 - 1], base, Cm.inv_step, Cm.num_rays + 1) : 0;
        }
        __syncwarp();
        uint32_t vis[RPL][W];
#pragma unroll
        for (int a = 0; a < RPL; ++a)
#pragma unroll
            for (int w = 0; w < W; ++w) vis[a][w] = 0;
        bool exact_used = false;
        for (int k = 0; k < n_cams; ++k) {
            const FastCam &Cm = cams[k];
            const double base = Cm.heading - Cm.fov * 0.5, inv_step = Cm.inv_step;
            const int NR = Cm.num_rays + 1, n_points = Cm.n_points;
            const double *P = Cm.P;
            int carry = Cm.carry;
            uint32_t acc[VC_ROWS / 2];
#pragma unroll
            for (int i = 0; i < VC_ROWS / 2; ++i) acc[i] = 0;
            bool more = true;
            for (int sb = Cm.s0; more; sb += 128) {
                double pv[4];   // four 32-segment passes in flight
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int s = sb + 32 * u + lane;
                    pv[u] = s < n_points ? P[s] : 1e300;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (!more) break;
                    const int s = sb + 32 * u + lane;
                    const int n_s = fast_nrays(pv[u], base, inv_step, NR);
                    int n_prev = __shfl_up_sync(0xffffffffu, n_s, 1);
                    if (lane == 0) n_prev = carry;
                    carry = __shfl_sync(0xffffffffu, n_s, 31);
                    if (n_s > n_prev) {
                        if (s & 1) {   // gap (s - 1) / 2: every ray inside marks the same tiles
                            const uint4 *mk = reinterpret_cast<const uint4 *>(Cm.MK + (size_t)(s >> 1) * VC_ROWS);
                            const uint4 m0 = __ldg(mk), m1 = __ldg(mk + 1);
                            acc[0] |= m0.x; acc[1] |= m0.y; acc[2] |= m0.z; acc[3] |= m0.w;
                            acc[4] |= m1.x; acc[5] |= m1.y; acc[6] |= m1.z; acc[7] |= m1.w;
                        } else {       // band: rays on (or within 1e-9 degree of) a rounding tie -> exact march
                            exact_used = true;
                            for (int ri = n_prev; ri < n_s; ++ri)
                                vc_ray(D, wall_g, Cm.row, Cm.col, Cm.fov, Cm.heading, Cm.num_rays, 2 * Cm.range, 0.5, ri,
                                       [&](int r, int c) {
                                           if (r == Cm.row && c == Cm.col) return;
                                           atomicOr(&xvis[r * D.W + (c >> 5)], 1u << (c & 31));
                                       });
                        }
                    }
                    if (carry >= NR) more = false;  // warp-uniform
                }
            }
#pragma unroll
            for (int i = 0; i < VC_ROWS / 2; ++i) acc[i] = __reduce_or_sync(0xffffffffu, acc[i]);
            if (lane == 0) {
                reinterpret_cast<uint4 *>(stage)[0] = make_uint4(acc[0], acc[1], acc[2], acc[3]);
                reinterpret_cast<uint4 *>(stage)[1] = make_uint4(acc[4], acc[5], acc[6], acc[7]);
            }
            __syncwarp();
            const uint16_t *rows = reinterpret_cast<const uint16_t *>(stage);
#pragma unroll
            for (int a = 0; a < RPL; ++a) {
                const int wr = lane + 32 * a - (Cm.row - Cm.range);
                if (wr >= 0 && wr <= 2 * Cm.range) fast_or_row<W>(vis[a], rows[wr], Cm.col - Cm.range);
            }
            __syncwarp();
        }
        const bool ex = __any_sync(0xffffffffu, exact_used);
        if (ex) __syncwarp();
#pragma unroll
        for (int a = 0; a < RPL; ++a) {
            const int r = lane + 32 * a;
            if (r < D.R) {
#pragma unroll
                for (int w = 0; w < W; ++w) {
                    uint32_t v = vis[a][w];
                    if (ex) { v |= xvis[r * D.W + w]; xvis[r * D.W + w] = 0; }
                    out[((size_t)t * D.N + env) * D.RW + r * D.W + w] = v;
                }
            }
        }
        __syncwarp();
    }
}

// Guards of the env at their current waypoints -> OR into vis (guard cones + own tiles, visibility.py:44-59).
template <int RPL, int W>
__device__ __forceinline__ void fast_guards(const Dev &D, int env, int lane, uint32_t (&vis)[RPL][W], const FastGuard *guards,
                                            int n_guards, uint32_t *xvis) {
    bool exact_used = false;
    for (int g = 0; g < n_guards; ++g) {
        const FastGuard &G = guards[g];
        const size_t o = (size_t)env * D.Kg + g;
        const int row = D.guard_path[(o * D.L + G.idx) * 2], col = D.guard_path[(o * D.L + G.idx) * 2 + 1];
        if (G.hslot >= 0) {
            const uint16_t *mk = D.vg_mask + ((o * D.L + G.idx) * (size_t)(D.L + 1) + G.hslot) * VC_ROWS;
#pragma unroll
            for (int a = 0; a < RPL; ++a) {
                const int wr = lane + 32 * a - (row - G.range);
                if (wr >= 0 && wr <= 2 * G.range) fast_or_row<W>(vis[a], mk[wr], col - G.range);
            }
        } else {   // a heading that is not one of the path's (state written by hand): march the whole cone
            exact_used = true;
            const uint32_t *wall_g = D.wall + (size_t)env * D.RW;
            for (int ri = lane; ri <= G.num_rays; ri += 32)
                vc_ray(D, wall_g, row, col, G.fov, G.heading, G.num_rays, G.range, 1.0, ri,
                       [&](int r, int c) { atomicOr(&xvis[r * D.W + (c >> 5)], 1u << (c & 31)); });
            if (lane == 0) atomicOr(&xvis[row * D.W + (col >> 5)], 1u << (col & 31));
        }
    }
    if (__any_sync(0xffffffffu, exact_used)) {
        __syncwarp();
#pragma unroll
        for (int a = 0; a < RPL; ++a) {
            const int r = lane + 32 * a;
            if (r < D.R) {
#pragma unroll
                for (int w = 0; w < W; ++w) { vis[a][w] |= xvis[r * D.W + w]; xvis[r * D.W + w] = 0; }
            }
        }
        __syncwarp();
    }
}

// T steps per launch (T = 1: HeistEnvironment.step), or -- with do_reset -- HeistEnvironment.reset for the masked
// envs.  cam_vis[t][env][RW] holds the camera part of tick t's visibility (k_cam_vis); when vis_out == cam_vis the
// trajectory is finished in place.
template <int RPL, int W>
__global__ void __launch_bounds__(FAST_WARPS * 32)
k_dyn(Dev D, const int8_t *__restrict__ actions, int T, int autoreset, float *__restrict__ reward,
      double *__restrict__ reward64, uint8_t *__restrict__ done, uint8_t *__restrict__ status_out,
      const uint32_t *cam_vis, uint32_t *vis_out, int do_reset, const uint8_t *__restrict__ mask) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int env = blockIdx.x * FAST_WARPS + warp;
    if (env >= D.N || !D.env_cached[env]) return;
    if (do_reset && mask && !mask[env]) return;
    unsigned char *sp = smem + (size_t)warp * dyn_warp_bytes(D.RW, D.Kg);
    FastGuard *guards = reinterpret_cast<FastGuard *>(sp);    sp += (size_t)D.Kg * sizeof(FastGuard);
    uint32_t *xvis = reinterpret_cast<uint32_t *>(sp);

    // ---- load ----
    const int4 es = *reinterpret_cast<const int4 *>(D.env_s + (size_t)env * 4);
    const int n_cams = es.x, n_guards = es.y;
    const int4 d0 = *reinterpret_cast<const int4 *>(D.env_d + (size_t)env * 8);
    const int4 d1 = *reinterpret_cast<const int4 *>(D.env_d + (size_t)env * 8 + 4);
    EnvRegs E;
    E.r = d0.x & 0xffff; E.c = d0.x >> 16; E.tick = d0.y; E.prev = d0.z; E.init = d0.w;
    E.flags = d1.x & 0xff; E.n_vault = d1.y; E.n_detect = d1.z; E.n_timeout = d1.w;
    uint32_t wall[RPL][W], vis[RPL][W], cam[RPL][W];
#pragma unroll
    for (int a = 0; a < RPL; ++a) {
        const int r = lane + 32 * a;
#pragma unroll
        for (int w = 0; w < W; ++w) {
            wall[a][w] = r < D.R ? D.wall[(size_t)env * D.RW + r * D.W + w] : 0xffffffffu;
            vis[a][w] = r < D.R ? D.vis[(size_t)env * D.RW + r * D.W + w] : 0u;
            cam[a][w] = r < D.R ? cam_vis[(size_t)env * D.RW + r * D.W + w] : 0u;   // tick 0
        }
    }
    for (int i = lane; i < D.RW; i += 32) xvis[i] = 0;
    if (lane < n_guards) {
        const size_t o = (size_t)env * D.Kg + lane;
        FastGuard &G = guards[lane];
        const int4 gi = *reinterpret_cast<const int4 *>(D.guard_i + o * 4);
        G.len = gi.x; G.speed = gi.y; G.range = gi.z; G.num_rays = gi.w;
        G.fov = D.guard_fov[o]; G.heading = D.guard_heading[o]; G.idx = D.guard_idx[o];
        G.nh = D.vg_nh[o];
        G.hslot = -1;
        for (int s = 0; s < G.nh; ++s)
            if (__double_as_longlong(D.vg_hval[o * (D.L + 1) + s]) == __double_as_longlong(G.heading)) { G.hslot = s; break; }
    }
    __syncwarp();

    int status = HEIST_RUNNING;
    int n_adv = 0;   // camera updates executed by this launch
    if (do_reset) {
        E.r = D.start_r; E.c = D.start_c; E.tick = 0; E.flags = 0;
        E.prev = abs(E.r - D.vault_r) + abs(E.c - D.vault_c); E.init = E.prev;
        if (lane < n_guards) guards[lane].idx = 0;
        __syncwarp();
#pragma unroll
        for (int a = 0; a < RPL; ++a)
#pragma unroll
            for (int w = 0; w < W; ++w) vis[a][w] = cam[a][w];
        fast_guards<RPL, W>(D, env, lane, vis, guards, n_guards, xvis);
    }
    bool pending_reset = false;
    for (int t = 0; t < T;) {
        const size_t o = (size_t)t * D.N + env;
        int kind = 0;  // 0: already done, 1: step, 2: auto-reset
        if (pending_reset) {   // the trainer's `if done: reset()` (environment.py:183-214)
            E.r = D.start_r; E.c = D.start_c; E.tick = 0; E.flags = 0;
            E.prev = abs(E.r - D.vault_r) + abs(E.c - D.vault_c); E.init = E.prev;
            if (lane < n_guards) guards[lane].idx = 0;
            kind = 2;
        } else if (!(E.flags & F_DONE)) {   // a done env is not mutated (:232-233)
            const int action = actions[o];
            // move (:239-246): blocked by the grid edge or a WALL tile
            const int nr = E.r + (action == 2) - (action == 1), nc = E.c + (action == 4) - (action == 3);
            if (nr >= 0 && nr < D.R && nc >= 0 && nc < D.C) {   // (warp-uniform)
                if (!fast_bit<RPL, W>(wall, nr, nc)) { E.r = nr; E.c = nc; }
            }
            ++n_adv;   // cameras rotate (:251-252): their cones for this tick are cam_vis[t]
            if (lane < n_guards) {
                FastGuard &G = guards[lane];
                if (G.len >= 2) {
                    const size_t go = ((size_t)env * D.Kg + lane) * D.L;
                    const int old = G.idx;
                    const double h = D.guard_head[go + old];
                    if (h == h) { G.heading = h; G.hslot = D.vg_hslot[go + old]; }  // NaN: the move is (0, 0)
                    G.idx = py_imod(old + G.speed, G.len);
                }
            }
            kind = 1;
        }
        __syncwarp();
        if (kind) {
#pragma unroll
            for (int a = 0; a < RPL; ++a)
#pragma unroll
                for (int w = 0; w < W; ++w) vis[a][w] = cam[a][w];
            fast_guards<RPL, W>(D, env, lane, vis, guards, n_guards, xvis);
        }
        if (kind != 2) {
            double rw = 0.0;
            status = HEIST_ALREADY_DONE;
            if (kind == 1) {
                // shaping (:261-269), detection (:273-281), vault (:284-288), timeout (:291-297)
                rw = D.reward_step;
                status = HEIST_RUNNING;
                const int curr = abs(E.r - D.vault_r) + abs(E.c - D.vault_c);
                rw = __dadd_rn(rw, __dmul_rn((double)(E.prev - curr), 0.1));
                E.prev = curr;
                if (curr <= 3 && E.init > 3) rw = __dadd_rn(rw, __dmul_rn(0.05, (double)(3 - curr)));
                if (fast_bit<RPL, W>(vis, E.r, E.c)) {
                    E.flags |= F_DETECTED | F_DONE;
                    rw = __dadd_rn(rw, D.reward_detection);
                    status = HEIST_DETECTED;
                }
                if (E.r == D.vault_r && E.c == D.vault_c) {
                    E.flags |= F_VAULT | F_DONE;
                    rw = __dadd_rn(rw, D.reward_vault);
                    status = HEIST_VAULT_REACHED;
                }
                E.tick += 1;
                if (E.tick >= D.max_steps) {
                    E.flags |= F_DONE;
                    status = HEIST_TIMEOUT;
                    double cf = __dsub_rn(1.0, __ddiv_rn((double)curr, (double)max(E.init, 1)));
                    if (!(cf > 0.0)) cf = 0.0;
                    rw = __dadd_rn(rw, __dmul_rn(cf, 2.0));
                }
                if (status == HEIST_VAULT_REACHED) E.n_vault++;        // training.py:535-540
                else if (status == HEIST_DETECTED) E.n_detect++;
                else if (status == HEIST_TIMEOUT) E.n_timeout++;
            }
            if (lane == 0) {
                if (reward) reward[o] = (float)rw;
                if (reward64) reward64[o] = rw;
                if (done) done[o] = (E.flags & F_DONE) ? 1 : 0;
                if (status_out) status_out[o] = (uint8_t)status;
            }
            pending_reset = autoreset && (E.flags & F_DONE);
        } else pending_reset = false;
        if (!pending_reset) {   // tick t is complete: its visibility map is final
            ++t;
#pragma unroll
            for (int a = 0; a < RPL; ++a) {
                const int r = lane + 32 * a;
                if (r < D.R) {
#pragma unroll
                    for (int w = 0; w < W; ++w) {
                        const size_t at = o * D.RW + r * D.W + w;
                        if (t < T) cam[a][w] = cam_vis[at + (size_t)D.N * D.RW];   // next tick's camera cones
                        if (vis_out) vis_out[at] = vis[a][w];
                    }
                }
            }
        }
    }

    // ---- store ----
    if (lane == 0) {
        *reinterpret_cast<int4 *>(D.env_d + (size_t)env * 8) = make_int4(E.r | (E.c << 16), E.tick, E.prev, E.init);
        *reinterpret_cast<int4 *>(D.env_d + (size_t)env * 8 + 4) =
            make_int4(E.flags | (status << 8), E.n_vault, E.n_detect, E.n_timeout);
    }
    if (lane < n_cams) {   // headings after the camera updates this launch executed
        const size_t co = (size_t)env * D.Kc + lane;
        double h = D.cam_heading[co];
        const double speed = D.cam_f[co * 2 + 1];
        for (int a = 0; a < n_adv; ++a) h = py_mod360(__dadd_rn(h, speed));
        D.cam_heading[co] = h;
    }
    if (lane < n_guards) {
        const size_t go = (size_t)env * D.Kg + lane;
        D.guard_heading[go] = guards[lane].heading;
        D.guard_idx[go] = guards[lane].idx;
    }
#pragma unroll
    for (int a = 0; a < RPL; ++a) {
        const int r = lane + 32 * a;
        if (r < D.R) {
#pragma unroll
            for (int w = 0; w < W; ++w) D.vis[(size_t)env * D.RW + r * D.W + w] = vis[a][w];
        }
    }
}
