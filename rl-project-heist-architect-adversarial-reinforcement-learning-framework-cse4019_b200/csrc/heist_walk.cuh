// heist_walk.cuh -- k_walk: the sequential half of the table-driven step, one WARP per env.
//
// Reference: HeistEnvironment.reset / step (environment.py:183-299), Guard.update (security.py:145-159),
// DynamicVisibilityMap.update (visibility.py:31-65; the guards' part: cones + own tiles).
//
// k_cam_vis (heist_fast.cuh) has already written the union of the CAMERA cones of every tick of the chunk into
// buf[t][env] (cameras never react to the Solver).  What is left is inherently sequential in t -- move, patrol,
// detection, vault / timeout, rewards, auto-reset -- but tiny, and independent between envs.  Round 1 ran it as one
// THREAD per env (k_seq: 128 warps for 4 096 envs, each alone on a scheduler, ~500 dependent warp-instructions per tick
// => a 75 us serial wall per 32-tick chunk) followed by a warp-per-env pass that OR-ed the guards in (k_finish).  Here
// a warp owns the env for the whole chunk:
//   * the scalar state (position, tick, guard indices) is warp-uniform; 4 096 warps interleave on the schedulers, so
//     the per-tick chain is hidden instead of exposed;
//   * lane = grid row: the env's wall rows live in registers (a move test is one shuffle), the camera rows of the
//     next ticks are prefetched (they do not depend on the state), the guards' cached cone rows are OR-ed in by the
//     lane that owns the row, detection is one shuffle of the Solver's row -- and the completed map is written back
//     in place, so there is no second pass and no (waypoint, slot) record traffic between kernels;
//   * lane = tick for the per-tick scalars: actions are fetched 32 ticks at a time (lane l holds tick t0 + l) and
//     reward / done / status are collected the same way and stored once per 32 ticks.
#pragma once
#include "heist_cache.cuh"
#include "heist_step.cuh"

#define WALK_WARPS 4
#define WALK_PF 2     // camera rows are requested this many ticks ahead

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

template <int RPL, int W>
__global__ void __launch_bounds__(WALK_WARPS * 32)
k_walk(Dev D, const int8_t *__restrict__ actions, int T, int autoreset, float *__restrict__ reward,
       double *__restrict__ reward64, uint8_t *__restrict__ done, uint8_t *__restrict__ status_out,
       uint32_t *buf, int write_traj, int do_reset, const uint8_t *__restrict__ mask, int store_heading) {
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int G = VC_MAX_GUARDS;
    const int lane = threadIdx.x & 31;
    const int env = blockIdx.x * WALK_WARPS + (threadIdx.x >> 5);
    if (env >= D.N || !D.env_cached[env]) return;
    if (do_reset && mask && !mask[env]) return;
    const int R = D.R, C = D.C, N = D.N, RW = D.RW, L = D.L, Kg = D.Kg;

    // ---- load (warp-uniform scalars are loaded by every lane: broadcast, one transaction) ----
    const int4 es = *reinterpret_cast<const int4 *>(D.env_s + (size_t)env * 4);
    const int n_cams = es.x, n_guards = es.y;
    const int4 d0 = *reinterpret_cast<const int4 *>(D.env_d + (size_t)env * 8);
    const int4 d1 = *reinterpret_cast<const int4 *>(D.env_d + (size_t)env * 8 + 4);
    EnvRegs E;
    E.r = d0.x & 0xffff; E.c = d0.x >> 16; E.tick = d0.y; E.prev = d0.z; E.init = d0.w;
    E.flags = d1.x & 0xff; E.n_vault = d1.y; E.n_detect = d1.z; E.n_timeout = d1.w;
    uint32_t wl[RPL][W], cur[RPL][W];   // wall rows / the env's current visibility map, rows lane (+ 32)
#pragma unroll
    for (int a = 0; a < RPL; ++a)
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const int r = lane + 32 * a;
            wl[a][w] = r < R ? D.wall[(size_t)env * RW + r * W + w] : 0xffffffffu;
            cur[a][w] = r < R ? D.vis[(size_t)env * RW + r * W + w] : 0u;
        }
    int gk[G], ghs[G], glen[G], gstp[G], grng[G];   // waypoint, heading slot, path length, stride, range
    unsigned pw[G], gw[G];                          // lane k: patrol word of waypoint k (row | col << 8 | slot << 16,
    const uint16_t *gmask[G];                       //   slot taken when LEAVING it, 255 = unchanged); gw: word at gk
#pragma unroll
    for (int g = 0; g < G; ++g) {
        gk[g] = ghs[g] = 0; glen[g] = 1; gstp[g] = 0; grng[g] = 0; pw[g] = gw[g] = 0; gmask[g] = D.vg_mask;
        if (g < n_guards) {
            const size_t o = (size_t)env * Kg + g;
            const int4 gi = *reinterpret_cast<const int4 *>(D.guard_i + o * 4);   // len, speed, range, num_rays
            glen[g] = gi.x; gstp[g] = gi.x >= 2 ? py_imod(gi.y, gi.x) : 0; grng[g] = gi.z;
            gk[g] = D.guard_idx[o];
            if (lane < gi.x)
                pw[g] = (unsigned)D.guard_path[(o * L + lane) * 2] | ((unsigned)D.guard_path[(o * L + lane) * 2 + 1] << 8) |
                        ((unsigned)D.vg_hslot[o * L + lane] << 16);
            // heading -> slot.  A heading that is none of the path's can only have been written by hand into the
            // state view; it is reported (ERR_STATE) and treated as the default heading.
            const long long hb = __double_as_longlong(D.guard_heading[o]);
            const int nh = D.vg_nh[o];
            const double *hv = D.vg_hval + o * (L + 1);   // at most L + 1 <= 33 distinct headings
            const unsigned m0 = __ballot_sync(FULL, lane < nh && __double_as_longlong(hv[lane]) == hb);
            const unsigned m1 = __ballot_sync(FULL, lane + 32 < nh && __double_as_longlong(hv[min(lane + 32, L)]) == hb);
            ghs[g] = m0 ? __ffs(m0) - 1 : (m1 ? 31 + __ffs(m1) : 0);
            if (!(m0 | m1) && lane == 0) atomicOr(D.err, ERR_STATE);
            gmask[g] = D.vg_mask + o * L * (size_t)(L + 1) * VC_ROWS;
            gw[g] = __shfl_sync(FULL, pw[g], gk[g]);
        }
    }

    // wall bit at (nr, nc), all arguments warp-uniform; outside the grid blocks (:242-245)
    auto blocked = [&](int nr, int nc) -> bool {
        if (nr < 0 || nr >= R || nc < 0 || nc >= C) return true;
        uint32_t mine = wl[0][0];
        if (W == 2 && (nc >> 5)) mine = wl[0][W - 1];
        if (RPL == 2 && nr >= 32) { mine = wl[RPL - 1][0]; if (W == 2 && (nc >> 5)) mine = wl[RPL - 1][W - 1]; }
        return (__shfl_sync(FULL, mine, nr & 31) >> (nc & 31)) & 1u;
    };
    // v |= the guards' cones and own tiles at their current (waypoint, heading slot)  (visibility.py:44-59)
    auto or_guards = [&](uint32_t (&v)[RPL][W]) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
            if (g < n_guards) {
                const int prow = gw[g] & 255, pcol = (gw[g] >> 8) & 255, rng = grng[g];
                const uint16_t *m = gmask[g] + (gk[g] * (L + 1) + ghs[g]) * VC_ROWS;
#pragma unroll
                for (int a = 0; a < RPL; ++a) {
                    const int wr = lane + 32 * a - (prow - rng);
                    if (wr >= 0 && wr <= 2 * rng) fast_or_row<W>(v[a], __ldg(m + wr), pcol - rng);
                }
            }
        }
    };
    auto reset_state = [&]() {   // HeistEnvironment.reset (:183-214): headings persist, guards back to waypoint 0
        E.r = D.start_r; E.c = D.start_c; E.tick = 0; E.flags = 0;
        E.prev = abs(E.r - D.vault_r) + abs(E.c - D.vault_c); E.init = E.prev;
#pragma unroll
        for (int g = 0; g < G; ++g)
            if (g < n_guards) { gk[g] = 0; gw[g] = __shfl_sync(FULL, pw[g], 0); }
    };
    auto load_rows = [&](uint32_t (&v)[RPL][W], int t) {
#pragma unroll
        for (int a = 0; a < RPL; ++a)
#pragma unroll
            for (int w = 0; w < W; ++w) {
                const int r = lane + 32 * a;
                v[a][w] = r < R ? buf[((size_t)t * N + env) * RW + r * W + w] : 0u;
            }
    };
    auto store_rows = [&](const uint32_t (&v)[RPL][W], int t) {
#pragma unroll
        for (int a = 0; a < RPL; ++a)
#pragma unroll
            for (int w = 0; w < W; ++w) {
                const int r = lane + 32 * a;
                if (r < R) buf[((size_t)t * N + env) * RW + r * W + w] = v[a][w];
            }
    };

    int status = HEIST_RUNNING;
    int n_adv = 0;   // camera updates executed by this launch
    if (do_reset) {
        reset_state();
        load_rows(cur, 0);   // camera cones at the (unchanged) headings
        or_guards(cur);
        T = 0;
    }
    // lane l holds the action of tick tb + l; the next 32 are requested one group ahead
    int a_now = 0, a_next = 0;
    if (lane < T) a_now = actions[(size_t)lane * N + env];
    if (32 + lane < T) a_next = actions[(size_t)(32 + lane) * N + env];
    uint32_t pre[WALK_PF][RPL][W];
#pragma unroll
    for (int j = 0; j < WALK_PF; ++j) if (j < T) load_rows(pre[j], j);
    float o_rw = 0.f; double o_rw64 = 0.0; int o_dn = 0, o_st = 0;
    for (int t0 = 0; t0 < T; t0 += WALK_PF) {
#pragma unroll
        for (int j = 0; j < WALK_PF; ++j) {
            const int t = t0 + j;
            if (t >= T) break;
            uint32_t v[RPL][W];
#pragma unroll
            for (int a = 0; a < RPL; ++a)
#pragma unroll
                for (int w = 0; w < W; ++w) v[a][w] = pre[j][a][w];
            if (t + WALK_PF < T) load_rows(pre[j], t + WALK_PF);
            const int act = __shfl_sync(FULL, a_now, t & 31);
            double rw = 0.0;
            bool rebuilt = false;
            status = HEIST_ALREADY_DONE;
            if (!(E.flags & F_DONE)) {   // a done env is not mutated (:232-233)
                // move (:239-246): blocked by the grid edge or a WALL tile
                const int nr = E.r + (act == 2) - (act == 1), nc = E.c + (act == 4) - (act == 3);
                if (!blocked(nr, nc)) { E.r = nr; E.c = nc; }
                ++n_adv;   // cameras rotate (:251-252): their cones for this tick are buf[t]
#pragma unroll
                for (int g = 0; g < G; ++g) {   // Guard.update (security.py:145-159)
                    if (g < n_guards && glen[g] >= 2) {
                        const int hsl = (gw[g] >> 16) & 255;
                        if (hsl != 255) ghs[g] = hsl;   // 255: the move is (0, 0), heading unchanged
                        gk[g] += gstp[g]; if (gk[g] >= glen[g]) gk[g] -= glen[g];
                        gw[g] = __shfl_sync(FULL, pw[g], gk[g]);
                    }
                }
                or_guards(v);   // visibility rebuild (:257-258): camera cones OR guard cones / own tiles
                uint32_t mine = v[0][0];
                if (W == 2 && (E.c >> 5)) mine = v[0][W - 1];
                if (RPL == 2 && E.r >= 32) { mine = v[RPL - 1][0]; if (W == 2 && (E.c >> 5)) mine = v[RPL - 1][W - 1]; }
                const bool detected = (__shfl_sync(FULL, mine, E.r & 31) >> (E.c & 31)) & 1u;
                // shaping (:261-269), detection (:273-281), vault (:284-288), timeout (:291-297)
                rw = D.reward_step;
                status = HEIST_RUNNING;
                const int curr = abs(E.r - D.vault_r) + abs(E.c - D.vault_c);
                rw = __dadd_rn(rw, __dmul_rn((double)(E.prev - curr), 0.1));
                E.prev = curr;
                if (curr <= 3 && E.init > 3) rw = __dadd_rn(rw, __dmul_rn(0.05, (double)(3 - curr)));
                if (detected) {
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
                rebuilt = true;
            }
            if (lane == (t & 31)) { o_rw = (float)rw; o_rw64 = rw; o_dn = (E.flags & F_DONE) ? 1 : 0; o_st = status; }
            if (autoreset && (E.flags & F_DONE)) {   // the trainer's `if done: reset()`: same cameras, guards at waypoint 0
                reset_state();
#pragma unroll
                for (int a = 0; a < RPL; ++a)
#pragma unroll
                    for (int w = 0; w < W; ++w) v[a][w] = 0u;
                load_rows(v, t);
                or_guards(v);
                rebuilt = true;
            }
            if (rebuilt) {
#pragma unroll
                for (int a = 0; a < RPL; ++a)
#pragma unroll
                    for (int w = 0; w < W; ++w) cur[a][w] = v[a][w];
            }
            if (write_traj) store_rows(cur, t);   // (a tick spent done without auto-reset keeps the frozen map)
            if ((t & 31) == 31 || t == T - 1) {   // lane l: outputs of tick tb + l; then the next group's actions
                const int tb = t & ~31;
                if (tb + lane <= t) {
                    const size_t o = (size_t)(tb + lane) * N + env;
                    if (reward) reward[o] = o_rw;
                    if (reward64) reward64[o] = o_rw64;
                    if (done) done[o] = (uint8_t)o_dn;
                    if (status_out) status_out[o] = (uint8_t)o_st;
                }
                a_now = a_next;
                if (tb + 64 + lane < T) a_next = actions[(size_t)(tb + 64 + lane) * N + env];
            }
        }
    }

    // ---- store ----
#pragma unroll
    for (int a = 0; a < RPL; ++a)
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const int r = lane + 32 * a;
            if (r < R) D.vis[(size_t)env * RW + r * W + w] = cur[a][w];
        }
    if (lane == 0) {
        *reinterpret_cast<int4 *>(D.env_d + (size_t)env * 8) = make_int4(E.r | (E.c << 16), E.tick, E.prev, E.init);
        *reinterpret_cast<int4 *>(D.env_d + (size_t)env * 8 + 4) =
            make_int4(E.flags | (status << 8), E.n_vault, E.n_detect, E.n_timeout);
    }
    if (store_heading && lane < n_cams) {   // headings after the camera updates this launch executed
        const size_t co = (size_t)env * D.Kc + lane;   // (otherwise stored by k_heads)
        double h = D.cam_heading[co];
        const double speed = D.cam_f[co * 2 + 1];
        for (int a = 0; a < n_adv; ++a) h = py_mod360(__dadd_rn(h, speed));
        D.cam_heading[co] = h;
    }
    if (lane < n_guards) {
        const size_t go = (size_t)env * Kg + lane;
        int k = 0, hs = 0;
#pragma unroll
        for (int g = 0; g < G; ++g) if (g == lane) { k = gk[g]; hs = ghs[g]; }
        D.guard_heading[go] = D.vg_hval[go * (L + 1) + hs];
        D.guard_idx[go] = k;
    }
}
