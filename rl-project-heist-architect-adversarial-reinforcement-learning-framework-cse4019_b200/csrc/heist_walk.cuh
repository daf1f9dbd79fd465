// heist_walk.cuh -- k_walk: the sequential half of the table-driven step, one WARP per env.
//
// Reference: HeistEnvironment.reset / step (environment.py:183-299), Guard.update (security.py:145-159),
// DynamicVisibilityMap.update (visibility.py:31-65; the guards' part: cones + own tiles).
//
// One launch = a few ticks (normally ONE: heist_step / heist_reset / heist_step_observe; long rollouts go through
// k_cam_vis_staged + k_seq + k_finish_or in heist_fast.cuh).  A warp owns the env:
//   * buf == nullptr (the fused single tick): the camera cones of the tick are scanned right here from the cache
//     tables in global memory (scan_window), the new headings are stored, and -- for heist_step_observe -- the dense
//     (3, R, C) state is written from the registers that hold the finished map; otherwise k_cam_vis has already
//     written the union of the camera cones of every tick into buf[t][env] (cameras never react to the Solver);
//   * the scalar state (position, tick, guard indices) is warp-uniform;
//   * lane = grid row: the env's wall rows live in registers (a move test is one shuffle), the guards' cached cone
//     rows are OR-ed in by the lane that owns the row, detection is one shuffle of the Solver's row;
//   * lane = tick for the per-tick scalars: actions are fetched 32 ticks at a time (lane l holds tick t0 + l) and
//     reward / done / status are collected the same way and stored once per 32 ticks.
// The tick is a chain of dependent round trips (~900 instructions, ~11 us even on an empty GPU), so the kernel is
// compiled for 7 CTAs per SM: all 4 096 warps of the bench batch are resident at once.
#pragma once
#include "heist_fast.cuh"

#define WALK_WARPS 4
#define WALK_PF 2     // camera rows are requested this many ticks ahead
__host__ __device__ inline size_t walk_warp_bytes(int RW, int Kc) {   // fused mode: per-camera constants + exact-ray rows
    return (((size_t)Kc * sizeof(FastCam) + (size_t)RW * 4) + 15) & ~(size_t)15;
}

template <int RPL, int W, bool FUSED>
__global__ void __launch_bounds__(WALK_WARPS * 32, 7)
k_walk(Dev D, const int8_t *__restrict__ actions, int T, int autoreset, float *__restrict__ reward,
       double *__restrict__ reward64, uint8_t *__restrict__ done, uint8_t *__restrict__ status_out,
       uint32_t *buf, int write_traj, int do_reset, const uint8_t *__restrict__ mask, int store_heading,
       float *__restrict__ state_out) {
    // FUSED (buf == nullptr): single tick (T <= 1) -- the camera cones of the tick are computed right here from the
    // cache tables (one launch instead of k_cam_vis + k_walk), the new headings are stored, and, if state_out is
    // given, the dense (3, R, C) state the policy reads next is written from the registers that hold the finished map.
    extern __shared__ __align__(16) unsigned char smem[];
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
    if (lane < n_guards) {   // the guards' records are read after the camera scan (their registers would only be
        const size_t o = (size_t)env * Kg + lane;   // spilled across it): requested now, so the reads below hit L1
        const void *pf[8] = {D.guard_i + o * 4, D.guard_path + o * L * 2, D.vg_hslot + o * L, D.guard_heading + o,
                             D.guard_idx + o, D.vg_nh + o, D.vg_hval + o * (L + 1), D.vg_reach + o * L};
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("prefetch.global.L1 [%0];" ::"l"(pf[i]));
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
    uint32_t camrows[RPL][W];   // fused mode: the camera cones of the tick
#pragma unroll
    for (int a = 0; a < RPL; ++a)
#pragma unroll
        for (int w = 0; w < W; ++w) camrows[a][w] = 0u;
    if (FUSED) {
        // Cameras of a live env rotate once before the cones are cast (:251-252); a reset, or the tick a done env
        // spends on the "already done" early-out, keeps the headings (:205-208, :232-233).
        const bool advance = !do_reset && T > 0 && !(E.flags & F_DONE);
        unsigned char *sp = smem + (size_t)(threadIdx.x >> 5) * walk_warp_bytes(RW, D.Kc);
        FastCam *cams = reinterpret_cast<FastCam *>(sp);
        uint32_t *xvis = reinterpret_cast<uint32_t *>(sp + (size_t)D.Kc * sizeof(FastCam));
        int my_s0 = 0, my_fx = 0;
        if (lane < n_cams) {
            const size_t o = (size_t)env * D.Kc + lane;
            FastCam &Cm = cams[lane];
            const int16_t *ci = D.cam_i + o * 4;
            Cm.fov = D.cam_f[o * 2]; Cm.speed = D.cam_f[o * 2 + 1];
            double h = D.cam_heading[o];
            if (advance) { h = py_mod360(__dadd_rn(h, Cm.speed)); D.cam_heading[o] = h; }
            Cm.h0 = h;
            Cm.row = ci[0]; Cm.col = ci[1]; Cm.range = ci[2]; Cm.num_rays = ci[3];
            { const double2 lf = *reinterpret_cast<const double2 *>(D.vc_lo + o * 2); Cm.dom_lo = lf.x; Cm.fx_scale = lf.y; }
            Cm.n_gaps = D.vc_meta[o * 2] >> 1;
            Cm.sh = D.vc_meta[o * 2 + 1];
            Cm.P2 = reinterpret_cast<const int2 *>(D.vc_p + o * VC_POINTS);
            Cm.MK4 = reinterpret_cast<const uint4 *>(D.vc_mask + o * (size_t)(VC_POINTS / 2) * VC_ROWS);
            const double base = h - Cm.fov * 0.5;
            const int q = max(0, min(VC_IDX - 1, (int)floor(base - Cm.dom_lo)));
            my_s0 = max(0, (int)D.vc_idx[o * VC_IDX + q] - 1) & ~1;
            my_fx = (int)fmax(-536870912.0, fmin(536870912.0, floor((base - Cm.dom_lo) * Cm.fx_scale)));
        }
        for (int i = lane; i < RW; i += 32) xvis[i] = 0;
        __syncwarp();
        bool exact_used = false;
        for (int k = 0; k < n_cams; ++k) {
            const FastCam &Cm = cams[k];
            const int s0 = __shfl_sync(FULL, my_s0, k);
            const int bias = ((1 << Cm.sh) - 1) - __shfl_sync(FULL, my_fx, k);
            uint32_t acc[VC_ROWS / 2];
            const unsigned bands = scan_window(Cm.P2, Cm.MK4, Cm.n_gaps, s0, bias, Cm.sh, Cm.num_rays + 1, lane, acc);
            place_rows<RPL, W>(acc, Cm.row - Cm.range, Cm.col - Cm.range, 2 * Cm.range, lane, camrows);
            if (bands) {   // warp-uniform
                cam_exact_scan(vc_geo(D), D.wall + (size_t)env * RW, xvis, &Cm, Cm.h0, s0, bias, lane);
                exact_used = true;
            }
        }
        if (exact_used) {
            __syncwarp();
#pragma unroll
            for (int a = 0; a < RPL; ++a)
#pragma unroll
                for (int w = 0; w < W; ++w) { const int r = lane + 32 * a; if (r < R) camrows[a][w] |= xvis[r * W + w]; }
        }
        store_heading = 0;   // stored above
    }
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
            if (lane == 0 && (!(m0 | m1) || !((D.vg_reach[o * L + min(max(gk[g], 0), L - 1)] >> min(ghs[g], 31)) & 1u)))
                atomicOr(D.err, ERR_STATE);   // (a heading of no waypoint, or a (waypoint, slot) cone that was never built)
            gmask[g] = D.vg_mask + o * L * (size_t)(L + 1) * VC_ROWS;
            gw[g] = __shfl_sync(FULL, pw[g], gk[g]);
        }
    }
    auto load_rows = [&](uint32_t (&v)[RPL][W], int t) {
#pragma unroll
        for (int a = 0; a < RPL; ++a)
#pragma unroll
            for (int w = 0; w < W; ++w) {
                const int r = lane + 32 * a;
                v[a][w] = FUSED ? camrows[a][w] : (r < R ? buf[((size_t)t * N + env) * RW + r * W + w] : 0u);
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
    // One env step (environment.py:216-299) on the camera cones v of the tick; returns whether the map was rebuilt.
    auto tick_step = [&](uint32_t (&v)[RPL][W], int act, double &rw) -> bool {
        rw = 0.0;
        status = HEIST_ALREADY_DONE;
        if (E.flags & F_DONE) return false;   // a done env is not mutated (:232-233)
        // move (:239-246): blocked by the grid edge or a WALL tile
        const int nr = E.r + (act == 2) - (act == 1), nc = E.c + (act == 4) - (act == 3);
        if (!blocked(nr, nc)) { E.r = nr; E.c = nc; }
        ++n_adv;   // cameras rotate (:251-252): their cones for this tick are v
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
        return true;
    };
    // the trainer's `if done: reset()`: same cameras (the cones of tick t), guards at waypoint 0
    auto tick_autoreset = [&](uint32_t (&v)[RPL][W], int t) {
        reset_state();
#pragma unroll
        for (int a = 0; a < RPL; ++a)
#pragma unroll
            for (int w = 0; w < W; ++w) v[a][w] = 0u;
        load_rows(v, t);
        or_guards(v);
    };
    if (FUSED) {   // exactly one tick (or none: a reset), nothing of the multi-tick machinery below
        if (T > 0) {
            uint32_t v[RPL][W];
            load_rows(v, 0);
            const int act = actions[env];   // (one address for the warp: a broadcast)
            double rw;
            bool rebuilt = tick_step(v, act, rw);
            const int dn = (E.flags & F_DONE) ? 1 : 0, st = status;
            if (autoreset && dn) { tick_autoreset(v, 0); rebuilt = true; }
            if (rebuilt) {
#pragma unroll
                for (int a = 0; a < RPL; ++a)
#pragma unroll
                    for (int w = 0; w < W; ++w) cur[a][w] = v[a][w];
            }
            if (lane == 0) {
                if (reward) reward[env] = (float)rw;
                if (reward64) reward64[env] = rw;
                if (done) done[env] = (uint8_t)dn;
                if (status_out) status_out[env] = (uint8_t)st;
            }
        }
    } else {
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
            double rw;
            bool rebuilt = tick_step(v, act, rw);
            if (lane == (t & 31)) { o_rw = (float)rw; o_rw64 = rw; o_dn = (E.flags & F_DONE) ? 1 : 0; o_st = status; }
            if (autoreset && (E.flags & F_DONE)) { tick_autoreset(v, t); rebuilt = true; }
            if (rebuilt) {
#pragma unroll
                for (int a = 0; a < RPL; ++a)
#pragma unroll
                    for (int w = 0; w < W; ++w) cur[a][w] = v[a][w];
            }
            if (write_traj && buf) store_rows(cur, t);   // (a tick spent done without auto-reset keeps the frozen map)
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
    }

    // ---- store ----
#pragma unroll
    for (int a = 0; a < RPL; ++a)
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const int r = lane + 32 * a;
            if (r < R) D.vis[(size_t)env * RW + r * W + w] = cur[a][w];
        }
    if (state_out) {
        // HeistEnvironment.get_state_tensor (environment.py:347-374) from the finished map in registers: channel 0 the
        // tile codes / 5, channel 1 the visibility map, channel 2 Solver +1 / vault -1 (vault wins) + distance
        // gradient; one float4 (4 cells, C % 4 == 0) per store, same arithmetic as k_observe_vec4.
        const int quads = D.RC >> 2, total = 3 * quads, vcell = D.vault_r * C + D.vault_c, scell = E.r * C + E.c;
        float4 *dst = reinterpret_cast<float4 *>(state_out) + (size_t)env * 3 * quads;
        // lane's quad q = q0 + lane walks the three channels 32 quads (128 cells) per trip: (channel, cell, row, col)
        // advance incrementally -- one integer division per lane up front instead of one per store
        int q = lane;
        int ch = (q >= quads) + (q >= 2 * quads);
        int cell = (q - ch * quads) << 2;
        int r = cell / C, c = cell - r * C;
        const int dr = 128 / C, dc = 128 - dr * C;
        for (int q0 = 0; q0 < total; q0 += 32) {   // (every lane takes every trip: the shuffles below need them all)
            const bool on = q < total;
            uint32_t word = 0;
            if (q0 + 31 >= quads && q0 < 2 * quads) {   // warp-uniform: some lane is in channel 1 (the visibility map)
#pragma unroll
                for (int a = 0; a < RPL; ++a)
#pragma unroll
                    for (int w2 = 0; w2 < W; ++w2) {
                        const uint32_t x = __shfl_sync(FULL, cur[a][w2], r & 31);
                        if ((r >> 5) == a && (c >> 5) == w2) word = x;
                    }
            }
            const int my_q = q, my_ch = ch, my_cell = cell, my_c = c;
            q += 32; cell += 128; c += dc; r += dr;
            if (c >= C) { c -= C; ++r; }
            while (cell >= D.RC) { cell -= D.RC; r -= R; ++ch; }
            if (!on) continue;
            float4 v;
            if (my_ch == 0) {
                const uchar4 t4 = *reinterpret_cast<const uchar4 *>(D.tile + (size_t)env * D.RC + my_cell);
                v.x = __fmul_rn((float)t4.x, 0.2f); v.y = __fmul_rn((float)t4.y, 0.2f);
                v.z = __fmul_rn((float)t4.z, 0.2f); v.w = __fmul_rn((float)t4.w, 0.2f);
            } else if (my_ch == 1) {
                const uint32_t bits = word >> (my_c & 31);
                v.x = (float)(bits & 1u); v.y = (float)((bits >> 1) & 1u); v.z = (float)((bits >> 2) & 1u); v.w = (float)((bits >> 3) & 1u);
            } else {
                const float4 g = *reinterpret_cast<const float4 *>(D.pos_tab + my_cell);
                const float b0 = (my_cell + 0 == vcell) ? -1.0f : ((my_cell + 0 == scell) ? 1.0f : 0.0f);
                const float b1 = (my_cell + 1 == vcell) ? -1.0f : ((my_cell + 1 == scell) ? 1.0f : 0.0f);
                const float b2 = (my_cell + 2 == vcell) ? -1.0f : ((my_cell + 2 == scell) ? 1.0f : 0.0f);
                const float b3 = (my_cell + 3 == vcell) ? -1.0f : ((my_cell + 3 == scell) ? 1.0f : 0.0f);
                v.x = __fadd_rn(b0, g.x); v.y = __fadd_rn(b1, g.y); v.z = __fadd_rn(b2, g.z); v.w = __fadd_rn(b3, g.w);
            }
            __stcs(dst + my_q, v);   // write-once stream
        }
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
