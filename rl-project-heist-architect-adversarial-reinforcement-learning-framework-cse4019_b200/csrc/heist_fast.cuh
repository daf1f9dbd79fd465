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
//   k_seq      thread per env, sequential in t: move, guard patrol, detection from one word of cam_vis[t] OR the
//              guards' cached masks (one per (waypoint, heading)) at the Solver's row, vault / timeout, rewards,
//              auto-reset; records each tick's guard state.  A serial chain, overlapped with k_cam_vis of the next
//              chunk of ticks (heist_b200.cu launch_fast);
//   k_finish   warp per (env, tick block): cam_vis[t] OR the guards' masks -> the visibility trajectory, in place.
// Rays that fall inside a tie band (or outside the cached angle domain) are marched exactly like the reference
// does, so results are bit-identical to heist_step.cuh's.
#pragma once
#include "heist_cache.cuh"
#include "heist_step.cuh"
#include "heist_walk.cuh"

#define FAST_WARPS 4
#define FAST_TB 8      // ticks per k_cam_vis warp

// Per-camera constants of a k_cam_vis warp (shared memory).
struct FastCam {
    double speed, fov, fx_scale, dom_lo;
    double h0;              // heading at the block's first tick
    const int2 *P2;         // gap g = boundary points 2g (its start), 2g + 1 (its end), fixed point (heist_cache.cuh)
    const uint4 *MK4;       // ... and its window mask: uint4 2g, 2g + 1
    int row, col, range, num_rays, n_gaps, sh;
};

__host__ __device__ inline size_t camvis_warp_bytes(int RW, int Kc) {
    return (size_t)Kc * sizeof(FastCam) + (((size_t)RW * 4 + 15) & ~(size_t)15) + (size_t)FAST_TB * Kc * 16;
}

// heads[b][env][k] = heading of camera k at the first tick of tick block b (FAST_TB ticks) of this launch; k_cam_vis
// advances it through the block.  write_final: with auto-reset every tick of the launch updates the cameras
// (environment.py:251-252), so the heading the launch ends on is the one of its last tick and is stored here;
// otherwise k_walk stores it (an env may stop stepping early).
__global__ void __launch_bounds__(128) k_heads(Dev D, int T, int do_reset, int write_final, double *__restrict__ heads) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= D.N * D.Kc) return;
    const int env = i / D.Kc, k = i - env * D.Kc;
    if (!D.env_cached[env] || k >= D.env_s[(size_t)env * 4]) return;
    double h = D.cam_heading[i];
    const double speed = D.cam_f[(size_t)i * 2 + 1];
    const int adv0 = fast_adv0(D, env, do_reset);
    for (int a = 0; a < adv0; ++a) h = py_mod360(__dadd_rn(h, speed));
    double last = h;
    for (int t = 0; t < T; ++t) {
        if (t % FAST_TB == 0) heads[(size_t)(t / FAST_TB) * D.N * D.Kc + i] = h;
        last = h;
        h = py_mod360(__dadd_rn(h, speed));
    }
    if (write_final && T > 0) D.cam_heading[i] = last;
}

// Rare path of k_cam_vis, kept out of line: the rays [r0, r1) of camera `cam` sit on (or within 1e-9 degree of) a
// rounding tie, or outside the cached angle domain -> march them exactly like the reference does
// (security.py:69-99), the whole warp on one ray: lane j evaluates sample j + 1 (at most 14 samples), a ballot finds
// the first blocked one.  Marks go to xvis (shared memory, OR-ed into the tick's map afterwards).
__device__ __noinline__ void cam_exact_rays(VcGeo D, const uint32_t *__restrict__ wall, uint32_t *xvis, const FastCam *cam,
                                            double heading, int r0, int r1, int lane) {
    const int row = cam->row, col = cam->col, nsamp = 2 * cam->range;
    const double fov = cam->fov;
    for (int ri = r0; ri < r1; ++ri) {
        const double angle_deg = __dadd_rn(__dsub_rn(heading, __ddiv_rn(fov, 2.0)),
                                           __ddiv_rn(__dmul_rn(fov, (double)ri), (double)cam->num_rays));
        double dx, dy;
        ray_dir(angle_deg, D.deg2rad, dx, dy);
        const double dist = 0.5 * (double)(lane + 1);   // step - 1 + sub: exact multiples of 0.5
        int r = 0, c = 0;
        bool blocked = true;
        if (lane < nsamp) {
            c = vc_rint_even(__dadd_rn((double)col, __dmul_rn(dx, dist)));
            r = vc_rint_even(__dadd_rn((double)row, __dmul_rn(dy, dist)));
            blocked = r < 0 || r >= D.R || c < 0 || c >= D.C;
            if (!blocked) blocked = (__ldg(wall + r * D.W + (c >> 5)) >> (c & 31)) & 1u;
        }
        const int first_blocked = __ffs(__ballot_sync(0xffffffffu, blocked)) - 1;   // lanes >= nsamp always are
        if (lane < first_blocked && !(r == row && c == col))   // (r, c) != (self.row, self.col), security.py:93
            atomicOr(&xvis[r * D.W + (c >> 5)], 1u << (c & 31));
    }
}

// ... for one (tick, camera) whose gap scan saw a band with a ray in it: the same scan again, boundary points only,
// marching the rays of every such band.  Out of line and outside the hot loop, so that the hot loop holds no call
// (a call there makes the compiler spill the loop's registers around it).
__device__ __noinline__ void cam_exact_scan(VcGeo D, const uint32_t *__restrict__ wall, uint32_t *xvis, const FastCam *cam,
                                            double heading, int s0, int bias, int lane) {
    const int sh = cam->sh, NR = cam->num_rays + 1;
    const int2 *P2 = cam->P2;
    int carry = 0;
    bool first = s0 > 0;
    for (int g = (s0 >> 1) + lane;; g += 32) {
        const int2 p = __ldg(P2 + min(g, VC_POINTS / 2 - 1));
        const int lo = max(0, min(NR, (p.x + bias) >> sh)), hi = max(0, min(NR, (p.y + bias) >> sh));
        int ph = __shfl_up_sync(0xffffffffu, hi, 1);
        if (lane == 0) ph = first ? lo : carry;
        first = false;
        carry = __shfl_sync(0xffffffffu, hi, 31);
        unsigned bh = __ballot_sync(0xffffffffu, lo > ph);
        while (bh) {
            const int src = __ffs(bh) - 1;
            bh &= bh - 1;
            const int r0 = __shfl_sync(0xffffffffu, ph, src), r1 = __shfl_sync(0xffffffffu, lo, src);
            cam_exact_rays(D, wall, xvis, cam, heading, r0, r1, lane);
        }
        if (carry >= NR) break;
    }
}

// Union of the camera cones of one env for FAST_TB consecutive ticks -> out[t][env][RW].
// One pass = 32 consecutive gaps of a camera's table, lane j = gap g0 + j: it loads the gap's two boundary points
// (one 8-byte load) and, independently, its 32-byte mask; the ray counts below the two points say whether the gap
// holds a ray (then the mask is OR-ed in, branch-free) and, with the upper count of lane j - 1, whether the band in
// front of it does (rare: those rays are marched exactly, out of line).  Two passes are in flight together (a
// window of the bench workload spans 30-60 gaps).  The 8 mask words are OR-reduced over the warp (REDUX: the result
// is warp-uniform) and every lane picks the 16-bit window row of the grid row it owns.
template <int RPL, int W>
__global__ void __launch_bounds__(FAST_WARPS * 32, 6)
k_cam_vis(Dev D, int T, int nblk, const double *__restrict__ heads, uint32_t *__restrict__ out, const uint8_t *__restrict__ mask,
          int do_reset) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr unsigned FULL = 0xffffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long wid = (long long)blockIdx.x * FAST_WARPS + warp;
    const int env = (int)(wid / nblk), b = (int)(wid - (long long)env * nblk);
    if (env >= D.N || !D.env_cached[env]) return;
    if (mask && !mask[env]) return;
    unsigned char *sp = smem + (size_t)warp * camvis_warp_bytes(D.RW, D.Kc);
    FastCam *cams = reinterpret_cast<FastCam *>(sp);        sp += (size_t)D.Kc * sizeof(FastCam);
    uint32_t *xvis = reinterpret_cast<uint32_t *>(sp);      sp += ((size_t)D.RW * 4 + 15) & ~(size_t)15;
    double *pre_head = reinterpret_cast<double *>(sp);       sp += (size_t)FAST_TB * D.Kc * 8;   // [tick][camera]
    int *pre_s0 = reinterpret_cast<int *>(sp);               sp += (size_t)FAST_TB * D.Kc * 4;
    int *pre_fx = reinterpret_cast<int *>(sp);
    const int n_cams = D.env_s[(size_t)env * 4];
    if (lane < n_cams) {
        double h0;
        const size_t o = (size_t)env * D.Kc + lane;
        FastCam &Cm = cams[lane];
        const int16_t *ci = D.cam_i + o * 4;
        Cm.fov = D.cam_f[o * 2]; Cm.speed = D.cam_f[o * 2 + 1];
        if (heads) h0 = heads[((size_t)b * D.N + env) * D.Kc + lane];
        else {   // single tick block: no k_heads launch, the heading of tick 0 is one update (or none) away
            h0 = D.cam_heading[o];
            if (fast_adv0(D, env, do_reset)) h0 = py_mod360(__dadd_rn(h0, Cm.speed));
        }
        Cm.h0 = h0;
        Cm.row = ci[0]; Cm.col = ci[1]; Cm.range = ci[2]; Cm.num_rays = ci[3];
        Cm.dom_lo = D.vc_lo[o];
        Cm.n_gaps = D.vc_meta[o * 2] >> 1;
        Cm.sh = D.vc_meta[o * 2 + 1];
        Cm.fx_scale = (1.0 / (Cm.fov / (double)Cm.num_rays)) * (double)(1 << Cm.sh);
        Cm.P2 = reinterpret_cast<const int2 *>(D.vc_p + o * VC_POINTS);
        Cm.MK4 = reinterpret_cast<const uint4 *>(D.vc_mask + o * (size_t)(VC_POINTS / 2) * VC_ROWS);
    }
    for (int i = lane; i < D.RW; i += 32) xvis[i] = 0;
    const int t_begin = b * FAST_TB, t_end = min(T, (b + 1) * FAST_TB);
    __syncwarp();
    // Per (tick, camera) of the block, in parallel lanes: the heading (the block's first, advanced tick by tick), the
    // window start from the coarse index (segments below IX[q] hold no ray of that tick; start one earlier and on an
    // even segment) and the first ray in fixed point (far outside the domain every ray is in a sentinel band anyway).
    for (int idx = lane; idx < (t_end - t_begin) * n_cams; idx += 32) {
        const int tt = idx / n_cams, k = idx - tt * n_cams;
        const FastCam &Cm = cams[k];
        const uint16_t *IX = D.vc_idx + ((size_t)env * D.Kc + k) * VC_IDX;
        double h = Cm.h0;
        for (int a = 0; a < tt; ++a) h = py_mod360(__dadd_rn(h, Cm.speed));
        const double base = h - Cm.fov * 0.5;
        const int q = max(0, min(VC_IDX - 1, (int)floor(base - Cm.dom_lo)));
        pre_head[tt * D.Kc + k] = h;
        pre_s0[tt * D.Kc + k] = max(0, (int)IX[q] - 1) & ~1;
        pre_fx[tt * D.Kc + k] = (int)fmax(-536870912.0, fmin(536870912.0, floor((base - Cm.dom_lo) * Cm.fx_scale)));
    }
    __syncwarp();
    for (int t = t_begin; t < t_end; ++t) {
        const int pi = (t - t_begin) * D.Kc;
        uint32_t vis[RPL][W];
#pragma unroll
        for (int a = 0; a < RPL; ++a)
#pragma unroll
            for (int w = 0; w < W; ++w) vis[a][w] = 0;
        bool exact_used = false;
        for (int k = 0; k < n_cams; ++k) {
            const FastCam &Cm = cams[k];
            const int sh = Cm.sh, NR = Cm.num_rays + 1, n_gaps = Cm.n_gaps;
            const int bias = ((1 << sh) - 1) - pre_fx[pi + k];   // rays below point p: clamp((p + bias) >> sh, 0, NR)
            const int s0 = pre_s0[pi + k];
            const int2 *P2 = Cm.P2;
            const uint4 *MK4 = Cm.MK4;
            int carry = 0;             // rays below the end of the previous gap
            bool first = s0 > 0;       // the band in front of the first gap looked at lies before the window: no rays
            unsigned bands = 0;
            uint32_t acc[VC_ROWS / 2];
#pragma unroll
            for (int i = 0; i < VC_ROWS / 2; ++i) acc[i] = 0;
            for (int g = (s0 >> 1) + lane;; g += 64) {
                const int2 p0 = __ldg(P2 + min(g, VC_POINTS / 2 - 1));        // padded above n_gaps with a point no ray reaches
                const int2 p1 = __ldg(P2 + min(g + 32, VC_POINTS / 2 - 1));
                uint4 ma0 = make_uint4(0, 0, 0, 0), ma1 = ma0, mb0 = ma0, mb1 = ma0;
                if (g < n_gaps) { ma0 = __ldg(MK4 + 2 * g); ma1 = __ldg(MK4 + 2 * g + 1); }
                if (g + 32 < n_gaps) { mb0 = __ldg(MK4 + 2 * g + 64); mb1 = __ldg(MK4 + 2 * g + 65); }
                {
                    const int lo = max(0, min(NR, (p0.x + bias) >> sh)), hi = max(0, min(NR, (p0.y + bias) >> sh));
                    int ph = __shfl_up_sync(FULL, hi, 1);
                    if (lane == 0) ph = first ? lo : carry;
                    carry = __shfl_sync(FULL, hi, 31);
                    const uint32_t sel = hi > lo ? 0xffffffffu : 0u;   // the gap holds a ray: every ray inside marks the same tiles
                    acc[0] |= ma0.x & sel; acc[1] |= ma0.y & sel; acc[2] |= ma0.z & sel; acc[3] |= ma0.w & sel;
                    acc[4] |= ma1.x & sel; acc[5] |= ma1.y & sel; acc[6] |= ma1.z & sel; acc[7] |= ma1.w & sel;
                    bands |= __ballot_sync(FULL, lo > ph);   // bands that hold a ray (rare): marched exactly below
                    if (carry >= NR) break;  // warp-uniform
                }
                {
                    const int lo = max(0, min(NR, (p1.x + bias) >> sh)), hi = max(0, min(NR, (p1.y + bias) >> sh));
                    int ph = __shfl_up_sync(FULL, hi, 1);
                    if (lane == 0) ph = carry;
                    carry = __shfl_sync(FULL, hi, 31);
                    const uint32_t sel = hi > lo ? 0xffffffffu : 0u;
                    acc[0] |= mb0.x & sel; acc[1] |= mb0.y & sel; acc[2] |= mb0.z & sel; acc[3] |= mb0.w & sel;
                    acc[4] |= mb1.x & sel; acc[5] |= mb1.y & sel; acc[6] |= mb1.z & sel; acc[7] |= mb1.w & sel;
                    bands |= __ballot_sync(FULL, lo > ph);
                    if (carry >= NR) break;
                }
                first = false;
            }
#pragma unroll
            for (int i = 0; i < VC_ROWS / 2; ++i) acc[i] = __reduce_or_sync(FULL, acc[i]);
            // lane = grid row: window row wr of the (now warp-uniform) mask, two 16-bit rows per word
            const int row0 = Cm.row - Cm.range, col0 = Cm.col - Cm.range, nrow = 2 * Cm.range;
#pragma unroll
            for (int a = 0; a < RPL; ++a) {
                const int wr = lane + 32 * a - row0;
                const uint32_t w01 = (wr & 2) ? acc[1] : acc[0], w23 = (wr & 2) ? acc[3] : acc[2];
                const uint32_t w45 = (wr & 2) ? acc[5] : acc[4], w67 = (wr & 2) ? acc[7] : acc[6];
                const uint32_t lo4 = (wr & 4) ? w23 : w01, hi4 = (wr & 4) ? w67 : w45;
                const uint32_t word = (wr & 8) ? hi4 : lo4;
                const unsigned bits = (wr & 1) ? (word >> 16) : (word & 0xffffu);
                if (wr >= 0 && wr <= nrow) fast_or_row<W>(vis[a], bits, col0);
            }
            if (bands) {   // warp-uniform
                cam_exact_scan(vc_geo(D), D.wall + (size_t)env * D.RW, xvis, &Cm, pre_head[pi + k], s0, bias, lane);
                exact_used = true;
            }
        }
        const bool ex = __any_sync(FULL, exact_used);
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
        if (ex) __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// k_seq: the sequential part of a launch, one THREAD per env (everything here is scalar per env: position,
// rewards, guard indices; a warp per env would execute it 32 times over).  Per tick: move, guards advance,
// detection from ONE word of cam_vis[t] OR the guards' cached masks at the Solver's row, vault / timeout,
// rewards, auto-reset.  What k_finish needs to complete the maps -- the guards' (waypoint, heading slot) of every
// tick -- is recorded.
//
// Few warps run this kernel (N / 32), each alone on its SM: its speed is the length of the per-tick dependency
// chain, not throughput.  So nothing on that chain touches memory: the three wall rows around the Solver, each
// guard's state, its current and next patrol word live in registers, and everything tick t + 1 will read from
// global memory -- the cam_vis word(s) and the guards' mask rows at the Solver's next row -- is requested while
// tick t is being decided (until an episode ends, the Solver's path and the patrols do not depend on what is
// seen, so the next state is known); actions are requested four ticks ahead.  Every thread runs exactly T
// iterations (an auto-reset is part of the tick that ended the episode), so the threads of a warp stay in step.
// ---------------------------------------------------------------------------------------------
#define SEQ_THREADS 32
__host__ __device__ inline size_t seq_thread_bytes(int RW, int L) { return (size_t)RW * 4 + (size_t)VC_MAX_GUARDS * L * 4; }

// guard cone + own tile of guard (o = env * Kg + g) at waypoint k with heading slot hs, grid row r
// (visibility.py:44-59) -> OR into the W words of that row
template <int W>
__device__ __forceinline__ void guard_row(const Dev &D, size_t o, int k, int hs, int prow, int pcol, int rng, int r,
                                          uint32_t (&v)[W]) {
    const int wr = r - (prow - rng);
    if (wr < 0 || wr > 2 * rng) return;
    const unsigned bits = D.vg_mask[((o * D.L + k) * (size_t)(D.L + 1) + hs) * VC_ROWS + wr];
    fast_or_row<W>(v, bits, pcol - rng);
}

template <int W>
__global__ void __launch_bounds__(SEQ_THREADS)
k_seq(Dev D, const int8_t *__restrict__ actions, int T, int autoreset, float *__restrict__ reward,
      double *__restrict__ reward64, uint8_t *__restrict__ done, uint8_t *__restrict__ status_out,
      const uint32_t *__restrict__ cam_vis, uint16_t *__restrict__ grec, uint8_t *__restrict__ fin,
      int32_t *__restrict__ last_t, int do_reset, const uint8_t *__restrict__ mask, int store_heading) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;
    const int env = blockIdx.x * SEQ_THREADS + tid;
    if (env >= D.N || !D.env_cached[env]) return;
    if (do_reset && mask && !mask[env]) { last_t[env] = -1; return; }
    constexpr int G = VC_MAX_GUARDS;
    const int R = D.R, C = D.C, N = D.N, RW = D.RW, L = D.L, Kg = D.Kg;
    // shared-memory planes [word][thread]: wall rows, then per guard the patrol words  row | col << 8 | slot << 16
    // (slot: heading slot taken when LEAVING the waypoint, 255 = unchanged)
    uint32_t *wall_s = reinterpret_cast<uint32_t *>(smem) + tid;
    uint32_t *pw_s = wall_s + RW * SEQ_THREADS;

    // ---- load ----
    const int4 es = *reinterpret_cast<const int4 *>(D.env_s + (size_t)env * 4);
    const int n_cams = es.x, n_guards = es.y;
    const int4 d0 = *reinterpret_cast<const int4 *>(D.env_d + (size_t)env * 8);
    const int4 d1 = *reinterpret_cast<const int4 *>(D.env_d + (size_t)env * 8 + 4);
    EnvRegs E;
    E.r = d0.x & 0xffff; E.c = d0.x >> 16; E.tick = d0.y; E.prev = d0.z; E.init = d0.w;
    E.flags = d1.x & 0xff; E.n_vault = d1.y; E.n_detect = d1.z; E.n_timeout = d1.w;
    for (int i = 0; i < RW; ++i) wall_s[i * SEQ_THREADS] = D.wall[(size_t)env * RW + i];
    int gk[G], ghs[G], glen[G], gstp[G], grng[G], gkn[G];   // waypoint, heading slot, path length, stride, range, next waypoint
    unsigned gw[G], gwn[G], gw0[G];                         // patrol word at gk, at gkn, at waypoint 0
    const uint16_t *gmask[G];                               // the guard's mask table
#pragma unroll
    for (int g = 0; g < G; ++g) {
        gk[g] = ghs[g] = gkn[g] = 0; glen[g] = 1; gstp[g] = 0; grng[g] = 0; gw[g] = gwn[g] = gw0[g] = 0; gmask[g] = D.vg_mask;
        if (g < n_guards) {
            const size_t o = (size_t)env * Kg + g;
            const int4 gi = *reinterpret_cast<const int4 *>(D.guard_i + o * 4);   // len, speed, range, num_rays
            glen[g] = gi.x; gstp[g] = gi.x >= 2 ? py_imod(gi.y, gi.x) : 0; grng[g] = gi.z;
            gk[g] = D.guard_idx[o];
            for (int k = 0; k < gi.x; ++k)
                pw_s[(g * L + k) * SEQ_THREADS] = (unsigned)D.guard_path[(o * L + k) * 2] | ((unsigned)D.guard_path[(o * L + k) * 2 + 1] << 8) |
                                                  ((unsigned)D.vg_hslot[o * L + k] << 16);
            // heading -> slot.  A heading that is none of the path's can only have been written by hand into the
            // state view; it is reported (ERR_STATE) and treated as the default heading.
            const double h = D.guard_heading[o];
            int hs = -1;
            const int nh = D.vg_nh[o];
            for (int s = 0; s < nh; ++s)
                if (__double_as_longlong(D.vg_hval[o * (L + 1) + s]) == __double_as_longlong(h)) { hs = s; break; }
            if (hs < 0) { atomicOr(D.err, ERR_STATE); hs = 0; }
            ghs[g] = hs;
            gmask[g] = D.vg_mask + o * L * (size_t)(L + 1) * VC_ROWS;
            gkn[g] = gk[g] + gstp[g]; if (gkn[g] >= glen[g]) gkn[g] -= glen[g];
            gw[g] = pw_s[(g * L + gk[g]) * SEQ_THREADS]; gwn[g] = pw_s[(g * L + gkn[g]) * SEQ_THREADS]; gw0[g] = pw_s[(g * L) * SEQ_THREADS];
        }
    }
    // wall rows E.r - 1, E.r, E.r + 1 (rows outside the grid block)
    uint32_t ww[3][W];
#define SEQ_WALL_ROW(dst, row)                                                                                 \
    do {                                                                                                       \
        const int r_ = (row);                                                                                  \
        _Pragma("unroll") for (int w = 0; w < W; ++w)                                                          \
            (dst)[w] = (r_ >= 0 && r_ < R) ? wall_s[(r_ * W + w) * SEQ_THREADS] : 0xffffffffu;                 \
    } while (0)
#define SEQ_WALL_AROUND() do { SEQ_WALL_ROW(ww[0], E.r - 1); SEQ_WALL_ROW(ww[1], E.r); SEQ_WALL_ROW(ww[2], E.r + 1); } while (0)
    // would the move (dr, dc) from (E.r, E.c) be accepted (:239-246)?  Uses the register rows only.
    auto wall_word = [&](int dr, int nc) -> uint32_t {   // selects, not indexing: the rows stay in registers
        uint32_t lo = dr < 0 ? ww[0][0] : (dr > 0 ? ww[2][0] : ww[1][0]);
        if (W == 2) { const uint32_t hi = dr < 0 ? ww[0][W - 1] : (dr > 0 ? ww[2][W - 1] : ww[1][W - 1]); if (nc >> 5) lo = hi; }
        return lo;
    };
#define SEQ_FREE(dr, nc) ((nc) >= 0 && (nc) < C && !((wall_word((dr), (nc)) >> ((nc) & 31)) & 1u))
    SEQ_WALL_AROUND();

    int status = HEIST_RUNNING;
    int n_adv = 0;           // camera updates executed by this launch
    int last = -1;           // last tick of this launch whose visibility map was rebuilt
    // HeistEnvironment.reset (environment.py:183-214): headings persist, guards back to waypoint 0.  The map after a
    // reset is completed by k_finish from the recorded guard state; nothing is detected on a reset.
#define SEQ_RESET_STATE()                                                                    \
    do {                                                                                     \
        E.r = D.start_r; E.c = D.start_c; E.tick = 0; E.flags = 0;                           \
        E.prev = abs(E.r - D.vault_r) + abs(E.c - D.vault_c); E.init = E.prev;               \
        SEQ_WALL_AROUND();                                                                   \
        _Pragma("unroll") for (int g = 0; g < G; ++g) {                                      \
            gk[g] = 0; gw[g] = gw0[g]; gkn[g] = gstp[g];                                     \
            gwn[g] = pw_s[(g * L + gkn[g]) * SEQ_THREADS];                                   \
        }                                                                                    \
    } while (0)
#define SEQ_RECORD(o_)                                                                                         \
    do {                                                                                                       \
        _Pragma("unroll") for (int g = 0; g < G; ++g)                                                          \
            if (g < n_guards) grec[(o_) * Kg + g] = (uint16_t)(gk[g] | (ghs[g] << 8));                         \
    } while (0)
    if (do_reset) {
        SEQ_RESET_STATE();
        fin[env] = 1; SEQ_RECORD((size_t)env); last = 0;
        T = 0;
    }
    // actions: four registers used round-robin (slot t & 3 holds tick t and is refilled for tick t + 4 as soon as
    // it has been read: no register is read in the iteration that follows its load -- they stream from HBM, one new
    // sector per tick, and a load takes longer than a tick)
    int a_cur = 0, a_q0 = 0, a_q1 = 0, a_q2 = 0, a_q3 = 0;
    if (T > 0) {
        a_cur = actions[env];
        if (1 < T) a_q1 = actions[(size_t)1 * N + env];
        if (2 < T) a_q2 = actions[(size_t)2 * N + env];
        if (3 < T) a_q3 = actions[(size_t)3 * N + env];
        if (4 < T) a_q0 = actions[(size_t)4 * N + env];
    }
    int pre_row = -1;
    uint32_t pre[W];
    unsigned pre_g[G], pre_key[G];   // mask row / (waypoint | slot << 8) it was fetched for
#pragma unroll
    for (int w = 0; w < W; ++w) pre[w] = 0;
#pragma unroll
    for (int g = 0; g < G; ++g) { pre_g[g] = 0; pre_key[g] = 0xffffffffu; }
    for (int t = 0; t < T; ++t) {
        const size_t o = (size_t)t * N + env;
        bool rebuilt = false;
        double rw = 0.0;
        status = HEIST_ALREADY_DONE;
        if (!(E.flags & F_DONE)) {   // a done env is not mutated (:232-233)
            // move (:239-246): blocked by the grid edge or a WALL tile
            const int dr = (a_cur == 2) - (a_cur == 1), nc = E.c + (a_cur == 4) - (a_cur == 3);
            if (SEQ_FREE(dr, nc)) {
                E.c = nc;
                if (dr) {   // rows shift by one; the new outer row is only needed from the next tick on
                    E.r += dr;
#pragma unroll
                    for (int w = 0; w < W; ++w) {
                        if (dr > 0) { ww[0][w] = ww[1][w]; ww[1][w] = ww[2][w]; } else { ww[2][w] = ww[1][w]; ww[1][w] = ww[0][w]; }
                    }
                    if (dr > 0) SEQ_WALL_ROW(ww[2], E.r + 1); else SEQ_WALL_ROW(ww[0], E.r - 1);
                }
            }
            ++n_adv;   // cameras rotate (:251-252): their cones for this tick are cam_vis[t]
#pragma unroll
            for (int g = 0; g < G; ++g) {   // Guard.update (security.py:145-159)
                if (g < n_guards && glen[g] >= 2) {
                    const int hsl = (gw[g] >> 16) & 255;
                    if (hsl != 255) ghs[g] = hsl;   // 255: the move is (0, 0), heading unchanged
                    gk[g] = gkn[g]; gw[g] = gwn[g];
                    gkn[g] = gk[g] + gstp[g]; if (gkn[g] >= glen[g]) gkn[g] -= glen[g];
                    gwn[g] = pw_s[(g * L + gkn[g]) * SEQ_THREADS];
                }
            }
            // visibility at the Solver's tile: camera cones OR guard cones / own tiles
            uint32_t v[W];
            const bool hit = E.r == pre_row;
            if (hit) {
#pragma unroll
                for (int w = 0; w < W; ++w) v[w] = pre[w];
            } else {
#pragma unroll
                for (int w = 0; w < W; ++w) v[w] = cam_vis[o * RW + E.r * W + w];
            }
#pragma unroll
            for (int g = 0; g < G; ++g) {
                if (g < n_guards) {
                    const int prow = gw[g] & 255, pcol = (gw[g] >> 8) & 255, wr = E.r - (prow - grng[g]);
                    if (wr >= 0 && wr <= 2 * grng[g]) {
                        const unsigned bits = (hit && pre_key[g] == (unsigned)(gk[g] | (ghs[g] << 8)))
                                                  ? pre_g[g] : gmask[g][(gk[g] * (L + 1) + ghs[g]) * VC_ROWS + wr];
                        fast_or_row<W>(v, bits, pcol - grng[g]);
                    }
                }
            }
            const bool detected = (v[(W == 2) ? (E.c >> 5) : 0] >> (E.c & 31)) & 1u;
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
        if (reward) reward[o] = (float)rw;
        if (reward64) reward64[o] = rw;
        if (done) done[o] = (E.flags & F_DONE) ? 1 : 0;
        if (status_out) status_out[o] = (uint8_t)status;
        if (autoreset && (E.flags & F_DONE)) { SEQ_RESET_STATE(); rebuilt = true; }   // the trainer's `if done: reset()`
        // the visibility map of tick t is final: tell k_finish how to complete it
        fin[o] = rebuilt ? 1 : 0;
        if (rebuilt) { last = t; SEQ_RECORD(o); }
        // requests for tick t + 1
        {   // a_cur <- tick t + 1 (slot (t + 1) & 3), then that slot <- tick t + 5
            const bool refill = t + 5 < T;
            const int8_t *nxt = actions + o + (size_t)5 * N;
            switch ((t + 1) & 3) {   // warp-uniform
                case 0: a_cur = a_q0; if (refill) a_q0 = *nxt; break;
                case 1: a_cur = a_q1; if (refill) a_q1 = *nxt; break;
                case 2: a_cur = a_q2; if (refill) a_q2 = *nxt; break;
                default: a_cur = a_q3; if (refill) a_q3 = *nxt; break;
            }
        }
        pre_row = -1;
        if (t + 1 < T && !(E.flags & F_DONE)) {   // tick t + 1 will be a step from exactly this state
            const int dr = (a_cur == 2) - (a_cur == 1), nc = E.c + (a_cur == 4) - (a_cur == 3);
            pre_row = SEQ_FREE(dr, nc) ? E.r + dr : E.r;
#pragma unroll
            for (int w = 0; w < W; ++w) pre[w] = cam_vis[(o + N) * RW + pre_row * W + w];
#pragma unroll
            for (int g = 0; g < G; ++g) {
                if (g < n_guards) {
                    int k = gk[g], hs = ghs[g];
                    unsigned word = gw[g];
                    if (glen[g] >= 2) {
                        const int hsl = (word >> 16) & 255;
                        if (hsl != 255) hs = hsl;
                        k = gkn[g]; word = gwn[g];
                    }
                    pre_key[g] = (unsigned)(k | (hs << 8));
                    const int wr = pre_row - ((int)(word & 255) - grng[g]);
                    pre_g[g] = 0;
                    if (wr >= 0 && wr <= 2 * grng[g]) pre_g[g] = gmask[g][(k * (L + 1) + hs) * VC_ROWS + wr];
                }
            }
        }
    }
#undef SEQ_RESET_STATE
#undef SEQ_RECORD
#undef SEQ_WALL_ROW
#undef SEQ_WALL_AROUND
#undef SEQ_FREE

    // ---- store ----
    last_t[env] = last;
    *reinterpret_cast<int4 *>(D.env_d + (size_t)env * 8) = make_int4(E.r | (E.c << 16), E.tick, E.prev, E.init);
    *reinterpret_cast<int4 *>(D.env_d + (size_t)env * 8 + 4) =
        make_int4(E.flags | (status << 8), E.n_vault, E.n_detect, E.n_timeout);
    for (int k = 0; k < (store_heading ? n_cams : 0); ++k) {   // headings after the camera updates this launch executed
        const size_t co = (size_t)env * D.Kc + k;              // (otherwise stored by k_heads)
        double h = D.cam_heading[co];
        const double speed = D.cam_f[co * 2 + 1];
        for (int a = 0; a < n_adv; ++a) h = py_mod360(__dadd_rn(h, speed));
        D.cam_heading[co] = h;
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
        if (g < n_guards) {
            const size_t go = (size_t)env * Kg + g;
            D.guard_heading[go] = D.vg_hval[go * (L + 1) + ghs[g]];
            D.guard_idx[go] = gk[g];
        }
    }
}

// k_finish: warp per (env, block of FIN_TB ticks), lane = grid row: visibility row = cam_vis row OR the guards'
// masks at the state k_seq recorded -> written in place (the buffer is the caller's trajectory, or scratch) and,
// for the last rebuilt tick of an env, to its current map D.vis.  The env's patrol words sit in registers (lane k
// holds waypoint k of every guard; max_path <= 32), so a tick costs one shuffle pair and one 2-byte mask load
// per guard.  only_last: no trajectory wanted -- one tick per env.
// Ticks an env spent done without auto-reset (fin == 0) keep the env's current map; they are filled by k_fill
// once D.vis is final.
#define FIN_TB 8
template <int RPL, int W>
__global__ void __launch_bounds__(256)
k_finish(Dev D, int T, uint32_t *buf, const uint16_t *__restrict__ grec, const uint8_t *__restrict__ fin,
         const int32_t *__restrict__ last_t, int only_last, const uint8_t *__restrict__ mask) {
    const int lane = threadIdx.x & 31;   // grid: x = env / 8, y = tick block
    const int env = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (env >= D.N || !D.env_cached[env] || (mask && !mask[env])) return;
    const int lt = last_t[env];
    int t0 = blockIdx.y * FIN_TB, t1 = min(T, t0 + FIN_TB);
    if (only_last) { t0 = lt; t1 = lt + 1; }
    if (t0 < 0) return;
    constexpr int G = VC_MAX_GUARDS;
    const int n_guards = D.env_s[(size_t)env * 4 + 1], L = D.L, Kg = D.Kg, RW = D.RW;
    unsigned pw[G];            // lane k: waypoint k of guard g  (row | col << 8 | range << 16)
    const uint16_t *gmask[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
        pw[g] = 0; gmask[g] = D.vg_mask;
        if (g < n_guards) {
            const size_t go = (size_t)env * Kg + g;
            if (lane < L) pw[g] = (unsigned)D.guard_path[(go * L + lane) * 2] | ((unsigned)D.guard_path[(go * L + lane) * 2 + 1] << 8) |
                                  ((unsigned)D.guard_i[go * 4 + 2] << 16);
            gmask[g] = D.vg_mask + go * L * (size_t)(L + 1) * VC_ROWS;
        }
    }
    for (int t = t0; t < t1; ++t) {
        const size_t o = (size_t)t * D.N + env;
        const unsigned f = fin[o];
        unsigned rec = 0;
        if (lane < n_guards) rec = grec[o * Kg + lane];
        uint32_t v[RPL][W];
#pragma unroll
        for (int a = 0; a < RPL; ++a)
#pragma unroll
            for (int w = 0; w < W; ++w) { const int r = lane + 32 * a; v[a][w] = r < D.R ? buf[o * RW + r * W + w] : 0u; }
        if (!f) continue;   // warp-uniform
#pragma unroll
        for (int g = 0; g < G; ++g) {
            if (g < n_guards) {
                const unsigned rg = __shfl_sync(0xffffffffu, rec, g);
                const int k = rg & 255, hs = rg >> 8;
                const unsigned word = __shfl_sync(0xffffffffu, pw[g], k);
                const int prow = word & 255, pcol = (word >> 8) & 255, rng = word >> 16;
#pragma unroll
                for (int a = 0; a < RPL; ++a) {
                    const int wr = lane + 32 * a - (prow - rng);
                    if (wr >= 0 && wr <= 2 * rng) fast_or_row<W>(v[a], gmask[g][(k * (L + 1) + hs) * VC_ROWS + wr], pcol - rng);
                }
            }
        }
#pragma unroll
        for (int a = 0; a < RPL; ++a) {
            const int r = lane + 32 * a;
            if (r < D.R) {
#pragma unroll
                for (int w = 0; w < W; ++w) {
                    if (!only_last) buf[o * RW + r * W + w] = v[a][w];
                    if (t == lt) D.vis[(size_t)env * RW + r * W + w] = v[a][w];
                }
            }
        }
    }
}

// k_fill (no auto-reset only): ticks an env spent done copy its current map, final once k_finish has run.
template <int W>
__global__ void __launch_bounds__(256)
k_fill(Dev D, int T, uint32_t *buf, const uint8_t *__restrict__ fin) {
    const int lane = threadIdx.x & 31;   // grid: x = env / 8, y = tick; lane strides over the RW words
    const int env = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (env >= D.N || !D.env_cached[env]) return;
    const size_t te = (size_t)blockIdx.y * D.N + env;
    if (fin[te]) return;
    for (int i = lane; i < D.RW; i += 32) buf[te * D.RW + i] = D.vis[(size_t)env * D.RW + i];
}
