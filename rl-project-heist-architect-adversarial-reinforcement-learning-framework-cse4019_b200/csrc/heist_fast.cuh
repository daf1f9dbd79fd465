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

#define FAST_WARPS 4

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

#define FAST_TB 8      // ticks per k_cam_vis warp

// Per-camera constants of a k_cam_vis warp (shared memory; a multiple of 16 bytes: the staged tables follow it).
struct alignas(16) FastCam {
    double speed, fov, fx_scale, dom_lo;
    double h0;              // heading at the block's first tick
    const int2 *P2;         // gap g = boundary points 2g (its start), 2g + 1 (its end), fixed point (heist_cache.cuh)
    const uint4 *MK4;       // ... and its window mask: uint4 2g, 2g + 1
    int row, col, range, num_rays, n_gaps, sh;
    int poff, moff;         // k_cam_vis_staged: where the table sits in the CTA's staging buffers (entries)
};

// Asynchronous global -> shared copies (LDGSTS): fire and forget, no register staging, so a burst of them is one
// memory round trip instead of one per iteration.
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void *dst_smem, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void or_and(uint32_t &acc, uint32_t m, uint32_t sel) {   // acc |= m & sel, one LOP3
    asm("lop3.b32 %0, %0, %1, %2, 0xF8;" : "+r"(acc) : "r"(m), "r"(sel));
}

__host__ __device__ inline size_t camvis_warp_bytes(int RW, int Kc) {
    return (size_t)Kc * sizeof(FastCam) + (((size_t)RW * 4 + 15) & ~(size_t)15) + (size_t)FAST_TB * Kc * 16;
}

// heads[b][env][k] = heading of camera k at the first tick of tick block b (FAST_TB ticks) of this stretch of T ticks;
// k_cam_vis advances it through the block.  A launch is cut into stretches (the pipelined chunks): `first` starts
// from the stored heading, later stretches continue from h_run, where each stretch leaves the heading of its next
// tick.  write_final: with auto-reset every tick of the launch updates the cameras (environment.py:251-252), so the
// heading the launch ends on is the one of its last tick and is stored here; otherwise k_seq / k_walk store it (an
// env may stop stepping early).
__global__ void __launch_bounds__(128) k_heads(Dev D, int T, int do_reset, int first, int write_final, double *__restrict__ heads,
                                               double *__restrict__ h_run) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= D.N * D.Kc) return;
    const int env = i / D.Kc, k = i - env * D.Kc;
    if (!D.env_cached[env] || k >= D.env_s[(size_t)env * 4]) return;
    const double speed = D.cam_f[(size_t)i * 2 + 1];
    double h;
    if (first) {
        h = D.cam_heading[i];
        const int adv0 = fast_adv0(D, env, do_reset);
        for (int a = 0; a < adv0; ++a) h = py_mod360(__dadd_rn(h, speed));
    } else h = h_run[i];
    double last = h;
    for (int t = 0; t < T; ++t) {
        if (t % FAST_TB == 0) heads[(size_t)(t / FAST_TB) * D.N * D.Kc + i] = h;
        last = h;
        h = py_mod360(__dadd_rn(h, speed));
    }
    if (h_run) h_run[i] = h;
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

// One (tick, camera) from the tables in global memory: OR of the masks of the gaps that hold a ray -> acc[8]
// (warp-reduced, i.e. identical in every lane); returns non-zero when a band holds a ray (see cam_exact_scan).
__device__ __forceinline__ unsigned scan_window(const int2 *__restrict__ P2, const uint4 *__restrict__ MK4, int n_gaps, int s0, int bias,
                                                int sh, int NR, int lane, uint32_t (&acc)[VC_ROWS / 2]) {
    constexpr unsigned FULL = 0xffffffffu;
    int carry = 0;             // rays below the end of the previous gap
    bool first = s0 > 0;       // the band in front of the first gap looked at lies before the window: no rays
    unsigned bands = 0;
#pragma unroll
    for (int i = 0; i < VC_ROWS / 2; ++i) acc[i] = 0;
    for (int g = (s0 >> 1) + lane;; g += 64) {
        const int g0 = min(g, VC_POINTS / 2 - 1), g1 = min(g + 32, VC_POINTS / 2 - 1);   // padded above n_gaps with a point no ray reaches
        const int2 p0 = __ldg(P2 + g0), p1 = __ldg(P2 + g1);
        const uint4 ma0 = __ldg(MK4 + 2 * g0), ma1 = __ldg(MK4 + 2 * g0 + 1), mb0 = __ldg(MK4 + 2 * g1), mb1 = __ldg(MK4 + 2 * g1 + 1);
        const int lo0 = max(0, min(NR, (p0.x + bias) >> sh)), hi0 = max(0, min(NR, (p0.y + bias) >> sh));
        const int lo1 = max(0, min(NR, (p1.x + bias) >> sh)), hi1 = max(0, min(NR, (p1.y + bias) >> sh));
        int ph0 = __shfl_up_sync(FULL, hi0, 1), ph1 = __shfl_up_sync(FULL, hi1, 1);
        const int c0 = __shfl_sync(FULL, hi0, 31), c1 = __shfl_sync(FULL, hi1, 31);
        if (lane == 0) { ph0 = first ? lo0 : carry; ph1 = c0; }
        const uint32_t sel0 = (hi0 > lo0 && g < n_gaps) ? 0xffffffffu : 0u;   // the gap holds a ray: every ray inside marks the same tiles
        const uint32_t sel1 = (hi1 > lo1 && g + 32 < n_gaps) ? 0xffffffffu : 0u;
        or_and(acc[0], ma0.x, sel0); or_and(acc[1], ma0.y, sel0); or_and(acc[2], ma0.z, sel0); or_and(acc[3], ma0.w, sel0);
        or_and(acc[4], ma1.x, sel0); or_and(acc[5], ma1.y, sel0); or_and(acc[6], ma1.z, sel0); or_and(acc[7], ma1.w, sel0);
        or_and(acc[0], mb0.x, sel1); or_and(acc[1], mb0.y, sel1); or_and(acc[2], mb0.z, sel1); or_and(acc[3], mb0.w, sel1);
        or_and(acc[4], mb1.x, sel1); or_and(acc[5], mb1.y, sel1); or_and(acc[6], mb1.z, sel1); or_and(acc[7], mb1.w, sel1);
        bands |= __ballot_sync(FULL, (lo0 > ph0) | (lo1 > ph1));   // bands that hold a ray (rare): marched exactly
        carry = c1;
        first = false;
        if (carry >= NR) break;  // warp-uniform
    }
#pragma unroll
    for (int i = 0; i < VC_ROWS / 2; ++i) acc[i] = __reduce_or_sync(FULL, acc[i]);
    return bands;
}

// lane = grid row (rows lane, lane + 32): OR window row wr of the warp-uniform mask acc (two 16-bit rows per word,
// window origin (row0, col0), rows 0..nrow) into the lane's row words
template <int RPL, int W>
__device__ __forceinline__ void place_rows(const uint32_t (&acc)[VC_ROWS / 2], int row0, int col0, int nrow, int lane, uint32_t (&rows)[RPL][W]) {
#pragma unroll
    for (int a = 0; a < RPL; ++a) {
        const int wr = lane + 32 * a - row0;
        const uint32_t w01 = (wr & 2) ? acc[1] : acc[0], w23 = (wr & 2) ? acc[3] : acc[2];
        const uint32_t w45 = (wr & 2) ? acc[5] : acc[4], w67 = (wr & 2) ? acc[7] : acc[6];
        const uint32_t lo4 = (wr & 4) ? w23 : w01, hi4 = (wr & 4) ? w67 : w45;
        const uint32_t word = (wr & 8) ? hi4 : lo4;
        const unsigned bits = (wr & 1) ? (word >> 16) : (word & 0xffffu);
        if (wr >= 0 && wr <= nrow) fast_or_row<W>(rows[a], bits, col0);
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
        { const double2 lf = *reinterpret_cast<const double2 *>(D.vc_lo + o * 2); Cm.dom_lo = lf.x; Cm.fx_scale = lf.y; }
        Cm.n_gaps = D.vc_meta[o * 2] >> 1;
        Cm.sh = D.vc_meta[o * 2 + 1];
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
            const int bias = ((1 << Cm.sh) - 1) - pre_fx[pi + k];   // rays below point p: clamp((p + bias) >> sh, 0, NR)
            const int s0 = pre_s0[pi + k];
            uint32_t acc[VC_ROWS / 2];
            const unsigned bands = scan_window(Cm.P2, Cm.MK4, Cm.n_gaps, s0, bias, Cm.sh, Cm.num_rays + 1, lane, acc);
            place_rows<RPL, W>(acc, Cm.row - Cm.range, Cm.col - Cm.range, 2 * Cm.range, lane, vis);
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
// k_cam_vis_staged: the many-tick variant.  A CTA owns one env and the tick blocks of one pipelined chunk (up to 8:
// cameras turn 5-35 degrees per tick, so such a stretch sweeps a camera's whole table); warp = tick block.  The CTA
// STAGES the tables of as many of the env's cameras as fit -- boundary points and gap masks, ~4 KB per camera, one
// burst of cp.async that is in flight while the warps compute their ticks' headings and window starts -- in shared
// memory, then every warp scans the windows of its 8 ticks from there, lane = (tick, quarter of the window's gaps).
// The scan's loads are shared-memory loads (~30 cycles instead of an L2 / HBM round trip per (tick, camera)) and a
// table is read from HBM once per chunk.  Rows are accumulated per tick in registers (one-word grids, lane = grid
// row) or shared memory.  Compiled for six CTAs per SM (40 registers): the kernel is latency-bound, occupancy pays.
// ---------------------------------------------------------------------------------------------
#define CVS_P2 640     // staged boundary-point pairs: VC_POINTS / 2 real slots + padding a scan can run into
#define CVS_GPL 4      // gaps a lane takes per pass (4 lanes per tick: 16 consecutive gaps per tick and pass; a window of the
                       // bench workload spans ~25, and what a pass covers beyond the window's end is wasted: 8 -> 4 = +3 %)

__host__ __device__ inline size_t camvis_staged_warp_bytes(int RW, int Kc) {
    // h0_s f64[Kc] | pre_fx i32[8][Kc] | pre_s0 u16[8][Kc] | (16-byte aligned) mask_s 8 x 32 B | vis_s u32[8][RW]
    return ((((size_t)Kc * (8 + FAST_TB * 6)) + 15) & ~(size_t)15) + (size_t)FAST_TB * 32 + (((size_t)FAST_TB * RW * 4 + 15) & ~(size_t)15);
}
#define CVS_MAX_WARPS 8
__host__ __device__ inline size_t camvis_staged_bytes(int RW, int Kc, int warps) {
    return (size_t)Kc * sizeof(FastCam) + (size_t)CVS_P2 * 8 + (size_t)(VC_POINTS / 2) * 32 +
           (size_t)warps * camvis_staged_warp_bytes(RW, Kc);
}

// cam_exact_scan for a staged table (same scan, boundary points from shared memory)
__device__ __noinline__ void cam_exact_scan_staged(VcGeo D, const uint32_t *__restrict__ wall, uint32_t *xvis, const FastCam *cam,
                                                   const int2 *P2s, double heading, int s0, int bias, int lane) {
    const int sh = cam->sh, NR = cam->num_rays + 1;
    int carry = 0;
    bool first = s0 > 0;
    for (int g = (s0 >> 1) + lane;; g += 32) {
        const int2 p = P2s[g];
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

template <int RPL, int W>
__global__ void __launch_bounds__(CVS_MAX_WARPS * 32, 6)
k_cam_vis_staged(Dev D, int T, int nblk, const double *__restrict__ heads, uint32_t *__restrict__ out, int rev) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nthr = blockDim.x;
    // rev: the chunks of a launch walk the envs in alternating directions -- what the previous chunk read last (the
    // tables of the envs it ended on) is what the L2 still holds when the next one starts
    const int env = rev ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x, b = blockIdx.y * (nthr >> 5) + warp;
    if (!D.env_cached[env]) return;   // CTA-uniform
    const int Kc = D.Kc, RW = D.RW;
    unsigned off = 0;   // (32-bit offsets: the carving is a handful of integer adds)
    auto take = [&](unsigned bytes) { unsigned char *p = smem + off; off += bytes; return p; };
    FastCam *cams = reinterpret_cast<FastCam *>(take((unsigned)Kc * (unsigned)sizeof(FastCam)));
    int2 *P2s = reinterpret_cast<int2 *>(take(CVS_P2 * 8));
    uint4 *M0s = reinterpret_cast<uint4 *>(take((VC_POINTS / 2) * 16));   // first / second half of the masks,
    uint4 *M1s = reinterpret_cast<uint4 *>(take((VC_POINTS / 2) * 16));   // split: conflict-free LDS.128
    const unsigned vis_bytes = ((unsigned)FAST_TB * (unsigned)RW * 4u + 15u) & ~15u;
    const unsigned pre_bytes = ((unsigned)Kc * (8u + FAST_TB * 6u) + 15u) & ~15u;
    off += (unsigned)warp * (pre_bytes + FAST_TB * 32u + vis_bytes);   // = camvis_staged_warp_bytes
    unsigned char *pre_base = take(pre_bytes);
    double *h0_s = reinterpret_cast<double *>(pre_base);                                            // heading at the block's first tick
    int *pre_fx = reinterpret_cast<int *>(pre_base + (unsigned)Kc * 8u);                            // [tick][camera] first ray, fixed point
    uint16_t *pre_s0 = reinterpret_cast<uint16_t *>(pre_base + (unsigned)Kc * (8u + FAST_TB * 4u));  // [tick][camera] window start (point index)
    uint32_t *mask_s = reinterpret_cast<uint32_t *>(take(FAST_TB * 32));                  // [tick] window mask of the camera in hand
    uint32_t *vis_s = reinterpret_cast<uint32_t *>(take(vis_bytes));                      // [tick][RW]
    const int n_cams = D.env_s[(size_t)env * 4];
    if (tid < n_cams) {
        const size_t o = (size_t)env * Kc + tid;
        FastCam &Cm = cams[tid];
        const int16_t *ci = D.cam_i + o * 4;
        Cm.fov = D.cam_f[o * 2]; Cm.speed = D.cam_f[o * 2 + 1];
        Cm.h0 = 0.0;
        Cm.row = ci[0]; Cm.col = ci[1]; Cm.range = ci[2]; Cm.num_rays = ci[3];
        { const double2 lf = *reinterpret_cast<const double2 *>(D.vc_lo + o * 2); Cm.dom_lo = lf.x; Cm.fx_scale = lf.y; }
        Cm.n_gaps = D.vc_meta[o * 2] >> 1;
        Cm.sh = D.vc_meta[o * 2 + 1];
        Cm.P2 = reinterpret_cast<const int2 *>(D.vc_p + o * VC_POINTS);
        Cm.MK4 = reinterpret_cast<const uint4 *>(D.vc_mask + o * (size_t)(VC_POINTS / 2) * VC_ROWS);
    }
    const bool active = b < nblk;
    const int t_begin = b * FAST_TB, t_end = active ? min(T, (b + 1) * FAST_TB) : t_begin;
    const int n_t = t_end - t_begin;
    if (active && lane < n_cams) h0_s[lane] = heads[((size_t)b * D.N + env) * Kc + lane];
    if (!(RPL == 1 && W == 1))   // (one-word grids keep their rows in registers; vis_s then only serves the exact path)
        for (int i = lane; i < FAST_TB * RW; i += 32) vis_s[i] = 0;
    __syncthreads();
    uint32_t vr[FAST_TB];   // one-word grids: the ticks' rows (lane = grid row) stay in registers
#pragma unroll
    for (int t2 = 0; t2 < FAST_TB; ++t2) vr[t2] = 0;
    bool any_exact = false;   // exactly marched rays put their tiles into vis_s
    // The tables of as many cameras as fit are staged TOGETHER (typically all of an env's: ~100 gaps each of 512
    // slots), back to back: one burst of copies and one barrier pair per group instead of per camera.
    auto stage_group = [&](int k0) -> int {   // issue the copies of cameras k0 .. k1 - 1 (as many as fit); returns k1
        int k1 = k0;
        for (int poff = 0, moff = 0; k1 < n_cams; ++k1) {
            const FastCam &Cs = cams[k1];
            const int n = Cs.n_gaps, need_p = ((n + 1) & ~1) + 96, need_m = max(n, 1);   // points: + the padding a scan can run into
            if (k1 > k0 && (poff + need_p > CVS_P2 || moff + need_m > VC_POINTS / 2)) break;
            const uint4 *src = reinterpret_cast<const uint4 *>(Cs.P2);
            uint4 *dst = reinterpret_cast<uint4 *>(P2s + poff);
            uint4 *m0 = M0s + moff, *m1 = M1s + moff;
            for (int i = tid; i < (n + 1) / 2; i += nthr) cp_async16(dst + i, src + i);
            for (int i = tid; i < 2 * n; i += nthr) cp_async16(((i & 1) ? m1 : m0) + (i >> 1), Cs.MK4 + i);
            for (int i = ((n + 1) & ~1) + tid; i < need_p; i += nthr) P2s[poff + i] = make_int2(0x3fffffff, 0x3fffffff);   // no ray reaches it
            if (tid == 0) { cams[k1].poff = poff; cams[k1].moff = moff; }
            poff += need_p; moff += need_m;
        }
        return k1;
    };
    int k0 = 0, k1 = stage_group(0);   // (cams visible: barrier above) -- the copies fly while the pre-phase computes
    // Per (tick, camera) of the block, in parallel lanes: heading, window start from the coarse index, first ray in
    // fixed point (see k_cam_vis).
    for (int idx = lane; idx < n_t * n_cams; idx += 32) {
        const int tt = idx / n_cams, k = idx - tt * n_cams;
        const FastCam &Cm = cams[k];
        const uint16_t *IX = D.vc_idx + ((size_t)env * Kc + k) * VC_IDX;
        double h = h0_s[k];
        for (int a = 0; a < tt; ++a) h = py_mod360(__dadd_rn(h, Cm.speed));
        const double base = h - Cm.fov * 0.5;
        const int q = max(0, min(VC_IDX - 1, (int)floor(base - Cm.dom_lo)));
        pre_s0[tt * Kc + k] = (uint16_t)(max(0, (int)IX[q] - 1) & ~1);
        pre_fx[tt * Kc + k] = (int)fmax(-536870912.0, fmin(536870912.0, floor((base - Cm.dom_lo) * Cm.fx_scale)));
    }
    while (k0 < n_cams) {   // one group of staged tables per trip
        cp_async_wait_all();
        __syncthreads();
        for (int k = k0; k < k1 && active; ++k) {   // (no barrier inside: idle warps go straight to the next group's barrier)
            const FastCam &Cm = cams[k];
            const int n_gaps = max(Cm.n_gaps, 1), n_pad = ((Cm.n_gaps + 1) & ~1) + 96;
            const int2 *P2c = P2s + Cm.poff;
            const uint4 *M0c = M0s + Cm.moff, *M1c = M1s + Cm.moff;
            const int sh = Cm.sh, NR = Cm.num_rays + 1;
            const int row0 = Cm.row - Cm.range, col0 = Cm.col - Cm.range, nrow = 2 * Cm.range;
            // The scan.  With the table in shared memory, per-lane addressing is cheap, so the warp scans the windows of ALL
            // its ticks at once: lane = (tick tt, q); in a pass the four lanes of a tick take the gaps gbase + 4 i + q,
            // i < CVS_GPL (16 consecutive gaps per tick and pass, the four lanes on four neighbouring table entries).  Per
            // gap: ray counts below its two boundary points from the tick's first-ray position; the gap holds a ray iff
            // they differ (its 32-byte mask is OR-ed in, branch-free).  Rays inside BANDS are found by counting: every ray
            // lies either in a gap or in a band, so the bands hold one iff the gaps' counts do not add up to all rays.
            const int tt = lane >> 2, q = lane & 3;
            const bool have = tt < n_t;
            int bias = 0, s0 = 0;
            if (have) { bias = ((1 << sh) - 1) - pre_fx[tt * Kc + k]; s0 = pre_s0[tt * Kc + k]; }   // rays below point p: clamp((p + bias) >> sh, 0, NR)
            uint32_t acc[VC_ROWS / 2];
#pragma unroll
            for (int i = 0; i < VC_ROWS / 2; ++i) acc[i] = 0;
            int in_gaps = 0;          // rays found inside my gaps (rays below the first gap looked at, or past the last real
            bool more = have;         //   one, are outside the window start's guarantee / the cached domain: they count as band rays)
            for (int gbase = s0 >> 1;; gbase += 4 * CVS_GPL) {
                const int g0 = min(gbase, n_pad - 4 * CVS_GPL) + q;   // (only a finished tick can be clamped)
                int hi_last = 0;
#pragma unroll
                for (int i = 0; i < CVS_GPL; ++i) {
                    const int g = g0 + 4 * i;
                    const int2 p = P2c[g];
                    const uint4 m0 = M0c[min(g, n_gaps - 1)], m1 = M1c[min(g, n_gaps - 1)];
                    const int lo = max(0, min(NR, (p.x + bias) >> sh)), hi = max(0, min(NR, (p.y + bias) >> sh));
                    const int cnt = more ? hi - lo : 0;
                    in_gaps += cnt;
                    const uint32_t sel = cnt > 0 ? 0xffffffffu : 0u;   // the gap holds a ray: every ray inside marks the same tiles
                    or_and(acc[0], m0.x, sel); or_and(acc[1], m0.y, sel); or_and(acc[2], m0.z, sel); or_and(acc[3], m0.w, sel);
                    or_and(acc[4], m1.x, sel); or_and(acc[5], m1.y, sel); or_and(acc[6], m1.z, sel); or_and(acc[7], m1.w, sel);
                    hi_last = hi;
                }
                const int carry = __shfl_sync(FULL, hi_last, lane | 3);   // rays below the end of the tick's last gap of this pass
                more = more && carry < NR;
                if (!__any_sync(FULL, more)) break;
            }
            in_gaps += __shfl_xor_sync(FULL, in_gaps, 1);
            in_gaps += __shfl_xor_sync(FULL, in_gaps, 2);
            const bool band = have && in_gaps != NR;
            // the four quarters of a tick -> one mask per tick, in shared memory
#pragma unroll
            for (int i = 0; i < VC_ROWS / 2; ++i) {
                acc[i] |= __shfl_xor_sync(FULL, acc[i], 1);
                acc[i] |= __shfl_xor_sync(FULL, acc[i], 2);
            }
            const unsigned bands = __ballot_sync(FULL, band);
            if (q == 0) {
                reinterpret_cast<uint4 *>(mask_s)[tt * 2] = make_uint4(acc[0], acc[1], acc[2], acc[3]);
                reinterpret_cast<uint4 *>(mask_s)[tt * 2 + 1] = make_uint4(acc[4], acc[5], acc[6], acc[7]);
            }
            __syncwarp();
            // lane = grid row: window row wr of every tick's mask (two 16-bit rows per word) into the tick's row words --
            // registers for grids of one word per row and lane (up to 32 x 32), shared memory otherwise
            if (RPL == 1 && W == 1) {
                const int wr = lane - row0;
                if (wr >= 0 && wr <= nrow) {
                    const uint16_t *mrow = reinterpret_cast<const uint16_t *>(mask_s) + wr;
#pragma unroll
                    for (int t2 = 0; t2 < FAST_TB; ++t2) {
                        const unsigned bits = mrow[t2 * VC_ROWS];   // (ticks beyond n_t hold stale masks: never written out)
                        vr[t2] |= col0 >= 0 ? (bits << col0) : (bits >> (-col0));
                    }
                }
            } else {
                for (int t2 = 0; t2 < n_t; ++t2) {
                    uint32_t *rows = vis_s + t2 * RW;
                    const uint16_t *mrow = reinterpret_cast<const uint16_t *>(mask_s) + t2 * VC_ROWS;
#pragma unroll
                    for (int a = 0; a < RPL; ++a) {
                        const int wr = lane + 32 * a - row0;
                        if (wr >= 0 && wr <= nrow) {
                            const unsigned bits = mrow[wr];
                            if (bits) {
                                uint32_t v[W];
#pragma unroll
                                for (int w = 0; w < W; ++w) v[w] = 0;
                                fast_or_row<W>(v, bits, col0);
#pragma unroll
                                for (int w = 0; w < W; ++w) rows[(lane + 32 * a) * W + w] |= v[w];
                            }
                        }
                    }
                }
            }
            if (bands) {   // warp-uniform; rare
                if (RPL == 1 && W == 1 && !any_exact)
                    for (int i = lane; i < FAST_TB * RW; i += 32) vis_s[i] = 0;
                any_exact = true;
                __syncwarp();
                for (int t2 = 0; t2 < n_t; ++t2) {
                    if (!((bands >> (4 * t2)) & 0xfu)) continue;
                    const int b2 = ((1 << sh) - 1) - pre_fx[t2 * Kc + k];
                    double h = h0_s[k];   // the tick's heading again (as in the pre-phase)
                    for (int a = 0; a < t2; ++a) h = py_mod360(__dadd_rn(h, Cm.speed));
                    cam_exact_scan_staged(vc_geo(D), D.wall + (size_t)env * RW, vis_s + t2 * RW, &Cm, P2c, h, (int)pre_s0[t2 * Kc + k], b2, lane);
                }
            }
            __syncwarp();
        }
        k0 = k1;
        if (k0 < n_cams) {
            __syncthreads();   // the group's tables are no longer read
            k1 = stage_group(k0);
        }
    }
    __syncwarp();
    {
        uint32_t *dst = out + ((size_t)t_begin * D.N + env) * RW;
        const size_t stride = (size_t)D.N * RW;
        if (RPL == 1 && W == 1) {
            if (lane < RW) {
                dst += lane;
                const uint32_t *xs = vis_s + lane;
#pragma unroll
                for (int tt = 0; tt < FAST_TB; ++tt, dst += stride, xs += RW)   // (pointer steps: no 64-bit multiply per store)
                    if (tt < n_t) *dst = any_exact ? (vr[tt] | *xs) : vr[tt];
            }
        } else {
            for (int tt = 0; tt < n_t; ++tt, dst += stride)
                for (int i = lane; i < RW; i += 32) dst[i] = vis_s[tt * RW + i];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// k_seq: the sequential part of a launch.  Everything here is scalar per env (position, rewards, guard indices), so
// a warp per env would execute it 32 times over; a THREAD per env (round 1) put four guards' worth of unrolled,
// divergent code on the per-tick dependency chain (~500 warp-instructions per tick).  Here a QUAD of lanes owns an
// env: the four lanes carry the Solver's state redundantly and lane j of the quad owns guard j, so the guards of an
// env advance and test the Solver's tile side by side, and one ballot joins the verdicts.  8 envs per warp.
// Per tick: move, guards advance, detection from ONE word of cam_vis[t] OR the guards' cached masks at the Solver's
// row, vault / timeout, rewards, auto-reset.  What k_finish needs to complete the maps -- the guards' (waypoint,
// heading slot) of every tick -- is recorded.
//
// Few warps run this kernel (N / 8), one per scheduler: its speed is the length of the per-tick dependency chain,
// not throughput.  So nothing on that chain waits for memory: the three wall rows around the Solver, the guard's
// state, its current and next patrol word live in registers, and what tick t + 1 will read from global memory --
// the cam_vis word and the guard's mask row at the Solver's next row -- is requested while tick t is being decided
// (until an episode ends, the Solver's path and the patrols do not depend on what is seen, so the next state is
// known); actions are requested four ticks ahead.  The camera rows themselves -- written to global memory by
// k_cam_vis just before -- are STAGED in shared memory a stretch of ticks at a time (one pipelined burst of loads per
// ~32 ticks), so the word the detection needs is a shared-memory load instead of an L2 round trip on every tick's
// chain.  Every lane runs exactly T iterations (an auto-reset is part of the tick that ended the episode), so the
// quads of a warp stay in step.
// ---------------------------------------------------------------------------------------------
#define SEQ_THREADS 32
#define SEQ_EPW 8   // envs per warp
#define SEQ_LIVE (1u << 22)      // per-tick record bits (k_seq pass A -> pass B)
#define SEQ_DET (1u << 23)
#define SEQ_VAULT (1u << 24)
#define SEQ_TOUT (1u << 25)
#define SEQ_REBUILT (1u << 26)
#define SEQ_STAGE_BYTES 8192    // camera rows of the next ticks staged in shared memory
__host__ __device__ inline int seq_stage_ticks(int RW) { return max(1, min(64, SEQ_STAGE_BYTES / (SEQ_EPW * RW * 4))); }
__host__ __device__ inline size_t seq_warp_bytes(int RW, int L) {
    return ((((size_t)RW * SEQ_EPW + (size_t)L * SEQ_THREADS) * 4 + 15) & ~(size_t)15) + (size_t)seq_stage_ticks(RW) * SEQ_EPW * RW * 4 +
           (size_t)seq_stage_ticks(RW) * SEQ_EPW * 4 + SEQ_EPW * 4 + (((size_t)seq_stage_ticks(RW) * SEQ_EPW + 15) & ~(size_t)15);
}

template <int W>
__global__ void __launch_bounds__(SEQ_THREADS)
k_seq(Dev D, const int8_t *__restrict__ actions, int T, int autoreset, float *__restrict__ reward,
      double *__restrict__ reward64, uint8_t *__restrict__ done, uint8_t *__restrict__ status_out,
      const uint32_t *__restrict__ cam_vis, uint16_t *__restrict__ grec, uint8_t *__restrict__ fin,
      int32_t *__restrict__ last_t, int do_reset, const uint8_t *__restrict__ mask, int store_heading) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x, q = lane >> 2, j = lane & 3;
    const int env_raw = blockIdx.x * SEQ_EPW + q;
    const bool exists = env_raw < D.N && D.env_cached[min(env_raw, D.N - 1)];
    const int env = min(env_raw, D.N - 1);
    const bool valid = exists && !(do_reset && mask && !mask[env]);   // (an idle quad still takes part in the ballots)
    if (do_reset && exists && !valid && j == 0) last_t[env] = -1;
    const int R = D.R, C = D.C, N = D.N, RW = D.RW, L = D.L, Kg = D.Kg;
    // shared memory: wall rows [word][quad]; patrol words [waypoint][lane]:  row | col << 8 | slot << 16
    // (slot: heading slot taken when LEAVING the waypoint, 255 = unchanged)
    uint32_t *wall_s = reinterpret_cast<uint32_t *>(smem) + q;
    uint32_t *pw_s = reinterpret_cast<uint32_t *>(smem) + RW * SEQ_EPW + lane;
    uint32_t *cam_s = reinterpret_cast<uint32_t *>(smem + ((((size_t)RW * SEQ_EPW + (size_t)L * SEQ_THREADS) * 4 + 15) & ~(size_t)15));
    const int TS = seq_stage_ticks(RW);
    unsigned *rec_s = cam_s + (size_t)TS * SEQ_EPW * RW;          // [tick of the stage][quad]
    int *init_s = reinterpret_cast<int *>(rec_s + TS * SEQ_EPW);  // [quad] initial distance (-1: quad idle)
    int8_t *act_s = reinterpret_cast<int8_t *>(init_s + SEQ_EPW);  // [tick of the stage][quad] actions
    const int env0 = blockIdx.x * SEQ_EPW, n_here = min(SEQ_EPW, N - env0);   // this warp's envs are contiguous

    // ---- load ----
    const int4 es = valid ? *reinterpret_cast<const int4 *>(D.env_s + (size_t)env * 4) : make_int4(0, 0, 0, 0);
    const int n_cams = es.x, n_guards = es.y;
    const int4 d0 = *reinterpret_cast<const int4 *>(D.env_d + (size_t)env * 8);
    const int4 d1 = *reinterpret_cast<const int4 *>(D.env_d + (size_t)env * 8 + 4);
    EnvRegs E;
    E.r = d0.x & 0xffff; E.c = d0.x >> 16; E.tick = d0.y; E.prev = d0.z; E.init = d0.w;
    E.flags = d1.x & 0xff; E.n_vault = d1.y; E.n_detect = d1.z; E.n_timeout = d1.w;
    if (j == 0) init_s[q] = valid ? E.init : -1;
    for (int i = j; i < RW; i += 4) wall_s[i * SEQ_EPW] = D.wall[(size_t)env * RW + i];
    // my guard: waypoint, heading slot, path length, stride, range, next waypoint; patrol word at gk, gkn, waypoint 0
    const bool has_g = valid && j < n_guards;
    int gk = 0, ghs = 0, glen = 1, gstp = 0, grng = 0, gkn = 0;
    unsigned gw = 0, gwn = 0, gw0 = 0;
    const uint16_t *gmask = D.vg_mask;
    const size_t go = (size_t)env * Kg + j;
    if (has_g) {
        const int4 gi = *reinterpret_cast<const int4 *>(D.guard_i + go * 4);   // len, speed, range, num_rays
        glen = gi.x; gstp = gi.x >= 2 ? py_imod(gi.y, gi.x) : 0; grng = gi.z;
        gk = D.guard_idx[go];
        for (int k = 0; k < gi.x; ++k)
            pw_s[k * SEQ_THREADS] = (unsigned)D.guard_path[(go * L + k) * 2] | ((unsigned)D.guard_path[(go * L + k) * 2 + 1] << 8) |
                                    ((unsigned)D.vg_hslot[go * L + k] << 16);
        // heading -> slot.  A heading that is none of the path's can only have been written by hand into the
        // state view; it is reported (ERR_STATE) and treated as the default heading.
        const double h = D.guard_heading[go];
        int hs = -1;
        const int nh = D.vg_nh[go];
        for (int s = 0; s < nh; ++s)
            if (__double_as_longlong(D.vg_hval[go * (L + 1) + s]) == __double_as_longlong(h)) { hs = s; break; }
        if (hs < 0) { atomicOr(D.err, ERR_STATE); hs = 0; }
        if (!((D.vg_reach[go * L + min(max(gk, 0), L - 1)] >> min(hs, 31)) & 1u)) atomicOr(D.err, ERR_STATE);   // a cone that was never built
        ghs = hs;
        gmask = D.vg_mask + go * L * (size_t)(L + 1) * VC_ROWS;
        gkn = gk + gstp; if (gkn >= glen) gkn -= glen;
    }
    __syncwarp();
    if (has_g) { gw = pw_s[gk * SEQ_THREADS]; gwn = pw_s[gkn * SEQ_THREADS]; gw0 = pw_s[0]; }
    // Pass A below is written branch-free (selects and predicated loads / stores): the quads of a warp are in
    // different situations every tick, and every divergent branch would put both of its sides -- plus the
    // reconvergence bookkeeping -- on the one dependency chain this kernel consists of.
    const unsigned gw_reset_next = has_g ? pw_s[gstp * SEQ_THREADS] : 0u;   // patrol word after the first move from waypoint 0
    const int start_dist = abs(D.start_r - D.vault_r) + abs(D.start_c - D.vault_c);
    // is (nr, nc) inside the grid and not a WALL tile (:242-245)?  One shared-memory word.
    auto free_tile = [&](int nr, int nc) -> bool {
        const bool inb = (unsigned)nr < (unsigned)R && (unsigned)nc < (unsigned)C;
        const int rr = inb ? nr : 0, cc = inb ? nc : 0;
        const uint32_t word = wall_s[(rr * W + (W == 2 ? (cc >> 5) : 0)) * SEQ_EPW];
        return inb && !((word >> (cc & 31)) & 1u);
    };

    int status = HEIST_RUNNING;
    unsigned rec_last = 0;   // record of the launch's last tick (-> status)
    bool stepped = false;
    int outcomes = 0;        // pass B: episodes of env (lane & 7) ended by vault | detection << 10 | timeout << 20
    int n_adv = 0;           // camera updates executed by this launch
    int last = -1;           // last tick of this launch whose visibility map was rebuilt
    if (do_reset) {
        // HeistEnvironment.reset (environment.py:183-214): headings persist, guards back to waypoint 0.  The map after
        // a reset is completed by k_finish from the recorded guard state; nothing is detected on a reset.
        if (valid) {
            E.r = D.start_r; E.c = D.start_c; E.tick = 0; E.flags = 0; E.prev = start_dist; E.init = start_dist;
            gk = 0; gw = gw0; gkn = gstp; gwn = gw_reset_next;
            if (j == 0) fin[env] = 1;
            if (has_g) grec[(size_t)env * Kg + j] = (uint16_t)(gk | (ghs << 8));
            last = 0;
        }
        T = 0;
    }
    const bool vec_ok = ((((size_t)N * RW * 4) | ((size_t)env0 * RW * 4) | (uintptr_t)cam_vis) & 15) == 0 && ((n_here * RW) & 3) == 0;
    const bool act_vec = (((size_t)N | (size_t)env0 | (uintptr_t)actions) & 3) == 0 && n_here == SEQ_EPW;
    const bool moving = has_g && glen >= 2;
    for (int ts0 = 0; ts0 < T; ts0 += TS) {
    const int n_st = min(TS, T - ts0);
    stepped = true;
    {   // stage the camera rows and the actions of ticks [ts0, ts0 + n_st) of this warp's (contiguous) envs
        const int words = n_here * RW;
        __syncwarp();
        if (vec_ok) {
            // 16-byte items, tick by tick: a lane keeps its column(s) of the tick's q4 items and steps both pointers by
            // one tick -- two adds per copy, no division, no multiply
            const int q4 = words >> 2, dstep = (SEQ_EPW * RW) >> 2;   // uint4 per tick: real / slot in shared memory
            const size_t sstep = ((size_t)N * RW) >> 2;               // ... and per tick of cam_vis
            const uint4 *src = reinterpret_cast<const uint4 *>(cam_vis + ((size_t)ts0 * N + env0) * RW) + lane;
            uint4 *dst = reinterpret_cast<uint4 *>(cam_s) + lane;
            if (q4 <= 64) {   // (one-word grids up to 32 rows: at most two items per lane and tick)
                const bool one = lane < q4, two = lane + 32 < q4;
                for (int tt = 0; tt < n_st; ++tt, src += sstep, dst += dstep) {
                    if (one) cp_async16(dst, src);
                    if (two) cp_async16(dst + 32, src + 32);
                }
            } else {
                for (int tt = 0; tt < n_st; ++tt, src += sstep, dst += dstep)
                    for (int i = lane; i < q4; i += 32) cp_async16(dst + i - lane, src + i - lane);
            }
        } else {
            for (int tt = 0; tt < n_st; ++tt) {
                const uint32_t *src = cam_vis + ((size_t)(ts0 + tt) * N + env0) * RW;
                uint32_t *dst = cam_s + (size_t)tt * SEQ_EPW * RW;
                for (int i = lane; i < words; i += 32) cp_async4(dst + i, src + i);
            }
        }
        if (act_vec) {   // 8 action bytes per tick = two words
            for (int i = lane; i < n_st * 2; i += 32)
                cp_async4(reinterpret_cast<uint32_t *>(act_s) + i, actions + (size_t)(ts0 + (i >> 1)) * N + env0 + 4 * (i & 1));
        } else {
            for (int i = lane; i < n_st * SEQ_EPW; i += 32)
                act_s[i] = (i & (SEQ_EPW - 1)) < n_here ? actions[(size_t)(ts0 + (i >> 3)) * N + env0 + (i & (SEQ_EPW - 1))] : (int8_t)0;
        }
        cp_async_wait_all();
        __syncwarp();
    }
    // ---- pass A, the chain: per tick only what the NEXT tick depends on; the rest goes into a one-word record ----
    uint16_t *grec_p = grec + ((size_t)ts0 * N + env) * Kg + j;   // (stepped per tick: no 64-bit multiply on the chain)
    for (int t = ts0; t < ts0 + n_st; ++t, grec_p += (size_t)N * Kg) {
        const uint32_t *cam_row = cam_s + ((size_t)(t - ts0) * SEQ_EPW + q) * RW;
        const int a_cur = act_s[(t - ts0) * SEQ_EPW + q];
        const bool live = valid && !(E.flags & F_DONE);   // a done env is not mutated (:232-233)
        // move (:239-246): blocked by the grid edge or a WALL tile
        // (actions 0..4 = stay, up, down, left, right: row / column step + 1 as 2-bit fields of a constant)
        const unsigned a2 = 2u * (unsigned)min(max(a_cur, 0), 7);
        const int nr = E.r + (int)((0x5561u >> a2) & 3u) - 1, nc = E.c + (int)((0x5615u >> a2) & 3u) - 1;
        const bool go_there = live && free_tile(nr, nc);
        E.r = go_there ? nr : E.r;
        E.c = go_there ? nc : E.c;
        n_adv += live;   // cameras rotate (:251-252): their cones for this tick are cam_vis[t]
        {   // Guard.update (security.py:145-159)
            const bool adv = live && moving;
            const int hsl = (gw >> 16) & 255;
            ghs = (adv && hsl != 255) ? hsl : ghs;   // 255: the move is (0, 0), heading unchanged
            gk = adv ? gkn : gk;
            gw = adv ? gwn : gw;
            int k2 = gk + gstp;
            k2 = k2 >= glen ? k2 - glen : k2;
            gkn = adv ? k2 : gkn;
            gwn = pw_s[gkn * SEQ_THREADS];
        }
        // visibility at the Solver's tile: camera cones (all four lanes agree) OR my guard's cone / own tile
        const uint32_t cw = cam_row[E.r * W + (W == 2 ? (E.c >> 5) : 0)];
        bool mine = (cw >> (E.c & 31)) & 1u;
        {
            const int prow = gw & 255, pcol = (gw >> 8) & 255;
            const int wr = E.r - (prow - grng), wc = E.c - (pcol - grng);
            const bool in_win = has_g && (unsigned)wr <= (unsigned)(2 * grng) && (unsigned)wc <= (unsigned)(2 * grng);
            unsigned bits = 0;
            if (in_win) bits = __ldg(gmask + (gk * (L + 1) + ghs) * VC_ROWS + wr);
            mine = mine || ((bits >> (wc & 15)) & 1u);
        }
        mine = mine && live;
        const bool detected = (__ballot_sync(FULL, mine) >> (lane & ~3)) & 0xfu;   // any lane of my quad
        // record of the tick for pass B: row | col << 7 | distance before the move << 14 | live, detected, vault,
        // timeout, map rebuilt << 22..26
        const bool at_vault = live && E.r == D.vault_r && E.c == D.vault_c;                     // :284-288
        E.tick += live;
        const bool tout = live && E.tick >= D.max_steps;                                         // :291-297
        unsigned rec = live ? ((unsigned)E.r | ((unsigned)E.c << 7) | ((unsigned)E.prev << 14) | SEQ_LIVE | SEQ_REBUILT) : 0u;
        rec |= (detected ? SEQ_DET : 0u) | (at_vault ? SEQ_VAULT : 0u) | (tout ? SEQ_TOUT : 0u);
        E.prev = live ? abs(E.r - D.vault_r) + abs(E.c - D.vault_c) : E.prev;
        E.flags |= (detected ? (F_DETECTED | F_DONE) : 0) | (at_vault ? (F_VAULT | F_DONE) : 0) | (tout ? F_DONE : 0);
        rec_last = rec;   // (the launch's final status and the outcome counters are derived off the chain: below / pass B)
        // the trainer's `if done: reset()` (environment.py:183-214): headings persist, guards back to waypoint 0
        const bool rs = valid && autoreset && (E.flags & F_DONE);
        E.r = rs ? D.start_r : E.r; E.c = rs ? D.start_c : E.c;
        E.tick = rs ? 0 : E.tick; E.prev = rs ? start_dist : E.prev; E.init = rs ? start_dist : E.init;
        E.flags = rs ? 0 : E.flags;
        gk = rs ? 0 : gk; gw = rs ? gw0 : gw; gkn = rs ? gstp : gkn; gwn = rs ? gw_reset_next : gwn;
        rec |= rs ? SEQ_REBUILT : 0u;
        // the visibility map of tick t is final: tell k_finish how to complete it
        const bool rebuilt = rec & SEQ_REBUILT;
        last = rebuilt ? t : last;
        if (rebuilt && has_g) *grec_p = (uint16_t)(gk | (ghs << 8));
        if (valid && j == 0) rec_s[(t - ts0) * SEQ_EPW + q] = rec;
    }
    __syncwarp();
    // ---- pass B, off the chain: rewards and outputs of the stage's ticks, one lane per (tick, env) ----
    for (int i = lane; i < n_st * SEQ_EPW; i += 32) {
        const int qq = i & (SEQ_EPW - 1);
        const int init = init_s[qq];
        if (init < 0) continue;   // no env / not taking part in this launch
        const unsigned rec = rec_s[i];
        const size_t o = (size_t)(ts0 + (i >> 3)) * N + env0 + qq;
        double rw = 0.0;
        int st = HEIST_ALREADY_DONE, dn = 1;
        if (rec & SEQ_LIVE) {
            // shaping (:261-269), detection (:273-281), vault (:284-288), timeout (:291-297): same operations, same order
            const int r = rec & 127, c = (rec >> 7) & 127, prev = (rec >> 14) & 255;
            const int curr = abs(r - D.vault_r) + abs(c - D.vault_c);
            rw = D.reward_step;
            rw = __dadd_rn(rw, __dmul_rn((double)(prev - curr), 0.1));
            if (curr <= 3 && init > 3) rw = __dadd_rn(rw, __dmul_rn(0.05, (double)(3 - curr)));
            st = HEIST_RUNNING;
            if (rec & SEQ_DET) { rw = __dadd_rn(rw, D.reward_detection); st = HEIST_DETECTED; }
            if (rec & SEQ_VAULT) { rw = __dadd_rn(rw, D.reward_vault); st = HEIST_VAULT_REACHED; }
            if (rec & SEQ_TOUT) {
                st = HEIST_TIMEOUT;
                double cf = __dsub_rn(1.0, __ddiv_rn((double)curr, (double)max(init, 1)));
                if (!(cf > 0.0)) cf = 0.0;
                rw = __dadd_rn(rw, __dmul_rn(cf, 2.0));
            }
            dn = (rec & (SEQ_DET | SEQ_VAULT | SEQ_TOUT)) ? 1 : 0;
            outcomes += (st == HEIST_VAULT_REACHED ? 1 : 0) | (st == HEIST_DETECTED ? 1 << 10 : 0) | (st == HEIST_TIMEOUT ? 1 << 20 : 0);   // training.py:535-540
        }
        if (reward) reward[o] = (float)rw;
        if (reward64) reward64[o] = rw;
        if (done) done[o] = (uint8_t)dn;
        if (status_out) status_out[o] = (uint8_t)st;
        fin[o] = (rec & SEQ_REBUILT) ? 1 : 0;
    }
    }

    // Outcome counters: pass B's lane l saw the ticks of env l & 7 -> join the four lanes of an env, hand the sum to
    // the env's quad (at most 256 ticks per launch: 10-bit fields).  Status of the launch's last tick from its record.
    outcomes += __shfl_xor_sync(FULL, outcomes, 8);
    outcomes += __shfl_xor_sync(FULL, outcomes, 16);
    outcomes = __shfl_sync(FULL, outcomes, q);
    E.n_vault += outcomes & 1023; E.n_detect += (outcomes >> 10) & 1023; E.n_timeout += (outcomes >> 20) & 1023;
    if (stepped)
        status = !(rec_last & SEQ_LIVE) ? HEIST_ALREADY_DONE
                 : ((rec_last & SEQ_TOUT) ? HEIST_TIMEOUT : ((rec_last & SEQ_VAULT) ? HEIST_VAULT_REACHED : ((rec_last & SEQ_DET) ? HEIST_DETECTED : HEIST_RUNNING)));
    // ---- store ----
    if (!valid) return;
    if (j == 0) {
        last_t[env] = last;
        *reinterpret_cast<int4 *>(D.env_d + (size_t)env * 8) = make_int4(E.r | (E.c << 16), E.tick, E.prev, E.init);
        *reinterpret_cast<int4 *>(D.env_d + (size_t)env * 8 + 4) =
            make_int4(E.flags | (status << 8), E.n_vault, E.n_detect, E.n_timeout);
    }
    for (int k = j; k < (store_heading ? n_cams : 0); k += 4) {   // headings after the camera updates this launch executed
        const size_t co = (size_t)env * D.Kc + k;                // (otherwise stored by k_heads)
        double h = D.cam_heading[co];
        const double speed = D.cam_f[co * 2 + 1];
        for (int a = 0; a < n_adv; ++a) h = py_mod360(__dadd_rn(h, speed));
        D.cam_heading[co] = h;
    }
    if (has_g) {
        D.guard_heading[go] = D.vg_hval[go * (L + 1) + ghs];
        D.guard_idx[go] = gk;
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
    if (n_guards == 0) {
        // no guards: the camera rows ARE the finished maps; only the env's current map has to be refreshed from
        // the last rebuilt tick (by the block that holds it)
        if (lt >= t0 && lt < t1)
            for (int i = lane; i < RW; i += 32) D.vis[(size_t)env * RW + i] = buf[((size_t)lt * D.N + env) * RW + i];
        return;
    }
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

// k_finish_or: k_finish for launches that auto-reset and keep the trajectory (the rollout path).  There every tick of
// every env is rebuilt, and the camera rows are already in place -- so instead of reading and re-writing whole maps
// (lane = grid row), a lane takes one (tick, guard) pair of the warp's env and ORs the guard's cone rows into the
// map with fire-and-forget atomics (RED.OR at L2; only rows with a lit tile are touched, and guards sharing a row
// need no ordering).  ~4x fewer instructions and no read of the maps.  The env's current map D.vis is refreshed
// from the finished last tick by the block that holds it (L2 loads: the atomics do not pass through L1).
template <int W>
__global__ void __launch_bounds__(256)
k_finish_or(Dev D, int T, uint32_t *buf, const uint16_t *__restrict__ grec, const uint8_t *__restrict__ fin,
            const int32_t *__restrict__ last_t) {
    const int lane = threadIdx.x & 31;   // grid: x = env / 8, y = tick block
    const int env = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (env >= D.N || !D.env_cached[env]) return;
    const int lt = last_t[env];
    const int t0 = blockIdx.y * FIN_TB, t1 = min(T, t0 + FIN_TB);
    const int n_guards = D.env_s[(size_t)env * 4 + 1], L = D.L, Kg = D.Kg, RW = D.RW;
    static_assert(FIN_TB * VC_MAX_GUARDS == 32, "lane = (tick, guard)");
    const int t = t0 + (lane >> 2), g = lane & 3;
    if (t < t1 && g < n_guards) {
        const size_t o = (size_t)t * D.N + env, go = (size_t)env * Kg + g;
        if (fin[o]) {
            const unsigned rec = grec[o * Kg + g];
            const int k = rec & 255, hs = rec >> 8;
            const int rng = D.guard_i[go * 4 + 2];
            const int row0 = (int)D.guard_path[(go * L + k) * 2] - rng, col0 = (int)D.guard_path[(go * L + k) * 2 + 1] - rng;
            const uint4 *m = reinterpret_cast<const uint4 *>(D.vg_mask + ((go * L + k) * (size_t)(L + 1) + hs) * VC_ROWS);
            const uint4 m0 = __ldg(m), m1 = __ldg(m + 1);
            const uint32_t words[VC_ROWS / 2] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
            uint32_t *rows = buf + o * RW;
#pragma unroll
            for (int wr = 0; wr < VC_ROWS; ++wr) {
                const unsigned bits = (wr & 1) ? (words[wr >> 1] >> 16) : (words[wr >> 1] & 0xffffu);
                const int r = row0 + wr;
                if (bits && (unsigned)r < (unsigned)D.R) {
                    uint32_t v[W];
#pragma unroll
                    for (int w = 0; w < W; ++w) v[w] = 0;
                    fast_or_row<W>(v, bits, col0);
#pragma unroll
                    for (int w = 0; w < W; ++w) if (v[w]) atomicOr(&rows[r * W + w], v[w]);
                }
            }
        }
    }
    if (lt >= t0 && lt < t1) {
        __syncwarp();
        for (int i = lane; i < RW; i += 32) D.vis[(size_t)env * RW + i] = __ldcg(buf + ((size_t)lt * D.N + env) * RW + i);
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
