// heist_common.cuh -- device-side state description and exact-arithmetic helpers.
//
// Everything that decides a bit of the visibility map / positions / done flags is written with
// explicit IEEE operations in the reference's order (CPython double arithmetic):
// separate multiply and add (no FMA contraction; the TU is also compiled with -fmad=false),
// round-half-even tile rounding, Python float modulo.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define HEIST_WARPS_PER_CTA 4
#define NICE_K 48                 // multiples of 30 degrees covered: |angle| <= 1440
#define NICE_N (2 * NICE_K + 1)
#define NICE_W 1e-11              // half width (degrees) of the window around each multiple taken from the host libm

// cos / -sin of every double within NICE_W degrees of a multiple of 30 degrees, evaluated by the HOST libm at
// heist_create (CPython's math.cos/sin call the same libm; security.py:71-75).  The last ulp of cos/sin can decide a
// tile only where col + cos(a) * dist lands within a rounding error (~1e-14 tile on a 64-wide grid) of an exact
// .5 tie, which happens structurally only at cos/sin = +-1/2, +-1 (Niven), i.e. for angles within ~7e-13 degree of
// a multiple of 30 -- the exact multiples people type (fov 60, speed 15, heading 0) and the values a hand-typed
// non-dyadic speed accumulates to (0.1 * 300 = 30.000000000000156).  Those angles are taken from the platform the
// reference runs on instead of the device's sincos: window k (angle ~ (k - NICE_K) * 30) is the c_nice_cnt[k]
// consecutive doubles of one sign starting at bit pattern c_nice_lo[k] (smallest magnitude first), stored at
// c_nice_tab[c_nice_off[k] ...] as (cos, -sin).  Around 0 the window is not enumerable (denormals); there
// cos(x) = 1.0 and sin(x) = x exactly for |x| < 2^-27 in glibc (s_sin.c) -- pinned by a CPU test.
__constant__ long long c_nice_lo[NICE_N];
__constant__ int c_nice_cnt[NICE_N];
__constant__ int c_nice_off[NICE_N];
__constant__ const double2 *c_nice_tab;

struct Dev {
    int N, R, C, W, RW, RC;
    int max_steps, start_r, start_c, vault_r, vault_c, budget;
    int Kw, Kc, Kg, L;
    double reward_vault, reward_detection, reward_step;
    double deg2rad;  // Py_MATH_PI / 180.0 evaluated on the host (math.radians)
    // static per layout
    uint8_t *tile;        // [N][RC]
    uint32_t *wall;       // [N][RW]
    int32_t *env_s;       // [N][4]  n_cams, n_guards, valid, spent
    uint8_t *wall_ok;     // [N][Kw] wall i of the request was accepted (HeistEnvironment.walls)
    double *cam_f;        // [N][Kc][2] fov, speed
    int16_t *cam_i;       // [N][Kc][4] row, col, range, num_rays
    double *guard_fov;    // [N][Kg]
    int32_t *guard_i;     // [N][Kg][4] len, speed, range, num_rays
    uint8_t *guard_path;  // [N][Kg][L][2]
    double *guard_head;   // [N][Kg][L] heading after leaving waypoint i (NaN: unchanged)
    // dynamic
    int32_t *env_d;       // [N][8]
    double *cam_heading;  // [N][Kc]
    double *guard_heading;// [N][Kg]
    int32_t *guard_idx;   // [N][Kg]
    uint32_t *vis;        // [N][RW]
    float *pos_tab;       // [RC] float32(-0.3 * (manhattan(cell, vault) / (R + C)))
    int32_t *cost;        // [N] ray-march samples per tick (load-balance estimate)
    int32_t *slot2env;    // [ceil(N/4)*4] warp slot -> env (-1: empty), cost-balanced
    int *err;             // sticky device error flags
    // angular visibility cache (heist_cache.cuh); vc_p == nullptr: cache disabled
    int32_t *vc_p;        // [N][Kc][VC_POINTS] sorted boundary points of the tie bands: (angle - lo) / ray pitch, fixed point
    uint16_t *vc_mask;    // [N][Kc][VC_POINTS/2][VC_ROWS] window bitmap of gap g
    uint16_t *vc_idx;     // [N][Kc][VC_IDX] coarse index: boundary points below each 1-degree bucket
    int32_t *vc_meta;     // [N][Kc][2] n_points (-1: not cacheable), fixed-point shift of vc_p
    double *vc_lo;        // [N][Kc][2] lower end of the cached angle domain; fixed-point units per degree (fx_scale)
    uint16_t *vg_mask;    // [N][Kg][L][L+1][VC_ROWS] guard cone per (waypoint, heading slot)
    double *vg_hval;      // [N][Kg][L+1] distinct headings a guard can carry
    uint8_t *vg_hslot;    // [N][Kg][L] slot of guard_head[k] (255: unchanged)
    int32_t *vg_nh;       // [N][Kg] number of heading slots (-1: not cacheable)
    uint32_t *vg_reach;   // [N][Kg][L] bit min(slot, 31): the guard can stand on this waypoint carrying this heading slot --
                          //   only those (waypoint, slot) cones are built
    uint8_t *env_cached;  // [N] every asset of the env is served by the cache
    int *n_uncached;      // [1] envs left to the ray-march kernel
    int skip_cached;      // launch flag: the ray-march kernel leaves cached envs to k_fast
    int strict_tables;    // launch flag (HEIST_MODE_TABLES): an env the cache does not cover is an error
};

enum { ERR_CAPACITY = 1, ERR_WAYPOINT = 2, ERR_RAYS = 4, ERR_BOUNDS = 8, ERR_STATE = 16, ERR_UNCOVERED = 32 };
enum { F_DONE = 1, F_DETECTED = 2, F_VAULT = 4 };

// Python `x % 360.0` (floatobject.c float_rem): fmod, then shift negative remainders up.
__device__ __forceinline__ double py_mod360(double x) {
    double m;
    if (fabs(x) < 360.0) m = x;                       // fmod is the identity here
    else if (x >= 360.0 && x < 720.0) m = x - 360.0;  // exact (Sterbenz), equals fmod
    else m = fmod(x, 360.0);                          // CUDA fmod is exact
    if (m != 0.0) { if (m < 0.0) m = __dadd_rn(m, 360.0); }
    else m = 0.0;                                     // copysign(0.0, 360.0)
    return m;
}

// Python int %: result has the divisor's sign (len > 0).
__device__ __forceinline__ int py_imod(int a, int n) { int m = a % n; return m < 0 ? m + n : m; }

// Ray direction for angle_deg: dx = cos(radians(a)), dy = -sin(radians(a))  (security.py:71-75).
__device__ __forceinline__ void ray_dir(double angle_deg, double deg2rad, double &dx, double &dy) {
    const double k = rint(angle_deg * (1.0 / 30.0));
    if (fabs(k) <= (double)NICE_K) {
        const int i = (int)k + NICE_K;
        if (i == NICE_K) {   // around 0: cos = 1, sin(x) = x (see above)
            if (fabs(angle_deg) <= NICE_W) { dx = 1.0; dy = -__dmul_rn(angle_deg, deg2rad); return; }
        } else {
            const unsigned long long d = (unsigned long long)(__double_as_longlong(fabs(angle_deg)) - c_nice_lo[i]);
            if (d < (unsigned long long)c_nice_cnt[i]) {
                const double2 v = c_nice_tab[c_nice_off[i] + (int)d];
                dx = v.x;
                dy = v.y;
                return;
            }
        }
    }
    double s, c;
    sincos(__dmul_rn(angle_deg, deg2rad), &s, &c);
    dx = c;
    dy = -s;
}
