// heist_cache.cuh -- angular visibility cache: the per-layout tables behind the table-driven step kernel.
//
// Reference: Camera.get_vision_cone_tiles (security.py:53-101), Guard.get_visible_tiles (security.py:161-192),
// DynamicVisibilityMap.update (visibility.py:31-65).
//
// The walls of a layout never move (environment.py:118-122), so the tile sequence marked by ONE ray of a camera
// is a function of the ray's angle alone, and that function is piecewise constant: it can only change where some
// sample `col + cos(a) * dist` / `row - sin(a) * dist` crosses a rounding tie (m + 0.5).  For every camera the
// build kernel below
//   1. encloses every tie crossing of every sample in a BAND of angles: the pre-image of [tie - 2mu, tie + 2mu]
//      (mu = 4e-12 tile, far above the ~5e-14 the reference's fp64 evaluation can deviate from the real value,
//      far below anything a generic ray comes near), padded by 1e-9 degree, over the domain
//      [-fov/2, 360 + fov/2) a ray angle `heading - fov/2 + fov * i / num_rays` can take (heading in [0, 360));
//   2. sorts and merges the bands; the GAPS between them are angle intervals on which every sample of a ray
//      stays at least mu away from every tie, so the reference's rounding -- whatever the last bit of the
//      platform's cos/sin -- is constant there;
//   3. marches one representative ray per gap through the reference's own arithmetic (ray_exact semantics, walls
//      included) and stores the marked tiles as a (2*range+1)^2 window bitmap: 16 rows of 16 bits;
//   4. drops every band that comes from a single tie crossing and separates two gaps with equal bitmaps (a ray
//      inside it takes one of the two neighbouring tile sequences, so it marks those same tiles);
//   5. stores the boundary points as fixed-point ray pitches (the per-tick "rays below this point" becomes an
//      integer subtract and shift) and a 1-degree coarse index over them.
// At run time (heist_fast.cuh) a tick of a camera is: find the segments its fov window covers, OR the masks of
// the gaps that contain at least one ray, and run the few rays that fall inside a band (structurally: rays on
// exact multiples of 30/45/90 degrees) through the exact ray-march.  A guard's cone depends only on its waypoint
// and on which of the path's headings it carries (security.py:145-159), so guards get one mask per
// (waypoint, heading) pair.
#pragma once
#include "heist_common.cuh"

#define VC_ROWS 16            // mask rows per window (u16 each): 2 * range + 1 <= 15
#define VC_MAX_RANGE 7
#define VC_MAX_GUARDS 4       // guards per env the table-driven kernels keep in registers
#define VC_POINTS 1024        // boundary points kept per camera (2 per merged band)
#define VC_RAW 1024           // raw bands sorted per camera (power of two)
#define VC_IDX 560            // 1-degree buckets of the coarse index (domain <= 540 degrees + slack)
#define VC_MU2 8e-12          // 2 * mu: position-space half width of a computed band (tiles)
#define VC_PAD 1e-9           // angular padding on each side of a band (degrees)
#define VC_BUILD_THREADS 128

// Tie bands per vision range (index = range 1 .. VC_MAX_RANGE), WITHOUT padding, over ray angles [-91, 451] degrees,
// sorted by start: (start, end) pairs built once per device by the host at heist_create (vc_upload_band_tables in
// heist_b200.cu).  The bands of a camera are this list cut to its angle domain and padded: the positions of the
// rounding ties depend on the sample distances only, not on the layout -- so no per-layout acos and no sort.
__constant__ const double2 *c_vcb_tab[VC_MAX_RANGE + 1];
__constant__ int c_vcb_n[VC_MAX_RANGE + 1];

#define RINT_MAGIC_C 6755399441055744.0  // 2^52 + 2^51
__device__ __forceinline__ int vc_rint_even(double x) { return __double2loint(__dadd_rn(x, RINT_MAGIC_C)); }

// One ray of the reference's ray-march on the bit maps (security.py:69-99 cameras, :170-190 guards), from
// sample j0 on.  `mark(r, c)` receives every visible tile (the camera's own tile is filtered by the caller's
// functor when needed).  Walls come from the row bitmaps in global memory: this is the rare / build-time path.
struct VcGeo {   // what a ray needs to know about the grid
    int R, C, W;
    double deg2rad;
};
__device__ __forceinline__ VcGeo vc_geo(const Dev &D) { VcGeo g; g.R = D.R; g.C = D.C; g.W = D.W; g.deg2rad = D.deg2rad; return g; }

// The angle of ray ri of a cone, in the reference's operation order (security.py:69-71).
__device__ __forceinline__ double vc_ray_angle_deg(double fov, double heading, int num_rays, int ri) {
    const double half_fov = __ddiv_rn(fov, 2.0);
    return __dadd_rn(__dsub_rn(heading, half_fov), __ddiv_rn(__dmul_rn(fov, (double)ri), (double)num_rays));
}
// ... and its march from (row, col) along (dx, dy): nsamp samples `unit` apart (security.py:77-99, :176-190).
template <typename Mark>
__device__ __forceinline__ void vc_march(const VcGeo &D, const uint32_t *__restrict__ wall, int row, int col, double dx, double dy,
                                         int nsamp, double unit, Mark mark) {
    const double dcol = (double)col, drow = (double)row;
    double dist = unit;
    for (int j = 1; j <= nsamp; ++j, dist += unit) {
        const int c = vc_rint_even(__dadd_rn(dcol, __dmul_rn(dx, dist)));
        const int r = vc_rint_even(__dadd_rn(drow, __dmul_rn(dy, dist)));
        if (r < 0 || r >= D.R || c < 0 || c >= D.C) return;
        if ((wall[r * D.W + (c >> 5)] >> (c & 31)) & 1u) return;
        mark(r, c);
    }
}
template <typename Mark>
__device__ __forceinline__ void vc_ray(const VcGeo &D, const uint32_t *__restrict__ wall, int row, int col, double fov,
                                       double heading, int num_rays, int nsamp, double unit, int ri, Mark mark) {
    double dx, dy;
    ray_dir(vc_ray_angle_deg(fov, heading, num_rays, ri), D.deg2rad, dx, dy);
    vc_march(D, wall, row, col, dx, dy, nsamp, unit, mark);
}

// Same march for the representative angle of a gap.  Inside a gap every sample of a ray stays at least mu = 4e-12
// tile away from every rounding tie, whatever the last bits of cos / sin, so any evaluation that is accurate to
// ~1e-13 gives the tiles the reference's arithmetic gives: sincospi (exact argument reduction, no slow path).
template <typename Mark>
__device__ __forceinline__ void vc_ray_angle(const VcGeo &D, const uint32_t *__restrict__ wall, int row, int col,
                                             double angle_deg, int nsamp, double unit, Mark mark) {
    double s, c;
    sincospi(angle_deg * (1.0 / 180.0), &s, &c);
    const double dx = c, dy = -s;
    const double dcol = (double)col, drow = (double)row;
    double dist = unit;
    for (int j = 1; j <= nsamp; ++j, dist += unit) {
        const int cc = vc_rint_even(__dadd_rn(dcol, __dmul_rn(dx, dist)));
        const int r = vc_rint_even(__dadd_rn(drow, __dmul_rn(dy, dist)));
        if (r < 0 || r >= D.R || cc < 0 || cc >= D.C) return;
        if ((wall[r * D.W + (cc >> 5)] >> (cc & 31)) & 1u) return;
        mark(r, cc);
    }
}

__device__ __forceinline__ bool vc_cam_cacheable(double fov, double heading, double speed, int range, int num_rays) {
    return range >= 1 && range <= VC_MAX_RANGE && fov > 0.0 && fov <= 180.0 && num_rays >= 1 && num_rays <= 32767 &&
           fabs(heading) < 1e6 && fabs(speed) < 1e6;
}
__device__ __forceinline__ bool vc_guard_cacheable(double fov, int range, int num_rays, int len, int L) {
    return range >= 0 && range <= VC_MAX_RANGE && fabs(fov) <= 1e6 && num_rays >= 1 && num_rays <= 4095 && len >= 1 &&
           len <= L;
}

struct VcSmem {
    double key[VC_RAW];   // merged band starts (sorted)
    double end[VC_RAW];   // ... and ends
    unsigned char first_flag[VC_BUILD_THREADS];   // merge scan: does the thread's first position open a band
    uint4 gm[VC_POINTS / 2][2];       // gap masks before compaction
    unsigned char members[VC_RAW];    // 1: merged band i comes from a single raw band (one tie crossing); saturates at 255
    unsigned short newidx[VC_RAW];    // band i -> index after dropping redundant bands (0xffff: dropped); before that,
                                      //   the merge scan's first position of merged band i; afterwards, the index marks
    uint32_t wall[HEIST_MAX_DIM * 2];   // the env's wall rows
    double red_d[VC_BUILD_THREADS / 32];
    int red_i[VC_BUILD_THREADS / 32];
    int n_raw, n_bands, ok;
};

// (j, tie) pairs with |tie| <= 0.5 * j: the tie crossings a sample at distance 0.5 * j can have (j = 1 .. 14)
#define VC_MAX_ITEMS (2 * (1 + 1 + 2 + 2 + 3 + 3 + 4 + 4 + 5 + 5 + 6 + 6 + 7 + 7))
__device__ __forceinline__ int vc_ties_upto(int j) {   // pairs with sample index < j:  sum_{i<j} 2 * ((i + 1) / 2)
    const int h = (j - 1) >> 1;             // full (odd, even) couples below j
    return 2 * (h * (h + 1)) + (((j - 1) & 1) ? 2 * (h + 1) : 0);
}

// block-wide EXCLUSIVE scans over VC_BUILD_THREADS threads (4 warps); `total` receives the reduction over all threads
__device__ __forceinline__ double vc_block_exscan_max(double v, double *red, int tid) {
    double inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const double u = __shfl_up_sync(0xffffffffu, inc, o); if ((tid & 31) >= o) inc = fmax(inc, u); }
    double ex = __shfl_up_sync(0xffffffffu, inc, 1);
    if ((tid & 31) == 0) ex = -1e308;
    if ((tid & 31) == 31) red[tid >> 5] = inc;
    __syncthreads();
    for (int w = 0; w < (tid >> 5); ++w) ex = fmax(ex, red[w]);
    __syncthreads();
    return ex;
}
__device__ __forceinline__ int vc_block_exscan_sum(int v, int *red, int tid, int &total) {
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if ((tid & 31) >= o) inc += u; }
    int ex = inc - v;
    if ((tid & 31) == 31) red[tid >> 5] = inc;
    __syncthreads();
    total = 0;
    for (int w = 0; w < VC_BUILD_THREADS / 32; ++w) { if (w < (tid >> 5)) ex += red[w]; total += red[w]; }
    __syncthreads();
    return ex;
}

// One CTA per env: tables of all its cameras and guards.  Launched after every k_set_layout.
__global__ void __launch_bounds__(VC_BUILD_THREADS, 6) k_build_cache(Dev D) {
    __shared__ VcSmem S;
    const int env = blockIdx.x, tid = threadIdx.x;
    const int n_cams = D.env_s[(size_t)env * 4 + 0], n_guards = D.env_s[(size_t)env * 4 + 1];
    for (int i = tid; i < D.RW; i += VC_BUILD_THREADS) S.wall[i] = D.wall[(size_t)env * D.RW + i];
    __syncthreads();
    const uint32_t *wall = S.wall;
    const VcGeo geo = vc_geo(D);
    bool env_ok = n_guards <= VC_MAX_GUARDS;
    constexpr int PER = VC_RAW / VC_BUILD_THREADS;   // sorted positions per thread in the merge scan

    for (int k = 0; k < n_cams; ++k) {
        const size_t o = (size_t)env * D.Kc + k;
        const double fov = D.cam_f[o * 2], speed = D.cam_f[o * 2 + 1], heading = D.cam_heading[o];
        const int16_t *ci = D.cam_i + o * 4;
        const int row = ci[0], col = ci[1], range = ci[2], num_rays = ci[3];
        if (!vc_cam_cacheable(fov, heading, speed, range, num_rays)) { env_ok = false; continue; }  // CTA-uniform
        const int nsamp = 2 * range;
        const double dom_lo = -0.5 * fov - 1e-6, dom_hi = 360.0 + 0.5 * fov + 1e-6;
        // Boundary points are stored in units of the ray pitch above dom_lo as fixed point below 2^29, so that the
        // per-tick ray count below a point is an integer subtract and shift (k_cam_vis).  The quantisation (under 2
        // units for point and base together) is absorbed by widening every band by 4 units.
        const double inv_step = 1.0 / (fov / (double)num_rays);
        const int sh = max(0, min(24, 28 - ilogb((dom_hi - dom_lo) * inv_step)));
        const double fx_scale = inv_step * (double)(1 << sh);
        const double pad = VC_PAD + 4.0 / fx_scale;
        // ---- 1. bands: the per-range list (sorted by start), cut to the domain, padded; everything outside the domain
        // is one band on each side (rays there -- odd initial headings -- take the exact path) ----
        const double2 *T = c_vcb_tab[range];
        const int nT = c_vcb_n[range];
        int i0, n_raw;
        {   // candidates: starts in [dom_lo - 2, dom_hi + 2] (no band is wider than 1e-3 degree)
            int lo = 0, hi = nT;
            while (lo < hi) { const int m = (lo + hi) >> 1; if (T[m].x < dom_lo - 2.0) lo = m + 1; else hi = m; }
            i0 = lo;
            hi = nT;
            while (lo < hi) { const int m = (lo + hi) >> 1; if (T[m].x <= dom_hi + 2.0) lo = m + 1; else hi = m; }
            n_raw = lo - i0 + 2;   // + the two outside bands
        }
        bool ok = n_raw <= VC_RAW;
        if (!ok) n_raw = 2;
        // ---- 2b. merge overlaps, in parallel: with the bands sorted by start, position p opens a new merged band iff its
        // start exceeds the running maximum of the ends before it (exclusive max-scan); the band index is the running
        // count of such positions (sum-scan); a band's end is the running maximum at its last position ----
        int nb0;
        {
            double st[PER], en[PER];
            double mine = -1e308;
#pragma unroll
            for (int u = 0; u < PER; ++u) {
                const int p = tid * PER + u;
                st[u] = 1e308; en[u] = -1e308;   // (a position without a band never opens one and leaves the maximum alone)
                if (p == 0) { st[u] = -1e300; en[u] = dom_lo; }
                else if (p == n_raw - 1) { st[u] = dom_hi; en[u] = 1e300; }
                else if (p < n_raw) {
                    const double2 be = T[i0 + p - 1];
                    const double s = be.x - pad, e = be.y + pad;
                    if (!(e < dom_lo || s > dom_hi)) { st[u] = s; en[u] = e; }
                }
                mine = fmax(mine, en[u]);
            }
            __syncthreads();   // (the previous camera's key / end are no longer read)
            const double before = vc_block_exscan_max(mine, S.red_d, tid);
            unsigned flags = 0;
            int cnt = 0;
            double pm = before;
#pragma unroll
            for (int u = 0; u < PER; ++u) {
                const int p = tid * PER + u;
                if (p < n_raw && st[u] < 1e308 && (p == 0 || st[u] > pm)) { flags |= 1u << u; ++cnt; }
                pm = fmax(pm, en[u]);
            }
            unsigned char *first_flag = S.first_flag;
            unsigned short *first_pos = S.newidx;
            first_flag[tid] = (unsigned char)(flags & 1u);
            int total;
            int bid = vc_block_exscan_sum(cnt, S.red_i, tid, total) - 1;   // band of the position before this thread's first
            const int bid0 = bid;
#pragma unroll
            for (int u = 0; u < PER; ++u) {
                const int p = tid * PER + u;
                if (p < n_raw && (flags & (1u << u))) { ++bid; S.key[bid] = st[u]; first_pos[bid] = (unsigned short)p; }
            }
            __syncthreads();
            bid = bid0;
            pm = before;
#pragma unroll
            for (int u = 0; u < PER; ++u) {
                const int p = tid * PER + u;
                if (p >= n_raw) break;
                if (flags & (1u << u)) ++bid;
                pm = fmax(pm, en[u]);
                const bool next_opens = (p + 1 >= n_raw) || (u + 1 < PER ? (flags >> (u + 1)) & 1u : first_flag[tid + 1]);
                if (next_opens) { S.end[bid] = pm; S.members[bid] = (unsigned char)min(255, p - (int)first_pos[bid] + 1); }
            }
            __syncthreads();
            nb0 = total;   // gap g lies between band g and band g + 1
        }
        if (2 * (nb0 - 1) > VC_POINTS - 2) ok = false;   // (one pair of padding points must remain: the scans rely on it)
        if (!ok) { env_ok = false; if (tid == 0) D.vc_meta[o * 2] = -1; __syncthreads(); continue; }
        // ---- 3. one representative ray per gap -> window mask ----
        for (int g = tid; g < nb0 - 1; g += VC_BUILD_THREADS) {
            const double mid = 0.5 * (S.end[g] + S.key[g + 1]);
            unsigned *rows = reinterpret_cast<unsigned *>(&S.gm[g][0]);   // the thread's own gap: plain read-modify-write
            S.gm[g][0] = make_uint4(0, 0, 0, 0);
            S.gm[g][1] = make_uint4(0, 0, 0, 0);
            vc_ray_angle(geo, wall, row, col, mid, nsamp, 0.5, [&](int r, int c) {
                if (r == row && c == col) return;  // (r, c) != (self.row, self.col), security.py:93
                const int wr = r - row + range, wc = c - col + range;
                rows[wr >> 1] |= 1u << (wc + 16 * (wr & 1));
            });
        }
        __syncthreads();
        // ---- 3b. drop redundant bands.  A band that comes from ONE tie crossing separates two gaps whose tile
        // sequences differ in that one sample only; a ray inside it takes one of the two sequences.  When both
        // gaps mark the same tiles the ray does too, whichever way it rounds: band and gaps merge into one gap.
        // (Bands merged from several crossings could mix the neighbours' sequences and are kept.)  Every band
        // decides for itself (the test reads the ORIGINAL gap masks); the new indices are a sum-scan of the keepers.
        {
            constexpr int PB = (VC_RAW + VC_BUILD_THREADS - 1) / VC_BUILD_THREADS;
            unsigned keepm = 0;
            int cnt = 0;
#pragma unroll
            for (int u = 0; u < PB; ++u) {
                const int bnd = tid * PB + u;
                if (bnd >= nb0) break;
                bool drop = false;
                if (bnd >= 1 && bnd < nb0 - 1 && S.members[bnd] == 1) {
                    const uint4 a0 = S.gm[bnd - 1][0], a1 = S.gm[bnd - 1][1], c0 = S.gm[bnd][0], c1 = S.gm[bnd][1];
                    drop = a0.x == c0.x && a0.y == c0.y && a0.z == c0.z && a0.w == c0.w && a1.x == c1.x && a1.y == c1.y &&
                           a1.z == c1.z && a1.w == c1.w;
                }
                if (!drop) { keepm |= 1u << u; ++cnt; }
            }
            int total;
            int ni = vc_block_exscan_sum(cnt, S.red_i, tid, total);
#pragma unroll
            for (int u = 0; u < PB; ++u) {
                const int bnd = tid * PB + u;
                if (bnd >= nb0) break;
                S.newidx[bnd] = (keepm & (1u << u)) ? (unsigned short)ni++ : (unsigned short)0xffff;
            }
            if (tid == 0) S.n_bands = total;
            __syncthreads();
        }
        const int nb = S.n_bands;
        const int n_points = 2 * (nb - 1);   // p[2g] = end of band g, p[2g+1] = start of band g + 1
        {   // compact bands and gap masks in place (new index <= old index; whole chunk read before it is written)
            for (int c0 = 0; c0 < nb0; c0 += VC_BUILD_THREADS) {
                const int b = c0 + tid;
                double ks = 0, ke = 0; uint4 m0 = make_uint4(0, 0, 0, 0), m1 = m0; int ni = 0xffff;
                if (b < nb0) { ks = S.key[b]; ke = S.end[b]; ni = S.newidx[b]; if (b < nb0 - 1) { m0 = S.gm[b][0]; m1 = S.gm[b][1]; } }
                __syncthreads();
                // gap b (right of band b) keeps its mask under band b's new index; a dropped band's right gap
                // equals its left one, which is already stored
                if (b < nb0 && ni != 0xffff) { S.key[ni] = ks; S.end[ni] = ke; if (b < nb0 - 1) { S.gm[ni][0] = m0; S.gm[ni][1] = m1; } }
                __syncthreads();
            }
        }
        int32_t *P = D.vc_p + o * VC_POINTS;
        uint4 *MK4 = reinterpret_cast<uint4 *>(D.vc_mask + o * (size_t)(VC_POINTS / 2) * VC_ROWS);
        for (int g = tid; g < nb - 1; g += VC_BUILD_THREADS) {
            P[2 * g] = (int32_t)floor((S.end[g] - dom_lo) * fx_scale); P[2 * g + 1] = (int32_t)floor((S.key[g + 1] - dom_lo) * fx_scale);
            MK4[2 * g] = S.gm[g][0]; MK4[2 * g + 1] = S.gm[g][1];
        }
        // padding: above every ray (real points are < 2^29, the first ray is clamped to +-2^29: no overflow in
        // point - first_ray + round_up, and (2^30 - 2^29) >> sh exceeds the ray count by construction of sh)
        for (int i = n_points + tid; i < VC_POINTS; i += VC_BUILD_THREADS) P[i] = 0x3fffffff;
        if (tid == 0) { D.vc_meta[o * 2] = n_points; D.vc_meta[o * 2 + 1] = sh; D.vc_lo[o * 2] = dom_lo; D.vc_lo[o * 2 + 1] = fx_scale; }
        // ---- 4. coarse index: points below each 1-degree bucket start ----
        // IX[q] = first index i with p[i] >= dom_lo + q  (p[i]: even i from S.end, odd from S.key; p[n_points] = +inf).
        // Point i answers the buckets q with p[i - 1] < dom_lo + q <= p[i] (the definition's own comparisons): it marks
        // the first of them, a running maximum over the buckets fills in the rest.
        uint16_t *IX = D.vc_idx + o * VC_IDX;
        unsigned short *opens = S.newidx;   // (free after the compaction; VC_IDX <= VC_RAW)
        for (int q = tid; q < VC_IDX; q += VC_BUILD_THREADS) opens[q] = 0;
        __syncthreads();
        for (int i = 1 + tid; i <= n_points; i += VC_BUILD_THREADS) {
            const double prev = ((i - 1) & 1) ? S.key[((i - 1) >> 1) + 1] : S.end[(i - 1) >> 1];
            const double cur = i == n_points ? 1e308 : ((i & 1) ? S.key[(i >> 1) + 1] : S.end[i >> 1]);
            int q = max(0, min(VC_IDX, (int)floor(prev - dom_lo) - 1));
            while (q < VC_IDX && !(dom_lo + (double)q > prev)) ++q;
            if (q < VC_IDX && dom_lo + (double)q <= cur) opens[q] = (unsigned short)i;
        }
        __syncthreads();
        {
            constexpr int QB = (VC_IDX + VC_BUILD_THREADS - 1) / VC_BUILD_THREADS;
            int mine = 0;
#pragma unroll
            for (int u = 0; u < QB; ++u) { const int q = tid * QB + u; if (q < VC_IDX) mine = max(mine, (int)opens[q]); }
            const double before = vc_block_exscan_max((double)mine, S.red_d, tid);
            int run = before > 0.0 ? (int)before : 0;
#pragma unroll
            for (int u = 0; u < QB; ++u) {
                const int q = tid * QB + u;
                if (q < VC_IDX) { run = max(run, (int)opens[q]); IX[q] = (uint16_t)run; }
            }
        }
        __syncthreads();
    }

    for (int g = 0; g < n_guards; ++g) {
        const size_t o = (size_t)env * D.Kg + g;
        const int4 gi = *reinterpret_cast<const int4 *>(D.guard_i + o * 4);  // len, speed, range, num_rays
        const double fov = D.guard_fov[o];
        const int HS = D.L + 1;
        if (!vc_guard_cacheable(fov, gi.z, gi.w, gi.x, D.L)) { env_ok = false; if (tid == 0) D.vg_nh[o] = -1; continue; }
        // distinct headings the guard can carry: 0.0 (Guard.heading default) + the per-waypoint headings
        if (tid == 0) {
            int nh = 1;
            D.vg_hval[o * HS] = 0.0;
            for (int k = 0; k < gi.x; ++k) {
                const double h = D.guard_head[o * D.L + k];
                int s = 255;
                if (h == h) {
                    for (s = 0; s < nh; ++s) if (__double_as_longlong(D.vg_hval[o * HS + s]) == __double_as_longlong(h)) break;
                    if (s == nh) { D.vg_hval[o * HS + nh] = h; ++nh; }
                }
                D.vg_hslot[o * D.L + k] = (uint8_t)s;
            }
            D.vg_nh[o] = nh;
            S.n_bands = nh;
            // Which (waypoint, slot) pairs can the guard ever be in?  It starts on waypoint 0 with the default heading
            // (k_set_layout), a reset puts it back there with whatever heading it carries (environment.py:205-208),
            // and a move from waypoint i lands on i + stride carrying slot hslot[i] (or the old one when the move is
            // no move, security.py:145-159): a fixed point over at most L words.  Only those cones are built (a
            // patrol ring: ~11 of 32); the step kernels report a state outside the set (ERR_STATE: hand-written).
            uint32_t *reach = reinterpret_cast<uint32_t *>(S.members);   // (free after the cameras)
            const int len = gi.x, stp = len >= 2 ? py_imod(gi.y, len) : 0;
            for (int k = 0; k < len; ++k) reach[k] = 0;
            reach[0] = nh >= 32 ? 0xffffffffu : (1u << nh) - 1u;
            for (bool changed = true; changed;) {
                changed = false;
                for (int k = 0; k < len; ++k) {
                    if (!reach[k]) continue;
                    int nxt = k + stp; if (nxt >= len) nxt -= len;
                    const int s = D.vg_hslot[o * D.L + k];
                    const uint32_t add = s == 255 ? reach[k] : 1u << min(s, 31);
                    if (add & ~reach[nxt]) { reach[nxt] |= add; changed = true; }
                }
            }
            int n_list = 0;
            for (int k = 0; k < len; ++k) {
                D.vg_reach[o * D.L + k] = reach[k];
                for (int s = 0; s < nh; ++s)
                    if (((reach[k] >> min(s, 31)) & 1u) && n_list < VC_RAW) S.newidx[n_list++] = (unsigned short)(k * 64 + s);
            }
            S.n_raw = n_list;   // (never more than ~300: a waypoint other than 0 carries every slot only behind a chain of
        }                       //  no-move steps from 0, and every such step is one slot less in the path: the list's 1 024 suffice)
        __syncthreads();
        const int nh = S.n_bands;
        // The masks of all (waypoint, heading slot) pairs at once: one work item per (pair, ray), OR-ed into shared
        // memory (two 16-bit window rows per word, the layout of vg_mask).  A ray's direction depends on the slot and
        // the ray only -- not on the waypoint -- so the directions are evaluated once per (slot, ray) and shared by
        // the waypoints (same arithmetic, same values).
        const int NRAY = gi.w + 1, n_pair = S.n_raw;
        const unsigned short *pair_list = S.newidx;   // waypoint * 64 + slot
        double2 *dirs = reinterpret_cast<double2 *>(S.key);               // key[] + end[]: 2 * VC_RAW doubles, free after the cameras
        unsigned *gmask = reinterpret_cast<unsigned *>(&S.gm[0][0]);      // [pair][VC_ROWS / 2]
        constexpr int PAIR_CAP = (VC_POINTS / 2) * 32 / (VC_ROWS * 2);
        const bool shared_dirs = nh * NRAY <= VC_RAW;
        __syncthreads();
        if (shared_dirs)
            for (int i = tid; i < nh * NRAY; i += VC_BUILD_THREADS) {
                const int hs = i / NRAY, ri = i - hs * NRAY;
                double dx, dy;
                ray_dir(vc_ray_angle_deg(fov, D.vg_hval[o * HS + hs], gi.w, ri), geo.deg2rad, dx, dy);
                dirs[i] = make_double2(dx, dy);
            }
        for (int p0 = 0; p0 < n_pair; p0 += PAIR_CAP) {
            const int np = min(PAIR_CAP, n_pair - p0);
            for (int i = tid; i < np * (VC_ROWS / 2); i += VC_BUILD_THREADS) gmask[i] = 0;
            __syncthreads();
            for (int pb = tid >> 5; pb < np; pb += VC_BUILD_THREADS / 32) {   // warp = pair, lane = ray
                const int idx = pair_list[p0 + pb] >> 6, hs = pair_list[p0 + pb] & 63;
                const int row = D.guard_path[(o * D.L + idx) * 2], col = D.guard_path[(o * D.L + idx) * 2 + 1];
                unsigned *m = gmask + pb * (VC_ROWS / 2);
                for (int ri = tid & 31; ri < NRAY; ri += 32) {
                    double dx, dy;
                    if (shared_dirs) { const double2 d = dirs[hs * NRAY + ri]; dx = d.x; dy = d.y; }
                    else ray_dir(vc_ray_angle_deg(fov, D.vg_hval[o * HS + hs], gi.w, ri), geo.deg2rad, dx, dy);
                    vc_march(geo, wall, row, col, dx, dy, gi.z, 1.0, [&](int r, int c) {
                        const int wr = r - row + gi.z, wc = c - col + gi.z;
                        atomicOr(&m[wr >> 1], 1u << (wc + 16 * (wr & 1)));
                    });
                }
            }
            // the guard's own tile is always lit (visibility.py:59)
            for (int pb = tid; pb < np; pb += VC_BUILD_THREADS) atomicOr(&gmask[pb * (VC_ROWS / 2) + (gi.z >> 1)], 1u << (gi.z + 16 * (gi.z & 1)));
            __syncthreads();
            for (int i = tid; i < np * (VC_ROWS / 2); i += VC_BUILD_THREADS) {
                const int pb = i / (VC_ROWS / 2), w = i - pb * (VC_ROWS / 2);
                const int idx = pair_list[p0 + pb] >> 6, hs = pair_list[p0 + pb] & 63;
                reinterpret_cast<uint32_t *>(D.vg_mask + ((o * D.L + idx) * HS + hs) * VC_ROWS)[w] = gmask[i];
            }
            __syncthreads();
        }
    }
    if (tid == 0) {
        D.env_cached[env] = env_ok ? 1 : 0;
        if (!env_ok) { atomicAdd(D.n_uncached, 1); if (D.strict_tables) atomicOr(D.err, ERR_UNCOVERED); }
    }
}
