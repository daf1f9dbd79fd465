// heist_step.cuh -- per-step dynamics: reset, step, step_many.
// A warp owns one env's scalar state; the 4 warps of a CTA march the rays of their 4 envs together.
//
// Reference: HeistEnvironment.reset / step (environment.py:183-299), Camera.update and
// get_vision_cone_tiles (security.py:49-101), Guard.update and get_visible_tiles
// (security.py:145-192), DynamicVisibilityMap.update (visibility.py:31-65).
//
// Visibility is the reference's sample-point ray-march, not a geometric rasterisation.  Every
// sample is decided by a FILTERED-EXACT scheme:
//   fast path  - ray direction from an fp32 polynomial sincos of the (fp64-reduced) angle (error
//                < 1e-7), sample positions as 8.24 fixed-point integers x_j = x_0 + j*sx (total
//                error < 1.3e-6 tile); a sample whose fractional part lies within 2^-16 tile of a
//                rounding boundary is declared ambiguous;
//   exact path - from the first ambiguous sample on, the ray is re-evaluated exactly as the reference
//                does it: fp64 cos/sin of radians(angle) (host-libm table at multiples of 30 degrees),
//                separate fp64 multiply and add, round-half-even.  Such rays are queued and processed
//                128 at a time after the fast pass (marks are ORs, so the order does not matter).
// Unambiguous samples round to the same tile in both paths, so the result is bit-identical to the
// all-fp64 evaluation (heist_set_mode(h, 1) forces the exact path everywhere; tests compare the two).
#pragma once
#include "heist_common.cuh"

#define FX_BITS 24
#define FX_ONE (1 << FX_BITS)
#define FX_EPS 256  // 2^-16 tile

#define CELL_BLOCK 1u  // WALL, or the out-of-bounds ring around the grid
#define CELL_VIS 2u    // under surveillance this tick

#define MAP_STRIDE 256  // bytes per map row: offset = (ytile << 8) | xtile is one PRMT of the 8.24 positions
#define RING 2          // blocking cells around the grid: 1 ends every ray, the 2nd keeps the load of the NEXT sample in the map

// Per-asset (camera or guard) working record in shared memory (64 bytes).
struct __align__(16) AssetW {
    double base;     // heading - fov/2 (refreshed every tick)
    double step;     // fov / num_rays: ray pitch for the fast path
    double fov;      // exact path
    double heading;  // current heading
    int x0, y0;      // 8.24 origin in the shared cell map (slot column offset, ring, +0.5 rounding bias)
    unsigned own;    // map offset of the asset's tile (cameras: excluded from marking), ~0 for guards
    int nsamp;       // 2*range (cameras, unit 0.5) or range (guards, unit 1)
    int shift;       // 23 for cameras, 24 for guards: sx = dx * 2^shift
    int num_rays;
    int row, col;
};

// Shared-memory geometry (CTA-uniform).  Cell maps have one byte per tile plus a one-tile ring of
// blocking cells (so an out-of-bounds sample needs no bounds test) and a 256-byte row stride; the envs
// of a CTA sit side by side in the rows: slot e uses columns [colbase, colbase + Sx) of map e >> lg_spr.
struct Geo {
    int Sx;         // slot width: 32, 64 or 128 bytes (>= C + 2 * RING)
    int lg_spr;     // log2(slots per 256-byte row): 3, 2 or 1
    int n_maps;     // maps needed for HEIST_WARPS_PER_CTA slots
    int map_bytes;  // (R + 2 * RING) * 256
};

__host__ __device__ inline Geo make_geo(int R, int C) {
    Geo g;
    g.Sx = (C + 2 * RING <= 32) ? 32 : ((C + 2 * RING <= 64) ? 64 : 128);
    g.lg_spr = (g.Sx == 32) ? 3 : ((g.Sx == 64) ? 2 : 1);
    const int spr = 1 << g.lg_spr;
    g.n_maps = (HEIST_WARPS_PER_CTA + spr - 1) / spr;
    g.map_bytes = (R + 2 * RING) * MAP_STRIDE;
    return g;
}

struct WarpCtx {
    AssetW *asset;    // [Kc+Kg] cameras first, then guards
    int4 *g_i;        // [Kg] len, speed, range, num_rays
    double *speed;    // [Kc] camera rotation speed
    int *rpre;        // [Kc+Kg+1] prefix sums of rays (num_rays+1) per asset
    int *g_idx;       // [Kg]
    uint8_t *cell;    // this slot's window of the shared cell map: cell[(r+RING)*256 + c + RING]
    unsigned cell_sa; // shared-space address of the MAP (not the window): map + (ytile<<8 | xtile)
    int colbase;      // column offset of this slot inside the map rows
};

__host__ __device__ inline size_t warp_ctx_bytes(int R, int C, int Kc, int Kg) {
    size_t b = 0;
    b += (size_t)(Kc + Kg) * sizeof(AssetW);
    b += (size_t)Kg * sizeof(int4);
    b += (size_t)Kc * sizeof(double);
    b += ((size_t)(Kc + Kg + 1) * sizeof(int) + (size_t)Kg * sizeof(int) + 15) & ~(size_t)15;
    return (b + 15) & ~(size_t)15;
}

#ifdef HEIST_DEBUG_BOUNDS
// Debug build (lib/libheist_b200_dbg.so): every cell-map access of the march is range-checked against the
// CTA's dynamic shared memory; a violation raises the sticky ERR_BOUNDS flag instead of touching memory.
// (compute-sanitizer is not available on the GPU pool; tests/test_gpu_parity.py runs this build.)
__device__ int *g_dbg_err;
__device__ unsigned g_dbg_smem_lo, g_dbg_smem_hi;
__device__ __forceinline__ bool dbg_ok(unsigned sa) {
    if (sa >= g_dbg_smem_lo && sa < g_dbg_smem_hi) return true;
    atomicOr(g_dbg_err, ERR_BOUNDS);
    return false;
}
#else
__device__ __forceinline__ bool dbg_ok(unsigned) { return true; }
#endif

__device__ __forceinline__ unsigned lds_u8(unsigned sa) {
    unsigned v = CELL_BLOCK;
    if (dbg_ok(sa)) asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(sa) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u8(unsigned sa, unsigned v) {
    if (dbg_ok(sa)) asm volatile("st.shared.u8 [%0], %1;" ::"r"(sa), "r"(v) : "memory");
}
// (ytile << 8) | xtile from the two 8.24 positions (tile = top byte; ytile < 128 so the sign-replicating
// selector 0xF yields zero bytes)
__device__ __forceinline__ unsigned tile_offset(unsigned x, unsigned y) {
    unsigned d;
    asm("prmt.b32 %0, %1, %2, 0xFF73;" : "=r"(d) : "r"(x), "r"(y));
    return d;
}

struct EnvRegs {
    int r, c, tick, prev, init, flags, n_vault, n_detect, n_timeout;
};

#define RINT_MAGIC 6755399441055744.0  // 2^52 + 2^51: (x + M) has rint(x) in its low 32 bits (RN-even)

__device__ __forceinline__ int rint_even(double x) { return __double2loint(__dadd_rn(x, RINT_MAGIC)); }

// Exact continuation of ray `ri` of asset `k` from sample j on (security.py:69-99 / 170-190):
// marks cells until the first out-of-bounds (ring) or WALL sample.
__device__ __noinline__ void ray_exact(const AssetW *asset, int k, unsigned map_sa, int colbase, double deg2rad, int ri, int j) {
    const AssetW A = asset[k];  // private copy: no re-reads of shared memory in the loop
    const double half_fov = __ddiv_rn(A.fov, 2.0);
    const double angle_deg =
        __dadd_rn(__dsub_rn(A.heading, half_fov), __ddiv_rn(__dmul_rn(A.fov, (double)ri), (double)A.num_rays));
    double dx, dy;
    ray_dir(angle_deg, deg2rad, dx, dy);
    const double unit = (A.shift == 23) ? 0.5 : 1.0;
    const double dcol = (double)A.col, drow = (double)A.row;
    double dist = unit * (double)j;  // exact (multiples of 0.5)
    for (; j <= A.nsamp; ++j, dist += unit) {
        int c = rint_even(__dadd_rn(dcol, __dmul_rn(dx, dist)));
        int r = rint_even(__dadd_rn(drow, __dmul_rn(dy, dist)));
        // consecutive samples move by at most one tile per axis, so the first out-of-bounds sample
        // always lands on the blocking ring
        unsigned off = (unsigned)(((r + RING) << 8) + c + RING + colbase);
        if (lds_u8(map_sa + off) == CELL_BLOCK) return;
        if (off != A.own) sts_u8(map_sa + off, CELL_VIS);
    }
}

// ---------------------------------------------------------------------------------------------
// CTA-cooperative ray-march.  The rays of all envs of the CTA are marched by all of its warps:
// every phase the owners publish how many 32-ray chunks their env needs, and the chunks of every env are
// dealt round-robin over the warps.  (Envs differ several-fold in rays per tick -- and a reset
// doubles a tick's work -- so a one-warp-per-env march leaves most of the SM idle behind its slowest
// warp.)
// ---------------------------------------------------------------------------------------------
#define PEND_CAP 480
struct CtaCtl {
    int taken;                      // exact-path queue entries already claimed by a draining warp
    int pend_n;                     // rays handed to the exact path this phase
    int cnt[HEIST_WARPS_PER_CTA];   // 32-ray chunks wanted by each env slot this phase (0: none)
    int n_assets[HEIST_WARPS_PER_CTA];
    int active[HEIST_WARPS_PER_CTA];  // owner still has ticks to run (step_many)
    int pad[2];
    unsigned pend[PEND_CAP];        // slot:2 | asset:6 | sample j:8 | ray:16
};

__host__ __device__ inline size_t cta_smem_bytes(int R, int C, int Kc, int Kg) {
    const Geo g = make_geo(R, C);
    return sizeof(CtaCtl) + (size_t)g.n_maps * g.map_bytes + HEIST_WARPS_PER_CTA * warp_ctx_bytes(R, C, Kc, Kg);
}

// smem layout: [CtaCtl][n_maps cell maps][slot 0 ctx][slot 1 ctx]...
__device__ __forceinline__ WarpCtx carve_warp_ctx(unsigned char *smem, const Geo &g, int slot, int R, int C, int Kc, int Kg) {
    WarpCtx S;
    unsigned char *maps = smem + sizeof(CtaCtl);
    unsigned char *p = maps + (size_t)g.n_maps * g.map_bytes + (size_t)slot * warp_ctx_bytes(R, C, Kc, Kg);
    S.asset = (AssetW *)p;  p += (size_t)(Kc + Kg) * sizeof(AssetW);
    S.g_i = (int4 *)p;      p += (size_t)Kg * sizeof(int4);
    S.speed = (double *)p;  p += (size_t)Kc * sizeof(double);
    S.rpre = (int *)p;      S.g_idx = S.rpre + (Kc + Kg + 1);
    p += ((size_t)(Kc + Kg + 1) * sizeof(int) + (size_t)Kg * sizeof(int) + 15) & ~(size_t)15;
    unsigned char *map = maps + (size_t)(slot >> g.lg_spr) * g.map_bytes;
    S.colbase = (slot & ((1 << g.lg_spr) - 1)) * g.Sx;
    S.cell = map + S.colbase;
    S.cell_sa = (unsigned)__cvta_generic_to_shared(map);
    return S;
}

__device__ __forceinline__ void dbg_init(const Dev &D, unsigned char *smem, const Geo &geo) {
#ifdef HEIST_DEBUG_BOUNDS
    if (threadIdx.x == 0) {  // every CTA writes the same values (shared-window addresses are CTA-independent)
        g_dbg_err = D.err;
        g_dbg_smem_lo = (unsigned)__cvta_generic_to_shared(smem) + (unsigned)sizeof(CtaCtl);
        g_dbg_smem_hi = g_dbg_smem_lo + (unsigned)(geo.n_maps * geo.map_bytes);
    }
    __syncthreads();
#endif
}

// One fast-path sample, software-pipelined by one: the tile offset, the cell byte and the ambiguity key of sample J
// were produced one step earlier (offn / celln / ambn); this step first issues the same for sample J+1 -- the
// second ring of blocking cells keeps that speculative load inside the map even when sample J is the one that
// ends the ray -- and only then decides sample J, so the shared-memory latency overlaps the decision.
// `J` is the 1-based sample index (a literal in the unrolled variants).  `break`s out of the enclosing loop when
// the ray ends: blocked, or ambiguous (amb_j = J; the rare hand-over to the exact path is done once, after the
// loop, to keep the unrolled code small).
#define HEIST_ADVANCE()                                                                                               \
    x += sx; y += sy;                                                                                                 \
    offn = tile_offset(x, y);                                                                                         \
    celln = lds_u8(map_sa + offn);                                                                                    \
    /* frac within 2^-16 of the rounding boundary <=> ((pos + EPS) << 8) < (2 EPS << 8) as u32 */                    \
    ambn = min(x * 256u + (FX_EPS << 8), y * 256u + (FX_EPS << 8));
#define HEIST_SAMPLE(J, CHECK_OWN, LAST)                                                                              \
    {                                                                                                                 \
        const unsigned offc = offn, cellc = celln, ambc = ambn;                                                       \
        if (!(LAST)) { HEIST_ADVANCE() }                                                                              \
        /* one exit branch per sample; blocking cells are never marked, so == suffices */                            \
        if ((ambc < (2u * FX_EPS << 8)) | (cellc == CELL_BLOCK)) { if (ambc < (2u * FX_EPS << 8)) amb_j = (J); break; } \
        if (!(CHECK_OWN) || offc != own) sts_u8(map_sa + offc, CELL_VIS);                                             \
    }

// March flattened rays [f0, f0+32) of env slot e (asset table `asset`, ray prefix sums `rpre`).
// `k` is the lane's asset cursor: a warp walks an env's chunks in increasing order, so the scan resumes.
template <bool EXACT_ONLY>
__device__ __forceinline__ void march_chunk(CtaCtl *ctl, int e, const AssetW *asset, const int *rpre, int total_rays, int f0,
                                            int lane, unsigned map_sa, int colbase, double deg2rad, int &k) {
    const int f = f0 + lane;
    if (f >= total_rays) return;
    while (f >= rpre[k + 1]) ++k;
    const int ri = f - rpre[k];
    if (EXACT_ONLY) { ray_exact(asset, k, map_sa, colbase, deg2rad, ri, 1); return; }
    const AssetW &A = asset[k];
    const int nsamp = A.nsamp;
    const unsigned own = A.own;
    // ---- fast path: direction (error < 1e-7) ----
    const double a = fma((double)ri, A.step, A.base);                // degrees
    const double t = fma(a, 1.0 / 90.0, RINT_MAGIC);
    const int q = __double2loint(t);                                 // nearest quadrant
    const float r = (float)(fma(t - RINT_MAGIC, -90.0, a) * 0.017453292519943295);  // |r| <= pi/4 (+eps)
    const float r2 = r * r;
    float sn = fmaf(r * r2, fmaf(r2, fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f), -1.6666654611e-1f), r);
    float cs = fmaf(r2 * r2, fmaf(r2, fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f), 4.166664568298827e-2f),
                    fmaf(r2, -0.5f, 1.0f));
    float c_a = (q & 1) ? sn : cs;   // cos(q*90 + r)
    float s_a = (q & 1) ? cs : sn;   // sin(q*90 + r)
    if ((q + 1) & 2) c_a = -c_a;
    if (q & 2) s_a = -s_a;
    const float scale = (A.shift == 23) ? 8388608.0f : 16777216.0f;
    const unsigned sx = (unsigned)__float2int_rn(c_a * scale);    // dx =  cos
    const unsigned sy = (unsigned)__float2int_rn(-s_a * scale);   // dy = -sin
    // ---- samples: 8.24 fixed point, tile = floor(pos + 0.5) unless within 2^-16 of a boundary ----
    // Only a camera's first sample (dist 0.5) can round to the camera's own tile (security.py:93).
    unsigned x = (unsigned)A.x0, y = (unsigned)A.y0;
    int amb_j = 0;
    unsigned offn, celln, ambn;
    if (nsamp >= 1) { HEIST_ADVANCE() }   // sample 1
    if (nsamp == 12) {          // camera, vision_range 6 (the Architect's cameras, networks.py:301)
        do {
            HEIST_SAMPLE(1, true, false) HEIST_SAMPLE(2, false, false) HEIST_SAMPLE(3, false, false)
            HEIST_SAMPLE(4, false, false) HEIST_SAMPLE(5, false, false) HEIST_SAMPLE(6, false, false)
            HEIST_SAMPLE(7, false, false) HEIST_SAMPLE(8, false, false) HEIST_SAMPLE(9, false, false)
            HEIST_SAMPLE(10, false, false) HEIST_SAMPLE(11, false, false) HEIST_SAMPLE(12, false, true)
        } while (0);
    } else if (nsamp == 4) {    // guard, vision_range 4 (networks.py:311)
        do {
            HEIST_SAMPLE(1, true, false) HEIST_SAMPLE(2, false, false) HEIST_SAMPLE(3, false, false)
            HEIST_SAMPLE(4, false, true)
        } while (0);
    } else if (nsamp >= 1) {
        for (int j = 1; j <= nsamp; ++j) { HEIST_SAMPLE(j, true, j == nsamp) }
    }
    if (amb_j) {  // queue the rest of the ray for the exact path (processed 128 rays at a time after the barrier)
        int slot = (amb_j < 256) ? atomicAdd(&ctl->pend_n, 1) : PEND_CAP;
        if (slot < PEND_CAP)
            ctl->pend[slot] = ((unsigned)e << 30) | ((unsigned)k << 24) | ((unsigned)amb_j << 16) | (unsigned)ri;
        else ray_exact(asset, k, map_sa, colbase, deg2rad, ri, amb_j);
    }
}

#define PEND_EMPTY 0xffffffffu
// Take queue entry i (spin until its producer has written it), hand it back as empty, run the exact path.
template <bool EXACT_ONLY>
__device__ __forceinline__ void drain_one(const Dev &D, const WarpCtx &S, const Geo &geo, int warp, int ctx_bytes,
                                          CtaCtl *ctl, unsigned map0_sa, int i) {
    volatile unsigned *pe = &ctl->pend[i];
    unsigned p;
    while ((p = *pe) == PEND_EMPTY) {}
    *pe = PEND_EMPTY;
    const int e = p >> 30, k = (p >> 24) & 63, j = (p >> 16) & 255, ri = p & 0xffff;
    const int delta = (e - warp) * ctx_bytes;
    const AssetW *asset = reinterpret_cast<const AssetW *>(reinterpret_cast<const unsigned char *>(S.asset) + delta);
    const unsigned map_sa = map0_sa + (unsigned)((e >> geo.lg_spr) * geo.map_bytes);
    ray_exact(asset, k, map_sa, (e & ((1 << geo.lg_spr) - 1)) * geo.Sx, D.deg2rad, ri, j);
}

// One cooperative phase: all warps of the CTA drain the chunk queue, then (after a barrier) the
// rays that were handed to the exact path.  Must be entered by every warp after a __syncthreads()
// that follows the owners' writes to ctl / cell maps; ends with a __syncthreads().
template <bool EXACT_ONLY>
__device__ __forceinline__ void march_phase(const Dev &D, const WarpCtx &S, const Geo &geo, int warp, int ctx_bytes,
                                            CtaCtl *ctl, int lane) {
    const unsigned map0_sa = S.cell_sa - (unsigned)((warp >> geo.lg_spr) * geo.map_bytes);
    // Static deal: chunk c of env slot e goes to warp (c + e) & 3.  No shared counter (ptxas turns a
    // shared atomicAdd into a ~20-instruction aggregate), no per-chunk slot decode, and the per-env pointers
    // and the lanes' asset cursors are set up once per env instead of once per chunk.
#pragma unroll 1
    for (int e = 0; e < HEIST_WARPS_PER_CTA; ++e) {
        const int cnt = ctl->cnt[e];
        if (cnt == 0) continue;
        // every slot's context has the same layout: shift this warp's own pointers by whole contexts
        const int delta = (e - warp) * ctx_bytes;
        const AssetW *asset = reinterpret_cast<const AssetW *>(reinterpret_cast<const unsigned char *>(S.asset) + delta);
        const int *rpre = reinterpret_cast<const int *>(reinterpret_cast<const unsigned char *>(S.rpre) + delta);
        const unsigned map_sa = map0_sa + (unsigned)((e >> geo.lg_spr) * geo.map_bytes);
        const int colbase = (e & ((1 << geo.lg_spr) - 1)) * geo.Sx;
        const int total_rays = rpre[ctl->n_assets[e]];
        int k = 0;
        for (int c = (warp - e) & (HEIST_WARPS_PER_CTA - 1); c < cnt; c += HEIST_WARPS_PER_CTA)
            march_chunk<EXACT_ONLY>(ctl, e, asset, rpre, total_rays, c * 32, lane, map_sa, colbase, D.deg2rad, k);
    }
    if (!EXACT_ONLY) {
        // Opportunistic drain of the exact-path queue by warps that ran out of chunks: the long fp64 dependency
        // chains of ray_exact overlap the other warps' marching instead of sitting between two barriers.
        // An entry is its own ready flag (PEND_EMPTY until the producer's store lands).
        for (;;) {
            int old = 0, cnt = 0;
            if (lane == 0) {
                old = *(volatile int *)&ctl->taken;
                cnt = min(32, min(*(volatile int *)&ctl->pend_n, PEND_CAP) - old);
                if (cnt > 0 && atomicCAS(&ctl->taken, old, old + cnt) != old) cnt = -1;  // lost the race: retry
            }
            old = __shfl_sync(0xffffffffu, old, 0);
            cnt = __shfl_sync(0xffffffffu, cnt, 0);
            if (cnt == 0) break;
            if (cnt > 0 && lane < cnt) drain_one<EXACT_ONLY>(D, S, geo, warp, ctx_bytes, ctl, map0_sa, old + lane);
        }
    }
    __syncthreads();
    if (!EXACT_ONLY) {
        // entries pushed after the last look above (CTA-uniform values: read after the barrier)
        const int P = min(ctl->pend_n, PEND_CAP), T0 = ctl->taken;
        for (int i = T0 + threadIdx.x; i < P; i += HEIST_WARPS_PER_CTA * 32)
            drain_one<EXACT_ONLY>(D, S, geo, warp, ctx_bytes, ctl, map0_sa, i);
        if (P > T0) __syncthreads();
    }
}

// Rows an env's rays can touch this rebuild: |row - asset row| <= vision range (a sample at distance d
// moves at most d rows).  Bit r of the mask = grid row r.
__device__ __forceinline__ unsigned long long dirty_rows(const WarpCtx &S, int R, int lane, int n_assets) {
    unsigned lo = 0, hi = 0;
    if (lane < n_assets) {
        const AssetW &A = S.asset[lane];
        const int rr = (A.shift == 23) ? (A.nsamp + 1) >> 1 : A.nsamp;
        const int r0 = max(A.row - rr, 0), r1 = min(A.row + rr, R - 1);
        const unsigned long long m = ((r1 - r0 + 1 >= 64) ? ~0ull : ((1ull << (r1 - r0 + 1)) - 1ull)) << r0;
        lo = (unsigned)m; hi = (unsigned)(m >> 32);
    }
    lo = __reduce_or_sync(0xffffffffu, lo);
    hi = __reduce_or_sync(0xffffffffu, hi);
    return ((unsigned long long)hi << 32) | lo;
}

// Owner-side preparation of a visibility rebuild (DynamicVisibilityMap.update, visibility.py:31-65):
// strip CELL_VIS from the rows the previous rebuild could have marked (`dirty`, in/out: replaced by the rows
// of this rebuild) and publish the chunk count.
template <bool BIG>
__device__ __forceinline__ void begin_visibility(const WarpCtx &S, const Geo &geo, int R, CtaCtl *ctl, int warp, int lane,
                                                 int n_assets, unsigned long long &dirty) {
    const int per_row = geo.Sx >> 4;  // int4 per window row
    for (int i = lane; i < R * per_row; i += 32) {
        const int row = i / per_row, q = i - row * per_row;
        if (BIG && !((dirty >> row) & 1ull)) continue;
        int4 *p = reinterpret_cast<int4 *>(S.cell + (row + RING) * MAP_STRIDE + q * 16);
        int4 v = *p;
        v.x &= 0x01010101; v.y &= 0x01010101; v.z &= 0x01010101; v.w &= 0x01010101;
        *p = v;
    }
    if (BIG) dirty = dirty_rows(S, R, lane, n_assets);  // grids up to 32 rows: nearly every row is in range anyway
    if (lane == 0) ctl->cnt[warp] = (S.rpre[n_assets] + 31) >> 5;
}

// ... and its owner-side end, after the cooperative march: the guards' own tiles are always lit
// (visibility.py:59).  Done last because a guard may stand on a WALL tile, which must keep blocking
// sight (the march tests cell == CELL_BLOCK).
__device__ __forceinline__ void end_visibility(const WarpCtx &S, int lane, int n_cams, int n_assets) {
    if (lane >= n_cams && lane < n_assets) {
        const AssetW &A = S.asset[lane];
        S.cell[(A.row + RING) * MAP_STRIDE + A.col + RING] |= CELL_VIS;
    }
    __syncwarp();
}

// per-tick refresh of an asset record after its heading / position changed
__device__ __forceinline__ void refresh_asset(const WarpCtx &S, AssetW &A, bool is_cam) {
    A.base = A.heading - A.fov * 0.5;
    A.x0 = (int)(((unsigned)(A.col + RING + S.colbase) << FX_BITS) + (FX_ONE >> 1));
    A.y0 = (int)(((unsigned)(A.row + RING) << FX_BITS) + (FX_ONE >> 1));
    A.own = is_cam ? (unsigned)(((A.row + RING) << 8) + A.col + RING + S.colbase) : ~0u;
}

__device__ __forceinline__ void load_env(const Dev &D, const WarpCtx &S, const Geo &geo, int env, int lane, EnvRegs &E,
                                         int &n_cams, int &n_guards) {
    const int4 es = *reinterpret_cast<const int4 *>(D.env_s + (size_t)env * 4);
    n_cams = es.x;
    n_guards = es.y;
    const int4 d0 = *reinterpret_cast<const int4 *>(D.env_d + (size_t)env * 8);
    const int4 d1 = *reinterpret_cast<const int4 *>(D.env_d + (size_t)env * 8 + 4);
    E.r = d0.x & 0xffff; E.c = d0.x >> 16; E.tick = d0.y; E.prev = d0.z; E.init = d0.w;
    E.flags = d1.x & 0xff; E.n_vault = d1.y; E.n_detect = d1.z; E.n_timeout = d1.w;
    // cell window with a blocking ring: CELL_BLOCK from the wall bitmap (everything outside the grid blocks),
    // CELL_VIS from the stored visibility.  One lane expands 32 window bytes at a time: the row bitmaps are
    // shifted by the ring width, cut into nibbles, and each nibble becomes 4 cell bytes with one multiply.
    const uint32_t *wall = D.wall + (size_t)env * D.RW, *vis = D.vis + (size_t)env * D.RW;
    const int gpr = geo.Sx >> 5;  // 32-byte groups per window row
    for (int u = lane; u < (D.R + 2 * RING) * gpr; u += 32) {
        const int row = u / gpr, g = u - row * gpr;
        const int rr = row - RING;
        uint32_t wbits = 0xffffffffu, vbits = 0u;
        if (rr >= 0 && rr < D.R) {
            unsigned long long w64 = wall[rr * D.W], v64 = vis[rr * D.W];
            if (D.W > 1) { w64 |= (unsigned long long)wall[rr * D.W + 1] << 32; v64 |= (unsigned long long)vis[rr * D.W + 1] << 32; }
            // 128-bit (hi:lo) = bitmap << RING, then bits outside [RING, RING + C) forced to "blocking, not visible"
            const unsigned long long wlo = w64 << RING, whi = w64 >> (64 - RING), vlo = v64 << RING, vhi = v64 >> (64 - RING);
            const int lo_bit = g * 32;  // window bytes [lo_bit, lo_bit + 32)
            const uint32_t wsel = lo_bit == 0 ? (uint32_t)wlo : lo_bit == 32 ? (uint32_t)(wlo >> 32) : lo_bit == 64 ? (uint32_t)whi : (uint32_t)(whi >> 32);
            const uint32_t vsel = lo_bit == 0 ? (uint32_t)vlo : lo_bit == 32 ? (uint32_t)(vlo >> 32) : lo_bit == 64 ? (uint32_t)vhi : (uint32_t)(vhi >> 32);
            // inside mask for this group: window positions p with RING <= p < RING + C
            const int a0 = max(RING - lo_bit, 0), a1 = min(RING + D.C - lo_bit, 32);
            const uint32_t inside = a1 > a0 ? ((a1 - a0 >= 32 ? 0xffffffffu : ((1u << (a1 - a0)) - 1u)) << a0) : 0u;
            wbits = (wsel & inside) | ~inside;
            vbits = vsel & inside;
        }
        uint32_t out[8];
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const uint32_t wn = (wbits >> (4 * n)) & 15u, vn = (vbits >> (4 * n)) & 15u;
            out[n] = ((wn * 0x00204081u) & 0x01010101u) | (((vn * 0x00204081u) & 0x01010101u) << 1);
        }
        uint4 *dst = reinterpret_cast<uint4 *>(S.cell + row * MAP_STRIDE + g * 32);
        dst[0] = make_uint4(out[0], out[1], out[2], out[3]);
        dst[1] = make_uint4(out[4], out[5], out[6], out[7]);
    }
    if (lane < n_cams) {
        size_t o = (size_t)env * D.Kc + lane;
        AssetW &A = S.asset[lane];
        const int16_t *ci = D.cam_i + o * 4;
        A.fov = D.cam_f[o * 2];
        S.speed[lane] = D.cam_f[o * 2 + 1];
        A.heading = D.cam_heading[o];
        A.row = ci[0]; A.col = ci[1]; A.nsamp = 2 * ci[2]; A.shift = 23; A.num_rays = ci[3];
        A.step = A.fov / (double)A.num_rays;
        refresh_asset(S, A, true);
    }
    if (lane < n_guards) {
        size_t o = (size_t)env * D.Kg + lane;
        AssetW &A = S.asset[n_cams + lane];
        const int4 gi = *reinterpret_cast<const int4 *>(D.guard_i + o * 4);
        S.g_i[lane] = gi;
        A.fov = D.guard_fov[o];
        A.heading = D.guard_heading[o];
        int idx = D.guard_idx[o];
        S.g_idx[lane] = idx;
        const uint8_t *p = D.guard_path + (o * D.L + idx) * 2;
        A.row = p[0]; A.col = p[1]; A.nsamp = gi.z; A.shift = 24; A.num_rays = gi.w;
        A.step = A.fov / (double)A.num_rays;
        refresh_asset(S, A, false);
    }
    __syncwarp();
    if (lane == 0) {
        int acc = 0;
        S.rpre[0] = 0;
        for (int k = 0; k < n_cams + n_guards; ++k) { acc += S.asset[k].num_rays + 1; S.rpre[k + 1] = acc; }
    }
    __syncwarp();
}

// bit i of the result = CELL_VIS of byte p[i], i < 32 (p 16-byte aligned)
__device__ __forceinline__ unsigned pack32(const uint8_t *p) {
    const uint4 a = *reinterpret_cast<const uint4 *>(p), b = *reinterpret_cast<const uint4 *>(p + 16);
    unsigned m = 0;
    // (w >> 1) & 0x01010101 isolates the CELL_VIS bit of 4 cells; * 0x01020408 gathers them into the top nibble
#define HEIST_NIB(w, sh) m |= (((((w) >> 1) & 0x01010101u) * 0x01020408u) >> 24) << (sh);
    HEIST_NIB(a.x, 0) HEIST_NIB(a.y, 4) HEIST_NIB(a.z, 8) HEIST_NIB(a.w, 12)
    HEIST_NIB(b.x, 16) HEIST_NIB(b.y, 20) HEIST_NIB(b.z, 24) HEIST_NIB(b.w, 28)
#undef HEIST_NIB
    return m;
}

// row bitmaps of the visibility bits of the cell window -> dst[RW]; lane = grid row, coalesced store.
// (The window's first byte is the ring, so the packed row is shifted down by one; ring and padding
// bytes never carry CELL_VIS.)  Rows outside `dirty` cannot be lit and are written as zero unread.
template <bool BIG>
__device__ __forceinline__ void pack_vis(const Dev &D, const WarpCtx &S, const Geo &geo, int lane, uint32_t *dst,
                                         unsigned long long dirty) {
    for (int r = lane; r < D.R; r += 32) {
        unsigned w0 = 0, w1 = 0;
        if (!BIG || ((dirty >> r) & 1ull)) {
            const uint8_t *row = S.cell + (r + RING) * MAP_STRIDE;
            const unsigned lo = pack32(row);
            const unsigned mid = geo.Sx > 32 ? pack32(row + 32) : 0u;
            w0 = (lo >> RING) | (mid << (32 - RING));
            if (D.W > 1) {
                const unsigned hi = geo.Sx > 64 ? pack32(row + 64) : 0u;
                w1 = (mid >> RING) | (hi << (32 - RING));
            }
        }
        dst[r * D.W] = w0;
        if (D.W > 1) dst[r * D.W + 1] = w1;
    }
}

template <bool BIG>
__device__ __forceinline__ void store_env(const Dev &D, const WarpCtx &S, const Geo &geo, int env, int lane,
                                          const EnvRegs &E, int status, int n_cams, int n_guards, unsigned long long dirty) {
    if (lane == 0) {
        int4 d0 = make_int4(E.r | (E.c << 16), E.tick, E.prev, E.init);
        int4 d1 = make_int4(E.flags | (status << 8), E.n_vault, E.n_detect, E.n_timeout);
        *reinterpret_cast<int4 *>(D.env_d + (size_t)env * 8) = d0;
        *reinterpret_cast<int4 *>(D.env_d + (size_t)env * 8 + 4) = d1;
    }
    if (lane < n_cams) D.cam_heading[(size_t)env * D.Kc + lane] = S.asset[lane].heading;
    if (lane < n_guards) {
        size_t o = (size_t)env * D.Kg + lane;
        D.guard_heading[o] = S.asset[n_cams + lane].heading;
        D.guard_idx[o] = S.g_idx[lane];
    }
    pack_vis<BIG>(D, S, geo, lane, D.vis + (size_t)env * D.RW, dirty);
}

// HeistEnvironment.reset (environment.py:183-214), owner part: solver to start, guards to waypoint 0;
// camera and guard headings persist.  The visibility rebuild follows as a cooperative phase.
__device__ __forceinline__ void reset_state(const Dev &D, const WarpCtx &S, int env, int lane, EnvRegs &E, int n_cams,
                                            int n_guards) {
    E.r = D.start_r; E.c = D.start_c; E.tick = 0;
    E.flags = 0;
    E.prev = abs(E.r - D.vault_r) + abs(E.c - D.vault_c);
    E.init = E.prev;
    if (lane < n_guards) {
        S.g_idx[lane] = 0;
        const uint8_t *p = D.guard_path + ((size_t)env * D.Kg + lane) * D.L * 2;
        AssetW &A = S.asset[n_cams + lane];
        A.row = p[0]; A.col = p[1];
        refresh_asset(S, A, false);
    }
    __syncwarp();
}

// HeistEnvironment.step (environment.py:216-299), owner part before the visibility rebuild:
// move (:239-246), cameras rotate (security.py:49-51), guards advance (security.py:145-159).
__device__ __forceinline__ void step_begin(const Dev &D, const WarpCtx &S, int env, int lane, EnvRegs &E, int n_cams,
                                           int n_guards, int action) {
    int nr = E.r + (action == 2) - (action == 1);
    int nc = E.c + (action == 4) - (action == 3);
    if (!(S.cell[(nr + RING) * MAP_STRIDE + nc + RING] & CELL_BLOCK)) { E.r = nr; E.c = nc; }  // ring = out of bounds
    if (lane < n_cams) {
        AssetW &A = S.asset[lane];
        A.heading = py_mod360(__dadd_rn(A.heading, S.speed[lane]));
        A.base = A.heading - A.fov * 0.5;
    }
    if (lane < n_guards) {
        int4 gi = S.g_i[lane];
        if (gi.x >= 2) {
            AssetW &A = S.asset[n_cams + lane];
            int old = S.g_idx[lane];
            int ni = py_imod(old + gi.y, gi.x);
            size_t o = ((size_t)env * D.Kg + lane) * D.L;
            double h = D.guard_head[o + old];
            if (h == h) A.heading = h;  // NaN <=> the move is (0,0): heading unchanged
            S.g_idx[lane] = ni;
            const uint8_t *p = D.guard_path + (o + ni) * 2;
            A.row = p[0]; A.col = p[1];
            refresh_asset(S, A, false);
        }
    }
    __syncwarp();
}

// ... and after it: shaping (:261-269), detection (:273-281), vault (:284-288), timeout (:291-297).
__device__ __forceinline__ int step_finish(const Dev &D, const WarpCtx &S, EnvRegs &E, double &reward_out) {
    double reward = D.reward_step;  // :235
    int status = HEIST_RUNNING;
    int curr = abs(E.r - D.vault_r) + abs(E.c - D.vault_c);
    reward = __dadd_rn(reward, __dmul_rn((double)(E.prev - curr), 0.1));
    E.prev = curr;
    if (curr <= 3 && E.init > 3) reward = __dadd_rn(reward, __dmul_rn(0.05, (double)(3 - curr)));
    if (S.cell[(E.r + RING) * MAP_STRIDE + E.c + RING] & CELL_VIS) {
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
// Warp slot -> env through D.slot2env (cost-balanced order built after every set_layout).
template <bool EXACT_ONLY, bool BIG>  // BIG: more than 32 grid rows (dirty-row tracking pays off)
__global__ void __launch_bounds__(HEIST_WARPS_PER_CTA * 32, 8)
k_step_many(Dev D, const int8_t *__restrict__ actions, int T, int autoreset, float *__restrict__ reward,
            double *__restrict__ reward64, uint8_t *__restrict__ done, uint8_t *__restrict__ status_out,
            uint32_t *__restrict__ vis_traj) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    CtaCtl *ctl = reinterpret_cast<CtaCtl *>(smem);
    const Geo geo = make_geo(D.R, D.C);
    const int ctx_bytes = (int)warp_ctx_bytes(D.R, D.C, D.Kc, D.Kg);
    dbg_init(D, smem, geo);
    for (int i = threadIdx.x; i < PEND_CAP; i += HEIST_WARPS_PER_CTA * 32) ctl->pend[i] = PEND_EMPTY;
    if (D.skip_cached && *D.n_uncached == 0) return;  // every env is served by k_fast (heist_fast.cuh)
    int env = D.slot2env[blockIdx.x * HEIST_WARPS_PER_CTA + warp];
    if (D.skip_cached && env >= 0 && D.env_cached[env]) env = -1;
    const bool have = env >= 0;
    const WarpCtx S = carve_warp_ctx(smem, geo, warp, D.R, D.C, D.Kc, D.Kg);
    EnvRegs E;
    E.flags = F_DONE;
    int n_cams = 0, n_guards = 0;
    if (have) load_env(D, S, geo, env, lane, E, n_cams, n_guards);
    const int n_assets = n_cams + n_guards;
    if (lane == 0) ctl->n_assets[warp] = n_assets;
    int status = HEIST_RUNNING;
    unsigned long long dirty = ~0ull;  // rows that may carry CELL_VIS (everything, for the map loaded from HBM)
    int work = 0;  // 32-ray chunks this env asked for during the launch (measured load-balance cost)
    const int chunks = have ? (S.rpre[n_assets] + 31) >> 5 : 0;
    // Every loop iteration is one cooperative visibility rebuild for the whole CTA, but the envs of a CTA are
    // NOT in lock-step: each owner advances its own tick `t` and asks for whatever rebuild it needs next -- the
    // one of its next step, or the one of the auto-reset that follows a finished episode (the trainer's
    // `if done: reset()`).  Iterations per launch = max over the CTA's envs of (T + resets) instead of
    // T + (ticks in which any of them resets), and the march code exists once (instruction cache).
    int t = 0;
    bool pending_reset = false;
    for (;;) {
        const bool active = have && t < T;
        int kind = 0;  // 0: no rebuild (env already done / finished), 1: step, 2: reset
        if (active) {
            if (pending_reset) { reset_state(D, S, env, lane, E, n_cams, n_guards); kind = 2; }
            else if (!(E.flags & F_DONE)) {  // a done env is not mutated (:232-233)
                step_begin(D, S, env, lane, E, n_cams, n_guards, actions[(size_t)t * D.N + env]);
                kind = 1;
            }
        }
        if (kind) { begin_visibility<BIG>(S, geo, D.R, ctl, warp, lane, n_assets, dirty); work += chunks; }
        else if (lane == 0) ctl->cnt[warp] = 0;
        if (lane == 0) ctl->active[warp] = active;
        if (threadIdx.x == 0) { ctl->pend_n = 0; ctl->taken = 0; }
        __syncthreads();
        if (!(ctl->active[0] | ctl->active[1] | ctl->active[2] | ctl->active[3])) break;  // CTA-uniform
        march_phase<EXACT_ONLY>(D, S, geo, warp, ctx_bytes, ctl, lane);
        if (active) {
            const size_t o = (size_t)t * D.N + env;
            if (kind) end_visibility(S, lane, n_cams, n_assets);
            if (kind != 2) {
                // ---- owner: rewards, termination, outputs of step t ----
                double rw = 0.0;
                status = HEIST_ALREADY_DONE;
                if (kind == 1) status = step_finish(D, S, E, rw);
                if (lane == 0) {
                    if (reward) reward[o] = (float)rw;
                    if (reward64) reward64[o] = rw;
                    if (done) done[o] = (E.flags & F_DONE) ? 1 : 0;
                    if (status_out) status_out[o] = (uint8_t)status;
                }
                pending_reset = autoreset && (E.flags & F_DONE);
            } else pending_reset = false;
            if (!pending_reset) {  // tick t is complete: its visibility map is final
                if (vis_traj) pack_vis<BIG>(D, S, geo, lane, vis_traj + o * D.RW, dirty);
                ++t;
            }
        }
    }
    if (have) {
        store_env<BIG>(D, S, geo, env, lane, E, status, n_cams, n_guards, dirty);
        // feed the measured work (resets included) back into the slot order of the next launch
        if (lane == 0 && T >= 8) D.cost[env] = (int)(((long long)work * 256) / T) + 16;
    }
}

template <bool EXACT_ONLY, bool BIG>
__global__ void __launch_bounds__(HEIST_WARPS_PER_CTA * 32, 8)
k_reset(Dev D, const uint8_t *__restrict__ mask) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    CtaCtl *ctl = reinterpret_cast<CtaCtl *>(smem);
    const Geo geo = make_geo(D.R, D.C);
    dbg_init(D, smem, geo);
    for (int i = threadIdx.x; i < PEND_CAP; i += HEIST_WARPS_PER_CTA * 32) ctl->pend[i] = PEND_EMPTY;
    const int ctx_bytes = (int)warp_ctx_bytes(D.R, D.C, D.Kc, D.Kg);
    if (D.skip_cached && *D.n_uncached == 0) return;
    const int env = D.slot2env[blockIdx.x * HEIST_WARPS_PER_CTA + warp];
    const bool have = env >= 0 && (!mask || mask[env]) && !(D.skip_cached && D.env_cached[env]);
    const WarpCtx S = carve_warp_ctx(smem, geo, warp, D.R, D.C, D.Kc, D.Kg);
    EnvRegs E;
    int n_cams = 0, n_guards = 0;
    unsigned long long dirty = ~0ull;
    if (have) {
        load_env(D, S, geo, env, lane, E, n_cams, n_guards);
        reset_state(D, S, env, lane, E, n_cams, n_guards);
        begin_visibility<BIG>(S, geo, D.R, ctl, warp, lane, n_cams + n_guards, dirty);
    } else if (lane == 0) ctl->cnt[warp] = 0;
    if (lane == 0) ctl->n_assets[warp] = n_cams + n_guards;
    if (threadIdx.x == 0) { ctl->pend_n = 0; ctl->taken = 0; }
    __syncthreads();
    march_phase<EXACT_ONLY>(D, S, geo, warp, ctx_bytes, ctl, lane);
    if (have) {
        end_visibility(S, lane, n_cams, n_cams + n_guards);
        store_env<BIG>(D, S, geo, env, lane, E, HEIST_RUNNING, n_cams, n_guards, dirty);
    }
}

// ---------------------------------------------------------------------------------------------
// Cost-balanced warp-slot order.  Envs differ several-fold in ray-march work, so slots are dealt in
// "snake" order over the envs sorted by cost: every CTA (4 envs) gets one env from each quartile and
// all CTAs carry about the same work.
// ---------------------------------------------------------------------------------------------
#define ORDER_BINS 1024
__global__ void __launch_bounds__(1024) k_build_order(Dev D, int n_ctas) {
    __shared__ int hist[ORDER_BINS];
    __shared__ int red[32];
    __shared__ int s_max;
    const int tid = threadIdx.x;
    int m = 0;
    for (int i = tid; i < D.N; i += 1024) m = max(m, D.cost[i]);
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((tid & 31) == 0) red[tid >> 5] = m;
    hist[tid] = 0;
    __syncthreads();
    if (tid < 32) {
        m = red[tid];
        for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (tid == 0) s_max = max(m, 1);
    }
    __syncthreads();
    const float inv = (float)(ORDER_BINS - 1) / (float)s_max;
    for (int i = tid; i < D.N; i += 1024) atomicAdd(&hist[min(ORDER_BINS - 1, (int)((float)D.cost[i] * inv))], 1);
    __syncthreads();
    // start[b] = number of envs in bins above b (descending cost)
    if (tid == 0) {
        int acc = 0;
        for (int b = ORDER_BINS - 1; b >= 0; --b) { int c = hist[b]; hist[b] = acc; acc += c; }
    }
    __syncthreads();
    for (int i = tid; i < D.N; i += 1024) {
        int b = min(ORDER_BINS - 1, (int)((float)D.cost[i] * inv));
        int rank = atomicAdd(&hist[b], 1);
        int rnd = rank / n_ctas, pos = rank - rnd * n_ctas;
        if (rnd & 1) pos = n_ctas - 1 - pos;
        D.slot2env[pos * HEIST_WARPS_PER_CTA + rnd] = i;
    }
}
