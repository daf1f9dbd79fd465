// heist_layout.cuh -- Architect-side decode, set_layout placement, BFS validity (one warp per env).
//
// Reference: ArchitectNetwork.generate_layout decode loop + _generate_patrol (networks.py:273-335),
// curriculum filter (training.py:464-467), HeistEnvironment.set_layout / _reset_layout /
// _is_valid_placement (environment.py:102-177), BudgetManager.purchase (budget.py:48-58),
// bfs_path_exists (utils.py:52-85).
#pragma once
#include "heist_common.cuh"

#define COST_WALL 1
#define COST_CAMERA 3
#define COST_GUARD 5

// k_set_layout shared memory per warp: tile codes [RC] (16-byte aligned) + wall row bitmaps [RW]
__host__ __device__ inline size_t layout_warp_bytes(int RC, int RW) {
    return ((((size_t)RC + 15) & ~(size_t)15) + (size_t)RW * 4 + 15) & ~(size_t)15;
}

// Mutable device copy of HeistLayoutArrays (decode output == set_layout input).
struct LayoutDev {
    int32_t *n_walls;  int16_t *wall_rc;
    int32_t *n_cams;   int16_t *cam_rc;  double *cam_f;  int32_t *cam_range;
    int32_t *n_guards; int32_t *guard_len; int16_t *guard_path; double *guard_head;
    int32_t *guard_speed; int32_t *guard_range; double *guard_fov;
};

// ---------------------------------------------------------------------------------------------
// Decode: row-major greedy budget scan of the sampled asset map (networks.py:283-318).
// All lanes run the scalar scan redundantly on ballot-compacted non-zero cells; lane 0 writes.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(HEIST_WARPS_PER_CTA * 32)
k_decode(Dev D, LayoutDev Lz, const int8_t *__restrict__ asset_map, const float *__restrict__ cam_params,
         const int32_t *__restrict__ budget, int allow_cameras, int allow_guards) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int env = blockIdx.x * HEIST_WARPS_PER_CTA + warp;
    if (env >= D.N) return;
    const int H = D.R, Wd = D.C;
    int remaining = budget ? budget[env] : D.budget;
    int nw = 0, ncam = 0, ng = 0;
    bool overflow = false;
    // .item() of fp32 tensors -> Python float: exact widening (networks.py:294-296)
    const double fov = (double)cam_params[(size_t)env * 3 + 0];
    const double speed = (double)cam_params[(size_t)env * 3 + 1];
    const double heading = (double)cam_params[(size_t)env * 3 + 2];
    const int8_t *am = asset_map + (size_t)env * D.RC;
    bool stop = false;
    for (int r = 1; r < H - 1 && !stop; ++r) {
        for (int half = 0; half * 32 < Wd && !stop; ++half) {
            int c0 = half * 32 + lane;
            int v = (c0 >= 1 && c0 < Wd - 1) ? (int)am[r * Wd + c0] : 0;
            unsigned nz = __ballot_sync(0xffffffffu, v != 0);
            while (nz) {
                int src = __ffs(nz) - 1;
                nz &= nz - 1;
                int t = __shfl_sync(0xffffffffu, v, src);
                int c = half * 32 + src;
                if (t == 1 && remaining >= COST_WALL) {
                    if (nw < D.Kw) { if (lane == 0) { int16_t *w = Lz.wall_rc + ((size_t)env * D.Kw + nw) * 2; w[0] = r; w[1] = c; } nw++; }
                    else overflow = true;
                    remaining -= COST_WALL;
                } else if (t == 2 && remaining >= COST_CAMERA) {
                    if (ncam < D.Kc) {
                        if (lane == 0) {
                            size_t o = (size_t)env * D.Kc + ncam;
                            Lz.cam_rc[o * 2] = r; Lz.cam_rc[o * 2 + 1] = c;
                            Lz.cam_f[o * 3] = fov; Lz.cam_f[o * 3 + 1] = heading; Lz.cam_f[o * 3 + 2] = speed;
                            Lz.cam_range[o] = 6;
                        }
                        ncam++;
                    } else overflow = true;
                    remaining -= COST_CAMERA;
                } else if (t == 3 && remaining >= COST_GUARD) {
                    if (ng < D.Kg && D.L >= 8) {
                        // _generate_patrol (networks.py:324-335): clamped 3x3 ring, 8 waypoints
                        if (lane < 8) {
                            const int offr[8] = {0, 0, 0, 1, 2, 2, 2, 1}, offc[8] = {0, 1, 2, 2, 2, 1, 0, 0};
                            int pr = max(1, min(H - 2, r + offr[lane] - 1));
                            int pc = max(1, min(Wd - 2, c + offc[lane] - 1));
                            int nx = (lane + 1) & 7;
                            int qr = max(1, min(H - 2, r + offr[nx] - 1));
                            int qc = max(1, min(Wd - 2, c + offc[nx] - 1));
                            int dr = qr - pr, dc = qc - pc;  // unit axis move or (0,0) after clamping
                            // degrees(atan2(-dr, dc)) % 360.0 for unit moves (security.py:159)
                            double h = (dr == 0 && dc == 0) ? __longlong_as_double(0x7ff8000000000000LL)
                                       : (dc == 1 ? 0.0 : (dr == -1 ? 90.0 : (dc == -1 ? 180.0 : 270.0)));
                            size_t o = ((size_t)env * D.Kg + ng) * D.L + lane;
                            Lz.guard_path[o * 2] = pr; Lz.guard_path[o * 2 + 1] = pc;
                            Lz.guard_head[o] = h;
                        }
                        if (lane == 0) {
                            size_t o = (size_t)env * D.Kg + ng;
                            Lz.guard_len[o] = 8; Lz.guard_speed[o] = 1; Lz.guard_range[o] = 4; Lz.guard_fov[o] = 90.0;
                        }
                        ng++;
                    } else overflow = true;
                    remaining -= COST_GUARD;
                }
                if (remaining <= 0) { stop = true; break; }
            }
        }
    }
    if (lane == 0) {
        Lz.n_walls[env] = nw;
        Lz.n_cams[env] = allow_cameras ? ncam : 0;   // training.py:464-467
        Lz.n_guards[env] = allow_guards ? ng : 0;
        if (overflow) atomicOr(D.err, ERR_CAPACITY);
    }
}

// ---------------------------------------------------------------------------------------------
// Bit-parallel flood fill, lane = grid row (rows lane and lane+32), 64-bit row words.
// ---------------------------------------------------------------------------------------------
// All cells reachable from the seeds s (s within pass) along runs of passable bits of one row.  Adding the
// seeds to the run mask ripples a carry from the lowest seed of a run to the run's top bit: the bits that
// flip are that stretch (plus the wall bit above, masked off); the other seeds of the run are OR-ed back.
// The downward direction is the same on the bit-reversed row.
__device__ __forceinline__ unsigned long long hfill(unsigned long long s, unsigned long long pass, unsigned long long rpass) {
    s &= pass;
    const unsigned long long up = (((pass + s) ^ pass) & pass) | s;
    const unsigned long long rs = __brevll(s);
    const unsigned long long dn = (((rpass + rs) ^ rpass) & rpass) | rs;
    return up | __brevll(dn);
}
__device__ __forceinline__ bool bfs_warp(const uint32_t *wallrows, int R, int C, int W, int lane, int sr, int sc,
                                         int gr, int gc) {
    if (sr == gr && sc == gc) return true;  // utils.py:65-66
    const unsigned long long colmask = (C >= 64) ? ~0ull : ((1ull << C) - 1ull);
    unsigned long long pass0 = 0, pass1 = 0;
    if (lane < R) {
        unsigned long long w = wallrows[lane * W];
        if (W > 1) w |= (unsigned long long)wallrows[lane * W + 1] << 32;
        pass0 = ~w & colmask;
    }
    if (lane + 32 < R) {
        unsigned long long w = wallrows[(lane + 32) * W];
        if (W > 1) w |= (unsigned long long)wallrows[(lane + 32) * W + 1] << 32;
        pass1 = ~w & colmask;
    }
    unsigned long long reach0 = (lane == sr) ? (1ull << sc) : 0ull;
    unsigned long long reach1 = (lane + 32 == sr) ? (1ull << sc) : 0ull;
    const unsigned long long gbit = 1ull << gc;
    const unsigned long long rpass0 = __brevll(pass0), rpass1 = __brevll(pass1);
    for (;;) {
        // complete flood along each row in one shot (carry trick), then a few cheap vertical exchanges
        unsigned long long n0 = hfill(reach0, pass0, rpass0), n1 = hfill(reach1, pass1, rpass1);
#pragma unroll 1
        for (int k = 0; k < 8; ++k) {
            unsigned long long up0 = __shfl_up_sync(0xffffffffu, n0, 1);
            unsigned long long dn0 = __shfl_down_sync(0xffffffffu, n0, 1);
            unsigned long long up1 = __shfl_up_sync(0xffffffffu, n1, 1);
            unsigned long long dn1 = __shfl_down_sync(0xffffffffu, n1, 1);
            unsigned long long n0_last = __shfl_sync(0xffffffffu, n0, 31);
            unsigned long long n1_first = __shfl_sync(0xffffffffu, n1, 0);
            if (lane == 0) { up0 = 0ull; up1 = n0_last; }
            if (lane == 31) { dn0 = n1_first; dn1 = 0ull; }
            n0 |= (up0 | dn0) & pass0;
            n1 |= (up1 | dn1) & pass1;
        }
        bool changed = (n0 != reach0) || (n1 != reach1);
        reach0 = n0; reach1 = n1;
        bool hit = (lane == gr && (reach0 & gbit)) || (lane + 32 == gr && (reach1 & gbit));
        if (__any_sync(0xffffffffu, hit)) return true;
        if (!__any_sync(0xffffffffu, changed)) return false;
    }
}

// ---------------------------------------------------------------------------------------------
// set_layout (environment.py:102-152): fresh grid, walls -> cameras -> guards with the env budget,
// wall bitmaps, BFS validity.  Lane 0 runs the (short, order-dependent) placement loops on the
// shared-memory tile grid; the rest is warp-parallel.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(HEIST_WARPS_PER_CTA * 32)
k_set_layout(Dev D, LayoutDev Lz, const int32_t *__restrict__ budget, uint8_t *__restrict__ valid_out) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int env = blockIdx.x * HEIST_WARPS_PER_CTA + warp;
    if (env >= D.N) return;
    const int R = D.R, C = D.C, W = D.W;
    const size_t per_warp = layout_warp_bytes(D.RC, D.RW);
    uint8_t *tile = smem + (size_t)warp * per_warp;
    uint32_t *wallrows = (uint32_t *)(tile + (((size_t)D.RC + 15) & ~(size_t)15));

    // _reset_layout (:169-177) / create_empty_grid (utils.py:131-139)
    if ((C & 3) == 0) {  // 4 cells per store
        const int qpr = C >> 2;
        for (int i = lane; i < (D.RC >> 2); i += 32) {
            const int r = i / qpr, q = i - r * qpr;
            unsigned v = (r == 0 || r == R - 1) ? 0x01010101u : 0u;
            if (q == 0) v |= 0x00000001u;
            if (q == qpr - 1) v |= 0x01000000u;
            reinterpret_cast<unsigned *>(tile)[i] = v;
        }
    } else {
        for (int i = lane; i < D.RC; i += 32) {
            int r = i / C, c = i - r * C;
            tile[i] = (r == 0 || r == R - 1 || c == 0 || c == C - 1) ? HEIST_WALL : HEIST_EMPTY;
        }
    }
    __syncwarp();
    int n_cams = 0, n_guards = 0, spent = 0, cost = 0;
    if (lane == 0) {
        tile[D.start_r * C + D.start_c] = HEIST_START;
        tile[D.vault_r * C + D.vault_c] = HEIST_VAULT;
        const int total = budget ? budget[env] : D.budget;
        int err = 0;
        // walls (:118-121)
        int nw = Lz.n_walls ? Lz.n_walls[env] : 0;
        if (nw > D.Kw) { nw = D.Kw; err |= ERR_CAPACITY; }
        for (int i = 0; i < nw; ++i) {
            const int16_t *w = Lz.wall_rc + ((size_t)env * D.Kw + i) * 2;
            int r = w[0], c = w[1];
            const bool ok = r > 0 && r < R - 1 && c > 0 && c < C - 1 && tile[r * C + c] == HEIST_EMPTY && total - spent >= COST_WALL;
            if (ok) {
                spent += COST_WALL;
                tile[r * C + c] = HEIST_WALL;
            }
            D.wall_ok[(size_t)env * D.Kw + i] = ok ? 1 : 0;
        }
        // cameras (:124-135)
        int nc = Lz.n_cams ? Lz.n_cams[env] : 0;
        if (nc > D.Kc) { nc = D.Kc; err |= ERR_CAPACITY; }
        for (int i = 0; i < nc; ++i) {
            size_t o = (size_t)env * D.Kc + i;
            int r = Lz.cam_rc[o * 2], c = Lz.cam_rc[o * 2 + 1];
            if (r > 0 && r < R - 1 && c > 0 && c < C - 1 && tile[r * C + c] == HEIST_EMPTY && total - spent >= COST_CAMERA) {
                spent += COST_CAMERA;
                tile[r * C + c] = HEIST_CAMERA;
                double fov = Lz.cam_f[o * 3], heading = Lz.cam_f[o * 3 + 1], speed = Lz.cam_f[o * 3 + 2];
                // num_rays = max(int(fov * 2), 30)  (security.py:67)
                double two = __dmul_rn(fov, 2.0);
                int num_rays = (two >= 32767.0) ? 32767 : __double2int_rz(two);
                if (two >= 32767.0) err |= ERR_RAYS;
                if (num_rays < 30) num_rays = 30;
                int rng = Lz.cam_range[o];
                if (rng > 32767) { rng = 32767; err |= ERR_RAYS; }
                size_t d = (size_t)env * D.Kc + n_cams;
                D.cam_f[d * 2] = fov; D.cam_f[d * 2 + 1] = speed;
                D.cam_heading[d] = heading;
                int16_t *ci = D.cam_i + d * 4;
                ci[0] = r; ci[1] = c; ci[2] = rng; ci[3] = num_rays;
                cost += (num_rays + 1) * (2 * max(rng, 0) + 4);
                n_cams++;
            }
        }
        // guards (:138-149): no placement check; the start waypoint tile is overwritten
        int ng = Lz.n_guards ? Lz.n_guards[env] : 0;
        if (ng > D.Kg) { ng = D.Kg; err |= ERR_CAPACITY; }
        for (int i = 0; i < ng; ++i) {
            size_t o = (size_t)env * D.Kg + i;
            int len = Lz.guard_len[o];
            if (len > D.L) { len = D.L; err |= ERR_CAPACITY; }
            if (len > 0 && total - spent >= COST_GUARD) {
                bool ok = true;
                for (int k = 0; k < len; ++k) {
                    int pr = Lz.guard_path[(o * D.L + k) * 2], pc = Lz.guard_path[(o * D.L + k) * 2 + 1];
                    if (pr < 0 || pr >= R || pc < 0 || pc >= C) ok = false;
                }
                if (!ok) { err |= ERR_WAYPOINT; continue; }
                spent += COST_GUARD;
                size_t d = (size_t)env * D.Kg + n_guards;
                double fov = Lz.guard_fov[o];
                double two = __dmul_rn(fov, 2.0);
                int num_rays = (two >= 32767.0) ? 32767 : __double2int_rz(two);
                if (two >= 32767.0) err |= ERR_RAYS;
                if (num_rays < 30) num_rays = 30;
                D.guard_fov[d] = fov;
                int32_t *gi = D.guard_i + d * 4;
                gi[0] = len; gi[1] = Lz.guard_speed[o]; gi[2] = Lz.guard_range[o]; gi[3] = num_rays;
                for (int k = 0; k < len; ++k) {
                    D.guard_path[(d * D.L + k) * 2] = (uint8_t)Lz.guard_path[(o * D.L + k) * 2];
                    D.guard_path[(d * D.L + k) * 2 + 1] = (uint8_t)Lz.guard_path[(o * D.L + k) * 2 + 1];
                    D.guard_head[d * D.L + k] = Lz.guard_head[o * D.L + k];
                }
                cost += (num_rays + 1) * (max(Lz.guard_range[o], 0) + 4);
                D.guard_heading[d] = 0.0;  // Guard.heading default (security.py:131)
                D.guard_idx[d] = 0;
                tile[Lz.guard_path[o * D.L * 2] * C + Lz.guard_path[o * D.L * 2 + 1]] = HEIST_GUARD;  // :148
                n_guards++;
            }
        }
        if (err) atomicOr(D.err, err);
    }
    __syncwarp();
    // wall_mask = (grid == WALL) as row bitmaps (environment.py:211,257)
    if ((C & 31) == 0) {  // lane = row: per-byte compare (__vcmpeq4) + multiply-gather, 4 tiles per step
        for (int r = lane; r < R; r += 32)
            for (int w = 0; w < W; ++w) {
                const unsigned *src = reinterpret_cast<const unsigned *>(tile + r * C + w * 32);
                unsigned m = 0;
#pragma unroll
                for (int n = 0; n < 8; ++n)
                    m |= (((__vcmpeq4(src[n], 0x01010101u) & 0x01010101u) * 0x01020408u) >> 24) << (4 * n);
                wallrows[r * W + w] = m;
            }
    } else {
        for (int r = 0; r < R; ++r)
            for (int w = 0; w < W; ++w) {
                int c = w * 32 + lane;
                unsigned m = __ballot_sync(0xffffffffu, c < C && tile[r * C + c] == HEIST_WALL);
                if (lane == 0) wallrows[r * W + w] = m;
            }
    }
    __syncwarp();
    bool valid = bfs_warp(wallrows, R, C, W, lane, D.start_r, D.start_c, D.vault_r, D.vault_c);
    // stores
    if ((D.RC & 15) == 0) {  // 16-byte stores (the per-env tile block is then 16-byte aligned too)
        int4 *dst = reinterpret_cast<int4 *>(D.tile + (size_t)env * D.RC);
        for (int i = lane; i < (D.RC >> 4); i += 32) dst[i] = reinterpret_cast<const int4 *>(tile)[i];
    } else {
        for (int i = lane; i < D.RC; i += 32) D.tile[(size_t)env * D.RC + i] = tile[i];
    }
    for (int i = lane; i < D.RW; i += 32) D.wall[(size_t)env * D.RW + i] = wallrows[i];
    if (lane == 0) {
        *reinterpret_cast<int4 *>(D.env_s + (size_t)env * 4) = make_int4(n_cams, n_guards, valid ? 1 : 0, spent);
        // outcome counters restart with a new layout; solver state is untouched (set_layout does not reset it)
        int32_t *d = D.env_d + (size_t)env * 8;
        d[5] = 0; d[6] = 0; d[7] = 0;
        if (valid_out) valid_out[env] = valid ? 1 : 0;
        D.cost[env] = cost + 64;
    }
}

// HeistEnvironment.__init__ solver state (environment.py:83-92)
__global__ void k_init_dyn(Dev D) {
    int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= D.N) return;
    int dist = abs(D.start_r - D.vault_r) + abs(D.start_c - D.vault_c);
    int32_t *d = D.env_d + (size_t)env * 8;
    d[0] = D.start_r | (D.start_c << 16); d[1] = 0; d[2] = dist; d[3] = dist;
    d[4] = 0; d[5] = 0; d[6] = 0; d[7] = 0;
}
