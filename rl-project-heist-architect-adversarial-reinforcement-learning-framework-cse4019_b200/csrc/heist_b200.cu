// heist_b200.cu -- C ABI (include/heist_b200.h) over the sm_100a kernels.
//
// Build (see build.py): nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false
//                       -shared -Xcompiler -fPIC -o libheist_b200.so heist_b200.cu
#include "../../include/heist_b200.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <utility>
#include <vector>

#include "heist_common.cuh"
#include "heist_layout.cuh"
#include "heist_step.cuh"
#include "heist_cache.cuh"
#include "heist_fast.cuh"
#include "heist_walk.cuh"
#include "heist_stream.cuh"

static thread_local std::string g_err;
static thread_local std::string g_warn;   // heist_last_warning: conditions that are not errors but cost performance

static int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                        \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) return fail((int)_e, "%s: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)

extern "C" int heist_destroy(HeistHandle *h);
// inside heist_create, once the handle exists: free it before reporting the failure
#define CUDA_TRY_H(expr)                                                                          \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) { heist_destroy(h); return fail((int)_e, "heist_create: %s: %s", #expr, cudaGetErrorString(_e)); } \
    } while (0)

struct HeistHandle {
    HeistParams p;
    int N, device;
    Dev d;
    LayoutDev lz;  // decode output buffers
    size_t step_smem, layout_smem;
    int mode;        // heist_set_mode: HEIST_MODE_*
    size_t camvis_smem, camvis_staged_smem, cache_bytes;
    long long launches;   // kernels launched by reset / step / step_many (heist_launch_count)
    int *n_unc_host;      // pinned copy of d.n_uncached, refreshed (async) after every cache build
    cudaEvent_t ev_unc;   // ... complete when this event is
    int all_cached;       // -1 unknown, 0 some envs need the ray-march kernels, 1 none does (their launch is skipped)
    double *heads;      size_t heads_cap;     // k_heads output, grow-only [tick blocks][N][Kc]
    double *h_run;                            // [N][Kc] running headings between the chunks of a pipelined launch
    cudaStream_t s_heads;  cudaEvent_t ev_heads[64];
    uint32_t *scratch;  size_t scratch_cap;   // cam_vis when the caller wants no visibility trajectory, grow-only
    uint16_t *grec;     size_t grec_cap;      // k_seq -> k_finish: guard (waypoint, heading slot) per tick [T][N][Kg]
    uint8_t *fin;       size_t fin_cap;       // k_seq -> k_finish: tick rebuilt its map [T][N]
    int32_t *last_t;    size_t last_cap;      // [chunks][N] last rebuilt tick of each chunk
    size_t seq_smem;
    cudaStream_t s_seq, s_fin, s_cam2;        // side streams of the pipelined launch
    cudaEvent_t ev_fork, ev_join, ev_join2, ev_cam[64], ev_seq[64];
    // heist_step_many_host: device staging of the host buffers, copy streams of the pipelined launch
    int8_t *st_act;  float *st_rew;  uint8_t *st_done, *st_status;  size_t st_cap;
    cudaStream_t s_h2d, s_d2h;
    cudaEvent_t ev_h2d[64], ev_join3;
    void *allocs[96];
    int n_allocs;
};

template <typename T>
static cudaError_t dalloc(HeistHandle *h, T **ptr, size_t count) {
    void *p = nullptr;
    size_t bytes = count * sizeof(T);
    if (bytes == 0) bytes = sizeof(T);
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return e;
    e = cudaMemset(p, 0, bytes);
    if (e != cudaSuccess) return e;
    h->allocs[h->n_allocs++] = p;
    *ptr = (T *)p;
    return cudaSuccess;
}

static cudaError_t build_cache(HeistHandle *h, cudaStream_t s);
static inline int env_blocks(int N) { return (N + HEIST_WARPS_PER_CTA - 1) / HEIST_WARPS_PER_CTA; }

extern "C" int heist_abi_version(void) { return HEIST_ABI_VERSION; }
extern "C" const char *heist_last_error(void) { return g_err.c_str(); }
extern "C" const char *heist_last_warning(void) { return g_warn.c_str(); }

extern "C" int heist_destroy(HeistHandle *h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < h->n_allocs; ++i) cudaFree(h->allocs[i]);
    if (h->heads) cudaFree(h->heads);
    if (h->scratch) cudaFree(h->scratch);
    if (h->h_run) cudaFree(h->h_run);
    if (h->grec) cudaFree(h->grec);
    if (h->fin) cudaFree(h->fin);
    if (h->last_t) cudaFree(h->last_t);
    if (h->n_unc_host) cudaFreeHost(h->n_unc_host);
    if (h->st_act) { cudaFree(h->st_act); cudaFree(h->st_rew); cudaFree(h->st_done); cudaFree(h->st_status); }
    if (h->ev_unc) cudaEventDestroy(h->ev_unc);
    // (a heist_create that failed half-way leaves some of these null)
    cudaStream_t streams[] = {h->s_seq, h->s_fin, h->s_cam2, h->s_h2d, h->s_d2h, h->s_heads};
    for (cudaStream_t st : streams) if (st) cudaStreamDestroy(st);
    cudaEvent_t events[] = {h->ev_join3, h->ev_fork, h->ev_join, h->ev_join2};
    for (cudaEvent_t ev : events) if (ev) cudaEventDestroy(ev);
    for (int i = 0; i < 64; ++i) {
        if (h->ev_h2d[i]) cudaEventDestroy(h->ev_h2d[i]);
        if (h->ev_heads[i]) cudaEventDestroy(h->ev_heads[i]);
        if (h->ev_cam[i]) cudaEventDestroy(h->ev_cam[i]);
        if (h->ev_seq[i]) cudaEventDestroy(h->ev_seq[i]);
    }
    cudaGetLastError();
    delete h;
    return 0;
}

// cos / -sin of every double within NICE_W degrees of a multiple of 30 degrees, from the HOST libm (heist_common.cuh).
// One table per device for the life of the process (handles share it; ~0.7 MB).
static cudaError_t upload_nice_table(int device, double deg2rad) {
    static double2 *tab_dev[64] = {nullptr};
    static long long lo[NICE_N];
    static int cnt[NICE_N], off[NICE_N];
    if (device < 0 || device >= 64) return cudaErrorInvalidDevice;
    if (!tab_dev[device]) {
        int total = 0;
        for (int i = 0; i < NICE_N; ++i) {
            const int k = i - NICE_K;
            lo[i] = 0; cnt[i] = 0; off[i] = total;
            if (k == 0) continue;
            const double c = fabs((double)k * 30.0);
            long long b0, b1;
            const double a0 = c - NICE_W, a1 = c + NICE_W;
            memcpy(&b0, &a0, 8); memcpy(&b1, &a1, 8);
            lo[i] = b0; cnt[i] = (int)(b1 - b0 + 1);
            total += cnt[i];
        }
        double2 *host = (double2 *)malloc(sizeof(double2) * (size_t)total);
        if (!host) return cudaErrorMemoryAllocation;
        for (int i = 0; i < NICE_N; ++i) {
            const int k = i - NICE_K;
            for (int j = 0; j < cnt[i]; ++j) {
                const long long b = lo[i] + j;
                double mag;
                memcpy(&mag, &b, 8);
                volatile double a = k < 0 ? -mag : mag;
                volatile double rad = a * deg2rad;   // math.radians
                host[off[i] + j].x = cos(rad);
                host[off[i] + j].y = -sin(rad);
            }
        }
        double2 *dev = nullptr;
        cudaError_t e = cudaMalloc(&dev, sizeof(double2) * (size_t)total);
        if (e == cudaSuccess) e = cudaMemcpy(dev, host, sizeof(double2) * (size_t)total, cudaMemcpyHostToDevice);
        free(host);
        if (e != cudaSuccess) { if (dev) cudaFree(dev); return e; }
        tab_dev[device] = dev;
    }
    cudaError_t e = cudaMemcpyToSymbol(c_nice_lo, lo, sizeof(lo));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_nice_cnt, cnt, sizeof(cnt));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_nice_off, off, sizeof(off));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_nice_tab, &tab_dev[device], sizeof(double2 *));
    return e;
}

// Tie bands per vision range for k_build_cache (heist_cache.cuh): for every sample distance d = 0.5 j and every
// rounding tie m + 0.5 with |tie| <= d, the angles at which col + cos(a) d (or row - sin(a) d) is within 2 mu of the
// tie -- unpadded (start, end) pairs over ray angles [-93, 453], sorted by start.  One set per device for the life of
// the process.
static cudaError_t vc_upload_band_tables(int device) {
    static double2 *tab_dev[64][VC_MAX_RANGE + 1];
    static int tab_n[64][VC_MAX_RANGE + 1];
    if (device < 0 || device >= 64) return cudaErrorInvalidDevice;
    if (!tab_dev[device][1]) {
        const double rad2deg = 180.0 / 3.14159265358979323846;
        for (int range = 1; range <= VC_MAX_RANGE; ++range) {
            std::vector<std::pair<double, double>> v;
            for (int j = 1; j <= 2 * range; ++j) {
                const double d = 0.5 * (double)j;
                for (int mi = -8; mi < 8; ++mi) {
                    const double tie = (double)mi + 0.5;
                    if (fabs(tie) > d + VC_MU2) continue;
                    const double t = tie / d, mu = VC_MU2 / d;
                    for (int axis = 0; axis < 2; ++axis) {
                        double c_lo = std::max(-1.0, t - mu), c_hi = std::min(1.0, t + mu);
                        if (axis) { const double a = -c_hi; c_hi = -c_lo; c_lo = a; }   // dy = -sin(a) = -cos(a - 90)
                        // cos(a - off) in [c_lo, c_hi]  <=>  a - off in +-[acos(c_hi), acos(c_lo)] + 360 n
                        const double a_lo = acos(c_hi) * rad2deg, a_hi = acos(c_lo) * rad2deg, off = axis ? 90.0 : 0.0;
                        for (int sgn = 0; sgn < 2; ++sgn) {
                            const double b0 = sgn ? off - a_hi : off + a_lo, b1 = sgn ? off - a_lo : off + a_hi;
                            for (int n = -1; n <= 2; ++n) {
                                const double s0 = b0 + 360.0 * n, e0 = b1 + 360.0 * n;
                                if (e0 < -93.0 || s0 > 453.0) continue;
                                v.push_back(std::make_pair(s0, e0));
                            }
                        }
                    }
                }
            }
            std::sort(v.begin(), v.end());
            std::vector<double2> h2(v.size());
            for (size_t i = 0; i < v.size(); ++i) { h2[i].x = v[i].first; h2[i].y = v[i].second; }
            double2 *dev = nullptr;
            cudaError_t e = cudaMalloc(&dev, sizeof(double2) * h2.size());
            if (e == cudaSuccess) e = cudaMemcpy(dev, h2.data(), sizeof(double2) * h2.size(), cudaMemcpyHostToDevice);
            if (e != cudaSuccess) { if (dev) cudaFree(dev); return e; }
            tab_dev[device][range] = dev;
            tab_n[device][range] = (int)h2.size();
        }
    }
    cudaError_t e = cudaMemcpyToSymbol(c_vcb_tab, tab_dev[device], sizeof(tab_dev[device]));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_vcb_n, tab_n[device], sizeof(tab_n[device]));
    return e;
}

extern "C" int heist_create(const HeistParams *params, int num_envs, int device, HeistHandle **out) {
    if (!params || !out) return fail(-1, "heist_create: null argument");
    const HeistParams &p = *params;
    if (num_envs <= 0) return fail(-2, "heist_create: num_envs must be positive");
    if (p.grid_rows < 3 || p.grid_cols < 3 || p.grid_rows > HEIST_MAX_DIM || p.grid_cols > HEIST_MAX_DIM)
        return fail(-3, "heist_create: grid %dx%d outside 3..%d", p.grid_rows, p.grid_cols, HEIST_MAX_DIM);
    if (p.start_row < 0 || p.start_row >= p.grid_rows || p.start_col < 0 || p.start_col >= p.grid_cols ||
        p.vault_row < 0 || p.vault_row >= p.grid_rows || p.vault_col < 0 || p.vault_col >= p.grid_cols)
        return fail(-4, "heist_create: start/vault outside the grid");
    if (p.max_walls < 0 || p.max_cams < 0 || p.max_guards < 0 || p.max_cams + p.max_guards > 32 || p.max_path < 1 ||
        p.max_path > 255)
        return fail(-5, "heist_create: capacities out of range (max_cams + max_guards <= 32; 1 <= max_path <= 255)");
    static_assert(HEIST_WARPS_PER_CTA == 4, "pend[] packs the env slot in 2 bits");
    if (p.max_steps < 1) return fail(-6, "heist_create: max_steps must be >= 1");
    CUDA_TRY(cudaSetDevice(device));

    HeistHandle *h = new HeistHandle();
    memset(h, 0, sizeof(*h));
    h->p = p; h->N = num_envs; h->device = device;
    Dev &d = h->d;
    d.N = num_envs; d.R = p.grid_rows; d.C = p.grid_cols; d.W = (p.grid_cols + 31) / 32;
    d.RW = d.R * d.W; d.RC = d.R * d.C;
    d.max_steps = p.max_steps; d.start_r = p.start_row; d.start_c = p.start_col;
    d.vault_r = p.vault_row; d.vault_c = p.vault_col; d.budget = p.architect_budget;
    d.Kw = p.max_walls > 0 ? p.max_walls : 1; d.Kc = p.max_cams > 0 ? p.max_cams : 1;
    d.Kg = p.max_guards > 0 ? p.max_guards : 1; d.L = p.max_path;
    d.reward_vault = p.reward_vault; d.reward_detection = p.reward_detection; d.reward_step = p.reward_step;
    d.deg2rad = 3.14159265358979323846 / 180.0;  // Py_MATH_PI / 180.0 (mathmodule.c degToRad)
    const size_t N = num_envs;
    cudaError_t e = cudaSuccess;
#define A(ptr, count) if (e == cudaSuccess) e = dalloc(h, &(ptr), (count))
    A(d.tile, N * d.RC); A(d.wall, N * d.RW); A(d.env_s, N * 4); A(d.wall_ok, N * d.Kw);
    A(d.cam_f, N * d.Kc * 2); A(d.cam_i, N * d.Kc * 4);
    A(d.guard_fov, N * d.Kg); A(d.guard_i, N * d.Kg * 4); A(d.guard_path, N * d.Kg * d.L * 2);
    A(d.guard_head, N * d.Kg * d.L);
    A(d.env_d, N * 8); A(d.cam_heading, N * d.Kc); A(d.guard_heading, N * d.Kg); A(d.guard_idx, N * d.Kg);
    A(d.vis, N * d.RW); A(d.pos_tab, (size_t)d.RC); A(d.err, (size_t)1);
    A(d.cost, N); A(d.slot2env, (size_t)env_blocks(num_envs) * HEIST_WARPS_PER_CTA);
    LayoutDev &z = h->lz;
    A(z.n_walls, N); A(z.wall_rc, N * d.Kw * 2); A(z.n_cams, N); A(z.cam_rc, N * d.Kc * 2);
    A(z.cam_f, N * d.Kc * 3); A(z.cam_range, N * d.Kc); A(z.n_guards, N); A(z.guard_len, N * d.Kg);
    A(z.guard_path, N * d.Kg * d.L * 2); A(z.guard_head, N * d.Kg * d.L); A(z.guard_speed, N * d.Kg);
    A(z.guard_range, N * d.Kg); A(z.guard_fov, N * d.Kg);
    A(d.env_cached, N); A(d.n_uncached, (size_t)1);
    if (e == cudaSuccess) e = cudaHostAlloc(&h->n_unc_host, sizeof(int), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_unc, cudaEventDisableTiming);
    h->all_cached = -1;
    if (e != cudaSuccess) { heist_destroy(h); return fail((int)e, "heist_create: cudaMalloc: %s", cudaGetErrorString(e)); }
    // Visibility cache (heist_cache.cuh): ~25 KB per camera slot.  Optional: when it does not fit, every env
    // stays on the ray-march kernel.
    {
        const char *off = getenv("HEIST_NO_VIS_CACHE");
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        const size_t HS = (size_t)d.L + 1;
        const size_t need = N * d.Kc * ((size_t)VC_POINTS * 4 + (size_t)(VC_POINTS / 2) * VC_ROWS * 2 + VC_IDX * 2 + 24) +
                            N * d.Kg * ((size_t)d.L * HS * VC_ROWS * 2 + HS * 8 + d.L * 5 + 4);
        const bool seq_fits = seq_warp_bytes(d.RW, d.L) <= (size_t)160 * 1024 &&
                              FAST_WARPS * camvis_warp_bytes(d.RW, d.Kc) <= (size_t)160 * 1024 && d.L <= 32;   // lane = waypoint
        const bool disabled = off && off[0] == '1';
        g_warn.clear();
        if (!disabled && need < free_b / 2 && seq_fits) {
            h->cache_bytes = need;
            A(d.vc_p, N * d.Kc * VC_POINTS); A(d.vc_mask, N * d.Kc * (VC_POINTS / 2) * VC_ROWS);
            A(d.vc_idx, N * d.Kc * VC_IDX); A(d.vc_meta, N * d.Kc * 2); A(d.vc_lo, N * d.Kc * 2);
            A(d.vg_mask, N * d.Kg * d.L * HS * VC_ROWS); A(d.vg_hval, N * d.Kg * HS);
            A(d.vg_hslot, N * d.Kg * d.L); A(d.vg_nh, N * d.Kg); A(d.vg_reach, N * d.Kg * d.L);
            if (e != cudaSuccess) { heist_destroy(h); return fail((int)e, "heist_create: cudaMalloc (visibility cache): %s", cudaGetErrorString(e)); }
        } else if (!disabled) {
            // not an error -- every env is ray-marched (same results, several times slower) -- but never silent
            char buf[384];
            if (!seq_fits)
                snprintf(buf, sizeof(buf), "heist_create: angular visibility cache disabled: max_path %d / max_cams %d / grid %dx%d do not fit "
                         "the table-driven kernels (max_path <= 32); every env takes the ray-march path", d.L, d.Kc, d.R, d.C);
            else
                snprintf(buf, sizeof(buf), "heist_create: angular visibility cache disabled: it needs %.2f GB for %d envs but only %.2f GB "
                         "are free (limit: half of the free memory); every env takes the ray-march path", need / 1e9, num_envs, free_b / 1e9);
            g_warn = buf;
            const char *req = getenv("HEIST_REQUIRE_VIS_CACHE");
            if (req && req[0] == '1') { heist_destroy(h); return fail(-11, "%s (HEIST_REQUIRE_VIS_CACHE=1)", buf); }
        }
    }
#undef A

    CUDA_TRY_H(upload_nice_table(device, d.deg2rad));
    CUDA_TRY_H(vc_upload_band_tables(device));

    h->step_smem = cta_smem_bytes(d.R, d.C, d.Kc, d.Kg);
    h->camvis_smem = FAST_WARPS * camvis_warp_bytes(d.RW, d.Kc);
    h->camvis_staged_smem = camvis_staged_bytes(d.RW, d.Kc, CVS_MAX_WARPS);
#define SET_FAST(RPL, W) \
    CUDA_TRY_H(cudaFuncSetAttribute(k_cam_vis<RPL, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->camvis_smem));
    SET_FAST(1, 1) SET_FAST(1, 2) SET_FAST(2, 1) SET_FAST(2, 2)
#undef SET_FAST
#define SET_FAST(RPL, W) \
    CUDA_TRY_H(cudaFuncSetAttribute(k_cam_vis_staged<RPL, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->camvis_staged_smem));
    SET_FAST(1, 1) SET_FAST(1, 2) SET_FAST(2, 1) SET_FAST(2, 2)
#undef SET_FAST
    h->seq_smem = seq_warp_bytes(d.RW, d.L);
    if (d.vc_p) {
        CUDA_TRY_H(cudaFuncSetAttribute(k_seq<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->seq_smem));
        CUDA_TRY_H(cudaFuncSetAttribute(k_seq<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->seq_smem));
    }
    if (d.vc_p) {   // single-tick buffers up front, so that step / reset never allocate (CUDA-graph capture)
        h->heads_cap = N * d.Kc; h->scratch_cap = N * d.RW;
        CUDA_TRY_H(cudaMalloc(&h->heads, h->heads_cap * sizeof(double)));
        CUDA_TRY_H(cudaMalloc(&h->scratch, h->scratch_cap * sizeof(uint32_t)));
        {   // k_seq is the serial chain of a launch: its blocks go first
            int lo = 0, hi = 0;
            CUDA_TRY_H(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            CUDA_TRY_H(cudaStreamCreateWithPriority(&h->s_seq, cudaStreamNonBlocking, hi));
        }
        CUDA_TRY_H(cudaMalloc(&h->h_run, N * d.Kc * sizeof(double)));
        CUDA_TRY_H(cudaStreamCreateWithFlags(&h->s_heads, cudaStreamNonBlocking));
        for (int i = 0; i < 64; ++i) CUDA_TRY_H(cudaEventCreateWithFlags(&h->ev_heads[i], cudaEventDisableTiming));
        CUDA_TRY_H(cudaStreamCreateWithFlags(&h->s_fin, cudaStreamNonBlocking));
        CUDA_TRY_H(cudaStreamCreateWithFlags(&h->s_cam2, cudaStreamNonBlocking));
        CUDA_TRY_H(cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking));
        CUDA_TRY_H(cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking));
        CUDA_TRY_H(cudaEventCreateWithFlags(&h->ev_join3, cudaEventDisableTiming));
        for (int i = 0; i < 64; ++i) CUDA_TRY_H(cudaEventCreateWithFlags(&h->ev_h2d[i], cudaEventDisableTiming));
        CUDA_TRY_H(cudaEventCreateWithFlags(&h->ev_join2, cudaEventDisableTiming));
        CUDA_TRY_H(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
        CUDA_TRY_H(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
        for (int i = 0; i < 64; ++i) {
            CUDA_TRY_H(cudaEventCreateWithFlags(&h->ev_cam[i], cudaEventDisableTiming));
            CUDA_TRY_H(cudaEventCreateWithFlags(&h->ev_seq[i], cudaEventDisableTiming));
        }
    }
    h->layout_smem = HEIST_WARPS_PER_CTA * layout_warp_bytes(d.RC, d.RW);
#define SET_SMEM(E, B)                                                                                                    \
    CUDA_TRY_H(cudaFuncSetAttribute(k_step_many<E, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->step_smem)); \
    CUDA_TRY_H(cudaFuncSetAttribute(k_reset<E, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->step_smem));
    SET_SMEM(false, false) SET_SMEM(false, true) SET_SMEM(true, false) SET_SMEM(true, true)
#undef SET_SMEM
    CUDA_TRY_H(cudaFuncSetAttribute(k_set_layout, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->layout_smem));

    // HeistEnvironment.__init__: bordered grid with START/VAULT, solver at start (environment.py:62-96)
    k_pos_table<<<(d.RC + 127) / 128, 128>>>(d);
    k_init_dyn<<<(num_envs + 127) / 128, 128>>>(d);
    LayoutDev none;
    memset(&none, 0, sizeof(none));
    k_set_layout<<<env_blocks(num_envs), HEIST_WARPS_PER_CTA * 32, h->layout_smem>>>(d, none, nullptr, nullptr);
    CUDA_TRY_H(build_cache(h, 0));
    CUDA_TRY_H(cudaMemsetAsync(d.slot2env, 0xFF, sizeof(int32_t) * env_blocks(num_envs) * HEIST_WARPS_PER_CTA, 0));
    k_build_order<<<1, 1024>>>(d, env_blocks(num_envs));
    CUDA_TRY_H(cudaGetLastError());
    CUDA_TRY_H(cudaDeviceSynchronize());
    *out = h;
    return 0;
}

// Per-layout visibility tables (heist_cache.cuh); without the cache every env is "uncached".
static cudaError_t build_cache(HeistHandle *h, cudaStream_t s) {
    if (!h->d.vc_p) {
        const int n = h->N;
        return cudaMemcpyAsync(h->d.n_uncached, &n, sizeof(int), cudaMemcpyHostToDevice, s);
    }
    cudaError_t e = cudaMemsetAsync(h->d.n_uncached, 0, sizeof(int), s);
    if (e != cudaSuccess) return e;
    Dev d = h->d;
    d.strict_tables = h->mode == HEIST_MODE_TABLES;
    k_build_cache<<<h->N, VC_BUILD_THREADS, 0, s>>>(d);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    h->all_cached = -1;
    if ((e = cudaMemcpyAsync(h->n_unc_host, h->d.n_uncached, sizeof(int), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
    return cudaEventRecord(h->ev_unc, s);
}

static int launch_set_layout(HeistHandle *h, const LayoutDev &lz, const int32_t *budget, uint8_t *valid_out,
                             cudaStream_t s) {
    k_set_layout<<<env_blocks(h->N), HEIST_WARPS_PER_CTA * 32, h->layout_smem, s>>>(h->d, lz, budget, valid_out);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(build_cache(h, s));
    CUDA_TRY(cudaMemsetAsync(h->d.slot2env, 0xFF, sizeof(int32_t) * env_blocks(h->N) * HEIST_WARPS_PER_CTA, s));
    k_build_order<<<1, 1024, 0, s>>>(h->d, env_blocks(h->N));
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int heist_decode_validate(HeistHandle *h, const int8_t *asset_map, const float *cam_params,
                                     const int32_t *budget, int allow_cameras, int allow_guards, uint8_t *valid_out,
                                     void *stream) {
    if (!h || !asset_map || !cam_params) return fail(-1, "heist_decode_validate: null argument");
    if (h->d.L < 8) return fail(-7, "heist_decode_validate: max_path must be >= 8 (patrol has 8 waypoints)");
    if (!budget) {   // the decode buys at most budget/cost assets of a kind (networks.py:283-318): they must fit the lists
        const int b = h->p.architect_budget;
        if (b / COST_CAMERA > h->p.max_cams || b / COST_GUARD > h->p.max_guards || b > h->p.max_walls)
            return fail(-12, "heist_decode_validate: budget %d can buy more assets than the capacities hold (max_walls %d, "
                        "max_cams %d, max_guards %d)", b, h->p.max_walls, h->p.max_cams, h->p.max_guards);
    }
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    k_decode<<<env_blocks(h->N), HEIST_WARPS_PER_CTA * 32, 0, s>>>(h->d, h->lz, asset_map, cam_params, budget,
                                                                   allow_cameras, allow_guards);
    CUDA_TRY(cudaGetLastError());
    return launch_set_layout(h, h->lz, budget, valid_out, s);
}

extern "C" int heist_set_layout_explicit(HeistHandle *h, const HeistLayoutArrays *a, const int32_t *budget,
                                         uint8_t *valid_out, void *stream) {
    if (!h || !a) return fail(-1, "heist_set_layout_explicit: null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    LayoutDev lz;
    lz.n_walls = (int32_t *)a->n_walls; lz.wall_rc = (int16_t *)a->wall_rc;
    lz.n_cams = (int32_t *)a->n_cams; lz.cam_rc = (int16_t *)a->cam_rc; lz.cam_f = (double *)a->cam_f;
    lz.cam_range = (int32_t *)a->cam_range;
    lz.n_guards = (int32_t *)a->n_guards; lz.guard_len = (int32_t *)a->guard_len;
    lz.guard_path = (int16_t *)a->guard_path; lz.guard_head = (double *)a->guard_head;
    lz.guard_speed = (int32_t *)a->guard_speed; lz.guard_range = (int32_t *)a->guard_range;
    lz.guard_fov = (double *)a->guard_fov;
    if (lz.n_walls && !lz.wall_rc) return fail(-8, "heist_set_layout_explicit: n_walls without wall_rc");
    if (lz.n_cams && (!lz.cam_rc || !lz.cam_f || !lz.cam_range)) return fail(-8, "heist_set_layout_explicit: camera arrays missing");
    if (lz.n_guards && (!lz.guard_len || !lz.guard_path || !lz.guard_head || !lz.guard_speed || !lz.guard_range || !lz.guard_fov))
        return fail(-8, "heist_set_layout_explicit: guard arrays missing");
    return launch_set_layout(h, lz, budget, valid_out, (cudaStream_t)stream);
}

// Table-driven path for the envs the visibility cache covers (default mode only): k_heads -> k_cam_vis -> k_dyn
// per chunk of ticks (heist_fast.cuh).  do_reset: HeistEnvironment.reset for the masked envs (T ignored).
// Grow-only launch scratch.  (cudaFree synchronises the device, so nothing in flight still uses the old buffer; a
// call that has to grow cannot be captured into a CUDA graph -- run it once un-captured first.)
template <typename T>
static cudaError_t grow(T **buf, size_t *cap, size_t need) {
    if (need <= *cap) return cudaSuccess;
    if (*buf) { cudaError_t e = cudaFree(*buf); *buf = nullptr; *cap = 0; if (e != cudaSuccess) return e; }
    cudaError_t e = cudaMalloc(buf, need * sizeof(T));
    if (e == cudaSuccess) *cap = need;
    return e;
}

// One chunk of ticks [t0, t0 + Tc) through the table-driven kernels.  s_cam / s_walk may be one stream (sequential)
// or two (pipelined: the camera cones of the next chunk are built while k_walk walks this one).
struct FastChunk {
    const int8_t *actions; float *reward; double *reward64; uint8_t *done, *status;
    uint32_t *cam; const double *heads; uint16_t *grec; uint8_t *fin; int32_t *last_t;
    int Tc, autoreset, do_reset, write_traj, store_heading, rev; const uint8_t *mask;
    float *state;   // fused single tick only: dense (3, R, C) state written by k_walk
};

static void launch_cam_vis(HeistHandle *h, const FastChunk &c, cudaStream_t s) {
    const Dev &d = h->d;
    const int nblk = (c.Tc + FAST_TB - 1) / FAST_TB;
    const unsigned g1 = (unsigned)(((long long)h->N * nblk + FAST_WARPS - 1) / FAST_WARPS);
    const uint8_t *m = c.do_reset ? c.mask : nullptr;
    if (c.heads && nblk >= 2 && !c.do_reset) {   // many ticks: per-camera tables staged in shared memory
        int wpc = nblk <= CVS_MAX_WARPS ? nblk : 4;   // warps (tick blocks) per CTA
        static const char *wpc_env = getenv("HEIST_CVS_WPC");   // debug knob
        if (wpc_env && atoi(wpc_env) >= 1 && atoi(wpc_env) <= CVS_MAX_WARPS) wpc = std::min(nblk, atoi(wpc_env));
        const dim3 g((unsigned)h->N, (unsigned)((nblk + wpc - 1) / wpc));
        const size_t sm = camvis_staged_bytes(d.RW, d.Kc, wpc);
#define GO(RPL, W) k_cam_vis_staged<RPL, W><<<g, wpc * 32, sm, s>>>(d, c.Tc, nblk, c.heads, c.cam, c.rev)
        if (d.R > 32) { if (d.C > 32) GO(2, 2); else GO(2, 1); }
        else { if (d.C > 32) GO(1, 2); else GO(1, 1); }
#undef GO
        h->launches += 1;
        return;
    }
#define GO(RPL, W) k_cam_vis<RPL, W><<<g1, FAST_WARPS * 32, h->camvis_smem, s>>>(d, c.Tc, nblk, c.heads, c.cam, m, c.do_reset)
    if (d.R > 32) { if (d.C > 32) GO(2, 2); else GO(2, 1); }
    else { if (d.C > 32) GO(1, 2); else GO(1, 1); }
#undef GO
    h->launches += 1;
}

static void launch_walk(HeistHandle *h, const FastChunk &c, cudaStream_t s) {
    const Dev &d = h->d;
    const unsigned g2 = (unsigned)((h->N + WALK_WARPS - 1) / WALK_WARPS);
    const size_t sm = WALK_WARPS * walk_warp_bytes(d.RW, d.Kc);
    // c.cam == nullptr: the fused single tick (cones + step + auto-reset + dense state, T <= 1)
#define GO1(RPL, W, F) k_walk<RPL, W, F><<<g2, WALK_WARPS * 32, sm, s>>>(d, c.actions, c.do_reset ? 0 : c.Tc, c.autoreset, c.reward, c.reward64, \
                                                                       c.done, c.status, c.cam, c.write_traj, c.do_reset, c.mask, c.store_heading, c.state)
#define GO(RPL, W) do { if (c.cam) GO1(RPL, W, false); else GO1(RPL, W, true); } while (0)
    if (d.R > 32) { if (d.C > 32) GO(2, 2); else GO(2, 1); }
    else { if (d.C > 32) GO(1, 2); else GO(1, 1); }
#undef GO
#undef GO1
    h->launches += 1;
}

static void launch_seq(HeistHandle *h, const FastChunk &c, cudaStream_t s) {
    const Dev &d = h->d;
    // a quad of lanes per env, 8 envs per warp (see k_seq)
    const unsigned g2 = (unsigned)((h->N + SEQ_EPW - 1) / SEQ_EPW);
#define GO(W) k_seq<W><<<g2, SEQ_THREADS, h->seq_smem, s>>>(d, c.actions, c.do_reset ? 0 : c.Tc, c.autoreset, c.reward, c.reward64, \
                                                              c.done, c.status, c.cam, c.grec, c.fin, c.last_t, c.do_reset, c.mask, c.store_heading)
    if (d.C > 32) GO(2); else GO(1);
#undef GO
    h->launches += 1;
}

static void launch_finish(HeistHandle *h, const FastChunk &c, cudaStream_t s) {
    const Dev &d = h->d;
    const dim3 g3((unsigned)((h->N + 7) / 8), (unsigned)(!c.write_traj ? 1 : (c.Tc + FIN_TB - 1) / FIN_TB)), g4((unsigned)((h->N + 7) / 8), (unsigned)c.Tc);
    const uint8_t *m = c.do_reset ? c.mask : nullptr;
    if (c.write_traj && c.autoreset && !c.do_reset) {   // the rollout path: guards OR-ed into the finished camera rows
        if (d.C > 32) k_finish_or<2><<<g3, 256, 0, s>>>(d, c.Tc, c.cam, c.grec, c.fin, c.last_t);
        else k_finish_or<1><<<g3, 256, 0, s>>>(d, c.Tc, c.cam, c.grec, c.fin, c.last_t);
        h->launches += 1;
        return;
    }
#define GO(RPL, W)                                                                                       \
    do {                                                                                                 \
        k_finish<RPL, W><<<g3, 256, 0, s>>>(d, c.Tc, c.cam, c.grec, c.fin, c.last_t, !c.write_traj, m);   \
        if (!!c.write_traj && !c.autoreset) k_fill<W><<<g4, 256, 0, s>>>(d, c.Tc, c.cam, c.fin);          \
    } while (0)
    if (d.R > 32) { if (d.C > 32) GO(2, 2); else GO(2, 1); }
    else { if (d.C > 32) GO(1, 2); else GO(1, 1); }
#undef GO
    h->launches += (!!c.write_traj && !c.autoreset) ? 2 : 1;
}

// Host buffers of heist_step_many_host (pinned): copied chunk by chunk next to the kernels of the pipelined launch.
struct HostIO {
    const int8_t *actions; float *reward; uint8_t *done, *status;
};

#define FAST_PIPE_BLOCKS 8     // tick blocks (of FAST_TB ticks) per pipelined chunk, at most
#define FAST_PIPE_MAX 64       // chunks (events) per launch

// A pipelined launch of `total` ticks is cut into n chunks of (almost) equal numbers of whole tick blocks, at most
// FAST_PIPE_BLOCKS each (200 ticks = 25 blocks -> 5 chunks of 40 ticks): chunk i covers blocks [b0, b0 + nb).
struct ChunkPlan {
    int n, total_blocks;
    int start[FAST_PIPE_MAX + 2];   // first tick block of chunk i; start[n] = total_blocks
    int first_block(int i) const { return start[i]; }
};
static ChunkPlan chunk_plan(int total) {
    ChunkPlan p;
    p.total_blocks = (total + FAST_TB - 1) / FAST_TB;
    // debug knob: explicit chunk sizes in tick blocks, e.g. HEIST_CHUNK_PLAN=7,7,6,3,2 (ignored unless they add up)
    static const char *forced = getenv("HEIST_CHUNK_PLAN");
    if (forced) {
        int sizes[FAST_PIPE_MAX], k = 0, sum = 0;
        for (const char *q = forced; *q && k < FAST_PIPE_MAX;) {
            const int v = atoi(q);
            if (v < 1 || v > FAST_PIPE_BLOCKS) { k = 0; break; }
            sizes[k++] = v; sum += v;
            while (*q && *q != ',') ++q;
            if (*q == ',') ++q;
        }
        if (k >= 1 && sum == p.total_blocks) {
            p.n = k; p.start[0] = 0;
            for (int i = 0; i < k; ++i) p.start[i + 1] = p.start[i] + sizes[i];
            return p;
        }
    }
    // The launch ends with the LAST chunk's k_seq and k_finish_or running alone (nothing left to overlap them with),
    // so the last chunk is short; the others take 7 tick blocks: the shorter chunks hand the sequential chain (k_seq, busy
    // back to back from the end of the first camera chunk on) its rows earlier and more evenly -- six camera CTAs per SM
    // either way (a 7-warp CTA is allocated like an 8-warp one).  Measured on config 2, T = 200
    // (25 blocks): 7,7,7,4 -> 2.07e9 env-steps/s; 8,8,6,3 -> 2.04e9; 7,7,6,5 / 6,7,7,5 / 7,6,6,6 -> 2.02e9; 8,8,5,4 / 8,7,6,4
    // -> 2.04-2.05e9; 7,7,7,3,1 / 7,7,7,2,2 -> 1.98-2.00e9 (finer tapers lose to their extra launches and small CTAs).
    const int B = p.total_blocks, M = FAST_PIPE_BLOCKS, FULL_BLOCKS = 7;
    p.start[0] = 0;
    if (B < 8 || B > FULL_BLOCKS * (FAST_PIPE_MAX - 2)) { p.n = B < 8 ? 1 : FAST_PIPE_MAX + 1; p.start[1] = B; return p; }   // (not pipelined: see fast_pipelined)
    int sizes[FAST_PIPE_MAX], k = 0;   // (filled last chunk first, reversed below)
    if (B <= M) { sizes[k++] = B / 2; sizes[k++] = B - B / 2; }   // (the larger half first)
    else {
        int left = B;
        int fwd[FAST_PIPE_MAX], nf = 0;
        while (left > FULL_BLOCKS + 3) { fwd[nf++] = FULL_BLOCKS; left -= FULL_BLOCKS; }   // 7, 7, 7, ...
        // what is left (4 .. 10 blocks): one short chunk, or two with the shorter one last
        if (left <= 5) fwd[nf++] = left;
        else { const int lastc = left >= 8 ? 4 : 3; fwd[nf++] = left - lastc; fwd[nf++] = lastc; }
        for (int i = nf - 1; i >= 0; --i) sizes[k++] = fwd[i];
    }
    p.n = k;
    for (int i = 0; i < k; ++i) p.start[i + 1] = p.start[i] + sizes[k - 1 - i];
    return p;
}

// Table-driven path for the envs the visibility cache covers (default mode only): k_heads -> k_cam_vis -> k_seq ->
// k_finish (heist_fast.cuh).  do_reset: HeistEnvironment.reset for the masked envs (T ignored).
static bool fast_pipelined(const HeistHandle *h, int total, int autoreset, int do_reset, const uint32_t *vis_traj) {
    const size_t N = h->N, NRW = N * h->d.RW;
    const ChunkPlan p = chunk_plan(total);
    return !do_reset && autoreset && p.n >= 2 && p.n <= FAST_PIPE_MAX && h->s_seq &&
           (vis_traj || (size_t)total * NRW * 4 <= ((size_t)1 << 30)) && (size_t)total * N * h->d.Kc * 8 <= ((size_t)4 << 30);
}

static int launch_fast(HeistHandle *h, const int8_t *actions, int T, int autoreset, float *reward, double *reward64,
                       uint8_t *done, uint8_t *status, uint32_t *vis_traj, int do_reset, const uint8_t *mask,
                       cudaStream_t s, const HostIO *io = nullptr, float *state = nullptr) {
    const Dev &d = h->d;
    const size_t N = h->N, NRW = N * d.RW;
    const int total = do_reset ? 1 : T;
    const unsigned gh = (unsigned)((N * d.Kc + 127) / 128);
    FastChunk c;
    c.autoreset = autoreset; c.do_reset = do_reset; c.write_traj = vis_traj ? 1 : 0; c.mask = mask; c.state = nullptr; c.rev = 0;

    // Pipelined: with auto-reset no env is ever left done at a chunk boundary, so the camera headings of the whole
    // launch are known up front (k_heads once) and k_cam_vis of chunk c + 1 does not wait for k_walk of chunk c.
    const ChunkPlan plan = chunk_plan(total);
    const int n_chunks = plan.n;
    const bool pipelined = fast_pipelined(h, total, autoreset, do_reset, vis_traj);
    if (pipelined) {
        CUDA_TRY(grow(&h->heads, &h->heads_cap, (size_t)((total + FAST_TB - 1) / FAST_TB) * N * d.Kc));
        CUDA_TRY(grow(&h->grec, &h->grec_cap, (size_t)total * N * d.Kg));
        CUDA_TRY(grow(&h->fin, &h->fin_cap, (size_t)total * N));
        CUDA_TRY(grow(&h->last_t, &h->last_cap, (size_t)n_chunks * N));
        if (!vis_traj) CUDA_TRY(grow(&h->scratch, &h->scratch_cap, (size_t)total * NRW));
        uint32_t *cam = vis_traj ? vis_traj : h->scratch;
        static const bool timing = getenv("HEIST_TIMING") != nullptr;   // debug: per-stage timeline of one launch
        cudaEvent_t te[3][FAST_PIPE_MAX + 1];
        if (timing) { for (int a = 0; a < 3; ++a) for (int i = 0; i <= n_chunks; ++i) cudaEventCreate(&te[a][i]); cudaEventRecord(te[0][0], s); }
        // headings: the first chunk's on the caller's stream, the later ones on a side stream while the cameras of
        // the chunks before them are being built (each continues from h_run where the previous one stopped)
        auto chunk_ticks = [&](int i, int &t0) { t0 = plan.first_block(i) * FAST_TB; return std::min(total, plan.first_block(i + 1) * FAST_TB) - t0; };
        {
            int t0; const int tc = chunk_ticks(0, t0);
            k_heads<<<gh, 128, 0, s>>>(d, tc, 0, 1, 0, h->heads, h->h_run);
            h->launches += 1;
        }
        c.store_heading = 0;
        CUDA_TRY(cudaEventRecord(h->ev_fork, s));
        CUDA_TRY(cudaStreamWaitEvent(h->s_heads, h->ev_fork, 0));
        for (int i = 1; i < n_chunks; ++i) {
            int t0; const int tc = chunk_ticks(i, t0);
            k_heads<<<gh, 128, 0, h->s_heads>>>(d, tc, 0, 0, i == n_chunks - 1, h->heads + (size_t)(t0 / FAST_TB) * N * d.Kc, h->h_run);
            h->launches += 1;
            CUDA_TRY(cudaEventRecord(h->ev_heads[i], h->s_heads));
        }
        CUDA_TRY(cudaStreamWaitEvent(h->s_seq, h->ev_fork, 0));
        CUDA_TRY(cudaStreamWaitEvent(h->s_fin, h->ev_fork, 0));
        CUDA_TRY(cudaStreamWaitEvent(h->s_cam2, h->ev_fork, 0));
        if (io) {
            CUDA_TRY(cudaStreamWaitEvent(h->s_h2d, h->ev_fork, 0));
            CUDA_TRY(cudaStreamWaitEvent(h->s_d2h, h->ev_fork, 0));
        }
        for (int i = 0; i < n_chunks; ++i) {
            // camera chunks are independent of each other: alternate two streams so that the tail of one chunk
            // overlaps the head of the next
            cudaStream_t sc = (i & 1) ? h->s_cam2 : s;
            int t0;
            c.Tc = chunk_ticks(i, t0);
            c.rev = i & 1;   // snake over the envs: see k_cam_vis_staged
            const size_t off = (size_t)t0 * N;
            if (i > 0) CUDA_TRY(cudaStreamWaitEvent(sc, h->ev_heads[i], 0));
            c.actions = actions + off; c.reward = reward ? reward + off : nullptr; c.reward64 = reward64 ? reward64 + off : nullptr;
            c.done = done ? done + off : nullptr; c.status = status ? status + off : nullptr;
            c.cam = cam + (size_t)t0 * NRW; c.heads = h->heads + (size_t)(t0 / FAST_TB) * N * d.Kc;
            c.grec = h->grec + off * d.Kg; c.fin = h->fin + off; c.last_t = h->last_t + (size_t)i * N;
            launch_cam_vis(h, c, sc);
            if (timing) cudaEventRecord(te[0][i + 1], sc);
            CUDA_TRY(cudaEventRecord(h->ev_cam[i], sc));
            CUDA_TRY(cudaStreamWaitEvent(h->s_seq, h->ev_cam[i], 0));
            if (io) {   // this chunk's actions: host -> device next to the camera kernel
                CUDA_TRY(cudaMemcpyAsync((void *)c.actions, io->actions + off, (size_t)c.Tc * N, cudaMemcpyHostToDevice, h->s_h2d));
                CUDA_TRY(cudaEventRecord(h->ev_h2d[i], h->s_h2d));
                CUDA_TRY(cudaStreamWaitEvent(h->s_seq, h->ev_h2d[i], 0));
            }
            launch_seq(h, c, h->s_seq);
            if (timing) cudaEventRecord(te[1][i + 1], h->s_seq);
            CUDA_TRY(cudaEventRecord(h->ev_seq[i], h->s_seq));
            if (io) {   // ... and its results: device -> host while the next chunk is walked
                CUDA_TRY(cudaStreamWaitEvent(h->s_d2h, h->ev_seq[i], 0));
                if (io->reward) CUDA_TRY(cudaMemcpyAsync(io->reward + off, c.reward, (size_t)c.Tc * N * 4, cudaMemcpyDeviceToHost, h->s_d2h));
                if (io->done) CUDA_TRY(cudaMemcpyAsync(io->done + off, c.done, (size_t)c.Tc * N, cudaMemcpyDeviceToHost, h->s_d2h));
                if (io->status) CUDA_TRY(cudaMemcpyAsync(io->status + off, c.status, (size_t)c.Tc * N, cudaMemcpyDeviceToHost, h->s_d2h));
            }
            CUDA_TRY(cudaStreamWaitEvent(h->s_fin, h->ev_seq[i], 0));
            launch_finish(h, c, h->s_fin);
            if (timing) cudaEventRecord(te[2][i + 1], h->s_fin);
        }
        CUDA_TRY(cudaEventRecord(h->ev_join, h->s_fin));   // s_fin's last kernel waited for s_seq's last
        CUDA_TRY(cudaStreamWaitEvent(s, h->ev_join, 0));
        CUDA_TRY(cudaEventRecord(h->ev_join2, h->s_cam2));   // (s_heads' kernels were all waited for by camera chunks)
        CUDA_TRY(cudaStreamWaitEvent(s, h->ev_join2, 0));
        if (io) {   // (s_h2d's copies were consumed by s_seq, which s_fin -- joined above -- waited for)
            CUDA_TRY(cudaEventRecord(h->ev_join3, h->s_d2h));
            CUDA_TRY(cudaStreamWaitEvent(s, h->ev_join3, 0));
        }
        CUDA_TRY(cudaGetLastError());
        if (timing) {
            cudaDeviceSynchronize();
            const char *nm[3] = {"cam_vis", "seq", "finish"};
            for (int a = 0; a < 3; ++a) {
                fprintf(stderr, "[heist timing] %-8s done at (us):", nm[a]);
                for (int i = 1; i <= n_chunks; ++i) { float ms = 0; cudaEventElapsedTime(&ms, te[0][0], te[a][i]); fprintf(stderr, " %7.1f", ms * 1e3f); }
                fprintf(stderr, "\n");
            }
            for (int a = 0; a < 3; ++a) for (int i = 0; i <= n_chunks; ++i) cudaEventDestroy(te[a][i]);
        }
        return 0;
    }

    // Sequential chunks on the caller's stream; ticks per chunk bounded by 256 MiB of cam_vis scratch when there is
    // no trajectory buffer to build the maps in.
    int cap = 256;
    if (!vis_traj) cap = (int)std::max<size_t>(1, std::min<size_t>(256, ((size_t)256 << 20) / (NRW * 4)));
    for (int t0 = 0; t0 < total; t0 += cap) {
        const size_t off = (size_t)t0 * N;
        c.Tc = std::min(cap, total - t0);
        const int nblk = (c.Tc + FAST_TB - 1) / FAST_TB;
        CUDA_TRY(grow(&h->heads, &h->heads_cap, (size_t)nblk * N * d.Kc));
        if (nblk > 1) {
            CUDA_TRY(grow(&h->grec, &h->grec_cap, (size_t)c.Tc * N * d.Kg));
            CUDA_TRY(grow(&h->fin, &h->fin_cap, (size_t)c.Tc * N));
            CUDA_TRY(grow(&h->last_t, &h->last_cap, N));
        }
        if (!vis_traj) CUDA_TRY(grow(&h->scratch, &h->scratch_cap, (size_t)c.Tc * NRW));
        c.actions = actions ? actions + off : nullptr; c.reward = reward ? reward + off : nullptr;
        c.reward64 = reward64 ? reward64 + off : nullptr; c.done = done ? done + off : nullptr;
        c.status = status ? status + off : nullptr;
        c.cam = vis_traj ? vis_traj + (size_t)t0 * NRW : h->scratch; c.heads = h->heads;
        c.grec = h->grec; c.fin = h->fin; c.last_t = h->last_t;
        const int heads_final = (autoreset && !do_reset) ? 1 : 0;   // k_heads can store the heading the chunk ends on
        if (nblk > 1) {
            k_heads<<<gh, 128, 0, s>>>(d, c.Tc, do_reset, 1, heads_final, h->heads, nullptr);
            h->launches += 1;
            c.store_heading = !heads_final;
        } else {   // a single tick block: k_cam_vis derives the headings itself, k_walk stores the last
            c.heads = nullptr;
            c.store_heading = 1;
        }
        if (c.Tc == 1 && !vis_traj) {   // one tick (step, step_observe, reset): cones + tick (+ dense state) fused in k_walk
            c.cam = nullptr; c.state = state;
            launch_walk(h, c, s);
            CUDA_TRY(cudaGetLastError());
            continue;
        }
        launch_cam_vis(h, c, s);
        if (nblk > 1) { launch_seq(h, c, s); launch_finish(h, c, s); }   // many ticks: quad-per-env chain + parallel completion
        else launch_walk(h, c, s);                                      // a few ticks: one warp-per-env kernel
        CUDA_TRY(cudaGetLastError());
    }
    return 0;
}

static inline bool use_cache(const HeistHandle *h) {
    return (h->mode == HEIST_MODE_DEFAULT || h->mode == HEIST_MODE_TABLES) && h->d.vc_p != nullptr;
}

// Does any env need the ray-march kernels next to the table-driven ones?  Learnt without a sync from the count the
// cache build copies to pinned memory; until that copy has landed (or while `s` is being captured) the answer is yes.
static bool march_needed(HeistHandle *h, cudaStream_t s) {
    if (!use_cache(h)) return true;
    if (h->mode == HEIST_MODE_TABLES) return false;   // the caller guarantees coverage (violations are reported by set_layout)
    // A launch that is being captured into a CUDA graph will be replayed after later set_layouts, whose envs may not
    // all be covered by the cache: it always carries the ray-march kernel (which exits at once when no env needs it).
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) { cudaGetLastError(); return true; }
    if (h->all_cached < 0) {
        if (cudaEventQuery(h->ev_unc) != cudaSuccess) { cudaGetLastError(); return true; }
        h->all_cached = *h->n_unc_host == 0 ? 1 : 0;
    }
    return h->all_cached != 1;
}

extern "C" int heist_reset(HeistHandle *h, const uint8_t *mask, void *stream) {
    if (!h) return fail(-1, "heist_reset: null handle");
    CUDA_TRY(cudaSetDevice(h->device));
    const int grid = env_blocks(h->N), block = HEIST_WARPS_PER_CTA * 32;
    cudaStream_t s = (cudaStream_t)stream;
    const bool big = h->d.R > 32, exact = h->mode == HEIST_MODE_EXACT;
    Dev d = h->d;
    d.skip_cached = use_cache(h);
    if (d.skip_cached) { int rc = launch_fast(h, nullptr, 0, 0, nullptr, nullptr, nullptr, nullptr, nullptr, 1, mask, s); if (rc) return rc; }
    if (!march_needed(h, s)) return 0;
#define GO(E, B) k_reset<E, B><<<grid, block, h->step_smem, s>>>(d, mask)
    if (exact) { if (big) GO(true, true); else GO(true, false); }
    else { if (big) GO(false, true); else GO(false, false); }
#undef GO
    h->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static int launch_step(HeistHandle *h, const int8_t *actions, int T, int autoreset, float *reward, double *reward64,
                       uint8_t *done, uint8_t *status, uint32_t *vis_traj, cudaStream_t s, float *state = nullptr) {
    const int grid = env_blocks(h->N), block = HEIST_WARPS_PER_CTA * 32;
    const bool big = h->d.R > 32, exact = h->mode == HEIST_MODE_EXACT;
    Dev d = h->d;
    d.skip_cached = use_cache(h);
    if (d.skip_cached) { int rc = launch_fast(h, actions, T, autoreset, reward, reward64, done, status, vis_traj, 0, nullptr, s, nullptr, state); if (rc) return rc; }
    if (!march_needed(h, s)) return 0;
#define GO(E, B) k_step_many<E, B><<<grid, block, h->step_smem, s>>>(d, actions, T, autoreset, reward, reward64, done, status, vis_traj)
    if (exact) { if (big) GO(true, true); else GO(true, false); }
    else { if (big) GO(false, true); else GO(false, false); }
#undef GO
    h->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int heist_launch_count(HeistHandle *h, int64_t *count) {
    if (!h || !count) return fail(-1, "heist_launch_count: null argument");
    *count = h->launches;
    return 0;
}

extern "C" int heist_cache_stats(HeistHandle *h, int32_t *envs_cached, int64_t *cache_bytes, void *stream) {
    if (!h) return fail(-1, "heist_cache_stats: null handle");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    if (envs_cached) {
        int n_unc = 0;
        CUDA_TRY(cudaMemcpy(&n_unc, h->d.n_uncached, sizeof(int), cudaMemcpyDeviceToHost));
        *envs_cached = h->N - n_unc;
    }
    if (cache_bytes) *cache_bytes = (int64_t)h->cache_bytes;
    return 0;
}

extern "C" int heist_set_mode(HeistHandle *h, int mode) {
    if (!h) return fail(-1, "heist_set_mode: null handle");
    if (mode < HEIST_MODE_DEFAULT || mode > HEIST_MODE_TABLES) return fail(-10, "heist_set_mode: unknown mode %d", mode);
    if (mode == HEIST_MODE_TABLES && !h->d.vc_p) return fail(-13, "heist_set_mode: HEIST_MODE_TABLES needs the visibility cache, which this handle does not have (%s)", g_warn.c_str());
    h->mode = mode;
    return 0;
}

extern "C" int heist_step(HeistHandle *h, const int8_t *actions, float *reward, double *reward64, uint8_t *done,
                          uint8_t *status, void *stream) {
    if (!h || !actions) return fail(-1, "heist_step: null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    return launch_step(h, actions, 1, 0, reward, reward64, done, status, nullptr, (cudaStream_t)stream);
}

extern "C" int heist_step_many(HeistHandle *h, const int8_t *actions, int T, int autoreset, float *reward,
                               uint8_t *done, uint8_t *status, uint32_t *vis_traj, void *stream) {
    if (!h || !actions) return fail(-1, "heist_step_many: null argument");
    if (T < 0) return fail(-9, "heist_step_many: negative T");
    if (T == 0) return 0;
    CUDA_TRY(cudaSetDevice(h->device));
    { int rc = launch_step(h, actions, T, autoreset, reward, nullptr, done, status, vis_traj, (cudaStream_t)stream); if (rc) return rc; }
    if (T >= 8 && !use_cache(h)) {  // re-deal the warp slots by the work each env actually did (its reset rate included)
        CUDA_TRY(cudaMemsetAsync(h->d.slot2env, 0xFF, sizeof(int32_t) * env_blocks(h->N) * HEIST_WARPS_PER_CTA,
                                 (cudaStream_t)stream));
        k_build_order<<<1, 1024, 0, (cudaStream_t)stream>>>(h->d, env_blocks(h->N));
        h->launches += 1;
        CUDA_TRY(cudaGetLastError());
    }
    return 0;
}

extern "C" int heist_step_many_host(HeistHandle *h, const int8_t *actions_host, int T, int autoreset, float *reward_host,
                                    uint8_t *done_host, uint8_t *status_host, uint32_t *vis_traj, void *stream) {
    if (!h || !actions_host) return fail(-1, "heist_step_many_host: null argument");
    if (T < 0) return fail(-9, "heist_step_many_host: negative T");
    if (T == 0) return 0;
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)T * h->N;
    if (n > h->st_cap) {
        if (h->st_act) { cudaFree(h->st_act); cudaFree(h->st_rew); cudaFree(h->st_done); cudaFree(h->st_status); h->st_act = nullptr; h->st_cap = 0; }
        CUDA_TRY(cudaMalloc(&h->st_act, n)); CUDA_TRY(cudaMalloc(&h->st_rew, n * 4));
        CUDA_TRY(cudaMalloc(&h->st_done, n)); CUDA_TRY(cudaMalloc(&h->st_status, n));
        h->st_cap = n;
    }
    // Copies ride along the pipelined launch when it serves every env; otherwise they bracket the call.
    if (use_cache(h) && !march_needed(h, s) && fast_pipelined(h, T, autoreset, 0, vis_traj)) {
        HostIO io = {actions_host, reward_host, done_host, status_host};
        return launch_fast(h, h->st_act, T, autoreset, h->st_rew, nullptr, h->st_done, h->st_status, vis_traj, 0, nullptr, s, &io);
    }
    CUDA_TRY(cudaMemcpyAsync(h->st_act, actions_host, n, cudaMemcpyHostToDevice, s));
    { int rc = heist_step_many(h, h->st_act, T, autoreset, h->st_rew, h->st_done, h->st_status, vis_traj, stream); if (rc) return rc; }
    if (reward_host) CUDA_TRY(cudaMemcpyAsync(reward_host, h->st_rew, n * 4, cudaMemcpyDeviceToHost, s));
    if (done_host) CUDA_TRY(cudaMemcpyAsync(done_host, h->st_done, n, cudaMemcpyDeviceToHost, s));
    if (status_host) CUDA_TRY(cudaMemcpyAsync(status_host, h->st_status, n, cudaMemcpyDeviceToHost, s));
    return 0;
}

static int launch_observe(HeistHandle *h, float *state, cudaStream_t s, int only_uncached) {
    Dev d = h->d;
    d.skip_cached = only_uncached;   // the fused tick has already written the states of the table-driven envs
    if (d.C % 4 == 0 && ((uintptr_t)state & 15) == 0) {
        const int nq = 3 * (d.RC / 4);
        const int bx = (nq + 255) / 256;                     // blocks per env, each <= 256 threads ...
        const int bdx = ((nq + bx - 1) / bx + 31) & ~31;     // ... evenly split and rounded up to a warp
        const int by = d.N < 148 * 16 ? d.N : 148 * 16;      // grid-stride over envs beyond that
        k_observe_vec4<<<dim3(bx, by), bdx, 0, s>>>(d, (float4 *)state);
    } else {
        long long total = (long long)d.N * 3 * d.RC;
        long long blocks = (total + 255) / 256;
        if (blocks > 148LL * 32) blocks = 148LL * 32;
        k_observe_scalar<<<(int)blocks, 256, 0, s>>>(d, state);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int heist_observe(HeistHandle *h, float *state, void *stream) {
    if (!h || !state) return fail(-1, "heist_observe: null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    return launch_observe(h, state, (cudaStream_t)stream, 0);
}

extern "C" int heist_step_observe(HeistHandle *h, const int8_t *actions, int autoreset, float *reward, uint8_t *done,
                                  uint8_t *status, float *state, void *stream) {
    if (!h || !actions || !state) return fail(-1, "heist_step_observe: null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    // Table-driven envs: ONE kernel does the tick, the auto-reset and the dense state (k_walk, fused mode).  A
    // separate observe pass runs only for ray-marched envs (it exits at once when there are none), or for everything
    // when the state cannot take 16-byte stores.
    const bool fused = use_cache(h) && h->d.C % 4 == 0 && ((uintptr_t)state & 15) == 0;
    const bool others = march_needed(h, s);
    { int rc = launch_step(h, actions, 1, autoreset, reward, nullptr, done, status, nullptr, s, fused ? state : nullptr); if (rc) return rc; }
    if (fused && !others) return 0;
    return launch_observe(h, state, s, fused ? 1 : 0);
}

extern "C" int heist_expand_states(HeistHandle *h, const uint32_t *vis_bits, const int32_t *pos, const int32_t *env_idx,
                                   int M, float *state, void *stream) {
    if (!h || !vis_bits || !pos || !env_idx || !state) return fail(-1, "heist_expand_states: null argument");
    if (M < 0) return fail(-9, "heist_expand_states: negative M");
    if (M == 0) return 0;
    CUDA_TRY(cudaSetDevice(h->device));
    const Dev &d = h->d;
    const int bx = (3 * d.RC + 255) / 256;
    const int by = M < 148 * 16 ? M : 148 * 16;
    k_expand_states<<<dim3(bx, by), 256, 0, (cudaStream_t)stream>>>(d, vis_bits, pos, env_idx, M, state);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int heist_observation_vectors(HeistHandle *h, float *obs_vec, void *stream) {
    if (!h || !obs_vec) return fail(-1, "heist_observation_vectors: null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    k_obs_vectors<<<(h->N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(h->d, obs_vec);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int heist_get_state(HeistHandle *h, HeistStateView *v) {
    if (!h || !v) return fail(-1, "heist_get_state: null argument");
    const Dev &d = h->d;
    v->tile = d.tile; v->wall_bits = d.wall; v->vis_bits = d.vis; v->env_static = d.env_s; v->env_dyn = d.env_d;
    v->cam_f = d.cam_f; v->cam_i = d.cam_i; v->cam_heading = d.cam_heading;
    v->guard_fov = d.guard_fov; v->guard_i = d.guard_i; v->guard_path = d.guard_path;
    v->guard_heading = d.guard_heading; v->guard_idx = d.guard_idx; v->wall_accepted = d.wall_ok;
    return 0;
}

extern "C" int heist_gae(const float *rew, const float *val, const uint8_t *done, int T, int n_cols, double gamma,
                         double gae_lambda, float *adv, float *ret, int device, void *stream) {
    if (!rew || !val || !done || !adv || !ret) return fail(-1, "heist_gae: null argument");
    if (T < 0 || n_cols < 0) return fail(-9, "heist_gae: negative size");
    if (T == 0 || n_cols == 0) return 0;
    CUDA_TRY(cudaSetDevice(device));
    const float g = (float)gamma;                  // torch casts the Python scalar to the tensor dtype
    const float gl = (float)(gamma * gae_lambda);  // self.gamma * self.gae_lambda is a Python double product
    // Columns are the only parallelism (the scan is sequential in t for bit-exactness).  With few columns use
    // one warp per CTA spread over all SMs and a deep load prefetch; with many, wider CTAs and less prefetch.
    const bool rows16 = (n_cols % GAE_CW) == 0 && ((((uintptr_t)rew | (uintptr_t)val) & 15) == 0) && (((uintptr_t)done & 7) == 0) &&
                        ((n_cols * 4) % 16) == 0;
    if (n_cols <= 16384 && rows16 && gae_staged_bytes(T) <= (size_t)200 * 1024) {   // few columns: blocks staged in shared memory
        static size_t attr_set[64];   // dynamic shared memory limit configured so far, per device
        const size_t sm = gae_staged_bytes(T);
        if (sm > 48 * 1024 && device >= 0 && device < 64 && !attr_set[device]) {
            CUDA_TRY(cudaFuncSetAttribute(k_gae_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)200 * 1024)));
            attr_set[device] = 1;
        }
        k_gae_staged<<<n_cols / GAE_CW, 32, sm, (cudaStream_t)stream>>>(rew, val, done, T, n_cols, g, gl, adv, ret);
    } else if (n_cols <= 32768)
        k_gae<32, 32><<<(n_cols + 31) / 32, 32, 0, (cudaStream_t)stream>>>(rew, val, done, T, n_cols, g, gl, adv, ret);
    else if (n_cols <= 131072)
        k_gae<16, 64><<<(n_cols + 63) / 64, 64, 0, (cudaStream_t)stream>>>(rew, val, done, T, n_cols, g, gl, adv, ret);
    else
        k_gae<8, 128><<<(n_cols + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rew, val, done, T, n_cols, g, gl, adv, ret);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int heist_architect_reward(HeistHandle *h, double *reward_out, double *solve_rate_out, void *stream) {
    if (!h || !reward_out) return fail(-1, "heist_architect_reward: null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    k_architect_reward<<<(h->N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(h->d, reward_out, solve_rate_out);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

__global__ void k_debug_ray_dirs(const double *__restrict__ a, int n, double deg2rad, double *__restrict__ dx, double *__restrict__ dy) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x, y;
    ray_dir(a[i], deg2rad, x, y);
    dx[i] = x; dy[i] = y;
}

extern "C" int heist_debug_ray_dirs(int device, const double *angles_deg, int n, double *dx, double *dy, void *stream) {
    if (!angles_deg || !dx || !dy) return fail(-1, "heist_debug_ray_dirs: null argument");
    if (n <= 0) return 0;
    CUDA_TRY(cudaSetDevice(device));
    k_debug_ray_dirs<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(angles_deg, n, 3.14159265358979323846 / 180.0, dx, dy);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int heist_check_errors(HeistHandle *h, void *stream) {
    if (!h) return fail(-1, "heist_check_errors: null handle");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    int flags = 0;
    CUDA_TRY(cudaMemcpy(&flags, h->d.err, sizeof(int), cudaMemcpyDeviceToHost));
    if (flags) {
        CUDA_TRY(cudaMemset(h->d.err, 0, sizeof(int)));
        return fail(-100 - flags, "device-side error:%s%s%s%s%s%s",
                    (flags & ERR_CAPACITY) ? " capacity exceeded (max_walls/max_cams/max_guards/max_path)" : "",
                    (flags & ERR_WAYPOINT) ? " guard waypoint outside the grid" : "",
                    (flags & ERR_RAYS) ? " fov/vision_range too large" : "",
                    (flags & ERR_BOUNDS) ? " cell-map access out of range (debug build)" : "",
                    (flags & ERR_STATE) ? " guard state in the state view (heading / waypoint) is not one the guard can reach on its path" : "",
                    (flags & ERR_UNCOVERED) ? " layout not covered by the visibility cache (HEIST_MODE_TABLES)" : "");
    }
    return 0;
}
