// heist_stream.cuh -- HBM-streaming kernels: dense observation expand, GAE/returns scan,
// architect reward.  Reference: HeistEnvironment.get_state_tensor / _get_observation
// (environment.py:305-374), SolverAgent._compute_gae (agents/solver.py:228-244),
// RewardCalculator.calculate_architect_reward (rewards.py:43-73).
#pragma once
#include "heist_common.cuh"

// pos_tab[cell] = float32(-0.3 * (manhattan(cell, vault) / (R + C)))   (environment.py:361-365;
// numpy >= 2: np.float32 += python float rounds the float to float32, then adds in float32)
__global__ void k_pos_table(Dev D) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= D.RC) return;
    int r = i / D.C, c = i - r * D.C;
    int d = abs(r - D.vault_r) + abs(c - D.vault_c);
    double v = __dmul_rn(-0.3, __ddiv_rn((double)d, (double)(D.R + D.C)));
    D.pos_tab[i] = (float)v;
}

// state[N][3][R][C] float32.  One thread per 4 consecutive cells of one channel (C % 4 == 0)
// -> one 16-byte store; inputs are bytes / bits and stay in L1/L2.  blockIdx.y strides over envs,
// blockIdx.x * blockDim.x + tid over the 3 * R*C/4 quads of one env (32-bit index math only; the block
// width is 3*R*C/4 rounded up to a warp so that small grids do not idle half a block).
__global__ void __launch_bounds__(1024) k_observe_vec4(Dev D, float4 *__restrict__ out) {
    const int quads = D.RC >> 2;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= 3 * quads) return;
    if (D.skip_cached && *D.n_uncached == 0) return;   // only the ray-marched envs are wanted, and there are none
    const int ch = (q >= quads) + (q >= 2 * quads);
    const int cell = (q - ch * quads) << 2;
    const int r = cell / D.C, c = cell - r * D.C;
    const int vcell = D.vault_r * D.C + D.vault_c;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ch == 2) g = *reinterpret_cast<const float4 *>(D.pos_tab + cell);
#pragma unroll 4
    for (int env = blockIdx.y; env < D.N; env += gridDim.y) {
        if (D.skip_cached && D.env_cached[env]) continue;
        float4 v;
        if (ch == 0) {  // occupancy = grid.astype(float32) / 5   (:319)
            uchar4 t = *reinterpret_cast<const uchar4 *>(D.tile + (size_t)env * D.RC + cell);
            // float32(code) / 5 == float32(code) * 0.2f bit-for-bit for the six tile codes 0..5 (checked
            // exhaustively; the parity tests compare every code): one multiply instead of an IEEE division
            v.x = __fmul_rn((float)t.x, 0.2f); v.y = __fmul_rn((float)t.y, 0.2f);
            v.z = __fmul_rn((float)t.z, 0.2f); v.w = __fmul_rn((float)t.w, 0.2f);
        } else if (ch == 1) {  // visibility (:322)
            uint32_t bits = D.vis[(size_t)env * D.RW + r * D.W + (c >> 5)] >> (c & 31);
            v.x = (float)(bits & 1u); v.y = (float)((bits >> 1) & 1u);
            v.z = (float)((bits >> 2) & 1u); v.w = (float)((bits >> 3) & 1u);
        } else {  // position channel (:356-365): solver +1, vault -1 (vault wins), plus gradient
            int pos = D.env_d[(size_t)env * 8];
            int scell = (pos & 0xffff) * D.C + (pos >> 16);
            float b0 = (cell + 0 == vcell) ? -1.0f : ((cell + 0 == scell) ? 1.0f : 0.0f);
            float b1 = (cell + 1 == vcell) ? -1.0f : ((cell + 1 == scell) ? 1.0f : 0.0f);
            float b2 = (cell + 2 == vcell) ? -1.0f : ((cell + 2 == scell) ? 1.0f : 0.0f);
            float b3 = (cell + 3 == vcell) ? -1.0f : ((cell + 3 == scell) ? 1.0f : 0.0f);
            v.x = __fadd_rn(b0, g.x); v.y = __fadd_rn(b1, g.y); v.z = __fadd_rn(b2, g.z); v.w = __fadd_rn(b3, g.w);
        }
        __stcs(out + (size_t)env * 3 * quads + q, v);  // write-once stream: do not keep it in L2
    }
}

// Minibatch assembly from a packed rollout buffer (agents/solver.py:155-169 gathers dense states; here the
// buffer keeps 4*R*W + 4 bytes per transition and the dense [M][3][R][C] states are rebuilt on demand):
// transition m = (visibility bitmap vis[m], solver position pos[m] = row | col << 16, env env_idx[m] whose
// static tile codes give channel 0).  Same arithmetic as k_observe_*; one thread per cell.
__global__ void __launch_bounds__(256)
k_expand_states(Dev D, const uint32_t *__restrict__ vis, const int32_t *__restrict__ pos, const int32_t *__restrict__ env_idx,
                int M, float *__restrict__ out) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= 3 * D.RC) return;
    const int ch = (q >= D.RC) + (q >= 2 * D.RC);
    const int cell = q - ch * D.RC;
    const int r = cell / D.C, c = cell - r * D.C;
    const int vcell = D.vault_r * D.C + D.vault_c;
    const float g = D.pos_tab[cell];
    for (int m = blockIdx.y; m < M; m += gridDim.y) {
        float v;
        if (ch == 0) v = __fmul_rn((float)D.tile[(size_t)env_idx[m] * D.RC + cell], 0.2f);  // == / 5 for codes 0..5
        else if (ch == 1) v = (float)((vis[(size_t)m * D.RW + r * D.W + (c >> 5)] >> (c & 31)) & 1u);
        else {
            const int p = pos[m];
            const int scell = (p & 0xffff) * D.C + (p >> 16);
            v = __fadd_rn((cell == vcell) ? -1.0f : ((cell == scell) ? 1.0f : 0.0f), g);
        }
        __stcs(out + (size_t)m * 3 * D.RC + q, v);
    }
}

// Scalar fallback for C % 4 != 0.
__global__ void __launch_bounds__(256) k_observe_scalar(Dev D, float *__restrict__ out) {
    const long long total = (long long)D.N * 3 * D.RC;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int env = (int)(idx / (3 * D.RC));
        int rem = (int)(idx - (long long)env * 3 * D.RC);
        int ch = rem / D.RC;
        int cell = rem - ch * D.RC;
        float v;
        if (ch == 0) v = __fmul_rn((float)D.tile[(size_t)env * D.RC + cell], 0.2f);  // == / 5 for codes 0..5
        else if (ch == 1) {
            int r = cell / D.C, c = cell - r * D.C;
            v = (float)((D.vis[(size_t)env * D.RW + r * D.W + (c >> 5)] >> (c & 31)) & 1u);
        } else {
            int pos = D.env_d[(size_t)env * 8];
            int scell = (pos & 0xffff) * D.C + (pos >> 16);
            int vcell = D.vault_r * D.C + D.vault_c;
            float b = (cell == vcell) ? -1.0f : ((cell == scell) ? 1.0f : 0.0f);
            v = __fadd_rn(b, D.pos_tab[cell]);
        }
        out[idx] = v;
    }
}

// solver_position, vault_direction, time_feature (environment.py:324-337): double divide -> float32.
__global__ void k_obs_vectors(Dev D, float *__restrict__ out) {
    int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= D.N) return;
    int pos = D.env_d[(size_t)env * 8], tick = D.env_d[(size_t)env * 8 + 1];
    int r = pos & 0xffff, c = pos >> 16;
    float *o = out + (size_t)env * 5;
    o[0] = (float)__ddiv_rn((double)r, (double)D.R);
    o[1] = (float)__ddiv_rn((double)c, (double)D.C);
    o[2] = (float)__ddiv_rn((double)(D.vault_r - r), (double)D.R);
    o[3] = (float)__ddiv_rn((double)(D.vault_c - c), (double)D.C);
    o[4] = (float)__ddiv_rn((double)tick, (double)D.max_steps);
}

// GAE + returns, one thread per time-major column, fp32 ops in the reference's order, no FMA:
//   delta = r + (g * next_v) * (1 - d) - v ;  A = delta + (gl * (1 - d)) * A ;  ret = A + v
// Loads do not depend on the recurrence, so each thread prefetches CH timesteps at a time.
template <int CH, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
k_gae(const float *__restrict__ rew, const float *__restrict__ val, const uint8_t *__restrict__ done, int T, int n,
      float g, float gl, float *__restrict__ adv, float *__restrict__ ret) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    float last = 0.0f, nv = 0.0f;
    int t = T - 1;
    while (t >= 0) {
        float r[CH], v[CH], om[CH];
        const int cnt = min(CH, t + 1);
#pragma unroll
        for (int k = 0; k < CH; ++k)
            if (k < cnt) {
                size_t i = (size_t)(t - k) * n + j;
                r[k] = rew[i]; v[k] = val[i]; om[k] = __fsub_rn(1.0f, (float)done[i]);
            }
#pragma unroll
        for (int k = 0; k < CH; ++k)
            if (k < cnt) {
                size_t i = (size_t)(t - k) * n + j;
                float delta = __fsub_rn(__fadd_rn(r[k], __fmul_rn(__fmul_rn(g, nv), om[k])), v[k]);
                float A = __fadd_rn(delta, __fmul_rn(__fmul_rn(gl, om[k]), last));
                adv[i] = A;
                ret[i] = __fadd_rn(A, v[k]);
                last = A;
                nv = v[k];
            }
        t -= cnt;
    }
}

// Few columns (a 4 096-env rollout): the scan is then latency-bound -- a handful of warps, each waiting for its own
// loads batch after batch.  k_gae_staged gives a warp only CW = 8 columns (4x the warps), first pulls the warp's
// whole [T][CW] block of rewards, values and done flags into shared memory with one burst of 16-byte asynchronous
// copies (every load in flight at once), then runs the same sequential recurrence from shared memory in lanes
// 0 .. CW-1 and streams the results out.  Needs n % CW == 0 and 16-byte aligned rows (else the plain kernels).
#define GAE_CW 8
__device__ __forceinline__ void gae_cp_async(void *dst_smem, const void *src, int bytes16) {
    if (bytes16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__host__ __device__ inline size_t gae_staged_bytes(int T) { return (size_t)T * GAE_CW * 9 + 16; }

__global__ void __launch_bounds__(32)
k_gae_staged(const float *__restrict__ rew, const float *__restrict__ val, const uint8_t *__restrict__ done, int T, int n,
             float g, float gl, float *__restrict__ adv, float *__restrict__ ret) {
    extern __shared__ __align__(16) unsigned char gsm[];
    float *rs = reinterpret_cast<float *>(gsm), *vs = rs + (size_t)T * GAE_CW;   // [T][CW]
    uint8_t *ds = reinterpret_cast<uint8_t *>(vs + (size_t)T * GAE_CW);           // [T][CW]
    const int lane = threadIdx.x, j0 = blockIdx.x * GAE_CW;
    // a row of the block is CW floats = 32 bytes = two 16-byte pieces (8 bytes of done flags): lane -> (row, piece)
    for (int i = lane; i < 2 * T; i += 32) {
        const int t = i >> 1, h = i & 1;
        gae_cp_async(rs + t * GAE_CW + 4 * h, rew + (size_t)t * n + j0 + 4 * h, 1);
        gae_cp_async(vs + t * GAE_CW + 4 * h, val + (size_t)t * n + j0 + 4 * h, 1);
    }
    for (int t = lane; t < T; t += 32) gae_cp_async(ds + t * GAE_CW, done + (size_t)t * n + j0, 0);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    if (lane >= GAE_CW) return;
    const int j = j0 + lane;
    float last = 0.0f, nv = 0.0f;
#pragma unroll 8
    for (int t = T - 1; t >= 0; --t) {
        const float r = rs[t * GAE_CW + lane], v = vs[t * GAE_CW + lane];
        const float om = __fsub_rn(1.0f, (float)ds[t * GAE_CW + lane]);
        const float delta = __fsub_rn(__fadd_rn(r, __fmul_rn(__fmul_rn(g, nv), om)), v);
        const float A = __fadd_rn(delta, __fmul_rn(__fmul_rn(gl, om), last));
        const size_t i = (size_t)t * n + j;
        adv[i] = A;
        ret[i] = __fadd_rn(A, v);
        last = A;
        nv = v;
    }
}

// calculate_architect_reward (rewards.py:43-73) with solve_rate = vault / finished episodes
// (training.py:535-550) per env.
__global__ void k_architect_reward(Dev D, double *__restrict__ reward, double *__restrict__ solve_rate_out) {
    int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= D.N) return;
    const int32_t *d = D.env_d + (size_t)env * 8;
    int valid = D.env_s[(size_t)env * 4 + 2];
    int nv = d[5], total = d[5] + d[6] + d[7];
    double sr = total > 0 ? __ddiv_rn((double)nv, (double)total) : 0.0;
    double rw;
    if (!valid) rw = -1.0;
    else {
        rw = __dmul_rn(__dsub_rn(1.0, sr), 1.0);
        rw = __dadd_rn(0.0, rw);
        if (sr > 0.8) rw = __dadd_rn(rw, -0.5);
        if (0.2 <= sr && sr <= 0.6) rw = __dadd_rn(rw, 0.2);
    }
    reward[env] = rw;
    if (solve_rate_out) solve_rate_out[env] = sr;
}
