// Tiled K1 / K3 for ANY geometry (seq_len, num_basis, degree, DoF) on the tokenizer's own times — the path of the
// reference's shipped configuration (train.sh / train/train_beast.py:34-36: 50 basis functions, degree 0, 1000 bins,
// actions [10, 32]: 1 600 tokens per trajectory, 1 280 B read and 19 200 B written per encode) and of every shape the
// T = 50 / nb = 10 bulk-copy kernels (spline_encode.cu, spline_decode.cu) do not cover.
//
// A CTA walks tiles of S trajectories through shared memory: global loads and stores are contiguous runs of a
// trajectory's samples / tokens / coefficients (every buffer of the reference's layout is contiguous per trajectory),
// the strided accesses — sample (t, dof) of a column, coefficient (slot, k) of a token — hit shared memory.  Sums run
// over the non-zero BAND of each projector / basis row (plan.cu): a degree-0 projector row touches the few samples of
// its interval (none at all for 40 of 50 basis functions when nb > T), a degree-p basis row p + 1 coefficients; the
// skipped terms are exact zeros, t / k ascending with fused multiply-adds as in the other kernels, so the results are
// bit-identical to them.
#include "common.cuh"

namespace beast {

constexpr int kTiledThreadsV1 = 512;

// encode: tile of trajectories -> coefficients, tokens, optional column min / max.
// Per CTA, once: tab[r] = k | slot << 16 for every token position r = k*D + slot of a trajectory, the quantiser
// constants of every column, the bands.  Per tile: (1) the samples, one contiguous run; (2) one TOKEN per thread and
// step in token order — consecutive threads store consecutive int64 tokens — each a sum over the band of projector row
// k (shared-memory samples), quantised exactly; the coefficient goes to its '(d t)' place in shared memory;
// (3) coefficients leave as one contiguous run.
__global__ void __launch_bounds__(kTiledThreadsV1)
encode_tiled_v1_kernel(const float* __restrict__ traj, long long B, int T, int D, int nb, int n_joint,
                    const int* __restrict__ slot_to_dof, const float* __restrict__ Pj, const float* __restrict__ Pg,
                    const int* __restrict__ bands, const float* __restrict__ w_min, const float* __restrict__ w_max,
                    float vm1, long long offset, float* __restrict__ params_out, long long* __restrict__ tokens_out,
                    float* __restrict__ bmin, float* __restrict__ bmax, int S) {
    extern __shared__ __align__(16) float smem_f[];
    const int row_in = T * D, row_out = D * nb;
    float* y = smem_f;                                          // [S][T][D]
    const int pitch = nb | 1, prow = D * pitch;                 // odd pitch: the '(d t)' staging is free of bank conflicts
    float* par = y + (((size_t)S * row_in + 3) & ~(size_t)3);   // [S][D][pitch] coefficients, slot major
    float* qtab = par + (((size_t)S * prow + 3) & ~(size_t)3);  // [D*nb][4] quantiser constants in TOKEN order, or [2][D*nb] min / max
    float* s_mn = qtab;
    float* s_mx = qtab + row_out;
    int* tab = (int*)(qtab + (size_t)4 * row_out);              // [D*nb] token position r -> k | slot << 16 | (joint ? 0 : 1 << 31)
    int* ctab = tab + row_out;                                  // [D*nb] column c = slot*nb + k -> slot*pitch + k
    int* ztok = ctab + row_out;                                 // [D*nb] token of a zero coefficient, token order (empty projector rows)
    int* s_band = ztok + row_out;                               // [2][2*nb]
    int* s_dof = s_band + 4 * nb;                               // [D]
    const bool want_mm = bmin != nullptr, want_par = params_out != nullptr, want_tok = tokens_out != nullptr;
    const int tid = threadIdx.x;
    for (int r = tid; r < row_out; r += kTiledThreadsV1) {
        const int k = r / D, slot = r - k * D;
        const bool grip_r = slot >= n_joint;
        const bool empty = bands[(grip_r ? 2 * nb : 0) + 2 * k] >= bands[(grip_r ? 2 * nb : 0) + 2 * k + 1];
        // bits 0-13 k, 14 = projector row k is all zero (the coefficient is exactly 0 for every trajectory), 16-30 slot, 31 gripper
        tab[r] = k | (empty ? 0x4000 : 0) | (slot << 16) | (grip_r ? (int)0x80000000u : 0);
        ctab[r] = (r / nb) * pitch + (r % nb);                  // r read as a column index here
        if (want_tok) {
            const float lo0 = w_min[slot * nb + k], hi0 = w_max[slot * nb + k];
            ztok[r] = (int)quantize_one(0.0f, lo0, hi0, quant_scale(lo0, hi0), vm1);
        }
        if (want_mm) { s_mn[r] = __int_as_float(0x7f800000); s_mx[r] = __int_as_float(0xff800000); }
        else if (want_tok) {
            QuantCol qc;
            qc.init(w_min[slot * nb + k], w_max[slot * nb + k]);
            qtab[4 * r] = qc.lo; qtab[4 * r + 1] = qc.hi; qtab[4 * r + 2] = qc.scale; qtab[4 * r + 3] = qc.rcp;
        }
    }
    for (int i = tid; i < 4 * nb; i += kTiledThreadsV1) s_band[i] = bands[i];
    for (int i = tid; i < D; i += kTiledThreadsV1) s_dof[i] = slot_to_dof[i];
    const long long n_tiles = (B + S - 1) / S;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long b0 = tile * S;
        const int ns = (int)((B - b0) < S ? (B - b0) : S);
        __syncthreads();                                      // tables ready / previous tile's staging drained
        {   // phase 1: the tile's samples, contiguous in global memory
            const float* src = traj + b0 * row_in;
            const int n = ns * row_in;
            if ((((uintptr_t)src) & 15u) == 0 && (n & 3) == 0) {
                for (int i = tid; i < (n >> 2); i += kTiledThreadsV1) ((float4*)y)[i] = __ldcs((const float4*)src + i);
            } else {
                for (int i = tid; i < n; i += kTiledThreadsV1) y[i] = __ldcs(src + i);
            }
        }
        __syncthreads();
        for (int tr = 0; tr < ns; ++tr) {                     // phase 2
            const float* ytr = y + (size_t)tr * row_in;
            float* ptr = par + (size_t)tr * prow;
            long long* ttr = want_tok ? tokens_out + (b0 + tr) * row_out : nullptr;
#pragma unroll 4
            for (int r = tid; r < row_out; r += kTiledThreadsV1) {
                const int e = tab[r];
                const int k = e & 0x3fff, slot = (e >> 16) & 0x7fff;
                if (e & 0x4000) {                             // empty projector row: coefficient 0, token a per-column constant
                    if (want_par || want_mm) ptr[slot * pitch + k] = 0.0f;
                    if (want_tok) ttr[r] = (long long)ztok[r] + offset;
                    continue;
                }
                const bool grip = e < 0;
                const float* Pk = (grip ? Pg : Pj) + k * T;
                const int t0 = s_band[(grip ? 2 * nb : 0) + 2 * k], t1 = s_band[(grip ? 2 * nb : 0) + 2 * k + 1];
                const float* col = ytr + s_dof[slot];
                float acc = 0.0f;
                for (int t = t0; t < t1; ++t) acc = fmaf(__ldg(Pk + t), col[t * D], acc);
                if (want_par || want_mm) ptr[slot * pitch + k] = acc;
                if (want_tok) {
                    const float4 q = *(const float4*)(qtab + 4 * r);
                    QuantCol qc;
                    qc.lo = q.x; qc.hi = q.y; qc.scale = q.z; qc.rcp = q.w;
                    ttr[r] = quantize_col(acc, qc, vm1) + offset;
                }
            }
        }
        if (want_par || want_mm) {
            __syncthreads();
            if (want_par) {                                   // phase 3: coefficients leave as contiguous rows
                for (int tr = 0; tr < ns; ++tr) {
                    float* dst = params_out + (b0 + tr) * row_out;
                    const float* src = par + (size_t)tr * prow;
#pragma unroll 4
                    for (int c = tid; c < row_out; c += kTiledThreadsV1) __stcs(dst + c, src[ctab[c]]);
                }
            }
            if (want_mm) {                                    // a thread owns columns c, c + blockDim, ...: no atomics needed
                for (int c = tid; c < row_out; c += kTiledThreadsV1) {
                    float mn = s_mn[c], mx = s_mx[c];
                    const int pc = ctab[c];
                    for (int tr = 0; tr < ns; ++tr) {
                        const float v = par[(size_t)tr * prow + pc];
                        mn = fminf(mn, v); mx = fmaxf(mx, v);
                    }
                    s_mn[c] = mn; s_mx[c] = mx;
                }
            }
        }
    }
    if (want_mm) {
        __syncthreads();
        for (int c = tid; c < row_out; c += kTiledThreadsV1)
            if (s_mn[c] <= s_mx[c]) { atomic_min_f32(bmin + c, s_mn[c]); atomic_max_f32(bmax + c, s_mx[c]); }
    }
}

// decode: tile of token rows (or coefficient rows) -> trajectories.
// Per tile: (1) one TOKEN per thread and step in token order (contiguous int64 loads), dequantised exactly into shared
// memory; (2) one output SAMPLE per thread and step in output order (contiguous fp32 stores): the sum over the band of
// basis row t against the slot's coefficients.
template <bool FROM_TOKENS>
__global__ void __launch_bounds__(kTiledThreadsV1, 4)
decode_tiled_v1_kernel(const long long* __restrict__ tokens, const float* __restrict__ params, long long B, int T, int D,
                    int nb, int n_joint, const int* __restrict__ slot_to_dof, const float* __restrict__ phi_j,
                    const float* __restrict__ phi_g, const int* __restrict__ bands, const float* __restrict__ w_min,
                    const float* __restrict__ w_max, float vm1, long long offset, const float* __restrict__ init_p,
                    float* __restrict__ out, int S) {
    extern __shared__ __align__(16) float smem_f[];
    const int row_in = D * nb, row_out = T * D;
    float* c_s = smem_f;                                        // [S][nb][D]   (token order: slot minor)
    float* lohi = c_s + (((size_t)S * row_in + 3) & ~(size_t)3);    // [D*nb][2] bounds in token order
    int* tab_in = (int*)(lohi + (size_t)2 * row_in);            // [D*nb] k | slot << 16
    int* tab_out = tab_in + row_in;                             // [T*D]  t | slot << 16 | grip << 31, output order [t][dof]
    int* s_band = tab_out + row_out;                            // [2][2*T]
    int* s_dof = s_band + 4 * T;                                // [D]
    const int tid = threadIdx.x;
    const float rcp_vm1 = __frcp_rn(vm1);                       // exact invariant division by V - 1 (common.cuh)
    for (int r = tid; r < row_in; r += kTiledThreadsV1) {
        const int k = r / D, slot = r - k * D;
        tab_in[r] = k | (slot << 16);
        if (FROM_TOKENS) { lohi[2 * r] = w_min[slot * nb + k]; lohi[2 * r + 1] = w_max[slot * nb + k]; }
    }
    for (int i = tid; i < D; i += kTiledThreadsV1) s_dof[i] = slot_to_dof[i];
    for (int i = tid; i < 4 * T; i += kTiledThreadsV1) s_band[i] = bands[4 * nb + i];
    __syncthreads();
    for (int q = tid; q < row_out; q += kTiledThreadsV1) {
        const int t = q / D, dof = q - t * D;
        int slot = 0;
        for (int sidx = 0; sidx < D; ++sidx) if (s_dof[sidx] == dof) slot = sidx;
        tab_out[q] = t | (slot << 16) | (slot < n_joint ? 0 : (int)0x80000000u);
    }
    const long long n_tiles = (B + S - 1) / S;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long b0 = tile * S;
        const int ns = (int)((B - b0) < S ? (B - b0) : S);
        __syncthreads();
        for (int tr = 0; tr < ns; ++tr) {                     // phase 1
            float* ctr = c_s + (size_t)tr * row_in;
            const long long* ttr = FROM_TOKENS ? tokens + (b0 + tr) * row_in : nullptr;
            const float* ptr = FROM_TOKENS ? nullptr : params + (b0 + tr) * row_in;
            const float* ip = init_p ? init_p + (b0 + tr) * D : nullptr;
#pragma unroll 4
            for (int r = tid; r < row_in; r += kTiledThreadsV1) {
                const int e = tab_in[r];
                const int k = e & 0xffff, slot = e >> 16;
                float v;
                if (FROM_TOKENS) v = dequantize_fast(__ldcs(ttr + r) - offset, lohi[2 * r], lohi[2 * r + 1], vm1, rcp_vm1);
                else v = ptr[slot * nb + k];
                if (k == 0 && ip && slot < n_joint) v = ip[s_dof[slot]];
                ctr[r] = v;
            }
        }
        __syncthreads();
        for (int tr = 0; tr < ns; ++tr) {                     // phase 2
            const float* ctr = c_s + (size_t)tr * row_in;
            float* otr = out + (b0 + tr) * row_out;
#pragma unroll 4
            for (int q = tid; q < row_out; q += kTiledThreadsV1) {
                const int e = tab_out[q];
                const int t = e & 0xffff, slot = (e >> 16) & 0x7fff;
                const bool grip = e < 0;
                const float* row = (grip ? phi_g : phi_j) + (size_t)t * nb;
                const int k0 = s_band[(grip ? 2 * T : 0) + 2 * t], k1 = s_band[(grip ? 2 * T : 0) + 2 * t + 1];
                const float* col = ctr + slot;
                float acc = 0.0f;
                for (int k = k0; k < k1; ++k) acc = fmaf(__ldg(row + k), col[(size_t)k * D], acc);
                __stcs(otr + q, acc);
            }
        }
    }
}

static int tile_rows_v1(size_t bytes_per_traj, size_t extra, int max_smem, size_t budget) {
    if (budget > (size_t)max_smem) budget = (size_t)max_smem;
    if (bytes_per_traj + extra > budget) budget = (size_t)max_smem;
    if (bytes_per_traj + extra > budget) return 0;
    size_t s = (budget - extra) / bytes_per_traj;
    if (s > 64) s = 64;
    return (int)s;
}

int launch_encode_tiled_v1(const Plan* p, const float* traj, long long B, const float* w_min, const float* w_max,
                        long long offset, float* params_out, long long* tokens_out, float* bmin, float* bmax,
                        cudaStream_t st) {
    const int T = p->T, D = p->D, nb = p->nb;
    const size_t per_traj = ((size_t)T * D + (size_t)D * (nb | 1)) * sizeof(float);
    if (bmin && tokens_out) return BEAST_E_UNSUPPORTED;      // the two tables share one shared-memory region
    if (nb > 0x3fff || D > 0x7fff || p->V > 0x7fffffff) return BEAST_E_UNSUPPORTED;
    // quantiser constants / min-max (4 floats per column), token-position, column and zero-token tables, bands, slot map
    const size_t extra = ((size_t)7 * D * nb + (size_t)4 * nb + D) * sizeof(float) + 64;
    const int S = tile_rows_v1(per_traj, extra, p->max_smem_optin, 104 * 1024);
    if (S < 1 || !p->bands_d) return BEAST_E_UNSUPPORTED;
    const size_t smem = (size_t)S * per_traj + extra;
    static size_t granted[kMaxDevices] = {};
    if (int rc = opt_in_smem(encode_tiled_v1_kernel, smem, granted)) return rc;
    const long long n_tiles = (B + S - 1) / S;
    long long per_sm = (long long)(220 * 1024) / (long long)(smem + 1024);
    if (per_sm > 2048 / kTiledThreadsV1) per_sm = 2048 / kTiledThreadsV1;
    if (per_sm < 1) per_sm = 1;
    long long grid = n_tiles < (long long)p->num_sms * per_sm ? n_tiles : (long long)p->num_sms * per_sm;
    encode_tiled_v1_kernel<<<(unsigned)grid, kTiledThreadsV1, smem, st>>>(traj, B, T, D, nb, p->n_joint, p->slot_to_dof_d,
                                                                    p->proj_joint_d, p->proj_grip_d, p->bands_d, w_min, w_max,
                                                                    (float)(p->V - 1), offset, params_out, tokens_out, bmin,
                                                                    bmax, S);
    count_launch();
    BEAST_CHECK_LAUNCH();
    return BEAST_OK;
}

int launch_decode_tiled_v1(const Plan* p, const long long* tokens, const float* params, long long B, const float* w_min,
                        const float* w_max, long long offset, const float* init_p, float* out, cudaStream_t st) {
    const int T = p->T, D = p->D, nb = p->nb;
    if (p->nc != nb || !p->bands_d) return BEAST_E_UNSUPPORTED;           // pinned control points: generic kernel
    if (nb > 0xffff || T > 0xffff || D > 0x7fff) return BEAST_E_UNSUPPORTED;
    const size_t per_traj = (size_t)D * nb * sizeof(float);
    const size_t extra = ((size_t)3 * D * nb + (size_t)T * D + (size_t)4 * T + D) * sizeof(float) + 64;
    const int S = tile_rows_v1(per_traj, extra, p->max_smem_optin, 54 * 1024);   // four CTAs per SM: the token loads need the warps
    if (S < 1) return BEAST_E_UNSUPPORTED;
    const size_t smem = (size_t)S * per_traj + extra;
    static size_t granted_t[kMaxDevices] = {}, granted_p[kMaxDevices] = {};
    if (int rc = tokens ? opt_in_smem(decode_tiled_v1_kernel<true>, smem, granted_t) : opt_in_smem(decode_tiled_v1_kernel<false>, smem, granted_p))
        return rc;
    const long long n_tiles = (B + S - 1) / S;
    long long per_sm = (long long)(220 * 1024) / (long long)(smem + 1024);
    if (per_sm > 2048 / kTiledThreadsV1) per_sm = 2048 / kTiledThreadsV1;
    if (per_sm < 1) per_sm = 1;
    long long grid = n_tiles < (long long)p->num_sms * per_sm ? n_tiles : (long long)p->num_sms * per_sm;
    if (tokens)
        decode_tiled_v1_kernel<true><<<(unsigned)grid, kTiledThreadsV1, smem, st>>>(tokens, nullptr, B, T, D, nb, p->n_joint,
            p->slot_to_dof_d, p->phi_joint_d, p->phi_grip_d, p->bands_d, w_min, w_max, (float)(p->V - 1), offset, init_p, out, S);
    else
        decode_tiled_v1_kernel<false><<<(unsigned)grid, kTiledThreadsV1, smem, st>>>(nullptr, params, B, T, D, nb, p->n_joint,
            p->slot_to_dof_d, p->phi_joint_d, p->phi_grip_d, p->bands_d, nullptr, nullptr, 0.0f, 0, init_p, out, S);
    count_launch();
    BEAST_CHECK_LAUNCH();
    return BEAST_OK;
}

}  // namespace beast
