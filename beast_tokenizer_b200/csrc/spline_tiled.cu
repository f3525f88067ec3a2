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

constexpr int kTiledThreads = 256;

// encode: tile of trajectories -> coefficients, tokens, optional column min / max
__global__ void __launch_bounds__(kTiledThreads)
encode_tiled_kernel(const float* __restrict__ traj, long long B, int T, int D, int nb, int n_joint,
                    const int* __restrict__ slot_to_dof, const float* __restrict__ Pj, const float* __restrict__ Pg,
                    const int* __restrict__ bands, const float* __restrict__ w_min, const float* __restrict__ w_max,
                    float vm1, long long offset, float* __restrict__ params_out, long long* __restrict__ tokens_out,
                    float* __restrict__ bmin, float* __restrict__ bmax, int S) {
    extern __shared__ __align__(16) float smem_f[];
    const int row_in = T * D, row_out = D * nb;
    float* y = smem_f;                                        // [S][T][D]
    float* par = y + (size_t)S * row_in;                      // [S][D][nb]   (coefficient layout '(d t)')
    float* s_mn = par + (size_t)S * row_out;                  // [D*nb] x 2 when bmin
    float* s_mx = s_mn + row_out;
    float* qtab = s_mn;                                       // ... or the quantiser constants [D*nb][4] when tokens
    const bool want_mm = bmin != nullptr, want_par = params_out != nullptr, want_tok = tokens_out != nullptr;
    const int tid = threadIdx.x;
    if (want_mm)
        for (int c = tid; c < row_out; c += kTiledThreads) { s_mn[c] = __int_as_float(0x7f800000); s_mx[c] = __int_as_float(0xff800000); }
    if (want_tok)                                             // per-column constants of the exact quantiser, once per CTA
        for (int c = tid; c < row_out; c += kTiledThreads) {
            QuantCol qc;
            qc.init(w_min[c], w_max[c]);
            qtab[4 * c] = qc.lo; qtab[4 * c + 1] = qc.hi; qtab[4 * c + 2] = qc.scale; qtab[4 * c + 3] = qc.rcp;
        }
    const long long n_tiles = (B + S - 1) / S;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long b0 = tile * S;
        const int ns = (int)((B - b0) < S ? (B - b0) : S);
        __syncthreads();                                      // previous tile's staging has been drained
        {   // phase 1: the tile's samples, contiguous in global memory
            const float* src = traj + b0 * row_in;
            const int n = ns * row_in;
            if ((((uintptr_t)src) & 15u) == 0 && (n & 3) == 0) {
                for (int i = tid; i < (n >> 2); i += kTiledThreads) ((float4*)y)[i] = __ldcs((const float4*)src + i);
            } else {
                for (int i = tid; i < n; i += kTiledThreads) y[i] = __ldcs(src + i);
            }
        }
        __syncthreads();
        // phase 2: one (trajectory, slot) column per thread, k ascending: for a fixed k the lanes of a warp hold
        // adjacent slots, so the int64 token stores of a step are contiguous runs; index arithmetic once per column
        const int n_out = ns * row_out;
        for (int col = tid; col < ns * D; col += kTiledThreads) {
            const int tr = col / D, slot = col - tr * D;
            const bool joint = slot < n_joint;
            const float* P = joint ? Pj : Pg;
            const int* band = bands + (joint ? 0 : 2 * nb);
            const float* ycol = y + (size_t)tr * row_in + slot_to_dof[slot];
            float* pcol = par + (size_t)tr * row_out + slot * nb;
            long long* tcol = want_tok ? tokens_out + (b0 + tr) * row_out + slot : nullptr;
            const float* qcol = qtab + (size_t)slot * nb * 4;
            for (int k = 0; k < nb; ++k) {
                const int t0 = band[2 * k], t1 = band[2 * k + 1];
                const float* Pk = P + (size_t)k * T;
                float acc = 0.0f;
                for (int t = t0; t < t1; ++t) acc = fmaf(__ldg(Pk + t), ycol[t * D], acc);
                if (want_par || want_mm) pcol[k] = acc;
                if (want_tok) {
                    QuantCol qc;
                    qc.lo = qcol[4 * k]; qc.hi = qcol[4 * k + 1]; qc.scale = qcol[4 * k + 2]; qc.rcp = qcol[4 * k + 3];
                    tcol[(size_t)k * D] = quantize_col(acc, qc, vm1) + offset;
                }
            }
        }
        if (want_par || want_mm) {
            __syncthreads();
            if (want_par) {                                   // phase 3: coefficients leave as contiguous rows
                float* dst = params_out + b0 * row_out;
                if ((((uintptr_t)dst) & 15u) == 0 && (n_out & 3) == 0) {
                    for (int i = tid; i < (n_out >> 2); i += kTiledThreads) __stcs((float4*)dst + i, ((const float4*)par)[i]);
                } else {
                    for (int i = tid; i < n_out; i += kTiledThreads) dst[i] = par[i];
                }
            }
            if (want_mm) {                                    // a thread owns columns c, c + 256, ...: no atomics needed
                for (int c = tid; c < row_out; c += kTiledThreads) {
                    float mn = s_mn[c], mx = s_mx[c];
                    for (int tr = 0; tr < ns; ++tr) {
                        const float v = par[(size_t)tr * row_out + c];
                        mn = fminf(mn, v); mx = fmaxf(mx, v);
                    }
                    s_mn[c] = mn; s_mx[c] = mx;
                }
            }
        }
    }
    if (want_mm) {
        __syncthreads();
        for (int c = tid; c < row_out; c += kTiledThreads)
            if (s_mn[c] <= s_mx[c]) { atomic_min_f32(bmin + c, s_mn[c]); atomic_max_f32(bmax + c, s_mx[c]); }
    }
}

// decode: tile of token rows (or coefficient rows) -> trajectories
template <bool FROM_TOKENS>
__global__ void __launch_bounds__(kTiledThreads)
decode_tiled_kernel(const long long* __restrict__ tokens, const float* __restrict__ params, long long B, int T, int D,
                    int nb, int n_joint, const int* __restrict__ slot_to_dof, const float* __restrict__ phi_j,
                    const float* __restrict__ phi_g, const int* __restrict__ bands, const float* __restrict__ w_min,
                    const float* __restrict__ w_max, float vm1, long long offset, const float* __restrict__ init_p,
                    float* __restrict__ out, int S) {
    extern __shared__ __align__(16) float smem_f[];
    const int row_in = D * nb, row_out = T * D;
    float* c_s = smem_f;                                      // [S][nb][D]   (token order: slot minor)
    float* o_s = c_s + (size_t)S * row_in;                    // [S][T][D]
    const int tid = threadIdx.x;
    const float rcp_vm1 = __frcp_rn(vm1);                     // exact invariant division by V - 1 (common.cuh)
    const long long n_tiles = (B + S - 1) / S;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long b0 = tile * S;
        const int ns = (int)((B - b0) < S ? (B - b0) : S);
        __syncthreads();
        // phase 1: one (trajectory, slot) column per thread, k ascending (adjacent lanes = adjacent slots: the int64
        // token loads of a step are contiguous runs); coefficients land in token order [k][slot]
        for (int col = tid; col < ns * D; col += kTiledThreads) {
            const int tr = col / D, slot = col - tr * D;
            const bool pin0 = init_p != nullptr && slot < n_joint;
            float* ccol = c_s + (size_t)tr * row_in + slot;
            if (FROM_TOKENS) {
                const long long* tcol = tokens + (b0 + tr) * row_in + slot;
                const float* lo = w_min + slot * nb;
                const float* hi = w_max + slot * nb;
                for (int k = 0; k < nb; ++k)
                    ccol[(size_t)k * D] = dequantize_fast(__ldcs(tcol + (size_t)k * D) - offset, __ldg(lo + k), __ldg(hi + k), vm1, rcp_vm1);
            } else {
                const float* pcol = params + (b0 + tr) * row_in + slot * nb;
                for (int k = 0; k < nb; ++k) ccol[(size_t)k * D] = pcol[k];
            }
            if (pin0) ccol[0] = init_p[(b0 + tr) * D + slot_to_dof[slot]];
        }
        __syncthreads();
        // phase 2: the same columns, t ascending over the non-zero band of every basis row
        const int n_out = ns * row_out;
        for (int col = tid; col < ns * D; col += kTiledThreads) {
            const int tr = col / D, slot = col - tr * D;
            const bool joint = slot < n_joint;
            const float* phi = joint ? phi_j : phi_g;
            const int* band = bands + 4 * nb + (joint ? 0 : 2 * T);
            const float* ccol = c_s + (size_t)tr * row_in + slot;
            float* ocol = o_s + (size_t)tr * row_out + slot_to_dof[slot];
            for (int t = 0; t < T; ++t) {
                const float* row = phi + (size_t)t * nb;
                float acc = 0.0f;
                for (int k = band[2 * t]; k < band[2 * t + 1]; ++k) acc = fmaf(__ldg(row + k), ccol[(size_t)k * D], acc);
                ocol[(size_t)t * D] = acc;
            }
        }
        __syncthreads();
        float* dst = out + b0 * row_out;                      // phase 3: contiguous rows out
        if ((((uintptr_t)dst) & 15u) == 0 && (n_out & 3) == 0) {
            for (int i = tid; i < (n_out >> 2); i += kTiledThreads) __stcs((float4*)dst + i, ((const float4*)o_s)[i]);
        } else {
            for (int i = tid; i < n_out; i += kTiledThreads) dst[i] = o_s[i];
        }
    }
}

static int tile_rows(size_t bytes_per_traj, size_t extra, int max_smem) {
    // three CTAs per SM when possible: ~70 KB each
    size_t budget = 104 * 1024;                          // two CTAs per SM
    if (budget > (size_t)max_smem) budget = (size_t)max_smem;
    if (bytes_per_traj + extra > budget) budget = (size_t)max_smem;
    if (bytes_per_traj + extra > budget) return 0;
    size_t s = (budget - extra) / bytes_per_traj;
    if (s > 64) s = 64;
    return (int)s;
}

int launch_encode_tiled(const Plan* p, const float* traj, long long B, const float* w_min, const float* w_max,
                        long long offset, float* params_out, long long* tokens_out, float* bmin, float* bmax,
                        cudaStream_t st) {
    const int T = p->T, D = p->D, nb = p->nb;
    const size_t per_traj = ((size_t)T * D + (size_t)D * nb) * sizeof(float);
    if (bmin && tokens_out) return BEAST_E_UNSUPPORTED;      // the two tables share one shared-memory region
    const size_t extra = (bmin ? (size_t)2 : tokens_out ? (size_t)4 : (size_t)0) * D * nb * sizeof(float);
    const int S = tile_rows(per_traj, extra, p->max_smem_optin);
    if (S < 1 || !p->bands_d) return BEAST_E_UNSUPPORTED;
    const size_t smem = (size_t)S * per_traj + extra;
    static size_t granted[kMaxDevices] = {};
    if (int rc = opt_in_smem(encode_tiled_kernel, smem, granted)) return rc;
    const long long n_tiles = (B + S - 1) / S;
    long long per_sm = (long long)(220 * 1024) / (long long)(smem + 1024);
    if (per_sm > 2048 / kTiledThreads) per_sm = 2048 / kTiledThreads;
    if (per_sm < 1) per_sm = 1;
    long long grid = n_tiles < (long long)p->num_sms * per_sm ? n_tiles : (long long)p->num_sms * per_sm;
    encode_tiled_kernel<<<(unsigned)grid, kTiledThreads, smem, st>>>(traj, B, T, D, nb, p->n_joint, p->slot_to_dof_d,
                                                                    p->proj_joint_d, p->proj_grip_d, p->bands_d, w_min, w_max,
                                                                    (float)(p->V - 1), offset, params_out, tokens_out, bmin,
                                                                    bmax, S);
    count_launch();
    BEAST_CHECK_LAUNCH();
    return BEAST_OK;
}

int launch_decode_tiled(const Plan* p, const long long* tokens, const float* params, long long B, const float* w_min,
                        const float* w_max, long long offset, const float* init_p, float* out, cudaStream_t st) {
    const int T = p->T, D = p->D, nb = p->nb;
    if (p->nc != nb || !p->bands_d) return BEAST_E_UNSUPPORTED;           // pinned control points: generic kernel
    const size_t per_traj = ((size_t)T * D + (size_t)D * nb) * sizeof(float);
    const int S = tile_rows(per_traj, 0, p->max_smem_optin);
    if (S < 1) return BEAST_E_UNSUPPORTED;
    const size_t smem = (size_t)S * per_traj;
    static size_t granted_t[kMaxDevices] = {}, granted_p[kMaxDevices] = {};
    if (int rc = tokens ? opt_in_smem(decode_tiled_kernel<true>, smem, granted_t) : opt_in_smem(decode_tiled_kernel<false>, smem, granted_p))
        return rc;
    const long long n_tiles = (B + S - 1) / S;
    long long per_sm = (long long)(220 * 1024) / (long long)(smem + 1024);
    if (per_sm > 2048 / kTiledThreads) per_sm = 2048 / kTiledThreads;
    if (per_sm < 1) per_sm = 1;
    long long grid = n_tiles < (long long)p->num_sms * per_sm ? n_tiles : (long long)p->num_sms * per_sm;
    if (tokens)
        decode_tiled_kernel<true><<<(unsigned)grid, kTiledThreads, smem, st>>>(tokens, nullptr, B, T, D, nb, p->n_joint,
            p->slot_to_dof_d, p->phi_joint_d, p->phi_grip_d, p->bands_d, w_min, w_max, (float)(p->V - 1), offset, init_p, out, S);
    else
        decode_tiled_kernel<false><<<(unsigned)grid, kTiledThreads, smem, st>>>(nullptr, params, B, T, D, nb, p->n_joint,
            p->slot_to_dof_d, p->phi_joint_d, p->phi_grip_d, p->bands_d, nullptr, nullptr, 0.0f, 0, init_p, out, S);
    count_launch();
    BEAST_CHECK_LAUNCH();
    return BEAST_OK;
}

}  // namespace beast
