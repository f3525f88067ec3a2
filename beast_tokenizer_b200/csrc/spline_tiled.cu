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
    const bool want_mm = bmin != nullptr, want_par = params_out != nullptr, want_tok = tokens_out != nullptr;
    const int tid = threadIdx.x;
    if (want_mm)
        for (int c = tid; c < row_out; c += kTiledThreads) { s_mn[c] = __int_as_float(0x7f800000); s_mx[c] = __int_as_float(0xff800000); }
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
        // phase 2: one output per thread and step, token order (k major, slot minor): coalesced token stores
        const int n_out = ns * row_out;
        for (int o = tid; o < n_out; o += kTiledThreads) {
            const int tr = o / row_out, r = o - tr * row_out;
            const int k = r / D, slot = r - k * D;
            const bool joint = slot < n_joint;
            const float* P = (joint ? Pj : Pg) + (size_t)k * T;
            const int* band = bands + (joint ? 0 : 2 * nb) + 2 * k;
            const float* col = y + (size_t)tr * row_in + slot_to_dof[slot];
            float acc = 0.0f;
            for (int t = band[0]; t < band[1]; ++t) acc = fmaf(__ldg(P + t), col[t * D], acc);
            const int c = slot * nb + k;
            if (want_par || want_mm) par[(size_t)tr * row_out + c] = acc;
            if (want_tok) {
                const float lo = __ldg(w_min + c), hi = __ldg(w_max + c);
                tokens_out[(b0 + tr) * row_out + r] = quantize_one(acc, lo, hi, quant_scale(lo, hi), vm1) + offset;
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
    const long long n_tiles = (B + S - 1) / S;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long b0 = tile * S;
        const int ns = (int)((B - b0) < S ? (B - b0) : S);
        __syncthreads();
        const int n_in = ns * row_in;
        for (int i = tid; i < n_in; i += kTiledThreads) {     // phase 1: coefficients of the tile, '(t d)' order
            const int tr = i / row_in, r = i - tr * row_in;
            const int k = r / D, slot = r - k * D;
            const int c = slot * nb + k;
            float v;
            if (FROM_TOKENS) v = dequantize_one(__ldcs(tokens + b0 * row_in + i) - offset, __ldg(w_min + c), __ldg(w_max + c), vm1);
            else v = params[(b0 + tr) * row_in + c];
            if (k == 0 && init_p && slot < n_joint) v = init_p[(b0 + tr) * D + slot_to_dof[slot]];
            c_s[i] = v;
        }
        __syncthreads();
        const int n_out = ns * row_out;
        for (int o = tid; o < n_out; o += kTiledThreads) {    // phase 2: one sample per thread, slot minor
            const int tr = o / row_out, r = o - tr * row_out;
            const int t = r / D, slot = r - t * D;
            const bool joint = slot < n_joint;
            const float* phi = (joint ? phi_j : phi_g) + (size_t)t * nb;
            const int* band = bands + 4 * nb + (joint ? 0 : 2 * T) + 2 * t;
            const float* col = c_s + (size_t)tr * row_in + slot;
            float acc = 0.0f;
            for (int k = band[0]; k < band[1]; ++k) acc = fmaf(__ldg(phi + k), col[k * D], acc);
            o_s[(size_t)tr * row_out + t * D + slot_to_dof[slot]] = acc;
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
    size_t budget = 70 * 1024;
    if (budget > (size_t)max_smem) budget = (size_t)max_smem;
    if (bytes_per_traj + extra > budget) budget = (size_t)max_smem;
    if (bytes_per_traj + extra > budget) return 0;
    size_t s = (budget - extra) / bytes_per_traj;
    if (s > 32) s = 32;
    return (int)s;
}

int launch_encode_tiled(const Plan* p, const float* traj, long long B, const float* w_min, const float* w_max,
                        long long offset, float* params_out, long long* tokens_out, float* bmin, float* bmax,
                        cudaStream_t st) {
    const int T = p->T, D = p->D, nb = p->nb;
    const size_t per_traj = ((size_t)T * D + (size_t)D * nb) * sizeof(float);
    const size_t extra = bmin ? (size_t)2 * D * nb * sizeof(float) : 0;
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
