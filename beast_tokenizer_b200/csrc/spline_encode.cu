// K1 — fused B-spline least-squares fit + quantisation  (replaces
// BEASTBsplineTokenizer.encode, beast/beast_bspline_tokenizer.py:399-428, and the
// UniformBSpline.learn_mp_params_from_trajs calls under it, mp/uni_bspline.py:471-602).
//
// Math: every trajectory shares self.times, so the reference's per-trajectory ridge solve
// collapses to  w[slot, k] = sum_t P[k, t] * y[t, dof(slot)]  with one P = (Phi^T Phi + 1e-9 I)^-1 Phi^T
// per spline (joint degree_p, gripper degree 0), computed once on the host.
//
// Fast kernel (seq_len = 50, num_basis = 10 — every BASELINE config): HBM-bound streaming.
//   * persistent CTAs, one per SM; tile = S trajectories (S*D <= 224 columns);
//   * warp 7 / lane 0 is the copy thread: 1-D bulk TMA (cp.async.bulk, UBLKCP) global->shared into a
//     3-deep ring, and bulk shared->global stores of the staged outputs — every HBM byte moves in
//     16-byte-aligned bursts, none through registers;
//   * warps 0-6: one thread per (trajectory, slot) column: 50 LDS + 500 FFMA whose P operand comes
//     straight from the constant bank (the table travels as a __grid_constant__ kernel parameter),
//     then the exact quantiser, staged '(t d)' int64 tokens + '(d t)' fp32 coefficients;
//   * full/empty mbarriers per stage, no CTA-wide barrier in the steady state.
// Generic kernel: any geometry, one thread per column, same accumulation order (bit-identical
// coefficients), plain loads/stores.  Also handles the ragged tail of the fast path.
#include <cstdlib>
#include "common.cuh"

namespace beast {

constexpr int kComputeWarps = 7;
constexpr int kComputeThreads = kComputeWarps * 32;
constexpr int kThreads = kComputeThreads + 32;

// Projector tables as the kernel sees them: [t][k] with k padded to a multiple of 4 so that one
// 16-byte uniform constant load (LDCU.128) feeds four FFMAs.
template <int T, int NB>
struct alignas(16) EncTables {
    static constexpr int NBP = (NB + 3) & ~3;
    float pj[T * NBP];
    float pg[T * NBP];
};

struct EncArgs {
    const float* traj;
    float* params_out;
    long long* tokens_out;
    const float* w_min;
    const float* w_max;
    long long offset;
    float vm1;
    int D, n_joint, S, n_tiles;
    int slot_to_dof[BEAST_MAX_SLOTS];
};

__host__ __device__ inline uint32_t round_up_128(uint32_t x) { return (x + 127u) & ~127u; }

template <int T, int NB, bool GRIP>
__device__ __forceinline__ void fit_column(const EncTables<T, NB>& tab, const float* __restrict__ y, int D,
                                           float (&acc)[NB]) {
    constexpr int NBP = EncTables<T, NB>::NBP;
#pragma unroll
    for (int k = 0; k < NB; ++k) acc[k] = 0.0f;
#pragma unroll
    for (int t = 0; t < T; ++t) {
        const float v = y[t * D];
#pragma unroll
        for (int k = 0; k < NB; ++k) acc[k] = fmaf(GRIP ? tab.pg[t * NBP + k] : tab.pj[t * NBP + k], v, acc[k]);
    }
}

template <int T, int NB, int NS, int DT>
__global__ void __launch_bounds__(kThreads, 1)
encode_fast_kernel(const __grid_constant__ EncTables<T, NB> tab, const __grid_constant__ EncArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int D = DT ? DT : a.D;          // compile-time DoF count for the common shapes (7, 14)
    const int S = a.S;
    const uint32_t in_bytes = (uint32_t)S * T * D * 4u;
    const uint32_t tok_bytes = (uint32_t)S * NB * D * 8u;
    const uint32_t par_bytes = (uint32_t)S * NB * D * 4u;
    const uint32_t in_stride = round_up_128(in_bytes);
    const uint32_t tok_stride = round_up_128(tok_bytes);
    const uint32_t out_stride = tok_stride + round_up_128(par_bytes);
    unsigned char* out_base = smem + NS * in_stride;
    uint64_t* bars = (uint64_t*)(out_base + 2 * out_stride);
    uint64_t* in_full = bars;
    uint64_t* in_empty = bars + NS;
    uint64_t* out_full = bars + 2 * NS;
    uint64_t* out_empty = bars + 2 * NS + 2;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&in_full[s], 1); mbar_init(&in_empty[s], kComputeWarps); }
        for (int s = 0; s < 2; ++s) { mbar_init(&out_full[s], kComputeWarps); mbar_init(&out_empty[s], 1); }
        mbar_fence_init();
    }
    __syncthreads();

    const int first = blockIdx.x, step = gridDim.x;
    const int n_my = first < a.n_tiles ? (a.n_tiles - first + step - 1) / step : 0;
    const size_t tile_in = (size_t)S * T * D, tile_out = (size_t)S * NB * D;

    if (warp == kComputeWarps) {
        // ---------------- copy thread ----------------
        if (lane == 0) {
            int issued = 0;
            for (; issued < NS && issued < n_my; ++issued) {
                const size_t tile = (size_t)first + (size_t)issued * step;
                mbar_arrive_expect_tx(&in_full[issued], in_bytes);
                bulk_g2s(smem + issued * in_stride, a.traj + tile * tile_in, in_bytes, &in_full[issued]);
            }
            for (int i = 0; i < n_my; ++i) {
                const int ob = i & 1;
                const size_t tile = (size_t)first + (size_t)i * step;
                mbar_wait(&out_full[ob], (i >> 1) & 1);
                if (a.tokens_out) bulk_s2g(a.tokens_out + tile * tile_out, out_base + ob * out_stride, tok_bytes);
                if (a.params_out)
                    bulk_s2g(a.params_out + tile * tile_out, out_base + ob * out_stride + tok_stride, par_bytes);
                bulk_commit();
                if (issued < n_my) {
                    const int s = issued % NS;
                    const size_t nt = (size_t)first + (size_t)issued * step;
                    mbar_wait(&in_empty[s], ((issued / NS) - 1) & 1);
                    mbar_arrive_expect_tx(&in_full[s], in_bytes);
                    bulk_g2s(smem + s * in_stride, a.traj + nt * tile_in, in_bytes, &in_full[s]);
                    ++issued;
                }
                bulk_wait_read<1>();                       // store of tile i-1 has left its staging buffer
                if (i >= 1) mbar_arrive(&out_empty[(i - 1) & 1]);
            }
            bulk_wait_all<0>();
        }
        return;
    }

    // ---------------- compute threads: one (trajectory, slot) column each ----------------
    const int nj = a.n_joint, ng = D - nj;
    const int ncols = S * D;
    const bool active = tid < ncols;
    int tl = 0, slot = 0;
    if (active) {
        if (tid < S * nj) { tl = tid / nj; slot = tid - tl * nj; }
        else { const int c = tid - S * nj; tl = c / ng; slot = nj + (c - tl * ng); }
    }
    const int dof = a.slot_to_dof[slot];
    const bool want_tok = a.tokens_out != nullptr, want_par = a.params_out != nullptr;
    float wmin[NB], wmax[NB], wscale[NB];
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        wmin[k] = want_tok ? a.w_min[slot * NB + k] : 0.0f;
        wmax[k] = want_tok ? a.w_max[slot * NB + k] : 0.0f;
        wscale[k] = quant_scale(wmin[k], wmax[k]);
    }

    for (int i = 0; i < n_my; ++i) {
        const int s = i % NS;
        mbar_wait(&in_full[s], (i / NS) & 1);
        float acc[NB];
        if (active) {
            const float* y = (const float*)(smem + s * in_stride) + tl * (T * D) + dof;
            if (slot < nj) fit_column<T, NB, false>(tab, y, D, acc);
            else fit_column<T, NB, true>(tab, y, D, acc);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&in_empty[s]);

        const int ob = i & 1;
        if (i >= 2) mbar_wait(&out_empty[ob], ((i >> 1) - 1) & 1);
        if (active) {
            unsigned char* o = out_base + ob * out_stride;
            if (want_tok) {
                long long* to = (long long*)o + tl * (NB * D) + slot;
#pragma unroll
                for (int k = 0; k < NB; ++k)
                    to[k * D] = quantize_one(acc[k], wmin[k], wmax[k], wscale[k], a.vm1) + a.offset;
            }
            if (want_par) {
                float* po = (float*)(o + tok_stride) + tl * (NB * D) + slot * NB;
                if (NB % 2 == 0) {
#pragma unroll
                    for (int k = 0; k < NB; k += 2) *(float2*)(po + k) = make_float2(acc[k], acc[k + 1]);
                } else {
#pragma unroll
                    for (int k = 0; k < NB; ++k) po[k] = acc[k];
                }
            }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&out_full[ob]);
    }
}

// One thread per (trajectory, slot) column; any geometry.  Accumulates in the same order
// (t ascending, fused multiply-add) as the fast kernel.
__global__ void __launch_bounds__(256)
encode_generic_kernel(const float* __restrict__ traj, long long ncol, int T, int D, int nb, int n_joint,
                      const int* __restrict__ slot_to_dof, const float* __restrict__ Pj,
                      const float* __restrict__ Pg, const float* __restrict__ w_min,
                      const float* __restrict__ w_max, float vm1, long long offset,
                      float* __restrict__ params_out, long long* __restrict__ tokens_out) {
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < ncol;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long b = idx / D;
        const int slot = (int)(idx - b * D);
        const int dof = slot_to_dof[slot];
        const float* P = slot < n_joint ? Pj : Pg;
        const float* y = traj + b * (long long)T * D + dof;
        for (int k = 0; k < nb; ++k) {
            float acc = 0.0f;
            const float* pk = P + (long long)k * T;
            for (int t = 0; t < T; ++t) acc = fmaf(__ldg(pk + t), __ldg(y + (long long)t * D), acc);
            if (params_out) params_out[b * (long long)D * nb + (long long)slot * nb + k] = acc;
            if (tokens_out) {
                const float lo = w_min[slot * nb + k], hi = w_max[slot * nb + k];
                tokens_out[b * (long long)D * nb + (long long)k * D + slot] =
                    quantize_one(acc, lo, hi, quant_scale(lo, hi), vm1) + offset;
            }
        }
    }
}

// params '(d t)' -> tokens '(t d)' (bit-exact quantiser) or normalised fp32 '(t d)'.
template <bool NORMALIZE>
__global__ void __launch_bounds__(256)
quantize_kernel(const float* __restrict__ params, long long n, int D, int nb, const float* __restrict__ w_min,
                const float* __restrict__ w_max, float vm1, long long offset, long long* __restrict__ tokens_out,
                float* __restrict__ norm_out) {
    const int row = D * nb;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long b = idx / row;
        const int r = (int)(idx - b * row);          // output position k*D + slot
        const int k = r / D, slot = r - k * D;
        const int c = slot * nb + k;
        const float w = params[b * row + c], lo = w_min[c], hi = w_max[c];
        if (NORMALIZE) {
            // beast/utils.py:32-34 with norm_min=-1, norm_max=1: (clip-lo)/scale * 2.0 + (-1.0)
            const float nrm = __fdiv_rn(__fsub_rn(clampf(w, lo, hi), lo), quant_scale(lo, hi));
            norm_out[idx] = __fadd_rn(__fmul_rn(nrm, 2.0f), -1.0f);
        } else {
            tokens_out[idx] = quantize_one(w, lo, hi, quant_scale(lo, hi), vm1) + offset;
        }
    }
}

static bool fast_disabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("BEAST_B200_DISABLE_FAST"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}

static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15u) == 0; }

static int grid_for(long long n, int block, int num_sms) {
    long long g = (n + block - 1) / block;
    const long long cap = (long long)num_sms * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

template <int T, int NB, int NS, int DT>
static int launch_fast(const Plan* p, const float* traj, long long n_tiles, int S, const float* w_min,
                       const float* w_max, long long offset, float* params_out, long long* tokens_out,
                       cudaStream_t st) {
    EncTables<T, NB> tab;
    constexpr int NBP = EncTables<T, NB>::NBP;
    for (int t = 0; t < T; ++t)
        for (int k = 0; k < NBP; ++k) {
            tab.pj[t * NBP + k] = k < NB ? p->proj_joint_h[k * T + t] : 0.0f;
            tab.pg[t * NBP + k] = (k < NB && p->proj_grip_h) ? p->proj_grip_h[k * T + t] : 0.0f;
        }
    EncArgs a;
    a.traj = traj; a.params_out = params_out; a.tokens_out = tokens_out;
    a.w_min = w_min; a.w_max = w_max; a.offset = offset; a.vm1 = (float)(p->V - 1);
    a.D = p->D; a.n_joint = p->n_joint; a.S = S; a.n_tiles = (int)n_tiles;
    for (int i = 0; i < BEAST_MAX_SLOTS; ++i) a.slot_to_dof[i] = i < p->D ? p->slot_to_dof[i] : 0;
    const uint32_t in_stride = round_up_128((uint32_t)S * T * p->D * 4u);
    const uint32_t out_stride = round_up_128((uint32_t)S * NB * p->D * 8u) + round_up_128((uint32_t)S * NB * p->D * 4u);
    const size_t smem = (size_t)NS * in_stride + 2 * (size_t)out_stride + (2 * NS + 4) * sizeof(uint64_t);
    if ((int)smem > p->max_smem_optin) return BEAST_E_UNSUPPORTED;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(encode_fast_kernel<T, NB, NS, DT>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, p->max_smem_optin);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int grid = (int)(n_tiles < p->num_sms ? n_tiles : p->num_sms);
    encode_fast_kernel<T, NB, NS, DT><<<grid, kThreads, smem, st>>>(tab, a);
    count_launch();
    BEAST_CHECK_LAUNCH();
    return BEAST_OK;
}

}  // namespace beast

using namespace beast;

extern "C" int beast_encode_f32(const beast_plan_t* plan, const float* traj, int64_t B, const float* w_min,
                                const float* w_max, int64_t offset, float* params_out, int64_t* tokens_out,
                                void* stream) {
    const Plan* p = (const Plan*)plan;
    if (!p || (B > 0 && !traj)) return BEAST_E_NULL;
    if (B < 0) return BEAST_E_SHAPE;
    if (!params_out && !tokens_out) return BEAST_E_NULL;
    if (tokens_out && (!w_min || !w_max)) return BEAST_E_NULL;
    if (B == 0) return BEAST_OK;
    if (((uintptr_t)traj & 3u) || ((uintptr_t)params_out & 3u) || ((uintptr_t)tokens_out & 7u)) return BEAST_E_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    const int T = p->T, D = p->D, nb = p->nb;
    long long done = 0;
    if (T == 50 && nb == 10 && !fast_disabled() && aligned16(traj) && (!params_out || aligned16(params_out)) &&
        (!tokens_out || aligned16(tokens_out))) {
        const int S = (kComputeThreads / D) & ~3;
        if (S >= 4 && B >= S) {
            const long long n_tiles = B / S;
            int rc;
            if (D == 14)
                rc = launch_fast<50, 10, 3, 14>(p, traj, n_tiles, S, w_min, w_max, offset, params_out,
                                                (long long*)tokens_out, st);
            else if (D == 7)
                rc = launch_fast<50, 10, 3, 7>(p, traj, n_tiles, S, w_min, w_max, offset, params_out,
                                               (long long*)tokens_out, st);
            else
                rc = launch_fast<50, 10, 3, 0>(p, traj, n_tiles, S, w_min, w_max, offset, params_out,
                                               (long long*)tokens_out, st);
            if (rc == BEAST_OK) done = n_tiles * S;
            else if (rc != BEAST_E_UNSUPPORTED) return rc;
        }
    }
    if (done < B) {
        const long long ncol = (B - done) * D;
        encode_generic_kernel<<<grid_for(ncol, 256, p->num_sms), 256, 0, st>>>(
            traj + done * (long long)T * D, ncol, T, D, nb, p->n_joint, p->slot_to_dof_d, p->proj_joint_d,
            p->proj_grip_d, w_min, w_max, (float)(p->V - 1), offset,
            params_out ? params_out + done * (long long)D * nb : nullptr,
            tokens_out ? (long long*)tokens_out + done * (long long)D * nb : nullptr);
        count_launch();
        BEAST_CHECK_LAUNCH();
    }
    return BEAST_OK;
}

extern "C" int beast_quantize_f32(const beast_plan_t* plan, const float* params, int64_t B, const float* w_min,
                                  const float* w_max, int64_t offset, int64_t* tokens_out, void* stream) {
    const Plan* p = (const Plan*)plan;
    if (!p || !w_min || !w_max || (B > 0 && (!params || !tokens_out))) return BEAST_E_NULL;
    if (B < 0) return BEAST_E_SHAPE;
    if (B == 0) return BEAST_OK;
    const long long n = (long long)B * p->D * p->nb;
    quantize_kernel<false><<<grid_for(n, 256, p->num_sms), 256, 0, (cudaStream_t)stream>>>(
        params, n, p->D, p->nb, w_min, w_max, (float)(p->V - 1), offset, (long long*)tokens_out, nullptr);
    count_launch();
    BEAST_CHECK_LAUNCH();
    return BEAST_OK;
}

extern "C" int beast_normalize_f32(const beast_plan_t* plan, const float* params, int64_t B, const float* w_min,
                                   const float* w_max, float* out, void* stream) {
    const Plan* p = (const Plan*)plan;
    if (!p || !w_min || !w_max || (B > 0 && (!params || !out))) return BEAST_E_NULL;
    if (B < 0) return BEAST_E_SHAPE;
    if (B == 0) return BEAST_OK;
    const long long n = (long long)B * p->D * p->nb;
    quantize_kernel<true><<<grid_for(n, 256, p->num_sms), 256, 0, (cudaStream_t)stream>>>(
        params, n, p->D, p->nb, w_min, w_max, 0.0f, 0, nullptr, out);
    count_launch();
    BEAST_CHECK_LAUNCH();
    return BEAST_OK;
}
