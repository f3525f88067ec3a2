// Tiled K1 / K3 for ANY geometry (seq_len, num_basis, degree, DoF) on the tokenizer's own times — the path of the
// reference's shipped configuration (train.sh / train/train_beast.py:34-36: 50 basis functions, degree 0, 1000 bins,
// actions [10, 32]: 1 600 tokens per trajectory, 1 280 B read and 19 200 B written per encode) and of every shape the
// T = 50 / nb = 10 bulk-copy kernels (spline_encode.cu, spline_decode.cu) do not cover.
//
// Sums run over the non-zero BAND of each projector / basis row (plan.cu): a degree-0 projector row touches the few
// samples of its interval (none at all for 40 of 50 basis functions when nb > T), a degree-p basis row p + 1
// coefficients; the skipped terms are exact zeros, t / k ascending with fused multiply-adds as in the other kernels, so
// the results are bit-identical to them.
//
// Encode: a persistent CTA walks tiles of S trajectories.  The tile's samples arrive by one bulk copy (double
// buffered: the next tile is in flight while this one is computed); tokens and coefficients are STAGED in shared memory
// in their final global layout and leave by one bulk copy each (a tile's tokens / coefficients are one contiguous run
// of global memory).  The staging rows are templates: a coefficient whose projector row is empty is exactly 0 for every
// trajectory and its token a per-column constant, both written ONCE per CTA; per tile only the positions of the plan's
// enc_list are recomputed (320 of 1 600 for the shipped shape), so the instruction stream per trajectory is the
// useful fit + quantiser and nothing else — the copy engine moves the bytes.
//
// Decode: only the tokens whose coefficient some basis row reads are loaded at all (plan's dec_list: with 50 degree-0
// basis functions over 10 samples that is one token in five, in 256-byte runs), dequantised into shared memory; then
// one output sample per thread in output order (contiguous fp32 stores).
#include "common.cuh"

namespace beast {

constexpr int kTiledThreads = 512;

struct EncTiledArgs {
    const float* traj;
    long long B;
    int T, D, nb, n_joint;
    const int* slot_to_dof;
    const float* Pj;
    const float* Pg;
    const int* bands;
    const int* list;
    int n_enc;
    const float* w_min;
    const float* w_max;
    float vm1;
    long long offset;
    float* params_out;
    long long* tokens_out;
    float* bmin;
    float* bmax;
    int S, G;                 // trajectories per tile; trajectory groups per list entry (work item = entry x group)
    uint32_t y_bytes, tok_off, par_off, q_off, list_off, band_off, dof_off, bar_off, p_off;
    int in_bulk, tok_bulk, par_bulk;   // base pointer 16-byte aligned and full tiles a multiple of 16 bytes
    int p_smem;                        // projectors staged in shared memory as [joint | gripper][t][k]
};

// R = trajectories per pass of a work item: 4 when the sums are long (one projector load then feeds four sums), 1 when
// they have one or two terms (the shipped shape: nothing to share, and the compute phase sits on the tile's critical path)
template <int R>
__global__ void __launch_bounds__(kTiledThreads, 2)
encode_tiled_kernel(const __grid_constant__ EncTiledArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int T = a.T, D = a.D, nb = a.nb, S = a.S, n_enc = a.n_enc;
    const int row_in = T * D, row_out = D * nb;
    float* const ybuf0 = (float*)smem;                          // [2][S][T][D] samples, double buffered
    float* const ybuf1 = (float*)(smem + a.y_bytes);
    long long* const tok_s = (long long*)(smem + a.tok_off);    // [S][D*nb] tokens, global layout '(t d)'
    float* const par_s = (float*)(smem + a.par_off);            // [S][D*nb] coefficients, global layout '(d t)'
    float* const qtab = (float*)(smem + a.q_off);               // [n_enc][4] quantiser constants, or [2][D*nb] min / max
    float* const s_mn = qtab;
    float* const s_mx = qtab + row_out;
    int* const lst = (int*)(smem + a.list_off);                 // [n_enc] k | slot << 16 | gripper << 31
    int* const s_band = (int*)(smem + a.band_off);              // [2][2*nb]
    int* const s_dof = (int*)(smem + a.dof_off);                // [D]
    uint64_t* const bar = (uint64_t*)(smem + a.bar_off);        // [2]
    float* const P_s = (float*)(smem + a.p_off);                // [2][T][nb] projectors, k minor (threads of a warp differ in k)
    const bool want_tok = a.tokens_out != nullptr, want_par = a.params_out != nullptr, want_mm = a.bmin != nullptr;
    const bool use_par = want_par || want_mm;
    const int tid = threadIdx.x;

    for (int i = tid; i < n_enc; i += kTiledThreads) {
        const int e = a.list[i];
        lst[i] = e;
        if (want_tok) {
            const int k = e & 0xffff, slot = (e >> 16) & 0x7fff;
            QuantCol qc;
            qc.init(a.w_min[slot * nb + k], a.w_max[slot * nb + k]);
            qtab[4 * i] = qc.lo; qtab[4 * i + 1] = qc.hi; qtab[4 * i + 2] = qc.scale; qtab[4 * i + 3] = qc.rcp;
        }
    }
    for (int i = tid; i < 4 * nb; i += kTiledThreads) s_band[i] = a.bands[i];
    for (int i = tid; i < D; i += kTiledThreads) s_dof[i] = a.slot_to_dof[i];
    if (a.p_smem)
        for (int i = tid; i < T * nb; i += kTiledThreads) {
            const int t = i / nb, k = i - t * nb;
            P_s[i] = a.Pj[k * T + t];
            if (a.Pg) P_s[T * nb + i] = a.Pg[k * T + t];
        }
    if (want_tok)                                               // templates: the token of a zero coefficient everywhere
        for (int r = tid; r < row_out; r += kTiledThreads) {
            const int k = r / D, slot = r - k * D;
            const float lo0 = a.w_min[slot * nb + k], hi0 = a.w_max[slot * nb + k];
            const long long z = quantize_one(0.0f, lo0, hi0, quant_scale(lo0, hi0), a.vm1) + a.offset;
            for (int tr = 0; tr < S; ++tr) tok_s[(size_t)tr * row_out + r] = z;
        }
    if (use_par)
        for (int i = tid; i < S * row_out; i += kTiledThreads) par_s[i] = 0.0f;
    if (want_mm)                                                // a column outside the list holds 0 for every trajectory
        for (int c = tid; c < row_out; c += kTiledThreads) { s_mn[c] = 0.0f; s_mx[c] = 0.0f; }
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (want_mm)
        for (int i = tid; i < n_enc; i += kTiledThreads) {
            const int e = lst[i];
            const int c = ((e >> 16) & 0x7fff) * nb + (e & 0xffff);
            s_mn[c] = __int_as_float(0x7f800000);
            s_mx[c] = __int_as_float(0xff800000);
        }

    const long long n_tiles = (a.B + S - 1) / S;
    auto rows_of = [&](long long tile) { const long long rem = a.B - tile * S; return (int)(rem < S ? rem : S); };
    auto bulk_in = [&](int ns) { return a.in_bulk && ((ns * row_in) & 3) == 0; };
    auto issue = [&](long long tile, int buf) {                 // thread 0: the tile's samples, one bulk copy
        const int ns = rows_of(tile);
        if (!bulk_in(ns)) return;
        const uint32_t bytes = (uint32_t)ns * row_in * 4u;
        mbar_arrive_expect_tx(&bar[buf], bytes);
        bulk_g2s(buf ? ybuf1 : ybuf0, a.traj + tile * S * row_in, bytes, &bar[buf]);
    };
    long long tile = blockIdx.x;
    if (tid == 0 && tile < n_tiles) issue(tile, 0);
    const int items = n_enc * a.G;
    // work item j = tid + m * blockDim -> (trajectory group g, list entry i) = (j / n_enc, j % n_enc), tracked by increments:
    // two divisions per tile instead of one per item
    const int ne = n_enc > 0 ? n_enc : 1;
    for (int it = 0; tile < n_tiles; ++it, tile += gridDim.x) {
        const int buf = it & 1;
        float* const y = buf ? ybuf1 : ybuf0;
        const long long b0 = tile * S;
        const int ns = rows_of(tile);
        // the other buffer was last read before the barrier that ended the previous tile's compute phase
        if (tid == 0 && tile + gridDim.x < n_tiles) issue(tile + gridDim.x, buf ^ 1);
        if (bulk_in(ns)) {
            // one thread polls (511 spinning threads would take issue slots from the CTAs that are computing); the block
            // barrier below orders everyone after its acquire.  Only a LAST tile can be loaded by hand: parity = use count
            if (tid == 0) mbar_wait(&bar[buf], (uint32_t)(it >> 1) & 1u);
        } else {
            const float* src = a.traj + b0 * row_in;
            for (int i = tid; i < ns * row_in; i += kTiledThreads) y[i] = __ldcs(src + i);
        }
        if (tid == 0) bulk_wait_read<0>();                      // the previous tile's stores have read the staging rows
        __syncthreads();
        const int g_step = kTiledThreads / ne, i_step = kTiledThreads - g_step * n_enc;
        int g = tid / ne, i = tid - g * n_enc;
        for (int j = tid; j < items; j += kTiledThreads, g += g_step, i += i_step) {
            if (i >= n_enc) { i -= n_enc; ++g; }
            const int e = lst[i];
            const int k = e & 0xffff, slot = (e >> 16) & 0x7fff;
            const bool grip = e < 0;
            // generic pointer + stride over t: the shared-memory copy [t][k] or the plan's global table [k][t]
            const float* Pk = a.p_smem ? P_s + (grip ? T * nb : 0) + k : (grip ? a.Pg : a.Pj) + k * T;
            const int pst = a.p_smem ? nb : 1;
            const int t0 = s_band[(grip ? 2 * nb : 0) + 2 * k], t1 = s_band[(grip ? 2 * nb : 0) + 2 * k + 1];
            const int r = k * D + slot, c = slot * nb + k;
            const float* col0 = y + s_dof[slot];
            QuantCol qc;
            if (want_tok) {
                const float4 q = *(const float4*)(qtab + 4 * i);
                qc.lo = q.x; qc.hi = q.y; qc.scale = q.z; qc.rcp = q.w;
            }
            float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
            auto emit = [&](int tr, float acc) {
                if (use_par) par_s[tr * row_out + c] = acc;
                if (want_tok) tok_s[tr * row_out + r] = quantize_col(acc, qc, a.vm1) + a.offset;
                mn = fminf(mn, acc); mx = fmaxf(mx, acc);
            };
            if constexpr (R == 4) {
                // four trajectories per pass: one projector load feeds four sums (each sum keeps its own t-ascending order)
                for (int tr = g; tr < ns; tr += 4 * a.G) {
                    const int tr1 = tr + a.G, tr2 = tr + 2 * a.G, tr3 = tr + 3 * a.G, last = ns - 1;
                    const float* c0 = col0 + tr * row_in;
                    const float* c1 = col0 + (tr1 < last ? tr1 : last) * row_in;
                    const float* c2 = col0 + (tr2 < last ? tr2 : last) * row_in;
                    const float* c3 = col0 + (tr3 < last ? tr3 : last) * row_in;
                    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
                    for (int t = t0; t < t1; ++t) {
                        const float pv = Pk[t * pst];
                        const int o = t * D;
                        a0 = fmaf(pv, c0[o], a0); a1 = fmaf(pv, c1[o], a1); a2 = fmaf(pv, c2[o], a2); a3 = fmaf(pv, c3[o], a3);
                    }
                    emit(tr, a0);
                    if (tr1 < ns) emit(tr1, a1);
                    if (tr2 < ns) emit(tr2, a2);
                    if (tr3 < ns) emit(tr3, a3);
                }
            } else {
                for (int tr = g; tr < ns; tr += a.G) {
                    const float* col = col0 + tr * row_in;
                    float acc = 0.0f;
                    for (int t = t0; t < t1; ++t) acc = fmaf(__ldg(Pk + t), col[t * D], acc);   // <1> never stages the projector
                    emit(tr, acc);
                }
            }
            if (want_mm && mn <= mx) { atomic_min_f32(&s_mn[c], mn); atomic_max_f32(&s_mx[c], mx); }
        }
        fence_async_smem();                                     // staging rows -> visible to the copy engine
        __syncthreads();
        if (want_tok) {
            long long* dst = a.tokens_out + b0 * row_out;
            if (a.tok_bulk && ((ns * row_out) & 1) == 0) {
                if (tid == 0) bulk_s2g(dst, tok_s, (uint32_t)ns * row_out * 8u);
            } else {
                for (int i = tid; i < ns * row_out; i += kTiledThreads) __stcs(dst + i, tok_s[i]);
            }
        }
        if (want_par) {
            float* dst = a.params_out + b0 * row_out;
            if (a.par_bulk && ((ns * row_out) & 3) == 0) {
                if (tid == 0) bulk_s2g(dst, par_s, (uint32_t)ns * row_out * 4u);
            } else {
                for (int i = tid; i < ns * row_out; i += kTiledThreads) __stcs(dst + i, par_s[i]);
            }
        }
        if (tid == 0) bulk_commit();
        // hand-written stores of this tile read the staging rows; the next tile's writes come after its barrier
    }
    if (tid == 0) bulk_wait_all<0>();
    if (want_mm) {
        __syncthreads();
        for (int c = tid; c < row_out; c += kTiledThreads)
            if (s_mn[c] <= s_mx[c]) { atomic_min_f32(a.bmin + c, s_mn[c]); atomic_max_f32(a.bmax + c, s_mx[c]); }
    }
}

struct DecTiledArgs {
    const long long* tokens;
    const float* params;
    long long B;
    int T, D, nb, n_joint;
    const int* slot_to_dof;
    const float* phi_j;
    const float* phi_g;
    const int* bands;
    const int* list;
    int n_dec;
    const float* w_min;
    const float* w_max;
    float vm1;
    long long offset;
    const float* init_p;
    float* out;
    int S, G1, G2;            // trajectories per tile; trajectory groups per token entry / per output sample
    int phi_smem;             // basis rows staged in shared memory as [joint | gripper][t][nb | 1]
};

// decode: tile of token rows (or coefficient rows) -> trajectories.
// Two 512-thread CTAs per SM, 64 registers per thread: with three CTAs (40 registers) ptxas could not keep the four token
// loads of a pass in flight together — it consumed the first before issuing the third — and the kernel waited on memory
// latency (shipped shape 87 us against 51 us).
template <bool FROM_TOKENS>
__global__ void __launch_bounds__(kTiledThreads, 2)
decode_tiled_kernel(const __grid_constant__ DecTiledArgs a) {
    extern __shared__ __align__(16) float smem_f[];
    const int T = a.T, D = a.D, nb = a.nb, S = a.S, n_dec = a.n_dec;
    const int row_in = D * nb, row_out = T * D;
    float* c_s = smem_f;                                        // [S][nb][D] coefficients, token order (slot minor)
    float* lohi = c_s + (((size_t)S * row_in + 3) & ~(size_t)3);    // [n_dec][2] bounds of the listed tokens
    int* lst = (int*)(lohi + (size_t)2 * n_dec);                // [n_dec] k | slot << 16 | gripper << 31
    int* tab_out = lst + n_dec;                                 // [T*D]  t | slot << 16 | gripper << 31, output order [t][dof]
    int* s_band = tab_out + row_out;                            // [2][2*T]
    int* s_dof = s_band + 4 * T;                                // [D]
    float* phi_s = (float*)(s_dof + D);                         // [2][T][pitch] basis rows, odd pitch
    const int pitch = nb | 1;
    const int tid = threadIdx.x;
    const float rcp_vm1 = __frcp_rn(a.vm1);                     // exact invariant division by V - 1 (common.cuh)
    for (int i = tid; i < n_dec; i += kTiledThreads) {
        const int e = a.list[i];
        lst[i] = e;
        if (FROM_TOKENS) {
            const int k = e & 0xffff, slot = (e >> 16) & 0x7fff;
            lohi[2 * i] = a.w_min[slot * nb + k];
            lohi[2 * i + 1] = a.w_max[slot * nb + k];
        }
    }
    for (int i = tid; i < D; i += kTiledThreads) s_dof[i] = a.slot_to_dof[i];
    for (int i = tid; i < 4 * T; i += kTiledThreads) s_band[i] = a.bands[4 * nb + i];
    if (a.phi_smem)
        for (int i = tid; i < T * nb; i += kTiledThreads) {
            const int t = i / nb, k = i - t * nb;
            phi_s[t * pitch + k] = a.phi_j[i];
            if (a.phi_g) phi_s[(T + t) * pitch + k] = a.phi_g[i];
        }
    __syncthreads();
    for (int q = tid; q < row_out; q += kTiledThreads) {
        const int t = q / D, dof = q - t * D;
        int slot = 0;
        for (int sidx = 0; sidx < D; ++sidx) if (s_dof[sidx] == dof) slot = sidx;
        tab_out[q] = t | (slot << 16) | (slot < a.n_joint ? 0 : (int)0x80000000u);
    }
    const long long n_tiles = (a.B + S - 1) / S;
    const int items1 = n_dec * a.G1, items2 = row_out * a.G2;
    // work item j = tid + m * blockDim -> (trajectory group, entry), tracked by increments inside a phase (two divisions
    // per phase and tile instead of one per item; recomputed per phase so that they do not occupy registers across it)
    const int nd = n_dec > 0 ? n_dec : 1;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long b0 = tile * S;
        const int ns = (int)((a.B - b0) < S ? (a.B - b0) : S);
        __syncthreads();                                        // tables ready / previous tile's coefficients consumed
        // tile bases: everything below indexes them with 32-bit offsets (a tile is at most S * row < 2^20 elements)
        const long long* tok_tile = FROM_TOKENS ? a.tokens + b0 * row_in : nullptr;
        const float* par_tile = FROM_TOKENS ? nullptr : a.params + b0 * row_in;
        const float* ip_tile = a.init_p ? a.init_p + b0 * D : nullptr;
        float* out_tile = a.out + b0 * row_out;
        const int g1_step = kTiledThreads / nd, i1_step = kTiledThreads - g1_step * n_dec;
        int g = tid / nd, i = tid - g * n_dec;
        for (int j = tid; j < items1; j += kTiledThreads, g += g1_step, i += i1_step) {     // phase 1: the listed tokens -> coefficients
            if (i >= n_dec) { i -= n_dec; ++g; }
            const int e = lst[i];
            const int k = e & 0xffff, slot = (e >> 16) & 0x7fff;
            const int r = k * D + slot;
            float lo = 0.0f, hi = 0.0f;
            if (FROM_TOKENS) { lo = lohi[2 * i]; hi = lohi[2 * i + 1]; }
            const int cpar = slot * nb + k;
            const int stride = a.G1 * row_in;
            // four trajectories per pass, all four loads issued before the first store (the token pointer is a plain
            // pointer: the compiler keeps a global load behind an earlier shared-memory store, which left ONE load in
            // flight per thread)
            for (int tr = g, off = g * row_in; tr < ns; tr += 4 * a.G1, off += 4 * stride) {
                const int o1 = off + stride, o2 = o1 + stride, o3 = o2 + stride;
                const bool h1 = tr + a.G1 < ns, h2 = tr + 2 * a.G1 < ns, h3 = tr + 3 * a.G1 < ns;
                float v0, v1 = 0.0f, v2 = 0.0f, v3 = 0.0f;
                if (FROM_TOKENS) {
                    const long long q0 = __ldcs(tok_tile + off + r);
                    const long long q1 = h1 ? __ldcs(tok_tile + o1 + r) : 0;
                    const long long q2 = h2 ? __ldcs(tok_tile + o2 + r) : 0;
                    const long long q3 = h3 ? __ldcs(tok_tile + o3 + r) : 0;
                    v0 = dequantize_fast(q0 - a.offset, lo, hi, a.vm1, rcp_vm1);
                    v1 = dequantize_fast(q1 - a.offset, lo, hi, a.vm1, rcp_vm1);
                    v2 = dequantize_fast(q2 - a.offset, lo, hi, a.vm1, rcp_vm1);
                    v3 = dequantize_fast(q3 - a.offset, lo, hi, a.vm1, rcp_vm1);
                } else {
                    v0 = par_tile[off + cpar];
                    if (h1) v1 = par_tile[o1 + cpar];
                    if (h2) v2 = par_tile[o2 + cpar];
                    if (h3) v3 = par_tile[o3 + cpar];
                }
                c_s[off + r] = v0;
                if (h1) c_s[o1 + r] = v1;
                if (h2) c_s[o2 + r] = v2;
                if (h3) c_s[o3 + r] = v3;
            }
            if (k == 0 && ip_tile != nullptr && slot < a.n_joint) {   // init_p pins the first joint coefficient: replaces it
                const int dof = s_dof[slot];
                for (int tr = g; tr < ns; tr += a.G1) c_s[tr * row_in + r] = ip_tile[tr * D + dof];
            }
        }
        __syncthreads();
        const int g2_step = kTiledThreads / row_out, q_step = kTiledThreads - g2_step * row_out;
        int g2 = tid / row_out, q = tid - g2 * row_out;
        for (int j = tid; j < items2; j += kTiledThreads, g2 += g2_step, q += q_step) {     // phase 2: one output sample per thread, output order
            if (q >= row_out) { q -= row_out; ++g2; }
            const int g = g2;
            const int e = tab_out[q];
            const int t = e & 0xffff, slot = (e >> 16) & 0x7fff;
            const bool grip = e < 0;
            const float* row = a.phi_smem ? phi_s + (size_t)((grip ? T : 0) + t) * pitch : (grip ? a.phi_g : a.phi_j) + (size_t)t * nb;
            const int k0 = s_band[(grip ? 2 * T : 0) + 2 * t], k1 = s_band[(grip ? 2 * T : 0) + 2 * t + 1];
            // four trajectories per pass: one basis load feeds four sums (each keeps its own k-ascending order)
            for (int tr = g; tr < ns; tr += 4 * a.G2) {
                const int tr1 = tr + a.G2, tr2 = tr + 2 * a.G2, tr3 = tr + 3 * a.G2, last = ns - 1;
                const float* c0 = c_s + tr * row_in + slot;
                const float* c1 = c_s + (tr1 < last ? tr1 : last) * row_in + slot;
                const float* c2 = c_s + (tr2 < last ? tr2 : last) * row_in + slot;
                const float* c3 = c_s + (tr3 < last ? tr3 : last) * row_in + slot;
                float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
                for (int k = k0; k < k1; ++k) {
                    const float pv = row[k];
                    const int o = k * D;
                    a0 = fmaf(pv, c0[o], a0); a1 = fmaf(pv, c1[o], a1); a2 = fmaf(pv, c2[o], a2); a3 = fmaf(pv, c3[o], a3);
                }
                __stcs(out_tile + tr * row_out + q, a0);
                if (tr1 < ns) __stcs(out_tile + tr1 * row_out + q, a1);
                if (tr2 < ns) __stcs(out_tile + tr2 * row_out + q, a2);
                if (tr3 < ns) __stcs(out_tile + tr3 * row_out + q, a3);
            }
        }
    }
}

static inline size_t up16(size_t v) { return (v + 15) & ~(size_t)15; }
static inline int groups_for(int entries, int S) {
    int g = entries > 0 ? (kTiledThreads + entries - 1) / entries : 1;
    if (g > S) g = S;
    return g < 1 ? 1 : g;
}

int launch_encode_tiled(const Plan* p, const float* traj, long long B, const float* w_min, const float* w_max,
                        long long offset, float* params_out, long long* tokens_out, float* bmin, float* bmax,
                        cudaStream_t st) {
    const int T = p->T, D = p->D, nb = p->nb;
    if (bmin && tokens_out) return BEAST_E_UNSUPPORTED;      // quantiser constants and min / max share one region
    if (nb > 0xffff || D > 0x7fff || p->V > 0x7fffffff || !p->bands_d || !p->enc_list_d) return BEAST_E_UNSUPPORTED;
    const size_t row_in = (size_t)T * D, row_out = (size_t)D * nb;
    const bool use_par = params_out || bmin;
    const size_t per_traj = 2 * row_in * 4 + (tokens_out ? row_out * 8 : 0) + (use_par ? row_out * 4 : 0);
    const size_t q_bytes = up16(bmin ? 2 * row_out * 4 : (tokens_out ? (size_t)p->n_enc * 16 : 0));
    const size_t p_bytes = up16((size_t)2 * T * nb * 4);
    // sums of one or two terms read every projector entry once per pass: no reuse to stage for (and the shipped shape
    // keeps a third trajectory per tile instead); tables beyond 32 KB are read through L1 from the plan's global copy
    const bool p_smem = p_bytes <= 32 * 1024 && p->enc_band_max > 2;
    const size_t extra = q_bytes + up16((size_t)p->n_enc * 4) + up16((size_t)4 * nb * 4) + up16((size_t)D * 4) + 16 + 4 * 16 +
                         (p_smem ? p_bytes : 0);
    // two CTAs per SM, 64 registers per thread (three CTAs at 40 registers measured 5 % slower on the shipped shape, 25 %
    // on long sums: ptxas had no room to overlap the shared-memory loads of a pass)
    size_t budget = 108 * 1024;
    if (per_traj + extra > budget) budget = (size_t)p->max_smem_optin;
    if (per_traj + extra > budget) return BEAST_E_UNSUPPORTED;
    int S = (int)((budget - extra) / per_traj);
    if (S > 64) S = 64;
    if (S >= 4) S &= ~3;                                      // full tiles a multiple of 16 bytes whatever the row sizes
    EncTiledArgs a;
    a.traj = traj; a.B = B; a.T = T; a.D = D; a.nb = nb; a.n_joint = p->n_joint;
    a.slot_to_dof = p->slot_to_dof_d; a.Pj = p->proj_joint_d; a.Pg = p->proj_grip_d; a.bands = p->bands_d;
    a.list = p->enc_list_d; a.n_enc = p->n_enc;
    a.w_min = w_min; a.w_max = w_max; a.vm1 = (float)(p->V - 1); a.offset = offset;
    a.params_out = params_out; a.tokens_out = tokens_out; a.bmin = bmin; a.bmax = bmax;
    a.S = S; a.G = groups_for(p->n_enc, S);
    a.y_bytes = (uint32_t)up16((size_t)S * row_in * 4);
    a.tok_off = 2 * a.y_bytes;
    a.par_off = a.tok_off + (uint32_t)(tokens_out ? up16((size_t)S * row_out * 8) : 0);
    a.q_off = a.par_off + (uint32_t)(use_par ? up16((size_t)S * row_out * 4) : 0);
    a.list_off = a.q_off + (uint32_t)q_bytes;
    a.band_off = a.list_off + (uint32_t)up16((size_t)p->n_enc * 4);
    a.dof_off = a.band_off + (uint32_t)up16((size_t)4 * nb * 4);
    a.bar_off = a.dof_off + (uint32_t)up16((size_t)D * 4);
    a.p_off = a.bar_off + 16;
    a.p_smem = p_smem ? 1 : 0;
    const size_t smem = (size_t)a.p_off + (p_smem ? p_bytes : 0);
    if (smem > (size_t)p->max_smem_optin) return BEAST_E_UNSUPPORTED;
    a.in_bulk = (((uintptr_t)traj & 15u) == 0 && ((size_t)S * row_in) % 4 == 0) ? 1 : 0;
    a.tok_bulk = (tokens_out && ((uintptr_t)tokens_out & 15u) == 0 && ((size_t)S * row_out) % 2 == 0) ? 1 : 0;
    a.par_bulk = (params_out && ((uintptr_t)params_out & 15u) == 0 && ((size_t)S * row_out) % 4 == 0) ? 1 : 0;
    const bool wide = p->enc_band_max > 2;
    static size_t granted1[kMaxDevices] = {}, granted4[kMaxDevices] = {};
    if (int rc = wide ? opt_in_smem(encode_tiled_kernel<4>, smem, granted4) : opt_in_smem(encode_tiled_kernel<1>, smem, granted1)) return rc;
    const long long n_tiles = (B + S - 1) / S;
    long long per_sm = (long long)(220 * 1024) / (long long)(smem + 1024);
    if (per_sm > 2048 / kTiledThreads) per_sm = 2048 / kTiledThreads;
    if (per_sm < 1) per_sm = 1;
    long long grid = n_tiles < (long long)p->num_sms * per_sm ? n_tiles : (long long)p->num_sms * per_sm;
    if (wide) encode_tiled_kernel<4><<<(unsigned)grid, kTiledThreads, smem, st>>>(a);
    else encode_tiled_kernel<1><<<(unsigned)grid, kTiledThreads, smem, st>>>(a);
    count_launch();
    BEAST_CHECK_LAUNCH();
    return BEAST_OK;
}

int launch_decode_tiled(const Plan* p, const long long* tokens, const float* params, long long B, const float* w_min,
                        const float* w_max, long long offset, const float* init_p, float* out, cudaStream_t st) {
    const int T = p->T, D = p->D, nb = p->nb;
    if (p->nc != nb || !p->bands_d || !p->dec_list_d) return BEAST_E_UNSUPPORTED;   // pinned control points: generic kernel
    if (nb > 0xffff || T > 0xffff || D > 0x7fff) return BEAST_E_UNSUPPORTED;
    const size_t row_in = (size_t)D * nb, row_out = (size_t)T * D;
    const size_t per_traj = row_in * sizeof(float);
    const size_t phi_bytes = (size_t)2 * T * (nb | 1) * sizeof(float);
    const bool phi_smem = phi_bytes <= 32 * 1024 && p->dec_band_max > 2;   // as in launch_encode_tiled
    const size_t extra = ((size_t)3 * p->n_dec + row_out + (size_t)4 * T + D) * sizeof(float) + 64 + (phi_smem ? phi_bytes : 0);
    size_t budget = 108 * 1024;                               // two CTAs per SM
    if (per_traj + extra > budget) budget = (size_t)p->max_smem_optin;
    if (per_traj + extra > budget) return BEAST_E_UNSUPPORTED;
    int S = (int)((budget - extra) / per_traj);
    if (S > 64) S = 64;
    const size_t smem = (size_t)S * per_traj + extra;
    static size_t granted_t[kMaxDevices] = {}, granted_p[kMaxDevices] = {};
    if (int rc = tokens ? opt_in_smem(decode_tiled_kernel<true>, smem, granted_t) : opt_in_smem(decode_tiled_kernel<false>, smem, granted_p))
        return rc;
    DecTiledArgs a;
    a.tokens = tokens; a.params = params; a.B = B; a.T = T; a.D = D; a.nb = nb; a.n_joint = p->n_joint;
    a.slot_to_dof = p->slot_to_dof_d; a.phi_j = p->phi_joint_d; a.phi_g = p->phi_grip_d; a.bands = p->bands_d;
    a.list = p->dec_list_d; a.n_dec = p->n_dec;
    a.w_min = w_min; a.w_max = w_max; a.vm1 = tokens ? (float)(p->V - 1) : 0.0f; a.offset = tokens ? offset : 0;
    a.init_p = init_p; a.out = out;
    a.S = S; a.G1 = groups_for(p->n_dec, S); a.G2 = groups_for((int)row_out, S);
    a.phi_smem = phi_smem ? 1 : 0;
    const long long n_tiles = (B + S - 1) / S;
    long long per_sm = (long long)(220 * 1024) / (long long)(smem + 1024);
    if (per_sm > 2048 / kTiledThreads) per_sm = 2048 / kTiledThreads;
    if (per_sm < 1) per_sm = 1;
    long long grid = n_tiles < (long long)p->num_sms * per_sm ? n_tiles : (long long)p->num_sms * per_sm;
    if (tokens) decode_tiled_kernel<true><<<(unsigned)grid, kTiledThreads, smem, st>>>(a);
    else decode_tiled_kernel<false><<<(unsigned)grid, kTiledThreads, smem, st>>>(a);
    count_launch();
    BEAST_CHECK_LAUNCH();
    return BEAST_OK;
}

}  // namespace beast
