// Word de-duplication in front of the BPE merge loop.
//
// HF's BpeTrainer — what the reference trains with (beast/beast_bpe_trainer.py:61-74) — counts every distinct
// pre-token ("word") once and carries its multiplicity: pair counts are sums over distinct words x count, merges
// are applied to the distinct words only.  On repetitive corpora (real robot data: the same motion segments over
// and over) that is the largest algorithmic saving of the trainer.  Here:
//
//   table    every word of the symbolised corpus is hashed (64 bits over its symbol ids and length) into an
//            open-addressing table: pass 1 claims the slot (atomicCAS on the key; the winner becomes the slot's
//            representative) and counts the occurrences; pass 2 finds the slot again, the representative appends
//            (location, count) to the list of distinct words, every other word is compared with the
//            representative SYMBOL BY SYMBOL (a 64-bit hash collision is reported, never trusted: the caller
//            then trains on the plain corpus);
//   pack     the distinct words are ordered by count (host side: torch sort / scans over U elements) and packed,
//            words of EQUAL count together, into pseudo-sequences of the same chunk-major layout the merge loop
//            already walks; one int32 weight per pseudo-sequence = the count of its words.
//
// The merge loop (csrc/bpe.cu) is unchanged apart from multiplying its count updates by that weight, so the
// learned table is the same function of the pair counts — bit-identical to the plain run.
#include "bpe_common.cuh"

namespace beast {

// packed location of a word: sequence (32 bits) | first symbol (16) | length in symbols (16)
__device__ __forceinline__ unsigned long long word_loc(long long seq, int start, int len) {
    return ((unsigned long long)seq << 32) | ((unsigned long long)(unsigned int)start << 16) | (unsigned long long)(unsigned int)len;
}
__device__ __forceinline__ unsigned long long hash_step(unsigned long long h, unsigned int id) {
    h = (h ^ (unsigned long long)(id + 1u)) * 0x9E3779B97F4A7C15ull;
    return h ^ (h >> 29);
}
__device__ __forceinline__ unsigned long long hash_final(unsigned long long h, int len) {
    h ^= (unsigned long long)len * 0xC2B2AE3D27D4EB4Full;
    h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
    return h ? h : 1ull;                                     // 0 marks an empty slot
}

__device__ __forceinline__ unsigned int sym_at(const uint16_t* __restrict__ sym, int p, long long seq, long long n_stride) {
    return sym[sym_index(p, seq, n_stride)];
}

// One thread per sequence walks its chunks (128-bit loads) and hands every finished word to `emit`.
template <typename Emit>
__device__ __forceinline__ void for_each_word(const uint16_t* __restrict__ sym, const int* __restrict__ len, long long seq,
                                              long long n_stride, Emit emit) {
    const int n = len[seq];
    unsigned long long h = 0x243F6A8885A308D3ull;
    int start = 0;
    for (int c = 0; c * kChunk < n; ++c) {
        const int4 q = *(const int4*)(sym + ((long long)c * n_stride + seq) * kChunk);
        const unsigned int w[4] = {(unsigned int)q.x, (unsigned int)q.y, (unsigned int)q.z, (unsigned int)q.w};
#pragma unroll
        for (int j = 0; j < kChunk; ++j) {
            const int p = c * kChunk + j;
            if (p >= n) break;
            const unsigned int cur = (w[j >> 1] >> ((j & 1) * 16)) & 0xffffu;
            if ((cur & kWordStart) && p > 0) {
                emit(hash_final(h, p - start), start, p - start);
                h = 0x243F6A8885A308D3ull;
                start = p;
            }
            h = hash_step(h, cur & kIdMask);
        }
    }
    if (n > 0) emit(hash_final(h, n - start), start, n - start);
}

__global__ void __launch_bounds__(256)
bpe_word_total_kernel(const uint16_t* __restrict__ sym, const int* __restrict__ len, long long N, long long n_stride,
                      unsigned long long* __restrict__ totals) {
    unsigned long long words = 0, symbols = 0;
    for (long long seq = (long long)blockIdx.x * blockDim.x + threadIdx.x; seq < N;
         seq += (long long)gridDim.x * blockDim.x) {
        const int n = len[seq];
        symbols += (unsigned long long)n;
        for (int c = 0; c * kChunk < n; ++c) {
            const int4 q = *(const int4*)(sym + ((long long)c * n_stride + seq) * kChunk);
            const unsigned int w[4] = {(unsigned int)q.x, (unsigned int)q.y, (unsigned int)q.z, (unsigned int)q.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                // padding (0xffff) has the flag bit set too: count only positions below n
                const int p = c * kChunk + 2 * k;
                if (p < n && (w[k] & 0x8000u)) ++words;
                if (p + 1 < n && (w[k] & 0x80000000u)) ++words;
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        words += __shfl_xor_sync(0xffffffffu, words, o);
        symbols += __shfl_xor_sync(0xffffffffu, symbols, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (words) atomicAdd(&totals[0], words);
        if (symbols) atomicAdd(&totals[1], symbols);
    }
}

// pass 1 (EMIT = false): claim slots (atomicCAS on the key; the winner stores its location as the slot's
// representative) and count occurrences.  pass 2 (EMIT = true): every word finds its slot again; the
// representative appends (location, count) to the list of distinct words, every other word is compared with the
// representative symbol by symbol.  flags[0]: 1 = hash collision, 2 = internal error, 3 = table too small;
// flags[1] (pass 1) = slots claimed; flags[2] (pass 2) = distinct words emitted.
constexpr int kMaxProbes = 1 << 12;
template <bool EMIT>
__global__ void __launch_bounds__(256)
bpe_word_table_kernel(const uint16_t* __restrict__ sym, const int* __restrict__ len, long long N, long long n_stride,
                      unsigned long long* __restrict__ keys, unsigned long long* __restrict__ rep,
                      int* __restrict__ count, unsigned long long mask, int* __restrict__ flags,
                      unsigned long long* __restrict__ out_loc, int* __restrict__ out_cnt) {
    int claimed = 0;
    const unsigned int lane = threadIdx.x & 31u;
    for (long long seq = (long long)blockIdx.x * blockDim.x + threadIdx.x; seq < N;
         seq += (long long)gridDim.x * blockDim.x) {
        for_each_word(sym, len, seq, n_stride, [&](unsigned long long key, int start, int wl) {
            const unsigned long long loc = word_loc(seq, start, wl);
            unsigned long long slot = key & mask;
            bool won = false;
            int probes = 0;
            for (;;) {
                unsigned long long k = keys[slot];
                if (!EMIT && k == 0ull) {
                    const unsigned long long old = atomicCAS(&keys[slot], 0ull, key);
                    won = old == 0ull;
                    k = won ? key : old;
                }
                if (k == key) break;
                if (EMIT && k == 0ull) { atomicExch(&flags[0], 2); return; }     // cannot happen after pass 1
                if (++probes > kMaxProbes) { atomicExch(&flags[0], 3); return; } // table (nearly) full
                slot = (slot + 1) & mask;
            }
            if (!EMIT) {
                if (won) { rep[slot] = loc; ++claimed; }
                atomicAdd(&count[slot], 1);
                return;
            }
            const unsigned long long r = rep[slot];
            if (r == loc) {                                  // the representative lists the distinct word
                const unsigned int m = __activemask();       // opportunistic warp aggregation of the append
                const int leader = __ffs(m) - 1;
                int base = 0;
                if ((int)lane == leader) base = atomicAdd(&flags[2], __popc(m));
                base = __shfl_sync(m, base, leader);
                const int idx = base + __popc(m & ((1u << lane) - 1u));
                out_loc[idx] = loc;
                out_cnt[idx] = count[slot];
                return;
            }
            const long long rseq = (long long)(r >> 32);     // another word owns the slot: it must be the SAME word
            const int rstart = (int)((r >> 16) & 0xffffu), rlen = (int)(r & 0xffffu);
            bool same = rlen == wl;
            for (int q = 0; same && q < wl; ++q)
                same = (sym_at(sym, start + q, seq, n_stride) & kIdMask) == (sym_at(sym, rstart + q, rseq, n_stride) & kIdMask);
            if (!same) atomicExch(&flags[0], 1);
        });
    }
    if (!EMIT && claimed) atomicAdd(&flags[1], claimed);
}

// One thread per distinct word: copy its symbols from the corpus into its pseudo-sequence.
__global__ void __launch_bounds__(256)
bpe_word_pack_kernel(const uint16_t* __restrict__ src, long long src_stride, const unsigned long long* __restrict__ loc,
                     const int* __restrict__ dst_seq, const int* __restrict__ dst_off, long long U,
                     uint16_t* __restrict__ dst, long long dst_stride) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < U; i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long r = loc[i];
        const long long seq = (long long)(r >> 32);
        const int start = (int)((r >> 16) & 0xffffu), wl = (int)(r & 0xffffu);
        const long long ds = dst_seq[i];
        const int off = dst_off[i];
        for (int q = 0; q < wl; ++q) {
            uint16_t v = src[sym_index(start + q, seq, src_stride)];
            v = q == 0 ? (uint16_t)(v | kWordStart) : (uint16_t)(v & kIdMask);
            dst[sym_index(off + q, ds, dst_stride)] = v;
        }
    }
}

// pad the last chunk of every pseudo-sequence (the merge loop expects 0xffff behind len)
__global__ void __launch_bounds__(256)
bpe_word_pad_kernel(uint16_t* __restrict__ dst, const int* __restrict__ len, long long P, long long dst_stride) {
    for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < P; s += (long long)gridDim.x * blockDim.x)
        for (int q = len[s]; q & 7; ++q) dst[sym_index(q, s, dst_stride)] = kPad;
}

static int dedup_grid(long long n) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long g = (n + 255) / 256;
    if (g > (long long)sms * 16) g = (long long)sms * 16;
    return g < 1 ? 1 : (int)g;
}

}  // namespace beast

using namespace beast;

extern "C" int bpe_word_totals(const uint16_t* sym, const int32_t* len, int64_t N, int64_t n_stride, uint64_t* totals,
                               void* stream) {
    if (!totals) return BEAST_E_NULL;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(totals, 0, 2 * sizeof(uint64_t), st);
    if (e != cudaSuccess) return (int)e;
    if (N == 0) return BEAST_OK;
    if (!sym || !len) return BEAST_E_NULL;
    if (N < 0 || n_stride < N) return BEAST_E_SHAPE;
    if ((uintptr_t)sym & 15u) return BEAST_E_ALIGN;
    bpe_word_total_kernel<<<dedup_grid(N), 256, 0, st>>>(sym, len, N, n_stride, (unsigned long long*)totals);
    count_launch();
    BEAST_CHECK_LAUNCH();
    return BEAST_OK;
}

static int word_table_args(const uint16_t* sym, const int32_t* len, int64_t N, int64_t n_stride, const void* keys,
                           const void* rep, const void* count, int64_t table_size, const void* flags) {
    if (!sym || !len || !keys || !rep || !count || !flags) return BEAST_E_NULL;
    if (N < 0 || n_stride < N || N > 0x7fffffffLL || table_size < 2 || (table_size & (table_size - 1))) return BEAST_E_SHAPE;
    if ((uintptr_t)sym & 15u) return BEAST_E_ALIGN;
    return BEAST_OK;
}

extern "C" int bpe_word_insert(const uint16_t* sym, const int32_t* len, int64_t N, int64_t n_stride, uint64_t* keys,
                               uint64_t* rep, int32_t* count, int64_t table_size, int32_t* flags, void* stream) {
    if (N == 0) return BEAST_OK;
    if (int rc = word_table_args(sym, len, N, n_stride, keys, rep, count, table_size, flags)) return rc;
    bpe_word_table_kernel<false><<<dedup_grid(N), 256, 0, (cudaStream_t)stream>>>(
        sym, len, N, n_stride, (unsigned long long*)keys, (unsigned long long*)rep, count,
        (unsigned long long)table_size - 1ull, flags, nullptr, nullptr);
    count_launch();
    BEAST_CHECK_LAUNCH();
    return BEAST_OK;
}

extern "C" int bpe_word_emit(const uint16_t* sym, const int32_t* len, int64_t N, int64_t n_stride, const uint64_t* keys,
                             const uint64_t* rep, const int32_t* count, int64_t table_size, int32_t* flags,
                             uint64_t* out_loc, int32_t* out_cnt, void* stream) {
    if (N == 0) return BEAST_OK;
    if (int rc = word_table_args(sym, len, N, n_stride, keys, rep, count, table_size, flags)) return rc;
    if (!out_loc || !out_cnt) return BEAST_E_NULL;
    bpe_word_table_kernel<true><<<dedup_grid(N), 256, 0, (cudaStream_t)stream>>>(
        sym, len, N, n_stride, (unsigned long long*)keys, (unsigned long long*)rep, (int*)count,
        (unsigned long long)table_size - 1ull, flags, (unsigned long long*)out_loc, out_cnt);
    count_launch();
    BEAST_CHECK_LAUNCH();
    return BEAST_OK;
}

extern "C" int bpe_word_pack(const uint16_t* src, int64_t src_stride, const uint64_t* loc, const int32_t* dst_seq,
                             const int32_t* dst_off, int64_t U, uint16_t* dst, const int32_t* dst_len, int64_t P,
                             int64_t dst_stride, void* stream) {
    if (U == 0 || P == 0) return BEAST_OK;
    if (!src || !loc || !dst_seq || !dst_off || !dst || !dst_len) return BEAST_E_NULL;
    if (U < 0 || P < 0 || dst_stride < P) return BEAST_E_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    bpe_word_pack_kernel<<<dedup_grid(U), 256, 0, st>>>(src, src_stride, (const unsigned long long*)loc, dst_seq, dst_off,
                                                        U, dst, dst_stride);
    bpe_word_pad_kernel<<<dedup_grid(P), 256, 0, st>>>(dst, dst_len, P, dst_stride);
    count_launch(2);
    BEAST_CHECK_LAUNCH();
    return BEAST_OK;
}
