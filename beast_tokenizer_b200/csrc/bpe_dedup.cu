// Word de-duplication in front of the BPE merge loop.
//
// HF's BpeTrainer — what the reference trains with (beast/beast_bpe_trainer.py:61-74) — counts every distinct
// pre-token ("word") once and carries its multiplicity: pair counts are sums over distinct words x count, merges
// are applied to the distinct words only.  On repetitive corpora (real robot data: the same motion segments over
// and over) that is the largest algorithmic saving of the trainer.  Here:
//
//   list     every word of the symbolised corpus is hashed (64 bits over its symbol ids and length) into a flat
//            list of (hash, location), one thread per sequence;
//   table    one thread per WORD probes an open-addressing table of 32-byte slots: an empty slot is claimed with
//            one 128-bit compare-and-swap of (hash, location) — the winner is the slot's representative — the
//            occurrences are counted, and every other word is compared with the representative SYMBOL BY SYMBOL
//            (a 64-bit hash collision is reported, never trusted: the caller then trains on the plain corpus);
//            the distinct words are then read off the table, slot by slot;
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

// One thread per sequence walks its chunks (128-bit loads) and hands every finished word to `emit`, together
// with its first six symbol ids (+1, 16 bits each: two in head, four in tail; 0 = no symbol).
constexpr int kInline = 6;
template <typename Emit>
__device__ __forceinline__ void for_each_word(const uint16_t* __restrict__ sym, const int* __restrict__ len, long long seq,
                                              long long n_stride, Emit emit) {
    const int n = len[seq];
    unsigned long long h = 0x243F6A8885A308D3ull, tail = 0ull;
    unsigned int head = 0u;
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
                emit(hash_final(h, p - start), start, p - start, head, tail);
                h = 0x243F6A8885A308D3ull;
                head = 0u;
                tail = 0ull;
                start = p;
            }
            const unsigned int id = cur & kIdMask;
            const int k = p - start;
            if (k < 2) head |= (id + 1u) << (16 * k);
            else if (k < kInline) tail |= (unsigned long long)(id + 1u) << (16 * (k - 2));
            h = hash_step(h, id);
        }
    }
    if (n > 0) emit(hash_final(h, n - start), start, n - start, head, tail);
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

// 128-bit accesses to a table slot's (key, representative) pair
__device__ __forceinline__ ulonglong2 slot_load(const unsigned long long* p) {
    ulonglong2 v;
    asm volatile("{\n\t.reg .b128 t;\n\tld.relaxed.gpu.global.b128 t, [%2];\n\tmov.b128 {%0, %1}, t;\n\t}"
                 : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ ulonglong2 slot_claim(unsigned long long* p, unsigned long long key, unsigned long long loc) {
    ulonglong2 old;
    asm volatile("{\n\t.reg .b128 c, v, o;\n\tmov.b128 c, {%3, %3};\n\tmov.b128 v, {%4, %5};\n\t"
                 "atom.relaxed.gpu.global.cas.b128 o, [%2], c, v;\n\tmov.b128 {%0, %1}, o;\n\t}"
                 : "=l"(old.x), "=l"(old.y) : "l"(p), "l"(0ull), "l"(key), "l"(loc) : "memory");
    return old;
}

// The flat word list: one thread per sequence counts its words (flag bits of its chunks), reserves a range of the
// list (warp-aggregated), then walks the sequence again and writes (hash, location) per word.  The order of the
// list is irrelevant (an entry is 32 bytes: hash, location, the word's first six symbols); what it buys is that
// the table pass below runs one THREAD PER WORD — every lane has its own
// independent probe in flight instead of a warp following one sequence's word boundaries (the per-sequence
// version of the table pass waited on memory 128 cycles per issued instruction).
__global__ void __launch_bounds__(256)
bpe_word_list_kernel(const uint16_t* __restrict__ sym, const int* __restrict__ len, long long N, long long n_stride,
                     ulonglong2* __restrict__ words, long long W, unsigned long long* __restrict__ cursor,
                     int* __restrict__ flags) {
    const unsigned int lane = threadIdx.x & 31u;
    const long long span = ((N + 31) / 32) * 32;             // whole warps: every lane takes part in the shuffles
    for (long long seq = (long long)blockIdx.x * blockDim.x + threadIdx.x; seq < span;
         seq += (long long)gridDim.x * blockDim.x) {
        const int n = seq < N ? len[seq] : 0;
        int nw = 0;
        for (int c = 0; c * kChunk < n; ++c) {
            const int4 q = *(const int4*)(sym + ((long long)c * n_stride + seq) * kChunk);
            const unsigned int w[4] = {(unsigned int)q.x, (unsigned int)q.y, (unsigned int)q.z, (unsigned int)q.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int p = c * kChunk + 2 * k;
                if (p < n && (w[k] & 0x8000u)) ++nw;
                if (p + 1 < n && (w[k] & 0x80000000u)) ++nw;
            }
        }
        if (n > 0 && !(sym[sym_index(0, seq, n_stride)] & kWordStart)) ++nw;   // a first symbol without the flag still opens a word
        int incl = nw;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o) incl += t;
        }
        unsigned long long base = 0;
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (lane == 31u && total) base = atomicAdd(cursor, (unsigned long long)total);
        base = __shfl_sync(0xffffffffu, base, 31) + (unsigned long long)(incl - nw);
        if (nw == 0) continue;
        if (base + (unsigned long long)nw > (unsigned long long)W) { atomicExch(&flags[0], 2); continue; }
        int k = 0;
        for_each_word(sym, len, seq, n_stride, [&](unsigned long long key, int start, int wl, unsigned int head,
                                                   unsigned long long tail) {
            if (k < nw) {
                words[2 * (base + k)] = make_ulonglong2(key, word_loc(seq, start, wl));
                words[2 * (base + k) + 1] = make_ulonglong2((unsigned long long)head << 32, tail);
            }
            ++k;
        });
        if (k != nw) atomicExch(&flags[0], 2);
    }
}

// Table pass, one thread per word.  Slot = 32 bytes = one memory sector: key, representative location, count and
// the representative's first six symbols.  An empty slot is claimed with ONE 128-bit compare-and-swap of
// (key, location), so whoever finds the key also finds its representative; the winner then stores its symbols
// next to the count.  Every other word is compared with the representative SYMBOL BY SYMBOL (a 64-bit hash
// collision is reported, never trusted): against the slot's copy when the word has at most six symbols and the
// copy agrees — symbols are stored +1, so a copy not written yet (or half written) can only fail to agree —
// and against the representative's place in the corpus otherwise.  A short word therefore costs one random
// memory sector.
// Word frequencies are very skewed (the most frequent word of the benchmark corpus is 7 % of all words): with
// every access going to L2 the slice that owns that one sector serialises millions of requests.  So (1) the
// slot is first read through L1 (key, location and symbols are write-once: a stale copy can only look emptier,
// in which case the slot is read again from L2), and (2) the counts are summed per block in a small shared-memory
// table first (direct-mapped on the slot index; what does not find a place goes to memory directly).
// flags[0]: 1 = hash collision, 2 = internal error, 3 = table too small; flags[1] = slots claimed.
constexpr int kMaxProbes = 1 << 12;
constexpr int kCountCache = 1024;
__global__ void __launch_bounds__(256)
bpe_word_table_kernel(const uint16_t* __restrict__ sym, long long n_stride, const ulonglong2* __restrict__ words,
                      long long W, unsigned long long* table, unsigned long long mask, int* __restrict__ flags) {
    __shared__ unsigned long long s_slot[kCountCache];
    __shared__ int s_cnt[kCountCache];
    for (int i = threadIdx.x; i < kCountCache; i += blockDim.x) { s_slot[i] = ~0ull; s_cnt[i] = 0; }
    __syncthreads();
    int claimed = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < W; i += (long long)gridDim.x * blockDim.x) {
        const ulonglong2 w = __ldcs(&words[2 * i]), ws = __ldcs(&words[2 * i + 1]);
        const unsigned long long key = w.x, loc = w.y;
        unsigned long long slot = key & mask;
        ulonglong2 e, es;
        int probes = 0;
        bool ok = true, won = false;
        for (;;) {
            unsigned long long* p = table + slot * 4;
            e = __ldca((const ulonglong2*)p);
            es = __ldca((const ulonglong2*)p + 1);
            if (e.x == 0ull) {                               // empty, or a stale line: ask L2
                e = slot_load(p);
                es = slot_load(p + 2);
                if (e.x == 0ull) {
                    e = slot_claim(p, key, loc);
                    es = make_ulonglong2(0ull, 0ull);        // whoever won has not published its symbols yet
                    if (e.x == 0ull) { e = make_ulonglong2(key, loc); won = true; }
                }
            }
            if (e.x == key) break;
            if (++probes > kMaxProbes) { atomicExch(&flags[0], 3); ok = false; break; }   // table (nearly) full
            slot = (slot + 1) & mask;
        }
        if (!ok) continue;
        {
            const unsigned int h = (unsigned int)((slot * 0x9E3779B97F4A7C15ull) >> 54);     // 10 bits
            const unsigned long long old = atomicCAS(&s_slot[h], ~0ull, slot);
            if (old == ~0ull || old == slot) atomicAdd(&s_cnt[h], 1);
            else atomicAdd((int*)(table + slot * 4 + 2), 1);
        }
        if (won) {
            ++claimed;
            ((unsigned int*)(table + slot * 4 + 2))[1] = (unsigned int)(ws.x >> 32);
            table[slot * 4 + 3] = ws.y;
            continue;
        }
        const int wl = (int)(loc & 0xffffu), rlen = (int)(e.y & 0xffffu);
        if (rlen == wl && wl <= kInline && (es.x >> 32) == (ws.x >> 32) && es.y == ws.y) continue;
        const long long seq = (long long)(loc >> 32), rseq = (long long)(e.y >> 32);
        const int start = (int)((loc >> 16) & 0xffffu), rstart = (int)((e.y >> 16) & 0xffffu);
        unsigned int diff = rlen != wl;
        for (int q = 0; q < wl && !diff; q += 4) {           // four symbols' loads in flight, no exit inside
#pragma unroll
            for (int t = 0; t < 4; ++t)
                if (q + t < wl)
                    diff |= (sym_at(sym, start + q + t, seq, n_stride) ^ sym_at(sym, rstart + q + t, rseq, n_stride)) & kIdMask;
        }
        if (diff) atomicExch(&flags[0], 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kCountCache; i += blockDim.x)
        if (s_cnt[i]) atomicAdd((int*)(table + s_slot[i] * 4 + 2), s_cnt[i]);
    for (int o = 16; o > 0; o >>= 1) claimed += __shfl_xor_sync(0xffffffffu, claimed, o);
    if ((threadIdx.x & 31) == 0 && claimed) atomicAdd(&flags[1], claimed);
}

// The distinct words, read off the table: (location, count) of every claimed slot; flags[2] = entries written.
__global__ void __launch_bounds__(256)
bpe_word_emit_kernel(const unsigned long long* __restrict__ table, long long table_size, int* __restrict__ flags,
                     unsigned long long* __restrict__ out_loc, int* __restrict__ out_cnt, long long capacity) {
    const unsigned int lane = threadIdx.x & 31u;
    const long long span = ((table_size + 31) / 32) * 32;
    for (long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x; slot < span;
         slot += (long long)gridDim.x * blockDim.x) {
        ulonglong2 e = make_ulonglong2(0ull, 0ull);
        if (slot < table_size) e = *(const ulonglong2*)(table + slot * 4);
        const bool used = e.x != 0ull;
        const unsigned int m = __ballot_sync(0xffffffffu, used);
        if (!m) continue;
        int base = 0;
        if (lane == 0u) base = atomicAdd(&flags[2], __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (!used) continue;
        const long long idx = base + __popc(m & ((1u << lane) - 1u));
        if (idx >= capacity) { atomicExch(&flags[0], 2); continue; }
        out_loc[idx] = e.y;
        out_cnt[idx] = (int)(table[slot * 4 + 2] & 0xffffffffull);
    }
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

extern "C" int bpe_word_list(const uint16_t* sym, const int32_t* len, int64_t N, int64_t n_stride, uint64_t* words,
                             int64_t W, uint64_t* cursor, int32_t* flags, void* stream) {
    if (!cursor || !flags) return BEAST_E_NULL;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(cursor, 0, sizeof(uint64_t), st);
    if (e != cudaSuccess) return (int)e;
    if (N == 0) return BEAST_OK;
    if (!sym || !len || !words) return BEAST_E_NULL;
    if (N < 0 || n_stride < N || N > 0x7fffffffLL || W < 0) return BEAST_E_SHAPE;
    if (((uintptr_t)sym | (uintptr_t)words) & 15u) return BEAST_E_ALIGN;
    bpe_word_list_kernel<<<dedup_grid(N), 256, 0, st>>>(sym, len, N, n_stride, (ulonglong2*)words, W,
                                                         (unsigned long long*)cursor, flags);
    count_launch();
    BEAST_CHECK_LAUNCH();
    return BEAST_OK;
}

extern "C" int bpe_word_insert(const uint16_t* sym, int64_t n_stride, const uint64_t* words, int64_t W, uint64_t* table,
                               int64_t table_size, int32_t* flags, void* stream) {
    if (W == 0) return BEAST_OK;
    if (!sym || !words || !table || !flags) return BEAST_E_NULL;
    if (W < 0 || n_stride < 1 || table_size < 2 || (table_size & (table_size - 1))) return BEAST_E_SHAPE;
    if (((uintptr_t)words & 15u) || ((uintptr_t)table & 31u)) return BEAST_E_ALIGN;
    bpe_word_table_kernel<<<dedup_grid(W), 256, 0, (cudaStream_t)stream>>>(
        sym, n_stride, (const ulonglong2*)words, W, (unsigned long long*)table, (unsigned long long)table_size - 1ull, flags);
    count_launch();
    BEAST_CHECK_LAUNCH();
    return BEAST_OK;
}

extern "C" int bpe_word_emit(const uint64_t* table, int64_t table_size, int32_t* flags, uint64_t* out_loc,
                             int32_t* out_cnt, int64_t capacity, void* stream) {
    if (!table || !flags) return BEAST_E_NULL;
    if (table_size < 2 || (table_size & (table_size - 1)) || capacity < 0) return BEAST_E_SHAPE;
    if (capacity > 0 && (!out_loc || !out_cnt)) return BEAST_E_NULL;
    if ((uintptr_t)table & 31u) return BEAST_E_ALIGN;
    bpe_word_emit_kernel<<<dedup_grid(table_size), 256, 0, (cudaStream_t)stream>>>(
        (const unsigned long long*)table, table_size, flags, (unsigned long long*)out_loc, out_cnt, capacity);
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
