// Corpus layout shared by the BPE kernels (csrc/bpe.cu, csrc/bpe_dedup.cu).
#pragma once
#include "common.cuh"

namespace beast {

constexpr int kBpeBlock = 128;
constexpr uint16_t kWordStart = 0x8000u;
constexpr uint16_t kIdMask = 0x7fffu;
constexpr int kCoopWord = 24;            // pre-tokens from this many symbols on are merged by the whole warp, one rank per step
constexpr int kMaxWordLong = 8192;       // longest pre-token (in symbols) of the thread-per-sequence encode kernel

// Corpus layout ("chunk-major"): the symbols of sequence `seq` live in 16-byte chunks of 8,
// chunk c of all sequences contiguous:  sym[((p >> 3) * n_stride + seq) * 8 + (p & 7)].
// One thread owns one sequence; a warp reading chunk c of its 32 sequences touches 512 contiguous
// bytes with one 128-bit load per lane.  Slots past len[seq] in the last chunk hold 0xffff.
constexpr int kChunk = 8;
constexpr uint16_t kPad = 0xffffu;
__device__ __forceinline__ long long sym_index(int p, long long seq, long long n_stride) {
    return ((long long)(p >> 3) * n_stride + seq) * kChunk + (p & 7);
}

}  // namespace beast
