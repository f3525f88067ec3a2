"""FIGBPE — trainer for byte-pair encoding over discretised BEAST tokens, on the GPU.

Drop-in for the reference trainer (beast/beast_bpe_trainer.py:32-160): same class names, arguments,
return type and exceptions.  The reference builds `chr(bin - min)` strings in Python and hands them to
HF `tokenizers`' Rust BpeTrainer; here the MP tokens never leave the GPU and the merge loop runs on
libbeast_b200.so kernels (csrc/bpe.cu): GPT-2 pre-tokenisation + byte-level expansion, a dense
L2-resident pair histogram, arg-max with the trainer's tie-break, in-place merges with compaction.

Sharded training: when torch.distributed is initialised with more than one rank (one process per
GPU), every rank passes ITS shard of the sequences; min/max, the seen-byte set, the initial histogram
and the per-merge 4 x V count deltas are all-reduced (NCCL), so every rank replays the same merges and
ends with the same vocabulary.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib
from .beast_bspline_tokenizer import BEASTBsplineTokenizer
from .bpe_model import B2U, MAX_SHIFT, B200ByteLevelBPE, class_table_device, utf8_len

try:
    from tqdm.auto import tqdm
except Exception:  # pragma: no cover
    tqdm = None  # type: ignore[assignment]

ArrayLike = Union[Sequence[int], np.ndarray, torch.Tensor]


@dataclass
class FIGBPEState:
    tokenizer: B200ByteLevelBPE
    min_token: int
    max_token: int


class _Collective:
    """The three reductions the sharded trainer needs; a no-op for a single rank."""

    def __init__(self, group=None, enabled=None):
        import torch.distributed as dist
        self.dist = dist
        self.on = bool(enabled) if enabled is not None else (dist.is_available() and dist.is_initialized()
                                                             and dist.get_world_size(group) > 1)
        self.group = group

    def reduce_(self, t: torch.Tensor, op: str):
        if self.on:
            self.dist.all_reduce(t, op={"sum": self.dist.ReduceOp.SUM, "min": self.dist.ReduceOp.MIN,
                                        "max": self.dist.ReduceOp.MAX}[op], group=self.group)
        return t


def build_alphabet(min_token: int, max_token: int, seen_bytes: Sequence[int]):
    """A.3: vocabulary = chr(0..max-min) U byte-level characters seen, ids by sorted codepoint."""
    chars = set(range(max_token - min_token + 1)) | {B2U[b] for b in range(256) if seen_bytes[b]}
    tokens = [chr(c) for c in sorted(chars)]
    index = {t: i for i, t in enumerate(tokens)}
    byte_to_id = np.array([index.get(chr(B2U[b]), -1) for b in range(256)], dtype=np.int16)
    return tokens, byte_to_id


class GpuBpeEngine:
    """Device state of one shard: chunk-major symbols, lengths, replicated V x V histogram."""

    def __init__(self, bins: torch.Tensor, min_token: int, byte_to_id: np.ndarray, V: int, max_shift: int = 255):
        self.lib = _lib.load()
        self.dev = bins.device
        self.N, self.L = bins.shape
        self.V = V
        self.stride = max(self.N, 1)
        cap = utf8_len(max_shift) * self.L                  # byte-level symbols per sequence
        with torch.cuda.device(self.dev):
            # chunk-major corpus: [ceil(cap / 8) chunks, sequences, 8 symbols] uint16 (see csrc/bpe.cu)
            self.sym = torch.empty(((cap + 7) // 8, self.stride, 8), device=self.dev, dtype=torch.int16)
            self.len = torch.zeros(self.stride, device=self.dev, dtype=torch.int32)
            self.hist = torch.zeros((V, V), device=self.dev, dtype=torch.int32)
            self.delta = torch.zeros(4 * V, device=self.dev, dtype=torch.int32)
            self.result = torch.zeros(256, device=self.dev, dtype=torch.int64)   # arg-max scratch (per-block maxima)
            self.work = torch.zeros(4 + 2 * self.stride, device=self.dev, dtype=torch.int32)   # scan -> rewrite work list
            err = torch.zeros(1, device=self.dev, dtype=torch.int32)
            b2i = torch.from_numpy(byte_to_id).to(self.dev)
            st = _lib.stream_ptr(self.dev)
            _lib.check(self.lib.bpe_symbolize(_lib.ptr(bins), self.N, self.L, int(min_token), _lib.ptr(b2i),
                                              _lib.ptr(class_table_device(self.dev)), _lib.ptr(self.sym),
                                              _lib.ptr(self.len), self.stride, _lib.ptr(err), st), "bpe_symbolize")
            used = np.unique(byte_to_id[byte_to_id >= 0]).astype(np.int16)   # ids before any merge (byte-level symbols)
            used_d = torch.from_numpy(used).to(self.dev)
            n_ids = int(used[-1]) + 1 if used.size else 0
            _lib.check(self.lib.bpe_count_pairs(_lib.ptr(self.sym), _lib.ptr(self.len), self.N, self.stride, V,
                                                n_ids, _lib.ptr(used_d), int(used.size), _lib.ptr(self.hist), st),
                       "bpe_count_pairs")
            if int(err.item()):
                raise ValueError("discrete tokens outside the representable range after subtracting min_token")

    def argmax(self, n_active: int):
        with torch.cuda.device(self.dev):
            _lib.check(self.lib.bpe_argmax(_lib.ptr(self.hist), self.V, n_active, _lib.ptr(self.result),
                                           _lib.stream_ptr(self.dev)), "bpe_argmax")
        key = int(self.result[0].item()) & 0xFFFFFFFFFFFFFFFF
        if key == 0:
            return 0, -1, -1
        flat = 0xFFFFFFFF - (key & 0xFFFFFFFF)
        return key >> 32, flat // self.V, flat % self.V

    def merge(self, a: int, b: int, c: int):
        with torch.cuda.device(self.dev):
            _lib.check(self.lib.bpe_apply_merge(_lib.ptr(self.sym), _lib.ptr(self.len), self.N, self.stride, a, b, c,
                                                self.V, _lib.ptr(self.delta), _lib.ptr(self.work),
                                                _lib.stream_ptr(self.dev)), "bpe_apply_merge")

    def apply_delta(self, a: int, b: int, c: int):
        with torch.cuda.device(self.dev):
            _lib.check(self.lib.bpe_apply_delta(_lib.ptr(self.hist), _lib.ptr(self.delta), a, b, c, self.V,
                                                _lib.stream_ptr(self.dev)), "bpe_apply_delta")

    def run_fast(self, coll: "_Collective", n_tokens: int, vocab_size: int, min_frequency: int):
        """Sync-free merge loop: arg-max, stop rules, id assignment and the merge log all stay on the
        device (bpe_train_step); the host only enqueues iterations (plus the NCCL all-reduce of the
        delta block when sharded) and reads the log once.  Returns [(a, b, new_id, count)] assuming every merge creates a NEW token string; the
        caller verifies that and falls back to the exact host-driven loop otherwise."""
        max_merges = int(vocab_size) - int(n_tokens)
        if max_merges <= 0:
            return []
        dev = self.dev
        with torch.cuda.device(dev):
            ctl = torch.zeros(8, device=dev, dtype=torch.int32)
            ctl[4] = n_tokens
            log = torch.zeros(4 * max_merges, device=dev, dtype=torch.int32)
            self.result.zero_()
            # Pair signatures (one 4-byte column read per merge tells which sequences can hold the pair).  The
            # first merges touch most sequences anyway, so the signatures are built after kSigStart merges and
            # rebuilt from the current corpus every kSigRebuild merges (0.6 ms at 1.6 M sequences) to drop stale bits.
            sig = torch.empty((int(self.lib.bpe_signature_words()), self.stride), device=dev, dtype=torch.int32)
            sig_start, sig_rebuild = 64, 256

            def build_sig():
                _lib.check(self.lib.bpe_build_signatures(_lib.ptr(self.sym), _lib.ptr(self.len), self.N, self.stride,
                                                         _lib.ptr(sig), _lib.stream_ptr(dev)), "bpe_build_signatures")

            def step(phase, use_sig, iters=1):
                _lib.check(self.lib.bpe_train_step(_lib.ptr(self.sym), _lib.ptr(self.len), self.N, self.stride, self.V,
                                                   _lib.ptr(self.hist), _lib.ptr(self.delta), _lib.ptr(ctl),
                                                   _lib.ptr(log), _lib.ptr(self.result), _lib.ptr(self.work),
                                                   int(vocab_size), int(min_frequency), max_merges, phase,
                                                   _lib.ptr(sig) if use_sig else None, int(iters),
                                                   _lib.stream_ptr(dev)),
                           "bpe_train_step")

            # plain stream launches: ~60 us of host enqueue per merge, never a sync (capturing the
            # iteration into a CUDA graph costs more to instantiate than 1 700 replays save)
            self.work[:4].zero_()
            i = 0
            while i < max_merges:
                if i >= sig_start and (i - sig_start) % sig_rebuild == 0:
                    build_sig()
                if coll.on:
                    step(0, i >= sig_start)             # fold previous delta + arg-max, pick, scan, rewrite
                    coll.reduce_(self.delta, "sum")
                    i += 1
                else:                                   # unsharded: enqueue up to the next signature rebuild in one call
                    nxt = sig_start if i < sig_start else i + sig_rebuild - (i - sig_start) % sig_rebuild
                    n = min(nxt, max_merges) - i
                    step(0, i >= sig_start, n)
                    i += n
            ctl_h = ctl.cpu().tolist()
            n = ctl_h[5]
            return log[:4 * n].cpu().view(-1, 4).tolist()


def scan_bins_gpu(bins: torch.Tensor, coll: _Collective):
    """Global min / max token and the set of UTF-8 bytes seen (A.1, A.3)."""
    lib = _lib.load()
    dev = bins.device
    with torch.cuda.device(dev):
        st = _lib.stream_ptr(dev)
        mm = torch.tensor([torch.iinfo(torch.int64).max, torch.iinfo(torch.int64).min], device=dev, dtype=torch.int64)
        if bins.numel():
            _lib.check(lib.bpe_scan_bins(_lib.ptr(bins), bins.numel(), 0, _lib.ptr(mm), None, None, 0, st), "bpe_scan_bins")
        lo, hi = mm[:1].clone(), mm[1:].clone()
        coll.reduce_(lo, "min")
        coll.reduce_(hi, "max")
        min_token, max_token = int(lo.item()), int(hi.item())
        if max_token - min_token > MAX_SHIFT:
            raise ValueError("BPE over more than 55 296 distinct bin values is not representable: chr() of the shifted "
                             "bins would enter the surrogate range")
        seen = torch.zeros(256, device=dev, dtype=torch.int32)
        err = torch.zeros(1, device=dev, dtype=torch.int32)
        if bins.numel():
            _lib.check(lib.bpe_scan_bins(_lib.ptr(bins), bins.numel(), min_token, None, _lib.ptr(seen), _lib.ptr(err), 1,
                                         st), "bpe_scan_bins")
        coll.reduce_(seen, "max")
    return min_token, max_token, seen.cpu().numpy()


def train_bpe(bins: torch.Tensor, vocab_size: int, min_frequency: int = 2, *, engine_factory=None,
              scan=None, coll: Optional[_Collective] = None, show_progress: bool = False):
    """The merge loop (A.4) over this rank's shard `bins` [N, L] int64.  Returns
    (B200ByteLevelBPE, min_token, max_token).  `engine_factory` / `scan` are injection points for the
    CPU-only multi-process tests of the orchestration; production uses the GPU engine."""
    coll = coll or _Collective()
    scan = scan or scan_bins_gpu
    engine_factory = engine_factory or GpuBpeEngine
    min_token, max_token, seen = scan(bins, coll)
    tokens, byte_to_id = build_alphabet(min_token, max_token, seen)
    V = max(int(vocab_size), len(tokens))
    if V > 32767:
        raise NotImplementedError("bpe_vocab_size above 32767 is not supported")
    make_engine = lambda: (engine_factory(bins, min_token, byte_to_id, V, max_token - min_token)
                           if engine_factory is GpuBpeEngine else engine_factory(bins, min_token, byte_to_id, V))
    eng = make_engine()
    coll.reduce_(eng.hist, "sum")                       # replicated global histogram
    index = {t: i for i, t in enumerate(tokens)}
    merges: List[tuple] = []
    if hasattr(eng, "run_fast") and not show_progress:
        log = eng.run_fast(coll, len(tokens), vocab_size, min_frequency)
        fast_tokens, fast_index, ok = list(tokens), dict(index), True
        for a, b, c, _count in log:
            new = fast_tokens[a] + fast_tokens[b]
            if new in fast_index or c != len(fast_tokens):
                ok = False                              # an existing string must keep its id: replay exactly
                break
            fast_index[new] = c
            fast_tokens.append(new)
        if ok:
            return B200ByteLevelBPE(fast_tokens, [(a, b, c) for a, b, c, _ in log]), min_token, max_token
        eng = make_engine()
        coll.reduce_(eng.hist, "sum")
    bar = tqdm(total=max(vocab_size - len(tokens), 0), desc="BPE merges", leave=False) if (show_progress and tqdm) else None
    while len(tokens) < vocab_size:
        count, a, b = eng.argmax(len(tokens))
        if a < 0 or count < 1 or count < min_frequency:
            break
        new = tokens[a] + tokens[b]
        c = index.get(new)
        if c is None:                                   # an existing string keeps its id (merge still recorded)
            c = len(tokens)
            tokens.append(new)
            index[new] = c
            if bar is not None:
                bar.update(1)
        merges.append((a, b, c))
        eng.merge(a, b, c)
        coll.reduce_(eng.delta, "sum")                  # per-merge count deltas: 4 x V int32
        eng.apply_delta(a, b, c)
    if bar is not None:
        bar.close()
    return B200ByteLevelBPE(tokens, merges), min_token, max_token


def _flatten_to_numpy(sequence: ArrayLike) -> np.ndarray:
    if isinstance(sequence, torch.Tensor):
        array = sequence.detach().cpu().numpy()
    else:
        array = np.asarray(sequence)
    if array.ndim > 1:
        array = array.reshape(-1)
    return array.astype(np.int64)


class FIGBPE:
    """Trainer for Byte Pair Encoding over discretised BEAST tokens (reference :39-160)."""

    def __init__(self, vocab_size: int = 1024, *, min_frequency: int = 2, special_tokens: Optional[Sequence[str]] = None,
                 show_progress: bool = True, max_token_length: int = 10000, device=None, process_group=None) -> None:
        self.vocab_size = vocab_size
        self.min_frequency = min_frequency
        self.special_tokens = list(special_tokens or [])
        if self.special_tokens:
            raise NotImplementedError("special tokens are not supported by the B200 BPE trainer")
        self.show_progress = show_progress
        self.max_token_length = max_token_length
        self.device = device
        self.process_group = process_group
        self.tokenizer: Optional[B200ByteLevelBPE] = None
        self.min_token: Optional[int] = None
        self.max_token: Optional[int] = None

    def _device(self):
        if self.device is not None:
            return _lib.require_cuda(self.device)
        if not torch.cuda.is_available():
            raise _lib.BeastB200Error("no CUDA device available: the BPE trainer has no CPU fallback")
        return torch.device("cuda", torch.cuda.current_device())

    def fit_from_bins(self, bins: torch.Tensor) -> FIGBPEState:
        """bins [N, L] int64 on the GPU (this rank's shard when sharded)."""
        # process_group=False forces a local (unsharded) fit even under torch.distributed
        coll = _Collective(enabled=False) if self.process_group is False else _Collective(self.process_group)
        tok, mn, mx = train_bpe(bins.to(self._device(), torch.int64).contiguous(), self.vocab_size, self.min_frequency,
                                coll=coll, show_progress=self.show_progress)
        self.tokenizer, self.min_token, self.max_token = tok, mn, mx
        return FIGBPEState(tokenizer=tok, min_token=mn, max_token=mx)

    def fit_from_sequences(self, sequences: Iterable[ArrayLike]) -> FIGBPEState:
        processed: List[np.ndarray] = []
        for seq in sequences:
            arr = _flatten_to_numpy(seq)
            if arr.size == 0:
                continue
            processed.append(arr)
        if not processed:
            raise ValueError("No non-empty sequences provided for BPE training.")
        lengths = {a.size for a in processed}
        if len(lengths) != 1:
            raise NotImplementedError("the B200 BPE trainer expects equal-length sequences (BEAST tokens are)")
        return self.fit_from_bins(torch.from_numpy(np.stack(processed)))

    def fit_from_trajectories(self, tokenizer: BEASTBsplineTokenizer, trajectories: Iterable[Union[ArrayLike, dict]], *,
                              update_bounds: bool = False, batch_key: str = "actions",
                              max_sequences: Optional[int] = None) -> FIGBPEState:
        chunks: List[torch.Tensor] = []
        collected = 0
        encode_fn = getattr(tokenizer, "encode_to_mp_tokens", None)
        if encode_fn is None:
            encode_fn = tokenizer.encode
        progress_bar = None
        if self.show_progress and tqdm is not None:
            progress_bar = tqdm(total=max_sequences, desc="Collecting BEAST sequences for BPE", unit="seq", leave=False)
        for batch in trajectories:
            if isinstance(batch, dict):
                if batch_key not in batch:
                    raise KeyError(f"Batch dictionary is missing required key '{batch_key}'.")
                data = batch[batch_key]
            else:
                data = batch
            if not torch.is_tensor(data):
                data = torch.as_tensor(data)
            tokens, _ = encode_fn(data, update_bounds=update_bounds)       # stays on the GPU
            if max_sequences is not None and collected + tokens.shape[0] > max_sequences:
                tokens = tokens[: max_sequences - collected]
            chunks.append(tokens)
            collected += tokens.shape[0]
            if progress_bar is not None:
                progress_bar.update(tokens.shape[0])
            if max_sequences is not None and collected >= max_sequences:
                break
        if progress_bar is not None:
            progress_bar.close()
        if not chunks or collected == 0:
            raise ValueError("No non-empty sequences provided for BPE training.")
        return self.fit_from_bins(torch.cat(chunks, 0))

    def get_state(self) -> FIGBPEState:
        if self.tokenizer is None or self.min_token is None or self.max_token is None:
            raise RuntimeError("BPE tokenizer has not been trained yet.")
        return FIGBPEState(tokenizer=self.tokenizer, min_token=self.min_token, max_token=self.max_token)
