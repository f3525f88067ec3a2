"""Fill ratio of the pair signatures on the synthetic corpus (fresh build)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beast_tokenizer_b200 import BEASTBsplineTokenizer, _lib
from beast_tokenizer_b200.beast_bpe_trainer import GpuBpeEngine, _Collective, scan_bins_gpu, build_alphabet
from beast_tokenizer_b200.synth import SyntheticLoader, synth_device
dev = torch.device("cuda", 0)
N = 200_000
tok = BEASTBsplineTokenizer(num_dof=14, num_basis=10, seq_len=50, vocab_size=256, gripper_zero_order=True, gripper_indices=[6, 13], device="cuda")
tok.fit_parameters(SyntheticLoader(100, 32, 50, 14, seed0=1), verbose=False)
bins = tok.encode(synth_device(N, 50, 14, 1000, dev))[0]
coll = _Collective(enabled=False)
mn, mx, seen = scan_bins_gpu(bins, coll)
tokens, b2i = build_alphabet(mn, mx, seen)
eng = GpuBpeEngine(bins, mn, b2i, 2048)
lib = eng.lib
W = int(lib.bpe_signature_words())
sig = torch.empty((W, eng.stride), device=dev, dtype=torch.int32)
_lib.check(lib.bpe_build_signatures(_lib.ptr(eng.sym), _lib.ptr(eng.len), eng.N, eng.stride, _lib.ptr(sig), _lib.stream_ptr(dev)), "sig")
torch.cuda.synchronize()
bits = torch.zeros(eng.stride, device=dev, dtype=torch.int64)
s64 = sig.to(torch.int64) & 0xFFFFFFFF
for b in range(32):
    bits += ((s64 >> b) & 1).sum(0)
print(f"mean bits set per sequence {bits.float().mean().item():.1f} of {W * 32} ({bits.float().mean().item() / (W * 32):.3f}); "
      f"mean symbols {eng.len.float().mean().item():.1f}")
col = (s64 != 0).float().mean(1)
print("fraction of sequences with any bit in a word column: min %.3f max %.3f" % (col.min().item(), col.max().item()))
perbit = torch.stack([((s64 >> b) & 1).float().mean(1) for b in range(32)], 1).flatten()
print("per-bit-position fill: min %.4f median %.4f max %.4f" % (perbit.min().item(), perbit.median().item(), perbit.max().item()))
