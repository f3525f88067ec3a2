"""Per-iteration GPU time of the sync-free BPE loop (events), 1.6 M sequences."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beast_tokenizer_b200 import BEASTBsplineTokenizer, _lib
from beast_tokenizer_b200.beast_bpe_trainer import GpuBpeEngine, _Collective, scan_bins_gpu, build_alphabet
from beast_tokenizer_b200.synth import SyntheticLoader, synth_device
dev = torch.device("cuda", 0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_600_000
tok = BEASTBsplineTokenizer(num_dof=14, num_basis=10, seq_len=50, vocab_size=256, gripper_zero_order=True, gripper_indices=[6, 13], device="cuda")
tok.fit_parameters(SyntheticLoader(100, 32, 50, 14, seed0=1), verbose=False)
bins = torch.cat([tok.encode(synth_device(N // 25, 50, 14, 1000 + c, dev))[0] for c in range(25)])
coll = _Collective(enabled=False)
mn, mx, seen = scan_bins_gpu(bins, coll)
tokens, b2i = build_alphabet(mn, mx, seen)
torch.cuda.synchronize(); t0 = time.perf_counter()
eng = GpuBpeEngine(bins, mn, b2i, 2048)
torch.cuda.synchronize(); print(f"symbolize+count: {time.perf_counter()-t0:.3f} s")
lib = eng.lib
ctl = torch.zeros(8, device=dev, dtype=torch.int32); ctl[4] = len(tokens)
mm = 2048 - len(tokens)
log = torch.zeros(4 * mm, device=dev, dtype=torch.int32); eng.result.zero_()
SIG = None
USE_SIG = os.environ.get("NOSIG") != "1"
sig_t = torch.empty((int(lib.bpe_signature_words()), eng.stride), device=dev, dtype=torch.int32)
def build_sig():
    _lib.check(lib.bpe_build_signatures(_lib.ptr(eng.sym), _lib.ptr(eng.len), eng.N, eng.stride, _lib.ptr(sig_t), _lib.stream_ptr(dev)), "sig")
def step(ph):
    _lib.check(lib.bpe_train_step(_lib.ptr(eng.sym), _lib.ptr(eng.len), eng.N, eng.stride, eng.V, _lib.ptr(eng.hist), _lib.ptr(eng.delta),
                                  _lib.ptr(ctl), _lib.ptr(log), _lib.ptr(eng.result), _lib.ptr(eng.work), 2048, 2, mm, ph, SIG, 1, _lib.stream_ptr(dev)), "s")
evs = [torch.cuda.Event(enable_timing=True) for _ in range(mm + 1)]
evs[0].record()
for i in range(mm):
    if USE_SIG and i >= 64 and (i - 64) % 256 == 0:
        build_sig(); SIG = _lib.ptr(sig_t)
    step(0); step(1); evs[i + 1].record()
torch.cuda.synchronize()
ts = [evs[i].elapsed_time(evs[i + 1]) for i in range(mm)]
lens = eng.len.sum().item()
print(f"total {sum(ts):.1f} ms; live symbols at end {lens/1e6:.1f} M")
for i in (0, 1, 10, 50, 100, 127, 128, 150, 200, 300, 500, 800, 1200, 1700):
    if i < mm: print(i, f"{ts[i]*1e3:.0f} us", log[4*i:4*i+4].tolist())
