"""Per-iteration GPU time of the sync-free BPE loop (events around every iteration), 1.6 M sequences by default."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beast_tokenizer_b200 import BEASTBsplineTokenizer
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
eng = GpuBpeEngine(bins, mn, b2i, 2048, mx - mn, None, "auto", seen)
torch.cuda.synchronize(); print(f"symbolize+count: {time.perf_counter()-t0:.3f} s")
run = eng.start_run(len(tokens), 2048, 2)
mm = run.max_merges
evs = [torch.cuda.Event(enable_timing=True) for _ in range(mm + 1)]
evs[0].record()
for i in range(mm):
    run.enqueue(limit=1); evs[i + 1].record()
torch.cuda.synchronize()
ts = [evs[i].elapsed_time(evs[i + 1]) for i in range(mm)]
log = run.finish()
print(f"total {sum(ts):.1f} ms; live symbols at end {eng.len.sum().item()/1e6:.1f} M; merges {len(log)}")
for i in (0, 1, 10, 50, 100, 127, 128, 150, 200, 300, 500, 800, 1200, 1700):
    if i < len(log): print(i, f"{ts[i]*1e3:.0f} us", log[i])
edges = [0, 16, 64, 128, 256, 512, 1024, len(ts)]
for a, b in zip(edges[:-1], edges[1:]):
    print(f"merges [{a:4d}, {b:4d}): {sum(ts[a:b]):7.2f} ms  = {1e3 * sum(ts[a:b]) / max(b - a, 1):6.1f} us per merge")
