"""torchrun --nproc-per-node N scripts/bpe_shard_times.py — where does a SHARDED FIGBPE fit spend its time?
Synchronised + barriered phase timings on rank 0 (1.6 M sequences over N ranks), best of 3 runs."""
import os, sys, time, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beast_tokenizer_b200 import BEASTBsplineTokenizer
from beast_tokenizer_b200 import beast_bpe_trainer as T
from beast_tokenizer_b200.synth import SyntheticLoader, synth_device
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
tok = BEASTBsplineTokenizer(num_dof=14, num_basis=10, seq_len=50, vocab_size=256, gripper_zero_order=True, gripper_indices=[6, 13], device=f"cuda:{local}")
tok.fit_parameters(SyntheticLoader(100, 32, 50, 14, seed0=1), verbose=False)
CH = 200
bins = torch.cat([tok.encode(synth_device(8000, 50, 14, 1000 + c, dev))[0] for c in range(CH) if c % world == rank])
def mark():
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize(); return time.perf_counter()
best = None
for rep in range(4):
    coll = T._Collective()
    t = [mark()]
    mn, mx, seen = T.scan_bins_gpu(bins, coll); t.append(mark())
    tokens, b2i = T.build_alphabet(mn, mx, seen)
    eng = T.GpuBpeEngine(bins, mn, b2i, 2048, mx - mn); t.append(mark())
    coll.reduce_(eng.hist, "sum"); t.append(mark())
    block = T._PeerBlock.get(dev, 2048, coll); peers = block.begin_run(coll); t.append(mark())
    run = eng.start_run(len(tokens), 2048, 2, peers, block.delta_view)
    t0 = time.perf_counter()
    while not run.finished:
        run.enqueue()
    t_enq = time.perf_counter() - t0
    t.append(mark())
    log = run.finish(); t.append(mark())
    names = ["scan_bins", "engine (symbolise, dedup, count)", "hist all-reduce", "peer block + run barriers", "merge loop", "read log"]
    d = [b - a for a, b in zip(t[:-1], t[1:])]
    if rep and (best is None or sum(d) < sum(best[0])):
        best = (d, t_enq, len(log), eng.dedup_stats)
if rank == 0:
    d, t_enq, n, stats = best
    print(f"world {world}: total {sum(d)*1e3:.1f} ms, merges {n}, host enqueue {t_enq*1e3:.1f} ms, dedup {stats and {k: stats[k] for k in ('distinct_symbol_ratio', 'pseudo_sequences')}}")
    for nme, v in zip(names, d):
        print(f"   {nme:36s} {v*1e3:8.2f} ms")
dist.destroy_process_group()
