"""torchrun --nproc-per-node N scripts/bpe_multi_check.py [n_sequences] — sharded BPE training over N
GPUs (NCCL) must reproduce the unsharded merge table; prints timings."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beast_tokenizer_b200 import BEASTBsplineTokenizer, FIGBPE
from beast_tokenizer_b200.synth import SyntheticLoader, synth_device

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
CH = 40
tok = BEASTBsplineTokenizer(num_dof=14, num_basis=10, seq_len=50, vocab_size=256, gripper_zero_order=True,
                            gripper_indices=[6, 13], device=f"cuda:{local}")
tok.fit_parameters(SyntheticLoader(100, 32, 50, 14, seed0=1), verbose=False)
chunk = lambda c: tok.encode(synth_device(N // CH, 50, 14, 1000 + c, dev))[0]
shard = torch.cat([chunk(c) for c in range(CH) if c % world == rank])
fig = FIGBPE(vocab_size=2048, show_progress=False, device=f"cuda:{local}")
fig.fit_from_bins(shard[:2048])
torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
st = fig.fit_from_bins(shard)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
txt = st.tokenizer.merges_txt()
objs = [None] * world
dist.all_gather_object(objs, txt)
if rank == 0:
    assert all(o == txt for o in objs), "ranks disagree on the merge table"
    full = torch.cat([chunk(c) for c in range(CH)])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ref = FIGBPE(vocab_size=2048, show_progress=False, device=f"cuda:{local}", process_group=False).fit_from_bins(full)
    torch.cuda.synchronize(); dt1 = time.perf_counter() - t0
    same = ref.tokenizer.merges_txt() == txt and ref.tokenizer.vocab_json() == st.tokenizer.vocab_json()
    print(f"world {world}: sharded {dt:.3f} s ({len(st.tokenizer.merges) / dt:.0f} merges/s), unsharded on one GPU {dt1:.3f} s "
          f"({len(ref.tokenizer.merges) / dt1:.0f} merges/s), identical={same}, merges={len(st.tokenizer.merges)}")
    assert same
dist.destroy_process_group()
