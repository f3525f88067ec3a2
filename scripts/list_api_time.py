"""Time the list-returning BPE API (reference return type List[List[int]]) at 65 536 trajectories, with a breakdown."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beast_tokenizer_b200 import BEASTBsplineTokenizer, BEASTBsplineBPETokenizer
import beast_tokenizer_b200.beast_bspline_bpe_tokenizer as M
from beast_tokenizer_b200.synth import synth
dev = torch.device("cuda", 0)
B = 65536
tok = BEASTBsplineTokenizer(num_dof=14, num_basis=10, seq_len=50, vocab_size=256, gripper_zero_order=True,
                            gripper_indices=[6, 13], device="cuda")
x = synth(B, 50, 14, seed=5, device=dev)
tok.update_weights_bounds(x[:4096])
btok = BEASTBsplineBPETokenizer.from_beast(tok, bpe_vocab_size=2048)
btok.fit_from_trajectories([{"actions": x}], show_progress=False)
print("helper:", M._pylists())
for mode in ("c", "py"):
    if mode == "py":
        M._PYLISTS = False
    btok.encode(x[:1024]); torch.cuda.synchronize()
    for rep in range(2):
        t0 = time.perf_counter(); ids, _ = btok.encode(x); t1 = time.perf_counter()
        rec = btok.reconstruct_traj(ids); torch.cuda.synchronize(); t2 = time.perf_counter()
        print(f"{mode} run {rep}: encode {B / (t1 - t0) / 1e6:.2f} M traj/s ({(t1 - t0) * 1e3:.1f} ms), "
              f"reconstruct_traj {B / (t2 - t1) / 1e6:.2f} M traj/s ({(t2 - t1) * 1e3:.1f} ms), ids/traj {sum(map(len, ids)) / B:.1f}")
    mp, _ = btok.encode_to_mp_tokens(x); torch.cuda.synchronize()
    t0 = time.perf_counter(); flat, off = btok._discrete_to_bpe_csr(mp); torch.cuda.synchronize(); t1 = time.perf_counter()
    fh, oh = flat.cpu().numpy(), off.cpu().numpy(); t2 = time.perf_counter()
    rows = M._split_rows(fh, oh); t3 = time.perf_counter()
    f2 = M._flatten_rows(rows); t4 = time.perf_counter()
    print(f"{mode} breakdown: csr kernels {(t1 - t0) * 1e3:.1f} ms, D2H {(t2 - t1) * 1e3:.1f} ms, split {(t3 - t2) * 1e3:.1f} ms, flatten {(t4 - t3) * 1e3:.1f} ms")
