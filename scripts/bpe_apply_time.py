"""Time BPE encode / decode (device CSR) on 1 M sequences: python scripts/bpe_apply_time.py [n_seq]"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beast_tokenizer_b200 import BEASTBsplineTokenizer, BEASTBsplineBPETokenizer  # noqa: E402
from beast_tokenizer_b200.synth import synth_device  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
dev = torch.device("cuda:0")
tok = BEASTBsplineTokenizer(num_dof=14, num_basis=10, seq_len=50, vocab_size=256, gripper_zero_order=True,
                            gripper_indices=[6, 13], device="cuda:0")
tok.update_weights_bounds(synth_device(100_000, 50, 14, 1, dev))
btok = BEASTBsplineBPETokenizer.from_beast(tok, bpe_vocab_size=2048)
btok.fit_from_trajectories([{"actions": synth_device(65536, 50, 14, 1000, dev)}], show_progress=False)
x = synth_device(n, 50, 14, 5, dev)
mp, _ = btok.encode_to_mp_tokens(x)
flat, offsets = btok._discrete_to_bpe_csr(mp)
back = btok._bpe_csr_to_discrete(flat, offsets)
assert torch.equal(back, mp)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for rep in range(3):
    torch.cuda.synchronize()
    ev[0].record()
    flat, offsets = btok._discrete_to_bpe_csr(mp)
    ev[1].record()
    back = btok._bpe_csr_to_discrete(flat, offsets)
    ev[2].record()
    torch.cuda.synchronize()
    e, d = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    print(f"encode {e:.2f} ms ({n / e / 1e3:.1f} M seq/s)   decode {d:.2f} ms ({n / d / 1e3:.1f} M seq/s)   "
          f"{flat.numel() / n:.1f} ids/seq")
