"""Short driver for ncu: a few encode (K1) + decode (K3) launches at the bench workload."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beast_tokenizer_b200 import BEASTBsplineTokenizer  # noqa: E402
from beast_tokenizer_b200.synth import synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
tok = BEASTBsplineTokenizer(num_dof=14, num_basis=10, seq_len=50, vocab_size=256, gripper_zero_order=True,
                            gripper_indices=[6, 13], device="cuda", llm_vocab_size=32000)
x = synth(B, 50, 14, seed=2, device="cuda")
tok.update_weights_bounds(x[:4096])
for _ in range(steps):
    tokens, _ = tok.encode(x)
    rec = tok.reconstruct_traj(tokens)
torch.cuda.synchronize()
print("ok", tokens.shape, rec.shape)
