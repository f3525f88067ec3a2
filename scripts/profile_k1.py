"""Driver for ncu: K1 / K3 on the bench shape (65 536 x 14 DoF) through the C ABI."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beast_tokenizer_b200 import BEASTBsplineTokenizer
from beast_tokenizer_b200.synth import SyntheticLoader, synth_device
dev = torch.device("cuda", 0)
tok = BEASTBsplineTokenizer(num_dof=14, num_basis=10, seq_len=50, vocab_size=256, gripper_zero_order=True, gripper_indices=[6, 13], device="cuda", llm_vocab_size=32000)
tok.fit_parameters(SyntheticLoader(100, 32, 50, 14, seed0=1), verbose=False)
x = synth_device(65536, 50, 14, 2, dev)
for _ in range(4):
    tokens, _ = tok.encode(x)
    rec = tok.reconstruct_traj(tokens)
    tok.update_weights_bounds(x)
torch.cuda.synchronize()
print("ok")
