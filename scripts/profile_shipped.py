"""Driver for ncu: the reference's shipped shape (nb=50, degree 0, V=1000, [10, 32]) through the tiled K1 / K3."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beast_tokenizer_b200 import BEASTBsplineTokenizer
from beast_tokenizer_b200.synth import synth_device
dev = torch.device("cuda", 0)
tok = BEASTBsplineTokenizer(num_dof=32, num_basis=50, seq_len=10, vocab_size=1000, degree_p=0, device="cuda")
x = synth_device(32768, 10, 32, 11, dev)
tok.update_weights_bounds(x)
for _ in range(3):
    tokens, _ = tok.encode(x)
    rec = tok.reconstruct_traj(tokens)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record(); tokens, _ = tok.encode(x); e[1].record(); rec = tok.reconstruct_traj(tokens); e[2].record(); torch.cuda.synchronize()
print("encode ms", e[0].elapsed_time(e[1]), "decode ms", e[1].elapsed_time(e[2]))
