"""Driver for ncu: one FIGBPE.fit_from_bins on an i.i.d. corpus (set-up kernels: symbolise, word tables, pair count)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beast_tokenizer_b200 import BEASTBsplineTokenizer, FIGBPE
from beast_tokenizer_b200.synth import SyntheticLoader, synth_device
dev = torch.device("cuda", 0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
tok = BEASTBsplineTokenizer(num_dof=14, num_basis=10, seq_len=50, vocab_size=256, gripper_zero_order=True, gripper_indices=[6, 13], device="cuda")
tok.fit_parameters(SyntheticLoader(100, 32, 50, 14, seed0=1), verbose=False)
bins = torch.cat([tok.encode(synth_device(N // 25, 50, 14, 1000 + c, dev))[0] for c in range(25)])
st = FIGBPE(vocab_size=300, show_progress=False, process_group=False, dedup="auto").fit_from_bins(bins)
torch.cuda.synchronize()
print("ok", st.tokenizer.dedup_stats)
