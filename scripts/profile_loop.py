"""Driver for an ncu launch list of the BPE merge loop on a small (repetitive or sharded-size) corpus."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beast_tokenizer_b200 import BEASTBsplineTokenizer, FIGBPE
from beast_tokenizer_b200.synth import SyntheticLoader, synth_device
dev = torch.device("cuda", 0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
tok = BEASTBsplineTokenizer(num_dof=14, num_basis=10, seq_len=50, vocab_size=256, gripper_zero_order=True, gripper_indices=[6, 13], device="cuda")
tok.fit_parameters(SyntheticLoader(100, 32, 50, 14, seed0=1), verbose=False)
bins = torch.cat([tok.encode(synth_device(N // 4, 50, 14, 1000 + c, dev))[0] for c in range(4)])
st = FIGBPE(vocab_size=2048, show_progress=False).fit_from_bins(bins)
torch.cuda.synchronize()
print("merges", len(st.tokenizer.merges), st.tokenizer.dedup_stats)
