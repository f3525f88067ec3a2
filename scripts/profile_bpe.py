"""Short driver for ncu: BPE training (few merges) + encode/decode on a mid-size corpus."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beast_tokenizer_b200 import BEASTBsplineTokenizer, FIGBPE
from beast_tokenizer_b200.synth import SyntheticLoader, synth_device

N = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
VOCAB = int(sys.argv[2]) if len(sys.argv) > 2 else 340
dev = torch.device("cuda", 0)
tok = BEASTBsplineTokenizer(num_dof=14, num_basis=10, seq_len=50, vocab_size=256, gripper_zero_order=True,
                            gripper_indices=[6, 13], device="cuda")
tok.fit_parameters(SyntheticLoader(20, 32, 50, 14, seed0=1), verbose=False)
bins = torch.cat([tok.encode(synth_device(N // 4, 50, 14, 1000 + c, dev))[0] for c in range(4)])
st = FIGBPE(vocab_size=VOCAB, show_progress=False).fit_from_bins(bins)
flat, offsets, status = st.tokenizer.encode_bins(bins[:262144], st.min_token, st.max_token)
back, s2, l2 = st.tokenizer.decode_ids(flat, offsets, 140, st.min_token)
torch.cuda.synchronize()
print("ok", len(st.tokenizer.merges), flat.numel(), bool(torch.equal(back, bins[:262144])))
