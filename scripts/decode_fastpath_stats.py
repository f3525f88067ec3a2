"""How often does the BPE decode fast path (per-token character table) apply?  Trains the bench table on a sample,
encodes 262 144 sequences and counts: tokens flagged slow, sequences holding one, sequences with a boundary mismatch."""
import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beast_tokenizer_b200 import BEASTBsplineTokenizer, FIGBPE
from beast_tokenizer_b200.synth import SyntheticLoader, synth_device
dev = torch.device("cuda", 0)
tok = BEASTBsplineTokenizer(num_dof=14, num_basis=10, seq_len=50, vocab_size=256, gripper_zero_order=True, gripper_indices=[6, 13], device="cuda")
tok.fit_parameters(SyntheticLoader(100, 32, 50, 14, seed0=1), verbose=False)
bins = torch.cat([tok.encode(synth_device(100000, 50, 14, 1000 + c, dev))[0] for c in range(4)])
st = FIGBPE(vocab_size=2048, show_progress=False).fit_from_bins(bins)
m = st.tokenizer
flat, off, status = m.encode_bins(bins[:262144], st.min_token, st.max_token)
t_ = m._tables(dev); tab = t_["tab"].view(-1, 2 if t_["tab_slots"] == 2 else 4)
meta = tab[:, -1]
slow = (meta & 0x80) != 0
nst, need, lead = meta & 7, (meta >> 3) & 3, (meta >> 5) & 3
print("vocab", len(m.tokens), "slow tokens", int(slow.sum()), "max chars", int(nst.max()), "tokens with need>0", int((need > 0).sum()),
      "tokens with lead_cont>0", int((lead > 0).sum()))
lens = torch.tensor([len(m.token_bytes(i)) for i in range(len(m.tokens))])
print("token bytes: max", int(lens.max()), "hist", torch.bincount(lens).tolist())
ids = flat.long()
seq_of = torch.repeat_interleave(torch.arange(off.numel() - 1, device=dev), (off[1:] - off[:-1]))
n_seq = off.numel() - 1
has_slow = torch.zeros(n_seq, device=dev, dtype=torch.bool).index_put_((seq_of[slow[ids]],), torch.tensor(True, device=dev))
nxt_lead = torch.zeros_like(ids)
nxt_lead[:-1] = lead[ids[1:]]
last = torch.zeros_like(ids, dtype=torch.bool); last[(off[1:] - 1)] = True
nxt_lead[last] = 0
mism = need[ids] != nxt_lead
has_mism = torch.zeros(n_seq, device=dev, dtype=torch.bool).index_put_((seq_of[mism],), torch.tensor(True, device=dev))
first_lead = lead[ids[off[:-1]]] != 0
print(f"sequences {n_seq}: with slow token {float(has_slow.float().mean()):.4f}, with boundary mismatch {float(has_mism.float().mean()):.4f}, "
      f"first token starts inside a char {float(first_lead.float().mean()):.4f}, any fallback {float((has_slow | has_mism | first_lead).float().mean()):.4f}")
for _ in range(3):
    back, s2, l2 = m.decode_ids(flat, off, 140, st.min_token)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); back, s2, l2 = m.decode_ids(flat, off, 140, st.min_token); e1.record(); torch.cuda.synchronize()
print("decode ms", e0.elapsed_time(e1), "seq/s", n_seq / e0.elapsed_time(e1) * 1e3, "exact", bool(torch.equal(back, bins[:262144])))
