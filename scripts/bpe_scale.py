"""BPE training at BASELINE configs[3] scale on one GPU (+ the HF CPU trainer on a subset)."""
import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beast_tokenizer_b200 import BEASTBsplineTokenizer, FIGBPE
from beast_tokenizer_b200.synth import SyntheticLoader, synth_device

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_600_000
VOCAB = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
CPU_N = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
dev = torch.device("cuda", 0)
tok = BEASTBsplineTokenizer(num_dof=14, num_basis=10, seq_len=50, vocab_size=256, gripper_zero_order=True,
                            gripper_indices=[6, 13], device="cuda")
tok.fit_parameters(SyntheticLoader(100, 32, 50, 14, seed0=1), verbose=False)
t0 = time.perf_counter()
chunks = []
for c in range((N + 65535) // 65536):
    n = min(65536, N - c * 65536)
    chunks.append(tok.encode(synth_device(n, 50, 14, 1000 + c, dev), respect_llm_vocab_size=False)[0])
bins = torch.cat(chunks)
torch.cuda.synchronize()
print(f"corpus: {bins.shape} in {time.perf_counter() - t0:.2f} s; bins min/max {int(bins.min())} {int(bins.max())}")
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    st = FIGBPE(vocab_size=VOCAB, show_progress=False).fit_from_bins(bins)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"GPU train {N} seqs vocab {VOCAB}: {dt:.3f} s, {len(st.tokenizer.merges)} merges, {len(st.tokenizer.merges) / dt:.1f} merges/s")
if CPU_N:
    sub = bins[:CPU_N]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    st_sub = FIGBPE(vocab_size=VOCAB, show_progress=False).fit_from_bins(sub)
    torch.cuda.synchronize(); dt_g = time.perf_counter() - t0
    from tokenizers import ByteLevelBPETokenizer
    from tokenizers.trainers import BpeTrainer
    sub_np = sub.cpu().numpy()
    mn, mx = int(sub_np.min()), int(sub_np.max())
    t0 = time.perf_counter()
    strings = ["".join(map(chr, (row - mn).astype(int))) for row in sub_np]
    t1 = time.perf_counter()
    hf = ByteLevelBPETokenizer()
    trainer = BpeTrainer(vocab_size=VOCAB, min_frequency=2, show_progress=False, special_tokens=[],
                         initial_alphabet=[chr(i) for i in range(mx - mn + 1)], max_token_length=10000)
    hf._tokenizer.train_from_iterator(strings, trainer=trainer)
    t2 = time.perf_counter()
    model = json.loads(hf._tokenizer.to_str())["model"]
    hf_merges = [m if isinstance(m, str) else " ".join(m) for m in model["merges"]]
    same = hf_merges == [f"{a} {b}" for a, b in st_sub.tokenizer.merge_strings()] and hf.get_vocab() == st_sub.tokenizer.get_vocab()
    print(f"subset {CPU_N}: GPU {dt_g:.3f} s ({len(st_sub.tokenizer.merges) / dt_g:.1f} merges/s); HF CPU strings {t1 - t0:.2f} s + train {t2 - t1:.2f} s "
          f"({len(hf_merges) / (t2 - t1):.1f} merges/s, {os.cpu_count()} cores); identical={same}")
