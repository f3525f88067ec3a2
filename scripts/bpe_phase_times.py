"""Phase timings of one FIGBPE.fit_from_bins (synchronised between phases; device 0):
scan, symbolise, word tables, pack, pair count, merge loop, host vocabulary.  `iid` = bench corpus, `rep` = repetitive."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beast_tokenizer_b200 import BEASTBsplineTokenizer, FIGBPE, _lib
from beast_tokenizer_b200 import beast_bpe_trainer as T
from beast_tokenizer_b200.synth import SyntheticLoader, synth_device

dev = torch.device("cuda", 0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_600_000
kind = sys.argv[2] if len(sys.argv) > 2 else "iid"
tok = BEASTBsplineTokenizer(num_dof=14, num_basis=10, seq_len=50, vocab_size=256, gripper_zero_order=True, gripper_indices=[6, 13], device="cuda")
tok.fit_parameters(SyntheticLoader(100, 32, 50, 14, seed0=1), verbose=False)
bins = torch.cat([tok.encode(synth_device(N // 25, 50, 14, 1000 + c, dev))[0] for c in range(25)])
if kind == "rep":
    g = torch.Generator(device=dev).manual_seed(4)
    bins = bins[:20000][torch.randint(0, 20000, (N,), generator=g, device=dev)].contiguous()

marks = []
def mark(name):
    torch.cuda.synchronize(); marks.append((name, time.perf_counter()))

# instrument the engine's phases by wrapping library calls
lib = _lib.load()
acc = {}
def timed(name):
    fn = getattr(lib, name)
    def wrap(*a):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        rc = fn(*a)
        torch.cuda.synchronize(); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
        return rc
    return wrap
class Lib:
    def __getattr__(self, k):
        if k.startswith("bpe_") and k not in ("bpe_signature_words",):
            return timed(k)
        return getattr(lib, k)
for dedup in ("auto", False):
    for rep in range(2):
        acc.clear()
        T._lib.load = lambda *a, **k: Lib()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        st = FIGBPE(vocab_size=2048, show_progress=False, process_group=False, dedup=dedup).fit_from_bins(bins)
        torch.cuda.synchronize(); total = time.perf_counter() - t0
        T._lib.load = _lib.load
    print(f"{kind} N={N} dedup={dedup}: total {total*1e3:.1f} ms (synchronised phases)  stats={st.tokenizer.dedup_stats}")
    for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
        print(f"   {k:28s} {v*1e3:8.2f} ms")
    print(f"   {'other (torch ops, host)':28s} {(total - sum(acc.values()))*1e3:8.2f} ms")
