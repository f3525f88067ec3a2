"""Why does the Python-API encode of a large batch take longer than the kernel?  Counts device allocations per call."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beast_tokenizer_b200 import BEASTBsplineTokenizer
from beast_tokenizer_b200.synth import synth_device
dev = torch.device("cuda", 0)
tok = BEASTBsplineTokenizer(num_dof=32, num_basis=50, seq_len=10, vocab_size=1000, degree_p=0, device="cuda")
x = synth_device(32768, 10, 32, 11, dev)
tok.update_weights_bounds(x)
for label in ("reassign", "del first"):
    tokens = pd = None
    for _ in range(3):
        tokens, pd = tok.encode(x)
    torch.cuda.synchronize()
    s0 = torch.cuda.memory_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(10):
        if label == "del first":
            tokens = pd = None
        tokens, pd = tok.encode(x)
    e1.record(); t_host = time.perf_counter() - t0; torch.cuda.synchronize()
    s1 = torch.cuda.memory_stats()
    print(label, "gpu ms/call", e0.elapsed_time(e1) / 10, "host enqueue ms/call", t_host * 100,
          "cudaMalloc calls", s1["num_device_alloc"] - s0["num_device_alloc"], "cudaFree calls", s1["num_device_free"] - s0["num_device_free"],
          "reserved GB", s1["reserved_bytes.all.current"] / 2**30)
