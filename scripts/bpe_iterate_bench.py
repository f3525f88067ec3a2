"""Isolate bpe_iterate_kernel: bpe_train_step with N = 0 launches only the iteration head."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beast_tokenizer_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda", 0)
V = 2048
for n_tok in (400, 1300, 2000):
    hist = torch.randint(0, 1000, (V, V), device=dev, dtype=torch.int32)
    delta = torch.zeros(4 * V, device=dev, dtype=torch.int32)
    ctl = torch.zeros(8, device=dev, dtype=torch.int32); ctl[4] = n_tok
    log = torch.zeros(4 * 4096, device=dev, dtype=torch.int32)
    result = torch.zeros(256, device=dev, dtype=torch.int64)
    work = torch.zeros(16, device=dev, dtype=torch.int32)
    def run(iters):
        _lib.check(lib.bpe_train_step(None, None, 0, 1, V, _lib.ptr(hist), _lib.ptr(delta), _lib.ptr(ctl), _lib.ptr(log),
                                      _lib.ptr(result), _lib.ptr(work), 1 << 20, 1, 4096, 0, None, iters, _lib.stream_ptr(dev)), "step")
    run(3)
    ctl[4] = n_tok; ctl[5] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); run(50); e1.record(); torch.cuda.synchronize()
    print(f"n_active {n_tok}..{int(ctl[4])}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per iterate launch")
