import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beast_tokenizer_b200 import BEASTBsplineTokenizer, _lib
from beast_tokenizer_b200.synth import synth_device
dev = torch.device("cuda", 0)
tok = BEASTBsplineTokenizer(num_dof=14, num_basis=10, seq_len=50, vocab_size=256, gripper_zero_order=True, gripper_indices=[6, 13], device="cuda")
x = synth_device(100_000, 50, 14, 7, dev)
plan = tok._plan(); lib = plan._lib
lo = torch.empty(140, device=dev); hi = torch.empty(140, device=dev)
def raw(): _lib.check(lib.beast_fit_minmax_f32(plan.handle, _lib.ptr(x), 100000, _lib.ptr(lo), _lib.ptr(hi), 0, _lib.stream_ptr(dev)), "mm")
for name, fn in (("raw C-ABI", raw), ("update_weights_bounds", lambda: tok.update_weights_bounds(x))):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(50): fn()
    e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"{name}: gpu {e0.elapsed_time(e1)/50*1e3:.1f} us/call, host wall {(t1-t0)/50*1e6:.1f} us/call")
