"""Event-time K1/K3 back to back (graph replay) — quick A/B of kernel variants."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beast_tokenizer_b200 import BEASTBsplineTokenizer, _lib
from beast_tokenizer_b200.synth import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
tok = BEASTBsplineTokenizer(num_dof=14, num_basis=10, seq_len=50, vocab_size=256, gripper_zero_order=True,
                            gripper_indices=[6, 13], device="cuda", llm_vocab_size=32000)
dev = torch.device("cuda", 0)
R = 4
xs = [synth(B, 50, 14, seed=2 + i, device=dev) for i in range(R)]
tok.update_weights_bounds(xs[0][:4096])
if os.environ.get("QT_TIGHT") == "1":      # 1 % / 99 % quantile bounds: 2 % of the coefficients clamp
    from beast_tokenizer_b200.synth import SyntheticLoader
    tok.fit_parameters(SyntheticLoader(100, 32, 50, 14, seed0=1), verbose=False)
plan = tok._plan(); lib = plan._lib; lo, hi = tok._bounds(dev)
toks = [torch.empty((B, 140), device=dev, dtype=torch.int64) for _ in range(R)]
pars = [torch.empty((B, 140), device=dev, dtype=torch.float32) for _ in range(R)]
outs = [torch.empty((B, 50, 14), device=dev, dtype=torch.float32) for _ in range(R)]
def enc(i): _lib.check(lib.beast_encode_f32(plan.handle, _lib.ptr(xs[i]), B, _lib.ptr(lo), _lib.ptr(hi), 31744, _lib.ptr(pars[i]), _lib.ptr(toks[i]), _lib.stream_ptr(dev)), "e")
def dec(i): _lib.check(lib.beast_decode_f32(plan.handle, _lib.ptr(toks[i]), B, _lib.ptr(lo), _lib.ptr(hi), 31744, None, _lib.ptr(outs[i]), _lib.stream_ptr(dev)), "d")
for i in range(R): enc(i); dec(i)
torch.cuda.synchronize()
K = 40
for name, fn in (("encode", enc), ("decode", dec)):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for j in range(K): fn(j % R)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / K
    nbytes = (4480 if name == "encode" else 3920) * B
    print(f"{name}: {us:.2f} us/launch  {nbytes / us / 1e3:.0f} GB/s  {B / us:.1f} M traj/s")

def timed(label, seq):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for fn, i in seq: fn(i)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    print(f"{label}: {e0.elapsed_time(e1) * 1e3 / (len(seq) / 2):.2f} us per (enc+dec) pair")

timed("alternating enc(j) dec(j+2)  ", [p for j in range(K) for p in ((enc, j % R), (dec, (j + 2) % R))])
timed("alternating enc(j) dec(j)    ", [p for j in range(K) for p in ((enc, j % R), (dec, j % R))])
timed("blocked 40 enc then 40 dec   ", [(enc, j % R) for j in range(K)] + [(dec, j % R) for j in range(K)])
timed("pairs of two: e e d d        ", [p for j in range(K // 2) for p in ((enc, (2 * j) % R), (enc, (2 * j + 1) % R), (dec, (2 * j + 2) % R), (dec, (2 * j + 3) % R))])

if os.environ.get("QT_SAMPLER") == "1":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import ClockSampler
    sm = ClockSampler(0); sm.start()
    import time; time.sleep(0.05)
    timed("with NVML sampler thread     ", [p for j in range(K) for p in ((enc, j % R), (dec, (j + 2) % R))])
    timed("with NVML sampler thread (2) ", [p for j in range(K) for p in ((enc, j % R), (dec, (j + 2) % R))])
    sm.stop_flag = True; sm.join()
    print(sm.summary())
    timed("sampler stopped              ", [p for j in range(K) for p in ((enc, j % R), (dec, (j + 2) % R))])
