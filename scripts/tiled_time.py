"""Time the tiled K1 / K3 (csrc/spline_tiled.cu) on several geometries through the C ABI on preallocated buffers and
print a digest of the outputs (sha-1: compare across builds / processes; profiles/r02_tiled_kernels_timing.txt)."""
import hashlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beast_tokenizer_b200 import BEASTBsplineTokenizer, _lib
from beast_tokenizer_b200.synth import synth_device
dev = torch.device("cuda", 0)
GEOMS = [
    ("shipped", dict(num_dof=32, num_basis=50, seq_len=10, vocab_size=1000, degree_p=0), 32768),
    ("d5_cubic_T33", dict(num_dof=5, num_basis=8, seq_len=33, vocab_size=1000, degree_p=3, gripper_zero_order=True, gripper_indices=[0]), 262144),
    ("d3_nb20_T64", dict(num_dof=3, num_basis=20, seq_len=64, vocab_size=512, degree_p=2), 262144),
    ("d14_T50_odd", dict(num_dof=14, num_basis=10, seq_len=50, vocab_size=256, degree_p=4, gripper_zero_order=True, gripper_indices=[6, 13]), 15),
]
def digest(t):
    return hashlib.sha1(t.cpu().numpy().tobytes()).hexdigest()[:12]
for name, geom, n in GEOMS:
    tok = BEASTBsplineTokenizer(device="cuda", **geom)
    T, D, nb = geom["seq_len"], geom["num_dof"], geom["num_basis"]
    x = synth_device(n, T, D, 11, dev)
    tok.update_weights_bounds(x)
    plan = tok._plan(); lo, hi = tok._bounds(dev)
    toks = torch.empty((n, nb * D), device=dev, dtype=torch.int64)
    pars = torch.empty((n, nb * D), device=dev, dtype=torch.float32)
    out = torch.empty((n, T, D), device=dev, dtype=torch.float32)
    def enc(): _lib.check(plan._lib.beast_encode_f32(plan.handle, _lib.ptr(x), n, _lib.ptr(lo), _lib.ptr(hi), 0, _lib.ptr(pars), _lib.ptr(toks), _lib.stream_ptr(dev)), "e")
    def dec(): _lib.check(plan._lib.beast_decode_f32(plan.handle, _lib.ptr(toks), n, _lib.ptr(lo), _lib.ptr(hi), 0, None, _lib.ptr(out), _lib.stream_ptr(dev)), "d")
    def mm(): tok.update_weights_bounds(x)
    enc(); dec(); torch.cuda.synchronize()
    reps = 10
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    for _ in range(reps): enc()
    ev[1].record()
    for _ in range(reps): dec()
    ev[2].record()
    for _ in range(reps): mm()
    ev[3].record()
    torch.cuda.synchronize()
    e_ms, d_ms, m_ms = (ev[i].elapsed_time(ev[i + 1]) / reps for i in range(3))
    eb, db = 4 * T * D + 12 * nb * D, 8 * nb * D + 4 * T * D
    print(f"{name:14s} n={n:7d} encode {e_ms * 1e3:8.1f} us ({eb * n / e_ms / 1e6:6.0f} GB/s)  decode {d_ms * 1e3:8.1f} us ({db * n / d_ms / 1e6:6.0f} GB/s)  "
          f"bounds {m_ms * 1e3:8.1f} us  digests tok {digest(toks)} par {digest(pars)} rec {digest(out)} lo {digest(tok.w_min)} hi {digest(tok.w_max)}")
