import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beast_tokenizer_b200 import BEASTBsplineTokenizer, _lib
from beast_tokenizer_b200.beast_bpe_trainer import GpuBpeEngine, _Collective, scan_bins_gpu, build_alphabet
from beast_tokenizer_b200.synth import SyntheticLoader, synth_device
dev = torch.device("cuda", 0)
tok = BEASTBsplineTokenizer(num_dof=14, num_basis=10, seq_len=50, vocab_size=256, gripper_zero_order=True, gripper_indices=[6, 13], device="cuda")
tok.fit_parameters(SyntheticLoader(100, 32, 50, 14, seed0=1), verbose=False)
for N in (65536, 400000):
    bins = tok.encode(synth_device(N, 50, 14, 1000, dev))[0]
    coll = _Collective(enabled=False)
    mn, mx, seen = scan_bins_gpu(bins, coll)
    tokens, b2i = build_alphabet(mn, mx, seen)
    for mode in ("plain", "graph"):
        eng = GpuBpeEngine(bins, mn, b2i, 2048)
        lib = eng.lib
        ctl = torch.zeros(8, device=dev, dtype=torch.int32); ctl[4] = len(tokens)
        mm = 2048 - len(tokens)
        log = torch.zeros(4 * mm, device=dev, dtype=torch.int32); eng.result.zero_()
        SIG = None
def step(ph):
            _lib.check(lib.bpe_train_step(_lib.ptr(eng.sym), _lib.ptr(eng.len), eng.N, eng.stride, eng.V, _lib.ptr(eng.hist), _lib.ptr(eng.delta),
                                          _lib.ptr(ctl), _lib.ptr(log), _lib.ptr(eng.result), _lib.ptr(eng.work), 2048, 2, mm, ph, SIG, 1, _lib.stream_ptr(dev)), "s")
        torch.cuda.synchronize(); t0 = time.perf_counter()
        if mode == "plain":
            for _ in range(mm): step(0); step(1)
            t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
            print(N, mode, f"enqueue {t1-t0:.3f}s total {t2-t0:.3f}s merges {int(ctl[5])}")
        else:
            step(0); step(1); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph(); tc = time.perf_counter()
            with torch.cuda.graph(g):
                for _ in range(8): step(0); step(1)
            torch.cuda.synchronize(); t1 = time.perf_counter()
            for _ in range((mm + 6) // 8): g.replay()
            torch.cuda.synchronize(); t2 = time.perf_counter()
            print(N, mode, f"capture {t1-tc:.3f}s replay {t2-t1:.3f}s merges {int(ctl[5])}")
