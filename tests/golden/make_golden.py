"""Generate golden fixtures from the LIVE reference (run in the build container).

    python tests/golden/make_golden.py

Imports the unmodified reference from /root/reference through
tests/golden/ref_harness.py, runs its public API on seeded synthetic inputs and
writes the inputs and outputs to tests/golden/*.npz (+ the reference's own
save_pretrained files).  The fixtures, not the reference, travel to the GPU box.
Library versions are recorded in tests/golden/VERSIONS.json.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from ref_harness import import_reference  # noqa: E402
from beast_tokenizer_b200.synth import synth, SyntheticLoader  # noqa: E402


def npy(x):
    return x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)


def spline_case(ref, name, *, num_dof, num_basis, seq_len, vocab_size, degree_p=4,
                gripper_zero_order=False, gripper_indices=None, llm_vocab_size=None,
                batch, seed, fit_batches=0, fit_seed0=0, custom_T=None):
    torch.manual_seed(0)
    tok = ref.BEASTBsplineTokenizer(
        num_dof=num_dof, num_basis=num_basis, seq_len=seq_len, vocab_size=vocab_size,
        degree_p=degree_p, gripper_zero_order=gripper_zero_order, gripper_indices=gripper_indices,
        device="cpu", llm_vocab_size=llm_vocab_size)
    out = {}
    x = synth(batch, seq_len, num_dof, seed)
    out["trajs"] = npy(x)
    out["times"] = npy(tok.times)
    out["phi_joint"] = npy(tok.mp.basis_gn.basis(tok.times))
    out["knots_joint"] = npy(tok.mp.basis_gn.knots_vec)
    if tok.gripper_mp is not None:
        out["phi_grip"] = npy(tok.gripper_mp.basis_gn.basis(tok.times))
    out["joint_indices"] = np.asarray(tok.joint_indices, dtype=np.int64)
    out["gripper_indices"] = np.asarray(tok.gripper_indices, dtype=np.int64)

    # default bounds (+-0.02)
    out["w_min_default"], out["w_max_default"] = npy(tok.w_min).copy(), npy(tok.w_max).copy()
    toks, pd = tok.encode(x)
    out["tokens_default"], out["params"] = npy(toks), npy(pd["params"])
    out["recon_default"] = npy(tok.reconstruct_traj(toks))

    # min/max bounds (update_weights_bounds) and per-batch hysteresis expansion
    tok.update_weights_bounds(x)
    out["w_min_minmax"], out["w_max_minmax"] = npy(tok.w_min).copy(), npy(tok.w_max).copy()
    x2 = synth(batch, seq_len, num_dof, seed + 7) * 1.5
    toks_ub, pd_ub = tok.encode(x2, update_bounds=True)
    out["trajs_ub"] = npy(x2)
    out["w_min_expand"], out["w_max_expand"] = npy(tok.w_min).copy(), npy(tok.w_max).copy()
    out["tokens_ub"] = npy(toks_ub)

    # quantile bounds (fit_parameters) over a seeded loader
    if fit_batches:
        loader = SyntheticLoader(fit_batches, 32, seq_len, num_dof, seed0=fit_seed0)
        tok.fit_parameters(loader, verbose=False)
        out["fit_batches"] = np.int64(fit_batches)
        out["fit_seed0"] = np.int64(fit_seed0)
        out["w_min_fit"], out["w_max_fit"] = npy(tok.w_min).copy(), npy(tok.w_max).copy()
        toks_f, pd_f = tok.encode(x)
        out["tokens_fit"] = npy(toks_f)
        toks_nf, _ = tok.encode(x, respect_llm_vocab_size=False)
        out["tokens_fit_nooffset"] = npy(toks_nf)
        out["recon_fit"] = npy(tok.reconstruct_traj(toks_f))
        init_p = x[:, 0, :] + 0.001
        out["init_p"] = npy(init_p)
        out["recon_fit_initp"] = npy(tok.reconstruct_traj(toks_f, init_p=init_p))
        out["decode_fit"] = npy(tok.decode(toks_f))
        ctoks, _ = tok.encode_continuous(x)
        out["cont_tokens_fit"] = npy(ctoks)
        # reconstruct_traj_continuous is broken upstream: beast/utils.py:42 calls
        # torch.clamp(float) -> TypeError.  Record that fact instead of an output.
        try:
            tok.reconstruct_traj_continuous(ctoks)
            out["recon_cont_raises"] = np.int64(0)
        except TypeError:
            out["recon_cont_raises"] = np.int64(1)
        if custom_T:
            g = torch.Generator().manual_seed(seed + 99)
            tt = torch.sort(torch.rand(batch, custom_T, generator=g) * float(tok.duration), dim=1)[0]
            tt[:, 0] = 0.0
            tt[:, -1] = float(tok.times[-1])
            out["custom_times"] = npy(tt)
            out["recon_fit_custom_times"] = npy(tok.reconstruct_traj(toks_f, times=tt))
        if llm_vocab_size is not None:
            out["llm_tokens"] = npy(tok.tokens_to_llm_tokens(toks_nf))
            out["mp_tokens_3d"] = npy(tok.llm_tokens_to_mp_tokens(toks_f))
            out["recon_from_llm"] = npy(tok.reconstruct_from_llm_tokens(toks_f))
        l2, l1 = tok.compute_reconstruction_error(x)
        out["recon_err"] = np.asarray([float(l2), float(l1)], dtype=np.float64)
        sd = os.path.join(HERE, f"{name}_pretrained")
        tok.save_pretrained(sd)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
    print(name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})
    return tok


def bpe_case(ref, name, *, bpe_vocab_size, fit_batches, fit_seed0, max_sequences=None, vocab_size=256):
    """BEASTBsplineBPETokenizer on the bimanual config: train (HF byte-level BPE under the hood),
    save, encode to ragged ids, decode back."""
    base = ref.BEASTBsplineTokenizer.__new__(ref.BEASTBsplineTokenizer)
    cfg = dict(num_dof=14, num_basis=10, seq_len=50, vocab_size=vocab_size, gripper_zero_order=True,
               gripper_indices=[6, 13], device="cpu")
    base = ref.BEASTBsplineTokenizer(**cfg)
    g2 = dict(np.load(os.path.join(HERE, "cfg2_d14.npz")))
    base.w_min.copy_(torch.from_numpy(g2["w_min_fit"]))
    base.w_max.copy_(torch.from_numpy(g2["w_max_fit"]))
    tok = ref.BEASTBsplineBPETokenizer.from_beast(base, bpe_vocab_size=bpe_vocab_size)
    loader = SyntheticLoader(fit_batches, 32, 50, 14, seed0=fit_seed0)
    state = tok.fit_from_trajectories(loader, show_progress=False, max_sequences=max_sequences)
    sd = os.path.join(HERE, f"{name}_pretrained")
    tok.save_pretrained(sd)
    x = torch.from_numpy(g2["trajs"])
    ids, pd, mp = tok.encode(x, return_mp_tokens=True)
    out = {
        "min_token": np.int64(state.min_token), "max_token": np.int64(state.max_token),
        "fit_batches": np.int64(fit_batches), "fit_seed0": np.int64(fit_seed0),
        "max_sequences": np.int64(-1 if max_sequences is None else max_sequences),
        "bpe_vocab_size": np.int64(bpe_vocab_size),
        "mp_tokens": npy(mp), "params": npy(pd["params"]),
        "ids_flat": np.asarray([i for row in ids for i in row], dtype=np.int64),
        "ids_len": np.asarray([len(r) for r in ids], dtype=np.int64),
        "bpe_to_mp": npy(tok.bpe_to_mp_tokens(ids)),
        "recon": npy(tok.reconstruct_traj(ids)),
        "decode": npy(tok.decode(ids)),
    }
    # the training corpus (MP tokens of the loader) so that trainers can be compared without refitting
    seqs = []
    for b in loader:
        t, _ = tok.encode_to_mp_tokens(b["actions"])
        seqs.append(npy(t))
    corpus = np.concatenate(seqs, 0)
    if max_sequences is not None:
        corpus = corpus[:max_sequences]
    out["corpus_bins"] = corpus.astype(np.uint8) if corpus.max() < 256 else corpus.astype(np.uint16)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
    print(name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


COND_ORDERS = [(1, 0), (2, 0), (0, 1), (0, 2), (1, 1), (2, 2)]
COND_ODD = dict(cfg=dict(num_dof=5, num_basis=8, seq_len=33, vocab_size=1000, degree_p=3, gripper_zero_order=True,
                         gripper_indices=[0]), orders=[(2, 1), (1, 2)], batch=9, seed=41, custom_T=11)


def cond_case(ref, name, *, batch=16, seed=31, custom_T=37, cfg=None, orders=None):
    """Non-zero init/end condition orders (MP_lite_PyTorch/mp_pytorch/mp/uni_bspline.py:500-537):
    the first/last control points of every JOINT spline are pinned to the trajectory's boundary
    position (order 1) and velocity (order 2).  The reference keeps those boundary control points
    as state of its MP object, so reconstruct_traj uses the ones of the LAST fit — recorded here
    with a second encode on other data in between ("stale" outputs)."""
    cfg = dict(cfg or dict(num_dof=14, num_basis=10, seq_len=50, vocab_size=256, gripper_zero_order=True,
                           gripper_indices=[6, 13]), device="cpu")
    orders = orders or COND_ORDERS
    x = synth(batch, cfg["seq_len"], cfg["num_dof"], seed)
    x2 = synth(batch, cfg["seq_len"], cfg["num_dof"], seed + 1)
    g = torch.Generator().manual_seed(seed + 99)
    out = {"trajs": npy(x), "trajs_other": npy(x2),
           "orders": np.asarray(orders, dtype=np.int64)}
    for io, eo in orders:
        tok = ref.BEASTBsplineTokenizer(init_cond_order=io, end_cond_order=eo, **cfg)
        k = f"o{io}{eo}_"
        out[k + "phi_joint"] = npy(tok.mp.basis_gn.basis(tok.times))
        out[k + "knots_joint"] = npy(tok.mp.basis_gn.knots_vec)
        tok.update_weights_bounds(x)
        out[k + "w_min"], out[k + "w_max"] = npy(tok.w_min).copy(), npy(tok.w_max).copy()
        toks, pd = tok.encode(x)
        out[k + "tokens"], out[k + "params"] = npy(toks), npy(pd["params"])
        for key in ("init_pos", "init_vel", "end_pos", "end_vel"):
            if pd[key] is not None:
                out[k + key] = npy(pd[key])
        out[k + "recon"] = npy(tok.reconstruct_traj(toks))
        init_p = x[:, 0, :] + 0.001
        out[k + "recon_initp"] = npy(tok.reconstruct_traj(toks, init_p=init_p))
        tt = torch.sort(torch.rand(batch, custom_T, generator=g) * float(tok.duration), dim=1)[0]
        tt[:, 0] = 0.0
        tt[:, -1] = float(tok.times[-1])
        out[k + "custom_times"] = npy(tt)
        out[k + "recon_custom_times"] = npy(tok.reconstruct_traj(toks, times=tt))
        ctoks, _ = tok.encode_continuous(x)
        out[k + "cont_tokens"] = npy(ctoks)
        tok.encode(x2)                                    # boundary state now belongs to x2
        out[k + "recon_stale"] = npy(tok.reconstruct_traj(toks))
        l2, l1 = (float(v) for v in tok.compute_reconstruction_error(x))
        out[k + "recon_err"] = np.asarray([l2, l1], dtype=np.float64)
    out["init_p"] = npy(x[:, 0, :] + 0.001)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
    print(name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


def main():
    ref = import_reference()
    import tokenizers
    if "--only-cond" in sys.argv:                         # add the newest cases without touching the others
        cond_case(ref, "cond_orders")
        cond_case(ref, "cond_odd", **COND_ODD)
        return
    with open(os.path.join(HERE, "VERSIONS.json"), "w") as f:
        json.dump({"torch": torch.__version__, "numpy": np.__version__,
                   "tokenizers": tokenizers.__version__,
                   "reference": "Dont4rootMe/beast_tokenizer @ /root/reference"}, f, indent=1)
    # BASELINE.json configs[0]: gripper_indices=[6] is ignored (gripper_zero_order False)
    spline_case(ref, "cfg1_d7", num_dof=7, num_basis=10, seq_len=50, vocab_size=256,
                gripper_indices=[6], batch=32, seed=0, fit_batches=20, fit_seed0=100, custom_T=23)
    # configs[1] (bimanual) at a size the reference finishes in seconds
    spline_case(ref, "cfg2_d14", num_dof=14, num_basis=10, seq_len=50, vocab_size=256,
                gripper_zero_order=True, gripper_indices=[6, 13], llm_vocab_size=32000,
                batch=96, seed=2, fit_batches=100, fit_seed0=1, custom_T=120)
    # a ragged / odd shape: odd DoF count, gripper first, cubic, different T / nb / V
    spline_case(ref, "odd_d5", num_dof=5, num_basis=8, seq_len=33, vocab_size=1000, degree_p=3,
                gripper_zero_order=True, gripper_indices=[0], batch=17, seed=11,
                fit_batches=12, fit_seed0=300, custom_T=9)
    # the CLI-default degenerate shape (train/train_beast.py:34-36): nb > T, degree 0
    spline_case(ref, "cli_default", num_dof=32, num_basis=50, seq_len=10, vocab_size=1000, degree_p=0,
                batch=8, seed=21, fit_batches=6, fit_seed0=400)
    # BPE on top of the bimanual tokenizer (configs[3]/[4] at a size the reference finishes in seconds)
    bpe_case(ref, "bpe_d14", bpe_vocab_size=640, fit_batches=40, fit_seed0=1000)
    bpe_case(ref, "bpe_d14_small", bpe_vocab_size=400, fit_batches=8, fit_seed0=2000, max_sequences=200)
    # a 1000-bin tokenizer (the CLI default vocab): shifted bins are codepoints up to U+03E7 (2-byte UTF-8)
    bpe_case(ref, "bpe_v1000", bpe_vocab_size=1600, fit_batches=24, fit_seed0=3000, vocab_size=1000)
    # init/end condition orders 1 and 2 on the bimanual config
    cond_case(ref, "cond_orders")
    # ... and on an odd geometry (cubic, gripper first, 1000 bins, T = 33): generic kernels on the GPU side
    cond_case(ref, "cond_odd", **COND_ODD)


if __name__ == "__main__":
    main()
