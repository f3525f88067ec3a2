"""Pin the numpy oracle (oracle/beast_oracle.py) to the golden vectors produced
by the live reference (tests/golden/make_golden.py).  CPU only."""
import math

import numpy as np
import pytest

from oracle import beast_oracle as O
from conftest import load_golden, rel_err

TOL = 1e-5          # north-star tolerance for coefficients / trajectories


def _layout(cfg):
    return O.slot_layout(cfg["num_dof"], cfg["gripper_zero_order"], cfg["gripper_indices"])


def _offset(cfg):
    return 0 if cfg["llm_vocab_size"] is None else cfg["llm_vocab_size"] - cfg["vocab_size"]


def test_times_knots_basis_bit_exact(golden_case):
    name, cfg, g = golden_case
    times = O.linspace_f32(0.0, 2 * math.pi, cfg["seq_len"])
    assert times.dtype == np.float32 and np.array_equal(times, g["times"])
    assert np.array_equal(O.knot_vector(cfg["num_basis"], cfg["degree_p"]), g["knots_joint"])
    phi = O.bspline_basis(times, 2 * math.pi, cfg["num_basis"], cfg["degree_p"])
    assert np.array_equal(phi, g["phi_joint"])
    if "phi_grip" in g:
        assert np.array_equal(O.bspline_basis(times, 2 * math.pi, cfg["num_basis"], 0), g["phi_grip"])
    joint, grip = _layout(cfg)
    assert joint == g["joint_indices"].tolist() and grip == g["gripper_indices"].tolist()


@pytest.mark.parametrize("literal", [False, True])
def test_fit_within_tolerance(golden_case, literal):
    name, cfg, g = golden_case
    if literal and name == "cli_default":
        pytest.skip("1600x1600 literal systems: covered by the per-DoF form")
    joint, grip = _layout(cfg)
    w = O.compute_weights(g["trajs"], g["times"], 2 * math.pi, cfg["num_basis"], cfg["degree_p"],
                          joint, grip, literal=literal)
    assert w.dtype == np.float32 and w.shape == g["params"].shape
    assert rel_err(w, g["params"]) <= TOL


def test_quantise_bit_exact_given_reference_coefficients(golden_case):
    """tokens are bit-exact given identical fp32 coefficients."""
    name, cfg, g = golden_case
    D, nb, V = cfg["num_dof"], cfg["num_basis"], cfg["vocab_size"]
    t = O.tokens_from_params(g["params"], g["w_min_default"], g["w_max_default"], V, D, nb, _offset(cfg))
    assert t.dtype == np.int64 and np.array_equal(t, g["tokens_default"])
    t = O.tokens_from_params(g["params"], g["w_min_fit"], g["w_max_fit"], V, D, nb, _offset(cfg))
    assert np.array_equal(t, g["tokens_fit"])
    t = O.tokens_from_params(g["params"], g["w_min_fit"], g["w_max_fit"], V, D, nb, 0)
    assert np.array_equal(t, g["tokens_fit_nooffset"])


def test_dequantise_bit_exact(golden_case):
    name, cfg, g = golden_case
    D, nb, V = cfg["num_dof"], cfg["num_basis"], cfg["vocab_size"]
    c = O.decode(g["tokens_fit"], g["w_min_fit"], g["w_max_fit"], V, D, nb, _offset(cfg))
    assert c.dtype == np.float32 and np.array_equal(c, g["decode_fit"])


def test_reconstruct_within_tolerance(golden_case):
    name, cfg, g = golden_case
    joint, grip = _layout(cfg)
    args = (2 * math.pi, cfg["num_basis"], cfg["degree_p"], joint, grip, g["w_min_fit"], g["w_max_fit"],
            cfg["vocab_size"], _offset(cfg))
    r = O.reconstruct_traj(g["tokens_fit"], g["times"], *args)
    assert rel_err(r, g["recon_fit"]) <= TOL
    r = O.reconstruct_traj(g["tokens_fit"], g["times"], *args, init_p=g["init_p"])
    assert rel_err(r, g["recon_fit_initp"]) <= TOL
    if "custom_times" in g:
        r = O.reconstruct_traj(g["tokens_fit"], g["custom_times"], *args)
        assert r.shape == g["recon_fit_custom_times"].shape
        assert rel_err(r, g["recon_fit_custom_times"]) <= TOL
    # default +-0.02 bounds
    r = O.reconstruct_traj(g["tokens_default"], g["times"], 2 * math.pi, cfg["num_basis"], cfg["degree_p"],
                           joint, grip, g["w_min_default"], g["w_max_default"], cfg["vocab_size"], _offset(cfg))
    assert rel_err(r, g["recon_default"]) <= TOL


def test_bounds(golden_case):
    name, cfg, g = golden_case
    # min/max and the hysteresis expansion are exact given the reference coefficients
    lo, hi = O.bounds_minmax(g["params"])
    assert np.array_equal(lo, g["w_min_minmax"]) and np.array_equal(hi, g["w_max_minmax"])
    joint, grip = _layout(cfg)
    w_ub = O.compute_weights(g["trajs_ub"], g["times"], 2 * math.pi, cfg["num_basis"], cfg["degree_p"], joint, grip)
    lo2, hi2 = O.bounds_expand(w_ub, lo, hi)
    assert rel_err(lo2, g["w_min_expand"]) <= TOL and rel_err(hi2, g["w_max_expand"]) <= TOL
    # same entries were replaced
    assert np.array_equal(lo2 != lo, g["w_min_expand"] != g["w_min_minmax"])
    assert np.array_equal(hi2 != hi, g["w_max_expand"] != g["w_max_minmax"])


def test_fit_parameters_quantile(golden_case):
    name, cfg, g = golden_case
    from beast_tokenizer_b200.synth import SyntheticLoader
    joint, grip = _layout(cfg)
    ws = []
    for b in SyntheticLoader(int(g["fit_batches"]), 32, cfg["seq_len"], cfg["num_dof"], seed0=int(g["fit_seed0"])):
        ws.append(O.compute_weights(b["actions"].numpy(), g["times"], 2 * math.pi, cfg["num_basis"],
                                    cfg["degree_p"], joint, grip))
    lo, hi = O.bounds_quantile(np.concatenate(ws, 0))
    assert rel_err(lo, g["w_min_fit"]) <= TOL and rel_err(hi, g["w_max_fit"]) <= TOL


def test_continuous_tokens(golden_case):
    name, cfg, g = golden_case
    D, nb = cfg["num_dof"], cfg["num_basis"]
    n = O.normalize_tensor(g["params"], g["w_min_fit"], g["w_max_fit"])
    n = n.reshape(-1, D, nb).transpose(0, 2, 1).reshape(-1, nb * D)
    assert np.array_equal(n, g["cont_tokens_fit"])
    assert int(g["recon_cont_raises"]) == 1      # upstream bug recorded (beast/utils.py:42)


def test_llm_offset_helpers():
    g = dict(np.load(__import__("os").path.join(__import__("conftest").GOLDEN, "cfg2_d14.npz")))
    off = 32000 - 256
    assert np.array_equal(g["tokens_fit_nooffset"] + off, g["tokens_fit"])
    assert np.array_equal(g["llm_tokens"], g["tokens_fit"])
    assert np.array_equal(g["mp_tokens_3d"].reshape(g["tokens_fit"].shape[0], -1), g["tokens_fit_nooffset"])


def test_torch_reference_port_matches_golden(golden_case):
    """The torch CPU port used as bench.py's CPU baseline reproduces the live reference."""
    import torch
    from oracle.reference_port_torch import ReferencePort
    name, cfg, g = golden_case
    port = ReferencePort(num_dof=cfg["num_dof"], num_basis=cfg["num_basis"], seq_len=cfg["seq_len"],
                         vocab_size=cfg["vocab_size"], degree_p=cfg["degree_p"],
                         gripper_zero_order=cfg["gripper_zero_order"], gripper_indices=cfg["gripper_indices"],
                         llm_vocab_size=cfg["llm_vocab_size"])
    assert np.array_equal(port.basis(port.times, cfg["degree_p"]).numpy(), g["phi_joint"])
    port.w_min, port.w_max = torch.from_numpy(g["w_min_fit"]), torch.from_numpy(g["w_max_fit"])
    tok, params = port.encode(torch.from_numpy(g["trajs"]))
    assert rel_err(params.numpy(), g["params"]) <= 1e-6
    assert (tok.numpy() != g["tokens_fit"]).mean() <= 1e-3
    rec = port.reconstruct_traj(torch.from_numpy(g["tokens_fit"]))
    assert rel_err(rec.numpy(), g["recon_fit"]) <= 1e-6
    rec = port.reconstruct_traj(torch.from_numpy(g["tokens_fit"]), init_p=torch.from_numpy(g["init_p"]))
    assert rel_err(rec.numpy(), g["recon_fit_initp"]) <= 1e-6


# ---------------------------------------------------------------- init / end condition orders
COND_CASES = {
    "cond_orders": dict(num_dof=14, num_basis=10, seq_len=50, vocab_size=256, degree_p=4, gripper_indices=[6, 13]),
    "cond_odd": dict(num_dof=5, num_basis=8, seq_len=33, vocab_size=1000, degree_p=3, gripper_indices=[0]),
}
COND_PARAMS = [("cond_orders", o) for o in [(1, 0), (2, 0), (0, 1), (0, 2), (1, 1), (2, 2)]] + \
              [("cond_odd", o) for o in [(2, 1), (1, 2)]]


@pytest.mark.parametrize("case,orders", COND_PARAMS)
def test_condition_orders_oracle_matches_reference(case, orders):
    """Pins the restatement of mp/uni_bspline.py:500-537 + 126-166 (pinned boundary control points)
    against outputs of the live reference, including its stateful reconstruct."""
    io, eo = orders
    cfg = COND_CASES[case]
    g = load_golden(case)
    assert list(orders) in g["orders"].tolist()
    k = f"o{io}{eo}_"
    D, nb, T, V, p = cfg["num_dof"], cfg["num_basis"], cfg["seq_len"], cfg["vocab_size"], cfg["degree_p"]
    tau = 2 * math.pi
    times = O.linspace_f32(0, tau, T)
    joint, grip = O.slot_layout(D, True, cfg["gripper_indices"])
    nc = nb + io + eo
    assert np.array_equal(O.knot_vector(nc, p), g[k + "knots_joint"])
    assert np.array_equal(O.bspline_basis(times, tau, nc, p), g[k + "phi_joint"])
    w, st = O.compute_weights_cond(g["trajs"], times, tau, nb, p, joint, grip, io, eo)
    assert rel_err(w, g[k + "params"]) <= 1e-5
    for key, have in (("init_pos", st["init_pos"]), ("init_vel", st["init_vel"]),
                      ("end_pos", st["ret_end_pos"]), ("end_vel", st["end_vel"])):
        if k + key in g:
            assert np.array_equal(have, g[k + key]), key          # plain fp32 elementwise: bit-exact
        else:
            assert have is None
    lo, hi = g[k + "w_min"], g[k + "w_max"]
    assert rel_err(w.min(0), lo) <= 1e-5 and rel_err(w.max(0), hi) <= 1e-5
    assert np.array_equal(O.tokens_from_params(g[k + "params"], lo, hi, V, D, nb), g[k + "tokens"])
    dec = O.decode(g[k + "tokens"], lo, hi, V, D, nb)
    args = (times, tau, nb, p, joint, grip, io, eo)
    assert rel_err(O.reconstruct_from_params_cond(dec, st, *args), g[k + "recon"]) <= 1e-5
    assert rel_err(O.reconstruct_from_params_cond(dec, st, *args, init_p=g["init_p"]), g[k + "recon_initp"]) <= 1e-5
    r = O.reconstruct_from_params_cond(dec, st, g[k + "custom_times"], tau, nb, p, joint, grip, io, eo)
    assert rel_err(r, g[k + "recon_custom_times"]) <= 1e-5
    # the reference reconstructs with the boundary state of its LAST fit
    _, st_other = O.compute_weights_cond(g["trajs_other"], times, tau, nb, p, joint, grip, io, eo)
    assert rel_err(O.reconstruct_from_params_cond(dec, st_other, *args), g[k + "recon_stale"]) <= 1e-5
    assert rel_err(O.reconstruct_from_params_cond(dec, st, *args), g[k + "recon_stale"]) > 1e-3
    cont = O.normalize_tensor(w, lo, hi).reshape(-1, D, nb).transpose(0, 2, 1).reshape(-1, D * nb)
    assert np.abs(cont - g[k + "cont_tokens"]).max() <= 2.0 * 1e-5 * np.abs(g[k + "params"]).max() / \
        max(float((hi - lo).min()), 1e-8) + 1e-6
