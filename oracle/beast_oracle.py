"""CPU oracle for the BEAST B-spline tokenizer hot path (numpy, fp32 op-by-op).

TEST INFRASTRUCTURE ONLY.  This module restates the reference's algorithm so
that the CUDA path can be checked against it.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import it; the product package `beast_tokenizer_b200` never does.

Parity status: the reference (Dont4rootMe/beast_tokenizer) ships no tests or
golden vectors for this path ("parity unpinned" upstream).  The
oracle is therefore pinned to outputs of the LIVE reference run in the build
container: `tests/golden/make_golden.py` imports the reference and commits
its tokens / coefficients / trajectories / bounds as fixtures, and
`tests/test_oracle_golden.py` checks every function below against them.

All citations are relative to /root/reference.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------
# times / phase / knots / basis
# --------------------------------------------------------------------------
def linspace_f32(start: float, end: float, steps: int) -> np.ndarray:
    """fp32 `torch.linspace(start, end, steps)` as used for `self.times`
    (beast/beast_bspline_tokenizer.py:113 -> MP_lite_PyTorch/mp_pytorch/util/
    util_matrix.py:116) and for the knot vector (basis_gn/uni_bspline_basis.py:
    50-51).  ATen computes step=(end-start)/(steps-1) in fp32 and fills the
    first half as start+step*i and the second half as end-step*(steps-1-i)."""
    if steps == 1:
        return np.array([start], dtype=F32)
    s, e = F32(start), F32(end)
    step = F32((e - s) / F32(steps - 1))
    idx = np.arange(steps)
    half = steps // 2
    # ATen's CPU kernel evaluates start + step*i / end - step*j with a fused
    # multiply-add; fp32*fp32 is exact in fp64, so one fp64 op rounded to fp32
    # reproduces it (checked bit-for-bit against torch.linspace in the tests).
    lo = (np.float64(s) + np.float64(step) * idx).astype(F32)
    hi = (np.float64(e) - np.float64(step) * (steps - 1 - idx)).astype(F32)
    return np.where(idx < half, lo, hi).astype(F32)


def linear_phase(times: np.ndarray, tau: float, delay: float = 0.0) -> np.ndarray:
    """`clip((t - delay) / tau, 0, 1)` in fp32
    (MP_lite_PyTorch/mp_pytorch/phase_gn/linear_phase.py:22-23; tau/delay are
    fp32 buffers, phase_generator.py:41-42)."""
    t = np.asarray(times, dtype=F32)
    ph = ((t - F32(delay)).astype(F32) / F32(tau)).astype(F32)
    return np.clip(ph, F32(0), F32(1)).astype(F32)


def knot_vector(num_basis: int, degree_p: int) -> np.ndarray:
    """Clamped uniform knots for init/end condition order 0
    (basis_gn/uni_bspline_basis.py:40-55): num_ctrlp = num_basis,
    `cat[zeros(p), linspace(0, 1, num_ctrlp + 1 - p), ones(p)]`."""
    num_knots = degree_p + 1 + num_basis
    inner = num_knots - 2 * degree_p
    return np.concatenate([
        np.zeros(degree_p, dtype=F32),
        linspace_f32(0.0, 1.0, inner),
        np.ones(degree_p, dtype=F32),
    ]).astype(F32)


def _basis_function(i: int, k: int, knots: np.ndarray, u: np.ndarray, num_ctrlp: int):
    """Cox-de Boor recursion, same fp32 op order as
    basis_gn/uni_bspline_basis.py:82-113 (last basis closed on the right,
    `0/0 -> 0` through the `denom == 0` checks)."""
    if k == 0:
        if i == num_ctrlp - 1:
            b0 = (u >= knots[i]) & (u <= knots[i + 1])
        else:
            b0 = (u >= knots[i]) & (u < knots[i + 1])
        return b0.astype(F32)
    denom1 = F32(knots[i + k] - knots[i])
    if denom1 == 0:
        term1 = None
    else:
        term1 = (((u - knots[i]).astype(F32) / denom1).astype(F32)
                 * _basis_function(i, k - 1, knots, u, num_ctrlp)).astype(F32)
    denom2 = F32(knots[i + k + 1] - knots[i + 1])
    if denom2 == 0:
        term2 = None
    else:
        term2 = (((knots[i + k + 1] - u).astype(F32) / denom2).astype(F32)
                 * _basis_function(i + 1, k - 1, knots, u, num_ctrlp)).astype(F32)
    if term1 is None and term2 is None:
        return np.zeros_like(u, dtype=F32)
    if term1 is None:
        return term2
    if term2 is None:
        return term1
    return (term1 + term2).astype(F32)


def bspline_basis(times: np.ndarray, tau: float, num_basis: int, degree_p: int) -> np.ndarray:
    """Phi[..., T, num_basis] (basis_gn/uni_bspline_basis.py:59-80)."""
    u = linear_phase(times, tau)
    knots = knot_vector(num_basis, degree_p)
    cols = [_basis_function(i, degree_p, knots, u, num_basis) for i in range(num_basis)]
    return np.stack(cols, axis=-1).astype(F32)


# --------------------------------------------------------------------------
# slot layout
# --------------------------------------------------------------------------
def slot_layout(num_dof: int, gripper_zero_order: bool, gripper_indices):
    """joint_indices then gripper_indices
    (beast/beast_bspline_tokenizer.py:56-70, 351-358)."""
    if gripper_indices is None or not gripper_zero_order:
        gripper_indices = []
    grip = sorted(gripper_indices)
    joint = sorted(set(range(num_dof)) - set(grip))
    return joint, grip


# --------------------------------------------------------------------------
# fit (ridge normal equations)
# --------------------------------------------------------------------------
def fit_mp_literal(phi: np.ndarray, trajs: np.ndarray, reg: float = 1e-9) -> np.ndarray:
    """The reference's algorithm as written: dense block-diagonal basis
    [B, d*T, d*nb], A = Phi_m^T Phi_m + reg*I, B = Phi_m^T y, batched fp32 LU
    solve (MP_lite_PyTorch/mp_pytorch/mp/uni_bspline.py:559-586,
    basis_gn/uni_bspline_basis.py:349-356).  O(B * (d*nb)^3): used as the timed
    CPU baseline and for small parity cases.

    phi [T, nb] fp32, trajs [B, T, d] fp32 -> params [B, d*nb] (d-major)."""
    trajs = np.asarray(trajs, dtype=F32)
    Bn, T, d = trajs.shape
    nb = phi.shape[1]
    phi_m = np.zeros((d * T, d * nb), dtype=F32)
    for i in range(d):
        phi_m[i * T:(i + 1) * T, i * nb:(i + 1) * nb] = phi
    phi_b = np.broadcast_to(phi_m, (Bn, d * T, d * nb))
    A = np.einsum('bki,bkj->bij', phi_b, phi_b, optimize=True).astype(F32)
    A = A + (np.eye(d * nb, dtype=F32) * F32(reg))[None]
    y = np.transpose(trajs, (0, 2, 1)).reshape(Bn, d * T)
    rhs = np.einsum('bki,bk->bi', phi_b, y, optimize=True).astype(F32)
    return np.linalg.solve(A.astype(F32), rhs[..., None].astype(F32))[..., 0].astype(F32)


def fit_mp(phi: np.ndarray, trajs: np.ndarray, reg: float = 1e-9) -> np.ndarray:
    """Same normal equations solved per DoF block (the blocks of the
    block-diagonal system are independent and identical for every trajectory,
    so partial-pivoting LU never leaves a block).  fp32 throughout, like the
    reference.  phi [T, nb], trajs [B, T, d] -> [B, d*nb]."""
    trajs = np.asarray(trajs, dtype=F32)
    Bn, T, d = trajs.shape
    nb = phi.shape[1]
    A = (phi.T @ phi).astype(F32) + np.eye(nb, dtype=F32) * F32(reg)
    rhs = np.einsum('tk,btd->kbd', phi, trajs).astype(F32).reshape(nb, Bn * d)
    w = np.linalg.solve(A.astype(F32), rhs).astype(F32)          # [nb, B*d]
    return np.transpose(w.reshape(nb, Bn, d), (1, 2, 0)).reshape(Bn, d * nb).astype(F32)


def compute_weights(trajs, times, tau, num_basis, degree_p, joint_idx, grip_idx,
                    literal: bool = False) -> np.ndarray:
    """`compute_weights` / the fit half of `encode`
    (beast/beast_bspline_tokenizer.py:344-360, 403-414): joints with degree_p,
    grippers with degree 0, concatenated joints-first, (d t) layout."""
    trajs = np.asarray(trajs, dtype=F32)
    fit = fit_mp_literal if literal else fit_mp
    phi_j = bspline_basis(times, tau, num_basis, degree_p)
    w = fit(phi_j, trajs[..., joint_idx])
    if len(grip_idx) > 0:
        phi_g = bspline_basis(times, tau, num_basis, 0)
        wg = fit(phi_g, trajs[..., grip_idx])
        w = np.concatenate([w, wg], axis=-1)
    return w.astype(F32)


# --------------------------------------------------------------------------
# quantise / dequantise  (bit-exact contract)
# --------------------------------------------------------------------------
def continuous_to_discrete(x, min_val, max_val, num_bins: int) -> np.ndarray:
    """beast/utils.py:4-17 — each step a separately rounded fp32 op;
    `torch.round` is round-half-to-even == np.rint; result int64."""
    x = np.asarray(x, dtype=F32)
    min_val = np.asarray(min_val, dtype=F32)
    max_val = np.asarray(max_val, dtype=F32)
    scale = np.maximum((max_val - min_val).astype(F32), F32(1e-8)).astype(F32)
    n = ((x - min_val).astype(F32) / scale).astype(F32)
    n = np.clip(n, F32(0), F32(1)).astype(F32)
    return np.rint((n * F32(num_bins - 1)).astype(F32)).astype(np.int64)


def discrete_to_continuous(tok, min_val, max_val, num_bins: int) -> np.ndarray:
    """beast/utils.py:20-26 — `float(tok)/(V-1)`, mul, add (two roundings), clamp."""
    min_val = np.asarray(min_val, dtype=F32)
    max_val = np.asarray(max_val, dtype=F32)
    n = (np.asarray(tok).astype(F32) / F32(num_bins - 1)).astype(F32)
    c = (n * (max_val - min_val).astype(F32)).astype(F32)
    c = (c + min_val).astype(F32)
    return np.minimum(np.maximum(c, min_val), max_val).astype(F32)


def normalize_tensor(x, w_min, w_max, norm_min=-1.0, norm_max=1.0) -> np.ndarray:
    """beast/utils.py:29-35."""
    x = np.asarray(x, dtype=F32)
    w_min = np.asarray(w_min, dtype=F32)
    w_max = np.asarray(w_max, dtype=F32)
    c = np.minimum(np.maximum(x, w_min), w_max).astype(F32)
    n = ((c - w_min).astype(F32) / np.maximum((w_max - w_min).astype(F32), F32(1e-8))).astype(F32)
    n = (n * F32(F32(norm_max) - F32(norm_min))).astype(F32)
    return (n + F32(norm_min)).astype(F32)


def denormalize_tensor(n, w_min, w_max, norm_min=-1.0, norm_max=1.0) -> np.ndarray:
    """beast/utils.py:38-44."""
    n = np.asarray(n, dtype=F32)
    w_min = np.asarray(w_min, dtype=F32)
    w_max = np.asarray(w_max, dtype=F32)
    c = np.clip(n, F32(norm_min), F32(norm_max)).astype(F32)
    d = ((c - F32(norm_min)).astype(F32)
         / np.maximum(F32(F32(norm_max) - F32(norm_min)), F32(1e-8))).astype(F32)
    d = (d * (w_max - w_min).astype(F32)).astype(F32)
    return (d + w_min).astype(F32)


# --------------------------------------------------------------------------
# encode / decode / reconstruct
# --------------------------------------------------------------------------
def tokens_from_params(params, w_min, w_max, vocab_size, num_dof, num_basis, offset=0):
    """clamp -> quantise -> 'b (d t) -> b (t d)' -> + offset
    (beast/beast_bspline_tokenizer.py:419-426)."""
    params = np.asarray(params, dtype=F32)
    w_min = np.asarray(w_min, dtype=F32)
    w_max = np.asarray(w_max, dtype=F32)
    clamped = np.minimum(np.maximum(params, w_min), w_max).astype(F32)
    q = continuous_to_discrete(clamped, w_min, w_max, vocab_size)
    Bn = q.shape[0]
    tok = q.reshape(Bn, num_dof, num_basis).transpose(0, 2, 1).reshape(Bn, num_basis * num_dof)
    return (tok + np.int64(offset)).astype(np.int64)


def encode(trajs, times, tau, num_basis, degree_p, joint_idx, grip_idx, w_min, w_max,
           vocab_size, offset=0, literal=False):
    """`BEASTBsplineTokenizer.encode` (beast/beast_bspline_tokenizer.py:399-428),
    without the optional bound update.  Returns (tokens int64 [B, nb*D],
    unclamped params fp32 [B, D*nb])."""
    D = len(joint_idx) + len(grip_idx)
    params = compute_weights(trajs, times, tau, num_basis, degree_p, joint_idx, grip_idx, literal)
    return tokens_from_params(params, w_min, w_max, vocab_size, D, num_basis, offset), params


def decode(tokens, w_min, w_max, vocab_size, num_dof, num_basis, offset=0):
    """`decode` (beast/beast_bspline_tokenizer.py:483-496): - offset,
    'b (t d) -> b (d t)', dequantise."""
    tokens = np.asarray(tokens).astype(np.int64)
    if tokens.ndim == 3:
        tokens = tokens.reshape(tokens.shape[0], -1)
    elif tokens.ndim != 2:
        raise ValueError(f"Unexpected token shape {tokens.shape}")
    tokens = tokens - np.int64(offset)
    Bn = tokens.shape[0]
    t = tokens.reshape(Bn, num_basis, num_dof).transpose(0, 2, 1).reshape(Bn, num_dof * num_basis)
    return discrete_to_continuous(t, w_min, w_max, vocab_size)


def reconstruct_from_params(params, times, tau, num_basis, degree_p, joint_idx, grip_idx,
                            init_p=None, use_init_pos=True):
    """`reconstruct_traj` after decode (beast/beast_bspline_tokenizer.py:503-536
    -> mp/uni_bspline.py:160-166): optional init_p overwrite of the first
    control point of every JOINT slot, Phi.w per slot, scatter to DoF order.
    `times` may be [T] (shared) or [B, T'] (per trajectory)."""
    params = np.array(params, dtype=F32, copy=True)
    Bn = params.shape[0]
    nj, ng = len(joint_idx), len(grip_idx)
    D = nj + ng
    p3 = params.reshape(Bn, D, num_basis)
    if use_init_pos and init_p is not None:
        init_p = np.asarray(init_p, dtype=F32)
        for i, j in enumerate(joint_idx):
            p3[:, i, 0] = init_p[:, j]
    times = np.asarray(times, dtype=F32)
    phi_j = bspline_basis(times, tau, num_basis, degree_p)          # [T,nb] or [B,T,nb]
    if phi_j.ndim == 2:
        joint_pos = np.einsum('tk,bdk->btd', phi_j, p3[:, :nj]).astype(F32)
    else:
        joint_pos = np.einsum('btk,bdk->btd', phi_j, p3[:, :nj]).astype(F32)
    T = joint_pos.shape[1]
    pos = np.zeros((Bn, T, D), dtype=F32)
    for i, j in enumerate(joint_idx):
        pos[..., j] = joint_pos[..., i]
    if ng > 0:
        phi_g = bspline_basis(times, tau, num_basis, 0)
        if phi_g.ndim == 2:
            gp = np.einsum('tk,bdk->btd', phi_g, p3[:, nj:]).astype(F32)
        else:
            gp = np.einsum('btk,bdk->btd', phi_g, p3[:, nj:]).astype(F32)
        for i, j in enumerate(grip_idx):
            pos[..., j] = gp[..., i]
    return pos


def reconstruct_traj(tokens, times, tau, num_basis, degree_p, joint_idx, grip_idx, w_min, w_max,
                     vocab_size, offset=0, init_p=None, use_init_pos=True):
    D = len(joint_idx) + len(grip_idx)
    params = decode(tokens, w_min, w_max, vocab_size, D, num_basis, offset)
    return reconstruct_from_params(params, times, tau, num_basis, degree_p, joint_idx, grip_idx,
                                   init_p, use_init_pos)


# --------------------------------------------------------------------------
# init / end condition orders 1 and 2 (joint spline only; goal basis and
# end order -1 are not reachable through BEASTBsplineTokenizer)
# --------------------------------------------------------------------------
def boundary_state(trajs_joint, times, init_order: int, end_order: int) -> dict:
    """Boundary position / finite-difference velocity the reference reads off
    the trajectory itself (mp/uni_bspline.py:507-531): init_pos = y[0],
    init_vel = (y[1]-y[0]) * (1/dt), end_pos = y[-1] (made RELATIVE to
    init_pos when init_order != 0, :90-91), end_vel = (y[-1]-y[-2]) * (1/dt),
    dt = times[1]-times[0]."""
    y = np.asarray(trajs_joint, dtype=F32)
    t = np.asarray(times, dtype=F32)
    if t.ndim == 1:
        t = np.broadcast_to(t, (y.shape[0], t.shape[0]))
    inv_dt = (F32(1) / (t[:, 1] - t[:, 0]).astype(F32)).astype(F32)[:, None]
    st = {"init_pos": None, "init_vel": None, "end_pos": None, "end_vel": None}
    if init_order != 0:
        st["init_pos"] = y[:, 0, :].copy()
        st["init_vel"] = ((y[:, 1, :] - y[:, 0, :]).astype(F32) * inv_dt).astype(F32)
    if end_order != 0:
        end_pos = y[:, -1, :].copy()
        if st["init_pos"] is not None:
            end_pos = (end_pos - st["init_pos"]).astype(F32)
        st["end_pos"] = end_pos
        st["end_vel"] = ((y[:, -1, :] - y[:, -2, :]).astype(F32) * inv_dt).astype(F32)
    return st


def boundary_ctrl_points(st: dict, tau: float, knots: np.ndarray, degree_p: int, num_ctrlp: int,
                         init_order: int, end_order: int):
    """Pinned control points (basis_gn/uni_bspline_basis.py:192-229 with
    init_pos = 0 as passed at mp/uni_bspline.py:76-78, and :231-274):
      init: [0, ((init_vel*tau) * (k[1+p]-k[1])) / p + 0]
      end : [end_pos - ((end_vel*tau) * (k[nc-1+p]-k[nc-1])) / p, end_pos]
    each [B, d, order]; None for order 0."""
    p_init = p_end = None
    if init_order != 0:
        zero = np.zeros_like(st["init_pos"], dtype=F32)
        p_init = zero[..., None]
        if init_order == 2:
            dk = F32(knots[1 + degree_p] - knots[1])
            v = ((((st["init_vel"] * F32(tau)).astype(F32) * dk).astype(F32) / F32(degree_p)).astype(F32)
                 + zero).astype(F32)
            p_init = np.concatenate([p_init, v[..., None]], axis=-1)
    if end_order != 0:
        ep = st["end_pos"]
        p_end = ep[..., None]
        if end_order == 2:
            dk = F32(knots[num_ctrlp - 1 + degree_p] - knots[num_ctrlp - 1])
            v = (ep - (((st["end_vel"] * F32(tau)).astype(F32) * dk).astype(F32) / F32(degree_p)).astype(F32)
                 ).astype(F32)
            p_end = np.concatenate([v[..., None], p_end], axis=-1)
    return p_init, p_end


def fit_joint_with_conditions(trajs_joint, times, tau, num_basis, degree_p, init_order, end_order,
                              reg: float = 1e-9):
    """learn_mp_params_from_trajs with pinned boundary control points
    (mp/uni_bspline.py:471-602): subtract the pinned points' contribution
    (and init_pos) from the trajectory, ridge-fit the remaining num_basis
    columns of the [T, num_ctrlp] basis.  Returns (params [B, d*nb], state)
    where state carries what the reference's MP object keeps for the next
    reconstruct: init_pos, ctrl_init, ctrl_end (+ the returned dict entries)."""
    y = np.asarray(trajs_joint, dtype=F32)
    Bn, T, d = y.shape
    nc = num_basis + init_order + end_order
    knots = knot_vector(nc, degree_p)
    phi_full = bspline_basis(times, tau, nc, degree_p)                    # [T, nc]
    st = boundary_state(y, times, init_order, end_order)
    p_init, p_end = boundary_ctrl_points(st, tau, knots, degree_p, nc, init_order, end_order)
    parts = []
    if p_init is not None:
        parts.append(p_init)
    parts.append(np.zeros((Bn, d, num_basis), dtype=F32))
    if p_end is not None:
        parts.append(p_end)
    dummy = np.concatenate(parts, axis=-1)                                # [B, d, nc]
    pos_det = np.einsum('tk,bdk->btd', phi_full, dummy).astype(F32)
    if init_order != 0:
        pos_det = (pos_det + st["init_pos"][:, None, :]).astype(F32)
    pos_w = (y - pos_det).astype(F32)
    phi_learn = phi_full[:, init_order:nc - end_order]
    params = fit_mp(phi_learn, pos_w, reg)
    state = dict(st)
    state["ctrl_init"], state["ctrl_end"] = p_init, p_end
    # the dict learn_mp_params_from_trajs returns (:597-602): end_pos made absolute again
    state["ret_end_pos"] = ((st["end_pos"] + st["init_pos"]).astype(F32)
                            if st["end_pos"] is not None and st["init_pos"] is not None else st["end_pos"])
    return params, state


def eval_joint_with_conditions(params_joint, state, times, tau, num_basis, degree_p, init_order, end_order):
    """get_traj_pos with pinned control points (mp/uni_bspline.py:126-166):
    cat[ctrl_init, params, ctrl_end] against the full basis, + init_pos.
    params_joint [B, d, nb] -> [B, T', d]."""
    nc = num_basis + init_order + end_order
    parts = []
    if state.get("ctrl_init") is not None:
        parts.append(state["ctrl_init"])
    parts.append(np.asarray(params_joint, dtype=F32))
    if state.get("ctrl_end") is not None:
        parts.append(state["ctrl_end"])
    full = np.concatenate(parts, axis=-1)
    phi = bspline_basis(np.asarray(times, dtype=F32), tau, nc, degree_p)
    sub = 'tk,bdk->btd' if phi.ndim == 2 else 'btk,bdk->btd'
    pos = np.einsum(sub, phi, full).astype(F32)
    if state.get("init_pos") is not None:
        pos = (pos + state["init_pos"][:, None, :]).astype(F32)
    return pos


def compute_weights_cond(trajs, times, tau, num_basis, degree_p, joint_idx, grip_idx, init_order, end_order):
    """compute_weights / the fit half of encode for non-zero condition orders:
    conditions apply to the joint MP only (beast_bspline_tokenizer.py:79-96)."""
    trajs = np.asarray(trajs, dtype=F32)
    w, state = fit_joint_with_conditions(trajs[..., joint_idx], times, tau, num_basis, degree_p,
                                         init_order, end_order)
    if len(grip_idx) > 0:
        wg = fit_mp(bspline_basis(times, tau, num_basis, 0), trajs[..., grip_idx])
        w = np.concatenate([w, wg], axis=-1)
    return w.astype(F32), state


def reconstruct_from_params_cond(params, state, times, tau, num_basis, degree_p, joint_idx, grip_idx,
                                 init_order, end_order, init_p=None, use_init_pos=True):
    """reconstruct_traj after decode (beast_bspline_tokenizer.py:503-536) when the joint MP
    carries pinned control points from its last fit (`state`)."""
    params = np.array(params, dtype=F32, copy=True)
    Bn = params.shape[0]
    nj, ng = len(joint_idx), len(grip_idx)
    D = nj + ng
    p3 = params.reshape(Bn, D, num_basis)
    if use_init_pos and init_p is not None:
        init_p = np.asarray(init_p, dtype=F32)
        for i, j in enumerate(joint_idx):
            p3[:, i, 0] = init_p[:, j]
    joint_pos = eval_joint_with_conditions(p3[:, :nj], state, times, tau, num_basis, degree_p,
                                           init_order, end_order)
    pos = np.zeros((Bn, joint_pos.shape[1], D), dtype=F32)
    for i, j in enumerate(joint_idx):
        pos[..., j] = joint_pos[..., i]
    if ng > 0:
        phi_g = bspline_basis(np.asarray(times, dtype=F32), tau, num_basis, 0)
        sub = 'tk,bdk->btd' if phi_g.ndim == 2 else 'btk,bdk->btd'
        gp = np.einsum(sub, phi_g, p3[:, nj:]).astype(F32)
        for i, j in enumerate(grip_idx):
            pos[..., j] = gp[..., i]
    return pos


# --------------------------------------------------------------------------
# bounds
# --------------------------------------------------------------------------
def bounds_minmax(params):
    """`update_weights_bounds` (beast/beast_bspline_tokenizer.py:377-378)."""
    params = np.asarray(params, dtype=F32)
    return params.min(axis=0), params.max(axis=0)


def bounds_expand(params, w_min, w_max, hyst: float = 1e-4):
    """`update_weights_bounds_per_batch` (beast/beast_bspline_tokenizer.py:
    379-389): replace only where batch_min < w_min - 1e-4 / batch_max > w_max + 1e-4
    (the 1e-4 is added to an fp32 tensor, i.e. rounded to fp32)."""
    params = np.asarray(params, dtype=F32)
    w_min = np.array(w_min, dtype=F32, copy=True)
    w_max = np.array(w_max, dtype=F32, copy=True)
    bmin, bmax = params.min(axis=0), params.max(axis=0)
    smaller = bmin < (w_min - F32(hyst)).astype(F32)
    larger = bmax > (w_max + F32(hyst)).astype(F32)
    w_min[smaller] = bmin[smaller]
    w_max[larger] = bmax[larger]
    return w_min, w_max


def bounds_quantile(params, lo: float = 0.01, hi: float = 0.99):
    """`fit_parameters` (beast/beast_bspline_tokenizer.py:211-220):
    np.quantile(params, 0.01 / 0.99, axis=0) on fp32 input (result fp32),
    copied into the fp32 buffers."""
    params = np.asarray(params, dtype=F32)
    return (np.quantile(params, lo, 0).astype(F32), np.quantile(params, hi, 0).astype(F32))
