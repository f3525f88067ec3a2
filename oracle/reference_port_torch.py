"""Op-for-op CPU port of the reference's encode / reconstruct_traj in torch (fp32, CPU).

TEST INFRASTRUCTURE ONLY (see oracle/beast_oracle.py for the rules).  The numpy oracle is the
parity checker; this module exists because the reference's CPU path *is* a sequence of torch
ops (einsum -> bmm, linalg.solve -> batched LU, elementwise clamp/round), so the honest CPU
baseline for bench.py is the same sequence of torch ops on the same host cores — including the
work the reference repeats on every call (basis rebuilt twice per fit, dense block-diagonal
basis, one (D*nb)^2 LU per trajectory).  Pinned in tests/test_oracle_golden.py against the live
reference's golden vectors: basis bit-identical, coefficients / trajectories within 1e-6 normwise (the
batched LU's thread schedule differs between runs), tokens equal on >= 99.9 % of the positions.

Citations relative to /root/reference.
"""
import math

import torch


class ReferencePort:
    def __init__(self, num_dof=14, num_basis=10, duration=2 * math.pi, seq_len=50, vocab_size=256, degree_p=4,
                 gripper_zero_order=False, gripper_indices=None, llm_vocab_size=None):
        # beast/beast_bspline_tokenizer.py:55-70, 113-116
        if gripper_indices is None or not gripper_zero_order:
            gripper_indices = []
        self.grip = sorted(gripper_indices)
        self.joint = sorted(set(range(num_dof)) - set(self.grip))
        self.D, self.nb, self.V, self.p = num_dof, num_basis, vocab_size, degree_p
        self.tau = torch.tensor(duration, dtype=torch.float32)
        self.delay = torch.tensor(0.0, dtype=torch.float32)
        self.times = torch.linspace(0, duration, seq_len)
        self.w_min = -0.02 * torch.ones(num_dof * num_basis)
        self.w_max = 0.02 * torch.ones(num_dof * num_basis)
        self.offset = 0 if llm_vocab_size is None else llm_vocab_size - vocab_size
        self.knots = {d: self._knots(d) for d in {degree_p, 0}}

    def _knots(self, p):
        # basis_gn/uni_bspline_basis.py:46-55
        inner = torch.linspace(0, 1, self.nb + 1 - p, dtype=torch.float32)
        return torch.cat([torch.zeros(p), inner, torch.ones(p)])

    def _bf(self, i, k, kn, u):
        # basis_gn/uni_bspline_basis.py:82-113 (recursive, recomputed on every call like upstream)
        if k == 0:
            if i == self.nb - 1:
                b0 = torch.where((u >= kn[i]) & (u <= kn[i + 1]), 1, 0)
            else:
                b0 = torch.where((u >= kn[i]) & (u < kn[i + 1]), 1, 0)
            return torch.as_tensor(b0, dtype=torch.float32)
        d1 = kn[i + k] - kn[i]
        t1 = 0.0 if d1 == 0 else (u - kn[i]) / d1 * self._bf(i, k - 1, kn, u)
        d2 = kn[i + k + 1] - kn[i + 1]
        t2 = 0.0 if d2 == 0 else (kn[i + k + 1] - u) / d2 * self._bf(i + 1, k - 1, kn, u)
        return t1 + t2

    def basis(self, times, p):
        # linear_phase.py:22-23 + uni_bspline_basis.py:59-80
        u = torch.clip((times - self.delay[..., None]) / self.tau[..., None], 0, 1)
        return torch.stack([self._bf(i, p, self.knots[p], u) for i in range(self.nb)], dim=-1)

    def _learn(self, times, trajs, p, reg=1e-9):
        # mp/uni_bspline.py:539-586 with condition orders 0
        d = trajs.shape[-1]
        Bn, Tn = times.shape
        basis_single = self.basis(times, p)                                   # :539 (used for pos_det = 0)
        dummy = torch.zeros(Bn, d, self.nb)
        pos_det = torch.einsum('...ik,...jk->...ij', basis_single, dummy)     # :544
        pos_det = torch.einsum('...ij->...ji', pos_det).reshape(Bn, -1)
        bs = self.basis(times, p)                                             # :559 -> basis_multi_dofs rebuilds it
        multi = torch.zeros(Bn, d * Tn, d * self.nb)                          # uni_bspline_basis.py:349-356
        for i in range(d):
            multi[..., i * Tn:(i + 1) * Tn, i * self.nb:(i + 1) * self.nb] = bs
        A = torch.einsum('...ki,...kj->...ij', multi, multi)                  # :564
        A += torch.eye(d * self.nb) * reg                                     # :566
        y = torch.einsum('...ij->...ji', trajs).reshape(Bn, -1)               # :572-575
        pos_w = y - pos_det                                                   # :578
        rhs = torch.einsum('...ki,...k->...i', multi, pos_w)                  # :583
        return torch.linalg.solve(A, rhs)                                     # :586

    @torch.no_grad()
    def compute_weights(self, trajs):
        trajs = trajs.to(torch.float32)
        times = self.times[None].expand(trajs.shape[0], -1)
        w = self._learn(times, trajs[..., self.joint], self.p)
        if self.grip:
            w = torch.cat([w, self._learn(times, trajs[..., self.grip], 0)], dim=-1)
        return w

    @torch.no_grad()
    def encode(self, trajs):
        # beast/beast_bspline_tokenizer.py:399-428, beast/utils.py:4-17
        params = self.compute_weights(trajs)
        clamped = torch.clamp(params, min=self.w_min, max=self.w_max)
        scale = torch.clamp(self.w_max - self.w_min, min=1e-8)
        n = torch.clamp((clamped - self.w_min) / scale, 0, 1)
        tok = torch.round(n * (self.V - 1)).to(torch.long)
        Bn = tok.shape[0]
        tok = tok.reshape(Bn, self.D, self.nb).transpose(1, 2).reshape(Bn, -1)
        return tok + self.offset, params

    @torch.no_grad()
    def reconstruct_traj(self, tokens, init_p=None):
        # :483-536, beast/utils.py:20-26, mp/uni_bspline.py:160-166
        Bn = tokens.shape[0]
        tok = (tokens - self.offset).reshape(Bn, self.nb, self.D).transpose(1, 2).reshape(Bn, -1)
        n = tok.float() / (self.V - 1)
        params = torch.clamp(n * (self.w_max - self.w_min) + self.w_min, self.w_min, self.w_max)
        if init_p is not None:
            p3 = params.reshape(Bn, self.D, self.nb).transpose(1, 2).clone()
            for i, j in enumerate(self.joint):
                p3[:, 0, i] = init_p[:, j]
            params = p3.transpose(1, 2).reshape(Bn, -1)
        times = self.times[None].expand(Bn, -1)
        nj = len(self.joint)
        jp = params[..., :nj * self.nb].reshape(Bn, nj, self.nb)
        joint_pos = torch.einsum('...ik,...jk->...ij', self.basis(times, self.p), jp)
        pos = torch.zeros(Bn, joint_pos.shape[1], self.D)
        for i, j in enumerate(self.joint):
            pos[..., j] = joint_pos[..., i]
        if self.grip:
            gp = params[..., nj * self.nb:].reshape(Bn, len(self.grip), self.nb)
            grip_pos = torch.einsum('...ik,...jk->...ij', self.basis(times, 0), gp)
            for i, j in enumerate(self.grip):
                pos[..., j] = grip_pos[..., i]
        return pos
