"""CPU ORACLE (test infrastructure, NOT the product): the Cartesian-stage terms in torch fp64
with autograd gradients -- cart_bonded-like springs, Ramachandran and omega terms evaluated
from coordinates.  These terms are stated approximations on both sides (constants in
include/trx_centroid_model.h); the oracle checks that the CUDA kernel computes exactly this
function and its analytic gradient.  PARITY UNPINNED against PyRosetta."""
from __future__ import annotations

import os
import re

import numpy as np
import torch

_HDR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "trx_centroid_model.h")


def _consts():
    src = open(_HDR).read()
    src_nc = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    env = {}
    for line in src_nc.splitlines():
        m = re.match(r"\s*#define\s+(TRX_[A-Z_0-9]+)\s+(.+?)\s*$", line)
        if not m:
            continue
        try:
            env[m.group(1)] = float(eval(m.group(2), {"__builtins__": {}}, env))
        except Exception:
            pass
    rama = re.search(r"TRX_RAMA\[2\]\[TRX_RAMA_NB\]\[5\] = \{(.*?)\};", src, flags=re.S).group(1)
    nums = [float(x) for x in re.findall(r"-?\d+\.\d+", rama)]
    env["RAMA"] = np.array(nums).reshape(2, 5, 5)
    off = re.search(r"TRX_RAMA_OFFSET\[2\] = \{(.*?)\};", src, flags=re.S).group(1)
    env["RAMA_OFFSET"] = np.array([float(x) for x in off.split(",")])
    return env


K = _consts()
N_, CA_, CB_, C_, O_ = 0, 1, 2, 3, 4


def _dih(p1, p2, p3, p4):
    b0, b1, b2 = p1 - p2, p3 - p2, p4 - p3
    b1 = b1 / b1.norm(dim=-1, keepdim=True)
    v = b0 - (b0 * b1).sum(-1, keepdim=True) * b1
    w = b2 - (b2 * b1).sum(-1, keepdim=True) * b1
    return torch.atan2((torch.cross(b1, v, dim=-1) * w).sum(-1), (v * w).sum(-1))


def _ang(a, b, c):
    u, v = a - b, c - b
    return torch.acos(((u * v).sum(-1) / (u.norm(dim=-1) * v.norm(dim=-1))).clamp(-1, 1))


def cart_terms(xyz, aa):
    """xyz (L,5,3) torch float64 [N,CA,CB,C,O]; aa (L,) residue types.  Returns dict of unweighted
    terms: cart (bonded springs), rama, omega."""
    n, ca, cb, c, o = (xyz[:, k] for k in range(5))
    KB, KA, KCB, KPL = K["TRX_CART_KB"], K["TRX_CART_KA"], K["TRX_CART_KCB"], K["TRX_CART_KPL"]
    e = KB * ((ca - n).norm(dim=-1) - K["TRX_B_N_CA"]).pow(2).sum()
    e = e + KB * ((c - ca).norm(dim=-1) - K["TRX_B_CA_C"]).pow(2).sum()
    e = e + KB * ((o - c).norm(dim=-1) - K["TRX_B_C_O"]).pow(2).sum()
    e = e + KB * ((n[1:] - c[:-1]).norm(dim=-1) - K["TRX_B_C_N"]).pow(2).sum()
    e = e + KA * (_ang(n, ca, c) - K["TRX_A_N_CA_C"]).pow(2).sum()
    e = e + KA * (_ang(ca, c, o) - K["TRX_A_CA_C_O"]).pow(2).sum()
    e = e + KA * (_ang(ca[:-1], c[:-1], n[1:]) - K["TRX_A_CA_C_N"]).pow(2).sum()
    e = e + KA * (_ang(o[:-1], c[:-1], n[1:]) - K["TRX_A_O_C_N"]).pow(2).sum()
    e = e + KA * (_ang(c[:-1], n[1:], ca[1:]) - K["TRX_A_C_N_CA"]).pow(2).sum()
    b, cc = ca - n, c - ca
    vcb = K["TRX_CB_A"] * torch.cross(b, cc, dim=-1) + K["TRX_CB_B"] * b + K["TRX_CB_C"] * cc + ca
    e = e + KCB * (cb - vcb).pow(2).sum()
    t = (torch.cross(ca[:-1] - c[:-1], n[1:] - c[:-1], dim=-1) * (o[:-1] - c[:-1])).sum(-1)
    e = e + KPL * t.pow(2).sum()
    # omega tether on residues 0..L-2, rama on 1..L-2 (same definitions as the torsion-space terms)
    omega = _dih(ca[:-1], c[:-1], n[1:], ca[1:])
    dev = omega - np.pi
    dev = dev - 2 * np.pi * torch.floor((dev + np.pi) / (2 * np.pi))
    e_omega = (K["TRX_OMEGA_K"] * (dev / K["TRX_DEG"]).pow(2)).sum()
    phi = _dih(c[:-2], n[1:-1], ca[1:-1], c[1:-1])
    psi = _dih(n[1:-1], ca[1:-1], c[1:-1], n[2:])
    cls = torch.as_tensor((np.asarray(aa)[1:-1] == 14).astype(np.int64))
    R = torch.as_tensor(K["RAMA"])[cls]                      # (L-2, 5, 5)
    dphi = phi[:, None] - R[:, :, 0] * K["TRX_DEG"]
    dpsi = psi[:, None] - R[:, :, 1] * K["TRX_DEG"]
    P = K["TRX_RAMA_FLOOR"] + (R[:, :, 4] * torch.exp(R[:, :, 2] * (torch.cos(dphi) - 1) + R[:, :, 3] * (torch.cos(dpsi) - 1))).sum(-1)
    off = torch.as_tensor(K["RAMA_OFFSET"])[cls]
    e_rama = (-torch.log(P) - off).sum()
    return {"cart": e, "rama": e_rama, "omega": e_omega}


def cart_energy_grad(xyz, aa, w_cart, w_rama, w_omega):
    """numpy in/out: -> (terms dict of floats, grad (L,5,3) of the weighted sum)."""
    x = torch.tensor(np.asarray(xyz, dtype=np.float64), requires_grad=True)
    t = cart_terms(x, aa)
    tot = w_cart * t["cart"] + w_rama * t["rama"] + w_omega * t["omega"]
    tot.backward()
    return {k: float(v.detach()) for k, v in t.items()}, x.grad.numpy()
