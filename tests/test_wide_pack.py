"""CPU test of the WIDE kernel's host-side weight repack (wrnn_wide_pack_host, csrc/wavernn_wide.cuh): decode the packed
per-CTA images with the index arithmetic the kernel's lanes use (pass_tile, pass4, cond_pass), re-run the step dataflow in
numpy float64 and compare the teacher-forced logits with the fp64 oracle.  Pins the algebraic folding, the
[warp][ig][ks][unit] weight layouts, the 176-wide conditioning K space and the row ownership without a GPU."""
import ctypes

import numpy as np
import pytest

from expressive_speech_synthesis_research_b200 import _lib
from oracle import c_oracle, synth

HID, NWORK, KC2 = 512, 128, 176
SV = dict(U1=0, B1=12, BHH1=24, U2=36, B2=48, BHH2=60, U3=72, B3=76, B4=80, B5=84)


def pack(sd, mode):
    C = sd["fc3.weight"].shape[0]
    cfg = _lib.Config(512, 512, 80, 32, C, _lib.MODE[mode], 0)
    L = _lib.lib()
    layout = (ctypes.c_int64 * 8)()
    n = L.wrnn_wide_packed_floats(ctypes.byref(cfg), layout)
    assert n > 0
    keep, w = [], _lib.Weights()
    for field, key in zip(_lib.Weights.FIELDS, _lib.Weights.KEYS):
        a = np.ascontiguousarray(sd[key].numpy(), dtype=np.float32)
        keep.append(a)
        setattr(w, field, a.ctypes.data)
    out = np.zeros(n, dtype=np.float32)
    _lib.check(L.wrnn_wide_pack_host(ctypes.byref(cfg), ctypes.byref(w), out.ctypes.data, n))
    return out.reshape(NWORK, -1), list(layout)


def gate_matrix(img, off):
    """RB = 3 layout -> [cta][gate 3][unit 4][k 512], read the way pass_tile does: lane (ks, u) of warp w, block ig, element
    ii*3 + gate multiplies x[k = 32 w + 2 (4 ig + ii) + ks]."""
    blk = img[:, off:off + 12 * HID].reshape(NWORK, 16, 4, 2, 4, 4, 3).astype(np.float64)   # [cta][w][ig][ks][u][ii][g]
    out = np.zeros((NWORK, 3, 4, HID))
    for w in range(16):
        for ig in range(4):
            for ks in range(2):
                for ii in range(4):
                    k = 32 * w + 2 * (4 * ig + ii) + ks
                    out[:, :, :, k] = blk[:, w, ig, ks, :, ii, :].transpose(0, 2, 1)
    return out


def fc_matrix(img, off):
    """fc layout [k 512][unit 4] -> [cta][unit 4][k 512]"""
    return img[:, off:off + 4 * HID].reshape(NWORK, HID, 4).astype(np.float64).transpose(0, 2, 1)


def cond_matrices(img, off):
    """[k' 176][rb 8][4] -> dense [cta][which 2][unit 4][row 4][208] over the REAL conditioning index, using cond_pass's map
    rk = k' (k' < 112) | which 0: k' + 32 | which 1: k' (k' < 144) else k' + 32."""
    blk = img[:, off:off + KC2 * 32].reshape(NWORK, KC2, 2, 4, 4).astype(np.float64)          # [cta][k'][which][u][row]
    out = np.zeros((NWORK, 2, 4, 4, 208))
    for kp in range(KC2):
        for which in range(2):
            rk = kp if kp < 112 else (kp + 32 if which == 0 else (kp if kp < 144 else kp + 32))
            out[:, which, :, :, rk] += blk[:, kp, which]
    return out


class Emu:
    def __init__(self, img, layout, C):
        per, ih2, hh1, hh2, fc1, fc2, fc3, wc = layout
        assert img.shape[1] == per
        self.C = C
        self.IH2, self.HH1, self.HH2 = gate_matrix(img, ih2), gate_matrix(img, hh1), gate_matrix(img, hh2)
        self.FC1, self.FC2, self.FC3 = fc_matrix(img, fc1), fc_matrix(img, fc2), fc_matrix(img, fc3)
        self.WC = cond_matrices(img, wc)
        self.sv = img[:, wc + KC2 * 32: wc + KC2 * 32 + 128].astype(np.float64)

    def run(self, mels, aux, forced):
        B, S, _ = mels.shape
        sv = self.sv
        v3 = lambda name: np.stack([sv[:, SV[name] + 4 * q: SV[name] + 4 * q + 4] for q in range(3)], 1)[..., None]   # [cta][3][4][1]
        v1 = lambda name: sv[:, SV[name]: SV[name] + 4][..., None]                                                    # [cta][4][1]
        sig = lambda v: 1.0 / (1.0 + np.exp(-v))
        gates = lambda M, v: np.einsum("cguk,kb->cgub", M, v)
        rows = lambda M, v: np.einsum("cuk,kb->cub", M, v)
        vec = lambda per_cta: per_cta.reshape(HID, -1)                    # unit = 4 cta + u
        h1 = np.zeros((HID, B)); h2 = np.zeros((HID, B)); x = np.zeros(B)
        gh1 = np.broadcast_to(v3("BHH1"), (NWORK, 3, 4, B)).copy()
        gh2 = np.broadcast_to(v3("BHH2"), (NWORK, 3, 4, B)).copy()
        out = np.zeros((S, B, self.C))
        for s in range(S):
            c = np.concatenate([mels[:, s].T, aux[:, s].T], 0)            # [208][B]
            P = np.einsum("cwurk,kb->cwurb", self.WC, c)                  # [cta][which][u][row][B]
            PA, PB = P[:, 0], P[:, 1]                                     # rows {P1 r,z,n,P3} / {P2 r,z,n,P4}
            gi = PA[:, :, :3].transpose(0, 2, 1, 3) + x * v3("U1") + v3("B1")
            r, z = sig(gi[:, 0] + gh1[:, 0]), sig(gi[:, 1] + gh1[:, 1])
            n = np.tanh(gi[:, 2] + r * gh1[:, 2])
            h1 = vec((1 - z) * n + z * h1.reshape(NWORK, 4, B))
            gi = gates(self.IH2, h1) + PB[:, :, :3].transpose(0, 2, 1, 3) + x * v3("U2") + v3("B2")
            gh1 = gates(self.HH1, h1) + v3("BHH1")
            f1 = rows(self.FC1, h1)
            r, z = sig(gi[:, 0] + gh2[:, 0]), sig(gi[:, 1] + gh2[:, 1])
            n = np.tanh(gi[:, 2] + r * gh2[:, 2])
            h2 = vec((1 - z) * n + z * h2.reshape(NWORK, 4, B))
            gh2 = gates(self.HH2, h2) + v3("BHH2")
            y1 = vec(np.maximum(rows(self.FC1, h2) + f1 + PA[:, :, 3] + x * v1("U3") + v1("B3"), 0))
            y2 = vec(np.maximum(rows(self.FC2, y1) + PB[:, :, 3] + v1("B4"), 0))
            lg = vec(rows(self.FC3, y2) + v1("B5"))                       # class = 4 cta + u
            out[s] = lg[:self.C].T
            x = forced[s]
        return out


@pytest.mark.parametrize("mode", ["RAW", "MOL"])
def test_wide_images_reproduce_oracle_logits(mode):
    sd = synth.make_state(mode, "ref", 3)
    C = sd["fc3.weight"].shape[0]
    img, layout = pack(sd, mode)
    rng = np.random.default_rng(5)
    B, S = 3, 6
    mels = rng.uniform(0, 1, (B, S, 80)).astype(np.float32)
    aux = rng.normal(0, 1, (B, S, 128)).astype(np.float32)
    forced = rng.uniform(-1, 1, (S, B)).astype(np.float32)
    U = np.zeros((S, B) if mode == "RAW" else (S, B, 11), np.float32) + 0.5
    want = c_oracle.generate_folds(sd, mode, mels, aux, U, forced_x=forced, want_logits=True, precision="fp64")["logits"]
    got = Emu(img, layout, C).run(mels.astype(np.float64), aux.astype(np.float64), forced.astype(np.float64))
    err = np.abs(got - want).max()
    assert err < 5e-6, err


def test_wide_pack_rejects_what_the_wide_kernel_does_not_serve():
    L = _lib.lib()
    for cfg in (_lib.Config(512, 512, 80, 32, 1024, 0, 0), _lib.Config(512, 512, 80, 32, 512, 0, 1), _lib.Config(512, 512, 80, 32, 512, 0, 2)):
        assert L.wrnn_wide_packed_floats(ctypes.byref(cfg), None) == -1
