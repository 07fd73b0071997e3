"""CPU test of the host-side weight repack (wrnn_pack_weights_host): re-run the kernel's
dataflow in numpy float64 FROM THE PACKED SHARED-MEMORY IMAGES ONLY and compare the logits
with the oracle.  This pins the algebraic folding of the I layer / conditioning terms, the
item / chunk / lane permutations and the per-CTA row ownership without needing a GPU."""
import ctypes

import numpy as np
import pytest

from expressive_speech_synthesis_research_b200 import _lib
from oracle import c_oracle, synth

HID, NCTA, ITEM = 512, 128, 512


def pack(sd, mode, precision=0):
    C = sd["fc3.weight"].shape[0]
    cfg = _lib.Config(512, 512, 80, 32, C, _lib.MODE[mode], precision)
    L = _lib.lib()
    n = L.wrnn_packed_floats(ctypes.byref(cfg))
    assert n > 0
    keep, w = [], _lib.Weights()
    for field, key in zip(_lib.Weights.FIELDS, _lib.Weights.KEYS):
        a = np.ascontiguousarray(sd[key].numpy(), dtype=np.float32)
        keep.append(a)
        setattr(w, field, a.ctypes.data)
    out = np.zeros(n, dtype=np.float32)
    _lib.check(L.wrnn_pack_weights_host(ctypes.byref(cfg), ctypes.byref(w), out.ctypes.data, n))
    out = out.reshape(NCTA, -1)
    if precision == 1:
        # bf16 images: every item weight is a bf16 (two per float slot), the 128 small-vector floats stay fp32
        items = (out.shape[1] - 128) * 2
        bits = out[:, :-128].copy().view(np.uint16).astype(np.uint32) << 16
        assert bits.shape[1] == items
        out = np.concatenate([bits.view(np.float32), out[:, -128:]], axis=1)
    return out


def item_rows(img, base, idx):
    """item image [4 slots][32 lanes][4] -> [4 rows][128] with k = lane + 32*i; slot s of lane l holds
    row s ^ (l >> 3) (the per-lane permutation behind the kernel's select-free reduce-scatter)"""
    it = img[:, base + idx * ITEM: base + (idx + 1) * ITEM].reshape(NCTA, 4, 32, 4)
    rows = np.empty_like(it)
    lanes = np.arange(32)
    for s in range(4):
        rows[:, s ^ (lanes >> 3), lanes, :] = it[:, s, lanes, :]
    return rows.transpose(0, 1, 3, 2).reshape(NCTA, 4, 128).astype(np.float64)


def matrix(img, base, rg, nchunk=4):
    return np.concatenate([item_rows(img, base, rg * nchunk + kc) for kc in range(nchunk)], axis=2)   # [cta][4][512]


class Emu:
    def __init__(self, img, C, mode):
        rows5 = C // 128 if C > 512 else 4
        self.rows5, self.C, self.mode = rows5, C, mode
        W_M2, W_M3 = 0, 24 * HID
        W_M4 = W_M3 + 16 * HID
        W_M5 = W_M4 + 4 * HID
        W_MC = W_M5 + rows5 * HID
        W_SV = W_MC + 12 * ITEM
        g = lambda base, rg: matrix(img, base, rg)
        self.A2 = [g(W_M2, q) for q in range(3)]          # Wih2x gates  [cta][4][512]
        self.H1 = [g(W_M2, 3 + q) for q in range(3)]      # Whh1 gates
        self.A3 = g(W_M3, 0)
        self.H2 = [g(W_M3, 1 + q) for q in range(3)]
        self.W4 = g(W_M4, 0)
        self.W5 = [g(W_M5, rg) for rg in range(rows5 // 4)]

        def cond(idx_chunks):                              # list of (item, chunk) -> [cta][4][256]
            out = np.zeros((NCTA, 4, 256))
            for it, ch in idx_chunks:
                out[:, :, ch * 128:(ch + 1) * 128] += item_rows(img, W_MC, it)
            return out
        self.C1 = [cond([(q, 0)]) for q in range(3)]
        self.C2 = [cond([(3 + 2 * q, 0), (4 + 2 * q, 1)]) for q in range(3)]
        self.C3 = cond([(9, 0), (10, 1)])
        self.C4 = cond([(11, 1)])
        sv = img[:, W_SV:W_SV + 128].astype(np.float64)
        self.sv = sv
        self.h1 = None

    def vec(self, per_cta):                                # [cta][4][B] -> [512][B]
        return per_cta.reshape(HID, -1)

    def run(self, mels, aux, forced):
        """mels [B,S,80], aux [B,S,128], forced [S,B] -> logits [S,B,C] (float64)."""
        B, S, _ = mels.shape
        sv = self.sv
        sl = lambda o, q: sv[:, o + 4 * q:o + 4 * q + 4][:, :, None]           # [cta][4][1]
        h1 = np.zeros((HID, B)); h2 = np.zeros((HID, B)); x = np.zeros(B)
        gh1 = [np.broadcast_to(sl(24, q), (NCTA, 4, B)).copy() for q in range(3)]
        gh2 = [np.broadcast_to(sl(60, q), (NCTA, 4, B)).copy() for q in range(3)]
        sig = lambda v: 1.0 / (1.0 + np.exp(-v))
        out = np.zeros((S, B, self.C))
        for s in range(S):
            c = np.zeros((256, B))
            c[:80] = mels[:, s].T
            c[80:208] = aux[:, s].T
            mm = lambda M, v: np.einsum("crk,kb->crb", M, v)
            gi = [mm(self.C1[q], c) + x * sl(0, q) + sl(12, q) for q in range(3)]
            r, z = sig(gi[0] + gh1[0]), sig(gi[1] + gh1[1])
            n = np.tanh(gi[2] + r * gh1[2])
            h1 = self.vec((1 - z) * n + z * h1.reshape(NCTA, 4, B))
            gi = [mm(self.A2[q], h1) + mm(self.C2[q], c) + x * sl(36, q) + sl(48, q) for q in range(3)]
            gh1 = [mm(self.H1[q], h1) + sl(24, q) for q in range(3)]
            r, z = sig(gi[0] + gh2[0]), sig(gi[1] + gh2[1])
            n = np.tanh(gi[2] + r * gh2[2])
            h2 = self.vec((1 - z) * n + z * h2.reshape(NCTA, 4, B))
            gh2 = [mm(self.H2[q], h2) + sl(60, q) for q in range(3)]
            y1 = self.vec(np.maximum(mm(self.A3, h1 + h2) + mm(self.C3, c) + x * sv[:, 72:76, None] + sv[:, 76:80, None], 0))
            y2 = self.vec(np.maximum(mm(self.W4, y1) + mm(self.C4, c) + sv[:, 80:84, None], 0))
            lg = np.concatenate([mm(W, y2) + sv[:, 84 + 4 * i:88 + 4 * i, None] for i, W in enumerate(self.W5)], axis=1)
            lg = lg.reshape(NCTA * self.rows5, B)                               # class = rows5*cta + row
            out[s] = lg[:self.C].T
            x = forced[s]
        return out


@pytest.mark.parametrize("mode,bits", [("RAW", 9), ("MOL", 9), ("RAW", 10), ("RAW", 8)])
def test_packed_images_reproduce_oracle_logits(mode, bits):
    sd = synth.make_state(mode, "ref", 3, bits=bits)
    C = sd["fc3.weight"].shape[0]
    img = pack(sd, mode)
    rng = np.random.default_rng(5)
    B, S = 3, 6
    mels = rng.uniform(0, 1, (B, S, 80)).astype(np.float32)
    aux = rng.normal(0, 1, (B, S, 128)).astype(np.float32)
    forced = rng.uniform(-1, 1, (S, B)).astype(np.float32)
    U = np.zeros((S, B) if mode == "RAW" else (S, B, 11), np.float32) + 0.5
    want = c_oracle.generate_folds(sd, mode, mels, aux, U, forced_x=forced, want_logits=True, precision="fp64")["logits"]
    got = Emu(img, C, mode).run(mels.astype(np.float64), aux.astype(np.float64), forced.astype(np.float64))
    err = np.abs(got - want).max()
    assert err < 5e-6, err


@pytest.mark.parametrize("mode", ["RAW", "MOL"])
def test_bf16_weight_images_stay_within_the_bf16_tolerance(mode):
    """precision=bf16 rounds the RESIDENT weights (after the fp64 folding) to bf16; activations and sums stay fp32.
    Tolerance from SURVEY.md 8c: teacher-forced logits within 3e-2 of the fp32 model."""
    sd = synth.make_state(mode, "ref", 3)
    C = sd["fc3.weight"].shape[0]
    img32, img16 = pack(sd, mode, 0), pack(sd, mode, 1)
    assert img16.shape == img32.shape
    w32, w16 = img32[:, :-128], img16[:, :-128]
    assert np.array_equal(img16[:, -128:], img32[:, -128:])
    nz = w32 != 0
    assert np.abs(w16[nz] / w32[nz] - 1).max() <= 2.0 ** -8          # bf16 keeps 8 significant bits
    rng = np.random.default_rng(5)
    B, S = 3, 6
    mels = rng.uniform(0, 1, (B, S, 80)).astype(np.float32)
    aux = rng.normal(0, 1, (B, S, 128)).astype(np.float32)
    forced = rng.uniform(-1, 1, (S, B)).astype(np.float32)
    U = np.zeros((S, B) if mode == "RAW" else (S, B, 11), np.float32) + 0.5
    want = c_oracle.generate_folds(sd, mode, mels, aux, U, forced_x=forced, want_logits=True, precision="fp64")["logits"]
    got = Emu(img16, C, mode).run(mels.astype(np.float64), aux.astype(np.float64), forced.astype(np.float64))
    err = np.abs(got - want).max()
    assert 1e-6 < err < 3e-2, err


def test_pack_rejects_unsupported_configs():
    L = _lib.lib()
    for cfg in (_lib.Config(256, 512, 80, 32, 512, 0, 0), _lib.Config(512, 512, 80, 32, 500, 0, 0),
                _lib.Config(512, 512, 80, 32, 30, 1, 7), _lib.Config(512, 512, 80, 32, 512, 7, 0)):
        assert L.wrnn_packed_floats(ctypes.byref(cfg)) == -1
