"""Test infrastructure: numpy replay of the tensor-core program of precision "bf16-dense"
(expressive_speech_synthesis_research_b200/csrc/wavernn_dense.cuh) FROM THE PACKED OPERAND STREAM AND BUNDLE TABLE
that wrnn_dense_pack_host emits -- the same bytes wrnn_load_weights uploads.  Every bundle is applied in issue
order to per-CTA accumulators, the epilogues run where the kernel's commits are, so a wrong tile, row map,
first-touch flag or B-operand offset shows up as a logits error on the CPU, before any GPU time is spent."""
import ctypes

import numpy as np

from expressive_speech_synthesis_research_b200 import _lib

SEG = np.dtype([("off16", "<u2"), ("rows", "<u2"), ("nk", "<u2"), ("bsrc16", "<u2"), ("dcol", "<u2"), ("first", "<u2")])
BUNDLE = np.dtype([("bytes", "<u4"), ("src_off", "<u4"), ("nseg", "<u2"), ("wait", "<u2"), ("commit", "<u2"), ("pad", "<u2"),
                   ("seg", SEG, (4,))])
W_COND = 5
C_G2, C_F1, C_F2, C_F3, C_G1, C_H2RD = 1, 2, 3, 4, 5, 6
D_G1_T0, D_G1_1H, D_G1_1I, D_G2_T0, D_G2_1H, D_G2_1I, D_F1, D_F2, D_F3 = 0, 32, 64, 96, 128, 160, 192, 224, 256
(B1R, U1R, B1Z, U1Z, B1NI, U1N, B1NH, B2R, U2R, B2Z, U2Z, B2NI, U2N, B2NH, B3, U3, B4, B5) = range(18)


def bf16_round(a):
    """fp32 round-to-nearest-even to bf16 (returned as float64)."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) >> 16 << 16
    return u.astype(np.uint32).view(np.float32).astype(np.float64)


class DenseReplay:
    def __init__(self, sd, mode="RAW"):
        C = sd["fc3.weight"].shape[0]
        cfg = _lib.Config(512, 512, 80, 32, C, _lib.MODE[mode], _lib.PRECISION["bf16-dense"])
        L = _lib.lib()
        lay = (ctypes.c_int64 * 8)()
        _lib.check(L.wrnn_dense_layout(ctypes.byref(cfg), lay))
        self.nb, self.stream_bytes, bsz, self.CL, self.UPC, self.BC, self.NSV = (int(v) for v in lay[:7])
        assert bsz == BUNDLE.itemsize
        keep, w = [], _lib.Weights()
        for field, key in zip(_lib.Weights.FIELDS, _lib.Weights.KEYS):
            a = np.ascontiguousarray(sd[key].numpy(), dtype=np.float32)
            keep.append(a)
            setattr(w, field, a.ctypes.data)
        stream = np.zeros(self.CL * self.stream_bytes, np.uint8)
        table = np.zeros(self.nb * bsz, np.uint8)
        sv = np.zeros(self.CL * self.NSV * self.UPC, np.float32)
        _lib.check(L.wrnn_dense_pack_host(ctypes.byref(cfg), ctypes.byref(w), stream.ctypes.data, table.ctypes.data, sv.ctypes.data))
        self.stream = stream.reshape(self.CL, self.stream_bytes)
        self.table = table.view(BUNDLE)
        self.sv = sv.reshape(self.CL, self.NSV, self.UPC).astype(np.float64)
        self.C = C
        self.chunk_b = self.BC * 16
        self.img_b = 64 * self.chunk_b

    def tile(self, rank, bundle, seg):
        """A operand of a segment as float64 [rows][nk * 16]."""
        rows, nk = int(seg["rows"]), int(seg["nk"])
        o = int(bundle["src_off"]) + int(seg["off16"]) * 16
        raw = self.stream[rank, o:o + rows * nk * 32].view(np.uint16).astype(np.uint32) << 16
        t = raw.view(np.float32).reshape(2 * nk, rows, 8)                   # [k chunk][row][8 k]
        return t.transpose(1, 0, 2).reshape(rows, nk * 16).astype(np.float64)

    def operand_b(self, seg, images, cond):
        off = int(seg["bsrc16"]) * 16
        nk = int(seg["nk"])
        if off >= 4 * self.img_b:
            k0 = (off - 4 * self.img_b) // self.chunk_b * 8
            return cond[k0:k0 + nk * 16]
        img, k0 = off // self.img_b, off % self.img_b // self.chunk_b * 8
        return images[img][k0:k0 + nk * 16]

    def run(self, mels, aux, forced, round_act=True):
        """mels [B,S,80], aux [B,S,128], forced [S,B] -> logits [S,B,C] float64 (teacher forced)."""
        B, S, _ = mels.shape
        CL, UPC = self.CL, self.UPC
        rnd = bf16_round if round_act else (lambda a: np.asarray(a, np.float64))
        sig = lambda v: 1.0 / (1.0 + np.exp(-v))
        images = [np.zeros((512, B)) for _ in range(4)]                       # h1, h2, y1, y2 as the MMA reads them
        acc = [dict() for _ in range(CL)]                                      # rank -> column -> [128][B]
        h1 = np.zeros((512, B)); h2 = np.zeros((512, B)); x = np.zeros(B)      # fp32 state in the kernel
        out = np.zeros((S, B, self.C))

        def cond_of(step):
            c = np.zeros((208, B))
            if step < S:
                c[:80] = mels[:, step].T
                c[80:] = aux[:, step].T
            return rnd(c)

        # accumulator lanes as the kernel reads them: tile 0 (M = 128, row = lane) holds r of unit 16q+j in lane 32q+j and its z
        # 16 lanes above; a 64-row tile (M = 64) leaves row i in lane 32 (i // 16) + i % 16, modelled here as plain row i
        uu = np.arange(UPC)
        r_lane = 32 * (uu // 16) + uu % 16

        def gru(rank, t0, c1h, c1i, sv, br, ur, bz, uz, bni, un, bnh, hprev):
            a0 = acc[rank][t0]
            r, z = a0[r_lane], a0[r_lane + 16]
            nh, ni = acc[rank][c1h][:UPC], acc[rank][c1i][:UPC]
            col = lambda i: sv[i][:, None]
            rr = sig(r + col(br) + x * col(ur))
            zz = sig(z + col(bz) + x * col(uz))
            nn = np.tanh(ni + col(bni) + x * col(un) + rr * (nh + col(bnh)))
            return nn + zz * (hprev - nn)

        cond = cond_of(0)
        for t in range(-1, S):
            pre = t < 0
            if not pre:                                                        # E1
                for rank in range(CL):
                    sl = slice(UPC * rank, UPC * rank + UPC)
                    h1[sl] = gru(rank, D_G1_T0, D_G1_1H, D_G1_1I, self.sv[rank], B1R, U1R, B1Z, U1Z, B1NI, U1N, B1NH, h1[sl])
                images[0] = rnd(h1)
            for b in range(self.nb):
                bd = self.table[b]
                if int(bd["wait"]) == W_COND:
                    cond = cond_of(t + 1)
                for rank in range(CL):
                    for s in range(int(bd["nseg"])):
                        sg = bd["seg"][s]
                        A = self.tile(rank, bd, sg)
                        prod = A @ self.operand_b(sg, images, cond)
                        rows, col = int(sg["rows"]), int(sg["dcol"])
                        if int(sg["first"]):
                            a = np.full((128, B), np.nan)
                            a[:rows] = prod
                            acc[rank][col] = a
                        else:
                            if col not in acc[rank]:
                                acc[rank][col] = np.full((128, B), np.nan)     # uninitialised tensor memory
                            acc[rank][col][:rows] += prod
                cm = int(bd["commit"])
                if pre or cm in (0, C_G1, C_H2RD):
                    continue
                for rank in range(CL):
                    sl = slice(UPC * rank, UPC * rank + UPC)
                    sv = self.sv[rank]
                    if cm == C_G2:
                        h2[sl] = gru(rank, D_G2_T0, D_G2_1H, D_G2_1I, sv, B2R, U2R, B2Z, U2Z, B2NI, U2N, B2NH, h2[sl])
                    elif cm == C_F1:
                        images[2][sl] = rnd(np.maximum(acc[rank][D_F1][:UPC] + sv[B3][:, None] + x * sv[U3][:, None], 0))
                    elif cm == C_F2:
                        images[3][sl] = rnd(np.maximum(acc[rank][D_F2][:UPC] + sv[B4][:, None], 0))
                    elif cm == C_F3:
                        lo, hi = UPC * rank, min(UPC * rank + UPC, self.C)          # MOL: only rows 0-29 of CTA 0 exist
                        if hi > lo:
                            out[t, :, lo:hi] = (acc[rank][D_F3][:hi - lo] + sv[B5][:hi - lo, None]).T
                if cm == C_G2:
                    images[1] = rnd(h2)
                if cm == C_F3:
                    x = np.asarray(forced[t], np.float64)
        return out
