"""Host-side mirror of the reference vocoder's interface for the generation hot path.

`WaveRNN` keeps the reference constructor keywords, sub-module names (=> identical
state_dict keys, reference checkpoints load unchanged) and the
`generate(mels, [save_path,] batched, target, overlap, mu_law)` contract of
WaveRNN/models/fatchord_version.py:89-243, but the autoregressive loop, the fold
gather and the crossfade/mu-law epilogue run in the sm_100a library behind
include/wavernn_b200.h.  PyTorch only runs the non-autoregressive conditioning
network (MelResNet / UpsampleNetwork, fatchord_version.py:10-86) and owns device memory.

There is no CPU path: generate() raises if CUDA or the compiled library is missing.
"""
import ctypes
import os
import time

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .wavio import WavWriter, decode_mu_law_host, save_wav


class _ResidualUnit(nn.Module):
    """1x1 conv / BN / ReLU / 1x1 conv / BN + skip (reference ResBlock, fatchord_version.py:10-25)."""

    def __init__(self, dims):
        super().__init__()
        self.conv1 = nn.Conv1d(dims, dims, kernel_size=1, bias=False)
        self.conv2 = nn.Conv1d(dims, dims, kernel_size=1, bias=False)
        self.batch_norm1 = nn.BatchNorm1d(dims)
        self.batch_norm2 = nn.BatchNorm1d(dims)

    def forward(self, x):
        y = F.relu(self.batch_norm1(self.conv1(x)))
        return self.batch_norm2(self.conv2(y)) + x


class MelResNet(nn.Module):
    """Frame-rate aux feature network (fatchord_version.py:28-45)."""

    def __init__(self, res_blocks, in_dims, compute_dims, res_out_dims, pad):
        super().__init__()
        self.conv_in = nn.Conv1d(in_dims, compute_dims, kernel_size=2 * pad + 1, bias=False)
        self.batch_norm = nn.BatchNorm1d(compute_dims)
        self.layers = nn.ModuleList(_ResidualUnit(compute_dims) for _ in range(res_blocks))
        self.conv_out = nn.Conv1d(compute_dims, res_out_dims, kernel_size=1)

    def forward(self, x):
        x = F.relu(self.batch_norm(self.conv_in(x)))
        for layer in self.layers:
            x = layer(x)
        return self.conv_out(x)


class _Repeat(nn.Module):
    """Nearest-neighbour stretch of the time axis (reference Stretch2d(x_scale, 1), :48-58)."""

    def __init__(self, scale):
        super().__init__()
        self.scale = scale

    def forward(self, x):
        return x.repeat_interleave(self.scale, dim=-1)


class UpsampleNetwork(nn.Module):
    """Mel -> sample-rate conditioning (fatchord_version.py:61-86): aux = repeat(resnet(m), hop);
    mel = three (repeat, box-filter conv) stages cropped by pad*hop on both sides."""

    def __init__(self, feat_dims, upsample_scales, compute_dims, res_blocks, res_out_dims, pad):
        super().__init__()
        self.total_scale = int(np.prod(upsample_scales))
        self.indent = pad * self.total_scale
        self.resnet = MelResNet(res_blocks, feat_dims, compute_dims, res_out_dims, pad)
        self.resnet_stretch = _Repeat(self.total_scale)
        self.up_layers = nn.ModuleList()
        for s in upsample_scales:
            conv = nn.Conv2d(1, 1, kernel_size=(1, 2 * s + 1), padding=(0, s), bias=False)
            conv.weight.data.fill_(1.0 / (2 * s + 1))
            self.up_layers.append(_Repeat(s))      # index 0, 2, 4 (no parameters)
            self.up_layers.append(conv)            # index 1, 3, 5 (state_dict keys up_layers.{1,3,5}.weight)

    def forward(self, m):
        aux = self.resnet_stretch(self.resnet(m))
        m = m.unsqueeze(1)
        for layer in self.up_layers:
            m = layer(m)
        m = m.squeeze(1)[:, :, self.indent:-self.indent]
        return m.transpose(1, 2), aux.transpose(1, 2)


class _Engine:
    """Owns one wrnn_handle (include/wavernn_b200.h) on one CUDA device."""

    def __init__(self, cfg: _lib.Config, device_index: int):
        self.lib = _lib.lib()
        self.handle = _lib.vp()
        self.cfg = cfg
        self.device_index = device_index
        _lib.check(self.lib.wrnn_create(ctypes.byref(cfg), device_index, ctypes.byref(self.handle)))
        self.weights_tag = None

    def load(self, state, tag):
        keep, w = [], _lib.Weights()
        for field, key in zip(_lib.Weights.FIELDS, _lib.Weights.KEYS):
            t = state[key].detach().to("cpu", torch.float32).contiguous()
            keep.append(t)
            setattr(w, field, t.data_ptr())
        _lib.check(self.lib.wrnn_load_weights(self.handle, ctypes.byref(w)))
        self.weights_tag = tag

    def info(self):
        out = _lib.Info()
        _lib.check(self.lib.wrnn_get_info(self.handle, ctypes.byref(out)))
        return out

    def set_kernel(self, choice):
        """-1: fp32 kernel by the fold count of each call (default); 0 grouped; 1 wide (wrnn_set_kernel)."""
        _lib.check(self.lib.wrnn_set_kernel(self.handle, int(choice)))

    def synchronize(self):
        """Wait for the generate call enqueued last; raises when an in-kernel watchdog fired (WRNN_ERR_TIMEOUT)."""
        _lib.check(self.lib.wrnn_synchronize(self.handle))

    def query(self):
        """(done, steps_done) of the call in flight, without blocking (the gen_display hook, fatchord_version.py:220)."""
        done, steps = _lib.i32(), _lib.i32()
        _lib.check(self.lib.wrnn_query(self.handle, ctypes.byref(done), ctypes.byref(steps)))
        return bool(done.value), int(steps.value)

    def stage_cycles(self, enable=None):
        """enable=True/False toggles in-kernel stage timing; None returns the last launch's counters
        as an int64 array [128, 32]."""
        if enable is not None:
            _lib.check(self.lib.wrnn_set_profiling(self.handle, int(bool(enable))))
            return None
        out = np.zeros((128, 32), dtype=np.int64)
        _lib.check(self.lib.wrnn_get_stage_cycles(self.handle, out.ctypes.data, out.size))
        return out

    def measure_exchange(self, iters=2000):
        us = ctypes.c_float()
        _lib.check(self.lib.wrnn_measure_exchange(self.handle, iters, ctypes.byref(us)))
        return float(us.value)

    def close(self):
        if self.handle:
            self.lib.wrnn_destroy(self.handle)
            self.handle = _lib.vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _Cond:
    """Owns one wrnn_cond (include/wavernn_b200.h): the frame-rate MelResNet of one model on one CUDA device."""

    def __init__(self, blob: np.ndarray, res_blocks: int, device_index: int):
        self.lib = _lib.lib()
        self.handle = _lib.vp()
        self.device_index = device_index
        blob = np.ascontiguousarray(blob, dtype=np.float32)
        _lib.check(self.lib.wrnn_cond_create(device_index, blob.ctypes.data_as(ctypes.c_void_p), blob.size, res_blocks, ctypes.byref(self.handle)))
        self.tag = None

    def launches(self):
        return int(self.lib.wrnn_cond_launches(self.handle))

    def frames(self, mel_frames, segments, aux_out):
        """mel_frames [rows, 80] (device, zero-padded segments), segments int32 [n, 3] host, aux_out [rows, 128] (device): one launch."""
        seg = np.ascontiguousarray(segments, dtype=np.int32)
        stream = torch.cuda.current_stream(mel_frames.device).cuda_stream
        _lib.check(self.lib.wrnn_cond_frames(self.handle, ctypes.c_void_p(mel_frames.data_ptr()), seg.ctypes.data_as(ctypes.c_void_p),
                                             seg.shape[0], ctypes.c_void_p(aux_out.data_ptr()), ctypes.c_void_p(stream)))

    def __del__(self):
        try:
            if self.handle:
                self.lib.wrnn_cond_destroy(self.handle)
                self.handle = _lib.vp()
        except Exception:
            pass


class WaveRNN(nn.Module):
    """Drop-in for models.fatchord_version.WaveRNN on the generation path."""

    def __init__(self, rnn_dims, fc_dims, bits, pad, upsample_factors, feat_dims, compute_dims,
                 res_out_dims, res_blocks, hop_length, sample_rate, mode='RAW'):
        super().__init__()
        self.mode = mode
        self.pad = pad
        if mode == 'RAW':
            self.n_classes = 2 ** bits
        elif mode == 'MOL':
            self.n_classes = 30
        else:
            raise RuntimeError("Unknown model mode value - %s" % (mode,))   # the reference builds this error but never raises it (:101)
        self.rnn_dims = rnn_dims
        self.aux_dims = res_out_dims // 4
        self.hop_length = hop_length
        self.sample_rate = sample_rate

        self.upsample = UpsampleNetwork(feat_dims, upsample_factors, compute_dims, res_blocks, res_out_dims, pad)
        self.I = nn.Linear(feat_dims + self.aux_dims + 1, rnn_dims)
        self.rnn1 = nn.GRU(rnn_dims, rnn_dims, batch_first=True)
        self.rnn2 = nn.GRU(rnn_dims + self.aux_dims, rnn_dims, batch_first=True)
        self.fc1 = nn.Linear(rnn_dims + self.aux_dims, fc_dims)
        self.fc2 = nn.Linear(fc_dims + self.aux_dims, fc_dims)
        self.fc3 = nn.Linear(fc_dims, self.n_classes)
        self.step = nn.Parameter(torch.zeros(1).long(), requires_grad=False)

        self._feat_dims = feat_dims
        self._fc_dims = fc_dims
        self._engines = {}
        self._conds = {}                    # device index -> _Cond (frame-rate MelResNet kernel)
        # "fp32" (default, the reference's precision) or "bf16": resident weights rounded to bf16 (after the fp64
        # folding of the input layer), activations and accumulation stay fp32.  Set before calling generate().
        self.precision = "fp32"
        # "host": decode_mu_law + tail fade in numpy on the crossfaded float64 signal (bit-exact with the reference, dsp.py:100-105);
        # "device": CUDA pow inside the epilogue kernel (within 2 ulp; for long-form audio); "auto": host below 2 M samples
        self.mu_law_decode = "auto"
        # called as progress(steps_done, steps_total) about ten times a second while a single-launch generate() runs (the
        # reference prints a progress line every 100 steps, fatchord_version.py:220,246-250); None = silent
        self.progress = None
        # precision "bf16-dense" only: expand the conditioning inside the kernel from frame-rate tensors instead
        # of materialising UpsampleNetwork's [samples, 208] output (832 B per sample).  False = materialise, as the other kernels do.
        self.expand_in_kernel = True
        self._interp_cache = None
        self._pinned_cache = {}
        self.last_stats = {}

    # ------------------------------------------------------------------ teacher-forced twin (training path)
    def forward(self, x, mels):
        """WaveRNN.forward, fatchord_version.py:119-148: x (B, T*hop) the previous samples (x[:, 0] = 0), mels
        (B, feat, T + 2*pad) -> logits (B, T*hop, n_classes).  PyTorch on whatever device the inputs live on (the
        reference forces .cuda()); autograd works, so train_wavernn.py:36 (`model(x, m)`) runs against this class."""
        self.step += 1
        dev = x.device
        h1 = torch.zeros(1, x.size(0), self.rnn_dims, device=dev, dtype=x.dtype)
        h2 = torch.zeros(1, x.size(0), self.rnn_dims, device=dev, dtype=x.dtype)
        mels, aux = self.upsample(mels)
        d = self.aux_dims
        a1, a2, a3, a4 = (aux[:, :, i * d:(i + 1) * d] for i in range(4))
        x = self.I(torch.cat([x.unsqueeze(-1), mels, a1], dim=2))
        res = x
        x, _ = self.rnn1(x, h1)
        x = x + res
        res = x
        x, _ = self.rnn2(torch.cat([x, a2], dim=2), h2)
        x = x + res
        x = F.relu(self.fc1(torch.cat([x, a3], dim=2)))
        x = F.relu(self.fc2(torch.cat([x, a4], dim=2)))
        return self.fc3(x)

    # ------------------------------------------------------------------ engine plumbing
    def _weights_tag(self):
        sd = self.state_dict()
        return tuple((sd[k].data_ptr(), sd[k]._version) for k in _lib.Weights.KEYS)

    def _engine(self, device: torch.device, precision=None) -> _Engine:
        if device.type != "cuda":
            raise RuntimeError("WaveRNN.generate runs on a CUDA sm_100 device only (no CPU path); got %s" % device)
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if precision is None:
            precision = "fp32" if self.precision == "auto" else self.precision
        if precision not in _lib.PRECISION:
            raise ValueError("precision must be 'auto' or one of %s, got %r" % (sorted(_lib.PRECISION), self.precision))
        key = (idx, precision)
        eng = self._engines.get(key)
        if eng is None:
            cfg = _lib.Config(self.rnn_dims, self._fc_dims, self._feat_dims, self.aux_dims, self.n_classes,
                              _lib.MODE[self.mode], _lib.PRECISION[precision])
            with torch.cuda.device(idx):                               # the engine's allocations land on its own device
                eng = self._engines[key] = _Engine(cfg, idx)
        tag = self._weights_tag()
        if eng.weights_tag != tag:
            eng.load(self.state_dict(), tag)
        return eng

    def _device(self):
        p = self.I.weight
        if p.is_cuda:
            return p.device
        if not torch.cuda.is_available():
            raise RuntimeError("CUDA is not available: the B200 WaveRNN path has no CPU fallback")
        return torch.device("cuda", torch.cuda.current_device())

    # ------------------------------------------------------------------ conditioning (PyTorch)
    def conditioning(self, mels):
        """generate() prologue, fatchord_version.py:162-165: (1, feat, T) -> (L, feat), (L, 4*aux)."""
        m = F.pad(mels, (self.pad, self.pad))                       # pad_tensor(side='both'), :260-270
        if mels.is_cuda and self.melresnet_native():
            # aux = repeat(MelResNet(m), hop) with the CUDA kernel of csrc/wavernn_cond.cuh (one launch instead of ~65); the three
            # (repeat, box filter) stages of the mel branch stay PyTorch convolutions in true fp32
            T = int(mels.shape[-1])
            af = torch.empty(T, 4 * self.aux_dims, dtype=torch.float32, device=mels.device)
            self._cond(mels.device).frames(m[0].t().contiguous(), np.array([[0, T, 0]], dtype=np.int32), af)
            aux = af.repeat_interleave(self.upsample.total_scale, dim=0)
            with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
                mu = m.unsqueeze(1)
                for layer in self.upsample.up_layers:
                    mu = layer(mu)
            mu = mu.squeeze(1)[:, :, self.upsample.indent:-self.upsample.indent]
            return mu[0].t().contiguous(), aux
        m, aux = self.upsample_fp32(m)
        return m[0].contiguous(), aux[0].contiguous()

    def upsample_fp32(self, m):
        """UpsampleNetwork in true fp32: cuDNN's default TF32 convolutions would put ~1e-3 relative
        error into the conditioning, visible as 4e-5 on the logits (measured on B200)."""
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            prev = torch.backends.cuda.matmul.allow_tf32
            torch.backends.cuda.matmul.allow_tf32 = False
            try:
                return self.upsample(m)
            finally:
                torch.backends.cuda.matmul.allow_tf32 = prev

    def _frames_mode(self, eng=None):
        dense = self.precision == "bf16-dense" if eng is None else eng.cfg.precision == _lib.PRECISION["bf16-dense"]
        return dense and self.expand_in_kernel

    # folds above which precision="auto" switches from the fp32 FFMA kernels (exact fp32, one launch of <= 21 folds per
    # ~15 us step) to the dense tcgen05 kernel (bf16 products, 32 folds per cluster, 480 in flight per ~18 us step)
    AUTO_DENSE_FOLDS = 64

    def _pick_engine(self, eng, device, nfolds):
        """Regime selection (north star: tensor cores only when the fold batch makes the step matmul dense).
        precision "auto": fp32 up to AUTO_DENSE_FOLDS folds, the dense kernel beyond.  An explicit fp32 / bf16 call with a large
        fold batch still runs (in serial launches) but says so loudly."""
        if self.precision == "auto":
            want = "bf16-dense" if nfolds > self.AUTO_DENSE_FOLDS and self._dense_supported() else "fp32"
            return self._engine(device, want)
        if self.precision in ("fp32", "bf16") and nfolds > self.AUTO_DENSE_FOLDS:
            import warnings
            warnings.warn("WaveRNN.generate: %d folds on the %s FFMA kernels run as %d serial launches; precision='bf16-dense' "
                          "(or 'auto') advances them together on the tensor cores at bf16 product precision"
                          % (nfolds, self.precision, -(-nfolds // max(1, int(eng.info().max_folds_per_launch)))), RuntimeWarning, stacklevel=4)
        return eng

    def _side_stream(self, device, name):
        """Side streams are kept across calls: the caching allocator keeps a pool per stream, and a fresh stream per call starts
        with an empty one (cudaMalloc for every conditioning tensor: ~20 ms per generate_many call)."""
        key = ("stream", name, device.index if device.index is not None else torch.cuda.current_device())
        st = self._pinned_cache.get(key)
        if st is None:
            st = self._pinned_cache[key] = torch.cuda.Stream(device)
        return st

    def _pinned(self, slot, n, dtype=torch.float64):
        """Pinned buffer number `slot` of at least n elements, kept across calls (cudaHostAlloc of a chunk's 80 MB costs
        ~25 ms, as much as its copy); the caller gets a fresh numpy COPY of the samples, so the buffer can be reused."""
        buf = self._pinned_cache.get(slot)
        if buf is None or buf.numel() < n or buf.dtype != dtype:
            buf = self._pinned_cache[slot] = torch.empty(max(n, 1), dtype=dtype, pin_memory=True)
        return buf[:n]

    def _dense_supported(self):
        return (self.mode == 'RAW' and self.n_classes == 512) or self.mode == 'MOL'

    def interp_table(self, device=None):
        """Composite response of UpsampleNetwork's three (Stretch2d, Conv2d box filter) stages (fatchord_version.py:70-77,83-85)
        as a [hop, 5] fp32 table: for a sample at phase r = (p + pad*hop) % hop of padded frame q = (p + pad*hop) // hop,
        mel_up[p] = sum_{j<4} table[r, j] * mel_pad[q + table[r, 4] + j].  Measured by pushing an impulse through the layers
        themselves, so trained filter weights are honoured.  The crop by pad*hop (:85) keeps the zero padding of the intermediate
        stages out of every retained sample, which is what makes the response phase-periodic."""
        convs = [self.upsample.up_layers[i] for i in (1, 3, 5)]
        device = convs[0].weight.device if device is None else device
        with torch.no_grad():                                               # 2*sum(scales)+3 floats: keyed on their values, not versions
            tag = (torch.cat([c.weight.detach().flatten() for c in convs]).cpu().numpy().tobytes(), str(device))
        if self._interp_cache is not None and self._interp_cache[0] == tag:
            return self._interp_cache[1]
        hop, n, c = self.hop_length, 9, 4
        with torch.no_grad():
            x = torch.zeros(1, 1, 1, n, dtype=torch.float32, device=convs[0].weight.device)
            x[..., c] = 1.0
            with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
                y = x
                for layer in self.upsample.up_layers:
                    y = layer(y)
            h = y.reshape(n, hop).cpu()                                     # h[k, r]: response at phase r of frame k
        full = torch.stack([h[c - j] for j in (-2, -1, 0, 1, 2)], dim=1)    # [hop, 5]: weight of frame q + j
        first = torch.where(full[:, 0] != 0, torch.full((hop,), -2.0), torch.full((hop,), -1.0))
        table = torch.zeros(hop, 5, dtype=torch.float32)
        for r in range(hop):
            s0 = int(first[r].item()) + 2
            if s0 == 0 and full[r, 4] != 0:
                raise RuntimeError("interpolation support wider than four frames at phase %d" % r)
            table[r, :4] = full[r, s0:s0 + 4]
            table[r, 4] = first[r]
        if float(h[:c - 2].abs().max()) != 0.0 or float(h[c + 3:].abs().max()) != 0.0:
            raise RuntimeError("interpolation support wider than the probe")
        table = table.to(device)
        self._interp_cache = (tag, table)
        return table

    # ------------------------------------------------------------------ frame-rate conditioning (csrc/wavernn_cond.cuh)
    MELRESNET_KEYS = ("conv_in.weight", "batch_norm.weight", "batch_norm.bias", "batch_norm.running_mean", "batch_norm.running_var",
                      "conv_out.weight", "conv_out.bias")

    def melresnet_native(self):
        """True when the MelResNet has the reference's dimensions (hparams.py:35-39: 80 mel channels, kernel 2 pad + 1 = 5, 128 compute
        and output channels), which is what the CUDA kernel is written for; other shapes run through PyTorch on the GPU."""
        rn = self.upsample.resnet
        return (rn.conv_in.in_channels == 80 and rn.conv_in.kernel_size[0] == 5 and rn.conv_in.out_channels == 128
                and rn.conv_out.out_channels == 128 and self.pad == 2 and os.environ.get("WRNN_MELRESNET", "native") != "torch")

    def pack_melresnet(self):
        """The weight blob of wrnn_cond_create (layout: include/wavernn_b200.h): convolution weights transposed to [in][out], eval-mode
        batch norm (fatchord_version.py:13-24, 36-44) folded in float64 to scale = w / sqrt(var + eps), shift = b - mean * scale."""
        rn = self.upsample.resnet
        sd = {k: v.detach().to("cpu", torch.float64).numpy() for k, v in rn.state_dict().items() if v.is_floating_point()}

        def bn(prefix, eps):
            scale = sd[prefix + ".weight"] / np.sqrt(sd[prefix + ".running_var"] + eps)
            return [scale, sd[prefix + ".bias"] - sd[prefix + ".running_mean"] * scale]

        parts = [sd["conv_in.weight"].transpose(2, 1, 0).reshape(-1)] + bn("batch_norm", rn.batch_norm.eps)     # [tap][in][out]
        for i, layer in enumerate(rn.layers):
            parts += [sd["layers.%d.conv1.weight" % i][:, :, 0].T.reshape(-1)] + bn("layers.%d.batch_norm1" % i, layer.batch_norm1.eps)
            parts += [sd["layers.%d.conv2.weight" % i][:, :, 0].T.reshape(-1)] + bn("layers.%d.batch_norm2" % i, layer.batch_norm2.eps)
        parts += [sd["conv_out.weight"][:, :, 0].T.reshape(-1), sd["conv_out.bias"]]
        return np.concatenate([np.asarray(x, dtype=np.float64).reshape(-1) for x in parts]).astype(np.float32)

    def _cond(self, device):
        idx = device.index if device.index is not None else torch.cuda.current_device()
        c = self._conds.get(idx)
        sd = self.upsample.resnet.state_dict()
        tag = tuple((v.data_ptr(), v._version) for v in sd.values())
        if c is None or c.tag != tag:
            with torch.cuda.device(idx):
                c = self._conds[idx] = _Cond(self.pack_melresnet(), len(self.upsample.resnet.layers), idx)
            c.tag = tag
        return c

    def conditioning_frames_many(self, mel_list, device):
        """Frame-rate prologue of SEVERAL utterances in one host-to-device copy and one kernel launch:
        [(1, feat, T_i)] -> mel frames [sum (T_i + 2 pad), feat] (zero-padded per utterance), MelResNet output [sum T_i, 4 aux],
        row offsets of every utterance in both.  A frame's result does not depend on what it is pooled with."""
        Ts = [int(m.shape[-1]) for m in mel_list]
        mel_rows = np.concatenate([[0], np.cumsum([T + 2 * self.pad for T in Ts])]).astype(np.int64)
        aux_rows = np.concatenate([[0], np.cumsum(Ts)]).astype(np.int64)
        af = torch.empty(int(aux_rows[-1]), 4 * self.aux_dims, dtype=torch.float32, device=device)
        seg = np.stack([mel_rows[:-1], np.asarray(Ts, dtype=np.int64), aux_rows[:-1]], axis=1)
        if all(m.is_cuda for m in mel_list):                       # inputs already resident in HBM: assemble on the device
            mf = torch.zeros(int(mel_rows[-1]), self._feat_dims, dtype=torch.float32, device=device)
            for i, m in enumerate(mel_list):
                mf[mel_rows[i] + self.pad: mel_rows[i] + self.pad + Ts[i]].copy_(m.detach()[0].t())
            self._cond(device).frames(mf, seg, af)
            return mf, af, mel_rows, aux_rows
        host = self._pinned("cond_frames", int(mel_rows[-1]) * self._feat_dims, dtype=torch.float32).view(int(mel_rows[-1]), self._feat_dims)
        hv = host.numpy()
        ev = self._pinned_cache.get("cond_frames_event")
        if ev is not None:
            ev.synchronize()                       # the previous call's copy has read the staging buffer
        hv[:] = 0.0
        for i, m in enumerate(mel_list):
            mm = m.detach()
            if mm.is_cuda:
                mm = mm.cpu()
            hv[mel_rows[i] + self.pad: mel_rows[i] + self.pad + Ts[i]] = mm.to(torch.float32).numpy()[0].T
        mf = host.to(device, non_blocking=True)
        ev = self._pinned_cache["cond_frames_event"] = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(device))
        self._cond(device).frames(mf, seg, af)
        return mf, af, mel_rows, aux_rows

    def conditioning_frames(self, mels):
        """Frame-rate view of the prologue (fatchord_version.py:162-165) for the in-kernel expansion:
        (1, feat, T) -> zero-padded mel frames [T + 2*pad, feat] and MelResNet output [T, 4*aux]."""
        if mels.is_cuda and self.melresnet_native():
            m = F.pad(mels.to(torch.float32), (self.pad, self.pad))[0].t().contiguous()
            T = int(mels.shape[-1])
            aux = torch.empty(T, 4 * self.aux_dims, dtype=torch.float32, device=mels.device)
            self._cond(mels.device).frames(m, np.array([[0, T, 0]], dtype=np.int32), aux)
            return m, aux
        m = F.pad(mels, (self.pad, self.pad))
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            prev = torch.backends.cuda.matmul.allow_tf32
            torch.backends.cuda.matmul.allow_tf32 = False
            try:
                aux = self.upsample.resnet(m)
            finally:
                torch.backends.cuda.matmul.allow_tf32 = prev
        return m[0].t().contiguous(), aux[0].t().contiguous()

    def _run_folds_frames(self, eng, device, mel_frames, aux_frames, geo, S, uniforms, seed, forced_x, return_logits, wait=True):
        """wrnn_generate_folds_frames: geo is a host int32 [B, 4] array (sample0, utterance samples, mel frame row, aux frame row)."""
        geo = np.ascontiguousarray(geo, dtype=np.int32)
        B = geo.shape[0]

        def dev(t, shape, what):
            if t is None:
                return None
            t = torch.as_tensor(t).to(device=device, dtype=torch.float32).contiguous()
            if tuple(t.shape) != shape:
                raise ValueError("%s must have shape %s, got %s" % (what, shape, tuple(t.shape)))
            return t

        n_u = 1 if self.mode == 'RAW' else self.n_classes // 3 + 1
        u = dev(uniforms, (S, B) if n_u == 1 else (S, B, n_u), "uniforms")
        fx = dev(forced_x, (S, B), "forced_x")
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        table = self.interp_table(device)
        samples = torch.empty(B, S, dtype=torch.float32, device=device)
        labels = torch.empty(B, S, dtype=torch.int32, device=device)
        logits = torch.empty(S, B, self.n_classes, dtype=torch.float32, device=device) if return_logits else None
        stream = torch.cuda.current_stream(device).cuda_stream
        ptr = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
        _lib.check(eng.lib.wrnn_generate_folds_frames(
            eng.handle, ptr(mel_frames), mel_frames.size(0), ptr(aux_frames), aux_frames.size(0), ptr(table),
            self.hop_length, self.pad, geo.ctypes.data_as(ctypes.c_void_p), B, S,
            ptr(u), ctypes.c_uint64(seed), ptr(fx), ptr(logits), ptr(samples), ptr(labels), ctypes.c_void_p(stream)))
        if wait:
            eng.synchronize()
        return dict(samples=samples, labels=labels, logits=logits)

    # ------------------------------------------------------------------ the hot path
    def generate(self, mels, *args, uniforms=None, seed=None, forced_x=None, return_logits=False,
                 return_samples=False, **kwargs):
        """generate(mels, batched, target, overlap, mu_law)            -- this repo (fatchord_version.py:150)
           generate(mels, save_path, batched, target, overlap, mu_law) -- upstream form, gen_wavernn.py:34,51

        Returns the waveform as float64 numpy of length (T-1)*hop_length.  Extra keyword-only
        arguments: `uniforms` (pre-drawn U[0,1): [S,B] RAW / [S,B,11] MOL), `seed` (in-kernel
        Philox when no uniforms are given), `forced_x` [S,B] + `return_logits` (teacher forcing),
        `return_samples` (also return the per-fold samples / labels).
        """
        names = ("batched", "target", "overlap", "mu_law")
        save_path = kwargs.pop("save_path", None)
        args = list(args)
        if args and (args[0] is None or isinstance(args[0], (str, os.PathLike))):
            save_path = args.pop(0)
        if len(args) > 4:
            raise TypeError("generate() takes at most 6 positional arguments")
        call = dict(zip(names, args))
        for k in names:
            if k in kwargs:
                if k in call:
                    raise TypeError("generate() got multiple values for argument '%s'" % k)
                call[k] = kwargs.pop(k)
        if kwargs:
            raise TypeError("generate() got unexpected keyword arguments %s" % sorted(kwargs))
        missing = [k for k in names if k not in call]
        if missing:
            raise TypeError("generate() missing required arguments: %s" % ", ".join(missing))
        batched, target, overlap, mu_law = (call[k] for k in names)

        mu_law = mu_law if self.mode == 'RAW' else False             # :152
        self.eval()                                                   # :154
        t_start = time.perf_counter()
        try:
            with torch.no_grad():
                device = self._device()
                with torch.cuda.device(device):
                    eng = self._engine(device)
                    out = self._generate_on_device(eng, device, mels, bool(batched), int(target), int(overlap),
                                                   bool(mu_law), uniforms, seed, forced_x, return_logits)
        finally:
            self.train()                                              # :241 (regardless of the prior mode)
        wav_dev, extras, host_mu = out
        wav = wav_dev.cpu().numpy()                                   # D2H, the reference's :223
        if host_mu:
            # decode_mu_law (dsp.py:100-105) and the tail fade (fatchord_version.py:235-237) with the reference's own numpy
            # expressions on the crossfaded float64 signal: bit-exact, where CUDA's pow is within 2 ulp of numpy's
            mu = self.n_classes - 1
            wav = decode_mu_law_host(wav, mu)
            wav[-20 * self.hop_length:] *= np.linspace(1, 0, 20 * self.hop_length)
        self.last_stats["wall_s"] = time.perf_counter() - t_start
        if save_path is not None:
            save_wav(wav, save_path, self.sample_rate)                # upstream :239
        if return_logits or return_samples:
            return wav, extras
        return wav

    def _generate_on_device(self, eng, device, mels, batched, target, overlap, mu_law, uniforms, seed,
                            forced_x, return_logits):
        if mels.dim() != 3 or mels.size(0) != 1:
            raise RuntimeError("mels must have shape (1, feat_dims, T), got %s" % (tuple(mels.shape),))
        mels = mels.to(device=device, dtype=torch.float32)            # H2D, :162
        if self.upsample.resnet.conv_in.weight.device != device:
            self.to(device)
        T = mels.size(-1)
        wave_len = (T - 1) * self.hop_length                          # :163
        tail = 20 * self.hop_length
        if wave_len < tail:
            raise ValueError("operands could not be broadcast together: wave_len %d < 20*hop_length %d "
                             "(the reference needs T >= 21 frames, fatchord_version.py:235-237)" % (wave_len, tail))
        if batched:
            nfolds, _ = _lib.fold_index(T * self.hop_length, target, overlap)
            eng = self._pick_engine(eng, device, nfolds)
        frames = self._frames_mode(eng)
        if frames:
            mel_fr, aux_fr = self.conditioning_frames(mels)           # :164 at frame rate; :165 happens inside the kernel
            L = T * self.hop_length
        else:
            m_up, aux = self.conditioning(mels)                       # :164-165
            L = m_up.size(0)
        if batched:
            B, _ = _lib.fold_index(L, target, overlap)                # :298-309
            if B <= 0:
                raise RuntimeError("fold_with_overlap yields no folds: total_len %d <= overlap %d" % (L, overlap))
            S = target + 2 * overlap
            starts = np.arange(B, dtype=np.int64) * (target + overlap)   # :315
        else:
            B, S = 1, L
            starts = np.zeros(1, dtype=np.int64)
        if frames:
            geo = np.stack([starts, np.full(B, L), np.zeros(B), np.zeros(B)], axis=1)
            res = self._run_folds_frames(eng, device, mel_fr, aux_fr, geo, S, uniforms, seed, forced_x, return_logits, wait=False)
        else:
            limits = np.full(B, L, dtype=np.int64)
            res = self._run_folds(eng, device, m_up, aux, starts, limits, S, uniforms, seed, forced_x, return_logits, wait=False)
        if self.mu_law_decode not in ("auto", "host", "device"):
            raise ValueError("mu_law_decode must be 'auto', 'host' or 'device', got %r" % (self.mu_law_decode,))
        host_mu = bool(mu_law) and (self.mu_law_decode == "host" or (self.mu_law_decode == "auto" and wave_len < 2_000_000))
        wav = torch.empty(wave_len, dtype=torch.float64, device=device)
        stream = torch.cuda.current_stream(device).cuda_stream
        # host_mu: the kernel stops after the crossfade + trim; mu-law decode and tail fade follow in numpy (generate())
        _lib.check(eng.lib.wrnn_xfade_unfold(res["samples"].data_ptr(), B, S, int(batched), overlap if batched else 0,
                                             self.n_classes if (mu_law and not host_mu) else 0, wave_len,
                                             0 if host_mu else tail, wav.data_ptr(), ctypes.c_void_p(stream)))
        if self.progress is not None:                                 # step loop + epilogue are enqueued: report while they run
            while True:
                done, steps_done = eng.query()
                if done:
                    break
                self.progress(steps_done, S)
                time.sleep(0.1)
            self.progress(S, S)
        eng.synchronize()                                             # a fired watchdog raises here
        info = eng.info()
        self.last_stats.update(folds=B, steps=S, wave_len=wave_len, kernel_ms=info.last_kernel_ms, kernel_kind=info.kernel_kind)
        return wav, res, host_mu

    def _run_folds(self, eng, device, m_up, aux, starts, limits, S, uniforms, seed, forced_x, return_logits, wait=True):
        """wrnn_generate_folds over conditioning rows; starts/limits are host int64 arrays."""
        B = len(starts)
        n_u = 1 if self.mode == 'RAW' else self.n_classes // 3 + 1

        def dev(t, shape, what):
            if t is None:
                return None
            t = torch.as_tensor(t).to(device=device, dtype=torch.float32).contiguous()
            if tuple(t.shape) != shape:
                raise ValueError("%s must have shape %s, got %s" % (what, shape, tuple(t.shape)))
            return t

        u = dev(uniforms, (S, B) if n_u == 1 else (S, B, n_u), "uniforms")
        fx = dev(forced_x, (S, B), "forced_x")
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())       # follows torch's global RNG like the reference
        samples = torch.empty(B, S, dtype=torch.float32, device=device)
        labels = torch.empty(B, S, dtype=torch.int32, device=device)
        logits = torch.empty(S, B, self.n_classes, dtype=torch.float32, device=device) if return_logits else None
        starts = np.ascontiguousarray(starts, dtype=np.int64)
        limits = np.ascontiguousarray(limits, dtype=np.int64)
        stream = torch.cuda.current_stream(device).cuda_stream
        ptr = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
        _lib.check(eng.lib.wrnn_generate_folds(
            eng.handle, ptr(m_up), ptr(aux), m_up.size(0),
            starts.ctypes.data_as(ctypes.c_void_p), limits.ctypes.data_as(ctypes.c_void_p), B, S,
            ptr(u), ctypes.c_uint64(seed), ptr(fx), ptr(logits), ptr(samples), ptr(labels),
            ctypes.c_void_p(stream)))
        if wait:
            eng.synchronize()
        return dict(samples=samples, labels=labels, logits=logits)    # u / fx return to torch's stream-ordered allocator: safe on this stream

    def generate_many(self, mel_list, target, overlap, mu_law, uniforms=None, seed=None, writer=None):
        """Batched generation of several utterances with their folds POOLED into one launch
        sequence: the reference's callers loop one sentence at a time
        (synthesize_sentences.py:63-73) and only ever see B = folds of one utterance.
        Returns a list of float64 waveforms, each equal to generate(mel, True, target, overlap, mu_law)
        given the matching slice of `uniforms` ([S, total_folds(,11)], utterance-major fold order).
        Chunks are pipelined: the conditioning network of chunk k+1 and the device-to-host copy of chunk k-1 run under the step
        loop of chunk k.  `writer(i, wav)`, if given, receives every utterance once its waveform is on the host (e.g. to stream
        it into a wavio.WavWriter)."""
        mu_law = mu_law if self.mode == 'RAW' else False
        self.eval()
        try:
            with torch.no_grad():
                device = self._device()
                with torch.cuda.device(device):
                    eng = self._engine(device)
                    if self.upsample.resnet.conv_in.weight.device != device:
                        self.to(device)
                    S = target + 2 * overlap
                    # plan: fold counts from the mel lengths alone (fold_with_overlap index arithmetic, :298-309)
                    plan = []
                    for mel in mel_list:
                        wave_len = (mel.size(-1) - 1) * self.hop_length
                        if wave_len < 20 * self.hop_length:
                            raise ValueError("utterance shorter than 21 frames")
                        L = mel.size(-1) * self.hop_length                                   # rows the conditioning network returns
                        B, _ = _lib.fold_index(L, target, overlap)
                        if B <= 0:
                            raise RuntimeError("utterance yields no folds")
                        plan.append((B, L, wave_len))
                    total_folds = sum(p[0] for p in plan)
                    if uniforms is not None:
                        uniforms = torch.as_tensor(uniforms)
                        if uniforms.size(1) != total_folds:
                            raise ValueError("uniforms must have %d fold columns, got %d" % (total_folds, uniforms.size(1)))
                    # Utterances are pooled in CHUNKS of about two waves of the step-loop kernel: the upsampled conditioning of a chunk
                    # (832 B per sample) is produced, consumed and its memory reused, instead of materialising it for the whole set
                    # (34 GB for BASELINE.json configs[3]); fold results do not depend on how folds are pooled (tests: pooling /
                    # chunking / placement invariance), so the waveforms are the ones of the unchunked call.
                    eng = self._pick_engine(eng, device, total_folds)
                    cap = max(1, int(eng.info().max_folds_per_launch))
                    goal = cap * max(1, round(960 / cap))
                    chunks, cur, cur_folds = [], [], 0
                    for idx, (B, L, wave_len) in enumerate(plan):
                        if cur and cur_folds + B > goal:
                            chunks.append(cur)
                            cur, cur_folds = [], 0
                        cur.append(idx)
                        cur_folds += B
                    if cur:
                        chunks.append(cur)
                    frames = self._frames_mode(eng)
                    main = torch.cuda.current_stream(device)
                    copy_stream = self._side_stream(device, "copy")
                    if self.mu_law_decode not in ("auto", "host", "device"):
                        raise ValueError("mu_law_decode must be 'auto', 'host' or 'device', got %r" % (self.mu_law_decode,))
                    total_samples = sum(p[2] for p in plan)
                    host_mu = bool(mu_law) and (self.mu_law_decode == "host" or (self.mu_law_decode == "auto" and total_samples < 2_000_000))
                    fold_base = np.concatenate([[0], np.cumsum([p[0] for p in plan])])

                    prep_stream = self._side_stream(device, "prep")

                    def prepare(ci):
                        # Conditioning of chunk ci (PyTorch) on a SIDE stream: the host work overlaps the previous chunk's step loop and
                        # its kernels run beside it (the dense kernel leaves 28 SMs free).  Copies are non-blocking: a blocking .to() of a
                        # pinned mel waits for everything queued on its stream, i.e. for the whole step loop (measured: 515 ms per chunk).
                        with torch.cuda.stream(prep_stream):
                            job = prepare_on_stream(ci)
                            job["ready"] = torch.cuda.Event()
                            job["ready"].record(prep_stream)
                        return job

                    def prepare_on_stream(ci):
                        chunk = chunks[ci]
                        f0, f1 = int(fold_base[chunk[0]]), int(fold_base[chunk[-1] + 1])
                        job = dict(chunk=chunk, u=None if uniforms is None else uniforms[:, f0:f1].contiguous(),
                                   seed=None if seed is None else int(seed) + ci)          # in-kernel draws: one Philox stream per chunk
                        if frames and self.melresnet_native():
                            mf, af, mrows, arows = self.conditioning_frames_many([mel_list[i] for i in chunk], device)
                            geo = []
                            for j, i in enumerate(chunk):
                                B, L, _ = plan[i]
                                g = np.zeros((B, 4), dtype=np.int64)
                                g[:, 0] = np.arange(B, dtype=np.int64) * (target + overlap)
                                g[:, 1], g[:, 2], g[:, 3] = L, mrows[j], arows[j]
                                geo.append(g)
                            job.update(m=mf, a=af, geo=np.concatenate(geo))
                        elif frames:
                            mfs, afs, geo, mb, ab = [], [], [], 0, 0
                            for i in chunk:
                                B, L, _ = plan[i]
                                mf, af = self.conditioning_frames(mel_list[i].to(device=device, dtype=torch.float32, non_blocking=True))
                                mfs.append(mf)
                                afs.append(af)
                                g = np.zeros((B, 4), dtype=np.int64)
                                g[:, 0] = np.arange(B, dtype=np.int64) * (target + overlap)
                                g[:, 1], g[:, 2], g[:, 3] = L, mb, ab
                                geo.append(g)
                                mb += mf.size(0)
                                ab += af.size(0)
                            job.update(m=torch.cat(mfs), a=torch.cat(afs), geo=np.concatenate(geo))
                        else:
                            rows = sum(plan[i][1] for i in chunk)
                            m_all = torch.empty(rows, self._feat_dims, dtype=torch.float32, device=device)
                            a_all = torch.empty(rows, 4 * self.aux_dims, dtype=torch.float32, device=device)
                            starts, limits, base = [], [], 0
                            for i in chunk:
                                B, L, _ = plan[i]
                                m_up, aux = self.conditioning(mel_list[i].to(device=device, dtype=torch.float32, non_blocking=True))
                                if m_up.size(0) != L:
                                    raise RuntimeError("conditioning network returned %d rows, expected %d" % (m_up.size(0), L))
                                m_all[base:base + L].copy_(m_up)
                                a_all[base:base + L].copy_(aux)
                                del m_up, aux
                                starts.append(base + np.arange(B, dtype=np.int64) * (target + overlap))
                                limits.append(np.full(B, base + L, dtype=np.int64))
                                base += L
                            job.update(m=m_all, a=a_all, starts=np.concatenate(starts), limits=np.concatenate(limits))
                        return job

                    def launch(job):
                        # Enqueue the step loop of a chunk, the epilogue of each of its utterances into one device buffer, and the
                        # device-to-host copy of that buffer (pinned, on a second stream) behind them.
                        chunk = job["chunk"]
                        main.wait_event(job["ready"])                      # the chunk's conditioning (side stream) is complete
                        for t in (job["m"], job["a"]):
                            t.record_stream(main)                          # allocated on the side stream, consumed on this one
                        if frames:
                            res = self._run_folds_frames(eng, device, job["m"], job["a"], job["geo"], S, job["u"], job["seed"], None, False, wait=False)
                        else:
                            res = self._run_folds(eng, device, job["m"], job["a"], job["starts"], job["limits"], S, job["u"], job["seed"], None, False, wait=False)
                        n = sum(plan[i][2] for i in chunk)
                        wav_dev = torch.empty(n, dtype=torch.float64, device=device)
                        b0 = w0 = 0
                        for i in chunk:
                            B, _, wave_len = plan[i]
                            _lib.check(eng.lib.wrnn_xfade_unfold(res["samples"][b0:b0 + B].data_ptr(), B, S, 1, overlap,
                                                                 self.n_classes if (mu_law and not host_mu) else 0, wave_len,
                                                                 0 if host_mu else 20 * self.hop_length, wav_dev[w0:w0 + wave_len].data_ptr(),
                                                                 ctypes.c_void_p(main.cuda_stream)))
                            b0 += B
                            w0 += wave_len
                        done = torch.cuda.Event()
                        done.record(main)
                        wav_host = self._pinned(len(jobs) % 3, n)            # three landing buffers in rotation: chunk k-2 has been collected
                        copied = torch.cuda.Event()
                        with torch.cuda.stream(copy_stream):
                            copy_stream.wait_event(done)
                            wav_host.copy_(wav_dev, non_blocking=True)
                            copied.record(copy_stream)
                        job.update(res=res, wav_dev=wav_dev, wav_host=wav_host, copied=copied)
                        return job

                    outs = [None] * len(plan)

                    def collect(job):
                        # host side of a finished chunk (runs while a later chunk's step loop occupies the GPU): wait for its copy,
                        # take the utterances out of the reused pinned buffer, host mu-law decode if selected, hand them to `writer`
                        job["copied"].synchronize()
                        flat, w0 = job["wav_host"].numpy(), 0
                        for i in job["chunk"]:
                            wave_len = plan[i][2]
                            if host_mu:                                  # the decode writes a fresh array: no copy out of the pinned buffer first
                                wav = decode_mu_law_host(flat[w0:w0 + wave_len], self.n_classes - 1)
                                wav[-20 * self.hop_length:] *= np.linspace(1, 0, 20 * self.hop_length)
                            else:
                                wav = flat[w0:w0 + wave_len].copy()
                            outs[i] = wav
                            if writer is not None:
                                writer(i, wav)
                            w0 += wave_len
                        job.pop("wav_dev", None)

                    trace = [] if os.environ.get("WRNN_TRACE") else None     # development: host timeline of the pipeline
                    def mark(what):
                        if trace is not None:
                            trace.append((what, time.perf_counter()))
                    mark("start")
                    jobs, kernel_ms = [], 0.0
                    nxt = prepare(0)
                    mark("prepare 0")
                    for ci in range(len(chunks)):
                        if ci > 0:
                            kernel_ms += eng.info().last_kernel_ms          # waits for the previous chunk's step loop (as the launch below would)
                            mark("wait kernel %d" % (ci - 1))
                        cur = launch(nxt)
                        jobs.append(cur)
                        mark("launch %d" % ci)
                        if ci + 1 < len(chunks):
                            nxt = prepare(ci + 1)                            # overlaps the step loop just enqueued
                            mark("prepare %d" % (ci + 1))
                        if ci > 0:                                           # the chunk before: its step loop is done, its inputs can go
                            for k in ("m", "a", "res", "u"):
                                jobs[ci - 1].pop(k, None)
                            collect(jobs[ci - 1])
                            mark("collect %d" % (ci - 1))
                    eng.synchronize()
                    kernel_ms += eng.info().last_kernel_ms
                    mark("wait kernel %d" % (len(chunks) - 1))
                    collect(jobs[-1])
                    mark("collect %d" % (len(chunks) - 1))
                    if trace is not None:
                        print("generate_many timeline (ms): " + ", ".join("%s +%.1f" % (w, (t - trace[i][1]) * 1e3) for i, (w, t) in enumerate(trace[1:])), flush=True)
                    self.last_stats.update(folds=total_folds, steps=S, kernel_ms=kernel_ms, chunks=len(chunks), kernel_kind=eng.info().kernel_kind)
                    return outs
        finally:
            self.train()

    # ------------------------------------------------------------------ reference housekeeping API
    def get_step(self):
        return self.step.data.item()

    def checkpoint(self, path):
        self.save('%s/checkpoint_%dk_steps.pyt' % (path, self.get_step() // 1000))

    def log(self, path, msg):
        with open(path, 'a') as f:
            print(msg, file=f)

    def restore(self, path):
        if not os.path.exists(path):
            self.save(path)
        else:
            self.load(path)

    def load(self, path):
        self.load_state_dict(torch.load(path, map_location="cpu"), strict=False)

    def save(self, path):
        torch.save(self.state_dict(), path)

    def num_params(self, print_out=True):
        n = sum(int(np.prod(p.size())) for p in self.parameters() if p.requires_grad) / 1_000_000
        if print_out:
            print('Trainable Parameters: %.3fM' % n)
        return n
