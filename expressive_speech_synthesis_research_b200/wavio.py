"""Wav output stage: writers standing in for the removed librosa.output.write_wav that the reference's
save_wav used (WaveRNN/utility/dsp.py:21-22 float32; fatchord_version.py:239) and for its 16-bit path (encode_16bits,
dsp.py:36-37; synthesize_sentences.py:72 writes int16 through scipy).  `WavWriter` streams: generate_many hands every
utterance over as soon as its epilogue has landed, so a sentence set never sits in memory as a whole."""
import os
import struct

import numpy as np


def encode_16bits(x):
    """dsp.py:36-37: clip(x * 2**15, -2**15, 2**15 - 1).astype(int16)."""
    return np.clip(np.asarray(x) * 2 ** 15, -2 ** 15, 2 ** 15 - 1).astype(np.int16)


class WavWriter:
    """Mono RIFF/WAVE writer that accepts samples in pieces; the sizes in the header are patched on close.
    encoding 'float32' (WAVE_FORMAT_IEEE_FLOAT, what librosa wrote for float input) or 'int16' (PCM via encode_16bits)."""

    def __init__(self, path, sample_rate, encoding="float32"):
        if encoding not in ("float32", "int16"):
            raise ValueError("encoding must be 'float32' or 'int16', got %r" % (encoding,))
        self.encoding, self.frames = encoding, 0
        self.width = 4 if encoding == "float32" else 2
        self.f = open(path, "wb")
        fmt = struct.pack('<HHIIHH', 3 if encoding == "float32" else 1, 1, int(sample_rate), int(sample_rate) * self.width,
                          self.width, 8 * self.width)
        self.f.write(b'RIFF' + struct.pack('<I', 0) + b'WAVE' + b'fmt ' + struct.pack('<I', len(fmt)) + fmt)
        if encoding == "float32":
            self.fact_at = self.f.tell() + 8
            self.f.write(b'fact' + struct.pack('<II', 4, 0))
        else:
            self.fact_at = None
        self.data_at = self.f.tell() + 4
        self.f.write(b'data' + struct.pack('<I', 0))

    def write(self, samples):
        x = np.asarray(samples)
        data = np.ascontiguousarray(x, dtype='<f4') if self.encoding == "float32" else encode_16bits(x).astype('<i2')
        self.f.write(data.tobytes())
        self.frames += data.size

    def close(self):
        if self.f is None:
            return
        nbytes = self.frames * self.width
        if nbytes & 1:
            self.f.write(b'\0')
        end = self.f.tell()
        self.f.seek(4)
        self.f.write(struct.pack('<I', end - 8))
        if self.fact_at is not None:
            self.f.seek(self.fact_at)
            self.f.write(struct.pack('<I', self.frames))
        self.f.seek(self.data_at)
        self.f.write(struct.pack('<I', nbytes))
        self.f.close()
        self.f = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def save_wav(x, path, sample_rate, encoding="float32"):
    """dsp.py:21-22 (`encoding='float32'`, the default) or the 16-bit form of synthesize_sentences.py:72 (`'int16'`)."""
    with WavWriter(path, sample_rate, encoding) as w:
        w.write(x)


def load_wav(path):
    """Inverse of save_wav (mono float32 or int16) -> (samples as stored, sample_rate)."""
    with open(path, 'rb') as f:
        blob = f.read()
    assert blob[:4] == b'RIFF' and blob[8:12] == b'WAVE'
    pos, rate, data, tag_fmt, bits = 12, None, None, None, None
    while pos + 8 <= len(blob):
        tag, size = blob[pos:pos + 4], struct.unpack('<I', blob[pos + 4:pos + 8])[0]
        chunk = blob[pos + 8:pos + 8 + size]
        if tag == b'fmt ':
            tag_fmt, _, rate = struct.unpack('<HHI', chunk[:8])
            bits = struct.unpack('<H', chunk[14:16])[0]
        elif tag == b'data':
            data = np.frombuffer(chunk, dtype='<f4' if (tag_fmt, bits) == (3, 32) else '<i2').copy()
        pos += 8 + size + (size & 1)
    return data, rate


# ---------------------------------------------------------------------------------------------
# bit-exact mu-law expansion on the host
# ---------------------------------------------------------------------------------------------
_DECODE_POOL = None


def decode_mu_law_host(y, mu):
    """decode_mu_law(y, mu + 1, from_labels=False), WaveRNN/utility/dsp.py:100-105, with numpy's own float64 `pow`: the expression
    is the reference's, evaluated slice by slice on a few threads (numpy releases the GIL inside the ufuncs; 220 000 samples cost
    5 ms on one thread, as much as 400 steps of the kernel).  Every slice but the last is a multiple of 64 elements long, so each
    element goes through the same vector / scalar loop of numpy as in one big call: the result is identical bit for bit
    (tests/test_wavio.py)."""
    global _DECODE_POOL
    y = np.ascontiguousarray(y, dtype=np.float64)
    n = y.size
    workers = min(8, os.cpu_count() or 1)
    if n < 32768 or workers < 2:
        return np.sign(y) / mu * ((1 + mu) ** np.abs(y) - 1)
    if _DECODE_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _DECODE_POOL = ThreadPoolExecutor(max_workers=workers, thread_name_prefix="wrnn-mulaw")
    out = np.empty(n, dtype=np.float64)
    step = -(-n // workers)
    step = -(-step // 64) * 64

    def part(a):
        b = min(n, a + step)
        out[a:b] = np.sign(y[a:b]) / mu * ((1 + mu) ** np.abs(y[a:b]) - 1)

    list(_DECODE_POOL.map(part, range(0, n, step)))
    return out


def parallel_copy(dst, src):
    """dst[:] = src for large 1-D arrays on a few threads (numpy releases the GIL while it copies): one thread moves ~6 GB/s, the
    106 MB waveform of a 10-minute utterance took 16 ms of a 250 ms call."""
    global _DECODE_POOL
    n = src.size
    workers = min(8, os.cpu_count() or 1)
    if n < (1 << 20) or workers < 2:
        np.copyto(dst, src)
        return dst
    if _DECODE_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _DECODE_POOL = ThreadPoolExecutor(max_workers=workers, thread_name_prefix="wrnn-mulaw")
    step = -(-n // workers)

    def part(a):
        np.copyto(dst[a:a + step], src[a:a + step])

    list(_DECODE_POOL.map(part, range(0, n, step)))
    return dst
