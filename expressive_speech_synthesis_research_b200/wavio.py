"""Wav output stage: float32 WAV writer standing in for the removed librosa.output.write_wav
that the reference's save_wav used (WaveRNN/utility/dsp.py:21-22; fatchord_version.py:239)."""
import struct

import numpy as np


def save_wav(x, path, sample_rate):
    """Write mono IEEE-float32 PCM (what librosa.output.write_wav produced for float input)."""
    data = np.ascontiguousarray(np.asarray(x, dtype=np.float32))
    payload = data.tobytes()
    fmt = struct.pack('<HHIIHH', 3, 1, int(sample_rate), int(sample_rate) * 4, 4, 32)   # WAVE_FORMAT_IEEE_FLOAT
    fact = struct.pack('<I', data.size)
    body = (b'WAVE' + b'fmt ' + struct.pack('<I', len(fmt)) + fmt + b'fact' + struct.pack('<I', 4) + fact
            + b'data' + struct.pack('<I', len(payload)) + payload)
    with open(path, 'wb') as f:
        f.write(b'RIFF' + struct.pack('<I', len(body)) + body)


def load_wav(path):
    """Inverse of save_wav (float32 mono) -> (samples float32, sample_rate)."""
    with open(path, 'rb') as f:
        blob = f.read()
    assert blob[:4] == b'RIFF' and blob[8:12] == b'WAVE'
    pos, rate, data = 12, None, None
    while pos + 8 <= len(blob):
        tag, size = blob[pos:pos + 4], struct.unpack('<I', blob[pos + 4:pos + 8])[0]
        chunk = blob[pos + 8:pos + 8 + size]
        if tag == b'fmt ':
            rate = struct.unpack('<HHI', chunk[:8])[2]
        elif tag == b'data':
            data = np.frombuffer(chunk, dtype='<f4').copy()
        pos += 8 + size + (size & 1)
    return data, rate
