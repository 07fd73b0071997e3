"""CPU tests of the wav output stage (dsp.py:21-22, 36-37; synthesize_sentences.py:72)."""
import numpy as np

from expressive_speech_synthesis_research_b200 import wavio


def test_float32_and_int16_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-1, 1, 1001), [1.0, -1.0, 1.5, -1.5, 0.0]])          # odd length, values beyond full scale
    p = str(tmp_path / "f.wav")
    wavio.save_wav(x, p, 22050)
    got, rate = wavio.load_wav(p)
    assert rate == 22050 and got.dtype == np.float32 and np.array_equal(got, x.astype(np.float32))
    p = str(tmp_path / "i.wav")
    wavio.save_wav(x, p, 16000, encoding="int16")
    got, rate = wavio.load_wav(p)
    want = np.clip(x * 2 ** 15, -2 ** 15, 2 ** 15 - 1).astype(np.int16)                  # encode_16bits, dsp.py:36-37
    assert rate == 16000 and got.dtype == np.int16 and np.array_equal(got, want)
    assert want.max() == 32767 and want.min() == -32768


def test_streaming_writer_equals_one_shot(tmp_path):
    rng = np.random.default_rng(1)
    parts = [rng.uniform(-1, 1, n) for n in (5, 1000, 1, 333)]
    for enc in ("float32", "int16"):
        a, b = str(tmp_path / ("a_%s.wav" % enc)), str(tmp_path / ("b_%s.wav" % enc))
        with wavio.WavWriter(a, 22050, enc) as w:
            for part in parts:
                w.write(part)
        wavio.save_wav(np.concatenate(parts), b, 22050, enc)
        assert open(a, "rb").read() == open(b, "rb").read()
