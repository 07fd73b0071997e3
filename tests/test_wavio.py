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


def test_threaded_mu_law_decode_is_bit_identical_to_the_reference_expression():
    """decode_mu_law (dsp.py:100-105) on several threads: same float64 values, bit for bit, as numpy's single call, at sizes around the
    slice boundaries; and fast enough to matter (the headline's end-to-end call spends 5 ms there on one thread)."""
    import time
    from expressive_speech_synthesis_research_b200.wavio import decode_mu_law_host
    rng = np.random.default_rng(0)
    mu = 511
    for n in (0, 1, 1000, 32767, 32768, 32769, 220550, 262144 + 7, 1_000_003):
        y = rng.uniform(-1, 1, n)
        y[::97] = 0.0
        want = np.sign(y) / mu * ((1 + mu) ** np.abs(y) - 1)
        got = decode_mu_law_host(y, mu)
        assert got.dtype == np.float64 and np.array_equal(got, want), n
    y = rng.uniform(-1, 1, 220550)
    decode_mu_law_host(y, mu)
    t0 = time.perf_counter()
    decode_mu_law_host(y, mu)
    t1 = time.perf_counter()
    np.sign(y) / mu * ((1 + mu) ** np.abs(y) - 1)
    t2 = time.perf_counter()
    print("mu-law decode of 220550 samples: %.2f ms threaded, %.2f ms one call" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))


def test_parallel_copy_copies_everything():
    from expressive_speech_synthesis_research_b200.wavio import parallel_copy
    rng = np.random.default_rng(1)
    for n in (0, 5, (1 << 20) - 1, (1 << 20) + 3, 3_000_001):
        src = rng.standard_normal(n)
        dst = np.full(n, np.nan)
        assert parallel_copy(dst, src) is dst and np.array_equal(dst, src)
