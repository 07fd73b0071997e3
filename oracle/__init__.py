"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the WaveRNN generation hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may
import anything from this package; the product package never does.
"""
