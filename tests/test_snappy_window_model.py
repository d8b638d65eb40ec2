"""The windowed Snappy algorithm of k_decompress.cu, as modelled lane by lane in tools/snappy_window_model.py, decodes
what pyarrow's Snappy encoder writes and never needs the serial fallback on valid input (CPU only)."""
import os
import random
import sys

import numpy as np
import pyarrow as pa
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))
import snappy_window_model as model


def _cases():
    rng = random.Random(1)
    words = ["the", "quick", "brown", "fox", "jumps", "over", "lazy", "dog", "furiously", "final", "requests", "pending",
             "deposits", "carefully", "ironic", "accounts"]
    yield "text", " ".join(rng.choice(words) for _ in range(3000)).encode()
    yield "random", bytes(rng.randrange(256) for _ in range(6000))              # one long literal
    yield "periods", b"a" * 700 + b"ab" * 500 + b"abc" * 400 + bytes(range(256)) * 6  # copies that overlap themselves
    yield "int64 column", np.array([rng.randrange(100) for _ in range(2500)], dtype=np.int64).tobytes()
    yield "two symbols", bytes(rng.choice(b"ab") for _ in range(8000))
    yield "runs", b"".join(bytes([rng.randrange(4)]) * rng.randrange(1, 70) for _ in range(600))
    yield "varints", b"".join(bytes([0x80 | rng.randrange(128), rng.randrange(8)]) if rng.random() < 0.6 else bytes([rng.randrange(128)])
                              for _ in range(9000))                              # far, 4-byte copies like decimal streams
    yield "empty", b""
    yield "one byte", b"x"


@pytest.mark.parametrize("name,data", list(_cases()), ids=[n for n, _ in _cases()])
def test_window_model_round_trip(name, data):
    comp = pa.Codec("snappy").compress(data, asbytes=True)
    for k in model.stats:
        model.stats[k] = 0
    assert model.decode(comp) == data
    if len(data) > 1000:
        assert model.stats["windows"] > 0 and model.stats["elements"] >= model.stats["windows"]
