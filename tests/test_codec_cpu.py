"""Host logic of the bitstream codec and the CPU restatement of its range coder (no GPU)."""
import random

import torch

from contextgs_b200.codec import frequency_tables
from oracle import codec_ref


def test_frequency_tables_are_strictly_increasing_16_bit():
    g = torch.Generator().manual_seed(0)
    for L in (2, 3, 17, 300):
        pmf = torch.rand(4, L, generator=g) ** 8        # very skewed, many near-zero entries
        pmf[0, L // 2] = 0.0
        tb = frequency_tables(pmf)
        assert tb.shape == (4, L + 1) and tb.dtype == torch.int32
        assert (tb[:, 0] == 0).all() and (tb[:, -1] == 65536).all()
        assert (tb[:, 1:] > tb[:, :-1]).all()


def test_range_coder_restatement_round_trips():
    rnd = random.Random(1)
    tables = [frequency_tables(torch.tensor([[0.9, 0.1]]))[0].tolist(),
              frequency_tables(torch.tensor([[0.02, 0.5, 0.3, 0.18]]))[0].tolist()]
    for n in (0, 1, 7, 5000):
        syms = [rnd.choices(range(2), weights=[0.9, 0.1])[0] if i % 2 == 0 else rnd.choices(range(4), weights=[2, 50, 30, 18])[0]
                for i in range(n)]
        data = codec_ref.encode(syms, lambda i: i % 2, tables)
        assert codec_ref.decode(data, n, lambda i: i % 2, tables) == syms
        if n == 5000:   # close to the entropy of the source (0.47 + 1.58 bits per pair)
            assert len(data) * 8 < 1.05 * n / 2 * (0.469 + 1.58) + 64


def test_phi_table_is_the_documented_function():
    """Host-only entry point: the coder's tabulated normal CDF is fp32(0.5 erfc(-z / sqrt 2)) on 4097 points of
    [-4.75, 4.75] with exact end points, as include/contextgs_b200.h documents (no GPU involved)."""
    import math
    import numpy as np
    from contextgs_b200 import codec
    T, z0, inv_h = codec.phi_table()
    want = np.array([0.5 * math.erfc(-(-4.75 + 9.5 * j / 4096) / math.sqrt(2.0)) for j in range(4097)]).astype(np.float32)
    want[0], want[-1] = 0.0, 1.0
    assert np.array_equal(T, want)
    assert z0 == -4.75 and inv_h == float(np.float32(4096 / 9.5))
    assert np.all(np.diff(T) >= 0) and abs(float(T[2048]) - 0.5) < 1e-7


def test_small_levels_get_short_chunks():
    from contextgs_b200 import codec
    mult = list(codec.ATTR_CHUNK_MULT)
    assert codec.level_chunk_rows(1_212_902, 8) == [8 * m for m in mult]          # enough chunks already
    assert codec.level_chunk_rows(227_251, 8) == [2 * m for m in mult]            # halved until 100 k feat chunks
    assert codec.level_chunk_rows(59_847, 8) == [codec.MIN_CHUNK_ROWS * m for m in mult]
    assert codec.level_chunk_rows(59_847, 8, adaptive=False) == [8 * m for m in mult]
    assert codec.level_chunk_rows(10, 1000) == [125 * m for m in mult]            # stops when the row count turns odd


def test_mask_table_is_the_same_from_a_tensor_and_from_the_stored_float():
    """The encoder builds the Bernoulli table from the fp32 device scalar, the decoder from the python float stored in the
    metadata: the two routes must give identical 16-bit frequencies for every probability."""
    from contextgs_b200 import codec
    g = torch.Generator().manual_seed(3)
    for p in torch.rand(200, generator=g).tolist() + [0.0, 1.0, 1e-12, 1 - 1e-7]:
        t = torch.tensor(p, dtype=torch.float32)
        a, b = codec.mask_table(t), codec.mask_table(float(t))
        assert torch.equal(a, b), p
        assert a[0, 0] == 0 and a[0, 2] == 65536 and 0 < a[0, 1] < 65536
