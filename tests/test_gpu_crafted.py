"""-m gpu: hand-crafted bitstreams (tests/drc_writer.py) through the CUDA path vs the oracle: irregular wrap
corrections, uncompressed integers, generic attributes, several attribute decoders, every max_bit_length,
degenerate alphabets, Tagged streams followed by more attributes."""
import numpy as np
import pytest

import drc_writer as W
from common import compare_with_oracle, gpu_decode_all

pytestmark = pytest.mark.gpu


def _attr(att_type, data_type, nc, seq_type, portable, xform=b"", uid=0):
    return dict(att_type=att_type, data_type=data_type, nc=nc, seq_type=seq_type, portable=portable, xform=xform, unique_id=uid)


def crafted_buffers():
    rng = np.random.default_rng(20261018)
    bufs = []
    n = 3000
    # 1-3: delta+wrap with corrections far outside (-max_diff, max_diff): the clamp fires, the modular
    # shortcut is invalid -> must fall back to the exact serial recurrence.  Raw, Tagged and uncompressed sources.
    for scheme, nb in (("raw", None), ("tagged", None), ("uncompressed", 2)):
        corr = rng.integers(-40, 41, size=n * 3)
        corr[rng.integers(0, n * 3, size=40)] = rng.integers(-3000, 3000, size=40)
        port = W.portable_int(corr, 3, 0, 1, scheme, W.wrap_data(-100, 155), num_bytes=nb)
        bufs.append(W.point_cloud(n, [_attr(0, 9, 3, 2, port, W.quant_params([0.5, -2.0, 10.0], 3.75, 8))]))
    # 4: regular corrections, Tagged + uncompressed + Raw attributes in ONE decoder (walk resumes twice on the device)
    c1 = rng.integers(-20, 21, size=n * 3)
    c2 = rng.integers(-5, 6, size=n * 2)
    c3 = rng.integers(-3, 4, size=n * 4)
    c4 = rng.integers(-50, 51, size=n * 1)
    attrs = [
        _attr(0, 9, 3, 2, W.portable_int(c1, 3, 0, 1, "tagged", W.wrap_data(0, 1023)), W.quant_params([0, 0, 0], 1.0, 10), 0),
        _attr(3, 9, 2, 2, W.portable_int(c2, 2, 0, 1, "tagged", W.wrap_data(0, 255)), W.quant_params([0.25, 0.5], 0.5, 8), 1),
        _attr(2, 2, 4, 1, W.portable_int(c3, 4, 0, 1, "uncompressed", W.wrap_data(0, 255), num_bytes=1), b"", 2),
        _attr(4, 4, 1, 1, W.portable_int(c4, 1, 0, 1, "raw", W.wrap_data(0, 65535)), b"", 3),
    ]
    bufs.append(W.point_cloud(n, attrs))
    # 5: the same attributes split over three attribute decoders, plus a generic (raw bytes) attribute
    gen = rng.integers(0, 256, size=n * 6, dtype=np.uint8).tobytes()
    attrs5 = attrs + [_attr(4, 3, 3, 0, gen, b"", 4)]
    bufs.append(W.point_cloud(n, attrs5, decoders=[[0, 4], [1, 2], [3]]))
    # 6: no prediction scheme (pred_method -2) and prediction without a usable transform: values = zig-zag(symbols)
    v = rng.integers(-1000, 1000, size=n * 2)
    bufs.append(W.point_cloud(n, [
        _attr(4, 5, 2, 1, W.portable_int(v, 2, -2, -1, "raw"), b"", 0),
        _attr(4, 3, 2, 1, W.portable_int(v, 2, 0, 0, "tagged"), b"", 1),      # transform 0 (delta): no scheme object
        _attr(4, 6, 2, 1, W.portable_int(v, 2, -2, -1, "uncompressed", num_bytes=4), b"", 2),
    ]))
    # 7..: every max_bit_length 1..18 (precision 12..20) with a small alphabet
    for mbl in range(1, 19):
        k = rng.integers(-3, 4, size=600)
        bufs.append(W.point_cloud(200, [_attr(0, 9, 3, 2, W.portable_int(k, 3, 0, 1, "raw", W.wrap_data(0, 4095), mbl=mbl),
                                              W.quant_params([0, 0, 0], 8.0, 12))]))
    # degenerate: one symbol with probability 2^12 (payload never consumed), as attribute 2 of house_04.obj.drc
    z = np.zeros(900, dtype=np.int64)
    bufs.append(W.point_cloud(300, [_attr(4, 2, 3, 1, W.portable_int(z, 3, 0, 1, "raw", W.wrap_data(0, 0)))]))
    bufs.append(W.point_cloud(300, [_attr(4, 2, 3, 1, W.portable_int(z, 3, 0, 1, "tagged", W.wrap_data(0, 0)))]))
    # wide alphabet: symbols up to 2^17 (dense/compact, u32 tables), 32-bit tagged fields
    big = rng.integers(-60000, 60000, size=2000 * 1)
    bufs.append(W.point_cloud(2000, [_attr(4, 5, 1, 1, W.portable_int(big, 1, 0, 1, "raw", W.wrap_data(-(1 << 20), 1 << 20)))]))
    huge = rng.integers(-(1 << 30), 1 << 30, size=500 * 2)
    bufs.append(W.point_cloud(500, [_attr(4, 5, 2, 1, W.portable_int(huge, 2, 0, 1, "tagged", W.wrap_data(-(1 << 30), (1 << 30) - 2)))]))
    mid = rng.integers(-(1 << 29), 1 << 29, size=500 * 2)
    bufs.append(W.point_cloud(500, [_attr(4, 5, 2, 1, W.portable_int(mid, 2, 0, 1, "tagged", W.wrap_data(-(1 << 29), (1 << 29) - 1)))]))
    # wrap range wider than 2^30: modular shortcut must not be used (int32 wrap-around semantics)
    bufs.append(W.point_cloud(500, [_attr(4, 5, 2, 1, W.portable_int(huge, 2, 0, 1, "tagged", W.wrap_data(-(1 << 31) + 5, (1 << 31) - 9)))]))
    # invalid wrap bounds, invalid quantization bits, bad scheme byte -> status codes
    bufs.append(W.point_cloud(10, [_attr(0, 9, 3, 2, W.portable_int(np.zeros(30), 3, 0, 1, "raw", W.wrap_data(5, 4)), W.quant_params([0, 0, 0], 1, 8))]))
    bufs.append(W.point_cloud(10, [_attr(0, 9, 3, 2, W.portable_int(np.zeros(30), 3, 0, 1, "raw", W.wrap_data(0, 4)), W.quant_params([0, 0, 0], 1, 31))]))
    bad = bytearray(W.point_cloud(10, [_attr(0, 9, 3, 2, W.portable_int(np.zeros(30), 3, 0, 1, "raw", W.wrap_data(0, 4)), W.quant_params([0, 0, 0], 1, 8))]))
    bad[11 + 1 + 1 + 5 + 1 + 3] = 7  # scheme byte
    bufs.append(bytes(bad))
    # integer attributes with MORE than 4 components (the reference loops over any nc,
    # SequentialIntegerAttributeDecoder.cs:144-152): Raw / Tagged / uncompressed sources, every store width, with and
    # without a prediction scheme, clamping corrections, 255 components
    m = 700
    for nc, dt, scheme, nb, lo, hi, pm in ((5, 2, "raw", None, 0, 255, 0), (8, 4, "tagged", None, 0, 65535, 0),
                                           (7, 5, "uncompressed", 2, -2000, 2000, 0), (6, 6, "raw", None, 0, 0, -2),
                                           (16, 1, "raw", None, -128, 127, 0), (255, 3, "tagged", None, -500, 500, 0)):
        k = rng.integers(-9, 10, size=m * nc)
        if nc == 16:
            k[rng.integers(0, m * nc, size=60)] = rng.integers(-700, 700, size=60)  # the clamp fires
        if pm == -2:
            port = W.portable_int(np.abs(k) * 37, nc, -2, -1, scheme)
        else:
            port = W.portable_int(k, nc, 0, 1, scheme, W.wrap_data(lo, hi), num_bytes=nb)
        bufs.append(W.point_cloud(m, [_attr(4, dt, nc, 1, port),
                                      _attr(0, 9, 3, 2, W.portable_int(rng.integers(-4, 5, size=m * 3), 3, 0, 1, "raw",
                                                                       W.wrap_data(0, 1023)), W.quant_params([0, 0, 0], 1.0, 10), 1)]))
    return bufs


def test_crafted_streams(gpu_decoder):
    bufs = crafted_buffers()
    gpu = gpu_decode_all(gpu_decoder, bufs)
    n_ok = compare_with_oracle(gpu, bufs)
    assert n_ok == len(bufs) - 4  # four of the buffers are invalid on purpose
    assert len(bufs) >= 39
