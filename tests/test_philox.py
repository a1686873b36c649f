"""Philox4x32-10: Random123 known-answer vectors on the host mirror and on the device."""
import numpy as np
import pytest

from oracle import philox

KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


@pytest.mark.parametrize('ctr,key,expect', KAT)
def test_kat_host(ctr, key, expect):
    out = philox.philox4x32_10(*ctr, *key)
    assert tuple(int(x) for x in out) == expect


def test_uniform_grid():
    u = philox.word_to_uniform(np.array([0, 1 << 9, 0xffffffff], dtype=np.uint32))
    assert u[0] == 0.0 and u[1] == np.float32(2.0 ** -23) and u[2] == np.float32(1 - 2.0 ** -23)
    w = philox.indicator_words(seed=12345, chain_id=3, iteration=7, n=10)
    q = philox.philox4x32_10(np.arange(3), 7, 3, 0, 12345, 0)
    assert np.array_equal(w, np.stack(q, 1).reshape(-1)[:10])


@pytest.mark.gpu
@pytest.mark.parametrize('ctr,key,expect', KAT)
def test_kat_device(ctr, key, expect):
    import ctypes as C
    import torch
    from basicrta_b200 import _cabi
    lib = _cabi.load()
    out = torch.zeros(4, dtype=torch.int32, device='cuda')
    seed = key[0] | (key[1] << 32)
    rc = lib.brta_philox_fill(C.c_void_p(out.data_ptr()), 1, ctr[0], ctr[1], ctr[2], ctr[3], seed, None)
    assert rc == 0
    torch.cuda.synchronize()
    got = tuple(int(x) & 0xffffffff for x in out.cpu().numpy())
    assert got == expect


@pytest.mark.gpu
def test_device_stream_matches_host_mirror():
    import ctypes as C
    import torch
    from basicrta_b200 import _cabi
    lib = _cabi.load()
    n = 1000
    out = torch.zeros(4 * n, dtype=torch.int32, device='cuda')
    assert lib.brta_philox_fill(C.c_void_p(out.data_ptr()), n, 0, 17, 5, 0, 0xdeadbeefcafe, None) == 0
    torch.cuda.synchronize()
    got = out.cpu().numpy().view(np.uint32)
    assert np.array_equal(got, philox.indicator_words(0xdeadbeefcafe, 5, 17, 4 * n))
