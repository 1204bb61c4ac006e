"""GPU: the binning stage's own sort / scan primitives (csrc/sort.cu) against torch's stable sort and cumsum -- the
specification cub::DeviceRadixSort::SortPairs / DeviceScan::InclusiveSum fulfil in the reference
(rasterizer_impl.cu:426,452-457): bit-exact, stable, every key width / pass count, ragged sizes, skewed digits."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _sort(keys, vals, bits, key_bytes):
    from ibgs_b200 import _native as N
    n = keys.numel()
    kdt = torch.int16 if key_bytes == 2 else torch.int32
    k_in = keys.to(kdt).contiguous()
    k_out = torch.empty_like(k_in)
    v_out = torch.empty(n, dtype=torch.int32, device="cuda")
    tb = int(N.lib.ibgs_sort_temp_bytes(n, bits, key_bytes))
    temp = torch.empty(max(tb, 1), dtype=torch.uint8, device="cuda")
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    N.check(N.lib.ibgs_sort_pairs(k_in.data_ptr(), vals.data_ptr() if vals is not None else None, k_out.data_ptr(),
                                  v_out.data_ptr(), n, bits, key_bytes, temp.data_ptr(), tb, stream), "ibgs_sort_pairs")
    return k_out, v_out


@pytest.mark.parametrize("n", [1, 31, 2048, 2049, 100_003, 3_000_017])
@pytest.mark.parametrize("bits,key_bytes", [(1, 2), (5, 2), (8, 2), (13, 2), (16, 2), (19, 4), (30, 4), (32, 4)])
def test_sort_pairs_is_the_stable_ascending_sort(n, bits, key_bytes):
    g = torch.Generator().manual_seed(n * 131 + bits)
    hi = 1 << bits
    keys = torch.randint(0, hi, (n,), generator=g, dtype=torch.int64)
    if n > 1000:   # skew: most keys in a few values (long runs of equal digits), a block of identical keys
        keys[: n // 3] = keys[: n // 3] % 7
        keys[n // 2: n // 2 + 5000] = hi - 1
    vals = torch.randint(0, 2 ** 31 - 1, (n,), generator=g, dtype=torch.int64).to(torch.int32)
    keys_d = keys.cuda()
    # reinterpret as the signed storage type of the same width
    store = keys_d.to(torch.int16) if key_bytes == 2 else keys_d.to(torch.int32)
    k_out, v_out = _sort(store, vals.cuda(), bits, key_bytes)
    order = torch.sort(keys_d, stable=True).indices
    want_k = keys_d[order]
    mask = 0xFFFF if key_bytes == 2 else 0xFFFFFFFF
    assert torch.equal(k_out.long() & mask, want_k)
    assert torch.equal(v_out, vals.cuda()[order])
    # index sort (values = positions): what the depth-order sort uses
    _, v_idx = _sort(store, None, bits, key_bytes)
    assert torch.equal(v_idx.long(), order)


def test_sort_only_looks_at_the_low_bits():
    g = torch.Generator().manual_seed(3)
    keys = torch.randint(0, 2 ** 31 - 1, (50_000,), generator=g, dtype=torch.int64)
    k_out, v_out = _sort(keys.cuda().to(torch.int32), None, 11, 4)
    order = torch.sort(keys.cuda() & 0x7FF, stable=True).indices
    assert torch.equal(v_out.long(), order)


@pytest.mark.parametrize("n", [1, 255, 4096, 4097, 1_000_003])
def test_gathered_inclusive_scan(n):
    from ibgs_b200 import _native as N
    g = torch.Generator().manual_seed(n)
    src = torch.randint(0, 40, (n,), generator=g, dtype=torch.int64).to(torch.int32).cuda()
    idx = torch.randperm(n, generator=g).to(torch.int32).cuda()
    out = torch.empty(n, dtype=torch.int32, device="cuda")
    tb = int(N.lib.ibgs_scan_temp_bytes(n))
    temp = torch.empty(tb, dtype=torch.uint8, device="cuda")
    N.check(N.lib.ibgs_scan_gather(n, idx.data_ptr(), src.data_ptr(), out.data_ptr(), temp.data_ptr(), tb,
                                   C.c_void_p(torch.cuda.current_stream().cuda_stream)), "ibgs_scan_gather")
    assert torch.equal(out.long(), torch.cumsum(src[idx.long()].long(), 0))


def test_sort_rejects_bad_arguments():
    from ibgs_b200 import _native as N
    assert N.lib.ibgs_sort_pairs(None, None, None, None, 10, 8, 3, None, 0, None) < 0 and "key_bytes" in N.last_error()
    assert N.lib.ibgs_sort_pairs(None, None, None, None, 10, 40, 4, None, 0, None) < 0 and "key_bits" in N.last_error()
    assert N.lib.ibgs_sort_pairs(None, None, None, None, 0, 8, 4, None, 0, None) == 0
