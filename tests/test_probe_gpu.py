"""GPU parity of the HBM probe kernels against the oracle (numpy + C restatements of the spec),
called through the C ABI (ctypes -> libgpushare_b200.so). Bit-exact: bytes, checksums, counts."""
import ctypes as C

import numpy as np
import pytest

from oracle import probe_oracle as po

pytestmark = pytest.mark.gpu

from gpushare_device_plugin_b200 import _abi  # noqa: E402

VARIANTS = [_abi.GSB_VARIANT_DIRECT, _abi.GSB_VARIANT_CPASYNC, _abi.GSB_VARIANT_BULK, _abi.GSB_VARIANT_BULKW, _abi.GSB_VARIANT_BULKD]
IDS = ["direct", "cpasync", "bulk", "bulkw", "bulkd"]
# (offset_bytes, n_bytes): tile-aligned, ragged head, ragged tail, tiny, one word, > one wave of tiles
WINDOWS = [
    (0, 1 << 20),
    (16 * 12345, 16 * 100003),
    (16 * 7, 16 * 1023),
    (16 * 1024 * 37, 16),
    (1 << 20, 48 << 20),
    (16 * 3, 16 * (4096 * 151 + 5)),
]


def read_words(gsb, offset, nbytes):
    return np.frombuffer(gsb.arena_read(0, offset, nbytes), dtype=np.uint32).reshape(-1, 4)


@pytest.mark.parametrize("variant", VARIANTS, ids=IDS)
@pytest.mark.parametrize("offset,nbytes", WINDOWS)
def test_fill_bytes_and_checksums_match_oracle(gsb, small_arena, variant, offset, nbytes):
    # background generation, so that "outside the window is untouched" is checkable
    gsb.probe(0, _abi.GSB_OP_FILL, variant=_abi.GSB_VARIANT_DIRECT, seed_write=11)
    r = gsb.probe(0, _abi.GSB_OP_FILL, variant=variant, offset=offset, nbytes=nbytes, seed_write=0xC0FFEE)
    assert r.variant in (variant, _abi.GSB_VARIANT_DIRECT)
    fw, nw = offset // 16, nbytes // 16
    want = po.pattern(fw, nw, 0xC0FFEE)
    got = read_words(gsb, offset, nbytes)
    assert np.array_equal(got, want)
    assert (r.checksum_xor, r.checksum_sum) == po.checksums(want)
    assert (r.bytes_walked, r.bytes_read, r.bytes_written) == (nbytes, 0, nbytes)
    assert r.mismatch_words == 0 and r.first_bad_offset == _abi.UINT64_MAX
    # 64 words either side still hold the background generation
    if offset >= 1024:
        assert np.array_equal(read_words(gsb, offset - 1024, 1024), po.pattern(fw - 64, 64, 11))
    assert np.array_equal(read_words(gsb, offset + nbytes, 1024), po.pattern(fw + nw, 64, 11))


@pytest.mark.parametrize("variant", VARIANTS, ids=IDS)
@pytest.mark.parametrize("offset,nbytes", WINDOWS)
def test_verify_refill_roundtrip(gsb, small_arena, variant, offset, nbytes):
    fw, nw = offset // 16, nbytes // 16
    gsb.probe(0, _abi.GSB_OP_FILL, variant=_abi.GSB_VARIANT_DIRECT, offset=offset, nbytes=nbytes, seed_write=5)
    want5 = po.pattern(fw, nw, 5)
    v = gsb.probe(0, _abi.GSB_OP_VERIFY, variant=variant, offset=offset, nbytes=nbytes, seed_expect=5)
    assert v.mismatch_words == 0 and v.mismatch_bits == 0
    assert (v.checksum_xor, v.checksum_sum) == po.checksums(want5)
    assert (v.bytes_read, v.bytes_written) == (nbytes, 0)
    r = gsb.probe(0, _abi.GSB_OP_VERIFY_REFILL, variant=variant, offset=offset, nbytes=nbytes, seed_expect=5,
                  seed_write=6)
    assert r.mismatch_words == 0
    assert (r.checksum_xor, r.checksum_sum) == po.checksums(want5)  # checksum is of what was READ
    assert (r.bytes_read, r.bytes_written) == (nbytes, nbytes)
    assert np.array_equal(read_words(gsb, offset, nbytes), po.pattern(fw, nw, 6))
    # the previous generation is now wrong everywhere: every word mismatches
    w = gsb.probe(0, _abi.GSB_OP_VERIFY, variant=variant, offset=offset, nbytes=nbytes, seed_expect=5)
    assert w.mismatch_words == nw and w.first_bad_offset == offset


@pytest.mark.parametrize("variant", VARIANTS, ids=IDS)
def test_injected_faults_are_counted_exactly(gsb, small_arena, variant):
    offset, nbytes = 16 * 999, 16 * 300007
    fw, nw = offset // 16, nbytes // 16
    gsb.probe(0, _abi.GSB_OP_FILL, variant=variant, offset=offset, nbytes=nbytes, seed_write=21)
    rng = np.random.default_rng(1234)
    words = np.sort(rng.choice(nw, size=37, replace=False))
    flipped = 0
    for i, w in enumerate(words):
        mask = np.zeros(4, dtype=np.uint32)
        if i % 3 == 0:
            mask[rng.integers(4)] = np.uint32(1) << np.uint32(rng.integers(32))  # single bit
        else:
            mask[:] = rng.integers(1, 1 << 32, size=4, dtype=np.uint64).astype(np.uint32)
        cur = read_words(gsb, offset + 16 * int(w), 16)[0]
        gsb.arena_write(0, offset + 16 * int(w), (cur ^ mask).tobytes())
        flipped += sum(bin(int(m)).count("1") for m in mask)
    observed = read_words(gsb, offset, nbytes)
    want = po.verify(observed, fw, 21)
    assert want["mismatch_words"] == 37 and want["mismatch_bits"] == flipped
    r = gsb.probe(0, _abi.GSB_OP_VERIFY_REFILL, variant=variant, offset=offset, nbytes=nbytes, seed_expect=21,
                  seed_write=22)
    assert r.mismatch_words == 37
    assert r.mismatch_bits == flipped
    assert r.first_bad_offset == offset + 16 * int(words[0]) == want["first_bad_offset"]
    assert (r.checksum_xor, r.checksum_sum) == (want["checksum_xor"], want["checksum_sum"])
    # the refill repaired the window
    again = gsb.probe(0, _abi.GSB_OP_VERIFY, variant=variant, offset=offset, nbytes=nbytes, seed_expect=22)
    assert again.mismatch_words == 0


def test_variants_agree_on_random_windows(gsb, small_arena):
    rng = np.random.default_rng(7)
    for _ in range(6):
        off = 16 * int(rng.integers(0, 1 << 20))
        nb = 16 * int(rng.integers(1, 1 << 21))
        seeds = [int(s) for s in rng.integers(0, 1 << 32, size=2)]
        outs = []
        for variant in VARIANTS:
            gsb.probe(0, _abi.GSB_OP_FILL, variant=variant, offset=off, nbytes=nb, seed_write=seeds[0])
            r = gsb.probe(0, _abi.GSB_OP_VERIFY_REFILL, variant=variant, offset=off, nbytes=nb,
                          seed_expect=seeds[0], seed_write=seeds[1])
            outs.append((r.mismatch_words, r.checksum_xor, r.checksum_sum, gsb.arena_read(0, off, min(nb, 1 << 16))))
        assert outs[0] == outs[1] == outs[2] == outs[3] == outs[4]
        assert outs[0][0] == 0


def test_empty_window_and_argument_errors(gsb, small_arena):
    r = gsb.probe(0, _abi.GSB_OP_VERIFY, offset=small_arena, nbytes=0, seed_expect=1)
    assert r.bytes_walked == 0 and r.mismatch_words == 0 and (r.checksum_xor, r.checksum_sum) == (0, 0)
    for kw in (dict(offset=8, nbytes=16), dict(offset=0, nbytes=24), dict(offset=small_arena, nbytes=16),
               dict(offset=0, nbytes=small_arena + 16)):
        bad = gsb.probe(0, _abi.GSB_OP_VERIFY, raise_on_error=False, **kw)
        assert bad.status == _abi.GSB_ERR_INVALID_ARGUMENT
    bad = gsb.probe(0, 9, raise_on_error=False)
    assert bad.status == _abi.GSB_ERR_INVALID_ARGUMENT
    bad = gsb.probe(99, _abi.GSB_OP_VERIFY, raise_on_error=False)
    assert bad.status == _abi.GSB_ERR_NO_DEVICE


def test_grid_override_gives_same_answer(gsb, small_arena):
    gsb.probe(0, _abi.GSB_OP_FILL, seed_write=3)
    ref = gsb.probe(0, _abi.GSB_OP_VERIFY, seed_expect=3)
    for grid in (1, 3, 148, 149, 1000):
        r = gsb.probe(0, _abi.GSB_OP_VERIFY, seed_expect=3, grid=grid)
        assert (r.checksum_xor, r.checksum_sum, r.mismatch_words) == (ref.checksum_xor, ref.checksum_sum, 0)
