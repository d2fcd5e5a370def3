"""Invariants of the 2-byte device layout of a posting row (usb_hostindex.h HostHalf, built by
make_half): U must not change, so every target appears exactly once, with its byte class given by the
entry position; padding only hits the dummy words; and the point of the order -- the 32 increments
one warp instruction issues fall into (nearly) distinct shared-memory banks."""
import ctypes as C

import numpy as np
import pytest

from usearch12_b200 import capi


def half_row(targets, n_targets):
    t = np.ascontiguousarray(targets, dtype=np.uint32)
    out = np.zeros(256 * (len(t) // 64 + 8), dtype=np.uint16)
    groups, dummy0 = C.c_uint32(), C.c_uint32()
    capi.check(capi.lib().usb_debug_half_row(t.ctypes.data, len(t), n_targets, out.ctypes.data, out.size, C.byref(groups),
                                             C.byref(dummy0)))
    return out[:groups.value * 256].reshape(groups.value, 32, 8), dummy0.value  # [group][lane][entry]


@pytest.mark.parametrize("n_targets,n", [(100000, 2146), (100000, 0), (100000, 1), (403, 403), (65537, 9000), (200000, 31)])
def test_half_row_holds_every_target_once(n_targets, n):
    rng = np.random.default_rng(n_targets + n)
    targets = np.sort(rng.choice(n_targets, size=n, replace=False)).astype(np.uint32)
    rows, dummy0 = half_row(targets, n_targets)
    assert dummy0 == ((n_targets + 15) // 16 * 16) // 4
    got = []
    for g in range(rows.shape[0]):
        for lane in range(32):
            for i in range(8):
                w = int(rows[g, lane, i])
                if w >= dummy0:
                    assert w == dummy0 + lane  # padding: the dummy word of this lane's bank
                else:
                    got.append(4 * w + i // 2)  # byte class implied by the entry position
    assert sorted(got) == list(targets)
    if n:
        per_class = np.bincount(targets & 3, minlength=4)
        assert rows.shape[0] == (per_class.max() + 63) // 64  # groups: the fullest byte class decides


def test_half_row_spreads_an_instruction_over_the_banks():
    rng = np.random.default_rng(7)
    worst, total, instr = 0, 0, 0
    for _ in range(20):
        targets = np.sort(rng.choice(100000, size=2146, replace=False)).astype(np.uint32)
        rows, dummy0 = half_row(targets, 100000)
        for g in range(rows.shape[0]):
            for i in range(8):  # one ATOMS instruction: entry i of every lane
                banks = rows[g, :, i].astype(np.int64) % 32
                deg = np.bincount(banks, minlength=32).max()
                worst = max(worst, deg)
                total += deg
                instr += 1
    # ascending rows walked as they are give about 3.4 wavefronts per instruction (32 random banks);
    # this order gives 1.8 here and ncu measured 1.7 on the B200 (profiles/README.md).  The rest comes
    # from the last rounds of a byte class, when the rarer banks have run out.
    assert total / instr < 2.0, total / instr
    assert worst <= 16  # the last instruction rows of a class collect what is left of the fullest banks


def test_half_row_rejects_bad_input():
    out = np.zeros(256, dtype=np.uint16)
    g, d = C.c_uint32(), C.c_uint32()
    t = np.array([5, 3], dtype=np.uint32)
    assert capi.lib().usb_debug_half_row(t.ctypes.data, 2, 100, out.ctypes.data, out.size, C.byref(g), C.byref(d)) == -1
    assert capi.lib().usb_debug_half_row(t.ctypes.data, 1, 300000, out.ctypes.data, out.size, C.byref(g), C.byref(d)) == -1
