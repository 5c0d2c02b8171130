"""CPU suite: the fp64 rescoring behind hulo_resect_acransac (host code of the library, no device
involved) -- radix sort with repair + bracketed NFA scan -- agrees bit for bit with a comparison
sort and the full log10 scan (hulo_selftest_rescoring) on seeded sets with ties, duplicates and
non-finite residuals, at the sizes either side of every switch in the code (N < 64 comparison
sort, the register-sort limit, large sets)."""
import ctypes as C

import pytest

from sfmlocalization_b200 import _lib


@pytest.mark.parametrize("n_points,n_cases", [(3, 2), (4, 50), (5, 50), (63, 200), (64, 200), (65, 200), (300, 300),
                                              (699, 300), (2048, 100), (4097, 40), (20000, 6)])
def test_fast_rescoring_equals_plain(n_points, n_cases):
    lib = _lib.load()
    n = C.c_size_t(12345)
    rc = lib.hulo_selftest_rescoring(1000 + n_points, n_points, n_cases, C.byref(n))
    assert rc == _lib.OK
    assert n.value == 0


def test_rescoring_selftest_rejects_bad_arguments():
    lib = _lib.load()
    assert lib.hulo_selftest_rescoring(1, 100, 1, None) == _lib.ERR_ARG
    n = C.c_size_t(0)
    assert lib.hulo_selftest_rescoring(1, 40000, 1, C.byref(n)) == _lib.ERR_ARG
