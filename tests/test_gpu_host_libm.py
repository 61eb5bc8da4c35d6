"""The device restatements of the host libm (include/snch_lbvh/core/host_libm.cuh: acosf, sinf, cosf, logf as glibc computes
them) against the libm of THIS host, bit for bit: a strided sweep of each function's domain (every 61st float; sinf / cosf:
|x| < 120, beyond which the device keeps CUDA's own function — the cone merge rotates by angles in [0, pi]), the
neighbourhood of every branch point, and the special values.  The cone refit and the 2-D sampling weights are bit-identical
to the reference's CPU build because these are."""
import ctypes as C

import numpy as np
import pytest

from oracle import host_libm

pytestmark = pytest.mark.gpu


def _device(pkg, which, x):
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(x)
    st = pkg.lib().snch_selftest_host_libm(which, x.ctypes.data_as(C.c_void_p), C.c_uint64(x.size), out.ctypes.data_as(C.c_void_p), 0)
    assert st == 0, pkg.lib().snch_last_error()
    return out


def _around(values, width=64):
    b = np.asarray(values, np.float32).view(np.uint32).astype(np.int64)
    return (b[:, None] + np.arange(-width, width + 1)[None, :]).ravel().clip(0, 0x7F7FFFFF).astype(np.uint32).view(np.float32)


def _sweep(lo_bits, hi_bits, stride=61):
    pos = np.arange(lo_bits, hi_bits, stride, dtype=np.uint32)
    return np.concatenate([pos, pos | np.uint32(0x80000000)]).view(np.float32)


@pytest.mark.parametrize("which, name, x", [
    (0, "acosf", lambda: np.concatenate([_sweep(0, 0x3F800001), _around([0.5, 1.0, 2.0 ** -26, 0.0]), -_around([0.5, 1.0, 2.0 ** -26]),
                                         np.array([1.0, -1.0, 0.0, -0.0, 1.5, -1.5, np.nan], np.float32)])),
    (1, "sinf", lambda: np.concatenate([_sweep(0, 0x42F00000), _around([np.pi / 4, 2.0 ** -12, np.pi / 2, np.pi, 2 * np.pi, 119.99]),
                                        np.array([0.0, -0.0, np.inf, np.nan], np.float32)])),
    (2, "cosf", lambda: np.concatenate([_sweep(0, 0x42F00000), _around([np.pi / 4, 2.0 ** -12, np.pi / 2, np.pi, 2 * np.pi, 119.99]),
                                        np.array([0.0, -0.0, np.inf, np.nan], np.float32)])),
    (3, "logf", lambda: np.concatenate([_sweep(0x00800000, 0x7F800000)[: (0x7F800000 - 0x00800000) // 61 + 1], _around([1.0, 0.7, 1.4, 1e-2, 1e-4]),
                                        np.array([0.0, 1.0, np.inf, -1.0, np.nan, 1e-40], np.float32)])),
])
def test_device_libm_equals_host_libm(pkg, which, name, x):
    x = x()
    dev, host = _device(pkg, which, x), host_libm(which, x)
    nan = np.isnan(host)
    assert np.array_equal(np.isnan(dev), nan), f"{name}: NaN sets differ"
    bad = np.nonzero((dev.view(np.uint32) != host.view(np.uint32)) & ~nan)[0]
    # sinf / cosf: glibc picks its FMA build at run time; on a CPU without FMA 12 / 22 of the 2.2e9 arguments round differently
    assert len(bad) == 0, f"{name}: {len(bad)} of {len(x)} differ, e.g. x={x[bad[:5]]} device={dev[bad[:5]]} host={host[bad[:5]]}"
