"""The product's restatement of glibc's acosf / sinf / cosf / logf (include/snch_lbvh/core/host_libm.cuh — ONE source for the
device kernels and for this host build) against the libm of this host, bit for bit, without a GPU: a strided sweep of each
function's domain (every 97th float) plus the neighbourhood of every branch point and the special values.  The exhaustive
comparison (every float of each domain, zero differences against glibc 2.39) is recorded in the header; the device build of
the same source is checked on the GPU box by tests/test_gpu_host_libm.py."""
import numpy as np
import pytest

from oracle import host_libm, restated_libm


def _around(values, width=64):
    b = np.asarray(values, np.float32).view(np.uint32).astype(np.int64)
    return (b[:, None] + np.arange(-width, width + 1)[None, :]).ravel().clip(0, 0x7F7FFFFF).astype(np.uint32).view(np.float32)


def _sweep(lo_bits, hi_bits, stride=97):
    pos = np.arange(lo_bits, hi_bits, stride, dtype=np.uint32)
    return np.concatenate([pos, pos | np.uint32(0x80000000)]).view(np.float32)


@pytest.mark.parametrize("which, name, x", [
    (0, "acosf", lambda: np.concatenate([_sweep(0, 0x3F800001), _around([0.5, 1.0, 2.0 ** -26, 0.0]), -_around([0.5, 1.0, 2.0 ** -26]),
                                         np.array([1.0, -1.0, 0.0, -0.0, 1.5, -1.5, np.nan], np.float32)])),
    (1, "sinf", lambda: np.concatenate([_sweep(0, 0x42F00000), _around([np.pi / 4, 2.0 ** -12, np.pi / 2, np.pi, 2 * np.pi, 119.99]),
                                        np.array([0.0, -0.0, 120.0, 1e6, np.inf, np.nan], np.float32)])),
    (2, "cosf", lambda: np.concatenate([_sweep(0, 0x42F00000), _around([np.pi / 4, 2.0 ** -12, np.pi / 2, np.pi, 2 * np.pi, 119.99]),
                                        np.array([0.0, -0.0, 120.0, 1e6, np.inf, np.nan], np.float32)])),
    (3, "logf", lambda: np.concatenate([np.arange(0x00800000, 0x7F800000, 97, dtype=np.uint32).view(np.float32), _around([1.0, 0.7, 1.4, 1e-2, 1e-4]),
                                        np.array([0.0, 1.0, np.inf, -1.0, np.nan, 1e-40], np.float32)])),
])
def test_restated_libm_equals_host_libm(which, name, x):
    x = x()
    mine, host = restated_libm(which, x), host_libm(which, x)
    nan = np.isnan(host)
    assert np.array_equal(np.isnan(mine), nan), f"{name}: NaN sets differ"
    bad = np.nonzero((mine.view(np.uint32) != host.view(np.uint32)) & ~nan)[0]
    # sinf / cosf follow glibc's FMA build (selected at run time on every x86-64 CPU since Haswell); on a CPU without FMA 12 / 22 of
    # the 2.2e9 arguments of the domain round differently
    assert len(bad) == 0, f"{name}: {len(bad)} of {len(x)} differ, e.g. x={x[bad[:5]]} restated={mine[bad[:5]]} host={host[bad[:5]]}"
