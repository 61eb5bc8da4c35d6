/* TEST INFRASTRUCTURE.  The HOST build of the product's restatement of glibc's acosf / sinf / cosf / logf
 * (include/snch_lbvh/core/host_libm.cuh: `*_glibc`, the very source the kernels compile for the device), exported so that
 * tests/test_host_libm_cpu.py can compare it with this host's libm without a GPU.  Built with -ffp-contract=off (the header's
 * host arithmetic is plain operators). */
#include "../include/snch_lbvh/core/host_libm.cuh"

extern "C" void orc_restated_libm(int which, const float *x, long n, float *out)
{
    for (long i = 0; i < n; ++i)
    {
        const float v = x[i];
        out[i] = which == 0   ? lbvh::detail::acosf_glibc(v)
                 : which == 1 ? lbvh::detail::sinf_glibc(v)
                 : which == 2 ? lbvh::detail::cosf_glibc(v)
                              : lbvh::detail::logf_glibc(v);
    }
}
