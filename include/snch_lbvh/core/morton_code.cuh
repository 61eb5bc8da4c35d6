// snch_lbvh/core/morton_code.cuh — Morton (Z-order) keys of the drop-in C++ API.
//
// Same names and values as the reference's core/morton_code.cuh: 10 bits per axis for the 32-bit code (x is the most
// significant axis), 21 bits per axis for the 64-bit one, a point is quantised by clamp(x * resolution, 0, resolution - 1)
// and truncation.  The bit spreading is written as a generic "insert two zero bits between consecutive bits" routine.
#ifndef SNCH_LBVH_B200_MORTON_CODE_CUH
#define SNCH_LBVH_B200_MORTON_CODE_CUH
#include "utility.cuh"
#include <cassert>

namespace lbvh
{
namespace detail
{
// b9..b0 -> b9 0 0 b8 0 0 ... b0 (magic-mask doubling: each step halves the run length of packed bits)
SNCH_LBVH_CALLABLE std::uint32_t spread3_u32(std::uint32_t v) noexcept
{
    v &= 0x3FFu;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
SNCH_LBVH_CALLABLE std::uint64_t spread3_u64(std::uint64_t v) noexcept
{
    v &= 0x3FFFFFull; // the reference accepts 22 significant bits (morton_code.cuh:42)
    v = (v | (v << 32)) & 0x003F00000000FFFFull;
    v = (v | (v << 16)) & 0x003F0000FF0000FFull;
    v = (v | (v << 8)) & 0x300F00F00F00F00Full;
    v = (v | (v << 4)) & 0x30C30C30C30C30C3ull;
    v = (v | (v << 2)) & 0x9249249249249249ull;
    return v;
}
template <typename T> SNCH_LBVH_CALLABLE T quantise(T x, T resolution) noexcept { return ::fmin(::fmax(x * resolution, T(0)), resolution - T(1)); }
} // namespace detail

// spread the low 10 (32-bit) / 22 (64-bit) bits of v so that two zero bits separate consecutive bits      morton_code.cuh:19-58
SNCH_LBVH_CALLABLE std::uint32_t expand_bits(std::uint32_t v) noexcept
{
    assert(v < (1u << 10));
    return detail::spread3_u32(v);
}
SNCH_LBVH_CALLABLE std::uint64_t expand_bits(std::uint64_t v) noexcept
{
    assert(v < (std::uint64_t(1) << 22));
    return detail::spread3_u64(v);
}

// point in the unit cube / square -> 30-bit code (3-D: xx*4 + yy*2 + zz) or 20 interleaved-by-3 bits (2-D: xx*2 + yy)  :61-101
template <typename V, detail::enable_real_vec<V> = 0>
SNCH_LBVH_CALLABLE std::uint32_t morton_code(V p, detail::scalar_of<V> resolution = detail::scalar_of<V>(1024)) noexcept
{
    static_assert(detail::vec_traits<V>::size == 2 || detail::vec_traits<V>::size == 3, "morton_code: 2-D or 3-D points");
    std::uint32_t code = 0;
    for (int i = 0; i < detail::vec_traits<V>::size; ++i)
        code = code * 2u + expand_bits(static_cast<std::uint32_t>(detail::quantise(detail::at(p, i), resolution)));
    return code;
}
// 21 bits per axis, 63-bit (3-D) / 42-bit (2-D) code.  Unused by the builder (the reference's duplicate-key path is
// (code << 32) | object index, bvh.cuh:464-476) but part of the public surface.                           :103-142
template <typename V, detail::enable_real_vec<V> = 0>
SNCH_LBVH_CALLABLE std::uint64_t morton_code64(V p, detail::scalar_of<V> resolution = detail::scalar_of<V>(1048576)) noexcept
{
    static_assert(detail::vec_traits<V>::size == 2 || detail::vec_traits<V>::size == 3, "morton_code64: 2-D or 3-D points");
    std::uint64_t code = 0;
    for (int i = 0; i < detail::vec_traits<V>::size; ++i)
    {
        using T = detail::scalar_of<V>;
        const T q = detail::min_of(detail::max_of(detail::at(p, i) * resolution, T(0)), resolution - T(1));
        code = (code << 1) | expand_bits(static_cast<std::uint64_t>(q));
    }
    return code;
}

// number of leading bits two keys share                                                                    :144-167
SNCH_LBVH_CALLABLE int common_upper_bits(const unsigned int lhs, const unsigned int rhs) noexcept
{
#ifdef __CUDA_ARCH__
    return ::__clz(lhs ^ rhs);
#else
    return (lhs ^ rhs) ? __builtin_clz(lhs ^ rhs) : 32;
#endif
}
SNCH_LBVH_CALLABLE int common_upper_bits(const unsigned long long int lhs, const unsigned long long int rhs) noexcept
{
#ifdef __CUDA_ARCH__
    return ::__clzll(lhs ^ rhs);
#else
    return (lhs ^ rhs) ? __builtin_clzll(lhs ^ rhs) : 64;
#endif
}
} // namespace lbvh
#endif // SNCH_LBVH_B200_MORTON_CODE_CUH
