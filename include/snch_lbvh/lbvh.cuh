// snch_lbvh/lbvh.cuh — umbrella header of the drop-in C++ API (same include path as the reference's lbvh.cuh).
#ifndef SNCH_LBVH_B200_LBVH_CUH
#define SNCH_LBVH_B200_LBVH_CUH
#include "core/utility.cuh"
#include "core/aabb.cuh"
#include "core/cone.cuh"
#include "core/morton_code.cuh"
#include "core/predicator.cuh"
#include "core/bvh.cuh"
#include "core/query.cuh"
#include "core/sample.cuh"
#endif // SNCH_LBVH_B200_LBVH_CUH
