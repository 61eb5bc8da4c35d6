/* TEST INFRASTRUCTURE ONLY — pre-include used when the reference headers under
 * /root/reference/include are compiled as plain C++ (Thrust CPP backend) by
 * oracle/Makefile.  It supplies the few CUDA built-ins the reference's
 * construct() uses on the device (bvh.cuh:531,548) and the unqualified
 * max/min it relies on (scene.cuh:429-430,609,1123).  Nothing under the
 * product path includes this file. */
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <iostream>
#include <stdexcept>
#include <unordered_map>
#include <vector>
using std::max;
using std::min;
/* serial stand-ins: the CPP backend runs for_each sequentially */
inline int atomicCAS(int *a, int c, int v)
{
    int o = *a;
    if (o == c) *a = v;
    return o;
}
inline void __threadfence() {}
