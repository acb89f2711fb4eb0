// Linear-space global aligners on the GPU (HirschbergSA, MyersMillerSA) -- see seqa_linspace_host.inl.
#pragma once
#include "seqa_common.cuh"
#include <vector>

struct LsState {
    int dummy = 0;
};
