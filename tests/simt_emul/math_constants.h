// TEST-ONLY stand-in for <math_constants.h> (see cuda_runtime.h in this directory).
#pragma once
#define CUDART_INF_F (__builtin_huge_valf())
#define CUDART_NAN_F (__builtin_nanf(""))
