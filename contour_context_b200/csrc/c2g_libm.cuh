// c2g_libm.cuh — libm calls whose result bits feed descriptors.  The reference calls glibc (std::exp in gaussPDF,
// include/tools/algos.h:53-56; std::atan2(float,float) in the BCI build, include/cont2/contour_mng.h:860; std::acos in
// checkConstellCorrespSim, :1191-1192).  CUDA's libdevice versions are within 1-2 ulp of those; until the glibc
// algorithms are ported bit-for-bit, every such call goes through this header so that the parity tests can bound and
// count the differences in one place (tests/test_ingest_gpu.py).
#pragma once
__device__ __forceinline__ float c2g_atan2f(float y, float x) { return atan2f(y, x); }
__device__ __forceinline__ float c2g_acosf(float x) { return acosf(x); }
__device__ __forceinline__ double c2g_exp(double x) { return exp(x); }
