/* TEST INFRASTRUCTURE ONLY (oracle/): stand-in for <nvector/nvector_openmp.h>
 * used when the reference is compiled with -D_OPENMP_ON (src/Model/Macros.hpp:11-14,
 * src/Model/f.cpp:8-9).  Same storage as the serial shim. */
#ifndef SHUD_B200_NVECTOR_OPENMP_SHIM_H
#define SHUD_B200_NVECTOR_OPENMP_SHIM_H
#include "nvector_serial.h"
#define NV_DATA_OMP(v) NV_DATA_S(v)
#define NV_Ith_OMP(v, i) NV_Ith_S(v, i)
#endif
