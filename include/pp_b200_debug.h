/* pp_b200_debug.h -- development entry points of libpp_b200.so.  NOT part of the product ABI: they exist only in a
 * library built with -DPP_DEBUG (python 3d-object-detection_b200/build.py --debug) and are used by scripts/*.py
 * (role timing, stage ablation of the tcgen05 kernels).  A reference-side binding never needs them. */
#ifndef PP_B200_DEBUG_H
#define PP_B200_DEBUG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* keys: "pfn_tc_timing" (1: CTA 0 of the tensor-core PFN kernels records per-role wait cycles),
 *       "pfn_tc_debug"  (bit mask: stages of the tensor-core kernels to skip, results are then invalid),
 *       "pad_reserve_sms" (SMs the persistent padding pass leaves free). */
int pp_debug_set(const char* key, int value);
/* 128 int64: [warp][wait0, wait1, busy / issue, role total] of the last instrumented launch. */
int pp_debug_tc_timing(int64_t* out64);

#ifdef __cplusplus
}
#endif
#endif
