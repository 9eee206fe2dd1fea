/* TEST INFRASTRUCTURE ONLY — not part of the product path.
 *
 * Builds oracle/_ref/libref.so = the UNMODIFIED reference native code, compiled from the
 * sources where they lie under /root/reference (nothing is copied into this repo):
 *   footprint_tools/modeling/predict.h    (fast_predict, free_result_t)
 *   footprint_tools/modeling/smoothing.h  (quickselect, trimmed_mean, windowed_trimmed_mean)
 *   footprint_tools/stats/windowing.h     (fast_sum/product/fishers_combined/stouffers_z,
 *                                          fast_windowing_func, fast_weighted_*)
 *   hcephes/src/...                       (compiled as separate objects by the Makefile)
 * The reference headers define their functions non-static, so including them in one
 * translation unit exports them from the shared object as-is. smoothing.h uses
 * malloc/calloc/memcpy without including their headers, hence the includes below.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include "footprint_tools/modeling/predict.h"
#include "footprint_tools/stats/windowing.h"
