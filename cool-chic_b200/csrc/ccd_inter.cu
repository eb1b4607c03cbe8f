// ccd_inter.cu -- P/B frame motion compensation (placeholder translation unit; the
// kernels land with SURVEY 8f1).
#include "ccd_internal.h"
