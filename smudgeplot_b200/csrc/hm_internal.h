/* hm_internal.h -- shared by the translation units of libhetmers_b200.so (not installed) */
#ifndef HM_INTERNAL_H
#define HM_INTERNAL_H

#include <stdarg.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* record a message for hm_last_error() and return `code` */
int hm_set_error(int code, const char *fmt, ...);

#ifdef __cplusplus
}
#endif

#ifdef __CUDACC__
#include <cuda_runtime.h>
int hm_cuda_fail(cudaError_t e, const char *what);
#define HM_CUDA(call)                                              \
  do { cudaError_t _e = (call);                                    \
       if (_e != cudaSuccess) return hm_cuda_fail(_e,#call);       \
     } while (0)
#endif

#endif
