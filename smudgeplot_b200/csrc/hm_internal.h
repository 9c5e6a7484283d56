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
/* chunk-wise index construction for the loader (hm_kernels.cu) */
int hm_build_bucket_index_range(const uint64_t *d_keys, int64_t n, int bits, void *d_bucket,
                                int idx64, int64_t i0, int64_t i1, void *stream);
int hm_build_filter_range(const uint64_t *d_keys, int filter_bits, uint32_t *d_filter,
                          int64_t i0, int64_t i1, void *stream);

#include <cuda_runtime.h>
int hm_cuda_fail(cudaError_t e, const char *what);
#define HM_CUDA(call)                                              \
  do { cudaError_t _e = (call);                                    \
       if (_e != cudaSuccess) return hm_cuda_fail(_e,#call);       \
     } while (0)
#endif

#endif
