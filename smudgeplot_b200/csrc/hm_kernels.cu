/*******************************************************************************************
 * hm_kernels.cu -- CUDA kernels (sm_100a) + layer-A entry points of include/hetmers_b200.h.
 *
 * One-substitution neighbour search over a sorted, device-resident k-mer table.
 * Integer / memory-bound work: no tensor cores (see DESIGN.md §5 for the roofline).
 *
 * What replaces what (reference file:line under /root/reference/src/lib):
 *   unpack_records_kernel   Next_Kmer_Entry + Current_Entry   libfastk.c:1159-1176,:1230-1269
 *   bucket_index_kernel     stub index + GoTo_Kmer_Entry       libfastk.c:1320-1409
 *   pass1_degree_kernel     analysis_in_core_1 / _thread_1     PloidyPlot.c:454-568,:168-301
 *   pass2_plot_kernel       analysis_in_core_2 / _thread_2     PloidyPlot.c:570-700,:303-452
 *   min_count_kernel        examine_table (trim half)          PloidyPlot.c:1171-1197
 *   find_keys_kernel        GoTo_Kmer_Entry exact-hit use      PloidyPlot.c:1213
 *******************************************************************************************/
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "hetmers_b200.h"
#include "hm_internal.h"

/* ------------------------------------------------------------------ error plumbing ------ */

static thread_local char g_err[1024] = "";

extern "C" const char *hm_last_error(void) { return g_err; }
extern "C" int hm_abi_version(void) { return 1; }

extern "C" int hm_set_error(int code, const char *fmt, ...)
{ va_list ap;
  va_start(ap,fmt);
  vsnprintf(g_err,sizeof(g_err),fmt,ap);
  va_end(ap);
  return code;
}

int hm_cuda_fail(cudaError_t e, const char *what)
{ return hm_set_error(HM_ECUDA,"%s: %s",what,cudaGetErrorString(e)); }

extern "C" int hm_device_count(void)
{ int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess)
    { cudaGetLastError(); return 0; }
  return n;
}

extern "C" int hm_device_info(int dev, char *name, int name_len, int *sm_count, int64_t *total_mem)
{ cudaDeviceProp p;
  cudaError_t e = cudaGetDeviceProperties(&p,dev);
  if (e != cudaSuccess)
    return hm_cuda_fail(e,"cudaGetDeviceProperties");
  if (name != NULL && name_len > 0)
    { strncpy(name,p.name,name_len-1); name[name_len-1] = 0; }
  if (sm_count != NULL) *sm_count = p.multiProcessorCount;
  if (total_mem != NULL) *total_mem = (int64_t) p.totalGlobalMem;
  return HM_OK;
}

extern "C" int hm_pick_bucket_bits(int64_t n)
{ int lg = 0;
  while (lg < 62 && ((int64_t) 1 << (lg+1)) <= n)
    lg += 1;                                  /* floor(log2 n) */
  int bits = lg-1;                            /* ~2-4 entries per bucket */
  if (bits > 28) bits = 28;
  if (bits < 2)  bits = 2;
  return bits;
}

/* ------------------------------------------------------------------ device helpers ------ */

template <typename IdxT> struct IdxNone { static constexpr IdxT value = (IdxT) ~(IdxT) 0; };

/* exact match of y inside its prefix bucket; -1 if absent */
template <typename IdxT>
__device__ __forceinline__ int64_t bucket_find(const uint64_t *__restrict__ keys,
                                               const IdxT *__restrict__ bucket,
                                               int bshift, uint64_t y)
{ uint64_t bk = y >> bshift;
  IdxT l = bucket[bk];
  IdxT r = bucket[bk+1];
  while (l < r)
    { IdxT     m = l + ((r-l)>>1);
      uint64_t v = __ldg(keys+m);
      if (v == y)
        return (int64_t) m;
      if (v < y) l = m+1; else r = m;
    }
  return -1;
}

/* ------------------------------------------------------------------------- unpack ------- */

__device__ __forceinline__ int upper_bound_index(const int64_t *__restrict__ index,
                                                 int l, int r, int64_t o)
{ /* first b in [l,r] with index[b] > o  (r is a valid answer bound) */
  while (l < r)
    { int m = (l+r)>>1;
      if (__ldg(index+m) > o) r = m; else l = m+1;
    }
  return l;
}

__global__ void __launch_bounds__(256)
unpack_records_kernel(const uint8_t *__restrict__ rec, int64_t n, int64_t first,
                      const int64_t *__restrict__ index, int ixlen, int ibyte, int hbyte,
                      uint64_t *__restrict__ keys, uint16_t *__restrict__ cnt)
{ __shared__ int s_blo, s_bhi;
  int64_t t0 = (int64_t) blockIdx.x * blockDim.x;
  if (threadIdx.x == 0)
    { int64_t o0 = first+t0;
      int64_t o1 = first + (t0+blockDim.x < n ? t0+blockDim.x : n) - 1;
      s_blo = upper_bound_index(index,0,ixlen-1,o0);
      s_bhi = upper_bound_index(index,s_blo,ixlen-1,o1);
    }
  __syncthreads();
  int64_t i = t0+threadIdx.x;
  if (i >= n)
    return;
  int pbyte = hbyte+2;
  uint64_t b = (uint64_t) upper_bound_index(index,s_blo,s_bhi,first+i);
  const uint8_t *r = rec + i*pbyte;
  uint64_t v = 0;
  for (int j = 0; j < hbyte; j++)
    v = (v<<8) | r[j];
  uint64_t key = b << (64-8*ibyte);
  if (hbyte > 0)
    key |= v << (64-8*(ibyte+hbyte));
  keys[i] = key;
  cnt[i]  = (uint16_t) (r[hbyte] | (r[hbyte+1]<<8));
}

extern "C" int hm_k_unpack_records(const uint8_t *d_rec, int64_t n, int64_t first,
                                   const int64_t *d_stub_index, int ibyte, int kmer,
                                   uint64_t *d_keys, uint16_t *d_cnt, void *stream)
{ if (kmer < 1 || kmer > HM_MAX_KMER)
    return hm_set_error(HM_EUNSUPPORTED,"k-mer length %d not supported (1..%d)",kmer,HM_MAX_KMER);
  int kbyte = (kmer+3)>>2;
  if (ibyte < 1 || ibyte > 3 || ibyte > kbyte)
    return hm_set_error(HM_EFORMAT,"prefix bytes ibyte=%d invalid for k=%d",ibyte,kmer);
  if (n <= 0)
    return HM_OK;
  int64_t nblk = (n+255)/256;
  unpack_records_kernel<<<(unsigned) nblk,256,0,(cudaStream_t) stream>>>
      (d_rec,n,first,d_stub_index,1<<(8*ibyte),ibyte,kbyte-ibyte,d_keys,d_cnt);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return hm_cuda_fail(e,"unpack_records_kernel");
  return HM_OK;
}

/* ------------------------------------------------------------------- bucket index ------- */

/* bucket[b] = first i with (keys[i] >> bshift) >= b, for b in [0, 2^bits]; thread i fills the
 * (normally 0 or 1) buckets that start at entry i.                                           */
template <typename IdxT>
__global__ void __launch_bounds__(256)
bucket_index_kernel(const uint64_t *__restrict__ keys, int64_t n, int bshift, int64_t nbuckets,
                    IdxT *__restrict__ bucket)
{ int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n)
    return;
  int64_t cur  = (i < n) ? (int64_t) (keys[i] >> bshift) : nbuckets;
  int64_t prev = (i > 0) ? (int64_t) (keys[i-1] >> bshift) : -1;
  for (int64_t b = prev+1; b <= cur; b++)
    bucket[b] = (IdxT) i;
}

extern "C" int hm_k_build_bucket_index(const uint64_t *d_keys, int64_t n, int bits,
                                       void *d_bucket, int idx64, void *stream)
{ if (bits < 1 || bits > 30)
    return hm_set_error(HM_EINVAL,"bucket bits %d out of range 1..30",bits);
  if (!idx64 && n >= 0xFFFFFFFFll)
    return hm_set_error(HM_EINVAL,"32-bit offsets need n < 2^32-1");
  int64_t nblk = (n+1+255)/256;
  if (idx64)
    bucket_index_kernel<uint64_t><<<(unsigned) nblk,256,0,(cudaStream_t) stream>>>
        (d_keys,n,64-bits,(int64_t) 1<<bits,(uint64_t *) d_bucket);
  else
    bucket_index_kernel<uint32_t><<<(unsigned) nblk,256,0,(cudaStream_t) stream>>>
        (d_keys,n,64-bits,(int64_t) 1<<bits,(uint32_t *) d_bucket);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return hm_cuda_fail(e,"bucket_index_kernel");
  return HM_OK;
}

/* -------------------------------------------------------------------------- pass 1 ------ */

/* One thread per table entry x = keys[i].  A neighbour y > x differing at base p shares x's
 * first p bases, and every entry between x and y in sorted order shares them too, so
 * p <= lcp(x, successor(x)): positions beyond that need no probe at all.  For the remaining
 * positions each larger base is tried by a prefix-bucket lookup + in-bucket bisection.  The
 * lower member of a pair does all the book-keeping: deg[x]+=1, deg[y]+=1 (byte-packed atomics),
 * up[x] = y.                                                                                 */
template <typename IdxT>
__global__ void __launch_bounds__(256)
pass1_degree_kernel(const uint64_t *__restrict__ keys, const uint16_t *__restrict__ cnt,
                    int64_t n, const IdxT *__restrict__ bucket, int bshift, int kmer,
                    int64_t lo, int64_t hi, uint32_t *__restrict__ deg32, IdxT *__restrict__ up)
{ int64_t i = lo + (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hi)
    return;
  uint64_t x = keys[i];
  int      pmax = -1;
  if (i+1 < n)
    { pmax = __clzll((long long) (x ^ keys[i+1])) >> 1;
      if (pmax > kmer-1) pmax = kmer-1;
    }
  int      cx   = cnt[i];
  unsigned udeg = 0;
  IdxT     upj  = IdxNone<IdxT>::value;

#pragma unroll 1
  for (int p = 0; p <= pmax; p++)
    { int sh = 62-2*p;
      int b  = (int) ((x >> sh) & 3);
      for (int d = 1; d <= 3-b; d++)
        { uint64_t y = x + ((uint64_t) d << sh);
          int64_t  j = bucket_find<IdxT>(keys,bucket,bshift,y);
          if (j >= 0 && cx + (int) __ldg(cnt+j) <= HM_SMAX)
            { udeg += 1;
              upj   = (IdxT) j;
              atomicAdd(deg32 + (j>>2), 1u << (8*(j&3)));
            }
        }
    }
  if (udeg != 0)
    atomicAdd(deg32 + (i>>2), udeg << (8*(i&3)));
  up[i-lo] = upj;
}

extern "C" int hm_k_pass1_degree(const uint64_t *d_keys, const uint16_t *d_cnt, int64_t n,
                                 const void *d_bucket, int bits, int idx64, int kmer,
                                 int64_t lo, int64_t hi, uint8_t *d_deg, void *d_up, void *stream)
{ if (kmer < 1 || kmer > HM_MAX_KMER)
    return hm_set_error(HM_EUNSUPPORTED,"k-mer length %d not supported (1..%d)",kmer,HM_MAX_KMER);
  if (lo < 0 || hi > n || lo > hi || bits < 1 || bits > 30)
    return hm_set_error(HM_EINVAL,"pass1: bad range [%lld,%lld) of %lld or bits %d",
                        (long long) lo,(long long) hi,(long long) n,bits);
  if (hi == lo)
    return HM_OK;
  int64_t nblk = (hi-lo+255)/256;
  if (idx64)
    pass1_degree_kernel<uint64_t><<<(unsigned) nblk,256,0,(cudaStream_t) stream>>>
        (d_keys,d_cnt,n,(const uint64_t *) d_bucket,64-bits,kmer,lo,hi,(uint32_t *) d_deg,(uint64_t *) d_up);
  else
    pass1_degree_kernel<uint32_t><<<(unsigned) nblk,256,0,(cudaStream_t) stream>>>
        (d_keys,d_cnt,n,(const uint32_t *) d_bucket,64-bits,kmer,lo,hi,(uint32_t *) d_deg,(uint32_t *) d_up);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return hm_cuda_fail(e,"pass1_degree_kernel");
  return HM_OK;
}

/* -------------------------------------------------------------------------- pass 2 ------ */

#define P2_TS 256      /* shared-memory tile: sums  < 256 */
#define P2_TM 128      /*                     mins  < 128 */
#define P2_THREADS 1024

/* deg[x] <= 1 and deg[y] <= 1 for a recorded qualifying pair means both are exactly 1, i.e. the
 * pair is isolated: one count in plot[cx+cy][min].  Persistent CTAs keep the dense corner of the
 * plot in shared memory (uint32) and flush once; the rest goes to 64-bit global atomics.      */
template <typename IdxT>
__global__ void __launch_bounds__(P2_THREADS,1)
pass2_plot_kernel(const uint16_t *__restrict__ cnt, const uint8_t *__restrict__ deg,
                  const IdxT *__restrict__ up, int64_t lo, int64_t hi,
                  unsigned long long *__restrict__ plot)
{ extern __shared__ uint32_t tile[];
  for (int t = threadIdx.x; t < P2_TS*P2_TM; t += blockDim.x)
    tile[t] = 0;
  __syncthreads();
  int64_t stride = (int64_t) gridDim.x * blockDim.x;
  for (int64_t i = lo + (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride)
    { if (deg[i] > 1)
        continue;
      IdxT j = up[i-lo];
      if (j == IdxNone<IdxT>::value || __ldg(deg+j) > 1)
        continue;
      int ci = cnt[i], cj = __ldg(cnt+j);
      int s  = ci+cj;
      int m  = ci < cj ? ci : cj;
      if (s < P2_TS && m < P2_TM)
        atomicAdd(tile + s*P2_TM + m, 1u);
      else
        atomicAdd(plot + s*HM_PLOT_W + m, 1ull);
    }
  __syncthreads();
  for (int t = threadIdx.x; t < P2_TS*P2_TM; t += blockDim.x)
    { uint32_t v = tile[t];
      if (v != 0)
        atomicAdd(plot + (t/P2_TM)*HM_PLOT_W + (t%P2_TM), (unsigned long long) v);
    }
}

extern "C" int hm_k_pass2_plot(const uint16_t *d_cnt, const uint8_t *d_deg, const void *d_up,
                               int idx64, int64_t lo, int64_t hi, unsigned long long *d_plot,
                               void *stream)
{ static int configured[64] = {0};
  if (lo > hi)
    return hm_set_error(HM_EINVAL,"pass2: bad range");
  if (hi == lo)
    return HM_OK;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms,cudaDevAttrMultiProcessorCount,dev);
  size_t smem = (size_t) P2_TS*P2_TM*sizeof(uint32_t);
  if (dev < 64 && !configured[dev])
    { cudaError_t e1 = cudaFuncSetAttribute(pass2_plot_kernel<uint32_t>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize,(int) smem);
      cudaError_t e2 = cudaFuncSetAttribute(pass2_plot_kernel<uint64_t>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize,(int) smem);
      if (e1 != cudaSuccess || e2 != cudaSuccess)
        return hm_cuda_fail(e1 != cudaSuccess ? e1 : e2,"cudaFuncSetAttribute(pass2)");
      configured[dev] = 1;
    }
  int64_t want = (hi-lo+P2_THREADS-1)/P2_THREADS;
  int     grid = (int) (want < sms ? want : sms);
  if (idx64)
    pass2_plot_kernel<uint64_t><<<grid,P2_THREADS,smem,(cudaStream_t) stream>>>
        (d_cnt,d_deg,(const uint64_t *) d_up,lo,hi,d_plot);
  else
    pass2_plot_kernel<uint32_t><<<grid,P2_THREADS,smem,(cudaStream_t) stream>>>
        (d_cnt,d_deg,(const uint32_t *) d_up,lo,hi,d_plot);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return hm_cuda_fail(e,"pass2_plot_kernel");
  return HM_OK;
}

/* ------------------------------------------------------------------------- examine ------ */

__global__ void __launch_bounds__(256)
min_count_kernel(const uint16_t *__restrict__ cnt, int64_t frst, int64_t last, int *out)
{ int     best = 0x8000;
  int64_t stride = (int64_t) gridDim.x * blockDim.x;
  for (int64_t i = frst + (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < last; i += stride)
    { int v = (int) (int16_t) cnt[i];          /* the reference reads counts as int16 (:1189) */
      if (v >= 1 && v < best) best = v;
    }
  for (int o = 16; o > 0; o >>= 1)
    { int w = __shfl_xor_sync(0xffffffffu,best,o);
      if (w < best) best = w;
    }
  if ((threadIdx.x & 31) == 0 && best < 0x8000)
    atomicMin(out,best);
}

extern "C" int hm_k_min_count(const uint16_t *d_cnt, int64_t frst, int64_t last, int *d_min,
                              void *stream)
{ if (last <= frst)
    return HM_OK;
  int64_t want = (last-frst+255)/256;
  int     grid = (int) (want < 148*8 ? want : 148*8);
  min_count_kernel<<<grid,256,0,(cudaStream_t) stream>>>(d_cnt,frst,last,d_min);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return hm_cuda_fail(e,"min_count_kernel");
  return HM_OK;
}

template <typename IdxT>
__global__ void __launch_bounds__(128)
find_keys_kernel(const uint64_t *__restrict__ keys, const IdxT *__restrict__ bucket, int bshift,
                 const uint64_t *__restrict__ query, int64_t nq, int64_t *__restrict__ pos)
{ int64_t q = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (q < nq)
    pos[q] = bucket_find<IdxT>(keys,bucket,bshift,query[q]);
}

extern "C" int hm_k_find_keys(const uint64_t *d_keys, int64_t n, const void *d_bucket, int bits,
                              int idx64, const uint64_t *d_query, int64_t nq, int64_t *d_pos,
                              void *stream)
{ (void) n;
  if (nq <= 0)
    return HM_OK;
  int64_t nblk = (nq+127)/128;
  if (idx64)
    find_keys_kernel<uint64_t><<<(unsigned) nblk,128,0,(cudaStream_t) stream>>>
        (d_keys,(const uint64_t *) d_bucket,64-bits,d_query,nq,d_pos);
  else
    find_keys_kernel<uint32_t><<<(unsigned) nblk,128,0,(cudaStream_t) stream>>>
        (d_keys,(const uint32_t *) d_bucket,64-bits,d_query,nq,d_pos);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return hm_cuda_fail(e,"find_keys_kernel");
  return HM_OK;
}
