/*******************************************************************************************
 * hm_condition.cu -- table conditioning on the GPU: the two things the reference delegates to
 * external FastK executables before it scans (PloidyPlot.c:1381-1426):
 *
 *   trim        `Logex -T<t> '<tmp>.trim=A[<L>-]' <table>`   keep entries with count >= L
 *   symmetrise  `Symmex -T<t> -P<dir> <table> <tmp>.symx`     add the reverse complement of every
 *                                                             k-mer (same count), keep the table sorted
 *
 * FastK's tools are not part of the reference tree and are not pinned to a version (SURVEY.md
 * §8c), so this restates their documented effect, not their code: parity for THIS step is pinned
 * only against a numpy restatement in tests/ ("parity unpinned" against the real tools).  The
 * executable uses it by default and falls back to the reference's shell-outs with
 * HETMERS_EXTERNAL_CONDITIONING=1.
 *
 * Not a hot path (it runs once, before the scan): selection and sorting use CUB's device-wide
 * primitives (library code); the reverse-complement / duplicate-flag kernels are ours.
 * Duplicates (palindromes for even k, or an input that already held both strands) keep the
 * ORIGINAL entry: the concatenation puts originals first and the radix sort is stable.
 *******************************************************************************************/
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <stdint.h>

#include "hetmers_b200.h"
#include "hm_internal.h"

__device__ __forceinline__ uint64_t rev2_64(uint64_t x)       /* reverse the 32 2-bit fields */
{ x = ((x >> 2)  & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
  x = ((x >> 4)  & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
  x = ((x >> 8)  & 0x00FF00FF00FF00FFull) | ((x & 0x00FF00FF00FF00FFull) << 8);
  x = ((x >> 16) & 0x0000FFFF0000FFFFull) | ((x & 0x0000FFFF0000FFFFull) << 16);
  return (x >> 32) | (x << 32);
}

__global__ void __launch_bounds__(256)
trim_flag_kernel(const uint16_t *__restrict__ cnt, int64_t n, int ethresh, uint8_t *__restrict__ flag)
{ int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    flag[i] = (cnt[i] >= ethresh);
}

/* out[0,n) = table, out[n,2n) = reverse complements with the same counts */
__global__ void __launch_bounds__(256)
append_revcomp_kernel(const uint64_t *__restrict__ hi, const uint64_t *__restrict__ lo,
                      const uint16_t *__restrict__ cnt, int64_t n, int kmer,
                      uint64_t *__restrict__ ohi, uint64_t *__restrict__ olo, uint16_t *__restrict__ ocnt)
{ int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  uint64_t x = hi[i];
  uint16_t c = cnt[i];
  ohi[i] = x; ocnt[i] = c; ocnt[n+i] = c;
  if (lo == NULL)
    { uint64_t r = rev2_64(~x);
      if (kmer < 32)
        r = (r & (((uint64_t) 1 << (2*kmer))-1)) << (64-2*kmer);
      ohi[n+i] = r;
    }
  else
    { uint64_t w = lo[i];
      uint64_t a = rev2_64(~w), b = rev2_64(~x);          /* the two words swap */
      int      sh = 2*(64-kmer);
      olo[i] = w;
      if (sh == 0) { ohi[n+i] = a; olo[n+i] = b; }
      else         { ohi[n+i] = (a << sh) | (b >> (64-sh)); olo[n+i] = b << sh; }
    }
}

__global__ void __launch_bounds__(256)
iota_kernel(uint32_t *__restrict__ idx, int64_t n)
{ int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) idx[i] = (uint32_t) i;
}

template <typename T>
__global__ void __launch_bounds__(256)
gather_kernel(const T *__restrict__ src, const uint32_t *__restrict__ idx, int64_t n, T *__restrict__ dst)
{ int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}

__global__ void __launch_bounds__(256)
first_of_run_kernel(const uint64_t *__restrict__ hi, const uint64_t *__restrict__ lo, int64_t n,
                    uint8_t *__restrict__ flag)
{ int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    flag[i] = (i == 0) || hi[i] != hi[i-1] || (lo != NULL && lo[i] != lo[i-1]);
}

#define GRID(n) ((unsigned) (((n)+255)/256))

template <typename T>
static int select_flagged(const T *in, const uint8_t *flag, T *out, int64_t n, int64_t *d_nsel,
                          void **tmp, size_t *tmp_bytes, cudaStream_t st)
{ size_t need = 0;
  HM_CUDA(cub::DeviceSelect::Flagged(NULL,need,in,flag,out,d_nsel,n,st));
  if (need > *tmp_bytes)
    { if (*tmp) cudaFree(*tmp);
      HM_CUDA(cudaMalloc(tmp,need));
      *tmp_bytes = need;
    }
  HM_CUDA(cub::DeviceSelect::Flagged(*tmp,need,in,flag,out,d_nsel,n,st));
  return HM_OK;
}

template <typename K, typename V>
static int sort_pairs(const K *kin, K *kout, const V *vin, V *vout, int64_t n, int begin_bit, int end_bit,
                      void **tmp, size_t *tmp_bytes, cudaStream_t st)
{ size_t need = 0;
  HM_CUDA(cub::DeviceRadixSort::SortPairs(NULL,need,kin,kout,vin,vout,n,begin_bit,end_bit,st));
  if (need > *tmp_bytes)
    { if (*tmp) cudaFree(*tmp);
      HM_CUDA(cudaMalloc(tmp,need));
      *tmp_bytes = need;
    }
  HM_CUDA(cub::DeviceRadixSort::SortPairs(*tmp,need,kin,kout,vin,vout,n,begin_bit,end_bit,st));
  return HM_OK;
}

/* Replace (*pk, *pl, *pc, *pn) by the conditioned table (new cudaMalloc'ed arrays with one spare
 * element; the old ones are freed).  *pl is NULL for k <= 32.                                  */
int hm_condition_arrays(int kmer, int ethresh, int do_trim, int do_symm,
                        uint64_t **pk, uint64_t **pl, uint16_t **pc, int64_t *pn, cudaStream_t st)
{ int64_t   n = *pn;
  int       two = (*pl != NULL);
  void     *tmp = NULL;
  size_t    tmp_bytes = 0;
  uint8_t  *flag = NULL;
  int64_t  *d_nsel = NULL, nsel = 0;
  int       rc = HM_OK;

  HM_CUDA(cudaMalloc(&d_nsel,sizeof(int64_t)));

  if (do_trim && n > 0)
    { uint64_t *k2 = NULL, *l2 = NULL; uint16_t *c2 = NULL;
      HM_CUDA(cudaMalloc(&flag,(size_t) n));
      HM_CUDA(cudaMalloc(&k2,sizeof(uint64_t)*(size_t) (n+1)));
      HM_CUDA(cudaMalloc(&c2,sizeof(uint16_t)*(size_t) (n+1)));
      if (two) HM_CUDA(cudaMalloc(&l2,sizeof(uint64_t)*(size_t) (n+1)));
      trim_flag_kernel<<<GRID(n),256,0,st>>>(*pc,n,ethresh,flag);
      if ((rc = select_flagged(*pk,flag,k2,n,d_nsel,&tmp,&tmp_bytes,st)) != HM_OK) return rc;
      if (two && (rc = select_flagged(*pl,flag,l2,n,d_nsel,&tmp,&tmp_bytes,st)) != HM_OK) return rc;
      if ((rc = select_flagged(*pc,flag,c2,n,d_nsel,&tmp,&tmp_bytes,st)) != HM_OK) return rc;
      HM_CUDA(cudaMemcpyAsync(&nsel,d_nsel,sizeof(int64_t),cudaMemcpyDeviceToHost,st));
      HM_CUDA(cudaStreamSynchronize(st));
      cudaFree(*pk); cudaFree(*pc); if (two) cudaFree(*pl);
      cudaFree(flag); flag = NULL;
      *pk = k2; *pc = c2; *pl = l2; n = nsel;
    }

  if (do_symm && n > 0)
    { int64_t   m = 2*n;
      uint64_t *h0 = NULL, *l0 = NULL, *h1 = NULL, *l1 = NULL;
      uint16_t *c0 = NULL, *c1 = NULL;
      if (m >= 0xFFFFFFF0ll)
        return hm_set_error(HM_EUNSUPPORTED,"symmetrising %lld entries needs 64-bit sort indices",(long long) n);
      HM_CUDA(cudaMalloc(&h0,sizeof(uint64_t)*(size_t) (m+1)));
      HM_CUDA(cudaMalloc(&h1,sizeof(uint64_t)*(size_t) (m+1)));
      HM_CUDA(cudaMalloc(&c0,sizeof(uint16_t)*(size_t) (m+1)));
      HM_CUDA(cudaMalloc(&c1,sizeof(uint16_t)*(size_t) (m+1)));
      if (two)
        { HM_CUDA(cudaMalloc(&l0,sizeof(uint64_t)*(size_t) (m+1)));
          HM_CUDA(cudaMalloc(&l1,sizeof(uint64_t)*(size_t) (m+1)));
        }
      append_revcomp_kernel<<<GRID(n),256,0,st>>>(*pk,*pl,*pc,n,kmer,h0,l0,c0);
      cudaFree(*pk); cudaFree(*pc); if (two) cudaFree(*pl);
      *pk = NULL; *pc = NULL; *pl = NULL;
      if (!two)
        { int bb = kmer < 32 ? 64-2*kmer : 0;
          if ((rc = sort_pairs(h0,h1,c0,c1,m,bb,64,&tmp,&tmp_bytes,st)) != HM_OK) return rc;
        }
      else
        { uint32_t *i0 = NULL, *i1 = NULL;
          HM_CUDA(cudaMalloc(&i0,sizeof(uint32_t)*(size_t) m));
          HM_CUDA(cudaMalloc(&i1,sizeof(uint32_t)*(size_t) m));
          iota_kernel<<<GRID(m),256,0,st>>>(i0,m);
          int bb = kmer < 64 ? 128-2*kmer : 0;
          /* least significant word first, then a stable sort on the most significant word */
          if ((rc = sort_pairs(l0,l1,i0,i1,m,bb,64,&tmp,&tmp_bytes,st)) != HM_OK) return rc;
          gather_kernel<uint64_t><<<GRID(m),256,0,st>>>(h0,i1,m,h1);          /* hi in lo-order   */
          if ((rc = sort_pairs(h1,l1,i1,i0,m,0,64,&tmp,&tmp_bytes,st)) != HM_OK) return rc;
          /* l1 = sorted hi, i0 = final permutation */
          gather_kernel<uint64_t><<<GRID(m),256,0,st>>>(l0,i0,m,h1);          /* h1 := lo sorted  */
          gather_kernel<uint16_t><<<GRID(m),256,0,st>>>(c0,i0,m,c1);
          /* arrange as (h1 = hi, l1 = lo) */
          uint64_t *t = h1; h1 = l1; l1 = t;
          cudaFree(i0); cudaFree(i1);
        }
      /* unique (first of every run of equal keys wins) back into h0/l0/c0 */
      HM_CUDA(cudaMalloc(&flag,(size_t) m));
      first_of_run_kernel<<<GRID(m),256,0,st>>>(h1,two ? l1 : NULL,m,flag);
      if ((rc = select_flagged(h1,flag,h0,m,d_nsel,&tmp,&tmp_bytes,st)) != HM_OK) return rc;
      if (two && (rc = select_flagged(l1,flag,l0,m,d_nsel,&tmp,&tmp_bytes,st)) != HM_OK) return rc;
      if ((rc = select_flagged(c1,flag,c0,m,d_nsel,&tmp,&tmp_bytes,st)) != HM_OK) return rc;
      HM_CUDA(cudaMemcpyAsync(&nsel,d_nsel,sizeof(int64_t),cudaMemcpyDeviceToHost,st));
      HM_CUDA(cudaStreamSynchronize(st));
      cudaFree(h1); cudaFree(c1); if (two) cudaFree(l1);
      cudaFree(flag); flag = NULL;
      *pk = h0; *pc = c0; *pl = l0; n = nsel;
    }

  cudaError_t e = cudaStreamSynchronize(st);
  if (tmp) cudaFree(tmp);
  cudaFree(d_nsel);
  if (e != cudaSuccess)
    return hm_cuda_fail(e,"conditioning");
  *pn = n;
  return HM_OK;
}
