/*******************************************************************************************
 * hm_scan.cu -- layer B of include/hetmers_b200.h: the whole hetmers path from HOST buffers.
 *
 *   hm_scan_create   H2D of the raw FastK part payloads (double-buffered, copy stream ||
 *                    unpack stream), SoA unpack, bucket index          ("T_load", device part)
 *   hm_scan_examine  trimmed? / symmetric? decisions of examine_table (PloidyPlot.c:1167-1230)
 *   hm_scan_run      pass 1 -> (degree exchange when >1 GPU) -> pass 2 -> plot D2H  ("T_scan")
 *
 * Every device holds a full replica of the table (180 GB HBM3e holds 16e9 k=31 entries); work is
 * sharded by contiguous index range [lo_g, hi_g).  With one GPU there is no exchange at all.
 * With several GPUs in this single process the loader gathers the shards over NVLink peer
 * copies and foreign degree bytes are reached through the owner's array (remote atomics / loads
 * fused into the two kernels; summed by a peer-memory kernel of hm_peer.cu if there are no
 * native NVLink atomics); the
 * one-process-per-GPU variant (torch.distributed / NCCL) lives in smudgeplot_b200/dist.py and
 * uses layer A directly.
 *******************************************************************************************/
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <pthread.h>
#include <unistd.h>

#include "hetmers_b200.h"
#include "hm_internal.h"

#define HM_MAX_GPUS 16
#define LOAD_CHUNK  (16ll<<20)        /* records per H2D/unpack chunk */

typedef struct
  { int                 dev;
    cudaStream_t        st, st_copy;
    uint64_t           *keys;
    uint64_t           *keys_lo;      /* second key word, k > 32 only */
    uint16_t           *cnt;
    uint8_t            *deg;          /* n rounded up to 4 */
    void               *bucket;
    uint32_t           *filter;       /* prefix presence bitmap */
    void               *up;           /* hi-lo entries */
    void               *p2scratch;    /* pass 2 defer list (multi-GPU peer mode) */
    int64_t             p2scratch_bytes;
    unsigned long long *plot;
    int64_t             lo, hi;       /* this device's work range */
  } DevTable;

struct hm_scan
  { int      kmer, ibyte, bits, fpos, idx64, ngpu;
    int64_t  n;
    DevTable d[HM_MAX_GPUS];
    double   ms_load, ms_alloc, ms_records, ms_index;
    int      ran, peer_mode;              /* state of the last hm_scan_run */
    hm_shards sh[HM_MAX_GPUS];
    int64_t  launches;
  };

static double now_ms(void)
{ struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC,&ts);
  return ts.tv_sec*1e3 + ts.tv_nsec*1e-6;
}

/* multi-GPU helpers (hm_peer.cu) */
int hm_peer_enable(const int *dev, int n);
int hm_peer_sum_deg(uint8_t **deg, const int64_t *lo, const int64_t *hi, const int *dev,
                    cudaStream_t *st, int n, int64_t nels);
int hm_peer_sum_plot(unsigned long long **plot, const int *dev, cudaStream_t *st, int n);
/* GPU trim / symmetrise (hm_condition.cu) */
int hm_condition_arrays(int kmer, int ethresh, int do_trim, int do_symm,
                        uint64_t **pk, uint64_t **pl, uint16_t **pc, int64_t *pn, cudaStream_t st);

/* Device allocations of a one-GPU scan come from the device's stream-ordered memory pool with a
 * release threshold of "never": a second hm_scan_create in the same process (bench e2e leg, a
 * service handling many tables) reuses the memory instead of paying cudaMalloc / cudaFree of
 * several GB every call (measured: ~20 ms of a 61 ms call).  Multi-GPU scans keep cudaMalloc:
 * their arrays are mapped by the peers.  HETMERS_NO_POOL=1 disables the pool.                  */
#define POOL_MAX 32
typedef struct { void *p[POOL_MAX]; int n, enabled; } PoolReg;
static PoolReg g_pool[64];

static void pool_setup(int dev, int enable)
{ static int configured[64] = {0};
  if (dev < 0 || dev >= 64) return;
  g_pool[dev].enabled = enable && getenv("HETMERS_NO_POOL") == NULL;
  if (g_pool[dev].enabled && !configured[dev])
    { cudaMemPool_t pool;
      unsigned long long never = ~0ull;
      if (cudaDeviceGetDefaultMemPool(&pool,dev) != cudaSuccess ||
          cudaMemPoolSetAttribute(pool,cudaMemPoolAttrReleaseThreshold,&never) != cudaSuccess)
        { cudaGetLastError(); g_pool[dev].enabled = 0; }
      configured[dev] = 1;
    }
}

static cudaError_t dalloc(int dev, cudaStream_t st, void **p, size_t bytes)
{ PoolReg *R = (dev >= 0 && dev < 64) ? g_pool+dev : NULL;
  if (R != NULL && R->enabled && R->n < POOL_MAX)
    { cudaError_t e = cudaMallocAsync(p,bytes,st);
      if (e == cudaSuccess)
        { R->p[R->n++] = *p; return e; }
      cudaGetLastError();
    }
  return cudaMalloc(p,bytes);
}

static void dfree(int dev, cudaStream_t st, void *p)
{ PoolReg *R = (dev >= 0 && dev < 64) ? g_pool+dev : NULL;
  if (p == NULL) return;
  if (R != NULL)
    for (int k = 0; k < R->n; k++)
      if (R->p[k] == p)
        { R->p[k] = R->p[--R->n];
          cudaFreeAsync(p,st);
          return;
        }
  cudaFree(p);
}

static void free_dev(DevTable *D)
{ cudaSetDevice(D->dev);
  dfree(D->dev,D->st,D->keys);  dfree(D->dev,D->st,D->keys_lo); dfree(D->dev,D->st,D->cnt);
  dfree(D->dev,D->st,D->deg);   dfree(D->dev,D->st,D->bucket);  dfree(D->dev,D->st,D->filter);
  dfree(D->dev,D->st,D->up);    dfree(D->dev,D->st,D->plot);
  if (D->p2scratch) cudaFree(D->p2scratch);
  if (D->st) cudaStreamSynchronize(D->st);
  if (D->st)      cudaStreamDestroy(D->st);
  if (D->st_copy) cudaStreamDestroy(D->st_copy);
  memset(D,0,sizeof(*D));
}

extern "C" void hm_scan_destroy(hm_scan *s)
{ if (s == NULL)
    return;
  for (int g = 0; g < s->ngpu; g++)
    free_dev(s->d+g);
  free(s);
}

/* ---- host-side staging for pageable sources (mmap'ed part files) -------------------------
 * cudaMemcpyAsync from pageable memory is staged by the driver on one thread (~3-6 GB/s).  The
 * executable's table lives in the page cache, so the loader copies each chunk into a pinned
 * buffer with a few host threads (this is what the reference's -T is for on the host side) while
 * the previous chunk is in flight to the GPU.                                                   */
static int g_io_threads = 0;

extern "C" void hm_set_io_threads(int n) { g_io_threads = n; }

typedef struct { uint8_t *dst; const uint8_t *src; size_t bytes; int fd; int64_t off; } CopyJob;

static void *copy_worker(void *arg)
{ CopyJob *j = (CopyJob *) arg;
  if (j->fd < 0)
    memcpy(j->dst,j->src,j->bytes);
  else
    { size_t got = 0;                                   /* page cache -> pinned buffer, no mapping */
      while (got < j->bytes)
        { ssize_t r = pread(j->fd,j->dst+got,j->bytes-got,j->off+(int64_t) got);
          if (r <= 0) break;
          got += (size_t) r;
        }
      if (got < j->bytes)
        memset(j->dst+got,0,j->bytes-got);
    }
  return NULL;
}

/* fill dst[0,bytes) from memory `src` (fd < 0) or from file `fd` at `off`, with the I/O threads */
static void parallel_fill(uint8_t *dst, const uint8_t *src, int fd, int64_t foff, size_t bytes)
{ int nt = g_io_threads;
  if (nt <= 0)
    { long c = sysconf(_SC_NPROCESSORS_ONLN);
      nt = c > 16 ? 16 : (c < 1 ? 1 : (int) c);
    }
  if (nt > 64) nt = 64;
  if (bytes < ((size_t) 4<<20)) nt = 1;
  pthread_t th[64];
  CopyJob   job[64];
  size_t    per = ((bytes/nt)+4095) & ~(size_t) 4095;
  int       started = 0;
  for (int k = 0; k < nt; k++)
    { size_t off = per*k;
      if (off >= bytes) break;
      job[k].dst = dst+off; job[k].src = src ? src+off : NULL;
      job[k].fd = fd; job[k].off = foff+(int64_t) off;
      job[k].bytes = bytes-off < per ? bytes-off : per;
      if (k == nt-1 || off+per >= bytes)
        { copy_worker(job+k); break; }                 /* the calling thread takes the last slice */
      if (pthread_create(th+k,NULL,copy_worker,job+k) != 0)
        { copy_worker(job+k); continue; }
      started = k+1;
    }
  for (int k = 0; k < started; k++)
    pthread_join(th[k],NULL);
}

static int is_pageable(const void *p)
{ cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a,p) != cudaSuccess)
    { cudaGetLastError(); return 1; }
  return (a.type == cudaMemoryTypeUnregistered);
}

/* Load ordinals [first, first+count) of the table onto device D (keys/cnt already allocated for
 * the full table).  Walks the parts, copies payload chunks H2D on st_copy into one of two device
 * staging buffers and unpacks them on st (copy of chunk c+1 overlaps the unpack of chunk c);
 * pageable sources additionally go through two pinned host buffers filled by host threads.      */
static int load_range(hm_scan *s, DevTable *D, const hm_host_table *t, const int64_t *d_index,
                      int64_t first, int64_t count)
{ int      kbyte = (t->kmer+3)>>2;
  int      pbyte = kbyte - t->ibyte + 2;
  uint8_t *stage[2] = { NULL, NULL };
  uint8_t *pin[2]   = { NULL, NULL };
  cudaEvent_t copied[2], unpacked[2];
  int64_t  chunk = LOAD_CHUNK;
  int      rc = HM_OK, b = 0, used[2] = {0,0};
  int      staged = 0;

  if (count <= 0)
    return HM_OK;
  for (int p = 0; p < t->nparts && !staged; p++)
    if (t->part_nels[p] > 0 &&
        ((t->part_fd != NULL && t->part_fd[p] >= 0) || is_pageable(t->part_rec[p])))
      staged = 1;
  if (staged)
    chunk = LOAD_CHUNK/2;
  if (chunk > count) chunk = count;
  for (int i = 0; i < 2; i++)
    { HM_CUDA(dalloc(D->dev,D->st,(void **) &stage[i],(size_t) chunk*pbyte));
      if (staged)
        HM_CUDA(cudaHostAlloc(&pin[i],(size_t) chunk*pbyte,cudaHostAllocDefault));
      HM_CUDA(cudaEventCreateWithFlags(&copied[i],cudaEventDisableTiming));
      HM_CUDA(cudaEventCreateWithFlags(&unpacked[i],cudaEventDisableTiming));
    }
  /* one GPU: the whole table arrives here in order, so the bucket index and the prefix filter are
   * built chunk by chunk right behind the unpack (hidden behind the next chunk's H2D)           */
  const int inc = (s->ngpu == 1 && first == 0 && count == s->n);
  if (inc)
    HM_CUDA(cudaMemsetAsync(D->filter,0,sizeof(uint32_t)*(size_t) hm_filter_words(s->fpos),D->st));
  HM_CUDA(cudaStreamSynchronize(D->st));         /* (pool) allocations are used on both streams */
  int64_t pstart = 0;                               /* ordinal of the part's first record */
  for (int p = 0; p < t->nparts && rc == HM_OK; p++)
    { int64_t pn   = t->part_nels[p];
      int64_t from = first > pstart ? first : pstart;
      int64_t to   = first+count < pstart+pn ? first+count : pstart+pn;
      for (int64_t o = from; o < to && rc == HM_OK; o += chunk)
        { int64_t        m   = to-o < chunk ? to-o : chunk;
          const uint8_t *src = t->part_rec[p] + (o-pstart)*pbyte;
          if (staged)
            { if (used[b])
                cudaEventSynchronize(copied[b]);      /* pin[b] has left for the GPU */
              if (t->part_fd != NULL && t->part_fd[p] >= 0)
                parallel_fill(pin[b],NULL,t->part_fd[p],t->part_fd_off[p]+(o-pstart)*pbyte,(size_t) m*pbyte);
              else
                parallel_fill(pin[b],src,-1,0,(size_t) m*pbyte);
              src = pin[b];
            }
          if (used[b])
            cudaStreamWaitEvent(D->st_copy,unpacked[b],0);
          cudaError_t e = cudaMemcpyAsync(stage[b],src,(size_t) m*pbyte,cudaMemcpyHostToDevice,D->st_copy);
          if (e != cudaSuccess) { rc = hm_cuda_fail(e,"cudaMemcpyAsync(H2D records)"); break; }
          cudaEventRecord(copied[b],D->st_copy);
          cudaStreamWaitEvent(D->st,copied[b],0);
          rc = hm_k_unpack_records(stage[b],m,o,d_index,t->ibyte,t->kmer,D->keys+o,
                                   D->keys_lo ? D->keys_lo+o : NULL,D->cnt+o,D->st);
          s->launches += 1;
          cudaEventRecord(unpacked[b],D->st);
          if (inc && rc == HM_OK)
            { rc = hm_build_bucket_index_range(D->keys,s->n,s->bits,D->bucket,s->idx64,o,o+m,D->st);
              if (rc == HM_OK)
                rc = hm_build_filter_range(D->keys,s->fpos,D->filter,o,o+m,D->st);
              s->launches += 2;
            }
          used[b] = 1;
          b ^= 1;
        }
      pstart += pn;
    }
  cudaStreamSynchronize(D->st_copy);
  cudaError_t e = cudaStreamSynchronize(D->st);
  for (int i = 0; i < 2; i++)
    { dfree(D->dev,D->st,stage[i]); cudaEventDestroy(copied[i]); cudaEventDestroy(unpacked[i]);
      if (pin[i] != NULL) cudaFreeHost(pin[i]);
    }
  if (rc == HM_OK && e != cudaSuccess)
    rc = hm_cuda_fail(e,"unpack");
  return rc;
}

extern "C" int hm_scan_create(const hm_host_table *t, const int *dev, int n_gpus, hm_scan **out)
{ double t0 = now_ms();
  if (t == NULL || out == NULL || n_gpus < 1 || n_gpus > HM_MAX_GPUS)
    return hm_set_error(HM_EINVAL,"hm_scan_create: bad arguments");
  if (t->kmer < 1 || t->kmer > HM_MAX_KMER)
    return hm_set_error(HM_EUNSUPPORTED,"k-mer length %d not supported by this build (1..%d)",
                        t->kmer,HM_MAX_KMER);
  int kbyte = (t->kmer+3)>>2;
  if (t->ibyte < 1 || t->ibyte > 3 || t->ibyte > kbyte)
    return hm_set_error(HM_EFORMAT,"table has ibyte=%d with k=%d",t->ibyte,t->kmer);
  if (hm_device_count() < 1)
    return hm_set_error(HM_ECUDA,"no CUDA device visible (this build has no CPU fallback)");

  hm_scan *s = (hm_scan *) calloc(1,sizeof(hm_scan));
  if (s == NULL)
    return hm_set_error(HM_ENOMEM,"out of host memory");
  s->kmer = t->kmer; s->ibyte = t->ibyte; s->n = t->nels; s->ngpu = n_gpus;
  s->bits  = hm_pick_bucket_bits(s->n);
  s->fpos  = hm_pick_filter_bits(s->n);
  s->idx64 = (s->n >= 0xFFFFFFF0ll);
  int64_t n  = s->n;
  size_t  ib = s->idx64 ? 8 : 4;
  int64_t ixlen = (int64_t) 1 << (8*t->ibyte);
  int     rc = HM_OK;

  if (n_gpus > 1 && (rc = hm_peer_enable(dev,n_gpus)) != HM_OK)
    { free(s); return rc; }

  for (int g = 0; g < n_gpus && rc == HM_OK; g++)
    { DevTable *D = s->d+g;
      D->dev = dev ? dev[g] : g;
      D->lo  = n*g/n_gpus;
      D->hi  = n*(g+1)/n_gpus;
      cudaError_t e;
#define TRY(call) if (rc == HM_OK && (e = (call)) != cudaSuccess) rc = hm_cuda_fail(e,#call)
      TRY(cudaSetDevice(D->dev));
      pool_setup(D->dev,n_gpus == 1);
      TRY(cudaStreamCreateWithFlags(&D->st,cudaStreamNonBlocking));
      TRY(cudaStreamCreateWithFlags(&D->st_copy,cudaStreamNonBlocking));
      TRY(dalloc(D->dev,D->st,(void **) &D->keys,sizeof(uint64_t)*(size_t) (n+1)));
      if (t->kmer > 32)
        TRY(dalloc(D->dev,D->st,(void **) &D->keys_lo,sizeof(uint64_t)*(size_t) (n+1)));
      TRY(dalloc(D->dev,D->st,(void **) &D->cnt,sizeof(uint16_t)*(size_t) (n+1)));
      TRY(dalloc(D->dev,D->st,(void **) &D->deg,(size_t) ((n+4)&~3ll)));
      TRY(dalloc(D->dev,D->st,(void **) &D->bucket,ib*(((size_t) 1<<s->bits)+1)));
      TRY(dalloc(D->dev,D->st,(void **) &D->filter,sizeof(uint32_t)*(size_t) hm_filter_words(s->fpos)));
      TRY(dalloc(D->dev,D->st,(void **) &D->up,ib*(size_t) (D->hi-D->lo+1)));
      TRY(dalloc(D->dev,D->st,(void **) &D->plot,sizeof(unsigned long long)*HM_PLOT_CELLS));
#undef TRY
    }

  double t_alloc = now_ms();
  /* each device unpacks its own shard from the host, then the shards are exchanged over peer
   * copies so that every device ends with the full table                                      */
  for (int g = 0; g < n_gpus && rc == HM_OK; g++)
    { DevTable *D = s->d+g;
      int64_t  *d_index = NULL;
      cudaSetDevice(D->dev);
      cudaError_t e = dalloc(D->dev,D->st,(void **) &d_index,sizeof(int64_t)*ixlen);
      if (e != cudaSuccess) { rc = hm_cuda_fail(e,"cudaMalloc(stub index)"); break; }
      e = cudaMemcpyAsync(d_index,t->index,sizeof(int64_t)*ixlen,cudaMemcpyHostToDevice,D->st);
      if (e != cudaSuccess) { rc = hm_cuda_fail(e,"cudaMemcpyAsync(stub index)"); dfree(D->dev,D->st,d_index); break; }
      cudaStreamSynchronize(D->st);          /* pool memory is about to be used on the copy stream too */
      rc = load_range(s,D,t,d_index,D->lo,D->hi-D->lo);
      dfree(D->dev,D->st,d_index);
    }
  double t_rec = now_ms();
  if (n_gpus > 1 && rc == HM_OK)
    { for (int g = 0; g < n_gpus && rc == HM_OK; g++)        /* all-gather by peer copies */
        for (int h = 0; h < n_gpus && rc == HM_OK; h++)
          if (h != g)
            { DevTable *S = s->d+h, *D = s->d+g;
              int64_t m = S->hi-S->lo;
              if (m <= 0) continue;
              cudaSetDevice(D->dev);
              cudaError_t e = cudaMemcpyPeerAsync(D->keys+S->lo,D->dev,S->keys+S->lo,S->dev,
                                                  sizeof(uint64_t)*(size_t) m,D->st);
              if (e == cudaSuccess && D->keys_lo != NULL)
                e = cudaMemcpyPeerAsync(D->keys_lo+S->lo,D->dev,S->keys_lo+S->lo,S->dev,
                                        sizeof(uint64_t)*(size_t) m,D->st);
              if (e == cudaSuccess)
                e = cudaMemcpyPeerAsync(D->cnt+S->lo,D->dev,S->cnt+S->lo,S->dev,
                                        sizeof(uint16_t)*(size_t) m,D->st);
              if (e != cudaSuccess) rc = hm_cuda_fail(e,"cudaMemcpyPeerAsync(table shard)");
            }
      for (int g = 0; g < n_gpus; g++)
        { cudaSetDevice(s->d[g].dev); cudaStreamSynchronize(s->d[g].st); }
    }
  for (int g = 0; g < n_gpus && rc == HM_OK && n_gpus > 1; g++)     /* (one GPU: done chunk-wise) */
    { DevTable *D = s->d+g;
      cudaSetDevice(D->dev);
      rc = hm_k_build_bucket_index(D->keys,n,s->bits,D->bucket,s->idx64,D->st);
      if (rc == HM_OK)
        rc = hm_k_build_filter(D->keys,n,s->fpos,D->filter,D->st);
      s->launches += 2;
    }
  for (int g = 0; g < n_gpus; g++)
    { cudaSetDevice(s->d[g].dev);
      cudaError_t e = cudaStreamSynchronize(s->d[g].st);
      if (rc == HM_OK && e != cudaSuccess) rc = hm_cuda_fail(e,"table load");
    }
  if (rc != HM_OK)
    { hm_scan_destroy(s); return rc; }
  s->ms_load = now_ms()-t0;
  s->ms_alloc = t_alloc-t0; s->ms_records = t_rec-t_alloc; s->ms_index = now_ms()-t_rec;
  *out = s;
  return HM_OK;
}

/* Trim (count >= ethresh) and / or symmetrise (add reverse complements) the device-resident table
 * in place: what the reference gets from `Logex` and `Symmex` (PloidyPlot.c:1381-1426), without
 * leaving the GPU.  Every device conditions its own replica (deterministic, identical results);
 * the index structures and work buffers are rebuilt for the new size.                          */
extern "C" int hm_scan_condition(hm_scan *s, int ethresh, int do_trim, int do_symm, int64_t *nels_out)
{ int     G = s->ngpu, rc = HM_OK;
  int64_t n_new = -1;
  if (!do_trim && !do_symm)
    { if (nels_out) *nels_out = s->n;
      return HM_OK;
    }
  for (int g = 0; g < G && rc == HM_OK; g++)
    { DevTable *D = s->d+g;
      int64_t   n = s->n;
      HM_CUDA(cudaSetDevice(D->dev));
      HM_CUDA(cudaStreamSynchronize(D->st));
      dfree(D->dev,D->st,D->deg);    D->deg = NULL;
      dfree(D->dev,D->st,D->up);     D->up = NULL;
      dfree(D->dev,D->st,D->bucket); D->bucket = NULL;
      dfree(D->dev,D->st,D->filter); D->filter = NULL;
      /* the table arrays are about to be replaced by plain cudaMalloc'ed ones: hand pooled ones back */
      { PoolReg *R = g_pool + (D->dev < 64 ? D->dev : 0);
        void *arr[3] = { D->keys, D->keys_lo, D->cnt };
        for (int a = 0; a < 3; a++)
          for (int k = 0; k < R->n; k++)
            if (arr[a] != NULL && R->p[k] == arr[a])
              { /* conditioning frees these with cudaFree, which is legal for pool memory */
                R->p[k] = R->p[--R->n];
              }
      }
      rc = hm_condition_arrays(s->kmer,ethresh,do_trim,do_symm,&D->keys,&D->keys_lo,&D->cnt,&n,D->st);
      s->launches += 6;
      if (rc != HM_OK) return rc;
      if (n_new >= 0 && n != n_new)
        return hm_set_error(HM_ECUDA,"conditioning gave %lld entries on GPU %d but %lld on GPU 0",
                            (long long) n,D->dev,(long long) n_new);
      n_new = n;
    }
  s->n     = n_new;
  s->bits  = hm_pick_bucket_bits(s->n);
  s->fpos  = hm_pick_filter_bits(s->n);
  s->idx64 = (s->n >= 0xFFFFFFF0ll);
  size_t ib = s->idx64 ? 8 : 4;
  for (int g = 0; g < G && rc == HM_OK; g++)
    { DevTable *D = s->d+g;
      int64_t   n = s->n;
      cudaError_t e;
      D->lo = n*g/G;
      D->hi = n*(g+1)/G;
#define TRY(call) if (rc == HM_OK && (e = (call)) != cudaSuccess) rc = hm_cuda_fail(e,#call)
      TRY(cudaSetDevice(D->dev));
      TRY(dalloc(D->dev,D->st,(void **) &D->deg,(size_t) ((n+4)&~3ll)));
      TRY(dalloc(D->dev,D->st,(void **) &D->bucket,ib*(((size_t) 1<<s->bits)+1)));
      TRY(dalloc(D->dev,D->st,(void **) &D->filter,sizeof(uint32_t)*(size_t) hm_filter_words(s->fpos)));
      TRY(dalloc(D->dev,D->st,(void **) &D->up,ib*(size_t) (D->hi-D->lo+1)));
#undef TRY
      if (rc == HM_OK)
        rc = hm_k_build_bucket_index(D->keys,n,s->bits,D->bucket,s->idx64,D->st);
      if (rc == HM_OK)
        rc = hm_k_build_filter(D->keys,n,s->fpos,D->filter,D->st);
      s->launches += 2;
    }
  for (int g = 0; g < G; g++)
    { cudaSetDevice(s->d[g].dev);
      cudaError_t e = cudaStreamSynchronize(s->d[g].st);
      if (rc == HM_OK && e != cudaSuccess) rc = hm_cuda_fail(e,"re-index after conditioning");
    }
  if (nels_out) *nels_out = s->n;
  return rc;
}

/* reverse complement of a left-aligned packed k-mer (k <= 32) */
static uint64_t revcomp64(uint64_t x, int k)
{ x = ~x;
  x = ((x >> 2)  & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
  x = ((x >> 4)  & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
  x = ((x >> 8)  & 0x00FF00FF00FF00FFull) | ((x & 0x00FF00FF00FF00FFull) << 8);
  x = ((x >> 16) & 0x0000FFFF0000FFFFull) | ((x & 0x0000FFFF0000FFFFull) << 16);
  x = (x >> 32) | (x << 32);
  if (k < 32)
    x = (x & (((uint64_t) 1 << (2*k))-1)) << (64-2*k);
  return x;
}

/* reverse complement of a left-aligned packed k-mer of 33..64 bases held in two words */
static void revcomp128(uint64_t hi, uint64_t lo, int k, uint64_t *rhi, uint64_t *rlo)
{ /* reversing all 64 slots swaps the words; the k real bases end up right-aligned over 128 bits */
  uint64_t a = revcomp64(lo,32), b = revcomp64(hi,32);      /* full-word reverse complements */
  int      sh = 2*(64-k);                                    /* pad slots now sit on top: shift them out */
  if (sh == 0) { *rhi = a; *rlo = b; }
  else         { *rhi = (a << sh) | (b >> (64-sh)); *rlo = b << sh; }
}

/* examine_table (PloidyPlot.c:1167-1230).  trim: smallest non-zero count among the middle <=1e8
 * entries >= ethresh.  symm: reverse complement of entry 1 (moving on past palindromes, where
 * the reference's loop would never terminate) is present.                                      */
extern "C" int hm_scan_examine(hm_scan *s, int ethresh, int *trim, int *symm)
{ DevTable *D = s->d;
  int64_t   n = s->n, frst, last;
  int       h_min = 0x8000, *d_min = NULL;
  uint64_t *d_q = NULL;
  int64_t  *d_pos = NULL;
  int       two = (D->keys_lo != NULL);

  HM_CUDA(cudaSetDevice(D->dev));
  if (n+3 < 100000000) { frst = 0; last = n; }
  else { frst = n/2-50000000; last = n/2+50000000; }
  HM_CUDA(cudaMalloc(&d_min,sizeof(int)));
  HM_CUDA(cudaMemcpyAsync(d_min,&h_min,sizeof(int),cudaMemcpyHostToDevice,D->st));
  int rc = hm_k_min_count(D->cnt,frst,last,d_min,D->st);
  s->launches += 1;
  if (rc == HM_OK)
    { cudaError_t e = cudaMemcpyAsync(&h_min,d_min,sizeof(int),cudaMemcpyDeviceToHost,D->st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(D->st);
      if (e != cudaSuccess) rc = hm_cuda_fail(e,"min_count");
    }
  cudaFree(d_min);
  if (rc != HM_OK)
    return rc;
  *trim = (h_min >= ethresh);

  *symm = 1;
  HM_CUDA(cudaMalloc(&d_q,2*sizeof(uint64_t)));
  HM_CUDA(cudaMalloc(&d_pos,sizeof(int64_t)));
  for (int64_t sidx = 1; sidx < n; sidx++)
    { uint64_t x, xw = 0, q[2];
      int64_t  pos;
      cudaError_t e = cudaMemcpyAsync(&x,D->keys+sidx,sizeof(uint64_t),cudaMemcpyDeviceToHost,D->st);
      if (e == cudaSuccess && two)
        e = cudaMemcpyAsync(&xw,D->keys_lo+sidx,sizeof(uint64_t),cudaMemcpyDeviceToHost,D->st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(D->st);
      if (e != cudaSuccess) { rc = hm_cuda_fail(e,"examine: key fetch"); break; }
      if (two) revcomp128(x,xw,s->kmer,q,q+1);
      else     { q[0] = revcomp64(x,s->kmer); q[1] = 0; }
      cudaMemcpyAsync(d_q,q,2*sizeof(uint64_t),cudaMemcpyHostToDevice,D->st);
      rc = hm_k_find_keys(D->keys,D->keys_lo,n,D->bucket,s->bits,s->idx64,d_q,two ? d_q+1 : NULL,1,d_pos,D->st);
      s->launches += 1;
      if (rc != HM_OK) break;
      e = cudaMemcpyAsync(&pos,d_pos,sizeof(int64_t),cudaMemcpyDeviceToHost,D->st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(D->st);
      if (e != cudaSuccess) { rc = hm_cuda_fail(e,"examine: lookup"); break; }
      if (pos < 0) { *symm = 0; break; }
      if (pos != sidx) { *symm = 1; break; }
    }
  cudaFree(d_q); cudaFree(d_pos);
  return rc;
}

extern "C" int hm_scan_run(hm_scan *s, int64_t *plot, hm_scan_stats *stats)
{ double      t0 = now_ms();
  int         G = s->ngpu, rc = HM_OK;
  int64_t     n = s->n, launches0 = s->launches;
  cudaEvent_t ev[HM_MAX_GPUS][4];
  float       ms1 = 0, ms2 = 0, msall = 0;

  /* several GPUs: foreign incidence bytes are reached through the owner's array (remote atomics
   * in pass 1, remote loads in pass 2) when every pair of GPUs has native NVLink atomics; otherwise
   * the partial arrays are summed by the peer-memory kernel of hm_peer.cu                        */
  int        peer_mode = (G > 1);
  hm_shards *sh = s->sh;
  for (int a = 0; a < G && peer_mode; a++)
    for (int b = a+1; b < G && peer_mode; b++)
      if (!hm_p2p_native_atomics(s->d[a].dev,s->d[b].dev))
        peer_mode = 0;
  if (getenv("HETMERS_DENSE_EXCHANGE") != NULL)
    peer_mode = 0;
  if (peer_mode)
    for (int g = 0; g < G; g++)
      { memset(&sh[g],0,sizeof(hm_shards));
        sh[g].n_shards = G; sh[g].self = g;
        for (int r = 0; r < G; r++)
          { sh[g].off[r] = s->d[r].lo; sh[g].deg[r] = s->d[r].deg; }
        sh[g].off[G] = n;
        DevTable *D = s->d+g;
        int64_t need = hm_pass2_scratch_bytes(D->hi-D->lo,s->idx64);
        if (D->p2scratch == NULL || D->p2scratch_bytes < need)
          { HM_CUDA(cudaSetDevice(D->dev));
            if (D->p2scratch) cudaFree(D->p2scratch);
            HM_CUDA(cudaMalloc(&D->p2scratch,(size_t) need));
            D->p2scratch_bytes = need;
          }
        sh[g].scratch = D->p2scratch; sh[g].scratch_bytes = D->p2scratch_bytes;
      }

  for (int g = 0; g < G; g++)
    { DevTable *D = s->d+g;
      HM_CUDA(cudaSetDevice(D->dev));
      for (int k = 0; k < 4; k++)
        HM_CUDA(cudaEventCreate(&ev[g][k]));
      HM_CUDA(cudaEventRecord(ev[g][0],D->st));
      HM_CUDA(cudaMemsetAsync(D->deg,0,(size_t) ((n+4)&~3ll),D->st));
      HM_CUDA(cudaMemsetAsync(D->plot,0,sizeof(unsigned long long)*HM_PLOT_CELLS,D->st));
    }
  if (peer_mode)                                /* nobody adds to a peer before it has been zeroed */
    for (int g = 0; g < G; g++)
      { HM_CUDA(cudaSetDevice(s->d[g].dev)); HM_CUDA(cudaStreamSynchronize(s->d[g].st)); }
  for (int g = 0; g < G; g++)
    { DevTable *D = s->d+g;
      HM_CUDA(cudaSetDevice(D->dev));
      rc = hm_k_pass1_degree(D->keys,D->keys_lo,D->cnt,n,D->bucket,s->bits,s->idx64,D->filter,s->fpos,s->kmer,
                             D->lo,D->hi,D->deg,D->up,peer_mode ? &sh[g] : NULL,D->st);
      if (rc != HM_OK) return rc;
      s->launches += (D->hi > D->lo);
      HM_CUDA(cudaEventRecord(ev[g][1],D->st));
    }
  if (G > 1 && !peer_mode)
    { uint8_t *deg[HM_MAX_GPUS]; int64_t lo[HM_MAX_GPUS], hi[HM_MAX_GPUS];
      int dev[HM_MAX_GPUS]; cudaStream_t st[HM_MAX_GPUS];
      for (int g = 0; g < G; g++)
        { deg[g] = s->d[g].deg; lo[g] = s->d[g].lo; hi[g] = s->d[g].hi;
          dev[g] = s->d[g].dev; st[g] = s->d[g].st;
        }
      rc = hm_peer_sum_deg(deg,lo,hi,dev,st,G,n);
      if (rc != HM_OK) return rc;
      s->launches += G;
    }
  if (peer_mode)                                /* every pass 1 (and its remote atomics) has landed */
    for (int g = 0; g < G; g++)
      { HM_CUDA(cudaSetDevice(s->d[g].dev)); HM_CUDA(cudaStreamSynchronize(s->d[g].st)); }
  for (int g = 0; g < G; g++)
    { DevTable *D = s->d+g;
      HM_CUDA(cudaSetDevice(D->dev));
      HM_CUDA(cudaEventRecord(ev[g][2],D->st));
      rc = hm_k_pass2_plot(D->cnt,D->deg,D->up,s->idx64,D->lo,D->hi,D->plot,
                           peer_mode ? &sh[g] : NULL,D->st);
      if (rc != HM_OK) return rc;
      s->launches += (D->hi > D->lo);
      HM_CUDA(cudaEventRecord(ev[g][3],D->st));
    }
  if (G > 1)
    { unsigned long long *pl[HM_MAX_GPUS]; int dev[HM_MAX_GPUS]; cudaStream_t st[HM_MAX_GPUS];
      for (int g = 0; g < G; g++)
        { pl[g] = s->d[g].plot; dev[g] = s->d[g].dev; st[g] = s->d[g].st; }
      rc = hm_peer_sum_plot(pl,dev,st,G);
      if (rc != HM_OK) return rc;
      s->launches += 1;
    }
  HM_CUDA(cudaSetDevice(s->d[0].dev));
  HM_CUDA(cudaMemcpyAsync(plot,s->d[0].plot,sizeof(int64_t)*HM_PLOT_CELLS,
                          cudaMemcpyDeviceToHost,s->d[0].st));
  for (int g = 0; g < G; g++)
    { HM_CUDA(cudaSetDevice(s->d[g].dev));
      HM_CUDA(cudaStreamSynchronize(s->d[g].st));
    }
  double t1 = now_ms();
  for (int g = 0; g < G; g++)
    { float a = 0, b = 0, c = 0;
      cudaSetDevice(s->d[g].dev);
      cudaEventElapsedTime(&a,ev[g][0],ev[g][1]);
      cudaEventElapsedTime(&b,ev[g][2],ev[g][3]);
      cudaEventElapsedTime(&c,ev[g][0],ev[g][3]);
      if (a > ms1) ms1 = a;
      if (b > ms2) ms2 = b;
      if (c > msall) msall = c;
      for (int k = 0; k < 4; k++)
        cudaEventDestroy(ev[g][k]);
    }
  s->ran = 1; s->peer_mode = peer_mode;
  if (stats != NULL)
    { stats->nels = n; stats->n_gpus = G; stats->bucket_bits = s->bits;
      stats->filter_bits = s->fpos; stats->reserved = 0;
      stats->ms_h2d_unpack = s->ms_load;
      stats->ms_pass1 = ms1; stats->ms_pass2 = ms2;
      stats->ms_scan = G > 1 ? (t1-t0) : msall;
      stats->ms_total = s->ms_load + (t1-t0);
      stats->kernel_launches = s->launches;
      stats->ms_alloc = s->ms_alloc; stats->ms_records = s->ms_records; stats->ms_index = s->ms_index;
    }
  (void) launches0;
  return HM_OK;
}

static int rec_cmp(const void *a, const void *b)
{ const hm_pair_rec *x = (const hm_pair_rec *) a, *y = (const hm_pair_rec *) b;
  if (x->smudge != y->smudge) return (x->smudge < y->smudge ? -1 : 1);
  if (x->key_hi != y->key_hi) return (x->key_hi < y->key_hi ? -1 : 1);
  if (x->key_lo != y->key_lo) return (x->key_lo < y->key_lo ? -1 : 1);
  if (x->pos != y->pos)       return (x->pos < y->pos ? -1 : 1);
  return ((int) x->alt - (int) y->alt);
}

/* extract_kmer_pairs' output as a list (PloidyList.c:425-450): needs the incidence array and the
 * recorded partners of a preceding hm_scan_run.  Two launches per GPU: count, then fill.        */
extern "C" int hm_scan_extract(hm_scan *s, const uint16_t *pixmap, hm_pair_rec **out, int64_t *n_out)
{ int G = s->ngpu, rc = HM_OK;
  if (!s->ran)
    return hm_set_error(HM_EINVAL,"hm_scan_extract needs a preceding hm_scan_run");
  int64_t      total = 0, cnts[HM_MAX_GPUS];
  hm_pair_rec *d_out[HM_MAX_GPUS];
  uint16_t    *d_pix[HM_MAX_GPUS];
  unsigned long long *d_cnt[HM_MAX_GPUS];
  memset(d_out,0,sizeof(d_out)); memset(d_pix,0,sizeof(d_pix)); memset(d_cnt,0,sizeof(d_cnt));
  for (int pass = 0; pass < 2 && rc == HM_OK; pass++)
    { for (int g = 0; g < G && rc == HM_OK; g++)
        { DevTable *D = s->d+g;
          HM_CUDA(cudaSetDevice(D->dev));
          if (pass == 0)
            { HM_CUDA(cudaMalloc(&d_pix[g],sizeof(uint16_t)*HM_PLOT_CELLS));
              HM_CUDA(cudaMalloc(&d_cnt[g],sizeof(unsigned long long)));
              HM_CUDA(cudaMemcpyAsync(d_pix[g],pixmap,sizeof(uint16_t)*HM_PLOT_CELLS,cudaMemcpyHostToDevice,D->st));
            }
          else if (cnts[g] > 0)
            HM_CUDA(cudaMalloc(&d_out[g],sizeof(hm_pair_rec)*(size_t) cnts[g]));
          HM_CUDA(cudaMemsetAsync(d_cnt[g],0,sizeof(unsigned long long),D->st));
          rc = hm_k_pass2_extract(D->keys,D->keys_lo,D->cnt,D->deg,D->up,s->idx64,D->lo,D->hi,d_pix[g],
                                  d_out[g],pass == 0 ? 0 : cnts[g],d_cnt[g],s->peer_mode ? &s->sh[g] : NULL,D->st);
          s->launches += (D->hi > D->lo);
        }
      for (int g = 0; g < G && rc == HM_OK; g++)
        { unsigned long long c = 0;
          HM_CUDA(cudaSetDevice(s->d[g].dev));
          HM_CUDA(cudaMemcpyAsync(&c,d_cnt[g],sizeof(c),cudaMemcpyDeviceToHost,s->d[g].st));
          HM_CUDA(cudaStreamSynchronize(s->d[g].st));
          if (pass == 0) { cnts[g] = (int64_t) c; total += cnts[g]; }
        }
    }
  hm_pair_rec *host = (hm_pair_rec *) malloc(sizeof(hm_pair_rec)*(size_t) (total > 0 ? total : 1));
  if (host == NULL && rc == HM_OK)
    rc = hm_set_error(HM_ENOMEM,"out of host memory for %lld pair records",(long long) total);
  int64_t at = 0;
  for (int g = 0; g < G; g++)
    { cudaSetDevice(s->d[g].dev);
      if (rc == HM_OK && cnts[g] > 0)
        { cudaError_t e = cudaMemcpy(host+at,d_out[g],sizeof(hm_pair_rec)*(size_t) cnts[g],cudaMemcpyDeviceToHost);
          if (e != cudaSuccess) rc = hm_cuda_fail(e,"cudaMemcpy(pair records)");
          at += cnts[g];
        }
      if (d_out[g]) cudaFree(d_out[g]);
      if (d_pix[g]) cudaFree(d_pix[g]);
      if (d_cnt[g]) cudaFree(d_cnt[g]);
    }
  if (rc != HM_OK)
    { free(host); return rc; }
  qsort(host,(size_t) total,sizeof(hm_pair_rec),rec_cmp);      /* deterministic order */
  *out = host; *n_out = total;
  return HM_OK;
}

extern "C" int hm_hetmers_host(const hm_host_table *t, const int *dev, int n_gpus,
                               int64_t *plot, hm_scan_stats *stats)
{ hm_scan *s = NULL;
  int rc = hm_scan_create(t,dev,n_gpus,&s);
  if (rc != HM_OK)
    return rc;
  rc = hm_scan_run(s,plot,stats);
  hm_scan_destroy(s);
  return rc;
}

extern "C" int hm_scan_download(hm_scan *s, uint64_t *keys, uint64_t *keys_lo, uint16_t *cnt, uint8_t *deg)
{ DevTable *D = s->d;
  HM_CUDA(cudaSetDevice(D->dev));
  HM_CUDA(cudaStreamSynchronize(D->st));
  if (keys != NULL)
    HM_CUDA(cudaMemcpy(keys,D->keys,sizeof(uint64_t)*(size_t) s->n,cudaMemcpyDeviceToHost));
  if (keys_lo != NULL && D->keys_lo != NULL)
    HM_CUDA(cudaMemcpy(keys_lo,D->keys_lo,sizeof(uint64_t)*(size_t) s->n,cudaMemcpyDeviceToHost));
  if (cnt != NULL)
    HM_CUDA(cudaMemcpy(cnt,D->cnt,sizeof(uint16_t)*(size_t) s->n,cudaMemcpyDeviceToHost));
  if (deg != NULL)                      /* every owner's slice (identical copies in dense mode) */
    for (int g = 0; g < s->ngpu; g++)
      { DevTable *O = s->d+g;
        HM_CUDA(cudaSetDevice(O->dev));
        HM_CUDA(cudaStreamSynchronize(O->st));
        if (O->hi > O->lo)
          HM_CUDA(cudaMemcpy(deg+O->lo,O->deg+O->lo,(size_t) (O->hi-O->lo),cudaMemcpyDeviceToHost));
      }
  return HM_OK;
}
