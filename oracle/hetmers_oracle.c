/*******************************************************************************************
 * hetmers_oracle.c  --  TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT.
 *
 * A single-threaded, plain-C restatement of what the reference `hetmers` backend computes
 * (KamilSJaron/smudgeplot src/lib/PloidyPlot.c + the Kmer_Stream reader of src/lib/libfastk.c),
 * used only as the checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
 * The CUDA product path (smudgeplot_b200/csrc) never includes, links or calls this file.
 *
 * Parity pinning: the reference ships NO golden vectors or tests for this path (SURVEY.md §4,
 * §8c).  This restatement is therefore pinned against outputs of the reference itself: the
 * unmodified reference C compiled by oracle/Makefile into oracle/_ref/hetmers, run on seeded
 * synthetic FastK tables by tests/golden/make_golden.py; the resulting (.ktab, .smu) pairs are
 * committed under tests/golden/ and tests/test_oracle.py requires this file to reproduce every
 * one of them byte for byte.
 *
 * What is restated (reference file:line):
 *   - FastK table layout: stub header + prefix index, hidden part files, suffix||count records
 *     (libfastk.c:786-908 Open_Kmer_Stream; :1230-1269 Current_Entry)        -> oracle_load_table
 *   - trimmed?/symmetric? decisions (PloidyPlot.c:1167-1230 examine_table)     -> oracle_examine
 *   - the two-pass, k-level, 4-list merge on the suffix after the varying base
 *     (PloidyPlot.c:454-700 analysis_in_core_1/2, :851-923 in_core_recursion); here as one
 *     recursion over the in-memory sorted table                               -> oracle_scan
 *   - uint8 incidence array `Pair` with wrap-around (PloidyPlot.c:163,:260-261)
 *   - SMAX/FMAX gates and the .smu writer (PloidyPlot.c:48-49,:1603-1617)      -> oracle_write_smu
 *   - extract_kmer_pairs: the same pass 2 writing labelled pairs as sequences
 *     (PloidyList.c:128-165 print_het, :425-450,:680-705, .sma parser :1288-1352) -> oracle_extract_file
 *
 * Works for any k (keys are kept as kbyte-byte big-endian strings), ibyte in {1,2,3}.
 *******************************************************************************************/

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <strings.h>

#define SMAX 1000   /* PloidyPlot.c:48 */
#define FMAX  500   /* PloidyPlot.c:49 */
#define PLOT_W (FMAX+1)
#define PLOT_N ((SMAX+1)*(FMAX+1))

typedef struct
  { int       kmer, kbyte, ibyte, nparts, minval;
    int64_t   nels;
    uint8_t  *keys;     /* nels * kbyte, ascending lexicographic */
    uint16_t *cnt;      /* nels */
  } OTable;

/* ---------------------------------------------------------------- FastK table reader ---- */

/* Split <name> the way PathTo/Root do (gene_core.c:64-114): dir part, root without ".ktab" */
static void split_name(const char *name, char *dir, char *root)
{ const char *slash = strrchr(name,'/');
  const char *base  = slash ? slash+1 : name;
  size_t      n;

  if (slash == NULL)
    strcpy(dir,".");
  else if (slash == name)
    strcpy(dir,"/");
  else
    { memcpy(dir,name,slash-name); dir[slash-name] = 0; }
  strcpy(root,base);
  n = strlen(root);
  if (n > 5 && strcasecmp(root+n-5,".ktab") == 0)
    root[n-5] = 0;
}

void oracle_free_table(OTable *T)
{ free(T->keys); free(T->cnt); T->keys = NULL; T->cnt = NULL; }

/* Returns 0 on success, -1 if the stub cannot be opened, -2 on a malformed/missing part.   */
int oracle_load_table(const char *name, OTable *T)
{ char    *dir, *root, *path;
  FILE    *f;
  int32_t  hdr[4];
  int64_t *index, ixlen, nels, n, at, b, i;
  int      p, kbyte, hbyte, pbyte, ibyte, j;
  int32_t  pk;

  memset(T,0,sizeof(*T));
  dir  = malloc(strlen(name)+8);
  root = malloc(strlen(name)+8);
  path = malloc(2*strlen(name)+64);
  split_name(name,dir,root);

  sprintf(path,"%s/%s.ktab",dir,root);
  f = fopen(path,"rb");
  if (f == NULL)
    { free(dir); free(root); free(path); return (-1); }
  if (fread(hdr,sizeof(int32_t),4,f) != 4)         /* kmer, nparts, minval, ibyte (libfastk.c:816-819) */
    { fclose(f); free(dir); free(root); free(path); return (-2); }
  T->kmer = hdr[0]; T->nparts = hdr[1]; T->minval = hdr[2]; T->ibyte = ibyte = hdr[3];
  T->kbyte = kbyte = (T->kmer+3)>>2;               /* libfastk.c:824 */
  hbyte = kbyte-ibyte;                             /* libfastk.c:827 */
  pbyte = hbyte+2;                                 /* libfastk.c:826 */
  ixlen = ((int64_t) 1) << (8*ibyte);
  index = malloc(sizeof(int64_t)*ixlen);
  if (fread(index,sizeof(int64_t),ixlen,f) != (size_t) ixlen)
    { fclose(f); free(index); free(dir); free(root); free(path); return (-2); }
  fclose(f);

  nels = 0;                                        /* part headers, libfastk.c:847-864 */
  for (p = 1; p <= T->nparts; p++)
    { sprintf(path,"%s/.%s.ktab.%d",dir,root,p);
      f = fopen(path,"rb");
      if (f == NULL)
        { free(index); free(dir); free(root); free(path); return (-2); }
      if (fread(&pk,sizeof(int32_t),1,f) != 1 || fread(&n,sizeof(int64_t),1,f) != 1 || pk != T->kmer)
        { fclose(f); free(index); free(dir); free(root); free(path); return (-2); }
      nels += n;
      fclose(f);
    }
  T->nels = nels;
  T->keys = malloc((size_t) (nels > 0 ? nels : 1) * kbyte);
  T->cnt  = malloc(sizeof(uint16_t) * (size_t) (nels > 0 ? nels : 1));

  at = 0;
  b  = 0;                                          /* current prefix bucket                */
  for (p = 1; p <= T->nparts; p++)
    { uint8_t *rec;
      sprintf(path,"%s/.%s.ktab.%d",dir,root,p);
      f = fopen(path,"rb");
      if (fread(&pk,sizeof(int32_t),1,f) != 1 || fread(&n,sizeof(int64_t),1,f) != 1)
        n = 0;
      rec = malloc((size_t) (n > 0 ? n : 1) * pbyte);
      if (fread(rec,pbyte,n,f) != (size_t) n)
        { fclose(f); free(rec); free(index); free(dir); free(root); free(path);
          oracle_free_table(T); return (-2);
        }
      fclose(f);
      for (i = 0; i < n; i++, at++)
        { uint8_t *k = T->keys + at*kbyte;
          uint8_t *r = rec + i*pbyte;
          /* prefix = first bucket whose end offset exceeds the ordinal (libfastk.c:1174-1175) */
          while (b < ixlen && index[b] <= at)
            b += 1;
          for (j = 0; j < ibyte; j++)              /* big-endian prefix bytes (libfastk.c:1246-1261) */
            k[j] = (uint8_t) (b >> (8*(ibyte-1-j)));
          memcpy(k+ibyte,r,hbyte);
          T->cnt[at] = (uint16_t) (r[hbyte] | (r[hbyte+1] << 8));   /* unaligned LE uint16 */
        }
      free(rec);
    }
  free(index); free(dir); free(root); free(path);
  return (0);
}

/* ------------------------------------------------------------------- packed k-mer ops ---- */

static inline int base_at(const uint8_t *k, int pos)        /* PloidyPlot.c:165-166,197 */
{ return ((k[pos>>2] >> (6-2*(pos&3))) & 0x3); }

/* compare two kbyte keys ignoring the base at position `level` */
static inline int cmp_masked(const uint8_t *a, const uint8_t *b, int kbyte, int level)
{ int j, lb = level>>2;
  uint8_t m = (uint8_t) ~(0x3 << (6-2*(level&3)));
  for (j = 0; j < kbyte; j++)
    { uint8_t x = a[j], y = b[j];
      if (j == lb) { x &= m; y &= m; }
      if (x != y)
        return (x < y ? -1 : 1);
    }
  return (0);
}

static int64_t find_key(const uint8_t *keys, int64_t n, int kbyte, const uint8_t *q)
{ int64_t l = 0, r = n;
  while (l < r)
    { int64_t m = (l+r)>>1;
      if (memcmp(keys+m*kbyte,q,kbyte) < 0) l = m+1; else r = m;
    }
  if (l < n && memcmp(keys+l*kbyte,q,kbyte) == 0)
    return (l);
  return (-1);
}

static void revcomp(const uint8_t *k, int kmer, int kbyte, uint8_t *out)
{ int i;
  memset(out,0,kbyte);
  for (i = 0; i < kmer; i++)
    { int c = 3-base_at(k,kmer-1-i);               /* comp[] table, PloidyPlot.c:1133-1141 */
      out[i>>2] |= (uint8_t) (c << (6-2*(i&3)));
    }
}

/* ------------------------------------------------------------------------- examine ------ */

/* PloidyPlot.c:1167-1230.  trim: smallest non-zero count (read as int16) among the middle
 * <=1e8 entries is >= ethresh.  symm: the reverse complement of the entry at index 1 (moving on
 * past palindromes) is present.  Where the reference has undefined behaviour (no non-zero
 * count; a palindrome makes its loop re-read the same k-mer for ever) we terminate instead.   */
void oracle_examine(const OTable *T, int ethresh, int *trim, int *symm)
{ int64_t frst, last, i, sidx;
  int     nz = 0x8000;
  uint8_t *rc;

  if (T->nels+3 < 100000000) { frst = 0; last = T->nels; }
  else { frst = T->nels/2 - 50000000; last = T->nels/2 + 50000000; }
  for (i = frst; i < last; i++)
    { int v = (int16_t) T->cnt[i];
      if (v >= 1 && v < nz) nz = v;
    }
  *trim = (nz >= ethresh);

  rc = malloc(T->kbyte);
  *symm = 1;
  for (sidx = 1; sidx < T->nels; sidx++)
    { int64_t at;
      revcomp(T->keys+sidx*T->kbyte,T->kmer,T->kbyte,rc);
      at = find_key(T->keys,T->nels,T->kbyte,rc);
      if (at < 0) { *symm = 0; break; }
      if (at != sidx) { *symm = 1; break; }
    }
  free(rc);
}

/* ---------------------------------------------------------------------------- scan ------ */

typedef struct
  { const uint8_t  *keys;
    const uint16_t *cnt;
    int             kmer, kbyte;
    int             pass1;
    uint8_t        *pair;      /* incidence array, uint8 with wrap (PloidyPlot.c:163) */
    int64_t        *plot;      /* [SMAX+1][FMAX+1] */
    const uint16_t *pix;       /* extract mode: pixel -> smudge label (PloidyList.c PLOT), else NULL */
    FILE          **out;       /* extract mode: out[label] */
  } Scan;

/* print_het (PloidyList.c:128-165): the k-mer in lower case with "(x/alt)" at position `half` */
static void print_het(const uint8_t *seq, int len, int half, int alt, FILE *f)
{ static const char dna[4] = { 'a', 'c', 'g', 't' };
  int i;
  for (i = 0; i < len; i++)
    { int b = (seq[i>>2] >> (6-2*(i&3))) & 0x3;
      if (i == half)
        fprintf(f,"(%c/%c)",dna[b],dna[alt]);
      else
        fputc(dna[b],f);
    }
  fputc('\n',f);
}

/* One node of the prefix trie: entries [lo,hi) share their first `level` bases.  Split them by
 * the base at `level` into 4 sorted lists, merge the lists on the remaining suffix, treat every
 * tie group of 2-4 heads as mutually one-away k-mers (PloidyPlot.c:494-565 pass 1, :611-697
 * pass 2), then descend into the 4 children (PloidyPlot.c:895-922).                           */
static void scan_node(Scan *S, int64_t lo, int64_t hi, int level)
{ int64_t bnd[5], ptr[4];
  int     a, kb = S->kbyte;

  if (hi-lo < 2 || level >= S->kmer)
    return;

  bnd[0] = lo; bnd[4] = hi;
  for (a = 1; a < 4; a++)                     /* first entry whose base at `level` is >= a */
    { int64_t l = bnd[a-1], r = hi;
      while (l < r)
        { int64_t m = (l+r)>>1;
          if (base_at(S->keys+m*kb,level) < a) l = m+1; else r = m;
        }
      bnd[a] = l;
    }
  for (a = 0; a < 4; a++)
    ptr[a] = bnd[a];

  while (1)
    { int in[4], itop = 0, i;
      int cnt[4];
      const uint8_t *mr = NULL;

      for (a = 0; a < 4; a++)
        if (ptr[a] < bnd[a+1])
          { const uint8_t *hr = S->keys + ptr[a]*kb;
            int v = (mr == NULL) ? -1 : cmp_masked(hr,mr,kb,level);
            if (v < 0) { mr = hr; in[0] = a; itop = 1; }
            else if (v == 0) in[itop++] = a;
          }
      if (itop == 0)
        break;

      if (itop > 1)
        { for (i = 0; i < itop; i++)
            cnt[i] = S->cnt[ptr[in[i]]];
          if (S->pass1)
            { for (i = 1; i < itop; i++)
                for (a = 0; a < i; a++)
                  if (cnt[a]+cnt[i] <= SMAX)                          /* :259 */
                    { S->pair[ptr[in[i]]] += 1;
                      S->pair[ptr[in[a]]] += 1;
                    }
            }
          else
            { for (i = 1; i < itop; i++)
                if (S->pair[ptr[in[i]]] <= 1)                         /* :403 */
                  for (a = 0; a < i; a++)
                    { int x = cnt[a]+cnt[i];
                      if (x <= SMAX && S->pair[ptr[in[a]]] <= 1)      /* :407 */
                        { int mn = (cnt[a] < cnt[i] ? cnt[a] : cnt[i]);
                          S->plot[x*PLOT_W + mn] += 1;
                          if (S->pix != NULL && S->pix[x*PLOT_W + mn] > 0)   /* PloidyList.c:431-447 */
                            { FILE *f = S->out[S->pix[x*PLOT_W + mn]];
                              if (cnt[a] < cnt[i])
                                print_het(S->keys + ptr[in[i]]*kb,S->kmer,level,in[a],f);
                              else
                                print_het(S->keys + ptr[in[a]]*kb,S->kmer,level,in[i],f);
                            }
                        }
                    }
            }
        }
      for (i = 0; i < itop; i++)
        ptr[in[i]] += 1;
    }

  for (a = 0; a < 4; a++)
    scan_node(S,bnd[a],bnd[a+1],level+1);
}

/* keys: n*kbyte sorted packed k-mers; cnt: n counts; plot: int64[1001*501] (zeroed here);
 * deg_out: optional n-byte copy of the incidence array after pass 1.                          */
int oracle_scan(const uint8_t *keys, const uint16_t *cnt, int64_t n, int kmer,
                int64_t *plot, uint8_t *deg_out)
{ Scan S;

  S.keys = keys; S.cnt = cnt; S.kmer = kmer; S.kbyte = (kmer+3)>>2;
  S.plot = plot; S.pix = NULL; S.out = NULL;
  S.pair = calloc((size_t) (n > 0 ? n : 1),1);
  if (S.pair == NULL)
    return (-1);
  memset(plot,0,sizeof(int64_t)*PLOT_N);
  for (S.pass1 = 1; S.pass1 >= 0; S.pass1--)          /* PloidyPlot.c:1489 */
    scan_node(&S,0,n,0);
  if (deg_out != NULL)
    memcpy(deg_out,S.pair,(size_t) n);
  free(S.pair);
  return (0);
}

/* .smu text: "min \t sum-min \t count", sum-major, min < FMAX (PloidyPlot.c:1612-1615) */
int oracle_write_smu(const char *path, const int64_t *plot)
{ FILE *f = fopen(path,"w");
  int   a, i;
  if (f == NULL)
    return (-1);
  for (a = 0; a <= SMAX; a++)
    for (i = 0; i < FMAX; i++)
      if (plot[a*PLOT_W+i] > 0)
        fprintf(f,"%i\t%i\t%lld\n",i,a-i,(long long) plot[a*PLOT_W+i]);
  fclose(f);
  return (0);
}

/* Whole path on a table on disk; returns 0 ok, 1 cannot open, 2 needs conditioning. */
int oracle_hetmers_file(const char *table, int ethresh, const char *smu_path,
                        int *trim, int *symm, int64_t *nels)
{ OTable   T;
  int64_t *plot;
  int      rc;

  if (oracle_load_table(table,&T) != 0)
    return (1);
  oracle_examine(&T,ethresh,trim,symm);
  if (nels != NULL) *nels = T.nels;
  if (!(*trim && *symm))
    { oracle_free_table(&T); return (2); }
  plot = malloc(sizeof(int64_t)*PLOT_N);
  oracle_scan(T.keys,T.cnt,T.nels,T.kmer,plot,NULL);
  rc = oracle_write_smu(smu_path,plot);
  free(plot);
  oracle_free_table(&T);
  return (rc == 0 ? 0 : 1);
}

/* extract_kmer_pairs on a conditioned table: parse <sma> (PloidyList.c:1288-1352), open
 * <out>.<a>A<b>B.txt per smudge, run both passes, write the labelled pairs (traversal order; the
 * reference's order depends on its thread schedule -- compare sorted).  0 ok, 1 cannot open table
 * or smudge file, 2 needs conditioning, 3 malformed .sma.                                      */
int oracle_extract_file(const char *table, int ethresh, const char *sma_path, const char *out_root)
{ OTable    T;
  Scan      S;
  int64_t  *plot;
  uint16_t *pix;
  FILE     *f, *out[65536];
  int       sa[4096], sb[4096], nsm = 0, trim, symm, i, j, a, b, s, pass;
  char      buf[1000], *name;

  f = fopen(sma_path,"r");
  if (f == NULL)
    return (1);
  pix  = calloc(PLOT_N,sizeof(uint16_t));
  name = malloc(strlen(out_root)+64);
  if (fgets(buf,1000,f) == NULL) buf[0] = 0;
  while (fgets(buf,1000,f) != NULL)
    { if (sscanf(buf," %d %d %*d %dA%dB",&i,&j,&a,&b) != 4 || a <= 0 || b <= 0 || a < b ||
          i < 0 || i > FMAX || j < i || i+j > SMAX)
        { fclose(f); free(pix); free(name); return (3); }
      for (s = 0; s < nsm; s++)
        if (sa[s] == a && sb[s] == b)
          break;
      if (s >= nsm)
        { if (nsm >= 4095) { fclose(f); free(pix); free(name); return (3); }
          sa[s] = a; sb[s] = b;
          sprintf(name,"%s.%dA%dB.txt",out_root,a,b);
          out[s+1] = fopen(name,"w");
          if (out[s+1] == NULL) { fclose(f); free(pix); free(name); return (1); }
          nsm += 1;
        }
      pix[(i+j)*PLOT_W+i] = (uint16_t) (s+1);
    }
  fclose(f);
  free(name);

  if (oracle_load_table(table,&T) != 0)
    { for (s = 1; s <= nsm; s++) fclose(out[s]);
      free(pix); return (1);
    }
  oracle_examine(&T,ethresh,&trim,&symm);
  if (!(trim && symm))
    { for (s = 1; s <= nsm; s++) fclose(out[s]);
      oracle_free_table(&T); free(pix); return (2);
    }
  plot = calloc(PLOT_N,sizeof(int64_t));
  S.keys = T.keys; S.cnt = T.cnt; S.kmer = T.kmer; S.kbyte = T.kbyte;
  S.plot = plot; S.pix = NULL; S.out = out;
  S.pair = calloc((size_t) (T.nels > 0 ? T.nels : 1),1);
  for (pass = 1; pass >= 0; pass--)
    { S.pass1 = pass;
      S.pix   = pass ? NULL : pix;
      scan_node(&S,0,T.nels,0);
    }
  for (s = 1; s <= nsm; s++)
    fclose(out[s]);
  free(S.pair); free(plot); free(pix);
  oracle_free_table(&T);
  return (0);
}

#ifdef ORACLE_MAIN
int main(int argc, char *argv[])
{ const char *out = NULL, *src = NULL;
  int   eth = 4, i, trim, symm, rc;
  char *smu;
  int64_t nels;

  for (i = 1; i < argc; i++)
    if (argv[i][0] == '-')
      { if (argv[i][1] == 'o') out = argv[i]+2;
        else if (argv[i][1] == 'e') eth = atoi(argv[i]+2);
      }
    else
      src = argv[i];
  if (src == NULL || out == NULL)
    { fprintf(stderr,"usage: hetmers_oracle -o<out> [-e<L>] <table>[.ktab]\n"); return (1); }
  smu = malloc(strlen(out)+8);
  sprintf(smu,"%s.smu",out);
  rc = oracle_hetmers_file(src,eth,smu,&trim,&symm,&nels);
  if (rc == 1) { fprintf(stderr,"hetmers_oracle: Cannot open k-mer table %s\n",src); return (1); }
  if (rc == 2) { fprintf(stderr,"hetmers_oracle: table needs conditioning (trim=%d symm=%d)\n",trim,symm); return (1); }
  return (0);
}
#endif
