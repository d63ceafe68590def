/*
 * tb_oracle.c — TEST INFRASTRUCTURE ONLY (never linked or loaded by the product path).
 *
 * A plain-C, single-threaded, literal restatement of the reference's hot path, working on the same
 * SoA windows as the CUDA library (include/tiebrush_b200.h). Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * Parity status: PINNED. tests/golden/ holds outputs of the UNMODIFIED compiled reference
 * (oracle/_ref, built by oracle/build_ref.sh) on the reference's own fixtures test/t1, test/t2 and on
 * randomized SAM inputs covering every mode/filter; tests/test_oracle.py checks this file against
 * all of them. (BigWig -W and the -s sample heat-map are out of scope and unpinned.)
 *
 * Each function cites the reference code it follows (paths relative to /root/reference).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#include "../include/tiebrush_b200.h"

typedef struct {
  int mode;           /* TB_MODE_* */
  uint32_t flag_mask; /* -F */
  int max_nh;         /* -N */
  int min_qual;       /* -Q */
  int keep_bits;      /* TB_KEEP_* */
  int collapse_same;  /* -A */
} tbo_opts;

typedef struct { uint32_t start, end; } seg_t;

/* ---- record model: GSamRecord::setupCoordinates, src/GSam.cpp:351-417 --------------------- */
typedef struct {
  uint32_t start, end; /* 1-based inclusive */
  int ex_off, n_ex;    /* into exon arena */
} rec_t;

static int setup_coordinates(int32_t pos, const uint32_t* cig, uint32_t ncig, seg_t* ex /*>= ncig+1*/,
                             uint32_t* out_start, uint32_t* out_end) {
  int l = 0, nex = 0;
  int exstart = pos;
  int intron = 0, ins = 0;
  for (uint32_t i = 0; i < ncig; ++i) {
    uint32_t op = cig[i] & 0xf, len = cig[i] >> 4;
    switch (op) {
      case 7: case 8: case 0: case 2: /* = X M D */
        l += (int)len; intron = 0; ins = 0; break;
      case 3: /* N */
        if (!ins || !intron) { ex[nex].end = (uint32_t)(pos + l); ex[nex].start = (uint32_t)(exstart + 1); nex++; }
        l += (int)len; exstart = pos + l; intron = 1; break;
      case 4: case 5: /* S H */
        intron = 0; ins = 0; break;
      case 1: /* I */
        ins = 1; break;
      default: /* P and unknown: nothing */
        break;
    }
  }
  ex[nex].start = (uint32_t)(exstart + 1); ex[nex].end = (uint32_t)(pos + l); nex++;
  *out_start = (uint32_t)(pos + 1);
  *out_end = (uint32_t)(pos + l);
  return nex;
}

/* ---- YD tracker: GSegList, src/tiebrush.cpp:122-253 ---------------------------------------- */
typedef struct segnode { uint32_t start, end; struct segnode* next; } segnode;
typedef struct { segnode* head; uint32_t last_pos; int last_dist; } seglist;

static void seglist_clear(seglist* L) { /* :140-149 */
  segnode* p = L->head;
  while (p) { segnode* n = p->next; free(p); p = n; }
  L->head = NULL;
}
static void seglist_reset(seglist* L) { seglist_clear(L); L->last_pos = 0; L->last_dist = -1; } /* :132-138 */

static segnode* newnode(uint32_t s, uint32_t e, segnode* nx) {
  segnode* n = (segnode*)malloc(sizeof(segnode)); n->start = s; n->end = e; n->next = nx; return n;
}

static void seglist_clear_to(seglist* L, segnode* to) { /* :151-165 */
  segnode* p = L->head;
  while (p && p != to) { segnode* n = p->next; free(p); p = n; }
  segnode* nx = to->next; free(to); L->head = nx;
}

static void seglist_merge(seglist* L, const seg_t* ex, int nex) { /* mergeRead :167-219 */
  if (L->head == NULL) {
    L->head = newnode(ex[0].start, ex[0].end, NULL);
    segnode* cn = L->head;
    for (int i = 1; i < nex; i++) { segnode* n = newnode(ex[i].start, ex[i].end, NULL); cn->next = n; cn = n; }
    return;
  }
  segnode* n = L->head; segnode* prev = NULL;
  for (int i = 0; i < nex; i++) {
    seg_t e = ex[i];
    while (n) {
      if (e.end < n->start) {
        segnode* nw = newnode(e.start, e.end, n);
        if (n == L->head) L->head = nw; else prev->next = nw;
        prev = nw;
        break;
      }
      if (e.start <= n->end) {
        if (e.start < n->start) n->start = e.start;
        if (e.end > n->end) n->end = e.end;
        segnode* next = n->next;
        while (next && next->start <= n->end) {
          uint32_t nend = next->end;
          n->next = next->next; free(next); next = n->next;
          if (nend > n->end) { n->end = nend; break; }
        }
        break;
      }
      prev = n; n = n->next;
    }
    /* n == NULL here: this and every later exon is dropped, as in the reference */
  }
}

static int seglist_process(seglist* L, uint32_t rstart, const seg_t* ex, int nex) { /* processRead :221-250 */
  if (L->last_pos == rstart) { seglist_merge(L, ex, nex); return L->last_dist; }
  int d = 0;
  segnode* node = L->head; segnode* prev = NULL;
  while (node && node->start < rstart) { prev = node; node = node->next; }
  if (prev) {
    if (prev->end >= rstart) d = (int)(rstart - prev->start);
    if (d == 0) seglist_clear_to(L, prev);
  }
  if (L->last_pos != rstart) { L->last_pos = rstart; L->last_dist = d; }
  seglist_merge(L, ex, nex);
  return d;
}

/* ---- group comparators: src/tiebrush.cpp:275-345, SPData::operator< :438-457 ---------------- */
typedef struct {
  const tb_soa_in* in; const tbo_opts* o; const rec_t* rec; const seg_t* exons;
} cmp_ctx;

static int cmp_flags(const cmp_ctx* c, int64_t a, int64_t b) { /* :275-283 */
  if (c->o->flag_mask == 0) return 0;
  if ((c->o->flag_mask & c->in->flag[a]) == (c->o->flag_mask & c->in->flag[b])) return 0;
  return 1;
}
static int cmp_cigar(const cmp_ctx* c, int64_t a, int64_t b) { /* :306-312 */
  int f = cmp_flags(c, a, b); if (f) return f;
  uint32_t na = c->in->cig_off[a + 1] - c->in->cig_off[a], nb = c->in->cig_off[b + 1] - c->in->cig_off[b];
  if (na != nb) return (int)na - (int)nb;
  if (na == 0) return 0;
  return memcmp(c->in->cigar + c->in->cig_off[a], c->in->cigar + c->in->cig_off[b], na * sizeof(uint32_t));
}
static int cmp_full(const cmp_ctx* c, int64_t a, int64_t b) { /* :285-304 */
  int f = cmp_flags(c, a, b); if (f) return f;
  uint32_t na = c->in->cig_off[a + 1] - c->in->cig_off[a], nb = c->in->cig_off[b + 1] - c->in->cig_off[b];
  if (na != nb) return (int)na - (int)nb;
  int cc = 0;
  if (na > 0) cc = memcmp(c->in->cigar + c->in->cig_off[a], c->in->cigar + c->in->cig_off[b], na * sizeof(uint32_t));
  if (cc) return cc;
  uint32_t la = c->in->md_off[a + 1] - c->in->md_off[a], lb = c->in->md_off[b + 1] - c->in->md_off[b];
  const char* am = la ? (const char*)c->in->md + c->in->md_off[a] : NULL;
  const char* bm = lb ? (const char*)c->in->md + c->in->md_off[b] : NULL;
  if (am == NULL || bm == NULL) { if (am == bm) return 0; if (am != NULL) return 1; return -1; }
  return strcmp(am, bm);
}
static int cmp_clip(const cmp_ctx* c, int64_t a, int64_t b) { /* :314-332 */
  int f = cmp_flags(c, a, b); if (f) return f;
  uint32_t al = c->in->cig_off[a + 1] - c->in->cig_off[a], bl = c->in->cig_off[b + 1] - c->in->cig_off[b];
  const uint32_t* as = c->in->cigar + c->in->cig_off[a]; const uint32_t* bs = c->in->cigar + c->in->cig_off[b];
  while (al > 0 && (as[0] & 0xf) == 4) { as++; al--; }
  while (al > 0 && (as[al - 1] & 0xf) == 4) al--;
  while (bl > 0 && (bs[0] & 0xf) == 4) { bs++; bl--; }
  while (bl > 0 && (bs[bl - 1] & 0xf) == 4) bl--;
  if (al != bl) return (int)al - (int)bl;
  if (al == 0) return 0;
  return memcmp(as, bs, al * sizeof(uint32_t));
}
static int cmp_exons(const cmp_ctx* c, int64_t a, int64_t b) { /* :334-345 */
  int f = cmp_flags(c, a, b); if (f) return f;
  const rec_t* ra = &c->rec[a]; const rec_t* rb = &c->rec[b];
  if (ra->n_ex != rb->n_ex) return ra->n_ex - rb->n_ex;
  for (int i = 0; i < ra->n_ex; i++) {
    seg_t ea = c->exons[ra->ex_off + i], eb = c->exons[rb->ex_off + i];
    if (ea.start != eb.start) return (int)ea.start - (int)eb.start;
    if (ea.end != eb.end) return (int)ea.end - (int)eb.end;
  }
  return 0;
}
static int sp_less(const cmp_ctx* c, int64_t a, int64_t b) { /* SPData::operator< :438-457 (tid equal inside a window) */
  const rec_t* ra = &c->rec[a]; const rec_t* rb = &c->rec[b];
  if (ra->start != rb->start) return ra->start < rb->start;
  char sa = (char)c->in->strand[a], sb = (char)c->in->strand[b];
  if (sa != sb) return sa < sb;
  if (ra->end != rb->end) return ra->end < rb->end;
  switch (c->o->mode) {
    case TB_MODE_FULL: return cmp_full(c, a, b) < 0;
    case TB_MODE_CLIP: return cmp_clip(c, a, b) < 0;
    case TB_MODE_EXON: return cmp_exons(c, a, b) < 0;
    default: return cmp_cigar(c, a, b) < 0;
  }
}
/* GList::DefaultCompareProc, include/gclib/GList.hh:85-90 */
static int sp_compare(const cmp_ctx* c, int64_t x, int64_t y) {
  if (sp_less(c, y, x)) return 1;
  if (sp_less(c, x, y)) return -1;
  return 0;
}

/* ---- groups (SPData, src/tiebrush.cpp:350-473) --------------------------------------------- */
typedef struct {
  int64_t rep;
  double accYC; int64_t accYX; int64_t maxYD;
  uint8_t* samples; /* one byte per file (GBitVec) */
} group_t;

/* GList<SPData>::Found, include/gclib/GList.hh:567-604 */
static int list_found(const cmp_ctx* c, group_t** L, int n, int64_t item, int* idx) {
  *idx = -1;
  if (n == 0) { *idx = 0; return 0; }
  if (sp_compare(c, L[0]->rep, item) > 0) { *idx = 0; return 0; }
  if (sp_compare(c, item, L[n - 1]->rep) > 0) { *idx = n; return 0; }
  int l = 0, h = n - 1;
  while (l <= h) {
    int i = l + ((h - l) >> 1);
    int cc = sp_compare(c, L[i]->rep, item);
    if (cc < 0) l = i + 1;
    else { h = i - 1; if (cc == 0) { *idx = i; return 1; } }
  }
  *idx = l;
  return 0;
}

/* ---- k-way merge heap: TInputRecord::operator< src/tmerge.h:28-50, TInputFiles::next tmerge.cpp:331-344 */
typedef struct { uint32_t start, end; int fidx; int64_t idx; } head_t;
static int head_less(const head_t* a, const head_t* b) { /* a pops before b */
  if (a->start != b->start) return a->start < b->start;
  if (a->end != b->end) return a->end < b->end;
  return a->fidx < b->fidx;
}
static void heap_push(head_t* h, int* n, head_t v) {
  int i = (*n)++; h[i] = v;
  while (i > 0) { int p = (i - 1) / 2; if (head_less(&h[i], &h[p])) { head_t t = h[i]; h[i] = h[p]; h[p] = t; i = p; } else break; }
}
static head_t heap_pop(head_t* h, int* n) {
  head_t top = h[0]; h[0] = h[--(*n)];
  int i = 0;
  for (;;) {
    int l = 2 * i + 1, r = l + 1, m = i;
    if (l < *n && head_less(&h[l], &h[m])) m = l;
    if (r < *n && head_less(&h[r], &h[m])) m = r;
    if (m == i) break;
    head_t t = h[i]; h[i] = h[m]; h[m] = t; i = m;
  }
  return top;
}

typedef struct {
  const tb_soa_in* in; const tbo_opts* o; cmp_ctx cc;
  rec_t* rec; seg_t* exons;
  group_t** list; int nlist, caplist;
  seglist* fsegs; seglist* rsegs;
  tb_groups_out* out;
} col_state;

static int passes_options(const col_state* S, int64_t i) { /* src/tiebrush.cpp:532-541 */
  uint16_t fl = S->in->flag[i];
  if (!(S->o->keep_bits & TB_KEEP_SUPP) && (fl & 0x800)) return 0;
  if (!(S->o->keep_bits & TB_KEEP_SECONDARY) && (fl & 0x100)) return 0;
  if (!(S->o->keep_bits & TB_KEEP_UNMAP) && (fl & 0x4)) return 0;
  if ((int)S->in->mapq[i] < S->o->min_qual) return 0;
  if ((int)S->in->nh[i] > S->o->max_nh) return 0;
  return 1;
}

static int pair_order(uint16_t fl) { return (fl & 0x40) ? 1 : ((fl & 0x80) ? 2 : 0); } /* GSam.h:315-321 */

static int flush_pdata(col_state* S) { /* src/tiebrush.cpp:501-530 */
  const tb_soa_in* in = S->in;
  for (int gi = 0; gi < S->nlist; gi++) {
    group_t* g = S->list[gi];
    int64_t accYX = g->accYX; int dS = 0;
    for (int s = 0; s < in->n_files; s++) dS += g->samples[s];
    accYX += dS;
    int64_t dmax = g->maxYD;
    const rec_t* r = &S->rec[g->rep];
    char ts = (char)in->strand[g->rep];
    for (int s = 0; s < in->n_files; s++) {
      if (!g->samples[s]) continue;
      if (ts == '+' || ts == '.') { int d = seglist_process(&S->fsegs[s], r->start, S->exons + r->ex_off, r->n_ex); if (d > dmax) dmax = d; }
      if (ts == '-' || ts == '.') { int d = seglist_process(&S->rsegs[s], r->start, S->exons + r->ex_off, r->n_ex); if (d > dmax) dmax = d; }
    }
    tb_groups_out* out = S->out;
    if (out->n_groups >= out->capacity) return 1;
    int64_t k = out->n_groups++;
    out->rep_index[k] = (uint32_t)g->rep;
    out->yc[k] = (float)g->accYC;           /* bam_aux_update_float: double -> float32 */
    out->yx[k] = (uint32_t)accYX;
    out->yd[k] = dmax > 0 ? (int32_t)dmax : 0;
    free(g->samples); free(g);
  }
  S->nlist = 0;
  return 0;
}

static void add_pdata(col_state* S, int64_t i, int fidx) { /* addPData :477-499, settle :378-406, dupAdd :408-436 */
  const tb_soa_in* in = S->in;
  int merged = in->file_merged ? in->file_merged[fidx] : 0;
  int idx = 0;
  if (S->nlist > 0 && list_found(&S->cc, S->list, S->nlist, i, &idx)) {
    group_t* g = S->list[idx];
    if (merged) {
      double v = in->yc_in ? (double)in->yc_in[i] : 0.0; if (v == 0.0) v = 1.0;
      g->accYC += v;
      g->accYX += in->yx_in ? in->yx_in[i] : 1;
      int64_t vyd = in->yd_in ? in->yd_in[i] : 0;
      if (vyd > g->maxYD) g->maxYD = vyd;
    } else {
      if (!S->o->collapse_same || !g->samples[fidx] || pair_order(in->flag[i]) != pair_order(in->flag[g->rep]) ||
          in->qhash == NULL || in->qhash[i] != in->qhash[g->rep]) {
        if (S->o->keep_bits & TB_STORE_FRAC) { int nh = in->nh[i] ? in->nh[i] : 1; g->accYC += 1.0 / nh; }
        else g->accYC += 1.0;
        g->samples[fidx] = 1;
      }
    }
    return;
  }
  /* new group inserted at idx (sortInsert) */
  group_t* g = (group_t*)calloc(1, sizeof(group_t));
  g->rep = i; g->samples = (uint8_t*)calloc((size_t)in->n_files, 1);
  if (merged) {
    g->accYC = in->yc_in ? (double)in->yc_in[i] : 0.0; if (g->accYC == 0.0) g->accYC = 1.0;
    g->accYX = in->yx_in ? in->yx_in[i] : 1;
    g->maxYD = in->yd_in ? in->yd_in[i] : 0;
  } else {
    if (S->o->keep_bits & TB_STORE_FRAC) { int nh = in->nh[i] ? in->nh[i] : 1; g->accYC = 1.0 / nh; }
    else g->accYC = 1.0;
    g->samples[fidx] = 1;
  }
  if (S->nlist == S->caplist) { S->caplist = S->caplist ? 2 * S->caplist : 64; S->list = (group_t**)realloc(S->list, sizeof(group_t*) * (size_t)S->caplist); }
  if (S->nlist == 0) idx = 0;
  memmove(&S->list[idx + 1], &S->list[idx], sizeof(group_t*) * (size_t)(S->nlist - idx));
  S->list[idx] = g; S->nlist++;
}

/* main loop: src/tiebrush.cpp:557-601 for one window (single tid; YD lists start empty) */
int tbo_collapse(const tb_soa_in* in, const tbo_opts* o, tb_groups_out* out) {
  if (in->on_device || out->on_device) return -1;
  int64_t n = in->n; int k = in->n_files;
  col_state S; memset(&S, 0, sizeof(S));
  S.in = in; S.o = o; S.out = out;
  out->n_groups = 0; out->n_kept = 0;
  S.rec = (rec_t*)malloc(sizeof(rec_t) * (size_t)(n > 0 ? n : 1));
  int64_t nex_cap = (int64_t)in->cig_off[n] + n + 1;
  S.exons = (seg_t*)malloc(sizeof(seg_t) * (size_t)nex_cap);
  int64_t eo = 0;
  for (int64_t i = 0; i < n; i++) {
    uint32_t nc = in->cig_off[i + 1] - in->cig_off[i];
    S.rec[i].ex_off = (int)eo;
    if (in->flag[i] & 0x4) { /* setupCoordinates returns early for unmapped: start=end=0, no exons */
      S.rec[i].start = 0; S.rec[i].end = 0; S.rec[i].n_ex = 0; continue;
    }
    S.rec[i].n_ex = setup_coordinates(in->pos[i], in->cigar + in->cig_off[i], nc, S.exons + eo, &S.rec[i].start, &S.rec[i].end);
    eo += S.rec[i].n_ex;
  }
  S.cc.in = in; S.cc.o = o; S.cc.rec = S.rec; S.cc.exons = S.exons;
  S.fsegs = (seglist*)calloc((size_t)(k > 0 ? k : 1), sizeof(seglist));
  S.rsegs = (seglist*)calloc((size_t)(k > 0 ? k : 1), sizeof(seglist));
  for (int f = 0; f < k; f++) { seglist_reset(&S.fsegs[f]); seglist_reset(&S.rsegs[f]); }
  head_t* heap = (head_t*)malloc(sizeof(head_t) * (size_t)(k > 0 ? k : 1)); int hn = 0;
  int64_t* cursor = (int64_t*)malloc(sizeof(int64_t) * (size_t)(k > 0 ? k : 1));
  for (int f = 0; f < k; f++) {
    cursor[f] = in->run_off[f];
    if (cursor[f] < in->run_off[f + 1]) { int64_t i = cursor[f]++; head_t h = { S.rec[i].start, S.rec[i].end, f, i }; heap_push(heap, &hn, h); }
  }
  int64_t prev_pos = -1; int rc = 0;
  while (hn > 0) {
    head_t h = heap_pop(heap, &hn);
    int f = h.fidx;
    if (cursor[f] < in->run_off[f + 1]) { int64_t i = cursor[f]++; head_t nh = { S.rec[i].start, S.rec[i].end, f, i }; heap_push(heap, &hn, nh); }
    if (!passes_options(&S, h.idx)) continue;
    out->n_kept++;
    int64_t pos = (int64_t)S.rec[h.idx].start;
    if (pos != prev_pos) { if (flush_pdata(&S)) { rc = 1; break; } prev_pos = pos; }
    add_pdata(&S, h.idx, f);
  }
  if (!rc && flush_pdata(&S)) rc = 1;
  for (int gi = 0; gi < S.nlist; gi++) { free(S.list[gi]->samples); free(S.list[gi]); }
  for (int f = 0; f < k; f++) { seglist_clear(&S.fsegs[f]); seglist_clear(&S.rsegs[f]); }
  free(S.fsegs); free(S.rsegs); free(heap); free(cursor); free(S.list); free(S.rec); free(S.exons);
  return rc;
}

/* ---- tiecov: src/tiecov.cpp:435-528 -------------------------------------------------------- */
typedef struct { int start, end; char strand; double dupcount; } cjunc;

static int cjunc_less(const cjunc* a, const cjunc* b) { /* CJunc::operator< :73-85 */
  if (a->start == b->start) { if (a->end == b->end) return a->strand < b->strand; return a->end < b->end; }
  return a->start < b->start;
}

typedef struct {
  double* bcov; int64_t bcount, bcap;
  cjunc* juncs; int nj, capj;
  tc_runs_out* runs; tc_juncs_out* jout;
} cov_state;

static void bcov_setcount(cov_state* S, int64_t n) { /* GVec::setCount(n, 0.0): new cells zeroed */
  if (n > S->bcap) { S->bcap = n * 2 + 16; S->bcov = (double*)realloc(S->bcov, sizeof(double) * (size_t)S->bcap); }
  for (int64_t i = S->bcount; i < n; i++) S->bcov[i] = 0.0;
  S->bcount = n;
}

static int flush_coverage(cov_state* S, int tid, int b_start) { /* flushCoverage(FILE*) :226-241 */
  if (tid < 0 || b_start <= 0 || !S->runs) return 0;
  int64_t i = 0; b_start--;
  while (i < S->bcount) {
    double iv = S->bcov[i]; int64_t j = i + 1;
    while (j < S->bcount && iv == S->bcov[j]) j++;
    if (iv != 0.0) {
      tc_runs_out* r = S->runs;
      if (r->n_runs >= r->capacity) return 1;
      int64_t k = r->n_runs++;
      r->tid[k] = tid; r->start0[k] = (int32_t)(b_start + i); r->end0[k] = (int32_t)(b_start + j); r->value[k] = iv;
    }
    i = j;
  }
  return 0;
}

static int flush_juncs(cov_state* S, int tid) { /* flushJuncs :114-120 */
  if (S->jout) for (int i = 0; i < S->nj; i++) {
    tc_juncs_out* o = S->jout;
    if (o->n_juncs >= o->capacity) return 1;
    int64_t k = o->n_juncs++;
    o->tid[k] = tid; o->start[k] = S->juncs[i].start; o->end[k] = S->juncs[i].end;
    o->strand[k] = (uint8_t)S->juncs[i].strand; o->value[k] = S->juncs[i].dupcount;
  }
  S->nj = 0;
  return 0;
}

static void add_junction(cov_state* S, cjunc j) { /* addJunction :100-112 via GArray::AddIfNew (sorted, unique) */
  int lo = 0, hi = S->nj;
  while (lo < hi) { int m = (lo + hi) / 2; if (cjunc_less(&S->juncs[m], &j)) lo = m + 1; else hi = m; }
  if (lo < S->nj && !cjunc_less(&j, &S->juncs[lo])) { S->juncs[lo].dupcount += j.dupcount; return; }
  if (S->nj == S->capj) { S->capj = S->capj ? 2 * S->capj : 64; S->juncs = (cjunc*)realloc(S->juncs, sizeof(cjunc) * (size_t)S->capj); }
  memmove(&S->juncs[lo + 1], &S->juncs[lo], sizeof(cjunc) * (size_t)(S->nj - lo));
  S->juncs[lo] = j; S->nj++;
}

/* returns 0 ok, 1 capacity, 2 unsupported CIGAR op (tiecov.cpp:219-220 aborts); *bad_rec = record */
int tbo_coverage(const tc_soa_in* in, tc_runs_out* runs, tc_juncs_out* jout, int64_t* bad_rec) {
  if (in->on_device) return -1;
  cov_state S; memset(&S, 0, sizeof(S));
  S.runs = runs; S.jout = jout;
  if (runs) runs->n_runs = 0;
  if (jout) jout->n_juncs = 0;
  int prev_tid = -1, b_end = 0, b_start = 0, rc = 0;
  int64_t maxc = 1;
  for (int64_t i = 0; i < in->n; i++) { int64_t c = (int64_t)in->cig_off[i + 1] - in->cig_off[i]; if (c > maxc) maxc = c; }
  seg_t* ex = (seg_t*)malloc(sizeof(seg_t) * (size_t)(maxc + 1));
  for (int64_t i = 0; i < in->n && !rc; i++) {
    const uint32_t* cig = in->cigar + in->cig_off[i]; uint32_t nc = in->cig_off[i + 1] - in->cig_off[i];
    uint32_t start, end; int nex = setup_coordinates(in->pos[i], cig, nc, ex, &start, &end);
    int tid = in->tid[i];
    if (tid != prev_tid || (int)start > b_end) { /* :443 */
      if (prev_tid >= 0) { if (flush_coverage(&S, prev_tid, b_start) || flush_juncs(&S, prev_tid)) { rc = 1; break; } }
      b_start = (int)start; b_end = (int)end;
      S.bcount = 0; bcov_setcount(&S, (int64_t)b_end - b_start + 1);
      prev_tid = tid;
    } else if (b_end < (int)end) { b_end = (int)end; bcov_setcount(&S, (int64_t)b_end - b_start + 1); }
    double acc = (double)in->yc[i];
    if (runs) { /* addCov :194-223 */
      if (nc >= 256) { rc = 2; if (bad_rec) *bad_rec = i; break; } /* uint8_t loop counter never terminates */
      int pos = in->pos[i]; int bs = b_start - 1;
      for (uint32_t c = 0; c < nc && !rc; c++) {
        uint32_t op = cig[c] & 0xf; int len = (int)(cig[c] >> 4);
        switch (op) {
          case 1: case 4: break;
          case 2: case 3: pos += len; break;
          case 0: for (int q = 0; q < len; q++) { S.bcov[pos - bs] += acc; pos++; } break;
          default: rc = 2; if (bad_rec) *bad_rec = i; break;
        }
      }
    }
    if (jout && nex > 1) {
      for (int e = 1; e < nex; e++) { cjunc j = { (int)ex[e - 1].end + 1, (int)ex[e].start - 1, (char)in->strand[i], acc }; add_junction(&S, j); }
    }
  }
  if (!rc && prev_tid >= 0) { if (flush_coverage(&S, prev_tid, b_start) || flush_juncs(&S, prev_tid)) rc = 1; }
  free(ex); free(S.bcov); free(S.juncs);
  return rc;
}


/* ---- tiecov -s (sample heat-map), src/tiecov.cpp:155-185 (addMean), :277-323 (flushCoverage of the pair vector, discretize,
 * normalize) and the main loop :443-499. GROUNDWORK for SURVEY §8f.2: the device path for -s does not exist yet; this
 * restatement is pinned against the compiled reference (tests/test_oracle.py) so that the next round starts from a
 * checked oracle. Per base of a bundle the reference keeps pair<float,uint64_t>{0,1}: for every covering record, in stream
 * order, first += ((int)YX - first) / second (float arithmetic, second converted to float), second++. At the bundle flush:
 * second = ceil(first), then runs of equal `second` (non-zero) are printed as chr, start0, end0, second, hval with
 * hval = ((float)second / n_samples) * (1.5f - 0.1f) + 0.1f. Outputs here: one row per run (tid, start0, end0, ival). */
int tbo_sample_heatmap(const tc_soa_in* in, const int32_t* yx, int64_t capacity, int64_t* n_out, int32_t* o_tid, int32_t* o_start,
                       int32_t* o_end, uint64_t* o_ival, int64_t* bad_rec) {
  if (in->on_device) return -1;
  float* mean = NULL; uint64_t* cnt = NULL; int64_t bcount = 0, bcap = 0;
  int prev_tid = -1, b_end = 0, b_start = 0, rc = 0;
  int64_t nout = 0;
  int64_t maxc = 1;
  for (int64_t i = 0; i < in->n; i++) { int64_t c = (int64_t)in->cig_off[i + 1] - in->cig_off[i]; if (c > maxc) maxc = c; }
  seg_t* ex = (seg_t*)malloc(sizeof(seg_t) * (size_t)(maxc + 1));
#define TBO_FLUSH_SAMPLE()                                                                                              \
  do {                                                                                                                  \
    int64_t x = 0;                                                                                                      \
    while (x < bcount) { /* discretize + flushCoverage(pair vector) */                                                  \
      uint64_t iv = (uint64_t)ceilf(mean[x]);                                                                           \
      int64_t y = x + 1;                                                                                                \
      while (y < bcount && (uint64_t)ceilf(mean[y]) == iv) y++;                                                         \
      if (iv != 0) {                                                                                                    \
        if (nout >= capacity) { rc = 1; break; }                                                                        \
        o_tid[nout] = prev_tid; o_start[nout] = (int32_t)(b_start - 1 + x); o_end[nout] = (int32_t)(b_start - 1 + y);   \
        o_ival[nout] = iv; nout++;                                                                                      \
      }                                                                                                                 \
      x = y;                                                                                                            \
    }                                                                                                                   \
  } while (0)
#define TBO_RESIZE_SAMPLE(newn)                                                                                         \
  do {                                                                                                                  \
    int64_t nn = (newn);                                                                                                \
    if (nn > bcap) { bcap = nn * 2 + 1024; mean = (float*)realloc(mean, sizeof(float) * (size_t)bcap);                  \
                     cnt = (uint64_t*)realloc(cnt, sizeof(uint64_t) * (size_t)bcap); }                                  \
    for (int64_t q = bcount; q < nn; q++) { mean[q] = 0.f; cnt[q] = 1; }                                                 \
    bcount = nn;                                                                                                        \
  } while (0)
  for (int64_t i = 0; i < in->n && !rc; i++) {
    const uint32_t* cig = in->cigar + in->cig_off[i]; uint32_t nc = in->cig_off[i + 1] - in->cig_off[i];
    uint32_t start, end; (void)setup_coordinates(in->pos[i], cig, nc, ex, &start, &end);
    int tid = in->tid[i];
    if (tid != prev_tid || (int)start > b_end) { /* :443 */
      if (prev_tid >= 0) { TBO_FLUSH_SAMPLE(); if (rc) break; }
      b_start = (int)start; b_end = (int)end;
      bcount = 0; TBO_RESIZE_SAMPLE((int64_t)b_end - b_start + 1);
      prev_tid = tid;
    } else if (b_end < (int)end) { b_end = (int)end; TBO_RESIZE_SAMPLE((int64_t)b_end - b_start + 1); }
    if (nc >= 256) { rc = 2; if (bad_rec) *bad_rec = i; break; } /* uint8_t loop counter, :158 */
    int val = yx[i];
    int pos = in->pos[i]; int bs = b_start - 1;
    for (uint32_t c = 0; c < nc && !rc; c++) { /* addMean :155-185 */
      uint32_t op = cig[c] & 0xf; int len = (int)(cig[c] >> 4);
      switch (op) {
        case 1: case 4: break;
        case 2: case 3: pos += len; break;
        case 0:
          for (int q = 0; q < len; q++) {
            int64_t x = pos - bs;
            mean[x] += ((float)val - mean[x]) / (float)cnt[x];   /* (val - first): int converted to float, / uint64 converted to float */
            cnt[x]++;
            pos++;
          }
          break;
        default: rc = 2; if (bad_rec) *bad_rec = i; break;
      }
    }
  }
  if (!rc && prev_tid >= 0) TBO_FLUSH_SAMPLE();
#undef TBO_FLUSH_SAMPLE
#undef TBO_RESIZE_SAMPLE
  free(ex); free(mean); free(cnt);
  *n_out = nout;
  return rc;
}
