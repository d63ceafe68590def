// collapse.cu — tiebrush's hot path on the device: k-way merge of the per-sample sorted runs, grouping of
// duplicate alignments per start position, YC / YX and the representative (reference src/tmerge.cpp:331-344,
// src/tmerge.h:28-50, src/tiebrush.cpp:275-345, 350-530, 532-541).
//
// Data flow for one file-major window of n records in k sorted runs:
//
//   C1 histogram   cnt[pos-pos_lo]++                                  (one u32 RED per record)
//   C2 scan        P[p] = #records with pos < p                       (merged-order rank of every position)
//   C3 slots       slot m owns the positions holding merged ranks [mT,(m+1)T): cut only BETWEEN positions,
//                  a position holding >= 2 multiples of T becomes a slot of its own (pile-ups)
//   C4 run offsets off[m][f] = first record of run f that falls in slot >= m   (merge-path style partition,
//                  written by the records that sit on a slot boundary)
//   C5 tile kernel one CTA per slot STREAMS its k run slices (file-major, so the per-file running max of
//                  `end` that defines the reference's merge order is a segmented scan), builds the mode key
//                  from the packed CIGAR, and groups records in a SHARED-MEMORY hash table whose slots are
//                  claimed by CAS with (tag32|owner index): a follower compares its real key bytes with the
//                  owner's, so grouping never rests on the hash. Per group: count, argmin of (E,fidx,ord) =
//                  representative, per-sample bitset. Groups are then bitonic-sorted with the reference's
//                  comparator (memcmp order of the little-endian CIGAR words etc.) and written in final order.
//   C6 compaction  per-slot group counts -> scan -> dense output
//   C7 YD          see yd section below (GSegList state machine, tiebrush.cpp:122-253)
#include "tb_common.cuh"

namespace {

enum {  // workspace slots in ctx->buf (shared numbering space with coverage.cu is fine: calls do not overlap)
  XB_HIST = 0,   // u32 [S+2]  counts, then exclusive scan P
  XB_AGG,        // scan aggregates
  XB_STATUS,     // i64 [16]
  XB_SLOTPOS,    // u32 [M+1]  p_m
  XB_OFF,        // u32 [(M+1)*k]
  XB_RUNOFF,     // i64 [k+1]
  XB_MERGED,     // u8  [k]
  XB_GCOUNT,     // u32 [M+1]
  XB_GBASE,      // u32 [M+1]
  XB_ST_REP,     // u32 [n] staged
  XB_ST_YC,      // f32 [n]
  XB_ST_YX,      // u32 [n]
  XB_ST_BITS,    // u32 [n*W]
  XB_BITS,       // u32 [G*W] compacted
  XB_GSTART,     // i32 [G]
  XB_GEND,       // i32 [G]
  XB_GKEY,       // u64 [G]
  XB_GPM,        // u64 [G]
  XB_BHEAD,      // u32 [G]  bundle start indices
  XB_YD,         // i32 [G]
  XB_COUNT_
};
enum { CS_ERR = 0, CS_ERRIDX, CS_NKEPT, CS_NGROUPS, CS_TABLE_OVERFLOW, CS_NBUNDLES, CS_YD_OVERFLOW, CS_N_ };
enum { ERR_POS_RANGE = 1, ERR_UNSORTED = 2 };

struct ColIn {
  int64_t n; int k; int mode; uint32_t flag_mask; int max_nh; int min_qual; int keep_bits;
  const int32_t* pos; const uint16_t* flag; const uint8_t* mapq; const uint8_t* strand; const uint16_t* nh;
  const uint32_t* cig_off; const uint32_t* cigar; const uint32_t* md_off; const uint8_t* md;
  int32_t pos_lo; uint32_t span;
};

// ---------------------------------------------------------------------------------------------------
// C1: histogram of start positions
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) col_hist_kernel(ColIn in, uint32_t* __restrict__ cnt, long long* __restrict__ status) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= in.n) return;
  int64_t rel = (int64_t)in.pos[i] - in.pos_lo;
  if (rel < 0 || rel >= (int64_t)in.span) { status[CS_ERR] = ERR_POS_RANGE; atomicMin((unsigned long long*)&status[CS_ERRIDX], (unsigned long long)i); return; }
  atomicAdd(&cnt[rel], 1u);
}

struct HistIn { const uint32_t* c; __device__ uint32_t operator()(int64_t i) const { return c[i]; } };
struct HistOut { uint32_t* p; __device__ void operator()(int64_t i, uint32_t exc, uint32_t) const { p[i] = exc; } };

// ---------------------------------------------------------------------------------------------------
// C3: slot geometry
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) col_slotpos_kernel(const uint32_t* __restrict__ P, uint32_t span, uint32_t n, uint32_t T, uint32_t M,
                                                          uint32_t* __restrict__ slotpos) {
  uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m > M) return;
  if (m == M) { slotpos[M] = span; return; }
  uint32_t target = m * T;  // merged rank; target < n
  // first p' in [0,span] with P[p'] > target  (P[span] = n > target)
  uint32_t lo = 0, hi = span;
  while (lo < hi) { uint32_t mid = lo + ((hi - lo) >> 1); if (P[mid] > target) hi = mid; else lo = mid + 1; }
  slotpos[m] = lo - 1;
}

__device__ __forceinline__ uint32_t slot_lo(const uint32_t* slotpos, uint32_t m, uint32_t M, uint32_t span) {
  if (m >= M) return span;
  uint32_t p = slotpos[m];
  return (m == 0 || slotpos[m - 1] != p) ? p : p + 1;
}
// slot of a non-empty position, from its merged-rank interval [a,b)
__device__ __forceinline__ uint32_t slot_of(uint32_t a, uint32_t b, uint32_t T) {
  uint32_t mlo = (a + T - 1) / T;
  return ((unsigned long long)mlo * T < b) ? mlo : mlo - 1;
}

// ---------------------------------------------------------------------------------------------------
// C4: per-run offsets of every slot
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) col_off_init_kernel(uint32_t* __restrict__ off, const long long* __restrict__ run_off, int k, uint64_t total) {
  uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= total) return;
  int f = (int)(x % (uint64_t)k);
  off[x] = (uint32_t)run_off[f + 1];
}

__global__ void __launch_bounds__(256) col_off_kernel(ColIn in, const long long* __restrict__ run_off, const uint32_t* __restrict__ P, uint32_t T,
                                                      uint32_t* __restrict__ off, long long* __restrict__ status) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= in.n) return;
  int lo = 0, hi = in.k;  // run containing i: last f with run_off[f] <= i
  while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (run_off[mid] <= i) lo = mid; else hi = mid; }
  int f = lo;
  uint32_t rel = (uint32_t)(in.pos[i] - in.pos_lo);
  if (rel >= in.span) return;  // already reported by C1
  uint32_t scur = slot_of(P[rel], P[rel + 1], T);
  int64_t sprev = -1;
  if (i > run_off[f]) {
    int32_t pp = in.pos[i - 1];
    if (pp > in.pos[i]) { status[CS_ERR] = ERR_UNSORTED; atomicMin((unsigned long long*)&status[CS_ERRIDX], (unsigned long long)i); return; }
    uint32_t prel = (uint32_t)(pp - in.pos_lo);
    if (prel >= in.span) return;
    if (prel == rel) return;
    sprev = slot_of(P[prel], P[prel + 1], T);
  }
  for (int64_t s = sprev + 1; s <= (int64_t)scur; ++s) off[(uint64_t)s * in.k + f] = (uint32_t)i;
}

// ---------------------------------------------------------------------------------------------------
// record model on the device: setupCoordinates (GSam.cpp:351-417) and the mode keys (tiebrush.cpp:275-345)
// ---------------------------------------------------------------------------------------------------
struct ExonIter {  // yields the exon chain of one record
  const uint32_t* cig; uint32_t c, c1; int pos, l, exstart; bool intron, ins, done;
  __device__ void init(const uint32_t* cigar, uint32_t c0, uint32_t cend, int p) { cig = cigar; c = c0; c1 = cend; pos = p; l = 0; exstart = p; intron = ins = false; done = false; }
  __device__ bool next(int& s, int& e) {
    if (done) return false;
    while (c < c1) {
      uint32_t w = cig[c++]; uint32_t op = w & 0xf; int len = (int)(w >> 4);
      switch (op) {
        case TB_OP_M: case TB_OP_EQ: case TB_OP_X: case TB_OP_D: l += len; intron = false; ins = false; break;
        case TB_OP_N: {
          bool emit = (!ins || !intron);
          int es = exstart + 1, ee = pos + l;
          l += len; exstart = pos + l; intron = true;
          if (emit) { s = es; e = ee; return true; }
          break;
        }
        case TB_OP_S: case TB_OP_H: intron = false; ins = false; break;
        case TB_OP_I: ins = true; break;
        default: break;
      }
    }
    done = true; s = exstart + 1; e = pos + l; return true;
  }
};

__device__ __forceinline__ uint64_t fold(uint64_t h, uint64_t w) { h = (h ^ w) * 0x9E3779B97F4A7C15ULL; return h ^ (h >> 29); }
__device__ __forceinline__ unsigned col_strand_code(uint8_t c) { return c == '+' ? 0u : (c == '-' ? 1u : 2u); }

// clipped range for -P: strip leading then trailing soft clips (cmpCigarClip :321-328)
__device__ __forceinline__ void clip_range(const uint32_t* cig, uint32_t& a, uint32_t& b) {
  while (a < b && (cig[a] & 0xf) == TB_OP_S) ++a;
  while (b > a && (cig[b - 1] & 0xf) == TB_OP_S) --b;
}

// one pass over the record's CIGAR: reference length and the mode-key hash
__device__ __forceinline__ void parse_record(const ColIn& in, int64_t i, int pos, int& reflen, uint64_t& khash) {
  uint32_t c0 = in.cig_off[i], c1 = in.cig_off[i + 1];
  int l = 0;
  uint64_t h;
  if (in.mode == TB_MODE_EXON) {
    ExonIter it; it.init(in.cigar, c0, c1, pos);
    int s, e, nex = 0; h = 0x1234567ULL;
    while (it.next(s, e)) { h = fold(h, ((uint64_t)(uint32_t)s << 32) | (uint32_t)e); ++nex; }
    h = fold(h, (uint64_t)nex);
    l = it.l;
  } else {
    uint32_t a = c0, b = c1;
    if (in.mode == TB_MODE_CLIP) clip_range(in.cigar, a, b);
    h = fold(0x9876543ULL, (uint64_t)(b - a));
    for (uint32_t c = c0; c < c1; ++c) {
      uint32_t w = in.cigar[c]; uint32_t op = w & 0xf;
      if (op == TB_OP_M || op == TB_OP_D || op == TB_OP_N || op == TB_OP_EQ || op == TB_OP_X) l += (int)(w >> 4);
      if (c >= a && c < b) h = fold(h, w);
    }
    if (in.mode == TB_MODE_FULL) {
      uint32_t m0 = in.md_off[i], m1 = in.md_off[i + 1];
      h = fold(h, (uint64_t)(m1 > m0));
      for (uint32_t q = m0; q < m1; ++q) { uint8_t ch = in.md[q]; if (ch == 0) break; h = fold(h, ch); }
    }
  }
  reflen = l; khash = h;
}

// exact mode comparison of two records, sign as in the reference's cmp* functions
__device__ int mode_cmp(const ColIn& in, uint32_t ia, uint32_t ib) {
  uint32_t a0 = in.cig_off[ia], a1 = in.cig_off[ia + 1], b0 = in.cig_off[ib], b1 = in.cig_off[ib + 1];
  if (in.mode == TB_MODE_EXON) {
    ExonIter x, y; x.init(in.cigar, a0, a1, in.pos[ia]); y.init(in.cigar, b0, b1, in.pos[ib]);
    // exon counts first (cmpExons :337)
    int na = 0, nb = 0, s, e;
    { ExonIter t = x; while (t.next(s, e)) ++na; }
    { ExonIter t = y; while (t.next(s, e)) ++nb; }
    if (na != nb) return na - nb;
    int sa, ea, sb, eb;
    while (x.next(sa, ea)) { y.next(sb, eb); if (sa != sb) return sa - sb; if (ea != eb) return ea - eb; }
    return 0;
  }
  if (in.mode == TB_MODE_CLIP) { clip_range(in.cigar, a0, a1); clip_range(in.cigar, b0, b1); }
  int na = (int)(a1 - a0), nb = (int)(b1 - b0);
  if (na != nb) return na - nb;
  for (int q = 0; q < na; ++q) {
    uint32_t wa = in.cigar[a0 + q], wb = in.cigar[b0 + q];
    if (wa != wb) {  // memcmp over little-endian bytes == numeric order of the byte-reversed words
      uint32_t ra = __byte_perm(wa, 0, 0x0123), rb = __byte_perm(wb, 0, 0x0123);
      return ra < rb ? -1 : 1;
    }
  }
  if (in.mode == TB_MODE_FULL) {
    uint32_t ma = in.md_off[ia], mae = in.md_off[ia + 1], mb = in.md_off[ib], mbe = in.md_off[ib + 1];
    bool pa = mae > ma, pb = mbe > mb;
    if (!pa || !pb) { if (pa == pb) return 0; return pa ? 1 : -1; }
    for (;; ++ma, ++mb) {  // strcmp
      uint8_t ca = ma < mae ? in.md[ma] : 0, cb = mb < mbe ? in.md[mb] : 0;
      if (ca != cb) return (int)ca - (int)cb;
      if (ca == 0) return 0;
    }
  }
  return 0;
}

__device__ __forceinline__ bool passes_options(const ColIn& in, uint16_t fl, uint8_t mq, uint16_t nh) {  // tiebrush.cpp:532-541
  if (!(in.keep_bits & TB_KEEP_SUPP) && (fl & 0x800)) return false;
  if (!(in.keep_bits & TB_KEEP_SECONDARY) && (fl & 0x100)) return false;
  if (!(in.keep_bits & TB_KEEP_UNMAP) && (fl & 0x4)) return false;
  if ((int)mq < in.min_qual) return false;
  if ((int)nh > in.max_nh) return false;
  return true;
}

// ---------------------------------------------------------------------------------------------------
// C5: the tile kernel
// ---------------------------------------------------------------------------------------------------
constexpr int TILE_THREADS = 512;
constexpr int TILE_WARPS = TILE_THREADS / 32;
constexpr uint32_t EMPTY32 = 0xffffffffu;
constexpr unsigned long long EMPTY64 = ~0ULL;

struct TileParams {
  uint32_t T, M, gcap, W;            // gcap power of two
  const uint32_t* P; const uint32_t* slotpos; const uint32_t* off;
  uint32_t* gcount; uint32_t* st_rep; float* st_yc; uint32_t* st_yx; uint32_t* st_bits;
  long long* status; uint64_t seed;
};

static size_t tile_smem_bytes_host(uint32_t k, uint32_t gcap, uint32_t W) {
  size_t b = 0;
  b += sizeof(uint64_t) * gcap * 3;      // slotword, rep, K1
  b += sizeof(uint32_t) * gcap;          // cnt
  b += sizeof(uint32_t) * (size_t)gcap * W;  // bitsets
  b += sizeof(uint16_t) * gcap;          // sort index
  b += sizeof(uint32_t) * (2 * (size_t)k + 2);  // s_lo, s_off
  b += 256;
  return b;
}

__global__ void __launch_bounds__(TILE_THREADS, 1) col_tile_kernel(ColIn in, TileParams tp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t gcap = tp.gcap, W = tp.W, k = (uint32_t)in.k;
  unsigned long long* g_word = (unsigned long long*)smem_raw;       // tag32<<32 | owner
  unsigned long long* g_rep = g_word + gcap;                       // min over members of (E-pos)<<32 | index
  unsigned long long* g_k1 = g_rep + gcap;                         // pos<<33 | strand<<31 | (end-pos)
  uint32_t* g_cnt = (uint32_t*)(g_k1 + gcap);
  uint32_t* g_bits = g_cnt + gcap;
  uint32_t* s_lo = g_bits + (size_t)gcap * W;
  uint32_t* s_off = s_lo + k;                                       // k+1 entries
  uint16_t* s_idx = (uint16_t*)(s_off + k + 2);
  __shared__ int s_wv[TILE_WARPS];
  __shared__ int s_wh[TILE_WARPS];
  __shared__ int s_carry;
  __shared__ uint32_t s_scan[33];
  __shared__ uint32_t s_ngroups, s_nkept, s_overflow;

  const uint32_t m = blockIdx.x;
  const uint32_t tid = threadIdx.x;
  const uint32_t blo = slot_lo(tp.slotpos, m, tp.M, in.span), bhi = slot_lo(tp.slotpos, m + 1, tp.M, in.span);
  const uint32_t rank0 = tp.P[blo];
  const uint32_t n_t = tp.P[bhi] - rank0;
  if (n_t == 0) { if (tid == 0) tp.gcount[m] = 0; return; }

  // ---- per-run slices of this slot, exclusive scan of their lengths ----
  if (tid == 0) { s_ngroups = 0; s_nkept = 0; s_overflow = 0; s_carry = 0; }
  {
    uint32_t carry = 0;
    for (uint32_t base = 0; base < k; base += TILE_THREADS) {
      uint32_t f = base + tid;
      uint32_t lo = 0, c = 0;
      if (f < k) { lo = tp.off[(uint64_t)m * k + f]; c = tp.off[(uint64_t)(m + 1) * k + f] - lo; s_lo[f] = lo; }
      uint32_t tot;
      uint32_t exc = tb_block_exscan<OpSumU32>(c, s_scan, &tot);
      if (f < k) s_off[f] = carry + exc;
      carry += tot;
    }
    if (tid == 0) s_off[k] = carry;
  }
  for (uint32_t s = tid; s < gcap; s += TILE_THREADS) { g_word[s] = EMPTY64; g_rep[s] = EMPTY64; g_cnt[s] = 0; }
  for (uint32_t s = tid; s < gcap * W; s += TILE_THREADS) g_bits[s] = 0;
  __syncthreads();

  // ---- stream the records ----
  uint32_t my_kept = 0;
  for (uint32_t base = 0; base < n_t; base += TILE_THREADS) {
    const uint32_t j = base + tid;
    const bool valid = j < n_t;
    uint32_t f = 0, i = 0; int pos = 0, reflen = 0; uint64_t kh = 0; bool head = true, pass = false; uint8_t sc = 0;
    if (valid) {
      uint32_t lo = 0, hi = k;  // last f with s_off[f] <= j
      while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (s_off[mid] <= j) lo = mid; else hi = mid; }
      f = lo;
      i = s_lo[f] + (j - s_off[f]);
      pos = in.pos[i];
      head = (j == s_off[f]) || (in.pos[i - 1] != pos);
      parse_record(in, i, pos, reflen, kh);
      pass = passes_options(in, in.flag[i], in.mapq[i], in.nh[i]);
      sc = (uint8_t)col_strand_code(in.strand[i]);
    }
    // segmented inclusive max-scan of `end - pos_lo` over (file,pos) segments == the running max that orders the PQ
    int v = valid ? (pos - in.pos_lo) + reflen : 0;
    int h = head ? 1 : 0;
    const int carry_in = s_carry;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int v2 = __shfl_up_sync(0xffffffffu, v, d), h2 = __shfl_up_sync(0xffffffffu, h, d);
      if (tb_lane() >= d) { if (!h) v = max(v, v2); h |= h2; }
    }
    if (tb_lane() == 31) { s_wv[tb_warp()] = v; s_wh[tb_warp()] = h; }
    __syncthreads();
    int E = v;
    if (!h) {
      int acc = 0; bool closed = false;
      for (int w = tb_warp() - 1; w >= 0; --w) { acc = max(acc, s_wv[w]); if (s_wh[w]) { closed = true; break; } }
      if (!closed) acc = max(acc, carry_in);
      E = max(v, acc);
    }
    __syncthreads();
    if (tid == TILE_THREADS - 1) s_carry = E;

    if (valid && pass) {
      ++my_kept;
      const uint32_t erel = (uint32_t)(reflen);  // end - pos
      const unsigned long long K1 = ((unsigned long long)(uint32_t)pos << 33) | ((unsigned long long)sc << 31) | erel;
      const unsigned long long hh = tb_mix64(K1 ^ tb_mix64(kh + tp.seed));
      const uint32_t tag = (uint32_t)(hh >> 32);
      uint32_t s = (uint32_t)hh & (gcap - 1);
      const unsigned long long mine = ((unsigned long long)tag << 32) | i;
      const unsigned long long repkey = ((unsigned long long)(uint32_t)(E - (pos - in.pos_lo)) << 32) | i;
      bool placed = false;
      for (uint32_t probe = 0; probe < gcap; ++probe) {
        unsigned long long old = atomicCAS(&g_word[s], EMPTY64, mine);
        bool match = false;
        if (old == EMPTY64) {
          g_k1[s] = K1;
          atomicAdd(&s_ngroups, 1u);
          match = true;
        } else if ((uint32_t)(old >> 32) == tag) {
          uint32_t o = (uint32_t)old;  // owner: compare the real keys
          match = (in.pos[o] == pos) && (col_strand_code(in.strand[o]) == sc) && (mode_cmp(in, i, o) == 0);
        }
        if (match) {
          atomicAdd(&g_cnt[s], 1u);
          atomicMin(&g_rep[s], repkey);
          atomicOr(&g_bits[(size_t)s * W + (f >> 5)], 1u << (f & 31));
          placed = true;
          break;
        }
        s = (s + 1) & (gcap - 1);
      }
      if (!placed) s_overflow = 1;
    }
    __syncthreads();
    if (s_ngroups > gcap - (gcap >> 3)) break;  // uniform: table (nearly) full
  }
  if (my_kept) atomicAdd(&s_nkept, my_kept);
  __syncthreads();
  const uint32_t G = s_ngroups;
  if (s_overflow || G > gcap - (gcap >> 3)) {
    if (tid == 0) { tp.status[CS_TABLE_OVERFLOW] = 1; tp.gcount[m] = 0; }
    return;
  }
  if (tid == 0 && s_nkept) atomicAdd((unsigned long long*)&tp.status[CS_NKEPT], (unsigned long long)s_nkept);

  // ---- compact occupied slots into s_idx, pad to a power of two ----
  uint32_t P2 = 1; while (P2 < G) P2 <<= 1;
  {
    uint32_t carry = 0;
    for (uint32_t base = 0; base < gcap; base += TILE_THREADS) {
      uint32_t s = base + tid;
      uint32_t occ = (s < gcap && g_word[s] != EMPTY64) ? 1u : 0u;
      uint32_t tot;
      uint32_t exc = tb_block_exscan<OpSumU32>(occ, s_scan, &tot);
      if (occ) s_idx[carry + exc] = (uint16_t)s;
      carry += tot;
    }
    for (uint32_t r = G + tid; r < P2; r += TILE_THREADS) s_idx[r] = 0xffff;
  }
  __syncthreads();

  // ---- bitonic sort of the groups by the reference's SPData::operator< (tiebrush.cpp:438-457) ----
  for (uint32_t kk = 2; kk <= P2; kk <<= 1) {
    for (uint32_t jj = kk >> 1; jj > 0; jj >>= 1) {
      for (uint32_t t = tid; t < P2; t += TILE_THREADS) {
        uint32_t p = t ^ jj;
        if (p > t) {
          uint16_t a = s_idx[t], b = s_idx[p];
          bool a_gt_b;  // is a strictly after b?
          if (a == 0xffff) a_gt_b = (b != 0xffff);
          else if (b == 0xffff) a_gt_b = false;
          else {
            unsigned long long ka = g_k1[a], kb = g_k1[b];
            if (ka != kb) a_gt_b = ka > kb;
            else a_gt_b = mode_cmp(in, (uint32_t)g_word[a], (uint32_t)g_word[b]) > 0;
          }
          bool up = ((t & kk) == 0);
          if (a_gt_b == up) { s_idx[t] = b; s_idx[p] = a; }
        }
      }
      __syncthreads();
    }
  }

  // ---- write the groups in final order (staged at the slot's merged rank) ----
  for (uint32_t r = tid; r < G; r += TILE_THREADS) {
    uint32_t s = s_idx[r];
    uint64_t o = (uint64_t)rank0 + r;
    tp.st_rep[o] = (uint32_t)g_rep[s];
    tp.st_yc[o] = (float)g_cnt[s];
    uint32_t yx = 0;
    for (uint32_t w = 0; w < W; ++w) { uint32_t b = g_bits[(size_t)s * W + w]; yx += __popc(b); tp.st_bits[o * W + w] = b; }
    tp.st_yx[o] = yx;
  }
  if (tid == 0) tp.gcount[m] = G;
}

// ---------------------------------------------------------------------------------------------------
// C6: compaction of the staged groups
// ---------------------------------------------------------------------------------------------------
struct GcIn { const uint32_t* c; __device__ uint32_t operator()(int64_t i) const { return c[i]; } };
struct GcOut { uint32_t* b; __device__ void operator()(int64_t i, uint32_t exc, uint32_t) const { b[i] = exc; } };

__global__ void __launch_bounds__(128) col_compact_kernel(TileParams tp, uint32_t span, const uint32_t* __restrict__ gbase, uint32_t* __restrict__ o_rep,
                                                          float* __restrict__ o_yc, uint32_t* __restrict__ o_yx, uint32_t* __restrict__ o_bits, int64_t capacity) {
  uint32_t m = blockIdx.x;
  uint32_t G = tp.gcount[m];
  if (G == 0) return;
  uint64_t src = tp.P[slot_lo(tp.slotpos, m, tp.M, span)];
  uint64_t dst = gbase[m];
  for (uint32_t r = threadIdx.x; r < G; r += blockDim.x) {
    if ((int64_t)(dst + r) >= capacity) break;
    o_rep[dst + r] = tp.st_rep[src + r];
    o_yc[dst + r] = tp.st_yc[src + r];
    o_yx[dst + r] = tp.st_yx[src + r];
  }
  for (uint64_t x = threadIdx.x; x < (uint64_t)G * tp.W; x += blockDim.x) o_bits[dst * tp.W + x] = tp.st_bits[src * tp.W + x];
}

__global__ void col_store_total_kernel(const uint32_t* tot, long long* status) { status[CS_NGROUPS] = *tot; }

// ---------------------------------------------------------------------------------------------------
// C7: YD — GSegList::processRead / mergeRead (tiebrush.cpp:151-250) per (sample, strand list).
// Groups are cut into global bundles (a group starts a bundle iff its start exceeds every earlier end):
// at such a gap every list is provably emptied by the next processRead (d==0 => clearTo(prev) with prev the
// last node), so bundles are independent. One thread owns one (bundle, sample) pair and walks the bundle's
// groups in output order, keeping the forward and reverse lists as small sorted arrays. Nodes lying before
// `prev` can never be touched again before they are freed (disjoint, sorted, all ends < r.start), so they are
// dropped eagerly; results are unchanged.
// ---------------------------------------------------------------------------------------------------
constexpr int YD_CAP = 64;      // live nodes per list (shared memory); more => CS_YD_OVERFLOW (fails loudly)
constexpr int YD_MAXEX = 16;    // exons per read staged in shared memory; longer chains are walked from global memory
struct SegArr {
  int n; uint32_t last_pos; int last_dist;
  uint32_t* st; uint32_t* en;   // YD_CAP entries each (shared memory)
  __device__ void reset() { n = 0; last_pos = 0; last_dist = -1; }
  __device__ void erase(int a, int b) {  // remove [a,b)
    int d = b - a; if (d <= 0) return;
    for (int q = b; q < n; ++q) { st[q - d] = st[q]; en[q - d] = en[q]; }
    n -= d;
  }
  __device__ bool insert(int at, uint32_t s, uint32_t e) {
    if (n >= YD_CAP) return false;
    for (int q = n; q > at; --q) { st[q] = st[q - 1]; en[q] = en[q - 1]; }
    st[at] = s; en[at] = e; ++n; return true;
  }
  // mergeRead :167-219 over an exon source; returns false on capacity overflow
  template <class Src>
  __device__ bool merge(Src& src) {
    int s, e;
    if (n == 0) {
      while (src.next(s, e)) { if (n >= YD_CAP) return false; st[n] = (uint32_t)s; en[n] = (uint32_t)e; ++n; }
      return true;
    }
    int cur = 0;
    while (src.next(s, e)) {
      uint32_t es = (uint32_t)s, ee = (uint32_t)e;
      while (cur < n) {
        if (ee < st[cur]) { if (!insert(cur, es, ee)) return false; ++cur; break; }  // inserted before n; n itself is now at cur
        if (es <= en[cur]) {
          if (es < st[cur]) st[cur] = es;
          if (ee > en[cur]) en[cur] = ee;
          int nx = cur + 1;
          while (nx < n && st[nx] <= en[cur]) {
            uint32_t nend = en[nx];
            erase(nx, nx + 1);
            if (nend > en[cur]) { en[cur] = nend; break; }
          }
          break;
        }
        ++cur;
      }
      if (cur >= n) return true;  // cursor ran off the list: this and all later exons are dropped (reference behaviour)
    }
    return true;
  }
  // processRead :221-250
  template <class Src>
  __device__ int process(Src& src, uint32_t rstart, bool& ok) {
    if (last_pos == rstart) { ok = merge(src) && ok; return last_dist; }
    int d = 0, prev = -1;
    for (int q = 0; q < n && st[q] < rstart; ++q) prev = q;
    if (prev >= 0) {
      if (en[prev] >= rstart) d = (int)(rstart - st[prev]);
      if (d == 0) erase(0, prev + 1);
      else erase(0, prev);  // eager drop of the inert nodes before prev
    }
    last_pos = rstart; last_dist = d;
    ok = merge(src) && ok;
    return d;
  }
};
struct SmemExons {  // exon source over a staged array
  const int2* ex; int n, i;
  __device__ bool next(int& s, int& e) { if (i >= n) return false; s = ex[i].x; e = ex[i].y; ++i; return true; }
};

// Y1: per-group start / end / strand of the representative
__global__ void __launch_bounds__(256) yd_group_info_kernel(ColIn in, const uint32_t* __restrict__ rep, int64_t G, int32_t* __restrict__ gstart,
                                                           int32_t* __restrict__ gend, uint8_t* __restrict__ gstrand) {
  int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  uint32_t r = rep[g];
  int pos = in.pos[r], l = 0;
  for (uint32_t c = in.cig_off[r]; c < in.cig_off[r + 1]; ++c) {
    uint32_t w = in.cigar[c]; uint32_t op = w & 0xf;
    if (op == TB_OP_M || op == TB_OP_D || op == TB_OP_N || op == TB_OP_EQ || op == TB_OP_X) l += (int)(w >> 4);
  }
  gstart[g] = pos + 1;
  gend[g] = pos + l;
  gstrand[g] = in.strand[r];
}

// A chain = (sample s, strand list): the groups, in output order, that contain s and whose strand feeds the list
// ('+','.' -> forward = 0; '-','.' -> reverse = 1). processRead is a sequential state machine along a chain and
// chains are independent (tiebrush.cpp:512-521), so the device builds every chain's member list with a stable
// counting scatter (Y2-Y4) and then walks each chain with one warp (Y5).
constexpr int YD_BLOCK = 1024;

// Y2: members per (block of groups, chain)
__global__ void __launch_bounds__(128) yd_count_kernel(const uint32_t* __restrict__ bits, uint32_t W, int k, int64_t G, const uint8_t* __restrict__ gstrand,
                                                      uint32_t* __restrict__ blkcnt, int64_t nblk) {
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= nblk * W) return;
  const int64_t b = warp / W; const uint32_t w = (uint32_t)(warp % W);
  const int lane = tb_lane();
  const int64_t g0 = b * YD_BLOCK, g1 = (g0 + YD_BLOCK < G) ? g0 + YD_BLOCK : G;
  uint32_t cf = 0, cr = 0;
  for (int64_t base = g0; base < g1; base += 32) {
    int64_t g = base + lane;
    uint32_t word = 0; uint8_t sc = 0;
    if (g < g1) { word = bits[(uint64_t)g * W + w]; sc = gstrand[g]; }
    unsigned any = __ballot_sync(0xffffffffu, word != 0);
    while (any) {
      int q = __ffs(any) - 1; any &= any - 1;
      uint32_t wq = __shfl_sync(0xffffffffu, word, q);
      int sq = __shfl_sync(0xffffffffu, (int)sc, q);
      if ((wq >> lane) & 1u) { cf += (sq != '-'); cr += (sq != '+'); }
    }
  }
  const int s = (int)(w * 32 + lane);
  if (s < k) { blkcnt[(b * 2 + 0) * k + s] = cf; blkcnt[(b * 2 + 1) * k + s] = cr; }
}

// Y3: exclusive prefix over blocks per chain column (in place), then chain base offsets
__global__ void __launch_bounds__(128) yd_prefix_kernel(uint32_t* __restrict__ blkcnt, int64_t nblk, int k, uint32_t* __restrict__ coltot) {
  int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= 2 * k) return;
  uint32_t run = 0;
  for (int64_t b = 0; b < nblk; ++b) {
    uint32_t* p = &blkcnt[b * 2 * k + col];
    uint32_t v = *p; *p = run; run += v;
  }
  coltot[col] = run;
}
__global__ void yd_colbase_kernel(const uint32_t* __restrict__ coltot, int k, unsigned long long* __restrict__ colbase) {
  unsigned long long run = 0;
  for (int c = 0; c < 2 * k; ++c) { colbase[c] = run; run += coltot[c]; }
  colbase[2 * k] = run;
}

// Y4: stable scatter of the group ids into the chain lists
__global__ void __launch_bounds__(128) yd_scatter_kernel(const uint32_t* __restrict__ bits, uint32_t W, int k, int64_t G, const uint8_t* __restrict__ gstrand,
                                                        const uint32_t* __restrict__ blkoff, const unsigned long long* __restrict__ colbase, int64_t nblk,
                                                        uint32_t* __restrict__ chain) {
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= nblk * W) return;
  const int64_t b = warp / W; const uint32_t w = (uint32_t)(warp % W);
  const int lane = tb_lane();
  const int s = (int)(w * 32 + lane);
  const int64_t g0 = b * YD_BLOCK, g1 = (g0 + YD_BLOCK < G) ? g0 + YD_BLOCK : G;
  unsigned long long pf = 0, pr = 0;
  if (s < k) { pf = colbase[s] + blkoff[(b * 2 + 0) * k + s]; pr = colbase[k + s] + blkoff[(b * 2 + 1) * k + s]; }
  for (int64_t base = g0; base < g1; base += 32) {
    int64_t g = base + lane;
    uint32_t word = 0; uint8_t sc = 0;
    if (g < g1) { word = bits[(uint64_t)g * W + w]; sc = gstrand[g]; }
    unsigned any = __ballot_sync(0xffffffffu, word != 0);
    while (any) {
      int q = __ffs(any) - 1; any &= any - 1;
      uint32_t wq = __shfl_sync(0xffffffffu, word, q);
      int sq = __shfl_sync(0xffffffffu, (int)sc, q);
      if ((wq >> lane) & 1u) {
        if (sq != '-') chain[pf++] = (uint32_t)(base + q);
        if (sq != '+') chain[pr++] = (uint32_t)(base + q);
      }
    }
  }
}

// Y4b: cut every chain where a member starts beyond every earlier end of its chain. There processRead finds only
// dead nodes (d==0 => clearTo(prev), prev = last node) and leaves exactly the read's own exons: the state no
// longer depends on history, so the pieces ("sub-chains") are independent. Prefix max of (chain<<32 | end) over
// the member array is a segmented prefix max because chain ids ascend.
struct MemberKeyIn {
  const uint32_t* chain; const int32_t* gend; const unsigned long long* colbase; int nchains;
  __device__ int chain_of(int64_t i) const {
    int lo = 0, hi = nchains;  // last c with colbase[c] <= i
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (colbase[mid] <= (unsigned long long)i) lo = mid; else hi = mid; }
    return lo;
  }
  __device__ unsigned long long operator()(int64_t i) const {
    return ((unsigned long long)(uint32_t)chain_of(i) << 32) | (uint32_t)gend[chain[i]];
  }
};
struct MemberPmOut { unsigned long long* pm; __device__ void operator()(int64_t i, unsigned long long exc, unsigned long long) const { pm[i] = exc; } };
struct SubHeadIn {
  MemberKeyIn mk; const int32_t* gstart; const unsigned long long* pm;
  __device__ uint32_t operator()(int64_t i) const {
    if (i == 0) return 1u;
    unsigned long long p = pm[i];
    int c = mk.chain_of(i);
    if ((int)(p >> 32) != c) return 1u;                       // first member of its chain
    return (uint32_t)gstart[mk.chain[i]] > (uint32_t)p ? 1u : 0u;
  }
};
struct SubHeadOut {
  uint32_t* heads;
  __device__ void operator()(int64_t i, uint32_t exc, uint32_t inc) const { if (inc != exc) heads[exc] = (uint32_t)i; }
};
__global__ void yd_subchain_total_kernel(const uint32_t* tot, uint32_t* heads, uint32_t n_members, unsigned long long* work) {
  heads[*tot] = n_members; work[0] = 0; work[1] = *tot;
}

// Y5: persistent warps pull sub-chains from a work counter. Lanes fetch the next 32 members and parse their exon
// chains into shared memory, lane 0 runs processRead/mergeRead over them, then the lanes publish the distances.
constexpr int YD_WARPS = 4;
__global__ void __launch_bounds__(YD_WARPS * 32) yd_chain_kernel(ColIn in, const uint32_t* __restrict__ rep, const int32_t* __restrict__ gstart,
                                                                const uint32_t* __restrict__ chain, const uint32_t* __restrict__ heads,
                                                                unsigned long long* work, int32_t* __restrict__ yd, long long* status) {
  __shared__ int2 s_ex[YD_WARPS][32][YD_MAXEX];
  __shared__ int s_nex[YD_WARPS][32];
  __shared__ int s_d[YD_WARPS][32];
  __shared__ uint32_t s_st[YD_WARPS][YD_CAP], s_en[YD_WARPS][YD_CAP];
  const int wl = tb_warp(), lane = tb_lane();
  const unsigned long long nsub = work[1];
  SegArr L; L.st = s_st[wl]; L.en = s_en[wl];
  bool ok = true;
  for (;;) {
    unsigned long long j = 0;
    if (lane == 0) j = atomicAdd(&work[0], 1ULL);
    j = __shfl_sync(0xffffffffu, j, 0);
    if (j >= nsub) break;
    const uint32_t c0 = heads[j], c1 = heads[j + 1];
    L.reset();
    for (uint32_t base = c0; base < c1; base += 32) {
      const uint32_t i = base + lane;
      uint32_t g = 0;
      if (i < c1) {
        g = chain[i];
        const uint32_t r = rep[g];
        ExonIter it; it.init(in.cigar, in.cig_off[r], in.cig_off[r + 1], gstart[g] - 1);
        int s, e, ne = 0;
        while (it.next(s, e)) { if (ne < YD_MAXEX) s_ex[wl][lane][ne] = make_int2(s, e); ++ne; }
        s_nex[wl][lane] = ne;
      }
      __syncwarp();
      if (lane == 0) {
        const int cnt = (int)((c1 - base) < 32u ? (c1 - base) : 32u);
        for (int q = 0; q < cnt; ++q) {
          const int ne = s_nex[wl][q];
          const int rsq = s_ex[wl][q][0].x;  // the first exon starts at the read start (setupCoordinates)
          int d;
          if (ne <= YD_MAXEX) { SmemExons src{s_ex[wl][q], ne, 0}; d = L.process(src, (uint32_t)rsq, ok); }
          else {
            const uint32_t gq = chain[base + q], rr = rep[gq];
            ExonIter it; it.init(in.cigar, in.cig_off[rr], in.cig_off[rr + 1], gstart[gq] - 1);
            d = L.process(it, (uint32_t)gstart[gq], ok);
          }
          s_d[wl][q] = d;
        }
      }
      __syncwarp();
      if (i < c1) { const int d = s_d[wl][lane]; if (d > 0) atomicMax(&yd[g], d); }
      __syncwarp();
    }
  }
  if (lane == 0 && !ok) status[CS_YD_OVERFLOW] = 1;
}

static inline unsigned grid_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

template <class T>
static int stage_in(tb_ctx* ctx, DevBuf& b, const T* src, size_t count, int on_device, const T** out) {
  if (on_device || src == nullptr) { *out = src; return 0; }
  TB_CUDA(b.ensure(count * sizeof(T) + 16));
  TB_CUDA(cudaMemcpyAsync(b.p, src, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  *out = (const T*)b.p;
  return 0;
}

}  // namespace

int tb_collapse_impl(tb_ctx* ctx, const tb_soa_in* hin, tb_groups_out* out) {
  TB_CUDA(cudaSetDevice(ctx->device));
  const int64_t n = hin->n;
  const int k = hin->n_files;
  out->n_groups = 0; out->n_kept = 0;
  if (n == 0) return 0;
  if (n >= (1LL << 31)) { ctx->set_error("tb_collapse_window: n=%lld too large for one window (< 2^31)", (long long)n); return 1; }
  if (k < 1 || k > ctx->n_samples) { ctx->set_error("tb_collapse_window: n_files=%d but the context was created for %d samples", k, ctx->n_samples); return 1; }
  if (ctx->flag_mask != 0) { ctx->set_error("tb_collapse_window: -F (flag mask) path is not implemented on the device yet"); return 1; }
  if (ctx->collapse_same) { ctx->set_error("tb_collapse_window: -A (collapse-same) path is not implemented on the device yet"); return 1; }
  if (hin->file_merged) for (int f = 0; f < k; ++f) if (hin->file_merged[f]) { ctx->set_error("tb_collapse_window: TieBrush-made inputs (re-collapse) are not implemented on the device yet"); return 1; }
  if (ctx->mode == TB_MODE_FULL && (!hin->md_off || !hin->md)) { ctx->set_error("tb_collapse_window: -L needs md_off/md"); return 1; }
  if (hin->pos_hi <= hin->pos_lo) { ctx->set_error("tb_collapse_window: pos_lo/pos_hi not set"); return 1; }
  if (out->capacity < 1) { ctx->set_error("tb_collapse_window: zero output capacity"); return 1; }
  cudaStream_t st = ctx->stream;
  DevBuf* B = ctx->buf;

  ColIn in; memset(&in, 0, sizeof(in));
  in.n = n; in.k = k; in.mode = ctx->mode; in.flag_mask = ctx->flag_mask; in.max_nh = ctx->max_nh; in.min_qual = ctx->min_qual; in.keep_bits = ctx->keep_bits;
  in.pos_lo = hin->pos_lo; in.span = (uint32_t)(hin->pos_hi - hin->pos_lo);
  if (stage_in(ctx, ctx->in_stage[0], hin->pos, (size_t)n, hin->on_device, &in.pos)) return 1;
  if (stage_in(ctx, ctx->in_stage[1], hin->flag, (size_t)n, hin->on_device, &in.flag)) return 1;
  if (stage_in(ctx, ctx->in_stage[2], hin->mapq, (size_t)n, hin->on_device, &in.mapq)) return 1;
  if (stage_in(ctx, ctx->in_stage[3], hin->strand, (size_t)n, hin->on_device, &in.strand)) return 1;
  if (stage_in(ctx, ctx->in_stage[4], hin->nh, (size_t)n, hin->on_device, &in.nh)) return 1;
  if (stage_in(ctx, ctx->in_stage[5], hin->cig_off, (size_t)n + 1, hin->on_device, &in.cig_off)) return 1;
  if (stage_in(ctx, ctx->in_stage[6], hin->cigar, (size_t)hin->n_cig, hin->on_device, &in.cigar)) return 1;
  if (ctx->mode == TB_MODE_FULL) {
    if (stage_in(ctx, ctx->in_stage[7], hin->md_off, (size_t)n + 1, hin->on_device, &in.md_off)) return 1;
    if (stage_in(ctx, ctx->in_stage[8], hin->md, (size_t)hin->n_md, hin->on_device, &in.md)) return 1;
  }
  for (int f = 0; f < k; ++f) if (hin->run_off[f] > hin->run_off[f + 1]) { ctx->set_error("tb_collapse_window: run_off not monotone"); return 1; }
  if (hin->run_off[0] != 0 || hin->run_off[k] != n) { ctx->set_error("tb_collapse_window: run_off must span [0,n]"); return 1; }

  // ---- geometry ----
  const uint32_t W = (uint32_t)((k + 31) / 32);
  uint32_t gcap = 4096;
  size_t smem_limit = ctx->smem_optin ? ctx->smem_optin : 232448;
  while (gcap > 64 && tile_smem_bytes_host((uint32_t)k, gcap, W) + 1024 > smem_limit) gcap >>= 1;
  if (tile_smem_bytes_host((uint32_t)k, gcap, W) + 1024 > smem_limit) { ctx->set_error("tb_collapse_window: %d samples do not fit the shared-memory group table", k); return 1; }
  const uint32_t T = gcap / 4;                       // a non-pile-up slot holds < 3T records (<= 3T groups < 7/8 gcap)
  const uint32_t M = (uint32_t)((n + T - 1) / T);
  const uint32_t S = in.span;

  TB_CUDA(B[XB_STATUS].ensure(sizeof(int64_t) * 16));
  TB_CUDA(ctx->pinned[0].ensure(sizeof(int64_t) * 16));
  long long* d_status = B[XB_STATUS].as<long long>();
  long long* h_status = ctx->pinned[0].as<long long>();
  memset(h_status, 0, sizeof(int64_t) * 16); h_status[CS_ERRIDX] = -1;
  TB_CUDA(cudaMemcpyAsync(d_status, h_status, sizeof(int64_t) * 16, cudaMemcpyHostToDevice, st));
  TB_CUDA(B[XB_HIST].ensure(sizeof(uint32_t) * ((size_t)S + 2)));
  TB_CUDA(B[XB_RUNOFF].ensure(sizeof(int64_t) * (k + 1)));
  TB_CUDA(cudaMemcpyAsync(B[XB_RUNOFF].p, hin->run_off, sizeof(int64_t) * (k + 1), cudaMemcpyHostToDevice, st));
  TB_CUDA(cudaStreamSynchronize(st));  // h_status / run_off host buffers are reused below
  {
    int64_t mx = (int64_t)S + 2; if ((int64_t)M + 2 > mx) mx = M + 2; if (n > mx) mx = n;
    TB_CUDA(B[XB_AGG].ensure((size_t)(tb_scan_blocks(mx) + 8) * sizeof(uint64_t)));
  }
  TB_CUDA(B[XB_SLOTPOS].ensure(sizeof(uint32_t) * ((size_t)M + 2)));
  TB_CUDA(B[XB_OFF].ensure(sizeof(uint32_t) * ((size_t)M + 1) * k));
  TB_CUDA(B[XB_GCOUNT].ensure(sizeof(uint32_t) * ((size_t)M + 2)));
  TB_CUDA(B[XB_GBASE].ensure(sizeof(uint32_t) * ((size_t)M + 2)));
  TB_CUDA(B[XB_ST_REP].ensure(sizeof(uint32_t) * n));
  TB_CUDA(B[XB_ST_YC].ensure(sizeof(float) * n));
  TB_CUDA(B[XB_ST_YX].ensure(sizeof(uint32_t) * n));
  TB_CUDA(B[XB_ST_BITS].ensure(sizeof(uint32_t) * (size_t)n * W));

  uint32_t* d_hist = B[XB_HIST].as<uint32_t>();
  const long long* d_runoff = B[XB_RUNOFF].as<long long>();
  // ---- C1, C2 ----
  if (ctx->profiling) TB_CUDA(cudaEventRecord(ctx->ev[4], st));
  TB_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(uint32_t) * ((size_t)S + 2), st));
  col_hist_kernel<<<grid_for(n, 256), 256, 0, st>>>(in, d_hist, d_status);
  ctx->launches++;
  TB_CUDA((tb_device_scan<OpSumU32>(ctx, HistIn{d_hist}, (int64_t)S + 1, B[XB_AGG].as<uint32_t>(), HistOut{d_hist})));
  const uint32_t* d_P = d_hist;  // P[0..S], P[S] = n
  // ---- C3, C4 ----
  if (ctx->profiling) TB_CUDA(cudaEventRecord(ctx->ev[5], st));
  col_slotpos_kernel<<<grid_for((int64_t)M + 1, 256), 256, 0, st>>>(d_P, S, (uint32_t)n, T, M, B[XB_SLOTPOS].as<uint32_t>());
  uint64_t off_total = ((uint64_t)M + 1) * k;
  col_off_init_kernel<<<grid_for((int64_t)off_total, 256), 256, 0, st>>>(B[XB_OFF].as<uint32_t>(), d_runoff, k, off_total);
  col_off_kernel<<<grid_for(n, 256), 256, 0, st>>>(in, d_runoff, d_P, T, B[XB_OFF].as<uint32_t>(), d_status);
  ctx->launches += 3;
  // ---- C5 ----
  TileParams tp; memset(&tp, 0, sizeof(tp));
  tp.T = T; tp.M = M; tp.gcap = gcap; tp.W = W; tp.P = d_P; tp.slotpos = B[XB_SLOTPOS].as<uint32_t>(); tp.off = B[XB_OFF].as<uint32_t>();
  tp.gcount = B[XB_GCOUNT].as<uint32_t>(); tp.st_rep = B[XB_ST_REP].as<uint32_t>(); tp.st_yc = B[XB_ST_YC].as<float>();
  tp.st_yx = B[XB_ST_YX].as<uint32_t>(); tp.st_bits = B[XB_ST_BITS].as<uint32_t>(); tp.status = d_status; tp.seed = 0x243F6A8885A308D3ULL;
  size_t smem = tile_smem_bytes_host((uint32_t)k, gcap, W);
  TB_CUDA(cudaFuncSetAttribute(col_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (ctx->profiling) TB_CUDA(cudaEventRecord(ctx->ev[0], st));
  col_tile_kernel<<<M, TILE_THREADS, smem, st>>>(in, tp);
  ctx->launches++;
  if (ctx->profiling) TB_CUDA(cudaEventRecord(ctx->ev[1], st));
  // ---- C6 ----
  if (ctx->profiling) TB_CUDA(cudaEventRecord(ctx->ev[6], st));
  TB_CUDA((tb_device_scan<OpSumU32>(ctx, GcIn{tp.gcount}, (int64_t)M, B[XB_AGG].as<uint32_t>(), GcOut{B[XB_GBASE].as<uint32_t>()})));
  col_store_total_kernel<<<1, 1, 0, st>>>(B[XB_AGG].as<uint32_t>() + tb_scan_blocks(M), d_status);
  ctx->launches++;
  TB_CUDA(cudaMemcpyAsync(h_status, d_status, sizeof(int64_t) * 16, cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaStreamSynchronize(st));
  if (ctx->profiling) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]) == cudaSuccess) ctx->last_ms[0] = ms;
    if (cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]) == cudaSuccess) ctx->last_ms[2] = ms;
    if (cudaEventElapsedTime(&ms, ctx->ev[5], ctx->ev[0]) == cudaSuccess) ctx->last_ms[3] = ms;
  }
  if (h_status[CS_ERR] == ERR_POS_RANGE) { ctx->set_error("tb_collapse_window: record %lld has pos outside [pos_lo,pos_hi)", h_status[CS_ERRIDX]); return 1; }
  if (h_status[CS_ERR] == ERR_UNSORTED) { ctx->set_error("tb_collapse_window: run not coordinate-sorted at record %lld", h_status[CS_ERRIDX]); return 1; }
  if (h_status[CS_TABLE_OVERFLOW]) { ctx->set_error("tb_collapse_window: more than %u distinct alignments at one start position (group table overflow; multi-pass fallback not implemented)", gcap - (gcap >> 3)); return 1; }
  const int64_t G = h_status[CS_NGROUPS];
  if (G > out->capacity) { ctx->set_error("tb_collapse_window: output capacity %lld < %lld groups", (long long)out->capacity, (long long)G); return 1; }
  out->n_kept = h_status[CS_NKEPT];
  out->n_groups = G;
  if (G == 0) return 0;

  uint32_t* o_rep = out->rep_index; float* o_yc = out->yc; uint32_t* o_yx = out->yx; int32_t* o_yd = out->yd;
  if (!out->on_device) {
    TB_CUDA(ctx->out_stage[0].ensure(sizeof(uint32_t) * G)); TB_CUDA(ctx->out_stage[1].ensure(sizeof(float) * G));
    TB_CUDA(ctx->out_stage[2].ensure(sizeof(uint32_t) * G)); TB_CUDA(ctx->out_stage[3].ensure(sizeof(int32_t) * G));
    o_rep = ctx->out_stage[0].as<uint32_t>(); o_yc = ctx->out_stage[1].as<float>(); o_yx = ctx->out_stage[2].as<uint32_t>(); o_yd = ctx->out_stage[3].as<int32_t>();
  }
  TB_CUDA(B[XB_BITS].ensure(sizeof(uint32_t) * (size_t)G * W));
  col_compact_kernel<<<M, 128, 0, st>>>(tp, S, B[XB_GBASE].as<uint32_t>(), o_rep, o_yc, o_yx, B[XB_BITS].as<uint32_t>(), G);
  ctx->launches++;
  if (ctx->profiling) TB_CUDA(cudaEventRecord(ctx->ev[7], st));

  // ---- C7: YD ----
  {
    const int64_t nblk = (G + YD_BLOCK - 1) / YD_BLOCK;
    TB_CUDA(B[XB_GSTART].ensure(sizeof(int32_t) * G));
    TB_CUDA(B[XB_YD].ensure(sizeof(int32_t) * G));                           // end per group
    TB_CUDA(B[XB_GKEY].ensure((size_t)G));                                  // strand char per group
    TB_CUDA(B[XB_GPM].ensure(sizeof(uint32_t) * (size_t)nblk * 2 * k));       // per (block, chain) counts -> offsets
    TB_CUDA(B[XB_GEND].ensure(sizeof(uint32_t) * 2 * k + sizeof(uint64_t) * (2 * k + 6)));
    TB_CUDA(cudaMemsetAsync(o_yd, 0, sizeof(int32_t) * G, st));
    int32_t* gstart = B[XB_GSTART].as<int32_t>(); int32_t* gend = B[XB_YD].as<int32_t>(); uint8_t* gstrand = B[XB_GKEY].as<uint8_t>();
    uint32_t* blkcnt = B[XB_GPM].as<uint32_t>();
    unsigned long long* colbase = B[XB_GEND].as<unsigned long long>();
    unsigned long long* work = colbase + 2 * k + 2;
    uint32_t* coltot = (uint32_t*)(colbase + 2 * k + 6);
    const uint32_t* d_bits = B[XB_BITS].as<uint32_t>();
    if (ctx->profiling) TB_CUDA(cudaEventRecord(ctx->ev[2], st));
    yd_group_info_kernel<<<grid_for(G, 256), 256, 0, st>>>(in, o_rep, G, gstart, gend, gstrand);
    const int64_t nwarps = nblk * W;
    yd_count_kernel<<<grid_for(nwarps * 32, 128), 128, 0, st>>>(d_bits, W, k, G, gstrand, blkcnt, nblk);
    yd_prefix_kernel<<<grid_for(2 * k, 128), 128, 0, st>>>(blkcnt, nblk, k, coltot);
    yd_colbase_kernel<<<1, 1, 0, st>>>(coltot, k, colbase);
    ctx->launches += 4;
    // total chain members (<= 2 x sum of YX) is only known on the device
    TB_CUDA(cudaMemcpyAsync(h_status, colbase + 2 * k, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    TB_CUDA(cudaStreamSynchronize(st));
    const int64_t n_members = h_status[0];
    if (n_members >= (1LL << 32)) { ctx->set_error("tb_collapse_window: %lld chain members exceed the 32-bit YD index", (long long)n_members); return 1; }
    if (n_members > 0) {
      uint32_t* chain = nullptr; uint32_t* heads = nullptr; unsigned long long* pm = nullptr;
      TB_CUDA(B[XB_BHEAD].ensure(sizeof(uint32_t) * ((size_t)n_members + 32)));
      TB_CUDA(B[XB_ST_REP].ensure(sizeof(uint32_t) * ((size_t)n_members + 32)));  // staging is dead by now: reuse
      TB_CUDA(B[XB_ST_BITS].ensure(sizeof(uint64_t) * ((size_t)n_members + 32)));
      TB_CUDA(B[XB_AGG].ensure((size_t)(tb_scan_blocks(n_members) + 8) * sizeof(uint64_t)));
      chain = B[XB_BHEAD].as<uint32_t>(); heads = B[XB_ST_REP].as<uint32_t>(); pm = B[XB_ST_BITS].as<unsigned long long>();
      yd_scatter_kernel<<<grid_for(nwarps * 32, 128), 128, 0, st>>>(d_bits, W, k, G, gstrand, blkcnt, colbase, nblk, chain);
      ctx->launches++;
      MemberKeyIn mk{chain, gend, colbase, 2 * k};
      TB_CUDA((tb_device_scan<OpMaxU64>(ctx, mk, n_members, B[XB_AGG].as<unsigned long long>(), MemberPmOut{pm})));
      TB_CUDA((tb_device_scan<OpSumU32>(ctx, SubHeadIn{mk, gstart, pm}, n_members, B[XB_AGG].as<uint32_t>(), SubHeadOut{heads})));
      yd_subchain_total_kernel<<<1, 1, 0, st>>>(B[XB_AGG].as<uint32_t>() + tb_scan_blocks(n_members), heads, (uint32_t)n_members, work);
      yd_chain_kernel<<<ctx->sm_count * 12, YD_WARPS * 32, 0, st>>>(in, o_rep, gstart, chain, heads, work, o_yd, d_status);
      ctx->launches += 2;
    }
    if (ctx->profiling) TB_CUDA(cudaEventRecord(ctx->ev[3], st));
  }
  TB_CUDA(cudaMemcpyAsync(h_status, d_status, sizeof(int64_t) * 16, cudaMemcpyDeviceToHost, st));
  if (!out->on_device) {
    TB_CUDA(cudaMemcpyAsync(out->rep_index, o_rep, sizeof(uint32_t) * G, cudaMemcpyDeviceToHost, st));
    TB_CUDA(cudaMemcpyAsync(out->yc, o_yc, sizeof(float) * G, cudaMemcpyDeviceToHost, st));
    TB_CUDA(cudaMemcpyAsync(out->yx, o_yx, sizeof(uint32_t) * G, cudaMemcpyDeviceToHost, st));
    TB_CUDA(cudaMemcpyAsync(out->yd, o_yd, sizeof(int32_t) * G, cudaMemcpyDeviceToHost, st));
  }
  TB_CUDA(cudaStreamSynchronize(st));
  if (ctx->profiling) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7]) == cudaSuccess) ctx->last_ms[4] = ms;
    if (cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]) == cudaSuccess) ctx->last_ms[5] = ms;
  }
  if (h_status[CS_YD_OVERFLOW]) { ctx->set_error("tb_collapse_window: YD segment list exceeded %d live nodes (spill path not implemented)", YD_CAP); return 1; }
  return 0;
}
