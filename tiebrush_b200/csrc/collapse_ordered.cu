// collapse_ordered.cu — exact front end of the collapse: the reference's per-position sorted unique list, emulated
// record by record in the reference's merge order.
//
// Needed whenever the result depends on arrival order or on the binary-search probe sequence:
//   -F          cmpFlags returns 1 in BOTH directions for different masked flags (src/tiebrush.cpp:275-283), so the
//               comparator is not a strict weak order and GList::Found (include/gclib/GList.hh:567-604) decides by
//               its probe sequence which existing group absorbs a record
//   -A          a duplicate is skipped iff its sample already contributed AND it has the pair order and QNAME of the
//               representative (src/tiebrush.cpp:421-434)
//   tbMerged    TieBrush-made inputs add their YC / YX tags and carry max YD (src/tiebrush.cpp:389-395,412-419)
//   --store-frac  YC += 1/NH in double, in arrival order (src/tiebrush.cpp:396-401)
// and as the fallback when one start position holds more distinct alignments than a shared-memory table.
//
//   O1  ref_len per record; segmented running maximum E of `end` per (file, position) over the file-major order
//       (a device-wide segmented max-scan): the reference's priority queue pops records in the order
//       (start, E, fidx, index in file) (src/tmerge.h:28-50)
//   O2  stable LSD radix sort of the records by (position, E): file-major input order supplies (fidx, index)
//   O3  one thread per start position walks its records in that order against a sorted list of groups kept in global
//       memory, with GList::Found's exact probe order (first, last, then midpoint bisection)
//   O4  stream compaction of the per-position group lists into the dense output
#include "collapse_internal.cuh"

namespace {

struct SegMax { uint32_t flag; uint32_t val; };
struct OpSegMax {
  typedef SegMax T;
  __host__ __device__ static T identity() { return SegMax{0u, 0u}; }
  __host__ __device__ static T combine(T a, T b) { return b.flag ? b : SegMax{a.flag, a.val > b.val ? a.val : b.val}; }
};

__global__ void __launch_bounds__(256) ord_reflen_kernel(ColIn in, uint32_t* __restrict__ reflen) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= in.n) return;
  reflen[i] = (uint32_t)tb_ref_len(in.cigar, in.cig_off[i], in.cig_off[i + 1]);
}

__device__ __forceinline__ int ord_file_of(const long long* __restrict__ run_off, int k, int64_t i) {
  int lo = 0, hi = k;  // last f with run_off[f] <= i
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (run_off[mid] <= i) lo = mid; else hi = mid; }
  return lo;
}

struct SegIn {
  ColIn in; const uint32_t* reflen; const long long* run_off;
  __device__ SegMax operator()(int64_t i) const {
    bool head = i == 0 || in.pos[i - 1] != in.pos[i];
    if (!head) { const int f = ord_file_of(run_off, in.k, i); head = (run_off[f] == i); }
    return SegMax{head ? 1u : 0u, reflen[i]};
  }
};
struct SegOut {
  ColIn in; unsigned long long* key; uint32_t* val;
  __device__ void operator()(int64_t i, SegMax, SegMax inc) const {
    key[i] = ((unsigned long long)(uint32_t)(in.pos[i] - in.pos_lo) << 32) | inc.val;
    val[i] = (uint32_t)i;
  }
};

struct OrdParams {
  const uint32_t* P; const uint32_t* order; const uint32_t* reflen; const long long* run_off; const uint8_t* merged;
  uint32_t W;
  uint32_t* list;       // [n] per-position sorted list of group slots (slot j of position p lives at P[p]+j)
  uint32_t* g_rep; double* g_yc; long long* g_yx; int32_t* g_yd; uint32_t* g_bits;   // [n] group slots
  uint32_t* st_rep; float* st_yc; uint32_t* st_yx; int32_t* st_yd; uint32_t* st_bits; uint32_t* valid;   // staged in list order
  long long* status;
};

// SPData::operator< (src/tiebrush.cpp:438-457) for two records of one start position
__device__ bool ord_less(const ColIn& in, const OrdParams& op, uint32_t a, uint32_t b) {
  const uint8_t sa = in.strand[a], sb = in.strand[b];
  if (sa != sb) return sa < sb;
  const uint32_t ea = op.reflen[a], eb = op.reflen[b];
  if (ea != eb) return ea < eb;
  return tb_mode_cmp_flags(in, a, b) < 0;
}
// GList::DefaultCompareProc (include/gclib/GList.hh:85-90)
__device__ int ord_compare(const ColIn& in, const OrdParams& op, uint32_t x, uint32_t y) {
  if (ord_less(in, op, y, x)) return 1;
  if (ord_less(in, op, x, y)) return -1;
  return 0;
}
__device__ __forceinline__ int ord_pair_order(uint16_t fl) { return (fl & 0x40) ? 1 : ((fl & 0x80) ? 2 : 0); }  // GSam.h:315-321

// GList::Found (GList.hh:567-604) of record i in the sorted list L[0..nl) of position slots based at b: first, last, then bisection
__device__ __forceinline__ bool ord_found(const ColIn& in, const OrdParams& op, uint32_t b, const uint32_t* L, uint32_t nl, uint32_t i, uint32_t& idx) {
  idx = 0;
  if (nl == 0) return false;
  if (ord_compare(in, op, op.g_rep[b + L[0]], i) > 0) { idx = 0; return false; }
  if (ord_compare(in, op, i, op.g_rep[b + L[nl - 1]]) > 0) { idx = nl; return false; }
  int l = 0, h = (int)nl - 1;
  while (l <= h) {
    const int mid = l + ((h - l) >> 1);
    const int c = ord_compare(in, op, op.g_rep[b + L[mid]], i);
    if (c < 0) l = mid + 1;
    else { h = mid - 1; if (c == 0) { idx = (uint32_t)mid; return true; } }
  }
  idx = (uint32_t)l;
  return false;
}
// dupAdd (tiebrush.cpp:408-436) of record i (file f) into group slot s
__device__ __forceinline__ void ord_dup_add(const ColIn& in, const OrdParams& op, uint32_t s, uint32_t i, int f, bool merged, uint16_t fl) {
  const uint32_t W = op.W;
  if (merged) {
    double v = (double)in.yc_in[i]; if (v == 0.0) v = 1.0;
    op.g_yc[s] += v;
    op.g_yx[s] += in.yx_in[i];
    const int32_t yd = in.yd_in[i];
    if (yd > op.g_yd[s]) op.g_yd[s] = yd;
  } else {
    const uint32_t bw = op.g_bits[(uint64_t)s * W + ((uint32_t)f >> 5)], bit = 1u << (f & 31);
    const uint32_t rep = op.g_rep[s];
    if (!in.collapse_same || !(bw & bit) || ord_pair_order(fl) != ord_pair_order(in.flag[rep]) || in.qhash[i] != in.qhash[rep]) {
      if (in.keep_bits & TB_STORE_FRAC) { const int nh = in.nh[i] ? in.nh[i] : 1; op.g_yc[s] += 1.0 / nh; }
      else op.g_yc[s] += 1.0;
      op.g_bits[(uint64_t)s * W + ((uint32_t)f >> 5)] = bw | bit;
    }
  }
}
// settle (tiebrush.cpp:378-406) + sortInsert of record i as new group j at list index idx
__device__ __forceinline__ void ord_settle(const ColIn& in, const OrdParams& op, uint32_t b, uint32_t* L, uint32_t& nl, uint32_t idx, uint32_t i, int f, bool merged) {
  const uint32_t W = op.W;
  const uint32_t j = nl;            // slots are handed out in arrival order
  const uint32_t s = b + j;
  op.g_rep[s] = i;
  for (uint32_t w = 0; w < W; ++w) op.g_bits[(uint64_t)s * W + w] = 0;
  if (merged) {
    double v = (double)in.yc_in[i]; if (v == 0.0) v = 1.0;
    op.g_yc[s] = v; op.g_yx[s] = in.yx_in[i]; op.g_yd[s] = in.yd_in[i];
  } else {
    if (in.keep_bits & TB_STORE_FRAC) { const int nh = in.nh[i] ? in.nh[i] : 1; op.g_yc[s] = 1.0 / nh; }
    else op.g_yc[s] = 1.0;
    op.g_yx[s] = 0; op.g_yd[s] = 0;
    op.g_bits[(uint64_t)s * W + ((uint32_t)f >> 5)] = 1u << (f & 31);
  }
  for (uint32_t q = nl; q > idx; --q) L[q] = L[q - 1];
  L[idx] = j; ++nl;
}
// flushPData order (tiebrush.cpp:501-530): the list order
__device__ __forceinline__ void ord_flush_one(const OrdParams& op, uint32_t b, const uint32_t* L, uint32_t x) {
  const uint32_t W = op.W;
  const uint32_t s = b + L[x], o = b + x;
  uint32_t pc = 0;
  for (uint32_t w = 0; w < W; ++w) { const uint32_t bw = op.g_bits[(uint64_t)s * W + w]; pc += __popc(bw); op.st_bits[(uint64_t)o * W + w] = bw; }
  op.st_rep[o] = op.g_rep[s];
  op.st_yc[o] = (float)op.g_yc[s];
  op.st_yx[o] = (uint32_t)((long long)pc + op.g_yx[s]);
  op.st_yd[o] = op.g_yd[s] > 0 ? op.g_yd[s] : 0;
  op.valid[o] = 1u;
}

// one THREAD per start position (positions of at least `deep` records are left to the warp kernel below)
__global__ void __launch_bounds__(128) ord_emulate_kernel(ColIn in, OrdParams op, uint32_t deep) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= in.span) return;
  const uint32_t b = op.P[p], e = op.P[p + 1];
  if (e == b || e - b >= deep) return;
  uint32_t* L = op.list + b;
  uint32_t nl = 0;
  long long kept = 0;
  for (uint32_t r = b; r < e; ++r) {
    const uint32_t i = op.order[r];
    const uint16_t fl = in.flag[i];
    if (!tb_passes_options(in, fl, in.mapq[i], in.nh[i])) continue;
    ++kept;
    const int f = ord_file_of(op.run_off, in.k, (int64_t)i);
    const bool merged = op.merged && op.merged[f];
    uint32_t idx;
    if (ord_found(in, op, b, L, nl, i, idx)) ord_dup_add(in, op, b + L[idx], i, f, merged, fl);
    else ord_settle(in, op, b, L, nl, idx, i, f, merged);
  }
  for (uint32_t x = 0; x < nl; ++x) ord_flush_one(op, b, L, x);
  if (kept) atomicAdd((unsigned long long*)&op.status[CS_NKEPT], (unsigned long long)kept);
}

// deep positions (pile-ups): their list of start positions
__global__ void __launch_bounds__(256) ord_deep_list_kernel(const uint32_t* __restrict__ P, uint32_t span, uint32_t deep, uint32_t* __restrict__ list, unsigned int* __restrict__ count) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= span) return;
  if (P[p + 1] - P[p] >= deep) list[atomicAdd(count, 1u)] = p;
}

// one WARP per deep position. The reference's loop is sequential because every record meets the list its predecessors
// left behind — but only an INSERT changes the list: the 32 lanes search 32 consecutive records of the merge order against
// the current list at once (the expensive part: ~log2(groups) comparator calls with global loads each); the records in
// front of the first one that is not found are then applied one after the other in merge order (dupAdd reads and writes
// group state, so -A / --store-frac / TieBrush-made inputs keep their order), that record is inserted, and the search
// restarts behind it. Exactly the sequence of list states of the one-thread loop.
__global__ void __launch_bounds__(128) ord_emulate_warp_kernel(ColIn in, OrdParams op, const uint32_t* __restrict__ deep_list, const unsigned int* __restrict__ deep_count) {
  const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wid >= *deep_count) return;
  const uint32_t p = deep_list[wid];
  const uint32_t b = op.P[p], e = op.P[p + 1];
  uint32_t* L = op.list + b;
  uint32_t nl = 0;
  long long kept = 0;
  uint32_t r0 = b;
  while (r0 < e) {
    const uint32_t r = r0 + lane;
    const bool have = r < e;
    uint32_t i = 0, idx = 0; uint16_t fl = 0; bool pass = false, found = false;
    if (have) {
      i = op.order[r]; fl = in.flag[i];
      pass = tb_passes_options(in, fl, in.mapq[i], in.nh[i]);
      if (pass) found = ord_found(in, op, b, L, nl, i, idx);
    }
    const unsigned need_insert = __ballot_sync(0xffffffffu, have && pass && !found);
    const unsigned in_range = __ballot_sync(0xffffffffu, have);
    const int stop = need_insert ? __ffs(need_insert) - 1 : 32;      // first lane whose record opens a new group
    // apply the found records in front of it, in merge order
    for (int q = 0; q < stop && ((in_range >> q) & 1u); ++q) {
      if ((int)lane == q && pass) {
        const int f = ord_file_of(op.run_off, in.k, (int64_t)i);
        ord_dup_add(in, op, b + L[idx], i, f, op.merged && op.merged[f], fl);
        ++kept;
      }
      __syncwarp();
    }
    if (stop < 32) {
      if ((int)lane == stop) {
        const int f = ord_file_of(op.run_off, in.k, (int64_t)i);
        ord_settle(in, op, b, L, nl, idx, i, f, op.merged && op.merged[f]);
        ++kept;
      }
      __threadfence_block();
      nl = __shfl_sync(0xffffffffu, nl, stop);
      r0 += (uint32_t)stop + 1u;
    } else {
      r0 += 32u;
    }
    __syncwarp();
  }
  for (uint32_t x = lane; x < nl; x += 32) ord_flush_one(op, b, L, x);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) kept += __shfl_xor_sync(0xffffffffu, kept, d);
  if (lane == 0 && kept) atomicAdd((unsigned long long*)&op.status[CS_NKEPT], (unsigned long long)kept);
}

struct ValidIn { const uint32_t* v; __device__ uint32_t operator()(int64_t i) const { return v[i]; } };
struct ValidOut {
  OrdParams op; uint32_t* o_rep; float* o_yc; uint32_t* o_yx; int32_t* o_yd; uint32_t* o_bits; long long capacity;
  __device__ void operator()(int64_t i, uint32_t exc, uint32_t inc) const {
    if (inc == exc || (long long)exc >= capacity) return;
    o_rep[exc] = op.st_rep[i]; o_yc[exc] = op.st_yc[i]; o_yx[exc] = op.st_yx[i]; o_yd[exc] = op.st_yd[i];
    for (uint32_t w = 0; w < op.W; ++w) o_bits[(uint64_t)exc * op.W + w] = op.st_bits[(uint64_t)i * op.W + w];
  }
};
__global__ void ord_store_total_kernel(const uint32_t* tot, long long* status) { status[CS_NGROUPS] = *tot; }

}  // namespace

int col_front_ordered(tb_ctx* ctx, const ColIn& in, const ColGeom& g, ColGroups& out, int64_t* n_groups, int64_t* n_kept) {
  cudaStream_t st = ctx->stream;
  DevBuf* B = ctx->buf;
  const int64_t n = g.n; const uint32_t W = g.W;
  TB_CUDA(B[XB_ORD_REFLEN].ensure(sizeof(uint32_t) * n));
  TB_CUDA(B[XB_ORD_KEY].ensure(sizeof(uint64_t) * n)); TB_CUDA(B[XB_ORD_KEY2].ensure(sizeof(uint64_t) * n));
  TB_CUDA(B[XB_ORD_VAL].ensure(sizeof(uint32_t) * n)); TB_CUDA(B[XB_ORD_VAL2].ensure(sizeof(uint32_t) * n));
  TB_CUDA(B[XB_ORD_TABLE].ensure(sizeof(uint32_t) * tb_radix_table_elems(n)));
  TB_CUDA(B[XB_ORD_AGG].ensure(sizeof(uint64_t) * (tb_radix_agg_elems(n) + tb_scan_blocks(n) + 16)));
  uint32_t* reflen = B[XB_ORD_REFLEN].as<uint32_t>();
  // ---- O1 ----
  ord_reflen_kernel<<<tb_grid_for(n, 256), 256, 0, st>>>(in, reflen);
  ctx->launches++;
  TB_CUDA((tb_device_scan<OpSegMax>(ctx, SegIn{in, reflen, g.d_runoff}, n, B[XB_ORD_AGG].as<SegMax>(),
                                   SegOut{in, B[XB_ORD_KEY].as<unsigned long long>(), B[XB_ORD_VAL].as<uint32_t>()})));
  // ---- O2 ----
  int pos_bits = 1; while (pos_bits < 32 && (1ull << pos_bits) < (unsigned long long)g.S) ++pos_bits;
  uint64_t* rk; uint32_t* rv;
  TB_CUDA(tb_radix_sort(ctx, B[XB_ORD_KEY].as<uint64_t>(), B[XB_ORD_VAL].as<uint32_t>(), B[XB_ORD_KEY2].as<uint64_t>(), B[XB_ORD_VAL2].as<uint32_t>(), n,
                        0, 32 + pos_bits, B[XB_ORD_TABLE].as<uint32_t>(), B[XB_ORD_AGG].as<uint32_t>(), &rk, &rv));
  // ---- O3 ----
  TB_CUDA(B[XB_ORD_LIST].ensure(sizeof(uint32_t) * n));
  TB_CUDA(B[XB_ORD_GREP].ensure(sizeof(uint32_t) * n));
  TB_CUDA(B[XB_ORD_GYC].ensure(sizeof(double) * n));
  TB_CUDA(B[XB_ORD_GYX].ensure(sizeof(long long) * n));
  TB_CUDA(B[XB_ORD_GYD].ensure(sizeof(int32_t) * n));
  TB_CUDA(B[XB_ORD_VALID].ensure(sizeof(uint32_t) * n));
  TB_CUDA(B[XB_ST_REP].ensure(sizeof(uint32_t) * n));
  TB_CUDA(B[XB_ST_YC].ensure(sizeof(float) * n));
  TB_CUDA(B[XB_ST_YX].ensure(sizeof(uint32_t) * n));
  TB_CUDA(B[XB_ST_YD].ensure(sizeof(int32_t) * n));
  TB_CUDA(B[XB_ST_BITS].ensure(sizeof(uint32_t) * (size_t)n * W));
  OrdParams op; memset(&op, 0, sizeof(op));
  op.P = g.P; op.order = rv; op.reflen = reflen; op.run_off = g.d_runoff; op.merged = g.d_merged; op.W = W;
  op.list = B[XB_ORD_LIST].as<uint32_t>(); op.g_rep = B[XB_ORD_GREP].as<uint32_t>(); op.g_yc = B[XB_ORD_GYC].as<double>();
  op.g_yx = B[XB_ORD_GYX].as<long long>(); op.g_yd = B[XB_ORD_GYD].as<int32_t>();
  op.st_rep = B[XB_ST_REP].as<uint32_t>(); op.st_yc = B[XB_ST_YC].as<float>(); op.st_yx = B[XB_ST_YX].as<uint32_t>();
  op.st_yd = B[XB_ST_YD].as<int32_t>(); op.st_bits = B[XB_ST_BITS].as<uint32_t>(); op.valid = B[XB_ORD_VALID].as<uint32_t>();
  op.status = g.d_status;
  TB_CUDA(B[XB_ORD_GBITS].ensure(sizeof(uint32_t) * (size_t)n * W));
  op.g_bits = B[XB_ORD_GBITS].as<uint32_t>();
  TB_CUDA(cudaMemsetAsync(op.valid, 0, sizeof(uint32_t) * n, st));
  // positions of at least `deep` records (pile-ups) get a warp each, the others a thread each
  uint32_t deep = 96;
  if (const char* e = getenv("TB_ORD_DEEP")) { const long v = atol(e); deep = v > 0 ? (uint32_t)v : 0xffffffffu; }
  TB_CUDA(B[XB_ORD_DEEP].ensure(sizeof(uint32_t) * ((size_t)(n / (deep ? deep : 1)) + 64)));
  unsigned int* d_deep_count = B[XB_ORD_DEEP].as<unsigned int>();
  uint32_t* d_deep_list = B[XB_ORD_DEEP].as<uint32_t>() + 16;
  TB_CUDA(cudaMemsetAsync(d_deep_count, 0, 64, st));
  ord_deep_list_kernel<<<tb_grid_for((int64_t)g.S, 256), 256, 0, st>>>(g.P, g.S, deep, d_deep_list, d_deep_count);
  ord_emulate_kernel<<<tb_grid_for((int64_t)g.S, 128), 128, 0, st>>>(in, op, deep);
  {
    unsigned int h_deep = 0;
    TB_CUDA(cudaMemcpyAsync(&h_deep, d_deep_count, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    TB_CUDA(cudaStreamSynchronize(st));
    if (h_deep > 0) { ord_emulate_warp_kernel<<<tb_grid_for((int64_t)h_deep * 32, 128), 128, 0, st>>>(in, op, d_deep_list, d_deep_count); ctx->launches++; }
    ctx->last_ord_deep = (int64_t)h_deep;
  }
  ctx->launches += 2;
  // ---- O4 ----
  TB_CUDA(B[XB_BITS].ensure(sizeof(uint32_t) * (size_t)n * W));   // dense bitsets; G <= n is only known after the scan
  out.bits = B[XB_BITS].as<uint32_t>();
  TB_CUDA(B[XB_AGG].ensure((size_t)(tb_scan_blocks(n) + 8) * sizeof(uint64_t)));
  TB_CUDA((tb_device_scan<OpSumU32>(ctx, ValidIn{op.valid}, n, B[XB_AGG].as<uint32_t>(),
                                   ValidOut{op, out.rep, out.yc, out.yx, out.yd, out.bits, (long long)out.capacity})));
  ord_store_total_kernel<<<1, 1, 0, st>>>(B[XB_AGG].as<uint32_t>() + tb_scan_blocks(n), g.d_status);
  ctx->launches++;
  long long* h_status = ctx->pinned[0].as<long long>();
  TB_CUDA(cudaMemcpyAsync(h_status, g.d_status, sizeof(int64_t) * 16, cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaStreamSynchronize(st));
  *n_groups = h_status[CS_NGROUPS]; *n_kept = h_status[CS_NKEPT];
  if (*n_groups > out.capacity) { ctx->set_error("tb_collapse_window: output capacity %lld < %lld groups", (long long)out.capacity, (long long)*n_groups); return 1; }
  return 0;
}
