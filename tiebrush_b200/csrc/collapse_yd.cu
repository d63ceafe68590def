// collapse_yd.cu — the YD tag: GSegList::processRead / mergeRead (src/tiebrush.cpp:151-250) driven by flushPData
// (src/tiebrush.cpp:512-524) over the collapsed groups in output order.
//
// The reference keeps, per (sample, strand list), a linked list of merged exon segments; for every emitted group and
// every sample that contributed to it, processRead returns the distance from the read start back to the start of the
// segment that reaches it, and YD = max over those samples (and over YD tags carried by TieBrush-made inputs).
// The list is a sequential state machine with binding quirks (touching segments are not merged; when the merge cursor
// runs off the list the current and all later exons are dropped), so it is emulated literally.
//
// A chain = (sample s, strand list): the groups, in output order, that contain s and whose strand feeds the list
// ('+','.' -> forward; '-','.' -> reverse). Chains are independent (tiebrush.cpp:512-521). Stages:
//   Y1  one 32-byte descriptor per group (start, strand, first three exons of the representative) + its end
//   Y2-Y4  member lists of all 2k chains by a stable counting scatter over the sample bitsets
//   Y5  every chain is cut where a member starts beyond every earlier end of its chain: there processRead finds only dead
//       nodes (d == 0 => clearTo(prev), prev = last node) and leaves exactly the read's own exons, so the pieces
//       ("sub-chains") are independent. Segmented prefix max over the member array (chain ids ascend).
//   Y6  persistent warps pull batches of 32 sub-chains: short ones run one per LANE (lists in shared memory, [node][lane]
//       layout), long ones are walked by the whole warp (lanes prefetch 32 descriptors, lane 0 runs the state machine).
// Nodes lying before `prev` can never be touched again before they are freed (disjoint, sorted, ends < read start), so
// they are dropped eagerly; results are unchanged. If a list outgrows shared memory the launch is repeated with lists in
// global memory.
#include "collapse_internal.cuh"

namespace {

constexpr int YD_INLINE_EX = 3;   // exons kept in the descriptor; longer chains are decoded from the CIGAR
struct __align__(16) GDesc {      // 32 bytes
  int32_t start;                  // 1-based start of the representative (== first exon start)
  uint32_t meta;                  // n_exons (low 16) | strand char << 16
  int32_t e[6];                   // ex0.end, ex1.start, ex1.end, ex2.start, ex2.end, rep record index
};

__global__ void __launch_bounds__(256) yd_desc_kernel(ColIn in, const uint32_t* __restrict__ rep, int64_t G, GDesc* __restrict__ desc,
                                                      uint32_t* __restrict__ gend, uint8_t* __restrict__ gstrand) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  const uint32_t r = rep[g];
  const int pos = in.pos[r];
  ExonIter it; it.init(in.cigar, in.cig_off[r], in.cig_off[r + 1], pos);
  GDesc d; d.start = pos + 1;
  int s, e, ne = 0;
#pragma unroll
  for (int q = 0; q < 6; ++q) d.e[q] = 0;
  while (it.next(s, e)) {
    if (ne == 0) d.e[0] = e;
    else if (ne == 1) { d.e[1] = s; d.e[2] = e; }
    else if (ne == 2) { d.e[3] = s; d.e[4] = e; }
    ++ne;
  }
  d.e[5] = (int32_t)r;
  const uint8_t sc = in.strand[r];
  d.meta = (uint32_t)(ne > 0xffff ? 0xffff : ne) | ((uint32_t)sc << 16);
  desc[g] = d;
  gend[g] = (uint32_t)(pos + it.l);
  gstrand[g] = sc;
}

constexpr int YD_BLOCK = 1024;   // groups per counting block

// Y2: members per (block of groups, chain). One warp per (block, 32-sample word).
__global__ void __launch_bounds__(128) yd_count_kernel(const uint32_t* __restrict__ bits, uint32_t W, int k, int64_t G, const uint8_t* __restrict__ gstrand,
                                                      uint32_t* __restrict__ blkcnt, int64_t nblk) {
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= nblk * W) return;
  const int64_t b = warp / W; const uint32_t w = (uint32_t)(warp % W);
  const int lane = tb_lane();
  const int64_t g0 = b * YD_BLOCK, g1 = (g0 + YD_BLOCK < G) ? g0 + YD_BLOCK : G;
  uint32_t cf = 0, cr = 0;
  for (int64_t base = g0; base < g1; base += 32) {
    const int64_t g = base + lane;
    uint32_t word = 0; uint8_t sc = 0;
    if (g < g1) { word = bits[(uint64_t)g * W + w]; sc = gstrand[g]; }
    unsigned any = __ballot_sync(0xffffffffu, word != 0);
    while (any) {
      const int q = __ffs(any) - 1; any &= any - 1;
      const uint32_t wq = __shfl_sync(0xffffffffu, word, q);
      const int sq = __shfl_sync(0xffffffffu, (int)sc, q);
      if ((wq >> lane) & 1u) { cf += (sq != '-'); cr += (sq != '+'); }
    }
  }
  const int s = (int)(w * 32 + lane);
  if (s < k) { blkcnt[(b * 2 + 0) * k + s] = cf; blkcnt[(b * 2 + 1) * k + s] = cr; }
}

// Y3: exclusive prefix over blocks per chain column (in place), then chain base offsets
__global__ void __launch_bounds__(128) yd_prefix_kernel(uint32_t* __restrict__ blkcnt, int64_t nblk, int k, uint32_t* __restrict__ coltot) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= 2 * k) return;
  uint32_t run = 0;
  const int side = col / k, s = col % k;   // column (side, sample) lives at blkcnt[(b*2+side)*k + s]
  for (int64_t b = 0; b < nblk; ++b) {
    uint32_t* p = &blkcnt[(b * 2 + side) * k + s];
    const uint32_t v = *p; *p = run; run += v;
  }
  coltot[col] = run;
}
__global__ void yd_colbase_kernel(const uint32_t* __restrict__ coltot, int k, unsigned long long* __restrict__ colbase) {
  unsigned long long run = 0;
  for (int c = 0; c < 2 * k; ++c) { colbase[c] = run; run += coltot[c]; }
  colbase[2 * k] = run;
}

// Y4: stable scatter of the group ids into the chain lists (chain c = side*k + sample)
__global__ void __launch_bounds__(128) yd_scatter_kernel(const uint32_t* __restrict__ bits, uint32_t W, int k, int64_t G, const uint8_t* __restrict__ gstrand,
                                                        const uint32_t* __restrict__ blkoff, const unsigned long long* __restrict__ colbase, int64_t nblk,
                                                        uint32_t* __restrict__ chain) {
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= nblk * W) return;
  const int64_t b = warp / W; const uint32_t w = (uint32_t)(warp % W);
  const int lane = tb_lane();
  const int s = (int)(w * 32 + lane);
  const int64_t g0 = b * YD_BLOCK, g1 = (g0 + YD_BLOCK < G) ? g0 + YD_BLOCK : G;
  unsigned long long pf = 0, pr = 0;
  if (s < k) { pf = colbase[s] + blkoff[(b * 2 + 0) * k + s]; pr = colbase[k + s] + blkoff[(b * 2 + 1) * k + s]; }
  for (int64_t base = g0; base < g1; base += 32) {
    const int64_t g = base + lane;
    uint32_t word = 0; uint8_t sc = 0;
    if (g < g1) { word = bits[(uint64_t)g * W + w]; sc = gstrand[g]; }
    unsigned any = __ballot_sync(0xffffffffu, word != 0);
    while (any) {
      const int q = __ffs(any) - 1; any &= any - 1;
      const uint32_t wq = __shfl_sync(0xffffffffu, word, q);
      const int sq = __shfl_sync(0xffffffffu, (int)sc, q);
      if ((wq >> lane) & 1u) {
        if (sq != '-') chain[pf++] = (uint32_t)(base + q);
        if (sq != '+') chain[pr++] = (uint32_t)(base + q);
      }
    }
  }
}

// Y5: sub-chain heads. Prefix max of (chain<<32 | end) over the member array is a segmented prefix max because chain
// ids ascend along the array.
struct MemberKeyIn {
  const uint32_t* chain; const uint32_t* gend; const unsigned long long* colbase; int nchains;
  __device__ int chain_of(int64_t i) const {
    int lo = 0, hi = nchains;  // last c with colbase[c] <= i
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (colbase[mid] <= (unsigned long long)i) lo = mid; else hi = mid; }
    return lo;
  }
  __device__ unsigned long long operator()(int64_t i) const { return ((unsigned long long)(uint32_t)chain_of(i) << 32) | gend[chain[i]]; }
};
struct MemberHeadOut {
  MemberKeyIn mk; const GDesc* desc; uint32_t* flag;
  __device__ void operator()(int64_t i, unsigned long long exc, unsigned long long) const {
    uint32_t h = 1u;
    if (i > 0 && (int)(exc >> 32) == mk.chain_of(i)) h = (uint32_t)desc[mk.chain[i]].start > (uint32_t)exc ? 1u : 0u;
    flag[i] = h;
  }
};
struct FlagIn { const uint32_t* f; __device__ uint32_t operator()(int64_t i) const { return f[i]; } };
struct HeadListOut { uint32_t* heads; __device__ void operator()(int64_t i, uint32_t exc, uint32_t inc) const { if (inc != exc) heads[exc] = (uint32_t)i; } };
__global__ void yd_subchain_total_kernel(const uint32_t* tot, uint32_t* heads, uint32_t n_members, unsigned long long* work) {
  heads[*tot] = n_members; work[0] = 0; work[1] = *tot;
}

// one segment list; node q lives at st[q*32], en[q*32] (the lane offset is folded into the pointers)
struct SegArr {
  int n; uint32_t last_pos; int last_dist; int cap;
  uint32_t* st; uint32_t* en;
  __device__ __forceinline__ uint32_t& S(int q) { return st[q << 5]; }
  __device__ __forceinline__ uint32_t& E(int q) { return en[q << 5]; }
  __device__ void reset() { n = 0; last_pos = 0; last_dist = -1; }
  __device__ void erase(int a, int b) {  // remove [a,b)
    const int d = b - a; if (d <= 0) return;
    for (int q = b; q < n; ++q) { S(q - d) = S(q); E(q - d) = E(q); }
    n -= d;
  }
  __device__ bool insert(int at, uint32_t s, uint32_t e) {
    if (n >= cap) return false;
    for (int q = n; q > at; --q) { S(q) = S(q - 1); E(q) = E(q - 1); }
    S(at) = s; E(at) = e; ++n; return true;
  }
  // mergeRead :167-219 over an exon source; returns false on capacity overflow
  template <class Src>
  __device__ bool merge(Src& src) {
    int s, e;
    if (n == 0) {
      while (src.next(s, e)) { if (n >= cap) return false; S(n) = (uint32_t)s; E(n) = (uint32_t)e; ++n; }
      return true;
    }
    int cur = 0;
    while (src.next(s, e)) {
      const uint32_t es = (uint32_t)s, ee = (uint32_t)e;
      while (cur < n) {
        if (ee < S(cur)) { if (!insert(cur, es, ee)) return false; ++cur; break; }  // inserted before the cursor node, which is now at cur
        if (es <= E(cur)) {
          if (es < S(cur)) S(cur) = es;
          if (ee > E(cur)) E(cur) = ee;
          const int nx = cur + 1;
          while (nx < n && S(nx) <= E(cur)) {
            const uint32_t nend = E(nx);
            erase(nx, nx + 1);
            if (nend > E(cur)) { E(cur) = nend; break; }
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
    for (int q = 0; q < n && S(q) < rstart; ++q) prev = q;
    if (prev >= 0) {
      if (E(prev) >= rstart) d = (int)(rstart - S(prev));
      if (d == 0) erase(0, prev + 1);
      else erase(0, prev);  // eager drop of the inert nodes before prev
    }
    last_pos = rstart; last_dist = d;
    ok = merge(src) && ok;
    return d;
  }
};

struct InlineExons {  // exon source over a descriptor
  int start, ne, i; int e0, s1, e1, s2, e2;
  __device__ bool next(int& s, int& e) {
    if (i >= ne) return false;
    if (i == 0) { s = start; e = e0; } else if (i == 1) { s = s1; e = e1; } else { s = s2; e = e2; }
    ++i; return true;
  }
};

__device__ __forceinline__ int yd_step(const ColIn& in, SegArr& L, const GDesc& d, bool& ok) {
  const int ne = (int)(d.meta & 0xffffu);
  if (ne <= YD_INLINE_EX) { InlineExons src{d.start, ne, 0, d.e[0], d.e[1], d.e[2], d.e[3], d.e[4]}; return L.process(src, (uint32_t)d.start, ok); }
  const uint32_t r = (uint32_t)d.e[5];
  ExonIter it; it.init(in.cigar, in.cig_off[r], in.cig_off[r + 1], d.start - 1);
  return L.process(it, (uint32_t)d.start, ok);
}

constexpr int YD_WARPS = 8;
constexpr int YD_CAP_SMEM = 24;      // live nodes per list in shared memory
constexpr int YD_CAP_GLOBAL = 4096;  // live nodes per list in the global-memory repeat
constexpr uint32_t YD_LONG = 96;     // sub-chains longer than this are walked by the whole warp

template <bool SMEM>
__global__ void __launch_bounds__(YD_WARPS * 32) yd_chain_kernel(ColIn in, const GDesc* __restrict__ desc, const uint32_t* __restrict__ chain,
                                                                const uint32_t* __restrict__ heads, unsigned long long* work, int32_t* __restrict__ yd,
                                                                uint32_t* gscratch, long long* status) {
  extern __shared__ __align__(16) uint32_t s_dyn[];   // SMEM: YD_WARPS * 2 * YD_CAP_SMEM * 32 words of lists
  __shared__ GDesc s_desc[YD_WARPS][32];
  __shared__ int s_d[YD_WARPS][32];
  const int wl = tb_warp(), lane = tb_lane();
  const int cap = SMEM ? YD_CAP_SMEM : YD_CAP_GLOBAL;
  uint32_t* base = SMEM ? &s_dyn[(size_t)wl * 2 * YD_CAP_SMEM * 32] : gscratch + ((size_t)blockIdx.x * YD_WARPS + wl) * 2 * (size_t)YD_CAP_GLOBAL * 32;
  SegArr L; L.cap = cap; L.st = base + lane; L.en = base + (size_t)cap * 32 + lane;
  SegArr L0; L0.cap = cap; L0.st = base; L0.en = base + (size_t)cap * 32;   // lane 0's list, used by the cooperative walk
  const unsigned long long nsub = work[1];
  bool ok = true;
  for (;;) {
    unsigned long long j0 = 0;
    if (lane == 0) j0 = atomicAdd(&work[0], 32ULL);
    j0 = __shfl_sync(0xffffffffu, j0, 0);
    if (j0 >= nsub) break;
    const unsigned long long j = j0 + lane;
    uint32_t c0 = 0, c1 = 0;
    if (j < nsub) { c0 = heads[j]; c1 = heads[j + 1]; }
    const uint32_t len = c1 - c0;
    // ---- short sub-chains: one per lane ----
    if (len > 0 && len <= YD_LONG) {
      L.reset();
      for (uint32_t i = c0; i < c1; ++i) {
        const uint32_t g = chain[i];
        const GDesc d = desc[g];
        const int dist = yd_step(in, L, d, ok);
        if (dist > 0) atomicMax(&yd[g], dist);
      }
    }
    // ---- long sub-chains: the whole warp walks them one after the other ----
    unsigned longs = __ballot_sync(0xffffffffu, len > YD_LONG);
    while (longs) {
      const int q = __ffs(longs) - 1; longs &= longs - 1;
      const uint32_t a = __shfl_sync(0xffffffffu, c0, q), b = __shfl_sync(0xffffffffu, c1, q);
      L0.reset();
      uint32_t gnext = 0; GDesc dnext; dnext.start = 0; dnext.meta = 0;
      if (a + lane < b) { gnext = chain[a + lane]; dnext = desc[gnext]; }
      for (uint32_t cb = a; cb < b; cb += 32) {
        const uint32_t g = gnext;
        s_desc[wl][lane] = dnext;
        __syncwarp();
        if (cb + 32 + lane < b) { gnext = chain[cb + 32 + lane]; dnext = desc[gnext]; }   // prefetch the next batch
        if (lane == 0) {
          const int cnt = (int)((b - cb) < 32u ? (b - cb) : 32u);
          for (int t = 0; t < cnt; ++t) s_d[wl][t] = yd_step(in, L0, s_desc[wl][t], ok);
        }
        __syncwarp();
        if (cb + lane < b) { const int dist = s_d[wl][lane]; if (dist > 0) atomicMax(&yd[g], dist); }
        __syncwarp();
      }
    }
  }
  if (!ok) status[CS_YD_OVERFLOW] = 1;
}

__global__ void __launch_bounds__(256) yd_combine_kernel(int32_t* __restrict__ yd, const int32_t* __restrict__ ydc, int64_t G) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < G) { const int32_t a = yd[g], b = ydc[g]; yd[g] = a > b ? a : b; }   // max(carried YD tags, chain distance); 0 = no tag
}

}  // namespace

int col_yd(tb_ctx* ctx, const ColIn& in, const ColGeom& g, const ColGroups& grp, int64_t G) {
  cudaStream_t st = ctx->stream;
  DevBuf* B = ctx->buf;
  const uint32_t W = g.W; const int k = g.k;
  const int64_t nblk = (G + YD_BLOCK - 1) / YD_BLOCK;
  TB_CUDA(B[XB_GDESC].ensure(sizeof(GDesc) * (size_t)G));
  TB_CUDA(B[XB_YDPM].ensure(sizeof(uint32_t) * (size_t)G + (size_t)G + 64));
  TB_CUDA(B[XB_YDBLK].ensure(sizeof(uint32_t) * (size_t)nblk * 2 * k + sizeof(uint32_t) * 2 * k + sizeof(uint64_t) * (2 * (size_t)k + 8)));
  TB_CUDA(B[XB_WORK].ensure(256));
  TB_CUDA(B[XB_YDC].ensure(sizeof(int32_t) * (size_t)G));
  GDesc* desc = B[XB_GDESC].as<GDesc>();
  uint32_t* gend = B[XB_YDPM].as<uint32_t>();
  uint8_t* gstrand = (uint8_t*)(gend + G);
  unsigned long long* colbase = B[XB_YDBLK].as<unsigned long long>();       // [2k+1] (+pad), 8-byte aligned at the buffer start
  uint32_t* coltot = (uint32_t*)(colbase + 2 * (size_t)k + 8);                // [2k]
  uint32_t* blkcnt = coltot + 2 * k;                                          // [nblk*2*k]
  unsigned long long* work = B[XB_WORK].as<unsigned long long>() + 8;         // the tile kernel's slot counter sits at +0
  int32_t* ydc = B[XB_YDC].as<int32_t>();
  long long* h_status = ctx->pinned[0].as<long long>();
  TB_CUDA(cudaMemsetAsync(ydc, 0, sizeof(int32_t) * (size_t)G, st));
  yd_desc_kernel<<<tb_grid_for(G, 256), 256, 0, st>>>(in, grp.rep, G, desc, gend, gstrand);
  const int64_t nwarps = nblk * W;
  yd_count_kernel<<<tb_grid_for(nwarps * 32, 128), 128, 0, st>>>(grp.bits, W, k, G, gstrand, blkcnt, nblk);
  yd_prefix_kernel<<<tb_grid_for(2 * k, 128), 128, 0, st>>>(blkcnt, nblk, k, coltot);
  yd_colbase_kernel<<<1, 1, 0, st>>>(coltot, k, colbase);
  ctx->launches += 4;
  // the number of chain members (<= 2 x sum of direct-sample counts) is only known on the device
  TB_CUDA(cudaMemcpyAsync(h_status, colbase + 2 * k, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaStreamSynchronize(st));
  const int64_t n_members = h_status[0];
  if (n_members >= (1LL << 32)) { ctx->set_error("tb_collapse_window: %lld chain members exceed the 32-bit YD index", (long long)n_members); return 1; }
  if (n_members > 0) {
    TB_CUDA(B[XB_YDCHAIN].ensure(sizeof(uint32_t) * ((size_t)n_members + 32)));
    TB_CUDA(B[XB_BHEAD].ensure(sizeof(uint32_t) * ((size_t)n_members + 32)));
    TB_CUDA(B[XB_YDFLAG].ensure(sizeof(uint32_t) * ((size_t)n_members + 32)));
    TB_CUDA(B[XB_AGG].ensure((size_t)(tb_scan_blocks(n_members) + 8) * sizeof(uint64_t)));
    uint32_t* chain = B[XB_YDCHAIN].as<uint32_t>(); uint32_t* heads = B[XB_BHEAD].as<uint32_t>(); uint32_t* flag = B[XB_YDFLAG].as<uint32_t>();
    yd_scatter_kernel<<<tb_grid_for(nwarps * 32, 128), 128, 0, st>>>(grp.bits, W, k, G, gstrand, blkcnt, colbase, nblk, chain);
    ctx->launches++;
    MemberKeyIn mk{chain, gend, colbase, 2 * k};
    TB_CUDA((tb_device_scan<OpMaxU64>(ctx, mk, n_members, B[XB_AGG].as<unsigned long long>(), MemberHeadOut{mk, desc, flag})));
    TB_CUDA((tb_device_scan<OpSumU32>(ctx, FlagIn{flag}, n_members, B[XB_AGG].as<uint32_t>(), HeadListOut{heads})));
    yd_subchain_total_kernel<<<1, 1, 0, st>>>(B[XB_AGG].as<uint32_t>() + tb_scan_blocks(n_members), heads, (uint32_t)n_members, work);
    ctx->launches++;
    const size_t smem = (size_t)YD_WARPS * 2 * YD_CAP_SMEM * 32 * sizeof(uint32_t);
    TB_CUDA(cudaFuncSetAttribute(yd_chain_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    yd_chain_kernel<true><<<(unsigned)ctx->sm_count * 4, YD_WARPS * 32, smem, st>>>(in, desc, chain, heads, work, ydc, nullptr, g.d_status);
    ctx->launches++;
    TB_CUDA(cudaMemcpyAsync(h_status, g.d_status, sizeof(int64_t) * 16, cudaMemcpyDeviceToHost, st));
    TB_CUDA(cudaStreamSynchronize(st));
    if (h_status[CS_YD_OVERFLOW]) {
      // a list outgrew shared memory (its later distances are wrong): repeat everything with lists in global memory
      const unsigned grid2 = (unsigned)ctx->sm_count;
      TB_CUDA(B[XB_YDSCRATCH].ensure((size_t)grid2 * YD_WARPS * 2 * YD_CAP_GLOBAL * 32 * sizeof(uint32_t)));
      TB_CUDA(cudaMemsetAsync(ydc, 0, sizeof(int32_t) * (size_t)G, st));
      TB_CUDA(cudaMemsetAsync(work, 0, sizeof(unsigned long long), st));
      TB_CUDA(cudaMemsetAsync(g.d_status + CS_YD_OVERFLOW, 0, sizeof(long long), st));
      yd_chain_kernel<false><<<grid2, YD_WARPS * 32, 0, st>>>(in, desc, chain, heads, work, ydc, B[XB_YDSCRATCH].as<uint32_t>(), g.d_status);
      ctx->launches++;
      TB_CUDA(cudaMemcpyAsync(h_status, g.d_status, sizeof(int64_t) * 16, cudaMemcpyDeviceToHost, st));
      TB_CUDA(cudaStreamSynchronize(st));
      if (h_status[CS_YD_OVERFLOW]) { ctx->set_error("tb_collapse_window: YD segment list exceeded %d live nodes", YD_CAP_GLOBAL); return 1; }
    }
  }
  yd_combine_kernel<<<tb_grid_for(G, 256), 256, 0, st>>>(grp.yd, ydc, G);
  ctx->launches++;
  return 0;
}
