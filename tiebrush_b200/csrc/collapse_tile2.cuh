// collapse_tile2.cuh — second-generation tile kernel of the collapse front end (included inside collapse_tile.cu's
// anonymous namespace). Same contract as col_tile_kernel: one slot = the start positions whose first merged rank falls into
// [mT,(m+1)T); groups of the slot go to the staging arrays at the slot's first merged rank, in final order
// (reference: src/tmerge.h:28-50 merge order, src/tiebrush.cpp:438-499 grouping, :501-530 output order).
//
// What is different from col_tile_kernel (profiles/r01f: 29 warp-instructions per record, 7 % DRAM throughput,
// every key verification two dependent L2 round trips):
//   * the slot's k file slices of pos / cig_off and their CIGAR words are brought into shared memory by the TMA unit
//     (cp.async.bulk 1-D copies issued by one thread per file, completion counted on an mbarrier); the copies of slot
//     i+1 are in flight while slot i's epilogue runs. The dependent chain pos -> cig_off -> cigar -> owner's cig_off ->
//     owner's cigar is shared-memory traffic only.
//   * ONE table per slot, hashed on (position, strand, end, mode key), sized for the groups expected (not for "every
//     record distinct"): slots are 2-3x larger for the same shared memory. The position order comes from a counting
//     sort of the occupied entries by the merged rank of their position (P[pos] - rank0 < T).
//   * representative = 32-bit (running max end - own end) << 13 | index in the slot's file-major list: native 32-bit
//     shared atomicMin; the 64-bit sort key is stored by the thread that inserts the group.
//   * anything irregular is DEFERRED to col_tile_kernel's full-size-table launch through the heavy list: a pile-up
//     position (slot > T records), CIGAR words beyond the staging arena, more groups than the table holds, a running
//     maximum more than 2^19 beyond the read's own end. Results are identical by construction (same staging arrays).

struct Tile2Params {
  uint32_t M, E, logE, W, T, rs_cap, cw_cap, k;
  const uint32_t* P; const uint32_t* slotpos; const uint32_t* off;
  uint32_t* gcount; uint32_t* st_rep; float* st_yc; uint32_t* st_yx; uint32_t* st_bits;
  long long* status; unsigned int* slot_counter; uint32_t seed;
  uint32_t* heavy_list;
  uint32_t n, n_cig;
};

// ---- mbarrier / bulk-copy wrappers (PTX ISA 8.x, sm_90+; SASS: SYNCS.*, UBLKCP) ----
__device__ __forceinline__ uint32_t t2_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void t2_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(t2_smem(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void t2_mbar_arrive_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(t2_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void t2_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(t2_smem(bar)) : "memory");
}
__device__ __forceinline__ bool t2_mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(t2_smem(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void t2_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(t2_smem(dst)), "l"(src), "r"(bytes), "r"(t2_smem(bar)) : "memory");
}
__device__ __forceinline__ void t2_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

static size_t tile2_smem_bytes(uint32_t k, uint32_t E, uint32_t W, uint32_t T, uint32_t rs_cap, uint32_t cw_cap) {
  return (size_t)rs_cap * 8 + (size_t)cw_cap * 4 + (size_t)E * (8 + 8 + 4 + 4 + 4 * (size_t)W + 2 + 2 + 2 + 2) + (size_t)(T + 4) * 4 +
         2 * (size_t)(4 * k + 4) * 4 + 128;
}

enum { T2_READY = 0, T2_EMPTY = 1, T2_DEFER = 2, T2_END = 3 };
struct Slot2 { uint32_t m, rank0, n_t, state, p0, p1; };

struct Tile2Smem {
  uint32_t* s_pos; uint32_t* s_co; uint32_t* s_cig;   // staged slices (TMA destinations, 16-byte aligned)
  unsigned long long* word;   // [E] tag32 << 32 | owner's shared CIGAR index << 16 | owner's staging slot
  unsigned long long* skey;   // [E] 64-bit sort key of the group (strand | ref_len | key length | first key bytes)
  uint32_t* rep;              // [E] streaming: (Erel - ref_len) << 13 | j ; epilogue: window index of the representative
  uint32_t* cnt;              // [E]
  uint32_t* bits;             // [E*W]
  uint32_t* pc;               // [T+2] epilogue: groups per position (by merged rank of the position), then their exclusive prefix
  uint32_t* fgeo;             // per-file geometry, double buffered: [2][4][k+1] = a | soff | radj | cadj
  uint16_t* prank;            // [E] merged rank of the entry's position inside the slot
  uint16_t* occ;              // [E] occupied entries
  uint16_t* ord;              // [E] occupied entries in position order
  uint16_t* tmp;              // [E] local index inside the position, then rank | saturated(14) | tied(15)
};

__device__ __forceinline__ uint32_t t2_fold(uint32_t h, uint32_t w) { h = (h ^ w) * 0x9E3779B1u; return h ^ (h >> 15); }

// one pass over the staged CIGAR of a record: reference length, running hash of the mode key, and the two variable fields
// of the 64-bit sort key (key length saturated to 8 bits, first three key bytes in memcmp order / first exon length)
template <int MODE>
__device__ __forceinline__ void tile2_parse(const ColIn& in, const uint32_t* cg, uint32_t c0, uint32_t c1, int pos, uint32_t gi,
                                            int& reflen, uint32_t& h, uint32_t& len8, uint32_t& w24) {
  if (MODE == TB_MODE_EXON) {
    ExonIter it; it.init(cg, c0, c1, pos);
    int s, e, nex = 0, e0 = 0; h = 0x1234567u;
    while (it.next(s, e)) { if (nex == 0) e0 = e; h = t2_fold(h, (uint32_t)s); h = t2_fold(h, (uint32_t)e); ++nex; }
    h = t2_fold(h, (uint32_t)nex);
    reflen = it.l;
    len8 = nex > 255 ? 255u : (uint32_t)nex;
    const uint32_t d = (uint32_t)(e0 - pos);
    w24 = d > 0xffffffu ? 0xffffffu : d;
    return;
  }
  uint32_t a = c0, b = c1;
  if (MODE == TB_MODE_CLIP) tb_clip_range(cg, a, b);
  const uint32_t nc = b - a;
  uint32_t hh = 0x9876543u ^ nc;
  int l = 0;
  for (uint32_t c = c0; c < c1; ++c) {
    const uint32_t w = cg[c];
    if (tb_op_ref(w & 0xf)) l += (int)(w >> 4);
    if (MODE != TB_MODE_CLIP || (c >= a && c < b)) hh = t2_fold(hh, w);
  }
  if (MODE == TB_MODE_FULL) {
    const uint32_t m0 = in.md_off[gi], m1 = in.md_off[gi + 1];
    hh = t2_fold(hh, (uint32_t)(m1 > m0));
    uint32_t acc = 0, nb = 0;
    for (uint32_t q = m0; q < m1; ++q) {
      const uint32_t ch = in.md[q];
      if (ch == 0) break;
      acc = (acc << 8) | ch;
      if (++nb == 4) { hh = t2_fold(hh, acc); acc = 0; nb = 0; }
    }
    hh = t2_fold(hh, acc ^ (nb << 29));
  }
  reflen = l; h = hh;
  len8 = nc > 255 ? 255u : nc;
  w24 = nc ? (__byte_perm(cg[a], 0, 0x0123) >> 8) : 0u;
}

// staging slot -> window index (binary search over the per-file slot bases; used off the hot path only)
__device__ __noinline__ uint32_t tile2_slot_to_global(const uint32_t* fa, const uint32_t* fsoff, const uint32_t* fradj, uint32_t k, uint32_t rs) {
  uint32_t lo = 0, hi = k;   // last f with slot base <= rs; slot base of f = fradj[f] - (fa[f] & 3) + fsoff[f]
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (fradj[mid] - (fa[mid] & 3u) + fsoff[mid] <= rs) lo = mid; else hi = mid;
  }
  return fa[lo] + (rs - fradj[lo] - fsoff[lo]);
}
__device__ __forceinline__ uint32_t tile2_j_to_global(const uint32_t* fa, const uint32_t* fsoff, uint32_t k, uint32_t j) {
  uint32_t lo = 0, hi = k;   // last f with fsoff[f] <= j (empty files share their successor's offset: the last one holds j)
  while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (fsoff[mid] <= j) lo = mid; else hi = mid; }
  return fa[lo] + (j - fsoff[lo]);
}

template <int THREADS, int MODE>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS) col_tile2_kernel(ColIn in_, Tile2Params tp) {
  constexpr int NW = THREADS / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint32_t s_scan[33];
  __shared__ uint32_t s_wtot[4][3];
  __shared__ uint32_t s_flags[4];          // [0] kept  [1] overflow  [2] groups inserted  [3] overflow that more passes cannot cure
  __shared__ uint32_t s_range[2];          // positions of the current pass, relative to pos_lo
  __shared__ Slot2 s_slot[2];
  __shared__ __align__(8) uint64_t s_bar;
  volatile uint32_t* vflags = s_flags;
  ColIn in = in_;
  in.mode = MODE;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t E = tp.E, W = tp.W, k = tp.k, emask = tp.E - 1;
  Tile2Smem sm;
  {
    unsigned char* p = smem_raw;
    sm.s_pos = (uint32_t*)p; p += (size_t)tp.rs_cap * 4;
    sm.s_co = (uint32_t*)p; p += (size_t)tp.rs_cap * 4;
    sm.s_cig = (uint32_t*)p; p += (size_t)tp.cw_cap * 4;
    sm.word = (unsigned long long*)p; p += (size_t)E * 8;
    sm.skey = (unsigned long long*)p; p += (size_t)E * 8;
    sm.rep = (uint32_t*)p; p += (size_t)E * 4;
    sm.cnt = (uint32_t*)p; p += (size_t)E * 4;
    sm.bits = (uint32_t*)p; p += (size_t)E * W * 4;
    sm.pc = (uint32_t*)p; p += (size_t)(tp.T + 4) * 4;
    sm.fgeo = (uint32_t*)p; p += (size_t)8 * (k + 1) * 4;
    sm.prank = (uint16_t*)p; p += (size_t)E * 2;
    sm.occ = (uint16_t*)p; p += (size_t)E * 2;
    sm.ord = (uint16_t*)p; p += (size_t)E * 2;
    sm.tmp = (uint16_t*)p;
  }
  if (tid == 0) {
    t2_mbar_init(&s_bar, k);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t phase = 0;        // parity of the mbarrier phase the next READY slot completes
  uint32_t kept_total = 0;   // thread 0 only

  // claim the next slot (one thread) ...
  auto claim = [&](Slot2& d) {
    Slot2 t; t.m = atomicAdd(tp.slot_counter, 1u); t.rank0 = 0; t.n_t = 0; t.state = T2_END; t.p0 = t.p1 = 0;
    if (t.m < tp.M) {
      const uint32_t p0 = tp.slotpos[t.m], p1 = tp.slotpos[t.m + 1];
      t.p0 = p0; t.p1 = p1;
      t.rank0 = tp.P[p0];
      t.n_t = tp.P[p1] - t.rank0;
      t.state = t.n_t == 0 ? T2_EMPTY : (t.n_t > tp.T ? T2_DEFER : T2_READY);
    }
    d = t;
  };
  // ... and start its copies (all threads; the staging area must be free). Leaves the per-file geometry in buffer `nb`.
  auto produce = [&](Slot2& d, int nb) {
    __syncthreads();                       // d is visible; every reader of the staging area is done
    const Slot2 s = d;
    if (s.state != T2_READY) return;       // block-uniform
    uint32_t a = 0, b = 0, ca = 0, cb = 0;
    const bool isf = tid < k;
    if (isf) {
      a = tp.off[(uint64_t)s.m * k + tid]; b = tp.off[(uint64_t)(s.m + 1) * k + tid];
      if (b > a) { ca = in.cig_off[a]; cb = in.cig_off[b]; }
    }
    const uint32_t len = b - a;
    const uint32_t a4 = a & ~3u, ca4 = ca & ~3u;
    const uint32_t nrs = len ? (((b + 4u) & ~3u) - a4) : 0u;      // slots covering records a..b (cig_off needs entry b too)
    const uint32_t ncw = len ? (((cb + 3u) & ~3u) - ca4) : 0u;
    // exclusive prefixes over the files of (len, nrs, ncw): the first four warps hold the k <= 128 files
    uint32_t x0 = len, x1 = nrs, x2 = ncw;
    if (warp < 4) {
#pragma unroll
      for (int dd = 1; dd < 32; dd <<= 1) {
        const uint32_t y0 = __shfl_up_sync(0xffffffffu, x0, dd), y1 = __shfl_up_sync(0xffffffffu, x1, dd), y2 = __shfl_up_sync(0xffffffffu, x2, dd);
        if ((int)lane >= dd) { x0 += y0; x1 += y1; x2 += y2; }
      }
      if (lane == 31) { s_wtot[warp][0] = x0; s_wtot[warp][1] = x1; s_wtot[warp][2] = x2; }
    }
    __syncthreads();
    uint32_t p0 = 0, p1 = 0, p2 = 0, t1 = 0, t2 = 0;
#pragma unroll
    for (uint32_t w = 0; w < 4; ++w) {
      if (w < warp) { p0 += s_wtot[w][0]; p1 += s_wtot[w][1]; p2 += s_wtot[w][2]; }
      t1 += s_wtot[w][1]; t2 += s_wtot[w][2];
    }
    if (t1 > tp.rs_cap || t2 > tp.cw_cap) {   // block-uniform: the slot does not fit the staging area
      __syncthreads();                        // everybody has read s_wtot and d
      if (tid == 0) d.state = T2_DEFER;
      return;
    }
    if (isf) {
      const uint32_t e0 = x0 - len + p0, rbase = x1 - nrs + p1, cbase = x2 - ncw + p2;
      uint32_t* ga = sm.fgeo + (size_t)nb * 4 * (k + 1);
      ga[tid] = a; ga[(k + 1) + tid] = e0;
      ga[2 * (k + 1) + tid] = rbase + (a & 3u) - e0;      // staging slot of compact index j: fradj + j
      ga[3 * (k + 1) + tid] = cbase + (ca & 3u) - ca;     // shared CIGAR index of window CIGAR offset c: fcadj + c
      if (tid == k - 1) ga[(k + 1) + k] = e0 + len;
      if (len) {
        // 16-byte chunks through the TMA unit; a chunk that would run past the end of its column is copied by hand
        const uint32_t pe = (b + 3u) & ~3u, plim = tp.n & ~3u, pb = pe < plim ? pe : plim;                 // pos: n entries
        const uint32_t oe = (b + 4u) & ~3u, olim = (tp.n + 1u) & ~3u, ob = oe < olim ? oe : olim;          // cig_off: n+1 entries
        const uint32_t we = (cb + 3u) & ~3u, wlim = tp.n_cig & ~3u, wb = we < wlim ? we : wlim;            // cigar: n_cig words
        for (uint32_t e = pb > a4 ? pb : a4; e < b; ++e) sm.s_pos[rbase + (e - a4)] = (uint32_t)in.pos[e];
        for (uint32_t e = ob > a4 ? ob : a4; e <= b; ++e) sm.s_co[rbase + (e - a4)] = in.cig_off[e];
        for (uint32_t e = wb > ca4 ? wb : ca4; e < cb; ++e) sm.s_cig[cbase + (e - ca4)] = in.cigar[e];
        const uint32_t by_p = pb > a4 ? (pb - a4) * 4u : 0u, by_o = ob > a4 ? (ob - a4) * 4u : 0u, by_w = wb > ca4 ? (wb - ca4) * 4u : 0u;
        t2_mbar_arrive_tx(&s_bar, by_p + by_o + by_w);
        if (by_p) t2_bulk_g2s(sm.s_pos + rbase, in.pos + a4, by_p, &s_bar);
        if (by_o) t2_bulk_g2s(sm.s_co + rbase, in.cig_off + a4, by_o, &s_bar);
        if (by_w) t2_bulk_g2s(sm.s_cig + cbase, in.cigar + ca4, by_w, &s_bar);
      } else {
        t2_mbar_arrive(&s_bar);
      }
    }
  };

  if (tid == 0) claim(s_slot[0]);
  produce(s_slot[0], 0);
  const uint32_t glimit = E - (E >> 3);
  for (uint32_t it = 0;; ++it) {
    const int cb_ = (int)(it & 1), nb_ = cb_ ^ 1;
    if (tid == 0) claim(s_slot[nb_]);        // its latency hides behind this slot's work
    __syncthreads();                         // s_slot[cb_].state is final (produce may have deferred it)
    const Slot2 cur = s_slot[cb_];
    if (cur.state == T2_END) break;
    if (cur.state != T2_READY) {
      if (tid == 0) {
        tp.gcount[cur.m] = 0;
        if (cur.state == T2_DEFER) { tp.heavy_list[atomicAdd((unsigned long long*)&tp.status[CS_NHEAVY], 1ULL)] = cur.m; atomicAdd((unsigned long long*)&tp.status[CS_T2_STAGE], 1ULL); }
      }
      produce(s_slot[nb_], nb_);
      continue;
    }
    const uint32_t n_t = cur.n_t, rank0 = cur.rank0;
    const uint32_t* fa = sm.fgeo + (size_t)cb_ * 4 * (k + 1); const uint32_t* fsoff = fa + (k + 1); const uint32_t* fradj = fsoff + (k + 1); const uint32_t* fcadj = fradj + (k + 1);
    // A slot whose groups outgrow the table (stretches of the genome where nearly every alignment is distinct) is redone in
    // npass passes over the staged records, each pass taking the positions of one quarter (eighth) of the slot's records:
    // a pass then holds at most n_t/npass records plus one position, i.e. fits unless that position is a pile-up.
    bool waited = false, produced = false, done = false;
    for (uint32_t npass = 1; !done; npass = npass == 1 ? 4u : npass * 2u) {
      if (npass > 8u) {   // block-uniform: a pile-up of distinct alignments at one position -> full-size table launch
        if (tid == 0) { tp.gcount[cur.m] = 0; tp.heavy_list[atomicAdd((unsigned long long*)&tp.status[CS_NHEAVY], 1ULL)] = cur.m; atomicAdd((unsigned long long*)&tp.status[CS_T2_TABLE], 1ULL); }
        break;
      }
      if (npass == 4u && tid == 0) atomicAdd((unsigned long long*)&tp.status[CS_T2_MULTI], 1ULL);
      uint32_t gbase = 0, kept_acc = 0;
      bool ovf = false;
      for (uint32_t ps = 0; ps < npass && !ovf; ++ps) {
        // ---- position range of this pass (relative to pos_lo): [pl, ph) ----
        __syncthreads();   // every thread has read the flags of the previous pass / attempt
        if (tid == 0) {
          uint32_t b2[2];
          for (int q = 0; q < 2; ++q) {
            const uint32_t target = (uint32_t)(((unsigned long long)n_t * (ps + q)) / npass);
            uint32_t lo = cur.p0, hi = cur.p1;   // first p in [p0,p1] with P[p]-rank0 >= target
            if (ps + q == 0) hi = lo; else if (ps + q == npass) lo = hi;
            while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if (tp.P[mid] - rank0 >= target) hi = mid; else lo = mid + 1; }
            b2[q] = lo;
          }
          s_range[0] = b2[0]; s_range[1] = b2[1];
          s_flags[0] = 0; s_flags[1] = 0; s_flags[2] = 0; s_flags[3] = 0;
        }
        // ---- clear the table (while the copies land) ----
        for (uint32_t s = tid; s < E; s += THREADS) { sm.word[s] = EMPTY64; sm.rep[s] = 0xffffffffu; sm.cnt[s] = 0; }
        {
          uint4* b4 = reinterpret_cast<uint4*>(sm.bits);
          const uint32_t n4 = (E * W) >> 2;   // E is a multiple of 4
          for (uint32_t s = tid; s < n4; s += THREADS) b4[s] = make_uint4(0u, 0u, 0u, 0u);
        }
        if (!waited) { while (!t2_mbar_try_wait(&s_bar, phase)) {} phase ^= 1u; waited = true; }
        __syncthreads();
        const uint32_t pl = s_range[0], pspan = s_range[1] - s_range[0];

        // ---- stream this warp's contiguous range of the slot's file-major record list ----
        const uint32_t j0 = (uint32_t)(((unsigned long long)n_t * warp) / NW), j1 = (uint32_t)(((unsigned long long)n_t * (warp + 1)) / NW);
        uint32_t carry_f = 0xffffffffu; int carry_pos = INT_MIN; int carry_max = 0;
        uint32_t my_kept = 0, f_lo = 0;
        if (j0 < j1) {
          uint32_t lo = 0, hi = k;
          while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (fsoff[mid] <= j0) lo = mid; else hi = mid; }
          f_lo = lo;
          if (j0 > fsoff[f_lo]) {   // the range starts inside a file slice: running maximum of the (file, position) run before it
            const uint32_t rs0 = fradj[f_lo] + j0, cad = fcadj[f_lo];
            const uint32_t nback = j0 - fsoff[f_lo];
            const int p = (int)sm.s_pos[rs0];
            if ((int)sm.s_pos[rs0 - 1] == p) {
              int mx = 0;
              for (uint32_t dn = 0; dn < nback; dn += 32) {
                const bool ok = dn + lane < nback;
                const uint32_t rs = ok ? rs0 - 1 - dn - lane : rs0;
                const bool same = ok && (int)sm.s_pos[rs] == p;
                const int rl = same ? tb_ref_len(sm.s_cig, cad + sm.s_co[rs], cad + sm.s_co[rs + 1]) : 0;
                const unsigned notsame = __ballot_sync(0xffffffffu, !same);
                const unsigned first = notsame ? (unsigned)(__ffs(notsame) - 1) : 32u;
                if (lane < first) mx = max(mx, rl);
                if (first < 32u) break;
              }
#pragma unroll
              for (int d = 16; d > 0; d >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
              carry_f = f_lo; carry_pos = p; carry_max = mx;
            }
          }
        }
        for (uint32_t jb = j0; jb < j1; jb += 32) {
          if (__any_sync(0xffffffffu, vflags[1] != 0)) break;   // warp-uniform: some record of the pass could not be placed
          const uint32_t j = jb + lane;
          const bool valid = j < j1;
          uint32_t f = 0xfffffffeu, rs = 0, cs0 = 0, h = 0, len8 = 0, w24 = 0, nc = 0;
          int pos = INT_MIN + 1 + (int)lane, reflen = 0; bool pass = false; unsigned sc = 0;
          if (valid) {
            f = f_lo;
            while (fsoff[f + 1] <= j) ++f;
            rs = fradj[f] + j;
            pos = (int)sm.s_pos[rs];
            if ((uint32_t)(pos - in.pos_lo) - pl < pspan) {   // a record of another pass only keeps its (file, position) run apart
              const uint32_t gi = fa[f] + (j - fsoff[f]);
              const uint16_t fl = in.flag[gi]; const uint8_t mq = in.mapq[gi]; const uint16_t nh = in.nh[gi];
              sc = tb_strand_code(in.strand[gi]);
              const uint32_t co = sm.s_co[rs], co1 = sm.s_co[rs + 1];
              cs0 = fcadj[f] + co; nc = co1 - co;
              tile2_parse<MODE>(in, sm.s_cig, cs0, cs0 + nc, pos, gi, reflen, h, len8, w24);
              pass = tb_passes_options(in, fl, mq, nh);
            }
          }
          // segmented inclusive max-scan of ref_len over (file, position) runs == the order of the reference's priority queue
          // (SURVEY §9.2); records the filters drop still take part
          uint32_t f_prev = __shfl_up_sync(0xffffffffu, f, 1); int pos_prev = __shfl_up_sync(0xffffffffu, pos, 1);
          if (lane == 0) { f_prev = carry_f; pos_prev = carry_pos; }
          int hd = (!valid || f != f_prev || pos != pos_prev) ? 1 : 0;
          int v = valid ? reflen : 0;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const int v2 = __shfl_up_sync(0xffffffffu, v, d), h2s = __shfl_up_sync(0xffffffffu, hd, d);
            if ((int)lane >= d) { if (!hd) v = max(v, v2); hd |= h2s; }
          }
          if (!hd) v = max(v, carry_max);
          const int Erel = v;
          carry_f = __shfl_sync(0xffffffffu, f, 31); carry_pos = __shfl_sync(0xffffffffu, pos, 31); carry_max = __shfl_sync(0xffffffffu, Erel, 31);
          f_lo = carry_f;   // only meaningful while lane 31 is valid, i.e. while another step follows

          uint32_t s = 0;
          if (pass) {
            ++my_kept;
            const uint32_t dE = (uint32_t)(Erel - reflen);
            if (dE >= (1u << 19)) { vflags[1] = 1; vflags[3] = 1; pass = false; }
            const uint32_t x = h ^ ((uint32_t)pos * 0x85EBCA77u) ^ ((uint32_t)reflen * 0xC2B2AE3Du) ^ tp.seed;
            const uint32_t tag = (sc << 30) | (tb_mix32(x) >> 2);
            s = (x * 0x9E3779B1u) >> (32u - tp.logE);
            bool placed = false;
            for (uint32_t t = 0; pass && t < E; ++t) {
              unsigned long long w = sm.word[s];
              if (w == EMPTY64) {
                if (vflags[2] > glimit) break;   // table (nearly) full: more passes
                const unsigned long long mine = ((unsigned long long)tag << 32) | ((unsigned long long)cs0 << 16) | rs;
                w = atomicCAS(&sm.word[s], EMPTY64, mine);
                if (w == EMPTY64) {
                  sm.skey[s] = ((unsigned long long)sc << 62) | ((unsigned long long)((uint32_t)reflen & 0x3fffffffu) << 32) | ((unsigned long long)len8 << 24) | w24;
                  sm.prank[s] = (uint16_t)(__ldg(&tp.P[(uint32_t)(pos - in.pos_lo)]) - rank0);
                  atomicAdd(&s_flags[2], 1u);
                  placed = true; break;
                }
              }
              if ((uint32_t)(w >> 32) == tag) {
                const uint32_t ors = (uint32_t)w & 0xffffu, ocs = ((uint32_t)w >> 16) & 0xffffu;
                if ((int)sm.s_pos[ors] == pos) {
                  const uint32_t onc = sm.s_co[ors + 1] - sm.s_co[ors];
                  bool same = onc == nc;
                  if (same) for (uint32_t q = 0; q < nc; ++q) if (sm.s_cig[cs0 + q] != sm.s_cig[ocs + q]) { same = false; break; }
                  if (MODE != TB_MODE_CIGAR && (MODE == TB_MODE_FULL || !same)) {   // identical raw CIGARs decide -P / -E; -L adds the MD bytes; the rest takes the exact comparator
                    const uint32_t gi = fa[f] + (j - fsoff[f]);
                    const uint32_t go = tile2_slot_to_global(fa, fsoff, fradj, k, ors);
                    same = tb_mode_cmp(in, gi, go) == 0;
                  }
                  if (same) { placed = true; break; }
                }
              }
              s = (s + 1u) & emask;
            }
            if (!placed) { vflags[1] = 1; pass = false; }
          }
          // merge the lanes of one file that hit one group: the lowest lane has the smallest (Erel, index) of them
          const uint32_t mkey = pass ? (s | (f << 11)) : (0x80000000u | lane);
          const unsigned peers = __match_any_sync(0xffffffffu, mkey);
          if (pass && (int)lane == __ffs(peers) - 1) {
            atomicAdd(&sm.cnt[s], (uint32_t)__popc(peers));
            const uint32_t repkey = ((uint32_t)(Erel - reflen) << 13) | j;
            if (repkey < sm.rep[s]) atomicMin(&sm.rep[s], repkey);
            const uint32_t bi = s * W + (f >> 5), bit = 1u << (f & 31);
            if (!(sm.bits[bi] & bit)) atomicOr(&sm.bits[bi], bit);
          }
        }
        if (my_kept) atomicAdd(&s_flags[0], my_kept);
        __syncthreads();                          // the table of the pass is complete
        ovf = vflags[1] != 0;                     // block-uniform
        if (ovf) {
          if (vflags[3]) npass = 8u;              // a running maximum beyond the 19-bit field: more passes cannot help
          break;
        }
        kept_acc += vflags[0];
        if (ps + 1 == npass) {
          // ---- nobody reads the staging area any more: the next slot's copies start now and land during the epilogue ----
          t2_fence_proxy_async();
          produce(s_slot[nb_], nb_);
          produced = true;
        }

        // ---- epilogue A: occupied entries; groups per position ----
        for (uint32_t p = tid; p <= n_t + 1; p += THREADS) sm.pc[p] = 0;
        uint32_t G = 0;
        {
          const uint32_t chunk = (((E + NW - 1) / NW) + 31u) & ~31u;
          const uint32_t e0 = warp * chunk, e1 = min(E, e0 + chunk);
          uint32_t mine = 0;
          for (uint32_t eb = e0; eb < e1; eb += 32) {
            const uint32_t e = eb + lane;
            mine += __popc(__ballot_sync(0xffffffffu, e < e1 && sm.word[e] != EMPTY64));
          }
          uint32_t tot;
          uint32_t base = tb_block_exscan<OpSumU32>(lane == 0 ? mine : 0u, s_scan, &tot);   // includes barriers: pc is cleared
          base = __shfl_sync(0xffffffffu, base, 0);
          for (uint32_t eb = e0; eb < e1; eb += 32) {
            const uint32_t e = eb + lane;
            const bool occf = e < e1 && sm.word[e] != EMPTY64;
            const unsigned bal = __ballot_sync(0xffffffffu, occf);
            if (occf) {
              sm.occ[base + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)e;
              sm.tmp[e] = (uint16_t)atomicAdd(&sm.pc[sm.prank[e]], 1u);   // arrival index inside the position
              sm.rep[e] = tile2_j_to_global(fa, fsoff, k, sm.rep[e] & 0x1fffu);
            }
            base += __popc(bal);
          }
          G = tot;
        }
        __syncthreads();
        // ---- epilogue B: exclusive prefix of the per-position counts (n_t + 1 entries), entries into position order ----
        {
          const uint32_t per = (n_t + 1 + THREADS - 1) / THREADS;
          const uint32_t b0 = tid * per, b1 = min(n_t + 1, b0 + per);
          uint32_t sum = 0;
          for (uint32_t p = b0; p < b1; ++p) sum += sm.pc[p];
          uint32_t run = tb_block_exscan<OpSumU32>(sum, s_scan, (uint32_t*)nullptr);
          for (uint32_t p = b0; p < b1; ++p) { const uint32_t c = sm.pc[p]; sm.pc[p] = run; run += c; }
        }
        __syncthreads();
        for (uint32_t r = tid; r < G; r += THREADS) {
          const uint32_t e = sm.occ[r];
          sm.ord[sm.pc[sm.prank[e]] + sm.tmp[e]] = (uint16_t)e;
        }
        __syncthreads();
        // ---- epilogue C: rank inside the position by the 64-bit keys; ties refined by (rank, next 6 bytes of the comparison
        // string) rounds, what still ties goes through the exact comparator (same scheme as col_tile_kernel) ----
        constexpr int TILE_REFINE = 6;
        for (int round = 0;; ++round) {
          int any_tie = 0;
          for (uint32_t r = tid; r < G; r += THREADS) {
            const uint32_t e = sm.occ[r];
            const uint32_t pr = sm.prank[e];
            const uint32_t q0 = sm.pc[pr], q1 = sm.pc[pr + 1];
            const unsigned long long ki = sm.skey[e];
            uint32_t rank = 0, ties = 0;
            for (uint32_t q = q0; q < q1; ++q) {
              const uint32_t eq = sm.ord[q];
              const unsigned long long kq = sm.skey[eq];
              rank += kq < ki; ties += (kq == ki && eq != e);
            }
            const uint16_t sat = round == 0 ? ((((uint32_t)(ki >> 24) & 0xffu) == 255u) ? 0x4000u : 0u) : (sm.tmp[e] & 0x4000u);
            sm.tmp[e] = (uint16_t)(rank | sat | (ties ? 0x8000u : 0u));
            any_tie |= (ties != 0 && !sat);
          }
          if (round >= TILE_REFINE || MODE == TB_MODE_CIGAR) { __syncthreads(); break; }
          if (!__syncthreads_or(any_tie)) break;
          for (uint32_t r = tid; r < G; r += THREADS) {
            const uint32_t e = sm.occ[r];
            const uint32_t t = sm.tmp[e];
            unsigned long long chunk = 0;
            if ((t & 0xC000u) == 0x8000u) chunk = tile_tail_chunk(in, sm.rep[e], (uint32_t)round * 6u);
            sm.skey[e] = ((unsigned long long)(t & 0x1fffu) << 48) | chunk;
          }
          __syncthreads();
        }
        for (uint32_t r = tid; r < G; r += THREADS) {
          const uint32_t e = sm.occ[r];
          const uint32_t pr = sm.prank[e];
          const uint32_t q0 = sm.pc[pr], q1 = sm.pc[pr + 1];
          const uint32_t t = sm.tmp[e];
          uint32_t rank = t & 0x1fffu;
          if (t & 0x8000u) {
            const unsigned long long ki = sm.skey[e];
            const uint32_t oi = sm.rep[e];
            for (uint32_t q = q0; q < q1; ++q) {
              const uint32_t eq = sm.ord[q];
              if (eq != e && sm.skey[eq] == ki && tb_mode_cmp(in, sm.rep[eq], oi) < 0) ++rank;
            }
          }
          const uint64_t o = (uint64_t)rank0 + gbase + q0 + rank;
          tp.st_rep[o] = sm.rep[e];
          tp.st_yc[o] = (float)sm.cnt[e];
          uint32_t yx = 0;
          for (uint32_t w = 0; w < W; ++w) { const uint32_t bw = sm.bits[e * W + w]; yx += __popc(bw); tp.st_bits[o * W + w] = bw; }
          tp.st_yx[o] = yx;
        }
        gbase += G;
        if (ps + 1 < npass) __syncthreads();      // the table is cleared again by the next pass
      }
      if (!ovf) {
        if (tid == 0) { tp.gcount[cur.m] = gbase; kept_total += kept_acc; }
        done = true;
      }
    }
    if (!produced) { t2_fence_proxy_async(); produce(s_slot[nb_], nb_); }
  }
  if (tid == 0 && kept_total) atomicAdd((unsigned long long*)&tp.status[CS_NKEPT], (unsigned long long)kept_total);
}
