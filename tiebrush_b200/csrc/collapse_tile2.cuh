// collapse_tile2.cuh — second-generation tile kernel of the collapse front end (included inside collapse_tile.cu's
// anonymous namespace). Same contract as col_tile_kernel: one slot = the start positions whose first merged rank falls into
// [mT,(m+1)T); groups of the slot go to the staging arrays at the slot's first merged rank, in final order
// (reference: src/tmerge.h:28-50 merge order, src/tiebrush.cpp:438-499 grouping, :501-530 output order).
//
// What is different from col_tile_kernel (profiles/r01f: 29 warp-instructions per record, 7 % DRAM throughput,
// every key verification two dependent L2 round trips):
//   * the slot's k file slices of pos / cig_off and their CIGAR words are brought into shared memory by the TMA unit
//     (cp.async.bulk 1-D copies issued by one thread per file, completion counted on an mbarrier), in chunks of T records
//     of the slot's file-major list: an ordinary slot is one chunk, a pile-up position as many as it takes. The first
//     chunk of slot i+1 is in flight while slot i's epilogue runs. The dependent chain pos -> cig_off -> cigar is
//     shared-memory traffic only.
//   * ONE table per slot, hashed on (position, strand, end, mode key), sized for the groups expected (not for "every
//     record distinct"). The key of a group (position + CIGAR words) is copied into a shared-memory arena when the group
//     is inserted, so verification never leaves shared memory and the table outlives the chunks. The position order
//     comes from a counting sort of the occupied entries by the merged rank of their position (P[pos] - rank0 < T).
//   * a slot with more groups than the table holds (stretches where nearly every alignment is distinct) is redone in 4
//     or 8 passes, each over the positions of a quarter / an eighth of its records.
//   * only what still does not fit is DEFERRED to col_tile_kernel's full-size-table launch through the heavy list: a
//     pile-up of more distinct alignments at ONE position than the table holds, or 32 records whose CIGARs outgrow the
//     staging area. Results are identical by construction (same staging arrays).

struct Tile2Params {
  uint32_t M, E, logE, W, T, rs_cap, cw_cap, arena, k;
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

static uint32_t tile2_arena_words(uint32_t E, int mode) { return (E - (E >> 3)) * (mode == TB_MODE_CIGAR ? 4u : 5u); }
static size_t tile2_smem_bytes(uint32_t k, uint32_t E, uint32_t W, uint32_t T, uint32_t rs_cap, uint32_t cw_cap, uint32_t arena) {
  return (size_t)rs_cap * 8 + (size_t)cw_cap * 4 + (size_t)arena * 4 + (size_t)E * (8 + 8 + 4 + 4 * (size_t)W + 2 + 2 + 2) + (size_t)(T + 4) * 4 +
         2 * (size_t)(4 * k + 4) * 4 + 128;
}

enum { T2_READY = 0, T2_EMPTY = 1, T2_END = 3 };
struct Slot2 { uint32_t m, rank0, n_t, state, p0, p1; };

struct Tile2Smem {
  uint32_t* s_pos; uint32_t* s_co; uint32_t* s_cig;   // staged chunk (TMA destinations, 16-byte aligned)
  uint32_t* arena;            // keys of the groups of the slot: [pos, (window index,) key CIGAR words] per group
  unsigned long long* word;   // [E] streaming: tag32 << 32 | n_cigar << 16 | arena offset ; epilogue: 64-bit sort key
  unsigned long long* rep;    // [E] min over members of (running max end) << 32 | window index
  uint32_t* cnt;              // [E]
  uint32_t* bits;             // [E*W]
  uint32_t* pc;               // [T+2] epilogue: groups per position (by merged rank of the position), then their exclusive prefix
  uint32_t* fgeo;             // per-file geometry, double buffered: [2][4][k+1] = a | soff (slot) | radj | cadj (chunk)
  uint16_t* prank;            // [E] merged rank of the entry's position inside the slot (< T)
  uint16_t* ord;              // [E] occupied entries in position order
  uint16_t* tmp;              // [E] arrival index inside the position (0xffff = empty), then rank | saturated(14) | tied(15)
};

__device__ __forceinline__ uint32_t t2_fold(uint32_t h, uint32_t w) { h = (h ^ w) * 0x9E3779B1u; return h ^ (h >> 15); }

// one pass over the CIGAR of a record (in shared memory): reference length, running hash of the mode key, and the two
// variable fields of the 64-bit sort key (key length saturated to 8 bits, first three key bytes in memcmp order / first exon length)
template <int MODE, bool HASH>
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
    if (HASH && (MODE != TB_MODE_CLIP || (c >= a && c < b))) hh = t2_fold(hh, w);
  }
  if (HASH && MODE == TB_MODE_FULL) {
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

template <int THREADS, int MODE>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS) col_tile2_kernel(ColIn in_, Tile2Params tp) {
  constexpr int NW = THREADS / 32;
  constexpr uint32_t HDR = MODE == TB_MODE_CIGAR ? 1u : 2u;   // arena record header: pos (, window index of the owner)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint32_t s_scan[33];
  __shared__ uint32_t s_wtot[4][3];
  __shared__ uint32_t s_flags[8];          // [0] kept  [1] overflow  [2] groups inserted  [3] overflow more passes cannot cure  [4] arena cursor
  __shared__ uint32_t s_range[2];          // positions of the current pass, relative to pos_lo
  __shared__ int s_cc[4];                  // (file, position, running max end) at the last record of the previous chunk
  __shared__ Slot2 s_slot[2];
  __shared__ __align__(8) uint64_t s_bar;
  volatile uint32_t* vflags = s_flags;
  ColIn in = in_;
  in.mode = MODE;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t E = tp.E, W = tp.W, k = tp.k, emask = tp.E - 1, C = tp.T;
  Tile2Smem sm;
  {
    unsigned char* p = smem_raw;
    sm.s_pos = (uint32_t*)p; p += (size_t)tp.rs_cap * 4;
    sm.s_co = (uint32_t*)p; p += (size_t)tp.rs_cap * 4;
    sm.s_cig = (uint32_t*)p; p += (size_t)tp.cw_cap * 4;
    sm.arena = (uint32_t*)p; p += (size_t)tp.arena * 4;
    sm.word = (unsigned long long*)p; p += (size_t)E * 8;
    sm.rep = (unsigned long long*)p; p += (size_t)E * 8;
    sm.cnt = (uint32_t*)p; p += (size_t)E * 4;
    sm.bits = (uint32_t*)p; p += (size_t)E * W * 4;
    sm.pc = (uint32_t*)p; p += (size_t)(tp.T + 4) * 4;
    sm.fgeo = (uint32_t*)p; p += (size_t)8 * (k + 1) * 4;
    sm.prank = (uint16_t*)p; p += (size_t)E * 2;
    sm.ord = (uint16_t*)p; p += (size_t)E * 2;
    sm.tmp = (uint16_t*)p;
  }
  if (tid == 0) {
    t2_mbar_init(&s_bar, k);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t phase = 0;        // parity of the mbarrier phase the next chunk completes
  uint32_t kept_total = 0;   // thread 0 only

  // exclusive prefixes over the files (thread f < k <= 128 holds file f) of up to three values; totals in s_wtot. One barrier.
  auto file_scan = [&](uint32_t& x0, uint32_t& x1, uint32_t& x2, uint32_t& t0, uint32_t& t1, uint32_t& t2) {
    const uint32_t v0 = x0, v1 = x1, v2 = x2;
    if (warp < 4) {
#pragma unroll
      for (int dd = 1; dd < 32; dd <<= 1) {
        const uint32_t y0 = __shfl_up_sync(0xffffffffu, x0, dd), y1 = __shfl_up_sync(0xffffffffu, x1, dd), y2 = __shfl_up_sync(0xffffffffu, x2, dd);
        if ((int)lane >= dd) { x0 += y0; x1 += y1; x2 += y2; }
      }
      if (lane == 31) { s_wtot[warp][0] = x0; s_wtot[warp][1] = x1; s_wtot[warp][2] = x2; }
    }
    __syncthreads();
    uint32_t p0 = 0, p1 = 0, p2 = 0; t0 = t1 = t2 = 0;
#pragma unroll
    for (uint32_t w = 0; w < 4; ++w) {
      if (w < warp) { p0 += s_wtot[w][0]; p1 += s_wtot[w][1]; p2 += s_wtot[w][2]; }
      t0 += s_wtot[w][0]; t1 += s_wtot[w][1]; t2 += s_wtot[w][2];
    }
    x0 = x0 - v0 + p0; x1 = x1 - v1 + p1; x2 = x2 - v2 + p2;
  };
  // claim the next slot (one thread)
  auto claim = [&](Slot2& d) {
    Slot2 t; t.m = atomicAdd(tp.slot_counter, 1u); t.rank0 = 0; t.n_t = 0; t.state = T2_END; t.p0 = t.p1 = 0;
    if (t.m < tp.M) {
      const uint32_t p0 = tp.slotpos[t.m], p1 = tp.slotpos[t.m + 1];
      t.p0 = p0; t.p1 = p1;
      t.rank0 = tp.P[p0];
      t.n_t = tp.P[p1] - t.rank0;
      t.state = t.n_t == 0 ? T2_EMPTY : T2_READY;
    }
    d = t;
  };
  // slot-level geometry into buffer nb: a[f] = first record of the slot in file f, soff[f] = exclusive prefix of the slice lengths
  auto setup = [&](const Slot2& d, int nb) {
    __syncthreads();                       // d is visible; buffer nb is not read any more
    const Slot2 s = d;
    if (s.state != T2_READY) return;       // block-uniform
    uint32_t a = 0, len = 0, z1 = 0, z2 = 0, t0, t1, t2;
    if (tid < k) { a = tp.off[(uint64_t)s.m * k + tid]; len = tp.off[(uint64_t)(s.m + 1) * k + tid] - a; }
    uint32_t e0 = len;
    file_scan(e0, z1, z2, t0, t1, t2);
    uint32_t* ga = sm.fgeo + (size_t)nb * 4 * (k + 1);
    if (tid < k) { ga[tid] = a; ga[(k + 1) + tid] = e0; if (tid == k - 1) ga[(k + 1) + k] = e0 + len; }
  };
  // start the copies of the chunk [cj0, ...) of the slot whose geometry is in buffer nb (all threads; the staging area must
  // be free). The chunk is C records unless their CIGAR words outgrow the arena: then it is halved until it fits.
  // Returns the end of the chunk (block-uniform), or 0 when not even 32 records fit.
  auto produce = [&](uint32_t n_t, uint32_t cj0, int nb) -> uint32_t {
    uint32_t* ga = sm.fgeo + (size_t)nb * 4 * (k + 1);
    uint32_t cj1 = min(n_t, cj0 + C);
    for (;;) {
      __syncthreads();                     // slot geometry visible; s_wtot free; every reader of the staging area is done
      uint32_t ra = 0, rb = 0, ca = 0, cb = 0, sof = 0;
      const bool isf = tid < k;
      if (isf) {
        const uint32_t a = ga[tid]; sof = ga[(k + 1) + tid];
        const uint32_t len = ga[(k + 1) + tid + 1] - sof;
        const uint32_t lo = cj0 > sof ? min(cj0 - sof, len) : 0u, hi = cj1 > sof ? min(cj1 - sof, len) : 0u;
        ra = a + lo; rb = a + hi;
        if (rb > ra) { ca = in.cig_off[ra]; cb = in.cig_off[rb]; }
      }
      const uint32_t len = rb - ra;
      const uint32_t a4 = ra & ~3u, ca4 = ca & ~3u;
      const uint32_t nrs = len ? (((rb + 4u) & ~3u) - a4) : 0u;      // slots covering records ra..rb (cig_off needs entry rb too)
      const uint32_t ncw = len ? (((cb + 3u) & ~3u) - ca4) : 0u;
      uint32_t x0 = 0, rbase = nrs, cbase = ncw, t0, t1, t2;
      file_scan(x0, rbase, cbase, t0, t1, t2);
      if (t1 > tp.rs_cap || t2 > tp.cw_cap) {   // block-uniform
        const uint32_t cl = (cj1 - cj0) >> 1;
        if (cl < 32u) return 0u;
        cj1 = cj0 + cl;
        continue;
      }
      if (isf) {
        ga[2 * (k + 1) + tid] = rbase + (ra & 3u) - (sof + (ra - ga[tid]));   // staging slot of compact index j: radj + j
        ga[3 * (k + 1) + tid] = cbase + (ca & 3u) - ca;                      // shared CIGAR index of window CIGAR offset c: cadj + c
        if (len) {
          // 16-byte chunks through the TMA unit; a chunk that would run past the end of its column is copied by hand
          const uint32_t pe = (rb + 3u) & ~3u, plim = tp.n & ~3u, pb = pe < plim ? pe : plim;                 // pos: n entries
          const uint32_t oe = (rb + 4u) & ~3u, olim = (tp.n + 1u) & ~3u, ob = oe < olim ? oe : olim;          // cig_off: n+1 entries
          const uint32_t we = (cb + 3u) & ~3u, wlim = tp.n_cig & ~3u, wb = we < wlim ? we : wlim;             // cigar: n_cig words
          for (uint32_t e = pb > a4 ? pb : a4; e < rb; ++e) sm.s_pos[rbase + (e - a4)] = (uint32_t)in.pos[e];
          for (uint32_t e = ob > a4 ? ob : a4; e <= rb; ++e) sm.s_co[rbase + (e - a4)] = in.cig_off[e];
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
      return cj1;
    }
  };

  const uint32_t glimit = E - (E >> 3);
  if (tid == 0) claim(s_slot[0]);
  setup(s_slot[0], 0);
  uint32_t pre_c1 = 0;                       // end of the chunk [0, pre_c1) in flight for the slot about to be processed (0 = none)
  if (s_slot[0].state == T2_READY) pre_c1 = produce(s_slot[0].n_t, 0u, 0);
  for (uint32_t it = 0;; ++it) {
    const int cb_ = (int)(it & 1), nb_ = cb_ ^ 1;
    if (tid == 0) claim(s_slot[nb_]);        // its latency hides behind this slot's work
    __syncthreads();
    const Slot2 cur = s_slot[cb_];
    if (cur.state == T2_END) break;
    bool produced = false;                   // the next slot's geometry and first chunk have been started
    auto prefetch_next = [&]() {
      t2_fence_proxy_async();
      setup(s_slot[nb_], nb_);
      const Slot2 nx = s_slot[nb_];
      pre_c1 = nx.state == T2_READY ? produce(nx.n_t, 0u, nb_) : 0u;
      produced = true;
    };
    if (cur.state != T2_READY) {
      if (tid == 0) tp.gcount[cur.m] = 0;
      prefetch_next();
      continue;
    }
    const uint32_t n_t = cur.n_t, rank0 = cur.rank0;
    const uint32_t* fa = sm.fgeo + (size_t)cb_ * 4 * (k + 1); const uint32_t* fsoff = fa + (k + 1); const uint32_t* fradj = fsoff + (k + 1); const uint32_t* fcadj = fradj + (k + 1);
    // A slot whose groups outgrow the table (stretches of the genome where nearly every alignment is distinct) is redone in
    // npass passes over its records, each pass taking the positions of one quarter (eighth) of the slot's records.
    uint32_t staged_c1 = pre_c1;             // chunk [0, staged_c1) is in the staging area or on its way (0 = nothing usable)
    bool staged_waited = false, done = false, hard = staged_c1 == 0u;
    for (uint32_t npass = 1; !done; npass = npass == 1 ? 4u : npass * 2u) {
      if (hard || npass > 8u) {   // block-uniform: full-size table launch of generation 1 (pile-up of distinct alignments, monster CIGARs)
        if (tid == 0) {
          tp.gcount[cur.m] = 0; tp.heavy_list[atomicAdd((unsigned long long*)&tp.status[CS_NHEAVY], 1ULL)] = cur.m;
          atomicAdd((unsigned long long*)&tp.status[CS_T2_TABLE], 1ULL);
          atomicAdd((unsigned long long*)&tp.status[CS_T2_STAGE], (unsigned long long)n_t);   // records of the deferred slots
        }
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
            const uint32_t target = (uint32_t)(((unsigned long long)min(n_t, C) * (ps + q)) / npass);
            uint32_t lo = cur.p0, hi = cur.p1;   // first p in [p0,p1] with P[p]-rank0 >= target
            if (ps + q == 0) hi = lo; else if (ps + q == npass) lo = hi;
            while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if (tp.P[mid] - rank0 >= target) hi = mid; else lo = mid + 1; }
            b2[q] = lo;
          }
          s_range[0] = b2[0]; s_range[1] = b2[1];
          s_flags[0] = 0; s_flags[1] = 0; s_flags[2] = 0; s_flags[3] = 0; s_flags[4] = 0;
          s_cc[0] = -1;
        }
        // ---- clear the table (while the copies land) ----
        for (uint32_t s = tid; s < E; s += THREADS) { sm.word[s] = EMPTY64; sm.rep[s] = EMPTY64; sm.cnt[s] = 0; }
        {
          uint4* b4 = reinterpret_cast<uint4*>(sm.bits);
          const uint32_t n4 = (E * W) >> 2;   // E is a multiple of 4
          for (uint32_t s = tid; s < n4; s += THREADS) b4[s] = make_uint4(0u, 0u, 0u, 0u);
        }
        uint32_t my_kept = 0;
        // ---- the chunks of the slot ----
        for (uint32_t cj0 = 0; cj0 < n_t && !ovf;) {
          uint32_t cj1;
          if (cj0 == 0 && staged_c1) {        // prefetched during the previous epilogue, or still there from the previous pass
            cj1 = staged_c1;
            if (!staged_waited) { while (!t2_mbar_try_wait(&s_bar, phase)) {} phase ^= 1u; staged_waited = true; }
          } else {
            t2_fence_proxy_async();
            cj1 = produce(n_t, cj0, cb_);     // starts with a barrier: the staging area is free
            if (cj1 == 0u) { hard = true; ovf = true; break; }   // block-uniform
            while (!t2_mbar_try_wait(&s_bar, phase)) {}
            phase ^= 1u;
            staged_c1 = 0u;
          }
          if (cj1 < n_t) staged_c1 = 0u;      // more than one chunk: the first one will not survive
          __syncthreads();                    // table cleared, s_range / s_cc visible, copies landed for everybody
          const uint32_t pl = s_range[0], pspan = s_range[1] - s_range[0];
          const uint32_t cn = cj1 - cj0;

          // ---- stream this warp's contiguous range of the chunk ----
          const uint32_t j0 = cj0 + (uint32_t)(((unsigned long long)cn * warp) / NW), j1 = cj0 + (uint32_t)(((unsigned long long)cn * (warp + 1)) / NW);
          uint32_t carry_f = 0xffffffffu; int carry_pos = INT_MIN; int carry_max = 0;
          uint32_t f_lo = 0;
          if (j0 < j1) {
            uint32_t lo = 0, hi = k;
            while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (fsoff[mid] <= j0) lo = mid; else hi = mid; }
            f_lo = lo;
            const uint32_t fstart = max(fsoff[f_lo], cj0);   // first record of this file inside the chunk
            const uint32_t rs0 = fradj[f_lo] + j0, cad = fcadj[f_lo];
            const int p = (int)sm.s_pos[rs0];
            const bool cont = cj0 > fsoff[f_lo] && s_cc[0] == (int)f_lo && s_cc[1] == p;   // the previous chunk ended inside this file at position p
            if (j0 > fstart) {   // the range starts inside a file slice: running maximum of the (file, position) run before it
              if ((int)sm.s_pos[rs0 - 1] == p) {
                const uint32_t nback = j0 - fstart;
                int mx = 0; bool all = true;
                for (uint32_t dn = 0; dn < nback; dn += 32) {
                  const bool ok = dn + lane < nback;
                  const uint32_t rs = ok ? rs0 - 1 - dn - lane : rs0;
                  const bool same = (int)sm.s_pos[rs] == p;
                  const int rl = (ok && same) ? tb_ref_len(sm.s_cig, cad + sm.s_co[rs], cad + sm.s_co[rs + 1]) : 0;
                  const unsigned notsame = __ballot_sync(0xffffffffu, ok && !same);
                  const unsigned first = notsame ? (unsigned)(__ffs(notsame) - 1) : 32u;
                  if (lane < first) mx = max(mx, rl);
                  if (notsame) { all = false; break; }
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
                if (all && fstart == cj0 && cont) mx = max(mx, s_cc[2]);   // the run reaches back into the previous chunk
                carry_f = f_lo; carry_pos = p; carry_max = mx;
              }
            } else if (j0 == cj0 && cont) {
              carry_f = f_lo; carry_pos = p; carry_max = s_cc[2];
            }
          }
          __syncwarp();
          for (uint32_t jb = j0; jb < j1; jb += 32) {
            if (__any_sync(0xffffffffu, vflags[1] != 0)) break;   // warp-uniform: some record of the pass could not be placed
            const uint32_t j = jb + lane;
            const bool valid = j < j1;
            uint32_t f = 0xfffffffeu, gi = 0, cs0 = 0, h = 0, len8 = 0, w24 = 0, nc = 0;
            int pos = INT_MIN + 1 + (int)lane, reflen = 0; bool pass = false; unsigned sc = 0;
            if (valid) {
              f = f_lo;
              while (fsoff[f + 1] <= j) ++f;
              const uint32_t rs = fradj[f] + j;
              pos = (int)sm.s_pos[rs];
              if ((uint32_t)(pos - in.pos_lo) - pl < pspan) {   // a record of another pass only keeps its (file, position) run apart
                gi = fa[f] + (j - fsoff[f]);
                const uint16_t fl = in.flag[gi]; const uint8_t mq = in.mapq[gi]; const uint16_t nh = in.nh[gi];
                sc = tb_strand_code(in.strand[gi]);
                const uint32_t co = sm.s_co[rs], co1 = sm.s_co[rs + 1];
                cs0 = fcadj[f] + co; nc = co1 - co;
                tile2_parse<MODE, true>(in, sm.s_cig, cs0, cs0 + nc, pos, gi, reflen, h, len8, w24);
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
            const int ll = (j1 - jb >= 32u) ? 31 : (int)(j1 - jb - 1u);   // last valid lane of the step
            carry_f = __shfl_sync(0xffffffffu, f, ll); carry_pos = __shfl_sync(0xffffffffu, pos, ll); carry_max = __shfl_sync(0xffffffffu, Erel, ll);
            f_lo = carry_f;

            uint32_t s = 0;
            if (pass) {
              ++my_kept;
              const uint32_t x = h ^ ((uint32_t)pos * 0x85EBCA77u) ^ ((uint32_t)reflen * 0xC2B2AE3Du) ^ tp.seed;
              const uint32_t tag = (sc << 30) | (tb_mix32(x) >> 2);
              s = (x * 0x9E3779B1u) >> (32u - tp.logE);
              bool placed = false;
              uint32_t ko = 0xffffffffu;
              for (uint32_t t = 0; t < E; ++t) {
                unsigned long long w = sm.word[s];
                if (w == EMPTY64) {
                  if (ko == 0xffffffffu) {   // the key goes into the arena before the entry is published
                    if (vflags[2] > glimit) break;                       // table (nearly) full: more passes
                    ko = atomicAdd(&s_flags[4], HDR + nc);
                    if (ko + HDR + nc > tp.arena || nc > 0xffffu) break;   // arena full: more passes (a 65536-op CIGAR cannot be in a BAM record)
                    sm.arena[ko] = (uint32_t)pos;
                    if (HDR == 2u) sm.arena[ko + 1] = gi;
                    for (uint32_t q = 0; q < nc; ++q) sm.arena[ko + HDR + q] = sm.s_cig[cs0 + q];
                    __threadfence_block();
                  }
                  const unsigned long long mine = ((unsigned long long)tag << 32) | ((unsigned long long)nc << 16) | ko;
                  w = atomicCAS(&sm.word[s], EMPTY64, mine);
                  if (w == EMPTY64) {
                    sm.prank[s] = (uint16_t)(__ldg(&tp.P[(uint32_t)(pos - in.pos_lo)]) - rank0);
                    atomicAdd(&s_flags[2], 1u);
                    placed = true; break;
                  }
                }
                if ((uint32_t)(w >> 32) == tag) {
                  const uint32_t oko = (uint32_t)w & 0xffffu, onc = ((uint32_t)w >> 16) & 0xffffu;
                  if ((int)sm.arena[oko] == pos) {
                    bool same = onc == nc;
                    if (same) for (uint32_t q = 0; q < nc; ++q) if (sm.s_cig[cs0 + q] != sm.arena[oko + HDR + q]) { same = false; break; }
                    if (MODE != TB_MODE_CIGAR && (MODE == TB_MODE_FULL || !same))   // identical raw CIGARs decide -P / -E; -L adds the MD bytes; the rest takes the exact comparator
                      same = tb_mode_cmp(in, gi, sm.arena[oko + 1]) == 0;
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
              const unsigned long long repkey = ((unsigned long long)(uint32_t)Erel << 32) | gi;
              if (repkey < sm.rep[s]) atomicMin(&sm.rep[s], repkey);
              const uint32_t bi = s * W + (f >> 5), bit = 1u << (f & 31);
              if (!(sm.bits[bi] & bit)) atomicOr(&sm.bits[bi], bit);
            }
          }
          if (j0 < j1 && j1 == cj1) { s_cc[0] = (int)carry_f; s_cc[1] = carry_pos; s_cc[2] = carry_max; }   // the warp that holds the chunk's last record
          __syncthreads();                          // the chunk is done: table and s_cc complete, staging area free
          ovf = vflags[1] != 0;                     // block-uniform
          cj0 = cj1;
        }
        if (my_kept) atomicAdd(&s_flags[0], my_kept);
        if (ovf && n_t > C) hard = true;   // a pile-up position with more distinct alignments than the table: more passes over positions cannot help
        if (ovf) break;
        __syncthreads();
        kept_acc += vflags[0];
        if (ps + 1 == npass) prefetch_next();       // nobody reads the staging area any more: the next slot's first chunk lands during the epilogue

        // ---- epilogue A: sort keys from the arena, groups per position ----
        for (uint32_t p = tid; p <= C + 1; p += THREADS) sm.pc[p] = 0;
        __syncthreads();
        for (uint32_t e = tid; e < E; e += THREADS) {
          const unsigned long long w = sm.word[e];
          uint16_t t = 0xffffu;
          if (w != EMPTY64) {
            const uint32_t ko = (uint32_t)w & 0xffffu, nc = ((uint32_t)w >> 16) & 0xffffu;
            int reflen; uint32_t h, len8, w24;
            tile2_parse<MODE, false>(in, sm.arena + ko + HDR, 0u, nc, (int)sm.arena[ko], 0u, reflen, h, len8, w24);
            sm.word[e] = ((unsigned long long)(w >> 62) << 62) | ((unsigned long long)((uint32_t)reflen & 0x3fffffffu) << 32) | ((unsigned long long)len8 << 24) | w24;
            t = (uint16_t)atomicAdd(&sm.pc[sm.prank[e]], 1u);   // arrival index inside the position
          }
          sm.tmp[e] = t;
        }
        __syncthreads();
        // ---- epilogue B: exclusive prefix of the per-position counts (a position's merged rank inside the slot is < C),
        // entries into position order ----
        {
          const uint32_t per = (C + 1 + THREADS - 1) / THREADS;
          const uint32_t b0 = tid * per, b1 = min(C + 1, b0 + per);
          uint32_t sum = 0;
          for (uint32_t p = b0; p < b1; ++p) sum += sm.pc[p];
          uint32_t run = tb_block_exscan<OpSumU32>(sum, s_scan, (uint32_t*)nullptr);
          for (uint32_t p = b0; p < b1; ++p) { const uint32_t c = sm.pc[p]; sm.pc[p] = run; run += c; }
        }
        __syncthreads();
        const uint32_t G = sm.pc[C];
        for (uint32_t e = tid; e < E; e += THREADS) {
          const uint32_t t = sm.tmp[e];
          if (t != 0xffffu) sm.ord[sm.pc[sm.prank[e]] + t] = (uint16_t)e;
        }
        __syncthreads();
        // ---- epilogue C: rank inside the position by the 64-bit keys; ties refined by (rank, next 6 bytes of the comparison
        // string) rounds, what still ties goes through the exact comparator (same scheme as col_tile_kernel) ----
        constexpr int TILE_REFINE = 6;
        for (int round = 0;; ++round) {
          int any_tie = 0;
          for (uint32_t r = tid; r < G; r += THREADS) {
            const uint32_t e = sm.ord[r];
            const uint32_t pr = sm.prank[e];
            const uint32_t q0 = sm.pc[pr], q1 = sm.pc[pr + 1];
            const unsigned long long ki = sm.word[e];
            uint32_t rank = 0, ties = 0;
            for (uint32_t q = q0; q < q1; ++q) {
              const uint32_t eq = sm.ord[q];
              const unsigned long long kq = sm.word[eq];
              rank += kq < ki; ties += (kq == ki && eq != e);
            }
            const uint16_t sat = round == 0 ? ((((uint32_t)(ki >> 24) & 0xffu) == 255u) ? 0x4000u : 0u) : (sm.tmp[e] & 0x4000u);
            sm.tmp[e] = (uint16_t)(rank | sat | (ties ? 0x8000u : 0u));
            any_tie |= (ties != 0 && !sat);
          }
          if (round >= TILE_REFINE || MODE == TB_MODE_CIGAR) { __syncthreads(); break; }
          if (!__syncthreads_or(any_tie)) break;
          for (uint32_t r = tid; r < G; r += THREADS) {
            const uint32_t e = sm.ord[r];
            const uint32_t t = sm.tmp[e];
            unsigned long long chunk = 0;
            if ((t & 0xC000u) == 0x8000u) chunk = tile_tail_chunk(in, (uint32_t)sm.rep[e], (uint32_t)round * 6u);
            sm.word[e] = ((unsigned long long)(t & 0x1fffu) << 48) | chunk;
          }
          __syncthreads();
        }
        for (uint32_t r = tid; r < G; r += THREADS) {
          const uint32_t e = sm.ord[r];
          const uint32_t pr = sm.prank[e];
          const uint32_t q0 = sm.pc[pr], q1 = sm.pc[pr + 1];
          const uint32_t t = sm.tmp[e];
          uint32_t rank = t & 0x1fffu;
          const uint32_t oi = (uint32_t)sm.rep[e];
          if (t & 0x8000u) {
            const unsigned long long ki = sm.word[e];
            for (uint32_t q = q0; q < q1; ++q) {
              const uint32_t eq = sm.ord[q];
              if (eq != e && sm.word[eq] == ki && tb_mode_cmp(in, (uint32_t)sm.rep[eq], oi) < 0) ++rank;
            }
          }
          const uint64_t o = (uint64_t)rank0 + gbase + q0 + rank;
          tp.st_rep[o] = oi;
          tp.st_yc[o] = (float)sm.cnt[e];
          uint32_t yx = 0;
          for (uint32_t w = 0; w < W; ++w) { const uint32_t bw = sm.bits[e * W + w]; yx += __popc(bw); tp.st_bits[o * W + w] = bw; }
          tp.st_yx[o] = yx;
        }
        gbase += G;
      }
      if (!ovf) {
        if (tid == 0) { tp.gcount[cur.m] = gbase; kept_total += kept_acc; }
        done = true;
      }
    }
    if (!produced) prefetch_next();
  }
  if (tid == 0 && kept_total) atomicAdd((unsigned long long*)&tp.status[CS_NKEPT], (unsigned long long)kept_total);
}
