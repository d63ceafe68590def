// tiebrush_gpu — the reference `tiebrush` command line with its hot loop running on a B200.
//
// This translation unit compiles the UNMODIFIED reference source (src/tiebrush.cpp, from where it lies under the
// reference checkout given with -I) with its main() renamed, and supplies a new main() that keeps everything the
// reference keeps on the host — option parsing (processOptions, tiebrush.cpp:603-680), header merge and the
// TieBrush-made-input detection (TInputFiles::start, tmerge.cpp:288-329), BGZF/BAM decode (GSamReader), tag patching
// (GSamRecord::add_*_tag -> bam_aux_update_*), BAM write (GSamWriter) and the final summary line — and replaces
//     TInputFiles::next() order + passes_options() + addPData() + flushPData()      (tiebrush.cpp:570-592)
// by a WINDOW PACKER and one tb_collapse_window() call per window (include/tiebrush_b200.h).
//
// Window = consecutive records of the merged stream on one tid, cut only where the next read starts beyond the end
// of every read seen so far on that tid (a global coverage gap): collapse groups never span a start position and
// every per-sample YD segment list is provably empty after the first read past such a gap (tiebrush.cpp:237-242), so
// a window needs no state from its predecessor. Records are handed over FILE-MAJOR (the device does the k-way
// merge itself); the raw GSamRecords of a window stay alive until rep_index[] comes back, then the representatives
// get their tags and are written, the rest are freed.
//
// Host pipeline: three stages on their own threads. The main thread decodes (TB_DECODE_THREADS readers)
// and cuts windows; a device thread per GPU owns a CUDA context and does pack -> tb_collapse_window; a writer thread patches
// the tags of the representatives and writes the BAM records strictly in hand-over order (BGZF compression on TB_IO_THREADS
// htslib worker threads) and hands the records to a background deleter. Decode of window w+2, pack + device of window w+1
// and tag + write of window w overlap, and CUDA start-up overlaps the first window's decode.
//
// No CPU fallback: without a CUDA device tb_create() fails and the tool exits 1 like any other GError.
#define main tb_reference_main_unused
#include "src/tiebrush.cpp"
#undef main

#include <vector>
#include <chrono>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <functional>
#include "tiebrush_b200.h"
#include <unistd.h>

// error exit from a worker thread: the reference's convention (message on stderr, exit code 1) without running exit handlers
// while the other threads of the pipeline are still alive
static void tb_die(const char* msg) { fputs(msg, stderr); fputc('\n', stderr); fflush(stderr); _exit(1); }

namespace {

struct TbWindow {   // what the reader thread hands to the device thread
  int tid = -1;
  long seq = 0;   // hand-over order: the windows are written in this order whichever GPU computed them
  size_t n = 0;
  std::vector<std::vector<GSamRecord*>> per_file;   // records of the window, per input file, in file order
  std::vector<uint8_t> file_merged;
  explicit TbWindow(int k) : per_file(k), file_merged(k, 0) {}
  void add(TInputRecord* irec) {
    per_file[irec->fidx].push_back(irec->brec);
    file_merged[irec->fidx] = irec->tbMerged ? 1 : 0;
    irec->disown();   // the window owns the record now (TInputFiles::next deletes its current record otherwise)
    ++n;
  }
};

// background deleter of the records of finished windows
struct Reaper {
  std::mutex m; std::condition_variable cv; std::deque<std::vector<GSamRecord*>> q; bool done = false; std::thread th;
  Reaper() : th([this] {
    for (;;) {
      std::vector<GSamRecord*> v;
      { std::unique_lock<std::mutex> lk(m); cv.wait(lk, [&] { return !q.empty() || done; }); if (q.empty()) return; v = std::move(q.front()); q.pop_front(); }
      const size_t n = v.size();
      int nt = n > 200000 ? 4 : 1;
      if (const char* e = getenv("TB_FREE_THREADS")) nt = atoi(e) > 0 ? atoi(e) : 1;
      if (nt == 1) { for (GSamRecord* r : v) delete r; continue; }
      std::vector<std::thread> w;
      for (int t = 0; t < nt; ++t) w.emplace_back([&, t] { for (size_t x = n * t / nt; x < n * (t + 1) / nt; ++x) delete v[x]; });
      for (auto& x : w) x.join();
    }
  }) {}
  void push(std::vector<GSamRecord*>&& v) { { std::lock_guard<std::mutex> lk(m); q.push_back(std::move(v)); } cv.notify_all(); }
  void finish() { { std::lock_guard<std::mutex> lk(m); done = true; } cv.notify_all(); th.join(); }
};

// What the device thread hands to the writer: the window's records and the groups tb_collapse_window returned for them.
struct WriteJob {
  long seq = 0; int64_t n_kept = 0, n_groups = 0;
  std::vector<GSamRecord*> held;
  std::vector<uint32_t> rep, yx; std::vector<float> yc; std::vector<int32_t> yd;
};

// The writer thread: windows may be computed on several GPUs at once (TB_DEVICES) and finish in any order, but they are tagged
// and written strictly in hand-over order (flushPData, tiebrush.cpp:506-527). A job is admitted when it is the next one to
// write or fewer than `cap` are waiting, so a slow output throttles the device threads without ever blocking the job the
// writer is waiting for.
struct Writer {
  htsFile* fp = NULL; sam_hdr_t* hdr = NULL; Reaper* reaper = NULL; int nt = 1; size_t cap = 2;
  std::mutex m; std::condition_variable cv; std::map<long, WriteJob> pend; long next = 0; bool done = false;
  double t_write = 0; std::thread th;
  void start() { th = std::thread([this] { run(); }); }
  void push(WriteJob&& j) {
    std::unique_lock<std::mutex> lk(m);
    const long seq = j.seq;
    cv.wait(lk, [&] { return seq == next || pend.size() < cap; });
    pend.emplace(seq, std::move(j));
    cv.notify_all();
  }
  void finish() { { std::lock_guard<std::mutex> lk(m); done = true; } cv.notify_all(); if (th.joinable()) th.join(); }
  void run() {
    for (;;) {
      WriteJob j;
      {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return pend.count(next) != 0 || (done && pend.empty()); });
        auto it = pend.find(next);
        if (it == pend.end()) return;
        j = std::move(it->second); pend.erase(it);
      }
      auto t0 = std::chrono::steady_clock::now();
      emit(j);
      t_write += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      { std::lock_guard<std::mutex> lk(m); ++next; }
      cv.notify_all();
    }
  }
  void emit(WriteJob& j) {
    inCounter += (uint64_t)j.n_kept;                                 // tiebrush.cpp:573
    // the tags of the representatives are patched on `nt` threads (records are independent), the BAM records are then written
    // in output order (BGZF compression on the htslib worker threads)
    const int64_t G = j.n_groups;
    auto patch = [&](int64_t g) {
      GSamRecord* r = j.held[j.rep[g]];
      r->add_double_tag("YC", (double)j.yc[g]);
      r->add_int_tag("YX", (int64_t)j.yx[g]);
      if (j.yd[g] > 0) r->add_int_tag("YD", j.yd[g]); else r->remove_tag("YD");
    };
    if (nt == 1 || G < 20000) { for (int64_t g = 0; g < G; ++g) patch(g); }
    else {
      std::vector<std::thread> w;
      for (int t = 0; t < nt; ++t) w.emplace_back([&, t] { for (int64_t g = G * t / nt; g < G * (t + 1) / nt; ++g) patch(g); });
      for (auto& x : w) x.join();
    }
    for (int64_t g = 0; g < G; ++g) {
      if (sam_write1(fp, hdr, j.held[j.rep[g]]->get_b()) < 0) GError("Error writing SAM record!\n");   // GSamWriter::write, GSam.h:648-653
      outCounter++;
    }
    // the window's records (millions of bam1_t + exon vectors) go back to the allocator on a background thread
    if (reaper) reaper->push(std::move(j.held));
    else for (GSamRecord* r : j.held) delete r;
  }
};

struct TbWindowPacker {
  int k = 0;
  Writer* writer = NULL;
  // SoA staging (reused between windows)
  std::vector<int64_t> run_off;
  std::vector<int32_t> pos, yx_in, yd_in;
  std::vector<uint16_t> flag, nh;
  std::vector<uint8_t> mapq, strand, md;
  std::vector<uint32_t> cig_off, cigar, md_off, cigar_ext;
  std::vector<uint8_t> n_cigar8; std::vector<uint16_t> cigar16;   // compact wire format of the CIGAR columns (tiebrush_b200.h)
  std::vector<uint64_t> qhash;
  std::vector<float> yc_in;
  std::vector<GSamRecord*> held;                    // window index -> record
  std::vector<uint32_t> o_rep, o_yx; std::vector<float> o_yc; std::vector<int32_t> o_yd;
  double t_pack = 0, t_device = 0;
  int64_t n_windows = 0;

  void init(int nfiles) { k = nfiles; }

  static uint64_t fnv1a(const char* s) {
    uint64_t h = 1469598103934665603ULL;
    for (; *s; ++s) { h ^= (unsigned char)*s; h *= 1099511628211ULL; }
    return h;
  }

  void flush(tb_ctx* ctx, TbWindow& w) {
    const size_t n = w.n; const int tid = w.tid;
    std::vector<std::vector<GSamRecord*>>& per_file = w.per_file;
    std::vector<uint8_t>& file_merged = w.file_merged;
    if (n == 0) { WriteJob e; e.seq = w.seq; writer->push(std::move(e)); return; }
    using clk = std::chrono::steady_clock;
    auto t0 = clk::now();
    const bool want_md = mrgStrategy == tMrgStratFull;
    const bool want_q = options.collapse_same;
    bool any_merged = false;
    for (int f = 0; f < k; ++f) any_merged |= file_merged[f] != 0;
    run_off.assign(k + 1, 0);
    pos.resize(n); flag.resize(n); mapq.resize(n); strand.resize(n); nh.resize(n); cig_off.resize(n + 1);
    held.resize(n);
    n_cigar8.resize(n);
    bool compact = getenv("TB_WIRE_WIDE") == NULL;   // falls back to the wide columns when a record has >= 256 ops
    if (want_md) md_off.resize(n + 1);
    if (want_q) qhash.resize(n);
    if (any_merged) { yc_in.resize(n); yx_in.resize(n); yd_in.resize(n); }
    // ---- two passes over the files, both on TB_PACK_THREADS threads (a share of the files each): sizes first (records,
    // CIGAR ops, MD bytes, escaped lengths per file), then every file fills its own slice of the columns ----
    int nt = (int)std::thread::hardware_concurrency(); if (nt > 16) nt = 16;
    if (const char* e = getenv("TB_PACK_THREADS")) nt = atoi(e);
    if (nt < 1) nt = 1;
    if (nt > k) nt = k;
    if (n < 50000) nt = 1;
    std::vector<size_t> f_rec(k + 1, 0), f_cig(k + 1, 0), f_md(k + 1, 0), f_ext(k + 1, 0);
    std::vector<uint8_t> f_long(k, 0);
    auto on_files = [&](const std::function<void(int)>& body) {
      if (nt == 1) { for (int f = 0; f < k; ++f) body(f); return; }
      std::vector<std::thread> th;
      for (int t = 0; t < nt; ++t) th.emplace_back([&, t] { for (int f = t; f < k; f += nt) body(f); });
      for (auto& x : th) x.join();
    };
    on_files([&](int f) {
      size_t c = 0, m = 0, x = 0; bool lg = false;
      for (GSamRecord* r : per_file[f]) {
        bam1_t* b = r->get_b();
        c += b->core.n_cigar;
        if (b->core.n_cigar >= 256) lg = true;
        const uint32_t* cg = bam_get_cigar(b);
        for (uint32_t q = 0; q < b->core.n_cigar; ++q) if ((cg[q] >> 4) >= 0xFFFu) ++x;
        if (want_md) { const char* s = r->tag_str("MD"); if (s) m += strlen(s) + 1; }
      }
      f_rec[f + 1] = per_file[f].size(); f_cig[f + 1] = c; f_md[f + 1] = m; f_ext[f + 1] = x; f_long[f] = lg;
    });
    for (int f = 0; f < k; ++f) { f_rec[f + 1] += f_rec[f]; f_cig[f + 1] += f_cig[f]; f_md[f + 1] += f_md[f]; f_ext[f + 1] += f_ext[f]; if (f_long[f]) compact = false; }
    if (f_cig[k] >= (1ULL << 32)) GError("Error: a window of %zu CIGAR operations does not fit 32-bit offsets (lower TB_WINDOW_RECORDS)\n", f_cig[k]);
    cigar.resize(f_cig[k]);
    if (compact) { cigar16.resize(f_cig[k]); cigar_ext.resize(f_ext[k]); }
    if (want_md) md.resize(f_md[k]);
    std::vector<int32_t> f_lo(k, 0x7fffffff), f_hi(k, 0);
    on_files([&](int f) {
      size_t i = f_rec[f], ci = f_cig[f], mi = f_md[f], xi = f_ext[f];
      for (GSamRecord* r : per_file[f]) {
        bam1_t* b = r->get_b();
        held[i] = r;
        pos[i] = (int32_t)b->core.pos;
        if (pos[i] < f_lo[f]) f_lo[f] = pos[i];
        if (pos[i] >= f_hi[f]) f_hi[f] = pos[i] + 1;
        flag[i] = b->core.flag;
        mapq[i] = b->core.qual;
        strand[i] = (uint8_t)r->spliceStrand();                      // GSam.cpp:464-475
        int64_t v = r->tag_int("NH", 0);                             // passes_options, tiebrush.cpp:537
        nh[i] = (uint16_t)(v < 0 ? 0 : (v > 65535 ? 65535 : v));
        cig_off[i] = (uint32_t)ci;
        const uint32_t* c = bam_get_cigar(b);
        const uint32_t nc = b->core.n_cigar;
        memcpy(cigar.data() + ci, c, sizeof(uint32_t) * nc);
        if (compact) {
          n_cigar8[i] = (uint8_t)nc;
          for (uint32_t q = 0; q < nc; ++q) {
            const uint32_t len = c[q] >> 4;
            if (len >= 0xFFFu) { cigar16[ci + q] = (uint16_t)((c[q] & 0xfu) | (0xFFFu << 4)); cigar_ext[xi++] = len; }
            else cigar16[ci + q] = (uint16_t)((c[q] & 0xfu) | (len << 4));
          }
        }
        ci += nc;
        if (want_md) {
          md_off[i] = (uint32_t)mi;
          const char* m = r->tag_str("MD");                          // cmpFull, tiebrush.cpp:285-304
          if (m) { const size_t l = strlen(m) + 1; memcpy(md.data() + mi, m, l); mi += l; }
        }
        if (want_q) qhash[i] = fnv1a(r->name());
        if (any_merged) {
          yc_in[i] = file_merged[f] ? (float)r->tag_float("YC") : 0.f;  // SPData::settle, tiebrush.cpp:389-395
          yx_in[i] = file_merged[f] ? (int32_t)r->tag_int("YX", 1) : 1;
          yd_in[i] = file_merged[f] ? (int32_t)r->tag_int("YD", 0) : 0;
        }
        ++i;
      }
    });
    int32_t pos_lo = 0x7fffffff, pos_hi = 0;
    for (int f = 0; f < k; ++f) { run_off[f] = (int64_t)f_rec[f]; if (f_lo[f] < pos_lo) pos_lo = f_lo[f]; if (f_hi[f] > pos_hi) pos_hi = f_hi[f]; }
    const size_t i = n;
    run_off[k] = (int64_t)i;
    cig_off[n] = (uint32_t)cigar.size();
    if (want_md) md_off[n] = (uint32_t)md.size();
    if (cigar.empty()) cigar.push_back(0);
    if (cigar16.empty()) cigar16.push_back(0);
    if (want_md && md.empty()) md.push_back(0);

    tb_soa_in in; memset(&in, 0, sizeof(in));
    in.n = (int64_t)n; in.n_files = k; in.tid = tid; in.run_off = run_off.data(); in.file_merged = any_merged ? file_merged.data() : NULL;
    in.pos = pos.data(); in.flag = flag.data(); in.mapq = mapq.data(); in.strand = strand.data(); in.nh = nh.data();
    in.n_cig = (int64_t)cig_off[n];
    if (compact) { in.n_cigar8 = n_cigar8.data(); in.cigar16 = cigar16.data(); in.cigar_ext = cigar_ext.empty() ? NULL : cigar_ext.data(); in.n_ext = (int64_t)cigar_ext.size(); }
    else { in.cig_off = cig_off.data(); in.cigar = cigar.data(); }
    if (want_md) { in.md_off = md_off.data(); in.md = md.data(); in.n_md = (int64_t)md_off[n]; }
    if (want_q) in.qhash = qhash.data();
    if (any_merged) { in.yc_in = yc_in.data(); in.yx_in = yx_in.data(); in.yd_in = yd_in.data(); }
    in.on_device = 0; in.pos_lo = pos_lo; in.pos_hi = pos_hi;
    o_rep.resize(n); o_yc.resize(n); o_yx.resize(n); o_yd.resize(n);
    tb_groups_out out; memset(&out, 0, sizeof(out));
    out.capacity = (int64_t)n; out.rep_index = o_rep.data(); out.yc = o_yc.data(); out.yx = o_yx.data(); out.yd = o_yd.data();
    auto t1 = clk::now();
    if (tb_collapse_window(ctx, &in, &out)) tb_die(tb_last_error(ctx));
    auto t2 = clk::now();
    WriteJob job;
    job.seq = w.seq; job.n_kept = out.n_kept; job.n_groups = out.n_groups;
    job.held = std::move(held); job.rep = std::move(o_rep); job.yc = std::move(o_yc); job.yx = std::move(o_yx); job.yd = std::move(o_yd);
    held = std::vector<GSamRecord*>(); o_rep = std::vector<uint32_t>(); o_yc = std::vector<float>(); o_yx = std::vector<uint32_t>(); o_yd = std::vector<int32_t>();
    writer->push(std::move(job));   // blocks while the writer is more than a window behind
    ++n_windows;
    t_pack += std::chrono::duration<double>(t1 - t0).count();
    t_device += std::chrono::duration<double>(t2 - t1).count();
  }
};

}  // namespace

// bounded hand-off queue between the reader thread and the device thread
struct TbQueue {
  size_t max_queued = 2;
  std::mutex m; std::condition_variable cv;
  std::deque<std::unique_ptr<TbWindow>> q; bool done = false; long n_popped = 0;
  void push(std::unique_ptr<TbWindow> w) {
    std::unique_lock<std::mutex> lk(m);
    cv.wait(lk, [&] { return q.size() < max_queued; });
    q.push_back(std::move(w)); cv.notify_all();
  }
  std::unique_ptr<TbWindow> pop() {
    std::unique_lock<std::mutex> lk(m);
    cv.wait(lk, [&] { return !q.empty() || done; });
    if (q.empty()) return nullptr;
    std::unique_ptr<TbWindow> w = std::move(q.front()); q.pop_front(); w->seq = n_popped++; cv.notify_all();
    return w;
  }
  void finish() { std::lock_guard<std::mutex> lk(m); done = true; cv.notify_all(); }
};

int main(int argc, char* argv[]) {
  using clk = std::chrono::steady_clock;
  auto t_begin = clk::now();
  inRecords.setup(VERSION, argc, argv);
  processOptions(argc, argv);
  int numSamples = inRecords.start();
  // output: same header and record bytes as GSamWriter (GSam.h:537-573), BGZF compression on worker threads
  int io_threads = 8;   // BGZF compression workers of the output (TB_IO_THREADS): 4 -> 8 took tag+write from 1.0 to 0.7 s on the 5M-record sample
  if (const char* e = getenv("TB_IO_THREADS")) io_threads = atoi(e);
  htsFile* out_fp = hts_open(outfname.chars(), "wb");
  if (out_fp == NULL) GError("Error: could not create output file %s\n", outfname.chars());
  if (io_threads > 1) hts_set_threads(out_fp, io_threads);
  sam_hdr_t* out_hdr = sam_hdr_dup(inRecords.header());
  if (sam_hdr_write(out_fp, out_hdr) < 0) GError("Error writing header data to file %s\n", outfname.chars());

  const int keep = (options.keep_supplementary ? TB_KEEP_SUPP : 0) | (options.keep_secondary ? TB_KEEP_SECONDARY : 0) |
                   (options.keep_unmapped ? TB_KEEP_UNMAP : 0) | (options.store_frac ? TB_STORE_FRAC : 0);
  const int mode = mrgStrategy == tMrgStratFull ? TB_MODE_FULL : mrgStrategy == tMrgStratClip ? TB_MODE_CLIP :
                   mrgStrategy == tMrgStratExon ? TB_MODE_EXON : TB_MODE_CIGAR;
  // records buffered before a coverage gap closes the window (TB_WINDOW_RECORDS; host memory ~ 400 B per record)
  size_t window_min = 1u << 19;   // 512 K: with three pipeline stages the last window's pack + device + write is the exposed tail
  if (const char* e = getenv("TB_WINDOW_RECORDS")) { long v = atol(e); if (v > 0) window_min = (size_t)v; }

  // The window is cut only at a coordinate no read covers; an input without such a gap (a deep locus, DNA-seq) would buffer a
  // whole contig on the host and then overflow tb_collapse_window's n < 2^31: stop with a clear message before that
  // (the reference streams such inputs in O(k) memory; carrying the YD segment lists across windows is not built yet).
  size_t window_max = 400u << 20;
  if (const char* e = getenv("TB_WINDOW_MAX_RECORDS")) { long long v = atoll(e); if (v > 0) window_max = (size_t)v; }
  if (window_max > 2000000000u) window_max = 2000000000u;
  const bool dry_run = getenv("TB_DRYRUN") != NULL;   // reader self-check: no device, no output records, window statistics only
  TbWindowPacker packer; packer.init(numSamples);
  Reaper reaper;
  Writer writer; writer.fp = out_fp; writer.hdr = out_hdr; writer.reaper = &reaper;
  {
    int nt = (int)std::thread::hardware_concurrency(); if (nt > 16) nt = 16;
    if (const char* e = getenv("TB_PACK_THREADS")) nt = atoi(e);
    writer.nt = nt < 1 ? 1 : nt;
  }
  packer.writer = &writer;
  TbQueue queue;
  double t_create = 0; int n_devices = 1;
  std::thread device_thread([&] {   // owns the CUDA context: one submitting host thread per context
    if (dry_run) {   // checks the reader's contract: file order kept, windows separated by coverage gaps, nothing lost
      long nwin = 0, nrec = 0, bad_order = 0, bad_gap = 0; int last_tid = -1; uint64_t last_end = 0;
      while (std::unique_ptr<TbWindow> w = queue.pop()) {
        uint64_t lo = ~0ULL, hi = 0;
        for (auto& v : w->per_file)
          for (size_t i = 0; i < v.size(); ++i) {
            if (i && v[i]->start < v[i - 1]->start) ++bad_order;
            if (v[i]->refId() != w->tid) ++bad_order;
            if (v[i]->start < lo) lo = v[i]->start;
            if (v[i]->end > hi) hi = v[i]->end;
          }
        for (auto& v : w->per_file) for (GSamRecord* r : v) delete r;
        if (w->tid == last_tid && lo <= last_end) ++bad_gap;
        if (w->tid < last_tid) ++bad_order;
        last_tid = w->tid; last_end = hi; ++nwin; nrec += (long)w->n;
      }
      fprintf(stderr, "tb_b200 dry run: %ld windows, %ld records, %ld order violations, %ld gap violations\n", nwin, nrec, bad_order, bad_gap);
      return;
    }
    // TB_DEVICES=0,1 / 0-3: one worker thread, context and packer per GPU; the windows (independent by construction) are
    // computed wherever a GPU is free and written in hand-over order. Default: TB_DEVICE (or 0) alone.
    std::vector<int> devs;
    if (const char* de = getenv("TB_DEVICES")) {
      for (const char* q = de; *q;) {
        char* endp; long a = strtol(q, &endp, 10);
        if (endp == q) break;
        long b = a;
        if (*endp == '-') { q = endp + 1; b = strtol(q, &endp, 10); }
        for (long d = a; d <= b; ++d) devs.push_back((int)d);
        q = *endp ? endp + 1 : endp;
      }
    }
    if (devs.empty()) { const char* dev_env = getenv("TB_DEVICE"); devs.push_back(dev_env ? atoi(dev_env) : 0); }
    const int nd = (int)devs.size();
    queue.max_queued = (size_t)nd + 1;
    writer.cap = (size_t)nd + 1;
    writer.start();
    std::vector<TbWindowPacker> packers(nd, packer);
    std::vector<std::thread> workers;
    std::vector<double> creates(nd, 0.0);
    for (int d = 0; d < nd; ++d)
      workers.emplace_back([&, d] {
        auto c0 = clk::now();
        tb_ctx* ctx = tb_create(devs[d], numSamples, mode, options.flags, options.max_nh, options.min_qual, keep, options.collapse_same ? 1 : 0);
        if (!ctx) tb_die(tb_last_error(NULL));
        creates[d] = std::chrono::duration<double>(clk::now() - c0).count();
        while (std::unique_ptr<TbWindow> w = queue.pop()) packers[d].flush(ctx, *w);
        tb_destroy(ctx);
      });
    for (auto& t : workers) t.join();
    writer.finish();
    for (int d = 0; d < nd; ++d) {
      packer.t_pack += packers[d].t_pack; packer.t_device += packers[d].t_device; packer.n_windows += packers[d].n_windows;
      if (creates[d] > t_create) t_create = creates[d];
    }
    n_devices = nd;
  });

  // ---- parallel decode: the files are read independently (the device does the k-way merge), TB_DECODE_THREADS readers
  // each own a share of the files. A chunk = every record of the current tid that starts at or before `bound`; the window
  // is cut at the LAST coordinate of the chunk that no read covers (depth array over the chunk), the records behind the
  // cut stay queued for the next window; without such a coordinate the chunk is extended.
  struct Cursor { GSamRecord* carry = NULL; bool merged = false; };
  std::vector<Cursor> cur(numSamples);
  while (inRecords.recs.Count() > 0) {   // TInputFiles::start() read the first record of every file (tmerge.cpp:319-327)
    TInputRecord* r0 = inRecords.recs.Pop();
    cur[r0->fidx].carry = r0->brec; cur[r0->fidx].merged = r0->tbMerged;
    r0->disown(); delete r0;
  }
  int n_threads = (int)std::thread::hardware_concurrency();
  if (n_threads > 16) n_threads = 16;
  if (const char* e = getenv("TB_DECODE_THREADS")) n_threads = atoi(e);
  if (n_threads < 1) n_threads = 1;
  if (n_threads > numSamples) n_threads = numSamples;
  std::vector<std::vector<GSamRecord*>> pend(numSamples);   // records read but not yet handed over (current tid), per file
  std::vector<int32_t> depth;                               // difference array over [chunk_lo, bound + 1]
  auto tr0 = clk::now();
  uint64_t span = 1u << 21;   // first chunk length in bases (TB_WINDOW_SPAN); later chunks follow the read density
  if (const char* e = getenv("TB_WINDOW_SPAN")) { long v = atol(e); if (v > 0) span = (uint64_t)v; }
  for (;;) {
    int cur_tid = 0x7fffffff;
    for (int f = 0; f < numSamples; ++f)
      if (cur[f].carry) {
        GSamRecord* c = cur[f].carry;
        if (c->isUnmapped() || c->refId() < 0)   // the reference's own loop does not survive these either
          GError("Error: unmapped read %s in the input (not supported by tiebrush)\n", c->name());
        if (c->refId() < cur_tid) cur_tid = c->refId();
      }
    if (cur_tid == 0x7fffffff) break;   // every file is exhausted
    uint64_t chunk_lo = ~0ULL;
    for (int f = 0; f < numSamples; ++f)
      if (cur[f].carry && cur[f].carry->refId() == cur_tid && cur[f].carry->start < chunk_lo) chunk_lo = cur[f].carry->start;
    uint64_t bound = chunk_lo + span;
    bool tid_done = false;
    while (!tid_done) {
      // ---- read every file up to `bound` ----
      std::vector<std::thread> readers;
      for (int t = 0; t < n_threads; ++t)
        readers.emplace_back([&, t] {
          for (int f = t; f < numSamples; f += n_threads) {
            GSamRecord* c = cur[f].carry;
            GSamReader* rd = inRecords.freaders[f]->samreader;
            while (c && c->refId() == cur_tid && (uint64_t)c->start <= bound) { pend[f].push_back(c); c = rd->next(); }
            cur[f].carry = c;
          }
        });
      for (auto& th : readers) th.join();
      tid_done = true;
      for (int f = 0; f < numSamples; ++f) if (cur[f].carry && cur[f].carry->refId() == cur_tid) tid_done = false;
      size_t n_pend = 0;
      for (int f = 0; f < numSamples; ++f) n_pend += pend[f].size();
      uint64_t cut = 0;   // window = records with start < cut; 0 = no cut yet
      if (tid_done) cut = ~0ULL;
      else if (n_pend >= window_min) {
        // ---- depth over [chunk_lo, bound]: +1 at start, -1 one past end (clipped to the chunk); a coordinate of depth zero
        // is covered by no read that starts before it, and every such read has been decoded (start <= bound) ----
        const size_t m = (size_t)(bound - chunk_lo) + 1;
        depth.assign(m + 2, 0);
        for (int f = 0; f < numSamples; ++f)
          for (GSamRecord* r : pend[f]) {
            uint64_t e = r->end; if (e > bound) e = bound;
            depth[(size_t)(r->start - chunk_lo)] += 1;
            depth[(size_t)(e - chunk_lo) + 1] -= 1;
          }
        int64_t run = 0;
        for (size_t x = 0; x < m; ++x) { run += depth[x]; if (run == 0 && x > 0) cut = chunk_lo + x; }   // keeps the last one
      }
      if (cut == 0) {   // too few records or no gap inside the chunk: extend it (span follows the read density)
        if (n_pend > window_max)
          GError("Error: no coverage gap within %zu alignments on reference %d (coordinates %llu-%llu): the window cannot be cut. Raise TB_WINDOW_MAX_RECORDS (host memory ~400 B per alignment, at most 2e9) or split the input.\n",
                 n_pend, cur_tid, (unsigned long long)chunk_lo, (unsigned long long)bound);
        if (n_pend > 0) {
          const double dens = (double)n_pend / (double)(bound - chunk_lo + 1);
          const double want = (double)window_min / (dens > 1e-9 ? dens : 1e-9);
          span = want < 65536.0 ? 65536u : (want > 67108864.0 ? 67108864u : (uint64_t)want);
        }
        bound += span;
        continue;
      }
      // ---- hand over the records in front of the cut; the rest stays queued ----
      std::unique_ptr<TbWindow> win(new TbWindow(numSamples));
      win->tid = cur_tid;
      for (int f = 0; f < numSamples; ++f) {
        std::vector<GSamRecord*>& v = pend[f];
        size_t idx = v.size();
        if (cut != ~0ULL) {
          size_t lo = 0, hi = v.size();   // first record with start >= cut (file order = coordinate order)
          while (lo < hi) { const size_t mid = (lo + hi) >> 1; if ((uint64_t)v[mid]->start >= cut) hi = mid; else lo = mid + 1; }
          idx = lo;
        }
        win->per_file[f].assign(v.begin(), v.begin() + idx);
        win->file_merged[f] = cur[f].merged ? 1 : 0;
        win->n += idx;
        v.erase(v.begin(), v.begin() + idx);
      }
      if (win->n) queue.push(std::move(win));
      if (tid_done) break;
      chunk_lo = cut;
      if (bound < chunk_lo) bound = chunk_lo;
      bound += span;
    }
  }
  const double t_read = std::chrono::duration<double>(clk::now() - tr0).count();
  queue.finish();
  device_thread.join();
  const double t_before_reap = std::chrono::duration<double>(clk::now() - t_begin).count();
  reaper.finish();
  inRecords.stop();
  if (hts_close(out_fp) < 0) GError("Error closing output file %s\n", outfname.chars());
  sam_hdr_destroy(out_hdr);

  double p = 100.00 - (double)(outCounter * 100.00) / (double)inCounter;
  GMessage("%ld input records written as %ld (%.2f%% reduction)\n", inCounter, outCounter, p);
  if (getenv("TB_TIMING"))
    fprintf(stderr, "tb_b200 timing: total %.3f s | decode+merge %.3f | pack %.3f | device (H2D+kernels+D2H) %.3f | tag+write %.3f | windows %ld on %d GPU(s) | cuda init %.3f (overlapped) | reader thread includes waiting for the device thread | all windows written at %.3f s, records freed on a background thread\n",
            std::chrono::duration<double>(clk::now() - t_begin).count(), t_read, packer.t_pack, packer.t_device, writer.t_write, (long)packer.n_windows, n_devices, t_create, t_before_reap);
  return 0;
}
