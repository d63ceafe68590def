// tiecov_gpu — the reference `tiecov` command line (-c coverage bedGraph, -j junction BED, -s sample heat-map) with its hot
// loops on a B200.
//
// Compiles the UNMODIFIED reference source (src/tiecov.cpp, from where it lies under the reference checkout given with
// -I) with its main() renamed, and supplies a new main() that keeps option parsing (processOptions, tiecov.cpp:533-581),
// BAM decode (GSamReader), output file naming / track lines (tiecov.cpp:352-411) and the text formatting
// ("%s\t%d\t%d\t%.3f", "JUNC%08d") on the host, and replaces
//     bundle logic + addCov + flushCoverage + addJunction + flushJuncs           (tiecov.cpp:435-512)
// by a window packer and one tc_coverage_window() call per window (include/tiebrush_b200.h).
// A window is cut only where a new bundle starts (tid change or start > running max end, tiecov.cpp:443), so windows
// hold whole bundles. -s (addMean / discretize / normalize / flushCoverage of the pair vector, tiecov.cpp:155-185, 277-323)
// goes through tc_sample_window(). -W (BigWig) is not on the device path (libBigWig is not vendored): asking for it exits
// with an error rather than silently running CPU code.
#define main tc_reference_main_unused
#include "src/tiecov.cpp"
#undef main

#include <vector>
#include <chrono>
#include <string>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <deque>
#include <memory>
#include "tiebrush_b200.h"
#include <unistd.h>

// error exit from a worker thread: the reference's convention (message on stderr, exit code 1) without running exit handlers
// while the other threads of the pipeline are still alive
static void tb_die(const char* msg) { fputs(msg, stderr); fputc('\n', stderr); fflush(stderr); _exit(1); }

namespace {

// rows [a, b) of a text track, formatted by `fmt(i, buf)` on TB_PRINT_THREADS threads and written in order (the text is what
// fprintf would have produced: the same format strings)
template <class F>
static void print_rows(FILE* f, int64_t n, F fmt) {
  if (n <= 0) return;
  int nt = 8;
  if (const char* e = getenv("TB_PRINT_THREADS")) nt = atoi(e);
  if (nt < 1) nt = 1;
  if (n < 20000) nt = 1;
  std::vector<std::string> parts(nt);
  std::vector<std::thread> th;
  for (int t = 0; t < nt; ++t)
    th.emplace_back([&, t] {
      const int64_t a = n * t / nt, b = n * (t + 1) / nt;
      std::string& out = parts[t];
      out.reserve((size_t)(b - a) * 40);
      char buf[512];
      for (int64_t i = a; i < b; ++i) { const int len = fmt(i, buf, sizeof(buf)); out.append(buf, (size_t)len); }
    });
  for (auto& x : th) x.join();
  for (int t = 0; t < nt; ++t) fwrite(parts[t].data(), 1, parts[t].size(), f);
}

struct TcWindow {   // what the reader thread hands to the device thread: SoA columns of whole bundles
  std::vector<int32_t> tid, pos;
  std::vector<float> yc;
  std::vector<uint8_t> strand;
  std::vector<uint32_t> cig_off, cigar;
  std::vector<int32_t> yx;                                          // -s: YX tag per record
  size_t n() const { return pos.size(); }
  void add(GSamRecord& brec) {
    bam1_t* b = brec.get_b();
    tid.push_back(b->core.tid); pos.push_back((int32_t)b->core.pos);
    double w = 1.0;                                                  // tiecov.cpp:482-485
    if (brec.find_tag("YC") != NULL) w = brec.tag_float("YC");
    yc.push_back((float)w);
    strand.push_back((uint8_t)brec.spliceStrand());
    if (soutf) yx.push_back((int32_t)brec.tag_int("YX", 1));          // tiecov.cpp:495
    cig_off.push_back((uint32_t)cigar.size());
    const uint32_t* c = bam_get_cigar(b);
    cigar.insert(cigar.end(), c, c + b->core.n_cigar);
  }
};

struct TcWindowPacker {
  int n_samples = 0;                                                // -s: @CO SAMPLE lines of the header
  std::vector<int32_t> r_tid, r_start, r_end, j_tid, j_start, j_end, s_tid, s_start, s_end; std::vector<double> r_val, j_val, s_val; std::vector<uint8_t> j_strand;
  double t_device = 0, t_print = 0; int64_t n_windows = 0;

  void flush(tb_ctx* ctx, sam_hdr_t* hdr, TcWindow& w) {
    std::vector<int32_t>& tid = w.tid; std::vector<int32_t>& pos = w.pos; std::vector<float>& yc = w.yc; std::vector<uint8_t>& strand = w.strand;
    std::vector<uint32_t>& cig_off = w.cig_off; std::vector<uint32_t>& cigar = w.cigar; std::vector<int32_t>& yx = w.yx;
    const size_t m = w.n();
    if (m == 0) return;
    using clk = std::chrono::steady_clock;
    auto t0 = clk::now();
    cig_off.push_back((uint32_t)cigar.size());
    const int64_t ncig = (int64_t)cigar.size();
    if (cigar.empty()) cigar.push_back(0);
    tc_soa_in in; memset(&in, 0, sizeof(in));
    in.n = (int64_t)m; in.tid = tid.data(); in.pos = pos.data(); in.yc = yc.data(); in.strand = strand.data();
    in.cig_off = cig_off.data(); in.cigar = cigar.data(); in.on_device = 0; in.n_cig = ncig;
    const int64_t cap_r = 2 * ncig + 16, cap_j = ncig + 16;
    tc_runs_out runs; memset(&runs, 0, sizeof(runs));
    tc_juncs_out js; memset(&js, 0, sizeof(js));
    if (coutf) {
      r_tid.resize(cap_r); r_start.resize(cap_r); r_end.resize(cap_r); r_val.resize(cap_r);
      runs.capacity = cap_r; runs.tid = r_tid.data(); runs.start0 = r_start.data(); runs.end0 = r_end.data(); runs.value = r_val.data();
    }
    if (joutf) {
      j_tid.resize(cap_j); j_start.resize(cap_j); j_end.resize(cap_j); j_val.resize(cap_j); j_strand.resize(cap_j);
      js.capacity = cap_j; js.tid = j_tid.data(); js.start = j_start.data(); js.end = j_end.data(); js.strand = j_strand.data(); js.value = j_val.data();
    }
    if (coutf || joutf) {
      const int rc = tc_coverage_window(ctx, &in, coutf ? &runs : NULL, joutf ? &js : NULL);
      if (rc) tb_die(tb_last_error(ctx));                             // incl. the "unknown opcode" abort of tiecov.cpp:219-220
    }
    tc_runs_out rows; memset(&rows, 0, sizeof(rows));
    if (soutf) {
      s_tid.resize(cap_r); s_start.resize(cap_r); s_end.resize(cap_r); s_val.resize(cap_r);
      rows.capacity = cap_r; rows.tid = s_tid.data(); rows.start0 = s_start.data(); rows.end0 = s_end.data(); rows.value = s_val.data();
      if (tc_sample_window(ctx, &in, yx.data(), &rows)) tb_die(tb_last_error(ctx));
    }
    auto t1 = clk::now();
    if (coutf)                                                        // flushCoverage, tiecov.cpp:237
      print_rows(coutf, runs.n_runs, [&](int64_t i, char* buf, size_t cap) {
        return snprintf(buf, cap, "%s\t%d\t%d\t%.3f\n", hdr->target_name[r_tid[i]], r_start[i], r_end[i], r_val[i]); });
    if (joutf) {                                                      // CJunc::write, tiecov.cpp:91-95 (global counter)
      const int base = juncCount;
      print_rows(joutf, js.n_juncs, [&](int64_t i, char* buf, size_t cap) {
        return snprintf(buf, cap, "%s\t%d\t%d\tJUNC%08d\t%.3f\t%c\n", hdr->target_name[j_tid[i]], j_start[i] - 1, j_end[i], base + (int)i + 1, j_val[i], (char)j_strand[i]); });
      juncCount += (int)js.n_juncs;
    }
    if (soutf) {   // normalize(bsam, 0.1, 1.5, n_samples) + flushCoverage of the pair vector (tiecov.cpp:293-318)
      const float mint = 0.1, maxt = 1.5;
      const float denom = n_samples, mult = (maxt - mint);
      print_rows(soutf, rows.n_runs, [&](int64_t i, char* buf, size_t cap) {
        const uint64_t ival = (uint64_t)s_val[i];
        const float hval = ((float)ival / denom) * mult + mint;
        return snprintf(buf, cap, "%s\t%d\t%d\t%ld\t%f\n", hdr->target_name[s_tid[i]], s_start[i], s_end[i], (long)ival, hval); });
    }
    ++n_windows;
    auto t2 = clk::now();
    t_device += std::chrono::duration<double>(t1 - t0).count();
    t_print += std::chrono::duration<double>(t2 - t1).count();
  }
};

}  // namespace

int main(int argc, char* argv[]) {
  using clk = std::chrono::steady_clock;
  auto t_begin = clk::now();
  processOptions(argc, argv);
  if (bigwig) GError("Error: -W (BigWig) is not available in tiecov_gpu (libBigWig is not part of the device path)\n");
  GSamReader samreader(infname.chars(), SAM_QNAME | SAM_FLAG | SAM_RNAME | SAM_POS | SAM_CIGAR | SAM_AUX);
  // output files: same naming rule (suffix appended unless already there), same track lines as tiecov.cpp:352-411
  auto open_track = [](GStr& name, const char* suffix, const char* track_line) -> FILE* {
    const int sl = (int)strlen(suffix);
    if (name.length() < sl || strcmp(name.chars() + name.length() - sl, suffix) != 0) name.append(suffix);
    FILE* f = fopen(name.chars(), "w");
    if (f == NULL) GError("Error creating file %s\n", name.chars());
    fputs(track_line, f);
    return f;
  };
  if (!covfname.is_empty())
    coutf = (covfname == "-" || covfname == "stdout") ? stdout : open_track(covfname, ".bedgraph", "track type=bedGraph\n");
  if (!jfname.is_empty()) joutf = open_track(jfname, ".bed", "track name=junctions\n");
  if (!sfname.is_empty())
    soutf = open_track(sfname, ".bedgraph", "track type=bedGraph name=\"Sample Count Heatmap\" description=\"Sample Count Heatmap\" visibility=full graphType=\"heatmap\" color=200,100,0 altColor=0,100,200\n");
  // ---- TB_DEVICES=0,1,... or 0-7: the stream sharded over several GPUs of the box (-c / -j). One thread per GPU, each with its
  // own context and NCCL rank; the reader hands over ROUNDS of whole bundles, a round is split by record count into one
  // slice per GPU (the cuts fall anywhere: tc_shard_coverage moves the records of a bundle that a cut separated from its
  // head to the GPU that holds the head), the rows are printed in GPU order with one running JUNC counter. ----
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
  if (devs.size() > 1) {
    if (soutf) GError("Error: -s is not available with TB_DEVICES (one GPU computes the sample heat-map)\n");
    const int W = (int)devs.size();
    unsigned char nccl_id[128];
    if (tb_comm_unique_id(nccl_id)) GError("%s\n", tb_last_error(NULL));
    size_t round_min = (size_t)W * (2u << 20);   // records per round (TB_WINDOW_RECORDS = per GPU)
    if (const char* e = getenv("TB_WINDOW_RECORDS")) { long v = atol(e); if (v > 0) round_min = (size_t)v * W; }
    struct RankOut { std::vector<int32_t> r_tid, r_start, r_end, j_tid, j_start, j_end; std::vector<double> r_val, j_val; std::vector<uint8_t> j_strand; int64_t nr = 0, nj = 0; };
    std::vector<RankOut> outs(W);
    std::mutex qm; std::condition_variable qcv; std::deque<std::unique_ptr<TcWindow>> q; bool q_done = false;
    // a reusable barrier over the W device threads
    std::mutex bm; std::condition_variable bcv; int b_wait = 0; long b_gen = 0;
    auto barrier = [&] {
      std::unique_lock<std::mutex> lk(bm);
      const long g = b_gen;
      if (++b_wait == W) { b_wait = 0; ++b_gen; bcv.notify_all(); }
      else bcv.wait(lk, [&] { return b_gen != g; });
    };
    std::unique_ptr<TcWindow> cur; bool all_done = false;
    double t_device = 0, t_print = 0; long n_rounds = 0;
    sam_hdr_t* hdr = samreader.header();
    auto device_main = [&](int r) {
      tb_ctx* ctx = tb_create(devs[r], 1, TB_MODE_CIGAR, 0, TB_NO_MAX_NH, -1, 0, 0);
      if (!ctx) tb_die(tb_last_error(NULL));
      if (tb_comm_init(ctx, r, W, nccl_id)) tb_die(tb_last_error(ctx));
      for (;;) {
        if (r == 0) {   // take the next round
          std::unique_lock<std::mutex> lk(qm);
          qcv.wait(lk, [&] { return !q.empty() || q_done; });
          if (q.empty()) all_done = true;
          else { cur = std::move(q.front()); q.pop_front(); qcv.notify_all(); }
        }
        barrier();
        if (all_done) break;
        auto t0 = clk::now();
        TcWindow& w = *cur;
        const int64_t n = (int64_t)w.n(), a = n * r / W, b = n * (r + 1) / W;
        tc_soa_in in; memset(&in, 0, sizeof(in));
        in.n = b - a; in.tid = w.tid.data() + a; in.pos = w.pos.data() + a; in.yc = w.yc.data() + a; in.strand = w.strand.data() + a;
        in.cig_off = w.cig_off.data() + a; in.cigar = w.cigar.data(); in.on_device = 0;
        in.n_cig = (int64_t)(w.cig_off[b] - w.cig_off[a]);
        RankOut& o = outs[r];
        // rows of this GPU: its own records' worth plus what the lead exchange may bring in from the right
        const int64_t cap_r = 2 * (int64_t)w.cigar.size() + 16, cap_j = (int64_t)w.cigar.size() + 16;
        tc_runs_out runs; memset(&runs, 0, sizeof(runs)); tc_juncs_out js; memset(&js, 0, sizeof(js));
        o.r_tid.resize(cap_r); o.r_start.resize(cap_r); o.r_end.resize(cap_r); o.r_val.resize(cap_r);
        runs.capacity = cap_r; runs.tid = o.r_tid.data(); runs.start0 = o.r_start.data(); runs.end0 = o.r_end.data(); runs.value = o.r_val.data();
        o.j_tid.resize(cap_j); o.j_start.resize(cap_j); o.j_end.resize(cap_j); o.j_val.resize(cap_j); o.j_strand.resize(cap_j);
        js.capacity = cap_j; js.tid = o.j_tid.data(); js.start = o.j_start.data(); js.end = o.j_end.data(); js.strand = o.j_strand.data(); js.value = o.j_val.data();
        const int rc = tc_shard_coverage(ctx, &in, 1, 1 << 26, coutf ? &runs : NULL, joutf ? &js : NULL);   // collective: every GPU thread is here
        if (rc) tb_die(tb_last_error(ctx));
        o.nr = coutf ? runs.n_runs : 0; o.nj = joutf ? js.n_juncs : 0;
        barrier();
        if (r == 0) {
          auto t1 = clk::now();
          for (int g = 0; g < W; ++g) {
            RankOut& og = outs[g];
            if (coutf) print_rows(coutf, og.nr, [&](int64_t i, char* buf, size_t cap) {
              return snprintf(buf, cap, "%s\t%d\t%d\t%.3f\n", hdr->target_name[og.r_tid[i]], og.r_start[i], og.r_end[i], og.r_val[i]); });
            if (joutf) {
              const int base = juncCount;
              print_rows(joutf, og.nj, [&](int64_t i, char* buf, size_t cap) {
                return snprintf(buf, cap, "%s\t%d\t%d\tJUNC%08d\t%.3f\t%c\n", hdr->target_name[og.j_tid[i]], og.j_start[i] - 1, og.j_end[i], base + (int)i + 1, og.j_val[i], (char)og.j_strand[i]); });
              juncCount += (int)og.nj;
            }
          }
          cur.reset();
          ++n_rounds;
          t_device += std::chrono::duration<double>(t1 - t0).count();
          t_print += std::chrono::duration<double>(clk::now() - t1).count();
        }
        barrier();
      }
      tb_destroy(ctx);
    };
    std::vector<std::thread> dth;
    for (int r = 0; r < W; ++r) dth.emplace_back(device_main, r);
    auto hand_over = [&](std::unique_ptr<TcWindow>& w) {
      if (w->n() == 0) return;
      w->cig_off.push_back((uint32_t)w->cigar.size());
      if (w->cigar.empty()) w->cigar.push_back(0);
      std::unique_lock<std::mutex> lk(qm);
      qcv.wait(lk, [&] { return q.size() < 2; });
      q.push_back(std::move(w)); qcv.notify_all();
      w.reset(new TcWindow());
    };
    std::unique_ptr<TcWindow> win(new TcWindow());
    int prev_tid = -1, b_end = 0;
    GSamRecord brec;
    while (samreader.next(brec)) {
      if (brec.isUnmapped()) continue;                                 // tiecov.cpp:436-438
      const bool new_bundle = brec.refId() != prev_tid || (int)brec.start > b_end;   // tiecov.cpp:443
      if (new_bundle) {
        if (win->n() >= round_min) hand_over(win);
        b_end = brec.end; prev_tid = brec.refId();
      } else if (b_end < (int)brec.end) b_end = brec.end;
      win->add(brec);
    }
    hand_over(win);
    { std::lock_guard<std::mutex> lk(qm); q_done = true; qcv.notify_all(); }
    for (auto& t : dth) t.join();
    if (coutf && coutf != stdout) fclose(coutf);
    if (joutf) fclose(joutf);
    if (getenv("TB_TIMING"))
      fprintf(stderr, "tb_b200 timing: total %.3f s | %d GPUs | device (H2D+halo exchange+kernels+D2H) %.3f | print %.3f | rounds %ld\n",
              std::chrono::duration<double>(clk::now() - t_begin).count(), W, t_device, t_print, n_rounds);
    return 0;
  }
  const char* dev_env = getenv("TB_DEVICE");
  tb_ctx* ctx = tb_create(dev_env ? atoi(dev_env) : 0, 1, TB_MODE_CIGAR, 0, TB_NO_MAX_NH, -1, 0, 0);
  if (!ctx) GError("%s\n", tb_last_error(NULL));
  size_t window_min = 8u << 20;   // records buffered before a bundle boundary closes the window (TB_WINDOW_RECORDS)
  if (const char* e = getenv("TB_WINDOW_RECORDS")) { long v = atol(e); if (v > 0) window_min = (size_t)v; }

  TcWindowPacker packer;
  if (soutf) {   // needs the @CO SAMPLE lines of a TieBrush-made header (aborts without them, like the reference)
    load_sample_info(samreader.header(), sample_info);
    packer.n_samples = (int)sample_info.size();
  }
  // host pipeline: this thread decodes and packs window w+1 while the device thread runs window w and
  // formats / writes its rows; at most two windows wait in the queue
  std::mutex qm; std::condition_variable qcv; std::deque<std::unique_ptr<TcWindow>> q; bool q_done = false;
  std::thread device_thread([&] {
    for (;;) {
      std::unique_ptr<TcWindow> w;
      {
        std::unique_lock<std::mutex> lk(qm);
        qcv.wait(lk, [&] { return !q.empty() || q_done; });
        if (q.empty()) return;
        w = std::move(q.front()); q.pop_front(); qcv.notify_all();
      }
      packer.flush(ctx, samreader.header(), *w);
    }
  });
  auto hand_over = [&](std::unique_ptr<TcWindow>& w) {
    if (w->n() == 0) return;
    std::unique_lock<std::mutex> lk(qm);
    qcv.wait(lk, [&] { return q.size() < 2; });
    q.push_back(std::move(w)); qcv.notify_all();
    w.reset(new TcWindow());
  };
  std::unique_ptr<TcWindow> win(new TcWindow());
  int prev_tid = -1, b_end = 0;
  GSamRecord brec;
  auto tr0 = clk::now();
  while (samreader.next(brec)) {
    if (brec.isUnmapped()) continue;                                 // tiecov.cpp:436-438
    const bool new_bundle = brec.refId() != prev_tid || (int)brec.start > b_end;   // tiecov.cpp:443
    if (new_bundle) {
      if (win->n() >= window_min) hand_over(win);
      b_end = brec.end; prev_tid = brec.refId();
    } else if (b_end < (int)brec.end) b_end = brec.end;
    win->add(brec);
  }
  hand_over(win);
  const double t_read = std::chrono::duration<double>(clk::now() - tr0).count();
  { std::lock_guard<std::mutex> lk(qm); q_done = true; qcv.notify_all(); }
  device_thread.join();
  if (coutf && coutf != stdout) fclose(coutf);
  if (joutf) fclose(joutf);
  if (soutf) fclose(soutf);
  tb_destroy(ctx);
  if (getenv("TB_TIMING"))
    fprintf(stderr, "tb_b200 timing: total %.3f s | decode+pack (reader thread, incl. waiting for the device thread) %.3f | device (H2D+kernels+D2H) %.3f | print %.3f | windows %ld\n",
            std::chrono::duration<double>(clk::now() - t_begin).count(), t_read, packer.t_device, packer.t_print, (long)packer.n_windows);
  return 0;
}
