/* TEST / BENCH INFRASTRUCTURE ONLY (built by oracle/build_ref.sh against the reference's vendored htslib into oracle/_ref/hts_tool).
 *   hts_tool tobam  in.sam out.bam   : SAM text -> BAM (the reference ships no converter and samtools is not in the image)
 *   hts_tool decode in.bam [...]     : decode-only pass (sam_read1 loop, htslib sam.c), prints "records seconds" per file:
 *                                      the host decode time SURVEY §8d asks to break out of the end-to-end numbers */
#include <stdio.h>
#include <string.h>
#include <time.h>
#include "htslib/sam.h"

static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

int main(int argc, char** argv) {
  if (argc >= 4 && strcmp(argv[1], "tobam") == 0) {
    htsFile* in = hts_open(argv[2], "r");
    if (!in) { fprintf(stderr, "hts_tool: cannot open %s\n", argv[2]); return 1; }
    sam_hdr_t* h = sam_hdr_read(in);
    htsFile* out = hts_open(argv[3], "wb");
    if (!h || !out || sam_hdr_write(out, h) < 0) { fprintf(stderr, "hts_tool: cannot write %s\n", argv[3]); return 1; }
    bam1_t* b = bam_init1();
    long n = 0;
    while (sam_read1(in, h, b) >= 0) { if (sam_write1(out, h, b) < 0) { fprintf(stderr, "hts_tool: write error\n"); return 1; } ++n; }
    bam_destroy1(b); sam_hdr_destroy(h); hts_close(in);
    if (hts_close(out) < 0) return 1;
    printf("%ld\n", n);
    return 0;
  }
  if (argc >= 3 && strcmp(argv[1], "decode") == 0) {
    long total = 0; const double t0 = now();
    for (int a = 2; a < argc; ++a) {
      htsFile* in = hts_open(argv[a], "r");
      if (!in) { fprintf(stderr, "hts_tool: cannot open %s\n", argv[a]); return 1; }
      sam_hdr_t* h = sam_hdr_read(in);
      bam1_t* b = bam_init1();
      while (sam_read1(in, h, b) >= 0) ++total;
      bam_destroy1(b); sam_hdr_destroy(h); hts_close(in);
    }
    printf("%ld %.6f\n", total, now() - t0);
    return 0;
  }
  fprintf(stderr, "usage: hts_tool tobam in.sam out.bam | hts_tool decode in.bam [...]\n");
  return 2;
}
