// collapse_internal.cuh — stage interfaces of the collapse pipeline (collapse.cu drives them).
#pragma once
#include "tb_record.cuh"

struct ColGeom {
  int64_t n; int k; uint32_t W;       // records, files, 32-bit words per sample bitset
  uint32_t S;                         // window span in positions
  const uint32_t* P;                  // [S+1] merged-order rank of every position (exclusive scan of the histogram)
  const long long* d_runoff;          // [k+1] device copy of run_off
  const uint8_t* d_merged;            // [k] device copy of file_merged (or nullptr)
  long long* d_status;                // CS_* block
  int64_t n_cig;                      // words in the CIGAR arena as the caller stated them (0 = unknown)
};

// dense group outputs of a front end, in final order
struct ColGroups {
  int64_t capacity;
  uint32_t* rep; float* yc; uint32_t* yx; int32_t* yd;   // yd is pre-set by the front end (0, or the carried YD tag maximum)
  uint32_t* bits;                                        // [G*W] direct-sample bitsets (XB_BITS)
};

// Fast path: shared-memory hash tiles. Returns 0 ok, 1 error, 2 = group table overflow (caller falls back to the ordered path).
int col_front_tile(tb_ctx* ctx, const ColIn& in, const ColGeom& g, ColGroups& out, int64_t* n_groups, int64_t* n_kept);
// Exact path: sequential emulation of the reference's per-position sorted list, in merge order.
int col_front_ordered(tb_ctx* ctx, const ColIn& in, const ColGeom& g, ColGroups& out, int64_t* n_groups, int64_t* n_kept);
// YD chains over the dense groups.
int col_yd(tb_ctx* ctx, const ColIn& in, const ColGeom& g, const ColGroups& grp, int64_t G);
