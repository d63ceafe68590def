#include "tb_common.cuh"
int tb_collapse_impl(tb_ctx* ctx, const tb_soa_in* in, tb_groups_out* out) { ctx->set_error("collapse: not built yet"); return 1; }
