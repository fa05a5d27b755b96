// fused_f32.cu -- placeholder until the fused kernels land.
#include "lnb_internal.h"
int lnb_fused_step(lnb_ctx *ctx, const lnb_mlp *mlp, const lnb_step_args *a, bool nerf)
{
    (void)mlp; (void)a; (void)nerf;
    ctx->err = "fused path: problem shape not supported";
    return LNB_ERR_UNSUPPORTED;
}
