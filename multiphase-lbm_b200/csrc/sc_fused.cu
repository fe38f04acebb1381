// sc_fused.cu -- fused plane-marching Shan-Chen step (placeholder: routes to the staged kernels
// until the marching kernel lands).
#include "clbm_internal.h"

namespace clbm {
int sc_psi_all(clbm_ctx *c);
int sc_collide_all(clbm_ctx *c);

int sc_fused_step(clbm_ctx *c)
{
    int rc = sc_psi_all(c);
    if (rc) return rc;
    rc = sc_collide_all(c);
    if (rc) return rc;
    c->parity = 1 - c->parity;
    return 0;
}
}  // namespace clbm
