// placeholder, replaced below
#include "common.cuh"
namespace satfill {
int build_hierarchy(sa_scene* s, const sa_options&) { return fail(s->ctx, SA_BAD_ARGUMENT, "multigrid not built"); }
void free_hierarchy(sa_scene*) {}
int apply_vcycle(sa_scene* s, const sa_options&, KernelTimer&) { return fail(s->ctx, SA_BAD_ARGUMENT, "multigrid not built"); }
}
