// Device-resident batched controller (placeholder until the ensemble driver lands).
#include "vk_internal.cuh"
struct EnsState { int dummy; };
namespace vk {
void ens_destroy(vk_column *c) { if (c->ens) { delete c->ens; c->ens = nullptr; } }
}
using namespace vk;
extern "C" {
int vk_ens_setup(vk_column *, const vk_ens_opts *) { set_error("ensemble driver not built"); return VK_ERR_UNSUPPORTED; }
int vk_ens_set_state(vk_column *, const double *, const double *) { set_error("ensemble driver not built"); return VK_ERR_UNSUPPORTED; }
int vk_ens_run(vk_column *, int) { set_error("ensemble driver not built"); return VK_ERR_UNSUPPORTED; }
int vk_ens_get_state(vk_column *, double *, double *, double *, int *, int *) { set_error("ensemble driver not built"); return VK_ERR_UNSUPPORTED; }
}
