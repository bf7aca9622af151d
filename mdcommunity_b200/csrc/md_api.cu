// Lock-step greedy rollout driver: the body of GetSol's `while not isTerminal()` loop
// (reference MultiDismantler_torch.py:759-784) for a whole batch, enqueued n_steps times on one stream.
#include "md_common.cuh"

extern int g_md_debug_flags;

extern "C" int md_rollout_steps(const md_graphs_t* g, const md_state_t* s, const md_weights_t* w, const md_work_t* k,
                                int32_t n_steps, md_stream_t stream) {
  if (!g || !k) return MD_EINVAL;
  const bool fused = k->fused_ws != nullptr && md_fused_applies(g) && !(g_md_debug_flags & 4);
  for (int i = 0; i < n_steps; ++i) {
    int rc;
    if (fused) {
      rc = md_embed_fused(g, s, w, k, stream);
      if (rc) return rc;
    } else {
      rc = md_embed_init(g, s, w, k, stream);
      if (rc) return rc;
      for (int it = 0; it < MD_BP_ITERS; ++it) {
        rc = md_embed_iter(g, s, w, k, it, stream);
        if (rc) return rc;
      }
    }
    rc = md_graph_embed(g, s, w, k, stream);
    if (rc) return rc;
    rc = md_qhead(g, s, w, k, stream);
    if (rc) return rc;
    rc = md_env_step(g, s, nullptr, 0, -1, stream);
    if (rc) return rc;
  }
  return 0;
}

extern "C" void md_abi_sizes(int32_t* out4) {
  out4[0] = (int32_t)sizeof(md_graphs_t);
  out4[1] = (int32_t)sizeof(md_state_t);
  out4[2] = (int32_t)sizeof(md_weights_t);
  out4[3] = (int32_t)sizeof(md_work_t);
}

extern "C" const char* md_version(void) { return "mdcommunity_b200 0.1 (sm_100a)"; }
