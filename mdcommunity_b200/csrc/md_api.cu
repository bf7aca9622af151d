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

// ---- CUDA graph of n lock-step iterations --------------------------------------------------------------------------
// The iteration is launch-bound towards the end of a rollout (5 short kernels per step): capturing a chunk of steps once
// and replaying it removes the per-launch gaps.  All kernel arguments are the struct VALUES (device pointers, sizes), so
// the graph stays valid as long as the caller keeps the same buffers; call md_rollout_steps once before capturing (the
// one-time cudaFuncSetAttribute calls must not fall into the capture).
extern "C" int md_rollout_graph_create(const md_graphs_t* g, const md_state_t* s, const md_weights_t* w, const md_work_t* k,
                                       int32_t n_steps, md_stream_t stream, void** exec_out) {
  if (!g || !s || !w || !k || !exec_out || n_steps < 1) return MD_EINVAL;
  if (!md_env_uses_small_path(g)) return MD_ETOOBIG;  // the cooperative large-graph env kernel is not captured
  (void)stream;  // the capture runs on a private stream (the legacy default stream cannot capture); nothing executes here
  cudaStream_t st = nullptr;
  cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  if (e != cudaSuccess) return (int)e;
  cudaGraph_t graph = nullptr;
  e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
  if (e != cudaSuccess) {
    cudaStreamDestroy(st);
    return (int)e;
  }
  const int rc = md_rollout_steps(g, s, w, k, n_steps, (md_stream_t)st);
  e = cudaStreamEndCapture(st, &graph);
  cudaStreamDestroy(st);
  if (rc || e != cudaSuccess) {
    if (graph) cudaGraphDestroy(graph);
    return rc ? rc : (int)e;
  }
  cudaGraphExec_t exec = nullptr;
  e = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) return (int)e;
  *exec_out = exec;
  return 0;
}

extern "C" int md_rollout_graph_launch(void* exec, md_stream_t stream) {
  if (!exec) return MD_EINVAL;
  return (int)cudaGraphLaunch((cudaGraphExec_t)exec, (cudaStream_t)stream);
}

extern "C" int md_rollout_graph_destroy(void* exec) {
  if (!exec) return 0;
  return (int)cudaGraphExecDestroy((cudaGraphExec_t)exec);
}

extern "C" void md_abi_sizes(int32_t* out4) {
  out4[0] = (int32_t)sizeof(md_graphs_t);
  out4[1] = (int32_t)sizeof(md_state_t);
  out4[2] = (int32_t)sizeof(md_weights_t);
  out4[3] = (int32_t)sizeof(md_work_t);
}

extern "C" const char* md_version(void) { return "mdcommunity_b200 0.1 (sm_100a)"; }
