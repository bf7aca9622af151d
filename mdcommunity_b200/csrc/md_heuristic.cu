// Model-free scorers on the same device-resident environment: the heuristic baselines the reference compares against.
// Replaces (reference, /root/reference/code/MultiDismantler_unit_cost/baseline):
//   HDA/hda_2max.py:50-69   delnode: key = max(deg_1(v), deg_2(v)) in the mutually pruned layers (hda_add.py: deg_1 + deg_2)
//   CI/ci_max.py:52-90      basic_ci: (deg(v) - 1) * sum_{u in N(v)} (deg(u) - 1), -1 for isolated nodes; key = max over
//                           layers (ci_add.py: sum)
// `sorted(..., reverse=True)[0]` is stable: the first node in G.nodes() order, i.e. the lowest id among the maxima.
// The kernel writes the same packed arg-max key md_qhead leaves in s->best, so md_env_step consumes it unchanged
// (remove the node, recompute the LMCC, accumulate the AUDC = critical_number's loop body, hda_2max.py:71-93).
// Only active nodes (alive degree > 0) are candidates: while LMCC > 1 some node has an alive edge, and every such
// node beats every isolated one under both keys.
#include "md_common.cuh"

namespace md {

__global__ void __launch_bounds__(128) heuristic_select_kernel(md_graphs_t G, md_state_t S, int kind, int combine) {
  const int tpg = G.stride / TILE_M;
  const int g = blockIdx.x / tpg, t = blockIdx.x - g * tpg;
  const int nact = S.n_act[g];
  if (t * TILE_M >= nact || S.done[g]) return;
  const int base = g * G.stride;
  const int p = t * TILE_M + threadIdx.x;
  unsigned long long key = 0ull;
  if (p < nact) {
    const int v = S.act_list[base + p];
    long long score = 0;
    for (int l = 0; l < G.n_layers; ++l) {
      const long long d = S.deg[l][v];
      long long s;
      if (kind == 0) {
        s = d;
      } else {
        long long nb = 0;
        const int32_t* __restrict__ acol = S.acol[l];
        for (int e = S.aptr[l][base + p], e1 = S.aptr[l][base + p + 1]; e < e1; ++e) nb += S.deg[l][acol[e]] - 1;
        s = d == 0 ? -1 : nb * (d - 1);
      }
      score = l == 0 ? s : (combine == 0 ? max(score, s) : score + s);
    }
    // key = (score + 1, saturated to 32 bits) << 32 | ~local id : larger score wins, then the lower id
    unsigned long long hi = (unsigned long long)max(score + 1, 0ll);
    if (hi > 0xffffffffull) hi = 0xffffffffull;
    key = (hi << 32) | (unsigned long long)(0xffffffffu - (uint32_t)(v - base));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
    key = other > key ? other : key;
  }
  if ((threadIdx.x & 31) == 0 && key != 0ull) atomicMax(S.best + g, key);
}

}  // namespace md

extern "C" int md_heuristic_select(const md_graphs_t* g, const md_state_t* s, int32_t kind, int32_t combine,
                                   md_stream_t stream) {
  if (!g || !s) return MD_EINVAL;
  if (g->n_layers < 1 || g->n_layers > MD_MAX_LAYERS) return MD_ELAYERS;
  if (g->stride % MD_TILE_M != 0 || g->n_total != g->n_graphs * g->stride) return MD_EINVAL;
  if (kind < 0 || kind > 1 || combine < 0 || combine > 1) return MD_EINVAL;
  md::heuristic_select_kernel<<<g->n_graphs * (g->stride / MD_TILE_M), 128, 0, (cudaStream_t)stream>>>(*g, *s, kind, combine);
  MD_CHECK_LAUNCH();
  return 0;
}
