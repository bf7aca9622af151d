"""The Q network of the DQN training step -- forward, loss, backward and Adam -- on the hand-written FP32 kernels of
csrc/md_train.cu + the CSR SpMM (csrc/md_spmm.cu).  No torch matmul / normalize / tanh / softmax / autograd formula and
no cuBLAS on this path: torch supplies device memory, index plumbing (building the CSR operands) and, for callers that
want `loss.backward()`, an autograd.Function each around the whole network and the whole loss that only hands the
hand-written gradients to the engine.

Reference arithmetic restated here (file:line under /root/reference/code/MultiDismantler_unit_cost):
  MultiDismantler_net_graphsage.py:92-241   train_forward (and :243-394 test_forward, same arithmetic on all nodes)
  MRGNN/mutil_layer_weight.py:266-313       BitwiseMultipyLogis / LogisticVector
  MultiDismantler_torch.py:410-431          calc_loss (MSE + Alpha * Laplacian reconstruction, global edge-count normaliser)
  MultiDismantler_torch.py:136              torch.optim.Adam(lr = LEARNING_RATE)

Layout.  The two layers share every weight, so they are stacked along the row dimension: with n compact node rows and
B graph rows ("virtual nodes", the y_* tensors of the reference) per layer, N = n + B, every per-layer matrix is one
[2N, 64] matrix (layer l = rows l*N .. (l+1)*N, node rows first, then the B graph rows).  Neighbour aggregation AND the
sub-graph sum (n2nsum_param / subgsum_param) are ONE block-diagonal CSR operand of 2N rows: node row -> its alive
neighbours, graph row -> the node rows of that graph.  So an embedding iteration is 1 SpMM + 4 GEMMs + 1 normalise for
both layers and both row kinds, and the weight gradients need no cross-layer accumulation.
"""
import ctypes as C

import torch

from . import _lib
from .ops import coo_to_csr, spmm_csr

EMBED = 64
PARAM_KEYS = ("w_n2l", "p_node_conv", "p_node_conv2", "p_node_conv3", "h1_weight", "h2_weight", "cross_product", "w_layer1",
              "w_layer2", "layerNodeAttention_weight.trans", "layerNodeAttention_weight.bias",
              "layerNodeAttention_weight.logis.parameter.weight", "layerNodeAttention_weight.logis.parameter.bias")

_WS = {}


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _st(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _chk(t):
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise ValueError("dense kernels take contiguous CUDA float32 tensors (no CPU fallback)")
    return t


def _ws(dev):
    w = _WS.get(dev)
    if w is None:
        w = _WS[dev] = torch.empty(64 * 128 * 128, dtype=torch.float32, device=dev)
    return w


# ---- thin wrappers (one C-ABI call each) ---------------------------------------------------------------------------
def gemm(a, b, ta=False, tb=False, out=None, acc=False, relu_a=False):
    _chk(a), _chk(b)
    M, K = (a.shape[1], a.shape[0]) if ta else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[0], b.shape[1]) if tb else (b.shape[1], b.shape[0])
    assert K == Kb, (a.shape, b.shape, ta, tb)
    if out is None:
        assert not acc
        out = torch.empty(M, N, dtype=torch.float32, device=a.device)
    assert _chk(out).numel() == M * N
    if M * N == 0:
        return out
    if K == 0:
        return out if acc else out.zero_()
    ws = _ws(a.device)
    _lib.check(_lib.lib().md_gemm_f32(_p(a), _p(b), _p(out), M, N, K, int(ta), int(tb), int(acc), int(relu_a), _p(ws), ws.numel(), _st(a)),
               "md_gemm_f32")
    return out


def norm_fwd(z, relu):
    h = torch.empty_like(_chk(z))
    if z.numel():
        _lib.check(_lib.lib().md_norm_fwd_f32(_p(z), _p(h), z.shape[0], z.shape[1], int(relu), _st(z)), "md_norm_fwd_f32")
    return h


def norm_bwd(z, dh, relu):
    dz = torch.empty_like(_chk(z))
    if z.numel():
        _lib.check(_lib.lib().md_norm_bwd_f32(_p(z), _p(_chk(dh)), _p(dz), z.shape[0], z.shape[1], int(relu), _st(z)), "md_norm_bwd_f32")
    return dz


def tanh_bias(z, b):
    t = torch.empty_like(_chk(z))
    if z.numel():
        _lib.check(_lib.lib().md_tanh_bias_fwd_f32(_p(z), _p(_chk(b)), _p(t), z.shape[0], z.shape[1], _st(z)), "md_tanh_bias_fwd_f32")
    return t


def tanh_bwd(t, dt):
    dz = torch.empty_like(_chk(t))
    if t.numel():
        _lib.check(_lib.lib().md_tanh_bwd_f32(_p(t), _p(_chk(dt)), _p(dz), t.shape[0], t.shape[1], _st(t)), "md_tanh_bwd_f32")
    return dz


def colsum(x, out):
    n, D = (x.shape[0], x.shape[1]) if x.dim() == 2 else (x.shape[0], 1)
    assert _chk(out).numel() == D
    if n == 0:
        return out.zero_()
    ws = _ws(x.device)
    _lib.check(_lib.lib().md_colsum_f32(_p(_chk(x)), _p(out), n, D, _p(ws), ws.numel(), _st(x)), "md_colsum_f32")
    return out


def rowscale(h, s, gid, out=None):
    out = torch.empty_like(_chk(h)) if out is None else _chk(out)
    if h.numel():
        _lib.check(_lib.lib().md_rowscale_f32(_p(h), _p(_chk(s)), _p(gid), _p(out), h.shape[0], h.shape[1], _st(h)), "md_rowscale_f32")
    return out


def rowdot(a, b):
    out = torch.empty(a.shape[0], dtype=torch.float32, device=a.device)
    if a.numel():
        _lib.check(_lib.lib().md_rowdot_f32(_p(_chk(a)), _p(_chk(b)), _p(out), a.shape[0], a.shape[1], _st(a)), "md_rowdot_f32")
    return out


def seg_reduce(x, y, off):
    B = off.numel() - 1
    out = torch.empty(B, dtype=torch.float32, device=off.device)
    if x.numel() == 0:
        return out.zero_()
    _lib.check(_lib.lib().md_seg_reduce_f32(_p(_chk(x)), _p(y), _p(off), _p(out), B, _st(x)), "md_seg_reduce_f32")
    return out


def axpby(a, b, alpha, beta, out=None):
    out = torch.empty_like(_chk(a)) if out is None else _chk(out)
    if a.numel():
        _lib.check(_lib.lib().md_axpby_f32(_p(a), _p(b), float(alpha), float(beta), _p(out), a.numel(), _st(a)), "md_axpby_f32")
    return out


def relu_mask(x, dx):
    if x.numel():
        _lib.check(_lib.lib().md_relu_mask_f32(_p(_chk(x)), _p(_chk(dx)), x.numel(), _st(x)), "md_relu_mask_f32")
    return dx


def dot(a, b):
    out = torch.zeros(1, dtype=torch.float32, device=a.device)
    if a.numel():
        _lib.check(_lib.lib().md_dot_f32(_p(_chk(a)), _p(_chk(b)), _p(out), a.numel(), _st(a)), "md_dot_f32")
    return out


# ---- operands --------------------------------------------------------------------------------------------------------
class Operands:
    """What the network reads besides its weights (all on the device):
      n, B, N = n + B          compact node rows / graphs / rows per layer
      X      [2N, 2]           node features (node rows) and ones (graph rows), both layers stacked
      rp, col, val             block CSR of 2N rows: aggregation + sub-graph sum;  val None = all ones
      rpT, colT, valT          its transpose (the backward of the aggregation)
      colL, valL               graph Laplacian D - A of the node rows on the row pointers rpT (calc_loss), or None
      gid [n] int32, off [B+1] int64   graph of every node row; node rows of graph g are off[g] .. off[g+1]
      aux    [2, B, 4]         per-layer aux features
    """

    def __init__(self, n, B, X, rp, col, val, rpT, colT, valT, gid, off, aux, colL=None, valL=None):
        self.n, self.B, self.N = n, B, n + B
        self.X, self.rp, self.col, self.val = X, rp, col, val
        self.rpT, self.colT, self.valT = rpT, colT, valT
        self.colL, self.valL = colL, valL
        self.gid, self.off, self.aux = gid, off, aux

    @staticmethod
    def from_alive_csr(n, B, gid, off, n_act, cnt, col, node_input, aux):
        """cnt[l] [n] alive degree of every compact row, col[l] the compact neighbour ids in row order (both int64)."""
        dev = gid.device
        N = n + B
        ar_n = torch.arange(n, device=dev)
        rp, cols, rpT, colsT, colsL, valsL, X = [], [], [], [], [], [], []
        base = baseT = 0
        for l in range(2):
            nnz = int(col[l].numel())
            r = torch.empty(N + 1, dtype=torch.long, device=dev)
            r[0] = 0
            r[1:n + 1] = torch.cumsum(cnt[l], 0)
            r[n + 1:] = nnz + torch.cumsum(n_act, 0)
            rp.append(r[(1 if l else 0):] + base)
            cols.append(torch.cat([col[l], ar_n]) + l * N)
            base += nnz + n
            # transpose / Laplacian: every node row has its neighbours + one extra slot (graph row / diagonal)
            endT = torch.cumsum(cnt[l] + 1, 0)
            rt = torch.empty(N + 1, dtype=torch.long, device=dev)
            rt[0] = 0
            rt[1:n + 1] = endT
            rt[n + 1:] = nnz + n
            rpT.append(rt[(1 if l else 0):] + baseT)
            row = torch.repeat_interleave(ar_n, cnt[l])
            ct = torch.empty(nnz + n, dtype=torch.long, device=dev)
            cl = torch.empty(nnz + n, dtype=torch.long, device=dev)
            vl = torch.full((nnz + n,), -1.0, dtype=torch.float32, device=dev)
            e = torch.arange(nnz, device=dev) + row
            ct[e] = col[l]
            cl[e] = col[l]
            ct[endT - 1] = n + gid.long()
            cl[endT - 1] = ar_n
            vl[endT - 1] = cnt[l].float()
            colsT.append(ct + l * N)
            colsL.append(cl + l * N)
            valsL.append(vl)
            baseT += nnz + n
            X.append(torch.cat([node_input[l].float(), torch.ones(B, 2, device=dev)]))
        i32 = lambda ts: torch.cat(ts).to(torch.int32).contiguous()
        return Operands(n, B, torch.cat(X).contiguous(), i32(rp), i32(cols), None, i32(rpT), i32(colsT), None,
                        gid.to(torch.int32).contiguous(), off.long().contiguous(),
                        aux.float().permute(1, 0, 2).contiguous(), i32(colsL), torch.cat(valsL).contiguous())

    @staticmethod
    def from_coo(n, B, n2n, subg, node_input, aux):
        """From the reference's own batch preparation (PrepareBatchGraph): per layer the n2nsum_param and subgsum_param
        COO operands ({"index", "value"}), general values.  Node rows of a graph must be consecutive (they are:
        PrepareBatchGraph numbers the batch graph by graph)."""
        dev = node_input[0].device
        N = n + B
        gid = torch.zeros(n, dtype=torch.long, device=dev)
        gid[subg[0]["index"][1].to(dev)] = subg[0]["index"][0].to(dev)
        if n > 1 and bool((gid[1:] < gid[:-1]).any()):
            raise ValueError("node rows must be numbered graph by graph (PrepareBatchGraph.py:308-340)")
        n_act = torch.bincount(gid, minlength=B)
        off = torch.zeros(B + 1, dtype=torch.long, device=dev)
        off[1:] = torch.cumsum(n_act, 0)
        rows, cols, vals, X = [], [], [], []
        for l in range(2):
            i1, v1 = n2n[l]["index"].to(dev).long(), n2n[l]["value"].to(dev).float()
            i2, v2 = subg[l]["index"].to(dev).long(), subg[l]["value"].to(dev).float()
            rows += [i1[0] + l * N, i2[0] + n + l * N]
            cols += [i1[1] + l * N, i2[1] + l * N]
            vals += [v1, v2]
            X.append(torch.cat([node_input[l].float(), torch.ones(B, 2, device=dev)]))
        index = torch.stack([torch.cat(rows), torch.cat(cols)])
        val = torch.cat(vals)
        rp, col, order = coo_to_csr(index, 2 * N)
        rpT, colT, orderT = coo_to_csr(torch.stack([index[1], index[0]]), 2 * N)
        return Operands(n, B, torch.cat(X).contiguous(), rp, col, val[order].contiguous(), rpT, colT, val[orderT].contiguous(),
                        gid.to(torch.int32).contiguous(), off, torch.as_tensor(aux, dtype=torch.float32, device=dev).permute(1, 0, 2).contiguous())


def _weights(params):
    """params: dict reference key -> contiguous f32 CUDA tensor (views into a flat buffer are fine)."""
    k = PARAM_KEYS
    w = {"w_n2l": params[k[0]], "c1": params[k[1]], "c2": params[k[2]], "c3": params[k[3]], "h1": params[k[4]], "h2": params[k[5]],
         "cross": params[k[6]], "wl1": params[k[7]], "wl2": params[k[8]], "trans": params[k[9]], "bias": params[k[10]],
         "lw": params[k[11]], "lb": params[k[12]]}
    for t in w.values():
        _chk(t)
    return w


# ---- forward ---------------------------------------------------------------------------------------------------------
def forward(params, o: Operands, bp_iters=3):
    """-> context dict; c["q"] [n] is Q of every node row, c["F"] [2N, 64] the fused, normalised embeddings of both
    layers (cur_message_layer of the reference = node rows of each layer)."""
    W = _weights(params)
    N, n, B, R = o.N, o.n, o.B, 2 * o.N
    Z = [gemm(o.X, W["w_n2l"])]                                           # :113-116 / :123-126
    H = [norm_fwd(Z[0], True)]
    P, M1, M2 = [], [], []
    c3a, c3b = W["c3"][:EMBED], W["c3"][EMBED:]
    for it in range(bp_iters):                                            # :136-166
        p = spmm_csr(o.rp, o.col, o.val, H[it], R)
        m1, m2 = gemm(p, W["c1"]), gemm(H[it], W["c2"])
        z = gemm(m1, c3a)
        gemm(m2, c3b, out=z, acc=True)
        P.append(p), M1.append(m1), M2.append(m2), Z.append(z), H.append(norm_fwd(z, True))
    E = H[-1]
    T = tanh_bias(gemm(E, W["trans"]), W["bias"])                         # mutil_layer_weight.py:270-276
    Xf = torch.empty(R, EMBED, dtype=torch.float32, device=E.device)
    saved = torch.empty(N, 8, dtype=torch.float32, device=E.device)
    _lib.check(_lib.lib().md_fuse2_fwd_f32(_p(T[:N]), _p(T[N:]), _p(W["lw"]), _p(W["lb"]), _p(Xf[:N]), _p(Xf[N:]), _p(saved), N, _st(E)),
               "md_fuse2_fwd_f32")
    F = norm_fwd(Xf, False)                                               # net_graphsage.py:182-185
    c = {"Z": Z, "H": H, "P": P, "M1": M1, "M2": M2, "T": T, "Xf": Xf, "saved": saved, "F": F, "heads": []}
    h2a, h2b = W["h2"][:W["h1"].shape[1]], W["h2"][W["h1"].shape[1]:]
    for l in range(2):                                                    # :187-241
        hf, yf = F[l * N:l * N + n], F[l * N + n:(l + 1) * N]
        s = gemm(yf, W["cross"])
        esa = rowscale(hf, s, o.gid)
        hp = gemm(esa, W["h1"])
        ql = gemm(hp, h2a, relu_a=True)
        aq = gemm(o.aux[l], h2b)
        g1 = gemm(yf, W["wl1"])
        gate = gemm(g1, W["wl2"], relu_a=True)
        c["heads"].append({"s": s, "esa": esa, "hp": hp, "q": ql, "aq": aq, "g1": g1, "gate": gate})
    q = torch.empty(n, dtype=torch.float32, device=E.device)
    h0, h1 = c["heads"]
    if n:
        _lib.check(_lib.lib().md_head_mix_fwd_f32(_p(h0["gate"]), _p(h1["gate"]), _p(h0["aq"]), _p(h1["aq"]), _p(h0["q"]), _p(h1["q"]),
                                                  _p(o.gid), _p(q), n, _st(E)), "md_head_mix_fwd_f32")
    c["q"] = q
    return c


# ---- backward --------------------------------------------------------------------------------------------------------
def backward(params, o: Operands, c, dq, dF, grads, zeroed=False):
    """Gradients of every parameter given dL/dq [n] (or None) and dL/dF [2N, 64] (or None), written into `grads`
    (dict reference key -> tensor of the parameter's shape; overwritten, not accumulated).  zeroed: the caller has
    already cleared them (views of one flat buffer: one memset instead of 13)."""
    W = _weights(params)
    G = _weights(grads)
    if not zeroed:
        for t in G.values():
            t.zero_()
    N, n, B, R = o.N, o.n, o.B, 2 * o.N
    dev = o.X.device
    F = c["F"]
    nh = W["h1"].shape[1]
    if dq is not None:
        h0, h1 = c["heads"]
        dsum_q = seg_reduce(dq, axpby(h0["q"].view(-1), h1["q"].view(-1), 1.0, -1.0), o.off)
        dqs = seg_reduce(dq, None, o.off)
        dql = [torch.empty(n, 1, dtype=torch.float32, device=dev) for _ in range(2)]
        dg = [torch.empty(B, 1, dtype=torch.float32, device=dev) for _ in range(2)]
        da = [torch.empty(B, 1, dtype=torch.float32, device=dev) for _ in range(2)]
        _lib.check(_lib.lib().md_head_mix_bwd_f32(_p(h0["gate"]), _p(h1["gate"]), _p(h0["aq"]), _p(h1["aq"]), _p(_chk(dq)), _p(dsum_q), _p(dqs),
                                                  _p(o.gid), _p(dql[0]), _p(dql[1]), _p(dg[0]), _p(dg[1]), _p(da[0]), _p(da[1]), n, B,
                                                  _st(dq)), "md_head_mix_bwd_f32")
        dFh = torch.empty(R, EMBED, dtype=torch.float32, device=dev)
        for l in range(2):
            hd = c["heads"][l]
            hf, yf = F[l * N:l * N + n], F[l * N + n:(l + 1) * N]
            dhf, dyf = dFh[l * N:l * N + n], dFh[l * N + n:(l + 1) * N]
            # layer gate: gate = relu(yf @ wl1) @ wl2
            gemm(hd["g1"], dg[l], ta=True, relu_a=True, out=G["wl2"], acc=True)
            dg1 = relu_mask(hd["g1"], gemm(dg[l], W["wl2"], tb=True))
            gemm(yf, dg1, ta=True, out=G["wl1"], acc=True)
            gemm(dg1, W["wl1"], tb=True, out=dyf)
            # aux part of the last layer: aq = aux @ h2[nh:]
            gemm(o.aux[l], da[l], ta=True, out=G["h2"][nh:], acc=True)
            # q = relu(esa @ h1) @ h2[:nh], esa = hf * (yf . cross)[gid]
            gemm(hd["hp"], dql[l], ta=True, relu_a=True, out=G["h2"][:nh], acc=True)
            dhp = relu_mask(hd["hp"], gemm(dql[l], W["h2"][:nh], tb=True))
            gemm(hd["esa"], dhp, ta=True, out=G["h1"], acc=True)
            desa = gemm(dhp, W["h1"], tb=True)
            rowscale(desa, hd["s"], o.gid, out=dhf)
            ds = seg_reduce(rowdot(desa, hf), None, o.off).view(B, 1)
            gemm(yf, ds, ta=True, out=G["cross"], acc=True)
            gemm(ds, W["cross"], tb=True, out=dyf, acc=True)
        if dF is not None:
            axpby(dFh, _chk(dF), 1.0, 1.0, out=dFh)
        dF = dFh
    assert dF is not None
    # fusion
    dXf = norm_bwd(c["Xf"], dF, False)
    T = c["T"]
    dT = torch.empty(R, EMBED, dtype=torch.float32, device=dev)
    glw = torch.empty(N, EMBED, dtype=torch.float32, device=dev)
    glb = torch.empty(N, dtype=torch.float32, device=dev)
    _lib.check(_lib.lib().md_fuse2_bwd_f32(_p(T[:N]), _p(T[N:]), _p(W["lw"]), _p(c["saved"]), _p(dXf[:N]), _p(dXf[N:]), _p(dT[:N]), _p(dT[N:]),
                                           _p(glw), _p(glb), N, _st(T)), "md_fuse2_bwd_f32")
    colsum(glw, G["lw"])
    colsum(glb, G["lb"])
    dU = tanh_bwd(T, dT)
    H, Z = c["H"], c["Z"]
    gemm(H[-1], dU, ta=True, out=G["trans"])
    colsum(dU, G["bias"])
    dH = gemm(dU, W["trans"], tb=True)
    # embedding iterations, last to first
    c3a, c3b = W["c3"][:EMBED], W["c3"][EMBED:]
    for it in range(len(c["P"]) - 1, -1, -1):
        dZ = norm_bwd(Z[it + 1], dH, True)
        gemm(c["M1"][it], dZ, ta=True, out=G["c3"][:EMBED], acc=True)
        gemm(c["M2"][it], dZ, ta=True, out=G["c3"][EMBED:], acc=True)
        dM1, dM2 = gemm(dZ, c3a, tb=True), gemm(dZ, c3b, tb=True)
        gemm(c["P"][it], dM1, ta=True, out=G["c1"], acc=True)
        gemm(H[it], dM2, ta=True, out=G["c2"], acc=True)
        dH = spmm_csr(o.rpT, o.colT, o.valT, gemm(dM1, W["c1"], tb=True), R)
        gemm(dM2, W["c2"], tb=True, out=dH, acc=True)
    dZ0 = norm_bwd(Z[0], dH, True)
    gemm(o.X, dZ0, ta=True, out=G["w_n2l"])
    return grads


# ---- loss ------------------------------------------------------------------------------------------------------------
def loss_forward(o: Operands, q_pred, F, target, batch_global, edge_num, alpha):
    """MSE(target, q_pred) (mean over the GLOBAL batch) + alpha * sum_l 2 trace(H_l^T (D - A)_l H_l) / edge_num[l]."""
    N, R = o.N, 2 * o.N
    diff = axpby(_chk(q_pred).reshape(-1), _chk(target).reshape(-1), 1.0, -1.0)
    loss = axpby(dot(diff, diff), None, 1.0 / float(batch_global), 0.0)
    LH = spmm_csr(o.rpT, o.colL, o.valL, F, R)
    for l in range(2):
        quad = dot(F[l * N:(l + 1) * N], LH[l * N:(l + 1) * N])
        loss = axpby(loss, quad, 1.0, 2.0 * alpha / max(float(edge_num[l]), 1.0))
    return loss.reshape(()), {"diff": diff, "LH": LH, "bg": float(batch_global), "en": [max(float(e), 1.0) for e in edge_num], "alpha": alpha}


def loss_backward(o: Operands, lc, g=1.0):
    N = o.N
    dq_pred = axpby(lc["diff"], None, 2.0 * g / lc["bg"], 0.0)
    dF = torch.empty_like(lc["LH"])
    for l in range(2):  # d (h^T L h) / dh = 2 L h (L symmetric)
        axpby(lc["LH"][l * N:(l + 1) * N], None, 4.0 * g * lc["alpha"] / lc["en"][l], 0.0, out=dF[l * N:(l + 1) * N])
    return dq_pred, dF


# ---- autograd adapters (callers that want loss.backward(); fit() calls forward / backward directly) --------------------
class _QNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, o, *params):
        pd = {k: p.detach() for k, p in zip(PARAM_KEYS, params)}
        c = forward(pd, o)
        ctx.o, ctx.c, ctx.pd = o, c, pd
        return c["q"].view(-1, 1), c["F"].view(2, o.N, EMBED)

    @staticmethod
    def backward(ctx, dq, dF):
        o = ctx.o
        grads = {k: torch.empty_like(v) for k, v in ctx.pd.items()}
        backward(ctx.pd, o, ctx.c, dq.contiguous().view(-1) if dq is not None else None,
                 dF.contiguous().view(2 * o.N, EMBED) if dF is not None else None, grads)
        return (None,) + tuple(grads[k] for k in PARAM_KEYS)


class _TakeRowsFn(torch.autograd.Function):
    """q[rows] with distinct rows: the backward is a scatter into zeros (data movement only)."""

    @staticmethod
    def forward(ctx, q, rows):
        ctx.n, ctx.rows = q.shape[0], rows
        return q.index_select(0, rows)

    @staticmethod
    def backward(ctx, g):
        out = torch.zeros((ctx.n,) + tuple(g.shape[1:]), dtype=g.dtype, device=g.device)
        out.index_copy_(0, ctx.rows, g)
        return out, None


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, o, q_pred, F, target, batch_global, edge_num, alpha):
        loss, lc = loss_forward(o, q_pred.detach().contiguous(), F.detach().contiguous().view(2 * o.N, EMBED), target, batch_global, edge_num, alpha)
        ctx.o, ctx.lc, ctx.qshape = o, lc, q_pred.shape
        return loss

    @staticmethod
    def backward(ctx, g):
        dq, dF = loss_backward(ctx.o, ctx.lc, float(g))
        return None, dq.view(ctx.qshape), dF.view(2, ctx.o.N, EMBED), None, None, None, None


def network(params_in_key_order, o):
    """(q [n, 1], F [2, N, 64]) with autograd through the hand-written backward."""
    return _QNetFn.apply(o, *params_in_key_order)


def take_rows(q, rows):
    return _TakeRowsFn.apply(q, rows)


def loss(o, q_pred, F, target, batch_global, edge_num, alpha):
    return _LossFn.apply(o, q_pred, F, target, batch_global, edge_num, alpha)


# ---- Adam on flat buffers --------------------------------------------------------------------------------------------
class FlatAdam:
    """torch.optim.Adam(params, lr) of the reference (betas 0.9 / 0.999, eps 1e-8) as ONE kernel over the flat parameter
    buffer.  `params` (dict order of `shapes`) are views into `flat` at `slices[key] = (offset, numel)`; their gradients
    are collected in `flat_grad` (also the all-reduce buffer)."""

    def __init__(self, flat, flat_grad, params, slices, shapes, lr, betas=(0.9, 0.999), eps=1e-8):
        self.flat, self.flat_grad, self.params = flat, flat_grad, list(params)
        self.lr, self.betas, self.eps, self.t = lr, betas, eps, 0
        self.m, self.v = torch.zeros_like(flat), torch.zeros_like(flat)
        self._views = [flat_grad[slices[k][0]:slices[k][0] + slices[k][1]].view(shapes[k]) for k in shapes]

    def zero_grad(self):
        self.flat_grad.zero_()
        for p in self.params:
            p.grad = None

    def adopt_flat_grads(self):
        """the gradients were written straight into flat_grad (dense.backward): expose them as p.grad."""
        for p, v in zip(self.params, self._views):
            p.grad = v

    def gather_grads(self):
        """gradients that autograd delivered as its own tensors are copied into the flat buffer."""
        for p, v in zip(self.params, self._views):
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad.reshape(v.shape))
                p.grad = v
        return self.flat_grad

    def step(self):
        self.gather_grads()
        self.t += 1
        _lib.check(_lib.lib().md_adam_step_f32(_p(self.flat), _p(self.flat_grad), _p(self.m), _p(self.v), self.flat.numel(), self.lr,
                                               self.betas[0], self.betas[1], self.eps, self.t, _st(self.flat)), "md_adam_step_f32")
