"""Model-level drop-in: `MultiDismantler_net` with the reference's module layout, state_dict keys and forward signatures.

Mirrors /root/reference/code/MultiDismantler_unit_cost/MultiDismantler_net_graphsage.py (class MultiDismantler_net:
__init__ :22-90, train_forward :92-241, test_forward :243-394), the degree-cost variant D/MultiDismantler_net.py and the
fusion module U/MRGNN/mutil_layer_weight.py:252-313 (BitwiseMultipyLogis / LogisticVector).

For callers that keep the reference's own batch preparation (PrepareBatchGraph -> COO dicts) and only swap the network:
the inputs are exactly what MultiDismantler.SetuppredAll / SetupTrain build (lists of two {"index","value","m","n"}
dicts, aux_input [B,2,4], node_input None or [2,n,2], dense adj / v_adj accepted and ignored: the degree features they
are used for are read off the sparse operands).  Neighbour aggregation runs in the CUDA SpMM kernel (`ops.spmm`,
forward and backward); the 31 k-parameter dense algebra is torch on the same device.  `load_state_dict` of the
reference's 14-key checkpoints works unchanged (`last_w` is the same Parameter as `h2_weight`, :69).
The device-resident rollout (`MultiDismantler.GetSol`, `BatchRollout`) does not go through this class: it runs the
hand-written forward kernels; this is the operator/model boundary of SURVEY 8(b).
"""
import torch
import torch.nn.functional as F

from . import ops

EMBEDDING_SIZE = 64   # U/MultiDismantler_torch.py:35
REG_HIDDEN = 32       # :53
AUX_DIM = 4           # :58
MAX_BP_ITER = 3       # :62


class LogisticVector(torch.nn.Module):  # mutil_layer_weight.py:304-313
    def __init__(self, n_feature, n_hidden):
        super().__init__()
        self.parameter = torch.nn.Linear(n_feature, n_hidden)

    def forward(self, x):
        return torch.sigmoid(self.parameter(x)).squeeze()


class BitwiseMultipyLogis(torch.nn.Module):  # mutil_layer_weight.py:252-301
    def __init__(self, features_num=EMBEDDING_SIZE, metapath_number=2):
        super().__init__()
        self.metapath_number = metapath_number
        self.logis = LogisticVector(features_num, 1)
        self.trans = torch.nn.Parameter(torch.eye(features_num))
        self.bias = torch.nn.Parameter(torch.zeros(features_num))

    def forward(self, node_features, nodes_ori, layer_predict):
        t = [torch.tanh(torch.matmul(node_features[k], self.trans) + self.bias) for k in range(self.metapath_number)]
        sem = torch.stack([t[k] * t[layer_predict] for k in range(self.metapath_number)], dim=1)  # layer_bitwise :294-301
        weight = F.softmax(self.logis(sem).reshape(-1, self.metapath_number), dim=1)
        x = t[layer_predict]
        for k in range(self.metapath_number):
            if k != layer_predict:
                x = x + weight[:, k].unsqueeze(1) * t[k]
        return x


class MultiDismantler_net(torch.nn.Module):
    def __init__(self, layerNodeAttention_weight=None, embedding_size=EMBEDDING_SIZE, w_initialization_std=1, reg_hidden=REG_HIDDEN,
                 max_bp_iter=MAX_BP_ITER, embeddingMethod=1, aux_dim=AUX_DIM, device=None, node_attr=False):
        super().__init__()
        if device is not None and torch.device(device).type != "cuda":
            raise RuntimeError("mdcommunity_b200.net needs a CUDA device (no CPU fallback)")
        self.embedding_size, self.reg_hidden, self.max_bp_iter, self.aux_dim = embedding_size, reg_hidden, max_bp_iter, aux_dim
        self.node_attr = node_attr
        rnd = lambda *shape: torch.nn.Parameter(torch.fmod(torch.randn(*shape), 2) * w_initialization_std)  # :30
        self.w_n2l = rnd(2, embedding_size)
        self.p_node_conv = rnd(embedding_size, embedding_size)
        self.p_node_conv2 = rnd(embedding_size, embedding_size)
        self.p_node_conv3 = rnd(2 * embedding_size, embedding_size)
        self.h1_weight = rnd(embedding_size, reg_hidden)
        self.h2_weight = rnd(reg_hidden + aux_dim, 1)
        self.last_w = self.h2_weight  # :69 -- one Parameter under two state_dict keys
        self.cross_product = rnd(embedding_size, 1)
        self.w_layer1 = rnd(embedding_size, 128)
        self.w_layer2 = rnd(128, 1)
        self.layerNodeAttention_weight = layerNodeAttention_weight if layerNodeAttention_weight is not None else BitwiseMultipyLogis(embedding_size)
        if device is not None:
            self.to(device)

    # ---- shared trunk ---------------------------------------------------------------------------------------------
    def _embed(self, node_input, subgsum_param, n2nsum_param):
        dev = self.w_n2l.device
        nodes_cnt, B = n2nsum_param[0]["m"], subgsum_param[0]["m"]
        gid = torch.zeros(nodes_cnt, dtype=torch.long, device=dev)
        gid[subgsum_param[0]["index"][1].to(dev)] = subgsum_param[0]["index"][0].to(dev)  # node -> graph (subg_construct :308-340)
        emb = []
        for l in range(2):
            index = n2nsum_param[l]["index"].to(dev)
            value = n2nsum_param[l]["value"].to(dev).float()
            if node_input is None:  # unit cost: degree / max degree of the graph (:100-112), read off the n2n operand
                deg = torch.zeros(nodes_cnt, device=dev).index_add_(0, index[0], torch.ones_like(value))
                dmax = torch.zeros(B, device=dev).scatter_reduce_(0, gid, deg, reduce="amax", include_self=True)
                d = deg / dmax.clamp_min(1.0)[gid]
                x = torch.stack([d, d], 1)
            else:
                x = torch.as_tensor(node_input[l], dtype=torch.float32, device=dev)
            h = F.normalize(torch.relu(torch.matmul(x, self.w_n2l)), p=2, dim=1)
            y = F.normalize(torch.relu(torch.matmul(torch.ones((B, 2), device=dev), self.w_n2l)), p=2, dim=1)
            sindex, svalue = subgsum_param[l]["index"].to(dev), subgsum_param[l]["value"].to(dev).float()
            for _ in range(self.max_bp_iter):
                n2npool = ops.spmm(index, value, nodes_cnt, nodes_cnt, h)           # torch_sparse.spmm call site :139
                y_n2npool = ops.spmm(sindex, svalue, B, nodes_cnt, h)               # :146
                merged = torch.cat([torch.matmul(n2npool, self.p_node_conv), torch.matmul(h, self.p_node_conv2)], 1)
                y_merged = torch.cat([torch.matmul(y_n2npool, self.p_node_conv), torch.matmul(y, self.p_node_conv2)], 1)
                h = F.normalize(torch.relu(torch.matmul(merged, self.p_node_conv3)), p=2, dim=1)
                y = F.normalize(torch.relu(torch.matmul(y_merged, self.p_node_conv3)), p=2, dim=1)
            emb.append(torch.cat([h, y], 0))
        fused = [F.normalize(self.layerNodeAttention_weight(emb, None, l), p=2, dim=1) for l in range(2)]
        return fused, nodes_cnt, B, gid

    def _heads(self, fused, nodes_cnt, gid, rows, aux_input):
        """Q of the node rows `rows` (None = all nodes), graph of each row = gid[rows]."""
        dev = self.w_n2l.device
        aux = torch.as_tensor(aux_input, dtype=torch.float32, device=dev)
        qs, gates = [], []
        for l in range(2):
            hf, yf = fused[l][:nodes_cnt], fused[l][nodes_cnt:]
            g_of = gid if rows is None else gid[rows]
            act = hf if rows is None else hf.index_select(0, rows)
            rep_y = yf.index_select(0, g_of)
            embed_s_a = act * torch.matmul(rep_y, self.cross_product)                # (act y^T) cross == act * (y . cross)
            hidden = torch.relu(torch.matmul(embed_s_a, self.h1_weight))
            qs.append(torch.matmul(torch.cat([hidden, aux[:, l, :].index_select(0, g_of)], 1), self.last_w))
            gates.append(torch.matmul(torch.relu(torch.matmul(rep_y, self.w_layer1)), self.w_layer2))
        sm = F.softmax(torch.cat(gates, 1), dim=1)
        return sm[:, 0:1] * qs[0] + sm[:, 1:2] * qs[1]

    # ---- reference entry points -----------------------------------------------------------------------------------
    def train_forward(self, node_input, subgsum_param, n2nsum_param, action_select, aux_input, adj=None, v_adj=None):
        """-> (q_pred [B,1], cur_message_layer [2,n,64])  (:92-241)."""
        fused, n, B, gid = self._embed(node_input, subgsum_param, n2nsum_param)
        dev = self.w_n2l.device
        a_index = action_select[0]["index"].to(dev)  # row = graph, col = node of the chosen action (:199-203)
        rows = torch.zeros(B, dtype=torch.long, device=dev)
        rows[a_index[0]] = a_index[1]
        q = self._heads(fused, n, gid, rows, aux_input)
        return q, torch.stack([fused[0][:n], fused[1][:n]])

    def test_forward(self, node_input, subgsum_param, n2nsum_param, rep_global, aux_input, adj=None, v_adj=None):
        """-> q_on_all [n,1]  (:243-394)."""
        fused, n, B, gid = self._embed(node_input, subgsum_param, n2nsum_param)
        return self._heads(fused, n, gid, None, aux_input)
