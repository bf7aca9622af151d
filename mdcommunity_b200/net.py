"""Model-level drop-in: `MultiDismantler_net` with the reference's module layout, state_dict keys and forward signatures.

Mirrors /root/reference/code/MultiDismantler_unit_cost/MultiDismantler_net_graphsage.py (class MultiDismantler_net:
__init__ :22-90, train_forward :92-241, test_forward :243-394), the degree-cost variant D/MultiDismantler_net.py and the
fusion module U/MRGNN/mutil_layer_weight.py:252-313 (BitwiseMultipyLogis / LogisticVector).

For callers that keep the reference's own batch preparation (PrepareBatchGraph -> COO dicts) and only swap the network:
the inputs are exactly what MultiDismantler.SetuppredAll / SetupTrain build (lists of two {"index","value","m","n"}
dicts, aux_input [B,2,4], node_input None or [2,n,2], dense adj / v_adj accepted and ignored: the degree features they
are used for are read off the sparse operands).  Neighbour aggregation AND the dense algebra of the 31 k-parameter
network run in the hand-written kernels of `dense.py` (csrc/md_train.cu, csrc/md_spmm.cu), forward and backward; torch
only builds the CSR operands and carries the autograd edge.  `load_state_dict` of the
reference's 14-key checkpoints works unchanged (`last_w` is the same Parameter as `h2_weight`, :69).
The device-resident rollout (`MultiDismantler.GetSol`, `BatchRollout`) does not go through this class: it runs the
hand-written forward kernels; this is the operator/model boundary of SURVEY 8(b).
"""
import torch
import torch.nn.functional as F

from . import dense

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
    def _params(self):
        law = self.layerNodeAttention_weight
        return [self.w_n2l, self.p_node_conv, self.p_node_conv2, self.p_node_conv3, self.h1_weight, self.h2_weight, self.cross_product,
                self.w_layer1, self.w_layer2, law.trans, law.bias, law.logis.parameter.weight, law.logis.parameter.bias]

    def _network(self, node_input, subgsum_param, n2nsum_param, aux_input):
        """(q of every node [n, 1], fused embeddings [2, n + B, 64], n) through the hand-written kernels (dense.py:
        forward AND backward; torch only builds the CSR operands from the COO dicts)."""
        dev = self.w_n2l.device
        n, B = n2nsum_param[0]["m"], subgsum_param[0]["m"]
        if self.max_bp_iter != MAX_BP_ITER or self.embedding_size != EMBEDDING_SIZE:
            raise ValueError("the kernels are built for embedding_size 64 and max_bp_iter 3 (U/MultiDismantler_torch.py:35,62)")
        x = []
        if node_input is None:  # unit cost: degree / max degree of the graph (:100-112), read off the n2n operand
            gid = torch.zeros(n, dtype=torch.long, device=dev)
            gid[subgsum_param[0]["index"][1].to(dev)] = subgsum_param[0]["index"][0].to(dev)  # node -> graph (subg_construct :308-340)
            for l in range(2):
                index = n2nsum_param[l]["index"].to(dev)
                deg = torch.zeros(n, device=dev).index_add_(0, index[0], torch.ones(index.shape[1], device=dev))
                dmax = torch.zeros(B, device=dev).scatter_reduce_(0, gid, deg, reduce="amax", include_self=True)
                d = deg / dmax.clamp_min(1.0)[gid]
                x.append(torch.stack([d, d], 1))
        else:
            x = [torch.as_tensor(node_input[l], dtype=torch.float32, device=dev) for l in range(2)]
        o = dense.Operands.from_coo(n, B, n2nsum_param, subgsum_param, x, aux_input)
        q, fused = dense.network(self._params(), o)
        return q, fused, n

    # ---- reference entry points -----------------------------------------------------------------------------------
    def train_forward(self, node_input, subgsum_param, n2nsum_param, action_select, aux_input, adj=None, v_adj=None):
        """-> (q_pred [B,1], cur_message_layer [2,n,64])  (:92-241)."""
        q, fused, n = self._network(node_input, subgsum_param, n2nsum_param, aux_input)
        dev = self.w_n2l.device
        a_index = action_select[0]["index"].to(dev)  # row = graph, col = node of the chosen action (:199-203)
        rows = torch.zeros(subgsum_param[0]["m"], dtype=torch.long, device=dev)
        rows[a_index[0]] = a_index[1]
        return dense.take_rows(q, rows), fused[:, :n]

    def test_forward(self, node_input, subgsum_param, n2nsum_param, rep_global, aux_input, adj=None, v_adj=None):
        """-> q_on_all [n,1]  (:243-394)."""
        return self._network(node_input, subgsum_param, n2nsum_param, aux_input)[0]
