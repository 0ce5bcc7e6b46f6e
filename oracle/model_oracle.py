"""ORACLE (test infrastructure, not product code): CPU PyTorch restatement of the reference's
device hot path, /root/reference/graphormer/{wrapper,collator,model_fqandtoyo,modelGNN,lr}.py.

The reference needs pytorch_lightning / torch_geometric / ogb (not installed) and its dataset files, so every function below
restates the cited lines with plain torch on CPU.  PARITY PIN: the integer preprocessing (preprocess_item) is pinned to the
compiled reference through oracle/algos_oracle.c + tests/golden/algos_golden.npz; the collator fields, the forward pass, the
losses and the metrics are pinned to the UNMODIFIED reference modules (wrapper.py, collator.py, model_fqandtoyo.py, modelGNN.py),
which tests/golden/make_model_golden.py executes on the CPU in the build container behind import stubs for the three missing
packages, for the toyotagraph / gowalla_nevda / foursquaregraph branches: tests/golden/model_golden_*.npz, metrics_golden.npz,
checked by tests/test_oracle_model_golden.py (fields bit for bit, logits / loss within 1e-5).  Not pinned: `poi_pos` binning
(the reference bins a distance pickle that is not shipped; PoiWorld.poi_pos_bins is a stand-in), the dropout masks (stochastic)
and the reference's fp16 AMP rounding (the oracle is fp32; tolerances are the north_star's).  Two bias modes: 'fp32' (= model.py:157-190 arithmetic + the live model's poi_pos
term; the 1e-5 target) and 'ref_half' (the live model's .half() round trips,
model_fqandtoyo.py:1178-1198; compared at the bf16 tolerance).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import copy
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

import algos_oracle

NODE_DIM = 2000          # model_fqandtoyo.py:567


# ------------------------------------------------------------------------------------ wrapper.py
def convert_to_single_emb(x, offset=512):
    """wrapper.py:18-22"""
    feature_num = x.size(1) if len(x.size()) > 1 else 1
    feature_offset = 1 + torch.arange(0, feature_num * offset, offset, dtype=torch.long)
    return x + feature_offset


class Item:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def preprocess_item(raw, hop_cap=None):
    """wrapper.py:25-102 on a raw dataset item (numpy fields as owndata.py:340-349 stores them).
    hop_cap: produce only the first hop_cap hop slots of edge_input (== the collator's slice,
    collator.py:323) instead of the reference's max_dist(=510)-deep temporary."""
    x = torch.as_tensor(raw.x, dtype=torch.long)
    edge_index = torch.as_tensor(raw.edge_index, dtype=torch.long)
    edge_attr = torch.as_tensor(raw.edge_attr, dtype=torch.long)
    N = x.size(0)
    x = convert_to_single_emb(x)                                   # :37
    user = convert_to_single_emb(torch.as_tensor(raw.user, dtype=torch.long))   # :39
    adj_orig = torch.zeros([N, N], dtype=torch.bool)               # :42-43
    adj_orig[edge_index[0, :], edge_index[1, :]] = True
    if len(edge_attr.size()) == 1:
        edge_attr = edge_attr[:, None]
    attn_edge_type = torch.zeros([N, N, edge_attr.size(-1)], dtype=torch.long)   # :49-53
    attn_edge_type[edge_index[0, :], edge_index[1, :]] = convert_to_single_emb(edge_attr) + 1
    M, path = algos_oracle.floyd_warshall(adj_orig.numpy())        # :55-57
    max_dist = int(np.amax(M)) if N > 0 else 0                     # :58
    edge_input = algos_oracle.gen_edge_input(max_dist, path, attn_edge_type.numpy(), hop_cap=hop_cap)  # :60
    rel_pos = torch.from_numpy(M).long()                           # :61
    attn_bias = torch.zeros([N + 1, N + 1], dtype=torch.float)     # :63-65
    adj = torch.zeros([N + 1, N + 1], dtype=torch.bool)
    adj[edge_index[0, :], edge_index[1, :]] = True
    adj1 = torch.zeros([N, N], dtype=torch.bool)
    adj1[edge_index[0, :], edge_index[1, :]] = True
    adj[N, :] = True
    adj[:, N] = True
    return Item(
        idx=raw.idx, x=x, user=user, adj=adj, adj1=adj1, attn_bias=attn_bias, attn_edge_type=attn_edge_type,
        rel_pos=rel_pos, in_degree=adj_orig.long().sum(dim=1).view(-1), out_degree=adj_orig.long().sum(dim=0).view(-1),
        edge_input=torch.from_numpy(edge_input).long(), edge_index=edge_index, y=torch.as_tensor(raw.y, dtype=torch.long),
        time=torch.as_tensor(raw.time, dtype=torch.long), time_normal=torch.as_tensor(raw.time_normal, dtype=torch.float),
        cat=torch.as_tensor(raw.cat, dtype=torch.long), max_dist=max_dist)


# ------------------------------------------------------------------------------------ collator.py
def _pad_1d(x, padlen):                      # pad_1d_unsqueeze :11-18
    x = x + 1
    out = x.new_zeros([padlen])
    out[:x.size(0)] = x
    return out.unsqueeze(0)


def _pad_2d_squeeze(x, padlen):              # :29-37
    x = x - 1
    out = x.new_zeros([padlen, x.size(1)])
    out[:x.size(0)] = x
    return out.unsqueeze(0)


def _pad_rows(x, padlen):                    # pad_time_unsqueeze :39-45
    out = x.new_zeros([padlen, x.size(1)])
    out[:x.size(0)] = x
    return out.unsqueeze(0)


def _pad_attn_bias(x, padlen):               # :57-64
    xlen = x.size(0)
    if xlen < padlen:
        new_x = x.new_zeros([padlen, padlen]).fill_(float("-inf"))
        new_x[:xlen, :xlen] = x
        new_x[xlen:, :xlen] = 0
        x = new_x
    return x.unsqueeze(0)


def _pad_sq(x, padlen, shift):               # pad_rel_pos_unsqueeze :76-83 / pad_2d_bool :48-54
    x = x + shift if shift else x
    out = x.new_zeros([padlen, padlen] + list(x.shape[2:]))
    out[:x.size(0), :x.size(1)] = x
    return out.unsqueeze(0)


def _pad_3d(x, p1, p2, p3):                  # pad_3d_unsqueeze :86-93
    x = x + 1
    out = x.new_zeros([p1, p2, p3, x.size(3)])
    out[:x.size(0), :x.size(1), :x.size(2)] = x
    return out.unsqueeze(0)


class Batch1(Item):
    """collator.py:149-215"""

    def __len__(self):
        return self.in_degree.size(0)


def collate(items, world, max_node=30000, multi_hop_max_dist=20, rel_pos_max=1024):
    """collator_foursquare / _gowalla / _toyota (collator.py:310-458, 460-608, 610-748): identical
    but for the distance pickle.  `world.poi_pos_bins` stands in for pickle + np.digitize (:428-437);
    the unused Laplacian eigen-decomposition (:394-410) is not restated (feature_matrix=None)."""
    items = [it for it in items if it is not None and it.x.size(0) <= max_node]
    for it in items:
        it.attn_bias[1:, 1:][it.rel_pos >= rel_pos_max] = float("-inf")       # :354-358
    N = max(it.x.size(0) for it in items)                                      # :361
    eis = [it.edge_input[:, :, :multi_hop_max_dist, :] for it in items]       # :323
    max_dist = max(e.size(-2) for e in eis)                                    # :366
    x = torch.cat([_pad_2d_squeeze(it.x, N) for it in items])
    B = len(items)
    indx = (x != 0).sum(dim=-2).view(-1)
    poi_pos = torch.cat([_pad_sq(it.rel_pos, N, 1) for it in items])          # :428 (overwritten below)
    for i in range(B):
        n = int(indx[i])
        ids = x[i, :n, 0].numpy()
        poi_pos[i, :n, :n] = torch.from_numpy(world.poi_pos_bins(ids, ids))   # :435-437
    return Batch1(
        idx=torch.LongTensor([it.idx for it in items]),
        attn_bias=torch.cat([_pad_attn_bias(it.attn_bias, N + 1) for it in items]),
        attn_edge_type=torch.cat([_pad_sq(it.attn_edge_type, N + 1, 0) for it in items]),
        rel_pos=torch.cat([_pad_sq(it.rel_pos, N, 1) for it in items]),
        in_degree=torch.cat([_pad_1d(it.in_degree, N) for it in items]),
        out_degree=torch.cat([_pad_1d(it.out_degree, N) for it in items]),
        x=x,
        edge_input=torch.cat([_pad_3d(e, N, N, max_dist) for e in eis]),
        y=torch.cat([it.y for it in items]),
        adj=torch.cat([_pad_sq(it.adj, N + 1, 0) for it in items]),
        adj1=torch.cat([_pad_sq(it.adj1, N, 0) for it in items]),
        time=torch.cat([_pad_rows(it.time, N) for it in items]),
        feature_matrix=None,
        time_normal=torch.cat([_pad_rows(it.time_normal, N) for it in items]),
        user=torch.cat([it.user for it in items]),
        cat=torch.cat([_pad_rows(it.cat, N) for it in items]),
        poi_pos=poi_pos)


# ------------------------------------------------------------------------------------ modelGNN.py
class GraphConvolution(nn.Module):
    """modelGNN.py:21-50"""

    def __init__(self, in_features, out_features):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(in_features, out_features))
        self.bias = nn.Parameter(torch.empty(out_features))
        stdv = 1.0 / math.sqrt(out_features)
        self.weight.data.uniform_(-stdv, stdv)
        self.bias.data.uniform_(-stdv, stdv)

    def forward(self, x, adj):
        s = torch.mm(x, self.weight)
        # modelGNN.py:40 torch.spmm(adj, support): dense mm in the reference; a sparse COO adj gives the same sums without
        # the P x P dense matrix (14 GB at P = 60 000) and is used for the large synthetic worlds
        return (torch.sparse.mm(adj, s) if adj.is_sparse else torch.mm(adj, s)) + self.bias


class GCN(nn.Module):
    """modelGNN.py:53-73"""

    def __init__(self, ninput, nhid, noutput, dropout):
        super().__init__()
        ch = [ninput] + nhid + [noutput]
        self.gcn = nn.ModuleList([GraphConvolution(ch[i], ch[i + 1]) for i in range(len(ch) - 1)])
        self.dropout = dropout

    def forward(self, x, adj):
        for i in range(len(self.gcn) - 1):
            x = F.leaky_relu(self.gcn[i](x, adj), 0.2)
        x = F.dropout(x, self.dropout, training=self.training)
        return self.gcn[-1](x, adj)


# ------------------------------------------------------------------------------------ model_fqandtoyo.py
class FuseEmbeddings(nn.Module):
    """:440-455"""

    def __init__(self, d1, d2):
        super().__init__()
        self.fuse_embed = nn.Linear(d1 + d2, d1 + d2)

    def forward(self, a, b):
        return F.leaky_relu(self.fuse_embed(torch.cat((a, b), a.dim() - 1)), 0.2)


class UserEmbeddings(nn.Module):
    def __init__(self, n, d):
        super().__init__()
        self.user_embedding = nn.Embedding(n, d)

    def forward(self, i):
        return self.user_embedding(i)


class LearnablePositionalEncoding(nn.Module):
    """:330-358 (its Dropout(0.1), :358, is applied by Graphormer.forward below when `pos_dropout` > 0)"""

    def __init__(self, d_model, max_len):
        super().__init__()
        self.pe = nn.Parameter(torch.empty(d_model, max_len))
        nn.init.uniform_(self.pe, -0.02, 0.02)


class FeedForwardNetwork(nn.Module):
    """:1644-1656"""

    def __init__(self, h, f):
        super().__init__()
        self.layer1 = nn.Linear(h, f)
        self.layer2 = nn.Linear(f, h)

    def forward(self, x):
        return self.layer2(F.gelu(self.layer1(x)))


class MultiHeadAttention(nn.Module):
    """:1659-1711"""

    def __init__(self, hidden, heads, attention_dropout_rate=0.0):
        super().__init__()
        self.att_dropout = nn.Dropout(attention_dropout_rate)                                 # :1674
        self.num_heads = heads
        self.att_size = hidden // heads
        self.scale = self.att_size ** -0.5
        self.linear_q = nn.Linear(hidden, heads * self.att_size)
        self.linear_k = nn.Linear(hidden, heads * self.att_size)
        self.linear_v = nn.Linear(hidden, heads * self.att_size)
        self.output_layer = nn.Linear(heads * self.att_size, hidden)

    def forward(self, q, k, v, attn_bias):
        b, d = q.size(0), self.att_size
        q = self.linear_q(q).view(b, -1, self.num_heads, d).transpose(1, 2)
        k = self.linear_k(k).view(b, -1, self.num_heads, d).transpose(1, 2).transpose(2, 3)
        v = self.linear_v(v).view(b, -1, self.num_heads, d).transpose(1, 2)
        x = torch.matmul(q * self.scale, k) + attn_bias
        x = self.att_dropout(torch.softmax(x, dim=3)).matmul(v)                               # :1703-1705
        x = x.transpose(1, 2).contiguous().view(b, -1, self.num_heads * d)
        return self.output_layer(x)


class EncoderLayer(nn.Module):
    """:1714-1743 — post-LN variant of the live model (self_attention_norm exists but is unused)"""

    def __init__(self, hidden, ffn, heads, dropout_rate=0.0, attention_dropout_rate=0.0):
        super().__init__()
        self.self_attention_norm = nn.LayerNorm(hidden)
        self.self_attention = MultiHeadAttention(hidden, heads, attention_dropout_rate)
        self.self_attention_dropout = nn.Dropout(dropout_rate)                                # :1724
        self.ffn_dropout = nn.Dropout(dropout_rate)                                           # :1729
        self.ffn_norm1 = nn.LayerNorm(hidden)
        self.ffn_norm2 = nn.LayerNorm(hidden)
        self.ffn = FeedForwardNetwork(hidden, ffn)

    def forward(self, x, attn_bias):
        x = x + self.self_attention_dropout(self.self_attention(x, x, x, attn_bias))          # :1731-1735
        x = x + self.ffn_dropout(self.ffn(self.ffn_norm1(x)))                                 # :1737-1741
        return self.ffn_norm2(x)


DATASET_TRAITS = {
    # time rows, time padding_idx, user rows (+1?), cat_decoder extra, out_proj extra, log_softmax
    "foursquaregraph": dict(time_rows=49, time_pad=0, user_extra=0, cat_extra=0, poi_extra=0, log_softmax=False),   # :781-900
    "gowalla_nevda": dict(time_rows=48, time_pad=0, user_extra=0, cat_extra=1, poi_extra=1, log_softmax=False),     # :636-779
    "gowalla_7day": dict(time_rows=48, time_pad=0, user_extra=0, cat_extra=1, poi_extra=1, log_softmax=False),
    "toyotagraph": dict(time_rows=48, time_pad=None, user_extra=1, cat_extra=0, poi_extra=1, log_softmax=True),     # :902-1029
}


def csr_to_dense(csr, n, sparse=False):
    crow, col, val = csr
    rows = np.repeat(np.arange(n), np.diff(crow))
    if sparse:
        idx = torch.from_numpy(np.stack([rows, np.asarray(col)]))
        return torch.sparse_coo_tensor(idx, torch.from_numpy(np.asarray(val)), (n, n)).coalesce()
    a = torch.zeros(n, n)
    a[torch.from_numpy(rows), torch.from_numpy(np.asarray(col))] = torch.from_numpy(np.asarray(val))
    return a


class Graphormer(nn.Module):
    """model_fqandtoyo.py:580-1121 (__init__) and :1123-1432 (forward), POI dataset branches."""

    def __init__(self, world, n_layers=6, num_heads=8, hidden_dim=128, ffn_dim=1024, multi_hop_max_dist=20,
                 dataset_name=None, dropout_rate=0.0, intput_dropout_rate=0.0, attention_dropout_rate=0.0, pos_dropout=0.0):
        """All dropout rates default to 0: parity is defined with dropout off.  The timed CPU arm (bench.py) passes the
        canonical rates (README.md:62 flags; LearnablePositionalEncoding's own Dropout(0.1), :334) so it does the same work."""
        super().__init__()
        self.input_dropout = nn.Dropout(intput_dropout_rate)                                  # :1041
        self.output_dropout = nn.Dropout(intput_dropout_rate)                                 # :770, :892, :1021 (same rate)
        self.pos_dropout = nn.Dropout(pos_dropout)                                            # :334, applied :358
        self.dataset_name = dataset_name or world.dataset_name
        tr = DATASET_TRAITS[self.dataset_name]
        self.traits = tr
        self.num_heads, self.hidden_dim, self.multi_hop_max_dist = num_heads, hidden_dim, multi_hop_max_dist
        H, C, P = num_heads, world.C, world.P
        self.time_embed_dim = self.cat_embed_dim = 32
        D = hidden_dim + 64
        self.edge_encoder = nn.Embedding(128, H, padding_idx=0)
        self.edge_dis_encoder = nn.Embedding(128 * H * H, 1)
        self.rel_pos_encoder = nn.Embedding(512, H, padding_idx=0)
        self.poi_distance_model = GCN(3 + C, [16, 64], hidden_dim, 0.3)
        self.poi_cat_model = GCN(C, [16, 64], 32, 0.1)
        self.user_embed_model = UserEmbeddings(world.U + tr["user_extra"], hidden_dim)
        self.time_embed_model_48 = nn.Embedding(tr["time_rows"], 32, padding_idx=tr["time_pad"])
        self.cat_decoder = nn.Linear(2 * hidden_dim + 64, C + tr["cat_extra"])
        self.embed_fuse_model2 = FuseEmbeddings(hidden_dim, 32)
        self.embed_fuse_model3 = FuseEmbeddings(hidden_dim, hidden_dim + 64)
        self.embed_fuse_model4 = FuseEmbeddings(hidden_dim + 32, 32)
        self.pos_embed = LearnablePositionalEncoding(NODE_DIM, D)
        self.in_degree_encoder = nn.Embedding(128, D, padding_idx=0)
        self.out_degree_encoder = nn.Embedding(128, D, padding_idx=0)
        self.fre_embed_model = nn.Embedding(int(world.check_freq.max()) + 1, D, padding_idx=0)
        self.poi_pos_encoder = nn.Embedding(world.num_bins, H, padding_idx=0)
        self.layers = nn.ModuleList([EncoderLayer(D, ffn_dim, H, dropout_rate, attention_dropout_rate) for _ in range(n_layers)])
        self.final_ln = nn.LayerNorm(2 * hidden_dim + 64)
        self.out_proj = nn.Linear(2 * hidden_dim + 64, P + tr["poi_extra"])
        self.graph_token = nn.Embedding(1, D)
        self.graph_token_virtual_distance = nn.Embedding(1, H)
        # non-parameter tables (model_fqandtoyo.py:791-832, 1106-1108)
        self.register_buffer("X", torch.from_numpy(world.X))
        self.register_buffer("C_X", torch.from_numpy(world.C_X))
        self.D_A = csr_to_dense(world.D_A, P, sparse=P > 10000)
        self.C_A = csr_to_dense(world.C_A, C)
        self.register_buffer("cat_of_poi", torch.from_numpy(world.cat_of_poi))     # poi_idx2cat_idx_dict

    # -------------------------------------------------------------------- A2: attention bias
    def attn_bias_build(self, b, mode="fp32"):
        H = self.num_heads
        attn_bias, rel_pos, poi_pos, edge_input = b.attn_bias, b.rel_pos, b.poi_pos, b.edge_input
        n_graph, n_node = b.x.size()[:2]
        g = attn_bias.clone().unsqueeze(1).repeat(1, H, 1, 1)                                # :1143-1147
        g[:, :, 1:, 1:] = g[:, :, 1:, 1:] + self.rel_pos_encoder(rel_pos).permute(0, 3, 1, 2) \
            + self.poi_pos_encoder(poi_pos).permute(0, 3, 1, 2)                                 # :1151-1158
        t = self.graph_token_virtual_distance.weight.view(1, H, 1).unsqueeze(-2)
        g[:, :, 1:, :1] = g[:, :, 1:, :1] + t                                                 # :1160-1165
        rp = rel_pos.clone()                                                                  # :1169-1175
        rp[rp == 0] = 1
        rp = torch.where(rp > 1, rp - 1, rp).clamp(0, self.multi_hop_max_dist)
        ei = edge_input[:, :, :, :self.multi_hop_max_dist, :]
        e = self.edge_encoder(ei).mean(-2)                                                    # :1177-1178
        if mode == "ref_half":
            e = e.half().float()
        max_dist = e.size(-2)
        flat = e.permute(3, 0, 1, 2, 4).reshape(max_dist, -1, H)
        W = self.edge_dis_encoder.weight.reshape(-1, H, H)[:max_dist]
        if mode == "ref_half":                                                                # :1184-1198
            flat = torch.bmm(flat.half().float(), W.half().float()).half().float()
        else:
            flat = torch.bmm(flat, W)                                                         # model.py:171-176
        e = flat.reshape(max_dist, n_graph, n_node, n_node, H).permute(1, 2, 3, 0, 4)
        e = (e.sum(-2) / rp.float().unsqueeze(-1)).permute(0, 3, 1, 2)                        # :1206-1208
        g[:, :, 1:, 1:] = g[:, :, 1:, 1:] + e                                                 # :1213-1215
        return g + attn_bias.unsqueeze(1)                                                     # :1216

    # -------------------------------------------------------------------- A4: node embeddings
    def gcn_tables(self):
        return self.poi_distance_model(self.X, self.D_A), self.poi_cat_model(self.C_X, self.C_A)   # :1236-1237

    def node_features(self, b, looped=False):
        x, time_normal = b.x, b.time_normal
        B, N = x.size()[:2]
        Gd, Gc = self.gcn_tables()
        indx = (x != 0).sum(dim=-2)                                                           # :1225-1227
        D = self.hidden_dim + 64
        nf = torch.zeros(B, N, D)
        cat_target = b.y.clone()
        if looped:                                                                            # :1257-1269 verbatim
            for p in range(B):
                L = int(indx[p][0])
                ce = Gc[torch.LongTensor([int(self.cat_of_poi[int(x[p][q]) - 1]) - 1 for q in range(L)])]
                te = self.time_embed_model_48((time_normal[p][:L] * 48).long()).squeeze(1)
                pe_ = Gd[x[p][:L] - 1].squeeze(1)
                f2 = self.embed_fuse_model2(pe_, te)
                nf[p, :L] = self.embed_fuse_model4(f2, ce).to(nf.dtype)      # (.to: a no-op in fp32; lets torch.autocast run this module)
        else:
            mask = (x[:, :, 0] != 0)
            xi = x[:, :, 0][mask]
            ce = Gc[self.cat_of_poi[xi - 1] - 1]
            te = self.time_embed_model_48((time_normal[:, :, 0][mask] * 48).long())
            f2 = self.embed_fuse_model2(Gd[xi - 1], te)
            nf[mask] = self.embed_fuse_model4(f2, ce).to(nf.dtype)
        for p in range(B):
            cat_target[p] = self.cat_of_poi[int(b.y[p]) - 1] - 1                              # :1265
        nf = nf + self.fre_embed_model(torch.zeros(B, N, dtype=torch.long)) \
            + self.in_degree_encoder(b.in_degree) + self.out_degree_encoder(b.out_degree)     # :1288-1298
        pe = self.pos_embed.pe
        out = nf.clone()
        for i in range(B):                                                                    # :348-351 'node_reverse'
            L = int(indx[i][0])
            out[i, :L] = nf[i, :L] + pe[1:L + 1]
        tok = self.graph_token.weight.unsqueeze(0).repeat(B, 1, 1) + pe[0]                    # :1338-1342
        return torch.cat([tok, out], dim=1), cat_target                                       # :1344

    # -------------------------------------------------------------------- forward
    def forward(self, b, bias_mode="fp32", looped=False, return_internals=False):
        bias = self.attn_bias_build(b, bias_mode)
        h, cat_target = self.node_features(b, looped)
        self.cat_target = cat_target
        h0 = h
        h = self.input_dropout(self.pos_dropout(h))                                           # :358, :1347
        for layer in self.layers:                                                             # :1348-1352
            h = layer(h, bias)
        user_embedding = self.user_embed_model(b.user - 1).squeeze(1)                         # :1239-1240
        B, N = b.x.size()[:2]
        if looped:                                                                            # :1353-1358 verbatim
            tmp = torch.zeros(B, N, 2 * self.hidden_dim + 64)
            for p in range(B):
                fused = [self.embed_fuse_model3(h[p][q], user_embedding[p]) for q in range(N)]
                tmp[p] = torch.cat(fused).reshape(N, -1)
            z = tmp[:, 0, :]
        else:   # only token 0 is consumed downstream (:1394-1396) -- equal to the looped form on [:,0,:]
            z = self.embed_fuse_model3(h[:, 0, :], user_embedding)
        z = self.output_dropout(F.elu(self.final_ln(z)))                                      # :1360-1364
        cat_logits = self.cat_decoder(z)
        poi_logits = self.out_proj(z)
        if self.traits["log_softmax"]:
            poi_logits = F.log_softmax(poi_logits, dim=1)                                     # :1425 (implicit dim=1 for 2-D)
        if return_internals:
            return [poi_logits, cat_logits], dict(bias=bias, h0=h0, h=h, z=z)
        return [poi_logits, cat_logits]

    # -------------------------------------------------------------------- losses :1434-1478
    def training_loss(self, b, **kw):
        y_out = self(b, **kw)
        if self.dataset_name == "toyotagraph":
            loss1 = gradient_tail_loss(y_out[1], self.cat_target, 0.1)
            loss2 = F.nll_loss(y_out[0], b.y, ignore_index=0)                                  # data.py:165
            return loss1 + loss2
        return gradient_tail_loss(y_out[0], b.y - 1, 0.2)


def gradient_tail_loss(inputs, targets, alpha=0.25, beta=1, k=1):
    """model_fqandtoyo.py:545-550"""
    one_hot = torch.zeros_like(inputs)
    one_hot.scatter_(1, targets[:len(inputs)].view(-1, 1), 1)
    prob = torch.sigmoid(inputs)
    loss = -alpha * (1 - prob) ** k * one_hot * torch.log(prob) - (1 - one_hot) * beta * prob ** k * torch.log(1 - prob)
    return loss.mean()


# ------------------------------------------------------------------------------------ metrics
def get_acc(target, scores):
    """model_fqandtoyo.py:48-90 verbatim semantics (including the `break` at the first target == 0)."""
    target = target.cpu().numpy()
    _, idxx = scores.topk(20, 1)
    predx = idxx.cpu().numpy()
    acc = np.zeros((4, 1))
    ndcg = np.zeros((4, 1))
    for i, p in enumerate(predx):
        t = target[i]
        if t != 0:
            for slot, kk in ((3, 20), (0, 10), (1, 5), (2, 1)):
                if t in p[:kk] and t > 0:
                    acc[slot] += 1
                    ndcg[slot] += 1.0 / np.log2(list(p[:kk]).index(t) + 2)
        else:
            break
    return acc, ndcg


def mrr_metric(target, scores):
    """model_fqandtoyo.py:122-131"""
    y_true = target.cpu().numpy()
    y_pred = scores.cpu().numpy()
    mrr = 0
    for j in range(len(y_pred)):
        rec_list = y_pred[j].argsort()[-len(y_pred[j]):][::-1]
        r_idx = np.where(rec_list == y_true[j])[0][0]
        mrr += 1 / (r_idx + 1)
    return mrr


def polynomial_decay_lr(step_count, warmup_updates, tot_updates, lr, end_lr, power=1.0):
    """lr.py:18-31"""
    if step_count <= warmup_updates:
        return step_count / float(warmup_updates) * lr
    if step_count >= tot_updates:
        return end_lr
    pct = 1 - (step_count - warmup_updates) / (tot_updates - warmup_updates)
    return (lr - end_lr) * pct ** power + end_lr
