"""CSPNet score network on libmatinvent_b200 (forward AND hand-written backward).

Host-side mirror of the reference's `CSPNet` (models/diffcsp/cspnet.py:93-294): same constructor
arguments, same `forward(t, atom_types, frac_coords, lattices, num_atoms, node2graph)` returning
`(lattice_out [B,3,3], coord_out [N,3], type_out [N,100])`, same `state_dict` names, so it drops in
under DiffCSPModule / models/suite.  Everything arithmetic runs in the CUDA library:

  * weights live in ONE flat fp32 buffer (flat Adam + a single gradient all-reduce); the first edge
    linear (cspnet.py:45, in = [h_i | h_j | vec(L L^T) | Phi]) is stored SPLIT as W_pq = [W_hi; W_hj],
    W_L, W_F so that  W1 . [h_i|h_j|ips|Phi] = P_i + Q_j + C_b + Phi_ij W_F^T  — the h_i / h_j parts
    become per-NODE GEMMs (n instead of n^2 rows) and the `[E,1801]` concat is never formed;
  * the Fourier basis (cspnet.py:12-24) is evaluated once per forward, not once per layer (:65-66);
  * edges are static per batch (graph.py) instead of block_diag + nonzero per forward (:238-242).
"""
import math
import os
from collections import OrderedDict

import torch
import torch.nn as nn

from ... import ops
from ...ops import ACT_DSILU, ACT_NONE, ACT_SILU
from .graph import CrystalGraph

MAX_ATOMIC_NUM = 100


def _pad8(n):
    return (n + 7) // 8 * 8      # 32-byte blocks: the fp16 split copies stay 16-byte aligned for TMA


class _Workspace:
    """Preallocated activations for one (graph size, mode); addresses are stable (CUDA-graph safe)."""

    def __init__(self, net, g, train):
        dev, f32 = g.device, torch.float32
        H, A, F6, L = net.hidden_dim, net.max_atoms, 6 * net.num_freqs, net.num_layers
        N, E, B = g.N, getattr(g, "E_cap", g.E), g.B
        nl = L if train else 1

        def buf(*shape):
            return torch.empty(*shape, device=dev, dtype=f32)

        self.train = train
        self.g = g                     # keeps the graph alive: the cache key must not be recycled under this workspace
        self._H, self._F6, self._dev = H, F6, dev
        self.phi = buf(E, F6)                               # fp32 basis: backward (dW_F) and the FFMA path
        self.phi_hi = torch.empty(E, F6, device=dev, dtype=torch.float16)
        self.phi_lo = torch.empty(E, F6, device=dev, dtype=torch.float16)
        self.ips = buf(B, 9)
        self.tb = buf(B, H)
        # [C_b | 0] of every layer: folded into P by the P|Q GEMM's epilogue (one launch fills all layers: the lattices
        # do not change inside a forward)
        self.cb = torch.zeros(net.num_layers, B, 3 * H, device=dev, dtype=f32)
        self.cb2 = self.cb[:, :, :2 * H]
        self.h0 = buf(N, H)
        # inference: the embedding output keeps its own buffer (the predictor forward of a reverse step reuses the
        # corrector's: same atom-type state, time and lattice), the layers update one shared buffer in place
        self.h = [buf(N, H) for _ in range(L + 1)] if train else [buf(N, H)] + [buf(N, H)] * L
        self.cat = [torch.zeros(N, 2 * H, device=dev, dtype=f32) for _ in range(nl)]
        # [P' | Q | R]: the per-node parts of the first edge linear and, third block, the LN(h) half of node_mlp.0
        # (forward_graph, "node path"); the FP32 / unfused path uses the first two blocks only
        self.pqr = buf(N, 3 * H)
        self.pq = self.pqr[:, :2 * H]
        # layer 0's [P'|Q|R] and its row maxima keep their own buffers: they depend on (atom types, time, lattice) only, so
        # the predictor evaluation of a reverse step reuses the corrector's (forward_graph(reuse_embedding=True))
        self.pqr0 = buf(N, 3 * H)
        self.amax_pqr0 = torch.zeros(N, device=dev, dtype=f32)
        self.hn_hi = torch.empty(N, H, device=dev, dtype=torch.float16)      # LN(h) as a pre-split tensor-core operand
        self.hn_lo = torch.empty(N, H, device=dev, dtype=torch.float16)
        # operand pairs written by the node chain's epilogues (mi_node_chain): xs = split(agg) / split(LN(h)), ys = split(an1)
        # (hi and lo are the two planes of ONE allocation: a single 3-D TMA operation loads both)
        self.xs = tuple(torch.empty(2, N, H, device=dev, dtype=torch.float16))
        self.ys = tuple(torch.empty(2, N, H, device=dev, dtype=torch.float16))
        self.a1 = [buf(E, H) for _ in range(nl)]
        self.a2 = buf(E, H)
        self.an1 = [buf(N, H) for _ in range(nl)]
        self.hf = buf(N, H)
        self.gmean = buf(B, H)
        self.lat9 = buf(B, 9)
        self.pred_l = buf(B, 3, 3)
        self.pred_x = buf(N, 3)
        self.pred_a = buf(N, A)
        # row-wise max |.| of the activations that feed tensor-core GEMMs (power-of-two row rescaling keeps the
        # fp16 split path inside fp32 dynamic range); zeroed at the start of every forward
        self.amax = torch.zeros(N + L * (E + 4 * N), device=dev, dtype=f32)
        self.amax_h0 = self.amax[:N]
        o = N
        self.amax_a1, self.amax_agg, self.amax_an1, self.amax_hn, self.amax_pqr = [], [], [], [], []
        for _ in range(L):
            self.amax_a1.append(self.amax[o:o + E]); o += E
            self.amax_agg.append(self.amax[o:o + N]); o += N
            self.amax_an1.append(self.amax[o:o + N]); o += N
            self.amax_hn.append(self.amax[o:o + N]); o += N
            self.amax_pqr.append(self.amax[o:o + N]); o += N      # row maxima of [P'|Q|R]: the bounds of the node chain
        if train:
            self.z1 = [buf(E, H) for _ in range(L)]
            self.z2 = [buf(E, H) for _ in range(L)]
            self.zn1 = [buf(N, H) for _ in range(L)]
            self.zn2 = [buf(N, H) for _ in range(L)]
            self.ln_mean = [buf(N) for _ in range(L + 1)]
            self.ln_rstd = [buf(N) for _ in range(L + 1)]
            # backward scratch
            self.dh = buf(N, H)
            self.dhf = buf(N, H)
            self.dcat = buf(N, 2 * H)
            self.dpq = buf(N, 2 * H)
            self.dzn = buf(N, H)
            self.dzn1 = buf(N, H)
            self.dz2 = buf(E, H)
            self.dz1 = buf(E, H)
            self.dcb = buf(B, H)
            self.dg = buf(B, H)
            self.dgn = buf(N, H)
            self.dlat9 = buf(B, 9)
            self.dtb = buf(B, H)
            # row maxima of the gradient operands of the tensor-core input-gradient GEMMs, per layer (gradients span
            # many binades: the fp16 split needs the power-of-two row rescaling); zeroed at the start of a backward
            self.gamax = torch.zeros(L * (3 * N + E), device=dev, dtype=f32)
            self.amax_dzn, self.amax_dzn1, self.amax_dz2, self.amax_dpq = [], [], [], []
            o = 0
            for _ in range(L):
                self.amax_dzn.append(self.gamax[o:o + N]); o += N
                self.amax_dzn1.append(self.gamax[o:o + N]); o += N
                self.amax_dpq.append(self.gamax[o:o + N]); o += N
                self.amax_dz2.append(self.gamax[o:o + E]); o += E


    def wgrad_scratch(self, K):
        """transposed operands of the tensor-core weight-gradient GEMMs (allocated on first use; K = reduction rows)"""
        sc = getattr(self, "_wg", None)
        Kp = (K + 7) // 8 * 8
        if sc is None or sc["Kp"] < Kp:
            H2, dev = 2 * self._H, self._dev
            F6p = max(H2, self._F6)
            sc = self._wg = dict(
                Kp=Kp, dyT=torch.zeros(H2, Kp, device=dev, dtype=torch.float32),
                xhi=torch.zeros(F6p, Kp, device=dev, dtype=torch.float16), xlo=torch.zeros(F6p, Kp, device=dev, dtype=torch.float16),
                phi_hi=torch.zeros(self._F6, Kp, device=dev, dtype=torch.float16),
                phi_lo=torch.zeros(self._F6, Kp, device=dev, dtype=torch.float16),
                inv=torch.zeros(F6p, device=dev, dtype=torch.float32), phi_inv=torch.zeros(self._F6, device=dev, dtype=torch.float32),
                amax=torch.zeros(H2 + F6p, device=dev, dtype=torch.float32))
        return sc


def an1_contig(ws):
    return ws.an1[0].is_contiguous()


class CSPNet(nn.Module):
    def __init__(self, hidden_dim=128, latent_dim=256, num_layers=4, max_atoms=100, act_fn="silu",
                 dis_emb="sin", num_freqs=10, edge_style="fc", cutoff=6.0, max_neighbors=20, ln=False,
                 ip=True, smooth=False, pred_type=False, pred_scalar=False, device=None):
        super().__init__()
        if act_fn != "silu" or dis_emb != "sin":
            raise NotImplementedError("matinvent_b200 CSPNet implements act_fn='silu', dis_emb='sin' (the MatInvent config)")
        if not (smooth and pred_type) or pred_scalar:
            raise NotImplementedError("matinvent_b200 CSPNet implements the smooth=True, pred_type=True generation head "
                                      "(models/diffcsp/diffusion.py:73)")
        if hidden_dim % 4 or latent_dim % 4 or max_atoms % 4:
            raise ValueError("hidden_dim, latent_dim and max_atoms must be multiples of 4")
        self.hidden_dim, self.latent_dim, self.num_layers = hidden_dim, latent_dim, num_layers
        self.max_atoms, self.num_freqs = max_atoms, num_freqs
        self.edge_style, self.cutoff, self.max_neighbors = edge_style, cutoff, max_neighbors
        self.ln, self.ip = ln, ip
        dev = torch.device(device if device is not None else "cuda")
        H, A, T, F6 = hidden_dim, max_atoms, latent_dim, 6 * num_freqs
        spec = [("emb_w", (H, A)), ("emb_b", (H,)), ("lat_w_h", (H, H)), ("lat_w_t", (H, T)), ("lat_b", (H,))]
        for i in range(num_layers):
            p = "l%d." % i
            spec += [(p + "ln_g", (H,)), (p + "ln_b", (H,)), (p + "w_pq", (2 * H, H)), (p + "w_l", (H, 9)),
                     (p + "w_f", (H, F6)), (p + "b1", (H,)), (p + "w2", (H, H)), (p + "b2", (H,)),
                     (p + "wn1", (H, 2 * H)), (p + "bn1", (H,)), (p + "wn2", (H, H)), (p + "bn2", (H,))]
        spec += [("fin_g", (H,)), ("fin_b", (H,)), ("coord_w", (3, H)), ("lattice_w", (9, H)),
                 ("type_w", (MAX_ATOMIC_NUM, H)), ("type_b", (MAX_ATOMIC_NUM,))]
        self._spec = spec
        off, self._slices = 0, OrderedDict()
        for name, shape in spec:
            n = math.prod(shape)
            self._slices[name] = (off, n, shape)
            off += _pad8(n)
        self.flat = nn.Parameter(torch.zeros(off, device=dev, dtype=torch.float32))
        self._views, self._gviews = {}, {}
        self._flat_grad = None
        self._ws = {}
        self._graphs = {}
        # tensor-core path (mi_tc_gemm): fp16 head / scaled tail of every weight, refreshed when the weights change
        self.use_tc = True
        self._flat_hi = self._flat_lo = None
        self._hi, self._lo = {}, {}
        # merged-format copies (per-row scaled, single-accumulator 128x256 tiles) of the per-edge weights: used when
        # the edge count fills the machine with 256-wide tiles (see forward_graph)
        self.use_merged = os.environ.get("MI_TC_MERGED", "1") != "0"
        self.force_merged = False      # merged tiles / CTA-pair kernels at any edge count (small parity runs on the benchmark's kernels)
        # per-edge blocks on CTA pairs (csrc/mi_edge.cu): inference, merged tiles, LayerNorm'd node path
        self.use_pair = os.environ.get("MI_EDGE_PAIR", "1") != "0"
        self.compose_embedding = os.environ.get("MI_COMPOSE_EMB", "1") != "0"     # composed embedding GEMM (inference)
        self.fused_heads = os.environ.get("MI_FUSED_HEADS", "1") != "0"    # one-launch output heads (inference)
        self.use_chain = os.environ.get("MI_NODE_CHAIN", "1") != "0"      # fused node-level chain (inference, H = 512)
        self._mhi, self._mlo, self._minv = {}, {}, {}
        self._pqr_hi, self._pqr_lo = {}, {}
        self._bounds = None
        self._emb_hl = self._emb_bias = None
        # transposed copies W^T (fp16 head / tail) of the weights whose input gradients run on the tensor cores
        # (dX = dY W is the forward kernel with W^T as its weight); built once a backward has asked for them
        self._hiT, self._loT, self._wT = {}, {}, {}
        self._need_T = False
        self._tc_version = None
        self.reset_parameters()

    # ------------------------------------------------------------------ parameters
    def _rebuild_views(self):
        d = self.flat.data
        self._views = {k: d[o:o + n].view(shape) for k, (o, n, shape) in self._slices.items()}
        if self._flat_grad is not None:
            g = self._flat_grad
            self._gviews = {k: g[o:o + n].view(shape) for k, (o, n, shape) in self._slices.items()}

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._flat_grad = None
        self._gviews = {}
        self._rebuild_views()
        self._ws, self._graphs = {}, {}
        self._flat_hi = self._flat_lo = None
        self._mhi, self._mlo, self._minv = {}, {}, {}
        self._pqr_hi, self._pqr_lo = {}, {}
        self._bounds = None
        self._emb_hl = self._emb_bias = None
        self._hiT, self._loT, self._wT = {}, {}, {}
        self._tc_version = None
        return r

    def weights_changed(self):
        """Call after the flat weight buffer was written outside torch (mi_adam_step): the fp16 split copies
        used by the tensor-core GEMMs are rebuilt on the next forward."""
        self._tc_version = None
        if self.use_tc and self._flat_hi is not None:
            self._refresh_tc()          # immediately: captured CUDA graphs read the split copies

    def _refresh_tc(self):
        ver = self.flat._version
        if self._tc_version == ver and self._flat_hi is not None:
            return
        if self._flat_hi is None:
            # hi and lo are the two planes of one allocation (operand pairs are loaded by single 3-D TMA operations)
            self._flat_hi, self._flat_lo = torch.empty(2, self.flat.numel(), device=self.flat.device, dtype=torch.float16)
            self._hi = {k: self._flat_hi[o:o + n].view(shape) for k, (o, n, shape) in self._slices.items()}
            self._lo = {k: self._flat_lo[o:o + n].view(shape) for k, (o, n, shape) in self._slices.items()}
        ops.f16_split(self.flat.data, self._flat_hi, self._flat_lo)
        if self.use_merged:
            for i in range(self.num_layers):
                for k in ("l%d.w_f" % i, "l%d.w2" % i):
                    w = self._views[k]
                    if k not in self._mhi:
                        self._mhi[k] = torch.empty_like(w, dtype=torch.float16)
                        self._mlo[k] = torch.empty_like(w, dtype=torch.float16)
                        self._minv[k] = torch.empty(w.shape[0], device=w.device, dtype=torch.float32)
                    ops.f16_split_rows(w, self._mhi[k], self._mlo[k], self._minv[k])
        # packed [W_hi; W_hj; node_mlp.0[:, :H]] rows (fp16 head / tail) of every layer: one per-node GEMM produces P, Q and
        # the LN(h) half of node_mlp.0 (forward_graph, node path)
        H = self.hidden_dim
        for i in range(self.num_layers):
            q = "l%d." % i
            if i not in self._pqr_hi:
                self._pqr_hi[i], self._pqr_lo[i] = torch.empty(2, 3 * H, H, device=self.flat.device, dtype=torch.float16)
            for dst, src in ((self._pqr_hi[i], self._hi), (self._pqr_lo[i], self._lo)):
                dst[:2 * H].copy_(src[q + "w_pq"])
                dst[2 * H:].copy_(src[q + "wn1"][:, :H])
        # inference: node_embedding and the h part of atom_latent_emb are two linear maps in a row (cspnet.py:264-271), composed
        # once per weight update:  h = a (W_h E)^T + [temb W_t^T + b_l + W_h b_e][crystal]   (one K = A GEMM instead of two)
        A8 = _pad8(self.max_atoms)
        if self._emb_hl is None:
            self._emb_hl = torch.zeros(2, H, A8, device=self.flat.device, dtype=torch.float16)
            self._emb_bias = torch.empty(H, device=self.flat.device, dtype=torch.float32)
        wh, E_ = self._views["lat_w_h"].double(), self._views["emb_w"].double()
        wae = torch.zeros(H, A8, device=self.flat.device, dtype=torch.float32)
        wae[:, :self.max_atoms] = (wh @ E_).float()
        ops.f16_split(wae, self._emb_hl[0], self._emb_hl[1])
        self._emb_bias.copy_((self._views["lat_b"].double() + wh @ self._views["emb_b"].double()).float())
        # a-priori bounds of the node chain's row scales (mi_node_chain), rounded up: {max_j ||W_b[j]||_1, max |b_n1|,
        # sqrt(H) max |gamma| + max |beta| of the NEXT layer's LayerNorm}
        if self._bounds is None:
            self._bounds = torch.zeros(self.num_layers, 4, device=self.flat.device, dtype=torch.float32)
        for i in range(self.num_layers):
            q = "l%d." % i
            self._bounds[i, 0] = self._views[q + "wn1"][:, H:].abs().sum(1).max() * 1.001
            self._bounds[i, 1] = self._views[q + "bn1"].abs().max() * 1.001
            # max |Phi W_F^T| <= sqrt(3F) max_j ||W_F[j]||_2: every sin / cos pair of the Fourier basis has unit norm
            self._bounds[i, 3] = math.sqrt(3 * self.num_freqs) * self._views[q + "w_f"].norm(dim=1).max() * 1.001
            if i + 1 < self.num_layers:
                qn = "l%d." % (i + 1)
                self._bounds[i, 2] = (math.sqrt(H) * self._views[qn + "ln_g"].abs().max() + self._views[qn + "ln_b"].abs().max()) * 1.001
        if self._need_T:
            for i in range(self.num_layers):
                for k in ("l%d.wn2" % i, "l%d.wn1" % i, "l%d.w2" % i, "l%d.w_pq" % i):
                    w = self._views[k]
                    if k not in self._hiT:
                        self._wT[k] = torch.empty(w.shape[1], w.shape[0], device=w.device, dtype=torch.float32)
                        self._hiT[k] = torch.empty_like(self._wT[k], dtype=torch.float16)
                        self._loT[k] = torch.empty_like(self._wT[k], dtype=torch.float16)
                    self._wT[k].copy_(w.t())
                    ops.f16_split(self._wT[k], self._hiT[k], self._loT[k])
        self._tc_version = ver

    def _linear(self, A, wname, C, M, **epi):
        """C = epilogue(A @ W^T): tensor cores (split FP16) when the operands are TMA-compatible, else FP32 FFMA."""
        W = self._views[wname]
        if self.use_tc and ops.tc_ok(A, self._hi[wname]):
            return ops.tc_gemm(A, self._hi[wname], self._lo[wname], C, M=M, **epi)
        return ops.sgemm(A, W, C, M=M, **epi)

    def tc_linear_view(self, A, wname, cols, C, M, **epi):
        """C = epilogue(A @ W[:, cols]^T) on the tensor cores (a column block of a weight: K = len(cols))"""
        return ops.tc_gemm(A, self._hi[wname][:, cols], self._lo[wname][:, cols], C, M=M, **epi)

    def _dgrad(self, dY, wname, C, M, a_amax, act=ACT_NONE, z_in=None, amax_out=None, accumulate=False):
        """C = dY @ W (* silu'(z_in)) (+ C): input gradient of y = x W^T.  Tensor cores (the forward kernel with W^T as
        its weight operand, rows of dY rescaled from their maxima) or the FP32 CUDA-core NN GEMM."""
        W = self._views[wname]
        N_out, K_in = W.shape
        if self.use_tc and wname in self._hiT and ops.tc_ok(dY, self._hiT[wname]) and C.stride(0) % 4 == 0:
            return ops.tc_gemm(dY, self._hiT[wname], self._loT[wname], C, M=M, act=act, z_in=z_in, a_amax=a_amax,
                               amax_out=amax_out, resid=C if accumulate else None)
        return ops.sgemm(dY, W, C, transB=False, M=M, N=K_in, K=N_out, act=act, z_in=z_in, beta=1.0 if accumulate else 0.0,
                         amax_out=amax_out)

    @property
    def device(self):
        return self.flat.device

    def w(self, name):
        """Native weight block (a view into the flat buffer).  After writing to it call weights_changed()."""
        return self._views[name]

    def flat_grad(self):
        """Flat fp32 gradient buffer, aliased by `self.flat.grad` (so torch.optim / autograd see it)."""
        if self._flat_grad is None:
            self._flat_grad = torch.zeros_like(self.flat.data)
            self._rebuild_views()
        cur = self.flat.grad
        if cur is None:                       # never attached, or dropped by zero_grad(set_to_none=True)
            self._flat_grad.zero_()
            self.flat.grad = self._flat_grad
        elif cur.data_ptr() != self._flat_grad.data_ptr():
            self._flat_grad.copy_(cur)
            self.flat.grad = self._flat_grad
        return self._flat_grad

    def reset_parameters(self, seed=None):
        """nn.Linear / nn.LayerNorm default init of the reference's modules (uniform +-1/sqrt(fan_in))."""
        gen = None
        if seed is not None:
            gen = torch.Generator().manual_seed(seed)
        sd = OrderedDict()
        H, A, T, F6 = self.hidden_dim, self.max_atoms, self.latent_dim, 6 * self.num_freqs

        def lin(name, out_f, in_f, bias=True):
            k = 1.0 / math.sqrt(in_f)
            sd[name + ".weight"] = (torch.rand(out_f, in_f, generator=gen) * 2 - 1) * k
            if bias:
                sd[name + ".bias"] = (torch.rand(out_f, generator=gen) * 2 - 1) * k

        lin("node_embedding", H, A)
        lin("atom_latent_emb", H, H + T)
        for i in range(self.num_layers):
            p = "csp_layer_%d." % i
            lin(p + "edge_mlp.0", H, 2 * H + 9 + F6)
            lin(p + "edge_mlp.2", H, H)
            lin(p + "node_mlp.0", H, 2 * H)
            lin(p + "node_mlp.2", H, H)
            if self.ln:
                sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"] = torch.ones(H), torch.zeros(H)
        lin("coord_out", 3, H, bias=False)
        lin("lattice_out", 9, H, bias=False)
        if self.ln:
            sd["final_layer_norm.weight"], sd["final_layer_norm.bias"] = torch.ones(H), torch.zeros(H)
        lin("type_out", MAX_ATOMIC_NUM, H)
        self._load_reference_state(sd)

    def _load_reference_state(self, sd, prefix=""):
        """Scatter reference-named tensors (cspnet.py:118-146 module names) into the flat buffer."""
        H = self.hidden_dim
        self._rebuild_views()
        v = self._views

        def put(dst, src):
            if tuple(src.shape) != tuple(dst.shape):
                raise ValueError("shape mismatch %s vs %s" % (tuple(src.shape), tuple(dst.shape)))
            dst.copy_(src.to(dst.device, torch.float32))

        g = lambda k: sd[prefix + k]
        put(v["emb_w"], g("node_embedding.weight")), put(v["emb_b"], g("node_embedding.bias"))
        wl = g("atom_latent_emb.weight")
        put(v["lat_w_h"], wl[:, :H]), put(v["lat_w_t"], wl[:, H:]), put(v["lat_b"], g("atom_latent_emb.bias"))
        for i in range(self.num_layers):
            p, q = "csp_layer_%d." % i, "l%d." % i
            w1 = g(p + "edge_mlp.0.weight")
            put(v[q + "w_pq"][:H], w1[:, :H]), put(v[q + "w_pq"][H:], w1[:, H:2 * H])
            put(v[q + "w_l"], w1[:, 2 * H:2 * H + 9]), put(v[q + "w_f"], w1[:, 2 * H + 9:])
            put(v[q + "b1"], g(p + "edge_mlp.0.bias"))
            put(v[q + "w2"], g(p + "edge_mlp.2.weight")), put(v[q + "b2"], g(p + "edge_mlp.2.bias"))
            put(v[q + "wn1"], g(p + "node_mlp.0.weight")), put(v[q + "bn1"], g(p + "node_mlp.0.bias"))
            put(v[q + "wn2"], g(p + "node_mlp.2.weight")), put(v[q + "bn2"], g(p + "node_mlp.2.bias"))
            if self.ln:
                put(v[q + "ln_g"], g(p + "layer_norm.weight")), put(v[q + "ln_b"], g(p + "layer_norm.bias"))
            else:
                v[q + "ln_g"].fill_(1.0), v[q + "ln_b"].zero_()
        put(v["coord_w"], g("coord_out.weight")), put(v["lattice_w"], g("lattice_out.weight"))
        if self.ln:
            put(v["fin_g"], g("final_layer_norm.weight")), put(v["fin_b"], g("final_layer_norm.bias"))
        else:
            v["fin_g"].fill_(1.0), v["fin_b"].zero_()
        put(v["type_w"], g("type_out.weight")), put(v["type_b"], g("type_out.bias"))
        self.weights_changed()

    def _reference_named(self, views):
        """Inverse of _load_reference_state: reference-named tensors from native blocks."""
        H = self.hidden_dim
        o = OrderedDict()
        o["node_embedding.weight"], o["node_embedding.bias"] = views["emb_w"].clone(), views["emb_b"].clone()
        o["atom_latent_emb.weight"] = torch.cat([views["lat_w_h"], views["lat_w_t"]], dim=1)
        o["atom_latent_emb.bias"] = views["lat_b"].clone()
        for i in range(self.num_layers):
            p, q = "csp_layer_%d." % i, "l%d." % i
            o[p + "edge_mlp.0.weight"] = torch.cat([views[q + "w_pq"][:H], views[q + "w_pq"][H:], views[q + "w_l"],
                                                    views[q + "w_f"]], dim=1)
            o[p + "edge_mlp.0.bias"] = views[q + "b1"].clone()
            o[p + "edge_mlp.2.weight"], o[p + "edge_mlp.2.bias"] = views[q + "w2"].clone(), views[q + "b2"].clone()
            o[p + "node_mlp.0.weight"], o[p + "node_mlp.0.bias"] = views[q + "wn1"].clone(), views[q + "bn1"].clone()
            o[p + "node_mlp.2.weight"], o[p + "node_mlp.2.bias"] = views[q + "wn2"].clone(), views[q + "bn2"].clone()
            if self.ln:
                o[p + "layer_norm.weight"], o[p + "layer_norm.bias"] = views[q + "ln_g"].clone(), views[q + "ln_b"].clone()
        o["coord_out.weight"], o["lattice_out.weight"] = views["coord_w"].clone(), views["lattice_w"].clone()
        if self.ln:
            o["final_layer_norm.weight"], o["final_layer_norm.bias"] = views["fin_g"].clone(), views["fin_b"].clone()
        o["type_out.weight"], o["type_out.bias"] = views["type_w"].clone(), views["type_b"].clone()
        return o

    def state_dict(self, destination=None, prefix="", keep_vars=False):
        d = destination if destination is not None else OrderedDict()
        for k, t in self._reference_named(self._views).items():
            d[prefix + k] = t
        return d

    def reference_named_grads(self):
        self.flat_grad()
        return self._reference_named(self._gviews)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        try:
            self._load_reference_state(state_dict, prefix)
        except KeyError as e:
            missing_keys.append(str(e))

    # ------------------------------------------------------------------ graphs / workspaces
    MAX_CACHED = 8      # batch topologies (and their workspaces) kept per network, least recently used first out

    def graph_for(self, num_atoms):
        key = tuple(int(v) for v in torch.as_tensor(num_atoms).reshape(-1).tolist())
        g = self._graphs.pop(key, None)
        if g is None:
            if self.edge_style == "fc":
                g = CrystalGraph(key, self.device)
            else:
                from .knn import KnnGraph
                g = KnnGraph(key, self.device, self.max_neighbors)
            while len(self._graphs) >= self.MAX_CACHED:
                self.release(next(iter(self._graphs.values())))
        self._graphs[key] = g           # most recently used last
        return g

    def workspace(self, g, train):
        """Activations for (batch topology, mode).  Keyed by the topology itself (not by the identity of the graph
        object, which can be recycled) and bounded: holders (a sampling run, a fine-tune group) keep their workspace
        alive themselves, the cache only avoids re-allocation between runs on the same batch shape."""
        key = (g.key(), bool(train))
        ws = self._ws.pop(key, None)
        if ws is None or ws.g.N != g.N or getattr(ws.g, "E_cap", ws.g.E) != getattr(g, "E_cap", g.E):
            ws = _Workspace(self, g, train)
            while len(self._ws) >= 2 * self.MAX_CACHED:
                self._ws.pop(next(iter(self._ws)))
        self._ws[key] = ws
        return ws

    def release(self, g):
        """drop a batch topology and its workspaces from the caches (memory is freed once no run holds them)"""
        key = g.key()
        self._graphs.pop(key, None)
        self._ws.pop((key, True), None)
        self._ws.pop((key, False), None)

    # ------------------------------------------------------------------ per-edge GEMMs (the dominant launches)
    def edge_mode(self, E):
        """(presplit, merged): Phi arrives pre-split from mi_edge_fourier when its rows are TMA-legal; the two per-edge
        GEMMs use 128x256 single-accumulator tiles once they fill every SM (fewer operand bytes per flop; ~2.5x the
        rounding error of the two-accumulator 128x128 tiles, both FP32-grade)."""
        H, F = self.hidden_dim, self.num_freqs
        presplit = self.use_tc and (6 * F) % 8 == 0
        # MI_TC_FORCE_MERGED=1: merged tiles at any edge count (parity runs of small goldens on the benchmark's format)
        fill = self.force_merged or os.environ.get("MI_TC_FORCE_MERGED", "0") == "1" or ((E + 127) // 128) * (H // 256) >= ops.sm_count()
        merged = presplit and self.use_merged and H % 256 == 0 and fill
        return presplit, merged

    def pair_mode(self, train, merged):
        """the CTA-pair kernels of csrc/mi_edge.cu serve the inference path with merged tiles; they take the row maxima of
        [P'|Q|R] the LayerNorm'd node path reports (ws.amax_pqr)"""
        return merged and not train and self.use_pair and self.use_tc and self.ln and self.hidden_dim % 256 == 0

    def edge_gemm1(self, i, ws, g, E, a1, train, presplit, merged, pqr=None, amax_pqr=None):
        """a1 = silu(Phi W_F^T + P'[src] + Q[dst])   (first edge linear, cspnet.py:59-72, per-edge part)"""
        H, q = self.hidden_dim, "l%d." % i
        pqr = ws.pqr if pqr is None else pqr
        if self.pair_mode(train, merged):
            a1h, a1l = a1.view(torch.float16).view(2, -1, H)[:, :a1.shape[0]]
            ops.edge_block1(E, ws.phi_hi, ws.phi_lo, self._mhi[q + "w_f"], self._mlo[q + "w_f"], self._minv[q + "w_f"], 2.0 ** -14,
                            pqr[:, :H], pqr[:, H:2 * H], g.edge_src, g.edge_dst, ws.amax_pqr[i] if amax_pqr is None else amax_pqr,
                            self._bounds[i, 3:4], a1h, a1l, ws.amax_a1[i])
            return
        epi1 = dict(gathers=[(pqr[:, :H], g.edge_src), (pqr[:, H:2 * H], g.edge_dst)],
                    z_out=ws.z1[i] if train else None, act=ACT_SILU, amax_out=ws.amax_a1[i])
        if merged:
            ops.tc_gemm_presplit(ws.phi_hi, ws.phi_lo, self._mhi[q + "w_f"], self._mlo[q + "w_f"], a1, M=E,
                                 alpha=2.0 ** -14, col_scale=self._minv[q + "w_f"], flags=ops.TC_MERGED, **epi1)
        elif presplit:
            ops.tc_gemm_presplit(ws.phi_hi, ws.phi_lo, self._hi[q + "w_f"], self._lo[q + "w_f"], a1, M=E, **epi1)
        else:
            self._linear(ws.phi, q + "w_f", a1, E, **epi1)

    def edge_gemm2(self, i, ws, g, E, a1, agg, train, merged):
        """agg = mean_j silu(a1 W_2^T + b_2)   (second edge linear + scatter-mean over the source node, cspnet.py:73-79).
        Merged tiles: the segment means are formed in the GEMM's epilogue (`agg` must be zeroed: the LayerNorm of the
        layer does it) and the [E, H] messages are never written; otherwise GEMM -> a2 -> segment_reduce."""
        N, H, q = g.N, self.hidden_dim, "l%d." % i
        if self.pair_mode(train, merged):
            a1h, a1l = a1.view(torch.float16).view(2, -1, H)[:, :a1.shape[0]]
            ops.edge_block2(E, a1h, a1l, ws.amax_a1[i], self._mhi[q + "w2"], self._mlo[q + "w2"], self._minv[q + "w2"],
                            self._views[q + "b2"], agg, g.edge_src, g.edge_w, ws.amax_agg[i])
            return
        epi2 = dict(bias=self._views[q + "b2"], z_out=ws.z2[i] if train else None, act=ACT_SILU, a_amax=ws.amax_a1[i])
        if merged:
            ops.tc_gemm(a1, self._mhi[q + "w2"], self._mlo[q + "w2"], None, M=E, col_scale=self._minv[q + "w2"],
                        flags=ops.TC_MERGED, scatter=(agg, g.edge_src, g.edge_w, ws.amax_agg[i]), **epi2)
        else:
            self._linear(a1, q + "w2", ws.a2, E, **epi2)
            ops.segment_reduce(ws.a2, g.seg_ptr, agg, N, H, mean=True, amax_out=ws.amax_agg[i])

    # ------------------------------------------------------------------ forward
    def _composed(self, a, train):
        """inference: the two embedding linears run as one tensor-core GEMM over the composed weight (see _refresh_tc)"""
        return (not train and self.use_tc and self.compose_embedding and a.stride(0) % 4 == 0 and a.data_ptr() % 16 == 0)

    def time_term_table(self, ttab, a):
        """[rows of ttab, H]: the per-crystal embedding term `temb W_t^T + bias` (cspnet.py:267-271) of EVERY timestep.  While
        sampling all crystals share the time, so the per-step [B, time_dim] x [time_dim, H] GEMM (12 us of latency per reverse
        step) becomes one GEMM per sampling run and a row gather per step (mi_sampler_step_begin on this table).  Same kernel,
        same row arithmetic as the per-step call; `a` is the atom-type state the forward will see (it decides the bias)."""
        if self.use_tc:
            self._refresh_tc()
        bias = self._emb_bias if self._composed(a, False) else self._views["lat_b"]
        out = torch.empty(ttab.shape[0], self.hidden_dim, device=ttab.device, dtype=torch.float32)
        self._linear(ttab.contiguous(), "lat_w_t", out, ttab.shape[0], bias=bias)
        return out

    def forward_graph(self, g, temb, a, x, l, train=False, heads=(True, True, True), ws=None, reuse_embedding=False,
                      tb_ready=False):
        """Score network on a prebuilt graph.  temb [B,T], a [N,A], x [N,3], l [B,3,3] fp32 CUDA.
        heads = which of (lattice, coord, type) outputs to compute.  Returns views into the workspace.
        reuse_embedding: temb, a and l are those of the previous call on this workspace (only x moved, as between the
        corrector and the predictor of one reverse step): the embedding GEMMs and the per-crystal lattice terms of
        that call are kept.
        tb_ready: the caller has already put the per-crystal time term into ws.tb (rows of time_term_table); temb is unused."""
        W, H, F = self._views, self.hidden_dim, self.num_freqs
        N, E, B, L = g.N, g.E, g.B, self.num_layers
        ws = ws or self.workspace(g, train)
        if self.edge_style != "fc":
            g.rebuild(x, l, need_dst=train)
            E = g.E
        if self.use_tc:
            self._refresh_tc()
        # embedding (cspnet.py:264-271):  h = [Lin_A(a) | temb_b] W^T + b
        ws.amax.zero_()
        reuse = reuse_embedding and not train
        # (the Fourier basis is compute-bound — 26 M sincosf per evaluation fill every SM — so running it on a side stream
        # next to the embedding GEMMs measured no gain: 98.4 vs 98.7 crystals/s)
        presplit, merged = self.edge_mode(E)
        ops.edge_fourier(x, g.edge_src, g.edge_dst, g.cell_off, E, F, None, ws.phi if (train or not presplit) else None,
                         ws.phi_hi if presplit else None, ws.phi_lo if presplit else None,
                         op_scale=2.0 ** 14 if merged else 1.0, lo_scale=1.0 if merged else 2048.0)
        composed = self._composed(a, train)
        if not reuse and composed:
            # inference: the two embedding linears as one tensor-core GEMM over the composed weight (see _refresh_tc); the
            # atom-type state has no producer that reports row maxima: one tiny kernel takes them
            ops.row_amax(a, N, self.max_atoms, ws.amax_h0)
            if not tb_ready:
                self._linear(temb, "lat_w_t", ws.tb, B, bias=self._emb_bias)
            ops.tc_gemm(a, self._emb_hl[0], self._emb_hl[1], ws.h[0], M=N, N=H, K=self.max_atoms,
                        gathers=[(ws.tb, g.node_graph)], a_amax=ws.amax_h0)
            ops.lattice_ip(l, ws.ips, B)
        elif not reuse:
            # the atom-type state is unbounded (no producer reports its row maxima) and K = 100: FP32 CUDA-core GEMM
            ops.sgemm(a, W["emb_w"], ws.h0, M=N, bias=W["emb_b"], amax_out=ws.amax_h0)
            if not tb_ready:
                self._linear(temb, "lat_w_t", ws.tb, B, bias=W["lat_b"])
            self._linear(ws.h0, "lat_w_h", ws.h[0], N, gathers=[(ws.tb, g.node_graph)], a_amax=ws.amax_h0)
            ops.lattice_ip(l, ws.ips, B)
        # per-crystal term C_b of the first edge linear, all layers in one launch (the layer blocks of the flat weight
        # buffer are equally spaced)
        if not reuse:
            lstride = (self._slices["l1.w_l"][0] - self._slices["l0.w_l"][0]) if L > 1 else 0
            ops.lattice_linear(l, W["l0.w_l"], W["l0.b1"], ws.cb[0, :, :H], B, H, n_sets=L, w_stride=lstride,
                               bias_stride=lstride, out_stride=ws.cb.stride(0))
        # node path: LayerNorm emits the pre-split operand of ONE per-node GEMM that produces P', Q and R = LN(h) W_a^T
        # (the LN(h) half of node_mlp.0, whose weight is [W_a | W_b] over [LN(h) | agg]); node_mlp.0 then runs on agg alone
        # (K = H instead of 2H) and adds R in its epilogue.  Needs the tensor-core path and LayerNorm.
        node_path = self.use_tc and self.ln and H % 8 == 0 and H <= 1024 and ops.tc_ok(ws.cat[0][:, H:], self._hi["l0.wn1"][:, H:])
        # inference: the node-level chain of every layer boundary (node_mlp.0 -> node_mlp.2 + residual -> next layer's
        # LayerNorm -> next P|Q|R GEMM) is ONE cluster launch (csrc/mi_node.cu) instead of four latency-bound ones
        chain = node_path and not train and H == 512 and self.use_chain and an1_contig(ws)
        for i in range(L):
            q = "l%d." % i
            k = i if train else 0
            cat, a1, an1 = ws.cat[k], ws.a1[k], ws.an1[k]
            h_in, h_out = ws.h[i], ws.h[i + 1]
            hn, agg = cat[:, :H], cat[:, H:]
            if chain:
                pqr_i, amax_i = (ws.pqr0, ws.amax_pqr0) if i == 0 else (ws.pqr, ws.amax_pqr[i])
                if i == 0 and not reuse:
                    # layer 0's LayerNorm and [P'|Q|R] see the embedding output only (atom types, time, lattice): the
                    # predictor evaluation of a reverse step reuses the corrector's
                    ws.amax_pqr0.zero_()
                    ops.layernorm_fwd_split(h_in, W[q + "ln_g"], W[q + "ln_b"], None, ws.hn_hi, ws.hn_lo, ws.amax_hn[i], N, H,
                                            zero_out=agg if merged else None, zero_cols=H)
                    ops.tc_gemm_presplit(ws.hn_hi, ws.hn_lo, self._pqr_hi[i], self._pqr_lo[i], ws.pqr0, M=N,
                                         gathers=[(ws.cb[i], g.node_graph)], a_amax=ws.amax_hn[i], amax_out=ws.amax_pqr0)
                self.edge_gemm1(i, ws, g, E, a1, train, presplit, merged, pqr=pqr_i, amax_pqr=amax_i)
                self.edge_gemm2(i, ws, g, E, a1, agg, train, merged)
                nxt = None
                if i + 1 < L:
                    qn = "l%d." % (i + 1)
                    nxt = (W[qn + "ln_g"], W[qn + "ln_b"], 1e-5, self._pqr_hi[i + 1], self._pqr_lo[i + 1], ws.cb[i + 1],
                           g.node_graph, ws.pqr, ws.amax_pqr[i + 1])
                # (agg is zeroed after every layer, the last one included: it is the zeroed destination of the next
                # evaluation's first fused scatter-mean, whose LayerNorm launch is skipped when the embedding is reused)
                ops.node_chain(N, H, agg, ws.amax_agg[i], merged, ws.xs, ws.ys, self._hi[q + "wn1"][:, H:],
                               self._lo[q + "wn1"][:, H:], W[q + "bn1"], pqr_i[:, 2 * H:], amax_i, self._bounds[i],
                               self._hi[q + "wn2"], self._lo[q + "wn2"], W[q + "bn2"], h_in, h_out, ln=nxt)
                continue
            # edge model (cspnet.py:59-75) with the first linear split into per-node / per-crystal / per-edge parts; the
            # per-crystal term C_b is folded into P (P'_i = P_i + C_b(i)) by the per-node GEMM's epilogue: the per-edge
            # GEMM then adds two gathered rows instead of three
            if node_path:
                ops.layernorm_fwd_split(h_in, W[q + "ln_g"], W[q + "ln_b"], hn if train else None, ws.hn_hi, ws.hn_lo,
                                        ws.amax_hn[i], N, H, zero_out=agg if merged else None, zero_cols=H,
                                        mean=ws.ln_mean[i] if train else None, rstd=ws.ln_rstd[i] if train else None)
                ops.tc_gemm_presplit(ws.hn_hi, ws.hn_lo, self._pqr_hi[i], self._pqr_lo[i], ws.pqr, M=N,
                                     gathers=[(ws.cb[i], g.node_graph)], a_amax=ws.amax_hn[i], amax_out=ws.amax_pqr[i])
            else:
                if self.ln:
                    # the row maximum of cat = [LN(h) | agg] is accumulated by both producers (LayerNorm here, the scatter below)
                    ops.layernorm_fwd(h_in, W[q + "ln_g"], W[q + "ln_b"], hn, N, H,
                                      ws.ln_mean[i] if train else None, ws.ln_rstd[i] if train else None,
                                      amax_out=ws.amax_agg[i])
                else:
                    hn.copy_(h_in)
                    torch.maximum(ws.amax_agg[i], h_in.abs().amax(dim=1), out=ws.amax_agg[i])
                if merged:
                    agg.zero_()
                # (amax_agg[i] holds the row maxima of LN(h) only at this point: the scatter adds its own further down)
                self._linear(hn, q + "w_pq", ws.pq, N, gathers=[(ws.cb2[i], g.node_graph)], a_amax=ws.amax_agg[i])
            self.edge_gemm1(i, ws, g, E, a1, train, presplit, merged)
            # second edge linear + scatter-mean over the source node (cspnet.py:73-79)
            self.edge_gemm2(i, ws, g, E, a1, agg, train, merged)
            # node model + residual (cspnet.py:77-91)
            if node_path:
                self.tc_linear_view(agg, q + "wn1", slice(H, 2 * H), an1, N, bias=W[q + "bn1"], gathers=[(ws.pqr[:, 2 * H:], None)],
                                    z_out=ws.zn1[i] if train else None, act=ACT_SILU, a_amax=ws.amax_agg[i],
                                    amax_out=ws.amax_an1[i])
            else:
                self._linear(cat, q + "wn1", an1, N, bias=W[q + "bn1"], z_out=ws.zn1[i] if train else None, act=ACT_SILU,
                             a_amax=ws.amax_agg[i], amax_out=ws.amax_an1[i])
            self._linear(an1, q + "wn2", h_out, N, bias=W[q + "bn2"], z_out=ws.zn2[i] if train else None,
                         act=ACT_SILU, resid=h_in, a_amax=ws.amax_an1[i])
        hL = ws.h[L]
        if not train and self.fused_heads and H in (128, 256, 512, 1024):
            # inference: final LayerNorm + the three heads in one launch, one CTA per crystal (cspnet.py:276-294)
            ops.output_heads(hL, g.node_off, B, H, W["fin_g"] if self.ln else None, W["fin_b"] if self.ln else None,
                             W["coord_w"], ws.pred_x if heads[1] else None, W["type_w"], W["type_b"],
                             ws.pred_a if heads[2] else None, W["lattice_w"], l, self.ip, ws.pred_l if heads[0] else None)
            return ws.pred_l, ws.pred_x, ws.pred_a
        hL = ws.h[L]
        if self.ln:
            ops.layernorm_fwd(hL, W["fin_g"], W["fin_b"], ws.hf, N, H,
                              ws.ln_mean[L] if train else None, ws.ln_rstd[L] if train else None)
            hf = ws.hf
        else:
            hf = hL
        ws.hf_used = hf
        if heads[1]:
            self._linear(hf, "coord_w", ws.pred_x, N)
        if heads[0]:
            ops.segment_reduce(hf, g.node_off, ws.gmean, B, H, mean=True)
            if self.ip:
                self._linear(ws.gmean, "lattice_w", ws.lat9, B)
                ops.bmm3(ws.lat9, l, ws.pred_l, B)
            else:
                self._linear(ws.gmean, "lattice_w", ws.pred_l.view(B, 9), B)
        if heads[2]:
            self._linear(hf, "type_w", ws.pred_a, N, bias=W["type_b"])
        return ws.pred_l, ws.pred_x, ws.pred_a

    # ------------------------------------------------------------------ backward
    @staticmethod
    def _splitk(M, N, K):
        tiles = ((M + 127) // 128) * ((N + 127) // 128)
        want = max(1, (2 * 148) // tiles)
        return max(2, min(want, (K + 255) // 256, 64))

    # reduction rows from which a weight gradient goes to the tensor cores (below, the FP32 split-K GEMM wins: the
    # three transposes cost more than they save); MI_WGRAD_TC_ROWS overrides it for tuning runs
    WGRAD_TC_ROWS = int(os.environ.get("MI_WGRAD_TC_ROWS", "20000"))

    def _wgrad(self, dY, X, gname, M, N, K, ws=None, xt=None):
        """grad[gname] [M,N] += dY[K,M]^T @ X[K,N]   (sum over K rows: nodes / edges / crystals).
        Tensor-core path for long reductions: dY^T (fp32, row maxima = column maxima of dY) and X^T (merged-format
        fp16 split, one scale per column of X; `xt` = an already transposed X) feed mi_tc_gemm with K cut into parts
        that are added to the gradient with 16-byte reductions."""
        if (ws is not None and self.use_tc and self.use_merged and K >= self.WGRAD_TC_ROWS and M % 128 == 0 and
                N % 256 == 0 and M <= 2 * self.hidden_dim):
            sc = ws.wgrad_scratch(K)
            sc["amax"].zero_()
            ca = sc["amax"][:M]
            ops.transpose_amax(dY, K, XT=sc["dyT"], col_amax=ca)
            if xt is None:
                xt = self._transpose_split(X, K, N, sc, sc["xhi"], sc["xlo"], sc["inv"])
            hi, lo, inv = xt
            tiles = (M // 128) * (N // 256)
            ks = max(1, min(ops.sm_count() // tiles, (K + 32 * 8 - 1) // (32 * 8)))
            ops.tc_gemm(sc["dyT"][:M], hi[:N], lo[:N], self._gviews[gname], M=M, N=N, K=K, a_amax=ca, col_scale=inv[:N],
                        flags=ops.TC_MERGED, splitk=ks)
            return
        ops.sgemm(dY, X, self._gviews[gname], transA=True, transB=False, M=M, N=N, K=K, beta=1.0,
                  splitk=self._splitk(M, N, K))

    def _transpose_split(self, X, K, N, sc, hi, lo, inv):
        cx = sc["amax"][2 * self.hidden_dim:2 * self.hidden_dim + N]
        ops.transpose_amax(X, K, XT=None, col_amax=cx)
        ops.transpose_split(X, K, cx, hi, lo, inv)
        return hi, lo, inv

    def backward_graph(self, g, temb, a, x, l, d_l, d_x, d_a, ws=None):
        """Accumulate d(objective)/d(weights) into the flat gradient buffer given the gradients w.r.t. the
        three outputs of the matching `forward_graph(..., train=True)` call (activations in `ws`)."""
        self.flat_grad()
        W, G, H = self._views, self._gviews, self.hidden_dim
        N, E, B, L, A = g.N, g.E, g.B, self.num_layers, MAX_ATOMIC_NUM
        F6, T = 6 * self.num_freqs, self.latent_dim
        ws = ws or self.workspace(g, True)
        hf = ws.hf if self.ln else ws.h[L]
        if self.use_tc and not self._need_T:
            self._need_T = True
            self._tc_version = None
            self._refresh_tc()
        ws.gamax.zero_()
        # transposed Fourier basis for the six dW_F GEMMs, once (long reductions go to the tensor cores: _wgrad)
        phi_t = None
        if self.use_tc and self.use_merged and E >= self.WGRAD_TC_ROWS and H % 128 == 0 and F6 % 256 == 0:
            sc = ws.wgrad_scratch(max(E, N))
            sc["amax"].zero_()
            phi_t = self._transpose_split(ws.phi, E, F6, sc, sc["phi_hi"], sc["phi_lo"], sc["phi_inv"])
        # ---- heads (cspnet.py:276-294)
        if self.ip:
            ops.bmm3(d_l, l, ws.dlat9, B, transL=True)
            dlat9 = ws.dlat9
        else:
            dlat9 = d_l.view(B, 9)
        self._wgrad(dlat9, ws.gmean, "lattice_w", 9, H, B)
        ops.sgemm(dlat9, W["lattice_w"], ws.dg, transB=False, M=B, N=H, K=9)
        ops.gather_rows_dsilu(ws.dg, g.node_graph, g.node_off, None, ws.dgn, N, H)
        self._wgrad(d_x, hf, "coord_w", 3, H, N)
        self._wgrad(d_a, hf, "type_w", A, H, N)
        ops.colsum(d_a, N, A, G["type_b"])
        ops.sgemm(d_a, W["type_w"], ws.dhf, transB=False, M=N, N=H, K=A, gathers=[(ws.dgn, None)])
        ops.sgemm(d_x, W["coord_w"], ws.dhf, transB=False, M=N, N=H, K=3, beta=1.0)
        if self.ln:
            ops.layernorm_bwd(ws.dhf, ws.h[L], W["fin_g"], ws.ln_mean[L], ws.ln_rstd[L], ws.dh, G["fin_g"], G["fin_b"],
                              N, H)
            dh = ws.dh
        else:
            ws.dh.copy_(ws.dhf)
            dh = ws.dh
        # ---- layers, reversed
        for i in reversed(range(L)):
            q = "l%d." % i
            cat, a1, an1 = ws.cat[i], ws.a1[i], ws.an1[i]
            # h_out = h_in + silu(zn2);  zn2 = an1 wn2^T + bn2;  an1 = silu(zn1);  zn1 = cat wn1^T + bn1
            ops.gather_rows_dsilu(dh, None, None, ws.zn2[i], ws.dzn, N, H, amax_out=ws.amax_dzn[i])
            self._wgrad(ws.dzn, an1, q + "wn2", H, H, N, ws=ws)
            ops.colsum(ws.dzn, N, H, G[q + "bn2"])
            self._dgrad(ws.dzn, q + "wn2", ws.dzn1, N, ws.amax_dzn[i], act=ACT_DSILU, z_in=ws.zn1[i],
                        amax_out=ws.amax_dzn1[i])
            self._wgrad(ws.dzn1, cat, q + "wn1", H, 2 * H, N, ws=ws)
            ops.colsum(ws.dzn1, N, H, G[q + "bn1"])
            self._dgrad(ws.dzn1, q + "wn1", ws.dcat, N, ws.amax_dzn1[i])
            # agg = mean_j a2 ; a2 = silu(z2) ; z2 = a1 w2^T + b2
            ops.gather_rows_dsilu(ws.dcat[:, H:], g.edge_src, g.seg_ptr, ws.z2[i], ws.dz2, E, H, amax_out=ws.amax_dz2[i])
            self._wgrad(ws.dz2, a1, q + "w2", H, H, E, ws=ws)
            ops.colsum(ws.dz2, E, H, G[q + "b2"])
            # a1 = silu(z1) ; z1 = Phi w_f^T + P[src] + Q[dst] + C[graph]
            self._dgrad(ws.dz2, q + "w2", ws.dz1, E, ws.amax_dz2[i], act=ACT_DSILU, z_in=ws.z1[i])
            self._wgrad(ws.dz1, ws.phi, q + "w_f", H, F6, E, ws=ws, xt=phi_t)
            ops.segment_reduce(ws.dz1, g.seg_ptr, ws.dpq[:, :H], N, H, mean=False, amax_out=ws.amax_dpq[i])
            ops.segment_reduce(ws.dz1, g.dst_ptr, ws.dpq[:, H:], N, H, perm=g.dst_perm, mean=False,
                               amax_out=ws.amax_dpq[i])
            ops.segment_reduce(ws.dpq[:, :H], g.node_off, ws.dcb, B, H, mean=False)
            ops.colsum(ws.dcb, B, H, G[q + "b1"])
            self._wgrad(ws.dcb, ws.ips, q + "w_l", H, 9, B)
            self._wgrad(ws.dpq, cat[:, :H], q + "w_pq", 2 * H, H, N, ws=ws)
            # d hn = dcat[:, :H] + dpq @ w_pq   (written over dcat[:, :H])
            self._dgrad(ws.dpq, q + "w_pq", ws.dcat[:, :H], N, ws.amax_dpq[i], accumulate=True)
            if self.ln:
                ops.layernorm_bwd(ws.dcat[:, :H], ws.h[i], W[q + "ln_g"], ws.ln_mean[i], ws.ln_rstd[i], dh,
                                  G[q + "ln_g"], G[q + "ln_b"], N, H, accumulate_dx=True)
            else:
                dh.add_(ws.dcat[:, :H])
        # ---- embedding
        self._wgrad(dh, ws.h0, "lat_w_h", H, H, N)
        ops.segment_reduce(dh, g.node_off, ws.dtb, B, H, mean=False)
        ops.colsum(ws.dtb, B, H, G["lat_b"])
        self._wgrad(ws.dtb, temb, "lat_w_t", H, T, B)
        ops.sgemm(dh, W["lat_w_h"], ws.dzn, transB=False, M=N, N=H, K=H)
        self._wgrad(ws.dzn, a, "emb_w", H, self.max_atoms, N)
        ops.colsum(ws.dzn, N, H, G["emb_b"])

    # ------------------------------------------------------------------ reference-facing forward
    def forward(self, t, atom_types, frac_coords, lattices, num_atoms, node2graph):
        """models/diffcsp/cspnet.py:260-294.  With autograd enabled and trainable weights the call is
        differentiable w.r.t. the weights (torch.autograd.Function over the hand-written backward)."""
        g = self.graph_for(num_atoms)
        temb = t.to(torch.float32).contiguous()
        a = atom_types.to(torch.float32).contiguous()
        x = frac_coords.to(torch.float32).contiguous()
        l = lattices.to(torch.float32).contiguous()
        if torch.is_grad_enabled() and self.flat.requires_grad:
            return _CSPNetFunction.apply(self.flat, self, g, temb, a, x, l)
        pl, px, pa = self.forward_graph(g, temb, a, x, l, train=False)
        return pl.clone(), px.clone(), pa.clone()


class _CSPNetFunction(torch.autograd.Function):
    """Autograd bridge for the plugin path (pipeline/mat_invent.py:152-164 calls `.backward()`)."""

    @staticmethod
    def forward(ctx, flat, net, g, temb, a, x, l):
        pl, px, pa = net.forward_graph(g, temb, a, x, l, train=True)
        ctx.net, ctx.g = net, g
        ctx.save_for_backward(temb, a, x, l)
        return pl.clone(), px.clone(), pa.clone()

    @staticmethod
    def backward(ctx, d_l, d_x, d_a):
        net, g = ctx.net, ctx.g
        temb, a, x, l = ctx.saved_tensors
        B, N = g.B, g.N

        def z(t, shape):
            return torch.zeros(shape, device=net.device) if t is None else t.contiguous()

        gbuf = net.flat_grad()
        before = gbuf.clone()
        net.backward_graph(g, temb, a, x, l, z(d_l, (B, 3, 3)), z(d_x, (N, 3)), z(d_a, (N, MAX_ATOMIC_NUM)))
        delta = gbuf - before          # autograd accumulates the returned gradient into flat.grad itself
        gbuf.copy_(before)
        return delta, None, None, None, None, None, None
