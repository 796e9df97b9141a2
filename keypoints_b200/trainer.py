"""Fused training step for KeyNet / TransporterNet / AutoEncoder (reference inner loops: keypoints.py:70-84,
transporter.py:75-89, autoencode.py:84-96): augment -> forward -> L2 loss -> backward -> gradient all-reduce -> Adam,
run as one chain of C-ABI kernel launches on static NHWC buffers and replayed as ONE CUDA graph.

Data parallel: one process per GPU; each rank trains on its own shard of the batch with
local BatchNorm statistics; parameters, Adam state and BatchNorm buffers are broadcast from rank 0 at construction.
The only exchange is a sum all-reduce of the flat fp32 gradient bucket, issued bucket by bucket AS backward finishes
them (decoder | deep layers of keypoint and encoder | their shallow layers) on NCCL's stream, so the transfer runs under
the remaining backward kernels; the collectives are captured in the same graph as the kernels and Adam (1/world
scaling folded into Adam).
"""
from __future__ import annotations

import math
import os
from typing import Optional

import torch

from . import engine, lib as L, parallel
from .engine import CachedAlloc, LayerGrads, LayerParams
from .models import knn
from .models.autoencoder import AutoEncoder
from .models.keynet import KeyNet
from .models.transporter import TransporterNet


DEFER_FOLD = os.environ.get('KP_DEFER_FOLD', '1') != '0'


class _UnitState:
    """Specs, parameter / gradient views into the flat buckets and static buffers of one Unit."""

    def __init__(self, name, unit: knn.Unit):
        self.name = name
        self.unit = unit
        self.specs, self.mods = unit.layers()
        self.alloc = CachedAlloc(name)
        self.params = None
        self.grads = None
        self.packs = None
        self.ctxs = None
        self.span = (0, 0)

    def tensors(self):
        return knn.trainable(self.mods)


class Trainer:
    def __init__(self, net, precision: str = 'bf16', lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 augment: Optional[dict] = None, use_graph: bool = True, process_group=None, device=None,
                 seed: int = 0, overlap_allreduce: bool = True):
        if not torch.cuda.is_available():
            raise RuntimeError('keypoints_b200.Trainer needs a CUDA (sm_100a) device; there is no CPU path')
        self.device = torch.device(device if device is not None else f'cuda:{torch.cuda.current_device()}')
        if self.device.index is None:
            self.device = torch.device('cuda', torch.cuda.current_device())
        with torch.cuda.device(self.device):
            L.device_info()                   # fails loudly on a non-sm_100 device
            self.ctx = L.Context()            # SM budget + tensor-map cache of this trainer's device
        self.net = net.to(self.device)
        if isinstance(net, KeyNet):
            self.kind = 'keynet'
        elif isinstance(net, TransporterNet):
            self.kind = 'transporter'
        elif isinstance(net, AutoEncoder):
            self.kind = 'autoencoder'
        else:
            raise TypeError('Trainer supports KeyNet, TransporterNet and AutoEncoder')
        self.combine = 'max'
        if self.kind == 'transporter':
            self.combine = net.combine_method
            if self.combine not in L.COMBINE:
                raise NotImplementedError(f"combine_method {self.combine!r}: 'pretrained_network' needs a MaskMaker network "
                                          "that the reference's make() never builds (models/transporter.py:112-128)")
        self.precision = precision
        self.T = engine.act_dtype(precision)
        self._lr, self._betas, self._eps = float(lr), tuple(betas), float(eps)
        self.sigma = float(getattr(getattr(net, 'key2map', None), 'sigma', 0.1))
        self.augment = augment               # dict(cntl_pts=4, variance=0.05, max_rotate=0.1) or None
        self.use_graph = use_graph
        self.overlap_allreduce = overlap_allreduce and os.environ.get('KP_AR_OVERLAP', '1') != '0'
        self.pg = process_group
        self.world = 1
        if process_group is False:           # single-process trainer inside a distributed job (checks, rank-0-only legs)
            self.pg = None
        elif process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        self.rank = torch.distributed.get_rank(self.pg) if self.world > 1 else 0
        # gradient exchange: 'p2p' = reduce-scatter + Adam + all-gather fused in one kernel over NVLink peer memory
        # (kp_dp_adam_step); 'nccl' = bucket-wise NCCL all-reduce overlapped with backward + replicated Adam
        self.dp_mode = os.environ.get('KP_DP', 'p2p') if self.world > 1 else 'none'
        self.peer = None
        # augmentation stream: base seed + rank, so the shards of a data-parallel job draw different perturbations
        self.aug_seed = (int(seed) * 1000003 + self.rank) & 0xFFFFFFFFFFFFFFFF
        first = net.feature if self.kind == 'transporter' else net.encoder
        # bucket order = order in which backward finishes units
        self.units = {'decoder': _UnitState('decoder', net.decoder)}
        if self.kind != 'autoencoder':
            self.units['keypoint'] = _UnitState('keypoint', net.keypoint)
        self.units['encoder'] = _UnitState('encoder', first)
        with torch.cuda.device(self.device):
            self._flatten()
            self.misc = CachedAlloc('misc')
            self.side = torch.cuda.Stream(device=self.device)     # encoder branch runs beside the keypoint branch
            self.step_dev = torch.zeros(1, dtype=torch.int32, device=self.device)
            self.loss_sum = torch.zeros(1, dtype=torch.float64, device=self.device)
            self.loss_ring = None             # see LossRing / attach_loss_ring
        self.two_streams = os.environ.get('KP_TWO_STREAMS', '1') != '0'
        # weight-gradient kernels on their own stream per unit, off the dgrad -> BatchNorm-backward chain the next layer waits
        # for: their tails fill with the other kernels (measured A/B/A/B: 14.29 -> 14.16 ms/step, 256x256: 26.40 -> 26.11)
        self.wgrad_streams = os.environ.get('KP_WGRAD_STREAM', '1') != '0'
        self._wg_side = {}
        self.graph = None
        self.graph_key = None
        self.steps_done = 0
        self.numel = 1
        self._stg, self._fold_table, self._fold_rows = {}, None, {}
        self._works = []
        if self.world > 1:
            self.broadcast_state()

    # hyper-parameters are baked into the captured Adam launch: changing one drops the graph (re-captured on the next step)
    def _set_hyper(self, name, value):
        if getattr(self, name) != value:
            setattr(self, name, value)
            self.graph = None

    lr = property(lambda self: self._lr, lambda self, v: self._set_hyper('_lr', float(v)))
    betas = property(lambda self: self._betas, lambda self, v: self._set_hyper('_betas', tuple(v)))
    eps = property(lambda self: self._eps, lambda self, v: self._set_hyper('_eps', float(v)))

    def broadcast_state(self, src: int = 0):
        """Make every replica identical to rank `src`: parameters, Adam moments and step, BatchNorm buffers (called at
        construction and after `load` in a data-parallel job; the gradient all-reduce alone only keeps replicas identical
        if they START identical)."""
        if self.world == 1:
            return
        import torch.distributed as dist
        gsrc = dist.get_global_rank(self.pg, src) if self.pg is not None else src
        for t in (self.flat_p, self.flat_m, self.flat_v, self.step_dev):
            dist.broadcast(t, gsrc, group=self.pg)
        for _, bn in self._all_mods():
            if bn is not None:
                for b in (bn.running_mean, bn.running_var, bn.num_batches_tracked):
                    dist.broadcast(b, gsrc, group=self.pg)
        self.steps_done = int(self.step_dev.item())

    # ------------------------------------------------------------------------------------------
    def _flatten(self):
        """One flat fp32 bucket for all trainable tensors (decoder | keypoint | encoder); every tensor starts on a
        128-byte boundary (the kernels use 16-byte vector loads on biases / BN vectors), padding stays zero."""
        ALIGN = 32
        tensors, offsets, off = [], [], 0
        for u in self.units.values():
            start = off
            u.layer_off = []                  # flat offset where each layer's tensors start
            for conv, bn in u.mods:
                u.layer_off.append(off)
                for t in knn.trainable([(conv, bn)]):
                    tensors.append(t)
                    offsets.append(off)
                    off += (t.numel() + ALIGN - 1) // ALIGN * ALIGN
            u.span = (start, off)
        n = off
        if self.dp_mode == 'p2p':
            try:
                self.peer = parallel.PeerBuckets(n, self.device, self.pg)
            except Exception as e:                # no peer mapping between these GPUs: NCCL carries the gradients instead
                import warnings
                warnings.warn(f'keypoints_b200: peer-memory data parallelism unavailable ({type(e).__name__}: {e}); using NCCL')
                self.dp_mode = 'nccl'
            if self.world > 1:                    # all replicas must agree on the mode
                ok = torch.tensor([1 if self.dp_mode == 'p2p' else 0], device=self.device)
                torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN, group=self.pg)
                if int(ok.item()) == 0:
                    self.dp_mode, self.peer = 'nccl', None
        if self.peer is not None:
            self.flat_p, self.flat_g = self.peer.p, self.peer.g
        else:
            self.flat_p = torch.zeros(n, dtype=torch.float32, device=self.device)
            self.flat_g = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.flat_m = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.flat_v = torch.zeros(n, dtype=torch.float32, device=self.device)
        gviews = {}
        for t, o in zip(tensors, offsets):
            k = t.numel()
            self.flat_p[o:o + k].copy_(t.data.reshape(-1))
            t.data = self.flat_p[o:o + k].view_as(t)          # the module now aliases the flat bucket
            gviews[id(t)] = self.flat_g[o:o + k].view_as(t)
        self.n_true_params = sum(t.numel() for t in tensors)
        for u in self.units.values():
            u.params = knn.layer_params(u.mods)
            u.grads = []
            for conv, bn in u.mods:
                g = LayerGrads(dw=gviews[id(conv.weight)], db=None if conv.bias is None else gviews[id(conv.bias)])
                if bn is not None:
                    g.dgamma, g.dbeta = gviews[id(bn.weight)], gviews[id(bn.bias)]
                u.grads.append(g)
        self.n_params = n

    # ------------------------------------------------------------------------------------------
    def _pack(self, shapes):
        """Repack every layer's master weights for the kernels: one multi-tensor launch over a device-resident table."""
        key = tuple((n, tuple(v)) for n, v in sorted(shapes.items()))
        if getattr(self, '_pack_key', None) != key:
            rows = []
            for u in self.units.values():
                u.packs = [engine.pack_layer(s, p, cp, self.precision, u.alloc, f'pk{i}', launch=False)
                           for i, (s, p, cp) in enumerate(zip(u.specs, u.params, shapes[u.name]))]
                rows += [pk['desc'] for pk in u.packs]
            self._pack_table = torch.tensor(rows, dtype=torch.int64).to(self.device)
            self._pack_key = key
        L.call('kp_pack_weights_multi', L.stream(), L.ptr(self._pack_table), self._pack_table.shape[0])

    def _pitches(self, u: _UnitState, cin_pitch):
        out, cp = [], cin_pitch
        for s in u.specs:
            out.append(cp)
            cp = engine.pitch(s.cout, self.precision)
        return out

    def _fwd_unit(self, u: _UnitState, x_pad, H, W, out, out_pad):
        # Co-resident BatchNorm hint (small-footprint forward passes that fit beside the other stream's persistent conv CTAs):
        # off by default since the fragment epilogue - the 256-wide conv tiles now hold 162 registers per thread, so a
        # 288-thread BatchNorm CTA no longer fits next to them, and with or without the hint the step measures the same
        # (4503 vs 4501 pairs/s, 4 interleaved runs each, gpurun_out/r3_b_f3c{1,0}_*.log).  KP_BN_CORESIDENT=1 turns it on.
        co = (self.two_streams and self.kind == 'keynet' and u.name != 'decoder'
              and os.environ.get('KP_BN_CORESIDENT', '0') == '1')
        u.ctxs = engine.unit_forward(u.specs, u.params, x_pad, H, W, self.precision, out, out_pad, alloc=u.alloc,
                                     training=True, packs=u.packs, tag='f', coresident=co)

    def _staging(self, shapes):
        """One fp32 arena for the tensor-core weight-gradient staging of every layer (zeroed once per step) and the
        device table kp_wgrad_finalize_multi folds it with."""
        key = tuple((n, tuple(v)) for n, v in sorted(shapes.items()))
        if getattr(self, '_stg_key', None) == key:
            return
        self._stg_key, self._stg, self._fold_table, self._fold_rows = key, {}, None, {}
        if self.precision != 'bf16' or not DEFER_FOLD:
            return
        sizes, off = [], 0
        for u in self.units.values():
            for i, (sp, cp) in enumerate(zip(u.specs, shapes[u.name])):
                if engine.uses_tc(sp, cp, self.precision):
                    n = sp.k * sp.k * sp.cout * cp
                    sizes.append((u, i, sp, cp, off, n))
                    off += (n + 63) // 64 * 64
        if not sizes:
            return
        self._stg_arena = torch.zeros(off, dtype=torch.float32, device=self.device)
        rows = []
        for u, i, sp, cp, o, n in sizes:
            self._stg.setdefault(u.name, [None] * len(u.specs))[i] = self._stg_arena[o:o + n]
            self._fold_rows[(u.name, i)] = len(rows)
            rows.append([self._stg_arena[o:o + n].data_ptr(), u.grads[i].dw.data_ptr(), sp.k * sp.k, sp.cout, sp.cin, cp,
                         1 if sp.cout % 128 == 0 else 0])
        self._fold_table = torch.tensor(rows, dtype=torch.int64).to(self.device)

    def _fork(self):
        """Run the following block on the side stream, ordered after everything issued so far (captured as a parallel
        branch of the CUDA graph): the memory-bound BatchNorm passes of one Unit overlap the tensor-core convolutions
        of the other."""
        if not self.two_streams:
            import contextlib
            return contextlib.nullcontext()
        self.side.wait_stream(torch.cuda.current_stream())
        return torch.cuda.stream(self.side)

    def _join(self):
        if self.two_streams:
            torch.cuda.current_stream().wait_stream(self.side)

    def _bottleneck_dims(self, H, W):
        h, w = H, W
        for s in self.units['encoder'].specs:
            h, w = engine.post_dims(s.post, h, w)
        return h, w

    # ------------------------------------------------------------------------------------------
    def _augment(self, x):
        """TpsAndRotate (data_augments.py:27-38) with parameters drawn on the device."""
        a = self.augment
        n, c, H, W = x.shape
        dev = x.device
        T = int(a.get('cntl_pts', 4))
        # both perturbations' parameters in one Philox launch keyed by (seed + rank, device step counter): a replayed graph
        # draws fresh parameters every step, and no ATen RNG kernel runs inside the step
        theta = self.misc('aug.theta', (2, n, T + 3, 2), torch.float32, dev)
        ctrl = self.misc('aug.ctrl', (2, n, T, 2), torch.float32, dev)
        rot = self.misc('aug.rot', (2, n), torch.float32, dev)
        L.call('kp_aug_draw', L.stream(), self.aug_seed, L.ptr(self.step_dev), 2, n, T, float(a.get('variance', 0.05)),
               float(a.get('max_rotate', 0.1)), L.ptr(theta), L.ptr(ctrl), L.ptr(rot))

        def perturb(src, dst, tmp, d):
            s_ = L.stream()
            L.call('kp_tps_warp', s_, L.ptr(src), L.ptr(tmp), L.ptr(theta[d]), L.ptr(ctrl[d]), n, c, H, W, T, 0)
            L.call('kp_rotate_warp', s_, L.ptr(tmp), L.ptr(dst), L.ptr(rot[d]), n, c, H, W)

        mk = lambda name: self.misc(name, (n, c, H, W), torch.float32, dev)
        tmp, x1, x2, m1, m2 = mk('aug.tmp'), mk('aug.x1'), mk('aug.x2'), mk('aug.m1'), mk('aug.m2')
        tmp_m = mk('aug.tmp_m')
        p1, p2 = 0, 1
        # the loss mask P2(P1(1)) is independent of the image chain P2(P1(x)): the small latency-bound warp kernels of the
        # two chains run side by side
        with self._fork():
            perturb(None, m1, tmp_m, p1)      # NULL source = the constant image 1
            perturb(m1, m2, tmp_m, p2)
        perturb(x, x1, tmp, p1)
        perturb(x1, x2, tmp, p2)
        self._join()
        return x1, x2, m2

    # ------------------------------------------------------------------------------------------
    def _forward_backward(self, xa, xb, mask):
        """xa: source / first image, xb: target (the loss compares the reconstruction with xb)."""
        n, c, H, W = xa.shape
        dev = self.device
        st = L.stream()
        enc, kp, dec = self.units['encoder'], self.units.get('keypoint'), self.units['decoder']
        h, w = self._bottleneck_dims(H, W)
        C = enc.specs[-1].cout
        K = kp.specs[-1].cout if kp is not None else 0
        f32 = torch.float32
        prec = self.precision
        cin_p = engine.pitch(c, prec)
        dec_cin = dec.specs[0].cin
        dec_cp = engine.pitch(dec_cin, prec)
        shapes = {'encoder': self._pitches(enc, cin_p), 'decoder': self._pitches(dec, dec_cp)}
        if kp is not None:
            shapes['keypoint'] = self._pitches(kp, cin_p)
        self._pack(shapes)
        self._staging(shapes)
        dec_in = self.misc('dec_in', (n, h + 2, w + 2, dec_cp), self.T, dev, zero=True)
        xhat = self.misc('xhat', (n, c, H, W), f32, dev)
        dxhat = self.misc('dxhat', (n, c, H, W), f32, dev)
        sig, eps = self.sigma, 1e-6
        mode = L.COMBINE[self.combine]
        if kp is not None:
            heat = self.misc('heat', (n, K, h, w), f32, dev)
            k_t = self.misc('k_t', (n, K, 2), f32, dev)
            p_h = self.misc('p_h', (n, K, h), f32, dev)
            p_w = self.misc('p_w', (n, K, w), f32, dev)

        if self.kind == 'autoencoder':          # z = encoder(x); x_hat = decoder(z)  (models/autoencoder.py:14-17)
            xe = engine.to_padded(xa, prec, self.misc, 'xe')
            self._fwd_unit(enc, xe, H, W, dec_in[..., :C], 1)
        elif self.kind == 'keynet':
            m = self.misc('m', (n, K, h, w), f32, dev)
            xe = engine.to_padded(xa, prec, self.misc, 'xe')
            xk = engine.to_padded(xb, prec, self.misc, 'xk')
            with self._fork():
                self._fwd_unit(enc, xe, H, W, dec_in[..., :C], 1)
            self._fwd_unit(kp, xk, H, W, heat.permute(0, 2, 3, 1), 0)
            L.call('kp_spatial_softmax_fwd', st, L.ptr(heat), n * K, h, w, L.ptr(k_t), L.ptr(p_h), L.ptr(p_w))
            L.call('kp_gaussian_fwd', st, L.ptr(k_t), n * K, h, w, sig, eps, L.ptr(m))
            L.call('kp_bn_act_fwd', st, L.nchw(m), L.view(dec_in[..., C:C + K]), None, None, L.ACT_NONE, L.POST_NONE, 1,
                   n, h, w, K)
            self._join()
        else:
            phi_s = self.misc('phi_s', (n, h, w, C), self.T, dev)
            phi_t = self.misc('phi_t', (n, h, w, C), self.T, dev)
            k_s = self.misc('k_s', (n, K, 2), f32, dev)
            mask_s = self.misc('mask_s', (n, 1, h, w), f32, dev)
            mask_t = self.misc('mask_t', (n, 1, h, w), f32, dev)
            amax = self.misc('amax', (n, h, w), torch.int32, dev)
            xs = engine.to_padded(xa, prec, self.misc, 'xs')
            xt = engine.to_padded(xb, prec, self.misc, 'xt')
            # source frame: constants, but its BatchNorm running statistics update (transporter.py:36-37)
            with self._fork():
                self._fwd_unit(enc, xs, H, W, phi_s, 0)
            self._fwd_unit(kp, xs, H, W, heat.permute(0, 2, 3, 1), 0)
            L.call('kp_spatial_softmax_fwd', st, L.ptr(heat), n * K, h, w, L.ptr(k_s), None, None)
            self._join()
            with self._fork():
                self._fwd_unit(enc, xt, H, W, phi_t, 0)
            self._fwd_unit(kp, xt, H, W, heat.permute(0, 2, 3, 1), 0)
            L.call('kp_spatial_softmax_fwd', st, L.ptr(heat), n * K, h, w, L.ptr(k_t), L.ptr(p_h), L.ptr(p_w))
            self._join()
            if self.combine == 'max':
                L.call('kp_transport_fwd', st, L.view(phi_s), L.view(phi_t), L.ptr(k_s), L.ptr(k_t), L.view(dec_in[..., :C]),
                       1, L.ptr(mask_s), L.ptr(mask_t), L.ptr(amax), n, h, w, C, K, sig, eps)
            else:                               # 'sum_and_clamp' / 'loop' (models/transporter.py:41-50)
                coef = self.misc('coef', (n, h, w, 2), f32, dev)
                L.call('kp_transport_mode_fwd', st, mode, L.view(phi_s), L.view(phi_t), L.ptr(k_s), L.ptr(k_t),
                       L.view(dec_in[..., :C]), 1, L.ptr(mask_s), L.ptr(mask_t), L.ptr(amax), L.ptr(coef), n, h, w, C, K,
                       sig, eps)
        self._fwd_unit(dec, dec_in, h, w, xhat.permute(0, 2, 3, 1), 0)

        self.numel = xhat.numel()
        L.zero(self.loss_sum)
        L.call('kp_l2_loss', st, L.ptr(xhat), L.ptr(xb), L.ptr(mask), self.numel, 1.0 / self.numel, L.ptr(self.loss_sum),
               L.ptr(dxhat))
        if self.loss_ring is not None:
            self.loss_ring.scale = 1.0 / self.numel
            self.loss_ring.push(self.loss_sum, self.step_dev)

        # ---- backward ----
        L.zero(self.flat_g)
        if self._fold_table is not None:
            L.zero(self._stg_arena)
        buckets = self._buckets()
        ddec = self._bwd_unit(dec, dxhat.permute(0, 2, 3, 1), 0, True, buckets)
        if self.kind == 'autoencoder':
            self._bwd_unit(enc, ddec[..., :C], 1, False, buckets)
        else:
            dk = self.misc('dk', (n, K, 2), f32, dev)
            dheat = self.misc('dheat', (n, K, h, w), f32, dev)
            if self.kind == 'keynet':
                L.call('kp_gaussian_bwd', st, L.view(ddec[..., C:C + K]), 1, L.ptr(k_t), None, n, K, h, w, sig, eps, L.ptr(dk))
                denc, denc_pad = ddec[..., :C], 1
            else:
                dphi = self.misc('dphi_t', (n, h, w, C), self.T, dev)
                if self.combine == 'max':
                    dmask = self.misc('dmask', (n, h, w, 1), f32, dev)
                    L.call('kp_transport_bwd', st, L.view(ddec[..., :C]), 1, L.view(phi_s), L.view(phi_t), L.ptr(mask_s),
                           L.ptr(mask_t), L.view(dphi), L.ptr(dmask), n, h, w, C)
                    L.call('kp_gaussian_bwd', st, L.view(dmask), 0, L.ptr(k_t), L.ptr(amax), n, K, h, w, sig, eps, L.ptr(dk))
                else:
                    dm_t = self.misc('dm_t', (n, K, h, w), f32, dev)
                    L.call('kp_transport_mode_bwd', st, mode, L.view(ddec[..., :C]), 1, L.view(phi_s), L.view(phi_t),
                           L.ptr(k_s), L.ptr(k_t), L.ptr(mask_s), L.ptr(mask_t), L.ptr(amax), L.ptr(coef), L.view(dphi),
                           L.ptr(dm_t), n, h, w, C, K, sig, eps)
                    L.call('kp_gaussian_bwd', st, L.nchw(dm_t), 0, L.ptr(k_t), None, n, K, h, w, sig, eps, L.ptr(dk))
                denc, denc_pad = dphi, 0
            L.call('kp_spatial_softmax_bwd', st, L.ptr(dk), L.ptr(k_t), L.ptr(p_h), L.ptr(p_w), n * K, h, w, L.ptr(dheat))
            with self._fork():
                self._bwd_unit(enc, denc, denc_pad, False, buckets)
            self._bwd_unit(kp, dheat.permute(0, 2, 3, 1), 0, False, buckets)
            self._join()
        if buckets is None:
            self._fold_wgrads()

    # ------------------------------------------------------------------------------------------
    # Gradient buckets.  A bucket is a contiguous span of the flat gradient buffer that backward completes at a
    # known point: the whole decoder when its last (input-side) layer is done, the deep layers of the keypoint / encoder
    # stacks (>= 85 % of their weights, the 16x16 / 32x32 layers that backward reaches first) and then their shallow rest.
    # When a bucket completes its staged tensor-core weight gradients are folded and its sum all-reduce is issued at once;
    # NCCL runs it on its own stream under the remaining backward kernels.
    def _bwd_unit(self, u: _UnitState, dout, dout_pad, need_dx, buckets=None):
        hook = None
        if buckets is not None:
            ends = buckets[u.name]            # {layer index at which a bucket completes: (first layer, last layer + 1)}

            def hook(i):
                if i in ends:
                    self._bucket_ready(u, *ends[i])
        ws = None
        if self.wgrad_streams and hook is None:
            ws = self._wg_side.get(u.name)
            if ws is None:
                ws = self._wg_side[u.name] = torch.cuda.Stream(device=self.device)
        return engine.unit_backward(u.specs, u.params, u.grads, u.ctxs, dout, dout_pad, self.precision, need_dx,
                                    alloc=u.alloc, tag='b', defer_stg=self._stg.get(u.name), after_wgrad=hook,
                                    wgrad_stream=ws)

    def _buckets(self):
        if self.world == 1 or not self.overlap_allreduce or self.dp_mode != 'nccl':
            return None
        if getattr(self, '_bucket_plan', None) is None:
            plan = {}
            for u in self.units.values():
                nl = len(u.specs)
                if u.name == 'decoder':
                    plan[u.name] = {0: (0, nl)}
                    continue
                sizes = [sum(t.numel() for t in knn.trainable([m])) for m in u.mods]
                total, tail, split = sum(sizes), 0, 0
                for i in range(nl - 1, 0, -1):
                    tail += sizes[i]
                    if tail >= 0.85 * total:
                        split = i
                        break
                if split == 0 or total < (1 << 20):       # small nets: one bucket per unit
                    plan[u.name] = {0: (0, nl)}
                else:
                    plan[u.name] = {split: (split, nl), 0: (0, split)}
            self._bucket_plan = plan
        return self._bucket_plan

    def _bucket_ready(self, u: _UnitState, lo: int, hi: int):
        """Layers [lo, hi) of unit u have all their gradients issued on the current stream: fold the staged weight
        gradients of those layers and start the bucket's sum all-reduce (asynchronous; `_wait_allreduce` joins)."""
        rows = [self._fold_rows[(u.name, i)] for i in range(lo, hi) if (u.name, i) in self._fold_rows]
        if rows:
            a, b = min(rows), max(rows) + 1
            assert b - a == len(rows)
            L.call('kp_wgrad_finalize_multi', L.stream(), L.ptr(self._fold_table[a:b]), b - a)
        start = u.layer_off[lo]
        end = u.layer_off[hi] if hi < len(u.specs) else u.span[1]
        self._works += parallel.allreduce_buckets(self.flat_g, [(start, end)], self.pg)

    def _fold_wgrads(self):
        if self._fold_table is not None:
            L.call('kp_wgrad_finalize_multi', L.stream(), L.ptr(self._fold_table), self._fold_table.shape[0])

    def _allreduce(self):
        """Join the bucket all-reduces issued during backward (or, without overlap, run them now): NCCL over
        NVLink/NVSwitch on the GPUs."""
        if self.world == 1 or self.dp_mode != 'nccl':
            return
        if self._buckets() is None:
            self._works += parallel.allreduce_buckets(self.flat_g, [u.span for u in self.units.values()], self.pg)
        parallel.wait_all(self._works)
        self._works = []

    def _adam(self):
        L.call('kp_adam_step', L.stream(), L.ptr(self.flat_p), L.ptr(self.flat_g), L.ptr(self.flat_m), L.ptr(self.flat_v),
               self.n_params, self.lr, self.betas[0], self.betas[1], self.eps, 0, 1.0 / self.world, L.ptr(self.step_dev))

    def _dp_adam(self):
        """Gradient reduce-scatter + Adam + parameter all-gather over NVLink peer memory, one fused kernel between two flag
        barriers (csrc/kp_dp.cu); replaces `_allreduce` + `_adam` when the replicas can map each other's buffers."""
        import ctypes
        pb = self.peer
        L.call('kp_dp_adam_step', L.stream(), ctypes.byref(pb.peers), pb.rank, pb.world, self.n_params, L.ptr(self.flat_m),
               L.ptr(self.flat_v), self.lr, self.betas[0], self.betas[1], self.eps, 1.0 / self.world, L.ptr(self.step_dev),
               L.ptr(pb.epoch), 1 if pb.multicast else 0)

    def _full_moments(self):
        """Adam moments of the whole bucket.  In 'p2p' mode every rank only holds (and updates) its own slice; the rest of
        its buffers is whatever was loaded / broadcast, so the slices are summed over the replicas (collective: every rank
        must call)."""
        if self.peer is None:
            return self.flat_m, self.flat_v
        lo, hi = self.peer.owned(self.n_params)
        out = []
        for t in (self.flat_m, self.flat_v):
            own = torch.zeros_like(t)
            own[lo:hi] = t[lo:hi]
            torch.distributed.all_reduce(own, group=self.pg)
            out.append(own)
        return out

    def _whole(self, xa, xb, mask):
        if self.kind == 'autoencoder':
            xb = xa                              # the target is the input (autoencode.py:88-90)
        elif self.augment is not None:
            xa, xb, mask = self._augment(xa)
        self._forward_backward(xa, xb, mask)
        if self.world > 1 and self.dp_mode == 'p2p':
            self._dp_adam()
        else:
            self._allreduce()
            self._adam()

    # ------------------------------------------------------------------------------------------
    def step(self, xa: torch.Tensor, xb: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None):
        """One training step on device tensors (NCHW fp32).  With ``augment`` set, ``xa`` is the clean batch and the
        (x, x_, loss_mask) triple is produced by the TPS+rotate kernels.  Returns the device scalar holding
        sum((xhat-x_)^2 mask); ``loss()`` converts it."""
        with torch.cuda.device(self.device):
            self.ctx.use()
            xa = xa.to(self.device, torch.float32).contiguous()
            paired = self.augment is None and self.kind != 'autoencoder'
            if paired:
                if xb is None:
                    raise ValueError('xb is required without augmentation')
                xb = xb.to(self.device, torch.float32).contiguous()
                if mask is not None:
                    mask = mask.to(self.device, torch.float32).contiguous()
            else:
                xb = mask = None
            if not self.use_graph:
                self._whole(xa, xb, mask)
                self.steps_done += 1
                return self.loss_sum
            key = (tuple(xa.shape), None if xb is None else tuple(xb.shape), None if mask is None else tuple(mask.shape))
            if self.graph is None or self.graph_key != key:
                self._capture(xa, xb, mask, key)
            self.in_a.copy_(xa, non_blocking=True)
            if xb is not None:
                self.in_b.copy_(xb, non_blocking=True)
            if mask is not None:
                self.in_m.copy_(mask, non_blocking=True)
            self.graph.replay()
            self.steps_done += 1
            return self.loss_sum

    def _capture(self, xa, xb, mask, key):
        self.in_a = xa.clone()
        self.in_b = None if xb is None else xb.clone()
        self.in_m = None if mask is None else mask.clone()
        # warm-up outside capture: allocates every static buffer, sets kernel attributes, builds nothing lazily later
        # (and, data parallel, runs the collectives once so NCCL has its channels before they are captured)
        state = self._snapshot()
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._whole(self.in_a, self.in_b, self.in_m)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self._restore(state)
        if self.world > 1:
            torch.distributed.barrier(group=self.pg)
            torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        before = L.launches
        # thread_local: NCCL's watchdog thread may touch the CUDA API while this thread captures
        with torch.cuda.graph(self.graph, capture_error_mode='thread_local'):
            self._whole(self.in_a, self.in_b, self.in_m)
        self.calls_per_step = L.launches - before
        self.graph_key = key

    def close(self):
        """Drop the captured graph.  Data parallel: call before ``destroy_process_group`` — NCCL cannot tear down a
        communicator while a live CUDA graph still holds its captured collectives (the destroy blocks forever)."""
        import gc
        torch.cuda.synchronize(self.device)
        self.graph = None
        self.graph_key = None
        gc.collect()

    def _snapshot(self):
        bufs = [b for _, bn in self._all_mods() if bn is not None for b in (bn.running_mean, bn.running_var, bn.num_batches_tracked)]
        return ([t.clone() for t in (self.flat_p, self.flat_m, self.flat_v, self.step_dev)], [(b, b.clone()) for b in bufs])

    def _restore(self, state):
        flats, bufs = state
        for dst, src in zip((self.flat_p, self.flat_m, self.flat_v, self.step_dev), flats):
            dst.copy_(src)
        for b, saved in bufs:
            b.copy_(saved)

    def _all_mods(self):
        for u in self.units.values():
            yield from u.mods

    # ------------------------------------------------------------------------------------------
    def loss(self) -> float:
        """Mean masked L2 loss of the last step (device -> host read)."""
        return float(self.loss_sum.item()) / self.numel

    def named_grads(self):
        """{parameter name (as in net.named_parameters()): view of its gradient in the flat bucket} of the last step."""
        out = {}
        for n, t in self.net.named_parameters():
            o = (t.data.data_ptr() - self.flat_p.data_ptr()) // 4
            if 0 <= o < self.n_params:
                out[n] = self.flat_g[o:o + t.numel()].view_as(t)
        return out

    def attach_loss_ring(self, ring):
        """Have every step append (step, mean loss) to a device ring (runlog.LossRing) — the non-blocking replacement of
        the per-step `loss.item()` (utils.py:124).  Part of the captured graph, so attaching re-captures."""
        self.loss_ring = ring
        self.graph = None

    def outputs(self):
        """Keypoints (N,K,2) (y,x) (None for the auto-encoder) and reconstruction of the last step (static buffers)."""
        return self.misc.bufs.get(('misc', 'k_t')), self.misc.bufs[('misc', 'xhat')]

    # ------------------------------------------------------------------------------------------
    # Resume support.  The reference checkpoints only the module weights (9 .mdl files, knn.py:143-167) and
    # restarts Adam from zero; `save` writes those same files through the module plus the optimiser state next to them.
    def state_dict(self):
        """Optimiser state of the fused trainer: Adam moments as tensors shaped like the parameters (keyed like
        ``net.state_dict()``), the step count and the hyper-parameters."""
        names = {id(p): n for n, p in self.net.named_parameters()}
        m, v = {}, {}
        flat_m, flat_v = self._full_moments()
        for u in self.units.values():
            for t in u.tensors():
                o = t.data.data_ptr() - self.flat_p.data_ptr()
                assert o % 4 == 0
                o //= 4
                k = t.numel()
                m[names[id(t)]] = flat_m[o:o + k].view_as(t).clone()
                v[names[id(t)]] = flat_v[o:o + k].view_as(t).clone()
        return {'exp_avg': m, 'exp_avg_sq': v, 'step': int(self.step_dev.item()), 'lr': self.lr, 'betas': tuple(self.betas),
                'eps': self.eps, 'precision': self.precision}

    def load_state_dict(self, sd):
        names = {n: p for n, p in self.net.named_parameters()}
        for key, dst in (('exp_avg', self.flat_m), ('exp_avg_sq', self.flat_v)):
            for n, src in sd[key].items():
                t = names[n]
                o = (t.data.data_ptr() - self.flat_p.data_ptr()) // 4
                dst[o:o + t.numel()].copy_(src.reshape(-1).to(self.device, torch.float32))
        self.step_dev.fill_(int(sd['step']))
        self.steps_done = int(sd['step'])
        self.lr, self.betas, self.eps = float(sd['lr']), tuple(sd['betas']), float(sd['eps'])
        self.graph = None                      # lr / betas are baked into the captured Adam launch

    def save(self, directory):
        """Module weights in the reference's layout (loadable by the reference's ``net.load``) + ``trainer.pt``.
        Data parallel: call on EVERY rank (gathering the sharded Adam moments is a collective); rank 0 writes."""
        import os as _os
        sd = self.state_dict()
        if self.rank == 0:
            self.net.save(directory)
            torch.save(sd, _os.path.join(directory, 'trainer.pt'))
        if self.world > 1:
            torch.distributed.barrier(group=self.pg)

    def load(self, directory, map_device=None):
        import os as _os
        # the module's parameters alias the flat bucket: load_state_dict copies in place, the aliasing survives
        self.net.load(directory, map_device=map_device or str(self.device))
        path = _os.path.join(directory, 'trainer.pt')
        if _os.path.exists(path):
            self.load_state_dict(torch.load(path, map_location=self.device))
        self.broadcast_state()                 # data parallel: every rank calls load(); rank 0's copy wins

    def activation_bytes(self):
        return sum(u.alloc.nbytes() for u in self.units.values()) + self.misc.nbytes()
