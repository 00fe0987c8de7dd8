"""TrainEngine: one KeypointsGauss training step (forward in train() mode + loss + backward) as a fixed sequence of
libhulk_sm100 kernels -- the B200-native replacement of what torch autograd executes for reference train.py:33-35
(zero_grad, forward(sample, model), loss.backward()).  SURVEY.md §8 rows a4/a9/a11/a12 and f1.

Numerics: bf16 NHWC activations and bf16 conv operands (tcgen05, fp32 accumulation in TMEM), fp32 BatchNorm statistics,
fp32 parameter gradients, fp64 loss.  BatchNorm uses the statistics of the local batch (per replica, like the reference,
which has no SyncBN) and updates running_mean / running_var / num_batches_tracked exactly as nn.BatchNorm2d does.

Per conv the forward is   y = conv(x) [raw, bf16]  ->  batch stats  ->  out = relu(gamma*xhat + beta [+ shortcut])
and the backward          d' = dout*[out>0]  ->  (dgamma, dbeta, dy)  ->  dW = wgrad(x, dy),  dx = conv(dy, W flipped).
All 110 parameter gradients land in ONE flat fp32 buffer laid out like FusedAdam's (one all-reduce, one update kernel);
the 996 dead rows of the 1000-row scoring conv keep a zero gradient, as in the reference.

Data layout in HBM (per engine, batch B of H x W): input (B,3,H,W) f32; per conv a raw output and an activation,
(B,h,w,C) bf16 NHWC; (B,K,H,W) f32 upsampled logits and their gradient; scratch gradients sized by stage.
The whole step is captured in one CUDA graph (~560 launches) and replayed.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Tuple

import torch

from . import ops
from ._lib import HK_BF16, check, lib, ptr, require_device, stream_ptr


class _ConvT:
    """One conv + its BatchNorm in training: parameters, packed operands, saved tensors."""

    def __init__(self, name, conv, bn, B, H, W, dev):
        self.name, self.conv, self.bn = name, conv, bn
        self.cout, self.cin, self.k, _ = conv.weight.shape
        self.stride, self.pad, self.dil = conv.stride[0], conv.padding[0], conv.dilation[0]
        self.H, self.W = H, W
        self.Ho, self.Wo = ops.conv_out_hw(H, W, self.k, self.stride, self.pad, self.dil)
        bf = torch.bfloat16
        self.w_fwd = torch.empty((self.cout, self.k, self.k, self.cin), device=dev, dtype=bf)
        self.w_dgrad = torch.empty((self.cin, self.k, self.k, self.cout), device=dev, dtype=bf)
        self.one_out = torch.empty(self.cout, device=dev)   # scale = 1 / bias = 0 of the raw forward conv (filled by the pack kernel)
        self.zero_out = torch.empty(self.cout, device=dev)
        self.one_in = torch.ones(self.cin, device=dev)      # same for the data-gradient conv (its "Cout" is cin)
        self.zero_in = torch.zeros(self.cin, device=dev)
        self.y = torch.empty((B, self.Ho, self.Wo, self.cout), device=dev, dtype=bf)   # raw conv output (saved for BN backward)
        self.mean, self.invstd, self.scale, self.shift = (torch.empty(self.cout, device=dev) for _ in range(4))
        self.relu_bits = torch.empty(self.y.numel() // 8, device=dev, dtype=torch.uint8)  # ReLU mask of the BN output, 1 bit / element


class TrainEngine:
    def __init__(self, model, B: int, H: int, W: int, sigma: float = 8.0):
        require_device()
        if H % 8 or W % 8 or H < 32 or W < 32:
            raise ValueError("TrainEngine needs H and W to be multiples of 8 (>= 32)")
        net = model.resnet.resnet34_8s
        dev = net.conv1.weight.device
        if dev.type != "cuda":
            raise RuntimeError("TrainEngine needs the model on a CUDA device (no CPU path)")
        self.model, self.net, self.device = model, net, dev
        self.B, self.H, self.W, self.K, self.sigma = B, H, W, int(model.num_keypoints), float(sigma)
        bf = torch.bfloat16
        # ---- flat gradient buffer, FusedAdam layout (each tensor starts on a 16-byte boundary) ----
        self.params: List[torch.nn.Parameter] = [p for p in model.parameters() if p.requires_grad]
        offs, total = [], 0
        for p in self.params:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4
        self.flat_grad = torch.zeros(total, device=dev, dtype=torch.float32)
        self.grad_of: Dict[int, torch.Tensor] = {id(p): self.flat_grad[o:o + p.numel()].view_as(p) for p, o in zip(self.params, offs)}
        # ---- buffers ----
        self.x = torch.empty((B, 3, H, W), device=dev, dtype=torch.float32)
        self.uv = torch.zeros((B, self.K, 2), device=dev, dtype=torch.float32)
        self.target: Optional[torch.Tensor] = None      # (B,K,H,W) fp64/fp32 when the caller supplies targets instead of labels
        h2, w2 = ops.conv_out_hw(H, W, 7, 2, 3, 1)
        h4, w4 = (h2 + 2 - 3) // 2 + 1, (w2 + 2 - 3) // 2 + 1
        self.h2, self.w2, self.h4, self.w4 = h2, w2, h4, w4
        self.stem_w = torch.empty(int(lib().hk_stem_packed_weight_bytes()) // 2, device=dev, dtype=bf)
        self.stem = _ConvT("stem", net.conv1, net.bn1, B, H, W, dev)
        self.a0 = torch.empty((B, h2, w2, 64), device=dev, dtype=bf)
        self.p0 = torch.empty((B, h4, w4, 64), device=dev, dtype=bf)
        self.pool_idx = torch.empty(B * h4 * w4 * 64, device=dev, dtype=torch.uint8)
        self.blocks = []
        h, w = h4, w4
        for i, blk in enumerate(net.blocks()):
            c1 = _ConvT(f"b{i}.c1", blk.conv1, blk.bn1, B, h, w, dev)
            c2 = _ConvT(f"b{i}.c2", blk.conv2, blk.bn2, B, c1.Ho, c1.Wo, dev)
            ds = _ConvT(f"b{i}.ds", blk.downsample[0], blk.downsample[1], B, h, w, dev) if blk.downsample is not None else None
            a1 = torch.empty_like(c1.y)
            sc = torch.empty_like(c2.y) if ds is not None else None
            out = torch.empty_like(c2.y)
            self.blocks.append((c1, c2, ds, a1, sc, out))
            h, w = c1.Ho, c1.Wo
        self.h8, self.w8 = h, w
        self.convs: List[_ConvT] = [self.stem] + [c for b in self.blocks for c in (b[0], b[1], b[2]) if c is not None]
        # BatchNorm reductions through exact fixed-point accumulators (hk_bn_stats_acc / hk_bn_bwd_acc): no finalize launches, 72 fewer
        # launches per step.  One flat buffer per direction, sliced per conv in self.convs order (stem, then the blocks in order), so the
        # early part of the two-phase backward (stem + blocks < split_block) is a prefix.  HK_BN_ACC=0: the three-launch path.
        self.use_bn_acc = os.environ.get("HK_BN_ACC", "1") != "0"
        # ... and the forward statistics gathered in the conv epilogues (hk_conv_bn_stats_fwd): 36 launches and one read of every raw conv
        # output fewer.  HK_CONV_STATS=0: stand-alone hk_bn_stats_acc.
        self.fuse_conv_stats = os.environ.get("HK_CONV_STATS", "1") != "0"
        # HK_CONV_STATS=all: every conv with a specialised tcgen05 kernel gathers its statistics in the epilogue (A/B switch; at small batch
        # the stand-alone reduction is a latency-bound launch, at large batch the epilogue-bound 64/128-channel layers lose more than it costs)
        self.fuse_conv_stats_all = os.environ.get("HK_CONV_STATS", "1") == "all"
        # max-pool forward records the winning taps (one pass over the stem map instead of two); HK_MAXPOOL_IDX=0: argmax pass in the backward
        self.maxpool_idx = os.environ.get("HK_MAXPOOL_IDX", "1") != "0"
        offs_acc, tot = [], 0
        for c in self.convs:
            offs_acc.append(tot)
            tot += ops.bn_acc_bytes(c.cout)
        self.acc_fwd = torch.zeros(tot, device=dev, dtype=torch.uint8)
        self.acc_bwd = torch.zeros(tot, device=dev, dtype=torch.uint8)
        for c, o in zip(self.convs, offs_acc):
            c.acc_f = self.acc_fwd[o:o + ops.bn_acc_bytes(c.cout)]
            c.acc_b = self.acc_bwd[o:o + ops.bn_acc_bytes(c.cout)]
        K = self.K
        self.logits_lr = torch.empty((B, K, h, w), device=dev, dtype=torch.float32)
        self.logits_up = torch.empty((B, K, H, W), device=dev, dtype=torch.float32)
        self.g_up = torch.empty((B, K, H, W), device=dev, dtype=torch.float32)
        self.dlogits_lr = torch.empty((B, K, h, w), device=dev, dtype=torch.float32)
        self.loss = torch.zeros((), device=dev, dtype=torch.float64)
        n = B * K * H * W
        self.bce_ws = torch.empty(int(lib().hk_bce_workspace_bytes(n)), device=dev, dtype=torch.uint8)
        self.bn_ws = ops.bn_workspace(512, dev)
        wg = max(ops.conv_wgrad_workspace_bytes(B, c.H, c.W, c.cin, c.cout, c.k, c.stride, c.pad, c.dil) for c in self.convs[1:])
        self.wgrad_ws = torch.empty(max(wg, int(lib().hk_stem_wgrad_workspace_bytes())), device=dev, dtype=torch.uint8)
        self.head_ws = torch.empty(int(lib().hk_head_bwd_workspace_bytes(B, K, 512, h, w)), device=dev, dtype=torch.uint8)
        # first block of layer3: everything from here to the end of the parameter list (layer3, layer4, fc) is the "late" part
        self.split_block = 7
        first_late = self.blocks[self.split_block][0].conv.weight
        self.acc_early_bytes = offs_acc[self.convs.index(self.blocks[self.split_block][0])]   # accumulators of the stem + early blocks
        self.late_offset = next(o for p, o in zip(self.params, offs) if p is first_late)
        self._bwd_state = None
        self._scratch: Dict[Tuple, torch.Tensor] = {}
        self._pack_key = None
        self._pack_items: Optional[torch.Tensor] = None
        self._bn_counters = [c.bn.num_batches_tracked for c in self.convs]
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self._graph_key = None
        self.use_cuda_graph = os.environ.get("HK_TRAIN_NO_GRAPH") is None   # eager launches for ncu launch lists
        # ReLU masks of the backward pass as 1 bit / element written by the forward apply pass (HK_BN_BITS=0: re-read the bf16 outputs)
        self.use_relu_bits = os.environ.get("HK_BN_BITS", "1") != "0"
        # Weight gradients are leaves of the backward chain (bn_bwd -> dgrad -> bn_bwd ...): they run on a second stream inside the same
        # graph, concurrently with the data-gradient chain (each conv then keeps its own dy buffer).  Measured on one box: per-GPU batch 4
        # 5.74 -> 5.37 ms per step, batch 32 25.40 -> 24.53 ms.  HK_WGRAD_STREAM=0 keeps everything on one stream.
        self.wgrad_stream = torch.cuda.Stream(device=dev) if os.environ.get("HK_WGRAD_STREAM", "1") != "0" else None
        # The 1x1 downsample branch of the three stride/channel-changing blocks (conv + BN forward; BN backward + dgrad) is independent
        # of the block's main branch until the residual add: it runs on a third stream with its own BN workspace (HK_AUX_STREAM=0: off).
        self.aux_stream = torch.cuda.Stream(device=dev) if os.environ.get("HK_AUX_STREAM", "1") != "0" else None
        self.bn_ws_aux = ops.bn_workspace(512, dev) if self.aux_stream is not None else None
        # HK_WGRAD_STREAM=0 with the aux stream on: the downsample conv's weight gradient then runs on the aux stream concurrently with
        # conv2's on the main stream, so it needs its own split-K partials buffer
        self.wgrad_ws_aux = None
        if self.aux_stream is not None and self.wgrad_stream is None:
            wga = max(ops.conv_wgrad_workspace_bytes(B, c.H, c.W, c.cin, c.cout, c.k, c.stride, c.pad, c.dil)
                      for b in self.blocks for c in (b[2],) if c is not None)
            self.wgrad_ws_aux = torch.empty(wga, device=dev, dtype=torch.uint8)
        self.launches = 0

    # ------------------------------------------------------------------ helpers
    def _buf(self, role: str, shape) -> torch.Tensor:
        key = (role, tuple(shape))
        t = self._scratch.get(key)
        if t is None:
            t = torch.empty(tuple(shape), device=self.device, dtype=torch.bfloat16)
            self._scratch[key] = t
        return t

    def _g(self, p) -> torch.Tensor:
        return self.grad_of[id(p)]

    def _pack_all(self) -> int:
        """Every conv weight -> bf16 forward + data-gradient layouts in one launch (hk_pack_conv_weights_many)."""
        import numpy as np
        key = tuple(c.conv.weight.data_ptr() for c in self.convs)
        if self._pack_items is None or key != self._pack_key:
            dt = np.dtype([("w", "<u8"), ("f", "<u8"), ("d", "<u8"), ("cout", "<i4"), ("cin", "<i4"), ("khw", "<i4"), ("r", "<i4")])
            arr = np.zeros(len(self.convs) - 1, dtype=dt)
            for i, c in enumerate(self.convs[1:]):
                arr[i] = (c.conv.weight.data_ptr(), c.w_fwd.data_ptr(), c.w_dgrad.data_ptr(), c.cout, c.cin, c.k * c.k, 0)
            self._pack_items = torch.from_numpy(arr.view(np.uint8).copy()).to(self.device)
            self._pack_key = key
            self._pack_max = max(c.cout * c.cin * c.k * c.k for c in self.convs[1:])
            self._pack_khw = max(c.k * c.k for c in self.convs[1:])
            # shared-memory tiled repack (contiguous reads, 64-byte writes) when every conv tiles 64 x 32; HK_PACK_TILED=0: element-wise kernel
            self._pack_tiled = (os.environ.get("HK_PACK_TILED", "1") != "0"
                                and all(c.cout % 64 == 0 and c.cin % 32 == 0 for c in self.convs[1:]))
            for c in self.convs:  # scale = 1 / bias = 0 of the raw convs: constant
                c.one_out.fill_(1.0)
                c.zero_out.zero_()
        if self._pack_tiled:
            check(lib().hk_pack_conv_weights_many_tiled(ptr(self._pack_items), len(self.convs) - 1, self._pack_khw, C.c_longlong(self._pack_max),
                                                        stream_ptr()), "hk_pack_conv_weights_many_tiled")
        else:
            check(lib().hk_pack_conv_weights_many(ptr(self._pack_items), len(self.convs) - 1, C.c_longlong(self._pack_max), stream_ptr()),
                  "hk_pack_conv_weights_many")
        return 1

    def _pack(self, c: _ConvT, with_dgrad: bool) -> int:
        w = c.conv.weight.data
        check(lib().hk_pack_conv_weights(ptr(w), None, None, None, None, C.c_float(1e-5), c.cout, c.cin, c.k, c.k, HK_BF16,
                                         ptr(c.w_fwd), ptr(c.one_out), ptr(c.zero_out), stream_ptr()), "hk_pack_conv_weights")
        n = 2
        if with_dgrad:
            ops.pack_conv_weights_dgrad(w, out=c.w_dgrad)
            n += 1
        return n

    def _conv_bn(self, c: _ConvT, x, out, relu: bool, residual=None, ws=None) -> int:
        """raw conv -> batch statistics (+ running stats) -> fused normalise (+ shortcut) (+ ReLU)."""
        bn = c.bn
        if self.use_bn_acc and self.fuse_conv_stats and ((c.k == 3 and c.cout >= 256) or (self.fuse_conv_stats_all and c is not self.stem)):
            # the conv epilogue gathers sum y / sum y^2 of the tile it stores: no statistics pass over the activation.  Only where the
            # epilogue hides under a long mainloop (3x3 convs of layers 3-4: 36 / 72 K blocks per tile); the 64/128-channel layers and
            # the 1x1 downsample convs are epilogue-bound and lose more than the stand-alone reduction costs (measured per launch)
            ops.conv_bn_stats(x, c.w_fwd, c.one_out, c.zero_out, c.acc_f, stride=c.stride, pad=c.pad, dil=c.dil, out=c.y)
            ops.bn_apply_acc(c.y, c.acc_f, bn.weight.data, bn.bias.data, bn.running_mean, bn.running_var, bn.momentum, bn.eps, c.mean, c.invstd,
                             relu=relu, residual=residual, out=out, relu_bits=c.relu_bits if relu and self.use_relu_bits else None)
            c.relu_out = out if relu else None
            return 2
        ops.conv_bn_act(x, c.w_fwd, c.one_out, c.zero_out, stride=c.stride, pad=c.pad, dil=c.dil, relu=False, out=c.y)
        if self.use_bn_acc:
            ops.bn_stats_acc(c.y, c.acc_f)
            ops.bn_apply_acc(c.y, c.acc_f, bn.weight.data, bn.bias.data, bn.running_mean, bn.running_var, bn.momentum, bn.eps, c.mean, c.invstd,
                             relu=relu, residual=residual, out=out, relu_bits=c.relu_bits if relu and self.use_relu_bits else None)
            c.relu_out = out if relu else None
            return 3
        ops.bn_train_stats(c.y, bn.weight.data, bn.bias.data, bn.running_mean, bn.running_var, bn.momentum, bn.eps, c.mean, c.invstd,
                           c.scale, c.shift, self.bn_ws if ws is None else ws)
        ops.bn_apply(c.y, c.scale, c.shift, relu=relu, residual=residual, out=out, relu_bits=c.relu_bits if relu and self.use_relu_bits else None)
        c.relu_out = out if relu else None
        return 4

    def _fork_aux(self):
        """aux stream waits for everything enqueued on the current stream so far; returns the stream (None: stay on the current one)."""
        if self.aux_stream is None:
            return None
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self.aux_stream.wait_event(ev)
        return self.aux_stream

    def _join_aux(self) -> None:
        if self.aux_stream is not None:
            ev = torch.cuda.Event()
            ev.record(self.aux_stream)
            torch.cuda.current_stream().wait_event(ev)

    def _bn_bwd(self, c: _ConvT, dout, relu: bool, dy, dmasked=None, ws=None) -> int:
        """relu: the BN output went through a ReLU; its mask is c.relu_bits (written by the forward apply pass)."""
        bn = c.bn
        mask = None if not relu else (c.relu_bits if self.use_relu_bits else c.relu_out)
        if self.use_bn_acc:
            ops.bn_bwd_acc(dout, mask, c.y, c.mean, c.invstd, bn.weight.data, c.acc_b, self._g(bn.weight), self._g(bn.bias), dy, dmasked=dmasked)
            return 2
        ops.bn_train_bwd(dout, mask, c.y, c.mean, c.invstd, bn.weight.data, self._g(bn.weight), self._g(bn.bias), dy,
                         self.bn_ws if ws is None else ws, dmasked=dmasked)
        return 3

    def _dy_buf(self, c: _ConvT, role: str = "dy") -> torch.Tensor:
        """dL/d(raw conv output) of conv c: shared scratch per shape, or (wgrads on their own stream) one buffer per conv, because the
        weight gradient may still be reading it while the data-gradient chain has moved on."""
        if self.wgrad_stream is None:
            return self._buf(role, c.y.shape)
        if getattr(c, "dy", None) is None:
            c.dy = torch.empty_like(c.y)
        return c.dy

    def _on_wgrad_stream(self, fn) -> None:
        if self.wgrad_stream is None:
            fn()
            return
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream())      # dy (just written on the main stream) is complete
        self.wgrad_stream.wait_event(ready)
        with torch.cuda.stream(self.wgrad_stream):
            fn()
        self._wgrad_pending = True

    def _join_wgrad_stream(self) -> None:
        if self.wgrad_stream is not None and getattr(self, "_wgrad_pending", False):
            done = torch.cuda.Event()
            done.record(self.wgrad_stream)
            torch.cuda.current_stream().wait_event(done)
            self._wgrad_pending = False

    def _wgrad(self, c: _ConvT, x, dy, ws=None) -> int:
        ws = self.wgrad_ws if ws is None else ws
        self._on_wgrad_stream(lambda: ops.conv_wgrad(x, dy, self._g(c.conv.weight), k=c.k, stride=c.stride, pad=c.pad, dil=c.dil, ws=ws))
        return 2

    def _dgrad(self, c: _ConvT, dy, dx, residual=None, scratch: str = "up") -> int:
        """dx = conv_transpose(dy, W) (+ residual) through the forward kernel on the flipped weights."""
        n = 1
        if c.stride == 2:
            up = self._buf(scratch, (self.B, c.H, c.W, c.cout))
            ops.zero_insert2x(dy, out=up)
            dy, n = up, 2
        ops.conv_bn_act(dy, c.w_dgrad, c.one_in, c.zero_in, stride=1, pad=c.dil * (c.k - 1) - c.pad, dil=c.dil, relu=False,
                        residual=residual, out=dx)
        return n

    # ------------------------------------------------------------------ the step
    def _enqueue(self) -> int:
        """pack + forward + fused loss + backward: the whole step."""
        return self._enqueue_forward() + self._enqueue_loss() + self._enqueue_backward()

    def _enqueue_forward(self) -> int:
        net, K = self.net, self.K
        n = 0
        # weights of this step (parameters change every optimizer step)
        check(lib().hk_stem_pack_weights(ptr(net.conv1.weight.data), ptr(self.stem_w), stream_ptr()), "hk_stem_pack_weights")
        self._on_wgrad_stream(self._pack_all)   # the block convs' weights are repacked while the stem, its BN and the maxpool run
        n += 2
        if self.use_bn_acc:
            self.acc_fwd.zero_()                # the forward accumulators of every BatchNorm of this step
            n += 1
        # ---------------- forward (train mode) ----------------
        st = self.stem
        ops.stem_conv(self.x, self.stem_w, st.one_out, st.zero_out, relu=False, out=st.y)
        bn = st.bn
        if self.use_bn_acc:
            ops.bn_stats_acc(st.y, st.acc_f)
            ops.bn_apply_acc(st.y, st.acc_f, bn.weight.data, bn.bias.data, bn.running_mean, bn.running_var, bn.momentum, bn.eps, st.mean,
                             st.invstd, relu=True, out=self.a0, relu_bits=st.relu_bits if self.use_relu_bits else None)
            n -= 1
        else:
            ops.bn_train_stats(st.y, bn.weight.data, bn.bias.data, bn.running_mean, bn.running_var, bn.momentum, bn.eps, st.mean, st.invstd,
                               st.scale, st.shift, self.bn_ws)
            ops.bn_apply(st.y, st.scale, st.shift, relu=True, out=self.a0, relu_bits=st.relu_bits if self.use_relu_bits else None)
        st.relu_out = self.a0
        if self.maxpool_idx:
            ops.maxpool3x3s2_fwd_idx(self.a0, out=self.p0, idx=self.pool_idx)   # forward + the winning taps the backward routes through
        else:
            ops.maxpool3x3s2(self.a0, out=self.p0)
        self._join_wgrad_stream()               # packed weights ready
        n += 5
        x = self.p0
        for (c1, c2, ds, a1, sc, out) in self.blocks:
            if ds is not None:
                aux = self._fork_aux()
                if aux is not None:
                    with torch.cuda.stream(aux):
                        n += self._conv_bn(ds, x, sc, relu=False, ws=self.bn_ws_aux)
                    n += self._conv_bn(c1, x, a1, relu=True)
                    self._join_aux()
                else:
                    n += self._conv_bn(c1, x, a1, relu=True)
                    n += self._conv_bn(ds, x, sc, relu=False)
                shortcut = sc
            else:
                n += self._conv_bn(c1, x, a1, relu=True)
                shortcut = x
            n += self._conv_bn(c2, a1, out, relu=True, residual=shortcut)
            x = out
        feat = x
        torch._foreach_add_(self._bn_counters, 1)
        fc_w = net.fc.weight.data[:K].view(K, 512)
        fc_b = net.fc.bias.data[:K]
        ops.head_logits(feat, fc_w, fc_b, self.H, self.W, out=self.logits_up, logits_ws=self.logits_lr)
        n += 3
        return n

    def _enqueue_loss(self) -> int:
        """sigmoid + BCE(mean) forward and d/dlogits in one pass."""
        K = self.K
        tgt = self.target
        tcode = 0 if tgt is None else ops.dtype_code(tgt.dtype)
        check(lib().hk_bce_fwd_bwd(ptr(self.logits_up), 1, ptr(tgt), tcode, None if tgt is not None else ptr(self.uv), self.B, K, self.H,
                                   self.W, C.c_float(self.sigma), ptr(self.loss), ptr(self.g_up), ptr(self.bce_ws), self.bce_ws.numel(),
                                   stream_ptr()), "hk_bce_fwd_bwd")
        return 2

    def _enqueue_backward(self, part: str = "all") -> int:
        """From self.g_up = dL/d(upsampled logits) to every parameter gradient.  part "late" = head + layers 4..3 (the tail of the
        flat gradient buffer, 91 % of its bytes), part "early" = layers 2..1 + stem: the data-parallel step all-reduces the tail
        while the early layers are still being differentiated."""
        net, K, st = self.net, self.K, self.stem
        n = 0
        nb = len(self.blocks)
        if self.use_bn_acc:   # the backward accumulators this part is going to fill
            e = self.acc_early_bytes
            (self.acc_bwd if part == "all" else self.acc_bwd[e:] if part == "late" else self.acc_bwd[:e]).zero_()
            n += 1
        if part in ("all", "late"):
            feat = self.blocks[-1][5]
            fc_w = net.fc.weight.data[:K].view(K, 512)
            d = self._buf("d0", feat.shape)
            ops.head_bwd(self.g_up, feat, fc_w, d, self._g(net.fc.weight)[:K].view(K, 512), self._g(net.fc.bias)[:K],
                         dlogits_ws=self.dlogits_lr, ws=self.head_ws)
            n += 5
            parity, hi = 1, nb
        else:
            d, parity, hi = self._bwd_state
        lo = self.split_block if part == "late" else 0
        for bi in range(hi - 1, lo - 1, -1):
            c1, c2, ds, a1, sc, out = self.blocks[bi]
            x_in = self.blocks[bi - 1][5] if bi > 0 else self.p0
            dy2 = self._dy_buf(c2)
            dm = self._buf("dm", c2.y.shape)
            n += self._bn_bwd(c2, d, True, dy2, dmasked=dm)            # d' = d*[out>0] also feeds the shortcut
            dxd = None
            if ds is not None:                                          # shortcut branch: needs only d'; joins at the residual add
                dyd = self._dy_buf(ds, "dyd")
                dxd = self._buf("dxd", x_in.shape)
                aux = self._fork_aux()
                with torch.cuda.stream(aux if aux is not None else torch.cuda.current_stream()):
                    n += self._bn_bwd(ds, dm, False, dyd, ws=self.bn_ws_aux)
                    n += self._wgrad(ds, x_in, dyd, ws=self.wgrad_ws_aux)
                    n += self._dgrad(ds, dyd, dxd, scratch="up_aux")
            n += self._wgrad(c2, a1, dy2)
            da1 = self._buf("da1", a1.shape)
            n += self._dgrad(c2, dy2, da1)
            dy1 = self._dy_buf(c1)
            n += self._bn_bwd(c1, da1, True, dy1)
            n += self._wgrad(c1, x_in, dy1)
            dx = self._buf(f"d{parity}", x_in.shape)
            if ds is not None:
                self._join_aux()
                n += self._dgrad(c1, dy1, dx, residual=dxd)
            else:
                n += self._dgrad(c1, dy1, dx, residual=dm)
            d, parity = dx, parity ^ 1
        if part == "late":
            # the hand-off gradient gets its own buffer: the early graph is executed more than once per capture (warm-up, capture,
            # replay) and its ping-pong scratch would otherwise overwrite its own input
            dh = self._buf("dh", d.shape)
            dh.copy_(d)
            self._bwd_state = (dh, parity, lo)
            self._join_wgrad_stream()
            return n + 1
        da0 = self._buf("da0", self.a0.shape)
        if self.maxpool_idx:
            ops.maxpool3x3s2_bwd_idx(d, self.pool_idx, self.a0.shape[1], self.a0.shape[2], dx=da0)
        else:
            ops.maxpool3x3s2_bwd(d, self.a0, dx=da0, idx_ws=self.pool_idx)
            n += 1
        dy0 = self._dy_buf(st)
        n += 1 + self._bn_bwd(st, da0, True, dy0)
        self._on_wgrad_stream(lambda: ops.stem_wgrad(self.x, dy0, self._g(net.conv1.weight), ws=self.wgrad_ws))
        self._join_wgrad_stream()
        n += 2
        return n

    def _key(self):
        # sigma is passed by value to hk_bce_fwd_bwd, i.e. baked into a captured graph: a new sigma needs a new capture
        return (tuple(p.data_ptr() for p in self.params), None if self.target is None else (self.target.data_ptr(), self.target.dtype),
                self.sigma)

    def forward_backward(self, img: torch.Tensor, uv: Optional[torch.Tensor] = None, target: Optional[torch.Tensor] = None) -> torch.Tensor:
        """img (B,3,H,W) fp32 CUDA; labels uv (B,K,2) = (x,y) [Gaussian targets generated on the fly, dataset.py:36-44] or a
        (B,K,H,W) fp64/fp32 target tensor (what KeypointsDataset returns).  Returns the fp64 mean-BCE loss (0-dim, a view of an
        engine buffer); the gradients of all parameters are in `self.flat_grad` / `grad(p)`."""
        if (uv is None) == (target is None):
            raise ValueError("pass exactly one of uv / target")
        if not img.is_cuda or tuple(img.shape) != (self.B, 3, self.H, self.W):
            raise ValueError(f"expected a CUDA image batch of shape {(self.B, 3, self.H, self.W)}, got {tuple(img.shape)} on {img.device}")
        self.model.mark_weights_changed()   # BatchNorm running stats are advanced through raw pointers (tensor versions do not move)
        with torch.no_grad():
            self.x.copy_(img)
            if uv is not None:
                self.uv.copy_(uv.reshape(self.B, self.K, 2))
                self.target = None
            else:
                if tuple(target.shape) != (self.B, self.K, self.H, self.W) or target.dtype not in (torch.float64, torch.float32):
                    raise ValueError("target must be (B,K,H,W) float64 or float32")
                if self.target is None or self.target.dtype != target.dtype:
                    self.target = torch.empty_like(target, memory_format=torch.contiguous_format)
                self.target.copy_(target)
            if not self.use_cuda_graph:
                self.launches = self._enqueue()
                return self.loss
            key = self._key()
            if self.graph is None or key != self._graph_key:
                # one eager step (lazy function attributes, scratch allocation), restoring the BN buffers it advanced
                saved = [(c.bn.running_mean.clone(), c.bn.running_var.clone(), c.bn.num_batches_tracked.clone()) for c in self.convs]
                self.launches = self._enqueue()
                torch.cuda.current_stream().synchronize()
                for c, (m, v, t) in zip(self.convs, saved):
                    c.bn.running_mean.copy_(m); c.bn.running_var.copy_(v); c.bn.num_batches_tracked.copy_(t)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._enqueue()
                self.graph, self._graph_key = g, key
            self.graph.replay()
        return self.loss

    # ------------------------------------------------------------------ split step: the caller owns the loss (unmodified train.py)
    def _run(self, name: str, fn, name_key: Optional[str] = None) -> None:
        """Eager, or capture-once / replay of one of the split graphs.  `name` "fwd" marks graphs that advance the BN buffers."""
        if name == "fwd":
            self.model.mark_weights_changed()   # BatchNorm running stats move (raw-pointer writes)
        gname = name_key or name
        counts = self.__dict__.setdefault("_split_launches", {})
        if not self.use_cuda_graph:
            counts[gname] = fn()
            return
        key = self._key()
        graphs = self.__dict__.setdefault("_split_graphs", {})
        ent = graphs.get(gname)
        if ent is None or ent[1] != key:
            if name == "fwd":  # the warm-up run advances the BN buffers once more than the captured replay: restore them
                saved = [(c.bn.running_mean.clone(), c.bn.running_var.clone(), c.bn.num_batches_tracked.clone()) for c in self.convs]
            counts[gname] = fn()
            torch.cuda.current_stream().synchronize()
            if name == "fwd":
                for c, (m, v, t) in zip(self.convs, saved):
                    c.bn.running_mean.copy_(m); c.bn.running_var.copy_(v); c.bn.num_batches_tracked.copy_(t)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            graphs[gname] = ent = (g, key)
        ent[0].replay()

    def forward_heatmaps(self, img: torch.Tensor) -> torch.Tensor:
        """Train-mode forward only: (B,3,H,W) -> sigmoid heatmaps (B,K,H,W) fp32 (model.py:19-22 under model.train()); activations
        are kept for backward_from_heatmap_grad().  Returns a view of an engine buffer."""
        if not img.is_cuda or tuple(img.shape) != (self.B, 3, self.H, self.W):
            raise ValueError(f"expected a CUDA image batch of shape {(self.B, 3, self.H, self.W)}, got {tuple(img.shape)} on {img.device}")
        if "heat" not in self.__dict__:
            self.heat = torch.empty_like(self.logits_up)
            self.g_heat = torch.empty_like(self.logits_up)
        n = self.logits_up.numel()

        def fn():
            k = self._enqueue_forward()
            check(lib().hk_sigmoid_fwd(ptr(self.logits_up), ptr(self.heat), C.c_longlong(n), stream_ptr()), "hk_sigmoid_fwd")
            return k + 1

        with torch.no_grad():
            self.x.copy_(img)
            self._run("fwd", fn)
        return self.heat

    def backward_from_heatmap_grad(self, grad_heat: torch.Tensor) -> None:
        """dL/dheat (B,K,H,W) fp32 -> all parameter gradients (self.flat_grad), through sigmoid, head and backbone."""
        n = self.logits_up.numel()

        def fn():
            check(lib().hk_sigmoid_bwd(ptr(self.heat), ptr(self.g_heat), ptr(self.g_up), C.c_longlong(n), stream_ptr()), "hk_sigmoid_bwd")
            return 1 + self._enqueue_backward()

        with torch.no_grad():
            self.g_heat.copy_(grad_heat)
            self._run("bwd", fn)

    # ------------------------------------------------------------------ two-phase step for comm/compute overlap
    def forward_backward_late(self, img: torch.Tensor, uv: torch.Tensor) -> torch.Tensor:
        """Phase 1 of the data-parallel step: forward + loss + backward of head, layer4, layer3 (gradients flat_grad[late_offset:])."""
        if not img.is_cuda or tuple(img.shape) != (self.B, 3, self.H, self.W):
            raise ValueError(f"expected a CUDA image batch of shape {(self.B, 3, self.H, self.W)}")

        def fn():
            return self._enqueue_forward() + self._enqueue_loss() + self._enqueue_backward("late")

        with torch.no_grad():
            self.x.copy_(img)
            self.uv.copy_(uv.reshape(self.B, self.K, 2))
            self.target = None
            self._run("fwd", fn, name_key="late")
        return self.loss

    def backward_early(self) -> None:
        """Phase 2: backward of layer2, layer1 and the stem (gradients flat_grad[:late_offset])."""
        with torch.no_grad():
            self._run("bwd_early", lambda: self._enqueue_backward("early"))

    def split_launches(self) -> int:
        """Kernel launches of the split graphs captured so far (two-phase data-parallel step: "late" + "bwd_early")."""
        return int(sum(v or 0 for v in self.__dict__.get("_split_launches", {}).values()))

    def grad(self, p: torch.nn.Parameter) -> torch.Tensor:
        return self.grad_of[id(p)]

    def grads_into_params(self) -> None:
        """p.grad = the engine's gradient (views of the flat buffer) for every parameter."""
        for p in self.params:
            p.grad = self.grad_of[id(p)]
