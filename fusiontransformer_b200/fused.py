"""Fused execution of the reference's ``Conv3d -> BatchNorm -> ReLU (-> + shortcut -> ReLU)`` module chains.

The reference model code (FusionTransformer/models/spvcnn.py:22-78) strings torchsparse modules together with
``nn.Sequential``; every link is a separate pass over the activation.  ``fuse(model)`` leaves the modules, their
parameters and state_dict untouched but re-routes the forward of

  * every ``nn.Sequential`` containing ``Conv3d, BatchNorm[, ReLU]`` runs (BasicConvolutionBlock :22-35,
    BasicDeconvolutionBlock :38-50, the stem :87-93, ResidualBlock.net / .downsample :57-75), and
  * every ResidualBlock-shaped module (``net``, ``downsample``, ``relu``; forward :77-79)

through ``conv_bn_act``: one autograd node per convolution whose forward is
    pair-major tcgen05 GEMM  ->  sorted scatter + BatchNorm statistics  ->  normalise/affine/(+shortcut)/ReLU
and which hands the next convolution its bf16 operand (``SparseTensor.F16``) instead of re-reading fp32.
Backward is  BN-backward reduce -> BN-backward apply (emits the bf16 dgrad/wgrad operand) -> dgrad -> wgrad.
Numerics are those of the unfused path (same kernels, same bf16 rounding points), checked by
tests/test_gpu_fused.py against the oracle's module-by-module execution.
"""
from __future__ import annotations

import types

import torch
from torch import nn

from . import conv_engine, ops
from . import nn as spnn
from .functional import conv_geometry
from .sparse_tensor import SparseTensor

__all__ = ["conv_bn_act", "fuse", "unfuse", "join_side_streams", "weight_packer"]


class _SideStream:
    """wgrad runs on a second stream: it only feeds the optimizer, so it need not sit on the dgrad critical path of
    the backward pass.  One side stream per device; ``join()`` (called by dp.GradSync before it reads gradients)
    makes the main stream wait for everything queued on it.  Operand tensors are kept referenced until the join, so
    the caching allocator cannot hand their memory to a main-stream kernel that runs before the wgrad has read it
    (cheaper than record_stream, and legal under CUDA-graph capture)."""
    streams = {}
    keep = []
    dirty = set()

    @classmethod
    def fork(cls, device):
        """-> raw cudaStream_t of the side stream, ordered after everything queued on the current stream so far."""
        st = cls.streams.get(device.index)
        if st is None:
            st = cls.streams[device.index] = torch.cuda.Stream(device=device)
        st.wait_stream(torch.cuda.current_stream(device))
        cls.dirty.add(device.index)
        return st.cuda_stream

    @classmethod
    def join(cls):
        for idx in list(cls.dirty):
            torch.cuda.current_stream(idx).wait_stream(cls.streams[idx])
        cls.dirty.clear()
        cls.keep.clear()


class _BranchStream:
    """The shortcut convolution of a ResidualBlock (kernel 1 conv + BatchNorm, models/spvcnn.py:69-75) does not depend
    on the block's main branch: it runs on a second stream, forward and (autograd replays a node on its forward stream)
    backward, and overlaps the main branch's conv chain.  ``FT3D_BRANCH_STREAM=0`` keeps everything on one stream."""
    streams = {}
    used = set()

    @classmethod
    def enabled(cls):
        import os
        return os.environ.get("FT3D_BRANCH_STREAM", "1") != "0"

    @classmethod
    def get(cls, device):
        st = cls.streams.get(device.index)
        if st is None:
            st = cls.streams[device.index] = torch.cuda.Stream(device=device)
        cls.used.add(device.index)
        return st


def join_side_streams():
    for idx in list(_BranchStream.used):
        torch.cuda.current_stream(idx).wait_stream(_BranchStream.streams[idx])
    _BranchStream.used.clear()
    _SideStream.join()


def _role(transpose: bool, grad: bool) -> str:
    if grad:
        return "dgrad_transposed" if transpose else "dgrad"
    return "transposed" if transpose else "forward"


class _ConvBNAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, kernel, gamma, beta, res, x16, kmap, transpose, relu, bn, skip=None, skip16=None):
        ctx.set_materialize_grads(False)         # no zero-filled gradient for the (non-differentiable) bf16 output
        training = bn.training or bn.running_mean is None
        rm, rv = (bn.running_mean, bn.running_var) if (bn.training and bn.track_running_stats) else (None, None)
        cin, cout = kernel.shape[-2], kernel.shape[-1]
        tc = conv_engine.pairs_ok(cin, cout)
        stat = None
        if tc:
            if x16 is None:
                x16 = ops.to_bf16(feats)
            if kmap is None:                                   # kernel_size 1: dense GEMM, rows are already final
                y = conv_engine.dense_conv(x16, kernel, w_transposed=False)
            elif conv_engine.os_enabled():                     # gather-GEMM, final rows and the BN statistics
                y, stat = conv_engine.os_conv(x16, kmap, kernel, _role(transpose, False),
                                              bn=(bn.eps, bn.momentum, rm, rv) if training else None)
            else:
                partial, ppos, ncols = conv_engine.pairs_partial(x16, kmap, kernel, _role(transpose, False))
                if training:
                    y, stat = ops.conv_reduce_bn(partial, ppos, ncols, bn.eps, bn.momentum, rm, rv)
                else:
                    y = ops.conv_reduce(partial, ppos, ncols)
            saved_in = x16
        else:
            if kmap is None:
                y = feats.matmul(kernel)
            else:
                table = kmap.nbrT if transpose else kmap.nbr
                y = conv_engine.gather_conv(feats, table, kmap, kernel, kflip=False, w_transposed=False)
            saved_in = feats
        if stat is None:
            if training:
                stat = ops.bn_stats(y, bn.eps, bn.momentum, rm, rv)
            else:
                stat = torch.stack([bn.running_mean, torch.rsqrt(bn.running_var + bn.eps)]).contiguous()
        # skip: torchsparse.cat([this block's output, skip]) (models/spvcnn.py:212-228) -- the normalise pass writes
        # into the left columns of the concatenation buffers, one copy launch fills the right columns
        extra = 0 if skip is None else skip.shape[1]
        z, z16 = ops.bn_apply(y, stat, gamma, beta, res, relu, want_f32=True, want_bf16=tc, extra_cols=extra)
        if extra:
            ops.copy_cols(skip, skip16, z, z16, cout)
        ctx.extra = extra
        ctx.kmap, ctx.transpose, ctx.relu, ctx.tc, ctx.training = kmap, transpose, relu, tc, training
        ctx.has_res = res is not None
        mask = (z16 if tc else z) if relu else None
        ctx.save_for_backward(saved_in, kernel, gamma, y, stat, mask)
        ctx.beta = beta
        if z16 is None:
            z16 = z.new_empty(0)
        ctx.mark_non_differentiable(z16)
        return z, z16

    @staticmethod
    def backward(ctx, gz, _unused):
        if gz is None:
            return (None,) * 12
        saved_in, kernel, gamma, y, stat, mask = ctx.saved_tensors
        beta = ctx.beta
        kmap, transpose, tc = ctx.kmap, ctx.transpose, ctx.tc
        gz = gz.contiguous()
        m16, m32 = (mask, None) if (mask is not None and mask.dtype == torch.bfloat16) else (None, mask)
        # Gradient arena (dp.GradSync): parameter gradients are accumulated by the kernels directly into the flat
        # buffer `.grad` views -- no zero-filled temporaries, no autograd accumulation launches.
        sink = getattr(kernel, "_ft3d_sink", None)
        if sink is not None and not (sink.owns(kernel) and sink.owns(gamma) and sink.owns(beta)):
            sink = None
        need_in, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        want_res = ctx.has_res and ctx.needs_input_grad[4] and mask is not None
        ldg = gz.shape[1] if ctx.extra else 0          # gz and the saved output are [n, cout + extra]: row pitch
        if sink is not None:
            red, dgamma, dbeta = ops.bn_bwd_reduce(gz, y, m16, m32, stat, gamma.grad, beta.grad, ldg=ldg)
        else:
            red, dgamma, dbeta = ops.bn_bwd_reduce(gz, y, m16, m32, stat, ldg=ldg)
        gy, gy16, gres = ops.bn_bwd_apply(gz, y, m16, m32, stat, gamma, red if ctx.training else None,
                                          want_f32=not tc, want_bf16=tc, want_res=want_res, ldg=ldg)
        gskip = gz[:, kernel.shape[-1]:] if (ctx.extra and ctx.needs_input_grad[10]) else None
        if ctx.has_res and ctx.needs_input_grad[4] and mask is None:
            gres = gz
        gin = gw = None
        cin, cout = kernel.shape[-2], kernel.shape[-1]
        if tc:
            if kmap is None:
                if need_in:
                    gin = conv_engine.dense_conv(gy16, kernel, w_transposed=True)
                if need_w:
                    gw = conv_engine.dense_wgrad(saved_in, gy16, cin, cout, into=kernel.grad if sink else None)
            else:
                if need_in and conv_engine.os_enabled():
                    gin = conv_engine.os_conv(gy16, kmap, kernel, _role(transpose, True))[0]
                elif need_in:
                    gin = conv_engine.pairs_conv(gy16, kmap, kernel, _role(transpose, True))
                if need_w and sink is not None and sink.side_wgrad:
                    # gradient goes straight into the arena, nothing downstream in autograd consumes it: run it
                    # beside the dgrad chain
                    raw = _SideStream.fork(gz.device)
                    conv_engine.pairs_wgrad(saved_in, gy16, kmap, cin, cout, transpose, into=kernel.grad, stream=raw)
                    _SideStream.keep.append((saved_in, gy16))
                elif need_w:
                    gw = conv_engine.pairs_wgrad(saved_in, gy16, kmap, cin, cout, transpose,
                                                 into=kernel.grad if sink else None)
            if sink is not None and need_w:
                gw = None
                sink.note(kernel)
        else:
            if kmap is None:
                if need_in:
                    gin = gy.matmul(kernel.t())
                if need_w:
                    gw = saved_in.t().matmul(gy)
            else:
                if need_in:
                    if transpose:
                        table, kflip = kmap.nbr, False
                    elif kmap.symmetric:
                        table, kflip = kmap.nbr, True
                    else:
                        table, kflip = kmap.nbrT, False
                    gin = conv_engine.gather_conv(gy, table, kmap, kernel, kflip=kflip, w_transposed=True)
                if need_w:
                    gw = conv_engine.wgrad(saved_in, gy, kmap, cin, cout, transpose)
        if sink is not None:
            sink.note(gamma)
            sink.note(beta)
        return gin, gw, dgamma, dbeta, gres, None, None, None, None, None, gskip, None


def _bn_forward(y, bn, training, rm, rv, gamma, beta, relu):
    """BatchNorm(+ReLU) of a plain fp32 [N,C] tensor -> (z, stat)."""
    if training:
        stat = ops.bn_stats(y, bn.eps, bn.momentum, rm, rv)
    else:
        stat = torch.stack([bn.running_mean, torch.rsqrt(bn.running_var + bn.eps)]).contiguous()
    z, _ = ops.bn_apply(y, stat, gamma, beta, None, relu, want_f32=True, want_bf16=False)
    return z, stat


def _bn_backward(gz, y, mask, stat, gamma, beta, sink, training, want_bf16, want_f32=True):
    """-> (gy f32 | None, gy16 | None, dgamma, dbeta); with a gradient arena the parameter gradients are added in place."""
    if sink is not None:
        red, dgamma, dbeta = ops.bn_bwd_reduce(gz, y, None, mask, stat, gamma.grad, beta.grad)
    else:
        red, dgamma, dbeta = ops.bn_bwd_reduce(gz, y, None, mask, stat)
    gy, gy16, _ = ops.bn_bwd_apply(gz, y, None, mask, stat, gamma, red if training else None, want_f32=want_f32,
                                   want_bf16=want_bf16, want_res=False)
    return gy, gy16, dgamma, dbeta


class _BNAct(torch.autograd.Function):
    """``relu?(batch_norm(y))`` on a plain [N,C] tensor: the point-branch ``Linear -> BatchNorm1d -> ReLU`` blocks
    (models/spvcnn.py:164-180, middle_fusion.py:18-22) on the same kernels as the voxel branch."""

    @staticmethod
    def forward(ctx, y, gamma, beta, relu, bn):
        training = bn.training or bn.running_mean is None
        rm, rv = (bn.running_mean, bn.running_var) if (bn.training and bn.track_running_stats) else (None, None)
        y = y.contiguous()
        z, stat = _bn_forward(y, bn, training, rm, rv, gamma, beta, relu)
        ctx.training, ctx.beta = training, beta
        ctx.save_for_backward(y, gamma, stat, z if relu else None)
        return z

    @staticmethod
    def backward(ctx, gz):
        y, gamma, stat, mask = ctx.saved_tensors
        beta = ctx.beta
        gz = gz.contiguous()
        sink = getattr(gamma, "_ft3d_sink", None)
        if sink is not None and not (sink.owns(gamma) and sink.owns(beta)):
            sink = None
        gy, _, dgamma, dbeta = _bn_backward(gz, y, mask, stat, gamma, beta, sink, ctx.training, want_bf16=False)
        if sink is not None:
            sink.note(gamma)
            sink.note(beta)
        return gy, dgamma, dbeta, None, None


class _LinearBNAct(torch.autograd.Function):
    """``relu?(batch_norm(x @ W^T + b))`` for the point-branch / fusion MLPs (models/spvcnn.py:164-180,
    middle_fusion.py:18-22) as one autograd node.  Forward and dX are library GEMMs (cuBLAS, as SURVEY 8(a14) allows);
    the weight gradient -- a [out,in] result reduced over ~50k points, a shape cuBLAS serves with a 70 us split-K
    kernel -- runs on the tcgen05 wgrad kernel (bf16 operands, fp32 accumulation) in tensor-core mode, and the bias
    gradient is a deterministic column sum; both go straight into the gradient arena when there is one."""

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, relu, bn):
        training = bn.training or bn.running_mean is None
        rm, rv = (bn.running_mean, bn.running_var) if (bn.training and bn.track_running_stats) else (None, None)
        x = x.contiguous()
        out_f, in_f = weight.shape
        tc = conv_engine.pairs_ok(in_f, out_f)
        if tc:       # a14 on the tcgen05 kernels: identity-gather GEMM, bf16 operands, fp32 accumulation
            x = ops.to_bf16(x)
            y = ops.conv_pairs_tc(x, None, None, 1, 0, x.shape[0], weight.detach().unsqueeze(0), True, owner=weight)
            if bias is not None:
                y += bias
        else:
            y = torch.addmm(bias, x, weight.t()) if bias is not None else x.matmul(weight.t())
        z, stat = _bn_forward(y, bn, training, rm, rv, gamma, beta, relu)
        ctx.training, ctx.beta, ctx.bias, ctx.tc = training, beta, bias, tc
        ctx.save_for_backward(x, weight, y, gamma, stat, z if relu else None)      # tc: x is the bf16 copy
        return z

    @staticmethod
    def backward(ctx, gz):
        x, weight, y, gamma, stat, mask = ctx.saved_tensors
        beta, bias = ctx.beta, ctx.bias
        gz = gz.contiguous()
        out_f, in_f = weight.shape
        sink = getattr(weight, "_ft3d_sink", None)
        if sink is not None and not (sink.owns(weight) and sink.owns(gamma) and sink.owns(beta)
                                     and (bias is None or sink.owns(bias))):
            sink = None
        tc = ctx.tc
        want_gb = bias is not None and ctx.needs_input_grad[2]
        gy, gy16, dgamma, dbeta = _bn_backward(gz, y, mask, stat, gamma, beta, sink, ctx.training, want_bf16=tc,
                                               want_f32=(not tc) or want_gb)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            if tc:       # gX = gy W : the same identity-gather GEMM with the untransposed image
                gx = ops.conv_pairs_tc(gy16, None, None, 1, 0, gy16.shape[0], weight.detach().unsqueeze(0), False,
                                       owner=weight)
            else:
                gx = gy.matmul(weight)
        if ctx.needs_input_grad[1]:
            if tc:       # gW[out,in] = gy^T x : the wgrad kernel with (a, b) = (gy, x) and an identity pair list
                gw = ops.conv_wgrad_pairs_tc(gy16, x, None, None, 1, 0, out_f, in_f, x.shape[0],
                                             into=weight.grad.view(1, out_f, in_f) if sink is not None else None)
                gw = None if sink is not None else gw.view(out_f, in_f)
            elif sink is not None:
                weight.grad.addmm_(gy.t(), x)
            else:
                gw = gy.t().matmul(x)
        if want_gb:
            gb = ops.col_sum(gy, into=bias.grad if sink is not None else None)
            if sink is not None:
                gb = None
        if sink is not None:
            for p in (weight, bias, gamma, beta):
                if p is not None:
                    sink.note(p)
        return gx, gw, gb, dgamma, dbeta, None, None


def linear_bn_act(x: torch.Tensor, lin, bn, relu: bool) -> torch.Tensor:
    out = _LinearBNAct.apply(x, lin.weight, lin.bias, bn.weight, bn.bias, relu, bn)
    if bn.training and bn.track_running_stats:
        bn._ft3d_pending_batches = getattr(bn, "_ft3d_pending_batches", 0) + 1
    return out


def bn_act(y: torch.Tensor, bn, relu: bool) -> torch.Tensor:
    out = _BNAct.apply(y, bn.weight, bn.bias, relu, bn)
    if bn.training and bn.track_running_stats:
        bn._ft3d_pending_batches = getattr(bn, "_ft3d_pending_batches", 0) + 1
    return out


def _bn1d_fusable(bn) -> bool:
    return (type(bn) is nn.BatchNorm1d and bn.affine and bn.momentum is not None and bn.num_features % 4 == 0)


def _fusable(conv, bn) -> bool:
    return (isinstance(conv, spnn.Conv3d) and isinstance(bn, nn.BatchNorm1d) and conv.bias is None and bn.affine
            and bn.momentum is not None and conv.out_channels % 4 == 0 and conv.d == 1)


def conv_bn_act(x: SparseTensor, conv, bn, relu: bool, res: torch.Tensor | None = None,
                skip: SparseTensor | None = None) -> SparseTensor:
    """``relu?(bn(conv(x)) [+ res])`` as one autograd node; ``res`` is an fp32 [N_out, Cout] tensor.  With ``skip`` the
    result is ``torchsparse.cat([that, skip])``: [N_out, Cout + skip channels]."""
    kmap, make_out = conv_geometry(x, conv.ks, conv.s, conv.d, conv.t)
    feats = x.F
    if not feats.is_contiguous():
        feats = feats.contiguous()
    x16 = x.F16
    if x16 is not None and (x16.shape != feats.shape or x16.dtype != torch.bfloat16):
        x16 = None
    if skip is None:
        z, z16 = _ConvBNAct.apply(feats, conv.kernel, bn.weight, bn.bias, res, x16, kmap, conv.t, relu, bn)
    else:
        s16 = skip.F16
        if s16 is not None and (s16.shape != skip.F.shape or s16.dtype != torch.bfloat16):
            s16 = None
        z, z16 = _ConvBNAct.apply(feats, conv.kernel, bn.weight, bn.bias, res, x16, kmap, conv.t, relu, bn,
                                  skip.F.contiguous(), s16)
    if bn.training and bn.track_running_stats:
        bn._ft3d_pending_batches = getattr(bn, "_ft3d_pending_batches", 0) + 1
    out = make_out(z)
    if z16.numel():
        out.F16 = z16
    return out


# ------------------------------------------------------------------------------------------------ module re-routing
def _flush_batches(module, *_):
    """num_batches_tracked is kept on the host between state_dict() calls (one fewer launch per BatchNorm per step)."""
    n = getattr(module, "_ft3d_pending_batches", 0)
    if n and module.num_batches_tracked is not None:
        module.num_batches_tracked += n
    module._ft3d_pending_batches = 0


def _sequential_forward(self, x):
    mods = list(self._modules.values())
    i, n = 0, len(mods)
    while i < n:
        m = mods[i]
        if (isinstance(x, SparseTensor) and i + 1 < n and _fusable(m, mods[i + 1]) and x.F.is_cuda):
            relu = i + 2 < n and type(mods[i + 2]) is spnn.ReLU
            x = conv_bn_act(x, m, mods[i + 1], relu)
            i += 3 if relu else 2
        elif (isinstance(x, torch.Tensor) and x.is_cuda and x.dim() == 2 and x.dtype == torch.float32
              and type(m) is nn.Linear and i + 1 < n and _bn1d_fusable(mods[i + 1])
              and m.in_features % 4 == 0 and m.out_features % 4 == 0):
            relu = i + 2 < n and type(mods[i + 2]) is nn.ReLU
            x = linear_bn_act(x, m, mods[i + 1], relu)
            i += 3 if relu else 2
        elif (isinstance(x, torch.Tensor) and x.is_cuda and x.dim() == 2 and x.dtype == torch.float32
              and _bn1d_fusable(m)):
            relu = i + 1 < n and type(mods[i + 1]) is nn.ReLU
            x = bn_act(x, m, relu)
            i += 2 if relu else 1
        else:
            x = m(x)
            i += 1
    return x


def _is_residual_block(m) -> bool:
    net, ds, relu = getattr(m, "net", None), getattr(m, "downsample", None), getattr(m, "relu", None)
    if not (isinstance(net, nn.Sequential) and isinstance(ds, nn.Sequential) and type(relu) is spnn.ReLU):
        return False
    nm, dm = list(net), list(ds)
    if len(nm) != 5 or not (_fusable(nm[0], nm[1]) and type(nm[2]) is spnn.ReLU and _fusable(nm[3], nm[4])):
        return False
    return len(dm) == 0 or (len(dm) == 2 and _fusable(dm[0], dm[1]))


def _residual_forward(self, x):
    if not (isinstance(x, SparseTensor) and x.F.is_cuda):
        return self.relu(self.net(x) + self.downsample(x))
    nm, dm = list(self.net), list(self.downsample)
    if dm and _BranchStream.enabled():
        main = torch.cuda.current_stream(x.F.device)
        side = _BranchStream.get(x.F.device)
        side.wait_stream(main)                       # x is ready
        with torch.cuda.stream(side):
            shortcut = conv_bn_act(x, dm[0], dm[1], False).F.contiguous()
        h = conv_bn_act(x, nm[0], nm[1], True)
        main.wait_stream(side)
        # allocated on the side stream, read on the main stream: keep it referenced until the streams are joined at the
        # end of the step so that the caching allocator cannot recycle it under the reader (same rule as the wgrad fork)
        _SideStream.keep.append(shortcut)
        return conv_bn_act(h, nm[3], nm[4], True, res=shortcut)
    h = conv_bn_act(x, nm[0], nm[1], True)
    shortcut = x.F if not dm else conv_bn_act(x, dm[0], dm[1], False).F
    return conv_bn_act(h, nm[3], nm[4], True, res=shortcut.contiguous())


def deconv_cat(block, y: SparseTensor, skip: SparseTensor) -> SparseTensor:
    """``torchsparse.cat([block(y), skip])`` for a Basic(De)ConvolutionBlock (models/spvcnn.py:212,216,224,228): when the
    block is fused, its BatchNorm epilogue writes straight into the concatenation buffer."""
    net = getattr(block, "net", None)
    mods = list(net) if isinstance(net, nn.Sequential) else []
    if (len(mods) == 3 and _fusable(mods[0], mods[1]) and type(mods[2]) is spnn.ReLU and y.F.is_cuda
            and getattr(net, "_ft3d_fused", False) and skip.F.shape[1] % 4 == 0
            and skip.F.dtype == torch.float32 and y.F.dtype == torch.float32):
        return conv_bn_act(y, mods[0], mods[1], True, skip=skip)
    from . import cat
    return cat([block(y), skip])


def weight_packer(model: nn.Module):
    """ops.WeightPacker over every tensor-core conv kernel of ``model``: call ``.pack()`` once per step (after the
    optimizer update) instead of one pack launch per layer and orientation."""
    ks = [m.kernel for m in model.modules() if isinstance(m, spnn.Conv3d)
          and conv_engine.pairs_ok(m.kernel.shape[-2], m.kernel.shape[-1])]
    # nn.Linear weights [out,in] of the point-branch / fusion MLPs (a14), read as a [1, out, in] kernel
    ks += [m.weight for m in model.modules() if type(m) is nn.Linear
           and conv_engine.pairs_ok(m.out_features, m.in_features)]
    return ops.WeightPacker(ks)


def fuse(model: nn.Module) -> nn.Module:
    """Re-route the conv/BN/ReLU chains of ``model`` (in place) through the fused kernels; returns ``model``."""
    for m in model.modules():
        if getattr(m, "_ft3d_fused", False):
            continue
        if _is_residual_block(m):
            m.forward = types.MethodType(_residual_forward, m)
            m._ft3d_fused = True
        elif isinstance(m, nn.Sequential):
            mods = list(m)
            if any(_fusable(a, b) for a, b in zip(mods, mods[1:])) or any(_bn1d_fusable(a) for a in mods):
                m.forward = types.MethodType(_sequential_forward, m)
                m._ft3d_fused = True
        if isinstance(m, nn.BatchNorm1d) and not getattr(m, "_ft3d_hooked", False):
            m.register_state_dict_pre_hook(_flush_batches)
            m._ft3d_hooked = True
    return model


def unfuse(model: nn.Module) -> nn.Module:
    for m in model.modules():
        if getattr(m, "_ft3d_fused", False):
            del m.forward
            m._ft3d_fused = False
        if isinstance(m, nn.BatchNorm1d):
            _flush_batches(m)
    return model
