"""TEST INFRASTRUCTURE: torch-CPU stand-ins for the `sniper_b200.ops` entry points that `model_mnv2` (and `model.Conv`)
call, each restating the documented contract of the C-ABI function in plain float64 torch (MultiProposalTarget and
DeformablePSROIPooling go through the C oracle).  tests/test_mnv2_wiring_cpu.py patches them over `sniper_b200.ops`
to execute the hand-scheduled forward / backward of the model on the CPU against the autograd oracle: that pins the
ORCHESTRATION (operand order, saved tensors, residual routing, channel padding, loss normalisation, checkpoint layout
mapping) in a container without a GPU.  The kernels themselves are pinned by the -m gpu tests.  Never imported by the
product path.
"""
import numpy as np
import torch
import torch.nn.functional as F

import oracle_lib as O

NCHW, NHWC = 0, 1


def _nchw(x):
    return x.permute(0, 3, 1, 2)


def _nhwc(x):
    return x.permute(0, 2, 3, 1)


# Storage semantics.  Every stand-in computes in float64 ("exact arithmetic") and rounds ONCE to the dtype the product
# stores the result in: bf16 tensors through float32 (the kernels accumulate in fp32 and convert; oracle/torch_graph._bf16
# does the same), float32 tensors to float32, float64 (the default dtype of the sharp tests) not at all.
def _d(t):
    return None if t is None else t.double()


# TF32[0] = True: the operands of every tensor-core contraction (gemm_nt, conv2d_nhwc, conv2d_wgrad_nhwc) that are NOT
# bf16 are reduced the way `tcgen05.mma kind::tf32` reads fp32 words -- fp32, low 13 mantissa bits dropped -- as
# oracle/torch_graph.py's MODE "tf32" does.
TF32 = [False]


def _mma(t):
    if t is None:
        return None
    if TF32[0] and t.dtype != torch.bfloat16:
        i = t.detach().to(torch.float32).contiguous().view(torch.int32)
        return (i & -8192).view(torch.float32).double()
    return t.double()


def _st(y, dtype):
    if dtype == torch.bfloat16:
        return y.float().to(torch.bfloat16)
    return y.to(dtype)


def _emit(y, out, dtype=None):
    if out is None:
        return _st(y, dtype if dtype is not None else y.dtype).contiguous()
    out.copy_(_st(y, out.dtype))
    return out


def _stats(y, stats, dtype):
    """sum / sum of squares of the STORED values"""
    if stats is not None:
        C = y.shape[-1]
        f = _st(y, dtype).reshape(-1, C).double()
        stats[:C] += f.sum(0)
        stats[C:2 * C] += (f * f).sum(0)


def _epi(y, scale, bias, residual, relu):
    if scale is not None:
        y = y * scale
    if bias is not None:
        y = y + bias
    if residual is not None:
        y = y + residual
    if relu:
        y = y.clamp(min=0)
    return y


def gemm_nt(a, b, *, out=None, scale=None, bias=None, residual=None, relu=False, accumulate=False, out_dtype=None,
            stats=None):
    y = _epi(_mma(a) @ _mma(b).t(), _d(scale), _d(bias), _d(residual), relu)
    if accumulate:
        y = y + _d(out)
    dt = out.dtype if out is not None else _exact(out_dtype or a.dtype)
    _stats(y, stats, dt)
    return _emit(y, out, dt)


def conv_taps(kh, kw, dil, pad):
    dh = [i * dil - pad for i in range(kh) for _ in range(kw)]
    dw = [j * dil - pad for _ in range(kh) for j in range(kw)]
    return dh, dw


def conv2d_nhwc(x, w, *, kh, kw, stride=1, dil=1, pad=0, out=None, scale=None, bias=None, residual=None, relu=False,
                accumulate=False, taps=None, out_hw=None, out_map=None, stats=None, out_dtype=None):
    """The documented contract of sniper_conv2d_nhwc: y[n,oh,ow,co] = sum_t sum_ci x[n, oh*stride + dh[t], ow*stride +
    dw[t], ci] * w[co, t*Cin + ci] (zero outside the map), epilogue *scale +bias +residual relu, written to the positions
    (oh*os + ooh, ow*os + oow) of an oH x oW map when out_map is given (the strided data gradients)."""
    NB, H, W, Cin = x.shape
    Cout = w.shape[0]
    dh, dw = taps if taps is not None else conv_taps(kh, kw, dil, pad)
    if out_hw is None:
        Ho = (H + 2 * pad - dil * (kh - 1) - 1) // stride + 1
        Wo = (W + 2 * pad - dil * (kw - 1) - 1) // stride + 1
    else:
        Ho, Wo = out_hw
    wd, xin = _mma(w), x
    x = _mma(x)
    scale, bias, residual = _d(scale), _d(bias), _d(residual)
    dt = out.dtype if (out is not None and out_map is None) else _exact(out_dtype or xin.dtype)
    y = x.new_zeros(NB, Ho, Wo, Cout)
    ah, aw = torch.arange(Ho) * stride, torch.arange(Wo) * stride
    for t in range(len(dh)):
        hi, wi = ah + dh[t], aw + dw[t]
        m = ((hi >= 0) & (hi < H))[:, None] & ((wi >= 0) & (wi < W))[None, :]
        xs = x[:, hi.clamp(0, H - 1)][:, :, wi.clamp(0, W - 1)] * m[None, :, :, None].to(x.dtype)
        y = y + xs @ wd[:, t * Cin:(t + 1) * Cin].t()
    if out_map is None:
        y = _epi(y, scale, bias, residual, relu)
        _stats(y, stats, dt)
        return _emit(y, out, dt)
    oH, oW, os_, ooh, oow = out_map
    res = None if residual is None else residual[:, ooh::os_, oow::os_][:, :Ho, :Wo]
    y = _epi(y, scale, bias, res, relu)
    assert out is not None and out.shape[1] == oH and out.shape[2] == oW
    _stats(y, stats, out.dtype)
    out[:, ooh::os_, oow::os_][:, :Ho, :Wo] = _st(y, out.dtype)
    return out


def conv2d_wgrad_nhwc(dy, x, *, kh, kw, stride=1, dil=1, pad=0, dw_out=None, splits=8, taps=None):
    Cout, Cin = dy.shape[3], x.shape[3]
    g = torch.nn.grad.conv2d_weight(_nchw(_mma(x)).contiguous(), (Cout, Cin, kh, kw), _nchw(_mma(dy)).contiguous(), stride, pad, dil)
    dw_out += g.permute(0, 2, 3, 1).reshape(Cout, kh * kw * Cin).to(dw_out.dtype)
    return dw_out


def weight_transpose_jobs(jobs, device):
    return list(jobs)


def weight_transpose_batched(table):
    for w, wt, sel, Cout, T, Cin in table:
        s = sel.long()
        wt.view(Cin, len(s), Cout).copy_(_st(_d(w).view(Cout, T, Cin)[:, s, :].permute(2, 1, 0), wt.dtype))


def bn_param_grad_jobs(states, device):
    return list(states)


def bn_param_grad_batched(table):
    for st in table:
        C = st.C
        if st.dbeta is not None:
            st.dbeta += st.sums[:C].to(st.dbeta.dtype)
            st.dgamma += st.sums[C:].to(st.dgamma.dtype)
        st.sums.zero_()
        st.sums_f.zero_()               # the forward statistics of the fused-apply path are cleared here


def _finish_stats(bn, mean, var, M, eps, momentum, fix_gamma, update_moving):
    g = torch.ones_like(bn.gamma).double() if fix_gamma else _d(bn.gamma)
    invstd = 1.0 / torch.sqrt(var + eps)
    bn.mean.copy_(mean); bn.invstd.copy_(invstd)
    bn.scale.copy_(g * invstd); bn.shift.copy_(_d(bn.beta) - mean * g * invstd)
    if update_moving:
        unb = var * M / (M - 1) if M > 1 else var
        bn.moving_mean.copy_(_d(bn.moving_mean) * momentum + mean * (1 - momentum))
        bn.moving_var.copy_(_d(bn.moving_var) * momentum + unb * (1 - momentum))


def bn_stats(x, bn, eps=2e-5, momentum=0.9, fix_gamma=False, update_moving=True):
    f = _d(x).reshape(-1, x.shape[-1])
    M = f.shape[0]
    _finish_stats(bn, f.mean(0), f.var(0, unbiased=False), M, eps, momentum, fix_gamma, update_moving)


def bn_finalize(bn, M, eps=2e-5, momentum=0.9, fix_gamma=False, update_moving=True):
    C = bn.C
    mean = bn.sums[:C] / M
    var = bn.sums[C:] / M - mean * mean
    _finish_stats(bn, mean, var.clamp(min=0), M, eps, momentum, fix_gamma, update_moving)
    bn.sums.zero_()


def bn_frozen(bn, eps=2e-5, fix_gamma=False):
    g = torch.ones_like(bn.gamma).double() if fix_gamma else _d(bn.gamma)
    sc = g / torch.sqrt(_d(bn.moving_var) + eps)
    bn.scale.copy_(sc)
    bn.shift.copy_(_d(bn.beta) - _d(bn.moving_mean) * sc)


def bn_apply_train(x, bn, eps=2e-5, momentum=0.9, relu=True, fix_gamma=False, update_moving=True, out=None):
    """finalisation from bn.sums_f (left as is) + apply, one launch in the product"""
    C = bn.C
    M = x.numel() // C
    mean = bn.sums_f[:C] / M
    var = (bn.sums_f[C:] / M - mean * mean).clamp(min=0)
    _finish_stats(bn, mean, var, M, eps, momentum, fix_gamma, update_moving)
    return affine_act(x, bn.scale, bn.shift, relu=relu, out=out)


def bn_relu_bwd(x, dy, bn, add=None, out=None, defer=False):
    return bn_act_bwd(x, dy, bn, 1, add=add, out=out, defer=defer)


def maxpool3x3s2(x):
    return _st(_nhwc(F.max_pool2d(_nchw(_d(x)), 3, 2, 1)), x.dtype).contiguous()


def stem_rows(w, dtype=None):
    """[64,7,7,3]: the stand-in stem convolves directly; bf16 rows in the mixed-precision configuration"""
    return _st(_d(w), torch.bfloat16) if dtype == torch.bfloat16 else w


def stem_conv_tc(x_nchw, rows, in_scale, in_shift, out_scale, out_shift, out_dtype=None):
    """bn_data -> conv0 7x7/2 pad 3 (zero padding of the NORMALISED image) -> bn0 -> relu, NHWC out.  The im2col buffer
    has the rows' dtype: bf16 rows = a bf16 im2col of bn_data(x); fp32 rows are read as TF32 by the GEMM."""
    xn = _d(x_nchw) * _d(in_scale).view(1, 3, 1, 1) + _d(in_shift).view(1, 3, 1, 1)
    if rows.dtype == torch.bfloat16:
        xn = _st(xn, torch.bfloat16).double()
    y = F.conv2d(_mma(xn) if rows.dtype != torch.bfloat16 else xn, _mma(rows).permute(0, 3, 1, 2), None, 2, 3)
    y = (y * _d(out_scale).view(1, -1, 1, 1) + _d(out_shift).view(1, -1, 1, 1)).clamp(min=0)
    return _st(_nhwc(y), _exact(out_dtype) if out_dtype is not None else x_nchw.dtype).contiguous()


stem_conv = stem_conv_tc


def _deform_cols(x, offset, dil, pad, dg):
    """Differentiable deformable im2col (deformable_im2col.cuh:78-113, 216-263), NHWC: x [N,H,W,C], offset
    [N,H,W,>=dg*18] with channel g*18 + 2*tap (+1) = (dy, dx) -> col [N,H,W,9,C] (same sampling rule as
    oracle/torch_graph.deform_conv2d)."""
    N, H, W, C = x.shape
    cpg = C // dg
    hh, ww = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    n_idx = torch.arange(N).view(N, 1, 1).expand(N, H, W)
    cols = []
    for tap in range(9):
        i, j = divmod(tap, 3)
        per_g = []
        for g in range(dg):
            oh, ow = offset[..., g * 18 + 2 * tap], offset[..., g * 18 + 2 * tap + 1]
            h_im = (hh - pad + i * dil).to(x.dtype) + oh
            w_im = (ww - pad + j * dil).to(x.dtype) + ow
            valid = (h_im >= 0) & (w_im >= 0) & (h_im < H) & (w_im < W)
            h_low = torch.floor(h_im).clamp(max=H - 1)
            w_low = torch.floor(w_im).clamp(max=W - 1)
            hc = torch.where(torch.floor(h_im) >= H - 1, h_low, h_im)
            wc = torch.where(torch.floor(w_im) >= W - 1, w_low, w_im)
            h_high, w_high = (h_low + 1).clamp(max=H - 1), (w_low + 1).clamp(max=W - 1)
            lh, lw = hc - h_low, wc - w_low
            xg = x[..., g * cpg:(g + 1) * cpg]
            at = lambda hi, wi: xg[n_idx, hi.long().clamp(0, H - 1), wi.long().clamp(0, W - 1)]
            v = ((1 - lh) * (1 - lw)).unsqueeze(-1) * at(h_low, w_low) + ((1 - lh) * lw).unsqueeze(-1) * at(h_low, w_high) \
                + (lh * (1 - lw)).unsqueeze(-1) * at(h_high, w_low) + (lh * lw).unsqueeze(-1) * at(h_high, w_high)
            per_g.append(v * valid.unsqueeze(-1).to(x.dtype))
        cols.append(torch.cat(per_g, -1))
    return torch.stack(cols, 3)


def deform_im2col(x, offset, *, kh=3, kw=3, stride=1, dil=1, pad=1, dgroups=4, out=None):
    N, H, W, C = x.shape
    return _st(_deform_cols(_d(x), _d(offset), dil, pad, dgroups).reshape(N * H * W, 9 * C), x.dtype)       # col has x's dtype


def deform_col2im(dcol, x, offset, *, kh=3, kw=3, stride=1, dil=1, pad=1, dgroups=4, dx=None, doffset=None):
    N, H, W, C = x.shape
    xr = _d(x).detach().clone().requires_grad_(True)
    orr = _d(offset).detach().clone().requires_grad_(True)
    col = _deform_cols(xr, orr, dil, pad, dgroups)
    gx, go = torch.autograd.grad(col, (xr, orr), _d(dcol).reshape(col.shape))
    wide = _exact(torch.float32)                      # the sums are fp32 tensors in the product whatever x's dtype
    dx = _st(gx, wide) if dx is None else dx + gx
    doffset = _st(go, wide) if doffset is None else doffset + go
    return dx, doffset


def affine_act(x, scale, shift, relu=True, out=None):
    y = _d(x) * _d(scale) + _d(shift)
    r = int(relu)
    if r:
        y = y.clamp(min=0)
    if r == 2:
        y = y.clamp(max=6)
    return _emit(y, out, x.dtype)


def bn_act_bwd(x, dy, bn, act, add=None, out=None, defer=False):
    C = bn.C
    xin = x
    x, dy = _d(x), _d(dy)
    scale, shift, mean, invstd = _d(bn.scale), _d(bn.shift), _d(bn.mean), _d(bn.invstd)
    y = x * scale + shift
    if act == 1:
        g = dy * (y > 0)
    elif act == 2:
        g = dy * ((y >= 0) & (y <= 6))
    else:
        g = dy
    xhat = (x - mean) * invstd
    gf, xf = g.reshape(-1, C), xhat.reshape(-1, C)
    M = gf.shape[0]
    S1, S2 = gf.sum(0), (gf * xf).sum(0)
    dx = scale * (g - S1 / M - xhat * (S2 / M))
    if add is not None:
        dx = dx + _d(add)
    if defer:
        bn.sums[:C] += S1
        bn.sums[C:] += S2
    else:
        bn.dbeta += S1.to(bn.dbeta.dtype)
        bn.dgamma += S2.to(bn.dgamma.dtype)
    return _emit(dx, out, xin.dtype)


def depthwise3x3(x, w, stride=1, out=None):
    C = x.shape[3]
    w4 = _d(w).t().reshape(C, 1, 3, 3)
    return _emit(_nhwc(F.conv2d(_nchw(_d(x)), w4, None, stride, 1, 1, C)), out, x.dtype)


def depthwise3x3_dgrad(dy, w, in_hw, stride=1, out=None):
    NB, Ho, Wo, C = dy.shape
    w4 = _d(w).t().reshape(C, 1, 3, 3)
    g = torch.nn.grad.conv2d_input((NB, C, in_hw[0], in_hw[1]), w4, _nchw(_d(dy)).contiguous(), stride, 1, 1, C)
    return _emit(_nhwc(g), out, dy.dtype)


def depthwise3x3_wgrad(x, dy, dw, stride=1):
    C = x.shape[3]
    g = torch.nn.grad.conv2d_weight(_nchw(_d(x)).contiguous(), (C, 1, 3, 3), _nchw(_d(dy)).contiguous(), stride, 1, 1, C)
    dw += g.reshape(C, 9).t().to(dw.dtype)
    return dw


def im2col3x3s2(x_nchw, Kp, dtype=None, out=None):
    NB, Cin, H, W = x_nchw.shape
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    u = F.unfold(x_nchw, 3, padding=1, stride=2).view(NB, Cin, 9, Ho, Wo).permute(0, 3, 4, 2, 1).reshape(NB, Ho, Wo, 9 * Cin)
    col = x_nchw.new_zeros(NB, Ho, Wo, Kp)
    col[..., :9 * Cin] = u
    return _st(col, dtype if dtype is not None else x_nchw.dtype) if dtype in (torch.bfloat16,) else col


def add_rows(a, b, out=None):
    return _emit(_d(a) + _d(b), out, a.dtype)


def _exact(dtype):
    """In the float64 tests "fp32" means "the exact type": a float32 request becomes float64 (a bf16 -> fp32 Cast is exact
    anyway), so that nothing but the bf16 storage rounds."""
    return torch.float64 if (dtype == torch.float32 and torch.get_default_dtype() == torch.float64) else dtype


def cast_rows(x, dtype=None, out=None):
    if out is not None:
        out.copy_(_st(_d(x), out.dtype))
        return out
    return _st(_d(x), _exact(dtype)).contiguous()


def relu_bwd(y, dy, out=None):
    return _emit(_d(dy) * (y > 0), out, dy.dtype)


def colsum_accum(x, out):
    out += x.reshape(-1, x.shape[-1]).sum(0)


def count_valid(label, out, ignore=-1):
    out += int((label.long() != ignore).sum())


def rpn_softmax_loss(score, label, A, grad_scale, valid_cnt, prob, dscore, loss_sum):
    B, H, W, _ = score.shape
    s0, s1 = score[..., :A], score[..., A:2 * A]
    p = torch.softmax(torch.stack([s0, s1], -1), -1)
    prob[..., :A] = p[..., 0]
    prob[..., A:2 * A] = p[..., 1]
    if dscore is None:
        return
    lab = label.view(B, A, H, W).permute(0, 2, 3, 1).long()
    norm = grad_scale / max(int(valid_cnt.item()), 1) if valid_cnt is not None else grad_scale
    valid = lab != -1
    dscore[..., :A] = torch.where(valid, (p[..., 0] - (lab == 0).to(p.dtype)) * norm, torch.zeros_like(s0))
    dscore[..., A:2 * A] = torch.where(valid, (p[..., 1] - (lab == 1).to(p.dtype)) * norm, torch.zeros_like(s0))
    pl = torch.where(lab == 1, p[..., 1], p[..., 0]).clamp(min=1e-14)
    loss_sum += -(torch.log(pl)[valid]).sum()


def _sl1(d):
    return torch.where(d.abs() < 1, 0.5 * d * d, d.abs() - 0.5), torch.where(d.abs() < 1, d, torch.sign(d))


def rpn_smooth_l1_loss(pred, target, weight, C4, grad_scale, dpred, loss_sum):
    d = pred[..., :C4] - _nhwc(target)
    v, g = _sl1(d)
    w = _nhwc(weight)
    dpred[..., :C4] = w * g * grad_scale
    loss_sum += (w * v).sum()


def softmax_ce(logits, label, K, grad_scale, valid_cnt, prob, grad, loss_sum, ignore=-1):
    p = torch.softmax(logits[:, :K], 1)
    if prob is not None:
        prob.copy_(p)
    if grad is None:
        return
    lab = label.long()
    valid = lab != ignore
    norm = grad_scale / max(int(valid_cnt.item()), 1) if valid_cnt is not None else grad_scale
    onehot = F.one_hot(lab.clamp(min=0), K).to(p.dtype)
    grad[:, :K] = torch.where(valid.unsqueeze(1), (p - onehot) * norm, torch.zeros_like(p))
    loss_sum += -(torch.log(p.gather(1, lab.clamp(min=0).unsqueeze(1)).squeeze(1).clamp(min=1e-14))[valid]).sum()


def smooth_l1_loss(pred, target, weight, C, grad_scale, grad, loss_sum):
    v, g = _sl1(pred - target)
    grad.copy_(weight * g * grad_scale)
    loss_sum += (weight * v).sum()


def multi_proposal_target(cls_prob, bbox_pred, im_info, gt_boxes, valid_ranges, *, feat_stride=16, scales=(), ratios=(),
                          rpn_post_nms_top_n=300, threshold=0.7, layout=NCHW, return_keep=False, return_fallback=False):
    A = len(scales) * len(ratios)
    assert layout == NHWC
    n = lambda t: np.ascontiguousarray(t.detach().cpu().numpy(), dtype=np.float32)
    res = O.multi_proposal_target(n(_nchw(cls_prob[..., :2 * A])), n(_nchw(bbox_pred[..., :4 * A])), n(im_info), n(gt_boxes),
                                  n(valid_ranges), feat_stride=feat_stride, scales=scales, ratios=ratios,
                                  post=rpn_post_nms_top_n)
    t = lambda a: torch.from_numpy(a).to(cls_prob.dtype)
    return t(res["rois"]), t(res["label"].reshape(-1)), t(res["bbox_target"]), t(res["bbox_weight"])


def multi_proposal(cls_prob, bbox_pred, im_info, *, feat_stride=16, scales=(), ratios=(), rpn_pre_nms_top_n=12000,
                   rpn_post_nms_top_n=300, threshold=0.7, suppress_anchor_types=False, fast_nms=False, roi_iou_thresh=0.3,
                   layout=NCHW, return_keep=False):
    A = len(scales) * len(ratios)
    assert layout == NHWC and not suppress_anchor_types and not fast_nms
    n = lambda t: np.ascontiguousarray(t.detach().cpu().numpy(), dtype=np.float32)
    res = O.multi_proposal(n(_nchw(cls_prob[..., :2 * A])), n(_nchw(bbox_pred[..., :4 * A])), n(im_info),
                           feat_stride=feat_stride, scales=scales, ratios=ratios, pre=rpn_pre_nms_top_n,
                           post=rpn_post_nms_top_n)
    return torch.from_numpy(res["rois"]).to(cls_prob.dtype), torch.from_numpy(res["scores"]).to(cls_prob.dtype)


_last_cnt = {}


def deform_psroi_fwd(data, rois, trans, *, spatial_scale, output_dim, group_size, pooled_size, part_size=0,
                     sample_per_part=1, trans_std=0.0, no_trans=False, layout=NCHW, want_count=True, want_sample_idx=False):
    assert layout == NHWC
    n = lambda t: np.ascontiguousarray(t.detach().cpu().numpy(), dtype=np.float32)
    out, cnt, _ = O.deform_psroi_fwd(n(_nchw(data)), n(rois), None if no_trans else n(trans), spatial_scale, output_dim,
                                     group_size, pooled_size, part_size, sample_per_part, trans_std, no_trans)
    _last_cnt[bool(no_trans)] = cnt
    return torch.from_numpy(out).to(data.dtype).permute(0, 2, 3, 1).contiguous(), None, None


def deform_psroi_bwd(top_diff, data, rois, trans, *, spatial_scale, output_dim, group_size, pooled_size, part_size=0,
                     sample_per_part=1, trans_std=0.0, no_trans=False, layout=NCHW, data_diff=None, trans_diff=None):
    assert layout == NHWC
    n = lambda t: np.ascontiguousarray(t.detach().cpu().numpy(), dtype=np.float32)
    dd, td = O.deform_psroi_bwd(n(top_diff.permute(0, 3, 1, 2)), _last_cnt[bool(no_trans)], n(_nchw(data)), n(rois),
                                None if no_trans else n(trans), spatial_scale, output_dim, group_size, pooled_size,
                                part_size, sample_per_part, trans_std, no_trans)
    g = torch.from_numpy(dd).to(data.dtype).permute(0, 2, 3, 1)
    if data_diff is None:
        data_diff = g.contiguous()
    else:
        data_diff += g
    if not no_trans:
        t = torch.from_numpy(td).to(data.dtype)
        trans_diff = t if trans_diff is None else trans_diff + t
    return data_diff, trans_diff


def sgd_mom_dev(w, mom, g, hyper, lr_mult, wd_mult, momentum, rescale=1.0, w_bf16=None):
    """optimizer_op-inl.h:279-300: mom = momentum*mom - lr*(rescale*g + wd*w); w += mom."""
    lr, wd = float(hyper[0]) * lr_mult, float(hyper[1]) * wd_mult
    mom.mul_(momentum).sub_(lr * (rescale * g + wd * w))
    w.add_(mom)
    if w_bf16 is not None:
        w_bf16.copy_(w)


PATCHED = [k for k, v in list(globals().items()) if callable(v) and not k.startswith("_") and k not in ("F", "O")]


def install(monkeypatch, ops_module):
    for k in PATCHED:
        if hasattr(ops_module, k):
            monkeypatch.setattr(ops_module, k, globals()[k])
