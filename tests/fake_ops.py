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


def _emit(y, out):
    if out is None:
        return y.contiguous()
    out.copy_(y)
    return out


def _stats(y, stats):
    if stats is not None:
        C = y.shape[-1]
        f = y.reshape(-1, C).double()
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
    y = _epi(a @ b.to(a.dtype).t(), scale, bias, residual, relu)      # (data-gradient operands are stored fp32)
    if accumulate:
        y = y + out
    _stats(y, stats)
    return _emit(y, out)


def conv2d_nhwc(x, w, *, kh, kw, stride=1, dil=1, pad=0, out=None, scale=None, bias=None, residual=None, relu=False,
                accumulate=False, taps=None, out_hw=None, out_map=None, stats=None, out_dtype=None):
    assert taps is None and out_hw is None and out_map is None, "strided data-gradient maps are not used by this model"
    Cout, Cin = w.shape[0], x.shape[3]
    w4 = w.to(x.dtype).view(Cout, kh, kw, Cin).permute(0, 3, 1, 2)
    y = _nhwc(F.conv2d(_nchw(x), w4, None, stride, pad, dil))
    y = _epi(y, scale, bias, residual, relu)
    _stats(y, stats)
    return _emit(y, out)


def conv2d_wgrad_nhwc(dy, x, *, kh, kw, stride=1, dil=1, pad=0, dw_out=None, splits=8, taps=None):
    Cout, Cin = dy.shape[3], x.shape[3]
    g = torch.nn.grad.conv2d_weight(_nchw(x).contiguous(), (Cout, Cin, kh, kw), _nchw(dy).contiguous(), stride, pad, dil)
    dw_out += g.permute(0, 2, 3, 1).reshape(Cout, kh * kw * Cin)
    return dw_out


def weight_transpose_jobs(jobs, device):
    return list(jobs)


def weight_transpose_batched(table):
    for w, wt, sel, Cout, T, Cin in table:
        s = sel.long()
        wt.view(Cin, len(s), Cout).copy_(w.view(Cout, T, Cin)[:, s, :].permute(2, 1, 0))


def bn_param_grad_jobs(states, device):
    return list(states)


def bn_param_grad_batched(table):
    for st in table:
        C = st.C
        st.dbeta += st.sums[:C].to(st.dbeta.dtype)
        st.dgamma += st.sums[C:].to(st.dgamma.dtype)
        st.sums.zero_()


def _finish_stats(bn, mean, var, M, eps, momentum, fix_gamma, update_moving):
    g = torch.ones_like(bn.gamma) if fix_gamma else bn.gamma
    invstd = 1.0 / torch.sqrt(var + eps)
    bn.mean.copy_(mean); bn.invstd.copy_(invstd)
    bn.scale.copy_(g * invstd); bn.shift.copy_(bn.beta - mean * g * invstd)
    if update_moving:
        unb = var * M / (M - 1) if M > 1 else var
        bn.moving_mean.copy_(bn.moving_mean * momentum + mean * (1 - momentum))
        bn.moving_var.copy_(bn.moving_var * momentum + unb * (1 - momentum))


def bn_stats(x, bn, eps=2e-5, momentum=0.9, fix_gamma=False, update_moving=True):
    f = x.reshape(-1, x.shape[-1])
    M = f.shape[0]
    _finish_stats(bn, f.mean(0), f.var(0, unbiased=False), M, eps, momentum, fix_gamma, update_moving)


def bn_finalize(bn, M, eps=2e-5, momentum=0.9, fix_gamma=False, update_moving=True):
    C = bn.C
    mean = (bn.sums[:C] / M).to(bn.mean.dtype)
    var = (bn.sums[C:] / M).to(bn.mean.dtype) - mean * mean
    _finish_stats(bn, mean, var.clamp(min=0), M, eps, momentum, fix_gamma, update_moving)
    bn.sums.zero_()


def bn_frozen(bn, eps=2e-5, fix_gamma=False):
    g = torch.ones_like(bn.gamma) if fix_gamma else bn.gamma
    bn.scale.copy_(g / torch.sqrt(bn.moving_var + eps))
    bn.shift.copy_(bn.beta - bn.moving_mean * bn.scale)


def affine_act(x, scale, shift, relu=True, out=None):
    y = x * scale + shift
    r = int(relu)
    if r:
        y = y.clamp(min=0)
    if r == 2:
        y = y.clamp(max=6)
    return _emit(y, out)


def bn_act_bwd(x, dy, bn, act, add=None, out=None, defer=False):
    C = bn.C
    y = x * bn.scale + bn.shift
    if act == 1:
        g = dy * (y > 0)
    elif act == 2:
        g = dy * ((y >= 0) & (y <= 6))
    else:
        g = dy
    xhat = (x - bn.mean) * bn.invstd
    gf, xf = g.reshape(-1, C), xhat.reshape(-1, C)
    M = gf.shape[0]
    S1, S2 = gf.sum(0), (gf * xf).sum(0)
    dx = bn.scale * (g - S1 / M - xhat * (S2 / M))
    if add is not None:
        dx = dx + add
    if defer:
        bn.sums[:C] += S1.double()
        bn.sums[C:] += S2.double()
    else:
        bn.dbeta += S1
        bn.dgamma += S2
    return _emit(dx, out)


def depthwise3x3(x, w, stride=1, out=None):
    C = x.shape[3]
    w4 = w.t().reshape(C, 1, 3, 3)
    return _emit(_nhwc(F.conv2d(_nchw(x), w4, None, stride, 1, 1, C)), out)


def depthwise3x3_dgrad(dy, w, in_hw, stride=1, out=None):
    NB, Ho, Wo, C = dy.shape
    w4 = w.t().reshape(C, 1, 3, 3)
    g = torch.nn.grad.conv2d_input((NB, C, in_hw[0], in_hw[1]), w4, _nchw(dy).contiguous(), stride, 1, 1, C)
    return _emit(_nhwc(g), out)


def depthwise3x3_wgrad(x, dy, dw, stride=1):
    C = x.shape[3]
    g = torch.nn.grad.conv2d_weight(_nchw(x).contiguous(), (C, 1, 3, 3), _nchw(dy).contiguous(), stride, 1, 1, C)
    dw += g.reshape(C, 9).t()
    return dw


def im2col3x3s2(x_nchw, Kp, dtype=None, out=None):
    NB, Cin, H, W = x_nchw.shape
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    u = F.unfold(x_nchw, 3, padding=1, stride=2).view(NB, Cin, 9, Ho, Wo).permute(0, 3, 4, 2, 1).reshape(NB, Ho, Wo, 9 * Cin)
    col = x_nchw.new_zeros(NB, Ho, Wo, Kp)
    col[..., :9 * Cin] = u
    return col


def add_rows(a, b, out=None):
    return _emit(a + b, out)


def cast_rows(x, dtype=None, out=None):
    if out is not None:
        out.copy_(x)
        return out
    return x.to(dtype).contiguous()


def relu_bwd(y, dy, out=None):
    return _emit(dy * (y > 0), out)


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


PATCHED = [k for k, v in list(globals().items()) if callable(v) and not k.startswith("_") and k not in ("F", "O")]


def install(monkeypatch, ops_module):
    for k in PATCHED:
        if hasattr(ops_module, k):
            monkeypatch.setattr(ops_module, k, globals()[k])
