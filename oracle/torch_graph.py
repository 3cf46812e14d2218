"""ORACLE -- test & measurement infrastructure only (never imported by the product path).

`get_symbol_rcnn(cfg, is_train=True)` of the reference (symbols/faster/resnet_mx_101_e2e.py:227-345, with
resnetc4 :394-420, resnetc5 :422-448, residual_unit :36-69, residual_unit_deform :106-145, get_rpn :147-155) restated
as a PyTorch autograd graph over tensors carrying the REFERENCE's parameter names and layouts (OIHW convolutions,
NCHW-flattened FullyConnected inputs, separate rpn_cls_score / rpn_bbox_pred and cls_score / bbox_pred heads).  Runs
in float64 (the parity oracle of tests/test_graph_parity_gpu.py) or float32 (the CPU baseline of bench.py).

Operators with no PyTorch equivalent:
  * DeformableConvolution: `deform_conv2d` below, a differentiable gather formulation of deformable_im2col.cuh:78-113,
    216-263 (+ the GEMM of deformable_convolution-inl.h:111-168);
  * DeformablePSROIPooling: `DeformPSROI`, an autograd.Function around the C oracle (oracle/psroi.c, restating
    deformable_psroi_pooling.cu:71-161 forward and :203-330 backward);
  * MultiProposalTarget: a callback (the tests pass the C oracle, oracle/mpt.c); its backward is a zero fill
    (multi_proposal_target.cu:591-615), so rois / labels / targets enter the graph as constants.
Gradient conventions follow the reference operators, not PyTorch's defaults: SoftmaxOutput(normalization='valid',
use_ignore) back-propagates (p - onehot) * grad_scale / #valid (softmax_output-inl.h:162-263) = the gradient of the
mean cross entropy over valid labels; MakeLoss back-propagates grad_scale * d(sum) (make_loss-inl.h).
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

UNITS = (3, 4, 23, 3)
FILTERS = (64, 256, 512, 1024, 2048)
FIXED = ("conv0", "bn0", "stage1", "bn_data")        # FIXED_PARAMS, sniper_res101_e2e.yml:22-25 (+ use_global_stats bn_data)


def is_fixed(name):
    return any(f in name for f in FIXED)


# ------------------------------------------------------------------------------------------------------------------
# Contractions.  MODE "exact": plain float64 / float32 arithmetic.  MODE "tf32": every operand of every tensor-core
# contraction (forward, data gradient, weight gradient) is first reduced to TF32 the way the tcgen05 `kind::tf32` MMA
# reads fp32 words from shared memory -- the low 13 mantissa bits are ignored (truncation, not rounding;
# tests/test_graph_parity_gpu.py measures that on the device) -- and the products are accumulated exactly (float64).
# This models the product path's only systematic deviation from real arithmetic, so the whole graph can be compared
# with a tolerance ~100x tighter than against MODE "exact" (random-init ResNet-101 amplifies each unit's TF32 error by
# ~4 % per residual unit: 1e-3 after stage 1 grows to ~1e-1 at c4).  conv0: STEM "tc" (the product's default: im2col +
# tcgen05 GEMM, operands reduced like every other contraction; in MODE "bf16" the im2col buffer and the weight rows are
# bf16) or "fma" (SNIPER_STEM_TC=0: exact fp32 FMA kernel).
# MODE "bf16": the mixed-precision path (Cfg.bf16).  Inside the backbone (LOWP) every tensor the product STORES --
# activations, the im2col buffer, the bf16 weight copies, and on the way back the gradients of those activations -- is
# rounded to bf16 (round-to-nearest-even), contractions accumulate exactly; the heads behave as in MODE "tf32".
MODE = ["exact"]
LOWP = [False]
STEM = ["tc"]


def _bf16(x):
    return x.detach().to(torch.float32).to(torch.bfloat16).to(x.dtype)


class _RoundStored(torch.autograd.Function):
    """a stored activation: rounded on the way forward, its gradient rounded on the way back"""
    @staticmethod
    def forward(ctx, x):
        return _bf16(x)

    @staticmethod
    def backward(ctx, g):
        return _bf16(g)


class _RoundWeight(torch.autograd.Function):
    """the bf16 copy of an fp32 master weight: rounded forward, gradient passed to the master unchanged"""
    @staticmethod
    def forward(ctx, w):
        return _bf16(w)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundGrad(torch.autograd.Function):
    """identity forward, bf16-rounded gradient: a point where the product stores a GRADIENT in bf16 although the forward
    tensor there is wide (the deformable units: the offset gradient is cast to bf16 for the offset convolution's backward
    GEMMs, and the im2col path's share of the activation gradient is cast before the offset path's share is added)"""
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return _bf16(g)


def qg(x):
    return _RoundGrad.apply(x) if (MODE[0] in ("bf16", "bf16x") and LOWP[0]) else x


def qs(x):
    return _RoundStored.apply(x) if (MODE[0] in ("bf16", "bf16x") and LOWP[0]) else x


def qw(w):
    return _RoundWeight.apply(w) if (MODE[0] in ("bf16", "bf16x") and LOWP[0]) else w


def tf32(x):
    """x -> fp32 -> TF32 (low 13 mantissa bits cleared) -> x.dtype."""
    i = x.detach().to(torch.float32).contiguous().view(torch.int32)
    return (i & -8192).view(torch.float32).to(x.dtype)


class _TF32Conv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, stride, padding, dilation):
        xt, wt = tf32(x), tf32(w)
        ctx.save_for_backward(xt, wt)
        ctx.cfg = (stride, padding, dilation)
        return F.conv2d(xt, wt, None, stride, padding, dilation)

    @staticmethod
    def backward(ctx, gy):
        xt, wt = ctx.saved_tensors
        stride, padding, dilation = ctx.cfg
        gt = tf32(gy)
        gx = torch.nn.grad.conv2d_input(xt.shape, wt, gt, stride, padding, dilation)
        gw = torch.nn.grad.conv2d_weight(xt, wt.shape, gt, stride, padding, dilation)
        return gx, gw, None, None, None


class _TF32MatmulNT(torch.autograd.Function):
    """y = a @ b^T with TF32 operands in all three contractions."""
    @staticmethod
    def forward(ctx, a, b):
        at, bt = tf32(a), tf32(b)
        ctx.save_for_backward(at, bt)
        return at @ bt.t()

    @staticmethod
    def backward(ctx, gy):
        at, bt = ctx.saved_tensors
        gt = tf32(gy)
        return gt @ bt, gt.t() @ at


def conv2d(x, w, b=None, stride=1, padding=0, dilation=1, exact=False):
    if MODE[0] == "exact" or exact:
        return F.conv2d(x, w, b, stride, padding, dilation)
    if MODE[0] in ("bf16", "bf16x") and LOWP[0]:          # operands are stored bf16 tensors already; exact accumulation
        return F.conv2d(x, qw(w), b, stride, padding, dilation)
    if MODE[0] == "bf16x":                     # "bf16x": bf16 storage between the Casts, EXACT arithmetic in the heads (the
        return F.conv2d(x, w, b, stride, padding, dilation)      # CPU orchestration tests; "bf16" models the TF32 heads too)
    y = _TF32Conv.apply(x, w, stride, padding, dilation)
    return y if b is None else y + b.view(1, -1, 1, 1)


def linear(x, w, b=None):
    if MODE[0] == "exact":
        return F.linear(x, w, b)
    if MODE[0] in ("bf16", "bf16x") and LOWP[0]:
        return F.linear(x, qw(w), b)
    if MODE[0] == "bf16x":
        return F.linear(x, w, b)
    y = _TF32MatmulNT.apply(x, w)
    return y if b is None else y + b


def deform_conv2d(x, offset, w, dil=2, pad=2, dg=4):
    """DeformableConvolution forward, kernel 3x3, stride 1 (resnet_mx_101_e2e.py:124-130).  x [N,C,H,W], offset
    [N, dg*2*9, H, W] with channel g*18 + 2*tap (+1) = (dy, dx) of tap (i,j)=divmod(tap,3) in deformable group g
    (deformable_im2col.cuh:237-246), w [Cout,C,3,3].  Sampling rule of deformable_im2col.cuh:78-113 + :247-258: a
    sample contributes iff 0 <= h < H and 0 <= w < W; `h_low >= H-1` clamps both corners to the last row (same for
    columns)."""
    N, C, H, W = x.shape
    cpg = C // dg
    dev = x.device
    hh, ww = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
    n_idx = torch.arange(N, device=dev).view(N, 1, 1).expand(N, H, W)
    xh = x.permute(0, 2, 3, 1)                                        # [N,H,W,C]
    cols = []
    for tap in range(9):
        i, j = divmod(tap, 3)
        per_g = []
        for g in range(dg):
            oh = offset[:, g * 18 + 2 * tap]
            ow = offset[:, g * 18 + 2 * tap + 1]
            h_im = (hh - pad + i * dil).to(x.dtype) + oh
            w_im = (ww - pad + j * dil).to(x.dtype) + ow
            valid = (h_im >= 0) & (w_im >= 0) & (h_im < H) & (w_im < W)
            h_low = torch.floor(h_im).clamp(max=H - 1)
            w_low = torch.floor(w_im).clamp(max=W - 1)
            hc = torch.where(torch.floor(h_im) >= H - 1, h_low, h_im)
            wc = torch.where(torch.floor(w_im) >= W - 1, w_low, w_im)
            h_high = (h_low + 1).clamp(max=H - 1)
            w_high = (w_low + 1).clamp(max=W - 1)
            lh, lw = hc - h_low, wc - w_low
            xg = xh[..., g * cpg:(g + 1) * cpg]

            def at(hi, wi, xg=xg):
                return xg[n_idx, hi.long().clamp(0, H - 1), wi.long().clamp(0, W - 1)]
            v = ((1 - lh) * (1 - lw)).unsqueeze(-1) * at(h_low, w_low) + ((1 - lh) * lw).unsqueeze(-1) * at(h_low, w_high) \
                + (lh * (1 - lw)).unsqueeze(-1) * at(h_high, w_low) + (lh * lw).unsqueeze(-1) * at(h_high, w_high)
            per_g.append(v * valid.unsqueeze(-1).to(x.dtype))
        cols.append(torch.cat(per_g, -1))
    col = qs(torch.stack(cols, 3))                                    # [N,H,W,9,C]; a stored (bf16) buffer in MODE "bf16"
    wt = w.permute(0, 2, 3, 1).reshape(w.shape[0], 9 * C)             # [Cout, (tap, C)]
    y = linear(col.reshape(N * H * W, 9 * C), wt)                     # the GEMM of deformable_convolution-inl.h:148-160
    return y.view(N, H, W, w.shape[0]).permute(0, 3, 1, 2)


class DeformPSROI(torch.autograd.Function):
    """DeformablePSROIPooling through the C oracle (float32 inside; rois are constants)."""

    @staticmethod
    def forward(ctx, data, trans, rois_np, kw):
        import oracle_lib as O
        d = data.detach().cpu().numpy().astype(np.float32)
        t = None if trans is None else trans.detach().cpu().numpy().astype(np.float32)
        out, cnt, _ = O.deform_psroi_fwd(d, rois_np, t, no_trans=t is None, **kw)
        ctx.saved = (d, t, rois_np, cnt, kw, data.dtype, data.device)
        return torch.from_numpy(out).to(device=data.device, dtype=data.dtype)

    @staticmethod
    def backward(ctx, g):
        import oracle_lib as O
        d, t, rois_np, cnt, kw, dt, dev = ctx.saved
        dd, td = O.deform_psroi_bwd(g.detach().cpu().numpy().astype(np.float32), cnt, d, rois_np, t, no_trans=t is None,
                                    **kw)
        gd = torch.from_numpy(dd).to(device=dev, dtype=dt)
        gt = None if t is None else torch.from_numpy(td).to(device=dev, dtype=dt)
        return gd, gt, None, None


PSROI_KW = dict(spatial_scale=0.0625, output_dim=256, group_size=1, pooled=7, part_size=7, spp=4, trans_std=0.1)


# ------------------------------------------------------------------------------------------------------------------
def params_to_torch(arg, aux, dtype=torch.float64, device="cpu"):
    """numpy (arg_params, aux_params) of a reference checkpoint -> torch; trainable tensors get requires_grad."""
    P = {}
    for k, v in arg.items():
        t = torch.from_numpy(np.ascontiguousarray(v)).to(device=device, dtype=dtype)
        P[k] = t.requires_grad_(not is_fixed(k))
    A = {k: torch.from_numpy(np.ascontiguousarray(v)).to(device=device, dtype=dtype) for k, v in aux.items()}
    return P, A


def _bn(P, A, x, name, eps, train, relu=True, fix_gamma=False):
    g, b = P[name + "_gamma"], P[name + "_beta"]
    if train:
        y = F.batch_norm(x, None, None, g, b, True, 0.0, eps)
    else:
        if fix_gamma:
            g = torch.ones_like(g)
        y = F.batch_norm(x, A[name + "_moving_mean"], A[name + "_moving_var"], g.detach(), b.detach(), False, 0.0, eps)
    return F.relu(y) if relu else y


def _unit(P, A, x, name, stride, dim_match, train, deform, eps, taps=None):
    """residual_unit :36-69 / residual_unit_deform :106-145.  The shortcut convolution reads act1, not data."""
    # Stored tensors (MODE "bf16"): a trainable unit stores every conv output (its BatchNorm needs batch statistics of
    # it); a frozen unit folds BN + ReLU into the producing conv's epilogue, so c1 / c2 never exist in memory there.
    qc = qs if train else (lambda t: t)
    a1 = qs(_bn(P, A, x, name + "_bn1", eps, train))
    # (the product stores conv1's data gradient in bf16 BEFORE the shortcut convolution's is added to it: qg)
    c1 = qc(conv2d(qg(a1) if train else a1, P[name + "_conv1_weight"]))
    a2 = qs(_bn(P, A, c1, name + "_bn2", eps, train))
    if deform:
        # offsets are kept fp32 by the product; their gradient goes through the offset convolution's backward GEMMs in
        # bf16 (the bias gradient is taken from the fp32 tensor), and the im2col path's activation gradient is stored in
        # bf16 before the offset path's is added
        off = qg(conv2d(a2, P[name + "_offset_weight"], None, 1, 2, 2)) + P[name + "_offset_bias"].view(1, -1, 1, 1)
        c2 = qc(deform_conv2d(qg(a2), off, P[name + "_conv2_weight"]))
    else:
        c2 = qc(conv2d(a2, P[name + "_conv2_weight"], None, stride, 1))
    a3 = qs(_bn(P, A, c2, name + "_bn3", eps, train))
    c3 = conv2d(a3, P[name + "_conv3_weight"])
    sc = x if dim_match else qs(conv2d(a1, P[name + "_sc_weight"], None, stride))
    out = qs(c3 + sc)                    # the residual is added in the fp32 epilogue, the sum stored once
    if taps is not None:
        taps[name] = dict(a1=a1, c1=c1, c2=c2, out=out)
    return out


def backbone(P, A, data, eps=2e-5, taps=None, is_train=True):
    """resnetc4 + resnetc5(deform=True) + Concat (:243-249).  Returns relu1 = cat(conv_feat, relut) [B,3072,H/16,W/16].
    is_train=False: every BatchNorm on its moving statistics (the test graph)."""
    x = _bn(P, A, data, "bn_data", eps, False, relu=False, fix_gamma=True)
    if STEM[0] == "fma" or MODE[0] == "exact":
        x = conv2d(x, P["conv0_weight"], None, 2, 3, exact=True)
    elif MODE[0] in ("bf16", "bf16x"):    # bf16 im2col buffer x bf16 weight rows, exact accumulation
        x = F.conv2d(_bf16(x), _bf16(P["conv0_weight"]), None, 2, 3)
    else:
        x = conv2d(x, P["conv0_weight"], None, 2, 3)
    LOWP[0] = True                        # the reference's Cast(float16) sits here (:405-406)
    x = qs(_bn(P, A, x, "bn0", eps, False))
    if taps is not None:
        taps["relu0"] = x
    x = F.max_pool2d(x, 3, 2, 1)
    if taps is not None:
        taps["pool0"] = x
    c4 = None
    for si, n in enumerate(UNITS):
        stage = si + 1
        for j in range(n):
            name = "stage%d_unit%d" % (stage, j + 1)
            stride = 2 if (j == 0 and stage in (2, 3)) else 1
            x = _unit(P, A, x, name, stride, j > 0, train=is_train and stage > 1, deform=stage == 4, eps=eps, taps=taps)
        if stage == 3:
            c4 = x
    c4 = qg(c4)                           # the concat's share of c4's gradient is cast to bf16 before stage 4's is added
    LOWP[0] = False                       # Cast(relu1, float32) (:250-252): the heads are fp32
    return torch.cat([c4, x], 1)


def forward_train(P, A, batch, proposals, batch_images, rpn_batch_size=256, num_anchors=21, num_classes=81,
                  grad_scale=1.0, eps=2e-5, taps=None):
    """The training graph.  batch: tensors named as MNIteratorE2E provides them (data NCHW, label [B,A*H*W],
    bbox_target / bbox_weight [B,4A,H,W], gt_boxes, valid_ranges, im_info).  proposals(rpn_cls_prob [B,2A,H,W],
    rpn_bbox_pred [B,4A,H,W]) -> dict(rois [N,5], label [N], bbox_target [N,4], bbox_weight [N,4]) as numpy arrays.
    Returns (objective, out) where d(objective)/d(param) is what the reference's backward leaves in the gradient
    arrays, and out holds the graph outputs + the un-normalised loss sums the product path reports."""
    data = batch["data"]
    B = data.shape[0]
    An = num_anchors
    relu1 = backbone(P, A, data, eps, taps)
    rpn = F.relu(conv2d(relu1, P["rpn_conv_3x3_weight"], P["rpn_conv_3x3_bias"], 1, 1))
    rpn_cls_score = conv2d(rpn, P["rpn_cls_score_weight"], P["rpn_cls_score_bias"])
    rpn_bbox_pred = conv2d(rpn, P["rpn_bbox_pred_weight"], P["rpn_bbox_pred_bias"])
    feat = F.relu(conv2d(relu1, P["conv_new_1_weight"], P["conv_new_1_bias"]))
    H, W = rpn_cls_score.shape[2], rpn_cls_score.shape[3]
    score2 = rpn_cls_score.reshape(B, 2, An * H, W)                       # rpn_cls_score_reshape (0, 2, -1, 0)
    rpn_label = batch["label"].reshape(B, An * H, W).long()
    logp = F.log_softmax(score2, 1)
    rpn_prob = logp.exp()
    valid = rpn_label != -1
    nll = -(logp.gather(1, rpn_label.clamp(min=0).unsqueeze(1)).squeeze(1))[valid]
    rpn_cls_sum = nll.sum()
    rpn_cls_obj = grad_scale * rpn_cls_sum / max(int(valid.sum()), 1)
    d = rpn_bbox_pred - batch["bbox_target"]
    sl1 = torch.where(d.abs() < 1, 0.5 * d * d, d.abs() - 0.5)
    rpn_bbox_sum = (batch["bbox_weight"] * sl1).sum()
    rpn_bbox_obj = (3.0 * grad_scale / float(batch_images * rpn_batch_size)) * rpn_bbox_sum

    prop = proposals(rpn_prob.reshape(B, 2 * An, H, W).detach(), rpn_bbox_pred.detach())
    rois = np.ascontiguousarray(prop["rois"], dtype=np.float32)
    N = rois.shape[0]
    dev, dt = data.device, data.dtype
    label = torch.from_numpy(np.asarray(prop["label"]).reshape(-1)).to(dev).long()
    bbox_target = torch.from_numpy(np.asarray(prop["bbox_target"])).to(device=dev, dtype=dt)
    bbox_weight = torch.from_numpy(np.asarray(prop["bbox_weight"])).to(device=dev, dtype=dt)

    offset_t = DeformPSROI.apply(feat, None, rois, PSROI_KW)             # [N,256,7,7]
    offset = linear(offset_t.reshape(N, -1), P["offset_weight"], P["offset_bias"])
    trans = offset.reshape(N, 2, 7, 7)
    pooled = DeformPSROI.apply(feat, trans, rois, PSROI_KW)
    fc1 = F.relu(linear(pooled.reshape(N, -1), P["fc_new_1_weight"], P["fc_new_1_bias"]))
    fc2 = F.relu(linear(fc1, P["fc_new_2_weight"], P["fc_new_2_bias"]))
    cls_score = linear(fc2, P["cls_score_weight"], P["cls_score_bias"])
    bbox_pred = linear(fc2, P["bbox_pred_weight"], P["bbox_pred_bias"])
    lp = F.log_softmax(cls_score, 1)
    cvalid = label != -1
    cls_sum = -(lp.gather(1, label.clamp(min=0).unsqueeze(1)).squeeze(1))[cvalid].sum()
    cls_obj = grad_scale * cls_sum / max(int(cvalid.sum()), 1)
    d2 = bbox_pred - bbox_target
    sl2 = torch.where(d2.abs() < 1, 0.5 * d2 * d2, d2.abs() - 0.5)
    bbox_sum = (bbox_weight * sl2).sum()
    bbox_obj = (grad_scale / (188.0 * 16.0)) * bbox_sum
    objective = rpn_cls_obj + rpn_bbox_obj + cls_obj + bbox_obj
    out = dict(rpn_cls_prob=rpn_prob.reshape(B, 2 * An, H, W), rpn_bbox_pred=rpn_bbox_pred, feat=feat, relu1=relu1,
               cls_prob=lp.exp(), bbox_pred=bbox_pred, trans=trans, pooled=pooled, offset_t=offset_t,
               loss_sums=torch.stack([rpn_cls_sum, rpn_bbox_sum, cls_sum, bbox_sum]).detach(), rois=rois, label=label)
    return objective, out


def forward_test(P, A, data, proposals, num_anchors=21, eps=2e-5, autofocus=False):
    """get_symbol_rcnn(cfg, is_train=False) (resnet_mx_101_e2e.py:227-267, 346-389): moving-statistics BatchNorm everywhere,
    SoftmaxActivation(mode=channel) over the reshaped RPN scores, MultiProposal (callback: (rpn_cls_prob [B,2A,H,W],
    rpn_bbox_pred) -> rois [N,5] numpy), the deformable R-FCN head, softmax; autofocus: the FocusPixel branch conv_new_2 ->
    relu -> conv_new_3 -> relu -> conv_new_out -> softmax over its two channels (:259-267, 385-386), channel 1 returned."""
    B = data.shape[0]
    An = num_anchors
    with torch.no_grad():
        relu1 = backbone(P, A, data, eps, is_train=False)
        rpn = F.relu(conv2d(relu1, P["rpn_conv_3x3_weight"], P["rpn_conv_3x3_bias"], 1, 1))
        score = conv2d(rpn, P["rpn_cls_score_weight"], P["rpn_cls_score_bias"])
        rpn_bbox_pred = conv2d(rpn, P["rpn_bbox_pred_weight"], P["rpn_bbox_pred_bias"])
        feat = F.relu(conv2d(relu1, P["conv_new_1_weight"], P["conv_new_1_bias"]))
        H, W = score.shape[2], score.shape[3]
        prob = F.softmax(score.reshape(B, 2, An * H, W), 1).reshape(B, 2 * An, H, W)
        rois = np.ascontiguousarray(proposals(prob, rpn_bbox_pred), dtype=np.float32)
        N = rois.shape[0]
        offset_t = DeformPSROI.apply(feat, None, rois, PSROI_KW)
        trans = linear(offset_t.reshape(N, -1), P["offset_weight"], P["offset_bias"]).reshape(N, 2, 7, 7)
        pooled = DeformPSROI.apply(feat, trans, rois, PSROI_KW)
        fc1 = F.relu(linear(pooled.reshape(N, -1), P["fc_new_1_weight"], P["fc_new_1_bias"]))
        fc2 = F.relu(linear(fc1, P["fc_new_2_weight"], P["fc_new_2_bias"]))
        out = dict(rois=rois, cls_prob=F.softmax(linear(fc2, P["cls_score_weight"], P["cls_score_bias"]), 1),
                   bbox_pred=linear(fc2, P["bbox_pred_weight"], P["bbox_pred_bias"]), rpn_cls_prob=prob, relu1=relu1)
        if autofocus:
            f = F.relu(conv2d(relu1, P["conv_new_2_weight"], P["conv_new_2_bias"], 1, 1))
            f = F.relu(conv2d(f, P["conv_new_3_weight"], P["conv_new_3_bias"]))
            f = conv2d(f, P["conv_new_out_weight"], P["conv_new_out_bias"])
            out["focus"] = F.softmax(f, 1)[:, 1]
    return out
