"""ORACLE -- test infrastructure only (never imported by the product path).

`mobilenetv2_e2e.get_symbol_rcnn(cfg, is_train=True)` of the reference (symbols/faster/mobilenetv2_e2e.py:171-305, with
mobilenet_unit :27-43, inverted_residual_unit :46-90, invresi_blocks :93-117, MNETV2_CONFIGS_MAP :120-135, get_rpn
:160-169) restated as a PyTorch autograd graph over tensors carrying the REFERENCE's parameter names and layouts (OIHW
convolutions incl. the (C,1,3,3) depthwise filters, separate rpn_cls_score / rpn_bbox_pred and cls_score / bbox_pred).
float64; arithmetic modes as in oracle/torch_graph.py (MODE "exact" | "tf32" | "bf16": the tensor-core contractions --
first layer, 1x1 convolutions, heads -- read TF32-truncated / bf16 operands, the depthwise layers are exact fp32 FMA
kernels in the product and plain float64 here; in "bf16" every tensor the product stores between the two Casts is
rounded to bf16, forward and backward).

Gradient conventions of the reference operators: BatchNorm on batch statistics (fix_bn=False), gamma / beta are
FIXED_PARAMS (sniper_mobilenetv2_e2e.yml:22-32: the only patterns that match a MobileNetV2 name) and get no gradient;
clip(0, 6) passes the gradient where 0 <= y <= 6 (tensor/matrix_op-inl.h:1319-1332 -- torch.clamp does the same);
rpn_cls_prob: SoftmaxOutput(normalization='valid'); cls_prob: SoftmaxOutput WITHOUT normalization, grad_scale
1 / (300 * BATCH_IMAGES) (:281-282); bbox_loss: MakeLoss grad_scale 1 / (188 * BATCH_IMAGES) (:285-286); rpn_bbox_loss:
3 / (BATCH_IMAGES * RPN_BATCH_SIZE) (:299-301).
"""
import numpy as np
import torch
import torch.nn.functional as F

import torch_graph as TG

BOTTLENECKS = ((1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1))
FIRST_C, LAST_C = 32, 1280
PSROI_KW = dict(spatial_scale=0.03125, output_dim=256, group_size=1, pooled=7, part_size=7, spp=4, trans_std=0.1)


def is_fixed(name):
    return name.endswith("_gamma") or name.endswith("_beta")


def layer_list():
    """[(prefix, cin, cout, kernel, stride, groups)] of every mobilenet_unit in graph order."""
    L = [("first-3x3-conv", 3, FIRST_C, 3, 2, 1)]
    in_c = FIRST_C
    for i, (t, c, n, s) in enumerate(BOTTLENECKS):
        for j in range(n):
            ci = in_c if j == 0 else c
            e = int(round(ci * t))
            p = "seq-%d-block%d" % (i, j)
            L += [(p + "-exp", ci, e, 1, 1, 1), (p + "-depthwise", e, e, 3, s if j == 0 else 1, e), (p + "-linear", e, c, 1, 1, 1)]
        in_c = c
    L.append(("last-1x1-conv", in_c, LAST_C, 1, 1, 1))
    return L


def make_params(seed=0, num_anchors=15, num_classes=81, fc_dim=512, rpn_dim=256):
    """Random parameters under the reference's names / shapes (numpy float32): He-normal backbone, N(0, 0.01) heads."""
    rng = np.random.RandomState(seed)
    arg, aux = {}, {}
    for p, ci, co, k, s, g in layer_list():
        fan = (ci // g) * k * k
        arg[p + "-conv2d_weight"] = (rng.randn(co, ci // g, k, k) * np.sqrt(2.0 / fan)).astype(np.float32)
        arg[p + "-batchnorm_gamma"] = rng.uniform(0.8, 1.2, co).astype(np.float32)
        arg[p + "-batchnorm_beta"] = (rng.randn(co) * 0.1).astype(np.float32)
        aux[p + "-batchnorm_moving_mean"] = np.zeros(co, np.float32)
        aux[p + "-batchnorm_moving_var"] = np.ones(co, np.float32)
    A = num_anchors
    for n, shape in (("rpn_conv_3x3", (rpn_dim, LAST_C, 3, 3)), ("rpn_cls_score", (2 * A, rpn_dim, 1, 1)),
                     ("rpn_bbox_pred", (4 * A, rpn_dim, 1, 1)), ("conv_new_1", (256, LAST_C, 1, 1)),
                     ("offset", (98, 256 * 49)), ("fc_new_1", (fc_dim, 256 * 49)), ("fc_new_2", (fc_dim, fc_dim)),
                     ("cls_score", (num_classes, fc_dim)), ("bbox_pred", (4, fc_dim))):
        arg[n + "_weight"] = (rng.randn(*shape) * (0.001 if n == "offset" else 0.01)).astype(np.float32)
        arg[n + "_bias"] = np.zeros(shape[0], np.float32)
    return arg, aux


def params_to_torch(arg, aux, dtype=torch.float64, device="cpu"):
    P = {k: torch.from_numpy(np.ascontiguousarray(v)).to(device=device, dtype=dtype).requires_grad_(not is_fixed(k))
         for k, v in arg.items()}
    A = {k: torch.from_numpy(np.ascontiguousarray(v)).to(device=device, dtype=dtype) for k, v in aux.items()}
    return P, A


def _unit(P, x, prefix, k, stride, groups, act, eps, first=False, aux=None):
    """mobilenet_unit: Convolution(no_bias) -> BatchNorm -> clip(0, 6) | identity.  aux = None: batch statistics (training
    graph); aux = the auxiliary states: moving statistics (is_train=False)."""
    w = P[prefix + "-conv2d_weight"]
    if groups > 1:
        c = F.conv2d(x, w, None, stride, 1, 1, groups)               # depthwise: exact FMA kernel in the product
    elif first and TG.MODE[0] in ("bf16", "bf16x"):
        c = F.conv2d(TG._bf16(x), TG._RoundWeight.apply(w), None, stride, 1)       # bf16 im2col buffer x bf16 weight rows
    else:
        c = TG.conv2d(x, w, None, stride, (k - 1) // 2)
    c = TG.qs(c)
    if aux is not None:
        y = F.batch_norm(c, aux[prefix + "-batchnorm_moving_mean"], aux[prefix + "-batchnorm_moving_var"],
                         P[prefix + "-batchnorm_gamma"], P[prefix + "-batchnorm_beta"], False, 0.0, eps)
    else:
        y = F.batch_norm(c, None, None, P[prefix + "-batchnorm_gamma"], P[prefix + "-batchnorm_beta"], True, 0.0, eps)
    if act:
        y = torch.clamp(y, 0.0, 6.0)
    return TG.qs(y)


def backbone(P, data, eps=1e-5, taps=None, aux=None):
    TG.LOWP[0] = True                      # the product's bf16 region starts at the first layer's im2col buffer
    x = _unit(P, data, "first-3x3-conv", 3, 2, 1, True, eps, first=True, aux=aux)
    if taps is not None:
        taps["first"] = x
    in_c = FIRST_C
    for i, (t, c, n, s) in enumerate(BOTTLENECKS):
        for j in range(n):
            p = "seq-%d-block%d" % (i, j)
            ci = in_c if j == 0 else c
            e = int(round(ci * t))
            a1 = _unit(P, x, p + "-exp", 1, 1, 1, True, eps, aux=aux)
            a2 = _unit(P, a1, p + "-depthwise", 3, s if j == 0 else 1, e, True, eps, aux=aux)
            y = _unit(P, a2, p + "-linear", 1, 1, 1, False, eps, aux=aux)
            x = TG.qs(y + x) if j > 0 else y
            if taps is not None:
                taps[p] = x
        in_c = c
    x = _unit(P, x, "last-1x1-conv", 1, 1, 1, True, eps, aux=aux)
    TG.LOWP[0] = False                     # Cast(float32) (:226)
    return x


def forward_train(P, A, batch, proposals, batch_images, rpn_batch_size=256, num_anchors=15, num_classes=81,
                  grad_scale=1.0, eps=1e-5, taps=None):
    """The training graph; same contract as torch_graph.forward_train (proposals = callback returning the
    MultiProposalTarget outputs as numpy arrays)."""
    data = batch["data"]
    B = data.shape[0]
    An = num_anchors
    fm = backbone(P, data, eps, taps)
    rpn = F.relu(TG.conv2d(fm, P["rpn_conv_3x3_weight"], P["rpn_conv_3x3_bias"], 1, 1))
    rpn_cls_score = TG.conv2d(rpn, P["rpn_cls_score_weight"], P["rpn_cls_score_bias"])
    rpn_bbox_pred = TG.conv2d(rpn, P["rpn_bbox_pred_weight"], P["rpn_bbox_pred_bias"])
    feat = F.relu(TG.conv2d(fm, P["conv_new_1_weight"], P["conv_new_1_bias"]))
    H, W = rpn_cls_score.shape[2], rpn_cls_score.shape[3]
    score2 = rpn_cls_score.reshape(B, 2, An * H, W)
    rpn_label = batch["label"].reshape(B, An * H, W).long()
    logp = F.log_softmax(score2, 1)
    rpn_prob = logp.exp()
    valid = rpn_label != -1
    rpn_cls_sum = -(logp.gather(1, rpn_label.clamp(min=0).unsqueeze(1)).squeeze(1))[valid].sum()
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

    offset_t = TG.DeformPSROI.apply(feat, None, rois, PSROI_KW)
    offset = TG.linear(offset_t.reshape(N, -1), P["offset_weight"], P["offset_bias"])
    trans = offset.reshape(N, 2, 7, 7)
    pooled = TG.DeformPSROI.apply(feat, trans, rois, PSROI_KW)
    fc1 = F.relu(TG.linear(pooled.reshape(N, -1), P["fc_new_1_weight"], P["fc_new_1_bias"]))
    fc2 = F.relu(TG.linear(fc1, P["fc_new_2_weight"], P["fc_new_2_bias"]))
    cls_score = TG.linear(fc2, P["cls_score_weight"], P["cls_score_bias"])
    bbox_pred = TG.linear(fc2, P["bbox_pred_weight"], P["bbox_pred_bias"])
    lp = F.log_softmax(cls_score, 1)
    cvalid = label != -1
    cls_sum = -(lp.gather(1, label.clamp(min=0).unsqueeze(1)).squeeze(1))[cvalid].sum()
    cls_obj = (grad_scale / (300.0 * batch_images)) * cls_sum
    d2 = bbox_pred - bbox_target
    sl2 = torch.where(d2.abs() < 1, 0.5 * d2 * d2, d2.abs() - 0.5)
    bbox_sum = (bbox_weight * sl2).sum()
    bbox_obj = (grad_scale / (188.0 * batch_images)) * bbox_sum
    objective = rpn_cls_obj + rpn_bbox_obj + cls_obj + bbox_obj
    out = dict(rpn_cls_prob=rpn_prob.reshape(B, 2 * An, H, W), rpn_bbox_pred=rpn_bbox_pred, feat=feat, last_fm=fm,
               cls_prob=lp.exp(), bbox_pred=bbox_pred, trans=trans, pooled=pooled,
               loss_sums=torch.stack([rpn_cls_sum, rpn_bbox_sum, cls_sum, bbox_sum]).detach(), rois=rois, label=label)
    return objective, out


def forward_test(P, A, data, proposals, num_anchors=15, eps=1e-5):
    """get_symbol_rcnn(cfg, is_train=False) (:306-362): moving-statistics BatchNorm, SoftmaxActivation(mode=channel) over
    the reshaped RPN scores, MultiProposal (callback: (rpn_cls_prob [B,2A,H,W], rpn_bbox_pred) -> rois [N,5] numpy), the
    deformable R-FCN head, softmax.  Returns dict(rois, cls_prob [N,K], bbox_pred [N,4], rpn_cls_prob, rpn_bbox_pred)."""
    B = data.shape[0]
    An = num_anchors
    with torch.no_grad():
        fm = backbone(P, data, eps, aux=A)
        rpn = F.relu(TG.conv2d(fm, P["rpn_conv_3x3_weight"], P["rpn_conv_3x3_bias"], 1, 1))
        score = TG.conv2d(rpn, P["rpn_cls_score_weight"], P["rpn_cls_score_bias"])
        rpn_bbox_pred = TG.conv2d(rpn, P["rpn_bbox_pred_weight"], P["rpn_bbox_pred_bias"])
        feat = F.relu(TG.conv2d(fm, P["conv_new_1_weight"], P["conv_new_1_bias"]))
        H, W = score.shape[2], score.shape[3]
        prob = F.softmax(score.reshape(B, 2, An * H, W), 1).reshape(B, 2 * An, H, W)
        rois = np.ascontiguousarray(proposals(prob, rpn_bbox_pred), dtype=np.float32)
        N = rois.shape[0]
        offset_t = TG.DeformPSROI.apply(feat, None, rois, PSROI_KW)
        trans = TG.linear(offset_t.reshape(N, -1), P["offset_weight"], P["offset_bias"]).reshape(N, 2, 7, 7)
        pooled = TG.DeformPSROI.apply(feat, trans, rois, PSROI_KW)
        fc1 = F.relu(TG.linear(pooled.reshape(N, -1), P["fc_new_1_weight"], P["fc_new_1_bias"]))
        fc2 = F.relu(TG.linear(fc1, P["fc_new_2_weight"], P["fc_new_2_bias"]))
        cls_prob = F.softmax(TG.linear(fc2, P["cls_score_weight"], P["cls_score_bias"]), 1)
        bbox_pred = TG.linear(fc2, P["bbox_pred_weight"], P["bbox_pred_bias"])
    return dict(rois=rois, cls_prob=cls_prob, bbox_pred=bbox_pred, rpn_cls_prob=prob, rpn_bbox_pred=rpn_bbox_pred, last_fm=fm)
