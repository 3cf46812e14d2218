"""Operator-registration surface of the reference, re-hosted on torch tensors over the C-ABI.

Mirrors python/mxnet/operator.py of the fork (CustomOp :426-470, CustomOpProp :472-672, register :692) as
used by lib/operator_py/box_annotator_ohem.py:19-120, so custom Python operators written against the
reference keep their shape: subclass `CustomOp` (forward / backward / assign), subclass `CustomOpProp`
(list_arguments, list_outputs, infer_shape, create_operator, declare_backward_dependency), decorate with
`@register('Name')`, invoke with `Custom(op_type='Name', **kwargs)` (kwargs arrive as strings, as in MXNet).

The SNIPER native operators are registered under the reference's names with the reference's argument
names and output order (multi_proposal_target-inl.h:55-177, deformable_psroi_pooling-inl.h:51-205,
psroi_pooling-inl.h, deformable_convolution-inl.h:59-96) and run the sm_100a kernels.
"""
import torch

from . import ops

_REGISTRY = {}


class CustomOp(object):
    """Base class of operators implemented in Python (python/mxnet/operator.py:426-470)."""

    def forward(self, is_train, req, in_data, out_data, aux):
        raise NotImplementedError()

    def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
        raise NotImplementedError()

    def assign(self, dst, req, src):
        """dst[:] = src honouring req in {'null','write','inplace','add'} (operator.py:458-470)."""
        if req == 'null':
            return
        if not torch.is_tensor(src):
            src = torch.as_tensor(src, dtype=dst.dtype, device=dst.device)
        if req in ('write', 'inplace'):
            dst.copy_(src.expand_as(dst) if src.dim() == 0 else src)
        elif req == 'add':
            dst.add_(src)
        else:
            raise ValueError("unknown req %r" % (req,))


class CustomOpProp(object):
    """Operator property (python/mxnet/operator.py:472-672)."""

    def __init__(self, need_top_grad=True):
        self.need_top_grad_ = need_top_grad

    def infer_shape(self, in_shape):
        return in_shape, (in_shape[0],) * len(self.list_outputs()), ()

    def infer_type(self, in_type):
        return in_type, [in_type[0]] * len(self.list_outputs()), [in_type[0]] * len(self.list_auxiliary_states())

    def list_outputs(self):
        return ['output']

    def list_arguments(self):
        return ['data']

    def list_auxiliary_states(self):
        return []

    def declare_backward_dependency(self, out_grad, in_data, out_data):
        deps = []
        if self.need_top_grad_:
            deps.extend(out_grad)
        deps.extend(in_data)
        deps.extend(out_data)
        return deps

    def create_operator(self, ctx, in_shapes, in_dtypes):
        return CustomOp()


def register(reg_name):
    """@register('Name') (python/mxnet/operator.py:692)."""
    def do_register(prop_cls):
        _REGISTRY[reg_name] = prop_cls
        return prop_cls
    return do_register


def get_all_registered_operators():
    return sorted(_REGISTRY)


def Custom(*args, **kwargs):
    """mx.sym.Custom / mx.nd.Custom: runs a registered operator eagerly.  Keyword tensors are matched to
    list_arguments(); every other kwarg is passed to the Prop constructor as a string."""
    op_type = kwargs.pop('op_type')
    is_train = kwargs.pop('is_train', True)
    if op_type not in _REGISTRY:
        raise KeyError("operator %r is not registered (known: %s)" % (op_type, ", ".join(sorted(_REGISTRY))))
    tens = {k: v for k, v in kwargs.items() if torch.is_tensor(v)}
    params = {k: (v if isinstance(v, str) else str(v)) for k, v in kwargs.items() if not torch.is_tensor(v) and k != 'name'}
    prop = _REGISTRY[op_type](**params)
    names = prop.list_arguments()
    in_data = list(args) + [tens[n] for n in names[len(args):] if n in tens]
    if len(in_data) != len(names):
        raise ValueError("%s expects arguments %s" % (op_type, names))
    in_shapes = [list(t.shape) for t in in_data]
    res = prop.infer_shape(in_shapes)
    out_shapes = res[1]
    op = prop.create_operator(in_data[0].device, in_shapes, [t.dtype for t in in_data])
    out_data = [torch.empty(tuple(int(d) for d in s), device=in_data[0].device) for s in out_shapes]
    op.forward(is_train, ['write'] * len(out_data), in_data, out_data, [])
    return out_data[0] if len(out_data) == 1 else out_data


def _tuple(v, cast=float):
    if isinstance(v, str):
        v = v.strip("()[] ").split(",")
    return tuple(cast(x) for x in v if str(x).strip() != "")


# ------------------------------------------------------------------------------------------------
# The SNIPER native operators under the reference's names.
# ------------------------------------------------------------------------------------------------
@register('MultiProposalTarget')
class MultiProposalTargetProp(CustomOpProp):
    """MultiProposalTargetParam (multi_proposal_target-inl.h:55-93): same keyword names and defaults."""

    def __init__(self, batch_size=16, rpn_pre_nms_top_n=6000, rpn_post_nms_top_n=300, threshold=0.7, rpn_min_size=16,
                 scales=(2, 4, 7, 10, 13, 16, 24), ratios=(0.5, 1, 2), feature_stride=16, bbox_scale=1.0,
                 output_score=False, iou_loss=False, workspace=256, crowd_boxes=None, layout=ops.NCHW):
        super(MultiProposalTargetProp, self).__init__(need_top_grad=False)
        self.post = int(rpn_post_nms_top_n)
        self.threshold = float(threshold)
        self.scales = _tuple(scales)
        self.ratios = _tuple(ratios)
        self.stride = int(feature_stride)
        self.layout = int(layout)
        # crowd_boxes is passed by mobilenetv2_e2e.py:247-255 but not declared by this fork's op: accept, ignore

    def list_arguments(self):
        return ['cls_prob', 'bbox_pred', 'im_info', 'gt_boxes', 'valid_ranges']

    def list_outputs(self):
        return ['rois', 'label', 'bbox_target', 'bbox_weight']

    def infer_shape(self, in_shape):
        n = in_shape[0][0] * self.post
        return in_shape, [[n, 5], [n, 1], [n, 4], [n, 4]], []

    def create_operator(self, ctx, shapes, dtypes):
        prop = self

        class _Op(CustomOp):
            def forward(self, is_train, req, in_data, out_data, aux):
                r = ops.multi_proposal_target(in_data[0].contiguous(), in_data[1].contiguous(), in_data[2], in_data[3],
                                              in_data[4], feat_stride=prop.stride, scales=prop.scales,
                                              ratios=prop.ratios, rpn_post_nms_top_n=prop.post,
                                              threshold=prop.threshold, layout=prop.layout)
                for i, t in enumerate(r):
                    self.assign(out_data[i], req[i], t.view(out_data[i].shape))

            def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
                for i in range(len(in_grad)):          # multi_proposal_target.cu:591-615
                    self.assign(in_grad[i], req[i], 0)
        return _Op()


def MultiProposalTarget(cls_prob, bbox_pred, im_info, gt_boxes, valid_ranges, **kw):
    """mx.sym.MultiProposalTarget(...) -> (rois, label, bbox_target, bbox_weight)."""
    return Custom(cls_prob=cls_prob, bbox_pred=bbox_pred, im_info=im_info, gt_boxes=gt_boxes,
                  valid_ranges=valid_ranges, op_type='MultiProposalTarget', **kw)


@register('MultiProposal')
class MultiProposalProp(CustomOpProp):
    """MultiProposalParam (multi_proposal-inl.h:55-100): same keyword names and defaults; outputs `output`, `score`
    (multi_proposal-inl.h:147-155).  The exact NMS of the CPU operator (multi_proposal.cc) runs by default;
    `fast_nms=True` selects the GPU build's FastNMS restricted to the anchor-overlap map of `roi_iou_thresh`
    (multi_proposal.cu:267-387) and `suppress_anchor_types=True` its anchor-type suppression (multi_proposal.cu:505-508)."""

    def __init__(self, batch_size=16, rpn_pre_nms_top_n=6000, rpn_post_nms_top_n=300, threshold=0.7, rpn_min_size=4,
                 scales=(2, 4, 7, 10, 13, 16, 24), ratios=(0.5, 1, 2), feature_stride=16, bbox_scale=1.0,
                 roi_iou_thresh=0.3, workspace=128, suppress_anchor_types=False, fast_nms=False, layout=ops.NCHW):
        super(MultiProposalProp, self).__init__(need_top_grad=False)
        self.fast_nms = str(fast_nms) in ("True", "true", "1")
        self.roi_iou_thresh = float(roi_iou_thresh)
        self.post = int(rpn_post_nms_top_n)
        self.threshold = float(threshold)
        self.scales = _tuple(scales)
        self.ratios = _tuple(ratios)
        self.stride = int(feature_stride)
        self.layout = int(layout)
        self.suppress = str(suppress_anchor_types) in ("True", "true", "1")
        # the reference ignores rpn_pre_nms_top_n and always sorts min(12000, A*H*W) rows (multi_proposal.cc:176)
        self.pre = 12000

    def list_arguments(self):
        return ['cls_prob', 'bbox_pred', 'im_info']

    def list_outputs(self):
        return ['output', 'score']

    def infer_shape(self, in_shape):
        n = in_shape[0][0] * self.post
        return in_shape, [[n, 5], [n, 1]], []

    def create_operator(self, ctx, shapes, dtypes):
        prop = self

        class _Op(CustomOp):
            def forward(self, is_train, req, in_data, out_data, aux):
                r = ops.multi_proposal(in_data[0].contiguous(), in_data[1].contiguous(), in_data[2],
                                       feat_stride=prop.stride, scales=prop.scales, ratios=prop.ratios,
                                       rpn_pre_nms_top_n=prop.pre, rpn_post_nms_top_n=prop.post,
                                       threshold=prop.threshold, suppress_anchor_types=prop.suppress,
                                       fast_nms=prop.fast_nms, roi_iou_thresh=prop.roi_iou_thresh, layout=prop.layout)
                for i, t in enumerate(r):
                    self.assign(out_data[i], req[i], t.view(out_data[i].shape))

            def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
                for i in range(len(in_grad)):          # multi_proposal.cc:376-394
                    self.assign(in_grad[i], req[i], 0)
        return _Op()


def MultiProposal(cls_prob, bbox_pred, im_info, **kw):
    """mx.sym.MultiProposal(...) -> (rois, score)."""
    return Custom(cls_prob=cls_prob, bbox_pred=bbox_pred, im_info=im_info, op_type='MultiProposal', **kw)


class _PoolFn(torch.autograd.Function):
    """Differentiable wrappers so user code written against autograd can call the pooling ops directly."""

    @staticmethod
    def forward(ctx, data, rois, trans, kw):
        no_trans = kw['no_trans']
        out, _, _ = ops.deform_psroi_fwd(data, rois, None if no_trans else trans, want_count=False, **kw)
        ctx.save_for_backward(data, rois, trans if trans is not None else rois)
        ctx.kw = kw
        return out

    @staticmethod
    def backward(ctx, g):
        data, rois, trans = ctx.saved_tensors
        kw = ctx.kw
        dd, dt = ops.deform_psroi_bwd(g.contiguous(), data, rois, None if kw['no_trans'] else trans, **kw)
        return dd, None, (None if kw['no_trans'] else dt), None


def DeformablePSROIPooling(data, rois, trans=None, spatial_scale=0.0625, output_dim=256, group_size=1, pooled_size=7,
                           part_size=0, sample_per_part=1, trans_std=0.0, no_trans=False, layout=ops.NCHW, **_):
    """mx.contrib.sym.DeformablePSROIPooling (deformable_psroi_pooling-inl.h:51-75) -> output."""
    kw = dict(spatial_scale=float(spatial_scale), output_dim=int(output_dim), group_size=int(group_size),
              pooled_size=int(pooled_size), part_size=int(part_size), sample_per_part=int(sample_per_part),
              trans_std=float(trans_std), no_trans=bool(no_trans) or trans is None, layout=int(layout))
    return _PoolFn.apply(data, rois, trans, kw)


def PSROIPooling(data, rois, spatial_scale, output_dim, pooled_size, group_size=0, layout=ops.NCHW, **_):
    """mx.contrib.sym.PSROIPooling (psroi_pooling-inl.h) -> output (forward only helper)."""
    g = int(group_size) or int(pooled_size)
    return ops.psroi_fwd(data, rois, spatial_scale=float(spatial_scale), output_dim=int(output_dim), group_size=g,
                         pooled_size=int(pooled_size), layout=int(layout))[0]


def DeformableConvolution(data, offset, weight, bias=None, kernel=(3, 3), stride=(1, 1), dilate=(1, 1), pad=(0, 0),
                          num_filter=None, num_group=1, num_deformable_group=1, no_bias=False, **_):
    """mx.contrib.symbol.DeformableConvolution (deformable_convolution-inl.h:59-96), NHWC tensors:
    data [N,H,W,C], offset [N,Ho,Wo,>=dg*2*kh*kw], weight [Cout, kh*kw*C] -> [N,Ho,Wo,Cout]."""
    assert int(num_group) == 1, "grouped deformable convolution is not used by the SNIPER symbols"
    kh, kw = _tuple(kernel, int)
    col = ops.deform_im2col(data, offset, kh=kh, kw=kw, stride=_tuple(stride, int)[0], dil=_tuple(dilate, int)[0],
                            pad=_tuple(pad, int)[0], dgroups=int(num_deformable_group))
    y = ops.gemm_nt(col, weight, bias=None if no_bias else bias)
    return y.view(data.shape[0], offset.shape[1], offset.shape[2], weight.shape[0])
