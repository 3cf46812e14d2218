"""The operator-registration surface mirrored from python/mxnet/operator.py / lib/operator_py."""
import numpy as np
import pytest
import torch

import oracle_lib as O
from sniper_b200 import operator_py as op, synth


def test_register_and_custom_roundtrip_cpu():
    @op.register('ScaleBy')
    class ScaleProp(op.CustomOpProp):
        def __init__(self, factor):
            super(ScaleProp, self).__init__(need_top_grad=True)
            self.factor = float(factor)               # kwargs arrive as strings, as in MXNet

        def list_arguments(self):
            return ['data']

        def infer_shape(self, in_shape):
            return in_shape, [in_shape[0]], []

        def create_operator(self, ctx, shapes, dtypes):
            f = self.factor

            class _Op(op.CustomOp):
                def forward(self, is_train, req, in_data, out_data, aux):
                    self.assign(out_data[0], req[0], in_data[0] * f)

                def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
                    self.assign(in_grad[0], req[0], out_grad[0] * f)
            return _Op()

    assert 'ScaleBy' in op.get_all_registered_operators()
    x = torch.arange(6.0).view(2, 3)
    y = op.Custom(data=x, op_type='ScaleBy', factor=2.5)
    assert torch.equal(y, x * 2.5)
    o = ScaleProp('2').create_operator(None, None, None)
    g = torch.ones(2, 3)
    o.assign(g, 'add', torch.ones(2, 3))
    assert float(g.sum()) == 12.0
    o.assign(g, 'null', torch.zeros(2, 3))
    assert float(g.sum()) == 12.0
    with pytest.raises(KeyError):
        op.Custom(data=x, op_type='NoSuchOp')
    for name in ('MultiProposalTarget', 'MultiProposal'):
        assert name in op.get_all_registered_operators()
    q = op.MultiProposalProp(batch_size='4', rpn_post_nms_top_n='300', roi_iou_thresh='0.3', fast_nms='True',
                             suppress_anchor_types='True')      # kwargs arrive as strings (box_annotator_ohem.py:86-120)
    assert q.list_arguments() == ['cls_prob', 'bbox_pred', 'im_info'] and q.list_outputs() == ['output', 'score']
    assert q.infer_shape([[4, 42, 32, 32], [4, 84, 32, 32], [4, 3]])[1] == [[1200, 5], [1200, 1]]
    assert q.fast_nms and q.suppress and q.pre == 12000 and abs(q.roi_iou_thresh - 0.3) < 1e-9
    p = op.MultiProposalTargetProp(batch_size='20', scales='(2,4,7,10,13,16,24)', ratios='(0.5,1,2)', crowd_boxes='x')
    assert p.list_arguments() == ['cls_prob', 'bbox_pred', 'im_info', 'gt_boxes', 'valid_ranges']
    assert p.list_outputs() == ['rois', 'label', 'bbox_target', 'bbox_weight']
    assert p.infer_shape([[20, 42, 32, 32]] + [[0]] * 4)[1] == [[6000, 5], [6000, 1], [6000, 4], [6000, 4]]


@pytest.mark.gpu
def test_reference_named_ops_on_gpu():
    inp = synth.mpt_inputs(5, 2, 21, 32, 32)
    res = O.multi_proposal_target(*inp)
    t = [torch.from_numpy(a).cuda() for a in inp]
    rois, label, bt, bw = op.MultiProposalTarget(*t, batch_size=2)
    assert rois.cpu().numpy().tobytes() == res["rois"].tobytes()
    assert label.shape == (600, 1) and np.array_equal(label.view(-1).cpu().numpy(), res["label"])
    # autograd-capable pooling wrapper
    rng = np.random.RandomState(1)
    data = torch.randn(2, 8, 16, 16, device="cuda", requires_grad=True)
    r = torch.from_numpy(synth.rois_for_pool(rng, 16, 2, chip=256)).cuda()
    out = op.DeformablePSROIPooling(data, r, None, spatial_scale=0.0625, output_dim=8, group_size=1, pooled_size=3,
                                    part_size=3, sample_per_part=2, no_trans=True)
    out.sum().backward()
    assert data.grad is not None and float(data.grad.abs().sum()) > 0
