"""Symbol-class facade (symbols/symbol.py + resnet_mx_101_e2e): argument / auxiliary names and shapes equal the oracle's
parameter dictionaries (reference names, OIHW), init_weight_rcnn fills exactly the layers the reference initialises, the
checkpoint callback writes the `*_test` copies."""
import os
import sys
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def _cfg():
    return SimpleNamespace(dataset=SimpleNamespace(NUM_CLASSES=81), network=SimpleNamespace(NUM_ANCHORS=21),
                           TRAIN=SimpleNamespace(AUTO_FOCUS=False))


def test_names_and_shapes_match_the_oracle_parameter_set():
    import cpu_train_step as CTS
    from sniper_b200 import symbols
    arg, aux = CTS.make_params(seed=5)
    inst = symbols.resnet_mx_101_e2e(n_proposals=300, momentum=0.995)
    sym = inst.get_symbol_rcnn(_cfg())
    assert inst.symbol is sym and inst.get_bbox_param_names() == ['bbox_pred_weight', 'bbox_pred_bias']
    data_shapes = {'data': (20, 3, 512, 512), 'im_info': (20, 3), 'gt_boxes': (20, 100, 5), 'valid_ranges': (20, 2),
                   'label': (20, 21504), 'bbox_target': (20, 84, 32, 32), 'bbox_weight': (20, 84, 32, 32)}
    inst.infer_shape(data_shapes)
    params = {k: v for k, v in inst.arg_shape_dict.items() if k not in data_shapes}
    assert set(params) == set(arg) and set(inst.aux_shape_dict) == set(aux)
    for k, v in arg.items():
        assert tuple(params[k]) == v.shape, k
    for k, v in aux.items():
        assert tuple(inst.aux_shape_dict[k]) == v.shape, k
    assert inst.arg_shape_dict['label'] == (20, 21504)
    assert list(inst.out_shape_dict.values()) == [(20, 2, 672, 32), (20, 84, 32, 32), (20, 300, 81), (20, 300, 4), (6000,)]
    inst.check_parameter_shapes(arg, aux, data_shapes)
    bad = dict(arg); bad['conv0_weight'] = np.zeros((64, 3, 3, 3), np.float32)
    try:
        inst.check_parameter_shapes(bad, aux, data_shapes)
        raise SystemExit("shape mismatch not detected")
    except AssertionError as e:
        assert 'conv0_weight' in str(e)


def test_init_weight_rcnn_fills_the_new_layers_only():
    from sniper_b200 import symbols
    inst = symbols.resnet_mx_101_e2e()
    inst.get_symbol_rcnn(_cfg())
    inst.infer_shape({'data': (2, 3, 512, 512)})
    arg, aux = {}, {}
    inst.init_weight_rcnn(_cfg(), arg, aux, seed=1)
    new = {'rpn_conv_3x3', 'rpn_cls_score', 'rpn_bbox_pred', 'conv_new_1', 'offset', 'fc_new_1', 'fc_new_2', 'cls_score',
           'bbox_pred', 'stage4_unit1_offset', 'stage4_unit2_offset', 'stage4_unit3_offset'}
    assert set(arg) == {n + s for n in new for s in ('_weight', '_bias')} and not aux
    for k, v in arg.items():
        assert v.dtype == np.float32 and v.shape == tuple(inst.arg_shape_dict[k])
        if k.endswith('_bias') or 'offset' in k:
            assert not v.any(), k                          # zero-initialised
        else:
            assert 0.009 < v.std() < 0.011, k              # N(0, 0.01)
    assert abs(inst.get_msra_std((256, 256, 3, 3)) - np.sqrt(2.0 / 2304)) < 1e-12


def test_test_symbol_and_checkpoint_callback(tmp_path):
    from sniper_b200 import checkpoint, symbols
    inst = symbols.resnet_mx_101_e2e(test_nbatch=4)
    sym = inst.get_symbol_rcnn(_cfg(), is_train=False)
    assert sym.list_arguments()[:4] == ['data', 'im_info', 'im_ids', 'chip_ids']
    inst.infer_shape({'data': (4, 3, 512, 512)})
    assert inst.out_shape_dict['rois_output'] == (1200, 5) and inst.out_shape_dict['cls_prob_reshape_output'] == (4, 300, 81)
    inst.check_parameter_shapes({k: np.zeros(s, np.float32) for k, s in inst.arg_shape_dict.items()},
                                {k: np.zeros(s, np.float32) for k, s in inst.aux_shape_dict.items()},
                                {'data': 0, 'im_info': 0, 'im_ids': 0, 'chip_ids': 0}, is_train=False)
    arg = {'bbox_pred_weight': np.arange(4096, dtype=np.float32).reshape(4, 1024), 'bbox_pred_bias': np.ones(4, np.float32),
           'conv0_weight': np.zeros((64, 3, 7, 7), np.float32)}
    aux = {'bn0_moving_mean': np.zeros(64, np.float32)}
    cb = symbols.checkpoint_callback(inst.get_bbox_param_names(), str(tmp_path / "m"), None, None)
    cb(6, sym, arg, aux)
    a2, x2 = checkpoint.load_param(str(tmp_path / "m"), 7)
    stds = np.array(checkpoint.BBOX_STDS_TEST, np.float32)
    assert np.array_equal(a2['bbox_pred_weight_test'], (arg['bbox_pred_weight'].T * stds).T)
    assert np.array_equal(a2['bbox_pred_bias_test'], stds) and 'bbox_pred_weight_test' not in arg
    a3, _ = checkpoint.load_param(str(tmp_path / "m"), 7, process=True)
    assert np.array_equal(a3['bbox_pred_bias'], stds) and np.array_equal(x2['bn0_moving_mean'], aux['bn0_moving_mean'])
